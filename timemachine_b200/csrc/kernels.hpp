// Host-callable launchers for every CUDA kernel of the hot path. Each launcher enqueues on `stream` and never
// synchronises; all control decisions that the reference takes on the host after a D2H copy (neighbour-list rebuild,
// nonbonded_all_pairs.cu:207-236) are taken on the device here.
#pragma once

#include "common.cuh"
#include "nb_types.cuh"

namespace tmb {

// ---------------------------------------------------------------------------------------------------------------
// Nonbonded tile kernel (reference k_nonbonded_unified, k_nonbonded.cuh:329)
template <typename Real> struct NbTileArgs {
    int K;    // gathered atoms (sorted slots); padding column entries are >= K
    int NR;   // row slots are [0, NR); NR == K means all-pairs (upper triangle, i < j)
    const unsigned int *tile_count;
    const int *tile_rows;
    const unsigned int *tile_cols;
    const Vec4<Real> *xw;
    const Vec4<Real> *qse;
    const double *box;
    double beta;
    double cutoff;
    // Fixed-point outputs in ATOM order, added to atomically (integer sums: order-free).  The kernels translate sorted
    // slots through perm when they flush a tile, so no scatter pass (reference k_scatter_accum k_nonbonded.cuh:86) runs.
    const unsigned int *perm; // [K] sorted slot -> atom index
    u64 *du_dx;               // [N,3] nullable
    u64 *du_dp;               // [N,4] nullable
    i128 *u_partials;
    unsigned int *ticket;
    i128 *d_u; // overwritten
    unsigned int *rebuild_flag; // cleared by this kernel (the build, if any, ran before it in stream order);
                                // rebuild_flag[2] counts the builds
    unsigned int tile_capacity;
    unsigned int *tile_cursor;  // dynamic tile scheduling counter, zeroed by k_nb_prepare of the same evaluation
    int grid_ctas = 0;              // persistent grid size; 0 = every CTA slot of the device
    unsigned int static_tiles = 0;  // tiles each warp takes up front (clamped to tile_count / warps); the rest is dynamic
    bool prefilter = true;          // f32 cq kernel: packed-half distance prefilter (k_nb_tiles_cq.cu); results identical
};
// CTA slots a machine-filling tile launch leaves free so that a small concurrent tile launch (the ligand-environment
// interaction group next to the environment all-pairs term) is not serialised behind it; TMB_NB_RESERVE overrides.
int nb_tiles_reserved_ctas();
// f32 "compaction queue" formulation (k_nb_tiles_cq.cu); launch_nb_tiles<float> uses it unless TMB_NB_RING=1
int nb_tiles_cq_max_grid();
void launch_nb_tiles_cq(const NbTileArgs<float> &args, bool with_u, bool with_dx, bool with_dp, cudaStream_t stream);
// the same tile evaluation behind an asynchronous staging pipeline (TMA bulk copy of the index vector + cp.async gather of
// the column atoms, k_nb_tiles_cq_async.cu); selected with TMB_NB_ASYNC=1 for A/B measurements
int nb_tiles_cq_async_max_grid();
void launch_nb_tiles_cq_async(const NbTileArgs<float> &args, bool with_u, bool with_dx, bool with_dp, cudaStream_t stream);
template <typename Real> int nb_tiles_max_grid();
template <typename Real>
void launch_nb_tiles(const NbTileArgs<Real> &args, bool with_u, bool with_dx, bool with_dp, cudaStream_t stream);

// Gather + cast + rebuild decision (reference k_gather_coords_and_params k_nonbonded.cuh:58 and
// k_check_rebuild_coords_and_box_gather :11, fused; the flag never leaves the device).
template <typename Real> struct NbPrepareArgs {
    int K;
    const unsigned int *perm; // [K] sorted slot -> atom index
    const double *x;          // [N,3]
    const double *p;          // [N,4]
    const double *box;        // [9]
    const Vec4<Real> *xw_build; // [K] sorted-order packed coordinates at the last neighbour-list build
    const double *box_build;  // [9]
    double padding;
    int force_rebuild;        // host-known (after a sort / set_atom_idxs)
    unsigned int *flag;       // [1] rebuild flag: set here, consumed by the build kernels, cleared by the tile kernel
    unsigned int *tile_cursor; // [1] zeroed here for the tile kernel's dynamic scheduler
    Vec4<Real> *xw;
    Vec4<Real> *qse;
    // Optional fusion of the block-bounds pass (all-pairs layout only: block b == sorted slots [32 b, 32 b + 32)).
    // Bounds are recomputed every step into ctr/ext (only read by a build); whoever raises the rebuild flag also
    // clears the tile counter, so a build needs no separate bounds/reset launch.
    Real *ctr = nullptr; // [ceil(K/32),3]
    Real *ext = nullptr;
    unsigned int *reset_count = nullptr;
    unsigned int *reset_overflow = nullptr;
};
template <typename Real> void launch_nb_prepare(const NbPrepareArgs<Real> &args, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// Neighbour list (reference k_neighborlist.cuh)
template <typename Real> struct BlockBoundsArgs {
    int num_blocks;
    int num_idxs;
    const unsigned int *idxs; // nullable: idx = base + i
    int base;
    const double *coords;     // exactly one of coords / xw is non-null
    const Vec4<Real> *xw;
    const double *box;
    Real *ctr; // [num_blocks,3]
    Real *ext; // [num_blocks,3]
    const unsigned int *flag; // nullable: skip all work when *flag == 0
    // Fused bookkeeping of a (re)build, all optional:
    unsigned int *reset_count;    // tile counter and
    unsigned int *reset_overflow; //   overflow flag to clear before the tile build
    Vec4<Real> *xw_build;         // snapshot: xw_build[slot] = xw[slot] for every slot this launch covers
    double *box_build; // box_build = box
};
template <typename Real> void launch_block_bounds(const BlockBoundsArgs<Real> &args, cudaStream_t stream);

template <typename Real> struct BuildTilesArgs {
    int N;  // sentinel for "no atom" (>= N)
    int NC; // number of column indices
    int NR; // number of row indices
    bool upper_triangular;
    const unsigned int *col_idxs; // nullable: col_base + i
    int col_base;
    const unsigned int *row_idxs; // nullable: row_base + i
    int row_base;
    const Real *col_ctr, *col_ext, *row_ctr, *row_ext;
    const double *coords; // exactly one of coords / xw
    const Vec4<Real> *xw;
    const double *box;
    double cutoff;
    TileList tiles;
    const unsigned int *flag; // nullable: skip all work when *flag == 0
    // Optional "state at build time" snapshot taken by the build itself (when the bounds pass was fused into
    // k_nb_prepare): xw_build[k] = xw[k] for k < snap_slots, box_build = box.
    Vec4<Real> *snap_xw_build = nullptr;
    double *snap_box_build = nullptr;
    int snap_slots = 0;
};
// The tile counter is reset by the bounds pass that precedes a build (k_block_bounds, or k_nb_prepare when fused).
template <typename Real> void launch_build_tiles(const BuildTilesArgs<Real> &args, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// Hilbert curve sort (reference hilbert_sort.cu, k_hilbert.cu)
void launch_hilbert_lut(unsigned int *lut /*[128^3]*/, cudaStream_t stream);
void launch_hilbert_keys(
    int n, const unsigned int *atom_idxs, const double *coords, const double *box, const unsigned int *lut,
    unsigned int *keys, unsigned int *vals, cudaStream_t stream);
size_t radix_sort_pairs_temp_bytes(int n);
void radix_sort_pairs(
    void *temp, size_t temp_bytes, const unsigned int *keys_in, unsigned int *keys_out, const unsigned int *vals_in,
    unsigned int *vals_out, int n, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// Explicit pair list (reference k_nonbonded_pair_list.cuh:19)
template <typename Real> struct PairListArgs {
    int M;
    const double *x;
    const double *p;
    const double *box;
    const int *pair_idxs; // [M,2]
    const double *scales; // [M,2] (charge, lj)
    double beta;
    double cutoff;
    bool negated;
    u64 *du_dx; // [N,3] nullable
    u64 *du_dp; // [N,4] nullable
    i128 *u_partials;
    unsigned int *ticket;
    i128 *d_u; // nullable
};
template <typename Real> void launch_pair_list(const PairListArgs<Real> &args, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// Bonded terms (reference k_harmonic_bond.cuh, k_harmonic_angle.cuh, k_periodic_torsion.cuh)
struct BondedArgs {
    int n_terms;
    const double *x;
    const double *p;
    const int *idxs;
    u64 *du_dx;
    u64 *du_dp;
    i128 *u_partials;
    unsigned int *ticket;
    i128 *d_u;
};
int bonded_grid(int n_terms);
// Restraints and the precomputed pair list (the rest of the HostGuestSystem potential set)
struct RestraintArgs {
    BondedArgs b;
    const double *box = nullptr; // flat-bottom bond, precomputed pairs
    const int *signs = nullptr;  // chiral bond restraint [R]
    double beta = 0;             // precomputed pairs; log flat-bottom bond (1 / kT)
    double cutoff = 0;
    bool full_gradient = false;  // precomputed pairs: keep the ES gradient of pairs without LJ (the reference drops it)
};
template <typename Real> void launch_flat_bottom_bond(const RestraintArgs &args, cudaStream_t stream);
template <typename Real> void launch_log_flat_bottom_bond(const RestraintArgs &args, cudaStream_t stream);
template <typename Real> void launch_chiral_atom_restraint(const RestraintArgs &args, cudaStream_t stream);
template <typename Real> void launch_chiral_bond_restraint(const RestraintArgs &args, cudaStream_t stream);
template <typename Real> void launch_nonbonded_precomputed(const RestraintArgs &args, cudaStream_t stream);
// CentroidRestraint: kb (|centroid_a - centroid_b| - b0)^2 over two atom groups (reference k_centroid_restraint.cuh:7-84)
struct CentroidArgs {
    const double *x;
    const int *group_a;
    const int *group_b;
    int n_a, n_b;
    double kb, b0;
    u64 *du_dx; // nullable
    i128 *d_u;  // nullable, overwritten
};
template <typename Real> void launch_centroid_restraint(const CentroidArgs &args, cudaStream_t stream);
template <typename Real> void launch_harmonic_bond(const BondedArgs &args, cudaStream_t stream);
template <typename Real> void launch_harmonic_angle(const BondedArgs &args, cudaStream_t stream);
template <typename Real> void launch_periodic_torsion(const BondedArgs &args, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// Langevin BAOAB update with in-kernel Philox noise (reference k_integrator.cuh:6-62)
struct BaoabArgs {
    int N;
    float ca;
    const unsigned int *idxs; // nullable
    const float *cbs;
    const float *ccs;
    const float *noise;       // nullable: externally supplied N x 3 normals (tests); else Philox(seed, step)
    unsigned long long seed;
    unsigned long long step;                // noise counter = step + (*step_base if step_base != nullptr)
    const unsigned long long *step_base;    // device-resident base so CUDA-graph replays draw fresh noise
    double *x;
    double *v;
    u64 *du_dx; // consumed and zeroed
    float dt;
};
void launch_baoab(const BaoabArgs &args, cudaStream_t stream);

// BAOAB fused with the NEXT evaluation's k_nb_prepare of an all-pairs potential (Context: steps 0..8 of a 10-step graph
// block).  The integrator walks the potential's sorted slots (a warp = a 32-atom block), integrates atom perm[k] and, with
// the new coordinates still in registers, writes the packed sorted copy xw[k], makes the displacement / rebuild decision
// and the block bounds exactly as k_nb_prepare would on the next step; atoms outside the potential's set come from
// `others`.  The potential then skips its own prepare launch: one streaming pass and one launch less on the critical path.
template <typename Real> struct FusedPrepareArgs {
    int K;                      // slots of the all-pairs potential
    const unsigned int *perm;   // [K] slot -> atom
    const unsigned int *others; // [n_others] atoms that are not in perm
    int n_others;
    Vec4<Real> *xw;             // [Kpad] in/out: w is kept (parameters do not change inside a block)
    const Vec4<Real> *xw_build;
    const double *box;
    const double *box_build;
    double padding;
    unsigned int *flag;
    unsigned int *tile_cursor;
    Real *ctr;
    Real *ext;
    unsigned int *reset_count;
    unsigned int *reset_overflow;
};
template <typename Real> void launch_baoab_prepare(const BaoabArgs &args, const FusedPrepareArgs<Real> &f, cudaStream_t stream);
// Velocity Verlet in f64 (reference k_integrator.cuh:64-130): mode 0 = full step, 1 = opening half kick + drift
// (VelocityVerletIntegrator::initialize), 2 = closing half kick (finalize)
struct VerletArgs {
    int N;
    const unsigned int *idxs; // nullable
    const double *cbs;        // -dt / mass (lib/__init__.py:31-34)
    double *x;
    double *v;
    u64 *du_dx; // consumed and zeroed
    double dt;
};
void launch_velocity_verlet(const VerletArgs &args, int mode, cudaStream_t stream);
void launch_fill_normal(float *out, int n, unsigned long long seed, unsigned long long step, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// Small utilities
void launch_sum_i128(const i128 *in, int n, i128 *out, cudaStream_t stream);
void launch_iota(unsigned int *out, int n, unsigned int base, cudaStream_t stream);

} // namespace tmb
