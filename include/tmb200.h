/*
 * tmb200.h - C ABI of libtmb200.so: the B200-native (sm_100a) force-evaluation + integration hot path of timemachine.
 *
 * Every entry point is what a binding for the reference's `timemachine.lib.custom_ops` module would call for this
 * path; the reference's own boundary is the pybind11 module in timemachine/cpp/src/wrap_kernels.cpp, and each
 * function below cites the lambda / method it replaces (file:line in /root/reference).  Conventions are the
 * reference's:
 *   - coords f64[N,3], params f64[...], box f64[3,3] row-major, index arrays int32 (uint32 where the reference has it);
 *   - outputs are the raw FIXED-POINT accumulators (uint64 two's complement, int128 energies as {lo,hi}); conversion to
 *     float happens on the caller's side exactly as wrap_kernels.cpp:83-89,1079-1096 does (tmb_*_fixed_to_float helps);
 *   - a function returns 0 on success, 1 on a std::runtime_error-class failure, 2 when the GPU/driver is unusable
 *     (the reference raises custom_ops.InvalidHardware, gpu_utils.cuh:31-53); tmb_last_error() gives the message,
 *     which matches the reference's message text where its tests assert on it;
 *   - not thread-safe, one CUDA device per process (potential.hpp:7 in the reference says the same).
 * No torch / pybind types appear here: plain pointers and sizes only.
 */
#ifndef TMB200_H
#define TMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tmb_i128 {
    uint64_t lo;
    int64_t hi;
} tmb_i128;

typedef void *tmb_potential;       /* std::shared_ptr<Potential>            (wrap_kernels.cpp:729  declare_potential) */
typedef void *tmb_bound_potential; /* std::shared_ptr<BoundPotential>       (wrap_kernels.cpp:1133 declare_bound_potential) */
typedef void *tmb_integrator;      /* std::shared_ptr<Integrator>           (wrap_kernels.cpp:692-729) */
typedef void *tmb_context;         /* Context                               (wrap_kernels.cpp:296) */
typedef void *tmb_neighborlist;    /* Neighborlist<float|double>            (wrap_kernels.cpp:113) */
typedef void *tmb_hilbert_sort;    /* HilbertSort                           (wrap_kernels.cpp:174) */
typedef void *tmb_sampler;         /* SegmentedWeightedRandomSampler        (wrap_kernels.cpp:196) */
typedef void *tmb_mover;           /* std::shared_ptr<Mover>                (wrap_kernels.cpp:1591 declare_mover) */

#define TMB_OK 0
#define TMB_ERROR 1
#define TMB_INVALID_HARDWARE 2

#define TMB_F32 32
#define TMB_F64 64

/* ---- library ------------------------------------------------------------------------------------------------- */
const char *tmb_last_error(void);
int tmb_version(void);
uint64_t tmb_fixed_exponent(void);          /* custom_ops.FIXED_EXPONENT, wrap_kernels.cpp:2311 */
int tmb_cuda_device_reset(void);            /* custom_ops.cuda_device_reset, wrap_kernels.cpp:2222 */
long long tmb_kernel_launch_count(void);    /* kernels launched by this library so far */
int tmb_set_stream(void *cuda_stream);      /* stream used by the host-buffer entry points (default: legacy stream 0) */
int tmb_device_synchronize(void);

/* ---- potentials: constructors (argument order = the reference's ctor order) ------------------------------------- */
/* HarmonicBond_{f32,f64}(bond_idxs i32[B,2])                       wrap_kernels.cpp:1311-1322 */
int tmb_harmonic_bond_create(int precision, const int32_t *bond_idxs, int n_values, tmb_potential *out);
/* HarmonicAngle_{f32,f64}(angle_idxs i32[A,3])                     wrap_kernels.cpp:1396-1408 */
int tmb_harmonic_angle_create(int precision, const int32_t *angle_idxs, int n_values, tmb_potential *out);
/* PeriodicTorsion_{f32,f64}(torsion_idxs i32[T,4])                 wrap_kernels.cpp:1432-1444 */
int tmb_periodic_torsion_create(int precision, const int32_t *torsion_idxs, int n_values, tmb_potential *out);
/* the rest of the HostGuestSystem potential set (fe/system.py:133-143)
 * FlatBottomBond_{f32,f64}(bond_idxs[B,2])                          wrap_kernels.cpp:1324-1335, flat_bottom_bond.cu
 * ChiralAtomRestraint_{f32,f64}(idxs[R,4])                          wrap_kernels.cpp:1366-1378, chiral_atom_restraint.cu
 * ChiralBondRestraint_{f32,f64}(idxs[R,4], signs[R] in {1,-1})      wrap_kernels.cpp:1380-1394, chiral_bond_restraint.cu
 * NonbondedPairListPrecomputed_{f32,f64}(pair_idxs[M,2], beta, cutoff); params [M,4] = (q_ij, sig_ij, eps_ij, w_ij)
 *                                                                   wrap_kernels.cpp:1351-1364, nonbonded_precomputed.cu */
int tmb_flat_bottom_bond_create(int precision, const int32_t *bond_idxs, int n_values, tmb_potential *out);
/* CentroidRestraint_{f32,f64}(group_a_idxs[A], group_b_idxs[B], kb, b0)   wrap_kernels.cpp:1410-1430, centroid_restraint.cu
 * kb (|centroid_a - centroid_b| - b0)^2 with geometric centroids; takes no parameters (P = 0) */
int tmb_centroid_restraint_create(
    int precision, const int32_t *group_a_idxs, int n_a, const int32_t *group_b_idxs, int n_b, double kb, double b0,
    tmb_potential *out);
/* LogFlatBottomBond_{f32,f64}(bond_idxs[B,2], beta)   wrap_kernels.cpp:1337-1349, log_flat_bottom_bond.cu; params [B,3] */
int tmb_log_flat_bottom_bond_create(int precision, const int32_t *bond_idxs, int n_values, double beta, tmb_potential *out);
int tmb_chiral_atom_restraint_create(int precision, const int32_t *idxs, int n_values, tmb_potential *out);
int tmb_chiral_bond_restraint_create(
    int precision, const int32_t *idxs, int n_values, const int32_t *signs, int n_signs, tmb_potential *out);
int tmb_nonbonded_pair_list_precomputed_create(
    int precision, const int32_t *pair_idxs, int n_values, double beta, double cutoff, tmb_potential *out);
/* NonbondedAllPairs_*(num_atoms, beta, cutoff, atom_idxs|None, disable_hilbert_sort, nblist_padding)
 *                                                                  wrap_kernels.cpp:1446-1478 ; n_atom_idxs < 0 = None */
int tmb_nonbonded_all_pairs_create(
    int precision, int num_atoms, double beta, double cutoff, const int32_t *atom_idxs, int n_atom_idxs,
    int disable_hilbert_sort, double nblist_padding, tmb_potential *out);
int tmb_nonbonded_all_pairs_set_atom_idxs(tmb_potential pot, const int32_t *atom_idxs, int n);
int tmb_nonbonded_all_pairs_get_num_atom_idxs(tmb_potential pot, int *out);
int tmb_nonbonded_all_pairs_get_atom_idxs(tmb_potential pot, int32_t *out /* [num_atom_idxs] */);
/* NonbondedInteractionGroup_*(num_atoms, row_atom_idxs, beta, cutoff, col_atom_idxs|None, disable_hilbert_sort,
 * nblist_padding)                                                  wrap_kernels.cpp:1480-1561 ; n_col < 0 = complement */
int tmb_nonbonded_interaction_group_create(
    int precision, int num_atoms, const int32_t *row_atom_idxs, int n_row, double beta, double cutoff,
    const int32_t *col_atom_idxs, int n_col, int disable_hilbert_sort, double nblist_padding, tmb_potential *out);
int tmb_nonbonded_interaction_group_set_atom_idxs(
    tmb_potential pot, const int32_t *row_atom_idxs, int n_row, const int32_t *col_atom_idxs, int n_col);
/* NonbondedPairList_* / NonbondedExclusions_*(pair_idxs i32[M,2], scales f64[M,2], beta, cutoff)
 *                                                                  wrap_kernels.cpp:1563-1589 ; negated = Exclusions */
int tmb_nonbonded_pair_list_create(
    int precision, int negated, const int32_t *pair_idxs, int n_pair_values, const double *scales, int n_scale_values,
    double beta, double cutoff, tmb_potential *out);
/* SummedPotential(potentials, params_sizes, parallel)              wrap_kernels.cpp:1661-1675 */
int tmb_summed_potential_create(
    const tmb_potential *potentials, int n_potentials, const int32_t *params_sizes, int n_sizes, int parallel,
    tmb_potential *out);
/* FanoutSummedPotential(potentials, parallel)                      wrap_kernels.cpp:1677-1691 */
int tmb_fanout_summed_potential_create(const tmb_potential *potentials, int n_potentials, int parallel, tmb_potential *out);
int tmb_potential_destroy(tmb_potential pot);

/* ---- potentials: evaluation --------------------------------------------------------------------------------------- */
/* Potential.execute -> execute_host: host buffers in, fixed-point host buffers out (any of the three may be NULL)
 *                                                                  wrap_kernels.cpp:1039-1105, potential.cu:224-292 */
int tmb_potential_execute(
    tmb_potential pot, int N, int P, const double *coords, const double *params, const double *box, uint64_t *du_dx,
    uint64_t *du_dp, tmb_i128 *u);
/* Potential.execute_batch                                          wrap_kernels.cpp:736-831, potential.cu:70-144 */
int tmb_potential_execute_batch(
    tmb_potential pot, int coord_batches, int N, int param_batches, int P, const double *coords, const double *params,
    const double *boxes, uint64_t *du_dx, uint64_t *du_dp, tmb_i128 *u);
/* Potential.execute_batch_sparse                                   wrap_kernels.cpp:871-986, potential.cu:146-222 */
int tmb_potential_execute_batch_sparse(
    tmb_potential pot, int coords_size, int N, int params_size, int P, int batch_size, const uint32_t *coords_batch_idxs,
    const uint32_t *params_batch_idxs, const double *coords, const double *params, const double *boxes, uint64_t *du_dx,
    uint64_t *du_dp, tmb_i128 *u);
/* virtual Potential::du_dp_fixed_to_float (per-column exponents for nonbonded)  potential.cu:322, nonbonded_all_pairs.cu:292 */
int tmb_potential_du_dp_fixed_to_float(tmb_potential pot, int N, int P, const uint64_t *du_dp, double *out);
/* Potential::execute_device: DEVICE pointers, accumulates into du_dx/du_dp, overwrites u, enqueues on `cuda_stream`
 * without synchronising                                            potential.hpp:87-96 */
int tmb_potential_execute_device(
    tmb_potential pot, int N, int P, const double *d_coords, const double *d_params, const double *d_box,
    uint64_t *d_du_dx, uint64_t *d_du_dp, tmb_i128 *d_u, void *cuda_stream);
/* number of 32x32 interaction tiles in the cached neighbour list of a NonbondedAllPairs / InteractionGroup */
int tmb_nonbonded_num_tiles(tmb_potential pot, unsigned int *out);
/* neighbour-list (re)builds since construction; the rebuild decision itself never leaves the device
 * (reference: host-side flag read every step, nonbonded_all_pairs.cu:217-235) */
int tmb_nonbonded_num_rebuilds(tmb_potential pot, unsigned int *out);
/* capacity (tiles) of the list buffer and the worst case the reference allocates up front (neighborlist.cu:22-28); the
 * buffer here starts at 4 tiles per atom and grows when a build needs more */
int tmb_nonbonded_tile_capacity(tmb_potential pot, unsigned long long *capacity, unsigned long long *worst_case);
/* measurement hooks (bench.py roofline): bracket each tile-kernel launch with CUDA events on its launch stream;
 * drain returns the per-launch durations in ms recorded since the previous drain (at most `capacity`). */
int tmb_nonbonded_set_kernel_timing(tmb_potential pot, int on);
int tmb_nonbonded_drain_kernel_times(tmb_potential pot, float *out_ms, int capacity, int *n_out);

/* ---- BoundPotential                                              wrap_kernels.cpp:1133-1309 ----------------------- */
int tmb_bound_potential_create(tmb_potential pot, const double *params, int n_params, tmb_bound_potential *out);
int tmb_bound_potential_destroy(tmb_bound_potential bp);
int tmb_bound_potential_set_params(tmb_bound_potential bp, const double *params, int n_params);
int tmb_bound_potential_size(tmb_bound_potential bp, int *out);
int tmb_bound_potential_execute(
    tmb_bound_potential bp, int N, const double *coords, const double *box, uint64_t *du_dx, tmb_i128 *u);
int tmb_bound_potential_execute_batch(
    tmb_bound_potential bp, int coord_batches, int N, const double *coords, const double *boxes, uint64_t *du_dx,
    tmb_i128 *u);
int tmb_bound_potential_set_params_device(tmb_bound_potential bp, const double *d_params, int n_params, void *cuda_stream);
int tmb_bound_potential_execute_device(
    tmb_bound_potential bp, int N, const double *d_coords, const double *d_box, uint64_t *d_du_dx, tmb_i128 *d_u,
    void *cuda_stream);

/* ---- LangevinIntegrator(masses, temperature, dt, friction, seed) wrap_kernels.cpp:698-715 ------------------------ */
int tmb_langevin_integrator_create(
    const double *masses, int N, double temperature, double dt, double friction, int seed, tmb_integrator *out);
int tmb_langevin_integrator_destroy(tmb_integrator intg);
/* test hook: N x 3 f32 normals used on every step instead of the in-kernel Philox stream (NULL restores Philox) */
int tmb_langevin_integrator_set_noise(tmb_integrator intg, const float *noise);
/* Position of the integrator in its counter-based noise stream: step s of atom a draws Philox(seed; a, s).  The
 * reference's cuRAND generator has no such handle (langevin_integrator.cu:35-37: its state is not exposed, SURVEY.md §5
 * "checkpoint / resume"); here a driver can give every replica its own sub-stream (high bits) and resume it exactly, which
 * makes HREX trajectories independent of how replicas are laid out over GPUs (timemachine_b200/hrex.py). */
int tmb_langevin_integrator_set_step(tmb_integrator intg, unsigned long long step);
int tmb_langevin_integrator_get_step(tmb_integrator intg, unsigned long long *step);

/* rmsd_align(x1[N,3], x2[N,3]) -> x2 rotated (no reflection) and shifted onto x1   wrap_kernels.cpp:1975-2001,
 * rmsd_align.cpp:11-61.  A host function in the reference too (Eigen JacobiSVD); f64. */
int tmb_rmsd_align(const double *x1, const double *x2, int N, double *x2_aligned);

/* ---- VelocityVerletIntegrator(dt, cbs)                          wrap_kernels.cpp:717-729, verlet_integrator.cu ------ */
/* cbs[N] = -dt / mass (the caller's sign convention, lib/__init__.py:25-37); all arithmetic in f64.  A context driven by it
 * brackets every multiple_steps call with the half kicks (tmb_context_initialize / _finalize for single steps).  The
 * tmb_langevin_integrator_* accessors fail on it with "integrator must be LangevinIntegrator." */
int tmb_velocity_verlet_integrator_create(double dt, const double *cbs, int N, tmb_integrator *out);
/* releases a handle of either integrator class (tmb_langevin_integrator_destroy is the same call) */
int tmb_integrator_destroy(tmb_integrator intg);

/* ---- Mover / MonteCarloBarostat           wrap_kernels.cpp:1591-1659, barostat.cu ------------- */
/* MonteCarloBarostat<float>(N, pressure[bar], temperature[K], group_idxs, interval, bps, seed, adaptive_scaling_enabled,
 * initial_volume_scale_factor).  group_idxs is passed flattened: atoms of group g are
 * group_atoms[group_offsets[g] .. group_offsets[g+1]).                wrap_kernels.cpp:1621-1653 */
int tmb_barostat_create(
    int N, double pressure, double temperature, const int *group_atoms, const int *group_offsets, int n_groups,
    int interval, const tmb_bound_potential *bps, int n_bps, int seed, int adaptive_scaling_enabled,
    double initial_volume_scale_factor, tmb_mover *out);
int tmb_mover_destroy(tmb_mover m);
int tmb_mover_set_interval(tmb_mover m, int interval); /* wrap_kernels.cpp:1596 */
int tmb_mover_get_interval(tmb_mover m, int *out);     /* :1597 */
int tmb_mover_set_step(tmb_mover m, int step);         /* :1598 */
/* Mover::move_host: HOST coords[N,3] / box[3,3] in, moved copies out   wrap_kernels.cpp:1599-1616, mover.cu:7-23 */
int tmb_mover_move_host(tmb_mover m, int N, const double *coords, const double *box, double *out_coords, double *out_box);
int tmb_barostat_set_volume_scale_factor(tmb_mover m, double v); /* :1654 */
int tmb_barostat_get_volume_scale_factor(tmb_mover m, double *out); /* :1655 */
int tmb_barostat_set_adaptive_scaling(tmb_mover m, int on);      /* :1656 */
int tmb_barostat_get_adaptive_scaling(tmb_mover m, int *out);    /* :1657 */
int tmb_barostat_set_pressure(tmb_mover m, double pressure);     /* :1658 */
/* introspection for the parity tests: the two cuRAND uniforms of the last attempted move; {attempted, accepted} */
int tmb_barostat_last_uniforms(tmb_mover m, float *out2);
int tmb_barostat_counters(tmb_mover m, int *out2);

/* ---- Water exchange by biased deletion      wrap_kernels.cpp:196-294, 1693-1900, 2004-2117 ----- */
/* Molecules are passed flattened like the barostat's groups: atoms of molecule m are mol_atoms[mol_offsets[m] ..
 * mol_offsets[m+1]).  Real-typed results are returned widened to double.  precision: TMB_F32 / TMB_F64. */
/* BDExchangeMove_{f32,f64}(N, target_mols, params[N,4], temperature, nb_beta, cutoff, seed, num_proposals_per_move,
 * interval, batch_size=1)                                            wrap_kernels.cpp:1732-1795, bd_exchange_move.cu */
int tmb_bd_exchange_move_create(
    int precision, int N, const int *mol_atoms, const int *mol_offsets, int n_mols, const double *params, int n_params,
    double temperature, double nb_beta, double cutoff, int seed, int num_proposals_per_move, int interval, int batch_size,
    tmb_mover *out);
int tmb_bd_exchange_move_num_target_mols(tmb_mover m, int *out);
int tmb_bd_exchange_move_batch_size(tmb_mover m, int *out);                      /* :1899 */
/* compute_initial_log_weights(coords, box) -> [num_target_mols]       :1862-1873 */
int tmb_bd_exchange_move_initial_log_weights(tmb_mover m, int N, const double *coords, const double *box, double *out);
/* compute_incremental_log_weights(coords, box, mol_idxs[B], quaternions[B,4], translations[B,3]) -> [B, num_target_mols]
 * (translations are used as given, not scaled by the box)             :1818-1861 */
int tmb_bd_exchange_move_incremental_log_weights(
    tmb_mover m, int N, const double *coords, const double *box, const int *mol_idxs, const double *quaternions,
    const double *translations, double *out);
int tmb_bd_exchange_move_get_params(tmb_mover m, double *out, int n_params);    /* :1874-1882 */
int tmb_bd_exchange_move_set_params(tmb_mover m, const double *params, int n_params); /* :1883-1888 */
int tmb_bd_exchange_move_last_log_probability(tmb_mover m, double *out);        /* :1889-1892 */
int tmb_bd_exchange_move_last_raw_log_probability(tmb_mover m, double *out);    /* :1893 */
int tmb_bd_exchange_move_n_accepted(tmb_mover m, unsigned long long *out);      /* :1894 */
int tmb_bd_exchange_move_n_proposed(tmb_mover m, unsigned long long *out);      /* :1895 */
int tmb_bd_exchange_move_before_log_weights(tmb_mover m, double *out);          /* [num_target_mols]      :1897 */
int tmb_bd_exchange_move_after_log_weights(tmb_mover m, double *out);           /* [B * num_target_mols]  :1898 */
/* TIBDExchangeMove_{f32,f64}(N, ligand_idxs, target_mols, params, temperature, nb_beta, cutoff, radius, seed,
 * num_proposals_per_move, interval, batch_size=1): targeted insertion / biased deletion between a sphere of `radius`
 * around the ligand centroid and the rest of the box.  The tmb_bd_exchange_move_* accessors apply to it as well.
 *                                                                    wrap_kernels.cpp:1902-1975, tibd_exchange_move.cu */
int tmb_tibd_exchange_move_create(
    int precision, int N, const int *ligand_idxs, int n_ligand, const int *mol_atoms, const int *mol_offsets, int n_mols,
    const double *params, int n_params, double temperature, double nb_beta, double cutoff, double radius, int seed,
    int num_proposals_per_move, int interval, int batch_size, tmb_mover *out);
/* inner_and_outer_mols_{f32,f64}(center_atoms, coords, box, group_idxs, radius): flags[m] = 1 when the centroid of
 * molecule m lies within `radius` of the centroid of center_atoms     wrap_kernels.cpp:2032-2048, exchange.cu:11-63 */
int tmb_inner_and_outer_mols(
    int precision, const int *center_atoms, int n_center, int N, const double *coords, const double *box, const int *mol_atoms,
    const int *mol_offsets, int n_mols, double radius, int *flags);
/* translations_inside_and_outside_sphere_host_{f32,f64}(n, box, center[3], radius, seed) -> [n, 2, 3]
 *                                                                    wrap_kernels.cpp:2117-2141, translations.cu:8-42 */
int tmb_translations_inside_and_outside_sphere(
    int precision, int n_translations, const double *box, const double *center, double radius, int seed, double *out);
/* NonbondedMolEnergyPotential_{f32,f64}(N, target_mols, beta, cutoff).execute(coords, params, box) -> [n_mols] fixed
 * point energies (int128, 2^36).  coords == NULL: construct and validate only.   wrap_kernels.cpp:234-294 */
int tmb_nonbonded_mol_energies(
    int precision, int N, const int *mol_atoms, const int *mol_offsets, int n_mols, double beta, double cutoff,
    const double *coords, const double *params, const double *box, tmb_i128 *out);
/* atom_by_atom_energies_{f32,f64}(target_atoms[T], coords, params, box, nb_beta, cutoff) -> [T, N]   :2004-2030 */
int tmb_atom_by_atom_energies(
    int precision, int N, const int *target_atoms, int T, const double *coords, const double *params, const double *box,
    double nb_beta, double cutoff, double *out);
/* SegmentedSumExp_{f32,f64}(max_vals_per_segment, num_segments).logsumexp(values): segment s holds
 * values[offsets[s] .. offsets[s+1])                                 :1693-1730, segmented_sumexp.cu:69-117 */
int tmb_segmented_logsumexp(
    int precision, int max_vals_per_segment, int num_segments_max, const double *values, const int *offsets,
    int num_segments, double *out);
/* SegmentedWeightedRandomSampler_{f32,f64}(max_vals_per_segment, segments, seed).sample(weights) -> [segments]
 * (the cuRAND stream advances by max_vals_per_segment * segments uniforms per call, as in the reference)   :196-232 */
int tmb_weighted_sampler_create(int precision, int max_vals_per_segment, int num_segments, int seed, tmb_sampler *out);
int tmb_weighted_sampler_destroy(tmb_sampler s);
int tmb_weighted_sampler_sample(tmb_sampler s, const double *weights, const int *offsets, int num_segments, int *out);
/* rotate_coords_{f32,f64}(coords[N,3], quaternions[R,4]) -> [N, R, 3]                :2050-2070, rotations.cu:12-37 */
int tmb_rotate_coords(int precision, int N, int n_rotations, const double *coords, const double *quaternions, double *out);
/* rotate_and_translate_mol_{f32,f64}(coords[N,3], box, quaternions[B,4], translations[B,3]) -> [B, N, 3]; the
 * translations are fractions of the box                                              :2072-2115, rotations.cu:39-92 */
int tmb_rotate_and_translate_mol(
    int precision, int N, int batch_size, const double *coords, const double *box, const double *quaternions,
    const double *translations, double *out);

/* ---- Context(x0, v0, box, integrator, bps)                       wrap_kernels.cpp:296-689, context.cu ------------- */
int tmb_context_create(
    const double *x0, const double *v0, const double *box, int N, tmb_integrator intg, const tmb_bound_potential *bps,
    int n_bps, tmb_context *out);
/* same with movers=[...] (Context ctor, context.cu:28-50); movers run after every integrator step (context.cu:261-277) */
int tmb_context_create_with_movers(
    const double *x0, const double *v0, const double *box, int N, tmb_integrator intg, const tmb_bound_potential *bps,
    int n_bps, const tmb_mover *movers, int n_movers, tmb_context *out);
int tmb_context_destroy(tmb_context ctx);
int tmb_context_step(tmb_context ctx);
/* Context::initialize / finalize (context.cu:250-260): the integrator's opening / closing half step on the context's state */
int tmb_context_initialize(tmb_context ctx);
int tmb_context_finalize(tmb_context ctx);
/* Context::multiple_steps(n_steps, n_samples, h_x[n_samples,N,3], h_box[n_samples,3,3])   context.cu:216-242 */
int tmb_context_multiple_steps(tmb_context ctx, int n_steps, int n_samples, double *h_x, double *h_box);
/* Local MD: Context.setup_local_md / multiple_steps_local / multiple_steps_local_selection (wrap_kernels.cpp:368-612,
 * context.cu:90-214).  h_x [n_samples, N, 3], h_box [n_samples, 3, 3].  tmb_context_local_md_free_idxs: the free-atom
 * array of the last local call, out[i] = i if atom i was free else N (introspection for tests). */
int tmb_context_setup_local_md(tmb_context ctx, double temperature, int freeze_reference);
int tmb_context_multiple_steps_local(
    tmb_context ctx, int n_steps, const int *local_idxs, int n_local_idxs, int n_samples, double radius, double k, int seed,
    double *h_x, double *h_box);
int tmb_context_multiple_steps_local_selection(
    tmb_context ctx, int n_steps, int reference_idx, const int *selection_idxs, int n_selection_idxs, int n_samples,
    double radius, double k, double *h_x, double *h_box);
int tmb_context_local_md_free_idxs(tmb_context ctx, unsigned int *out);
int tmb_context_set_x_t(tmb_context ctx, const double *x);
int tmb_context_set_v_t(tmb_context ctx, const double *v);
int tmb_context_set_box(tmb_context ctx, const double *box);
int tmb_context_get_x_t(tmb_context ctx, double *x);
int tmb_context_get_v_t(tmb_context ctx, double *v);
int tmb_context_get_box(tmb_context ctx, double *box);
int tmb_context_num_atoms(tmb_context ctx, int *out);
int tmb_context_set_stream(tmb_context ctx, void *cuda_stream); /* run the MD loop on the caller's stream */
int tmb_context_set_use_graphs(tmb_context ctx, int on);        /* CUDA-graph replay of step blocks (default on) */
int tmb_context_device_state(tmb_context ctx, double **d_x, double **d_v, double **d_box);

/* ---- Neighborlist_{f32,f64}(N)                                   wrap_kernels.cpp:113-172 ------------------------- */
int tmb_neighborlist_create(int precision, int N, tmb_neighborlist *out);
int tmb_neighborlist_destroy(tmb_neighborlist nb);
/* get_nblist in two calls: build returns the sizes, fetch copies CSR (offsets[n_row_blocks+1], atoms[n_entries]) */
int tmb_neighborlist_build(
    tmb_neighborlist nb, int N, const double *coords, const double *box, double cutoff, int *n_row_blocks, int *n_entries);
int tmb_neighborlist_fetch(tmb_neighborlist nb, int32_t *offsets, int32_t *atoms);
int tmb_neighborlist_compute_block_bounds(
    tmb_neighborlist nb, int N, const double *coords, const double *box, double *ctrs, double *exts);
int tmb_neighborlist_set_row_idxs(tmb_neighborlist nb, const uint32_t *idxs, int n);
int tmb_neighborlist_reset_row_idxs(tmb_neighborlist nb);
int tmb_neighborlist_resize(tmb_neighborlist nb, int size);
int tmb_neighborlist_get_tile_ixn_count(tmb_neighborlist nb, unsigned int *out);
int tmb_neighborlist_get_max_ixn_count(tmb_neighborlist nb, int *out);
int tmb_neighborlist_get_num_row_idxs(tmb_neighborlist nb, int *out);

/* ---- HilbertSort(size).sort(coords, box) -> perm u32[N]          wrap_kernels.cpp:174-194 ------------------------- */
int tmb_hilbert_sort_create(int size, tmb_hilbert_sort *out);
int tmb_hilbert_sort_destroy(tmb_hilbert_sort hs);
int tmb_hilbert_sort_sort(tmb_hilbert_sort hs, int N, const double *coords, const double *box, uint32_t *perm);

/* ---- HREX host loop: a batch of neighbour-swap attempts, `_run_neighbor_swaps` of timemachine/md/hrex.py:50-129 --------
 * (a jitted lax.scan there).  Attempt i takes the state pair neighbor_pairs[pair_idxs[i]] = (s_a, s_b) held by replicas
 * (r_a, r_b) and is accepted iff uniform_samples[i] < exp(min(log_q[r_a,s_b] + log_q[r_b,s_a] - log_q[r_a,s_a] -
 * log_q[r_b,s_b], 0)); NaN differences reject.  replica_idx_by_state[n_states] is updated in place; proposed / accepted
 * [n_pairs] are overwritten.  Host only: runs without a GPU.  Every rank calls it on the all-gathered matrix. */
int tmb_hrex_run_neighbor_swaps(
    int n_states, int n_replicas, int n_pairs, const int32_t *neighbor_pairs, const double *log_q_kl, int n_attempts,
    const int32_t *pair_idxs, const double *uniform_samples, int32_t *replica_idx_by_state, uint32_t *proposed,
    uint32_t *accepted);

/* ---- test utilities ---------------------------------------------------------------------------------------------- */
/* N x 3 standard normals from the integrator's Philox stream for (seed, step) */
int tmb_fill_normal(float *out, int n_atoms, uint64_t seed, uint64_t step);

#ifdef __cplusplus
}
#endif

#endif /* TMB200_H */
