// Nonbonded 32x32 tile kernel: LJ 6-12 + switched-erfc electrostatics with 4D decoupling, fixed-point du/dx, du/dp
// and int128 energy.  Replaces reference k_nonbonded_unified / v_nonbonded_unified (k_nonbonded.cuh:109-432).
//
// B200 design:
//  * inputs are the packed, pre-cast Hilbert-ordered working set (two 128-bit loads per atom);
//  * a persistent grid sized from the SM count; every warp owns a CONTIGUOUS slice of the tile list.  The list is
//    emitted in runs of equal row block, so the row atoms and their accumulators stay in registers across a run and
//    the row-side atomics are issued once per run instead of once per tile;
//  * column state rotates around the warp with shuffles (no shared memory traffic, no divergence between lanes);
//  * accumulators are SoA in sorted order so the 64-bit reductions of a warp land on contiguous 256 B segments;
//  * energy: int128 per thread -> warp shuffle -> CTA -> last-CTA-done sum, no extra launch.
#include "fixed_point.cuh"
#include "kernels.hpp"
#include "nb_math.cuh"
#include "reduce.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace tmb {

constexpr int NB_THREADS = 256;
constexpr int NB_WARPS = NB_THREADS / WARP;
constexpr unsigned int NB_CHUNK = 4; // tiles claimed per atomicAdd

template <typename Real, bool WITH_DP> struct LaneAtom {
    Real x, y, z, w;
    Real q, sig, eps;
    u64 gx, gy, gz;         // du/dx accumulators (fixed point)
    u64 gq, gsig, geps, gw; // du/dp accumulators (fixed point)
};

template <typename Real, bool DP>
__device__ __forceinline__ void
load_lane_atom(LaneAtom<Real, DP> &a, const Vec4<Real> *__restrict__ xw, const Vec4<Real> *__restrict__ qse, int slot, bool valid) {
    Vec4<Real> c = {0, 0, 0, 0};
    Vec4<Real> p = {0, 0, 0, 0};
    if (valid) {
        c = xw[slot];
        p = qse[slot];
    }
    a.x = c.x;
    a.y = c.y;
    a.z = c.z;
    a.w = c.w;
    a.q = p.x;
    a.sig = p.y;
    a.eps = p.z;
    a.gx = a.gy = a.gz = 0;
    a.gq = a.gsig = a.geps = a.gw = 0;
}

template <typename Real, bool X, bool P>
__device__ __forceinline__ void flush_lane_atom(
    const LaneAtom<Real, P> &a, int slot, bool valid, const unsigned int *__restrict__ perm, u64 *__restrict__ du_dx,
    u64 *__restrict__ du_dp) {
    if (!valid) {
        return;
    }
    const size_t atom = perm[slot];
    if (X) {
        atomicAdd(du_dx + atom * 3 + 0, a.gx);
        atomicAdd(du_dx + atom * 3 + 1, a.gy);
        atomicAdd(du_dx + atom * 3 + 2, a.gz);
    }
    if (P) {
        atomicAdd(du_dp + atom * P_PER_ATOM + P_CHARGE, a.gq);
        atomicAdd(du_dp + atom * P_PER_ATOM + P_SIG, a.gsig);
        atomicAdd(du_dp + atom * P_PER_ATOM + P_EPS, a.geps);
        atomicAdd(du_dp + atom * P_PER_ATOM + P_W, a.gw);
    }
}

// One 32x32 tile: lane l holds row atom i_l for the whole tile and, in round r, the column atom that started in lane
// (l + r) % 32.  Every pair term is converted to fixed point before accumulation (SURVEY.md §8a cheat-sheet item 4).
template <typename Real, bool ALCH, bool U, bool X, bool P>
__device__ __forceinline__ void tile_rounds(
    const BoxCache<Real> &box,
    const Real cutoff2,
    const Real beta,
    const bool triangular,
    const int K,
    const bool i_valid,
    const int i_slot,
    LaneAtom<Real, P> &ai,
    int j_slot,
    LaneAtom<Real, P> &aj,
    i128 &energy) {
    const unsigned int full = 0xffffffffu;
    const int src = (threadIdx.x + 1) & 31;
#pragma unroll 2
    for (int round = 0; round < WARP; round++) {
        Real dx = min_image(ai.x - aj.x, box.x, box.inv_x);
        Real dy = min_image(ai.y - aj.y, box.y, box.inv_y);
        Real dz = min_image(ai.z - aj.z, box.z, box.inv_z);
        Real d2 = dist2_3d(dx, dy, dz);
        Real dw = 0;
        if (ALCH) {
            dw = ai.w - aj.w;
            d2 = fma_(dw, dw, d2);
        }
        const bool valid = i_valid && (j_slot < K) && (!triangular || i_slot < j_slot);
        // strict '<': atoms parked at w == cutoff must not interact (reference k_nonbonded.cuh:215-218)
        if (valid && d2 < cutoff2) {
            PairTerms<Real> t = pair_terms<Real, U>(
                static_cast<Real>(1), static_cast<Real>(1), ai.q, aj.q, ai.sig, aj.sig, ai.eps, aj.eps, d2, beta);
            if (X) {
                const Real rx = t.prefactor * dx, ry = t.prefactor * dy, rz = t.prefactor * dz;
                ai.gx += to_fixed_force(rx);
                ai.gy += to_fixed_force(ry);
                ai.gz += to_fixed_force(rz);
                // the column side converts -v like the reference (k_nonbonded.cuh:252-254): fixed(-v) == -fixed(v) inside the
                // int64 range (round-half-even is symmetric), and beyond it these are the reference's bits
                aj.gx += to_fixed_force(-rx);
                aj.gy += to_fixed_force(-ry);
                aj.gz += to_fixed_force(-rz);
            }
            if (P) {
                ai.gq += to_fixed<FIXED_EXPONENT_DU_DCHARGE>(aj.q * t.inv_d * t.damping);
                aj.gq += to_fixed<FIXED_EXPONENT_DU_DCHARGE>(ai.q * t.inv_d * t.damping);
                if (t.lj) {
                    u64 fs = to_fixed<FIXED_EXPONENT_DU_DSIG>(t.sig_grad);
                    ai.gsig += fs;
                    aj.gsig += fs;
                    ai.geps += to_fixed<FIXED_EXPONENT_DU_DEPS>(t.eps_grad * aj.eps);
                    aj.geps += to_fixed<FIXED_EXPONENT_DU_DEPS>(t.eps_grad * ai.eps);
                }
                if (ALCH) {
                    const Real vw = t.prefactor * dw;
                    ai.gw += to_fixed<FIXED_EXPONENT_DU_DW>(vw);
                    aj.gw += to_fixed<FIXED_EXPONENT_DU_DW>(-vw);
                }
            }
            if (U) {
                energy += energy_to_fixed<Real>(t.u);
            }
        }
        // rotate the column state one lane
        j_slot = __shfl_sync(full, j_slot, src);
        aj.x = __shfl_sync(full, aj.x, src);
        aj.y = __shfl_sync(full, aj.y, src);
        aj.z = __shfl_sync(full, aj.z, src);
        aj.q = __shfl_sync(full, aj.q, src);
        aj.sig = __shfl_sync(full, aj.sig, src);
        aj.eps = __shfl_sync(full, aj.eps, src);
        if (ALCH) {
            aj.w = __shfl_sync(full, aj.w, src);
        }
        if (X) {
            aj.gx = __shfl_sync(full, aj.gx, src);
            aj.gy = __shfl_sync(full, aj.gy, src);
            aj.gz = __shfl_sync(full, aj.gz, src);
        }
        if (P) {
            aj.gq = __shfl_sync(full, aj.gq, src);
            aj.gsig = __shfl_sync(full, aj.gsig, src);
            aj.geps = __shfl_sync(full, aj.geps, src);
            if (ALCH) {
                aj.gw = __shfl_sync(full, aj.gw, src);
            }
        }
    }
    // after 32 rotations every lane holds its original column atom again
}

template <typename Real, bool U, bool X, bool P>
__global__ void __launch_bounds__(NB_THREADS) k_nb_tiles(const NbTileArgs<Real> a) {
    __shared__ i128 scratch[NB_WARPS];

    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const unsigned int total_warps = gridDim.x * NB_WARPS;
    const unsigned int gwarp = blockIdx.x * NB_WARPS + warp_in_block;

    const BoxCache<Real> box = load_box<Real>(a.box);
    const Real cutoff = static_cast<Real>(a.cutoff);
    const Real cutoff2 = cutoff * cutoff;
    const Real beta = static_cast<Real>(a.beta);
    const bool triangular = (a.NR == a.K);

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (*a.rebuild_flag != 0) {
            a.rebuild_flag[2] += 1; // builds since construction (introspection only)
            *a.rebuild_flag = 0;
        }
    }
    const unsigned int T = min(*a.tile_count, a.tile_capacity);
    (void)total_warps;
    (void)gwarp;

    i128 energy = 0;
    LaneAtom<Real, P> ai;
    int cur_row = -1;
    int i_slot = 0;
    bool i_valid = false;

    // Dynamic scheduling: warps claim chunks of consecutive tiles from a device counter (reset by k_nb_prepare).  Tile
    // cost varies with fill by ~3x, a static split left ~25 % of the warp-slots idle in the tail (profiles/, round 1).
    // Consecutive tiles mostly share their row block, so the row atoms still stay in registers across a chunk.
    for (;;) {
    unsigned int chunk_begin = 0;
    if (lane == 0) {
        chunk_begin = atomicAdd(a.tile_cursor, NB_CHUNK);
    }
    chunk_begin = __shfl_sync(0xffffffffu, chunk_begin, 0);
    if (chunk_begin >= T) {
        break;
    }
    const unsigned int chunk_end = min(T, chunk_begin + NB_CHUNK);
    for (unsigned int t = chunk_begin; t < chunk_end; t++) {
        const int row = a.tile_rows[t];
        if (row != cur_row) {
            if (cur_row >= 0) {
                flush_lane_atom<Real, X, P>(ai, i_slot, i_valid, a.perm, a.du_dx, a.du_dp);
            }
            cur_row = row;
            i_slot = row * TILE + lane;
            i_valid = i_slot < a.NR;
            load_lane_atom(ai, a.xw, a.qse, i_slot, i_valid);
        }
        const int j_slot = static_cast<int>(min(a.tile_cols[t * TILE + lane], static_cast<unsigned int>(a.K)));
        const bool j_valid = j_slot < a.K;
        LaneAtom<Real, P> aj;
        load_lane_atom(aj, a.xw, a.qse, j_slot, j_valid);

        // tiles whose 64 atoms all sit at w == 0 skip the 4D terms; adding 0*0 is exact so results are identical
        const bool vanilla = __all_sync(0xffffffffu, ai.w == static_cast<Real>(0) && aj.w == static_cast<Real>(0));
        if (vanilla) {
            tile_rounds<Real, false, U, X, P>(box, cutoff2, beta, triangular, a.K, i_valid, i_slot, ai, j_slot, aj, energy);
        } else {
            tile_rounds<Real, true, U, X, P>(box, cutoff2, beta, triangular, a.K, i_valid, i_slot, ai, j_slot, aj, energy);
        }
        flush_lane_atom<Real, X, P>(aj, j_slot, j_valid, a.perm, a.du_dx, a.du_dp);
    }
    }
    if (cur_row >= 0) {
        flush_lane_atom<Real, X, P>(ai, i_slot, i_valid, a.perm, a.du_dx, a.du_dp);
    }

    if (U) {
        grid_finish_energy(energy, scratch, a.u_partials, a.ticket, a.d_u);
    }
}

template <typename Real> static int nb_tiles_grid_impl() {
    static int cached = 0;
    if (cached == 0) {
        int per_sm = 0;
        // the du/dx-only variant is the MD hot loop; size the persistent grid for it
        TMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, k_nb_tiles<Real, false, true, false>, NB_THREADS, 0));
        if (per_sm < 1) {
            per_sm = 1;
        }
        cached = sm_count() * per_sm;
    }
    return cached;
}

static bool use_async_staging() {
    static const bool on = [] {
        const char *e = std::getenv("TMB_NB_ASYNC");
        return e != nullptr && e[0] == '1';
    }();
    return on;
}

static bool use_ring_for_f32() {
    static const bool ring = [] {
        const char *e = std::getenv("TMB_NB_RING");
        return e != nullptr && e[0] == '1';
    }();
    return ring;
}

int nb_tiles_reserved_ctas() {
    static int cached = -1;
    if (cached < 0) {
        const char *env = std::getenv("TMB_NB_RESERVE");
        cached = env != nullptr ? std::max(0, std::atoi(env)) : 16;
    }
    return cached;
}

template <typename Real> int nb_tiles_max_grid() { return nb_tiles_grid_impl<Real>(); }
template <> int nb_tiles_max_grid<float>() { return std::max(nb_tiles_grid_impl<float>(), nb_tiles_cq_max_grid()); }

template <typename Real>
void launch_nb_tiles(const NbTileArgs<Real> &args, bool with_u, bool with_dx, bool with_dp, cudaStream_t stream) {
    if (std::is_same<Real, float>::value && !use_ring_for_f32()) {
        // f32 (the MD path): compaction-queue kernel; the ring kernel below stays as the f64 path and as an A/B reference
        if (use_async_staging()) {
            launch_nb_tiles_cq_async(reinterpret_cast<const NbTileArgs<float> &>(args), with_u, with_dx, with_dp, stream);
        } else {
            launch_nb_tiles_cq(reinterpret_cast<const NbTileArgs<float> &>(args), with_u, with_dx, with_dp, stream);
        }
        return;
    }
    int grid = nb_tiles_grid_impl<Real>();
    if (args.grid_ctas > 0) {
        grid = std::min(grid, args.grid_ctas);
    }
    const int sel = (with_u ? 4 : 0) | (with_dx ? 2 : 0) | (with_dp ? 1 : 0);
    switch (sel) {
    case 0:
        // nothing requested: the reference still launches an empty variant; there is nothing to compute
        break;
    case 1:
        TMB_LAUNCH((k_nb_tiles<Real, false, false, true>), grid, NB_THREADS, 0, stream, args);
        break;
    case 2:
        TMB_LAUNCH((k_nb_tiles<Real, false, true, false>), grid, NB_THREADS, 0, stream, args);
        break;
    case 3:
        TMB_LAUNCH((k_nb_tiles<Real, false, true, true>), grid, NB_THREADS, 0, stream, args);
        break;
    case 4:
        TMB_LAUNCH((k_nb_tiles<Real, true, false, false>), grid, NB_THREADS, 0, stream, args);
        break;
    case 5:
        TMB_LAUNCH((k_nb_tiles<Real, true, false, true>), grid, NB_THREADS, 0, stream, args);
        break;
    case 6:
        TMB_LAUNCH((k_nb_tiles<Real, true, true, false>), grid, NB_THREADS, 0, stream, args);
        break;
    case 7:
        TMB_LAUNCH((k_nb_tiles<Real, true, true, true>), grid, NB_THREADS, 0, stream, args);
        break;
    }
}

template int nb_tiles_max_grid<double>();
template void launch_nb_tiles<float>(const NbTileArgs<float> &, bool, bool, bool, cudaStream_t);
template void launch_nb_tiles<double>(const NbTileArgs<double> &, bool, bool, bool, cudaStream_t);

} // namespace tmb
