// f32 nonbonded tile kernel, "compaction queue" formulation (round-1 optimisation of k_nb_tiles.cu, same results bit
// for bit: every pair term goes through the same pair_terms<> and is rounded to fixed point before it is summed).
//
// Why: in the ring formulation (k_nb_tiles.cu, and the reference's v_nonbonded_unified) the ~95-instruction pair
// evaluation runs whenever ANY lane of the warp has a pair inside the cutoff, with only ~34 % of the lanes active
// (tile fill measured on the Hilbert-sorted list; ncu: 17.4 of 32 threads active per instruction, the
// kernel is issue-bound).  Here a tile is processed in two decoupled phases:
//   A. 32 cheap rounds: minimum-image distance of lane i against column (i + r) % 32, read from shared memory; hits are
//      appended (ballot + prefix popcount) to a per-warp ring queue {dx, dy, dz, d2, dw, i, j} in shared memory;
//   B. whenever 32 hits are queued, one fully populated warp evaluates them: parameters are fetched from the per-warp
//      shared copy of the 64 atoms, the force is converted to fixed point, split into three 22-bit limbs and added with
//      native 32-bit shared-memory atomics (no-return ATOMS.ADD: 1.0-1.8 SM-cycles per warp instruction on B200,
//      profiles/microbench_r1.txt; a 64-bit shared atomicAdd is a CAS loop at 9-16 cycles) to the row atom and,
//      negated, to the column atom.  Limb sums of <= 32 terms cannot overflow 32 bits; they are folded back into
//      64-bit registers / global reductions once per tile.
// The expensive instructions therefore run at ~95 % lane utilisation instead of ~34 %, and the XU-pipe work (MUFU,
// F2I) drops by the same factor.  Accumulation order changes, results do not (integer sums).
#include "fixed_point.cuh"
#include "kernels.hpp"
#include "nb_math.cuh"
#include "reduce.cuh"

namespace tmb {

constexpr int CQ_THREADS = 256;
constexpr int CQ_WARPS = CQ_THREADS / WARP;
constexpr unsigned int CQ_CHUNK = 4; // tiles claimed per atomicAdd on the tile cursor
constexpr int CQ_QUEUE = 64;         // ring capacity (a round appends <= 32, a batch removes 32)
constexpr int LIMB_BITS = 22;
constexpr unsigned int LIMB_MASK = (1u << LIMB_BITS) - 1u;

template <bool P> struct AccP {
    int v[4][3][64]; // [param][limb][atom]
};
template <> struct AccP<false> {};

template <bool P> struct alignas(16) WarpShared {
    // the tile's 64 atoms: [0,32) row block, [32,64) column atoms
    float x[64], y[64], z[64], w[64], q[64], sig[64], eps[64];
    int jslot[32];
    // fixed-point accumulators as 3 x 22-bit limbs (signed 32-bit sums)
    int accx[3][3][64]; // [component][limb][atom]
    AccP<P> accp;
    // ring queue of pairs inside the cutoff
    float qdx[CQ_QUEUE], qdy[CQ_QUEUE], qdz[CQ_QUEUE], qd2[CQ_QUEUE], qdw[CQ_QUEUE];
    unsigned short qidx[CQ_QUEUE];
};

__device__ __forceinline__ void limbs_add(int *acc /*[3][64]*/, int atom, u64 v) {
    const unsigned int lo = static_cast<unsigned int>(v) & LIMB_MASK;
    const unsigned int mid = static_cast<unsigned int>(v >> LIMB_BITS) & LIMB_MASK;
    const int hi = static_cast<int>(static_cast<i64>(v) >> (2 * LIMB_BITS));
    atomicAdd(acc + 0 * 64 + atom, static_cast<int>(lo));
    atomicAdd(acc + 1 * 64 + atom, static_cast<int>(mid));
    atomicAdd(acc + 2 * 64 + atom, hi);
}
__device__ __forceinline__ void limbs_sub(int *acc, int atom, u64 v) {
    const unsigned int lo = static_cast<unsigned int>(v) & LIMB_MASK;
    const unsigned int mid = static_cast<unsigned int>(v >> LIMB_BITS) & LIMB_MASK;
    const int hi = static_cast<int>(static_cast<i64>(v) >> (2 * LIMB_BITS));
    atomicAdd(acc + 0 * 64 + atom, -static_cast<int>(lo));
    atomicAdd(acc + 1 * 64 + atom, -static_cast<int>(mid));
    atomicAdd(acc + 2 * 64 + atom, -hi);
}
// fold the three signed limb sums of `atom` into a 64-bit value and clear them
__device__ __forceinline__ u64 limbs_take(int *acc, int atom) {
    const i64 lo = acc[0 * 64 + atom];
    const i64 mid = acc[1 * 64 + atom];
    const i64 hi = acc[2 * 64 + atom];
    acc[0 * 64 + atom] = 0;
    acc[1 * 64 + atom] = 0;
    acc[2 * 64 + atom] = 0;
    return static_cast<u64>(lo) + (static_cast<u64>(mid) << LIMB_BITS) + (static_cast<u64>(hi) << (2 * LIMB_BITS));
}

// Phase B: evaluate `count` (<= 32) queued pairs, one per lane.
template <bool ALCH, bool U, bool X, bool P>
__device__ __forceinline__ void cq_process(WarpShared<P> &S, const int head, const int count, const float beta, i128 &energy) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane < count) {
        const int k = (head + lane) & (CQ_QUEUE - 1);
        const float dx = S.qdx[k], dy = S.qdy[k], dz = S.qdz[k], d2 = S.qd2[k];
        const unsigned int idx = S.qidx[k];
        const int i = idx & 31;
        const int j = 32 + (idx >> 5);
        const float qi = S.q[i], qj = S.q[j];
        const float ei = S.eps[i], ej = S.eps[j];
        const PairTerms<float> t = pair_terms<float, U>(1.0f, 1.0f, qi, qj, S.sig[i], S.sig[j], ei, ej, d2, beta);
        if (X) {
            const u64 fx = to_fixed_force(t.prefactor * dx);
            const u64 fy = to_fixed_force(t.prefactor * dy);
            const u64 fz = to_fixed_force(t.prefactor * dz);
            limbs_add(&S.accx[0][0][0], i, fx);
            limbs_add(&S.accx[1][0][0], i, fy);
            limbs_add(&S.accx[2][0][0], i, fz);
            // fixed(-v) == -fixed(v): the column atom receives the exact negation
            limbs_sub(&S.accx[0][0][0], j, fx);
            limbs_sub(&S.accx[1][0][0], j, fy);
            limbs_sub(&S.accx[2][0][0], j, fz);
        }
        if (P) {
            AccP<true> &A = reinterpret_cast<AccP<true> &>(S.accp);
            limbs_add(&A.v[P_CHARGE][0][0], i, to_fixed<FIXED_EXPONENT_DU_DCHARGE>(qj * t.inv_d * t.damping));
            limbs_add(&A.v[P_CHARGE][0][0], j, to_fixed<FIXED_EXPONENT_DU_DCHARGE>(qi * t.inv_d * t.damping));
            if (t.lj) {
                const u64 fs = to_fixed<FIXED_EXPONENT_DU_DSIG>(t.sig_grad);
                limbs_add(&A.v[P_SIG][0][0], i, fs);
                limbs_add(&A.v[P_SIG][0][0], j, fs);
                limbs_add(&A.v[P_EPS][0][0], i, to_fixed<FIXED_EXPONENT_DU_DEPS>(t.eps_grad * ej));
                limbs_add(&A.v[P_EPS][0][0], j, to_fixed<FIXED_EXPONENT_DU_DEPS>(t.eps_grad * ei));
            }
            if (ALCH) {
                const u64 fw = to_fixed<FIXED_EXPONENT_DU_DW>(t.prefactor * S.qdw[k]);
                limbs_add(&A.v[P_W][0][0], i, fw);
                limbs_sub(&A.v[P_W][0][0], j, fw);
            }
        }
        if (U) {
            energy += energy_to_fixed<float>(t.u);
        }
    }
    __syncwarp();
}

// Phase A + B for one tile whose atoms are already in S.
template <bool ALCH, bool DIAG, bool U, bool X, bool P>
__device__ __forceinline__ void cq_tile(
    WarpShared<P> &S, const BoxCache<float> &box, const float cutoff2, const float beta, const bool i_valid, const int i_slot,
    const unsigned int jvalid_mask, i128 &energy) {
    const int lane = threadIdx.x & 31;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const float xi = S.x[lane], yi = S.y[lane], zi = S.z[lane];
    const float wi = ALCH ? S.w[lane] : 0.0f;
    int head = 0;  // ring position of the oldest queued pair
    int count = 0; // queued pairs
#pragma unroll 2
    for (int round = 0; round < WARP; round++) {
        const int jp = (lane + round) & 31;
        const float dx = min_image(xi - S.x[32 + jp], box.x, box.inv_x);
        const float dy = min_image(yi - S.y[32 + jp], box.y, box.inv_y);
        const float dz = min_image(zi - S.z[32 + jp], box.z, box.inv_z);
        float d2 = dist2_3d(dx, dy, dz);
        float dw = 0.0f;
        if (ALCH) {
            dw = wi - S.w[32 + jp];
            d2 = fma_(dw, dw, d2);
        }
        bool hit = i_valid && ((jvalid_mask >> jp) & 1u) && (d2 < cutoff2); // strict '<', see k_nb_tiles.cu
        if (DIAG) {
            hit = hit && (i_slot < S.jslot[jp]); // all-pairs: count each pair once
        }
        const unsigned int ballot = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int k = (head + count + __popc(ballot & lt_mask)) & (CQ_QUEUE - 1);
            S.qdx[k] = dx;
            S.qdy[k] = dy;
            S.qdz[k] = dz;
            S.qd2[k] = d2;
            if (ALCH) {
                S.qdw[k] = dw;
            }
            S.qidx[k] = static_cast<unsigned short>(lane | (jp << 5));
        }
        count += __popc(ballot);
        if (count >= WARP) {
            cq_process<ALCH, U, X, P>(S, head, WARP, beta, energy);
            head = (head + WARP) & (CQ_QUEUE - 1);
            count -= WARP;
        }
    }
    if (count > 0) {
        cq_process<ALCH, U, X, P>(S, head, count, beta, energy);
    }
}

template <bool U, bool X, bool P> __global__ void __launch_bounds__(CQ_THREADS) k_nb_tiles_cq(const NbTileArgs<float> a) {
    extern __shared__ __align__(16) unsigned char cq_smem_raw[];
    __shared__ i128 scratch[CQ_WARPS];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpShared<P> &S = reinterpret_cast<WarpShared<P> *>(cq_smem_raw)[warp];

    const BoxCache<float> box = load_box<float>(a.box);
    const float cutoff = static_cast<float>(a.cutoff);
    const float cutoff2 = cutoff * cutoff;
    const float beta = static_cast<float>(a.beta);
    const bool triangular = (a.NR == a.K);

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *a.rebuild_flag = 0;
    }
    // clear this warp's limb accumulators
    for (int c = 0; c < 9; c++) {
        (&S.accx[0][0][0])[c * 64 + lane] = 0;
        (&S.accx[0][0][0])[c * 64 + 32 + lane] = 0;
    }
    if (P) {
        AccP<true> &A = reinterpret_cast<AccP<true> &>(S.accp);
        for (int c = 0; c < 12; c++) {
            (&A.v[0][0][0])[c * 64 + lane] = 0;
            (&A.v[0][0][0])[c * 64 + 32 + lane] = 0;
        }
    }
    __syncwarp();

    const unsigned int T = min(*a.tile_count, a.tile_capacity);
    i128 energy = 0;
    int cur_row = -1;
    int i_slot = 0;
    bool i_valid = false;
    u64 gi[3] = {0, 0, 0};    // row-atom du/dx accumulated over a run of tiles with the same row block
    u64 gpi[4] = {0, 0, 0, 0}; // row-atom du/dp

    auto flush_row = [&]() {
        if (cur_row >= 0 && i_valid) {
            if (X) {
                atomicAdd(a.acc_dx + 0 * a.Kpad + i_slot, gi[0]);
                atomicAdd(a.acc_dx + 1 * a.Kpad + i_slot, gi[1]);
                atomicAdd(a.acc_dx + 2 * a.Kpad + i_slot, gi[2]);
            }
            if (P) {
                for (int c = 0; c < 4; c++) {
                    atomicAdd(a.acc_dp + c * a.Kpad + i_slot, gpi[c]);
                }
            }
        }
        gi[0] = gi[1] = gi[2] = 0;
        gpi[0] = gpi[1] = gpi[2] = gpi[3] = 0;
    };

    for (;;) {
        unsigned int chunk_begin = 0;
        if (lane == 0) {
            chunk_begin = atomicAdd(a.tile_cursor, CQ_CHUNK);
        }
        chunk_begin = __shfl_sync(0xffffffffu, chunk_begin, 0);
        if (chunk_begin >= T) {
            break;
        }
        const unsigned int chunk_end = min(T, chunk_begin + CQ_CHUNK);
        for (unsigned int t = chunk_begin; t < chunk_end; t++) {
            const int row = a.tile_rows[t];
            if (row != cur_row) {
                flush_row();
                cur_row = row;
                i_slot = row * TILE + lane;
                i_valid = i_slot < a.NR;
                Vec4<float> c = {0.f, 0.f, 0.f, 0.f}, p = {0.f, 0.f, 0.f, 0.f};
                if (i_valid) {
                    c = a.xw[i_slot];
                    p = a.qse[i_slot];
                }
                S.x[lane] = c.x;
                S.y[lane] = c.y;
                S.z[lane] = c.z;
                S.w[lane] = c.w;
                S.q[lane] = p.x;
                S.sig[lane] = p.y;
                S.eps[lane] = p.z;
            }
            const int j_slot = static_cast<int>(min(a.tile_cols[t * TILE + lane], static_cast<unsigned int>(a.K)));
            const bool j_valid = j_slot < a.K;
            {
                Vec4<float> c = {0.f, 0.f, 0.f, 0.f}, p = {0.f, 0.f, 0.f, 0.f};
                if (j_valid) {
                    c = a.xw[j_slot];
                    p = a.qse[j_slot];
                }
                S.x[32 + lane] = c.x;
                S.y[32 + lane] = c.y;
                S.z[32 + lane] = c.z;
                S.w[32 + lane] = c.w;
                S.q[32 + lane] = p.x;
                S.sig[32 + lane] = p.y;
                S.eps[32 + lane] = p.z;
                S.jslot[lane] = j_slot;
            }
            const unsigned int jvalid_mask = __ballot_sync(0xffffffffu, j_valid);
            // 4D terms only where some atom of the tile has w != 0 (adding 0*0 is exact: identical results)
            const bool vanilla = __all_sync(0xffffffffu, S.w[lane] == 0.0f && S.w[32 + lane] == 0.0f);
            // the i < j rule only bites when the column atoms overlap the row block itself
            const bool diag = triangular && __any_sync(0xffffffffu, j_valid && j_slot < (row + 1) * TILE);
            __syncwarp();
            if (vanilla) {
                if (diag) {
                    cq_tile<false, true, U, X, P>(S, box, cutoff2, beta, i_valid, i_slot, jvalid_mask, energy);
                } else {
                    cq_tile<false, false, U, X, P>(S, box, cutoff2, beta, i_valid, i_slot, jvalid_mask, energy);
                }
            } else {
                if (diag) {
                    cq_tile<true, true, U, X, P>(S, box, cutoff2, beta, i_valid, i_slot, jvalid_mask, energy);
                } else {
                    cq_tile<true, false, U, X, P>(S, box, cutoff2, beta, i_valid, i_slot, jvalid_mask, energy);
                }
            }
            // fold this tile's limb sums: row atoms into registers, column atoms straight to the sorted accumulators
            if (X) {
                for (int c = 0; c < 3; c++) {
                    gi[c] += limbs_take(&S.accx[c][0][0], lane);
                    const u64 gj = limbs_take(&S.accx[c][0][0], 32 + lane);
                    if (j_valid && gj != 0) {
                        atomicAdd(a.acc_dx + c * a.Kpad + j_slot, gj);
                    }
                }
            }
            if (P) {
                AccP<true> &A = reinterpret_cast<AccP<true> &>(S.accp);
                for (int c = 0; c < 4; c++) {
                    gpi[c] += limbs_take(&A.v[c][0][0], lane);
                    const u64 gj = limbs_take(&A.v[c][0][0], 32 + lane);
                    if (j_valid && gj != 0) {
                        atomicAdd(a.acc_dp + c * a.Kpad + j_slot, gj);
                    }
                }
            }
            __syncwarp();
        }
    }
    flush_row();

    if (U) {
        grid_finish_energy(energy, scratch, a.u_partials, a.ticket, a.d_u);
    }
}

template <bool U, bool X, bool P> static void cq_launch(const NbTileArgs<float> &args, int grid, cudaStream_t stream) {
    const size_t smem = CQ_WARPS * sizeof(WarpShared<P>);
    static bool configured = false;
    if (!configured) {
        TMB_CUDA(cudaFuncSetAttribute(k_nb_tiles_cq<U, X, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured = true;
    }
    TMB_LAUNCH((k_nb_tiles_cq<U, X, P>), grid, CQ_THREADS, smem, stream, args);
}

int nb_tiles_cq_max_grid() {
    static int cached = 0;
    if (cached == 0) {
        const size_t smem = CQ_WARPS * sizeof(WarpShared<false>);
        TMB_CUDA(cudaFuncSetAttribute(
            k_nb_tiles_cq<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        int per_sm = 0;
        TMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_nb_tiles_cq<false, true, false>, CQ_THREADS, smem));
        cached = sm_count() * (per_sm < 1 ? 1 : per_sm);
    }
    return cached;
}

void launch_nb_tiles_cq(const NbTileArgs<float> &args, bool with_u, bool with_dx, bool with_dp, cudaStream_t stream) {
    const int grid = nb_tiles_cq_max_grid();
    const int sel = (with_u ? 4 : 0) | (with_dx ? 2 : 0) | (with_dp ? 1 : 0);
    switch (sel) {
    case 0:
        break;
    case 1:
        cq_launch<false, false, true>(args, grid, stream);
        break;
    case 2:
        cq_launch<false, true, false>(args, grid, stream);
        break;
    case 3:
        cq_launch<false, true, true>(args, grid, stream);
        break;
    case 4:
        cq_launch<true, false, false>(args, grid, stream);
        break;
    case 5:
        cq_launch<true, false, true>(args, grid, stream);
        break;
    case 6:
        cq_launch<true, true, false>(args, grid, stream);
        break;
    case 7:
        cq_launch<true, true, true>(args, grid, stream);
        break;
    }
}

} // namespace tmb
