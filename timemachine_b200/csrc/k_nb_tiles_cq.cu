// f32 nonbonded tile kernel, "compaction queue" formulation (round-1 optimisation of k_nb_tiles.cu, same results bit
// for bit: every pair term goes through the same pair_terms<> and is rounded to fixed point before it is summed).
//
// Why: in the ring formulation (k_nb_tiles.cu, and the reference's v_nonbonded_unified) the ~95-instruction pair
// evaluation runs whenever ANY lane of the warp has a pair inside the cutoff, with only ~34 % of the lanes active
// (tile fill measured on the Hilbert-sorted list; ncu: 17.4 of 32 threads active per instruction, the
// kernel is issue-bound).  Here a tile is processed in two decoupled phases:
//   A. 32 cheap rounds: minimum-image distance of lane i against column (i + r) % 32, read from shared memory; hits are
//      appended (ballot + prefix popcount) to a per-warp ring queue {dx, dy, dz, i|j} (+ dw) in shared memory with one
//      16-byte store;
//   B. whenever 32 hits are queued, one fully populated warp evaluates them: parameters are fetched from the per-warp
//      shared copy of the 64 atoms, the force is converted to fixed point, split into a 27-bit low limb and a signed
//      high limb and added with native 32-bit shared-memory atomics (no-return ATOMS.ADD: 1.0-1.8 SM-cycles per warp
//      instruction on B200, profiles/microbench_r1.txt; a 64-bit shared atomicAdd is a CAS loop at 9-16 cycles) to the
//      row atom's and the column atom's accumulators.  Two limbs hold any term with |v| < 2^53 (|force| < 1.3e5
//      kJ/mol/nm per pair) and sums of <= 32 such terms cannot overflow; the rare larger term (clashing atoms) goes
//      straight to the global 64-bit accumulator.  Limb sums are folded back into 64-bit registers / global
//      reductions once per tile (the column side is negated there: fixed(-v) == -fixed(v)).
// The expensive instructions therefore run at ~90 % lane utilisation instead of ~34 %, and the XU-pipe work (MUFU,
// F2I) drops by the same factor.  Accumulation order changes, results do not (integer sums).
//
// Invalid atoms (padding of the last block / of a partial tile) are given NaN coordinates in the shared copy: every
// comparison d2 < cutoff2 is then false, so no validity logic is needed in the inner loop.
//
// Prefilter (phase A of most tiles): the 32 x 32 distance tests do not need f32.  For a tile with no 4D atoms and no
// i < j rule, the 64 atoms are re-expressed relative to row atom 0 (minimum image, f32), rounded to half precision and
// tested two columns at a time with packed f16x2 arithmetic against a slightly LARGER threshold; only the 5-bit (i, j)
// codes of the candidates are queued, and phase B recomputes the exact f32 minimum-image displacement from the exact
// shared copy and applies the reference's strict d2 < cutoff2 before evaluating.  The set of evaluated pairs and every
// evaluated term are therefore bit-identical to the exact path; the prefilter only has to be conservative:
//   * same periodic image: all |rel| < A = b/2 - cutoff/2 - 0.01 per axis (checked per tile with one vote), so
//     |rel_i - rel_j| < b - cutoff - 0.02 and the difference of the relative coordinates IS the minimum-image vector of
//     any pair closer than the cutoff;
//   * rounding: |rel| < 4 -> conversion error <= 2^-10 per coordinate; a half difference below 2 rounds with <= 2^-11;
//     so each component is off by <= 0.00245, which moves d2 by <= 0.00245 (2 sqrt(3) 1.5 + 0.0074) = 0.0128 for
//     d <= 1.5 nm; the three half roundings of the squared sum add <= 3 * 2^-10.  Total < 0.0158: the threshold is
//     cutoff^2 + 0.02 rounded up to half, and the prefilter is only used for cutoff <= 1.5 nm (else: exact path).
// The exact 38-instruction round per 32 pairs becomes ~29 instructions per 64 pairs (ncu r1 -> r2 in profiles/).
#include <algorithm>

#include <cuda_fp16.h>

#include "fixed_point.cuh"
#include "kernels.hpp"
#include "nb_math.cuh"
#include "reduce.cuh"

namespace tmb {

#ifndef CQ_MIN_CTAS
#define CQ_MIN_CTAS 4
#endif
constexpr int CQ_THREADS = 256;
constexpr int CQ_WARPS = CQ_THREADS / WARP;
// Scheduling: every warp first takes args.static_tiles consecutive tiles (runs of equal row block stay together), the
// rest of the list is handed out one tile at a time from the device cursor.  A warp only processes ~6 tiles per launch
// at the 30k-atom benchmark size, so coarser dynamic chunks leave a long tail (ncu r1: 41 % of warp slots active).
constexpr int CQ_QUEUE = 64; // ring capacity (a round appends <= 32, a batch removes 32)
constexpr int LIMB_BITS = 27;
constexpr unsigned int LIMB_MASK = (1u << LIMB_BITS) - 1u;

// Per-warp shared memory, as offsets (in 4-byte words) into one flat array.
// atoms: [0,32) row block, [32,64) column atoms
constexpr int S_X = 0, S_Y = 64, S_Z = 128, S_W = 192, S_Q = 256, S_SIG = 320, S_EPS = 384;
constexpr int S_JSLOT = 448;                // int[32]
constexpr int S_ACCX = 480;                 // int[3 comps][2 limbs][64 atoms]: rows [0,32) get +, columns [32,64) get + (negated at fold)
constexpr int S_Q4 = S_ACCX + 6 * 64;       // queue: float4 {dx, dy, dz, bits(i | j << 5)} (16-byte aligned: 864 words)
constexpr int S_QDW = S_Q4 + 4 * CQ_QUEUE;  // queue: dw (alchemical tiles only)
// prefilter tiles reuse the (then idle) exact queue for the half-precision column coordinates: f16x2 words, per
// component the 32 pairs P[k] = {c_k, c_k+1 mod 32} stored twice in a row (64 words), so that lane i reads the columns
// (i + 2R, i + 2R + 1) of double round R as word i + 2R: 32 distinct banks, no wrap
constexpr int S_H2 = S_Q4;                  // [3 comps][64 words]
static_assert(S_H2 + 192 <= S_QDW + CQ_QUEUE, "prefilter coordinates must fit in the exact queue");
// and their own linear queue of 16-bit candidate codes (i + (j << 5), j not reduced mod 32): half a tile (8 double
// rounds) appends at most 512 behind at most 31 left over from the previous half
constexpr int CQ_CODES = 544;
constexpr int S_QC = S_QDW + CQ_QUEUE;      // u16[CQ_CODES] = 272 words
constexpr int S_ACCP = S_QC + CQ_CODES / 2;    // int[4 params][2 limbs][64 atoms] (du/dp variants only)
constexpr int S_WORDS_X = S_ACCP;           // 1456 words = 5824 B per warp
constexpr int S_WORDS_P = S_ACCP + 8 * 64;  // 1968 words = 7872 B per warp (3 CTAs per SM)
static_assert(S_Q4 % 4 == 0 && S_WORDS_X % 4 == 0 && S_WORDS_P % 4 == 0, "float4 queue alignment");

// |v| < 2^53 as a signed 64-bit number (tested on the high word only)
__device__ __forceinline__ bool limb_small(u64 v) {
    const int hi32 = static_cast<int>(v >> 32);
    return static_cast<unsigned int>(hi32 + (1 << 21)) < (1u << 22);
}
__device__ __forceinline__ void limb_add(int *acc /*[2][64]*/, int atom, u64 v) {
    atomicAdd(acc + atom, static_cast<int>(static_cast<unsigned int>(v) & LIMB_MASK));
    atomicAdd(acc + 64 + atom, static_cast<int>(static_cast<i64>(v) >> LIMB_BITS));
}
// fold the two limb sums of `atom` into a 64-bit value and clear them
__device__ __forceinline__ u64 limbs_take(int *acc, int atom) {
    const unsigned int lo = static_cast<unsigned int>(acc[atom]);
    const i64 hi = acc[64 + atom];
    acc[atom] = 0;
    acc[64 + atom] = 0;
    return static_cast<u64>(lo) + (static_cast<u64>(hi) << LIMB_BITS);
}

// Where the rare term too large for the limb accumulators goes (the global sorted-order accumulators).
struct CqSink {
    const unsigned int *perm;
    u64 *du_dx;
    u64 *du_dp;
    int row_base; // sorted slot of row atom 0 of the current tile
};

// One pair inside the cutoff: evaluate, round every term to fixed point, add the limbs to the row atom i (0..31) and
// the column atom j (32..63) of the warp's shared block.
template <bool ALCH, bool U, bool X, bool P>
__device__ __forceinline__ void cq_pair(
    float *S, const int i, const int j, const float dx, const float dy, const float dz, const float dw, const float d2,
    const float beta, const CqSink &sink, i128 &energy) {
    int *SI = reinterpret_cast<int *>(S);
    const float qi = S[S_Q + i], qj = S[S_Q + j];
    const float ei = S[S_EPS + i], ej = S[S_EPS + j];
    const PairTerms<float> t = pair_terms<float, U>(1.0f, 1.0f, qi, qj, S[S_SIG + i], S[S_SIG + j], ei, ej, d2, beta);
    const int si = sink.row_base + i; // sorted slots, translated to atoms on the (rare) direct path only
    const int sj = SI[S_JSLOT + j - 32];
    if (X) {
        const float rx = t.prefactor * dx, ry = t.prefactor * dy, rz = t.prefactor * dz;
        const u64 fx = to_fixed_force(rx);
        const u64 fy = to_fixed_force(ry);
        const u64 fz = to_fixed_force(rz);
        int *acc = SI + S_ACCX;
        // all three fixed-point values below 2^52 in magnitude (two limbs hold 2^53); NaN / inf compare false
        if (fabsf(rx) + fabsf(ry) + fabsf(rz) < 65536.0f) {
            limb_add(acc + 0 * 128, i, fx);
            limb_add(acc + 1 * 128, i, fy);
            limb_add(acc + 2 * 128, i, fz);
            // the column atom receives the exact negation; it is accumulated positively and negated once at the fold
            limb_add(acc + 0 * 128, j, fx);
            limb_add(acc + 1 * 128, j, fy);
            limb_add(acc + 2 * 128, j, fz);
        } else {
            // clashing atoms: too large for two limbs, add to the global accumulators directly
            u64 *gi = sink.du_dx + static_cast<size_t>(sink.perm[si]) * 3;
            u64 *gj = sink.du_dx + static_cast<size_t>(sink.perm[sj]) * 3;
            atomicAdd(gi + 0, fx);
            atomicAdd(gi + 1, fy);
            atomicAdd(gi + 2, fz);
            atomicAdd(gj + 0, 0ull - fx);
            atomicAdd(gj + 1, 0ull - fy);
            atomicAdd(gj + 2, 0ull - fz);
        }
    }
    if (P) {
        int *acc = SI + S_ACCP;
        const u64 pqi = to_fixed<FIXED_EXPONENT_DU_DCHARGE>(qj * t.inv_d * t.damping);
        const u64 pqj = to_fixed<FIXED_EXPONENT_DU_DCHARGE>(qi * t.inv_d * t.damping);
        u64 psig = 0, pei = 0, pej = 0, pw = 0;
        if (t.lj) {
            psig = to_fixed<FIXED_EXPONENT_DU_DSIG>(t.sig_grad);
            pei = to_fixed<FIXED_EXPONENT_DU_DEPS>(t.eps_grad * ej);
            pej = to_fixed<FIXED_EXPONENT_DU_DEPS>(t.eps_grad * ei);
        }
        if (ALCH) {
            pw = to_fixed<FIXED_EXPONENT_DU_DW>(t.prefactor * dw); // antisymmetric: the column atom gets -pw
        }
        if (limb_small(pqi) && limb_small(pqj) && limb_small(psig) && limb_small(pei) && limb_small(pej) &&
            limb_small(pw)) {
            limb_add(acc + P_CHARGE * 128, i, pqi);
            limb_add(acc + P_CHARGE * 128, j, pqj);
            if (t.lj) {
                limb_add(acc + P_SIG * 128, i, psig);
                limb_add(acc + P_SIG * 128, j, psig);
                limb_add(acc + P_EPS * 128, i, pei);
                limb_add(acc + P_EPS * 128, j, pej);
            }
            if (ALCH) {
                limb_add(acc + P_W * 128, i, pw);
                limb_add(acc + P_W * 128, j, 0ull - pw);
            }
        } else {
            u64 *gi = sink.du_dp + static_cast<size_t>(sink.perm[si]) * P_PER_ATOM;
            u64 *gj = sink.du_dp + static_cast<size_t>(sink.perm[sj]) * P_PER_ATOM;
            atomicAdd(gi + P_CHARGE, pqi);
            atomicAdd(gj + P_CHARGE, pqj);
            atomicAdd(gi + P_SIG, psig);
            atomicAdd(gj + P_SIG, psig);
            atomicAdd(gi + P_EPS, pei);
            atomicAdd(gj + P_EPS, pej);
            atomicAdd(gi + P_W, pw);
            atomicAdd(gj + P_W, 0ull - pw);
        }
    }
    if (U) {
        energy += energy_to_fixed<float>(t.u);
    }
}

// Phase B: evaluate `count` (<= 32) queued pairs, one per lane.
template <bool ALCH, bool U, bool X, bool P>
__device__ __forceinline__ void cq_process(
    float *S, const int head, const int count, const float beta, const CqSink &sink, i128 &energy) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane < count) {
        const int k = (head + lane) & (CQ_QUEUE - 1);
        const float4 item = reinterpret_cast<const float4 *>(S + S_Q4)[k];
        const float dx = item.x, dy = item.y, dz = item.z;
        const int idx = __float_as_int(item.w);
        // same expression as phase A: the queue does not carry d2
        float d2 = dist2_3d(dx, dy, dz);
        float dw = 0.0f;
        if (ALCH) {
            dw = S[S_QDW + k];
            d2 = fma_(dw, dw, d2);
        }
        cq_pair<ALCH, U, X, P>(S, idx & 31, 32 + (idx >> 5), dx, dy, dz, dw, d2, beta, sink, energy);
    }
    __syncwarp();
}

struct CqBox {
    float bx, by, bz, inv_bx, inv_by, inv_bz;
};

// Phase B of a prefilter tile: one queued candidate code per lane (`q` points at this lane's entry, `active` says
// whether there is one).  The exact f32 displacement is formed here with the very expressions of the exact phase A
// (min_image, dist2_3d) and the reference's strict test.
template <bool U, bool X, bool P>
__device__ __forceinline__ void cq_process_codes(
    float *S, const unsigned short *q, const bool active, const CqBox &b, const float cutoff2, const float beta,
    const CqSink &sink, i128 &energy) {
    if (active) {
        const int code = *q;
        const int i = code & 31;
        const int j = 32 + ((code >> 5) & 31);
        const float dx = min_image(S[S_X + i] - S[S_X + j], b.bx, b.inv_bx);
        const float dy = min_image(S[S_Y + i] - S[S_Y + j], b.by, b.inv_by);
        const float dz = min_image(S[S_Z + i] - S[S_Z + j], b.bz, b.inv_bz);
        const float d2 = dist2_3d(dx, dy, dz);
        if (d2 < cutoff2) {
            cq_pair<false, U, X, P>(S, i, j, dx, dy, dz, 0.0f, d2, beta, sink, energy);
        }
    }
}

__device__ __forceinline__ unsigned int h2_bits(const __half2 v) { return *reinterpret_cast<const unsigned int *>(&v); }
__device__ __forceinline__ __half2 bits_h2(const unsigned int v) { return *reinterpret_cast<const __half2 *>(&v); }

// Prefilter tile: phase A in packed half precision on the relative coordinates staged at S_H2 (see the file header),
// two columns per lane and round, in halves of 8 double rounds: the 8 rounds are fully unrolled (immediate offsets, no
// evaluation code inside, few live registers) and append their candidates to the linear queue; then full batches are
// evaluated and the < 32 left over move to the front of the queue for the next half.  Halves [h0, h1) are processed:
// {0, 2} is the whole tile, {h, h + 1} one half of a split tile; the last half also evaluates the partial batch.
// hx/hy/hz: this lane's row atom, each value duplicated in both halves.  thr2: the enlarged threshold, duplicated.
template <bool U, bool X, bool P>
__device__ __forceinline__ void cq_tile_prefilter(
    float *S, const CqBox &b, const float cutoff2, const unsigned int thr2, const float beta, const __half2 hx,
    const __half2 hy, const __half2 hz, const int h0, const int h1, const CqSink &sink, i128 &energy) {
    const int lane = threadIdx.x & 31;
    unsigned short *Q = reinterpret_cast<unsigned short *>(S + S_QC);
    int count = 0; // queued candidates, at [0, count)
#pragma unroll 1
    for (int h = h0; h < h1; h++) {
        {
            const unsigned int lane_bit = 1u << lane;
            const unsigned int lt_mask = lane_bit - 1u;
            const unsigned int *H = reinterpret_cast<const unsigned int *>(S) + S_H2 + lane + 16 * h;
            const int code = lane + ((lane + 16 * h) << 5); // i + (j << 5), first column of round 0; j mod 32 later
            unsigned short *tail = Q + count;
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const __half2 dx = __hsub2(hx, bits_h2(H[2 * r]));
                const __half2 dy = __hsub2(hy, bits_h2(H[64 + 2 * r]));
                const __half2 dz = __hsub2(hz, bits_h2(H[128 + 2 * r]));
                const __half2 d2 = __hfma2(dz, dz, __hfma2(dy, dy, __hmul2(dx, dx)));
                unsigned int b0, b1; // lanes whose first / second column passes (NaN padding compares false)
                asm volatile(
                    "{\n\t"
                    ".reg .pred p, q;\n\t"
                    "setp.lt.f16x2 p|q, %2, %3;\n\t"
                    "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
                    "vote.sync.ballot.b32 %1, q, 0xffffffff;\n\t"
                    "}"
                    : "=r"(b0), "=r"(b1)
                    : "r"(h2_bits(d2)), "r"(thr2));
                const int n0 = __popc(b0);
                if (b0 & lane_bit) {
                    tail[__popc(b0 & lt_mask)] = static_cast<unsigned short>(code + 64 * r);
                }
                if (b1 & lane_bit) {
                    tail[n0 + __popc(b1 & lt_mask)] = static_cast<unsigned short>(code + 64 * r + 32);
                }
                tail += n0 + __popc(b1);
            }
            count = static_cast<int>(tail - Q);
        }
        __syncwarp();
        // full batches; the last half also evaluates the partial one (one call site: one inlined copy of the evaluation)
        const bool last = (h == h1 - 1);
        const int end = last ? count : (count & ~(WARP - 1));
        const unsigned short *q = Q + lane;
        for (int base = 0; base < end; base += WARP, q += WARP) {
            cq_process_codes<U, X, P>(S, q, base + lane < end, b, cutoff2, beta, sink, energy);
        }
        const int head = end; // entries consumed
        count -= end;
        if (!last && head > 0 && count > 0) {
            // what is left (< 32) moves to the front; every lane has finished reading [0, head) once this converges
            __syncwarp();
            const unsigned short v = Q[head + (lane < count ? lane : 0)];
            __syncwarp();
            if (lane < count) {
                Q[lane] = v;
            }
        }
    }
    __syncwarp();
}

// Phase A + B for one tile whose atoms are already in the warp's shared block.
// ROUNDS == 32: the whole tile.  ROUNDS == 16: half of it, rounds [round0, round0 + 16) - used for the last tiles of
// the list so that the SMs finish closer together (the trip count stays a compile-time constant: a run-time round
// range cost ~10 instructions per round in an earlier experiment).
template <bool ALCH, bool DIAG, bool U, bool X, bool P, int ROUNDS>
__device__ __forceinline__ void cq_tile(
    float *S, const float bx, const float by, const float bz, const float inv_bx, const float inv_by, const float inv_bz,
    const float cutoff2, const float beta, const int i_slot, const int round0, const CqSink &sink, i128 &energy) {
    const int lane = threadIdx.x & 31;
    const unsigned int lt_mask = (1u << lane) - 1u;
    int *SI = reinterpret_cast<int *>(S);
    const float xi = S[S_X + lane], yi = S[S_Y + lane], zi = S[S_Z + lane];
    const float wi = ALCH ? S[S_W + lane] : 0.0f;
    const float *jx = S + S_X + 32;
    const float *jy = S + S_Y + 32;
    const float *jz = S + S_Z + 32;
    const float *jw = S + S_W + 32;
    int head = 0;  // ring position of the oldest queued pair
    int count = 0; // queued pairs
    int jp = (lane + round0) & 31; // column position met in this round: (lane + round) % 32
#pragma unroll 2
    for (int round = 0; round < ROUNDS; round++) {
        const float dx = min_image(xi - jx[jp], bx, inv_bx);
        const float dy = min_image(yi - jy[jp], by, inv_by);
        const float dz = min_image(zi - jz[jp], bz, inv_bz);
        float d2 = dist2_3d(dx, dy, dz);
        float dw = 0.0f;
        if (ALCH) {
            dw = wi - jw[jp];
            d2 = fma_(dw, dw, d2);
        }
        // strict '<' (atoms parked at w == cutoff must not interact); NaN coordinates of padding atoms compare false
        bool hit = d2 < cutoff2;
        if (DIAG) {
            hit = hit && (i_slot < SI[S_JSLOT + jp]); // all-pairs: each pair once
        }
        const unsigned int ballot = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int k = (head + count + __popc(ballot & lt_mask)) & (CQ_QUEUE - 1);
            reinterpret_cast<float4 *>(S + S_Q4)[k] = make_float4(dx, dy, dz, __int_as_float(lane | (jp << 5)));
            if (ALCH) {
                S[S_QDW + k] = dw;
            }
        }
        count += __popc(ballot);
        if (count >= WARP) {
            cq_process<ALCH, U, X, P>(S, head, WARP, beta, sink, energy);
            head = (head + WARP) & (CQ_QUEUE - 1);
            count -= WARP;
        }
        jp = (jp + 1) & 31;
    }
    if (count > 0) {
        cq_process<ALCH, U, X, P>(S, head, count, beta, sink, energy);
    }
}

template <bool U, bool X, bool P> __global__ void __launch_bounds__(CQ_THREADS, CQ_MIN_CTAS) k_nb_tiles_cq(const NbTileArgs<float> a) {
    extern __shared__ __align__(16) float cq_smem[];
    __shared__ i128 scratch[CQ_WARPS];
    constexpr int WORDS = P ? S_WORDS_P : S_WORDS_X;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float *S = cq_smem + warp * WORDS;
    int *SI = reinterpret_cast<int *>(S);

    const float bx = static_cast<float>(a.box[0]), by = static_cast<float>(a.box[4]), bz = static_cast<float>(a.box[8]);
    const float inv_bx = 1.0f / bx, inv_by = 1.0f / by, inv_bz = 1.0f / bz;
    const float cutoff = static_cast<float>(a.cutoff);
    const float cutoff2 = cutoff * cutoff;
    const float beta = static_cast<float>(a.beta);
    const bool triangular = (a.NR == a.K);
    const float nan = __int_as_float(0x7fc00000);
    const CqBox cqbox = {bx, by, bz, inv_bx, inv_by, inv_bz};
    // half-precision prefilter (file header): enlarged threshold, 0 = off
    unsigned int thr2 = 0;
    if (a.prefilter && cutoff <= 1.5f) {
        thr2 = h2_bits(__half2half2(__float2half_ru(cutoff2 + 0.02f)));
    }

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (*a.rebuild_flag != 0) {
            a.rebuild_flag[2] += 1; // builds since construction (introspection only)
            *a.rebuild_flag = 0;
        }
    }
    // clear this warp's limb accumulators
    for (int c = 0; c < 6; c++) {
        SI[S_ACCX + c * 64 + lane] = 0;
        SI[S_ACCX + c * 64 + 32 + lane] = 0;
    }
    if (P) {
        for (int c = 0; c < 8; c++) {
            SI[S_ACCP + c * 64 + lane] = 0;
            SI[S_ACCP + c * 64 + 32 + lane] = 0;
        }
    }
    __syncwarp();

    const unsigned int T = min(*a.tile_count, a.tile_capacity);
    i128 energy = 0;
    int cur_row = -1;
    int i_slot = 0;
    bool i_valid = false;
    u64 gi[3] = {0, 0, 0};     // row-atom du/dx accumulated over a run of tiles with the same row block
    u64 gpi[4] = {0, 0, 0, 0}; // row-atom du/dp

    auto flush_row = [&]() {
        if (cur_row >= 0 && i_valid) {
            const size_t atom = a.perm[i_slot];
            if (X) {
                atomicAdd(a.du_dx + atom * 3 + 0, gi[0]);
                atomicAdd(a.du_dx + atom * 3 + 1, gi[1]);
                atomicAdd(a.du_dx + atom * 3 + 2, gi[2]);
            }
            if (P) {
                for (int c = 0; c < 4; c++) {
                    atomicAdd(a.du_dp + atom * P_PER_ATOM + c, gpi[c]);
                }
            }
        }
        gi[0] = gi[1] = gi[2] = 0;
        gpi[0] = gpi[1] = gpi[2] = gpi[3] = 0;
    };

    const unsigned int total_warps = gridDim.x * CQ_WARPS;
    const unsigned int n_static = min(a.static_tiles, T / total_warps);
    const unsigned int static_end = n_static * total_warps;
    CqSink sink = {a.perm, a.du_dx, a.du_dp, 0};
    // Dynamic part: whole tiles, then the last total_warps tiles of the list as two half-tiles each.  A tile takes
    // ~16 us of wall time on a fully loaded SM; handing out whole tiles to the very end left the SMs finishing up to
    // one tile apart (ncu r1: SMs active 88 % of the launch, 3.6 ns/tile at 90k atoms but 4.1 ns/tile at 30k).
    const unsigned int n_dynamic = T - static_end;
    const unsigned int n_split = min(n_dynamic, total_warps);
    const unsigned int n_whole = n_dynamic - n_split;
    for (bool first = true;; first = false) {
        unsigned int chunk_begin, chunk_end;
        int half = -1; // -1: whole tile; 0 / 1: that half of a split tile
        if (first) {
            chunk_begin = (blockIdx.x * CQ_WARPS + warp) * n_static;
            chunk_end = chunk_begin + n_static;
        } else {
            unsigned int next = 0;
            if (lane == 0) {
                next = atomicAdd(a.tile_cursor, 1u);
            }
            next = __shfl_sync(0xffffffffu, next, 0);
            if (next < n_whole) {
                chunk_begin = static_end + next;
            } else if (next < n_whole + 2u * n_split) {
                next -= n_whole;
                chunk_begin = static_end + n_whole + (next >> 1);
                half = static_cast<int>(next & 1u);
            } else {
                break;
            }
            chunk_end = chunk_begin + 1;
        }
        for (unsigned int t = chunk_begin; t < chunk_end; t++) {
            const int row = a.tile_rows[t];
            if (row != cur_row) {
                flush_row();
                cur_row = row;
                sink.row_base = row * TILE;
                i_slot = row * TILE + lane;
                i_valid = i_slot < a.NR;
                Vec4<float> c = {nan, nan, nan, 0.f}, p = {0.f, 0.f, 0.f, 0.f};
                if (i_valid) {
                    c = a.xw[i_slot];
                    p = a.qse[i_slot];
                }
                S[S_X + lane] = c.x;
                S[S_Y + lane] = c.y;
                S[S_Z + lane] = c.z;
                S[S_W + lane] = c.w;
                S[S_Q + lane] = p.x;
                S[S_SIG + lane] = p.y;
                S[S_EPS + lane] = p.z;
            }
            const int j_slot = static_cast<int>(min(a.tile_cols[t * TILE + lane], static_cast<unsigned int>(a.K)));
            const bool j_valid = j_slot < a.K;
            const size_t j_atom = j_valid ? a.perm[j_slot] : 0;
            {
                Vec4<float> c = {nan, nan, nan, 0.f}, p = {0.f, 0.f, 0.f, 0.f};
                if (j_valid) {
                    c = a.xw[j_slot];
                    p = a.qse[j_slot];
                }
                S[S_X + 32 + lane] = c.x;
                S[S_Y + 32 + lane] = c.y;
                S[S_Z + 32 + lane] = c.z;
                S[S_W + 32 + lane] = c.w;
                S[S_Q + 32 + lane] = p.x;
                S[S_SIG + 32 + lane] = p.y;
                S[S_EPS + 32 + lane] = p.z;
                SI[S_JSLOT + lane] = j_slot;
            }
            // 4D terms only where some atom of the tile has w != 0 (adding 0*0 is exact: identical results)
            const bool vanilla = __all_sync(0xffffffffu, S[S_W + lane] == 0.0f && S[S_W + 32 + lane] == 0.0f);
            // the i < j rule only bites when the column atoms overlap the row block itself
            const bool diag = triangular && __any_sync(0xffffffffu, j_valid && j_slot < (row + 1) * TILE);
            __syncwarp();
            bool prefilter = false;
            __half2 hx, hy, hz;
            if (thr2 != 0 && vanilla && !diag) {
                // coordinates relative to row atom 0 (always a real atom); usable when both atoms of every pair inside
                // the cutoff are guaranteed to sit in the same periodic image (file header)
                const float rx = S[S_X], ry = S[S_Y], rz = S[S_Z];
                const float ix = min_image(S[S_X + lane] - rx, bx, inv_bx);
                const float iy = min_image(S[S_Y + lane] - ry, by, inv_by);
                const float iz = min_image(S[S_Z + lane] - rz, bz, inv_bz);
                const float jx = min_image(S[S_X + 32 + lane] - rx, bx, inv_bx);
                const float jy = min_image(S[S_Y + 32 + lane] - ry, by, inv_by);
                const float jz = min_image(S[S_Z + 32 + lane] - rz, bz, inv_bz);
                const float ax = fminf(0.5f * (bx - cutoff) - 0.01f, 3.9f);
                const float ay = fminf(0.5f * (by - cutoff) - 0.01f, 3.9f);
                const float az = fminf(0.5f * (bz - cutoff) - 0.01f, 3.9f);
                // NaN (padding) passes: it can never produce a candidate
                const bool ok = !(fabsf(ix) >= ax) && !(fabsf(iy) >= ay) && !(fabsf(iz) >= az) && !(fabsf(jx) >= ax) &&
                                !(fabsf(jy) >= ay) && !(fabsf(jz) >= az);
                prefilter = __all_sync(0xffffffffu, ok);
                if (prefilter) {
                    hx = __half2half2(__float2half_rn(ix));
                    hy = __half2half2(__float2half_rn(iy));
                    hz = __half2half2(__float2half_rn(iz));
                    // lane j builds the f16x2 word P[j] = {c_j, c_j+1} of each component and stores it twice (S_H2 layout)
                    const unsigned int cxy = h2_bits(__floats2half2_rn(jx, jy));
                    const unsigned int czz = h2_bits(__floats2half2_rn(jz, jz));
                    const unsigned int nxy = __shfl_sync(0xffffffffu, cxy, (lane + 1) & 31);
                    const unsigned int nzz = __shfl_sync(0xffffffffu, czz, (lane + 1) & 31);
                    unsigned int *HW = reinterpret_cast<unsigned int *>(S) + S_H2 + lane;
                    const unsigned int wx = __byte_perm(cxy, nxy, 0x5410);
                    const unsigned int wy = __byte_perm(cxy, nxy, 0x7632);
                    const unsigned int wz = __byte_perm(czz, nzz, 0x5410);
                    HW[0] = wx;
                    HW[32] = wx;
                    HW[64] = wy;
                    HW[64 + 32] = wy;
                    HW[128] = wz;
                    HW[128 + 32] = wz;
                    __syncwarp();
                }
            }
            if (prefilter) {
                cq_tile_prefilter<U, X, P>(
                    S, cqbox, cutoff2, thr2, beta, hx, hy, hz, half >= 0 ? half : 0, half >= 0 ? half + 1 : 2, sink, energy);
            } else if (vanilla && !diag) {
                if (half >= 0) {
                    cq_tile<false, false, U, X, P, 16>(S, bx, by, bz, inv_bx, inv_by, inv_bz, cutoff2, beta, i_slot, half * 16, sink, energy);
                } else {
                    cq_tile<false, false, U, X, P, 32>(S, bx, by, bz, inv_bx, inv_by, inv_bz, cutoff2, beta, i_slot, 0, sink, energy);
                }
            } else if (half <= 0) { // the rarer tile kinds are never split: the first half-item does the whole tile
                if (vanilla) {
                    cq_tile<false, true, U, X, P, 32>(S, bx, by, bz, inv_bx, inv_by, inv_bz, cutoff2, beta, i_slot, 0, sink, energy);
                } else if (diag) {
                    cq_tile<true, true, U, X, P, 32>(S, bx, by, bz, inv_bx, inv_by, inv_bz, cutoff2, beta, i_slot, 0, sink, energy);
                } else {
                    cq_tile<true, false, U, X, P, 32>(S, bx, by, bz, inv_bx, inv_by, inv_bz, cutoff2, beta, i_slot, 0, sink, energy);
                }
            }
            // fold this tile's limb sums: row atoms into registers, column atoms (negated) to the sorted accumulators
            if (X) {
                for (int c = 0; c < 3; c++) {
                    gi[c] += limbs_take(SI + S_ACCX + c * 128, lane);
                    const u64 gj = 0ull - limbs_take(SI + S_ACCX + c * 128, 32 + lane);
                    if (j_valid && gj != 0) {
                        atomicAdd(a.du_dx + j_atom * 3 + c, gj);
                    }
                }
            }
            if (P) {
                for (int c = 0; c < 4; c++) {
                    gpi[c] += limbs_take(SI + S_ACCP + c * 128, lane);
                    const u64 gj = limbs_take(SI + S_ACCP + c * 128, 32 + lane);
                    if (j_valid && gj != 0) {
                        atomicAdd(a.du_dp + j_atom * P_PER_ATOM + c, gj);
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
    const size_t smem = CQ_WARPS * (P ? S_WORDS_P : S_WORDS_X) * sizeof(float);
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
        const size_t smem = CQ_WARPS * S_WORDS_X * sizeof(float);
        TMB_CUDA(cudaFuncSetAttribute(
            k_nb_tiles_cq<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        int per_sm = 0;
        TMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_nb_tiles_cq<false, true, false>, CQ_THREADS, smem));
        cached = sm_count() * (per_sm < 1 ? 1 : per_sm);
    }
    return cached;
}

void launch_nb_tiles_cq(const NbTileArgs<float> &args, bool with_u, bool with_dx, bool with_dp, cudaStream_t stream) {
    int grid = nb_tiles_cq_max_grid();
    if (args.grid_ctas > 0) {
        grid = std::min(grid, args.grid_ctas);
    }
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
