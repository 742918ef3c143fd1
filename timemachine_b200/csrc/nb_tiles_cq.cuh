// Device-side building blocks of the f32 "compaction queue" tile kernels (k_nb_tiles_cq.cu: the production kernel;
// k_nb_tiles_cq_async.cu: the same tile evaluation behind an asynchronous staging pipeline).  See k_nb_tiles_cq.cu for
// the description of the formulation.
#pragma once

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
// ROWREG: the row atom's force limbs are added to `racc` (registers: {x lo, x hi, y lo, y hi, z lo, z hi}) instead of the
// shared accumulators - the caller walks a row's pairs consecutively and flushes once per run (cq_tile_prefilter_runs).
template <bool ALCH, bool U, bool X, bool P, bool ROWREG = false>
__device__ __forceinline__ void cq_pair(
    float *S, const int i, const int j, const float dx, const float dy, const float dz, const float dw, const float d2,
    const float beta, const CqSink &sink, i128 &energy, unsigned int *racc = nullptr) {
    int *SI = reinterpret_cast<int *>(S);
    const float qi = S[S_Q + i], qj = S[S_Q + j];
    const float ei = S[S_EPS + i], ej = S[S_EPS + j];
    const PairTerms<float> t = pair_terms<float, U>(1.0f, 1.0f, qi, qj, S[S_SIG + i], S[S_SIG + j], ei, ej, d2, beta);
    const int si = sink.row_base + i; // sorted slots, translated to atoms on the (rare) direct path only
    const int sj = SI[S_JSLOT + j - 32];
    if (X) {
        // The fixed-point scale goes into the prefactor: (p 2^36) d == (p d) 2^36 bit for bit, scaling by a power of two
        // commutes with rounding - except where p d is subnormal (both forms then convert to 0) or p 2^36 overflows (the
        // test below fails and the rare path starts again from the unscaled product).  Two multiplies less per pair.
        const float ps = t.prefactor * static_cast<float>(FIXED_EXPONENT);
        const float sx = ps * dx, sy = ps * dy, sz = ps * dz;
        int *acc = SI + S_ACCX;
        // all three fixed-point values below 2^52 in magnitude (two limbs hold 2^53); NaN / inf compare false
        if (fabsf(sx) + fabsf(sy) + fabsf(sz) < 4503599627370496.0f) {
            const u64 fx = static_cast<u64>(round_to_i64_in_range(sx));
            const u64 fy = static_cast<u64>(round_to_i64_in_range(sy));
            const u64 fz = static_cast<u64>(round_to_i64_in_range(sz));
            if (ROWREG) {
                // unsigned: the low-limb sum of a row's <= 32 terms needs all 32 bits, the high limbs are two's complement
                racc[0] += static_cast<unsigned int>(fx) & LIMB_MASK;
                racc[1] += static_cast<unsigned int>(static_cast<i64>(fx) >> LIMB_BITS);
                racc[2] += static_cast<unsigned int>(fy) & LIMB_MASK;
                racc[3] += static_cast<unsigned int>(static_cast<i64>(fy) >> LIMB_BITS);
                racc[4] += static_cast<unsigned int>(fz) & LIMB_MASK;
                racc[5] += static_cast<unsigned int>(static_cast<i64>(fz) >> LIMB_BITS);
            } else {
                limb_add(acc + 0 * 128, i, fx);
                limb_add(acc + 1 * 128, i, fy);
                limb_add(acc + 2 * 128, i, fz);
            }
            // the column atom receives the exact negation; it is accumulated positively and negated once at the fold
            limb_add(acc + 0 * 128, j, fx);
            limb_add(acc + 1 * 128, j, fy);
            limb_add(acc + 2 * 128, j, fz);
        } else {
            // clashing atoms: too large for two limbs, add to the global accumulators directly.  Terms beyond the int64 range
            // are converted exactly as the reference converts them, the column atom's from -v like the reference's
            // (k_nonbonded.cuh:248-254): fixed(-v) == -fixed(v) only holds in range
            const float rx = t.prefactor * dx, ry = t.prefactor * dy, rz = t.prefactor * dz;
            u64 *gi = sink.du_dx + static_cast<size_t>(sink.perm[si]) * 3;
            u64 *gj = sink.du_dx + static_cast<size_t>(sink.perm[sj]) * 3;
            atomicAdd(gi + 0, to_fixed_force(rx));
            atomicAdd(gi + 1, to_fixed_force(ry));
            atomicAdd(gi + 2, to_fixed_force(rz));
            atomicAdd(gj + 0, to_fixed_force(-rx));
            atomicAdd(gj + 1, to_fixed_force(-ry));
            atomicAdd(gj + 2, to_fixed_force(-rz));
        }
    }
    if (P) {
        int *acc = SI + S_ACCP;
        const float vqi = qj * t.inv_d * t.damping, vqj = qi * t.inv_d * t.damping;
        const float vei = t.eps_grad * ej, vej = t.eps_grad * ei, vw = t.prefactor * dw;
        const u64 pqi = to_fixed_in_range<FIXED_EXPONENT_DU_DCHARGE>(vqi);
        const u64 pqj = to_fixed_in_range<FIXED_EXPONENT_DU_DCHARGE>(vqj);
        u64 psig = 0, pei = 0, pej = 0, pw = 0;
        if (t.lj) {
            psig = to_fixed_in_range<FIXED_EXPONENT_DU_DSIG>(t.sig_grad);
            pei = to_fixed_in_range<FIXED_EXPONENT_DU_DEPS>(vei);
            pej = to_fixed_in_range<FIXED_EXPONENT_DU_DEPS>(vej);
        }
        if (ALCH) {
            pw = to_fixed_in_range<FIXED_EXPONENT_DU_DW>(vw); // antisymmetric: the column atom gets -pw
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
            // some term is too large for two limbs, possibly for int64: the reference's conversion, term by term
            atomicAdd(gi + P_CHARGE, to_fixed<FIXED_EXPONENT_DU_DCHARGE>(vqi));
            atomicAdd(gj + P_CHARGE, to_fixed<FIXED_EXPONENT_DU_DCHARGE>(vqj));
            if (t.lj) {
                atomicAdd(gi + P_SIG, to_fixed<FIXED_EXPONENT_DU_DSIG>(t.sig_grad));
                atomicAdd(gj + P_SIG, to_fixed<FIXED_EXPONENT_DU_DSIG>(t.sig_grad));
                atomicAdd(gi + P_EPS, to_fixed<FIXED_EXPONENT_DU_DEPS>(vei));
                atomicAdd(gj + P_EPS, to_fixed<FIXED_EXPONENT_DU_DEPS>(vej));
            }
            if (ALCH) {
                atomicAdd(gi + P_W, to_fixed<FIXED_EXPONENT_DU_DW>(vw));
                atomicAdd(gj + P_W, to_fixed<FIXED_EXPONENT_DU_DW>(-vw));
            }
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
        // every lane has read its entries of [0, head) before any lane overwrites the queue again (the move below, or the next
        // half's appends when nothing is left over: found by compute-sanitizer racecheck, profiles/r2_sanitizer.txt)
        __syncwarp();
        if (!last && head > 0 && count > 0) {
            // what is left (< 32) moves to the front
            const unsigned short v = Q[head + (lane < count ? lane : 0)];
            __syncwarp();
            if (lane < count) {
                Q[lane] = v;
            }
        }
    }
    __syncwarp();
}

// "Row runs" (round 2): the same prefilter tile with the candidates kept as one 32-bit mask per row atom instead of a
// queue of codes.  Phase A is then 11 instructions per double round (no ballots, no prefix popcounts, no queue stores:
// HSET2.BM gives 0xffff per passing half and one LOP3 files both results under bits R and 16 + R of the lane's mask).  An
// exclusive scan of the popcounts makes the masks an implicit ROW-SORTED list of n candidates; lane l evaluates the K =
// ceil(n / 32) consecutive entries [l K, (l + 1) K) - perfectly balanced - and because consecutive entries share the row
// atom, its force limbs are summed in registers and flushed with shared atomics only where the row changes (~2 flushes per
// lane and tile instead of K): per evaluated batch 6 nearly conflict-free ATOMS wavefronts on the row side instead of
// 6 x 2.8 (profiles/r2_summary.md section 5 has the conflict model), in a kernel bound by that pipe.  The column side and
// every rounded operation are unchanged: bitwise the same sums.
// Bit b of row i's mask <-> column position (i + 2 (b & 15) + (b >> 4)) mod 32: a row's candidates are walked even offsets
// first, then odd.  (Staging the column words as {c_k, c_(k+16)} makes the bit number the offset itself and saves four
// instructions per step, but walking in plain offset order raises the column-side conflicts from 3.9 to 4.25 passes per
// ATOMS - profiles/study_row_runs.py - and measured 1.9 % slower, profiles/r2_summary.md section 10.)
constexpr int S_RM = S_QC;      // u32[32]: candidate mask per row
constexpr int S_RO = S_QC + 32; // u32[32]: exclusive prefix sum of the candidate counts
constexpr int S_RG = S_QC + 64; // u64[3][32]: row-atom force sums over a run of tiles (k_nb_tiles_cq.cu)
static_assert(64 + 192 <= CQ_CODES / 2 && S_RG % 2 == 0 && S_WORDS_X % 2 == 0 && S_WORDS_P % 2 == 0, "row-run tables live in the code queue's space, u64 aligned");

template <bool U, bool X, bool P>
__device__ __forceinline__ void cq_tile_prefilter_runs(
    float *S, const CqBox &b, const float cutoff2, const unsigned int thr2, const float beta, const __half2 hx,
    const __half2 hy, const __half2 hz, const int h0, const int h1, const CqSink &sink, i128 &energy) {
    const int lane = threadIdx.x & 31;
    unsigned int *SU = reinterpret_cast<unsigned int *>(S);
    int *SI = reinterpret_cast<int *>(S);
    // ---- phase A: this lane's row atom against all (or one half of the) column positions
    unsigned int m = 0;
    const __half2 thr = bits_h2(thr2);
#pragma unroll 1
    for (int h = h0; h < h1; h++) {
        const unsigned int *H = SU + S_H2 + lane + 16 * h;
        unsigned int mh = 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const __half2 dx = __hsub2(hx, bits_h2(H[2 * r]));
            const __half2 dy = __hsub2(hy, bits_h2(H[64 + 2 * r]));
            const __half2 dz = __hsub2(hz, bits_h2(H[128 + 2 * r]));
            const __half2 d2 = __hfma2(dz, dz, __hfma2(dy, dy, __hmul2(dx, dx)));
            mh |= __hlt2_mask(d2, thr) & (0x00010001u << r); // NaN padding compares false
        }
        m |= mh << (8 * h);
    }
    // ---- the implicit row-sorted candidate list
    const int c = __popc(m);
    int incl = c;
#pragma unroll
    for (int d = 1; d < WARP; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) {
            incl += up;
        }
    }
    const int n = __shfl_sync(0xffffffffu, incl, WARP - 1);
    if (n == 0) {
        return;
    }
    const unsigned int nonempty = __ballot_sync(0xffffffffu, c != 0);
    SU[S_RM + lane] = m;
    SU[S_RO + lane] = static_cast<unsigned int>(incl - c);
    __syncwarp();
    const int K = (n + WARP - 1) >> 5;
    const int e0 = lane * K;
    const int todo = min(K, n - e0); // <= 0: nothing for this lane
    int i = 0;
    unsigned int cur = 0;
    if (todo > 0) {
        // the row holding entry e0: the last row whose prefix is <= e0 (empty rows tie with the next non-empty one)
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            if (SU[S_RO + i + step] <= static_cast<unsigned int>(e0)) {
                i += step;
            }
        }
        int skip = e0 - static_cast<int>(SU[S_RO + i]);
        cur = SU[S_RM + i];
        // drop the `skip` lowest candidates of that row (select by rank, five halvings)
        int pos = 0;
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
            const int cnt = __popc((cur >> pos) & ((1u << w) - 1u));
            if (cnt <= skip) {
                skip -= cnt;
                pos += w;
            }
        }
        cur = (cur >> pos) << pos;
    }
    // ---- phase B: K steps, one candidate per lane and step
    unsigned int racc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        if (k < todo) {
            const int bpos = __ffs(static_cast<int>(cur)) - 1;
            const int j = 32 + ((i + ((bpos & 15) << 1) + (bpos >> 4)) & 31);
            const float dx = min_image(S[S_X + i] - S[S_X + j], b.bx, b.inv_bx);
            const float dy = min_image(S[S_Y + i] - S[S_Y + j], b.by, b.inv_by);
            const float dz = min_image(S[S_Z + i] - S[S_Z + j], b.bz, b.inv_bz);
            const float d2 = dist2_3d(dx, dy, dz);
            if (d2 < cutoff2) {
                cq_pair<false, U, X, P, true>(S, i, j, dx, dy, dz, 0.0f, d2, beta, sink, energy, racc);
            }
            cur &= cur - 1u;
            if (cur == 0u || k == todo - 1) {
                if (X) {
                    int *acc = SI + S_ACCX + i;
#pragma unroll
                    for (int q = 0; q < 6; q++) {
                        atomicAdd(acc + q * 64, static_cast<int>(racc[q])); // [comp][limb][atom]: x lo, x hi, y lo, ...
                        racc[q] = 0;
                    }
                }
                const unsigned int above = nonempty & ~((2u << i) - 1u);
                if (above != 0u) {
                    i = __ffs(static_cast<int>(above)) - 1;
                    cur = SU[S_RM + i];
                }
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

} // namespace tmb
