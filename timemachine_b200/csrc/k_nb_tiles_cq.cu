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
//
// Row runs (round 2, the production form of a prefilter tile; nb_tiles_cq.cuh: cq_tile_prefilter_runs): instead of queueing
// codes, every lane keeps the candidates of its row atom as one 32-bit mask (11 instructions per 64 pairs), a scan of the
// popcounts makes the masks an implicit row-sorted list, the lanes walk it in equal contiguous chunks and keep the row atom's
// force limbs in registers until the row changes.  Same pairs, same terms, same sums; 112.8 -> 109.7 us per launch
// (profiles/r2_summary.md section 10).  -DCQ_ROW_RUNS=0 builds the code-queue form for A/B runs.
#include "nb_tiles_cq.cuh"

#ifndef CQ_ROW_RUNS
#define CQ_ROW_RUNS 1 // 0: the code-queue phase A / B of round 1 (cq_tile_prefilter), kept for A/B measurements
#endif

namespace tmb {

// per-warp shared words
__host__ __device__ constexpr int cq_words(bool, bool p) { return p ? S_WORDS_P : S_WORDS_X; }

template <bool U, bool X, bool P> __global__ void __launch_bounds__(CQ_THREADS, CQ_MIN_CTAS) k_nb_tiles_cq(const NbTileArgs<float> a) {
    extern __shared__ __align__(16) float cq_smem[];
    __shared__ i128 scratch[CQ_WARPS];
    constexpr int WORDS = cq_words(X, P);
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
    // row-atom du/dx accumulated over a run of tiles with the same row block
#if CQ_ROW_RUNS
    // touched once per tile, so it lives in shared memory (lane-private u64[3][32] slots in the part of the code queue's space
    // the row-run tables leave free) and leaves its six registers to the phase-B loop.  NOT in extra shared memory: 6 KB more
    // per CTA moves the L1 / shared carve-out from 196 to 228 KB and costs more than the registers gain (measured).
    u64 *G = reinterpret_cast<u64 *>(S + S_RG);
#define CQ_GI(c) G[(c) * 32 + lane]
#else
    u64 gi[3];
#define CQ_GI(c) gi[c]
#endif
    if (X) {
        CQ_GI(0) = 0;
        CQ_GI(1) = 0;
        CQ_GI(2) = 0;
    }
    u64 gpi[4] = {0, 0, 0, 0}; // row-atom du/dp

    auto flush_row = [&]() {
        if (cur_row >= 0 && i_valid) {
            const size_t atom = a.perm[i_slot];
            if (X) {
                atomicAdd(a.du_dx + atom * 3 + 0, CQ_GI(0));
                atomicAdd(a.du_dx + atom * 3 + 1, CQ_GI(1));
                atomicAdd(a.du_dx + atom * 3 + 2, CQ_GI(2));
            }
            if (P) {
                for (int c = 0; c < 4; c++) {
                    atomicAdd(a.du_dp + atom * P_PER_ATOM + c, gpi[c]);
                }
            }
        }
        if (X) {
            CQ_GI(0) = 0;
            CQ_GI(1) = 0;
            CQ_GI(2) = 0;
        }
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
#if CQ_ROW_RUNS
                cq_tile_prefilter_runs<U, X, P>(
                    S, cqbox, cutoff2, thr2, beta, hx, hy, hz, half >= 0 ? half : 0, half >= 0 ? half + 1 : 2, sink, energy);
#else
                cq_tile_prefilter<U, X, P>(
                    S, cqbox, cutoff2, thr2, beta, hx, hy, hz, half >= 0 ? half : 0, half >= 0 ? half + 1 : 2, sink, energy);
#endif
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
                    CQ_GI(c) += limbs_take(SI + S_ACCX + c * 128, lane);
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
    const size_t smem = CQ_WARPS * cq_words(X, P) * sizeof(float);
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
        const size_t smem = CQ_WARPS * cq_words(true, false) * sizeof(float);
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
