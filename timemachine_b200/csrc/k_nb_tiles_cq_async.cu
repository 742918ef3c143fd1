// f32 tile kernel with an ASYNCHRONOUS STAGING PIPELINE (north_star: "TMA-staged coordinate/parameter tiles"): the same
// tile evaluation as k_nb_tiles_cq.cu - every pair term goes through the same code in nb_tiles_cq.cuh, results are bit
// for bit the same - but what a warp needs for its NEXT work items is brought into shared memory while it evaluates the
// current one:
//
//   item k + 2   the 128-byte vector of column-atom slots of its tile: one 1-D TMA bulk copy
//                (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes) into a two-slot ring, completion
//                signalled on an mbarrier;
//   item k + 1   its 32 column atoms {x,y,z,w}, {q,sigma,eps,0}: a GATHER through that index vector, which a bulk /
//                tensor-map copy cannot express (32 rows of 16 B at arbitrary addresses), so per-lane 16-byte
//                cp.async (LDGSTS) into a 1 KB landing buffer, tracked with cp.async groups;
//   item k       evaluated from the warp's structure-of-arrays block exactly as in k_nb_tiles_cq.cu (the landing buffer
//                is transposed into it first: the evaluation reads atoms component-wise, 32 banks conflict-free).
//
// The production kernel instead issues the index load and the two dependent 128-bit atom loads when it reaches a tile and
// relies on the other 31 warps of the SM to cover the two L2 round trips.  Which of the two is faster is a measurement:
// profiles/r2_summary.md has the A/B (TMB_NB_ASYNC=1 selects this kernel); DESIGN.md section 4 states the outcome.
//
// Work items are handed out exactly as in the production kernel (static tiles per warp, then the device cursor, the last
// wave as half tiles) but two items ahead of the evaluation.
#include "nb_tiles_cq.cuh"

namespace tmb {

// extra per-warp shared memory behind the production layout: landing buffer, index ring, two mbarriers
constexpr int A_COLBUF = 0;            // float4[64]: xw[32] then qse[32]
constexpr int A_RING = 256;            // u32[2][32]
constexpr int A_MBAR = A_RING + 64;    // u64[2] (8-byte aligned: word offset even)
constexpr int A_WORDS = A_MBAR + 4;
static_assert(A_WORDS % 4 == 0, "keeps the next warp's block 16-byte aligned");

__device__ __forceinline__ unsigned int smem_u32(const void *p) { return static_cast<unsigned int>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA: `bytes` (multiple of 16) from global to shared, completion counted on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned int bytes, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void cp_async_16(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct CqItem {
    unsigned int t; // tile
    int half;       // -1: whole tile; 0 / 1: that half of a split tile
    bool valid;
};

template <bool U, bool X, bool P>
__global__ void __launch_bounds__(CQ_THREADS, CQ_MIN_CTAS) k_nb_tiles_cq_async(const NbTileArgs<float> a) {
    extern __shared__ __align__(16) float cq_smem[];
    __shared__ i128 scratch[CQ_WARPS];
    constexpr int WORDS = (P ? S_WORDS_P : S_WORDS_X) + A_WORDS;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float *S = cq_smem + warp * WORDS;
    int *SI = reinterpret_cast<int *>(S);
    float *A = S + (P ? S_WORDS_P : S_WORDS_X);
    Vec4<float> *colbuf = reinterpret_cast<Vec4<float> *>(A + A_COLBUF);
    unsigned int *ring = reinterpret_cast<unsigned int *>(A + A_RING);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(A + A_MBAR);

    const float bx = static_cast<float>(a.box[0]), by = static_cast<float>(a.box[4]), bz = static_cast<float>(a.box[8]);
    const float inv_bx = 1.0f / bx, inv_by = 1.0f / by, inv_bz = 1.0f / bz;
    const float cutoff = static_cast<float>(a.cutoff);
    const float cutoff2 = cutoff * cutoff;
    const float beta = static_cast<float>(a.beta);
    const bool triangular = (a.NR == a.K);
    const float nan = __int_as_float(0x7fc00000);
    const CqBox cqbox = {bx, by, bz, inv_bx, inv_by, inv_bz};
    unsigned int thr2 = 0;
    if (a.prefilter && cutoff <= 1.5f) {
        thr2 = h2_bits(__half2half2(__float2half_ru(cutoff2 + 0.02f)));
    }

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (*a.rebuild_flag != 0) {
            a.rebuild_flag[2] += 1;
            *a.rebuild_flag = 0;
        }
    }
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
    if (lane == 0) {
        mbar_init(mbar + 0, 1);
        mbar_init(mbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncwarp();

    const unsigned int T = min(*a.tile_count, a.tile_capacity);
    i128 energy = 0;
    int cur_row = -1;
    int i_slot = 0;
    bool i_valid = false;
    u64 gi[3] = {0, 0, 0};
    u64 gpi[4] = {0, 0, 0, 0};

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

    // ---- work items: the production kernel's schedule, as a generator -----------------------------------------------
    const unsigned int total_warps = gridDim.x * CQ_WARPS;
    const unsigned int n_static = min(a.static_tiles, T / total_warps);
    const unsigned int static_end = n_static * total_warps;
    const unsigned int n_dynamic = T - static_end;
    const unsigned int n_split = min(n_dynamic, total_warps);
    const unsigned int n_whole = n_dynamic - n_split;
    unsigned int static_pos = (blockIdx.x * CQ_WARPS + warp) * n_static;
    const unsigned int static_stop = static_pos + n_static;
    bool exhausted = false;
    auto next_item = [&]() -> CqItem {
        CqItem it = {0u, -1, false};
        if (static_pos < static_stop) {
            it.t = static_pos++;
            it.valid = true;
            return it;
        }
        if (exhausted) {
            return it;
        }
        unsigned int next = 0;
        if (lane == 0) {
            next = atomicAdd(a.tile_cursor, 1u);
        }
        next = __shfl_sync(0xffffffffu, next, 0);
        if (next < n_whole) {
            it.t = static_end + next;
            it.valid = true;
        } else if (next < n_whole + 2u * n_split) {
            next -= n_whole;
            it.t = static_end + n_whole + (next >> 1);
            it.half = static_cast<int>(next & 1u);
            it.valid = true;
        } else {
            exhausted = true;
        }
        return it;
    };

    // ---- staging primitives ---------------------------------------------------------------------------------------------
    auto issue_index_copy = [&](const CqItem &it, int slot) { // 128 B of column slots -> ring[slot], via TMA
        if (lane == 0) {
            fence_proxy_async(); // the ring slot was last read through the generic proxy
            mbar_expect_tx(mbar + slot, 128u);
            tma_bulk_g2s(ring + slot * 32, a.tile_cols + static_cast<size_t>(it.t) * TILE, 128u, mbar + slot);
        }
    };
    auto issue_gather = [&](int j_slot) { // this lane's column atom -> landing buffer (nothing for padding slots)
        if (j_slot < a.K) {
            cp_async_16(colbuf + lane, a.xw + j_slot);
            cp_async_16(colbuf + 32 + lane, a.qse + j_slot);
        }
        cp_async_commit();
    };

    CqSink sink = {a.perm, a.du_dx, a.du_dp, 0};
    unsigned int phases = 0u; // bit s: the parity the next wait on mbarrier s expects

    // prologue: item 0 is staged synchronously, item 1's indices start their trip
    CqItem cur = next_item();
    int j_cur = 0;
    if (cur.valid) {
        j_cur = static_cast<int>(min(a.tile_cols[static_cast<size_t>(cur.t) * TILE + lane], static_cast<unsigned int>(a.K)));
        issue_gather(j_cur);
    }
    CqItem nxt = cur.valid ? next_item() : cur;
    if (nxt.valid) {
        issue_index_copy(nxt, 1);
    }
    for (unsigned int k = 0; cur.valid; k++) {
        const unsigned int t = cur.t;
        const int half = cur.half;
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
        const int j_slot = j_cur;
        const bool j_valid = j_slot < a.K;
        const size_t j_atom = j_valid ? a.perm[j_slot] : 0;
        // the column atoms of this item have been on their way since the previous iteration
        cp_async_wait_all();
        __syncwarp();
        {
            Vec4<float> c = {nan, nan, nan, 0.f}, p = {0.f, 0.f, 0.f, 0.f};
            if (j_valid) {
                c = colbuf[lane];
                p = colbuf[32 + lane];
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
        __syncwarp(); // every lane is done with the landing buffer
        // next item: its index vector has landed in the ring; start the gather of its atoms
        int j_nxt = 0;
        if (nxt.valid) {
            const int slot = (k + 1) & 1;
            mbar_wait(mbar + slot, (phases >> slot) & 1u);
            phases ^= 1u << slot;
            j_nxt = static_cast<int>(min(ring[slot * 32 + lane], static_cast<unsigned int>(a.K)));
            issue_gather(j_nxt);
        }
        // the item after that: its index vector goes into the slot this item's came from
        CqItem nn = nxt.valid ? next_item() : nxt;
        if (nn.valid) {
            __syncwarp();
            issue_index_copy(nn, k & 1);
        }

        // ---- evaluation: identical to k_nb_tiles_cq.cu ------------------------------------------------------------------
        const bool vanilla = __all_sync(0xffffffffu, S[S_W + lane] == 0.0f && S[S_W + 32 + lane] == 0.0f);
        const bool diag = triangular && __any_sync(0xffffffffu, j_valid && j_slot < (row + 1) * TILE);
        __syncwarp();
        bool prefilter = false;
        __half2 hx, hy, hz;
        if (thr2 != 0 && vanilla && !diag) {
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
            const bool ok = !(fabsf(ix) >= ax) && !(fabsf(iy) >= ay) && !(fabsf(iz) >= az) && !(fabsf(jx) >= ax) &&
                            !(fabsf(jy) >= ay) && !(fabsf(jz) >= az);
            prefilter = __all_sync(0xffffffffu, ok);
            if (prefilter) {
                hx = __half2half2(__float2half_rn(ix));
                hy = __half2half2(__float2half_rn(iy));
                hz = __half2half2(__float2half_rn(iz));
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
        } else if (half <= 0) {
            if (vanilla) {
                cq_tile<false, true, U, X, P, 32>(S, bx, by, bz, inv_bx, inv_by, inv_bz, cutoff2, beta, i_slot, 0, sink, energy);
            } else if (diag) {
                cq_tile<true, true, U, X, P, 32>(S, bx, by, bz, inv_bx, inv_by, inv_bz, cutoff2, beta, i_slot, 0, sink, energy);
            } else {
                cq_tile<true, false, U, X, P, 32>(S, bx, by, bz, inv_bx, inv_by, inv_bz, cutoff2, beta, i_slot, 0, sink, energy);
            }
        }
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
        cur = nxt;
        j_cur = j_nxt;
        nxt = nn;
    }
    flush_row();

    if (U) {
        grid_finish_energy(energy, scratch, a.u_partials, a.ticket, a.d_u);
    }
}

template <bool U, bool X, bool P> static void cq_async_launch(const NbTileArgs<float> &args, int grid, cudaStream_t stream) {
    const size_t smem = CQ_WARPS * ((P ? S_WORDS_P : S_WORDS_X) + A_WORDS) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        TMB_CUDA(cudaFuncSetAttribute(
            k_nb_tiles_cq_async<U, X, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured = true;
    }
    TMB_LAUNCH((k_nb_tiles_cq_async<U, X, P>), grid, CQ_THREADS, smem, stream, args);
}

int nb_tiles_cq_async_max_grid() {
    static int cached = 0;
    if (cached == 0) {
        const size_t smem = CQ_WARPS * (S_WORDS_X + A_WORDS) * sizeof(float);
        TMB_CUDA(cudaFuncSetAttribute(
            k_nb_tiles_cq_async<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        int per_sm = 0;
        TMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_nb_tiles_cq_async<false, true, false>, CQ_THREADS, smem));
        cached = sm_count() * (per_sm < 1 ? 1 : per_sm);
    }
    return cached;
}

void launch_nb_tiles_cq_async(const NbTileArgs<float> &args, bool with_u, bool with_dx, bool with_dp, cudaStream_t stream) {
    int grid = nb_tiles_cq_async_max_grid();
    if (args.grid_ctas > 0) {
        grid = std::min(grid, args.grid_ctas);
    }
    const int sel = (with_u ? 4 : 0) | (with_dx ? 2 : 0) | (with_dp ? 1 : 0);
    switch (sel) {
    case 0:
        break;
    case 1:
        cq_async_launch<false, false, true>(args, grid, stream);
        break;
    case 2:
        cq_async_launch<false, true, false>(args, grid, stream);
        break;
    case 3:
        cq_async_launch<false, true, true>(args, grid, stream);
        break;
    case 4:
        cq_async_launch<true, false, false>(args, grid, stream);
        break;
    case 5:
        cq_async_launch<true, false, true>(args, grid, stream);
        break;
    case 6:
        cq_async_launch<true, true, false>(args, grid, stream);
        break;
    case 7:
        cq_async_launch<true, true, true>(args, grid, stream);
        break;
    }
}

} // namespace tmb
