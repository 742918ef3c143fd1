// Neighbour-list construction: periodic-aware block bounding boxes and the 32-row x 32-column interaction tile list.
// Membership rule follows the reference (timemachine/cpp/src/kernels/k_neighborlist.cuh:11-458, itself after OpenMM's
// findInteractingBlocks): column atom j is listed for row block b iff j's column block passes the box-box test against
// b and some row atom i of b passes the atom-box test and |x_i - x_j|_pbc < cutoff, all evaluated in `Real`.
//
// What is different (B200 design):
//  * ONE kernel per build: a CTA of 4 warps owns a row block, the warps stride over 32-column-block chunks, and the
//    per-warp leftovers are merged inside the CTA (the reference needs a second kernel, k_compact_trim_atoms, and a
//    grid of row_blocks x Y single-warp CTAs);
//  * finished tiles are staged in shared memory and published in batches with one atomicAdd per batch, so the list
//    is made of RUNS of tiles with equal row block - the tile kernel keeps the row atoms in registers across a run;
//  * the whole build is skipped on the device when the rebuild flag is clear (no host round trip).
#include "block_bounds.cuh"
#include "kernels.hpp"

#include <algorithm>

namespace tmb {

constexpr int BB_THREADS = 128;

template <typename Real>
__device__ __forceinline__ void
load_pos(const double *__restrict__ coords, const Vec4<Real> *__restrict__ xw, unsigned int atom, Real &x, Real &y, Real &z) {
    if (xw != nullptr) {
        const Vec4<Real> v = xw[atom];
        x = v.x;
        y = v.y;
        z = v.z;
    } else {
        x = static_cast<Real>(coords[static_cast<size_t>(atom) * 3 + 0]);
        y = static_cast<Real>(coords[static_cast<size_t>(atom) * 3 + 1]);
        z = static_cast<Real>(coords[static_cast<size_t>(atom) * 3 + 2]);
    }
}

// Tight periodic bounding box of 32 consecutive indices.  The running re-centring is order dependent, so the atoms are
// visited in the reference's order (1, 2, ..., 31, 0 relative to the block start; k_neighborlist.cuh:84-106).
template <typename Real> __global__ void __launch_bounds__(BB_THREADS) k_block_bounds(const BlockBoundsArgs<Real> a) {
    if (a.flag != nullptr && *a.flag == 0) {
        return;
    }
    const int block = (blockIdx.x * blockDim.x + threadIdx.x) / WARP;
    if (block >= a.num_blocks) {
        return;
    }
    const int lane = threadIdx.x & 31;
    const int i = block * TILE + lane;
    const bool valid = i < a.num_idxs;

    const Real bx = static_cast<Real>(a.box[0]);
    const Real by = static_cast<Real>(a.box[4]);
    const Real bz = static_cast<Real>(a.box[8]);
    const Real inv_bx = 1 / bx;
    const Real inv_by = 1 / by;
    const Real inv_bz = 1 / bz;

    if (blockIdx.x == 0 && threadIdx.x < 9) {
        if (threadIdx.x == 0 && a.reset_count != nullptr) {
            *a.reset_count = 0;
            *a.reset_overflow = 0;
        }
        if (a.box_build != nullptr) {
            a.box_build[threadIdx.x] = a.box[threadIdx.x];
        }
    }
    Real px = 0, py = 0, pz = 0;
    if (valid) {
        const unsigned int atom = a.idxs != nullptr ? a.idxs[i] : static_cast<unsigned int>(a.base + i);
        load_pos<Real>(a.coords, a.xw, atom, px, py, pz);
        if (a.xw_build != nullptr) {
            // remember where this slot's atom was when the list was built (reference nonbonded_all_pairs.cu:241-242)
            a.xw_build[atom] = a.xw[atom];
        }
    }
    Real ctr[3], ext[3];
    warp_block_bounds<Real>(px, py, pz, a.num_idxs - block * TILE, bx, by, bz, inv_bx, inv_by, inv_bz, ctr, ext);
    if (lane == 0) {
        for (int c = 0; c < 3; c++) {
            a.ctr[block * 3 + c] = ctr[c];
            a.ext[block * 3 + c] = ext[c];
        }
    }
}

template <typename Real> void launch_block_bounds(const BlockBoundsArgs<Real> &args, cudaStream_t stream) {
    if (args.num_blocks <= 0) {
        return;
    }
    TMB_LAUNCH(k_block_bounds<Real>, ceil_div(args.num_blocks * WARP, BB_THREADS), BB_THREADS, 0, stream, args);
}
template void launch_block_bounds<float>(const BlockBoundsArgs<float> &, cudaStream_t);
template void launch_block_bounds<double>(const BlockBoundsArgs<double> &, cudaStream_t);

// ---------------------------------------------------------------------------------------------------------------
#ifndef BT_WARPS_N
#define BT_WARPS_N 8 // measured on the 30k-atom all-pairs build: 4 -> 79 us, 8 -> 66 us, 16 -> 70.5 us (the heaviest row blocks bound the launch)
#endif
constexpr int BT_WARPS = BT_WARPS_N;
constexpr int BT_THREADS = BT_WARPS * WARP;
constexpr int BT_STAGE = 8; // tiles staged per warp before publishing

struct WarpTileWriter {
    int *buf;            // [64] compaction buffer of column atoms
    unsigned int *stage; // [BT_STAGE][32] finished tiles
    int nbuf;
    int nstage;
    int row_block;
    int sentinel;
    TileList tiles;

    __device__ __forceinline__ void publish() {
        const int lane = threadIdx.x & 31;
        if (nstage == 0) {
            return;
        }
        unsigned int base = 0;
        if (lane == 0) {
            base = atomicAdd(tiles.count, static_cast<unsigned int>(nstage));
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int s = 0; s < nstage; s++) {
            const unsigned int t = base + s;
            if (t < tiles.capacity) {
                if (lane == 0) {
                    tiles.rows[t] = row_block;
                }
                tiles.cols[static_cast<size_t>(t) * TILE + lane] = stage[s * TILE + lane];
            } else if (lane == 0) {
                *tiles.overflow = 1;
                atomicMax(tiles.overflow + 1, base + static_cast<unsigned int>(nstage)); // how much room was needed
            }
        }
        __syncwarp();
        nstage = 0;
    }

    // move buf[0..31] to the stage as a finished tile (entries >= count are padding) and slide buf[32..63] down
    __device__ __forceinline__ void emit(int count) {
        const int lane = threadIdx.x & 31;
        __syncwarp();
        stage[nstage * TILE + lane] = lane < count ? static_cast<unsigned int>(buf[lane]) : static_cast<unsigned int>(sentinel);
        const int upper = buf[TILE + lane];
        __syncwarp();
        buf[lane] = upper;
        buf[TILE + lane] = sentinel;
        __syncwarp();
        nstage++;
        if (nstage == BT_STAGE) {
            publish();
        }
    }

    // append the atoms of lanes with `take` set, in lane order
    __device__ __forceinline__ void append(bool take, int atom) {
        const int lane = threadIdx.x & 31;
        const unsigned int mask = __ballot_sync(0xffffffffu, take);
        if (take) {
            buf[nbuf + __popc(mask & ((1u << lane) - 1u))] = atom;
        }
        nbuf += __popc(mask);
        if (nbuf > TILE) {
            emit(TILE);
            nbuf -= TILE;
        }
    }
};

template <typename Real, bool TRI> __global__ void __launch_bounds__(BT_THREADS) k_build_tiles(const BuildTilesArgs<Real> a) {
    if (a.flag != nullptr && *a.flag == 0) {
        return;
    }
    if (a.snap_xw_build != nullptr) {
        // remember where every atom was when the list was built (reference nonbonded_all_pairs.cu:241-242)
        const int n_threads = gridDim.x * gridDim.y * BT_THREADS;
        const int tid = (blockIdx.y * gridDim.x + blockIdx.x) * BT_THREADS + threadIdx.x;
        for (int k = tid; k < a.snap_slots; k += n_threads) {
            a.snap_xw_build[k] = a.xw[k];
        }
        if (tid < 9) {
            a.snap_box_build[tid] = a.box[tid];
        }
    }
    __shared__ int s_buf[BT_WARPS][2 * TILE];
    __shared__ unsigned int s_stage[BT_WARPS][BT_STAGE * TILE];
    __shared__ int s_left[BT_WARPS];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int row_block = blockIdx.x;
    const int N = a.N;

    s_buf[warp][lane] = N;
    s_buf[warp][TILE + lane] = N;
    __syncwarp();

    WarpTileWriter w;
    w.buf = s_buf[warp];
    w.stage = s_stage[warp];
    w.nbuf = 0;
    w.nstage = 0;
    w.row_block = row_block;
    w.sentinel = N;
    w.tiles = a.tiles;

    const int ri = row_block * TILE + lane;
    const unsigned int atom_i =
        ri < a.NR ? (a.row_idxs != nullptr ? a.row_idxs[ri] : static_cast<unsigned int>(a.row_base + ri)) : static_cast<unsigned int>(N);

    const Real row_ctr_x = a.row_ctr[row_block * 3 + 0];
    const Real row_ctr_y = a.row_ctr[row_block * 3 + 1];
    const Real row_ctr_z = a.row_ctr[row_block * 3 + 2];
    const Real row_ext_x = a.row_ext[row_block * 3 + 0];
    const Real row_ext_y = a.row_ext[row_block * 3 + 1];
    const Real row_ext_z = a.row_ext[row_block * 3 + 2];

    Real raw_i_x = 0, raw_i_y = 0, raw_i_z = 0;
    if (atom_i < static_cast<unsigned int>(N)) {
        load_pos<Real>(a.coords, a.xw, atom_i, raw_i_x, raw_i_y, raw_i_z);
    }
    Real pos_i_x = raw_i_x, pos_i_y = raw_i_y, pos_i_z = raw_i_z;

    const Real bx = static_cast<Real>(a.box[0]);
    const Real by = static_cast<Real>(a.box[4]);
    const Real bz = static_cast<Real>(a.box[8]);
    const Real inv_bx = 1 / bx;
    const Real inv_by = 1 / by;
    const Real inv_bz = 1 / bz;
    const Real half = static_cast<Real>(0.5);
    const Real zero = static_cast<Real>(0);

    // If the row block plus the cutoff fits in half a box, every candidate can be imaged once around the row-block
    // centre and plain (non-periodic) distances used (reference k_neighborlist.cuh:262-279).
    const bool single_box =
        (half * bx - row_ext_x >= a.cutoff && half * by - row_ext_y >= a.cutoff && half * bz - row_ext_z >= a.cutoff);
    Real np_i = 0;
    if (single_box) {
        pos_i_x -= bx * nearbyint((pos_i_x - row_ctr_x) * inv_bx);
        pos_i_y -= by * nearbyint((pos_i_y - row_ctr_y) * inv_by);
        pos_i_z -= bz * nearbyint((pos_i_z - row_ctr_z) * inv_bz);
        np_i = half * (pos_i_x * pos_i_x + pos_i_y * pos_i_y + pos_i_z * pos_i_z);
    }
    const Real cutoff2 = static_cast<Real>(a.cutoff) * static_cast<Real>(a.cutoff);

    const int num_col_blocks = (a.NC + TILE - 1) / TILE;
    const int num_chunks = (num_col_blocks + WARP - 1) / WARP;
    const int first_chunk = TRI ? row_block / WARP : 0;

    // Row atoms {x, y, z, |x|^2 / 2} (imaged around the row-block centre when single_box) in shared memory: the
    // atom-atom test below reads them with broadcast 128-bit loads instead of four shuffles per row atom.
    __shared__ Vec4<Real> s_row[TILE];
    if (warp == 0) {
        // a padding slot of the last row block is infinitely far from everything (the atom test below looks at every
        // atom of a group)
        const bool real_i = atom_i < static_cast<unsigned int>(N);
        s_row[lane] = Vec4<Real>{pos_i_x, pos_i_y, pos_i_z, real_i ? np_i : static_cast<Real>(INFINITY)};
    }
    __syncthreads();
    const Real half_cutoff2 = half * cutoff2;

    // Candidate column blocks are found a window of 32 chunks (32 x 32 blocks) at a time: the warps first run the
    // box-box test for the window (phase 1, one chunk per warp pass), then the candidates are dealt round-robin to the
    // warps (phase 2) - spatially sorted atoms put a row block's candidates into very few chunks, and handing whole
    // chunks to warps left most of them idle.  gridDim.y > 1 (few row blocks, e.g. a ligand against the whole
    // environment): the chunks are dealt over several CTAs per row block; each CTA closes its own partial tile.
    __shared__ unsigned int s_flags[WARP];
    const int n_set = max(0, (num_chunks - first_chunk - static_cast<int>(blockIdx.y) + static_cast<int>(gridDim.y) - 1) /
                                 static_cast<int>(gridDim.y));
    auto chunk_of = [&](int i) { return first_chunk + static_cast<int>(blockIdx.y) + i * static_cast<int>(gridDim.y); };

    for (int win = 0; win < n_set; win += WARP) {
        const int n_win = min(WARP, n_set - win);
        __syncthreads(); // previous window's flags are no longer needed
        for (int i = warp; i < n_win; i += BT_WARPS) {
            const int my_col_block = chunk_of(win + i) * WARP + lane;
            bool include = (my_col_block < num_col_blocks) && (!TRI || my_col_block >= row_block);
            if (include) {
                Real ddx = row_ctr_x - a.col_ctr[my_col_block * 3 + 0];
                Real ddy = row_ctr_y - a.col_ctr[my_col_block * 3 + 1];
                Real ddz = row_ctr_z - a.col_ctr[my_col_block * 3 + 2];
                ddx -= bx * nearbyint(ddx * inv_bx);
                ddy -= by * nearbyint(ddy * inv_by);
                ddz -= bz * nearbyint(ddz * inv_bz);
                ddx = max(zero, fabs(ddx) - row_ext_x - a.col_ext[my_col_block * 3 + 0]);
                ddy = max(zero, fabs(ddy) - row_ext_y - a.col_ext[my_col_block * 3 + 1]);
                ddz = max(zero, fabs(ddz) - row_ext_z - a.col_ext[my_col_block * 3 + 2]);
                include = (ddx * ddx + ddy * ddy + ddz * ddz) < cutoff2;
            }
            const unsigned int flags = __ballot_sync(0xffffffffu, include);
            if (lane == 0) {
                s_flags[i] = flags;
            }
        }
        __syncthreads();
        // lane i: candidates in window chunk i and before it (inclusive prefix) - identical in every warp
        const unsigned int my_flags = lane < n_win ? s_flags[lane] : 0u;
        int incl = __popc(my_flags);
        for (int d = 1; d < WARP; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) {
                incl += up;
            }
        }
        const int n_cand = __shfl_sync(0xffffffffu, incl, WARP - 1);

        // candidate number o (in chunk order, then block order) -> column block
        auto locate = [&](int o) {
            const int ci = __ffs(__ballot_sync(0xffffffffu, incl > o)) - 1;
            const unsigned int flags = __shfl_sync(0xffffffffu, my_flags, ci);
            const int before = __shfl_sync(0xffffffffu, incl - __popc(my_flags), ci);
            return chunk_of(win + ci) * WARP + static_cast<int>(__fns(flags, 0, o - before + 1));
        };
        // column block data {atoms, box} of the next candidate are fetched while the current one is tested
        struct Cand {
            unsigned int atom;
            Real x, y, z, cc_x, cc_y, cc_z, ce_x, ce_y, ce_z;
        };
        auto fetch = [&](int col_block) {
            Cand cd;
            const int cj = col_block * TILE + lane;
            cd.atom = cj < a.NC ? (a.col_idxs != nullptr ? a.col_idxs[cj] : static_cast<unsigned int>(a.col_base + cj))
                                : static_cast<unsigned int>(N);
            cd.x = cd.y = cd.z = 0;
            if (cd.atom < static_cast<unsigned int>(N)) {
                load_pos<Real>(a.coords, a.xw, cd.atom, cd.x, cd.y, cd.z);
            }
            cd.cc_x = a.col_ctr[col_block * 3 + 0];
            cd.cc_y = a.col_ctr[col_block * 3 + 1];
            cd.cc_z = a.col_ctr[col_block * 3 + 2];
            cd.ce_x = a.col_ext[col_block * 3 + 0];
            cd.ce_y = a.col_ext[col_block * 3 + 1];
            cd.ce_z = a.col_ext[col_block * 3 + 2];
            return cd;
        };
        Cand next;
        next.atom = N;
        if (warp < n_cand) {
            next = fetch(locate(warp));
        }
        for (int o = warp; o < n_cand; o += BT_WARPS) {
            const Cand cur = next;
            if (o + BT_WARPS < n_cand) {
                next = fetch(locate(o + BT_WARPS));
            }
            const unsigned int atom_j = cur.atom;
            Real pos_j_x = cur.x, pos_j_y = cur.y, pos_j_z = cur.z;
            const bool j_real = atom_j < static_cast<unsigned int>(N);
            const Real cc_x = cur.cc_x, cc_y = cur.cc_y, cc_z = cur.cc_z;
            const Real ce_x = cur.ce_x, ce_y = cur.ce_y, ce_z = cur.ce_z;

            // row atom vs column box (uses the un-imaged row coordinates)
            Real abx = raw_i_x - cc_x;
            Real aby = raw_i_y - cc_y;
            Real abz = raw_i_z - cc_z;
            abx -= bx * nearbyint(abx * inv_bx);
            aby -= by * nearbyint(aby * inv_by);
            abz -= bz * nearbyint(abz * inv_bz);
            abx = max(zero, fabs(abx) - ce_x);
            aby = max(zero, fabs(aby) - ce_y);
            abz = max(zero, fabs(abz) - ce_z);
            const bool row_near = atom_i < static_cast<unsigned int>(N) && (abx * abx + aby * aby + abz * abz) < cutoff2;
            const unsigned int row_flags = __ballot_sync(0xffffffffu, row_near);
            if (row_flags == 0) {
                continue; // no row atom reaches this column box: nothing of it is listed
            }

            bool interacts = false;
            if (single_box) {
                pos_j_x -= bx * nearbyint((pos_j_x - row_ctr_x) * inv_bx);
                pos_j_y -= by * nearbyint((pos_j_y - row_ctr_y) * inv_by);
                pos_j_z -= bz * nearbyint((pos_j_z - row_ctr_z) * inv_bz);
                const Real np_j = half * (pos_j_x * pos_j_x + pos_j_y * pos_j_y + pos_j_z * pos_j_z);
                // four row atoms per step, straight-line.  Groups none of whose atoms reaches the column box are skipped;
                // inside a group every atom is tested (one that does not reach the box cannot be within the cutoff of an
                // atom inside it, so membership is still "some row atom within the cutoff" as in the reference).  The
                // distance is three explicit FMAs: this file is compiled without contraction.
                for (int base = 0; base < TILE; base += 4) {
                    const unsigned int bits = (row_flags >> base) & 0xFu;
                    if (bits == 0) {
                        continue;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const Vec4<Real> r = s_row[base + u];
                        const Real half_d2 = fma(-r.z, pos_j_z, fma(-r.y, pos_j_y, fma(-r.x, pos_j_x, r.w + np_j)));
                        interacts |= half_d2 < half_cutoff2;
                    }
                    // once every column atom is known to interact there is nothing left to learn
                    if (__all_sync(0xffffffffu, interacts)) {
                        break;
                    }
                }
            } else {
                unsigned int todo = row_flags;
                while (todo != 0) {
                    const int row_atom = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const Vec4<Real> r = s_row[row_atom];
                    Real dx = r.x - pos_j_x;
                    Real dy = r.y - pos_j_y;
                    Real dz = r.z - pos_j_z;
                    dx -= bx * nearbyint(dx * inv_bx);
                    dy -= by * nearbyint(dy * inv_by);
                    dz -= bz * nearbyint(dz * inv_bz);
                    interacts |= (dx * dx + dy * dy + dz * dz) < cutoff2;
                    if (__all_sync(0xffffffffu, interacts)) {
                        break;
                    }
                }
            }
            w.append(interacts && j_real, static_cast<int>(atom_j));
        }
    }

    // merge the leftovers (<= 32 per warp) of warps 1..3 into warp 0 and finish the row block
    if (warp != 0) {
        w.publish();
        if (lane == 0) {
            s_left[warp] = w.nbuf;
        }
    }
    __syncthreads();
    if (warp == 0) {
        for (int other = 1; other < BT_WARPS; other++) {
            const int n = s_left[other];
            const int v = lane < n ? s_buf[other][lane] : N;
            w.append(lane < n, v);
        }
        if (w.nbuf > 0) {
            w.emit(w.nbuf);
            w.nbuf = 0;
        }
        w.publish();
    }
}

template <typename Real> void launch_build_tiles(const BuildTilesArgs<Real> &args, cudaStream_t stream) {
    const int row_blocks = ceil_div(args.NR, TILE);
    if (row_blocks <= 0) {
        return;
    }
    if (args.upper_triangular) {
        TMB_LAUNCH((k_build_tiles<Real, true>), row_blocks, BT_THREADS, 0, stream, args);
    } else {
        const int num_chunks = ceil_div(ceil_div(args.NC, TILE), WARP);
        int ny = ceil_div(2 * sm_count(), row_blocks);
        ny = std::max(1, std::min(ny, ceil_div(num_chunks, BT_WARPS)));
        TMB_LAUNCH((k_build_tiles<Real, false>), dim3(row_blocks, ny), BT_THREADS, 0, stream, args);
    }
}
template void launch_build_tiles<float>(const BuildTilesArgs<float> &, cudaStream_t);
template void launch_build_tiles<double>(const BuildTilesArgs<double> &, cudaStream_t);

} // namespace tmb
