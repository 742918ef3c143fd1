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
        if (a.x_build != nullptr) {
            // remember where this atom was when the list was built (reference nonbonded_all_pairs.cu:241-242)
            const size_t src = static_cast<size_t>(a.perm[atom]) * 3;
            a.x_build[src + 0] = a.x_src[src + 0];
            a.x_build[src + 1] = a.x_src[src + 1];
            a.x_build[src + 2] = a.x_src[src + 2];
        }
    }
    // lane 0 of a block is always a real atom
    Real min_x = __shfl_sync(0xffffffffu, px, 0), max_x = min_x;
    Real min_y = __shfl_sync(0xffffffffu, py, 0), max_y = min_y;
    Real min_z = __shfl_sync(0xffffffffu, pz, 0), max_z = min_z;
    const Real half = static_cast<Real>(0.5);
    for (int it = 1; it <= WARP; it++) {
        const int src = it & 31;
        const Real qx = __shfl_sync(0xffffffffu, px, src);
        const Real qy = __shfl_sync(0xffffffffu, py, src);
        const Real qz = __shfl_sync(0xffffffffu, pz, src);
        const bool src_valid = (block * TILE + src) < a.num_idxs;
        if (src_valid) {
            Real im = qx - bx * nearbyint((qx - half * (max_x + min_x)) * inv_bx);
            min_x = min(min_x, im);
            max_x = max(max_x, im);
            im = qy - by * nearbyint((qy - half * (max_y + min_y)) * inv_by);
            min_y = min(min_y, im);
            max_y = max(max_y, im);
            im = qz - bz * nearbyint((qz - half * (max_z + min_z)) * inv_bz);
            min_z = min(min_z, im);
            max_z = max(max_z, im);
        }
    }
    if (lane == 0) {
        a.ctr[block * 3 + 0] = half * (max_x + min_x);
        a.ctr[block * 3 + 1] = half * (max_y + min_y);
        a.ctr[block * 3 + 2] = half * (max_z + min_z);
        a.ext[block * 3 + 0] = half * (max_x - min_x);
        a.ext[block * 3 + 1] = half * (max_y - min_y);
        a.ext[block * 3 + 2] = half * (max_z - min_z);
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

__global__ void k_reset_tile_count(unsigned int *count, unsigned int *overflow, const unsigned int *flag) {
    if (flag != nullptr && *flag == 0) {
        return;
    }
    *count = 0;
    *overflow = 0;
}

void launch_reset_tile_count(const TileList &tiles, const unsigned int *flag, cudaStream_t stream) {
    TMB_LAUNCH(k_reset_tile_count, 1, 1, 0, stream, tiles.count, tiles.overflow, flag);
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int BT_WARPS = 4;
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

    // gridDim.y > 1 (few row blocks, e.g. a ligand against the whole environment): the column chunks are dealt over
    // several CTAs per row block; each CTA then closes its own partial tile
    for (int chunk = first_chunk + blockIdx.y * BT_WARPS + warp; chunk < num_chunks; chunk += BT_WARPS * gridDim.y) {
        const int col_block_base = chunk * WARP;
        const int my_col_block = col_block_base + lane;
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
        unsigned int block_flags = __ballot_sync(0xffffffffu, include);

        while (block_flags != 0) {
            const int offset = __ffs(block_flags) - 1;
            block_flags &= block_flags - 1;
            const int col_block = col_block_base + offset;
            const int cj = col_block * TILE + lane;
            const unsigned int atom_j =
                cj < a.NC ? (a.col_idxs != nullptr ? a.col_idxs[cj] : static_cast<unsigned int>(a.col_base + cj))
                          : static_cast<unsigned int>(N);
            const bool j_real = atom_j < static_cast<unsigned int>(N);

            const Real cc_x = a.col_ctr[col_block * 3 + 0];
            const Real cc_y = a.col_ctr[col_block * 3 + 1];
            const Real cc_z = a.col_ctr[col_block * 3 + 2];
            const Real ce_x = a.col_ext[col_block * 3 + 0];
            const Real ce_y = a.col_ext[col_block * 3 + 1];
            const Real ce_z = a.col_ext[col_block * 3 + 2];

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
            unsigned int row_flags = __ballot_sync(0xffffffffu, row_near);

            Real pos_j_x = 0, pos_j_y = 0, pos_j_z = 0;
            if (j_real) {
                load_pos<Real>(a.coords, a.xw, atom_j, pos_j_x, pos_j_y, pos_j_z);
            }
            Real np_j = 0;
            if (single_box) {
                pos_j_x -= bx * nearbyint((pos_j_x - row_ctr_x) * inv_bx);
                pos_j_y -= by * nearbyint((pos_j_y - row_ctr_y) * inv_by);
                pos_j_z -= bz * nearbyint((pos_j_z - row_ctr_z) * inv_bz);
                np_j = half * (pos_j_x * pos_j_x + pos_j_y * pos_j_y + pos_j_z * pos_j_z);
            }

            bool interacts = false;
            while (row_flags != 0) {
                const int row_atom = __ffs(row_flags) - 1;
                row_flags &= row_flags - 1;
                const Real rx = __shfl_sync(0xffffffffu, pos_i_x, row_atom);
                const Real ry = __shfl_sync(0xffffffffu, pos_i_y, row_atom);
                const Real rz = __shfl_sync(0xffffffffu, pos_i_z, row_atom);
                if (!single_box) {
                    Real dx = rx - pos_j_x;
                    Real dy = ry - pos_j_y;
                    Real dz = rz - pos_j_z;
                    dx -= bx * nearbyint(dx * inv_bx);
                    dy -= by * nearbyint(dy * inv_by);
                    dz -= bz * nearbyint(dz * inv_bz);
                    interacts |= (dx * dx + dy * dy + dz * dz) < cutoff2;
                } else {
                    const Real ci = __shfl_sync(0xffffffffu, np_i, row_atom);
                    const Real half_d2 = ci + np_j - rx * pos_j_x - ry * pos_j_y - rz * pos_j_z;
                    interacts |= half_d2 < (half * cutoff2);
                }
                // once every column atom is known to interact there is nothing left to learn
                if (__all_sync(0xffffffffu, interacts)) {
                    break;
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
