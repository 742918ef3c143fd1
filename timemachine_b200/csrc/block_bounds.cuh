// Tight periodic bounding box of the (up to) 32 atoms held one per lane by a warp.
//
// The running re-centring is order dependent, so the atoms are visited in the reference's order (1, 2, ..., 31, 0
// relative to the block start; timemachine/cpp/src/kernels/k_neighborlist.cuh:84-106) - block bounds, hence the tile
// list, come out identical to the reference's.  Shared by k_block_bounds (stand-alone Neighborlist) and k_nb_prepare
// (fused into the per-step gather of the all-pairs potential).
#pragma once

#include "common.cuh"

namespace tmb {

// n_valid: number of real atoms in this block (lanes [0, n_valid) hold them; lane 0 always does). All lanes must call.
template <typename Real>
__device__ __forceinline__ void warp_block_bounds(
    Real px, Real py, Real pz, int n_valid, Real bx, Real by, Real bz, Real inv_bx, Real inv_by, Real inv_bz, Real ctr[3],
    Real ext[3]) {
    Real min_x = __shfl_sync(0xffffffffu, px, 0), max_x = min_x;
    Real min_y = __shfl_sync(0xffffffffu, py, 0), max_y = min_y;
    Real min_z = __shfl_sync(0xffffffffu, pz, 0), max_z = min_z;
    const Real half = static_cast<Real>(0.5);
    for (int it = 1; it <= WARP; it++) {
        const int src = it & 31;
        const Real qx = __shfl_sync(0xffffffffu, px, src);
        const Real qy = __shfl_sync(0xffffffffu, py, src);
        const Real qz = __shfl_sync(0xffffffffu, pz, src);
        if (src < n_valid) {
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
    ctr[0] = half * (max_x + min_x);
    ctr[1] = half * (max_y + min_y);
    ctr[2] = half * (max_z + min_z);
    ext[0] = half * (max_x - min_x);
    ext[1] = half * (max_y - min_y);
    ext[2] = half * (max_z - min_z);
}

} // namespace tmb
