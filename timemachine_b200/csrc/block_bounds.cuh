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

// Same purpose, O(log 32) depth instead of a 32-step dependent chain: every atom is imaged next to lane 0's atom and
// the extrema are taken with butterfly shuffles.  The box encloses the same periodic images whenever the block spans
// less than half the box (always, for spatially sorted blocks), but it is not bit-for-bit the box of the sequential
// recentring above; it is used where only validity matters (the per-step bounds of the all-pairs potential, whose
// results do not depend on which valid boxes the list was culled with), not by the stand-alone Neighborlist.
template <typename Real>
__device__ __forceinline__ void warp_block_bounds_anchor(
    Real px, Real py, Real pz, int n_valid, Real bx, Real by, Real bz, Real inv_bx, Real inv_by, Real inv_bz, Real ctr[3],
    Real ext[3]) {
    const int lane = threadIdx.x & 31;
    const Real ax = __shfl_sync(0xffffffffu, px, 0);
    const Real ay = __shfl_sync(0xffffffffu, py, 0);
    const Real az = __shfl_sync(0xffffffffu, pz, 0);
    const bool valid = lane < n_valid;
    const Real qx = valid ? px - bx * nearbyint((px - ax) * inv_bx) : ax;
    const Real qy = valid ? py - by * nearbyint((py - ay) * inv_by) : ay;
    const Real qz = valid ? pz - bz * nearbyint((pz - az) * inv_bz) : az;
    Real lo_x = qx, hi_x = qx, lo_y = qy, hi_y = qy, lo_z = qz, hi_z = qz;
    for (int d = 16; d > 0; d >>= 1) {
        lo_x = min(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, d));
        hi_x = max(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, d));
        lo_y = min(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, d));
        hi_y = max(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, d));
        lo_z = min(lo_z, __shfl_xor_sync(0xffffffffu, lo_z, d));
        hi_z = max(hi_z, __shfl_xor_sync(0xffffffffu, hi_z, d));
    }
    const Real half = static_cast<Real>(0.5);
    ctr[0] = half * (hi_x + lo_x);
    ctr[1] = half * (hi_y + lo_y);
    ctr[2] = half * (hi_z + lo_z);
    ext[0] = half * (hi_x - lo_x);
    ext[1] = half * (hi_y - lo_y);
    ext[2] = half * (hi_z - lo_z);
}

} // namespace tmb
