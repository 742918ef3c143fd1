// Deterministic int128 energy reduction: warp shuffles -> per-CTA partial -> last CTA sums all partials.
// Integer addition is associative, so the result is independent of scheduling (the reference gets the same
// property from cub::BlockReduce + cub::DeviceReduce over int128, k_nonbonded.cuh:419-431).
#pragma once

#include "common.cuh"

namespace tmb {

__device__ __forceinline__ i128 shfl_down_i128(i128 v, int offset) {
    u64 lo = static_cast<u64>(v);
    u64 hi = static_cast<u64>(static_cast<unsigned __int128>(v) >> 64);
    lo = __shfl_down_sync(0xffffffffu, lo, offset);
    hi = __shfl_down_sync(0xffffffffu, hi, offset);
    return static_cast<i128>((static_cast<unsigned __int128>(hi) << 64) | lo);
}

__device__ __forceinline__ i128 warp_sum_i128(i128 v) {
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) {
        v += shfl_down_i128(v, offset);
    }
    return v; // valid in lane 0
}

// Sum `v` over the CTA; result valid in thread 0. `scratch` must hold blockDim.x/32 entries.
__device__ __forceinline__ i128 block_sum_i128(i128 v, i128 *scratch) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    v = warp_sum_i128(v);
    __syncthreads(); // scratch may be reused across calls
    if (lane == 0) {
        scratch[warp] = v;
    }
    __syncthreads();
    i128 total = 0;
    if (warp == 0) {
        total = lane < nwarps ? scratch[lane] : static_cast<i128>(0);
        total = warp_sum_i128(total);
    }
    return total;
}

// Each CTA deposits its partial; the CTA that takes the last ticket sums every partial and OVERWRITES *d_u
// (execute_device overwrites energies, reference SURVEY cheat-sheet item 7). Resets the ticket for the next launch.
// Must be called by every thread of every CTA of the grid. `scratch`: blockDim.x/32 int128 of shared memory.
__device__ __forceinline__ void grid_finish_energy(
    i128 thread_val, i128 *scratch, i128 *__restrict__ partials, unsigned int *__restrict__ ticket, i128 *__restrict__ d_u) {
    __shared__ bool is_last;
    i128 block_total = block_sum_i128(thread_val, scratch);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_total;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        i128 acc = 0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
            acc += __ldcg(reinterpret_cast<const unsigned long long *>(partials + i)) |
                   (static_cast<i128>(__ldcg(reinterpret_cast<const long long *>(partials + i) + 1)) << 64);
        }
        i128 total = block_sum_i128(acc, scratch);
        if (threadIdx.x == 0) {
            *d_u = total;
            *ticket = 0;
        }
    }
}

} // namespace tmb
