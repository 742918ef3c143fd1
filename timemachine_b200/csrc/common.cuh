// Common host/device utilities for the timemachine_b200 hot path (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace tmb {

typedef unsigned long long u64;
typedef long long i64;
typedef __int128 i128;

constexpr int WARP = 32;
constexpr int TILE = 32; // atoms per neighbour-list block; one warp lane per atom

// Nonbonded per-atom parameter layout at the API boundary (f64[N,4]):
// reference timemachine/cpp/src/nonbonded_common.hpp:12
enum { P_CHARGE = 0, P_SIG = 1, P_EPS = 2, P_W = 3, P_PER_ATOM = 4 };

// Raised for device-missing class errors (reference gpu_utils.cuh:31-53 -> custom_ops.InvalidHardware)
struct InvalidHardware : public std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t code, const char *what, const char *file, int line) {
    if (code == cudaSuccess)
        return;
    std::string msg = std::string("CUDA error: ") + cudaGetErrorString(code) + " (" + what + ") at " + file + ":" +
                      std::to_string(line);
    switch (code) {
    case cudaErrorInvalidDevice:
    case cudaErrorInsufficientDriver:
    case cudaErrorNoDevice:
    case cudaErrorStartupFailure:
    case cudaErrorInvalidPtx:
    case cudaErrorUnsupportedPtxVersion:
    case cudaErrorDevicesUnavailable:
    case cudaErrorNoKernelImageForDevice:
        throw InvalidHardware(msg);
    default:
        throw std::runtime_error(msg);
    }
}

#define TMB_CUDA(expr) ::tmb::cuda_check((expr), #expr, __FILE__, __LINE__)

// Count of kernels this library has launched (bench.py reports it as gpu_launches).
extern std::atomic<long long> g_kernel_launches;

// Bumped by every host-side mutation that changes what a step ENQUEUES (kernel arguments, grid sizes, buffer addresses):
// set_atom_idxs, a BoundPotential changing its parameter count, the integrator's noise source, kernel-timing hooks.  A
// Context remembers the value its CUDA graph was captured under and re-captures when it has moved (context.cu).
extern std::atomic<long long> g_launch_generation;
inline void bump_launch_generation() { g_launch_generation.fetch_add(1, std::memory_order_relaxed); }

#define TMB_LAUNCH(kernel, grid, block, smem, stream, ...)                                                             \
    do {                                                                                                               \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                                    \
        ::tmb::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);                                              \
        TMB_CUDA(cudaPeekAtLastError());                                                                               \
    } while (0)

template <typename T> inline T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T> inline T round_up(T a, T b) { return ceil_div(a, b) * b; }

int sm_count();

} // namespace tmb
