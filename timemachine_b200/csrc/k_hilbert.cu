// Hilbert-curve spatial sort: cell -> curve-index table, coordinate -> key kernel, stable radix sort.
// Reference: hilbert_sort.cu:13-81, kernels/k_hilbert.cu:9-54.  The LUT is generated on the device in one launch
// (the reference fills 2M entries on the host through the vendored C routine and uploads 8 MB).
#include "hilbert_curve.h"
#include "kernels.hpp"

#include <cub/cub.cuh>

namespace tmb {

__global__ void k_hilbert_lut(unsigned int *__restrict__ lut) {
    const unsigned int cell = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr unsigned int G = HILBERT_GRID_DIM;
    if (cell >= G * G * G) {
        return;
    }
    const unsigned int i = cell / (G * G);
    const unsigned int j = (cell / G) % G;
    const unsigned int k = cell % G;
    lut[cell] = hilbert3d_index(i, j, k, HILBERT_N_BITS);
}

void launch_hilbert_lut(unsigned int *lut, cudaStream_t stream) {
    constexpr int G = HILBERT_GRID_DIM;
    TMB_LAUNCH(k_hilbert_lut, ceil_div(G * G * G, 256), 256, 0, stream, lut);
}

// key = curve index of the 128^3 cell containing the home-box image of the atom.  Double precision on purpose:
// imaging with floor() in f32 can land outside the home box for large coordinates (reference k_hilbert.cu:5-8).
__global__ void k_hilbert_keys(
    const int n,
    const unsigned int *__restrict__ atom_idxs,
    const double *__restrict__ coords,
    const double *__restrict__ box,
    const unsigned int *__restrict__ lut,
    unsigned int *__restrict__ keys,
    unsigned int *__restrict__ vals) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) {
        return;
    }
    const double bx = box[0];
    const double by = box[4];
    const double bz = box[8];
    const double inv_bx = 1 / bx;
    const double inv_by = 1 / by;
    const double inv_bz = 1 / bz;
    const double inv_bin_width = min(min(inv_bx, inv_by), inv_bz) * (HILBERT_GRID_DIM - 1.0);

    const unsigned int atom = atom_idxs[idx];
    double x = coords[static_cast<size_t>(atom) * 3 + 0];
    double y = coords[static_cast<size_t>(atom) * 3 + 1];
    double z = coords[static_cast<size_t>(atom) * 3 + 2];
    x -= bx * floor(x * inv_bx);
    y -= by * floor(y * inv_by);
    z -= bz * floor(z * inv_bz);
    const unsigned int bin_x = static_cast<unsigned int>(x * inv_bin_width);
    const unsigned int bin_y = static_cast<unsigned int>(y * inv_bin_width);
    const unsigned int bin_z = static_cast<unsigned int>(z * inv_bin_width);
    constexpr unsigned int G = HILBERT_GRID_DIM;
    // clamp defensively: a coordinate exactly on the upper face maps to bin 127 by construction, NaNs map to 0
    const unsigned int cell = min(bin_x, G - 1) * G * G + min(bin_y, G - 1) * G + min(bin_z, G - 1);
    keys[idx] = lut[cell];
    vals[idx] = atom;
}

void launch_hilbert_keys(
    int n, const unsigned int *atom_idxs, const double *coords, const double *box, const unsigned int *lut,
    unsigned int *keys, unsigned int *vals, cudaStream_t stream) {
    if (n <= 0) {
        return;
    }
    TMB_LAUNCH(k_hilbert_keys, ceil_div(n, 128), 128, 0, stream, n, atom_idxs, coords, box, lut, keys, vals);
}

// Keys use 21 bits (7 bits per axis: the grid is 128^3); a stable LSD radix sort over those bits gives the same
// permutation as the reference's 32-bit sort (hilbert_sort.cu:69-80) because the upper bits are all zero.
constexpr int HILBERT_KEY_BITS = 21;

size_t radix_sort_pairs_temp_bytes(int n) {
    size_t bytes = 0;
    unsigned int *nul = nullptr;
    TMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, nul, nul, nul, nul, n, 0, HILBERT_KEY_BITS));
    return bytes;
}

void radix_sort_pairs(
    void *temp, size_t temp_bytes, const unsigned int *keys_in, unsigned int *keys_out, const unsigned int *vals_in,
    unsigned int *vals_out, int n, cudaStream_t stream) {
    TMB_CUDA(cub::DeviceRadixSort::SortPairs(
        temp, temp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, HILBERT_KEY_BITS, stream));
}

} // namespace tmb
