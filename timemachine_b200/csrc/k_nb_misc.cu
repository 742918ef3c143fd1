// Gather/cast + device-side rebuild decision and small utilities.
#include "block_bounds.cuh"
#include "fixed_point.cuh"
#include "kernels.hpp"
#include "reduce.cuh"

namespace tmb {

constexpr int MISC_THREADS = 128;

// One thread per sorted slot: gather atom perm[k] from the f64 API buffers, cast to Real and pack for 128-bit loads.
// In the same pass decide whether the neighbour list must be rebuilt: any atom moved more than padding/2 since the
// last build, or the box changed (reference k_check_rebuild_coords_and_box_gather, k_nonbonded.cuh:11-56).  The flag
// stays on the device: the reference copies it to the host and blocks on an event EVERY step
// (nonbonded_all_pairs.cu:217-235).  The flag is only ever SET here; the tile kernel of the same evaluation, which
// runs after the (conditional) build in stream order, clears it.
template <typename Real> __global__ void __launch_bounds__(MISC_THREADS) k_nb_prepare(const NbPrepareArgs<Real> a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    bool rebuild = false;
    if (k == 0) {
        *a.tile_cursor = 0;
        rebuild = a.force_rebuild != 0;
    }
    // Box changes (barostat proposals) are treated like displacements instead of forcing a rebuild as the reference does
    // (k_nonbonded.cuh:24-30): a periodic image moves by the atom's displacement plus the change of the box vector, so
    // a pair can approach by at most d_i + d_j + |dbox| and the list stays complete while every atom has moved less than
    // (padding - |dbox|) / 2.  Only a change of at least the padding, or an off-diagonal element, rebuilds by itself.
    double dbox = 0; // |change of the image vector (+-1, +-1, +-1) . box|, the largest any minimum image can shift by
    for (int c = 0; c < 3; c++) {
        const double dc = a.box[c * 4] - a.box_build[c * 4];
        dbox += dc * dc;
    }
    dbox = sqrt(dbox);
    rebuild = rebuild || !(dbox < a.padding); // also catches NaN
    if (k < 9 && (k % 4) != 0) {
        rebuild = rebuild || (a.box[k] != a.box_build[k]);
    }
    const double half_room = 0.5 * (a.padding - dbox);
    Vec4<Real> c = {0, 0, 0, 0};
    if (k < a.K) {
        const unsigned int atom = a.perm[k];
        const double x = a.x[atom * 3 + 0];
        const double y = a.x[atom * 3 + 1];
        const double z = a.x[atom * 3 + 2];
        const double *p = a.p + static_cast<size_t>(atom) * P_PER_ATOM;

        Vec4<Real> q;
        c.x = static_cast<Real>(x);
        c.y = static_cast<Real>(y);
        c.z = static_cast<Real>(z);
        c.w = static_cast<Real>(p[P_W]);
        q.x = static_cast<Real>(p[P_CHARGE]);
        q.y = static_cast<Real>(p[P_SIG]);
        q.z = static_cast<Real>(p[P_EPS]);
        q.w = static_cast<Real>(0);
        a.xw[k] = c;
        a.qse[k] = q;

        // the snapshot is kept in sorted order, already cast: one coalesced 16-byte load (a new permutation always
        // forces a rebuild, so slot k holds the same atom as at build time whenever the comparison matters)
        const Vec4<Real> o = a.xw_build[k];
        const Real dx = o.x - c.x;
        const Real dy = o.y - c.y;
        const Real dz = o.z - c.z;
        const Real d2 = dx * dx + dy * dy + dz * dz;
        rebuild = rebuild || (static_cast<double>(d2) > half_room * half_room);
    }
    if (rebuild) {
        // benign races: every writer stores the same values, and nothing reads them before the kernel ends
        *a.flag = 1;
        if (a.reset_count != nullptr) {
            *a.reset_count = 0;
            *a.reset_overflow = 0;
        }
    }
    if (a.ctr != nullptr) {
        // warp w of the grid holds block w (MISC_THREADS is a multiple of 32, k is lane-contiguous)
        const int block = k / WARP;
        if (block * WARP < a.K) {
            const Real bx = static_cast<Real>(a.box[0]);
            const Real by = static_cast<Real>(a.box[4]);
            const Real bz = static_cast<Real>(a.box[8]);
            Real ctr[3], ext[3];
            warp_block_bounds_anchor<Real>(c.x, c.y, c.z, a.K - block * WARP, bx, by, bz, 1 / bx, 1 / by, 1 / bz, ctr, ext);
            if ((threadIdx.x & 31) == 0) {
                for (int d = 0; d < 3; d++) {
                    a.ctr[block * 3 + d] = ctr[d];
                    a.ext[block * 3 + d] = ext[d];
                }
            }
        }
    }
}

template <typename Real> void launch_nb_prepare(const NbPrepareArgs<Real> &args, cudaStream_t stream) {
    const int n = args.K > 9 ? args.K : 9;
    TMB_LAUNCH(k_nb_prepare<Real>, ceil_div(n, MISC_THREADS), MISC_THREADS, 0, stream, args);
}
template void launch_nb_prepare<float>(const NbPrepareArgs<float> &, cudaStream_t);
template void launch_nb_prepare<double>(const NbPrepareArgs<double> &, cudaStream_t);

// single-CTA int128 sum (used by Summed/Fanout potentials over a handful of child energies)
__global__ void __launch_bounds__(256) k_sum_i128(const i128 *__restrict__ in, const int n, i128 *__restrict__ out) {
    __shared__ i128 scratch[8];
    i128 acc = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        acc += in[i];
    }
    i128 total = block_sum_i128(acc, scratch);
    if (threadIdx.x == 0) {
        *out = total;
    }
}

void launch_sum_i128(const i128 *in, int n, i128 *out, cudaStream_t stream) {
    TMB_LAUNCH(k_sum_i128, 1, 256, 0, stream, in, n, out);
}

__global__ void k_iota(unsigned int *__restrict__ out, const int n, const unsigned int base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        out[i] = base + i;
    }
}

void launch_iota(unsigned int *out, int n, unsigned int base, cudaStream_t stream) {
    if (n <= 0) {
        return;
    }
    TMB_LAUNCH(k_iota, ceil_div(n, MISC_THREADS), MISC_THREADS, 0, stream, out, n, base);
}

} // namespace tmb
