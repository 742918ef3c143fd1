// Langevin BAOAB update (rotated by half a step) with the thermostat noise generated inside the kernel.
//
// Arithmetic is the reference's mixed-precision sequence exactly as written (k_integrator.cuh:32-46; SURVEY.md §8a
// cheat-sheet item 9): f32 coefficients and force, f64 state.
//     F      = -(float)( (float)(int64)du_dx / 2^36 )
//     v_mid  = (float)( v_f64 + (double)(cb * F) )
//     v_new  = fma(ca, v_mid, cc * xi)                      (f32; the reference's source line is contracted by nvcc)
//     x_f64 += (double)(0.5f * dt) * ( (double)v_mid + v_new_f64 )
// and du_dx is zeroed for the next step.
//
// Noise: the reference refills a 10-step cuRAND (XORWOW) buffer from the host every 10 steps
// (langevin_integrator.cu:74-79).  Here a counter-based Philox4x32-10 keyed by (seed; atom, step) feeds Box-Muller
// in registers: no noise buffer in HBM, no extra launches, and a trajectory is reproducible from (seed, step) alone.
// The two generators cannot produce the same stream, so integrator parity is tested at friction = 0 (cc = 0), as the
// reference's own tests/test_md.py:174-178 does, or with an externally supplied noise buffer.
#include "block_bounds.cuh"
#include "fixed_point.cuh"
#include "kernels.hpp"

namespace tmb {

constexpr int INT_THREADS = 128;

struct Philox4 {
    unsigned int x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(Philox4 c, unsigned int k0, unsigned int k1) {
    constexpr unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned int hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const unsigned int hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = {hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
        k0 += W0;
        k1 += W1;
    }
    return c;
}

__device__ __forceinline__ float u32_to_open_unit(unsigned int v) {
    // 24 random bits, centred in their cell: strictly inside (0, 1)
    return (static_cast<float>(v >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// three independent standard normals for (atom, step)
__device__ __forceinline__ void normal3(unsigned long long seed, unsigned long long step, unsigned int atom, float &n0, float &n1, float &n2) {
    Philox4 ctr = {atom, static_cast<unsigned int>(step), static_cast<unsigned int>(step >> 32), 0x4c616e67u};
    Philox4 r = philox4x32_10(ctr, static_cast<unsigned int>(seed), static_cast<unsigned int>(seed >> 32));
    const float two_pi = 6.283185307179586f;
    float ra = sqrtf(-2.0f * logf(u32_to_open_unit(r.x)));
    float rb = sqrtf(-2.0f * logf(u32_to_open_unit(r.z)));
    float sa, ca, sb, cb;
    sincosf(two_pi * u32_to_open_unit(r.y), &sa, &ca);
    sincosf(two_pi * u32_to_open_unit(r.w), &sb, &cb);
    n0 = ra * ca;
    n1 = ra * sa;
    n2 = rb * cb;
    (void)sb;
}

// One atom: the mixed-precision update of the file header; returns the new coordinates.
__device__ __forceinline__ void baoab_atom(const BaoabArgs &a, const int atom, const unsigned long long step, double out[3]) {
    const float cb = a.cbs[atom];
    const float cc = a.ccs[atom];
    float xi[3];
    if (a.noise != nullptr) {
        xi[0] = a.noise[atom * 3 + 0];
        xi[1] = a.noise[atom * 3 + 1];
        xi[2] = a.noise[atom * 3 + 2];
    } else {
        normal3(a.seed, step, static_cast<unsigned int>(atom), xi[0], xi[1], xi[2]);
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const int q = atom * 3 + d;
        const float force = -fixed_to_real<float>(a.du_dx[q]);
        const float v_mid = static_cast<float>(a.v[q] + static_cast<double>(cb * force));
        // one FMA, as nvcc compiles the reference's `ca * v_mid + ccs * noise` (FFMA in its k_update_forward_baoab<float>)
        const float v_new = __fmaf_rn(a.ca, v_mid, cc * xi[d]);
        const double v_new_d = static_cast<double>(v_new);
        a.v[q] = v_new_d;
        const double x_new = a.x[q] + static_cast<double>(0.5f * a.dt) * (static_cast<double>(v_mid) + v_new_d);
        a.x[q] = x_new;
        out[d] = x_new;
        a.du_dx[q] = 0;
    }
}

__global__ void __launch_bounds__(INT_THREADS) k_baoab(const BaoabArgs a) {
    const unsigned long long step = a.step + (a.step_base != nullptr ? *a.step_base : 0ull);
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < a.N; tid += gridDim.x * blockDim.x) {
        const int atom = a.idxs == nullptr ? tid : static_cast<int>(a.idxs[tid]);
        if (atom < a.N) {
            double xn[3];
            baoab_atom(a, atom, step, xn);
        } else if (a.idxs != nullptr) {
            // local-MD convention: idxs[tid] == N marks a frozen atom; its force slot still has to be cleared
            a.du_dx[tid * 3 + 0] = 0;
            a.du_dx[tid * 3 + 1] = 0;
            a.du_dx[tid * 3 + 2] = 0;
        }
    }
}

// BAOAB over the sorted slots of an all-pairs potential + that potential's next k_nb_prepare (kernels.hpp).  Thread k of
// the first Kpad threads owns slot k (warp = 32-atom block, as in k_nb_prepare); the threads behind them integrate the
// atoms that are not in the potential's set.  The prepare half is k_nb_prepare's code path for force_rebuild == 0 with the
// parameter-derived fields (w, qse) left as the block's first prepare wrote them.
template <typename Real> __global__ void __launch_bounds__(INT_THREADS) k_baoab_prepare(const BaoabArgs a, const FusedPrepareArgs<Real> f) {
    const unsigned long long step = a.step + (a.step_base != nullptr ? *a.step_base : 0ull);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int kpad = (f.K + WARP - 1) / WARP * WARP;
    if (k >= kpad) {
        const int o = k - kpad;
        if (o < f.n_others) {
            double xn[3];
            baoab_atom(a, static_cast<int>(f.others[o]), step, xn);
        }
        return;
    }
    bool rebuild = false;
    if (k == 0) {
        *f.tile_cursor = 0;
    }
    double dbox = 0;
    for (int c = 0; c < 3; c++) {
        const double dc = f.box[c * 4] - f.box_build[c * 4];
        dbox += dc * dc;
    }
    dbox = sqrt(dbox);
    rebuild = rebuild || !(dbox < f.padding);
    if (k < 9 && (k % 4) != 0) {
        rebuild = rebuild || (f.box[k] != f.box_build[k]);
    }
    const double half_room = 0.5 * (f.padding - dbox);
    Vec4<Real> c = {0, 0, 0, 0};
    if (k < f.K) {
        double xn[3];
        baoab_atom(a, static_cast<int>(f.perm[k]), step, xn);
        c.x = static_cast<Real>(xn[0]);
        c.y = static_cast<Real>(xn[1]);
        c.z = static_cast<Real>(xn[2]);
        c.w = f.xw[k].w;
        f.xw[k] = c;
        const Vec4<Real> o = f.xw_build[k];
        const Real dx = o.x - c.x;
        const Real dy = o.y - c.y;
        const Real dz = o.z - c.z;
        const Real d2 = dx * dx + dy * dy + dz * dz;
        rebuild = rebuild || (static_cast<double>(d2) > half_room * half_room);
    }
    if (rebuild) {
        *f.flag = 1;
        *f.reset_count = 0;
        *f.reset_overflow = 0;
    }
    const int block = k / WARP;
    const Real bx = static_cast<Real>(f.box[0]);
    const Real by = static_cast<Real>(f.box[4]);
    const Real bz = static_cast<Real>(f.box[8]);
    Real ctr[3], ext[3];
    warp_block_bounds_anchor<Real>(c.x, c.y, c.z, f.K - block * WARP, bx, by, bz, 1 / bx, 1 / by, 1 / bz, ctr, ext);
    if ((threadIdx.x & 31) == 0) {
        for (int d = 0; d < 3; d++) {
            f.ctr[block * 3 + d] = ctr[d];
            f.ext[block * 3 + d] = ext[d];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Velocity Verlet, all f64 (reference k_integrator.cuh:64-130).  MODE 0: a whole step  v += cbs F, x += dt v;
// MODE 1: the opening half kick + drift  v += (cbs / 2) F, x += dt v;  MODE 2: the closing half kick.  Each update is the
// single FMA the reference's `a += b * c` statements compile to.  The force buffer is consumed and cleared (the reference
// clears it with a memset before every evaluation; here the next evaluation finds it zero like the Langevin path does).
template <int MODE> __global__ void __launch_bounds__(INT_THREADS) k_velocity_verlet(const VerletArgs a) {
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < a.N; tid += gridDim.x * blockDim.x) {
        const int atom = a.idxs == nullptr ? tid : static_cast<int>(a.idxs[tid]);
        if (atom < a.N) {
            const double cb = MODE == 0 ? a.cbs[atom] : 0.5 * a.cbs[atom];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const int q = atom * 3 + d;
                const double force = fixed_to_real<double>(a.du_dx[q]);
                const double v_new = fma(cb, force, a.v[q]);
                a.v[q] = v_new;
                if (MODE != 2) {
                    a.x[q] = fma(a.dt, v_new, a.x[q]);
                }
                a.du_dx[q] = 0;
            }
        } else if (a.idxs != nullptr) {
            a.du_dx[tid * 3 + 0] = 0;
            a.du_dx[tid * 3 + 1] = 0;
            a.du_dx[tid * 3 + 2] = 0;
        }
    }
}

void launch_velocity_verlet(const VerletArgs &args, int mode, cudaStream_t stream) {
    if (args.N <= 0) {
        return;
    }
    const int blocks = ceil_div(args.N, INT_THREADS);
    if (mode == 0) {
        TMB_LAUNCH(k_velocity_verlet<0>, blocks, INT_THREADS, 0, stream, args);
    } else if (mode == 1) {
        TMB_LAUNCH(k_velocity_verlet<1>, blocks, INT_THREADS, 0, stream, args);
    } else {
        TMB_LAUNCH(k_velocity_verlet<2>, blocks, INT_THREADS, 0, stream, args);
    }
}

void launch_baoab(const BaoabArgs &args, cudaStream_t stream) {
    if (args.N <= 0) {
        return;
    }
    TMB_LAUNCH(k_baoab, ceil_div(args.N, INT_THREADS), INT_THREADS, 0, stream, args);
}

template <typename Real> void launch_baoab_prepare(const BaoabArgs &args, const FusedPrepareArgs<Real> &f, cudaStream_t stream) {
    const int kpad = (f.K + WARP - 1) / WARP * WARP;
    const int threads = kpad + f.n_others;
    TMB_LAUNCH(k_baoab_prepare<Real>, ceil_div(threads, INT_THREADS), INT_THREADS, 0, stream, args, f);
}
template void launch_baoab_prepare<float>(const BaoabArgs &, const FusedPrepareArgs<float> &, cudaStream_t);
template void launch_baoab_prepare<double>(const BaoabArgs &, const FusedPrepareArgs<double> &, cudaStream_t);

// N x 3 standard normals with the same generator (used by tests to inspect the noise distribution)
__global__ void k_fill_normal(float *__restrict__ out, const int n_atoms, const unsigned long long seed, const unsigned long long step) {
    const int atom = blockIdx.x * blockDim.x + threadIdx.x;
    if (atom >= n_atoms) {
        return;
    }
    float n0, n1, n2;
    normal3(seed, step, static_cast<unsigned int>(atom), n0, n1, n2);
    out[atom * 3 + 0] = n0;
    out[atom * 3 + 1] = n1;
    out[atom * 3 + 2] = n2;
}

void launch_fill_normal(float *out, int n, unsigned long long seed, unsigned long long step, cudaStream_t stream) {
    if (n <= 0) {
        return;
    }
    TMB_LAUNCH(k_fill_normal, ceil_div(n, INT_THREADS), INT_THREADS, 0, stream, out, n, seed, step);
}

} // namespace tmb
