// Biased-deletion water exchange and per-molecule nonbonded energies (see exchange.hpp for what follows the reference
// and what does not).
//
// Compiled WITHOUT --fmad=false (build.py), like the reference build: logf / expf are inlined from libdevice and must
// contract the same way.  Every other floating-point operation that feeds a decision is an explicit round-to-nearest
// intrinsic in the order ptxas emits for the reference's kernels (read off oracle/_ref/obj/{all_atom_energies,
// k_rotations,nonbonded_mol_energy}.o built for sm_100a):
//   * d^2 = fma(dw, dw, fma(dz, dz, fma(dx, dx, dy * dy))), minimum image fma(-b, rint(d / b), d);
//   * the energy-only pair term of compute_electrostatics / compute_lj <Real, true> (k_nonbonded_common.cuh:184-246);
//   * the two Hamilton products of rotate_coordinates_by_quaternion (k_rotations.cu:9-48), whose contraction pattern
//     differs between the first and the second product.
#include "exchange.hpp"
#include "fixed_point.cuh"

#include <cooperative_groups.h>
#include <curand.h>
#include <curand_kernel.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace cg = cooperative_groups;

namespace tmb {

static const double BOLTZ = 0.008314462618; // reference constants.hpp:5

#define TMB_CURAND(expr)                                                                                               \
    do {                                                                                                               \
        curandStatus_t _st = (expr);                                                                                   \
        if (_st != CURAND_STATUS_SUCCESS) {                                                                            \
            throw std::runtime_error(std::string("cuRAND error ") + std::to_string(static_cast<int>(_st)) + " at " +   \
                                     __FILE__ + ":" + std::to_string(__LINE__));                                       \
        }                                                                                                              \
    } while (0)

static curandStatus_t gen_uniform(curandGenerator_t g, float *out, size_t n) { return curandGenerateUniform(g, out, n); }
static curandStatus_t gen_uniform(curandGenerator_t g, double *out, size_t n) { return curandGenerateUniformDouble(g, out, n); }
static curandStatus_t gen_normal(curandGenerator_t g, float *out, size_t n) { return curandGenerateNormal(g, out, n, 0.0f, 1.0f); }
static curandStatus_t gen_normal(curandGenerator_t g, double *out, size_t n) { return curandGenerateNormalDouble(g, out, n, 0.0, 1.0); }

// ---------------------------------------------------------------------------------------------------------------
// molecule bookkeeping (reference mol_utils.cpp)
void verify_mols_contiguous(const std::vector<std::vector<int>> &group_idxs) {
    int last_end = group_idxs[0][0] - 1;
    for (size_t i = 0; i < group_idxs.size(); i++) {
        std::vector<int> atoms = group_idxs[i];
        if (atoms[0] != last_end + 1) {
            throw std::runtime_error("Molecules are not contiguous: mol " + std::to_string(i));
        }
        std::sort(atoms.begin(), atoms.end());
        for (size_t j = 1; j < atoms.size(); j++) {
            if (atoms[j - 1] + 1 != atoms[j]) {
                throw std::runtime_error("Molecule " + std::to_string(i) + "is not sequential in atom indices");
            }
        }
        last_end = atoms.back();
    }
}

MolLayout flatten_mols(const std::vector<std::vector<int>> &group_idxs) {
    MolLayout l;
    for (size_t i = 0; i < group_idxs.size(); i++) {
        std::vector<int> atoms = group_idxs[i];
        std::sort(atoms.begin(), atoms.end());
        l.mol_offsets.push_back(static_cast<int>(l.atom_idxs.size()));
        for (int a : atoms) {
            l.atom_idxs.push_back(a);
            l.mol_idxs.push_back(static_cast<int>(i));
        }
    }
    l.mol_offsets.push_back(static_cast<int>(l.atom_idxs.size()));
    return l;
}

// ---------------------------------------------------------------------------------------------------------------
// arithmetic helpers: explicit rounding, never contracted
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmad_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fmad_(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double div_(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float rsqrt_(float a) { return rsqrtf(a); }
__device__ __forceinline__ double rsqrt_(double a) { return rsqrt(a); }
__device__ __forceinline__ float rint_(float a) { return rintf(a); }
__device__ __forceinline__ double rint_(double a) { return rint(a); }
__device__ __forceinline__ float floor_(float a) { return floorf(a); }
__device__ __forceinline__ double floor_(double a) { return floor(a); }
__device__ __forceinline__ float log_(float a) { return logf(a); }
__device__ __forceinline__ double log_(double a) { return log(a); }
__device__ __forceinline__ float exp_(float a) { return expf(a); }
__device__ __forceinline__ double exp_(double a) { return exp(a); }
__device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
template <typename Real> __device__ __forceinline__ Real inf_() { return static_cast<Real>(INFINITY); }

template <typename Real> struct Box3 {
    Real x, y, z, ix, iy, iz;
};
template <typename Real> __device__ __forceinline__ Box3<Real> load_box3(const double *__restrict__ box) {
    Box3<Real> b;
    b.x = static_cast<Real>(box[0]);
    b.y = static_cast<Real>(box[4]);
    b.z = static_cast<Real>(box[8]);
    b.ix = div_(static_cast<Real>(1), b.x);
    b.iy = div_(static_cast<Real>(1), b.y);
    b.iz = div_(static_cast<Real>(1), b.z);
    return b;
}

// ---- energy-only pair term --------------------------------------------------------------------------------------
// switching function cos^3(pi/2 (d / 1.2)^8), reference k_nonbonded_common.cuh:16-35 (f64), :96-130 (f32)
__device__ __forceinline__ float switch_fn_(float d) {
    constexpr float cutoff = 1.2f;
    if (d >= cutoff) {
        return 0.0f;
    }
    constexpr float pi = static_cast<float>(3.141592653589793115997963468544185161);
    constexpr float inv_c = 1.0f / cutoff;
    constexpr float k2 = inv_c * inv_c;
    constexpr float k4 = k2 * k2;
    constexpr float k8 = k4 * k4;
    constexpr float half_pi = 0.5f * pi;
    const float d2 = mul_(d, d);
    const float d4 = mul_(d2, d2);
    const float d8 = mul_(d4, d4);
    const float c = __cosf(mul_(half_pi, mul_(d8, k8)));
    return mul_(c, mul_(c, c));
}
__device__ __forceinline__ double switch_fn_(double d) {
    constexpr double cutoff = 1.2;
    if (d >= cutoff) {
        return 0.0;
    }
    constexpr double inv_c = 1 / cutoff;
    const double k = mul_(d, inv_c);
    const double k2 = mul_(k, k);
    const double k4 = mul_(k2, k2);
    const double k8 = mul_(k4, k4);
    const double c = cos(mul_(0.5 * 3.141592653589793115997963468544185161, k8));
    return mul_(mul_(c, c), c);
}
// erfc: A&S 7.1.26 with __expf / __frcp_rn in f32 (reference k_nonbonded_common.cuh:132-160), erfc() in f64
__device__ __forceinline__ float erfc_(float x) {
    const float e = __expf(-mul_(x, x));
    const float t = __frcp_rn(fmad_(0.3275911f, x, 1.0f));
    float p = fmad_(1.061405429f, t, -1.453152027f);
    p = fmad_(p, t, 1.421413741f);
    p = fmad_(p, t, -0.284496736f);
    p = fmad_(p, t, 0.254829592f);
    return mul_(e, mul_(p, t));
}
__device__ __forceinline__ double erfc_(double x) { return erfc(x); }

// u_ij for d2 already inside the cutoff (compute_electrostatics<Real, true> + compute_lj<Real, true> with scales of 1)
template <typename Real>
__device__ __forceinline__ Real pair_u(Real qi, Real qj, Real si, Real sj, Real ei, Real ej, Real d2, Real beta) {
    const Real inv_d = rsqrt_(d2);
    const Real d = mul_(d2, inv_d);
    const Real damping = mul_(erfc_(mul_(beta, d)), switch_fn_(d));
    Real u = mul_(mul_(mul_(qi, qj), inv_d), damping);
    if (ei != static_cast<Real>(0) && ej != static_cast<Real>(0)) {
        const Real eps4 = mul_(static_cast<Real>(4), mul_(ei, ej));
        const Real s1 = mul_(add_(si, sj), inv_d);
        const Real s2 = mul_(s1, s1);
        const Real s4 = mul_(s2, s2);
        const Real s6 = mul_(s4, s2);
        u = fmad_(s6, mul_(eps4, sub_(s6, static_cast<Real>(1))), u);
    }
    return u;
}

// fixed-point energy of atom (xi, pi) with atom (xj, pj); 0 outside the cutoff (reference k_nonbonded.cuh:520-551, 667-697)
template <typename Real>
__device__ __forceinline__ Real
pair_u_or_zero(const Vec4<Real> &xi, const Vec4<Real> &pi, const Vec4<Real> &xj, const Vec4<Real> &pj, const Box3<Real> &b, Real cutoff2, Real beta) {
    Real dx = sub_(xi.x, xj.x);
    Real dy = sub_(xi.y, xj.y);
    Real dz = sub_(xi.z, xj.z);
    const Real dw = sub_(xi.w, xj.w);
    dx = fmad_(-b.x, rint_(mul_(dx, b.ix)), dx);
    dy = fmad_(-b.y, rint_(mul_(dy, b.iy)), dy);
    dz = fmad_(-b.z, rint_(mul_(dz, b.iz)), dz);
    const Real d2 = fmad_(dw, dw, fmad_(dz, dz, fmad_(dx, dx, mul_(dy, dy))));
    if (!(d2 < cutoff2)) {
        return static_cast<Real>(0);
    }
    return pair_u(pi.x, pj.x, pi.y, pj.y, pi.z, pj.z, d2, beta);
}
template <typename Real>
__device__ __forceinline__ i128
pair_e(const Vec4<Real> &xi, const Vec4<Real> &pi, const Vec4<Real> &xj, const Vec4<Real> &pj, const Box3<Real> &b, Real cutoff2, Real beta) {
    Real dx = sub_(xi.x, xj.x);
    Real dy = sub_(xi.y, xj.y);
    Real dz = sub_(xi.z, xj.z);
    const Real dw = sub_(xi.w, xj.w);
    dx = fmad_(-b.x, rint_(mul_(dx, b.ix)), dx);
    dy = fmad_(-b.y, rint_(mul_(dy, b.iy)), dy);
    dz = fmad_(-b.z, rint_(mul_(dz, b.iz)), dz);
    const Real d2 = fmad_(dw, dw, fmad_(dz, dz, fmad_(dx, dx, mul_(dy, dy))));
    if (!(d2 < cutoff2)) {
        return 0;
    }
    return energy_to_fixed<Real>(pair_u(pi.x, pj.x, pi.y, pj.y, pi.z, pj.z, d2, beta));
}

// squared minimum-image distance in 3-D, for skip decisions only (plain arithmetic: compared against a padded bound)
template <typename Real> __device__ __forceinline__ Real anchor_d2(const Vec4<Real> &p, const Vec4<Real> &q, const Box3<Real> &b) {
    Real dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
    dx -= b.x * rint_(dx * b.ix);
    dy -= b.y * rint_(dy * b.iy);
    dz -= b.z * rint_(dz * b.iz);
    return dx * dx + dy * dy + dz * dz;
}

// 128-bit accumulate through two 64-bit atomics; the final value does not depend on the order of the adds
__device__ __forceinline__ void atomic_add_i128(i128 *addr, i128 v) {
    u64 *p = reinterpret_cast<u64 *>(addr);
    const u64 lo = static_cast<u64>(v);
    u64 hi = static_cast<u64>(static_cast<unsigned __int128>(v) >> 64);
    const u64 old = atomicAdd(p, lo);
    if (old + lo < old) {
        hi += 1;
    }
    if (hi != 0) {
        atomicAdd(p + 1, hi);
    }
}

__device__ __forceinline__ bool fixed_overflow(i128 v) { return v >= static_cast<i128>(LLONG_MAX) || v <= static_cast<i128>(LLONG_MIN); }

// beta * E, +inf for an energy outside the int64 range (reference k_exchange.cu:32-42)
template <typename Real> __device__ __forceinline__ Real log_weight(i128 e, Real beta) {
    return fixed_overflow(e) ? inf_<Real>() : mul_(beta, fixed_to_real<Real>(static_cast<u64>(e)));
}
template <typename Real> __device__ __forceinline__ Real nan_to_inf(Real v) { return isnan(v) ? inf_<Real>() : v; }

// ---------------------------------------------------------------------------------------------------------------
// staging: coordinates + w, and (q, sig, eps) per atom in Real (the reference casts per use: same values)
template <typename Real>
__global__ void k_stage_atoms(int N, const double *__restrict__ coords, const double *__restrict__ params, Vec4<Real> *__restrict__ xr, Vec4<Real> *__restrict__ pr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) {
        return;
    }
    Vec4<Real> x, p;
    x.x = static_cast<Real>(coords[i * 3 + 0]);
    x.y = static_cast<Real>(coords[i * 3 + 1]);
    x.z = static_cast<Real>(coords[i * 3 + 2]);
    x.w = static_cast<Real>(params[i * 4 + P_W]);
    p.x = static_cast<Real>(params[i * 4 + P_CHARGE]);
    p.y = static_cast<Real>(params[i * 4 + P_SIG]);
    p.z = static_cast<Real>(params[i * 4 + P_EPS]);
    p.w = static_cast<Real>(0);
    xr[i] = x;
    pr[i] = p;
}

// ---------------------------------------------------------------------------------------------------------------
// energies of all target molecules (reference k_compute_nonbonded_target_atom_energies + k_accumulate_..., brute force)
// One thread per target atom, a slice of all atoms per blockIdx.y staged through shared memory, 128-bit atomics per
// molecule.  out must be zeroed.
constexpr int ME_THREADS = 128;
template <typename Real>
__global__ void __launch_bounds__(ME_THREADS) k_mol_energies(
    int N, int num_targets, const int4 *__restrict__ targets, const Vec4<Real> *__restrict__ xr, const Vec4<Real> *__restrict__ pr,
    const double *__restrict__ box, Real beta, Real cutoff2, i128 *__restrict__ out) {
    __shared__ Vec4<Real> sx[ME_THREADS], sp[ME_THREADS];
    const Box3<Real> b = load_box3<Real>(box);
    const int t = blockIdx.x * ME_THREADS + threadIdx.x;
    const bool live = t < num_targets;
    int4 tg = make_int4(0, 0, 0, -1);
    Vec4<Real> xi = {}, pi = {};
    if (live) {
        tg = targets[t];
        xi = xr[tg.x];
        pi = pr[tg.x];
    }
    const int per = (N + static_cast<int>(gridDim.y) - 1) / static_cast<int>(gridDim.y);
    const int j0 = blockIdx.y * per;
    const int j1 = min(N, j0 + per);
    i128 acc = 0;
    for (int base = j0; base < j1; base += ME_THREADS) {
        const int j = base + threadIdx.x;
        __syncthreads();
        if (j < j1) {
            sx[threadIdx.x] = xr[j];
            sp[threadIdx.x] = pr[j];
        }
        __syncthreads();
        const int n = min(ME_THREADS, j1 - base);
        if (live) {
            for (int k = 0; k < n; k++) {
                const int jj = base + k;
                if (jj >= tg.z && jj <= tg.w) {
                    continue; // same molecule
                }
                acc += pair_e(xi, pi, sx[k], sp[k], b, cutoff2, beta);
            }
        }
    }
    if (live && acc != 0) {
        atomic_add_i128(out + tg.y, acc);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// all-molecule energies in Hilbert order: the brute-force row x column sweep above, but rows and columns are visited
// along the Hilbert curve of the molecules, 128 atoms per chunk with a bounding box each, and a row chunk skips every
// column chunk whose box is further than the cutoff away (exact zeros).  Same integers as k_mol_energies.
template <typename Real>
__global__ void __launch_bounds__(ME_THREADS) k_me_chunks(
    int N, int M, int S, int first, const unsigned int *__restrict__ mol_order, const Vec4<Real> *__restrict__ xr, const Vec4<Real> *__restrict__ pr,
    const double *__restrict__ box, Vec4<Real> *__restrict__ xs, Vec4<Real> *__restrict__ ps, int *__restrict__ col_atom,
    Vec4<Real> *__restrict__ chunk_ctr, Vec4<Real> *__restrict__ chunk_ext) {
    __shared__ Real lo[3][ME_THREADS], hi[3][ME_THREADS];
    const Box3<Real> b = load_box3<Real>(box);
    const int c = blockIdx.x * ME_THREADS + threadIdx.x;
    Real w[3] = {0, 0, 0};
    const bool live = c < N;
    if (live) {
        int atom;
        if (c < M * S) {
            atom = static_cast<int>(mol_order[c / S]) + c % S;
        } else {
            const int k = c - M * S;
            atom = k < first ? k : k + M * S;
        }
        const Vec4<Real> x = xr[atom];
        xs[c] = x;
        ps[c] = pr[atom];
        col_atom[c] = atom;
        // home-box image, for the bounding box only
        w[0] = x.x - b.x * floor_(x.x * b.ix);
        w[1] = x.y - b.y * floor_(x.y * b.iy);
        w[2] = x.z - b.z * floor_(x.z * b.iz);
    }
    for (int d = 0; d < 3; d++) {
        // a NaN coordinate poisons nothing here: the comparisons below drop it, and its pairs fail every cutoff test
        lo[d][threadIdx.x] = live ? w[d] : inf_<Real>();
        hi[d][threadIdx.x] = live ? w[d] : -inf_<Real>();
    }
    __syncthreads();
    for (int s2 = ME_THREADS / 2; s2 > 0; s2 >>= 1) {
        if (threadIdx.x < s2) {
            for (int d = 0; d < 3; d++) {
                if (lo[d][threadIdx.x + s2] < lo[d][threadIdx.x]) {
                    lo[d][threadIdx.x] = lo[d][threadIdx.x + s2];
                }
                if (hi[d][threadIdx.x + s2] > hi[d][threadIdx.x]) {
                    hi[d][threadIdx.x] = hi[d][threadIdx.x + s2];
                }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Vec4<Real> ctr, ext;
        const Real half = static_cast<Real>(0.5), pad = static_cast<Real>(1e-4);
        ctr.x = half * (lo[0][0] + hi[0][0]);
        ctr.y = half * (lo[1][0] + hi[1][0]);
        ctr.z = half * (lo[2][0] + hi[2][0]);
        ctr.w = 0;
        ext.x = half * (hi[0][0] - lo[0][0]) + pad;
        ext.y = half * (hi[1][0] - lo[1][0]) + pad;
        ext.z = half * (hi[2][0] - lo[2][0]) + pad;
        ext.w = 0;
        chunk_ctr[blockIdx.x] = ctr;
        chunk_ext[blockIdx.x] = ext;
    }
}

template <typename Real>
__global__ void __launch_bounds__(ME_THREADS) k_mol_energies_sorted(
    int N, int M, int S, int first, const Vec4<Real> *__restrict__ xs, const Vec4<Real> *__restrict__ ps, const int *__restrict__ col_atom,
    const Vec4<Real> *__restrict__ chunk_ctr, const Vec4<Real> *__restrict__ chunk_ext, const double *__restrict__ box, Real beta, Real cutoff2,
    i128 *__restrict__ out) {
    __shared__ Vec4<Real> sx[ME_THREADS], sp[ME_THREADS];
    __shared__ int sa[ME_THREADS];
    const Box3<Real> b = load_box3<Real>(box);
    const int r = blockIdx.x * ME_THREADS + threadIdx.x; // row = one of the first M * S sorted atoms
    const bool live = r < M * S;
    Vec4<Real> xi = {}, pi = {};
    int lo_atom = 0, hi_atom = -1, mol = 0;
    if (live) {
        xi = xs[r];
        pi = ps[r];
        lo_atom = col_atom[r] - r % S;
        hi_atom = lo_atom + S - 1;
        mol = (lo_atom - first) / S;
    }
    const Vec4<Real> rc = chunk_ctr[blockIdx.x], re = chunk_ext[blockIdx.x];
    const int n_chunks = (N + ME_THREADS - 1) / ME_THREADS;
    i128 acc = 0;
    for (int c = blockIdx.y; c < n_chunks; c += gridDim.y) {
        const Vec4<Real> cc = chunk_ctr[c], ce = chunk_ext[c];
        Real dx = rc.x - cc.x, dy = rc.y - cc.y, dz = rc.z - cc.z;
        dx -= b.x * rint_(dx * b.ix);
        dy -= b.y * rint_(dy * b.iy);
        dz -= b.z * rint_(dz * b.iz);
        const Real gx = fmax(fabs(dx) - re.x - ce.x, static_cast<Real>(0));
        const Real gy = fmax(fabs(dy) - re.y - ce.y, static_cast<Real>(0));
        const Real gz = fmax(fabs(dz) - re.z - ce.z, static_cast<Real>(0));
        if (gx * gx + gy * gy + gz * gz >= cutoff2) {
            continue; // block-uniform: no atom of the column chunk is within the cutoff of any row atom
        }
        const int base = c * ME_THREADS;
        const int n = min(ME_THREADS, N - base);
        __syncthreads();
        if (static_cast<int>(threadIdx.x) < n) {
            sx[threadIdx.x] = xs[base + threadIdx.x];
            sp[threadIdx.x] = ps[base + threadIdx.x];
            sa[threadIdx.x] = col_atom[base + threadIdx.x];
        }
        __syncthreads();
        if (live) {
            for (int k = 0; k < n; k++) {
                const int jj = sa[k];
                if (jj >= lo_atom && jj <= hi_atom) {
                    continue; // same molecule
                }
                acc += pair_e(xi, pi, sx[k], sp[k], b, cutoff2, beta);
            }
        }
    }
    if (live && acc != 0) {
        atomic_add_i128(out + mol, acc);
    }
}

// reference k_atom_by_atom_energies (k_nonbonded.cuh:604-700): [T, N] pair energies in Real
template <typename Real>
__global__ void k_atom_by_atom(
    int N, int T, const int *__restrict__ target_atoms, const Vec4<Real> *__restrict__ xr, const Vec4<Real> *__restrict__ pr,
    const double *__restrict__ box, Real beta, Real cutoff2, Real *__restrict__ out) {
    const Box3<Real> b = load_box3<Real>(box);
    const int row = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= T || j >= N) {
        return;
    }
    const int i = target_atoms[row];
    out[static_cast<size_t>(row) * N + j] = pair_u_or_zero(xr[i], pr[i], xr[j], pr[j], b, cutoff2, beta);
}

// ---------------------------------------------------------------------------------------------------------------
// block-level reductions (EX_THREADS threads).  No __restrict__ on their inputs: the persistent kernel rewrites those
// buffers between grid barriers and the loads must stay ordinary (coherent) loads.  The order of the floating-point sum is fixed by the thread count and
// the element count only, so a log-sum-exp is bitwise reproducible whichever batch slot computes it.
constexpr int EX_THREADS = 256;

template <typename Real> struct LseScratch {
    Real v[EX_THREADS];
    int i[EX_THREADS];
};

template <typename Real> __device__ Real block_max(const Real *vals, int n, LseScratch<Real> &s) {
    Real m = -inf_<Real>();
#pragma unroll 8
    for (int k = threadIdx.x; k < n; k += EX_THREADS) {
        const Real v = vals[k];
        m = (v > m) ? v : m;
    }
    s.v[threadIdx.x] = m;
    __syncthreads();
    for (int w = EX_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            const Real o = s.v[threadIdx.x + w];
            if (o > s.v[threadIdx.x]) {
                s.v[threadIdx.x] = o;
            }
        }
        __syncthreads();
    }
    const Real r = s.v[0];
    __syncthreads();
    return r;
}

// (max, sum exp(x - max))  (reference SegmentedSumExp::sum_device)
template <typename Real> __device__ void block_sumexp(const Real *vals, int n, LseScratch<Real> &s, Real &out_max, Real &out_sum) {
    const Real m = block_max(vals, n, s);
    Real acc = static_cast<Real>(0);
#pragma unroll 8
    for (int k = threadIdx.x; k < n; k += EX_THREADS) { // loads and exps overlap; the adds stay in order
        acc = add_(acc, exp_(sub_(vals[k], m)));
    }
    s.v[threadIdx.x] = acc;
    __syncthreads();
    for (int w = EX_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            s.v[threadIdx.x] = add_(s.v[threadIdx.x], s.v[threadIdx.x + w]);
        }
        __syncthreads();
    }
    out_max = m;
    out_sum = s.v[0];
    __syncthreads();
}

// arg max of log_weights[k] - log(-log(noise[k])) with the lowest index on ties (Gumbel-max trick, reference
// k_sampling.cu:10-75 + cub ArgMax)
template <typename Real> __device__ int block_gumbel_argmax(const Real *logw, const Real *noise, int n, LseScratch<Real> &s) {
    Real best = -inf_<Real>();
    int best_i = 0x7fffffff;
#pragma unroll 8
    for (int k = threadIdx.x; k < n; k += EX_THREADS) {
        const Real g = -log_(-log_(noise[k]));
        const Real v = add_(logw[k], g);
        if (v > best || best_i == 0x7fffffff) {
            best = v;
            best_i = k;
        }
    }
    s.v[threadIdx.x] = best;
    s.i[threadIdx.x] = best_i;
    __syncthreads();
    for (int w = EX_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            const Real ov = s.v[threadIdx.x + w];
            const int oi = s.i[threadIdx.x + w];
            const Real mv = s.v[threadIdx.x];
            const int mi = s.i[threadIdx.x];
            if (oi != 0x7fffffff && (mi == 0x7fffffff || ov > mv || (ov == mv && oi < mi))) {
                s.v[threadIdx.x] = ov;
                s.i[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    const int r = s.i[0];
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// rotation of (x, y, z) by the normalised quaternion q: q (0, v) q*   (reference k_rotations.cu:9-48)
template <typename Real> __device__ __forceinline__ void rotate_by_quaternion(const Real *q_raw, Real &cx, Real &cy, Real &cz) {
    const Real q0 = q_raw[0], q1 = q_raw[1], q2 = q_raw[2], q3 = q_raw[3];
    const Real n2 = fmad_(q3, q3, fmad_(q2, q2, fmad_(q0, q0, mul_(q1, q1))));
    const Real inv = rsqrt_(n2);
    const Real w = mul_(q0, inv), x = mul_(q1, inv), y = mul_(q2, inv), z = mul_(q3, inv);
    const Real zero = static_cast<Real>(0);
    // q (0, c)
    const Real i0 = fmad_(cz, -z, fmad_(cy, -y, fmad_(zero, w, -mul_(cx, x))));
    const Real i1 = fmad_(cy, -z, fmad_(cz, y, fmad_(cx, w, mul_(zero, x))));
    const Real i2 = fmad_(cx, z, fmad_(zero, y, fmad_(cy, w, -mul_(cz, x))));
    const Real i3 = fmad_(zero, z, fmad_(cx, -y, fmad_(cz, w, mul_(cy, x))));
    // (i) q*
    cx = fmad_(y, i3, fmad_(-z, i2, fmad_(w, i1, -mul_(x, i0))));
    cy = fmad_(-x, i3, fmad_(w, i2, fmad_(z, i1, -mul_(y, i0))));
    cz = fmad_(w, i3, fmad_(x, i2, fmad_(-y, i1, -mul_(z, i0))));
}

// rotate a molecule about its centroid and put the centroid at the translation (imaged into the home box; scaled by the
// box first when SCALE)   (reference k_rotate_and_translate_mols, k_rotations.cu:112-193)
template <typename Real>
__device__ void rotate_and_translate(
    int num_atoms, const Vec4<Real> *mol_x, const Real *quat, const Real *trans, const Box3<Real> &b, bool scale, Vec4<Real> *out) {
    Real tx = trans[0], ty = trans[1], tz = trans[2];
    if (scale) {
        tx = mul_(tx, b.x);
        ty = mul_(ty, b.y);
        tz = mul_(tz, b.z);
    }
    tx = fmad_(b.x, -floor_(mul_(tx, b.ix)), tx);
    ty = fmad_(b.y, -floor_(mul_(ty, b.iy)), ty);
    tz = fmad_(b.z, -floor_(mul_(tz, b.iz)), tz);
    u64 ax = 0, ay = 0, az = 0;
    for (int i = 0; i < num_atoms; i++) {
        ax += to_fixed<FIXED_EXPONENT>(mol_x[i].x);
        ay += to_fixed<FIXED_EXPONENT>(mol_x[i].y);
        az += to_fixed<FIXED_EXPONENT>(mol_x[i].z);
    }
    const Real n = static_cast<Real>(num_atoms);
    const Real gx = div_(fixed_to_real<Real>(ax), n);
    const Real gy = div_(fixed_to_real<Real>(ay), n);
    const Real gz = div_(fixed_to_real<Real>(az), n);
    for (int i = 0; i < num_atoms; i++) {
        Real cx = sub_(mol_x[i].x, gx);
        Real cy = sub_(mol_x[i].y, gy);
        Real cz = sub_(mol_x[i].z, gz);
        rotate_by_quaternion(quat, cx, cy, cz);
        Vec4<Real> o;
        o.x = add_(cx, tx);
        o.y = add_(cy, ty);
        o.z = add_(cz, tz);
        o.w = mol_x[i].w;
        out[i] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the biased-deletion move, phase by phase
enum { ST_OFFSET = 0, ST_SELECTED = 1, ST_ITER = 2, ST_INNER_USED = 3, ST_WORDS = 4 };

// targeted insertion (TIBDExchangeMove): molecules are split into those inside a sphere around the ligand centroid and
// the rest; a proposal deletes from one region and inserts into the other.  enabled == 0 for plain biased deletion.
template <typename Real> struct TIDevice {
    int enabled;
    Real inner_volume;
    const Real *box_volume;   // [1]
    const Real *uniform;      // [P] which region a proposal targets
    int *inner_flags;         // [M] 1 inside the sphere
    int *partition;           // [M] inner molecules ascending, then outer molecules descending (cub::DevicePartition::Flagged order)
    int *inner_count;         // [1]
    int *targeting;           // [B] 1: the proposal inserts into the sphere
    Real *src_logw, *dest_logw;       // [B, M] weights of the region deleted from / inserted into
    Real *lse_src_max, *lse_src_sum;  // [B]
};

template <typename Real> struct BDDevice {
    int N, M, S, B, P;   // atoms, target molecules, atoms per molecule, batch size, proposals per move
    int first;           // first atom of molecule 0
    int sample, scale;   // draw the molecule with the Gumbel-max trick / scale translations by the box
    Real nb_beta, beta, cutoff2;
    double *coords;      // [N, 3] the caller's coordinates (updated when a proposal is accepted)
    const double *box;
    Vec4<Real> *xr;      // [N] staged copy of coords (+ w), kept in step with coords
    const Vec4<Real> *pr;
    Vec4<Real> *prop;    // [B, S] proposed positions
    Vec4<Real> *prop_old; // [B, S] positions the proposal moves away from (a copy: xr changes when a proposal is accepted)
    i128 *before_E, *total;           // [M], [B]: molecule energies of the current state; energy of the moved molecule per proposal
    Real *logw_before, *logw_after;   // [M], [B, M]
    Real *lse_before;                 // {max, sum}
    Real *lse_after_max, *lse_after_sum; // [B]
    int *samples;                     // [B]
    const unsigned int *mol_order;    // [M] first atoms of the molecules in Hilbert order (work order of the pair phase)
    const Real *r_bound;              // [1] no atom of a target molecule is further than this from the molecule's first atom
    int *state;                       // ST_*
    u64 *num_accepted;
    const Real *quat, *trans, *sample_noise, *mh; // [P, 4], [P, 3] ([P, 2, 3] targeted), [P, M], [P]
    TIDevice<Real> ti;
};

template <typename Real> struct BDShared {
    LseScratch<Real> red;
    int sel;
};

// inner molecules in ascending order at the front, outer molecules in descending order behind them: the order
// cub::DevicePartition::Flagged produces in the reference (tibd_exchange_move.cu:214-226); the proposals index into it
template <typename Real> __device__ void ti_partition(const TIDevice<Real> &t, int M, int *scratch /* [EX_THREADS] */) {
    const int per = (M + EX_THREADS - 1) / EX_THREADS;
    const int lo = min(M, static_cast<int>(threadIdx.x) * per), hi = min(M, lo + per);
    int mine = 0;
    for (int m = lo; m < hi; m++) {
        mine += t.inner_flags[m];
    }
    scratch[threadIdx.x] = mine;
    __syncthreads();
    int inner_before = 0;
    for (int k = 0; k < static_cast<int>(threadIdx.x); k++) {
        inner_before += scratch[k];
    }
    for (int m = lo; m < hi; m++) {
        if (t.inner_flags[m]) {
            t.partition[inner_before++] = m;
        } else {
            t.partition[M - 1 - (m - inner_before)] = m; // m - inner_before outer molecules precede m
        }
    }
    if (threadIdx.x == EX_THREADS - 1) {
        t.inner_count[0] = inner_before;
    }
    __syncthreads();
}

// Most molecules are out of reach of a proposal and keep their weight: the proposal's row starts as a copy of the current
// weights (coalesced) and the pair phase rewrites only the molecules near the old or the new position.
template <typename Real> __device__ __forceinline__ void bd_seed_after_weights(const BDDevice<Real> &a, int b) {
    Real *row = a.logw_after + static_cast<size_t>(b) * a.M;
    // eight loads in flight per thread (the two arrays never overlap, which the compiler cannot know)
    int m = threadIdx.x;
    for (; m + 7 * EX_THREADS < a.M; m += 8 * EX_THREADS) {
        Real v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            v[u] = a.logw_before[m + u * EX_THREADS];
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            row[m + u * EX_THREADS] = v[u];
        }
    }
    for (; m < a.M; m += EX_THREADS) {
        row[m] = a.logw_before[m];
    }
}

// branch-free variant of pair_e for the unrolled molecule-molecule block below: the term is always evaluated and
// selected afterwards, so that the SC * SC independent evaluations of a block can overlap (the pair phase is latency
// bound: ncu r1s3, 21 % of issue slots used at 16 warps per SM).  Same arithmetic, same bits.
template <typename Real>
__device__ __forceinline__ i128
pair_e_select(const Vec4<Real> &xi, const Vec4<Real> &pi, const Vec4<Real> &xj, const Vec4<Real> &pj, const Box3<Real> &b, Real cutoff2, Real beta) {
    Real dx = sub_(xi.x, xj.x);
    Real dy = sub_(xi.y, xj.y);
    Real dz = sub_(xi.z, xj.z);
    const Real dw = sub_(xi.w, xj.w);
    dx = fmad_(-b.x, rint_(mul_(dx, b.ix)), dx);
    dy = fmad_(-b.y, rint_(mul_(dy, b.iy)), dy);
    dz = fmad_(-b.z, rint_(mul_(dz, b.iz)), dz);
    const Real d2 = fmad_(dw, dw, fmad_(dz, dz, fmad_(dx, dx, mul_(dy, dy))));
    const Real inv_d = rsqrt_(d2);
    const Real d = mul_(d2, inv_d);
    const Real damping = mul_(erfc_(mul_(beta, d)), switch_fn_(d));
    const Real u_es = mul_(mul_(mul_(pi.x, pj.x), inv_d), damping);
    const Real eps4 = mul_(static_cast<Real>(4), mul_(pi.z, pj.z));
    const Real s1 = mul_(add_(pi.y, pj.y), inv_d);
    const Real s2 = mul_(s1, s1);
    const Real s4 = mul_(s2, s2);
    const Real s6 = mul_(s4, s2);
    const Real u_lj = fmad_(s6, mul_(eps4, sub_(s6, static_cast<Real>(1))), u_es);
    const bool lj = pi.z != static_cast<Real>(0) && pj.z != static_cast<Real>(0);
    const i128 e = energy_to_fixed<Real>(lj ? u_lj : u_es);
    return (d2 < cutoff2) ? e : static_cast<i128>(0);
}

// energy change of molecule `mol` (first atom j0) when molecule s goes from xold to xnew; e_new_out: its energy with the
// new position alone.  Exact zeros when the first atoms are further apart than reach (see the pair phase).
// SC: atoms per molecule as a compile-time constant (fully unrolled, branch-free block) or 0 for the generic loop.
template <typename Real, int SC>
__device__ __forceinline__ bool bd_mol_delta_impl(
    const BDDevice<Real> &a, const Box3<Real> &bx, const Vec4<Real> *xold, const Vec4<Real> *xnew, const Vec4<Real> *pmol, int j0,
    Real reach2, i128 &delta, i128 &e_new_out) {
    const Vec4<Real> anchor = a.xr[j0];
    const bool near_old = anchor_d2(xold[0], anchor, bx) < reach2;
    const bool near_new = anchor_d2(xnew[0], anchor, bx) < reach2;
    if (!(near_old || near_new)) {
        return false;
    }
    if constexpr (SC > 0) {
        Vec4<Real> xj[SC], pj[SC];
#pragma unroll
        for (int j = 0; j < SC; j++) {
            xj[j] = a.xr[j0 + j];
            pj[j] = a.pr[j0 + j];
        }
        if (near_new) {
            i128 acc = 0;
#pragma unroll
            for (int i = 0; i < SC; i++) {
                const Vec4<Real> xn = xnew[i], pi = pmol[i];
#pragma unroll
                for (int j = 0; j < SC; j++) {
                    acc += pair_e_select(xn, pi, xj[j], pj[j], bx, a.cutoff2, a.nb_beta);
                }
            }
            delta += acc;
            e_new_out += acc;
        }
        if (near_old) {
            i128 acc = 0;
#pragma unroll
            for (int i = 0; i < SC; i++) {
                const Vec4<Real> xo = xold[i], pi = pmol[i];
#pragma unroll
                for (int j = 0; j < SC; j++) {
                    acc += pair_e_select(xo, pi, xj[j], pj[j], bx, a.cutoff2, a.nb_beta);
                }
            }
            delta -= acc;
        }
    } else {
        for (int i = 0; i < a.S; i++) {
            const Vec4<Real> xo = xold[i], xn = xnew[i], pi = pmol[i];
            for (int j = j0; j < j0 + a.S; j++) {
                const Vec4<Real> xj = a.xr[j], pj = a.pr[j];
                if (near_new) {
                    const i128 e_new = pair_e(xn, pi, xj, pj, bx, a.cutoff2, a.nb_beta);
                    delta += e_new;
                    e_new_out += e_new;
                }
                if (near_old) {
                    delta -= pair_e(xo, pi, xj, pj, bx, a.cutoff2, a.nb_beta);
                }
            }
        }
    }
    return true;
}
template <typename Real>
__device__ __forceinline__ bool bd_mol_delta(
    const BDDevice<Real> &a, const Box3<Real> &bx, const Vec4<Real> *xold, const Vec4<Real> *xnew, const Vec4<Real> *pmol, int j0,
    Real reach2, i128 &delta, i128 &e_new_out) {
    if (a.S == 3) { // water
        return bd_mol_delta_impl<Real, 3>(a, bx, xold, xnew, pmol, j0, reach2, delta, e_new_out);
    }
    return bd_mol_delta_impl<Real, 0>(a, bx, xold, xnew, pmol, j0, reach2, delta, e_new_out);
}

// phase S: choose the molecule of every live batch slot and build its proposal
template <typename Real> __device__ void bd_phase_sample(const BDDevice<Real> &a, BDShared<Real> &sh, int off, int nblocks, int block) {
    const Box3<Real> bx = load_box3<Real>(a.box);
    const int live = min(a.B, a.P - off);
    if (block == 0 && threadIdx.x == 0 && a.sample && !a.ti.enabled) {
        // the previous batch's accepted proposal becomes the "before" state (reference k_store_accepted_log_probability)
        const int sel = a.state[ST_SELECTED];
        if (a.state[ST_ITER] > 0 && sel < a.B) {
            a.lse_before[0] = a.lse_after_max[sel];
            a.lse_before[1] = a.lse_after_sum[sel];
        }
    }
    if (a.ti.enabled) {
        // (tibd_exchange_move.cu:228-307: k_decide_targeted_moves, k_separate_weights_for_targeted, the targeted Gumbel
        // set-up, arg max, the source region's log-sum-exp, k_adjust_sample_idxs)
        const int inner = a.ti.inner_count[0], outer = a.M - inner;
        if (block == 0 && threadIdx.x == 0) {
            a.state[ST_INNER_USED] = inner; // phases L and A read this copy: block 0 re-partitions while the others still test
        }
        for (int b = block; b < live; b += nblocks) {
            int flag;
            if (outer == 0) {
                flag = 0;
            } else if (inner == 0) {
                flag = 1;
            } else {
                flag = a.ti.uniform[off + b] < static_cast<Real>(0.5) ? 1 : 0;
            }
            const int count = flag ? outer : inner, offset = flag ? inner : 0;
            Real *src = a.ti.src_logw + static_cast<size_t>(b) * a.M;
            for (int k = threadIdx.x; k < count; k += EX_THREADS) {
                src[k] = a.logw_before[a.ti.partition[k + offset]];
            }
            __syncthreads();
            const int pick = block_gumbel_argmax(src, a.sample_noise + static_cast<size_t>(off + b) * a.M, count, sh.red);
            Real m, sum;
            block_sumexp(src, count, sh.red, m, sum);
            if (threadIdx.x == 0) {
                const int s = a.ti.partition[pick + offset];
                a.samples[b] = s;
                a.total[b] = 0;
                a.ti.targeting[b] = flag;
                a.ti.lse_src_max[b] = m;
                a.ti.lse_src_sum[b] = sum;
                // first translation of the pair lies inside the sphere, the second outside; already absolute positions
                rotate_and_translate(
                    a.S, a.xr + a.first + s * a.S, a.quat + static_cast<size_t>(off + b) * 4,
                    a.trans + static_cast<size_t>(off + b) * 6 + (flag ? 0 : 3), bx, false, a.prop + static_cast<size_t>(b) * a.S);
                for (int i = 0; i < a.S; i++) {
                    a.prop_old[static_cast<size_t>(b) * a.S + i] = a.xr[a.first + s * a.S + i];
                }
            }
            bd_seed_after_weights(a, b);
        }
        return;
    }
    for (int b = block; b < live; b += nblocks) {
        int s;
        if (a.sample) {
            s = block_gumbel_argmax(a.logw_before, a.sample_noise + static_cast<size_t>(off + b) * a.M, a.M, sh.red);
        } else {
            s = a.samples[b];
        }
        if (threadIdx.x == 0) {
            a.samples[b] = s;
            a.total[b] = 0;
            rotate_and_translate(
                a.S, a.xr + a.first + s * a.S, a.quat + static_cast<size_t>(off + b) * 4, a.trans + static_cast<size_t>(off + b) * 3, bx,
                a.scale != 0, a.prop + static_cast<size_t>(b) * a.S);
            for (int i = 0; i < a.S; i++) {
                a.prop_old[static_cast<size_t>(b) * a.S + i] = a.xr[a.first + s * a.S + i];
            }
        }
        bd_seed_after_weights(a, b);
    }
}

// phase E: pair energies of the chosen molecule, at its old and at its proposed position, with every atom
//   other molecule m:  E_after[b, m] = E_before[m] + sum_{i in s, j in m} (e_new(i, j) - e_old(i, j))
//   chosen molecule:   total[b]      = sum_{i in s, j not in s} e_new(i, j)
// (the reference's k_atom_by_atom_energies x 2 + k_adjust_energies x 2 + k_set_sampled_energy_block/_reduce)
template <typename Real> __device__ void bd_phase_energies(const BDDevice<Real> &a, BDShared<Real> &sh, int off, int nblocks, int block) {
    const Box3<Real> bx = load_box3<Real>(a.box);
    const int live = min(a.B, a.P - off);
    const int n_other = a.N - a.M * a.S;
    const int items = a.M + n_other;
    // Work is handed out per warp (32 consecutive items of one proposal) with no block-level step in between: the few
    // warps whose molecules lie near the moved molecule run long, latency-bound pair evaluations and must not hold up
    // the others.  Molecules are visited in Hilbert order, so a warp is either near or far as a whole.
    const int chunks = (items + WARP - 1) / WARP;
    const long long tasks = static_cast<long long>(live) * chunks;
    const int warps_per_block = EX_THREADS / WARP;
    const int lane = threadIdx.x % WARP;
    const Real reach_mol = sqrt(a.cutoff2) + 2 * a.r_bound[0], reach_atom = sqrt(a.cutoff2) + a.r_bound[0];
    for (long long task = static_cast<long long>(block) * warps_per_block + threadIdx.x / WARP; task < tasks;
         task += static_cast<long long>(nblocks) * warps_per_block) {
        const int b = static_cast<int>(task / chunks);
        const int item = static_cast<int>(task % chunks) * WARP + lane;
        const int s = a.samples[b];
        const Vec4<Real> *xold = a.prop_old + static_cast<size_t>(b) * a.S;
        const Vec4<Real> *pmol = a.pr + a.first + s * a.S;
        const Vec4<Real> *xnew = a.prop + static_cast<size_t>(b) * a.S;
        i128 acc_new = 0;
        // Anchor test: a pair of atoms is at least (distance of the molecules' first atoms) - 2 r_bound apart, so a
        // molecule whose anchor is further than cutoff + 2 r_bound from the moved molecule's anchor contributes exact
        // zeros and is skipped (pairs outside the cutoff are exact zeros anyway: the result is bit-identical).
        if (item < a.M) {
            const int j0 = static_cast<int>(a.mol_order[item]);
            const int mol = (j0 - a.first) / a.S;
            if (mol != s) {
                i128 delta = 0;
                if (bd_mol_delta(a, bx, xold, xnew, pmol, j0, reach_mol * reach_mol, delta, acc_new)) {
                    a.logw_after[static_cast<size_t>(b) * a.M + mol] = log_weight<Real>(a.before_E[mol] + delta, a.beta);
                }
            }
        } else if (item < items) {
            const int k = item - a.M;
            const int j = k < a.first ? k : k + a.M * a.S;
            const Vec4<Real> xj = a.xr[j], pj = a.pr[j];
            if (anchor_d2(xnew[0], xj, bx) < reach_atom * reach_atom) {
                for (int i = 0; i < a.S; i++) {
                    acc_new += pair_e(xnew[i], pmol[i], xj, pj, bx, a.cutoff2, a.nb_beta);
                }
            }
        }
        if (acc_new != 0) {
            atomic_add_i128(a.total + b, acc_new); // a few hundred per proposal: the atoms inside the cutoff of the new position
        }
    }
}

// phase L: energy and weight of the chosen molecule, log-sum-exp of the proposal's weights
template <typename Real> __device__ void bd_phase_lse(const BDDevice<Real> &a, BDShared<Real> &sh, int off, int nblocks, int block) {
    const int live = min(a.B, a.P - off);
    for (int b = block; b < live; b += nblocks) {
        if (threadIdx.x == 0) {
            const int s = a.samples[b];
            a.logw_after[static_cast<size_t>(b) * a.M + s] = log_weight<Real>(a.total[b], a.beta);
        }
        __syncthreads();
        Real m, sum;
        if (a.ti.enabled) {
            // weights of the region inserted into, plus the moved molecule's (k_setup_destination_weights_for_targeted)
            const int inner = a.state[ST_INNER_USED];
            const int flag = a.ti.targeting[b];
            const int count = flag ? inner : a.M - inner, offset = flag ? 0 : inner;
            const Real *after = a.logw_after + static_cast<size_t>(b) * a.M;
            Real *dest = a.ti.dest_logw + static_cast<size_t>(b) * a.M;
            for (int k = threadIdx.x; k < count; k += EX_THREADS) {
                dest[k] = after[a.ti.partition[k + offset]];
            }
            if (threadIdx.x == 0) {
                dest[count] = after[a.samples[b]];
            }
            __syncthreads();
            block_sumexp(dest, count + 1, sh.red, m, sum);
        } else {
            block_sumexp(a.logw_after + static_cast<size_t>(b) * a.M, a.M, sh.red, m, sum);
        }
        if (threadIdx.x == 0) {
            a.lse_after_max[b] = m;
            a.lse_after_sum[b] = sum;
        }
    }
}

// k_exchange.cuh:44-73 compute_raw_log_probability_targeted
template <typename Real>
__host__ __device__ inline Real ti_raw_log_probability(
    int targeting_inner, Real inner_volume, Real outer_volume, int inner_count, int M, Real src_max, Real src_sum, Real dest_max, Real dest_sum) {
    const Real log_vol = targeting_inner == 1 ? log(inner_volume) - log(outer_volume) : log(outer_volume) - log(inner_volume);
    const int src = targeting_inner == 1 ? M - inner_count : inner_count;
    const int dest = targeting_inner == 1 ? inner_count : M - inner_count;
    const Real half = static_cast<Real>(log(0.5)), one = static_cast<Real>(0);
    const Real log_fwd = (src > 0 && dest > 0) ? half : one;
    const Real log_rev = (src - 1 > 0 && dest + 1 > 0) ? half : one;
    Real before = src_max + log(src_sum);
    Real after = dest_max + log(dest_sum);
    before = isnan(before) ? static_cast<Real>(INFINITY) : before;
    after = isnan(after) ? static_cast<Real>(INFINITY) : after;
    return before - after + log_vol + (log_rev - log_fwd);
}

// phase A: Metropolis test per slot, the first accepted slot wins; store its state (reference k_accept_first_valid_move
// + k_store_exchange_move + k_convert_energies_to_log_weights).  Every block finds the winner for itself.
template <typename Real> __device__ void bd_phase_accept(const BDDevice<Real> &a, BDShared<Real> &sh, int off, int nblocks, int block) {
    const int live = min(a.B, a.P - off);
    if (threadIdx.x == 0) {
        sh.sel = a.B;
    }
    __syncthreads();
    if (a.ti.enabled) {
        const int inner = a.state[ST_INNER_USED];
        const Real outer_volume = a.ti.box_volume[0] - a.ti.inner_volume;
        for (int b = threadIdx.x; b < live; b += EX_THREADS) {
            const Real raw = ti_raw_log_probability<Real>(
                a.ti.targeting[b], a.ti.inner_volume, outer_volume, inner, a.M, a.ti.lse_src_max[b], a.ti.lse_src_sum[b], a.lse_after_max[b],
                a.lse_after_sum[b]);
            if (a.mh[off + b] < exp_(min_(raw, static_cast<Real>(0)))) {
                atomicMin(&sh.sel, b);
                break;
            }
        }
    } else {
        const Real before = nan_to_inf<Real>(add_(a.lse_before[0], log_(a.lse_before[1])));
        for (int b = threadIdx.x; b < live; b += EX_THREADS) {
            const Real after = nan_to_inf<Real>(add_(a.lse_after_max[b], log_(a.lse_after_sum[b])));
            const Real log_acc = min_(sub_(before, after), static_cast<Real>(0));
            if (a.mh[off + b] < exp_(log_acc)) {
                atomicMin(&sh.sel, b);
                break;
            }
        }
    }
    __syncthreads();
    const int sel = sh.sel;
    if (sel < a.B) {
        // the accepted proposal becomes the state: energies of the molecules it touches are re-derived (the same integers the
        // pair phase formed; no [B, M] energy matrix is kept), the weights are its row
        const Box3<Real> bx = load_box3<Real>(a.box);
        const int s = a.samples[sel];
        const Vec4<Real> *xold = a.prop_old + static_cast<size_t>(sel) * a.S;
        const Vec4<Real> *xnew = a.prop + static_cast<size_t>(sel) * a.S;
        const Vec4<Real> *pmol = a.pr + a.first + s * a.S;
        const Real reach_mol = sqrt(a.cutoff2) + 2 * a.r_bound[0];
        for (int item = block * EX_THREADS + threadIdx.x; item < a.M; item += nblocks * EX_THREADS) {
            const int j0 = static_cast<int>(a.mol_order[item]);
            const int mol = (j0 - a.first) / a.S;
            if (mol != s) {
                i128 delta = 0, unused = 0;
                if (bd_mol_delta(a, bx, xold, xnew, pmol, j0, reach_mol * reach_mol, delta, unused)) {
                    a.before_E[mol] += delta;
                }
            }
            a.logw_before[item] = a.logw_after[static_cast<size_t>(sel) * a.M + item];
        }
    }
    if (block == 0 && threadIdx.x == 0) {
        if (sel < a.B) {
            const int s = a.samples[sel];
            a.before_E[s] = a.total[sel];
            for (int i = 0; i < a.S; i++) {
                const Vec4<Real> p = a.prop[static_cast<size_t>(sel) * a.S + i];
                const int atom = a.first + s * a.S + i;
                a.coords[atom * 3 + 0] = static_cast<double>(p.x);
                a.coords[atom * 3 + 1] = static_cast<double>(p.y);
                a.coords[atom * 3 + 2] = static_cast<double>(p.z);
                a.xr[atom] = p;
            }
            a.state[ST_OFFSET] = off + sel + 1;
            a.num_accepted[0] += 1;
            if (a.ti.enabled) {
                a.ti.inner_flags[s] ^= 1; // the molecule changed region
            }
        } else {
            a.state[ST_OFFSET] = off + a.B;
        }
        a.state[ST_SELECTED] = sel;
        a.state[ST_ITER] += 1;
    }
    __syncthreads();
    if (a.ti.enabled && block == 0 && sel < a.B) {
        ti_partition(a.ti, a.M, sh.red.i); // for the next batch; the other blocks do not touch it in this phase
    }
}

// one phase per launch (host-driven loop, and compute_incremental_log_weights)
template <typename Real> __global__ void __launch_bounds__(EX_THREADS) k_bd_phase(BDDevice<Real> a, int phase) {
    __shared__ BDShared<Real> sh;
    const int off = a.state[ST_OFFSET];
    if (off >= a.P) {
        return;
    }
    switch (phase) {
    case 0:
        bd_phase_sample(a, sh, off, gridDim.x, blockIdx.x);
        break;
    case 1:
        bd_phase_energies(a, sh, off, gridDim.x, blockIdx.x);
        break;
    case 2:
        bd_phase_lse(a, sh, off, gridDim.x, blockIdx.x);
        break;
    default:
        bd_phase_accept(a, sh, off, gridDim.x, blockIdx.x);
        break;
    }
}

// all batches of a move in one cooperative launch
template <typename Real> __global__ void __launch_bounds__(EX_THREADS) k_bd_move(BDDevice<Real> a) {
    __shared__ BDShared<Real> sh;
    cg::grid_group grid = cg::this_grid();
    // every pass consumes at least one proposal, so P passes is an upper bound that cannot be reached unless the state is corrupt
    for (int pass = 0; pass < a.P; pass++) {
        const int off = *reinterpret_cast<volatile int *>(a.state + ST_OFFSET);
        if (off >= a.P) {
            break;
        }
        bd_phase_sample(a, sh, off, gridDim.x, blockIdx.x);
        grid.sync();
        bd_phase_energies(a, sh, off, gridDim.x, blockIdx.x);
        grid.sync();
        bd_phase_lse(a, sh, off, gridDim.x, blockIdx.x);
        grid.sync();
        bd_phase_accept(a, sh, off, gridDim.x, blockIdx.x);
        grid.sync();
    }
}

// log weights of all molecules, their log-sum-exp, and the bound on the size of a molecule (one block)
template <typename Real>
__global__ void __launch_bounds__(EX_THREADS) k_bd_initial_weights(
    int M, int S, int first, Real beta, const i128 *__restrict__ E, const Vec4<Real> *__restrict__ xr, Real *__restrict__ logw,
    Real *__restrict__ lse, Real *__restrict__ r_bound) {
    __shared__ LseScratch<Real> red;
    Real r2 = 0;
    for (int m = threadIdx.x; m < M; m += EX_THREADS) {
        logw[m] = log_weight<Real>(E[m], beta);
        const Vec4<Real> a0 = xr[first + m * S];
        for (int i = 1; i < S; i++) {
            const Vec4<Real> ai = xr[first + m * S + i];
            const Real dx = ai.x - a0.x, dy = ai.y - a0.y, dz = ai.z - a0.z;
            const Real d2 = dx * dx + dy * dy + dz * dz;
            r2 = d2 > r2 ? d2 : r2; // a NaN coordinate never raises the bound; its pairs are outside every cutoff test too
        }
    }
    red.v[threadIdx.x] = r2;
    __syncthreads();
    for (int w = EX_THREADS / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w && red.v[threadIdx.x + w] > red.v[threadIdx.x]) {
            red.v[threadIdx.x] = red.v[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // padded for the f32 rounding of the staged coordinates and of the rigid rotations applied during the move
        r_bound[0] = sqrt(red.v[0]) * static_cast<Real>(1.001) + static_cast<Real>(1e-4);
    }
    __syncthreads();
    Real mx, sum;
    block_sumexp(logw, M, red, mx, sum);
    if (threadIdx.x == 0) {
        lse[0] = mx;
        lse[1] = sum;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// SegmentedSumExp / SegmentedWeightedRandomSampler kernels (one block per segment)
template <typename Real>
__global__ void __launch_bounds__(EX_THREADS) k_segmented_sumexp(int num_segments, const int *__restrict__ offsets, const Real *__restrict__ vals, Real *__restrict__ out_max, Real *__restrict__ out_sum) {
    __shared__ LseScratch<Real> red;
    for (int s = blockIdx.x; s < num_segments; s += gridDim.x) {
        Real m, sum;
        block_sumexp(vals + offsets[s], offsets[s + 1] - offsets[s], red, m, sum);
        if (threadIdx.x == 0) {
            out_max[s] = m;
            out_sum[s] = sum;
        }
    }
}
template <typename Real>
__global__ void __launch_bounds__(EX_THREADS) k_segmented_gumbel_argmax(int num_segments, const int *__restrict__ offsets, const Real *__restrict__ logw, const Real *__restrict__ noise, int *__restrict__ out) {
    __shared__ LseScratch<Real> red;
    for (int s = blockIdx.x; s < num_segments; s += gridDim.x) {
        const int r = block_gumbel_argmax(logw + offsets[s], noise + offsets[s], offsets[s + 1] - offsets[s], red);
        if (threadIdx.x == 0) {
            out[s] = r;
        }
    }
}

template <typename Real> __global__ void k_rotate_coordinates(int N, int n_rot, const double *__restrict__ coords, const Real *__restrict__ quats, double *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (r >= n_rot || c >= N) {
        return;
    }
    Real x = static_cast<Real>(coords[c * 3 + 0]), y = static_cast<Real>(coords[c * 3 + 1]), z = static_cast<Real>(coords[c * 3 + 2]);
    rotate_by_quaternion(quats + r * 4, x, y, z);
    double *o = out + (static_cast<size_t>(c) * n_rot + r) * 3;
    o[0] = x;
    o[1] = y;
    o[2] = z;
}
template <typename Real>
__global__ void k_rotate_and_translate_mol(int N, int batch, const Vec4<Real> *__restrict__ xr, const double *__restrict__ box, const Real *__restrict__ quats, const Real *__restrict__ trans, Vec4<Real> *__restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) {
        return;
    }
    const Box3<Real> bx = load_box3<Real>(box);
    rotate_and_translate(N, xr, quats + b * 4, trans + b * 3, bx, true, out + static_cast<size_t>(b) * N);
}

// ===============================================================================================================
// host side
template <typename Real>
NonbondedMolEnergyPotential<Real>::NonbondedMolEnergyPotential(int N, const std::vector<std::vector<int>> &target_mols, double beta, double cutoff)
    : N_(N), num_target_mols_(static_cast<int>(target_mols.size())), beta_(static_cast<Real>(beta)),
      cutoff_squared_(static_cast<Real>(cutoff * cutoff)) {
    verify_group_idxs(N_, target_mols);
    if (num_target_mols_ <= 0) {
        throw std::runtime_error("must provide at least one target mol");
    }
    const MolLayout l = flatten_mols(target_mols);
    std::vector<int4> t(l.atom_idxs.size());
    for (size_t k = 0; k < t.size(); k++) {
        const int m = l.mol_idxs[k];
        t[k] = make_int4(l.atom_idxs[k], m, l.atom_idxs[l.mol_offsets[m]], l.atom_idxs[l.mol_offsets[m + 1] - 1]);
    }
    num_target_atoms_ = static_cast<int>(t.size());
    d_targets_.realloc(t.size());
    d_targets_.copy_from(t.data());
}

template <typename Real>
void NonbondedMolEnergyPotential<Real>::mol_energies_staged(const Vec4<Real> *d_xr, const Vec4<Real> *d_pr, const double *d_box, i128 *d_out, cudaStream_t stream) {
    TMB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(i128) * num_target_mols_, stream));
    if (num_target_atoms_ == 0) {
        return;
    }
    const int bx = ceil_div(num_target_atoms_, ME_THREADS);
    // enough slices of the atom range to fill the machine a few times over
    const int by = std::max(1, std::min(ceil_div(N_, ME_THREADS), ceil_div(4 * sm_count(), bx)));
    TMB_LAUNCH(k_mol_energies<Real>, dim3(bx, by), ME_THREADS, 0, stream, N_, num_target_atoms_, d_targets_.data, d_xr, d_pr, d_box, beta_, cutoff_squared_, d_out);
}

template <typename Real>
void NonbondedMolEnergyPotential<Real>::mol_energies_device(
    int N, int target_mols, const double *d_coords, const double *d_params, const double *d_box, i128 *d_out, cudaStream_t stream) {
    if (N != N_) {
        throw std::runtime_error("N != N_");
    }
    if (target_mols != num_target_mols_) {
        throw std::runtime_error("target_mols != num_target_mols_");
    }
    if (d_xr_.length != static_cast<size_t>(N_)) {
        d_xr_.realloc(N_);
        d_pr_.realloc(N_);
    }
    TMB_LAUNCH(k_stage_atoms<Real>, ceil_div(N_, 256), 256, 0, stream, N_, d_coords, d_params, d_xr_.data, d_pr_.data);
    mol_energies_staged(d_xr_.data, d_pr_.data, d_box, d_out, stream);
}

template <typename Real>
std::vector<i128> NonbondedMolEnergyPotential<Real>::mol_energies_host(int N, int P, const double *h_coords, const double *h_params, const double *h_box) {
    DeviceBuffer<double> d_coords(static_cast<size_t>(N) * 3), d_params(P), d_box(9);
    d_coords.copy_from(h_coords);
    d_params.copy_from(h_params);
    d_box.copy_from(h_box);
    DeviceBuffer<i128> d_out(num_target_mols_);
    cudaStream_t stream = main_stream();
    mol_energies_device(N, num_target_mols_, d_coords.data, d_params.data, d_box.data, d_out.data, stream);
    TMB_CUDA(cudaStreamSynchronize(stream));
    std::vector<i128> out(num_target_mols_);
    d_out.copy_to(out.data());
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
template <typename Real>
SegmentedSumExp<Real>::SegmentedSumExp(int max_vals_per_segment, int num_segments)
    : max_vals_per_segment_(max_vals_per_segment), num_segments_(num_segments) {}

template <typename Real> std::vector<Real> SegmentedSumExp<Real>::logsumexp_host(const std::vector<std::vector<Real>> &vals) {
    const int num_segments = static_cast<int>(vals.size());
    std::vector<int> offsets(num_segments + 1, 0);
    std::vector<Real> flat;
    for (int i = 0; i < num_segments; i++) {
        if (vals[i].empty()) {
            throw std::runtime_error("empty array not allowed");
        }
        flat.insert(flat.end(), vals[i].begin(), vals[i].end());
        offsets[i + 1] = static_cast<int>(flat.size());
    }
    const int total = offsets[num_segments];
    if (total > max_vals_per_segment_ * num_segments_) {
        throw std::runtime_error(
            "SegmentedSumExp::total values is greater than buffer size:  total_values=" + std::to_string(total) +
            ", buffer_size=" + std::to_string(max_vals_per_segment_ * num_segments_));
    }
    if (num_segments > num_segments_) {
        throw std::runtime_error(
            "SegmentedSumExp::number of segments must be less than or equal: num_segments=" + std::to_string(num_segments) +
            ", num_segments_=" + std::to_string(num_segments_));
    }
    std::vector<Real> out(num_segments);
    if (num_segments == 0) {
        return out;
    }
    DeviceBuffer<Real> d_vals(flat.size()), d_max(num_segments), d_sum(num_segments);
    DeviceBuffer<int> d_off(offsets.size());
    d_vals.copy_from(flat.data());
    d_off.copy_from(offsets.data());
    cudaStream_t stream = main_stream();
    TMB_LAUNCH(k_segmented_sumexp<Real>, std::min(num_segments, 4 * sm_count()), EX_THREADS, 0, stream, num_segments, d_off.data, d_vals.data, d_max.data, d_sum.data);
    TMB_CUDA(cudaStreamSynchronize(stream));
    std::vector<Real> h_max(num_segments), h_sum(num_segments);
    d_max.copy_to(h_max.data());
    d_sum.copy_to(h_sum.data());
    for (int i = 0; i < num_segments; i++) {
        out[i] = h_max[i] + std::log(h_sum[i]);
    }
    return out;
}

template <typename Real>
SegmentedWeightedRandomSampler<Real>::SegmentedWeightedRandomSampler(int max_vals_per_segment, int num_segments, int seed)
    : max_vals_per_segment_(max_vals_per_segment), num_segments_(num_segments),
      d_noise_(static_cast<size_t>(max_vals_per_segment) * num_segments) {
    TMB_CURAND(curandCreateGenerator(&rng_, CURAND_RNG_PSEUDO_DEFAULT));
    TMB_CURAND(curandSetPseudoRandomGeneratorSeed(rng_, seed));
}
template <typename Real> SegmentedWeightedRandomSampler<Real>::~SegmentedWeightedRandomSampler() {
    if (rng_ != nullptr) {
        curandDestroyGenerator(rng_);
    }
}

template <typename Real> std::vector<int> SegmentedWeightedRandomSampler<Real>::sample_host(const std::vector<std::vector<Real>> &weights) {
    const int num_segments = static_cast<int>(weights.size());
    std::vector<int> offsets(num_segments + 1, 0);
    std::vector<Real> log_probs;
    const Real inf = std::numeric_limits<Real>::infinity();
    for (int i = 0; i < num_segments; i++) {
        if (weights[i].empty()) {
            throw std::runtime_error("empty probability distribution not allowed");
        }
        for (Real w : weights[i]) {
            if (w == inf) {
                throw std::runtime_error("unable to use infinity as a weight");
            } else if (std::isnan(w)) {
                throw std::runtime_error("unable to use nan as a weight");
            } else if (w < static_cast<Real>(0)) {
                throw std::runtime_error("unable to use negative values as a weight");
            }
            log_probs.push_back(std::log(w));
        }
        offsets[i + 1] = static_cast<int>(log_probs.size());
    }
    const int total = offsets[num_segments];
    if (total > max_vals_per_segment_ * num_segments_) {
        throw std::runtime_error(
            "SegmentedWeightedRandomerSampler::total values is greater than buffer size:  vals_per_segment * num_segments=" +
            std::to_string(total) + ", buffer_size=" + std::to_string(max_vals_per_segment_ * num_segments_));
    }
    if (num_segments != num_segments_) {
        throw std::runtime_error(
            "SegmentedWeightedRandomerSampler::number of segments don't match: num_segments=" + std::to_string(num_segments) +
            ", num_segments_=" + std::to_string(num_segments_));
    }
    DeviceBuffer<Real> d_logp(log_probs.size());
    DeviceBuffer<int> d_off(offsets.size()), d_out(num_segments);
    d_logp.copy_from(log_probs.data());
    d_off.copy_from(offsets.data());
    cudaStream_t stream = main_stream();
    TMB_CURAND(curandSetStream(rng_, stream));
    // the whole noise buffer is drawn per call, as in the reference (sample_device): same stream position afterwards
    TMB_CURAND(gen_uniform(rng_, d_noise_.data, d_noise_.length));
    TMB_LAUNCH(k_segmented_gumbel_argmax<Real>, std::min(num_segments, 4 * sm_count()), EX_THREADS, 0, stream, num_segments, d_off.data, d_logp.data, d_noise_.data, d_out.data);
    TMB_CUDA(cudaStreamSynchronize(stream));
    std::vector<int> out(num_segments);
    d_out.copy_to(out.data());
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
static size_t round_up_even(size_t n) { return n + (n % 2); }

template <typename Real>
BDExchangeMove<Real>::BDExchangeMove(
    int N, const std::vector<std::vector<int>> &target_mols, const std::vector<double> &params, double temperature, double nb_beta,
    double cutoff, int seed, int num_proposals_per_move, int interval, int batch_size)
    : BDExchangeMove(
          N, target_mols, params, temperature, nb_beta, cutoff, seed, num_proposals_per_move, interval, batch_size,
          static_cast<size_t>(3) * std::max(num_proposals_per_move, 0)) {}

template <typename Real>
BDExchangeMove<Real>::BDExchangeMove(
    int N, const std::vector<std::vector<int>> &target_mols, const std::vector<double> &params, double temperature, double nb_beta,
    double cutoff, int seed, int num_proposals_per_move, int interval, int batch_size, size_t translation_buffer_size)
    : Mover(interval), N_(N), mol_size_(static_cast<int>(target_mols[0].size())), num_proposals_per_move_(num_proposals_per_move),
      num_target_mols_(static_cast<int>(target_mols.size())), nb_beta_(static_cast<Real>(nb_beta)),
      beta_(static_cast<Real>(1.0 / (BOLTZ * temperature))), cutoff_squared_(static_cast<Real>(cutoff * cutoff)), batch_size_(batch_size),
      first_atom_(target_mols[0].empty() ? 0 : *std::min_element(target_mols[0].begin(), target_mols[0].end())),
      mol_potential_(N, target_mols, nb_beta, cutoff), sorter_(N), d_anchor_atoms_(num_target_mols_), d_mol_order_(num_target_mols_),
      d_r_bound_(1), d_xs_(N), d_ps_(N), d_col_atom_(N), d_chunk_ctr_(ceil_div(std::max(N, 1), ME_THREADS)),
      d_chunk_ext_(ceil_div(std::max(N, 1), ME_THREADS)), d_params_(params.size()), d_xr_(N), d_pr_(N),
      d_prop_(static_cast<size_t>(batch_size) * std::max(1, mol_size_)), d_prop_old_(static_cast<size_t>(batch_size) * std::max(1, mol_size_)), d_before_E_(num_target_mols_),
      d_total_(batch_size), d_logw_before_(num_target_mols_),
      d_logw_after_(static_cast<size_t>(batch_size) * num_target_mols_), d_lse_before_(2), d_lse_after_max_(batch_size),
      d_lse_after_sum_(batch_size), d_samples_(batch_size), d_state_(ST_WORDS), d_num_accepted_(1),
      d_quat_(round_up_even(static_cast<size_t>(4) * std::max(num_proposals_per_move, 0))), d_trans_(translation_buffer_size),
      d_sample_noise_(static_cast<size_t>(num_target_mols_) * std::max(num_proposals_per_move, 0)), d_mh_(std::max(num_proposals_per_move, 0)) {
    if (num_proposals_per_move_ <= 0) {
        throw std::runtime_error("proposals per move must be greater than 0");
    }
    if (mol_size_ == 0) {
        throw std::runtime_error("must provide non-empty molecule indices");
    }
    verify_mols_contiguous(target_mols);
    for (const auto &m : target_mols) {
        if (static_cast<int>(m.size()) != mol_size_) {
            throw std::runtime_error("only support running with mols with constant size, got mixed sizes");
        }
    }
    if (static_cast<int>(params.size()) != N * P_PER_ATOM) {
        throw std::runtime_error("Number of parameters must match N");
    }
    d_params_.copy_from(params.data());
    {
        std::vector<unsigned int> anchors(num_target_mols_);
        for (int m = 0; m < num_target_mols_; m++) {
            anchors[m] = static_cast<unsigned int>(first_atom_ + m * mol_size_);
        }
        d_anchor_atoms_.copy_from(anchors.data());
        d_mol_order_.copy_from(anchors.data());
    }
    d_r_bound_.zero();
    d_lse_before_.zero();
    d_lse_after_max_.zero();
    d_lse_after_sum_.zero();
    d_num_accepted_.zero();
    d_state_.zero();
    d_logw_before_.zero();
    d_logw_after_.zero();
    // four generators so that the sequences do not depend on the batch size (reference bd_exchange_move.cu:96-108)
    TMB_CURAND(curandCreateGenerator(&rng_quat_, CURAND_RNG_PSEUDO_DEFAULT));
    TMB_CURAND(curandSetPseudoRandomGeneratorSeed(rng_quat_, seed));
    TMB_CURAND(curandCreateGenerator(&rng_trans_, CURAND_RNG_PSEUDO_DEFAULT));
    TMB_CURAND(curandSetPseudoRandomGeneratorSeed(rng_trans_, seed + 1));
    TMB_CURAND(curandCreateGenerator(&rng_samples_, CURAND_RNG_PSEUDO_DEFAULT));
    TMB_CURAND(curandSetPseudoRandomGeneratorSeed(rng_samples_, seed + 2));
    TMB_CURAND(curandCreateGenerator(&rng_mh_, CURAND_RNG_PSEUDO_DEFAULT));
    TMB_CURAND(curandSetPseudoRandomGeneratorSeed(rng_mh_, seed + 3));

    const char *mode = std::getenv("TMB_BD_LOOP");
    host_loop_ = mode != nullptr && std::strcmp(mode, "host") == 0;
    int per_sm = 0;
    TMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bd_move<Real>, EX_THREADS, 0));
    if (per_sm < 1) {
        throw std::runtime_error("BDExchangeMove: the move kernel does not fit on an SM");
    }
    const char *bps = std::getenv("TMB_BD_BLOCKS_PER_SM"); // measurement knob
    coop_blocks_ = sm_count() * std::min(per_sm, bps != nullptr ? std::max(1, std::atoi(bps)) : 3); // 3 measured best in f32 (80 registers)
    TMB_CUDA(cudaDeviceSynchronize());
}

template <typename Real> BDExchangeMove<Real>::~BDExchangeMove() {
    for (curandGenerator_t g : {rng_quat_, rng_trans_, rng_samples_, rng_mh_}) {
        if (g != nullptr) {
            curandDestroyGenerator(g);
        }
    }
}

template <typename Real> BDDevice<Real> BDExchangeMove<Real>::device_args(double *d_coords, const double *d_box, bool scale, bool sample) {
    BDDevice<Real> a;
    a.N = N_;
    a.M = num_target_mols_;
    a.S = mol_size_;
    a.B = batch_size_;
    a.P = num_proposals_per_move_;
    a.first = first_atom_;
    a.sample = sample ? 1 : 0;
    a.scale = scale ? 1 : 0;
    a.nb_beta = nb_beta_;
    a.beta = beta_;
    a.cutoff2 = cutoff_squared_;
    a.coords = d_coords;
    a.box = d_box;
    a.xr = d_xr_.data;
    a.pr = d_pr_.data;
    a.prop = d_prop_.data;
    a.before_E = d_before_E_.data;
    a.prop_old = d_prop_old_.data;
    a.total = d_total_.data;
    a.logw_before = d_logw_before_.data;
    a.logw_after = d_logw_after_.data;
    a.lse_before = d_lse_before_.data;
    a.lse_after_max = d_lse_after_max_.data;
    a.lse_after_sum = d_lse_after_sum_.data;
    a.samples = d_samples_.data;
    a.mol_order = d_mol_order_.data;
    a.r_bound = d_r_bound_.data;
    a.state = d_state_.data;
    a.num_accepted = d_num_accepted_.data;
    a.quat = d_quat_.data;
    a.trans = d_trans_.data;
    a.sample_noise = d_sample_noise_.data;
    a.mh = d_mh_.data;
    std::memset(&a.ti, 0, sizeof(a.ti));
    return a;
}

template <typename Real> void BDExchangeMove<Real>::initial_log_weights_device(double *d_coords, const double *d_box, cudaStream_t stream) {
    TMB_LAUNCH(k_stage_atoms<Real>, ceil_div(N_, 256), 256, 0, stream, N_, d_coords, d_params_.data, d_xr_.data, d_pr_.data);
    // molecules along the Hilbert curve of their first atoms: the order of the all-molecule energy sweep below and of the pair
    // phase of the move (entries of molecules moved during the move go stale, which only costs coherence)
    sorter_.sort_device(num_target_mols_, d_anchor_atoms_.data, d_coords, d_box, d_mol_order_.data, stream);
    const int n_chunks = ceil_div(N_, ME_THREADS);
    const int row_chunks = ceil_div(num_target_mols_ * mol_size_, ME_THREADS);
    TMB_LAUNCH(
        k_me_chunks<Real>, n_chunks, ME_THREADS, 0, stream, N_, num_target_mols_, mol_size_, first_atom_, d_mol_order_.data, d_xr_.data,
        d_pr_.data, d_box, d_xs_.data, d_ps_.data, d_col_atom_.data, d_chunk_ctr_.data, d_chunk_ext_.data);
    TMB_CUDA(cudaMemsetAsync(d_before_E_.data, 0, d_before_E_.bytes(), stream));
    const int by = std::max(1, std::min(n_chunks, ceil_div(6 * sm_count(), row_chunks)));
    TMB_LAUNCH(
        k_mol_energies_sorted<Real>, dim3(row_chunks, by), ME_THREADS, 0, stream, N_, num_target_mols_, mol_size_, first_atom_, d_xs_.data,
        d_ps_.data, d_col_atom_.data, d_chunk_ctr_.data, d_chunk_ext_.data, d_box, nb_beta_, cutoff_squared_, d_before_E_.data);
    TMB_LAUNCH(
        k_bd_initial_weights<Real>, 1, EX_THREADS, 0, stream, num_target_mols_, mol_size_, first_atom_, beta_, d_before_E_.data, d_xr_.data,
        d_logw_before_.data, d_lse_before_.data, d_r_bound_.data);
}

template <typename Real> void BDExchangeMove<Real>::run_phase(int phase, const BDDevice<Real> &a, cudaStream_t stream) {
    TMB_LAUNCH(k_bd_phase<Real>, coop_blocks_, EX_THREADS, 0, stream, a, phase);
}

template <typename Real> void BDExchangeMove<Real>::move(int N, double *d_coords, double *d_box, cudaStream_t stream) {
    if (N != N_) {
        throw std::runtime_error("N != N_");
    }
    step_++;
    if (step_ % interval_ != 0) {
        return;
    }
    set_generator_streams(stream);
    TMB_CUDA(cudaMemsetAsync(d_state_.data, 0, d_state_.bytes(), stream));

    initial_log_weights_device(d_coords, d_box, stream);

    // all noise of the move up front, one draw per generator (reference bd_exchange_move.cu:197-201)
    TMB_CURAND(gen_normal(rng_quat_, d_quat_.data, d_quat_.length));
    TMB_CURAND(gen_uniform(rng_trans_, d_trans_.data, d_trans_.length));
    TMB_CURAND(gen_uniform(rng_samples_, d_sample_noise_.data, d_sample_noise_.length));
    TMB_CURAND(gen_uniform(rng_mh_, d_mh_.data, d_mh_.length));

    BDDevice<Real> a = device_args(d_coords, d_box, true, true);
    run_proposals(a, stream);
    num_attempted_ += num_proposals_per_move_;
}

template <typename Real> void BDExchangeMove<Real>::set_generator_streams(cudaStream_t stream) {
    TMB_CURAND(curandSetStream(rng_quat_, stream));
    TMB_CURAND(curandSetStream(rng_trans_, stream));
    TMB_CURAND(curandSetStream(rng_samples_, stream));
    TMB_CURAND(curandSetStream(rng_mh_, stream));
}

template <typename Real> void BDExchangeMove<Real>::run_proposals(BDDevice<Real> &a, cudaStream_t stream) {
    if (host_loop_) {
        int off = 0;
        while (off < num_proposals_per_move_) {
            for (int phase = 0; phase < 4; phase++) {
                run_phase(phase, a, stream);
            }
            TMB_CUDA(cudaMemcpyAsync(&off, d_state_.data + ST_OFFSET, sizeof(int), cudaMemcpyDeviceToHost, stream));
            TMB_CUDA(cudaStreamSynchronize(stream));
        }
    } else {
        // a grid barrier costs in proportion to the CTAs that meet at it: small batches get a grid sized for their pair phase
        // (about two 32-molecule tasks per warp) instead of the whole machine
        const int chunks = ceil_div(num_target_mols_ + (N_ - num_target_mols_ * mol_size_), WARP);
        const long long want = std::max<long long>(batch_size_, ceil_div<long long>(static_cast<long long>(batch_size_) * chunks, 2 * (EX_THREADS / WARP)));
        const int blocks = static_cast<int>(std::min<long long>(coop_blocks_, std::max<long long>(want, 8)));
        void *args[] = {&a};
        TMB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_bd_move<Real>), dim3(blocks), dim3(EX_THREADS), args, 0, stream));
        g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    }
}

template <typename Real>
std::vector<std::vector<Real>> BDExchangeMove<Real>::compute_incremental_log_weights_host(
    int N, const double *h_coords, const double *h_box, const int *h_mol_idxs, const Real *h_quaternions, const Real *h_translations) {
    if (N != N_) {
        throw std::runtime_error("N != N_");
    }
    DeviceBuffer<double> d_coords(static_cast<size_t>(N) * 3), d_box(9);
    d_coords.copy_from(h_coords);
    d_box.copy_from(h_box);
    // the caller's batch goes to the head of the noise buffers (P >= B)
    TMB_CUDA(cudaMemcpy(d_quat_.data, h_quaternions, sizeof(Real) * 4 * batch_size_, cudaMemcpyHostToDevice));
    TMB_CUDA(cudaMemcpy(d_trans_.data, h_translations, sizeof(Real) * 3 * batch_size_, cudaMemcpyHostToDevice));
    TMB_CUDA(cudaMemcpy(d_samples_.data, h_mol_idxs, sizeof(int) * batch_size_, cudaMemcpyHostToDevice));
    cudaStream_t stream = main_stream();
    TMB_CUDA(cudaMemsetAsync(d_state_.data, 0, d_state_.bytes(), stream));
    initial_log_weights_device(d_coords.data, d_box.data, stream);
    // translations are used as given, the molecules are the caller's (reference bd_exchange_move.cu:437-487)
    BDDevice<Real> a = device_args(d_coords.data, d_box.data, false, false);
    for (int phase = 0; phase < 3; phase++) {
        run_phase(phase, a, stream);
    }
    TMB_CUDA(cudaStreamSynchronize(stream));
    std::vector<Real> flat = get_after_log_weights();
    std::vector<std::vector<Real>> out(batch_size_);
    for (int b = 0; b < batch_size_; b++) {
        out[b].assign(flat.begin() + static_cast<size_t>(b) * num_target_mols_, flat.begin() + static_cast<size_t>(b + 1) * num_target_mols_);
    }
    return out;
}

template <typename Real> std::vector<Real> BDExchangeMove<Real>::compute_initial_log_weights_host(int N, const double *h_coords, const double *h_box) {
    if (N != N_) {
        throw std::runtime_error("N != N_");
    }
    DeviceBuffer<double> d_coords(static_cast<size_t>(N) * 3), d_box(9);
    d_coords.copy_from(h_coords);
    d_box.copy_from(h_box);
    cudaStream_t stream = main_stream();
    initial_log_weights_device(d_coords.data, d_box.data, stream);
    TMB_CUDA(cudaStreamSynchronize(stream));
    return get_before_log_weights();
}

template <typename Real> std::vector<Real> BDExchangeMove<Real>::get_before_log_weights() {
    std::vector<Real> out(d_logw_before_.length);
    d_logw_before_.copy_to(out.data());
    return out;
}
template <typename Real> std::vector<Real> BDExchangeMove<Real>::get_after_log_weights() {
    std::vector<Real> out(d_logw_after_.length);
    d_logw_after_.copy_to(out.data());
    return out;
}

template <typename Real> static Real host_nan_to_inf(Real v) { return std::isnan(v) ? std::numeric_limits<Real>::infinity() : v; }

template <typename Real> double BDExchangeMove<Real>::raw_log_probability_host() {
    Real before[2], after_max, after_sum;
    d_lse_before_.copy_to(before);
    TMB_CUDA(cudaMemcpy(&after_max, d_lse_after_max_.data, sizeof(Real), cudaMemcpyDeviceToHost));
    TMB_CUDA(cudaMemcpy(&after_sum, d_lse_after_sum_.data, sizeof(Real), cudaMemcpyDeviceToHost));
    const Real b = host_nan_to_inf<Real>(before[0] + std::log(before[1]));
    const Real a = host_nan_to_inf<Real>(after_max + std::log(after_sum));
    return static_cast<double>(b - a);
}
template <typename Real> double BDExchangeMove<Real>::log_probability_host() { return std::fmin(raw_log_probability_host(), 0.0); } // NaN-safe like the CUDA host min()

template <typename Real> size_t BDExchangeMove<Real>::n_accepted() const {
    u64 v = 0;
    d_num_accepted_.copy_to(&v);
    return static_cast<size_t>(v);
}

template <typename Real> std::vector<double> BDExchangeMove<Real>::get_params() {
    std::vector<double> out(d_params_.length);
    d_params_.copy_to(out.data());
    return out;
}
template <typename Real> void BDExchangeMove<Real>::set_params(const std::vector<double> &params) {
    if (d_params_.length != params.size()) {
        throw std::runtime_error("number of params don't match");
    }
    d_params_.copy_from(params.data());
}
template <typename Real> void BDExchangeMove<Real>::set_params_device(int size, const double *d_p, cudaStream_t stream) {
    if (d_params_.length != static_cast<size_t>(size)) {
        throw std::runtime_error("number of params don't match");
    }
    TMB_CUDA(cudaMemcpyAsync(d_params_.data, d_p, d_params_.bytes(), cudaMemcpyDeviceToDevice, stream));
}

// ---------------------------------------------------------------------------------------------------------------
template <typename Real>
std::vector<Real> atom_by_atom_energies(
    int N, const std::vector<int> &target_atoms, const double *coords, const double *params, const double *box, Real nb_beta, Real cutoff) {
    const int T = static_cast<int>(target_atoms.size());
    std::vector<Real> out(static_cast<size_t>(T) * N);
    if (T == 0 || N == 0) {
        return out;
    }
    for (int t : target_atoms) {
        if (t < 0 || t >= N) {
            throw std::runtime_error("target atoms must be between 0 and N");
        }
    }
    DeviceBuffer<double> d_coords(static_cast<size_t>(N) * 3), d_params(static_cast<size_t>(N) * 4), d_box(9);
    DeviceBuffer<int> d_t(T);
    DeviceBuffer<Vec4<Real>> d_xr(N), d_pr(N);
    DeviceBuffer<Real> d_out(out.size());
    d_coords.copy_from(coords);
    d_params.copy_from(params);
    d_box.copy_from(box);
    d_t.copy_from(target_atoms.data());
    cudaStream_t stream = main_stream();
    TMB_LAUNCH(k_stage_atoms<Real>, ceil_div(N, 256), 256, 0, stream, N, d_coords.data, d_params.data, d_xr.data, d_pr.data);
    const Real cutoff2 = cutoff * cutoff; // formed in Real (reference all_atom_energies.cu:23)
    TMB_LAUNCH(k_atom_by_atom<Real>, dim3(ceil_div(N, 256), T), 256, 0, stream, N, T, d_t.data, d_xr.data, d_pr.data, d_box.data, nb_beta, cutoff2, d_out.data);
    TMB_CUDA(cudaStreamSynchronize(stream));
    d_out.copy_to(out.data());
    return out;
}

template <typename Real> void rotate_coordinates_host(int N, int n_rotations, const double *coords, const Real *quaternions, double *out) {
    if (N == 0 || n_rotations == 0) {
        return;
    }
    DeviceBuffer<double> d_coords(static_cast<size_t>(N) * 3), d_out(static_cast<size_t>(N) * n_rotations * 3);
    DeviceBuffer<Real> d_q(static_cast<size_t>(n_rotations) * 4);
    d_coords.copy_from(coords);
    d_q.copy_from(quaternions);
    cudaStream_t stream = main_stream();
    TMB_LAUNCH(k_rotate_coordinates<Real>, dim3(ceil_div(n_rotations, 256), N), 256, 0, stream, N, n_rotations, d_coords.data, d_q.data, d_out.data);
    TMB_CUDA(cudaStreamSynchronize(stream));
    d_out.copy_to(out);
}

template <typename Real>
void rotate_and_translate_mol_host(
    int N, int batch_size, const double *mol_coords, const double *box, const Real *quaternions, const Real *translations, double *out) {
    if (N == 0 || batch_size == 0) {
        return;
    }
    std::vector<Vec4<Real>> h_x(N);
    for (int i = 0; i < N; i++) {
        h_x[i].x = static_cast<Real>(mol_coords[i * 3 + 0]);
        h_x[i].y = static_cast<Real>(mol_coords[i * 3 + 1]);
        h_x[i].z = static_cast<Real>(mol_coords[i * 3 + 2]);
        h_x[i].w = 0;
    }
    DeviceBuffer<Vec4<Real>> d_x(N), d_out(static_cast<size_t>(N) * batch_size);
    DeviceBuffer<double> d_box(9);
    DeviceBuffer<Real> d_q(static_cast<size_t>(batch_size) * 4), d_t(static_cast<size_t>(batch_size) * 3);
    d_x.copy_from(h_x.data());
    d_box.copy_from(box);
    d_q.copy_from(quaternions);
    d_t.copy_from(translations);
    cudaStream_t stream = main_stream();
    TMB_LAUNCH(k_rotate_and_translate_mol<Real>, ceil_div(batch_size, 128), 128, 0, stream, N, batch_size, d_x.data, d_box.data, d_q.data, d_t.data, d_out.data);
    TMB_CUDA(cudaStreamSynchronize(stream));
    std::vector<Vec4<Real>> h_out(d_out.length);
    d_out.copy_to(h_out.data());
    for (size_t k = 0; k < h_out.size(); k++) {
        out[k * 3 + 0] = static_cast<double>(h_out[k].x);
        out[k * 3 + 1] = static_cast<double>(h_out[k].y);
        out[k * 3 + 2] = static_cast<double>(h_out[k].z);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// targeted insertion (reference tibd_exchange_move.cu, kernels/k_exchange.cu:440-545, k_translations.cuh, gpu_utils.cu:24-31)
constexpr int TI_STATE_THREADS = 128; // the reference's DEFAULT_THREADS_PER_BLOCK: one cuRAND state per thread of one block

__global__ void k_ti_init_states(int count, int seed, curandState_t *states) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) {
        curand_init(seed + idx, 0, 0, &states[idx]);
    }
}

// Pairs of translations: even slots uniformly inside the sphere (direction from three normals, radius^(1/3) law),
// odd slots anywhere in the box but outside the sphere (rejection).  The state a thread ends with is written back
// rotated so that the sequence does not depend on how many translations a call draws (k_translations.cuh:7-93).
template <typename Real>
__global__ void k_ti_translations(int num_translations, const double *__restrict__ box, const Real *__restrict__ center, Real radius, curandState_t *states, Real *__restrict__ out) {
    const int stride = gridDim.x * blockDim.x;
    const Real cx = center[0], cy = center[1], cz = center[2];
    curandState_t local = states[threadIdx.x];
    const Real bx = box[0], by = box[4], bz = box[8];
    const Real ibx = 1 / bx, iby = 1 / by, ibz = 1 / bz;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < num_translations * 2; idx += stride) {
        if (idx % 2 == 0) {
            Real x = curand_normal(&local);
            Real y = curand_normal(&local);
            Real z = curand_normal(&local);
            Real rad = curand_uniform(&local);
            const Real norm = sqrt((x * x) + (y * y) + (z * z));
            x /= norm;
            y /= norm;
            z /= norm;
            rad = cbrt(rad);
            out[idx * 3 + 0] = (x * rad * radius) + cx;
            out[idx * 3 + 1] = (y * rad * radius) + cy;
            out[idx * 3 + 2] = (z * rad * radius) + cz;
        } else {
            const Real r2 = radius * radius;
            for (int it = 0; it < 1000; it++) {
                const Real x = curand_normal(&local) * bx;
                const Real y = curand_normal(&local) * by;
                const Real z = curand_normal(&local) * bz;
                Real dx = x - cx, dy = y - cy, dz = z - cz;
                dx -= bx * nearbyint(dx * ibx);
                dy -= by * nearbyint(dy * iby);
                dz -= bz * nearbyint(dz * ibz);
                const Real dist = (dx * dx) + (dy * dy) + (dz * dz);
                if (dist >= r2) {
                    out[idx * 3 + 0] = x;
                    out[idx * 3 + 1] = y;
                    out[idx * 3 + 2] = z;
                    break;
                }
            }
        }
    }
    int slot = static_cast<int>(threadIdx.x) - ((num_translations * 2) % stride);
    slot = slot >= 0 ? slot : stride + slot;
    states[slot] = local;
}

// centroid of the ligand atoms (fixed-point sum), box volume, inside/outside flag of every molecule and the partition
// (k_compute_centroid_of_atoms, k_compute_box_volume, k_flag_mols_inner_outer; one block)
template <typename Real>
__global__ void __launch_bounds__(EX_THREADS) k_ti_setup(
    int num_center_atoms, const int *__restrict__ center_atoms, int M, int S, int first, const int *__restrict__ mol_first /* or null */,
    const int *__restrict__ mol_sizes /* or null */, const double *__restrict__ coords, const double *__restrict__ box, Real square_radius,
    Real *center, Real *box_volume, TIDevice<Real> t) {
    __shared__ unsigned long long acc[3];
    __shared__ int scratch[EX_THREADS];
    if (threadIdx.x < 3) {
        acc[threadIdx.x] = 0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < num_center_atoms; k += EX_THREADS) {
        const int atom = center_atoms[k];
        atomicAdd(acc + 0, to_fixed<FIXED_EXPONENT>(static_cast<Real>(coords[atom * 3 + 0])));
        atomicAdd(acc + 1, to_fixed<FIXED_EXPONENT>(static_cast<Real>(coords[atom * 3 + 1])));
        atomicAdd(acc + 2, to_fixed<FIXED_EXPONENT>(static_cast<Real>(coords[atom * 3 + 2])));
    }
    __syncthreads();
    const Real n = static_cast<Real>(num_center_atoms);
    const Real cx = fixed_to_real<Real>(acc[0]) / n, cy = fixed_to_real<Real>(acc[1]) / n, cz = fixed_to_real<Real>(acc[2]) / n;
    if (threadIdx.x == 0) {
        center[0] = cx;
        center[1] = cy;
        center[2] = cz;
        if (box_volume != nullptr) {
            const Real v = box[0] * box[4] * box[8]; // product in double, then rounded (k_exchange.cu:219-229)
            box_volume[0] = v;
        }
    }
    const Real bx = box[0], by = box[4], bz = box[8];
    const Real ibx = 1 / bx, iby = 1 / by, ibz = 1 / bz;
    for (int m = threadIdx.x; m < M; m += EX_THREADS) {
        const int start = mol_first != nullptr ? mol_first[m] : first + m * S;
        const int size = mol_sizes != nullptr ? mol_sizes[m] : S;
        u64 ax = 0, ay = 0, az = 0;
        for (int atom = start; atom < start + size; atom++) {
            ax += to_fixed<FIXED_EXPONENT>(static_cast<Real>(coords[atom * 3 + 0]));
            ay += to_fixed<FIXED_EXPONENT>(static_cast<Real>(coords[atom * 3 + 1]));
            az += to_fixed<FIXED_EXPONENT>(static_cast<Real>(coords[atom * 3 + 2]));
        }
        const Real ns = static_cast<Real>(size);
        Real dx = fixed_to_real<Real>(ax) / ns, dy = fixed_to_real<Real>(ay) / ns, dz = fixed_to_real<Real>(az) / ns;
        dx -= cx;
        dy -= cy;
        dz -= cz;
        dx -= bx * nearbyint(dx * ibx);
        dy -= by * nearbyint(dy * iby);
        dz -= bz * nearbyint(dz * ibz);
        const Real dist = (dx * dx) + (dy * dy) + (dz * dz);
        t.inner_flags[m] = dist < square_radius ? 1 : 0;
    }
    __syncthreads();
    if (t.partition != nullptr) {
        ti_partition(t, M, scratch);
    }
}

template <typename Real>
TIBDExchangeMove<Real>::TIBDExchangeMove(
    int N, const std::vector<int> &ligand_idxs, const std::vector<std::vector<int>> &target_mols, const std::vector<double> &params,
    double temperature, double nb_beta, double cutoff, double radius, int seed, int num_proposals_per_move, int interval, int batch_size)
    : BDExchangeMove<Real>(
          N, target_mols, params, temperature, nb_beta, cutoff, seed, num_proposals_per_move, interval, batch_size,
          static_cast<size_t>(6) * std::max(num_proposals_per_move, 0)),
      radius_(static_cast<Real>(radius)), inner_volume_(static_cast<Real>((4.0 / 3.0) * M_PI * std::pow(radius, 3))),
      d_ligand_idxs_(ligand_idxs.size()), d_inner_flags_(this->num_target_mols_), d_partition_(this->num_target_mols_), d_inner_count_(1),
      d_targeting_(batch_size), d_center_(3), d_box_volume_(1), d_uniform_(std::max(num_proposals_per_move, 0)),
      d_src_logw_(static_cast<size_t>(batch_size) * this->num_target_mols_), d_dest_logw_(static_cast<size_t>(batch_size) * this->num_target_mols_),
      d_lse_src_max_(batch_size), d_lse_src_sum_(batch_size) {
    if (radius <= 0.0) {
        throw std::runtime_error("radius must be greater than 0.0");
    }
    for (int idx : ligand_idxs) {
        if (idx < 0 || idx >= N) {
            throw std::runtime_error("ligand indices must be between 0 and N");
        }
    }
    d_ligand_idxs_.copy_from(ligand_idxs.data());
    d_box_volume_.zero();
    d_lse_src_max_.zero();
    d_lse_src_sum_.zero();
    d_inner_count_.zero();
    d_inner_flags_.zero();
    d_src_logw_.zero();
    d_dest_logw_.zero();
    // targeting the sphere with nothing inside it: log probability 0 before the first move (tibd_exchange_move.cu:104-108)
    std::vector<int> ones(batch_size, 1);
    d_targeting_.copy_from(ones.data());
    TMB_CUDA(cudaMalloc(&d_rand_states_, sizeof(curandState_t) * TI_STATE_THREADS));
    // seed + 4: the first four seeds belong to the host generators (tibd_exchange_move.cu:66-71)
    TMB_LAUNCH(k_ti_init_states, 1, TI_STATE_THREADS, 0, nullptr, TI_STATE_THREADS, seed + 4, reinterpret_cast<curandState_t *>(d_rand_states_));
    TMB_CUDA(cudaDeviceSynchronize());
}

template <typename Real> TIBDExchangeMove<Real>::~TIBDExchangeMove() {
    if (d_rand_states_ != nullptr) {
        cudaFree(d_rand_states_);
    }
}

template <typename Real> void TIBDExchangeMove<Real>::move(int N, double *d_coords, double *d_box, cudaStream_t stream) {
    if (N != this->N_) {
        throw std::runtime_error("N != N_");
    }
    this->step_++;
    if (this->step_ % this->interval_ != 0) {
        return;
    }
    this->set_generator_streams(stream);
    TMB_CUDA(cudaMemsetAsync(this->d_state_.data, 0, this->d_state_.bytes(), stream));
    this->initial_log_weights_device(d_coords, d_box, stream);

    BDDevice<Real> a = this->device_args(d_coords, d_box, false, true);
    a.ti.enabled = 1;
    a.ti.inner_volume = inner_volume_;
    a.ti.box_volume = d_box_volume_.data;
    a.ti.uniform = d_uniform_.data;
    a.ti.inner_flags = d_inner_flags_.data;
    a.ti.partition = d_partition_.data;
    a.ti.inner_count = d_inner_count_.data;
    a.ti.targeting = d_targeting_.data;
    a.ti.src_logw = d_src_logw_.data;
    a.ti.dest_logw = d_dest_logw_.data;
    a.ti.lse_src_max = d_lse_src_max_.data;
    a.ti.lse_src_sum = d_lse_src_sum_.data;

    TMB_LAUNCH(
        k_ti_setup<Real>, 1, EX_THREADS, 0, stream, static_cast<int>(d_ligand_idxs_.length), d_ligand_idxs_.data, this->num_target_mols_,
        this->mol_size_, this->first_atom_, static_cast<const int *>(nullptr), static_cast<const int *>(nullptr), d_coords, d_box,
        radius_ * radius_, d_center_.data, d_box_volume_.data, a.ti);

    // all noise of the move up front (tibd_exchange_move.cu:157-170); the region choice draws from the translations generator
    TMB_CURAND(gen_uniform(this->rng_mh_, this->d_mh_.data, this->d_mh_.length));
    TMB_CURAND(gen_uniform(this->rng_trans_, d_uniform_.data, d_uniform_.length));
    TMB_CURAND(gen_normal(this->rng_quat_, this->d_quat_.data, this->d_quat_.length));
    TMB_CURAND(gen_uniform(this->rng_samples_, this->d_sample_noise_.data, this->d_sample_noise_.length));
    TMB_LAUNCH(
        k_ti_translations<Real>, 1, TI_STATE_THREADS, 0, stream, this->num_proposals_per_move_, d_box, d_center_.data, radius_,
        reinterpret_cast<curandState_t *>(d_rand_states_), this->d_trans_.data);

    this->run_proposals(a, stream);
    this->num_attempted_ += this->num_proposals_per_move_;
}

template <typename Real> std::array<std::vector<double>, 2> TIBDExchangeMove<Real>::move_host(int N, const double *h_x, const double *h_box) {
    const double box_vol = h_box[0] * h_box[4] * h_box[8];
    if (box_vol <= inner_volume_) {
        throw std::runtime_error("volume of inner radius greater than box volume");
    }
    return Mover::move_host(N, h_x, h_box);
}

template <typename Real> double TIBDExchangeMove<Real>::raw_log_probability_host() {
    Real src_max, src_sum, dest_max, dest_sum, box_vol;
    int targeting, state[ST_WORDS];
    TMB_CUDA(cudaMemcpy(&src_max, d_lse_src_max_.data, sizeof(Real), cudaMemcpyDeviceToHost));
    TMB_CUDA(cudaMemcpy(&src_sum, d_lse_src_sum_.data, sizeof(Real), cudaMemcpyDeviceToHost));
    TMB_CUDA(cudaMemcpy(&dest_max, this->d_lse_after_max_.data, sizeof(Real), cudaMemcpyDeviceToHost));
    TMB_CUDA(cudaMemcpy(&dest_sum, this->d_lse_after_sum_.data, sizeof(Real), cudaMemcpyDeviceToHost));
    TMB_CUDA(cudaMemcpy(&targeting, d_targeting_.data, sizeof(int), cudaMemcpyDeviceToHost));
    d_box_volume_.copy_to(&box_vol);
    this->d_state_.copy_to(state);
    const Real outer_vol = box_vol - inner_volume_;
    return static_cast<double>(ti_raw_log_probability<Real>(
        targeting, inner_volume_, outer_vol, state[ST_INNER_USED], this->num_target_mols_, src_max, src_sum, dest_max, dest_sum));
}

template <typename Real>
std::array<std::vector<int>, 2> inner_and_outer_mols(
    const std::vector<int> &center_atoms, int N, const double *coords, const double *box, const std::vector<std::vector<int>> &group_idxs,
    Real radius) {
    const int M = static_cast<int>(group_idxs.size());
    std::array<std::vector<int>, 2> out;
    if (M == 0) {
        return out;
    }
    const MolLayout l = flatten_mols(group_idxs);
    std::vector<int> first(M), sizes(M);
    for (int m = 0; m < M; m++) {
        sizes[m] = l.mol_offsets[m + 1] - l.mol_offsets[m];
        first[m] = sizes[m] > 0 ? l.atom_idxs[l.mol_offsets[m]] : 0;
    }
    DeviceBuffer<double> d_coords(static_cast<size_t>(N) * 3), d_box(9);
    DeviceBuffer<int> d_center_atoms(center_atoms.size()), d_first(M), d_sizes(M), d_flags(M);
    DeviceBuffer<Real> d_center(3);
    d_coords.copy_from(coords);
    d_box.copy_from(box);
    d_center_atoms.copy_from(center_atoms.data());
    d_first.copy_from(first.data());
    d_sizes.copy_from(sizes.data());
    TIDevice<Real> t;
    std::memset(&t, 0, sizeof(t));
    t.inner_flags = d_flags.data;
    cudaStream_t stream = main_stream();
    TMB_LAUNCH(
        k_ti_setup<Real>, 1, EX_THREADS, 0, stream, static_cast<int>(center_atoms.size()), d_center_atoms.data, M, 0, 0, d_first.data,
        d_sizes.data, d_coords.data, d_box.data, radius * radius, d_center.data, static_cast<Real *>(nullptr), t);
    TMB_CUDA(cudaStreamSynchronize(stream));
    std::vector<int> flags(M);
    d_flags.copy_to(flags.data());
    for (int m = 0; m < M; m++) {
        out[flags[m] == 1 ? 0 : 1].push_back(m);
    }
    return out;
}

template <typename Real>
std::vector<Real> translations_inside_and_outside_sphere_host(int n_translations, const double *box, const Real *center, Real radius, int seed) {
    std::vector<Real> out(static_cast<size_t>(n_translations) * 6);
    if (n_translations <= 0) {
        return out;
    }
    DeviceBuffer<double> d_box(9);
    DeviceBuffer<Real> d_center(3), d_out(out.size());
    DeviceBuffer<curandState_t> d_states(TI_STATE_THREADS);
    d_box.copy_from(box);
    d_center.copy_from(center);
    cudaStream_t stream = main_stream();
    TMB_LAUNCH(k_ti_init_states, 1, TI_STATE_THREADS, 0, stream, TI_STATE_THREADS, seed, d_states.data);
    TMB_LAUNCH(k_ti_translations<Real>, 1, TI_STATE_THREADS, 0, stream, n_translations, d_box.data, d_center.data, radius, d_states.data, d_out.data);
    TMB_CUDA(cudaStreamSynchronize(stream));
    d_out.copy_to(out.data());
    return out;
}

template class NonbondedMolEnergyPotential<float>;
template class NonbondedMolEnergyPotential<double>;
template class SegmentedSumExp<float>;
template class SegmentedSumExp<double>;
template class SegmentedWeightedRandomSampler<float>;
template class SegmentedWeightedRandomSampler<double>;
template class BDExchangeMove<float>;
template class BDExchangeMove<double>;
template class TIBDExchangeMove<float>;
template class TIBDExchangeMove<double>;
template std::array<std::vector<int>, 2> inner_and_outer_mols<float>(const std::vector<int> &, int, const double *, const double *, const std::vector<std::vector<int>> &, float);
template std::array<std::vector<int>, 2> inner_and_outer_mols<double>(const std::vector<int> &, int, const double *, const double *, const std::vector<std::vector<int>> &, double);
template std::vector<float> translations_inside_and_outside_sphere_host<float>(int, const double *, const float *, float, int);
template std::vector<double> translations_inside_and_outside_sphere_host<double>(int, const double *, const double *, double, int);
template std::vector<float> atom_by_atom_energies<float>(int, const std::vector<int> &, const double *, const double *, const double *, float, float);
template std::vector<double> atom_by_atom_energies<double>(int, const std::vector<int> &, const double *, const double *, const double *, double, double);
template void rotate_coordinates_host<float>(int, int, const double *, const float *, double *);
template void rotate_coordinates_host<double>(int, int, const double *, const double *, double *);
template void rotate_and_translate_mol_host<float>(int, int, const double *, const double *, const float *, const float *, double *);
template void rotate_and_translate_mol_host<double>(int, int, const double *, const double *, const double *, const double *, double *);

} // namespace tmb
