// Bonded terms: harmonic bond, harmonic angle (4-D regularised, Kahan-stable), periodic torsion.
// One thread per term, fixed-point atomics for du/dx and du/dp, int128 energy reduced in-kernel.
// Functional forms follow reference k_harmonic_bond.cuh:7-60, k_harmonic_angle.cuh:10-146, k_periodic_torsion.cuh:20-131.
// Compiled with --fmad=false: products and sums are individually rounded, which is what gives the reference's
// bitwise i<->k (angle) and cross-product anti-commutativity (torsion) symmetries without special intrinsics.
#include "fixed_point.cuh"
#include "kernels.hpp"
#include "reduce.cuh"

namespace tmb {

constexpr int BD_THREADS = 128;

int bonded_grid(int n_terms) {
    int g = ceil_div(n_terms, BD_THREADS);
    const int cap = sm_count() * 8;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

template <typename Real> struct V3 {
    Real x, y, z;
};

// x[a] - x[b]: the difference is formed in double and THEN rounded to Real, as the reference's bonded kernels do
// (`RealType delta = coords[src * 3 + d] - coords[dst * 3 + d]`, k_harmonic_bond.cuh:26, k_harmonic_angle.cuh:43-44,
// k_periodic_torsion.cuh:48-50).  Rounding each coordinate first would cost ~ulp(|x|) ~ 1e-6 nm at 8 nm from the origin
// (coordinates are not wrapped into the box) - 100x the reference's error on a stiff bond - and translation invariance.
template <typename Real> __device__ __forceinline__ V3<Real> delta3(const double *__restrict__ x, int a, int b) {
    V3<Real> r;
    r.x = static_cast<Real>(x[a * 3 + 0] - x[b * 3 + 0]);
    r.y = static_cast<Real>(x[a * 3 + 1] - x[b * 3 + 1]);
    r.z = static_cast<Real>(x[a * 3 + 2] - x[b * 3 + 2]);
    return r;
}
// a.b as nvcc contracts the reference's `a[0]*b[0] + a[1]*b[1] + a[2]*b[2]` (the left product of each sum is fused)
template <typename Real> __device__ __forceinline__ Real dot3(const V3<Real> &a, const V3<Real> &b) {
    return fma(a.z, b.z, fma(a.x, b.x, a.y * b.y));
}
template <typename Real> __device__ __forceinline__ V3<Real> cross3(const V3<Real> &a, const V3<Real> &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ void add_force(u64 *__restrict__ du_dx, int atom, u64 fx, u64 fy, u64 fz) {
    atomicAdd(du_dx + atom * 3 + 0, fx);
    atomicAdd(du_dx + atom * 3 + 1, fy);
    atomicAdd(du_dx + atom * 3 + 2, fz);
}

// ---- u = k/2 (|x_i - x_j| - b0)^2 -----------------------------------------------------------------------------
template <typename Real> __global__ void __launch_bounds__(BD_THREADS) k_harmonic_bond(const BondedArgs a) {
    __shared__ i128 scratch[BD_THREADS / WARP];
    i128 energy = 0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < a.n_terms; b += gridDim.x * blockDim.x) {
        const int src = a.idxs[b * 2 + 0];
        const int dst = a.idxs[b * 2 + 1];
        const V3<Real> d = delta3<Real>(a.x, src, dst);
        const Real kb = static_cast<Real>(a.p[b * 2 + 0]);
        const Real b0 = static_cast<Real>(a.p[b * 2 + 1]);
        // the reference's `d2ij += delta * delta` loop is contracted by nvcc into this chain (default -fmad=true there, off
        // here): with it the bond term is the same sequence of rounded operations as the reference's
        const Real r = sqrt(fma(d.z, d.z, fma(d.y, d.y, d.x * d.x)));
        const Real db = r - b0;
        if (a.du_dx != nullptr) {
            const Real inv_r = 1 / r;
            // b0 == 0 is the "spring to a point" case: the gradient is k * delta without the 0/0
            const Real gx = b0 != 0 ? kb * db * d.x * inv_r : kb * d.x;
            const Real gy = b0 != 0 ? kb * db * d.y * inv_r : kb * d.y;
            const Real gz = b0 != 0 ? kb * db * d.z * inv_r : kb * d.z;
            add_force(a.du_dx, src, to_fixed_force(gx), to_fixed_force(gy), to_fixed_force(gz));
            add_force(a.du_dx, dst, to_fixed_force(-gx), to_fixed_force(-gy), to_fixed_force(-gz));
        }
        if (a.du_dp != nullptr) {
            // du/dk is formed in double even for the f32 kernel (reference k_harmonic_bond.cuh:52: `0.5 * db * db`)
            atomicAdd(a.du_dp + b * 2 + 0, to_fixed_force(0.5 * static_cast<double>(db) * static_cast<double>(db)));
            atomicAdd(a.du_dp + b * 2 + 1, to_fixed_force(-kb * db));
        }
        if (a.d_u != nullptr) {
            energy += energy_to_fixed<Real>(kb / 2 * db * db);
        }
    }
    if (a.d_u != nullptr) {
        grid_finish_energy(energy, scratch, a.u_partials, a.ticket, a.d_u);
    }
}

// ---- u = k/2 (theta - a0)^2, theta from the stable 2 atan2(|n_jk r_ji - n_ji r_jk|, |n_jk r_ji + n_ji r_jk|) ---
// The vectors carry a 4th component eps (a parameter) so that collinear geometries stay differentiable.
template <typename Real> __global__ void __launch_bounds__(BD_THREADS) k_harmonic_angle(const BondedArgs a) {
    __shared__ i128 scratch[BD_THREADS / WARP];
    i128 energy = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < a.n_terms; t += gridDim.x * blockDim.x) {
        const int i = a.idxs[t * 3 + 0];
        const int j = a.idxs[t * 3 + 1];
        const int k = a.idxs[t * 3 + 2];
        const Real ka = static_cast<Real>(a.p[t * 3 + 0]);
        const Real a0 = static_cast<Real>(a.p[t * 3 + 1]);
        const Real eps = static_cast<Real>(a.p[t * 3 + 2]);

        const V3<Real> vji = delta3<Real>(a.x, i, j);
        const V3<Real> vjk = delta3<Real>(a.x, k, j);
        const Real rji[4] = {vji.x, vji.y, vji.z, eps};
        const Real rjk[4] = {vjk.x, vjk.y, vjk.z, eps};
        // From here to the gradient every operation is written out with its rounding, in the order nvcc emits for the
        // reference's k_harmonic_angle (default -fmad there; read off the SASS of oracle/_ref/obj/harmonic_angle.o): squares
        // are separate products (they are reused by the dot products), sums of squares are plain adds, every other
        // sum of products is an FMA chain that fuses the LEFT product of each `x*y + z*w`.  With it the f32 angle term is
        // the same sequence of rounded operations as the reference's.
        const Real eps2 = eps * eps;
        const Real sji = ((rji[0] * rji[0] + rji[1] * rji[1]) + rji[2] * rji[2]) + eps2; // = a.a below, bit for bit
        const Real sjk = ((rjk[0] * rjk[0] + rjk[1] * rjk[1]) + rjk[2] * rjk[2]) + eps2; // = b.b
        const Real nji = sqrt(sji);
        const Real njk = sqrt(sjk);

        Real hi = 0, lo = 0;
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const Real p = njk * rji[d];
            const Real q = nji * rjk[d];
            const Real m = p - q;
            const Real s = p + q;
            hi = fma(m, m, hi);
            lo = fma(s, s, lo);
        }
        const Real theta = 2 * atan2(sqrt(hi), sqrt(lo));
        const Real delta = theta - a0;

        // gradient direction via a x (b x c) = b (a.c) - c (a.b), in 4-D
        const Real a_dot_b = eps2 + fma(rji[2], rjk[2], fma(rji[0], rjk[0], rji[1] * rjk[1]));
        const Real a_dot_a = sji;
        const Real b_dot_b = sjk;
        Real aab[4], bba[4];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            aab[d] = fma(rji[d], a_dot_b, -(rjk[d] * a_dot_a));
            bba[d] = fma(rjk[d], a_dot_b, -(rji[d] * b_dot_b));
        }
        const Real eps_adb = eps * a_dot_b;
        aab[3] = fma(eps, -a_dot_a, eps_adb);
        bba[3] = fma(eps, -b_dot_b, eps_adb);
        const Real aab2 = fma(aab[3], aab[3], fma(aab[2], aab[2], fma(aab[0], aab[0], aab[1] * aab[1])));
        const Real bba2 = fma(bba[3], bba[3], fma(bba[2], bba[2], fma(bba[0], bba[0], bba[1] * bba[1])));
        const Real aab_n = sqrt(aab2);
        const Real bba_n = sqrt(bba2);
        const Real pref = ka * delta;
        const Real coeff_i = pref * (1 / nji);
        const Real coeff_k = pref * (1 / njk);

        if (a.du_dx != nullptr) {
            Real gi[3], gk[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                gi[d] = coeff_i * (aab_n == 0 ? static_cast<Real>(0) : aab[d] / aab_n);
                gk[d] = coeff_k * (bba_n == 0 ? static_cast<Real>(0) : bba[d] / bba_n);
            }
            add_force(a.du_dx, i, to_fixed_force(gi[0]), to_fixed_force(gi[1]), to_fixed_force(gi[2]));
            add_force(a.du_dx, k, to_fixed_force(gk[0]), to_fixed_force(gk[1]), to_fixed_force(gk[2]));
            add_force(
                a.du_dx, j, to_fixed_force(-gi[0] - gk[0]), to_fixed_force(-gi[1] - gk[1]), to_fixed_force(-gi[2] - gk[2]));
        }
        if (a.du_dp != nullptr) {
            atomicAdd(a.du_dp + t * 3 + 0, to_fixed_force(delta * delta / 2));
            atomicAdd(a.du_dp + t * 3 + 1, to_fixed_force(-delta * ka));
            const Real e0 = aab_n == 0 ? static_cast<Real>(0) : coeff_i * aab[3] / aab_n;
            const Real e1 = bba_n == 0 ? static_cast<Real>(0) : coeff_k * bba[3] / bba_n;
            atomicAdd(a.du_dp + t * 3 + 2, to_fixed_force(e0 + e1));
        }
        if (a.d_u != nullptr) {
            energy += energy_to_fixed<Real>((ka / 2) * delta * delta);
        }
    }
    if (a.d_u != nullptr) {
        grid_finish_energy(energy, scratch, a.u_partials, a.ticket, a.d_u);
    }
}

// ---- u = k (1 + cos(n phi - phi0)), phi = atan2((n1 x n2).r_kj/|r_kj|, n1.n2) ----------------------------------
template <typename Real> __global__ void __launch_bounds__(BD_THREADS) k_periodic_torsion(const BondedArgs a) {
    __shared__ i128 scratch[BD_THREADS / WARP];
    i128 energy = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < a.n_terms; t += gridDim.x * blockDim.x) {
        const int i = a.idxs[t * 4 + 0];
        const int j = a.idxs[t * 4 + 1];
        const int k = a.idxs[t * 4 + 2];
        const int l = a.idxs[t * 4 + 3];
        const V3<Real> rij = delta3<Real>(a.x, j, i);
        V3<Real> rkj = delta3<Real>(a.x, j, k);
        const V3<Real> rkl = delta3<Real>(a.x, l, k);

        // As in the angle kernel: the operation sequence nvcc emits for the reference's k_periodic_torsion (read off the SASS of
        // oracle/_ref/obj/periodic_torsion.o).  Cross products are products then a subtraction (the reference forces that
        // with rmul_rn for bitwise anticommutativity); every dot product is fma(z, z', fma(x, x', y * y')); |r_kj|^2 is formed
        // twice by the reference's source, in two different associations, and both are kept.
        const Real rkj_yy = rkj.y * rkj.y;
        const Real rkj2 = fma(rkj.z, rkj.z, rkj.x * rkj.x + rkj_yy);
        const Real rkj_n = sqrt(rkj2);
        const V3<Real> n1 = cross3(rij, rkj);
        const V3<Real> n2 = cross3(rkj, rkl);
        const Real n1_2 = dot3(n1, n1);
        const Real n2_2 = dot3(n2, n2);
        const V3<Real> n3 = cross3(n1, n2);
        const Real rij_rkj = dot3(rij, rkj);
        const Real rkl_rkj = dot3(rkl, rkj);

        const Real c0 = rkj_n / n1_2;
        const Real c3 = -rkj_n / n2_2;
        const Real q1m1 = rij_rkj / rkj2 - 1;
        const Real q2m1 = rkl_rkj / rkj2 - 1;
        const Real n1v[3] = {n1.x, n1.y, n1.z};
        const Real n2v[3] = {n2.x, n2.y, n2.z};
        Real dR0[3], dR1[3], dR2[3], dR3[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            dR0[d] = c0 * n1v[d];
            dR3[d] = c3 * n2v[d];
            dR1[d] = fma(q1m1, dR0[d], -((dR3[d] * rkl_rkj) / rkj2));
            dR2[d] = fma(q2m1, dR3[d], -((dR0[d] * rij_rkj) / rkj2));
        }
        const Real rkj_n2 = sqrt(fma(rkj.z, rkj.z, fma(rkj.x, rkj.x, rkj_yy)));
        rkj.x /= rkj_n2;
        rkj.y /= rkj_n2;
        rkj.z /= rkj_n2;
        const Real phi = atan2(dot3(n3, rkj), dot3(n1, n2));

        const Real kt = static_cast<Real>(a.p[t * 3 + 0]);
        const Real phase = static_cast<Real>(a.p[t * 3 + 1]);
        const Real period = static_cast<Real>(a.p[t * 3 + 2]);
        const Real arg = fma(period, phi, -phase);
        const Real pref = kt * sin(arg) * period;

        if (a.du_dx != nullptr) {
            add_force(a.du_dx, i, to_fixed_force(dR0[0] * pref), to_fixed_force(dR0[1] * pref), to_fixed_force(dR0[2] * pref));
            add_force(a.du_dx, j, to_fixed_force(dR1[0] * pref), to_fixed_force(dR1[1] * pref), to_fixed_force(dR1[2] * pref));
            add_force(a.du_dx, k, to_fixed_force(dR2[0] * pref), to_fixed_force(dR2[1] * pref), to_fixed_force(dR2[2] * pref));
            add_force(a.du_dx, l, to_fixed_force(dR3[0] * pref), to_fixed_force(dR3[1] * pref), to_fixed_force(dR3[2] * pref));
        }
        if (a.du_dp != nullptr) {
            atomicAdd(a.du_dp + t * 3 + 0, to_fixed_force(1 + cos(arg)));
            atomicAdd(a.du_dp + t * 3 + 1, to_fixed_force(kt * sin(arg)));
            atomicAdd(a.du_dp + t * 3 + 2, to_fixed_force(-kt * sin(arg) * phi));
        }
        if (a.d_u != nullptr) {
            energy += energy_to_fixed<Real>(kt * (1 + cos(arg)));
        }
    }
    if (a.d_u != nullptr) {
        grid_finish_energy(energy, scratch, a.u_partials, a.ticket, a.d_u);
    }
}

template <typename Real> void launch_harmonic_bond(const BondedArgs &args, cudaStream_t stream) {
    TMB_LAUNCH(k_harmonic_bond<Real>, bonded_grid(args.n_terms), BD_THREADS, 0, stream, args);
}
template <typename Real> void launch_harmonic_angle(const BondedArgs &args, cudaStream_t stream) {
    TMB_LAUNCH(k_harmonic_angle<Real>, bonded_grid(args.n_terms), BD_THREADS, 0, stream, args);
}
template <typename Real> void launch_periodic_torsion(const BondedArgs &args, cudaStream_t stream) {
    TMB_LAUNCH(k_periodic_torsion<Real>, bonded_grid(args.n_terms), BD_THREADS, 0, stream, args);
}
template void launch_harmonic_bond<float>(const BondedArgs &, cudaStream_t);
template void launch_harmonic_bond<double>(const BondedArgs &, cudaStream_t);
template void launch_harmonic_angle<float>(const BondedArgs &, cudaStream_t);
template void launch_harmonic_angle<double>(const BondedArgs &, cudaStream_t);
template void launch_periodic_torsion<float>(const BondedArgs &, cudaStream_t);
template void launch_periodic_torsion<double>(const BondedArgs &, cudaStream_t);

} // namespace tmb
