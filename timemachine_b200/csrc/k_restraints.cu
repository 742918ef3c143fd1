// Restraints and the precomputed nonbonded pair list: the potentials that complete the reference's HostGuestSystem
// (fe/system.py:133-143) next to the hot path.  One thread per term, fixed-point atomics,
// int128 energy reduced in-kernel (no per-term energy buffer + CUB pass as in the reference).
//
// Functional forms: flat-bottom bond k_flat_bottom_bond.cuh:7-22,75-171; chiral volumes chiral_utils.cuh:94-181 and
// restraints k_chiral_restraint.cuh:7-181; precomputed pairs k_nonbonded_precomputed.cuh:11-184 with the f64
// electrostatics of k_nonbonded_common.cuh:16-94 for double and the shared f32 forms of nb_math.cuh for float.
#include "fixed_point.cuh"
#include "kernels.hpp"
#include "nb_math.cuh"
#include "reduce.cuh"

namespace tmb {

constexpr int RS_THREADS = 128;

template <typename Real> struct Vec3 {
    Real x, y, z;
    __device__ Vec3 operator+(const Vec3 &o) const { return {x + o.x, y + o.y, z + o.z}; }
    __device__ Vec3 operator-(const Vec3 &o) const { return {x - o.x, y - o.y, z - o.z}; }
    __device__ Vec3 operator-() const { return {-x, -y, -z}; }
    __device__ Real dot(const Vec3 &o) const { return x * o.x + y * o.y + z * o.z; }
    __device__ Real norm() const { return sqrt(x * x + y * y + z * z); }
    __device__ Vec3 unit() const {
        const Real n = norm();
        return {x / n, y / n, z / n};
    }
};
template <typename Real> struct Mat3 { // row r, column c: m[r][c]
    Real m[3][3];
};
template <typename Real> __device__ __forceinline__ Vec3<Real> load_vec(const double *__restrict__ x, int idx) {
    return {static_cast<Real>(x[idx * 3 + 0]), static_cast<Real>(x[idx * 3 + 1]), static_cast<Real>(x[idx * 3 + 2])};
}
template <typename Real> __device__ __forceinline__ Vec3<Real> cross(const Vec3<Real> &a, const Vec3<Real> &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// row vector times matrix
template <typename Real> __device__ __forceinline__ Vec3<Real> vec_mat(const Vec3<Real> &v, const Mat3<Real> &a) {
    return {
        v.x * a.m[0][0] + v.y * a.m[1][0] + v.z * a.m[2][0], v.x * a.m[0][1] + v.y * a.m[1][1] + v.z * a.m[2][1],
        v.x * a.m[0][2] + v.y * a.m[1][2] + v.z * a.m[2][2]};
}
// d(v / |v|) / dv = (I - u u^T) / |v|
template <typename Real> __device__ __forceinline__ Mat3<Real> unit_jacobian(const Vec3<Real> &v) {
    const Real n = v.norm();
    const Vec3<Real> u = v.unit();
    const Real c[3] = {u.x, u.y, u.z};
    Mat3<Real> j;
    for (int r = 0; r < 3; r++) {
        for (int k = 0; k < 3; k++) {
            j.m[r][k] = ((r == k ? static_cast<Real>(1) : static_cast<Real>(0)) - c[r] * c[k]) / n;
        }
    }
    return j;
}
// d(a x b)/da and d(a x b)/db, laid out so that (row vector g) * J is the pull-back of g
template <typename Real>
__device__ __forceinline__ void cross_jacobians(const Vec3<Real> &a, const Vec3<Real> &b, Mat3<Real> &ja, Mat3<Real> &jb) {
    const Real z = 0;
    const Mat3<Real> A = {{{z, b.z, -b.y}, {-b.z, z, b.x}, {b.y, -b.x, z}}};
    const Mat3<Real> B = {{{z, -a.z, a.y}, {a.z, z, -a.x}, {-a.y, a.x, z}}};
    // row r holds d(a x b)_r / d(a or b) (chiral_utils.cuh:69-81), so that vec_mat(g, J) pulls g back
    ja = A;
    jb = B;
}

__device__ __forceinline__ void add3(u64 *__restrict__ du_dx, int atom, u64 fx, u64 fy, u64 fz) {
    atomicAdd(du_dx + atom * 3 + 0, fx);
    atomicAdd(du_dx + atom * 3 + 1, fy);
    atomicAdd(du_dx + atom * 3 + 2, fz);
}
template <typename Real> __device__ __forceinline__ void add_grad(u64 *du_dx, int atom, const Vec3<Real> &g, Real prefactor) {
    add3(du_dx, atom, to_fixed_force(g.x * prefactor), to_fixed_force(g.y * prefactor), to_fixed_force(g.z * prefactor));
}

// ---- flat-bottom bond: u = k/4 (r - rmax)^4 for r > rmax, k/4 (r - rmin)^4 for r < rmin, periodic ------------------
template <typename Real> __global__ void __launch_bounds__(RS_THREADS) k_flat_bottom_bond(const RestraintArgs a) {
    __shared__ i128 scratch[RS_THREADS / WARP];
    i128 energy = 0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < a.b.n_terms; b += gridDim.x * blockDim.x) {
        const int src = a.b.idxs[b * 2 + 0], dst = a.b.idxs[b * 2 + 1];
        const Real k = static_cast<Real>(a.b.p[b * 3 + 0]);
        const Real rmin = static_cast<Real>(a.b.p[b * 3 + 1]);
        const Real rmax = static_cast<Real>(a.b.p[b * 3 + 2]);
        Real dx[3];
        Real r2 = 0;
        for (int d = 0; d < 3; d++) {
            // the displacement and its minimum image are formed in double, as the reference does (:108-113)
            double delta = a.b.x[src * 3 + d] - a.b.x[dst * 3 + d];
            const double bd = a.box[d * 3 + d];
            // the two fused multiply-adds nvcc makes of the reference's `delta -= box * n` and `r2 += delta * delta`
            delta = fma(-bd, nearbyint(delta / bd), delta);
            dx[d] = static_cast<Real>(delta);
            r2 = static_cast<Real>(fma(delta, delta, static_cast<double>(r2)));
        }
        const Real r = sqrt(r2);
        const Real above = static_cast<Real>(r > rmax), below = static_cast<Real>(r < rmin);
        const Real d_min = r - rmin, d_max = r - rmax;
        if (a.b.d_u != nullptr) {
            const Real d_min2 = d_min * d_min, d_max2 = d_max * d_max;
            energy += energy_to_fixed<Real>((k / 4) * ((below * (d_min2 * d_min2)) + (above * (d_max2 * d_max2))));
        }
        const Real d_min3 = d_min * d_min * d_min, d_max3 = d_max * d_max * d_max;
        if (a.b.du_dp != nullptr) {
            atomicAdd(a.b.du_dp + b * 3 + 0, to_fixed_force(above * ((d_max3 * d_max) / 4) + below * ((d_min3 * d_min) / 4)));
            atomicAdd(a.b.du_dp + b * 3 + 1, to_fixed_force(below * (-k * d_min3)));
            atomicAdd(a.b.du_dp + b * 3 + 2, to_fixed_force(above * (-k * d_max3)));
        }
        if (a.b.du_dx != nullptr) {
            const Real du_dr = k * ((above * d_max3) + (below * d_min3));
            const Real inv_r = 1 / r;
            for (int d = 0; d < 3; d++) {
                const Real g = du_dr * dx[d] * inv_r;
                atomicAdd(a.b.du_dx + src * 3 + d, to_fixed_force(g));
                atomicAdd(a.b.du_dx + dst * 3 + d, to_fixed_force(-g));
            }
        }
    }
    if (a.b.d_u != nullptr) {
        grid_finish_energy(energy, scratch, a.b.u_partials, a.b.ticket, a.b.d_u);
    }
}

// ---- log flat-bottom bond: u = -log(1 - exp(-beta u_fb)) / beta (reference k_log_flat_bottom_bond.cuh:7-117) ---------------
// The restraint local MD puts on the FROZEN shell when its reference atom moves: it diverges where the flat-bottom energy
// vanishes.  log(1 - exp(-x)) is evaluated as log(-expm1(-x)) below log 2 and log1p(-exp(-x)) above, like the reference.
template <typename Real> __device__ __forceinline__ Real log_1_exp_neg(Real x) {
    return x < static_cast<Real>(0.693147180559945309417232121) ? log(-expm1(-x)) : log1p(-exp(-x));
}

template <typename Real> __global__ void __launch_bounds__(RS_THREADS) k_log_flat_bottom_bond(const RestraintArgs a) {
    __shared__ i128 scratch[RS_THREADS / WARP];
    i128 energy = 0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < a.b.n_terms; b += gridDim.x * blockDim.x) {
        const int src = a.b.idxs[b * 2 + 0], dst = a.b.idxs[b * 2 + 1];
        const Real k = static_cast<Real>(a.b.p[b * 3 + 0]);
        const Real rmin = static_cast<Real>(a.b.p[b * 3 + 1]);
        const Real rmax = static_cast<Real>(a.b.p[b * 3 + 2]);
        Real dx[3];
        Real r2 = 0;
        for (int d = 0; d < 3; d++) {
            double delta = a.b.x[src * 3 + d] - a.b.x[dst * 3 + d];
            const double bd = a.box[d * 3 + d];
            // the two fused multiply-adds nvcc makes of the reference's `delta -= box * n` and `r2 += delta * delta`
            delta = fma(-bd, nearbyint(delta / bd), delta);
            dx[d] = static_cast<Real>(delta);
            r2 = static_cast<Real>(fma(delta, delta, static_cast<double>(r2)));
        }
        const Real r = sqrt(r2);
        const Real above = static_cast<Real>(r > rmax), below = static_cast<Real>(r < rmin);
        const Real d_min = r - rmin, d_max = r - rmax;
        const Real d_min2 = d_min * d_min, d_max2 = d_max * d_max;
        const Real nrg = (k / 4) * ((below * (d_min2 * d_min2)) + (above * (d_max2 * d_max2)));
        // beta is a double in the reference's kernel (k_log_flat_bottom_bond.cuh:21,64,69): both transcendental expressions are
        // evaluated in f64 whatever Real is, and rounded to Real afterwards
        if (a.b.d_u != nullptr) {
            energy += energy_to_fixed<Real>(static_cast<Real>(-log_1_exp_neg<double>(a.beta * static_cast<double>(nrg)) / a.beta));
        }
        // d/d(nrg) of the log form: -exp(-beta nrg) / (1 - exp(-beta nrg)); every flat-bottom gradient is scaled by it
        Real pre = static_cast<Real>(-exp(-a.beta * static_cast<double>(nrg)));
        pre = pre / (static_cast<Real>(1) + pre);
        const Real d_min3 = d_min2 * d_min, d_max3 = d_max2 * d_max;
        if (a.b.du_dp != nullptr) {
            atomicAdd(a.b.du_dp + b * 3 + 0, to_fixed_force((above * ((d_max3 * d_max) / 4) + below * ((d_min3 * d_min) / 4)) * pre));
            atomicAdd(a.b.du_dp + b * 3 + 1, to_fixed_force((below * (-k * d_min3)) * pre));
            atomicAdd(a.b.du_dp + b * 3 + 2, to_fixed_force((above * (-k * d_max3)) * pre));
        }
        if (a.b.du_dx != nullptr) {
            const Real du_dr = k * ((above * d_max3) + (below * d_min3));
            const Real inv_r = 1 / r;
            for (int d = 0; d < 3; d++) {
                const Real g = du_dr * dx[d] * inv_r;
                atomicAdd(a.b.du_dx + src * 3 + d, to_fixed_force(pre * g));
                atomicAdd(a.b.du_dx + dst * 3 + d, to_fixed_force(pre * (-g)));
            }
        }
    }
    if (a.b.d_u != nullptr) {
        grid_finish_energy(energy, scratch, a.b.u_partials, a.b.ticket, a.b.d_u);
    }
}

// ---- chiral atom restraint: vol = (x^ x y^) . z^ around a centre; u = k vol^2 where vol > 0 ------------------------
template <typename Real> __global__ void __launch_bounds__(RS_THREADS) k_chiral_atom_restraint(const RestraintArgs a) {
    __shared__ i128 scratch[RS_THREADS / WARP];
    i128 energy = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < a.b.n_terms; t += gridDim.x * blockDim.x) {
        const int ic = a.b.idxs[t * 4 + 0], i1 = a.b.idxs[t * 4 + 1], i2 = a.b.idxs[t * 4 + 2], i3 = a.b.idxs[t * 4 + 3];
        const Vec3<Real> xc = load_vec<Real>(a.b.x, ic);
        const Vec3<Real> xx = load_vec<Real>(a.b.x, i1) - xc, yy = load_vec<Real>(a.b.x, i2) - xc, zz = load_vec<Real>(a.b.x, i3) - xc;
        const Vec3<Real> x = xx.unit(), y = yy.unit(), z = zz.unit();
        const Vec3<Real> xy = cross(x, y);
        Mat3<Real> dc_dx, dc_dy;
        cross_jacobians(x, y, dc_dx, dc_dy);
        const Vec3<Real> g1 = vec_mat(vec_mat(z, dc_dx), unit_jacobian(xx));
        const Vec3<Real> g2 = vec_mat(vec_mat(z, dc_dy), unit_jacobian(yy));
        const Vec3<Real> g3 = vec_mat(xy, unit_jacobian(zz));
        const Real vol = xy.dot(z);
        const Real k = static_cast<Real>(a.b.p[t]);
        if (a.b.d_u != nullptr && vol > 0) {
            energy += static_cast<i128>(static_cast<i64>(to_fixed_force(k * vol * vol)));
        }
        if (k == 0 || vol <= 0) {
            continue;
        }
        if (a.b.du_dx != nullptr) {
            const Real pre = 2 * k * vol;
            add_grad(a.b.du_dx, ic, -g1 - g2 - g3, pre);
            add_grad(a.b.du_dx, i1, g1, pre);
            add_grad(a.b.du_dx, i2, g2, pre);
            add_grad(a.b.du_dx, i3, g3, pre);
        }
        if (a.b.du_dp != nullptr) {
            atomicAdd(a.b.du_dp + t, to_fixed_force(vol * vol));
        }
    }
    if (a.b.d_u != nullptr) {
        grid_finish_energy(energy, scratch, a.b.u_partials, a.b.ticket, a.b.d_u);
    }
}

// ---- chiral bond restraint: vol = (x^ x y^) . (y^ x z^) along a torsion; u = k vol^2 where sign * vol > 0 ----------
template <typename Real> __global__ void __launch_bounds__(RS_THREADS) k_chiral_bond_restraint(const RestraintArgs a) {
    __shared__ i128 scratch[RS_THREADS / WARP];
    i128 energy = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < a.b.n_terms; t += gridDim.x * blockDim.x) {
        const int i0 = a.b.idxs[t * 4 + 0], i1 = a.b.idxs[t * 4 + 1], i2 = a.b.idxs[t * 4 + 2], i3 = a.b.idxs[t * 4 + 3];
        const Vec3<Real> p0 = load_vec<Real>(a.b.x, i0), p1 = load_vec<Real>(a.b.x, i1), p2 = load_vec<Real>(a.b.x, i2),
                         p3 = load_vec<Real>(a.b.x, i3);
        const Vec3<Real> xx = p1 - p0, yy = p1 - p2, zz = p3 - p2;
        const Vec3<Real> x = xx.unit(), y = yy.unit(), z = zz.unit();
        const Vec3<Real> xy = cross(x, y), yz = cross(y, z);
        Mat3<Real> dc0_dx, dc0_dy, dc1_dy, dc1_dz;
        cross_jacobians(x, y, dc0_dx, dc0_dy);
        cross_jacobians(y, z, dc1_dy, dc1_dz);
        const Mat3<Real> jy = unit_jacobian(yy);
        const Vec3<Real> gx = vec_mat(vec_mat(yz, dc0_dx), unit_jacobian(xx));
        const Vec3<Real> gy = vec_mat(vec_mat(yz, dc0_dy), jy) + vec_mat(vec_mat(xy, dc1_dy), jy);
        const Vec3<Real> gz = vec_mat(vec_mat(xy, dc1_dz), unit_jacobian(zz));
        const Real vol = xy.dot(yz);
        const Real k = static_cast<Real>(a.b.p[t]);
        const int sign = a.signs[t];
        if (a.b.d_u != nullptr && sign * vol > 0) {
            energy += static_cast<i128>(static_cast<i64>(to_fixed_force(k * vol * vol)));
        }
        if (k == 0 || sign * vol <= 0) {
            continue;
        }
        if (a.b.du_dx != nullptr) {
            const Real pre = 2 * k * vol;
            add_grad(a.b.du_dx, i0, -gx, pre);
            add_grad(a.b.du_dx, i1, gx + gy, pre);
            add_grad(a.b.du_dx, i2, -gy - gz, pre);
            add_grad(a.b.du_dx, i3, gz, pre);
        }
        if (a.b.du_dp != nullptr) {
            atomicAdd(a.b.du_dp + t, to_fixed_force(vol * vol));
        }
    }
    if (a.b.d_u != nullptr) {
        grid_finish_energy(energy, scratch, a.b.u_partials, a.b.ticket, a.b.d_u);
    }
}

// ---- nonbonded on precomputed pairs: params [M,4] = (q_ij, sig_ij, eps_ij, w offset) -------------------------------
// f32: the rounded-operation sequence of the compiled reference (read off the SASS of k_nonbonded_precomputed<float> under
// oracle/_ref, as for the tile kernel): d^2 = fma(dw,dw, fma(dz,dz, fma(dx,dx, dy*dy))), IEEE sqrt and reciprocal,
// damping' = fma(erfc, S', -(2/sqrt(pi) e^(-b^2d^2) b) S), es_factor = fma(damping, -(1/d)^2, damping' (1/d)), every
// force term converted to fixed point for atom i and, from the negated float, for atom j.
template <typename Real> __global__ void __launch_bounds__(RS_THREADS) k_nonbonded_precomputed(const RestraintArgs a) {
    __shared__ i128 scratch[RS_THREADS / WARP];
    i128 energy = 0;
    const BoxCache<Real> box = load_box<Real>(a.box);
    const Real beta = static_cast<Real>(a.beta);
    const Real cutoff2 = static_cast<Real>(a.cutoff * a.cutoff);
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < a.b.n_terms; m += gridDim.x * blockDim.x) {
        const int i = a.b.idxs[m * 2 + 0], j = a.b.idxs[m * 2 + 1];
        const Real q = static_cast<Real>(a.b.p[m * 4 + P_CHARGE]);
        const Real sig = static_cast<Real>(a.b.p[m * 4 + P_SIG]);
        const Real eps = static_cast<Real>(a.b.p[m * 4 + P_EPS]);
        const Real dw = static_cast<Real>(a.b.p[m * 4 + P_W]);
        const Vec3<Real> xi = load_vec<Real>(a.b.x, i), xj = load_vec<Real>(a.b.x, j);
        const Real dx = min_image(xi.x - xj.x, box.x, box.inv_x);
        const Real dy = min_image(xi.y - xj.y, box.y, box.inv_y);
        const Real dz = min_image(xi.z - xj.z, box.z, box.inv_z);
        const Real d2 = fma_(dw, dw, dist2_3d(dx, dy, dz));
        if (!(d2 < cutoff2)) {
            continue;
        }
        const Real d = sqrt(d2);
        const Real inv_d = 1 / d;
        // each term is rounded to fixed point on its own, as in the reference (:95-105,150-158)
        u64 fix = 0, fiy = 0, fiz = 0, fjx = 0, fjy = 0, fjz = 0, g_q = 0, g_sig = 0, g_eps = 0, g_w = 0;
        if (q != 0) {
            Real debd, dsdr;
            const Real ebd = erfc_and_deriv(beta * d, debd);
            debd = beta * debd;
            const Real sr = switch_and_deriv(d, dsdr);
            const Real damping_prime = fma_(ebd, dsdr, debd * sr);
            const Real damping = ebd * sr;
            const Real es_factor = fma_(damping, -(inv_d * inv_d), damping_prime * inv_d);
            if (a.b.d_u != nullptr) {
                energy += energy_to_fixed<Real>(damping * (q * inv_d));
            }
            const Real pre = (q * es_factor) * inv_d;
            fix += to_fixed_force(dx * pre);
            fiy += to_fixed_force(dy * pre);
            fiz += to_fixed_force(dz * pre);
            fjx += to_fixed_force(-dx * pre);
            fjy += to_fixed_force(-dy * pre);
            fjz += to_fixed_force(-dz * pre);
            g_q = to_fixed<FIXED_EXPONENT_DU_DCHARGE>(damping * inv_d);
            g_w += to_fixed<FIXED_EXPONENT_DU_DW>(dw * pre);
        }
        const bool lj = eps != 0 && sig != 0;
        if (lj) {
            const Real s1 = sig * inv_d;
            const Real s2 = s1 * s1;
            const Real s6 = s2 * s2 * s2;
            const Real du_de = static_cast<Real>(4) * (s6 - 1) * s6;
            if (a.b.d_u != nullptr) {
                energy += energy_to_fixed<Real>(eps * du_de);
            }
            const Real d6 = d2 * d2 * d2;
            const Real sig2 = sig * sig;
            const Real sig6 = sig2 * sig2 * sig2;
            const Real d12 = d6 * d6;
            const Real du_dr = eps * static_cast<Real>(24) * sig6 * (d6 - static_cast<Real>(2) * sig6) / (d12 * d);
            const Real pre = du_dr * inv_d;
            fix += to_fixed_force(dx * pre);
            fiy += to_fixed_force(dy * pre);
            fiz += to_fixed_force(dz * pre);
            fjx += to_fixed_force(-dx * pre);
            fjy += to_fixed_force(-dy * pre);
            fjz += to_fixed_force(-dz * pre);
            g_w += to_fixed<FIXED_EXPONENT_DU_DW>(dw * pre);
            g_eps = to_fixed<FIXED_EXPONENT_DU_DEPS>(du_de);
            g_sig = to_fixed<FIXED_EXPONENT_DU_DSIG>(
                static_cast<Real>(-24) * eps * (sig2 * sig2 * sig) * (d6 - static_cast<Real>(2) * sig6) / d12);
        }
        // The compiled reference adds a pair's gradients from inside its Lennard-Jones branch
        // (k_nonbonded_precomputed.cuh:150-181): a pair with eps_ij == 0 (or sig_ij == 0) contributes its electrostatic
        // ENERGY there but neither force nor du/dp.  A drop-in returns what the reference returns, so that is the default;
        // TMB_PRECOMPUTED_FULL_GRADIENT=1 keeps the electrostatic gradient of such pairs, as the reference's Python
        // potential (potentials/nonbonded.py:403-446) does.
        if (!lj && !a.full_gradient) {
            continue;
        }
        if (a.b.du_dx != nullptr) {
            add3(a.b.du_dx, i, fix, fiy, fiz);
            add3(a.b.du_dx, j, fjx, fjy, fjz);
        }
        if (a.b.du_dp != nullptr) {
            atomicAdd(a.b.du_dp + m * 4 + P_CHARGE, g_q);
            atomicAdd(a.b.du_dp + m * 4 + P_SIG, g_sig);
            atomicAdd(a.b.du_dp + m * 4 + P_EPS, g_eps);
            atomicAdd(a.b.du_dp + m * 4 + P_W, g_w);
        }
    }
    if (a.b.d_u != nullptr) {
        grid_finish_energy(energy, scratch, a.b.u_partials, a.b.ticket, a.b.d_u);
    }
}

// ---- centroid restraint (reference k_centroid_restraint.cuh:7-84) ---------------------------------------------------
// The reference clears two accumulators, sums the fixed-point coordinates of both groups with global atomics in one
// launch and evaluates in a second.  The groups are a ligand and a pocket (tens to hundreds of atoms), so here ONE CTA
// does all of it: fixed-point centroid sums in shared memory (integer, hence the same whatever the order), a barrier,
// then every thread evaluates the atoms it owns with the reference's operation sequence (mixed f32 / f64 where the
// reference mixes `RealType` with its double kb / b0; the accumulation of |delta|^2 is the FMA chain nvcc emits).
constexpr int CR_THREADS = 256;
template <typename Real> __global__ void __launch_bounds__(CR_THREADS) k_centroid_restraint(const CentroidArgs a) {
    __shared__ unsigned long long s_sum[6];
    if (threadIdx.x < 6) {
        s_sum[threadIdx.x] = 0;
    }
    __syncthreads();
    const int n = a.n_a + a.n_b;
    for (int t = threadIdx.x; t < n; t += CR_THREADS) {
        const bool in_a = t < a.n_a;
        const int atom = in_a ? a.group_a[t] : a.group_b[t - a.n_a];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            atomicAdd(s_sum + (in_a ? 0 : 3) + d, to_fixed_force<Real>(static_cast<Real>(a.x[atom * 3 + d])));
        }
    }
    __syncthreads();
    Real deltas[3];
    Real dij = 0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        deltas[d] = fixed_to_real<Real>(s_sum[d]) / static_cast<Real>(a.n_a) - fixed_to_real<Real>(s_sum[3 + d]) / static_cast<Real>(a.n_b);
        dij = fma(deltas[d], deltas[d], dij);
    }
    dij = sqrt(dij);
    if (threadIdx.x == 0 && a.d_u != nullptr) {
        const Real nrg = static_cast<Real>(a.kb * (static_cast<double>(dij) - a.b0) * (static_cast<double>(dij) - a.b0));
        *a.d_u = energy_to_fixed<Real>(nrg);
    }
    if (a.du_dx == nullptr) {
        return;
    }
    for (int t = threadIdx.x; t < n; t += CR_THREADS) {
        const bool in_a = t < a.n_a;
        const int atom = in_a ? a.group_a[t] : a.group_b[t - a.n_a];
        const Real count = static_cast<Real>(in_a ? a.n_a : a.n_b);
        const Real sign = in_a ? static_cast<Real>(1) : static_cast<Real>(-1);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            Real delta;
            if (a.b0 != 0) {
                const Real du_ddij = static_cast<Real>(2 * a.kb * (static_cast<double>(dij) - a.b0));
                const Real ddij_dxi = deltas[d] / dij;
                delta = sign * du_ddij * ddij_dxi / count;
            } else {
                delta = static_cast<Real>(static_cast<double>(sign * 2) * a.kb * static_cast<double>(deltas[d]) / static_cast<double>(count));
            }
            atomicAdd(a.du_dx + atom * 3 + d, to_fixed_force<Real>(delta));
        }
    }
}

template <typename Real> void launch_centroid_restraint(const CentroidArgs &args, cudaStream_t stream) {
    TMB_LAUNCH(k_centroid_restraint<Real>, 1, CR_THREADS, 0, stream, args);
}
template void launch_centroid_restraint<float>(const CentroidArgs &, cudaStream_t);
template void launch_centroid_restraint<double>(const CentroidArgs &, cudaStream_t);

#define TMB_RESTRAINT_LAUNCHER(name, kernel)                                                                          \
    template <typename Real> void name(const RestraintArgs &args, cudaStream_t stream) {                              \
        TMB_LAUNCH(kernel<Real>, bonded_grid(args.b.n_terms), RS_THREADS, 0, stream, args);                           \
    }                                                                                                                  \
    template void name<float>(const RestraintArgs &, cudaStream_t);                                                   \
    template void name<double>(const RestraintArgs &, cudaStream_t);

TMB_RESTRAINT_LAUNCHER(launch_flat_bottom_bond, k_flat_bottom_bond)
TMB_RESTRAINT_LAUNCHER(launch_log_flat_bottom_bond, k_log_flat_bottom_bond)
TMB_RESTRAINT_LAUNCHER(launch_chiral_atom_restraint, k_chiral_atom_restraint)
TMB_RESTRAINT_LAUNCHER(launch_chiral_bond_restraint, k_chiral_bond_restraint)
TMB_RESTRAINT_LAUNCHER(launch_nonbonded_precomputed, k_nonbonded_precomputed)

} // namespace tmb
