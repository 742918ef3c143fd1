// Explicit pair-list nonbonded kernel with per-pair (charge, LJ) scales; with `negated` it is the exclusion term that
// must cancel the all-pairs contribution bit-exactly (reference k_nonbonded_pair_list.cuh:19-190).  It calls the same
// pair_terms<> as the tile kernel on operands produced by the same casts, so every fixed-point term is identical.
#include "fixed_point.cuh"
#include "kernels.hpp"
#include "nb_math.cuh"
#include "reduce.cuh"

namespace tmb {

constexpr int PL_THREADS = 128;

template <bool NEG> __device__ __forceinline__ void accum(u64 *addr, u64 v) { atomicAdd(addr, NEG ? (0ull - v) : v); }

template <typename Real, bool NEG> __global__ void __launch_bounds__(PL_THREADS) k_pair_list(const PairListArgs<Real> a) {
    __shared__ i128 scratch[PL_THREADS / WARP];
    i128 energy = 0;
    const BoxCache<Real> box = load_box<Real>(a.box);
    const Real cutoff = static_cast<Real>(a.cutoff);
    const Real cutoff2 = cutoff * cutoff;
    const Real beta = static_cast<Real>(a.beta);

    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < a.M; m += gridDim.x * blockDim.x) {
        const int i = a.pair_idxs[m * 2 + 0];
        const int j = a.pair_idxs[m * 2 + 1];
        const Real xi = static_cast<Real>(a.x[i * 3 + 0]);
        const Real yi = static_cast<Real>(a.x[i * 3 + 1]);
        const Real zi = static_cast<Real>(a.x[i * 3 + 2]);
        const Real xj = static_cast<Real>(a.x[j * 3 + 0]);
        const Real yj = static_cast<Real>(a.x[j * 3 + 1]);
        const Real zj = static_cast<Real>(a.x[j * 3 + 2]);
        const double *pi = a.p + static_cast<size_t>(i) * P_PER_ATOM;
        const double *pj = a.p + static_cast<size_t>(j) * P_PER_ATOM;
        const Real qi = static_cast<Real>(pi[P_CHARGE]), qj = static_cast<Real>(pj[P_CHARGE]);
        const Real si = static_cast<Real>(pi[P_SIG]), sj = static_cast<Real>(pj[P_SIG]);
        const Real ei = static_cast<Real>(pi[P_EPS]), ej = static_cast<Real>(pj[P_EPS]);
        const Real wi = static_cast<Real>(pi[P_W]), wj = static_cast<Real>(pj[P_W]);
        const Real q_scale = static_cast<Real>(a.scales[m * 2 + 0]);
        const Real lj_scale = static_cast<Real>(a.scales[m * 2 + 1]);

        const Real dx = min_image(xi - xj, box.x, box.inv_x);
        const Real dy = min_image(yi - yj, box.y, box.inv_y);
        const Real dz = min_image(zi - zj, box.z, box.inv_z);
        const Real dw = wi - wj;
        const Real d2 = fma_(dw, dw, dist2_3d(dx, dy, dz));

        if (d2 < cutoff2) {
            PairTerms<Real> t = pair_terms<Real, true>(q_scale, lj_scale, qi, qj, si, sj, ei, ej, d2, beta);
            if (a.du_dx != nullptr) {
                // j gets fixed(-v), not -fixed(v), like the reference (k_nonbonded_pair_list.cuh:146-152): the same inside the
                // int64 range, the reference's bits beyond it (then exclusions still cancel the all-pairs term exactly)
                const Real rx = t.prefactor * dx, ry = t.prefactor * dy, rz = t.prefactor * dz;
                accum<NEG>(a.du_dx + i * 3 + 0, to_fixed_force(rx));
                accum<NEG>(a.du_dx + i * 3 + 1, to_fixed_force(ry));
                accum<NEG>(a.du_dx + i * 3 + 2, to_fixed_force(rz));
                accum<NEG>(a.du_dx + j * 3 + 0, to_fixed_force(-rx));
                accum<NEG>(a.du_dx + j * 3 + 1, to_fixed_force(-ry));
                accum<NEG>(a.du_dx + j * 3 + 2, to_fixed_force(-rz));
            }
            if (a.du_dp != nullptr) {
                u64 *gi = a.du_dp + static_cast<size_t>(i) * P_PER_ATOM;
                u64 *gj = a.du_dp + static_cast<size_t>(j) * P_PER_ATOM;
                accum<NEG>(gi + P_CHARGE, to_fixed<FIXED_EXPONENT_DU_DCHARGE>(q_scale * qj * t.inv_d * t.damping));
                accum<NEG>(gj + P_CHARGE, to_fixed<FIXED_EXPONENT_DU_DCHARGE>(q_scale * qi * t.inv_d * t.damping));
                if (t.lj) {
                    const u64 fs = to_fixed<FIXED_EXPONENT_DU_DSIG>(t.sig_grad);
                    accum<NEG>(gi + P_SIG, fs);
                    accum<NEG>(gj + P_SIG, fs);
                    accum<NEG>(gi + P_EPS, to_fixed<FIXED_EXPONENT_DU_DEPS>(t.eps_grad * ej));
                    accum<NEG>(gj + P_EPS, to_fixed<FIXED_EXPONENT_DU_DEPS>(t.eps_grad * ei));
                }
                const Real vw = t.prefactor * dw;
                accum<NEG>(gi + P_W, to_fixed<FIXED_EXPONENT_DU_DW>(vw));
                accum<NEG>(gj + P_W, to_fixed<FIXED_EXPONENT_DU_DW>(-vw));
            }
            // negate AFTER the conversion: -fixed(u), never fixed(-u) (an overflowed term is LLONG_MAX either sign)
            const i128 e = energy_to_fixed<Real>(t.u);
            energy += NEG ? -e : e;
        }
    }
    if (a.d_u != nullptr) {
        grid_finish_energy(energy, scratch, a.u_partials, a.ticket, a.d_u);
    }
}

int pair_list_grid(int M) {
    int g = ceil_div(M, PL_THREADS);
    const int cap = sm_count() * 8;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

template <typename Real> void launch_pair_list(const PairListArgs<Real> &args, cudaStream_t stream) {
    const int grid = pair_list_grid(args.M);
    if (args.negated) {
        TMB_LAUNCH((k_pair_list<Real, true>), grid, PL_THREADS, 0, stream, args);
    } else {
        TMB_LAUNCH((k_pair_list<Real, false>), grid, PL_THREADS, 0, stream, args);
    }
}
template void launch_pair_list<float>(const PairListArgs<float> &, cudaStream_t);
template void launch_pair_list<double>(const PairListArgs<double> &, cudaStream_t);

} // namespace tmb
