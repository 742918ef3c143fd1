// Host classes for HarmonicBond / HarmonicAngle / PeriodicTorsion
// (reference harmonic_bond.cu:11-77, harmonic_angle.cu:11-79, periodic_torsion.cu:11-82).
#include "potential.hpp"

namespace tmb {

template <BondedKind KIND> struct BondedTraits;
template <> struct BondedTraits<BondedKind::Bond> {
    static constexpr int ARITY = 2;
    static constexpr int PARAMS = 2;
};
template <> struct BondedTraits<BondedKind::Angle> {
    static constexpr int ARITY = 3;
    static constexpr int PARAMS = 3;
};
template <> struct BondedTraits<BondedKind::Torsion> {
    static constexpr int ARITY = 4;
    static constexpr int PARAMS = 3;
};

template <BondedKind KIND> static void validate_bonded(const std::vector<int> &idxs) {
    constexpr int A = BondedTraits<KIND>::ARITY;
    if (idxs.size() % A != 0) {
        if (KIND == BondedKind::Bond) {
            throw std::runtime_error("bond_idxs.size() must be exactly 2*k!");
        } else if (KIND == BondedKind::Angle) {
            throw std::runtime_error("angle_idxs.size() must be exactly 3*A");
        } else {
            throw std::runtime_error("torsion_idxs.size() must be exactly 4*k");
        }
    }
    const int n = static_cast<int>(idxs.size() / A);
    for (int t = 0; t < n; t++) {
        for (int a = 0; a < A; a++) {
            for (int b = a + 1; b < A; b++) {
                if (idxs[t * A + a] == idxs[t * A + b]) {
                    if (KIND == BondedKind::Bond) {
                        throw std::runtime_error("src == dst");
                    } else if (KIND == BondedKind::Angle) {
                        throw std::runtime_error("angle triplets must be unique");
                    } else {
                        throw std::runtime_error("torsion quads must be unique");
                    }
                }
            }
        }
    }
}

template <typename Real, BondedKind KIND>
BondedPotential<Real, KIND>::BondedPotential(const std::vector<int> &idxs)
    : n_terms_(static_cast<int>(idxs.size() / BondedTraits<KIND>::ARITY)) {
    validate_bonded<KIND>(idxs);
    d_idxs_.realloc(idxs.size());
    d_idxs_.copy_from(idxs.data());
    d_partials_.realloc(bonded_grid(n_terms_));
    d_ticket_.realloc(1);
    d_ticket_.zero();
    TMB_CUDA(cudaDeviceSynchronize());
}

template <typename Real, BondedKind KIND>
void BondedPotential<Real, KIND>::execute_device(
    int, int P, const double *d_x, const double *d_p, const double *, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
    cudaStream_t stream) {
    constexpr int NP = BondedTraits<KIND>::PARAMS;
    if (KIND == BondedKind::Bond) {
        if (P != NP * n_terms_) {
            throw std::runtime_error(
                "HarmonicBond::execute_device(): expected P == 2*B, got P=" + std::to_string(P) +
                ", 2*B=" + std::to_string(2 * n_terms_));
        }
    }
    if (n_terms_ <= 0) {
        return; // the caller's energy slot stays at its zero-initialised value, as in the reference
    }
    if (KIND == BondedKind::Angle && P != NP * n_terms_) {
        throw std::runtime_error(
            "HarmonicAngle::execute_device(): expected P == 3*A_, got P=" + std::to_string(P) +
            ", 3*A_=" + std::to_string(3 * n_terms_));
    }
    if (KIND == BondedKind::Torsion && P != NP * n_terms_) {
        throw std::runtime_error(
            "PeriodicTorsion::execute_device(): expected P == 3*T_, got P=" + std::to_string(P) +
            ", 3*T_=" + std::to_string(3 * n_terms_));
    }
    BondedArgs a;
    a.n_terms = n_terms_;
    a.x = d_x;
    a.p = d_p;
    a.idxs = d_idxs_.data;
    a.du_dx = d_du_dx;
    a.du_dp = d_du_dp;
    a.u_partials = d_partials_.data;
    a.ticket = d_ticket_.data;
    a.d_u = d_u;
    if (KIND == BondedKind::Bond) {
        launch_harmonic_bond<Real>(a, stream);
    } else if (KIND == BondedKind::Angle) {
        launch_harmonic_angle<Real>(a, stream);
    } else {
        launch_periodic_torsion<Real>(a, stream);
    }
}

template class BondedPotential<float, BondedKind::Bond>;
template class BondedPotential<double, BondedKind::Bond>;
template class BondedPotential<float, BondedKind::Angle>;
template class BondedPotential<double, BondedKind::Angle>;
template class BondedPotential<float, BondedKind::Torsion>;
template class BondedPotential<double, BondedKind::Torsion>;

} // namespace tmb
