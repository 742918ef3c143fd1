// Host classes of the restraints and the precomputed pair list.  Reference:
// flat_bottom_bond.cu:11-94, chiral_atom_restraint.cu:11-70, chiral_bond_restraint.cu:11-85,
// nonbonded_precomputed.cu:12-109; argument checks and messages follow those files.
#include "fixed_point.cuh"
#include "potential.hpp"

#include <cstdlib>

namespace tmb {

template <typename Real, RestraintKind KIND>
RestraintPotential<Real, KIND>::RestraintPotential(
    const std::vector<int> &idxs, const std::vector<int> &signs, double beta, double cutoff)
    : n_terms_(static_cast<int>(idxs.size() / ARITY)), beta_(beta), cutoff_(cutoff) {
    if (KIND == RestraintKind::LogFlatBottomBond && beta <= 0) {
        throw std::runtime_error("beta must be positive"); // log_flat_bottom_bond.cu:15-17
    }
    if (idxs.size() % ARITY != 0) {
        switch (KIND) {
        case RestraintKind::FlatBottomBond:
        case RestraintKind::LogFlatBottomBond:
            throw std::runtime_error("bond_idxs.size() must be exactly 2*k!");
        case RestraintKind::ChiralAtom:
            throw std::runtime_error("idxs.size() must be exactly 4*k!");
        case RestraintKind::ChiralBond:
            throw std::runtime_error("idxs.size() must be exactly 4*R!");
        case RestraintKind::PrecomputedPairs:
            throw std::runtime_error("idxs.size() must be exactly 2*B!");
        }
    }
    if (KIND == RestraintKind::ChiralBond) {
        if (static_cast<size_t>(n_terms_) != signs.size()) {
            throw std::runtime_error("signs.size() must be exactly R!");
        }
        for (int s : signs) {
            if (s != 1 && s != -1) {
                throw std::runtime_error("signs must be comprised exclusively of 1 or -1");
            }
        }
    }
    if (IS_FB || KIND == RestraintKind::PrecomputedPairs) {
        for (int t = 0; t < n_terms_; t++) {
            const int src = idxs[t * 2 + 0], dst = idxs[t * 2 + 1];
            if (src == dst) {
                if (IS_FB) {
                    throw std::runtime_error("src == dst");
                }
                throw std::runtime_error("illegal pair with src == dst: " + std::to_string(src) + ", " + std::to_string(dst));
            }
            if (IS_FB && (src < 0 || dst < 0)) {
                throw std::runtime_error("idxs must be non-negative");
            }
        }
    }
    d_idxs_.realloc(std::max<size_t>(1, idxs.size()));
    if (!idxs.empty()) {
        TMB_CUDA(cudaMemcpy(d_idxs_.data, idxs.data(), idxs.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (KIND == RestraintKind::ChiralBond) {
        d_signs_.realloc(std::max<size_t>(1, signs.size()));
        if (!signs.empty()) {
            TMB_CUDA(cudaMemcpy(d_signs_.data, signs.data(), signs.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
    }
    d_partials_.realloc(bonded_grid(n_terms_));
    d_ticket_.realloc(1);
    d_ticket_.zero();
    TMB_CUDA(cudaDeviceSynchronize());
}

template <typename Real, RestraintKind KIND>
void RestraintPotential<Real, KIND>::execute_device(
    int, int P, const double *d_x, const double *d_p, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
    cudaStream_t stream) {
    const int expected = PARAMS * n_terms_;
    if (P != expected) {
        switch (KIND) {
        case RestraintKind::FlatBottomBond:
            throw std::runtime_error(
                "FlatBottomBond::execute_device(): expected P == " + std::to_string(expected) + ", got P=" + std::to_string(P));
        case RestraintKind::LogFlatBottomBond:
            throw std::runtime_error(
                "LogFlatBottomBond::execute_device(): expected P == " + std::to_string(expected) + ", got P=" + std::to_string(P));
        case RestraintKind::ChiralAtom:
            throw std::runtime_error(
                "ChiralAtomRestraint::execute_device(): expected P == R, got P=" + std::to_string(P) +
                ", R=" + std::to_string(n_terms_));
        case RestraintKind::ChiralBond:
            throw std::runtime_error(
                "ChiralBondRestraint::execute_device(): expected P == R, got P=" + std::to_string(P) +
                ", R=" + std::to_string(n_terms_));
        case RestraintKind::PrecomputedPairs:
            throw std::runtime_error(
                "NonbondedPairListPrecomputed::execute_device(): expected P == 4*B, got P=" + std::to_string(P) +
                ", 4*B=" + std::to_string(expected));
        }
    }
    if (n_terms_ <= 0) {
        return;
    }
    RestraintArgs a;
    a.b.n_terms = n_terms_;
    a.b.x = d_x;
    a.b.p = d_p;
    a.b.idxs = d_idxs_.data;
    a.b.du_dx = d_du_dx;
    a.b.du_dp = d_du_dp;
    a.b.u_partials = d_partials_.data;
    a.b.ticket = d_ticket_.data;
    a.b.d_u = d_u;
    a.box = d_box;
    a.signs = d_signs_.data;
    a.beta = beta_;
    a.cutoff = cutoff_;
    if (KIND == RestraintKind::PrecomputedPairs) {
        static const bool full = [] {
            const char *e = std::getenv("TMB_PRECOMPUTED_FULL_GRADIENT");
            return e != nullptr && e[0] == '1';
        }();
        a.full_gradient = full;
    }
    switch (KIND) {
    case RestraintKind::FlatBottomBond:
        launch_flat_bottom_bond<Real>(a, stream);
        break;
    case RestraintKind::LogFlatBottomBond:
        launch_log_flat_bottom_bond<Real>(a, stream);
        break;
    case RestraintKind::ChiralAtom:
        launch_chiral_atom_restraint<Real>(a, stream);
        break;
    case RestraintKind::ChiralBond:
        launch_chiral_bond_restraint<Real>(a, stream);
        break;
    case RestraintKind::PrecomputedPairs:
        launch_nonbonded_precomputed<Real>(a, stream);
        break;
    }
}

template <typename Real, RestraintKind KIND>
void RestraintPotential<Real, KIND>::du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const {
    if (KIND == RestraintKind::PrecomputedPairs) {
        nonbonded_du_dp_fixed_to_float(n_terms_, P, du_dp, out); // per-column scales, rows are pairs
    } else {
        Potential::du_dp_fixed_to_float(N, P, du_dp, out);
    }
}

// ---- CentroidRestraint (reference centroid_restraint.cu:10-71) -----------------------------------------------------------
template <typename Real>
CentroidRestraint<Real>::CentroidRestraint(
    const std::vector<int> &group_a_idxs, const std::vector<int> &group_b_idxs, double kb, double b0)
    : n_a_(static_cast<int>(group_a_idxs.size())), n_b_(static_cast<int>(group_b_idxs.size())), kb_(kb), b0_(b0) {
    d_group_a_.realloc(std::max<size_t>(1, group_a_idxs.size()));
    d_group_b_.realloc(std::max<size_t>(1, group_b_idxs.size()));
    if (n_a_ > 0) {
        TMB_CUDA(cudaMemcpy(d_group_a_.data, group_a_idxs.data(), group_a_idxs.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (n_b_ > 0) {
        TMB_CUDA(cudaMemcpy(d_group_b_.data, group_b_idxs.data(), group_b_idxs.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
}

template <typename Real>
void CentroidRestraint<Real>::execute_device(
    int, int, const double *d_x, const double *, const double *, u64 *d_du_dx, u64 *, i128 *d_u, cudaStream_t stream) {
    if (n_a_ + n_b_ <= 0) {
        if (d_u != nullptr) {
            TMB_CUDA(cudaMemsetAsync(d_u, 0, sizeof(i128), stream)); // execute_device overwrites the energy
        }
        return;
    }
    if (d_du_dx == nullptr && d_u == nullptr) {
        return; // no parameters: nothing to do for a du/dp-only request
    }
    CentroidArgs a;
    a.x = d_x;
    a.group_a = d_group_a_.data;
    a.group_b = d_group_b_.data;
    a.n_a = n_a_;
    a.n_b = n_b_;
    a.kb = kb_;
    a.b0 = b0_;
    a.du_dx = d_du_dx;
    a.d_u = d_u;
    launch_centroid_restraint<Real>(a, stream);
}
template class CentroidRestraint<float>;
template class CentroidRestraint<double>;

#define TMB_INSTANTIATE(KIND)                                                                                          \
    template class RestraintPotential<float, KIND>;                                                                    \
    template class RestraintPotential<double, KIND>;
TMB_INSTANTIATE(RestraintKind::FlatBottomBond)
TMB_INSTANTIATE(RestraintKind::ChiralAtom)
TMB_INSTANTIATE(RestraintKind::ChiralBond)
TMB_INSTANTIATE(RestraintKind::PrecomputedPairs)
TMB_INSTANTIATE(RestraintKind::LogFlatBottomBond)

} // namespace tmb
