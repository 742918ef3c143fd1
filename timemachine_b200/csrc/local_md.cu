// Local MD (reference local_md_potentials.cu:27-335, kernels/k_flat_bottom_bond.cuh:22-70,
// context.cu:90-214): a shell of atoms around one reference atom is simulated, everything else is frozen.
//
//   selection   atom i is free with probability exp(-U_fb(|x_i - x_ref|; k, radius) / kT) against one uniform number per
//               atom; the reference atom is always frozen.  The uniforms are cuRAND XORWOW (seed, offset 0) and the
//               reference atom is drawn with std::mt19937 / std::uniform_int_distribution exactly as in the reference, so
//               a seed selects the same shell in both implementations.
//   potentials  every potential of the context as it is (bonded terms, exclusions: forces on frozen atoms are computed
//               and ignored), the NonbondedAllPairs restricted to the free atoms, plus a NonbondedInteractionGroup
//               free x frozen and a FlatBottomBond (k, 0, radius) from every free atom to the reference atom.
//               freeze_reference = false: the reference atom moves too, and a LogFlatBottomBond (beta = 1 / kT, same
//               k and radius) between it and every frozen atom keeps it from dragging the shell over frozen atoms.
//   integrator  BAOAB over the free atoms only (k_baoab with an index array; frozen entries are N).
//
// The reference keeps all index bookkeeping on the device (cub::DevicePartition) but synchronises the stream to read the
// number of free atoms (local_md_potentials.cu:225-232); here the selection flags are read back in that same
// synchronisation and the row / column lists are formed on the host - setup runs once per call of hundreds of steps.
#include "potential.hpp"

#include <curand.h>

#include <random>

namespace tmb {

#define TMB_CURAND(call)                                                                                               \
    do {                                                                                                               \
        curandStatus_t st_ = (call);                                                                                   \
        if (st_ != CURAND_STATUS_SUCCESS) {                                                                            \
            throw std::runtime_error(std::string("cuRAND error ") + std::to_string(static_cast<int>(st_)) + " at " +   \
                                     __FILE__ + ":" + std::to_string(__LINE__));                                       \
        }                                                                                                              \
    } while (0)

// reference k_log_probability_selection<float> (k_flat_bottom_bond.cuh:22-70): f64 coordinate difference rounded to f32,
// minimum image, flat-bottom energy (k/4)(r - radius)^4 beyond the radius, probability exp(-energy / kT) in f64
__global__ void k_local_md_select(
    const int N, const double kBT, const float radius, const float k, const unsigned int reference_idx,
    const double *__restrict__ coords, const double *__restrict__ box, const float *__restrict__ uniforms,
    unsigned int *__restrict__ selected) {
    const unsigned int idx = blockDim.x * blockIdx.x + threadIdx.x;
    if (idx >= static_cast<unsigned int>(N)) {
        return;
    }
    const float radius_sq = radius * radius;
    const float bx = static_cast<float>(box[0]), by = static_cast<float>(box[4]), bz = static_cast<float>(box[8]);
    const float inv_bx = 1.0f / bx, inv_by = 1.0f / by, inv_bz = 1.0f / bz;
    float dx = static_cast<float>(coords[idx * 3 + 0] - coords[reference_idx * 3 + 0]);
    float dy = static_cast<float>(coords[idx * 3 + 1] - coords[reference_idx * 3 + 1]);
    float dz = static_cast<float>(coords[idx * 3 + 2] - coords[reference_idx * 3 + 2]);
    dx = __fmaf_rn(-bx, rintf(dx * inv_bx), dx);
    dy = __fmaf_rn(-by, rintf(dy * inv_by), dy);
    dz = __fmaf_rn(-bz, rintf(dz * inv_bz), dz);
    const float distance_sq = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
    float prob = 1.0f;
    if (distance_sq >= radius_sq) {
        const float d = sqrtf(distance_sq) - radius;
        const float d2 = d * d;
        const float energy = (k / 4.0f) * (d2 * d2);
        prob = static_cast<float>(exp(static_cast<double>(-energy) / kBT));
    }
    selected[idx] = (idx != reference_idx && prob >= uniforms[idx]) ? idx : static_cast<unsigned int>(N);
}

namespace {

// the NonbondedAllPairs of the system with a host copy of its parameters (reference nonbonded_common.cpp:76-124)
struct FoundAllPairs {
    std::shared_ptr<Potential> potential;
    std::vector<double> params;
};

bool is_all_pairs(const std::shared_ptr<Potential> &p) {
    return std::dynamic_pointer_cast<NonbondedAllPairs<float>>(p) != nullptr ||
           std::dynamic_pointer_cast<NonbondedAllPairs<double>>(p) != nullptr;
}

void find_all_pairs(const std::shared_ptr<Potential> &pot, const std::vector<double> &params, std::vector<FoundAllPairs> &out) {
    if (auto fan = std::dynamic_pointer_cast<FanoutSummedPotential>(pot)) {
        for (auto &child : fan->get_potentials()) {
            find_all_pairs(child, params, out);
        }
    } else if (auto sum = std::dynamic_pointer_cast<SummedPotential>(pot)) {
        size_t offset = 0;
        const auto &sizes = sum->get_parameter_sizes();
        for (size_t i = 0; i < sum->get_potentials().size(); i++) {
            std::vector<double> slice(params.begin() + offset, params.begin() + offset + sizes[i]);
            find_all_pairs(sum->get_potentials()[i], slice, out);
            offset += sizes[i];
        }
    } else if (is_all_pairs(pot)) {
        out.push_back({pot, params});
    }
}

} // namespace

LocalMD::LocalMD(int N, const std::vector<std::shared_ptr<BoundPotential>> &bps, bool freeze_reference_, double temperature_)
    : freeze_reference(freeze_reference_), temperature(temperature_), N_(N), base_(bps), d_uniforms_(N), d_selected_(N),
      h_selected_(N) {
    if (temperature <= 0.0) {
        throw std::runtime_error("temperature must be greater than 0");
    }
    if (N < 2) {
        throw std::runtime_error("N must be greater than 1");
    }
    std::vector<FoundAllPairs> found;
    for (auto &bp : bps) {
        std::vector<double> h(bp->size);
        if (bp->size > 0) {
            bp->d_p.copy_to(h.data());
        }
        find_all_pairs(bp->potential, h, found);
    }
    if (found.size() > 1) {
        throw std::runtime_error("found multiple NonbondedAllPairs potentials");
    }
    if (found.size() != 1) {
        throw std::runtime_error("unable to find a NonbondedAllPairs potential");
    }
    all_pairs_ = found[0].potential;
    const std::vector<int> rows{0}, cols{1};
    std::shared_ptr<Potential> ixn;
    if (auto f = std::dynamic_pointer_cast<NonbondedAllPairs<float>>(all_pairs_)) {
        original_idxs_ = f->get_atom_idxs();
        ixn = std::make_shared<NonbondedInteractionGroup<float>>(N, rows, cols, f->get_beta(), f->get_cutoff(), false, f->get_nblist_padding());
    } else {
        auto d = std::dynamic_pointer_cast<NonbondedAllPairs<double>>(all_pairs_);
        original_idxs_ = d->get_atom_idxs();
        ixn = std::make_shared<NonbondedInteractionGroup<double>>(N, rows, cols, d->get_beta(), d->get_cutoff(), false, d->get_nblist_padding());
    }
    ixn_group_ = std::make_shared<BoundPotential>(ixn, found[0].params);
    TMB_CURAND(curandCreateGenerator(&rng_, CURAND_RNG_PSEUDO_DEFAULT));
}

LocalMD::~LocalMD() { curandDestroyGenerator(rng_); }

void LocalMD::set_all_pairs_idxs(const std::vector<int> &idxs) {
    if (auto f = std::dynamic_pointer_cast<NonbondedAllPairs<float>>(all_pairs_)) {
        f->set_atom_idxs(idxs);
    } else {
        std::dynamic_pointer_cast<NonbondedAllPairs<double>>(all_pairs_)->set_atom_idxs(idxs);
    }
}

void LocalMD::setup_from_idxs(
    const double *d_x, const double *d_box, const std::vector<int> &local_idxs, int seed, double radius, double k, cudaStream_t stream) {
    TMB_CURAND(curandSetStream(rng_, stream));
    TMB_CURAND(curandSetPseudoRandomGeneratorSeed(rng_, seed));
    TMB_CURAND(curandSetGeneratorOffset(rng_, 0)); // same seed -> same numbers (local_md_potentials.cu:121-123)
    TMB_CURAND(curandGenerateUniform(rng_, d_uniforms_.data, d_uniforms_.length));
    std::mt19937 rng;
    rng.seed(seed);
    std::uniform_int_distribution<unsigned int> pick(0, static_cast<unsigned int>(local_idxs.size()) - 1);
    const unsigned int reference_idx = static_cast<unsigned int>(local_idxs[pick(rng)]);
    const double kBT = 0.008314462618 * temperature;
    TMB_LAUNCH(k_local_md_select, ceil_div(N_, 128), 128, 0, stream, N_, kBT, static_cast<float>(radius), static_cast<float>(k),
               reference_idx, d_x, d_box, static_cast<const float *>(d_uniforms_.data), d_selected_.data);
    TMB_CUDA(cudaMemcpyAsync(h_selected_.data(), d_selected_.data, N_ * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    TMB_CUDA(cudaStreamSynchronize(stream));
    configure(reference_idx, radius, k);
}

void LocalMD::setup_from_selection(int reference_idx, const std::vector<int> &selection_idxs, double radius, double k, cudaStream_t stream) {
    std::fill(h_selected_.begin(), h_selected_.end(), static_cast<unsigned int>(N_));
    for (int i : selection_idxs) {
        h_selected_[i] = static_cast<unsigned int>(i);
    }
    TMB_CUDA(cudaMemcpyAsync(d_selected_.data, h_selected_.data(), N_ * sizeof(unsigned int), cudaMemcpyHostToDevice, stream));
    TMB_CUDA(cudaStreamSynchronize(stream));
    configure(static_cast<unsigned int>(reference_idx), radius, k);
}

void LocalMD::configure(unsigned int reference_idx, double radius, double k) {
    if (!freeze_reference) {
        // the reference atom is integrated as well (local_md_potentials.cu:184-189)
        h_selected_[reference_idx] = reference_idx;
        TMB_CUDA(cudaMemcpy(d_selected_.data + reference_idx, &reference_idx, sizeof(unsigned int), cudaMemcpyHostToDevice));
    }
    // free atoms that the all-pairs term knows about become its (and the interaction group's row) atoms; the rest of its
    // atoms (the reference among them when it is frozen) are the columns (local_md_potentials.cu:190-300)
    std::vector<int> rows, cols;
    for (int a : original_idxs_) {
        (h_selected_[a] < static_cast<unsigned int>(N_) ? rows : cols).push_back(a);
    }
    if (rows.empty()) {
        throw std::runtime_error("LocalMDPotentials setup has no free particles selected");
    }
    const int n_rows = static_cast<int>(rows.size());
    if (n_rows == N_ - 1 || (!freeze_reference && n_rows == N_)) {
        fprintf(stderr, "LocalMDPotentials setup has entire system selected\n");
    }
    set_all_pairs_idxs(rows);
    modified_ = true;
    if (!cols.empty()) {
        if (auto f = std::dynamic_pointer_cast<NonbondedInteractionGroup<float>>(ixn_group_->potential)) {
            f->set_atom_idxs(rows, cols);
        } else {
            std::dynamic_pointer_cast<NonbondedInteractionGroup<double>>(ixn_group_->potential)->set_atom_idxs(rows, cols);
        }
    }
    auto bonds_to_reference = [&](const std::vector<int> &atoms, std::vector<int> &bonds, std::vector<double> &params) {
        for (int a : atoms) {
            if (a == static_cast<int>(reference_idx)) {
                continue; // the reference builds a (reference, reference) bond here, which contributes nothing
            }
            bonds.push_back(static_cast<int>(reference_idx));
            bonds.push_back(a);
            params.push_back(k);
            params.push_back(0.0);
            params.push_back(radius);
        }
    };
    std::vector<int> bonds;
    std::vector<double> params;
    bonds_to_reference(rows, bonds, params);
    auto restraint = std::make_shared<FlatBottomBond<float>>(bonds, std::vector<int>{}, 0.0, 0.0);
    free_restraint_ = std::make_shared<BoundPotential>(restraint, params);
    active_ = base_;
    active_.push_back(free_restraint_);
    if (!cols.empty()) {
        active_.push_back(ixn_group_);
    }
    if (!freeze_reference && !cols.empty()) {
        std::vector<int> fbonds;
        std::vector<double> fparams;
        bonds_to_reference(cols, fbonds, fparams);
        auto frozen = std::make_shared<LogFlatBottomBond<float>>(fbonds, std::vector<int>{}, 1.0 / (0.008314462618 * temperature), 0.0);
        frozen_restraint_ = std::make_shared<BoundPotential>(frozen, fparams);
        active_.push_back(frozen_restraint_);
    }
    num_free_ = n_rows;
}

void LocalMD::reset() {
    if (modified_) {
        set_all_pairs_idxs(original_idxs_);
        modified_ = false;
    }
}

} // namespace tmb
