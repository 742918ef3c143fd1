// Water-exchange Monte Carlo (biased deletion) and the per-molecule energy machinery behind it  (SURVEY.md §8f rank 4,
// second half).
//
// Reference: timemachine/cpp/src/bd_exchange_move.{hpp,cu}, nonbonded_mol_energy.cu, segmented_sumexp.cu,
// segmented_weighted_random_sampler.cu, kernels/k_exchange.cu, k_sampling.cu, k_rotations.cu, k_logsumexp.cu,
// k_nonbonded.cuh:434-700, all_atom_energies.cu, rotations.cu, mol_utils.cpp.
//
// What is kept exactly: every quantity that feeds a decision is the reference's - per-pair f32/f64 energies rounded to
// 2^36 fixed point and summed as integers (so the per-molecule energies are bit-identical and independent of any
// summation order), log weights beta * fixed_to_float(E), the Gumbel-max selection of the molecule, the rotation
// about the fixed-point centroid and the translation imaged into the home box, the Metropolis test on the difference
// of log-sum-exps, "first accepted proposal of a batch wins", and the four cuRAND XORWOW streams (seed .. seed + 3)
// drawn with the reference's counts - a run with the same seed proposes and accepts the same moves.
//
// What is different (B200-first): the reference runs ~12 launches, two [batch, mol_size, N] float matrices, CUB
// segmented reductions and a blocking device->host copy of the noise offset PER BATCH of proposals.  Here a move() is
// three small launches (stage + all-molecule energies + initial weights) followed by ONE persistent cooperative
// kernel that walks all batches of proposals on the device: sample -> propose -> pair energies of the moved molecule
// against every atom (old and new position, straight into the int128 molecule energies, no N-wide matrices) ->
// log-sum-exp -> accept/store, separated by grid barriers, never returning to the host.
#pragma once

#include "nb_types.cuh"
#include "potential.hpp"

namespace tmb {

// reference mol_utils.cpp:36-106
void verify_mols_contiguous(const std::vector<std::vector<int>> &group_idxs);
struct MolLayout {
    std::vector<int> atom_idxs, mol_idxs, mol_offsets;
};
MolLayout flatten_mols(const std::vector<std::vector<int>> &group_idxs);

// reference nonbonded_mol_energy.{hpp,cu}: energy of every target molecule with all atoms outside it
template <typename Real> class NonbondedMolEnergyPotential {
public:
    NonbondedMolEnergyPotential(int N, const std::vector<std::vector<int>> &target_mols, double beta, double cutoff);
    // d_xr / d_pr: staged (x, y, z, w) and (q, sig, eps, -) per atom in Real
    void mol_energies_staged(const Vec4<Real> *d_xr, const Vec4<Real> *d_pr, const double *d_box, i128 *d_out, cudaStream_t stream);
    void mol_energies_device(int N, int target_mols, const double *d_coords, const double *d_params, const double *d_box, i128 *d_out, cudaStream_t stream);
    std::vector<i128> mol_energies_host(int N, int P, const double *h_coords, const double *h_params, const double *h_box);
    int num_mols() const { return num_target_mols_; }

private:
    const int N_;
    const int num_target_mols_;
    const Real beta_;
    const Real cutoff_squared_;
    int num_target_atoms_ = 0;
    DeviceBuffer<int4> d_targets_; // per target atom: {atom, mol, first atom of mol, last atom of mol}
    DeviceBuffer<Vec4<Real>> d_xr_, d_pr_;
};

// reference segmented_sumexp.{hpp,cu}
template <typename Real> class SegmentedSumExp {
public:
    SegmentedSumExp(int max_vals_per_segment, int num_segments);
    std::vector<Real> logsumexp_host(const std::vector<std::vector<Real>> &vals);

private:
    const int max_vals_per_segment_, num_segments_;
};

// reference segmented_weighted_random_sampler.{hpp,cu}
template <typename Real> class SegmentedWeightedRandomSampler {
public:
    SegmentedWeightedRandomSampler(int max_vals_per_segment, int num_segments, int seed);
    ~SegmentedWeightedRandomSampler();
    std::vector<int> sample_host(const std::vector<std::vector<Real>> &weights);

private:
    const int max_vals_per_segment_, num_segments_;
    curandGenerator_st *rng_ = nullptr;
    DeviceBuffer<Real> d_noise_;
};

template <typename Real> struct BDDevice; // kernel argument block (exchange.cu)

// reference bd_exchange_move.{hpp,cu}
template <typename Real> class BDExchangeMove : public Mover {
public:
    BDExchangeMove(
        int N, const std::vector<std::vector<int>> &target_mols, const std::vector<double> &params, double temperature,
        double nb_beta, double cutoff, int seed, int num_proposals_per_move, int interval, int batch_size);
    ~BDExchangeMove() override;

    void move(int N, double *d_coords, double *d_box, cudaStream_t stream) override;

    std::vector<std::vector<Real>> compute_incremental_log_weights_host(
        int N, const double *h_coords, const double *h_box, const int *h_mol_idxs, const Real *h_quaternions, const Real *h_translations);
    std::vector<Real> compute_initial_log_weights_host(int N, const double *h_coords, const double *h_box);
    std::vector<Real> get_before_log_weights();
    std::vector<Real> get_after_log_weights();
    virtual double raw_log_probability_host();
    double log_probability_host();
    size_t n_accepted() const;
    size_t n_proposed() const { return num_attempted_; }
    double acceptance_fraction() const { return static_cast<double>(n_accepted()) / static_cast<double>(n_proposed()); }
    size_t batch_size() const { return batch_size_; }
    int num_target_mols() const { return num_target_mols_; }
    std::vector<double> get_params();
    void set_params(const std::vector<double> &params);
    void set_params_device(int size, const double *d_p, cudaStream_t stream);

protected:
    BDExchangeMove(
        int N, const std::vector<std::vector<int>> &target_mols, const std::vector<double> &params, double temperature,
        double nb_beta, double cutoff, int seed, int num_proposals_per_move, int interval, int batch_size,
        size_t translation_buffer_size);
    const int N_, mol_size_, num_proposals_per_move_, num_target_mols_;
    const Real nb_beta_, beta_, cutoff_squared_;
    const int batch_size_;
    const int first_atom_; // molecules are contiguous and of one size: molecule m owns atoms first + m S .. first + (m+1) S - 1
    size_t num_attempted_ = 0;
    bool host_loop_ = false; // TMB_BD_LOOP=host: one launch per phase and a host round trip per batch (A/B and debugging)
    int coop_blocks_ = 0;

    NonbondedMolEnergyPotential<Real> mol_potential_;
    HilbertSort sorter_;                                  // molecules in Hilbert order: spatially coherent warps in the pair phase
    DeviceBuffer<unsigned int> d_anchor_atoms_, d_mol_order_; // first atom of every molecule; the same, Hilbert-sorted
    DeviceBuffer<Real> d_r_bound_;                        // [1] bound on the distance of a molecule's atoms from its first atom
    DeviceBuffer<Vec4<Real>> d_xs_, d_ps_;                // atoms in sweep order (Hilbert-sorted molecules, then the other atoms)
    DeviceBuffer<int> d_col_atom_;                        // atom index of every sweep slot
    DeviceBuffer<Vec4<Real>> d_chunk_ctr_, d_chunk_ext_;  // bounding box of every 128 sweep slots (home-box images)
    DeviceBuffer<double> d_params_;
    DeviceBuffer<Vec4<Real>> d_xr_, d_pr_, d_prop_, d_prop_old_;
    DeviceBuffer<i128> d_before_E_, d_total_;
    DeviceBuffer<Real> d_logw_before_, d_logw_after_;
    DeviceBuffer<Real> d_lse_before_, d_lse_after_max_, d_lse_after_sum_;
    DeviceBuffer<int> d_samples_, d_state_;
    DeviceBuffer<u64> d_num_accepted_;
    DeviceBuffer<Real> d_quat_, d_trans_, d_sample_noise_, d_mh_;
    curandGenerator_st *rng_quat_ = nullptr, *rng_trans_ = nullptr, *rng_samples_ = nullptr, *rng_mh_ = nullptr;

    BDDevice<Real> device_args(double *d_coords, const double *d_box, bool scale, bool sample);
    void initial_log_weights_device(double *d_coords, const double *d_box, cudaStream_t stream);
    void run_phase(int phase, const BDDevice<Real> &a, cudaStream_t stream);
    void run_proposals(BDDevice<Real> &a, cudaStream_t stream); // all batches of a move: device loop or host loop
    void set_generator_streams(cudaStream_t stream);
};

struct curandStateXORWOW;

// reference tibd_exchange_move.{hpp,cu}: targeted insertion / biased deletion between a sphere around the ligand
// centroid and the rest of the box
template <typename Real> class TIBDExchangeMove : public BDExchangeMove<Real> {
public:
    TIBDExchangeMove(
        int N, const std::vector<int> &ligand_idxs, const std::vector<std::vector<int>> &target_mols, const std::vector<double> &params,
        double temperature, double nb_beta, double cutoff, double radius, int seed, int num_proposals_per_move, int interval,
        int batch_size);
    ~TIBDExchangeMove() override;

    void move(int N, double *d_coords, double *d_box, cudaStream_t stream) override;
    std::array<std::vector<double>, 2> move_host(int N, const double *h_x, const double *h_box) override;
    double raw_log_probability_host() override;

private:
    const Real radius_, inner_volume_;
    DeviceBuffer<int> d_ligand_idxs_, d_inner_flags_, d_partition_, d_inner_count_, d_targeting_;
    DeviceBuffer<Real> d_center_, d_box_volume_, d_uniform_, d_src_logw_, d_dest_logw_, d_lse_src_max_, d_lse_src_sum_;
    curandStateXORWOW *d_rand_states_ = nullptr; // [TI_STATE_THREADS]
};

// reference exchange.cu:11-63 (molecules inside / outside a sphere around the centroid of center_atoms) and
// translations.cu:8-42 (n pairs of translations: inside the sphere, outside it)
template <typename Real>
std::array<std::vector<int>, 2> inner_and_outer_mols(
    const std::vector<int> &center_atoms, int N, const double *coords, const double *box, const std::vector<std::vector<int>> &group_idxs,
    Real radius);
template <typename Real>
std::vector<Real> translations_inside_and_outside_sphere_host(int n_translations, const double *box, const Real *center, Real radius, int seed);

// reference all_atom_energies.cu:8-48 (pair energies of target atoms with every atom, [T, N])
template <typename Real>
std::vector<Real> atom_by_atom_energies(
    int N, const std::vector<int> &target_atoms, const double *coords, const double *params, const double *box, Real nb_beta, Real cutoff);

// reference rotations.cu:12-92
template <typename Real>
void rotate_coordinates_host(int N, int n_rotations, const double *coords, const Real *quaternions, double *out);
template <typename Real>
void rotate_and_translate_mol_host(
    int N, int batch_size, const double *mol_coords, const double *box, const Real *quaternions, const Real *translations, double *out);

} // namespace tmb
