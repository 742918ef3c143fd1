// Host-side classes of the hot path.  They mirror the reference's C++ layer (timemachine/cpp/src/*.hpp) class for
// class - Potential, BoundPotential, Summed/FanoutSummedPotential, HarmonicBond/Angle, PeriodicTorsion,
// NonbondedAllPairs/InteractionGroup/PairList, Neighborlist, HilbertSort, LangevinIntegrator, Context - so the
// Python shim (timemachine_b200/custom_ops.py) can present the reference's custom_ops API unchanged.
#pragma once

#include <memory>
#include <optional>
#include <set>
#include <vector>

#include "common.cuh"
#include "kernels.hpp"

struct curandGenerator_st; // <curand.h>, only local_md.cu and barostat.cu include it

namespace tmb {

// The stream host-level entry points run on (default: the legacy default stream, like the reference's
// execute_host / Context::multiple_steps, potential.cu:254, context.cu:225).
cudaStream_t main_stream();
void set_main_stream(cudaStream_t s);

// ---------------------------------------------------------------------------------------------------------------
template <typename T> class DeviceBuffer {
public:
    T *data = nullptr;
    size_t length = 0;

    DeviceBuffer() = default;
    explicit DeviceBuffer(size_t n) { realloc(n); }
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    ~DeviceBuffer() { release(); }

    void release() {
        if (data != nullptr) {
            cudaFree(data); // never throw from a destructor
            data = nullptr;
            length = 0;
        }
    }
    void realloc(size_t n) {
        release();
        if (n > 0) {
            TMB_CUDA(cudaMalloc(&data, n * sizeof(T)));
        }
        length = n;
    }
    size_t bytes() const { return length * sizeof(T); }
    void zero(cudaStream_t s = nullptr) {
        if (length) {
            TMB_CUDA(cudaMemsetAsync(data, 0, bytes(), s));
        }
    }
    void copy_from(const T *host) {
        if (length) {
            TMB_CUDA(cudaMemcpy(data, host, bytes(), cudaMemcpyHostToDevice));
        }
    }
    void copy_to(T *host) const {
        if (length) {
            TMB_CUDA(cudaMemcpy(host, data, bytes(), cudaMemcpyDeviceToHost));
        }
    }
};

// Temporaries of the host-buffer entry points (execute / execute_batch / execute_batch_sparse copy their arguments to
// the device on every call, reference potential.cu:70-320).  The reference cudaMallocs and cudaFrees them per call; both
// are device-wide synchronisation points whose cost varies by an order of magnitude from process to process (measured:
// 2 ms vs 56 ms per execute_batch_sparse of the 30k-atom system next to instantiated CUDA graphs).  These buffers come
// from a process-wide cache of freed blocks instead: after the first call of a given size no driver allocation happens.
void *scratch_alloc(size_t bytes);
void scratch_free(void *ptr, size_t bytes);

template <typename T> class ScratchBuffer {
public:
    T *data = nullptr;
    size_t length = 0;

    ScratchBuffer() = default;
    explicit ScratchBuffer(size_t n) { realloc(n); }
    ScratchBuffer(const ScratchBuffer &) = delete;
    ScratchBuffer &operator=(const ScratchBuffer &) = delete;
    ~ScratchBuffer() { release(); }

    void release() {
        if (data != nullptr) {
            scratch_free(data, length * sizeof(T));
            data = nullptr;
            length = 0;
        }
    }
    void realloc(size_t n) {
        release();
        if (n > 0) {
            data = static_cast<T *>(scratch_alloc(n * sizeof(T)));
        }
        length = n;
    }
    size_t bytes() const { return length * sizeof(T); }
    void zero(cudaStream_t s = nullptr) {
        if (length) {
            TMB_CUDA(cudaMemsetAsync(data, 0, bytes(), s));
        }
    }
    void copy_from(const T *host) {
        if (length) {
            TMB_CUDA(cudaMemcpy(data, host, bytes(), cudaMemcpyHostToDevice));
        }
    }
    void copy_to(T *host) const {
        if (length) {
            TMB_CUDA(cudaMemcpy(host, data, bytes(), cudaMemcpyDeviceToHost));
        }
    }
};

// Fork/join of per-child streams (reference stream_manager.cu:18-56)
class StreamFan {
public:
    StreamFan() = default;
    ~StreamFan();
    cudaStream_t fork(int i, cudaStream_t parent); // child stream i now waits for everything enqueued on parent
    void join(int i, cudaStream_t parent);         // parent now waits for everything enqueued on child i
private:
    std::vector<cudaStream_t> streams_;
    std::vector<cudaEvent_t> events_;
    cudaEvent_t parent_event_ = nullptr;
    void ensure(int i);
};

// ---------------------------------------------------------------------------------------------------------------
// What the integrator needs to run an all-pairs potential's next k_nb_prepare inside its own kernel (kernels.hpp,
// launch_baoab_prepare); filled by NonbondedAllPairs::fused_prepare_hook.
struct FusedPrepareHook {
    bool is_double = false;
    FusedPrepareArgs<float> f32{};
    FusedPrepareArgs<double> f64{};
};

class Potential {
public:
    virtual ~Potential() = default;
    static constexpr int D = 3;

    // Accumulates into d_du_dx / d_du_dp (callers zero them), OVERWRITES d_u.  Any output may be null.
    virtual void execute_device(
        int N, int P, const double *d_x, const double *d_p, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
        cudaStream_t stream) = 0;

    virtual void du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const;

    // How many consecutive execute_device calls from now enqueue an identical launch sequence with no host-side
    // decision in between, i.e. may be recorded into a CUDA graph and replayed (0: the next call is special, e.g. a
    // Hilbert re-sort).
    virtual int capturable_steps() const { return 1 << 30; }
    // Advance host-side cadence counters as if execute_device had been called `n` times (graph replays).
    virtual void advance(int n) { (void)n; }
    // Tile-list potentials size their list buffer from measured counts (neighborlist.cu); a build that ran out of room
    // leaves a record on the device.  Called by the host after a synchronisation: true = the buffer was too small during
    // the work just synchronised, it has been grown, and that work must be redone (its results miss interactions).
    virtual bool recover_overflow() { return false; }
    // Fusion of the integrator with the next evaluation's gather / rebuild decision (Context, inside CUDA-graph blocks):
    // a potential that can hand its prepare pass to the integrator fills `hook` and returns true; skip_next_prepare()
    // tells it that its NEXT execute_device finds that pass already done.
    virtual bool fused_prepare_hook(FusedPrepareHook &hook) {
        (void)hook;
        return false;
    }
    virtual void skip_next_prepare() {}
    virtual void cancel_skip_prepare() {}

    void execute_host(
        int N, int P, const double *h_x, const double *h_p, const double *h_box, u64 *h_du_dx, u64 *h_du_dp, i128 *h_u);
    void execute_batch_host(
        int coord_batch, int N, int param_batch, int P, const double *h_x, const double *h_p, const double *h_box,
        u64 *h_du_dx, u64 *h_du_dp, i128 *h_u);
    void execute_batch_sparse_host(
        int coords_size, int N, int params_size, int P, int batch_size, const unsigned int *coords_idxs,
        const unsigned int *params_idxs, const double *h_x, const double *h_p, const double *h_box, u64 *h_du_dx,
        u64 *h_du_dp, i128 *h_u);
};

class BoundPotential {
public:
    BoundPotential(std::shared_ptr<Potential> potential, const std::vector<double> &params);
    int size;
    DeviceBuffer<double> d_p;
    std::shared_ptr<Potential> potential;

    void set_params(const std::vector<double> &params);
    void set_params_device(int size, const double *d_params, cudaStream_t stream);
    void execute_device(int N, const double *d_x, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u, cudaStream_t s);
    void execute_host(int N, const double *h_x, const double *h_box, u64 *h_du_dx, i128 *h_u);
    void execute_batch_host(int coord_batch, int N, const double *h_x, const double *h_box, u64 *h_du_dx, i128 *h_u);
};

class SummedPotential : public Potential {
public:
    SummedPotential(std::vector<std::shared_ptr<Potential>> potentials, std::vector<int> params_sizes, bool parallel);
    const std::vector<std::shared_ptr<Potential>> &get_potentials() const { return potentials_; }
    const std::vector<int> &get_parameter_sizes() const { return params_sizes_; }
    void execute_device(int, int, const double *, const double *, const double *, u64 *, u64 *, i128 *, cudaStream_t) override;
    void du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const override;
    int capturable_steps() const override;
    void advance(int n) override;
    bool recover_overflow() override;
private:
    std::vector<std::shared_ptr<Potential>> potentials_;
    std::vector<int> params_sizes_;
    int P_;
    bool parallel_;
    DeviceBuffer<i128> d_u_children_;
    StreamFan fan_;
};

class FanoutSummedPotential : public Potential {
public:
    FanoutSummedPotential(std::vector<std::shared_ptr<Potential>> potentials, bool parallel);
    const std::vector<std::shared_ptr<Potential>> &get_potentials() const { return potentials_; }
    void execute_device(int, int, const double *, const double *, const double *, u64 *, u64 *, i128 *, cudaStream_t) override;
    void du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const override;
    int capturable_steps() const override;
    void advance(int n) override;
    bool recover_overflow() override;
private:
    std::vector<std::shared_ptr<Potential>> potentials_;
    bool parallel_;
    DeviceBuffer<i128> d_u_children_;
    StreamFan fan_;
};

// ---------------------------------------------------------------------------------------------------------------
enum class BondedKind { Bond, Angle, Torsion };

template <typename Real, BondedKind KIND> class BondedPotential : public Potential {
public:
    explicit BondedPotential(const std::vector<int> &idxs);
    void execute_device(int, int, const double *, const double *, const double *, u64 *, u64 *, i128 *, cudaStream_t) override;
    int num_terms() const { return n_terms_; }
private:
    int n_terms_;
    DeviceBuffer<int> d_idxs_;
    DeviceBuffer<i128> d_partials_;
    DeviceBuffer<unsigned int> d_ticket_;
};
template <typename Real> using HarmonicBond = BondedPotential<Real, BondedKind::Bond>;
template <typename Real> using HarmonicAngle = BondedPotential<Real, BondedKind::Angle>;
template <typename Real> using PeriodicTorsion = BondedPotential<Real, BondedKind::Torsion>;

// Restraints + precomputed pair list (reference flat_bottom_bond.hpp, chiral_atom_restraint.hpp,
// chiral_bond_restraint.hpp, nonbonded_precomputed.hpp)
enum class RestraintKind { FlatBottomBond, ChiralAtom, ChiralBond, PrecomputedPairs, LogFlatBottomBond };

template <typename Real, RestraintKind KIND> class RestraintPotential : public Potential {
public:
    static constexpr bool IS_FB = KIND == RestraintKind::FlatBottomBond || KIND == RestraintKind::LogFlatBottomBond;
    static constexpr int ARITY = (IS_FB || KIND == RestraintKind::PrecomputedPairs) ? 2 : 4;
    static constexpr int PARAMS = IS_FB ? 3 : (KIND == RestraintKind::PrecomputedPairs ? 4 : 1);
    RestraintPotential(const std::vector<int> &idxs, const std::vector<int> &signs, double beta, double cutoff);
    void execute_device(int, int, const double *, const double *, const double *, u64 *, u64 *, i128 *, cudaStream_t) override;
    void du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const override;
    int num_terms() const { return n_terms_; }
private:
    int n_terms_;
    double beta_, cutoff_;
    DeviceBuffer<int> d_idxs_, d_signs_;
    DeviceBuffer<i128> d_partials_;
    DeviceBuffer<unsigned int> d_ticket_;
};
template <typename Real> using FlatBottomBond = RestraintPotential<Real, RestraintKind::FlatBottomBond>;
// LogFlatBottomBond(bond_idxs, beta): reference log_flat_bottom_bond.cu (the (idxs, signs, beta, cutoff) ctor with beta)
template <typename Real> using LogFlatBottomBond = RestraintPotential<Real, RestraintKind::LogFlatBottomBond>;
template <typename Real> using ChiralAtomRestraint = RestraintPotential<Real, RestraintKind::ChiralAtom>;
template <typename Real> using ChiralBondRestraint = RestraintPotential<Real, RestraintKind::ChiralBond>;
template <typename Real> using NonbondedPairListPrecomputed = RestraintPotential<Real, RestraintKind::PrecomputedPairs>;

// CentroidRestraint(group_a_idxs, group_b_idxs, kb, b0): kb (|<x_a> - <x_b>| - b0)^2, geometric centroids, no parameters
// (reference centroid_restraint.cu:10-71)
template <typename Real> class CentroidRestraint : public Potential {
public:
    CentroidRestraint(const std::vector<int> &group_a_idxs, const std::vector<int> &group_b_idxs, double kb, double b0);
    void execute_device(int, int, const double *, const double *, const double *, u64 *, u64 *, i128 *, cudaStream_t) override;
private:
    int n_a_, n_b_;
    double kb_, b0_;
    DeviceBuffer<int> d_group_a_, d_group_b_;
};

// ---------------------------------------------------------------------------------------------------------------
class HilbertSort {
public:
    explicit HilbertSort(int N);
    void sort_device(int n, const unsigned int *d_atom_idxs, const double *d_coords, const double *d_box,
                     unsigned int *d_perm_out, cudaStream_t stream);
    std::vector<unsigned int> sort_host(int n, const double *h_coords, const double *h_box);
private:
    int N_;
    std::shared_ptr<DeviceBuffer<unsigned int>> lut_; // shared by every sorter in the process (constant table)
    DeviceBuffer<unsigned int> keys_in_, keys_out_, vals_in_;
    DeviceBuffer<char> temp_;
};

template <typename Real> class Neighborlist {
public:
    explicit Neighborlist(int N);
    void set_row_idxs(std::vector<unsigned int> row_idxs);
    void reset_row_idxs();
    void resize(int size);
    // device-side index management used by the potentials (null idx arrays mean "contiguous")
    void set_contiguous_split(int NR, int NC); // rows = [0,NR), cols = [NR, NR+NC) of the gathered set
    void set_all_pairs(int K);

    unsigned int num_tile_ixns();
    int max_ixn_count() const;
    int get_num_row_idxs() const { return NR_; }
    std::vector<std::vector<int>> get_nblist_host(int N, const double *h_coords, const double *h_box, double cutoff);
    void compute_block_bounds_host(int N, const double *h_coords, const double *h_box, double *h_ctr, double *h_ext);

    // Enqueue bounds + tile build.  Exactly one of d_coords / d_xw is given; `flag` (nullable) gates all work on
    // the device.
    struct Snapshot { // where to record "coordinates/box at build time" (fused into the bounds kernel)
        Vec4<Real> *xw_build; // sorted-order packed coordinates at build time
        double *box_build;
        // all-pairs layout only: column-block bounds were already written to col_ctr()/col_ext() and the tile counter
        // reset by k_nb_prepare; the build kernel takes the snapshot itself and no bounds kernel is launched
        bool bounds_done = false;
        int slots = 0;
    };
    Real *col_ctr() { return d_col_ctr_.data; }
    Real *col_ext() { return d_col_ext_.data; }
    void build_device(const double *d_coords, const Vec4<Real> *d_xw, const double *d_box, double cutoff,
                      const unsigned int *flag, cudaStream_t stream, const Snapshot *snap = nullptr);

    const TileList &tiles() const { return tiles_; }
    // Host side of the measured sizing of the tile buffer: true when a build since the last call did not fit; the buffer
    // has then been grown (to twice what was needed) and the caller must rebuild and redo whatever it evaluated.
    bool recover_overflow();
    size_t capacity() const { return tiles_.capacity; }
    size_t worst_case_capacity() const { return worst_case_; }
    bool upper_triangular() const { return NR_ == N_ && NC_ == N_; }
    int num_row_blocks() const { return ceil_div(NR_, TILE); }
    int num_col_blocks() const { return ceil_div(NC_, TILE); }

private:
    void set_capacity(size_t tiles);
    size_t worst_case_ = 0;
    const int max_size_;
    int N_, NC_, NR_;
    bool contiguous_ = true;
    int row_base_ = 0, col_base_ = 0;
    DeviceBuffer<unsigned int> d_row_idxs_, d_col_idxs_;
    DeviceBuffer<Real> d_row_ctr_, d_row_ext_, d_col_ctr_, d_col_ext_;
    DeviceBuffer<unsigned int> d_count_, d_overflow_, d_cols_;
    DeviceBuffer<int> d_rows_;
    TileList tiles_;
};

// ---------------------------------------------------------------------------------------------------------------
// Shared machinery of the two tile-list nonbonded potentials.
template <typename Real> class NonbondedTiled : public Potential {
public:
    NonbondedTiled(int N, double beta, double cutoff, bool disable_hilbert, double nblist_padding, int steps_per_sort);
    void du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const override;
    int capturable_steps() const override {
        if (timing_) {
            return 0; // per-launch events are recorded eagerly
        }
        const long long r = steps_since_last_sort_ % steps_per_sort_;
        return r == 0 ? 0 : static_cast<int>(steps_per_sort_ - r);
    }
    // Measurement hook for bench.py: bracket every tile-kernel launch with CUDA events on its launch stream.
    void set_kernel_timing(bool on);
    std::vector<float> drain_kernel_times(); // milliseconds per launch since the last drain (synchronises)
    void advance(int n) override { steps_since_last_sort_ += n; }
    bool recover_overflow() override;
    size_t tile_capacity() const { return nblist_.capacity(); }
    size_t tile_worst_case() const { return nblist_.worst_case_capacity(); }
    double get_cutoff() const { return cutoff_; }
    double get_beta() const { return beta_; }
    double get_nblist_padding() const { return nblist_padding_; }
    unsigned int num_tiles();
    unsigned int num_rebuilds(); // neighbour-list builds since construction (device counter; synchronises)

protected:
    const int N_;
    int K_ = 0;  // gathered atoms
    int NR_ = 0; // row slots [0, NR_)
    const double beta_, cutoff_, nblist_padding_;
    const bool disable_hilbert_;
    const int steps_per_sort_;
    long long steps_since_last_sort_ = 0;
    bool force_rebuild_ = true;
    bool skip_prepare_once_ = false; // the integrator ran this evaluation's prepare pass (fused_prepare_hook)
    bool timing_ = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events_;
    size_t timing_used_ = 0;

    DeviceBuffer<unsigned int> d_perm_;
    DeviceBuffer<Vec4<Real>> d_xw_, d_qse_;
    DeviceBuffer<Vec4<Real>> d_xw_build_;
    DeviceBuffer<double> d_box_build_;
    DeviceBuffer<unsigned int> d_flags_;
    DeviceBuffer<i128> d_partials_;
    DeviceBuffer<unsigned int> d_ticket_;
    Neighborlist<Real> nblist_;
    std::unique_ptr<HilbertSort> hilbert_;

    bool needs_sort() const { return steps_since_last_sort_ % steps_per_sort_ == 0; }
    virtual void sort(const double *d_x, const double *d_box, cudaStream_t stream) = 0;
    void run(int N, const double *d_x, const double *d_p, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
             cudaStream_t stream);
    int Kpad() const { return round_up(N_, TILE); }
};

template <typename Real> class NonbondedAllPairs : public NonbondedTiled<Real> {
public:
    NonbondedAllPairs(int N, double beta, double cutoff, const std::optional<std::set<int>> &atom_idxs,
                      bool disable_hilbert_sort, double nblist_padding);
    void execute_device(int, int, const double *, const double *, const double *, u64 *, u64 *, i128 *, cudaStream_t) override;
    void set_atom_idxs(const std::vector<int> &atom_idxs);
    std::vector<int> get_atom_idxs();
    int get_num_atom_idxs() const { return this->K_; }
    bool fused_prepare_hook(FusedPrepareHook &hook) override;
    void skip_next_prepare() override { this->skip_prepare_once_ = true; }
    void cancel_skip_prepare() override { this->skip_prepare_once_ = false; }
protected:
    void sort(const double *d_x, const double *d_box, cudaStream_t stream) override;
private:
    DeviceBuffer<unsigned int> d_atom_idxs_;
    DeviceBuffer<unsigned int> d_other_idxs_; // atoms NOT in the set (the integrator's fused pass covers them too)
    int n_others_ = 0;
};

template <typename Real> class NonbondedInteractionGroup : public NonbondedTiled<Real> {
public:
    NonbondedInteractionGroup(int N, const std::vector<int> &row_atom_idxs, const std::vector<int> &col_atom_idxs,
                              double beta, double cutoff, bool disable_hilbert_sort, double nblist_padding);
    void execute_device(int, int, const double *, const double *, const double *, u64 *, u64 *, i128 *, cudaStream_t) override;
    void set_atom_idxs(const std::vector<int> &row_atom_idxs, const std::vector<int> &col_atom_idxs);
protected:
    void sort(const double *d_x, const double *d_box, cudaStream_t stream) override;
private:
    int NC_ = 0;
    DeviceBuffer<unsigned int> d_row_atom_idxs_, d_col_atom_idxs_;
    static void validate_idxs(int N, const std::vector<int> &rows, const std::vector<int> &cols, bool allow_empty);
};

template <typename Real, bool Negated> class NonbondedPairList : public Potential {
public:
    NonbondedPairList(const std::vector<int> &pair_idxs, const std::vector<double> &scales, double beta, double cutoff);
    void execute_device(int, int, const double *, const double *, const double *, u64 *, u64 *, i128 *, cudaStream_t) override;
    void du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const override;
private:
    int M_;
    double beta_, cutoff_;
    DeviceBuffer<int> d_pair_idxs_;
    DeviceBuffer<double> d_scales_;
    DeviceBuffer<i128> d_partials_;
    DeviceBuffer<unsigned int> d_ticket_;
};

void verify_atom_idxs(int N, const std::vector<int> &atom_idxs, bool allow_empty = false);
void nonbonded_du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out);

// ---------------------------------------------------------------------------------------------------------------
// What Context drives (reference integrator.hpp:9-39).  graph_offset / fusable: see LangevinIntegrator::step_fwd.
class Integrator {
public:
    virtual ~Integrator() {}
    virtual int num_atoms() const = 0;
    virtual void step_fwd(std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box,
                          unsigned int *d_idxs, cudaStream_t stream, int graph_offset = -1, Potential *fusable = nullptr) = 0;
    // bracket a run of steps (velocity Verlet's half kicks; nothing for Langevin)
    virtual void initialize(std::vector<std::shared_ptr<BoundPotential>> &, double *, double *, double *, unsigned int *, cudaStream_t) {}
    virtual void finalize(std::vector<std::shared_ptr<BoundPotential>> &, double *, double *, double *, unsigned int *, cudaStream_t) {}
    // CUDA-graph blocks of steps: only an integrator whose step has no host-side state beyond these counters
    virtual bool graph_capable() const { return false; }
    virtual void publish_step_base(cudaStream_t) {}
    virtual long long step_count() const { return 0; }
    virtual void advance(int) {}
    virtual void set_step(long long) {}
};

class LangevinIntegrator : public Integrator {
public:
    LangevinIntegrator(int N, const double *masses, double temperature, double dt, double friction, int seed);
    int num_atoms() const override { return N_; }
    double get_temperature() const { return temperature_; }
    bool graph_capable() const override { return true; }
    // One force evaluation + BAOAB update on `stream` (reference langevin_integrator.cu:55-88)
    // graph_offset >= 0: the call is being captured into a CUDA graph; the noise counter is then
    // *d_step_base + graph_offset so that replays draw fresh noise.
    // fusable: a potential whose next prepare pass this step's update kernel takes over (Potential::fused_prepare_hook), or
    // null
    void step_fwd(std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box,
                  unsigned int *d_idxs, cudaStream_t stream, int graph_offset = -1, Potential *fusable = nullptr) override;
    void publish_step_base(cudaStream_t stream) override; // *d_step_base = step_
    void set_external_noise(const float *h_noise); // tests: N x 3 normals reused every step; nullptr restores Philox
    long long step_count() const override { return step_; }
    void advance(int n) override { step_ += n; }
    // reposition the counter-based noise stream (graph replays read the counter from the device: no re-capture needed)
    void set_step(long long step) override { step_ = step; }
private:
    int N_;
    double temperature_;
    float dt_;
    float ca_;
    unsigned long long seed_;
    long long step_ = 0;
    DeviceBuffer<float> d_cbs_, d_ccs_, d_noise_;
    bool external_noise_ = false;
    DeviceBuffer<u64> d_du_dx_;
    DeviceBuffer<unsigned long long> d_step_base_;
    StreamFan fan_;
    friend class Context;
};

// Velocity Verlet in f64 (reference verlet_integrator.cu:10-110): initialize() = half kick + drift, step_fwd() = kick +
// drift, finalize() = half kick; cbs = -dt / mass is the caller's (lib/__init__.py:25-37).
class VelocityVerletIntegrator : public Integrator {
public:
    VelocityVerletIntegrator(int N, double dt, const double *h_cbs);
    int num_atoms() const override { return N_; }
    void step_fwd(std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box,
                  unsigned int *d_idxs, cudaStream_t stream, int graph_offset = -1, Potential *fusable = nullptr) override;
    void initialize(std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box,
                    unsigned int *d_idxs, cudaStream_t stream) override;
    void finalize(std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box,
                  unsigned int *d_idxs, cudaStream_t stream) override;
private:
    int N_;
    double dt_;
    bool initialized_ = false;
    DeviceBuffer<double> d_cbs_;
    DeviceBuffer<u64> d_du_dx_;
    StreamFan fan_;
    void forces_then(int mode, std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box,
                     unsigned int *d_idxs, cudaStream_t stream);
};

} // namespace tmb
#include "barostat.hpp"
namespace tmb {

// Local MD: potentials and free-atom selection for simulating a shell around one reference atom (local_md.cu)
class LocalMD {
public:
    LocalMD(int N, const std::vector<std::shared_ptr<BoundPotential>> &bps, bool freeze_reference, double temperature);
    ~LocalMD();
    const bool freeze_reference;
    const double temperature;
    // reference LocalMDPotentials::setup_from_idxs / setup_from_selection (local_md_potentials.cu:108-176)
    void setup_from_idxs(const double *d_x, const double *d_box, const std::vector<int> &local_idxs, int seed, double radius,
                         double k, cudaStream_t stream);
    void setup_from_selection(int reference_idx, const std::vector<int> &selection_idxs, double radius, double k, cudaStream_t stream);
    std::vector<std::shared_ptr<BoundPotential>> &potentials() { return active_; }
    unsigned int *free_idxs() { return d_selected_.data; } // [N]: atom index if free, N if frozen (the integrator's idxs)
    const std::vector<unsigned int> &selected_host() const { return h_selected_; }
    int num_free() const { return num_free_; }
    void reset(); // give the NonbondedAllPairs its original atoms back (local_md_potentials.cu:330-335)
private:
    int N_;
    std::vector<std::shared_ptr<BoundPotential>> base_, active_;
    std::shared_ptr<Potential> all_pairs_;
    std::vector<int> original_idxs_;
    std::shared_ptr<BoundPotential> ixn_group_, free_restraint_, frozen_restraint_;
    DeviceBuffer<float> d_uniforms_;
    DeviceBuffer<unsigned int> d_selected_;
    std::vector<unsigned int> h_selected_;
    ::curandGenerator_st *rng_ = nullptr; // curandGenerator_t
    int num_free_ = 0;
    bool modified_ = false;
    void set_all_pairs_idxs(const std::vector<int> &idxs);
    void configure(unsigned int reference_idx, double radius, double k);
};

class Context {
public:
    Context(int N, const double *x0, const double *v0, const double *box0, std::shared_ptr<Integrator> intg,
            std::vector<std::shared_ptr<BoundPotential>> bps, std::vector<std::shared_ptr<Mover>> movers = {});
    ~Context();
    int num_atoms() const { return N_; }
    void step();
    void initialize(); // reference context.cu:250-260: the integrator's initialize / finalize on the context's state
    void finalize();
    // n_samples frames of x/box are written to h_x/h_box (reference context.cu:216-242)
    void multiple_steps(int n_steps, int n_samples, double *h_x, double *h_box);
    // local MD (reference context.cu:90-214); movers do not run during local steps
    void setup_local_md(double temperature, bool freeze_reference);
    void multiple_steps_local(int n_steps, const std::vector<int> &local_idxs, int n_samples, double radius, double k, int seed,
                              double *h_x, double *h_box);
    void multiple_steps_local_selection(int n_steps, int reference_idx, const std::vector<int> &selection_idxs, int n_samples,
                                        double radius, double k, double *h_x, double *h_box);
    const LocalMD *local_md() const { return local_md_.get(); }
    void set_x_t(const double *h);
    void set_v_t(const double *h);
    void set_box(const double *h);
    void get_x_t(double *h) const;
    void get_v_t(double *h) const;
    void get_box(double *h) const;
    double *d_x() { return d_x_.data; }
    double *d_v() { return d_v_.data; }
    double *d_box() { return d_box_.data; }
    std::shared_ptr<Integrator> get_integrator() const { return intg_; }
    const std::vector<std::shared_ptr<BoundPotential>> &get_potentials() const { return bps_; }
    const std::vector<std::shared_ptr<Mover>> &get_movers() const { return movers_; }
    std::shared_ptr<MonteCarloBarostat<float>> get_barostat() const; // reference context.cu:311-320
    void set_use_graphs(bool on) { use_graphs_ = on; }
    // Run the MD loop on a caller-owned stream (e.g. torch's current stream) instead of the context's own one.
    void set_stream(cudaStream_t s);

private:
    int N_;
    DeviceBuffer<double> d_x_, d_v_, d_box_;
    std::shared_ptr<Integrator> intg_;
    double langevin_temperature() const; // reference context.cu:80-88
    std::vector<std::shared_ptr<BoundPotential>> bps_;
    std::vector<std::shared_ptr<Mover>> movers_;
    std::vector<double> nb_cutoffs_with_padding_;
    bool use_graphs_ = true;
    cudaStream_t stream_ = nullptr;      // owned, non-blocking
    cudaStream_t user_stream_ = nullptr; // optional override
    cudaStream_t active_stream() const { return user_stream_ ? user_stream_ : stream_; }
    // CUDA-graph replay of blocks of steps
    cudaStream_t graph_stream_ = nullptr;
    cudaGraphExec_t graph_exec_ = nullptr;
    long long graph_kernels_ = 0;
    long long graph_generation_ = -1; // g_launch_generation at capture time
    void run_steps(int n, cudaStream_t stream);
    void eager_step(cudaStream_t stream); // integrator step, then every mover (reference context.cu:261-277)
    void verify_frame(const double *h_x, const double *h_box) const;
    void check_list_overflow();
    void destroy_graph();
    std::unique_ptr<LocalMD> local_md_;
    void run_local_steps(int n_steps, int n_samples, double *h_x, double *h_box, cudaStream_t stream);
};

// x2 moved rigidly (optimal rotation, no reflection; centroid onto x1's) onto x1: reference rmsd_align.cpp:11-61 (host, f64)
void rmsd_align_host(int N, const double *x1, const double *x2, double *x2_aligned);

void collect_nonbonded_cutoffs(const std::shared_ptr<Potential> &pot, std::vector<double> &out);

} // namespace tmb
