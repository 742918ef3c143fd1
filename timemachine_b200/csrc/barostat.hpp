// Movers that Context runs after every integrator step, and the Monte Carlo barostat (NPT).
//
// Reference: timemachine/cpp/src/mover.hpp:11-50, barostat.hpp:15-93, barostat.cu:18-262, kernels/k_barostat.cuh:7-191
// (itself after OpenMM's MonteCarloBarostat).  SURVEY.md §8f rank 1: the first row after the force + integration path.
//
// What is kept exactly: the proposal (molecule centroids scaled about the box centre, molecules moved rigidly and
// re-imaged into the scaled home box), the float arithmetic of every quantity that feeds the accept test, the
// acceptance rule, the adaptive volume-scale rule, and the random stream (cuRAND XORWOW, same seed, same batching of
// 2000 uniforms per 1000 moves) - so a run with the same seed takes the same accept/reject decisions and produces the
// same coordinates as the reference.
//
// What is different: one proposal kernel (warp per molecule: fixed-point centroid by shuffles, rescale, write; no
// memset, no atomics, no separate setup / copy / centroid / rescale launches) and one decision kernel that also sums
// the per-potential energies (no CUB reductions); the potentials' energy-only kernel variants provide U.
//
// Included by potential.hpp (after DeviceBuffer, StreamFan and BoundPotential are defined); not a stand-alone header.
#pragma once

#include <array>
#include <memory>
#include <vector>

typedef struct curandGenerator_st *curandGenerator_t;

namespace tmb {

class Mover {
public:
    virtual ~Mover() {}
    // reference mover.hpp:25-41
    void set_step(int step);
    void set_interval(int interval);
    int get_interval() const { return interval_; }
    // May modify d_x and d_box.  Counts calls: acts on every interval-th one.
    virtual void move(int N, double *d_x, double *d_box, cudaStream_t stream) = 0;
    virtual std::array<std::vector<double>, 2> move_host(int N, const double *h_x, const double *h_box);
    // Number of upcoming move() calls that are guaranteed to do nothing but count (Context replays that many plain
    // MD steps from a CUDA graph and calls skip()).
    int idle_steps() const { return interval_ - 1 - (step_ % interval_); }
    void skip(int n) { step_ += n; }

protected:
    explicit Mover(int interval) : interval_(interval), step_(0) {}
    int interval_;
    int step_;
};

template <typename Real> class MonteCarloBarostat : public Mover {
public:
    MonteCarloBarostat(
        int N, double pressure /*bar*/, double temperature /*K*/, std::vector<std::vector<int>> group_idxs, int interval,
        std::vector<std::shared_ptr<BoundPotential>> bps, int seed, bool adaptive_scaling_enabled,
        double initial_volume_scale_factor);
    ~MonteCarloBarostat() override;

    void move(int N, double *d_x, double *d_box, cudaStream_t stream) override;

    double get_volume_scale_factor();
    void set_volume_scale_factor(double volume_scale_factor);
    void set_pressure(double pressure);
    void set_adaptive_scaling(bool on) { adaptive_ = on; }
    bool get_adaptive_scaling() const { return adaptive_; }
    // introspection (tests): the two uniforms of the most recent attempted move, and the attempt/accept counters
    std::array<float, 2> last_uniforms();
    std::array<int, 2> counters();

private:
    const int N_;
    bool adaptive_;
    std::vector<std::shared_ptr<BoundPotential>> bps_;
    Real pressure_;
    const Real temperature_;
    const int seed_;
    int num_mols_ = 0;
    int num_grouped_ = 0;
    int num_ungrouped_ = 0;
    int last_offset_ = 0;

    curandGenerator_t rng_ = nullptr;
    DeviceBuffer<Real> d_rand_;
    DeviceBuffer<int> d_counters_;       // {attempted, accepted}
    DeviceBuffer<i128> d_u_init_, d_u_final_; // one entry per bound potential
    DeviceBuffer<Real> d_volume_;        // {volume, volume_delta}
    DeviceBuffer<double> d_volume_scale_;
    DeviceBuffer<double> d_x_proposed_, d_box_proposed_;
    DeviceBuffer<int> d_atom_idxs_, d_mol_offsets_, d_ungrouped_;
    StreamFan fan_;

    void reset_counters();
    void energies(const double *d_x, const double *d_box, i128 *d_u, cudaStream_t stream);
};

void verify_group_idxs(int N, const std::vector<std::vector<int>> &group_idxs);

} // namespace tmb
