// LangevinIntegrator and Context: the MD loop lives in C++ (reference langevin_integrator.cu:13-111, context.cu:28-322).
//
// B200 design: the steady-state step has no host decision in it (the neighbour-list rebuild is decided on the device),
// so blocks of GRAPH_STEPS steps are captured once into a CUDA graph and replayed; the only eager steps are the ones
// that re-sort atoms along the Hilbert curve (every 100th evaluation) and the remainders.
#include "potential.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace tmb {

constexpr double BOLTZ = 0.008314462618; // kJ/mol/K, reference timemachine/cpp/src/constants.hpp:5

__global__ void k_set_u64(unsigned long long *p, unsigned long long v) { *p = v; }

LangevinIntegrator::LangevinIntegrator(int N, const double *masses, double temperature, double dt, double friction, int seed)
    : N_(N), temperature_(temperature), dt_(static_cast<float>(dt)), seed_(static_cast<unsigned long long>(static_cast<long long>(seed))),
      d_cbs_(N), d_ccs_(N), d_du_dx_(static_cast<size_t>(N) * 3), d_step_base_(1) {
    ca_ = static_cast<float>(std::exp(-friction * dt));
    const double kT = BOLTZ * temperature;
    const double ccs_adjustment = std::sqrt(1 - std::exp(-2 * friction * dt));
    std::vector<float> h_cbs(N), h_ccs(N);
    for (int i = 0; i < N; i++) {
        // dt_ is already rounded to f32 here, exactly as the reference does (langevin_integrator.cu:27)
        h_cbs[i] = static_cast<float>(dt_ / masses[i]);
        h_ccs[i] = static_cast<float>(ccs_adjustment * std::sqrt(kT / masses[i]));
    }
    d_cbs_.copy_from(h_cbs.data());
    d_ccs_.copy_from(h_ccs.data());
    d_du_dx_.zero(); // the update kernel re-zeroes the forces every step
    d_step_base_.zero();
    TMB_CUDA(cudaDeviceSynchronize());
}

void LangevinIntegrator::set_external_noise(const float *h_noise) {
    bump_launch_generation(); // the noise pointer (or its absence) is an argument of the captured k_baoab launches
    if (h_noise == nullptr) {
        external_noise_ = false;
        return;
    }
    d_noise_.realloc(static_cast<size_t>(N_) * 3);
    d_noise_.copy_from(h_noise);
    external_noise_ = true;
}

void LangevinIntegrator::step_fwd(
    std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box, unsigned int *d_idxs,
    cudaStream_t stream, int graph_offset, Potential *fusable) {
    const int n = static_cast<int>(bps.size());
    // one stream per bound potential, forked from and joined back into `stream`
    // (reference streamed_potential_runner.cu:10-29)
    for (int i = 0; i < n; i++) {
        cudaStream_t s = n > 1 ? fan_.fork(i, stream) : stream;
        bps[i]->execute_device(N_, d_x, d_box, d_du_dx_.data, nullptr, nullptr, s);
    }
    if (n > 1) {
        for (int i = 0; i < n; i++) {
            fan_.join(i, stream);
        }
    }
    BaoabArgs a;
    a.N = N_;
    a.ca = ca_;
    a.idxs = d_idxs;
    a.cbs = d_cbs_.data;
    a.ccs = d_ccs_.data;
    a.noise = external_noise_ ? d_noise_.data : nullptr;
    a.seed = seed_;
    if (graph_offset >= 0) {
        a.step = static_cast<unsigned long long>(graph_offset);
        a.step_base = d_step_base_.data;
    } else {
        a.step = static_cast<unsigned long long>(step_);
        a.step_base = nullptr;
    }
    a.x = d_x;
    a.v = d_v;
    a.du_dx = d_du_dx_.data;
    a.dt = dt_;
    FusedPrepareHook hook;
    if (fusable != nullptr && d_idxs == nullptr && fusable->fused_prepare_hook(hook)) {
        // integrate in the potential's sorted order and leave its next prepare pass done (k_baoab_prepare)
        if (hook.is_double) {
            hook.f64.box = d_box;
            launch_baoab_prepare<double>(a, hook.f64, stream);
        } else {
            hook.f32.box = d_box;
            launch_baoab_prepare<float>(a, hook.f32, stream);
        }
        fusable->skip_next_prepare();
    } else {
        launch_baoab(a, stream);
    }
    step_++;
}

// ---------------------------------------------------------------------------------------------------------------
VelocityVerletIntegrator::VelocityVerletIntegrator(int N, double dt, const double *h_cbs)
    : N_(N), dt_(dt), d_cbs_(N), d_du_dx_(static_cast<size_t>(N) * 3) {
    d_cbs_.copy_from(h_cbs);
    d_du_dx_.zero(); // the update kernels re-zero the forces
    TMB_CUDA(cudaDeviceSynchronize());
}

void VelocityVerletIntegrator::forces_then(
    int mode, std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box, unsigned int *d_idxs,
    cudaStream_t stream) {
    const int n = static_cast<int>(bps.size());
    for (int i = 0; i < n; i++) {
        cudaStream_t s = n > 1 ? fan_.fork(i, stream) : stream;
        bps[i]->execute_device(N_, d_x, d_box, d_du_dx_.data, nullptr, nullptr, s);
    }
    if (n > 1) {
        for (int i = 0; i < n; i++) {
            fan_.join(i, stream);
        }
    }
    VerletArgs a;
    a.N = N_;
    a.idxs = d_idxs;
    a.cbs = d_cbs_.data;
    a.x = d_x;
    a.v = d_v;
    a.du_dx = d_du_dx_.data;
    a.dt = dt_;
    launch_velocity_verlet(a, mode, stream);
}

void VelocityVerletIntegrator::step_fwd(
    std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box, unsigned int *d_idxs,
    cudaStream_t stream, int, Potential *) {
    forces_then(0, bps, d_x, d_v, d_box, d_idxs, stream);
}

void VelocityVerletIntegrator::initialize(
    std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box, unsigned int *d_idxs,
    cudaStream_t stream) {
    if (initialized_) {
        throw std::runtime_error("initialized twice");
    }
    forces_then(1, bps, d_x, d_v, d_box, d_idxs, stream);
    initialized_ = true;
}

void VelocityVerletIntegrator::finalize(
    std::vector<std::shared_ptr<BoundPotential>> &bps, double *d_x, double *d_v, double *d_box, unsigned int *d_idxs,
    cudaStream_t stream) {
    if (!initialized_) {
        throw std::runtime_error("not initialized");
    }
    forces_then(2, bps, d_x, d_v, d_box, d_idxs, stream);
    initialized_ = false;
}

void LangevinIntegrator::publish_step_base(cudaStream_t stream) {
    TMB_LAUNCH(k_set_u64, 1, 1, 0, stream, d_step_base_.data, static_cast<unsigned long long>(step_));
}

// ---------------------------------------------------------------------------------------------------------------
constexpr int GRAPH_STEPS = 10;

// The first potential (depth first) that can hand its prepare pass to the integrator: the all-pairs term of the system.
static Potential *find_fusable(const std::shared_ptr<Potential> &pot) {
    if (auto s = std::dynamic_pointer_cast<SummedPotential>(pot)) {
        for (auto &c : s->get_potentials()) {
            if (Potential *p = find_fusable(c)) {
                return p;
            }
        }
        return nullptr;
    }
    if (auto f = std::dynamic_pointer_cast<FanoutSummedPotential>(pot)) {
        for (auto &c : f->get_potentials()) {
            if (Potential *p = find_fusable(c)) {
                return p;
            }
        }
        return nullptr;
    }
    if (std::dynamic_pointer_cast<NonbondedAllPairs<float>>(pot) || std::dynamic_pointer_cast<NonbondedAllPairs<double>>(pot)) {
        return pot.get();
    }
    return nullptr;
}

static bool fuse_prepare_enabled() {
    static const bool on = [] {
        // Measured and NOT adopted (profiles/r2_summary.md section 7): integrating in the potential's sorted order makes the
        // 72 B per atom of x / v / du_dx traffic a gather (32-byte sectors, 24-byte rows), and the fused launch costs more than
        // the two coalesced ones it replaces: 145.2 against 139.2 us per MD step.  TMB_FUSE_PREPARE=1 selects it for A/B runs.
        const char *e = std::getenv("TMB_FUSE_PREPARE");
        return e != nullptr && e[0] == '1';
    }();
    return on;
}

Context::Context(
    int N, const double *x0, const double *v0, const double *box0, std::shared_ptr<Integrator> intg,
    std::vector<std::shared_ptr<BoundPotential>> bps, std::vector<std::shared_ptr<Mover>> movers)
    : N_(N), d_x_(static_cast<size_t>(N) * 3), d_v_(static_cast<size_t>(N) * 3), d_box_(9), intg_(std::move(intg)),
      bps_(std::move(bps)), movers_(std::move(movers)) {
    if (intg_->num_atoms() != N) {
        throw std::runtime_error("integrator N != x0 N");
    }
    d_x_.copy_from(x0);
    d_v_.copy_from(v0);
    d_box_.copy_from(box0);
    for (auto &bp : bps_) {
        collect_nonbonded_cutoffs(bp->potential, nb_cutoffs_with_padding_);
    }
    TMB_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
}

Context::~Context() {
    destroy_graph();
    if (stream_) {
        cudaStreamDestroy(stream_);
    }
}

void Context::destroy_graph() {
    if (graph_exec_) {
        cudaGraphExecDestroy(graph_exec_);
        graph_exec_ = nullptr;
    }
}

void Context::set_stream(cudaStream_t s) {
    destroy_graph();
    user_stream_ = s;
}

void Context::run_steps(int n, cudaStream_t stream) {
    int remaining = n;
    while (remaining > 0) {
        int cap = 1 << 30;
        for (auto &bp : bps_) {
            cap = std::min(cap, bp->potential->capturable_steps());
        }
        for (auto &m : movers_) {
            cap = std::min(cap, m->idle_steps()); // a graph block must not span a step on which a mover acts
        }
        if (use_graphs_ && intg_->graph_capable() && cap >= GRAPH_STEPS && remaining >= GRAPH_STEPS) {
            intg_->publish_step_base(stream);
            const long long generation = g_launch_generation.load();
            if (graph_exec_ == nullptr || graph_stream_ != stream || graph_generation_ != generation) {
                destroy_graph();
                cudaGraph_t graph = nullptr;
                const long long launches_before = g_kernel_launches.load();
                const long long intg_step_before = intg_->step_count();
                TMB_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed));
                int captured = 0;
                // inside the block the integrator of step s also runs the all-pairs potential's prepare pass of step
                // s + 1 (the block's last step does not: what follows the block is not known here)
                Potential *fusable = nullptr;
                if (fuse_prepare_enabled()) {
                    for (auto &bp : bps_) {
                        if ((fusable = find_fusable(bp->potential)) != nullptr) {
                            break;
                        }
                    }
                }
                try {
                    for (int s = 0; s < GRAPH_STEPS; s++) {
                        intg_->step_fwd(
                            bps_, d_x_.data, d_v_.data, d_box_.data, nullptr, stream, s, s + 1 < GRAPH_STEPS ? fusable : nullptr);
                        captured++;
                    }
                    TMB_CUDA(cudaStreamEndCapture(stream, &graph));
                    TMB_CUDA(cudaGraphInstantiate(&graph_exec_, graph, 0));
                } catch (...) {
                    // nothing ran: end the capture (it is invalid if a forked stream was left un-joined), put the host-side
                    // cadence counters back where they were so that the context stays in step with the device
                    cudaGraph_t dead = nullptr;
                    cudaStreamEndCapture(stream, &dead);
                    if (dead) {
                        cudaGraphDestroy(dead);
                    }
                    if (graph && graph != dead) {
                        cudaGraphDestroy(graph);
                    }
                    cudaGetLastError();
                    graph_exec_ = nullptr;
                    if (fusable != nullptr) {
                        fusable->cancel_skip_prepare();
                    }
                    for (auto &bp : bps_) {
                        bp->potential->advance(-captured);
                    }
                    intg_->set_step(intg_step_before);
                    g_kernel_launches.store(launches_before);
                    throw;
                }
                TMB_CUDA(cudaGraphDestroy(graph));
                graph_stream_ = stream;
                graph_generation_ = generation;
                // the capture pass already advanced the host-side counters of the potentials and the integrator, and
                // counted its kernels once (they run on the first launch below)
                graph_kernels_ = g_kernel_launches.load() - launches_before;
            } else {
                for (auto &bp : bps_) {
                    bp->potential->advance(GRAPH_STEPS);
                }
                intg_->advance(GRAPH_STEPS);
                g_kernel_launches.fetch_add(graph_kernels_, std::memory_order_relaxed);
            }
            TMB_CUDA(cudaGraphLaunch(graph_exec_, stream));
            for (auto &m : movers_) {
                m->skip(GRAPH_STEPS);
            }
            remaining -= GRAPH_STEPS;
        } else {
            eager_step(stream);
            remaining -= 1;
        }
    }
}

void Context::eager_step(cudaStream_t stream) {
    intg_->step_fwd(bps_, d_x_.data, d_v_.data, d_box_.data, nullptr, stream, -1);
    for (auto &m : movers_) {
        m->move(N_, d_x_.data, d_box_.data, stream); // may modify coordinates and box
    }
}

void Context::step() {
    cudaStream_t stream = active_stream();
    eager_step(stream);
    TMB_CUDA(cudaStreamSynchronize(stream));
    check_list_overflow();
}

void Context::initialize() {
    cudaStream_t stream = active_stream();
    intg_->initialize(bps_, d_x_.data, d_v_.data, d_box_.data, nullptr, stream);
    TMB_CUDA(cudaStreamSynchronize(stream));
    check_list_overflow();
}

void Context::finalize() {
    cudaStream_t stream = active_stream();
    intg_->finalize(bps_, d_x_.data, d_v_.data, d_box_.data, nullptr, stream);
    TMB_CUDA(cudaStreamSynchronize(stream));
    check_list_overflow();
}

double Context::langevin_temperature() const {
    if (auto langevin = std::dynamic_pointer_cast<LangevinIntegrator>(intg_)) {
        return langevin->get_temperature();
    }
    throw std::runtime_error("integrator must be LangevinIntegrator.");
}

// The tile-list buffers are sized from measured counts (4x the liquid-density count, neighborlist.cu).  Steps cannot be
// redone the way a single evaluation can, so a build that ran out of room during them is an error: the buffers have been
// grown by the time this throws, the caller restores its state (set_x_t / set_v_t / set_box) and runs again.
void Context::check_list_overflow() {
    bool any = false;
    for (auto &bp : bps_) {
        any = bp->potential->recover_overflow() || any;
    }
    if (any) {
        destroy_graph();
        throw std::runtime_error(
            "neighborlist tile buffer overflow during MD steps: the buffer has been enlarged, but the steps of this call "
            "used an incomplete interaction list - restore the state and run them again");
    }
}

std::shared_ptr<MonteCarloBarostat<float>> Context::get_barostat() const {
    for (auto &m : movers_) {
        if (auto b = std::dynamic_pointer_cast<MonteCarloBarostat<float>>(m)) {
            return b;
        }
    }
    return nullptr;
}

void Context::verify_frame(const double *h_x, const double *h_box) const {
    if (nb_cutoffs_with_padding_.empty()) {
        return;
    }
    for (double c : nb_cutoffs_with_padding_) {
        for (int i = 0; i < 3; i++) {
            if (h_box[i * 3 + i] < 2 * c) {
                throw std::runtime_error(
                    "cutoff with padding is more than half of the box width, neighborlist is no longer reliable");
            }
        }
    }
    const double max_box = std::max(h_box[0], std::max(h_box[4], h_box[8]));
    const auto mm = std::minmax_element(h_x, h_x + static_cast<size_t>(N_) * 3);
    if (max_box * 100.0 < (*mm.second - *mm.first)) {
        throw std::runtime_error(
            "simulation unstable: dimensions of coordinates two orders of magnitude larger than max box dimension");
    }
}

void Context::multiple_steps(int n_steps, int n_samples, double *h_x, double *h_box) {
    if (n_samples < 0) {
        throw std::runtime_error("n_samples < 0");
    }
    const int interval = n_samples > 0 ? n_steps / n_samples : n_steps + 1;
    cudaStream_t stream = active_stream();
    intg_->initialize(bps_, d_x_.data, d_v_.data, d_box_.data, nullptr, stream); // reference context.cu:227
    int done = 0;
    int stored = 0;
    while (done < n_steps) {
        const int next_store = (done / interval + 1) * interval;
        const int chunk = std::min(n_steps, next_store) - done;
        run_steps(chunk, stream);
        done += chunk;
        if (done % interval == 0 && stored < n_samples) {
            double *xp = h_x + static_cast<size_t>(stored) * N_ * 3;
            double *bp = h_box + static_cast<size_t>(stored) * 9;
            TMB_CUDA(cudaMemcpyAsync(xp, d_x_.data, d_x_.bytes(), cudaMemcpyDeviceToHost, stream));
            TMB_CUDA(cudaMemcpyAsync(bp, d_box_.data, d_box_.bytes(), cudaMemcpyDeviceToHost, stream));
            TMB_CUDA(cudaStreamSynchronize(stream));
            verify_frame(xp, bp);
            stored++;
        }
    }
    intg_->finalize(bps_, d_x_.data, d_v_.data, d_box_.data, nullptr, stream); // reference context.cu:239
    TMB_CUDA(cudaStreamSynchronize(stream));
    check_list_overflow();
}

// ---------------------------------------------------------------------------------------------------------------
// Local MD (reference context.cu:90-214)
void Context::setup_local_md(double temperature, bool freeze_reference) {
    if (local_md_ != nullptr) {
        if (local_md_->temperature != temperature || local_md_->freeze_reference != freeze_reference) {
            throw std::runtime_error(
                "local md configured with different parameters, current parameters: Temperature " +
                std::to_string(local_md_->temperature) + " Freeze Reference " + std::to_string(local_md_->freeze_reference));
        }
        return;
    }
    local_md_.reset(new LocalMD(N_, bps_, freeze_reference, temperature));
}

void Context::run_local_steps(int n_steps, int n_samples, double *h_x, double *h_box, cudaStream_t stream) {
    const int interval = n_samples > 0 ? n_steps / n_samples : n_steps + 1;
    // the all-pairs potential now holds other atoms: a graph captured for the full system must not be replayed later
    destroy_graph();
    std::vector<std::shared_ptr<BoundPotential>> &pots = local_md_->potentials();
    try {
        intg_->initialize(pots, d_x_.data, d_v_.data, d_box_.data, local_md_->free_idxs(), stream); // reference context.cu:141
        for (int i = 1; i <= n_steps; i++) {
            intg_->step_fwd(pots, d_x_.data, d_v_.data, d_box_.data, local_md_->free_idxs(), stream, -1);
            if (i % interval == 0) {
                double *xp = h_x + static_cast<size_t>(i / interval - 1) * N_ * 3;
                double *bp = h_box + static_cast<size_t>(i / interval - 1) * 9;
                TMB_CUDA(cudaMemcpyAsync(xp, d_x_.data, d_x_.bytes(), cudaMemcpyDeviceToHost, stream));
                TMB_CUDA(cudaMemcpyAsync(bp, d_box_.data, d_box_.bytes(), cudaMemcpyDeviceToHost, stream));
                TMB_CUDA(cudaStreamSynchronize(stream));
                verify_frame(xp, bp);
            }
        }
        intg_->finalize(pots, d_x_.data, d_v_.data, d_box_.data, local_md_->free_idxs(), stream); // reference context.cu:152
    } catch (...) {
        cudaStreamSynchronize(stream);
        local_md_->reset();
        throw;
    }
    TMB_CUDA(cudaStreamSynchronize(stream));
    bool local_overflow = false;
    for (auto &bp : pots) {
        local_overflow = bp->potential->recover_overflow() || local_overflow;
    }
    local_md_->reset();
    if (local_overflow) {
        throw std::runtime_error("neighborlist tile buffer overflow during local MD steps: buffer enlarged, run them again");
    }
    check_list_overflow();
}

void Context::multiple_steps_local(
    int n_steps, const std::vector<int> &local_idxs, int n_samples, double radius, double k, int seed, double *h_x, double *h_box) {
    if (n_samples < 0) {
        throw std::runtime_error("n_samples < 0");
    }
    if (local_md_ == nullptr) {
        setup_local_md(langevin_temperature(), true);
    }
    cudaStream_t stream = active_stream();
    local_md_->setup_from_idxs(d_x_.data, d_box_.data, local_idxs, seed, radius, k, stream);
    run_local_steps(n_steps, n_samples, h_x, h_box, stream);
}

void Context::multiple_steps_local_selection(
    int n_steps, int reference_idx, const std::vector<int> &selection_idxs, int n_samples, double radius, double k, double *h_x,
    double *h_box) {
    if (n_samples < 0) {
        throw std::runtime_error("n_samples < 0");
    }
    if (local_md_ == nullptr) {
        setup_local_md(langevin_temperature(), true);
    }
    cudaStream_t stream = active_stream();
    local_md_->setup_from_selection(reference_idx, selection_idxs, radius, k, stream);
    run_local_steps(n_steps, n_samples, h_x, h_box, stream);
}

void Context::set_x_t(const double *h) { d_x_.copy_from(h); }
void Context::set_v_t(const double *h) { d_v_.copy_from(h); }
void Context::set_box(const double *h) { d_box_.copy_from(h); }
void Context::get_x_t(double *h) const { d_x_.copy_to(h); }
void Context::get_v_t(double *h) const { d_v_.copy_to(h); }
void Context::get_box(double *h) const { d_box_.copy_to(h); }

} // namespace tmb
