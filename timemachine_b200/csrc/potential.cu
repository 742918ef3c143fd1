// Potential base-class host paths, BoundPotential, Summed/Fanout composition, stream fan-out.
// Reference: potential.cu:10-328, bound_potential.cu:6-147, summed_potential.cu:33-99, fanout_summed_potential.cu:23-70.
#include "potential.hpp"

#include <mutex>
#include "fixed_point.cuh"

#include <algorithm>
#include <numeric>

namespace tmb {

std::atomic<long long> g_kernel_launches{0};
std::atomic<long long> g_launch_generation{0};

static cudaStream_t g_main_stream = nullptr;
cudaStream_t main_stream() { return g_main_stream; }
void set_main_stream(cudaStream_t s) { g_main_stream = s; }

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0;
        TMB_CUDA(cudaGetDevice(&dev));
        TMB_CUDA(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
    }
    return cached;
}

// ---------------------------------------------------------------------------------------------------------------
StreamFan::~StreamFan() {
    for (auto s : streams_) {
        cudaStreamDestroy(s);
    }
    for (auto e : events_) {
        cudaEventDestroy(e);
    }
    if (parent_event_) {
        cudaEventDestroy(parent_event_);
    }
}

void StreamFan::ensure(int i) {
    if (parent_event_ == nullptr) {
        TMB_CUDA(cudaEventCreateWithFlags(&parent_event_, cudaEventDisableTiming));
    }
    while (static_cast<int>(streams_.size()) <= i) {
        cudaStream_t s;
        cudaEvent_t e;
        TMB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        TMB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        streams_.push_back(s);
        events_.push_back(e);
    }
}

cudaStream_t StreamFan::fork(int i, cudaStream_t parent) {
    ensure(i);
    TMB_CUDA(cudaEventRecord(parent_event_, parent));
    TMB_CUDA(cudaStreamWaitEvent(streams_[i], parent_event_, 0));
    return streams_[i];
}

void StreamFan::join(int i, cudaStream_t parent) {
    TMB_CUDA(cudaEventRecord(events_[i], streams_[i]));
    TMB_CUDA(cudaStreamWaitEvent(parent, events_[i], 0));
}

// ---------------------------------------------------------------------------------------------------------------
void Potential::du_dp_fixed_to_float(int, int P, const u64 *du_dp, double *out) const {
    for (int i = 0; i < P; i++) {
        out[i] = fixed_to_real<double>(du_dp[i]);
    }
}

// ---- cache of freed device blocks behind ScratchBuffer (potential.hpp) ------------------------------------------------
namespace {
struct ScratchCache {
    std::mutex mu;
    std::vector<std::pair<size_t, void *>> free_blocks; // (capacity, ptr)
    size_t cached_bytes = 0;
    ~ScratchCache() {
        for (auto &b : free_blocks) {
            cudaFree(b.second);
        }
    }
};
ScratchCache &scratch_cache() {
    static ScratchCache c;
    return c;
}
size_t scratch_capacity_for(size_t bytes) {
    // size classes: powers of two from 256 B, so that a block serves every request of its class
    size_t cap = 256;
    while (cap < bytes) {
        cap <<= 1;
    }
    return cap;
}
} // namespace

void *scratch_alloc(size_t bytes) {
    const size_t cap = scratch_capacity_for(bytes);
    ScratchCache &c = scratch_cache();
    {
        std::lock_guard<std::mutex> lock(c.mu);
        for (size_t i = 0; i < c.free_blocks.size(); i++) {
            if (c.free_blocks[i].first == cap) {
                void *p = c.free_blocks[i].second;
                c.free_blocks[i] = c.free_blocks.back();
                c.free_blocks.pop_back();
                c.cached_bytes -= cap;
                return p;
            }
        }
    }
    void *p = nullptr;
    TMB_CUDA(cudaMalloc(&p, cap));
    return p;
}

void scratch_free(void *ptr, size_t bytes) {
    const size_t cap = scratch_capacity_for(bytes);
    ScratchCache &c = scratch_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.cached_bytes + cap > (size_t(4) << 30)) { // keep at most 4 GiB parked
        cudaFree(ptr);
        return;
    }
    c.free_blocks.emplace_back(cap, ptr);
    c.cached_bytes += cap;
}

// After a synchronised host-level evaluation: did a neighbour-list build run out of room (Potential::recover_overflow)?
// Then the buffers have been grown and the evaluation is repeated from zeroed outputs.
static bool redo_after_overflow(Potential &pot, int attempt) {
    if (!pot.recover_overflow()) {
        return false;
    }
    if (attempt >= 8) {
        throw std::runtime_error("neighborlist tile buffer overflow");
    }
    return true;
}

void Potential::execute_host(
    int N, int P, const double *h_x, const double *h_p, const double *h_box, u64 *h_du_dx, u64 *h_du_dp, i128 *h_u) {
    ScratchBuffer<double> d_x(static_cast<size_t>(N) * D), d_box(D * D), d_p(P);
    d_x.copy_from(h_x);
    d_box.copy_from(h_box);
    if (P > 0) {
        d_p.copy_from(h_p);
    }
    cudaStream_t stream = main_stream();
    ScratchBuffer<u64> d_du_dx, d_du_dp;
    ScratchBuffer<i128> d_u;
    // the kernels accumulate: outputs must start at zero
    if (h_du_dx) {
        d_du_dx.realloc(static_cast<size_t>(N) * D);
    }
    if (h_du_dp) {
        d_du_dp.realloc(P);
    }
    if (h_u) {
        d_u.realloc(1);
    }
    for (int attempt = 0;; attempt++) {
        d_du_dx.zero(stream);
        d_du_dp.zero(stream);
        d_u.zero(stream);
        this->execute_device(
            N, P, d_x.data, P > 0 ? d_p.data : nullptr, d_box.data, h_du_dx ? d_du_dx.data : nullptr,
            (h_du_dp && P > 0) ? d_du_dp.data : nullptr, h_u ? d_u.data : nullptr, stream);
        TMB_CUDA(cudaStreamSynchronize(stream));
        if (!redo_after_overflow(*this, attempt)) {
            break;
        }
    }
    if (h_du_dx) {
        d_du_dx.copy_to(h_du_dx);
    }
    if (h_du_dp) {
        d_du_dp.copy_to(h_du_dp);
    }
    if (h_u) {
        d_u.copy_to(h_u);
    }
}

void Potential::execute_batch_host(
    int coord_batch, int N, int param_batch, int P, const double *h_x, const double *h_p, const double *h_box,
    u64 *h_du_dx, u64 *h_du_dp, i128 *h_u) {
    ScratchBuffer<double> d_p(static_cast<size_t>(param_batch) * P), d_box(static_cast<size_t>(coord_batch) * D * D),
        d_x(static_cast<size_t>(coord_batch) * N * D);
    if (P > 0) {
        d_p.copy_from(h_p);
    }
    d_box.copy_from(h_box);
    d_x.copy_from(h_x);
    const size_t total = static_cast<size_t>(coord_batch) * param_batch;
    cudaStream_t stream = main_stream();
    ScratchBuffer<u64> d_du_dx, d_du_dp;
    ScratchBuffer<i128> d_u;
    if (h_du_dx) {
        d_du_dx.realloc(total * N * D);
    }
    if (h_du_dp) {
        d_du_dp.realloc(total * P);
    }
    if (h_u) {
        d_u.realloc(total);
    }
    for (int attempt = 0;; attempt++) {
        d_du_dx.zero(stream);
        d_du_dp.zero(stream);
        d_u.zero(stream);
        // serial loop over (coords_i x params_j): the potentials are stateful, coordinates vary slowest so a cached
        // neighbour list is reused across parameter sets (reference potential.cu:10-38)
        for (int i = 0; i < coord_batch; i++) {
            for (int j = 0; j < param_batch; j++) {
                const size_t k = static_cast<size_t>(i) * param_batch + j;
                this->execute_device(
                    N, P, d_x.data + static_cast<size_t>(i) * N * D, P > 0 ? d_p.data + static_cast<size_t>(j) * P : nullptr,
                    d_box.data + static_cast<size_t>(i) * D * D, h_du_dx ? d_du_dx.data + k * N * D : nullptr,
                    (h_du_dp && P > 0) ? d_du_dp.data + k * P : nullptr, h_u ? d_u.data + k : nullptr, stream);
            }
        }
        TMB_CUDA(cudaStreamSynchronize(stream));
        if (!redo_after_overflow(*this, attempt)) {
            break;
        }
    }
    if (h_du_dx) {
        d_du_dx.copy_to(h_du_dx);
    }
    if (h_du_dp) {
        d_du_dp.copy_to(h_du_dp);
    }
    if (h_u) {
        d_u.copy_to(h_u);
    }
}

void Potential::execute_batch_sparse_host(
    int coords_size, int N, int params_size, int P, int batch_size, const unsigned int *coords_idxs,
    const unsigned int *params_idxs, const double *h_x, const double *h_p, const double *h_box, u64 *h_du_dx,
    u64 *h_du_dp, i128 *h_u) {
    ScratchBuffer<double> d_p(static_cast<size_t>(params_size) * P), d_box(static_cast<size_t>(coords_size) * D * D),
        d_x(static_cast<size_t>(coords_size) * N * D);
    // sparse batch: only the parameter sets and coordinate sets some pair refers to cross PCIe (an HREX rank evaluates its
    // replica under 2-3 of K states: 3.4 MB per state at 30k atoms)
    std::vector<char> p_used(params_size, 0), c_used(coords_size, 0);
    for (int k = 0; k < batch_size; k++) {
        if (coords_idxs[k] >= static_cast<unsigned int>(coords_size) || params_idxs[k] >= static_cast<unsigned int>(params_size)) {
            throw std::runtime_error("batch index out of range");
        }
        c_used[coords_idxs[k]] = 1;
        p_used[params_idxs[k]] = 1;
    }
    for (int j = 0; j < params_size && P > 0; j++) {
        if (p_used[j]) {
            TMB_CUDA(cudaMemcpy(d_p.data + static_cast<size_t>(j) * P, h_p + static_cast<size_t>(j) * P, sizeof(double) * P, cudaMemcpyHostToDevice));
        }
    }
    for (int i = 0; i < coords_size; i++) {
        if (c_used[i]) {
            TMB_CUDA(cudaMemcpy(d_x.data + static_cast<size_t>(i) * N * D, h_x + static_cast<size_t>(i) * N * D, sizeof(double) * N * D, cudaMemcpyHostToDevice));
        }
    }
    d_box.copy_from(h_box);
    cudaStream_t stream = main_stream();
    ScratchBuffer<u64> d_du_dx, d_du_dp;
    ScratchBuffer<i128> d_u;
    if (h_du_dx) {
        d_du_dx.realloc(static_cast<size_t>(batch_size) * N * D);
    }
    if (h_du_dp) {
        d_du_dp.realloc(static_cast<size_t>(batch_size) * P);
    }
    if (h_u) {
        d_u.realloc(batch_size);
    }
    for (int attempt = 0;; attempt++) {
        d_du_dx.zero(stream);
        d_du_dp.zero(stream);
        d_u.zero(stream);
        for (int k = 0; k < batch_size; k++) {
            const size_t ic = coords_idxs[k];
            const size_t ip = params_idxs[k];
            this->execute_device(
                N, P, d_x.data + ic * N * D, P > 0 ? d_p.data + ip * P : nullptr, d_box.data + ic * D * D,
                h_du_dx ? d_du_dx.data + static_cast<size_t>(k) * N * D : nullptr,
                (h_du_dp && P > 0) ? d_du_dp.data + static_cast<size_t>(k) * P : nullptr, h_u ? d_u.data + k : nullptr, stream);
        }
        TMB_CUDA(cudaStreamSynchronize(stream));
        if (!redo_after_overflow(*this, attempt)) {
            break;
        }
    }
    if (h_du_dx) {
        d_du_dx.copy_to(h_du_dx);
    }
    if (h_du_dp) {
        d_du_dp.copy_to(h_du_dp);
    }
    if (h_u) {
        d_u.copy_to(h_u);
    }
}

// ---------------------------------------------------------------------------------------------------------------
BoundPotential::BoundPotential(std::shared_ptr<Potential> potential, const std::vector<double> &params)
    : size(static_cast<int>(params.size())), d_p(params.size()), potential(std::move(potential)) {
    set_params(params);
}

void BoundPotential::set_params(const std::vector<double> &params) {
    if (params.size() != d_p.length) {
        throw std::runtime_error(
            "parameter size is not equal to device buffer size: " + std::to_string(params.size()) +
            " != " + std::to_string(d_p.length));
    }
    d_p.copy_from(params.data());
    size = static_cast<int>(params.size());
}

void BoundPotential::set_params_device(int new_size, const double *d_params, cudaStream_t stream) {
    if (static_cast<size_t>(new_size) > d_p.length) {
        throw std::runtime_error(
            "parameter size is greater than device buffer size: " + std::to_string(new_size) + " > " +
            std::to_string(d_p.length));
    }
    TMB_CUDA(cudaMemcpyAsync(d_p.data, d_params, new_size * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    if (size != new_size) {
        bump_launch_generation(); // P is a kernel argument of every captured launch
    }
    size = new_size;
}

void BoundPotential::execute_device(
    int N, const double *d_x, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u, cudaStream_t s) {
    potential->execute_device(N, size, d_x, size > 0 ? d_p.data : nullptr, d_box, d_du_dx, d_du_dp, d_u, s);
}

void BoundPotential::execute_host(int N, const double *h_x, const double *h_box, u64 *h_du_dx, i128 *h_u) {
    ScratchBuffer<double> d_x(static_cast<size_t>(N) * 3), d_box(9);
    d_x.copy_from(h_x);
    d_box.copy_from(h_box);
    cudaStream_t stream = main_stream();
    ScratchBuffer<u64> d_du_dx;
    ScratchBuffer<i128> d_u;
    if (h_du_dx) {
        d_du_dx.realloc(static_cast<size_t>(N) * 3);
    }
    if (h_u) {
        d_u.realloc(1);
    }
    for (int attempt = 0;; attempt++) {
        d_du_dx.zero(stream);
        d_u.zero(stream);
        execute_device(N, d_x.data, d_box.data, h_du_dx ? d_du_dx.data : nullptr, nullptr, h_u ? d_u.data : nullptr, stream);
        TMB_CUDA(cudaStreamSynchronize(stream));
        if (!redo_after_overflow(*potential, attempt)) {
            break;
        }
    }
    if (h_du_dx) {
        d_du_dx.copy_to(h_du_dx);
    }
    if (h_u) {
        d_u.copy_to(h_u);
    }
}

void BoundPotential::execute_batch_host(int coord_batch, int N, const double *h_x, const double *h_box, u64 *h_du_dx, i128 *h_u) {
    ScratchBuffer<double> d_x(static_cast<size_t>(coord_batch) * N * 3), d_box(static_cast<size_t>(coord_batch) * 9);
    d_x.copy_from(h_x);
    d_box.copy_from(h_box);
    cudaStream_t stream = main_stream();
    ScratchBuffer<u64> d_du_dx;
    ScratchBuffer<i128> d_u;
    if (h_du_dx) {
        d_du_dx.realloc(static_cast<size_t>(coord_batch) * N * 3);
    }
    if (h_u) {
        d_u.realloc(coord_batch);
    }
    for (int attempt = 0;; attempt++) {
        d_du_dx.zero(stream);
        d_u.zero(stream);
        for (int i = 0; i < coord_batch; i++) {
            execute_device(
                N, d_x.data + static_cast<size_t>(i) * N * 3, d_box.data + static_cast<size_t>(i) * 9,
                h_du_dx ? d_du_dx.data + static_cast<size_t>(i) * N * 3 : nullptr, nullptr, h_u ? d_u.data + i : nullptr, stream);
        }
        TMB_CUDA(cudaStreamSynchronize(stream));
        if (!redo_after_overflow(*potential, attempt)) {
            break;
        }
    }
    if (h_du_dx) {
        d_du_dx.copy_to(h_du_dx);
    }
    if (h_u) {
        d_u.copy_to(h_u);
    }
}

// ---------------------------------------------------------------------------------------------------------------
SummedPotential::SummedPotential(
    std::vector<std::shared_ptr<Potential>> potentials, std::vector<int> params_sizes, bool parallel)
    : potentials_(std::move(potentials)), params_sizes_(std::move(params_sizes)),
      P_(std::accumulate(params_sizes_.begin(), params_sizes_.end(), 0)), parallel_(parallel),
      d_u_children_(potentials_.size()) {
    if (potentials_.size() != params_sizes_.size()) {
        throw std::runtime_error("number of potentials != number of parameter sizes");
    }
}

void SummedPotential::execute_device(
    int N, int P, const double *d_x, const double *d_p, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
    cudaStream_t stream) {
    if (P != P_) {
        throw std::runtime_error(
            "SummedPotential::execute_device(): expected " + std::to_string(P_) + " parameters, got " + std::to_string(P));
    }
    const int n = static_cast<int>(potentials_.size());
    if (d_u) {
        d_u_children_.zero(stream);
    }
    int offset = 0;
    for (int i = 0; i < n; i++) {
        cudaStream_t s = parallel_ ? fan_.fork(i, stream) : stream;
        potentials_[i]->execute_device(
            N, params_sizes_[i], d_x, d_p + offset, d_box, d_du_dx, d_du_dp == nullptr ? nullptr : d_du_dp + offset,
            d_u == nullptr ? nullptr : d_u_children_.data + i, s);
        offset += params_sizes_[i];
    }
    if (parallel_) {
        for (int i = 0; i < n; i++) {
            fan_.join(i, stream);
        }
    }
    if (d_u) {
        launch_sum_i128(d_u_children_.data, n, d_u, stream);
    }
}

void SummedPotential::du_dp_fixed_to_float(int N, int, const u64 *du_dp, double *out) const {
    int offset = 0;
    for (size_t i = 0; i < potentials_.size(); i++) {
        potentials_[i]->du_dp_fixed_to_float(N, params_sizes_[i], du_dp + offset, out + offset);
        offset += params_sizes_[i];
    }
}

int SummedPotential::capturable_steps() const {
    int n = 1 << 30;
    for (auto &p : potentials_) {
        n = std::min(n, p->capturable_steps());
    }
    return n;
}
void SummedPotential::advance(int n) {
    for (auto &p : potentials_) {
        p->advance(n);
    }
}
bool SummedPotential::recover_overflow() {
    bool any = false;
    for (auto &p : potentials_) {
        any = p->recover_overflow() || any; // every child gets to grow
    }
    return any;
}

FanoutSummedPotential::FanoutSummedPotential(std::vector<std::shared_ptr<Potential>> potentials, bool parallel)
    : potentials_(std::move(potentials)), parallel_(parallel), d_u_children_(potentials_.size()) {}

void FanoutSummedPotential::execute_device(
    int N, int P, const double *d_x, const double *d_p, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
    cudaStream_t stream) {
    const int n = static_cast<int>(potentials_.size());
    if (d_u) {
        d_u_children_.zero(stream);
    }
    for (int i = 0; i < n; i++) {
        cudaStream_t s = parallel_ ? fan_.fork(i, stream) : stream;
        potentials_[i]->execute_device(
            N, P, d_x, d_p, d_box, d_du_dx, d_du_dp, d_u == nullptr ? nullptr : d_u_children_.data + i, s);
    }
    if (parallel_) {
        for (int i = 0; i < n; i++) {
            fan_.join(i, stream);
        }
    }
    if (d_u) {
        launch_sum_i128(d_u_children_.data, n, d_u, stream);
    }
}

void FanoutSummedPotential::du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const {
    if (!potentials_.empty()) {
        potentials_[0]->du_dp_fixed_to_float(N, P, du_dp, out);
    }
}
int FanoutSummedPotential::capturable_steps() const {
    int n = 1 << 30;
    for (auto &p : potentials_) {
        n = std::min(n, p->capturable_steps());
    }
    return n;
}
void FanoutSummedPotential::advance(int n) {
    for (auto &p : potentials_) {
        p->advance(n);
    }
}
bool FanoutSummedPotential::recover_overflow() {
    bool any = false;
    for (auto &p : potentials_) {
        any = p->recover_overflow() || any;
    }
    return any;
}

void collect_nonbonded_cutoffs(const std::shared_ptr<Potential> &pot, std::vector<double> &out) {
    if (auto s = std::dynamic_pointer_cast<SummedPotential>(pot)) {
        for (auto &c : s->get_potentials()) {
            collect_nonbonded_cutoffs(c, out);
        }
    } else if (auto f = std::dynamic_pointer_cast<FanoutSummedPotential>(pot)) {
        for (auto &c : f->get_potentials()) {
            collect_nonbonded_cutoffs(c, out);
        }
    } else if (auto a = std::dynamic_pointer_cast<NonbondedAllPairs<float>>(pot)) {
        out.push_back(a->get_cutoff() + a->get_nblist_padding());
    } else if (auto b = std::dynamic_pointer_cast<NonbondedAllPairs<double>>(pot)) {
        out.push_back(b->get_cutoff() + b->get_nblist_padding());
    }
}

} // namespace tmb
