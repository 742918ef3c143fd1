// Monte Carlo barostat and the Mover base (see barostat.hpp for what follows the reference and what does not).
//
// This translation unit is compiled WITHOUT --fmad=false (build.py), like the reference: cbrtf / logf / exp are inlined
// from libdevice and must be contracted the same way.  Every other floating-point operation is an explicit
// round-to-nearest intrinsic, so the arithmetic below is exactly the sequence ptxas emits for the reference's
// k_setup_barostat_move / k_rescale_positions / k_decide_move <float> (read off oracle/_ref/obj/barostat.o); e.g. the
// final shift of a molecule is mul-then-sub for x and y but a fused multiply-add for z.
#include "fixed_point.cuh"
#include "kernels.hpp"
#include "potential.hpp"

#include <curand.h>

#include <algorithm>
#include <iostream>
#include <set>

namespace tmb {

// reference constants.hpp:5-6
static const double BOLTZ = 0.008314462618;
static const double AVOGADRO = 6.0221367e23;
static const int RANDOM_BATCH_SIZE = 1000; // moves per cuRAND batch (two uniforms each), reference barostat.cu:14

#define TMB_CURAND(expr)                                                                                               \
    do {                                                                                                               \
        curandStatus_t _st = (expr);                                                                                   \
        if (_st != CURAND_STATUS_SUCCESS) {                                                                            \
            throw std::runtime_error(std::string("cuRAND error ") + std::to_string(static_cast<int>(_st)) + " at " +   \
                                     __FILE__ + ":" + std::to_string(__LINE__));                                       \
        }                                                                                                              \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
void Mover::set_step(int step) {
    if (step < 0) {
        throw std::runtime_error("step must be at least 0");
    }
    step_ = step;
}

void Mover::set_interval(int interval) {
    if (interval <= 0) {
        throw std::runtime_error("interval must be greater than 0");
    }
    interval_ = interval;
    step_ = 0; // the mover acts `interval` calls from now
}

std::array<std::vector<double>, 2> Mover::move_host(int N, const double *h_x, const double *h_box) {
    DeviceBuffer<double> d_x(static_cast<size_t>(N) * 3), d_box(9);
    d_x.copy_from(h_x);
    d_box.copy_from(h_box);
    cudaStream_t stream = main_stream();
    this->move(N, d_x.data, d_box.data, stream);
    TMB_CUDA(cudaStreamSynchronize(stream));
    std::vector<double> x(d_x.length), box(d_box.length);
    d_x.copy_to(x.data());
    d_box.copy_to(box.data());
    return {x, box};
}

void verify_group_idxs(int N, const std::vector<std::vector<int>> &group_idxs) { // reference mol_utils.cpp:8-27
    size_t n_grouped = 0;
    std::set<int> seen;
    for (const auto &g : group_idxs) {
        n_grouped += g.size();
        for (int idx : g) {
            if (idx < 0 || idx >= N) {
                throw std::runtime_error("Grouped indices must be between 0 and N");
            }
            seen.insert(idx);
        }
    }
    if (seen.size() != n_grouped) {
        throw std::runtime_error("All grouped indices must be unique");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// arithmetic helpers: explicit rounding, never contracted
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmad_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fmad_(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float cbrt_(float a) { return cbrtf(a); }
__device__ __forceinline__ double cbrt_(double a) { return cbrt(a); }
__device__ __forceinline__ float log_(float a) { return logf(a); }
__device__ __forceinline__ double log_(double a) { return log(a); }
__device__ __forceinline__ float floor_(float a) { return floorf(a); }
__device__ __forceinline__ double floor_(double a) { return floor(a); }

template <typename Real> struct ProposeArgs {
    int N;
    int num_mols;
    int num_ungrouped;
    bool adaptive;
    const Real *rand;        // [2]; rand[0] picks the volume change
    const double *box;       // [9]
    const int *mol_offsets;  // [num_mols + 1] into atom_idxs
    const int *atom_idxs;    // [num_grouped]
    const int *ungrouped;    // [num_ungrouped] atoms that belong to no group: copied unchanged
    const double *x;         // [N,3]
    double *x_proposed;      // [N,3]
    double *box_proposed;    // [9]
    double *volume_scale;    // [1] read; written only when the adaptive default (1 % of the volume) is installed
    Real *volume;            // [2] {volume, volume_delta} for the decision kernel
};

constexpr int BARO_THREADS = 128;

// One warp per molecule.  Reference k_setup_barostat_move (k_barostat.cuh:92-115), k_find_group_centroids (:71-88) and
// k_rescale_positions (:11-68) in one pass; every warp recomputes the few scalars of the setup.
template <typename Real> __global__ void __launch_bounds__(BARO_THREADS) k_barostat_propose(const ProposeArgs<Real> a) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;

    const double b0 = a.box[0], b4 = a.box[4], b8 = a.box[8];
    const Real volume = static_cast<Real>(mul_(mul_(b0, b4), b8));
    double scale_factor = *a.volume_scale;
    const bool install_default = a.adaptive && scale_factor == 0.0;
    if (install_default) {
        scale_factor = mul_(static_cast<double>(volume), 0.01);
    }
    const Real delta_volume =
        static_cast<Real>(mul_(add_(scale_factor, scale_factor), add_(static_cast<double>(a.rand[0]), -0.5)));
    const Real new_volume = add_(volume, delta_volume);
    const Real scale = cbrt_(new_volume / volume);
    const double scale_d = static_cast<double>(scale);

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (install_default) {
            *a.volume_scale = scale_factor; // racing readers see 0 or this value and compute the same thing
        }
        a.volume[0] = volume;
        a.volume[1] = delta_volume;
        for (int i = 0; i < 9; i++) {
            a.box_proposed[i] = a.box[i];
        }
        a.box_proposed[0] = mul_(b0, scale_d);
        a.box_proposed[4] = mul_(b4, scale_d);
        a.box_proposed[8] = mul_(b8, scale_d);
    }

    const Real center[3] = {
        static_cast<Real>(mul_(b0, 0.5)), static_cast<Real>(mul_(b4, 0.5)), static_cast<Real>(mul_(b8, 0.5))};
    const Real scaled_box[3] = {
        static_cast<Real>(mul_(b0, scale_d)), static_cast<Real>(mul_(b4, scale_d)), static_cast<Real>(mul_(b8, scale_d))};
    const Real inv_fixed = static_cast<Real>(1.0 / 68719476736.0); // 2^-36, exact

    for (int m = gwarp; m < a.num_mols; m += n_warps) {
        const int begin = a.mol_offsets[m], end = a.mol_offsets[m + 1];
        // centroid: coordinates are rounded to Real, then to fixed point, and summed as integers (order-free)
        u64 sum[3] = {0, 0, 0};
        for (int k = begin + lane; k < end; k += WARP) {
            const size_t atom = a.atom_idxs[k];
            for (int c = 0; c < 3; c++) {
                sum[c] += to_fixed<FIXED_EXPONENT>(static_cast<Real>(a.x[atom * 3 + c]));
            }
        }
        for (int d = 16; d > 0; d >>= 1) {
            for (int c = 0; c < 3; c++) {
                sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], d);
            }
        }
        const Real n_atoms = static_cast<Real>(end - begin);
        Real shift[3];
        for (int c = 0; c < 3; c++) {
            const Real centroid = mul_(static_cast<Real>(static_cast<i64>(sum[c])), inv_fixed) / n_atoms;
            const Real displacement = add_(fmad_(scale, add_(centroid, -center[c]), center[c]), -centroid);
            const Real moved = add_(displacement, centroid);
            // bring the moved centroid back into the scaled home box
            const Real cells = floor_(moved / scaled_box[c]);
            shift[c] = c < 2 ? add_(displacement, -mul_(scaled_box[c], cells)) : fmad_(scaled_box[c], -cells, displacement);
        }
        for (int k = begin + lane; k < end; k += WARP) {
            const size_t atom = a.atom_idxs[k];
            for (int c = 0; c < 3; c++) {
                a.x_proposed[atom * 3 + c] = add_(static_cast<double>(shift[c]), a.x[atom * 3 + c]);
            }
        }
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < a.num_ungrouped; k += gridDim.x * blockDim.x) {
        const size_t atom = a.ungrouped[k];
        for (int c = 0; c < 3; c++) {
            a.x_proposed[atom * 3 + c] = a.x[atom * 3 + c];
        }
    }
}

template <typename Real> struct DecideArgs {
    int N;
    bool adaptive;
    int num_mols;
    int num_bps;
    double kt;
    double pressure;
    const Real *rand;   // [2]; rand[1] is the Metropolis draw
    const Real *volume; // {volume, volume_delta}
    double *volume_scale;
    const i128 *u_init;  // [num_bps]
    const i128 *u_final; // [num_bps]
    double *box;
    const double *box_proposed;
    double *x;
    const double *x_proposed;
    int *counters; // {attempted, accepted}
};

__device__ __forceinline__ bool energy_overflowed(i128 v) {
    return v >= static_cast<i128>(LLONG_MAX) || v <= static_cast<i128>(LLONG_MIN);
}

// Reference k_decide_move (k_barostat.cuh:120-189); the per-potential energies are summed here instead of by CUB.
template <typename Real> __global__ void __launch_bounds__(BARO_THREADS) k_barostat_decide(const DecideArgs<Real> a) {
    const Real volume = a.volume[0];
    const Real volume_delta = a.volume[1];
    const Real new_volume = add_(volume, volume_delta);
    i128 u0 = 0, u1 = 0;
    for (int i = 0; i < a.num_bps; i++) {
        u0 += a.u_init[i];
        u1 += a.u_final[i];
    }
    Real energy_delta = static_cast<Real>(INFINITY);
    if (!energy_overflowed(u1) && !energy_overflowed(u0)) {
        energy_delta = mul_(static_cast<Real>(static_cast<i64>(u1 - u0)), static_cast<Real>(1.0 / 68719476736.0));
    }
    // w = dU + P dV - N kT ln(V'/V), evaluated in double from Real inputs and rounded to Real
    const double nkt = mul_(static_cast<double>(a.num_mols), a.kt);
    const double pv = fmad_(static_cast<double>(volume_delta), a.pressure, static_cast<double>(energy_delta));
    const Real w = static_cast<Real>(fmad_(-nkt, static_cast<double>(log_(new_volume / volume)), pv));
    const bool rejected = w > 0 && static_cast<double>(a.rand[1]) > exp(static_cast<double>(-w) / a.kt);

    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid == 0) {
        int attempted = a.counters[0] + 1;
        int accepted = a.counters[1] + (rejected ? 0 : 1);
        if (a.adaptive && attempted >= 10) {
            if (accepted < 0.25 * attempted) {
                *a.volume_scale = *a.volume_scale / 1.1;
                attempted = accepted = 0;
            } else if (accepted > 0.75 * attempted) {
                *a.volume_scale = min(*a.volume_scale * 1.1, static_cast<double>(volume) * 0.3);
                attempted = accepted = 0;
            }
        }
        a.counters[0] = attempted;
        a.counters[1] = accepted;
    }
    if (rejected) {
        return;
    }
    if (tid < 9) {
        a.box[tid] = a.box_proposed[tid];
    }
    for (int i = tid; i < a.N * 3; i += gridDim.x * blockDim.x) {
        a.x[i] = a.x_proposed[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
template <typename Real>
MonteCarloBarostat<Real>::MonteCarloBarostat(
    int N, double pressure, double temperature, std::vector<std::vector<int>> group_idxs, int interval,
    std::vector<std::shared_ptr<BoundPotential>> bps, int seed, bool adaptive_scaling_enabled,
    double initial_volume_scale_factor)
    : Mover(interval), N_(N), adaptive_(adaptive_scaling_enabled), bps_(std::move(bps)),
      pressure_(static_cast<Real>(pressure)), temperature_(static_cast<Real>(temperature)), seed_(seed) {
    this->set_interval(interval_); // validates
    if (temperature < 100.0) {
        std::cout << "warning temperature less than 100K" << std::endl;
    }
    if (pressure > 10.0) {
        std::cout << "warning pressure more than 10bar" << std::endl;
    }
    verify_group_idxs(N, group_idxs);

    // flatten: atoms of molecule m are atom_idxs[offsets[m] .. offsets[m+1]) (reference mol_utils.cpp:53-85)
    std::vector<int> atom_idxs, offsets;
    std::vector<char> grouped(N, 0);
    for (auto g : group_idxs) {
        std::sort(g.begin(), g.end());
        offsets.push_back(static_cast<int>(atom_idxs.size()));
        for (int idx : g) {
            atom_idxs.push_back(idx);
            grouped[idx] = 1;
        }
    }
    offsets.push_back(static_cast<int>(atom_idxs.size()));
    std::vector<int> ungrouped;
    for (int i = 0; i < N; i++) {
        if (!grouped[i]) {
            ungrouped.push_back(i);
        }
    }
    num_mols_ = static_cast<int>(group_idxs.size());
    num_grouped_ = static_cast<int>(atom_idxs.size());
    num_ungrouped_ = static_cast<int>(ungrouped.size());

    TMB_CURAND(curandCreateGenerator(&rng_, CURAND_RNG_PSEUDO_DEFAULT));
    TMB_CURAND(curandSetPseudoRandomGeneratorSeed(rng_, seed_));
    d_rand_.realloc(RANDOM_BATCH_SIZE * 2);
    d_rand_.zero();
    d_counters_.realloc(2);
    d_u_init_.realloc(std::max<size_t>(1, bps_.size()));
    d_u_final_.realloc(std::max<size_t>(1, bps_.size()));
    // a potential with nothing to evaluate (no terms, empty pair list, empty interaction group) returns without writing its
    // energy slot, and k_barostat_decide sums every slot: they must read as zero, not as whatever cudaMalloc handed out
    d_u_init_.zero();
    d_u_final_.zero();
    d_volume_.realloc(2);
    d_volume_scale_.realloc(1);
    d_volume_scale_.copy_from(&initial_volume_scale_factor);
    d_x_proposed_.realloc(static_cast<size_t>(N) * 3);
    d_box_proposed_.realloc(9);
    d_atom_idxs_.realloc(std::max<size_t>(1, atom_idxs.size()));
    d_mol_offsets_.realloc(offsets.size());
    d_ungrouped_.realloc(std::max<size_t>(1, ungrouped.size()));
    if (!atom_idxs.empty()) {
        TMB_CUDA(cudaMemcpy(d_atom_idxs_.data, atom_idxs.data(), atom_idxs.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    d_mol_offsets_.copy_from(offsets.data());
    if (!ungrouped.empty()) {
        TMB_CUDA(cudaMemcpy(d_ungrouped_.data, ungrouped.data(), ungrouped.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    reset_counters();
    TMB_CUDA(cudaDeviceSynchronize());
}

template <typename Real> MonteCarloBarostat<Real>::~MonteCarloBarostat() {
    if (rng_) {
        curandDestroyGenerator(rng_);
    }
}

template <typename Real> void MonteCarloBarostat<Real>::reset_counters() {
    TMB_CUDA(cudaMemset(d_counters_.data, 0, d_counters_.bytes()));
}

template <typename Real> double MonteCarloBarostat<Real>::get_volume_scale_factor() {
    TMB_CUDA(cudaDeviceSynchronize());
    double h = 0;
    d_volume_scale_.copy_to(&h);
    return h;
}

template <typename Real> void MonteCarloBarostat<Real>::set_volume_scale_factor(double volume_scale_factor) {
    TMB_CUDA(cudaDeviceSynchronize());
    d_volume_scale_.copy_from(&volume_scale_factor);
    reset_counters();
}

template <typename Real> void MonteCarloBarostat<Real>::set_pressure(double pressure) {
    TMB_CUDA(cudaDeviceSynchronize());
    pressure_ = static_cast<Real>(pressure);
    reset_counters(); // the adaptation history belongs to the old pressure (reference barostat.cu:252-257)
}

template <typename Real> std::array<float, 2> MonteCarloBarostat<Real>::last_uniforms() {
    TMB_CUDA(cudaDeviceSynchronize());
    Real h[2];
    TMB_CUDA(cudaMemcpy(h, d_rand_.data + last_offset_, sizeof(h), cudaMemcpyDeviceToHost));
    return {static_cast<float>(h[0]), static_cast<float>(h[1])};
}

template <typename Real> std::array<int, 2> MonteCarloBarostat<Real>::counters() {
    TMB_CUDA(cudaDeviceSynchronize());
    int h[2];
    d_counters_.copy_to(h);
    return {h[0], h[1]};
}

template <typename Real>
void MonteCarloBarostat<Real>::energies(const double *d_x, const double *d_box, i128 *d_u, cudaStream_t stream) {
    // every bound potential on its own stream, energy only (reference StreamedPotentialRunner, barostat.cu:218-226)
    const int n = static_cast<int>(bps_.size());
    for (int i = 0; i < n; i++) {
        cudaStream_t s = fan_.fork(i, stream);
        bps_[i]->execute_device(N_, d_x, d_box, nullptr, nullptr, d_u + i, s);
    }
    for (int i = 0; i < n; i++) {
        fan_.join(i, stream);
    }
}

static curandStatus_t generate_uniform(curandGenerator_t g, float *out, size_t n) { return curandGenerateUniform(g, out, n); }
static curandStatus_t generate_uniform(curandGenerator_t g, double *out, size_t n) {
    return curandGenerateUniformDouble(g, out, n);
}

template <typename Real> void MonteCarloBarostat<Real>::move(int N, double *d_x, double *d_box, cudaStream_t stream) {
    if (N != N_) {
        throw std::runtime_error("N != N_");
    }
    step_++;
    if (step_ % interval_ != 0) {
        return;
    }
    // two uniforms per attempted move, drawn in batches (reference barostat.cu:166-176)
    const int offset = (((step_ / interval_) * 2) - 2) % (RANDOM_BATCH_SIZE * 2);
    if (offset == 0) {
        TMB_CURAND(curandSetStream(rng_, stream));
        TMB_CURAND(generate_uniform(rng_, d_rand_.data, RANDOM_BATCH_SIZE * 2));
    }
    last_offset_ = offset;

    ProposeArgs<Real> pa;
    pa.N = N_;
    pa.num_mols = num_mols_;
    pa.num_ungrouped = num_ungrouped_;
    pa.adaptive = adaptive_;
    pa.rand = d_rand_.data + offset;
    pa.box = d_box;
    pa.mol_offsets = d_mol_offsets_.data;
    pa.atom_idxs = d_atom_idxs_.data;
    pa.ungrouped = d_ungrouped_.data;
    pa.x = d_x;
    pa.x_proposed = d_x_proposed_.data;
    pa.box_proposed = d_box_proposed_.data;
    pa.volume_scale = d_volume_scale_.data;
    pa.volume = d_volume_.data;
    const int warps = std::max(1, std::max(num_mols_, ceil_div(num_ungrouped_, WARP)));
    const int grid = std::min(ceil_div(warps * WARP, BARO_THREADS), 16 * sm_count());
    TMB_LAUNCH(k_barostat_propose<Real>, grid, BARO_THREADS, 0, stream, pa);

    energies(d_x, d_box, d_u_init_.data, stream);
    energies(d_x_proposed_.data, d_box_proposed_.data, d_u_final_.data, stream);

    DecideArgs<Real> da;
    da.N = N_;
    da.adaptive = adaptive_;
    da.num_mols = num_mols_;
    da.num_bps = static_cast<int>(bps_.size());
    da.kt = BOLTZ * temperature_;
    da.pressure = pressure_ * AVOGADRO * 1e-25; // bar -> kJ/mol/nm^3
    da.rand = d_rand_.data + offset;
    da.volume = d_volume_.data;
    da.volume_scale = d_volume_scale_.data;
    da.u_init = d_u_init_.data;
    da.u_final = d_u_final_.data;
    da.box = d_box;
    da.box_proposed = d_box_proposed_.data;
    da.x = d_x;
    da.x_proposed = d_x_proposed_.data;
    da.counters = d_counters_.data;
    TMB_LAUNCH(k_barostat_decide<Real>, std::min(ceil_div(N_ * 3, BARO_THREADS), 8 * sm_count()), BARO_THREADS, 0, stream, da);
}

template class MonteCarloBarostat<float>;
template class MonteCarloBarostat<double>;

} // namespace tmb
