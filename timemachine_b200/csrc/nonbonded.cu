// Host orchestration of the nonbonded potentials: Hilbert sort cadence, gather, device-side rebuild decision,
// tile kernel.  Reference: nonbonded_all_pairs.cu:21-313, nonbonded_interaction_group.cu:20-386,
// nonbonded_pair_list.cu:12-121, nonbonded_common.cpp:45-63.
//
// One evaluation enqueues (no host synchronisation anywhere):
//   [every steps_per_sort-th call: hilbert keys + radix sort -> perm, force rebuild]
//   k_nb_prepare      gather/cast into packed sorted working set, decide rebuild (device flag)
//   k_block_bounds    } return immediately unless the flag is set (bounds also resets the tile counter and
//   k_build_tiles     } snapshots the build coordinates)
//   k_nb_tiles(_cq)   the hot loop: adds straight into the caller's atom-order du_dx / du_dp (slot -> atom through
//                     perm when a tile is flushed; the reference's separate scatter pass does not exist) and clears
//                     the rebuild flag
#include "fixed_point.cuh"
#include "potential.hpp"

#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <type_traits>

namespace tmb {

void verify_atom_idxs(int N, const std::vector<int> &atom_idxs, bool allow_empty) {
    if (atom_idxs.empty()) {
        if (allow_empty) {
            return;
        }
        throw std::runtime_error("indices can't be empty");
    }
    std::set<int> unique_idxs(atom_idxs.begin(), atom_idxs.end());
    if (unique_idxs.size() != atom_idxs.size()) {
        throw std::runtime_error("atom indices must be unique");
    }
    if (*std::max_element(atom_idxs.begin(), atom_idxs.end()) >= N) {
        throw std::runtime_error("index values must be less than N(" + std::to_string(N) + ")");
    }
    if (*std::min_element(atom_idxs.begin(), atom_idxs.end()) < 0) {
        throw std::runtime_error("index values must be greater or equal to zero");
    }
}

void nonbonded_du_dp_fixed_to_float(int N, int, const u64 *du_dp, double *out) {
    for (int i = 0; i < N; i++) {
        const int b = i * P_PER_ATOM;
        out[b + P_CHARGE] = static_cast<double>(static_cast<i64>(du_dp[b + P_CHARGE])) / FIXED_EXPONENT_DU_DCHARGE;
        out[b + P_SIG] = static_cast<double>(static_cast<i64>(du_dp[b + P_SIG])) / FIXED_EXPONENT_DU_DSIG;
        out[b + P_EPS] = static_cast<double>(static_cast<i64>(du_dp[b + P_EPS])) / FIXED_EXPONENT_DU_DEPS;
        out[b + P_W] = static_cast<double>(static_cast<i64>(du_dp[b + P_W])) / FIXED_EXPONENT_DU_DW;
    }
}

// ---------------------------------------------------------------------------------------------------------------
template <typename Real>
NonbondedTiled<Real>::NonbondedTiled(
    int N, double beta, double cutoff, bool disable_hilbert, double nblist_padding, int steps_per_sort)
    : N_(N), beta_(beta), cutoff_(cutoff), nblist_padding_(nblist_padding), disable_hilbert_(disable_hilbert),
      steps_per_sort_(steps_per_sort), d_perm_(N), d_xw_(round_up(N, TILE)), d_qse_(round_up(N, TILE)),
      d_xw_build_(round_up(N, TILE)), d_box_build_(9), d_flags_(3),
      d_partials_(nb_tiles_max_grid<Real>()), d_ticket_(1), nblist_(N) {
    d_xw_.zero();
    d_qse_.zero();
    d_xw_build_.zero(); // nonsensical positions: the first evaluation always rebuilds anyway
    d_box_build_.zero();
    d_flags_.zero();
    d_ticket_.zero();
    if (!disable_hilbert_) {
        hilbert_.reset(new HilbertSort(N));
    }
    TMB_CUDA(cudaDeviceSynchronize());
}

template <typename Real> void NonbondedTiled<Real>::du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const {
    nonbonded_du_dp_fixed_to_float(N, P, du_dp, out);
}

template <typename Real> unsigned int NonbondedTiled<Real>::num_tiles() { return nblist_.num_tile_ixns(); }

template <typename Real> bool NonbondedTiled<Real>::recover_overflow() {
    if (!nblist_.recover_overflow()) {
        return false;
    }
    force_rebuild_ = true;
    bump_launch_generation(); // the list moved: captured launches hold the old addresses
    return true;
}
template <typename Real> unsigned int NonbondedTiled<Real>::num_rebuilds() {
    TMB_CUDA(cudaDeviceSynchronize());
    unsigned int n = 0;
    TMB_CUDA(cudaMemcpy(&n, d_flags_.data + 2, sizeof(n), cudaMemcpyDeviceToHost));
    return n;
}

template <typename Real> void NonbondedTiled<Real>::set_kernel_timing(bool on) {
    constexpr size_t CAPACITY = 4096;
    if (on && timing_events_.empty()) {
        timing_events_.resize(CAPACITY);
        for (auto &p : timing_events_) {
            TMB_CUDA(cudaEventCreate(&p.first));
            TMB_CUDA(cudaEventCreate(&p.second));
        }
    }
    timing_ = on;
    timing_used_ = 0;
    bump_launch_generation();
}

template <typename Real> std::vector<float> NonbondedTiled<Real>::drain_kernel_times() {
    std::vector<float> out;
    for (size_t i = 0; i < timing_used_; i++) {
        TMB_CUDA(cudaEventSynchronize(timing_events_[i].second));
        float ms = 0.f;
        TMB_CUDA(cudaEventElapsedTime(&ms, timing_events_[i].first, timing_events_[i].second));
        out.push_back(ms);
    }
    timing_used_ = 0;
    return out;
}

template <typename Real>
void NonbondedTiled<Real>::run(
    int N, const double *d_x, const double *d_p, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
    cudaStream_t stream) {
    if (d_du_dx == nullptr && d_du_dp == nullptr && d_u == nullptr) {
        // nothing requested: enqueue nothing.  (A prepare/build without the tile kernel that clears the rebuild flag
        // would leave the flag raised and the next evaluation would append to a stale tile list.)
        return;
    }
    int force = force_rebuild_ ? 1 : 0;
    if (needs_sort()) {
        // a new permutation invalidates the tile list (reference nonbonded_all_pairs.cu:153-164)
        this->sort(d_x, d_box, stream);
        force = 1;
    }
    force_rebuild_ = false;
    // the integrator of the previous step already ran this evaluation's prepare pass (only honoured for a plain evaluation)
    const bool prepared = skip_prepare_once_ && force == 0;
    skip_prepare_once_ = false;

    NbPrepareArgs<Real> pa;
    pa.K = K_;
    pa.perm = d_perm_.data;
    pa.x = d_x;
    pa.p = d_p;
    pa.box = d_box;
    pa.xw_build = d_xw_build_.data;
    pa.box_build = d_box_build_.data;
    pa.padding = nblist_padding_;
    pa.force_rebuild = force;
    pa.flag = d_flags_.data;
    pa.tile_cursor = d_flags_.data + 1;
    pa.xw = d_xw_.data;
    pa.qse = d_qse_.data;
    // all-pairs layout (rows == columns == every gathered slot): block bounds, the counter reset and the build-time
    // snapshot ride along with the gather and the build, the per-step chain is prepare -> build -> tiles
    const bool fuse_bounds = (NR_ == K_);
    if (fuse_bounds) {
        pa.ctr = nblist_.col_ctr();
        pa.ext = nblist_.col_ext();
        pa.reset_count = nblist_.tiles().count;
        pa.reset_overflow = nblist_.tiles().overflow;
    }
    if (!prepared) {
        launch_nb_prepare<Real>(pa, stream);
    }

    const unsigned int *flag = d_flags_.data;
    typename Neighborlist<Real>::Snapshot snap{d_xw_build_.data, d_box_build_.data, fuse_bounds, K_};
    nblist_.build_device(nullptr, d_xw_.data, d_box, cutoff_ + nblist_padding_, flag, stream, &snap);
    (void)N;

    const TileList &tl = nblist_.tiles();
    NbTileArgs<Real> ta;
    ta.K = K_;
    ta.NR = NR_;
    ta.tile_count = tl.count;
    ta.tile_rows = tl.rows;
    ta.tile_cols = tl.cols;
    ta.xw = d_xw_.data;
    ta.qse = d_qse_.data;
    ta.box = d_box;
    ta.beta = beta_;
    ta.cutoff = cutoff_;
    ta.perm = d_perm_.data;
    ta.du_dx = d_du_dx;
    ta.du_dp = d_du_dp;
    ta.u_partials = d_partials_.data;
    ta.ticket = d_ticket_.data;
    ta.d_u = d_u;
    ta.rebuild_flag = d_flags_.data;
    ta.tile_capacity = tl.capacity;
    ta.tile_cursor = d_flags_.data + 1;
    {
        // Persistent-grid policy.  A launch that can fill the machine leaves a few CTA slots free (the tile kernel
        // otherwise owns every register of every SM and a concurrent small tile launch - the ligand-environment
        // interaction group next to the environment all-pairs term - would only start when it ends); a small launch
        // is sized from the upper bound of its tile count and scheduled dynamically, since its CTAs may become
        // resident one after another.
        const long row_blocks = ceil_div(NR_, TILE);
        const long col_blocks = ceil_div(NR_ == K_ ? K_ : K_ - NR_, TILE);
        const long max_tiles = NR_ == K_ ? row_blocks * (row_blocks + 1) / 2 : row_blocks * col_blocks;
        const int full = nb_tiles_max_grid<Real>();
        const int warps_per_cta = 8;
        const int roomy = std::max(1, full - (full > 4 * nb_tiles_reserved_ctas() ? nb_tiles_reserved_ctas() : 0));
        if (max_tiles >= 2L * full * warps_per_cta) {
            ta.grid_ctas = roomy;
            static const int static_tiles = [] {
                const char *env = std::getenv("TMB_NB_STATIC"); // tuning knob: tiles each warp takes before the dynamic part
                return env != nullptr ? std::max(0, std::atoi(env)) : 3; // 2-3 measured best at 30k atoms (r1: 0 -> 124.7, 2 -> 122.3, 3 -> 121.5, 4 -> 126.5, 6 -> 139.6 us)
            }();
            ta.static_tiles = static_tiles;
        } else {
            ta.grid_ctas = static_cast<int>(std::min<long>(roomy, std::max<long>(1, ceil_div(max_tiles, 2L * warps_per_cta))));
            ta.static_tiles = 0;
        }
    }
    {
        static const bool prefilter = [] {
            const char *env = std::getenv("TMB_NB_PREFILTER"); // A/B knob: 0 = exact f32 distance rounds in every tile
            return env == nullptr || std::atoi(env) != 0;
        }();
        ta.prefilter = prefilter;
    }
    const bool timed = timing_ && timing_used_ < timing_events_.size();
    if (timed) {
        TMB_CUDA(cudaEventRecord(timing_events_[timing_used_].first, stream));
    }
    launch_nb_tiles<Real>(ta, d_u != nullptr, d_du_dx != nullptr, d_du_dp != nullptr, stream);
    if (timed) {
        TMB_CUDA(cudaEventRecord(timing_events_[timing_used_].second, stream));
        timing_used_++;
    }

    steps_since_last_sort_++;
}

// ---------------------------------------------------------------------------------------------------------------
static const int ALL_PAIRS_STEPS_PER_SORT = 100; // reference nonbonded_all_pairs.cu:16
static const int IXN_GROUP_STEPS_PER_SORT = 200; // reference nonbonded_interaction_group.cu:17

template <typename Real>
NonbondedAllPairs<Real>::NonbondedAllPairs(
    int N, double beta, double cutoff, const std::optional<std::set<int>> &atom_idxs, bool disable_hilbert_sort,
    double nblist_padding)
    : NonbondedTiled<Real>(N, beta, cutoff, disable_hilbert_sort, nblist_padding, ALL_PAIRS_STEPS_PER_SORT),
      d_atom_idxs_(N) {
    std::vector<int> idxs;
    if (atom_idxs) {
        idxs.assign(atom_idxs->begin(), atom_idxs->end());
    } else {
        idxs.resize(N);
        std::iota(idxs.begin(), idxs.end(), 0);
    }
    set_atom_idxs(idxs);
}

template <typename Real> void NonbondedAllPairs<Real>::set_atom_idxs(const std::vector<int> &atom_idxs) {
    verify_atom_idxs(this->N_, atom_idxs);
    std::vector<unsigned int> u(atom_idxs.begin(), atom_idxs.end());
    TMB_CUDA(cudaMemcpy(d_atom_idxs_.data, u.data(), u.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    {
        std::vector<char> in_set(this->N_, 0);
        for (unsigned int a : u) {
            in_set[a] = 1;
        }
        std::vector<unsigned int> others;
        for (int a = 0; a < this->N_; a++) {
            if (!in_set[a]) {
                others.push_back(static_cast<unsigned int>(a));
            }
        }
        n_others_ = static_cast<int>(others.size());
        d_other_idxs_.realloc(others.size());
        d_other_idxs_.copy_from(others.data());
    }
    this->K_ = static_cast<int>(u.size());
    this->NR_ = this->K_;
    this->nblist_.set_all_pairs(this->K_);
    this->steps_since_last_sort_ = 0; // forces a sort, hence a rebuild, on the next evaluation
    this->force_rebuild_ = true;
    bump_launch_generation(); // K_, NR_ and the grid sizes are baked into captured launches
}

template <typename Real> bool NonbondedAllPairs<Real>::fused_prepare_hook(FusedPrepareHook &hook) {
    if (this->timing_ || this->force_rebuild_ || this->needs_sort()) {
        return false; // the next evaluation is not a plain one
    }
    FusedPrepareArgs<Real> f;
    f.K = this->K_;
    f.perm = this->d_perm_.data;
    f.others = d_other_idxs_.data;
    f.n_others = n_others_;
    f.xw = this->d_xw_.data;
    f.xw_build = this->d_xw_build_.data;
    f.box = nullptr; // filled by the caller (the Context's box)
    f.box_build = this->d_box_build_.data;
    f.padding = this->nblist_padding_;
    f.flag = this->d_flags_.data;
    f.tile_cursor = this->d_flags_.data + 1;
    f.ctr = this->nblist_.col_ctr();
    f.ext = this->nblist_.col_ext();
    f.reset_count = this->nblist_.tiles().count;
    f.reset_overflow = this->nblist_.tiles().overflow;
    if constexpr (std::is_same<Real, double>::value) {
        hook.is_double = true;
        hook.f64 = f;
    } else {
        hook.is_double = false;
        hook.f32 = f;
    }
    return true;
}

template <typename Real> std::vector<int> NonbondedAllPairs<Real>::get_atom_idxs() {
    std::vector<unsigned int> u(this->K_);
    TMB_CUDA(cudaMemcpy(u.data(), d_atom_idxs_.data, u.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    return std::vector<int>(u.begin(), u.end());
}

template <typename Real> void NonbondedAllPairs<Real>::sort(const double *d_x, const double *d_box, cudaStream_t stream) {
    if (!this->disable_hilbert_) {
        this->hilbert_->sort_device(this->K_, d_atom_idxs_.data, d_x, d_box, this->d_perm_.data, stream);
    } else {
        TMB_CUDA(cudaMemcpyAsync(
            this->d_perm_.data, d_atom_idxs_.data, this->K_ * sizeof(unsigned int), cudaMemcpyDeviceToDevice, stream));
    }
}

template <typename Real>
void NonbondedAllPairs<Real>::execute_device(
    int N, int P, const double *d_x, const double *d_p, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
    cudaStream_t stream) {
    if (N != this->N_) {
        throw std::runtime_error(
            "NonbondedAllPairs::execute_device(): expected N == N_, got N=" + std::to_string(N) +
            ", N_=" + std::to_string(this->N_));
    }
    if (P != this->N_ * P_PER_ATOM) {
        throw std::runtime_error(
            "NonbondedAllPairs::execute_device(): expected P == N_*" + std::to_string(P_PER_ATOM) + ", got P=" +
            std::to_string(P) + ", N_*" + std::to_string(P_PER_ATOM) + "=" + std::to_string(this->N_ * P_PER_ATOM));
    }
    this->run(N, d_x, d_p, d_box, d_du_dx, d_du_dp, d_u, stream);
}

// ---------------------------------------------------------------------------------------------------------------
template <typename Real>
NonbondedInteractionGroup<Real>::NonbondedInteractionGroup(
    int N, const std::vector<int> &row_atom_idxs, const std::vector<int> &col_atom_idxs, double beta, double cutoff,
    bool disable_hilbert_sort, double nblist_padding)
    : NonbondedTiled<Real>(N, beta, cutoff, disable_hilbert_sort, nblist_padding, IXN_GROUP_STEPS_PER_SORT),
      d_row_atom_idxs_(N), d_col_atom_idxs_(N) {
    validate_idxs(N, row_atom_idxs, col_atom_idxs, false);
    set_atom_idxs(row_atom_idxs, col_atom_idxs);
}

template <typename Real>
void NonbondedInteractionGroup<Real>::validate_idxs(
    int N, const std::vector<int> &rows, const std::vector<int> &cols, bool allow_empty) {
    if (!allow_empty) {
        if (rows.empty()) {
            throw std::runtime_error("row_atom_idxs must be nonempty");
        }
        if (cols.empty()) {
            throw std::runtime_error("col_atom_idxs must be nonempty");
        }
        if (rows.size() == static_cast<size_t>(N)) {
            throw std::runtime_error("must be less then N(" + std::to_string(N) + ") row indices");
        }
        if (cols.size() == static_cast<size_t>(N)) {
            throw std::runtime_error("must be less then N(" + std::to_string(N) + ") col indices");
        }
    }
    verify_atom_idxs(N, rows, allow_empty);
    verify_atom_idxs(N, cols, allow_empty);
    std::set<int> unique_rows(rows.begin(), rows.end());
    for (int c : cols) {
        if (unique_rows.count(c)) {
            throw std::runtime_error("row and col indices must be disjoint");
        }
    }
}

template <typename Real>
void NonbondedInteractionGroup<Real>::set_atom_idxs(const std::vector<int> &row_atom_idxs, const std::vector<int> &col_atom_idxs) {
    validate_idxs(this->N_, row_atom_idxs, col_atom_idxs, true);
    // rows are stored in ascending order (the reference passes them through a std::set)
    std::set<unsigned int> row_set(row_atom_idxs.begin(), row_atom_idxs.end());
    std::vector<unsigned int> rows(row_set.begin(), row_set.end());
    std::vector<unsigned int> cols(col_atom_idxs.begin(), col_atom_idxs.end());
    const int NR = static_cast<int>(rows.size());
    const int NC = static_cast<int>(cols.size());
    if (NR + NC > this->N_) {
        throw std::runtime_error("number of idxs must be less than or equal to N");
    }
    if (NR > 0 && NC > 0) {
        TMB_CUDA(cudaMemcpy(d_row_atom_idxs_.data, rows.data(), NR * sizeof(unsigned int), cudaMemcpyHostToDevice));
        TMB_CUDA(cudaMemcpy(d_col_atom_idxs_.data, cols.data(), NC * sizeof(unsigned int), cudaMemcpyHostToDevice));
        this->nblist_.resize(NR + NC);
        this->nblist_.set_contiguous_split(NR, NC);
    }
    this->NR_ = NR;
    NC_ = NC;
    this->K_ = NR + NC;
    this->steps_since_last_sort_ = 0;
    this->force_rebuild_ = true;
    bump_launch_generation();
}

template <typename Real>
void NonbondedInteractionGroup<Real>::sort(const double *d_x, const double *d_box, cudaStream_t stream) {
    unsigned int *perm = this->d_perm_.data;
    if (!this->disable_hilbert_) {
        // rows and columns are sorted separately; the gathered set is [rows | cols]
        this->hilbert_->sort_device(this->NR_, d_row_atom_idxs_.data, d_x, d_box, perm, stream);
        this->hilbert_->sort_device(NC_, d_col_atom_idxs_.data, d_x, d_box, perm + this->NR_, stream);
    } else {
        TMB_CUDA(cudaMemcpyAsync(perm, d_row_atom_idxs_.data, this->NR_ * sizeof(unsigned int), cudaMemcpyDeviceToDevice, stream));
        TMB_CUDA(cudaMemcpyAsync(
            perm + this->NR_, d_col_atom_idxs_.data, NC_ * sizeof(unsigned int), cudaMemcpyDeviceToDevice, stream));
    }
}

template <typename Real>
void NonbondedInteractionGroup<Real>::execute_device(
    int N, int P, const double *d_x, const double *d_p, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
    cudaStream_t stream) {
    if (N != this->N_) {
        throw std::runtime_error(
            "NonbondedInteractionGroup::execute_device(): expected N == N_, got N=" + std::to_string(N) +
            ", N_=" + std::to_string(this->N_));
    }
    if (P != this->N_ * P_PER_ATOM) {
        throw std::runtime_error(
            "NonbondedInteractionGroup::execute_device(): expected P == N_*" + std::to_string(P_PER_ATOM) +
            ", got P=" + std::to_string(P) + ", N_*" + std::to_string(P_PER_ATOM) + "=" +
            std::to_string(this->N_ * P_PER_ATOM));
    }
    if (this->NR_ == 0 || NC_ == 0) {
        return; // nothing interacts; the caller's energy buffer was zeroed (reference :176-179)
    }
    this->run(N, d_x, d_p, d_box, d_du_dx, d_du_dp, d_u, stream);
}

// ---------------------------------------------------------------------------------------------------------------
template <typename Real, bool Negated>
NonbondedPairList<Real, Negated>::NonbondedPairList(
    const std::vector<int> &pair_idxs, const std::vector<double> &scales, double beta, double cutoff)
    : M_(static_cast<int>(pair_idxs.size() / 2)), beta_(beta), cutoff_(cutoff), d_pair_idxs_(pair_idxs.size()),
      d_scales_(scales.size()), d_partials_(sm_count() * 8), d_ticket_(1) {
    if (pair_idxs.size() % 2 != 0) {
        throw std::runtime_error("pair_idxs.size() must be even, but got " + std::to_string(pair_idxs.size()));
    }
    for (int i = 0; i < M_; i++) {
        const int src = pair_idxs[i * 2 + 0];
        const int dst = pair_idxs[i * 2 + 1];
        if (src == dst) {
            throw std::runtime_error("illegal pair with src == dst: " + std::to_string(src) + ", " + std::to_string(dst));
        }
    }
    if (static_cast<int>(scales.size() / 2) != M_) {
        throw std::runtime_error(
            "expected same number of pairs and scale tuples, but got " + std::to_string(M_) +
            " != " + std::to_string(scales.size() / 2));
    }
    d_pair_idxs_.copy_from(pair_idxs.data());
    d_scales_.copy_from(scales.data());
    d_ticket_.zero();
    TMB_CUDA(cudaDeviceSynchronize());
}

template <typename Real, bool Negated>
void NonbondedPairList<Real, Negated>::execute_device(
    int, int, const double *d_x, const double *d_p, const double *d_box, u64 *d_du_dx, u64 *d_du_dp, i128 *d_u,
    cudaStream_t stream) {
    if (M_ <= 0) {
        return;
    }
    PairListArgs<Real> a;
    a.M = M_;
    a.x = d_x;
    a.p = d_p;
    a.box = d_box;
    a.pair_idxs = d_pair_idxs_.data;
    a.scales = d_scales_.data;
    a.beta = beta_;
    a.cutoff = cutoff_;
    a.negated = Negated;
    a.du_dx = d_du_dx;
    a.du_dp = d_du_dp;
    a.u_partials = d_partials_.data;
    a.ticket = d_ticket_.data;
    a.d_u = d_u;
    launch_pair_list<Real>(a, stream);
}

template <typename Real, bool Negated>
void NonbondedPairList<Real, Negated>::du_dp_fixed_to_float(int N, int P, const u64 *du_dp, double *out) const {
    nonbonded_du_dp_fixed_to_float(N, P, du_dp, out);
}

template class NonbondedTiled<float>;
template class NonbondedTiled<double>;
template class NonbondedAllPairs<float>;
template class NonbondedAllPairs<double>;
template class NonbondedInteractionGroup<float>;
template class NonbondedInteractionGroup<double>;
template class NonbondedPairList<float, true>;
template class NonbondedPairList<float, false>;
template class NonbondedPairList<double, true>;
template class NonbondedPairList<double, false>;

} // namespace tmb
