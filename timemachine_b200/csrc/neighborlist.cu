// Neighborlist and HilbertSort host classes (reference neighborlist.cu:14-381, hilbert_sort.cu:13-104).
#include "hilbert_curve.h"
#include "potential.hpp"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <numeric>

namespace tmb {

// ---------------------------------------------------------------------------------------------------------------
static std::shared_ptr<DeviceBuffer<unsigned int>> shared_hilbert_lut() {
    static std::mutex mu;
    static std::weak_ptr<DeviceBuffer<unsigned int>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto sp = cache.lock();
    if (!sp) {
        constexpr size_t G = HILBERT_GRID_DIM;
        sp = std::make_shared<DeviceBuffer<unsigned int>>(G * G * G);
        launch_hilbert_lut(sp->data, main_stream());
        TMB_CUDA(cudaStreamSynchronize(main_stream()));
        cache = sp;
    }
    return sp;
}

HilbertSort::HilbertSort(int N) : N_(N), lut_(shared_hilbert_lut()), keys_in_(N), keys_out_(N), vals_in_(N) {
    temp_.realloc(radix_sort_pairs_temp_bytes(N));
}

void HilbertSort::sort_device(
    int n, const unsigned int *d_atom_idxs, const double *d_coords, const double *d_box, unsigned int *d_perm_out,
    cudaStream_t stream) {
    if (n > N_) {
        throw std::runtime_error("number of idxs to sort must be less than or equal to N");
    }
    if (n <= 0) {
        return;
    }
    launch_hilbert_keys(n, d_atom_idxs, d_coords, d_box, lut_->data, keys_in_.data, vals_in_.data, stream);
    radix_sort_pairs(temp_.data, temp_.length, keys_in_.data, keys_out_.data, vals_in_.data, d_perm_out, n, stream);
}

std::vector<unsigned int> HilbertSort::sort_host(int n, const double *h_coords, const double *h_box) {
    std::vector<unsigned int> idxs(n);
    std::iota(idxs.begin(), idxs.end(), 0u);
    DeviceBuffer<double> d_coords(static_cast<size_t>(n) * 3), d_box(9);
    DeviceBuffer<unsigned int> d_idxs(n), d_perm(n);
    d_coords.copy_from(h_coords);
    d_box.copy_from(h_box);
    d_idxs.copy_from(idxs.data());
    cudaStream_t stream = main_stream();
    sort_device(n, d_idxs.data, d_coords.data, d_box.data, d_perm.data, stream);
    TMB_CUDA(cudaStreamSynchronize(stream));
    d_perm.copy_to(idxs.data());
    return idxs;
}

// ---------------------------------------------------------------------------------------------------------------
static size_t worst_case_tiles(int max_size) {
    // every row block against every column block at or after it, plus one partial tile per row block
    const size_t nb = ceil_div(max_size, TILE);
    return nb * (nb + 1) / 2 + nb;
}

template <typename Real>
Neighborlist<Real>::Neighborlist(int N)
    : max_size_(N), N_(N), NC_(N), NR_(N), d_row_idxs_(N > 0 ? N : 0), d_col_idxs_(N > 0 ? N : 0) {
    if (N <= 0) {
        throw std::runtime_error("Neighborlist N must be at least 1");
    }
    const size_t nb = ceil_div(N, TILE);
    d_row_ctr_.realloc(nb * 3);
    d_row_ext_.realloc(nb * 3);
    d_col_ctr_.realloc(nb * 3);
    d_col_ext_.realloc(nb * 3);
    d_count_.realloc(1);
    d_overflow_.realloc(2);
    d_count_.zero();
    d_overflow_.zero();
    // Worst case for an all-pairs list is nb(nb+1)/2 tiles; a row/column split can reach nb_r * nb_c <= (nb+1)^2/4+nb:
    // O((N/32)^2), what the reference allocates (neighborlist.cu:22-28: 58 MB at 30k atoms, 522 MB at 90k).  A list at
    // liquid density with cutoff + padding = 1.3 nm holds about ONE tile per atom (measured: 1.03 N at 23k-90k atoms), so
    // the buffer starts at 4 tiles per atom (4x the measured count: 16 MB at 30k, 48 MB at 90k) and never above the worst
    // case; a build that needs more records how much (sticky word of `overflow`) and the host grows the buffer and redoes
    // the evaluation at its next synchronisation point (recover_overflow).
    worst_case_ = std::max(worst_case_tiles(N), (nb + 1) * (nb + 1) / 4 + nb + 1);
    size_t per_atom = 4;
    if (const char *env = std::getenv("TMB_NBLIST_TILES_PER_ATOM_X100")) { // tests: start small to exercise the growth path
        per_atom = 0;
        set_capacity(std::max<size_t>(1, static_cast<size_t>(std::atoll(env)) * N / 100));
    }
    if (per_atom > 0) {
        set_capacity(std::min(worst_case_, per_atom * static_cast<size_t>(N) + 1024));
    }
    tiles_.count = d_count_.data;
    tiles_.overflow = d_overflow_.data;
    TMB_CUDA(cudaDeviceSynchronize());
}

template <typename Real> void Neighborlist<Real>::set_capacity(size_t cap) {
    cap = std::min(cap, worst_case_);
    d_rows_.realloc(cap);
    d_cols_.realloc(cap * TILE);
    tiles_.rows = d_rows_.data;
    tiles_.cols = d_cols_.data;
    tiles_.capacity = static_cast<unsigned int>(cap);
}

template <typename Real> bool Neighborlist<Real>::recover_overflow() {
    unsigned int needed = 0;
    TMB_CUDA(cudaMemcpy(&needed, d_overflow_.data + 1, sizeof(needed), cudaMemcpyDeviceToHost));
    if (needed == 0) {
        return false;
    }
    TMB_CUDA(cudaDeviceSynchronize());
    set_capacity(std::max<size_t>(2 * static_cast<size_t>(needed), 2 * static_cast<size_t>(tiles_.capacity)));
    d_overflow_.zero();
    d_count_.zero();
    TMB_CUDA(cudaDeviceSynchronize());
    return true;
}

template <typename Real> void Neighborlist<Real>::set_all_pairs(int K) {
    if (K <= 0) {
        throw std::runtime_error("size is must be at least 1");
    }
    if (K > max_size_) {
        throw std::runtime_error("size is greater than max size: " + std::to_string(K) + " > " + std::to_string(max_size_));
    }
    N_ = K;
    NR_ = K;
    NC_ = K;
    contiguous_ = true;
    row_base_ = 0;
    col_base_ = 0;
}

template <typename Real> void Neighborlist<Real>::resize(int size) { set_all_pairs(size); }

template <typename Real> void Neighborlist<Real>::reset_row_idxs() { set_all_pairs(N_); }

template <typename Real> void Neighborlist<Real>::set_contiguous_split(int NR, int NC) {
    if (NC + NR != N_) {
        throw std::runtime_error("Total of indices must equal N");
    }
    if (NC == 0 || NR == 0) {
        throw std::runtime_error("Number of column and row indices must be non-zero");
    }
    NR_ = NR;
    NC_ = NC;
    contiguous_ = true;
    row_base_ = 0;
    col_base_ = NR;
}

template <typename Real> void Neighborlist<Real>::set_row_idxs(std::vector<unsigned int> row_idxs) {
    if (row_idxs.empty()) {
        throw std::runtime_error("idxs can't be empty");
    }
    std::set<unsigned int> unique_idxs(row_idxs.begin(), row_idxs.end());
    if (unique_idxs.size() != row_idxs.size()) {
        throw std::runtime_error("atom indices must be unique");
    }
    if (static_cast<int>(row_idxs.size()) >= N_) {
        throw std::runtime_error("number of idxs must be less than N");
    }
    if (*std::max_element(row_idxs.begin(), row_idxs.end()) >= static_cast<unsigned int>(N_)) {
        throw std::runtime_error("indices values must be less than N");
    }
    // columns = complement of the rows, ascending
    std::vector<unsigned int> cols;
    cols.reserve(N_ - row_idxs.size());
    for (unsigned int i = 0; i < static_cast<unsigned int>(N_); i++) {
        if (!unique_idxs.count(i)) {
            cols.push_back(i);
        }
    }
    TMB_CUDA(cudaMemcpy(d_row_idxs_.data, row_idxs.data(), row_idxs.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    TMB_CUDA(cudaMemcpy(d_col_idxs_.data, cols.data(), cols.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    NR_ = static_cast<int>(row_idxs.size());
    NC_ = static_cast<int>(cols.size());
    contiguous_ = false;
}

template <typename Real> unsigned int Neighborlist<Real>::num_tile_ixns() {
    unsigned int h = 0;
    TMB_CUDA(cudaMemcpy(&h, d_count_.data, sizeof(h), cudaMemcpyDeviceToHost));
    return std::min(h, tiles_.capacity);
}

template <typename Real> int Neighborlist<Real>::max_ixn_count() const {
    const int nb = ceil_div(max_size_, TILE);
    return (nb * (nb + 1)) / 2 * TILE; // reference neighborlist.cu:369-376
}

template <typename Real>
void Neighborlist<Real>::build_device(
    const double *d_coords, const Vec4<Real> *d_xw, const double *d_box, double cutoff, const unsigned int *flag,
    cudaStream_t stream, const Snapshot *snap) {
    const bool tri = upper_triangular();
    const bool fused = snap != nullptr && snap->bounds_done;
    if (fused && !(tri && contiguous_ && col_base_ == 0)) {
        throw std::runtime_error("Neighborlist: fused bounds need the all-pairs layout");
    }
    BlockBoundsArgs<Real> ba;
    ba.num_blocks = num_col_blocks();
    ba.num_idxs = NC_;
    ba.idxs = contiguous_ ? nullptr : d_col_idxs_.data;
    ba.base = col_base_;
    ba.coords = d_coords;
    ba.xw = d_xw;
    ba.box = d_box;
    ba.ctr = d_col_ctr_.data;
    ba.ext = d_col_ext_.data;
    ba.flag = flag;
    // the first bounds launch also clears the tile counter and snapshots the build-time state
    ba.reset_count = tiles_.count;
    ba.reset_overflow = tiles_.overflow;
    const bool can_snapshot = snap != nullptr && contiguous_ && d_xw != nullptr;
    ba.xw_build = can_snapshot ? snap->xw_build : nullptr;
    ba.box_build = can_snapshot ? snap->box_build : nullptr;
    if (!fused) {
        launch_block_bounds<Real>(ba, stream);
    }
    if (!tri) {
        ba.num_blocks = num_row_blocks();
        ba.num_idxs = NR_;
        ba.idxs = contiguous_ ? nullptr : d_row_idxs_.data;
        ba.base = row_base_;
        ba.ctr = d_row_ctr_.data;
        ba.ext = d_row_ext_.data;
        ba.reset_count = nullptr;
        ba.reset_overflow = nullptr;
        ba.box_build = nullptr;
        launch_block_bounds<Real>(ba, stream);
    }

    BuildTilesArgs<Real> ta;
    ta.N = N_;
    ta.NC = NC_;
    ta.NR = NR_;
    ta.upper_triangular = tri;
    ta.col_idxs = contiguous_ ? nullptr : d_col_idxs_.data;
    ta.col_base = col_base_;
    ta.row_idxs = contiguous_ ? nullptr : d_row_idxs_.data;
    ta.row_base = row_base_;
    ta.col_ctr = d_col_ctr_.data;
    ta.col_ext = d_col_ext_.data;
    ta.row_ctr = tri ? d_col_ctr_.data : d_row_ctr_.data;
    ta.row_ext = tri ? d_col_ext_.data : d_row_ext_.data;
    ta.coords = d_coords;
    ta.xw = d_xw;
    ta.box = d_box;
    ta.cutoff = cutoff;
    ta.tiles = tiles_;
    ta.flag = flag;
    if (fused) {
        ta.snap_xw_build = snap->xw_build;
        ta.snap_box_build = snap->box_build;
        ta.snap_slots = snap->slots;
    }
    launch_build_tiles<Real>(ta, stream);
}

template <typename Real>
std::vector<std::vector<int>>
Neighborlist<Real>::get_nblist_host(int N, const double *h_coords, const double *h_box, double cutoff) {
    if (N != N_) {
        throw std::runtime_error("N != N_");
    }
    DeviceBuffer<double> d_coords(static_cast<size_t>(N) * 3), d_box(9);
    d_coords.copy_from(h_coords);
    d_box.copy_from(h_box);
    cudaStream_t stream = main_stream();
    for (int attempt = 0;; attempt++) {
        build_device(d_coords.data, nullptr, d_box.data, cutoff, nullptr, stream);
        TMB_CUDA(cudaStreamSynchronize(stream));
        if (!recover_overflow()) {
            break;
        }
        if (attempt >= 8) {
            throw std::runtime_error("neighborlist tile buffer overflow");
        }
    }
    const unsigned int T = num_tile_ixns();
    std::vector<int> rows(T);
    std::vector<unsigned int> cols(static_cast<size_t>(T) * TILE);
    if (T > 0) {
        TMB_CUDA(cudaMemcpy(rows.data(), d_rows_.data, T * sizeof(int), cudaMemcpyDeviceToHost));
        TMB_CUDA(cudaMemcpy(cols.data(), d_cols_.data, cols.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    }
    std::vector<std::vector<int>> ixn_list(num_row_blocks());
    for (unsigned int t = 0; t < T; t++) {
        for (int j = 0; j < TILE; j++) {
            const unsigned int atom = cols[static_cast<size_t>(t) * TILE + j];
            if (atom < static_cast<unsigned int>(N)) {
                ixn_list[rows[t]].push_back(static_cast<int>(atom));
            }
        }
    }
    return ixn_list;
}

template <typename Real>
void Neighborlist<Real>::compute_block_bounds_host(
    int N, const double *h_coords, const double *h_box, double *h_ctr, double *h_ext) {
    DeviceBuffer<double> d_coords(static_cast<size_t>(N) * 3), d_box(9);
    d_coords.copy_from(h_coords);
    d_box.copy_from(h_box);
    cudaStream_t stream = main_stream();
    BlockBoundsArgs<Real> ba;
    ba.num_blocks = num_col_blocks();
    ba.num_idxs = NC_;
    ba.idxs = contiguous_ ? nullptr : d_col_idxs_.data;
    ba.base = col_base_;
    ba.coords = d_coords.data;
    ba.xw = nullptr;
    ba.box = d_box.data;
    ba.ctr = d_col_ctr_.data;
    ba.ext = d_col_ext_.data;
    ba.flag = nullptr;
    ba.reset_count = nullptr;
    ba.reset_overflow = nullptr;
    ba.xw_build = nullptr;
    ba.box_build = nullptr;
    launch_block_bounds<Real>(ba, stream);
    TMB_CUDA(cudaStreamSynchronize(stream));
    const size_t n = static_cast<size_t>(num_col_blocks()) * 3;
    std::vector<Real> ctr(n), ext(n);
    TMB_CUDA(cudaMemcpy(ctr.data(), d_col_ctr_.data, n * sizeof(Real), cudaMemcpyDeviceToHost));
    TMB_CUDA(cudaMemcpy(ext.data(), d_col_ext_.data, n * sizeof(Real), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) {
        h_ctr[i] = ctr[i];
        h_ext[i] = ext[i];
    }
}

template class Neighborlist<float>;
template class Neighborlist<double>;

} // namespace tmb
