// Test-infrastructure stub (NOT product code).
// The reference's rmsd_align.cpp needs Eigen, which is not available offline and is
// off the hot path. wrap_kernels.cpp still references the symbol, so the
// out-of-tree oracle build links this throwing stand-in instead.
#include <stdexcept>
namespace timemachine {
void rmsd_align_cpu(const int, const double *, const double *, double *) {
    throw std::runtime_error("rmsd_align is not built into the oracle copy of custom_ops (needs Eigen)");
}
} // namespace timemachine
