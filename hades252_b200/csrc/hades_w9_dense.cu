// Width-9 kernels, dense schedule = the reference's round structure (A/B baseline; own constant bank).
#define HADES_W 9
#define HADES_ALGO 0
#include "width_impl.cuh"
namespace hades {
const WidthOps* width_ops_9_dense() { return &kOps; }
const WidthOps* width_ops_9_opt();
const WidthOps* width_ops_9_ccf();
const WidthOps* width_ops_9(int algo) {
    return algo == 0 ? width_ops_9_dense() : algo == 1 ? width_ops_9_opt() : width_ops_9_ccf();
}
}  // namespace hades
