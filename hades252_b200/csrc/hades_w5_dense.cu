// Width-5 kernels, dense schedule = the reference's round structure (A/B baseline; own constant bank).
#define HADES_W 5
#define HADES_ALGO 0
#include "width_impl.cuh"
namespace hades {
const WidthOps* width_ops_5_dense() { return &kOps; }
const WidthOps* width_ops_5_opt();
const WidthOps* width_ops_5_ccf();
const WidthOps* width_ops_5(int algo) {
    return algo == 0 ? width_ops_5_dense() : algo == 1 ? width_ops_5_opt() : width_ops_5_ccf();
}
}  // namespace hades
