// Width-3 kernels, dense schedule = the reference's round structure (A/B baseline; own constant bank).
#define HADES_W 3
#define HADES_ALGO 0
#include "width_impl.cuh"
namespace hades {
const WidthOps* width_ops_3_dense() { return &kOps; }
const WidthOps* width_ops_3_opt();
const WidthOps* width_ops_3_ccf();
const WidthOps* width_ops_3(int algo) {
    return algo == 0 ? width_ops_3_dense() : algo == 1 ? width_ops_3_opt() : width_ops_3_ccf();
}
}  // namespace hades
