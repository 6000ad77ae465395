// Width-3 kernels, canonical-form schedule (own translation unit = own 64 KB __constant__ bank).
#define HADES_W 3
#define HADES_ALGO 2
#include "width_impl.cuh"
namespace hades {
const WidthOps* width_ops_3_ccf() { return &kOps; }
}  // namespace hades
