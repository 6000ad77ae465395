// Width-3 kernels of the Hades252 engine (own translation unit = own 64 KB __constant__ bank).
#define HADES_W 3
#include "width_impl.cuh"
namespace hades {
const WidthOps* width_ops_3() { return &kOps; }
}  // namespace hades
