// Strip-kernel instantiations: float, NGLL 6 (see launch_strip_case in strip_kernels.cuh).
#define S2D_STRIP_CASES
#include "strip_kernels.cuh"
namespace s2d {
#ifndef S2D_ONLY_N5
template void launch_strip_case<float, 6>(const StripGeom&, const StripIO<float>&, cudaStream_t);
#endif
}  // namespace s2d
