// Strip-kernel instantiations: double, NGLL 7, 8 (see launch_strip_case in strip_kernels.cuh).
#define S2D_STRIP_CASES
#include "strip_kernels.cuh"
namespace s2d {
#ifndef S2D_ONLY_N5
template void launch_strip_case<double, 7>(const StripGeom&, const StripIO<double>&, cudaStream_t);
#endif
#ifndef S2D_ONLY_N5
template void launch_strip_case<double, 8>(const StripGeom&, const StripIO<double>&, cudaStream_t);
#endif
}  // namespace s2d
