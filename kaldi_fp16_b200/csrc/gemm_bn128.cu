// tcgen05 GEMM kernels with a 128-column tile (see gemm_launch.cuh / gemm_sm100.cuh)
#include "gemm_launch.cuh"
namespace kfp16 {
template bool launch_gemm_bn<128>(kfp16_ctx*, const GemmParams&, const GemmLaunch&);
}
