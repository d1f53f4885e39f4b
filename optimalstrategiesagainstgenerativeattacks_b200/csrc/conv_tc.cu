// placeholder until the tcgen05 path lands (next commit): reports "unsupported" so GIM_ALGO_AUTO takes the CUDA-core path
#include "common.cuh"
namespace gim {
bool conv_tc_supported(int, int, int, int, int, int, int) { return false; }
bool wgrad_tc_supported(int, int, int, int, int, int, int) { return false; }
int conv_fwd_tc(const void*, const void*, const float*, void*, int, int, int, int, int, int, cudaStream_t) { return fail(GIM_E_UNSUPPORTED, "tcgen05 conv not built"); }
int conv_wgrad_tc(const void*, const void*, float*, int, int, int, int, int, int, cudaStream_t) { return fail(GIM_E_UNSUPPORTED, "tcgen05 wgrad not built"); }
}
