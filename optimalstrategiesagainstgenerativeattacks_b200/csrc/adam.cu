// Fused multi-tensor Adam: one launch updates every parameter of an optimizer (torch.optim.Adam semantics, no weight decay /
// amsgrad; gim_img_trainer.py:50-58, gim_gaussian_trainer.py:48-49).  HBM bound: reads p,g,m,v (16 B/param), writes p,m,v (12 B).
#include "common.cuh"

namespace gim {

__global__ void __launch_bounds__(256) adam_multi_kernel(const gim_adam_tensor* __restrict__ table, const float* __restrict__ lrs,
                                                         const long long* __restrict__ step, float beta1, float beta2, float eps, float grad_scale) {
    __shared__ float s_step_size, s_bc2_sqrt;
    const gim_adam_tensor t = table[blockIdx.y];
    if ((long long)blockIdx.x * blockDim.x >= t.numel) return;
    if (threadIdx.x == 0) {
        double k = (double)(*step + 1);
        double bc1 = 1.0 - pow((double)beta1, k);
        double bc2 = 1.0 - pow((double)beta2, k);
        s_step_size = (float)((double)lrs[t.group] / bc1);
        s_bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool vec = ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0);
    const long long nv = vec ? t.numel / 4 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        float4 p = reinterpret_cast<float4*>(t.p)[i];
        float4 g = reinterpret_cast<const float4*>(t.g)[i];
        float4 m = reinterpret_cast<float4*>(t.m)[i];
        float4 v = reinterpret_cast<float4*>(t.v)[i];
        float* pp = &p.x; float* gg = &g.x; float* mm = &m.x; float* vv = &v.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gr = gg[j] * grad_scale;
            mm[j] = mm[j] + (gr - mm[j]) * omb1;
            vv[j] = vv[j] * beta2 + omb2 * gr * gr;
            float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
            pp[j] = pp[j] - step_size * (mm[j] / denom);
        }
        reinterpret_cast<float4*>(t.p)[i] = p;
        reinterpret_cast<float4*>(t.m)[i] = m;
        reinterpret_cast<float4*>(t.v)[i] = v;
    }
    for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.numel; i += stride) {
        float gr = t.g[i] * grad_scale;
        float m = t.m[i] + (gr - t.m[i]) * omb1;
        float v = t.v[i] * beta2 + omb2 * gr * gr;
        float denom = sqrtf(v) / bc2_sqrt + eps;
        t.p[i] = t.p[i] - step_size * (m / denom);
        t.m[i] = m;
        t.v[i] = v;
    }
}

__global__ void adam_step_inc_kernel(long long* step) { *step += 1; }

// zero every gradient of the optimizer in one launch (optimizer.zero_grad() with gradients kept in place)
__global__ void __launch_bounds__(256) zero_grads_multi_kernel(const gim_adam_tensor* __restrict__ table) {
    const gim_adam_tensor t = table[blockIdx.y];
    if ((long long)blockIdx.x * blockDim.x >= t.numel) return;
    const long long stride = (long long)gridDim.x * blockDim.x;
    float* g = const_cast<float*>(t.g);                  // the table is shared with the Adam kernel, which only reads gradients
    const bool vec = (((uintptr_t)g & 15) == 0);
    const long long nv = vec ? t.numel / 4 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = nv * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.numel; i += stride) g[i] = 0.f;
}

}  // namespace gim

using namespace gim;

extern "C" int gim_adam_multi(const gim_adam_tensor* table, int n_tensors, long long max_numel, const float* lrs, long long* step, float beta1,
                              float beta2, float eps, float grad_scale, gim_stream_t s) {
    if (n_tensors <= 0) return GIM_OK;
    GIM_REQUIRE(n_tensors <= 65535, "adam: too many tensors in one launch");
    long long gx = (max_numel + 256 * 4 * 4 - 1) / (256 * 4 * 4);
    if (gx < 1) gx = 1;
    if (gx > 128) gx = 128;
    dim3 grid((unsigned)gx, n_tensors);
    adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(table, lrs, step, beta1, beta2, eps, grad_scale);
    int rc = check_launch("adam_multi");
    if (rc != GIM_OK) return rc;
    adam_step_inc_kernel<<<1, 1, 0, (cudaStream_t)s>>>(step);
    return check_launch("adam_step_inc");
}

extern "C" int gim_zero_grads_multi(const gim_adam_tensor* table, int n_tensors, long long max_numel, gim_stream_t s) {
    if (n_tensors <= 0) return GIM_OK;
    GIM_REQUIRE(n_tensors <= 65535, "zero_grads: too many tensors in one launch");
    long long gx = (max_numel + 256 * 4 * 4 - 1) / (256 * 4 * 4);
    if (gx < 1) gx = 1;
    if (gx > 128) gx = 128;
    zero_grads_multi_kernel<<<dim3((unsigned)gx, n_tensors), 256, 0, (cudaStream_t)s>>>(table);
    return check_launch("zero_grads_multi");
}
