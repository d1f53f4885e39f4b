// InstanceNorm2d / ada_in as (statistics) + (per-(n,c) affine + LeakyReLU) so the apply step can later be folded into
// the consumer conv's operand producer.  Statistics are always fp32; NHWC so a warp reads 32 adjacent channels.
#include "common.cuh"

namespace gim {

// grid (ceil(c/32), n); block (32, 8).  ONE pass over the plane with pivot-shifted sums, pivot = the plane's first pixel:
//   d = x - pivot,  mean = pivot + sum(d)/hw,  M2 = sum(d^2) - sum(d)^2/hw
// (the shift keeps the subtraction in M2 benign -- d is of the order of the spread, not of the mean -- and halves the DRAM traffic of
// the two-pass form, whose second pass missed L2: ncu measured 1.92x the algorithmic bytes).  On a spatially CONSTANT plane every d is
// exactly zero, so mean == x, M2 == 0 and the centred values are exactly 0 -- as in exact arithmetic.  That case is not exotic: with the
// reference's initialisation (InstanceNorm bias 0) the right branch of every EnvDecoder block is spatially constant, the variance is
// 0 and rsqrt(eps) = 316 multiplies whatever round-off the mean carries (a plain running sum gives 3x != x + x + x).
template <typename T>
__global__ void __launch_bounds__(256) norm_stats_kernel(const T* __restrict__ x, float* __restrict__ mean, float* __restrict__ m2, int hw, int c) {
    __shared__ float sh[8][33], sh2[8][33];
    int ch = blockIdx.x * 32 + threadIdx.x;
    long long img = blockIdx.y;
    const T* xi = x + img * (long long)hw * c;
    const float pivot = ch < c ? to_f<T>(xi[ch]) : 0.f;
    float s = 0.f, q = 0.f;
    if (ch < c) {
        int p = threadIdx.y;
        for (; p + 24 < hw; p += 32) {                     // four independent loads in flight per thread
            const float d0 = to_f<T>(xi[(long long)p * c + ch]) - pivot, d1 = to_f<T>(xi[(long long)(p + 8) * c + ch]) - pivot;
            const float d2 = to_f<T>(xi[(long long)(p + 16) * c + ch]) - pivot, d3 = to_f<T>(xi[(long long)(p + 24) * c + ch]) - pivot;
            s += (d0 + d1) + (d2 + d3);
            q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
        for (; p < hw; p += 8) { const float d = to_f<T>(xi[(long long)p * c + ch]) - pivot; s += d; q += d * d; }
    }
    sh[threadIdx.y][threadIdx.x] = s;
    sh2[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { ts += sh[j][threadIdx.x]; tq += sh2[j][threadIdx.x]; }
        mean[img * c + ch] = pivot + ts / (float)hw;
        m2[img * c + ch] = fmaxf(tq - ts * ts / (float)hw, 0.f);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) affine_act_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ a,
                                                         const float* __restrict__ b, T* __restrict__ y, long long total, int hw, int c, float slope) {
    long long stride = (long long)gridDim.x * blockDim.x;
    long long plane = (long long)hw * c;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % c);
        long long k = (i / plane) * c + ch;
        float v = a[k] * (to_f<T>(x[i]) - mean[k]) + b[k];      // centred form: exact for the degenerate 1x1 map
        y[i] = from_f<T>(lrelu_f(v, slope));
    }
}

// fp32, c % 4 == 0: grid (chunks of the plane, n); one float4 of adjacent channels per thread and step
__global__ void __launch_bounds__(256) affine_act_vec4_kernel(const float4* __restrict__ x, const float4* __restrict__ mean, const float4* __restrict__ a,
                                                              const float4* __restrict__ b, float4* __restrict__ y, int plane4, int c4, float slope) {
    const long long base = (long long)blockIdx.y * plane4;
    const float4* mi = mean + (long long)blockIdx.y * c4;
    const float4* ai = a + (long long)blockIdx.y * c4;
    const float4* bi = b + (long long)blockIdx.y * c4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane4; i += gridDim.x * blockDim.x) {
        const int k = i % c4;
        const float4 xv = __ldg(x + base + i), m = __ldg(mi + k), av = __ldg(ai + k), bv = __ldg(bi + k);
        float4 o;
        o.x = lrelu_f(av.x * (xv.x - m.x) + bv.x, slope);
        o.y = lrelu_f(av.y * (xv.y - m.y) + bv.y, slope);
        o.z = lrelu_f(av.z * (xv.z - m.z) + bv.z, slope);
        o.w = lrelu_f(av.w * (xv.w - m.w) + bv.w, slope);
        y[base + i] = o;
    }
}
__global__ void __launch_bounds__(256) norm_bwd_apply_vec4_kernel(const float4* __restrict__ gy, const float4* __restrict__ x, const float4* __restrict__ y,
                                                                  const float4* __restrict__ mean, const float4* __restrict__ A, const float4* __restrict__ B,
                                                                  const float4* __restrict__ C, float4* __restrict__ gx, int plane4, int c4, float slope,
                                                                  const float4* __restrict__ ca, const float4* __restrict__ cb) {
    const long long base = (long long)blockIdx.y * plane4;
    const long long kb = (long long)blockIdx.y * c4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane4; i += gridDim.x * blockDim.x) {
        const int k = i % c4;
        float4 g = __ldg(gy + base + i);
        const float4 xv = __ldg(x + base + i), m = __ldg(mean + kb + k), av = __ldg(A + kb + k), bv = __ldg(B + kb + k), cv = __ldg(C + kb + k);
        if (y != nullptr) {
            const float4 yv = __ldg(y + base + i);
            if (!(yv.x > 0.f)) g.x *= slope;
            if (!(yv.y > 0.f)) g.y *= slope;
            if (!(yv.z > 0.f)) g.z *= slope;
            if (!(yv.w > 0.f)) g.w *= slope;
        } else if (ca != nullptr) {                  // the forward output was never materialised: recompute its sign
            const float4 fa = __ldg(ca + kb + k), fb = __ldg(cb + kb + k);
            if (!(fa.x * (xv.x - m.x) + fb.x > 0.f)) g.x *= slope;
            if (!(fa.y * (xv.y - m.y) + fb.y > 0.f)) g.y *= slope;
            if (!(fa.z * (xv.z - m.z) + fb.z > 0.f)) g.z *= slope;
            if (!(fa.w * (xv.w - m.w) + fb.w > 0.f)) g.w *= slope;
        }
        float4 o;
        o.x = av.x * g.x + bv.x * (xv.x - m.x) + cv.x;
        o.y = av.y * g.y + bv.y * (xv.y - m.y) + cv.y;
        o.z = av.z * g.z + bv.z * (xv.z - m.z) + cv.z;
        o.w = av.w * g.w + bv.w * (xv.w - m.w) + cv.w;
        gx[base + i] = o;
    }
}

// The norm -> activation -> (nearest x2 upsample) -> bf16 conv operand chain of the attacker's blocks (reference model_blocks.py:760-768,
// 805-811, 851-861) in ONE pass: out[n, (2)h, (2)w, c] = bf16(LeakyReLU(a (x - mean) + b)).  The fp32 normalised activation never exists.
// grid (chunks of the plane, n); thread = 8 adjacent channels (two 16-byte loads, one 16-byte store -- four when upsampling).
__global__ void __launch_bounds__(256) norm_act_operand_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ a,
                                                               const float* __restrict__ b, bf16* __restrict__ out, int h, int w, int c, float slope, int upsample) {
    const int c8 = c >> 3, plane8 = h * w * c8;
    const float* xi = x + (long long)blockIdx.y * h * w * c;
    const long long kb = (long long)blockIdx.y * c;
    bf16* oi = out + (long long)blockIdx.y * h * w * c * (upsample ? 4 : 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane8; i += gridDim.x * blockDim.x) {
        const int k = (i % c8) * 8, pix = i / c8;
        float v[8];
        {
            const float4 x0 = __ldg(reinterpret_cast<const float4*>(xi + (long long)pix * c + k)), x1 = __ldg(reinterpret_cast<const float4*>(xi + (long long)pix * c + k + 4));
            const float4 m0 = __ldg(reinterpret_cast<const float4*>(mean + kb + k)), m1 = __ldg(reinterpret_cast<const float4*>(mean + kb + k + 4));
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(a + kb + k)), a1 = __ldg(reinterpret_cast<const float4*>(a + kb + k + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + kb + k)), b1 = __ldg(reinterpret_cast<const float4*>(b + kb + k + 4));
            v[0] = a0.x * (x0.x - m0.x) + b0.x; v[1] = a0.y * (x0.y - m0.y) + b0.y; v[2] = a0.z * (x0.z - m0.z) + b0.z; v[3] = a0.w * (x0.w - m0.w) + b0.w;
            v[4] = a1.x * (x1.x - m1.x) + b1.x; v[5] = a1.y * (x1.y - m1.y) + b1.y; v[6] = a1.z * (x1.z - m1.z) + b1.z; v[7] = a1.w * (x1.w - m1.w) + b1.w;
        }
        uint4 o;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(lrelu_f(v[0], slope), lrelu_f(v[1], slope)), p1 = __floats2bfloat162_rn(lrelu_f(v[2], slope), lrelu_f(v[3], slope));
        __nv_bfloat162 p2 = __floats2bfloat162_rn(lrelu_f(v[4], slope), lrelu_f(v[5], slope)), p3 = __floats2bfloat162_rn(lrelu_f(v[6], slope), lrelu_f(v[7], slope));
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        if (!upsample) {
            *reinterpret_cast<uint4*>(oi + (long long)pix * c + k) = o;
        } else {
            const int py = pix / w, px = pix - py * w;
            bf16* q = oi + ((long long)(2 * py) * (2 * w) + 2 * px) * c + k;
            *reinterpret_cast<uint4*>(q) = o;
            *reinterpret_cast<uint4*>(q + c) = o;
            *reinterpret_cast<uint4*>(q + (long long)2 * w * c) = o;
            *reinterpret_cast<uint4*>(q + (long long)2 * w * c + c) = o;
        }
    }
}

// activation mask of the backward: from the saved forward output y, or -- when y was never materialised (norm_act_operand) --
// recomputed from the normalisation coefficients: sign(a (x - mean) + b)
__device__ __forceinline__ bool act_positive(const float* y_or_null, long long i, float xv, float mu, const float* a, const float* b, long long k) {
    if (y_or_null != nullptr) return y_or_null[i] > 0.f;
    return a[k] * (xv - mu) + b[k] > 0.f;
}

template <typename T>
__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ y,
                                                              const float* __restrict__ mean, float* __restrict__ s1, float* __restrict__ s2,
                                                              int hw, int c, float slope, const float* __restrict__ ca, const float* __restrict__ cb) {
    __shared__ float sh1[8][33], sh2[8][33];
    int ch = blockIdx.x * 32 + threadIdx.x;
    long long img = blockIdx.y;
    long long base = img * (long long)hw * c;
    float a1 = 0.f, a2 = 0.f;
    if (ch < c) {
        float mu = mean[img * c + ch];
        for (int p = threadIdx.y; p < hw; p += 8) {
            long long i = base + (long long)p * c + ch;
            float g = to_f<T>(gy[i]);
            const float xv = to_f<T>(x[i]);
            if (y != nullptr) { if (!(to_f<T>(y[i]) > 0.f)) g *= slope; }
            else if (ca != nullptr) { if (!(ca[img * c + ch] * (xv - mu) + cb[img * c + ch] > 0.f)) g *= slope; }
            a1 += g;
            a2 += g * (xv - mu);
        }
    }
    sh1[threadIdx.y][threadIdx.x] = a1;
    sh2[threadIdx.y][threadIdx.x] = a2;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { t1 += sh1[j][threadIdx.x]; t2 += sh2[j][threadIdx.x]; }
        s1[img * c + ch] = t1;
        s2[img * c + ch] = t2;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const T* __restrict__ gy, const T* __restrict__ x, const T* __restrict__ y,
                                                             const float* __restrict__ mean, const float* __restrict__ A, const float* __restrict__ B,
                                                             const float* __restrict__ C, T* __restrict__ gx, long long total, int hw, int c, float slope,
                                                             const float* __restrict__ ca, const float* __restrict__ cb) {
    long long stride = (long long)gridDim.x * blockDim.x;
    long long plane = (long long)hw * c;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % c);
        long long k = (i / plane) * c + ch;
        float g = to_f<T>(gy[i]);
        const float xc = to_f<T>(x[i]) - mean[k];
        if (y != nullptr) { if (!(to_f<T>(y[i]) > 0.f)) g *= slope; }
        else if (ca != nullptr) { if (!(ca[k] * xc + cb[k] > 0.f)) g *= slope; }
        gx[i] = from_f<T>(A[k] * g + B[k] * xc + C[k]);
    }
}

__global__ void norm_coeffs_kernel(int mode, const float* __restrict__ mean, const float* __restrict__ m2, const float* __restrict__ p_scale,
                                   const float* __restrict__ p_shift, float* __restrict__ a, float* __restrict__ b, int n, int hw, int c, float eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * c) return;
    int ch = i % c;
    float sc, sh, r;
    if (mode == 0) {
        r = rsqrtf(m2[i] / (float)hw + eps);
        sc = p_scale[ch];
        sh = p_shift[ch];
    } else {
        r = 1.f / (sqrtf(m2[i] / (float)(hw - 1)) + eps);
        sc = p_scale[i];
        sh = p_shift[i];
    }
    a[i] = sc * r;
    b[i] = sh;
}

__global__ void norm_bwd_coeffs_kernel(int mode, const float* __restrict__ m2, const float* __restrict__ s1, const float* __restrict__ s2,
                                       const float* __restrict__ p_scale, float* __restrict__ A, float* __restrict__ B, float* __restrict__ C,
                                       float* __restrict__ g_scale, float* __restrict__ g_shift, int n, int hw, int c, float eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * c) return;
    int ch = i % c;
    if (mode == 0) {
        float r = rsqrtf(m2[i] / (float)hw + eps);
        float w = p_scale[ch];
        A[i] = w * r;
        B[i] = -w * r * r * r * s2[i] / (float)hw;
        C[i] = -w * r * s1[i] / (float)hw;
    } else {
        float sd = sqrtf(m2[i] / (float)(hw - 1));
        float r = 1.f / (sd + eps);
        float s = p_scale[i];
        A[i] = s * r;
        B[i] = sd > 0.f ? -r * r * s * s2[i] / ((float)(hw - 1) * sd) : 0.f;
        C[i] = -s * r * s1[i] / (float)hw;
        g_scale[i] = r * s2[i];
        g_shift[i] = s1[i];
    }
}

// InstanceNorm affine parameter gradients: reduce over the image axis (one thread per channel; n*c is tiny)
// grid (ceil(c/32)), block (32, 32): lanes = adjacent channels (coalesced rows of the [n][c] arrays), 32 image-lanes reduced in smem
// (the loop over images is a chain of dependent-latency loads: 1024 threads keep it at n/32 steps)
__global__ void __launch_bounds__(1024) in_param_grad_kernel(const float* __restrict__ m2, const float* __restrict__ s1, const float* __restrict__ s2,
                                                            float* __restrict__ g_scale, float* __restrict__ g_shift, int n, int hw, int c, float eps) {
    __shared__ float sh1[32][33], sh2[32][33];
    int ch = blockIdx.x * 32 + threadIdx.x;
    float gs = 0.f, gb = 0.f;
    if (ch < c)
        for (int img = threadIdx.y; img < n; img += 32) {
            int i = img * c + ch;
            gs += rsqrtf(m2[i] / (float)hw + eps) * s2[i];
            gb += s1[i];
        }
    sh1[threadIdx.y][threadIdx.x] = gs;
    sh2[threadIdx.y][threadIdx.x] = gb;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { a += sh1[j][threadIdx.x]; b += sh2[j][threadIdx.x]; }
        g_scale[ch] = a;
        g_shift[ch] = b;
    }
}

}  // namespace gim

using namespace gim;

extern "C" {

int gim_norm_stats(const void* x, float* mean, float* m2, int n, int hw, int c, int dtype, gim_stream_t s) {
    if (n <= 0 || c <= 0) return GIM_OK;
    GIM_REQUIRE(hw >= 1 && n <= 65535, "norm_stats: bad shape");
    dim3 grid((c + 31) / 32, n), block(32, 8);
    GIM_DISPATCH_DTYPE(dtype, (norm_stats_kernel<T><<<grid, block, 0, (cudaStream_t)s>>>((const T*)x, mean, m2, hw, c)));
    return check_launch("norm_stats");
}
int gim_affine_act_fwd(const void* x, const float* mean, const float* a, const float* b, void* y, int n, int hw, int c, float slope, int dtype,
                       gim_stream_t s) {
    long long total = (long long)n * hw * c;
    if (total <= 0) return GIM_OK;
    const long long plane4 = (long long)hw * c / 4;
    if (dtype == GIM_F32 && c % 4 == 0 && n <= 65535 && plane4 < (1LL << 30) && ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)mean | (uintptr_t)a | (uintptr_t)b) & 15) == 0)) {
        int gx = (int)((plane4 + 1023) / 1024);                      // four float4 per thread
        const int cap = (num_sms() * 16 + n - 1) / n;
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        affine_act_vec4_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)s>>>((const float4*)x, (const float4*)mean, (const float4*)a, (const float4*)b, (float4*)y,
                                                                        (int)plane4, c / 4, slope);
        return check_launch("affine_act_fwd");
    }
    GIM_DISPATCH_DTYPE(dtype, (affine_act_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, mean, a, b, (T*)y, total, hw, c, slope)));
    return check_launch("affine_act_fwd");
}
int gim_norm_act_operand(const float* x, const float* mean, const float* a, const float* b, void* out_bf16, int n, int h, int wd, int c, float slope,
                         int upsample, gim_stream_t s) {
    if (n <= 0 || h <= 0 || wd <= 0 || c <= 0) return GIM_OK;
    GIM_REQUIRE(c % 8 == 0 && n <= 65535 && (long long)h * wd * c < (1LL << 31), "norm_act_operand: needs c % 8 == 0");
    GIM_REQUIRE(((((uintptr_t)x | (uintptr_t)out_bf16 | (uintptr_t)mean | (uintptr_t)a | (uintptr_t)b) & 15) == 0), "norm_act_operand: 16-byte alignment");
    const int plane8 = h * wd * (c / 8);
    int gx = (plane8 + 511) / 512;
    const int cap = (num_sms() * 16 + n - 1) / n;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    norm_act_operand_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)s>>>(x, mean, a, b, (bf16*)out_bf16, h, wd, c, slope, upsample);
    return check_launch("norm_act_operand");
}
int gim_norm_bwd_reduce(const void* gy, const void* x, const void* y, const float* mean, const float* act_a, const float* act_b, float* s1, float* s2,
                        int n, int hw, int c, float slope, int dtype, gim_stream_t s) {
    if (n <= 0 || c <= 0) return GIM_OK;
    GIM_REQUIRE(hw >= 1 && n <= 65535, "norm_bwd_reduce: bad shape");
    dim3 grid((c + 31) / 32, n), block(32, 8);
    GIM_DISPATCH_DTYPE(dtype, (norm_bwd_reduce_kernel<T><<<grid, block, 0, (cudaStream_t)s>>>((const T*)gy, (const T*)x, (const T*)y, mean, s1, s2, hw, c, slope, act_a, act_b)));
    return check_launch("norm_bwd_reduce");
}
int gim_norm_bwd_apply(const void* gy, const void* x, const void* y, const float* mean, const float* act_a, const float* act_b, const float* A, const float* B,
                       const float* C, void* gx, int n, int hw, int c, float slope, int dtype, gim_stream_t s) {
    long long total = (long long)n * hw * c;
    if (total <= 0) return GIM_OK;
    const long long plane4 = (long long)hw * c / 4;
    if (dtype == GIM_F32 && c % 4 == 0 && n <= 65535 && plane4 < (1LL << 30) &&
        ((((uintptr_t)gy | (uintptr_t)x | (uintptr_t)y | (uintptr_t)gx | (uintptr_t)mean | (uintptr_t)A | (uintptr_t)B | (uintptr_t)C | (uintptr_t)act_a | (uintptr_t)act_b) & 15) == 0)) {
        int g = (int)((plane4 + 1023) / 1024);
        const int cap = (num_sms() * 16 + n - 1) / n;
        if (g > cap) g = cap;
        if (g < 1) g = 1;
        norm_bwd_apply_vec4_kernel<<<dim3(g, n), 256, 0, (cudaStream_t)s>>>((const float4*)gy, (const float4*)x, (const float4*)y, (const float4*)mean, (const float4*)A,
                                                                          (const float4*)B, (const float4*)C, (float4*)gx, (int)plane4, c / 4, slope,
                                                                          (const float4*)act_a, (const float4*)act_b);
        return check_launch("norm_bwd_apply");
    }
    GIM_DISPATCH_DTYPE(dtype, (norm_bwd_apply_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>((const T*)gy, (const T*)x, (const T*)y, mean, A, B, C,
                                                                                                       (T*)gx, total, hw, c, slope, act_a, act_b)));
    return check_launch("norm_bwd_apply");
}
int gim_norm_coeffs(int mode, const float* mean, const float* m2, const float* p_scale, const float* p_shift, float* a, float* b, int n, int hw, int c,
                    float eps, gim_stream_t s) {
    if (n * c <= 0) return GIM_OK;
    GIM_REQUIRE(mode == 0 || hw > 1, "ada_in needs more than one pixel (unbiased std)");
    norm_coeffs_kernel<<<(n * c + 255) / 256, 256, 0, (cudaStream_t)s>>>(mode, mean, m2, p_scale, p_shift, a, b, n, hw, c, eps);
    return check_launch("norm_coeffs");
}
int gim_norm_bwd_coeffs(int mode, const float* m2, const float* s1, const float* s2, const float* p_scale, float* A, float* B, float* C,
                        float* g_scale, float* g_shift, int n, int hw, int c, float eps, gim_stream_t s) {
    if (n * c <= 0) return GIM_OK;
    norm_bwd_coeffs_kernel<<<(n * c + 255) / 256, 256, 0, (cudaStream_t)s>>>(mode, m2, s1, s2, p_scale, A, B, C, g_scale, g_shift, n, hw, c, eps);
    int rc = check_launch("norm_bwd_coeffs");
    if (rc != GIM_OK || mode != 0) return rc;
    in_param_grad_kernel<<<(c + 31) / 32, dim3(32, 32), 0, (cudaStream_t)s>>>(m2, s1, s2, g_scale, g_shift, n, hw, c, eps);
    return check_launch("in_param_grad");
}

}  // extern "C"
