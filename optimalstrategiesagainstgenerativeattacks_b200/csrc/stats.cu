// Permutation-invariant set statistics over the sample axis (the n leaked / m generated / k registration samples), the encoder's
// global max, the per-episode losses.  All fp32, HBM/latency bound: one thread per (episode, feature), coalesced over features.
#include "common.cuh"

namespace gim {

__global__ void __launch_bounds__(256) set_stats_fwd_kernel(const float* __restrict__ x, float* __restrict__ out_sum, float* __restrict__ out_std, int ld,
                                                            int b, int s, int d, float scale, float eps) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)b * d) return;
    int j = (int)(i % d);
    long long e = i / d;
    const float* xe = x + e * (long long)s * d + j;
    float sum = 0.f;
    for (int k = 0; k < s; ++k) sum += xe[(long long)k * d];
    if (out_sum) out_sum[e * ld + j] = scale * sum;
    if (out_std) {
        float sd = 0.f;
        if (s > 1) {
            float mu = sum / (float)s, q = 0.f;
            for (int k = 0; k < s; ++k) { float t = xe[(long long)k * d] - mu; q += t * t; }
            sd = sqrtf(q / (float)(s - 1) + eps);
        }
        out_std[e * ld + j] = sd;
    }
}

__global__ void __launch_bounds__(256) set_stats_bwd_kernel(const float* __restrict__ g_sum, const float* __restrict__ g_std, int ld,
                                                            const float* __restrict__ x, float* __restrict__ gx, int b, int s, int d, float scale, float eps) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)b * d) return;
    int j = (int)(i % d);
    long long e = i / d;
    const float* xe = x + e * (long long)s * d + j;
    float* ge = gx + e * (long long)s * d + j;
    float gs = g_sum ? scale * g_sum[e * ld + j] : 0.f;
    float coef = 0.f, mu = 0.f;
    if (g_std && s > 1) {
        float sum = 0.f, q = 0.f;
        for (int k = 0; k < s; ++k) sum += xe[(long long)k * d];
        mu = sum / (float)s;
        for (int k = 0; k < s; ++k) { float t = xe[(long long)k * d] - mu; q += t * t; }
        float sd = sqrtf(q / (float)(s - 1) + eps);
        coef = g_std[e * ld + j] / ((float)(s - 1) * sd);
    }
    if (coef != 0.f) {
        for (int k = 0; k < s; ++k) ge[(long long)k * d] = gs + coef * (xe[(long long)k * d] - mu);
    } else {
        for (int k = 0; k < s; ++k) ge[(long long)k * d] = gs;      // x is not read (set-broadcast passes no x)
    }
}

__global__ void __launch_bounds__(256) set_std_bwd_bwd_kernel(const float* __restrict__ ggx, const float* __restrict__ g_std, int ld,
                                                              const float* __restrict__ x, float* __restrict__ gg_std, float* __restrict__ g_x, int b, int s,
                                                              int d, float eps) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)b * d) return;
    int j = (int)(i % d);
    long long e = i / d;
    const float* xe = x + e * (long long)s * d + j;
    const float* qe = ggx + e * (long long)s * d + j;
    float* oe = g_x + e * (long long)s * d + j;
    if (s <= 1) {
        gg_std[e * d + j] = 0.f;
        for (int k = 0; k < s; ++k) oe[(long long)k * d] = 0.f;
        return;
    }
    float sum = 0.f, qs = 0.f, q = 0.f, dot = 0.f;
    for (int k = 0; k < s; ++k) { sum += xe[(long long)k * d]; qs += qe[(long long)k * d]; }
    float mu = sum / (float)s, mq = qs / (float)s;
    for (int k = 0; k < s; ++k) {
        float t = xe[(long long)k * d] - mu;
        q += t * t;
        dot += qe[(long long)k * d] * t;
    }
    float sd = sqrtf(q / (float)(s - 1) + eps);
    float sm1 = (float)(s - 1);
    gg_std[e * d + j] = dot / (sm1 * sd);
    float gsd = g_std[e * ld + j] / sm1;
    float c2 = dot / (sm1 * sd * sd * sd);
    for (int k = 0; k < s; ++k) oe[(long long)k * d] = gsd * ((qe[(long long)k * d] - mq) / sd - c2 * (xe[(long long)k * d] - mu));
}

__global__ void __launch_bounds__(256) set_center_add_kernel(const float* __restrict__ x, const float* __restrict__ add, float* __restrict__ y, int b, int s,
                                                             int d, int center) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)b * d) return;
    int j = (int)(i % d);
    long long e = i / d;
    const float* xe = x + e * (long long)s * d + j;
    float* ye = y + e * (long long)s * d + j;
    float shift = add ? add[e * d + j] : 0.f;
    if (center) {
        float sum = 0.f;
        for (int k = 0; k < s; ++k) sum += xe[(long long)k * d];
        shift -= sum / (float)s;
    }
    for (int k = 0; k < s; ++k) ye[(long long)k * d] = xe[(long long)k * d] + shift;
}

// y[e, k, :] = alpha * x[e, k, :] + beta * add[e, :]: the Gaussian episode synthesis mu + sigma * noise (reference
// training/gim_gaussian_training.py:71-86) from standard-normal draws, one read + one write per element
__global__ void __launch_bounds__(256) affine_rows_kernel(const float* x, const float* __restrict__ add, float* y /* may alias x */, long long total,
                                                          int s, int d, float alpha, float beta) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int j = (int)(i % d);
        const long long e = i / ((long long)s * d);
        y[i] = fmaf(alpha, x[i], beta * add[e * d + j]);
    }
}

// ImgAttention blend (reference models/model_blocks.py:596-608), one thread per pixel, c = image channels (1 or 3):
//   s1 = <q1, k1>, s2 = <q2, k2> over channels; (a1, a2) = softmax(s1, s2); out = a1 * x1 + a2 * v2
__global__ void __launch_bounds__(256) img_att_blend_fwd_kernel(const float* __restrict__ q1, const float* __restrict__ k1, const float* __restrict__ q2,
                                                                const float* __restrict__ k2, const float* __restrict__ x1, const float* __restrict__ v2,
                                                                float* __restrict__ out, float* __restrict__ att, long long pixels, int c) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pixels) return;
    const long long o = p * c;
    float s1 = 0.f, s2 = 0.f;
    for (int j = 0; j < c; ++j) { s1 = fmaf(q1[o + j], k1[o + j], s1); s2 = fmaf(q2[o + j], k2[o + j], s2); }
    const float m = fmaxf(s1, s2), e1 = expf(s1 - m), e2 = expf(s2 - m), inv = 1.f / (e1 + e2);
    const float a1 = e1 * inv, a2 = e2 * inv;
    att[2 * p] = a1;
    att[2 * p + 1] = a2;
    for (int j = 0; j < c; ++j) out[o + j] = a1 * x1[o + j] + a2 * v2[o + j];
}
// d a1 = <g, x1>, d a2 = <g, v2>;  d s1 = a1 (d a1 - (a1 d a1 + a2 d a2)), d s2 likewise;  d q1 = d s1 k1, ...;  d x1 = a1 g, d v2 = a2 g
__global__ void __launch_bounds__(256) img_att_blend_bwd_kernel(const float* __restrict__ g, const float* __restrict__ q1, const float* __restrict__ k1,
                                                                const float* __restrict__ q2, const float* __restrict__ k2, const float* __restrict__ x1,
                                                                const float* __restrict__ v2, const float* __restrict__ att, float* __restrict__ gq1,
                                                                float* __restrict__ gk1, float* __restrict__ gq2, float* __restrict__ gk2,
                                                                float* __restrict__ gx1, float* __restrict__ gv2, long long pixels, int c) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pixels) return;
    const long long o = p * c;
    const float a1 = att[2 * p], a2 = att[2 * p + 1];
    float d1 = 0.f, d2 = 0.f;
    for (int j = 0; j < c; ++j) { d1 = fmaf(g[o + j], x1[o + j], d1); d2 = fmaf(g[o + j], v2[o + j], d2); }
    const float dot = a1 * d1 + a2 * d2, ds1 = a1 * (d1 - dot), ds2 = a2 * (d2 - dot);
    for (int j = 0; j < c; ++j) {
        gq1[o + j] = ds1 * k1[o + j];
        gk1[o + j] = ds1 * q1[o + j];
        gq2[o + j] = ds2 * k2[o + j];
        gk2[o + j] = ds2 * q2[o + j];
        if (gx1) gx1[o + j] = a1 * g[o + j];
        gv2[o + j] = a2 * g[o + j];
    }
}

// grid (ceil(c/32), n), block (32, 8): max over the hw pixels + first arg-max
template <typename T>
__global__ void __launch_bounds__(256) gmax_fwd_kernel(const T* __restrict__ x, float* __restrict__ y, int32_t* __restrict__ idx, int hw, int c, float slope) {
    __shared__ float shv[8][33];
    __shared__ int shi[8][33];
    int ch = blockIdx.x * 32 + threadIdx.x;
    long long img = blockIdx.y;
    const T* xi = x + img * (long long)hw * c;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    if (ch < c)
        for (int p = threadIdx.y; p < hw; p += 8) {
            float v = to_f<T>(xi[(long long)p * c + ch]);
            if (v > best || (v == best && p < bi)) { best = v; bi = p; }
        }
    shv[threadIdx.y][threadIdx.x] = best;
    shi[threadIdx.y][threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            float v = shv[j][threadIdx.x];
            int p = shi[j][threadIdx.x];
            if (v > best || (v == best && p < bi)) { best = v; bi = p; }
        }
        y[img * c + ch] = lrelu_f(best, slope);          // the encoder's output LeakyReLU (gim_img_models.py:55-56) rides along; slope 1 = none
        idx[img * c + ch] = bi;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) gather_idx_kernel(const T* __restrict__ x, const int32_t* __restrict__ idx, float* __restrict__ y, int n, int hw, int c) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * c) return;
    int ch = (int)(i % c);
    long long img = i / c;
    y[i] = to_f<T>(x[(img * hw + idx[i]) * (long long)c + ch]);
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_idx_kernel(const float* __restrict__ g, const int32_t* __restrict__ idx, T* __restrict__ gx, long long total,
                                                          int hw, int c) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % c);
        long long r = i / c;
        int p = (int)(r % hw);
        long long k = (r / hw) * c + ch;
        gx[i] = from_f<T>(idx[k] == p ? g[k] : 0.f);
    }
}

__global__ void bce_fwd_kernel(const float* __restrict__ x, float t, float* __restrict__ loss, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i];
    loss[i] = fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v)));
}
__global__ void bce_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, float t, float* __restrict__ gx, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = x[i];
    float sg = 1.f / (1.f + expf(-v));
    gx[i] = g[i] * (sg - t);
}

template <typename T>
__global__ void __launch_bounds__(256) rows_sqsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long l) {
    __shared__ float sh[33];
    const T* xr = x + (long long)blockIdx.y * l;
    float acc = 0.f;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < l; j += (long long)gridDim.x * blockDim.x) {
        float v = to_f<T>(xr[j]);
        acc += v * v;
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(&out[blockIdx.y], acc);
}
template <typename T>
__global__ void __launch_bounds__(256) rows_scale_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ y, long long l, float alpha) {
    float sc = alpha * s[blockIdx.y];
    const T* xr = x + (long long)blockIdx.y * l;
    T* yr = y + (long long)blockIdx.y * l;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < l; j += (long long)gridDim.x * blockDim.x) yr[j] = from_f<T>(sc * to_f<T>(xr[j]));
}

}  // namespace gim

using namespace gim;

extern "C" {

int gim_set_stats_fwd(const float* x, float* out_sum, float* out_std, int ld_out, int b, int s, int d, float scale, float eps, gim_stream_t st) {
    long long total = (long long)b * d;
    if (total <= 0) return GIM_OK;
    GIM_REQUIRE(s >= 1, "set_stats: empty sample axis");
    set_stats_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)st>>>(x, out_sum, out_std, ld_out, b, s, d, scale, eps);
    return check_launch("set_stats_fwd");
}
int gim_set_stats_bwd(const float* g_sum, const float* g_std, int ld_g, const float* x, float* gx, int b, int s, int d, float scale, float eps,
                      gim_stream_t st) {
    long long total = (long long)b * d;
    if (total <= 0) return GIM_OK;
    set_stats_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)st>>>(g_sum, g_std, ld_g, x, gx, b, s, d, scale, eps);
    return check_launch("set_stats_bwd");
}
int gim_set_std_bwd_bwd(const float* ggx, const float* g_std, int ld_g, const float* x, float* gg_std, float* g_x, int b, int s, int d, float eps,
                        gim_stream_t st) {
    long long total = (long long)b * d;
    if (total <= 0) return GIM_OK;
    set_std_bwd_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)st>>>(ggx, g_std, ld_g, x, gg_std, g_x, b, s, d, eps);
    return check_launch("set_std_bwd_bwd");
}
int gim_set_center_add(const float* x, const float* add, float* y, int b, int s, int d, int center, gim_stream_t st) {
    long long total = (long long)b * d;
    if (total <= 0) return GIM_OK;
    set_center_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)st>>>(x, add, y, b, s, d, center);
    return check_launch("set_center_add");
}
int gim_affine_rows(const float* x, const float* add, float* y, int b, int s, int d, float alpha, float beta, gim_stream_t st) {
    long long total = (long long)b * s * d;
    if (total <= 0) return GIM_OK;
    affine_rows_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)st>>>(x, add, y, total, s, d, alpha, beta);
    return check_launch("affine_rows");
}
int gim_img_att_blend_fwd(const float* q1, const float* k1, const float* q2, const float* k2, const float* x1, const float* v2, float* out, float* att,
                          long long pixels, int c, gim_stream_t st) {
    if (pixels <= 0 || c <= 0) return GIM_OK;
    img_att_blend_fwd_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, (cudaStream_t)st>>>(q1, k1, q2, k2, x1, v2, out, att, pixels, c);
    return check_launch("img_att_blend_fwd");
}
int gim_img_att_blend_bwd(const float* g, const float* q1, const float* k1, const float* q2, const float* k2, const float* x1, const float* v2,
                          const float* att, float* gq1, float* gk1, float* gq2, float* gk2, float* gx1, float* gv2, long long pixels, int c, gim_stream_t st) {
    if (pixels <= 0 || c <= 0) return GIM_OK;
    img_att_blend_bwd_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, (cudaStream_t)st>>>(g, q1, k1, q2, k2, x1, v2, att, gq1, gk1, gq2, gk2, gx1, gv2, pixels, c);
    return check_launch("img_att_blend_bwd");
}
int gim_gmax_fwd(const void* x, float* y, int32_t* idx, int n, int hw, int c, float slope, int dtype, gim_stream_t st) {
    if (n <= 0 || c <= 0) return GIM_OK;
    GIM_REQUIRE(hw >= 1 && n <= 65535, "gmax: bad shape");
    dim3 grid((c + 31) / 32, n), block(32, 8);
    GIM_DISPATCH_DTYPE(dtype, (gmax_fwd_kernel<T><<<grid, block, 0, (cudaStream_t)st>>>((const T*)x, y, idx, hw, c, slope)));
    return check_launch("gmax_fwd");
}
int gim_gather_idx(const void* x, const int32_t* idx, float* y, int n, int hw, int c, int dtype, gim_stream_t st) {
    long long total = (long long)n * c;
    if (total <= 0) return GIM_OK;
    GIM_DISPATCH_DTYPE(dtype, (gather_idx_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)st>>>((const T*)x, idx, y, n, hw, c)));
    return check_launch("gather_idx");
}
int gim_scatter_idx(const float* g, const int32_t* idx, void* gx, int n, int hw, int c, int dtype, gim_stream_t st) {
    long long total = (long long)n * hw * c;
    if (total <= 0) return GIM_OK;
    GIM_DISPATCH_DTYPE(dtype, (scatter_idx_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)st>>>(g, idx, (T*)gx, total, hw, c)));
    return check_launch("scatter_idx");
}
int gim_bce_logits_fwd(const float* x, float target, float* loss, long long n, gim_stream_t st) {
    if (n <= 0) return GIM_OK;
    bce_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)st>>>(x, target, loss, n);
    return check_launch("bce_fwd");
}
int gim_bce_logits_bwd(const float* g, const float* x, float target, float* gx, long long n, gim_stream_t st) {
    if (n <= 0) return GIM_OK;
    bce_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)st>>>(g, x, target, gx, n);
    return check_launch("bce_bwd");
}
int gim_rows_sqsum(const void* x, float* out, int b, long long l, int dtype, gim_stream_t st) {
    if (b <= 0) return GIM_OK;
    GIM_REQUIRE(b <= 65535, "rows_sqsum: too many rows");
    if (cudaMemsetAsync(out, 0, sizeof(float) * (size_t)b, (cudaStream_t)st) != cudaSuccess) return fail(GIM_E_CUDA, "rows_sqsum memset");
    if (l <= 0) return GIM_OK;
    int gx = (int)((l + 256 * 16 - 1) / (256 * 16));
    if (gx > 64) gx = 64;
    if (deterministic()) gx = 1;
    dim3 grid(gx, b);
    GIM_DISPATCH_DTYPE(dtype, (rows_sqsum_kernel<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)x, out, l)));
    return check_launch("rows_sqsum");
}
int gim_rows_scale(const void* x, const float* s, void* y, int b, long long l, float alpha, int dtype, gim_stream_t st) {
    if (b <= 0 || l <= 0) return GIM_OK;
    GIM_REQUIRE(b <= 65535, "rows_scale: too many rows");
    int gx = (int)((l + 256 * 4 - 1) / (256 * 4));
    if (gx > 256) gx = 256;
    dim3 grid(gx, b);
    GIM_DISPATCH_DTYPE(dtype, (rows_scale_kernel<T><<<grid, 256, 0, (cudaStream_t)st>>>((const T*)x, s, (T*)y, l, alpha)));
    return check_launch("rows_scale");
}

}  // extern "C"
