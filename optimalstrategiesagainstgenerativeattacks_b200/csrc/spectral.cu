// Spectral normalisation of the conv weights: one power iteration per train-mode forward call, exactly the state machine of
// torch.nn.utils.spectral_norm (old-style hook) used at model_blocks.py:492-495, 522-526, 750-751, 792-793, 836-840.
// Batch independent and tiny (<= 5.3 M floats per weight): plain coalesced GEMV kernels, fp32 throughout.
#include "common.cuh"

namespace gim {

// t[j] = sum_co W[co][j] * u[co]
// grid (ceil(J/32), row_splits), block (32, 8): each CTA reduces a slab of rows for 32 columns; slabs combine with atomicAdd (t zeroed first)
__global__ void __launch_bounds__(256) sn_wtu_kernel(const float* __restrict__ W, const float* __restrict__ u, float* __restrict__ t, int cout, int J,
                                                     int rows_per_split) {
    __shared__ float sh[8][33];
    int j = blockIdx.x * 32 + threadIdx.x;
    int r0 = blockIdx.y * rows_per_split, r1 = min(cout, r0 + rows_per_split);
    float acc = 0.f;
    if (j < J)
        for (int co = r0 + threadIdx.y; co < r1; co += 8) acc = fmaf(W[(long long)co * J + j], u[co], acc);
    sh[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && j < J) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) a += sh[q][threadIdx.x];
        atomicAdd(&t[j], a);
    }
}

// single CTA: v = t / max(||t||, eps) (power_iter) ; v_used = v
__global__ void __launch_bounds__(1024) sn_vnorm_kernel(const float* __restrict__ t, float* __restrict__ v, float* __restrict__ v_used, int J, float eps,
                                                        int power_iter) {
    __shared__ float sh[33];
    if (power_iter) {
        float acc = 0.f;
        for (int j = threadIdx.x; j < J; j += blockDim.x) acc += t[j] * t[j];
        float nrm = sqrtf(block_sum(acc, sh));
        float inv = 1.f / fmaxf(nrm, eps);
        for (int j = threadIdx.x; j < J; j += blockDim.x) {
            float val = t[j] * inv;
            v[j] = val;
            v_used[j] = val;
        }
    } else {
        for (int j = threadIdx.x; j < J; j += blockDim.x) v_used[j] = v[j];
    }
}

// s[co] = sum_j W[co][j] * v[j]; one CTA per row
__global__ void __launch_bounds__(128) sn_wv_kernel(const float* __restrict__ W, const float* __restrict__ v, float* __restrict__ s, int J) {
    __shared__ float sh[33];
    const float* row = W + (long long)blockIdx.x * J;
    float acc = 0.f;
    for (int j = threadIdx.x; j < J; j += blockDim.x) acc = fmaf(row[j], v[j], acc);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) s[blockIdx.x] = acc;
}

// single CTA: u = s / max(||s||, eps) (power_iter); sigma = u . s ; u_used = u
__global__ void __launch_bounds__(512) sn_unorm_kernel(const float* __restrict__ s, float* __restrict__ u, float* __restrict__ u_used,
                                                       float* __restrict__ sigma, int cout, float eps, int power_iter) {
    __shared__ float sh[33];
    float inv = 0.f;
    if (power_iter) {
        float acc = 0.f;
        for (int i = threadIdx.x; i < cout; i += blockDim.x) acc += s[i] * s[i];
        inv = 1.f / fmaxf(sqrtf(block_sum(acc, sh)), eps);
    }
    float dot = 0.f;
    for (int i = threadIdx.x; i < cout; i += blockDim.x) {
        float ui = power_iter ? s[i] * inv : u[i];
        if (power_iter) u[i] = ui;
        u_used[i] = ui;
        dot += ui * s[i];
    }
    dot = block_sum(dot, sh);
    if (threadIdx.x == 0) *sigma = dot;
}

// w_sn[t][co][ci] = W[co][ci][t] / sigma
__global__ void __launch_bounds__(256) sn_pack_kernel(const float* __restrict__ W, const float* __restrict__ sigma, float* __restrict__ w_sn, int cout,
                                                      int cin, int taps) {
    float inv = 1.f / *sigma;
    long long total = (long long)taps * cout * cin;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ci = (int)(i % cin);
        long long r = i / cin;
        int co = (int)(r % cout);
        int t = (int)(r / cout);
        w_sn[i] = W[((long long)co * cin + ci) * taps + t] * inv;
    }
}

// c += sum G*W with G unpacked from [t][co][ci]
__global__ void __launch_bounds__(256) sn_bwd_dot_kernel(const float* __restrict__ g, const float* __restrict__ W, float* __restrict__ c, int cout, int cin,
                                                         int taps) {
    __shared__ float sh[33];
    long long total = (long long)taps * cout * cin;
    long long stride = (long long)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ci = (int)(i % cin);
        long long r = i / cin;
        int co = (int)(r % cout);
        int t = (int)(r / cout);
        acc = fmaf(g[i], W[((long long)co * cin + ci) * taps + t], acc);
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(c, acc);
}

// gW[co][ci][t] = G/sigma - (c/sigma^2) u[co] v[ci*taps+t]
__global__ void __launch_bounds__(256) sn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ u, const float* __restrict__ v,
                                                           const float* __restrict__ sigma, const float* __restrict__ c, float* __restrict__ gW, int cout,
                                                           int cin, int taps) {
    float inv = 1.f / *sigma;
    float k = (*c) * inv * inv;
    long long total = (long long)taps * cout * cin;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
        int t = (int)(o % taps);
        long long r = o / taps;
        int ci = (int)(r % cin);
        int co = (int)(r / cin);
        gW[o] = g[((long long)t * cout + co) * cin + ci] * inv - k * u[co] * v[ci * taps + t];
    }
}

}  // namespace gim

using namespace gim;

extern "C" {

int gim_sn_forward(const float* weight_orig, float* u, float* v, int power_iter, float eps, float* w_sn, float* sigma, float* u_used, float* v_used,
                   float* scratch, int cout, int cin, int ksize, gim_stream_t s) {
    GIM_REQUIRE(cout > 0 && cin > 0 && ksize > 0, "sn_forward: bad shape");
    cudaStream_t st = (cudaStream_t)s;
    int taps = ksize * ksize, J = cin * taps;
    float* t = scratch;
    float* sv = scratch + J;
    int rc;
    if (power_iter) {
        if (cudaMemsetAsync(t, 0, sizeof(float) * (size_t)J, st) != cudaSuccess) return fail(GIM_E_CUDA, "sn_forward memset");
        int col_blocks = (J + 31) / 32;
        int splits = (2 * num_sms() + col_blocks - 1) / col_blocks;          // enough CTAs to cover the chip ~2x
        if (splits > (cout + 31) / 32) splits = (cout + 31) / 32;
        if (splits < 1) splits = 1;
        int rows_per_split = (cout + splits - 1) / splits;
        sn_wtu_kernel<<<dim3(col_blocks, (cout + rows_per_split - 1) / rows_per_split), dim3(32, 8), 0, st>>>(weight_orig, u, t, cout, J, rows_per_split);
        if ((rc = check_launch("sn_wtu")) != GIM_OK) return rc;
    }
    sn_vnorm_kernel<<<1, 1024, 0, st>>>(t, v, v_used, J, eps, power_iter);
    if ((rc = check_launch("sn_vnorm")) != GIM_OK) return rc;
    sn_wv_kernel<<<cout, 128, 0, st>>>(weight_orig, v_used, sv, J);
    if ((rc = check_launch("sn_wv")) != GIM_OK) return rc;
    sn_unorm_kernel<<<1, 512, 0, st>>>(sv, u, u_used, sigma, cout, eps, power_iter);
    if ((rc = check_launch("sn_unorm")) != GIM_OK) return rc;
    long long total = (long long)taps * cout * cin;
    sn_pack_kernel<<<ew_grid(total, 256), 256, 0, st>>>(weight_orig, sigma, w_sn, cout, cin, taps);
    return check_launch("sn_pack");
}

int gim_sn_backward(const float* g_w_sn, const float* weight_orig, const float* u_used, const float* v_used, const float* sigma, float* g_weight_orig,
                    float* scratch, int cout, int cin, int ksize, gim_stream_t s) {
    GIM_REQUIRE(cout > 0 && cin > 0 && ksize > 0, "sn_backward: bad shape");
    cudaStream_t st = (cudaStream_t)s;
    int taps = ksize * ksize;
    long long total = (long long)taps * cout * cin;
    if (cudaMemsetAsync(scratch, 0, sizeof(float), st) != cudaSuccess) return fail(GIM_E_CUDA, "sn_backward memset");
    int grid = ew_grid(total, 256, 8);
    sn_bwd_dot_kernel<<<grid, 256, 0, st>>>(g_w_sn, weight_orig, scratch, cout, cin, taps);
    int rc = check_launch("sn_bwd_dot");
    if (rc != GIM_OK) return rc;
    sn_bwd_apply_kernel<<<ew_grid(total, 256), 256, 0, st>>>(g_w_sn, u_used, v_used, sigma, scratch, g_weight_orig, cout, cin, taps);
    return check_launch("sn_bwd_apply");
}

}  // extern "C"
