// Small dense algebra of the heads: nn.Linear, the SelfAttention bmm pair and its softmax.  These are <1 % of the FLOPs
//  and latency-bound; a strided, mixed-dtype FFMA GEMM with fp32 accumulation covers every transpose
// combination the forward, backward and double-backward need.
#include "common.cuh"

namespace gim {

constexpr int GK = 16;

__global__ void __launch_bounds__(256) gemm_strided_kernel(const void* __restrict__ A, int dtA, long long sAb, long long sAm, long long sAk,
                                                           const void* __restrict__ B, int dtB, long long sBb, long long sBk, long long sBn,
                                                           void* __restrict__ C, int dtC, long long sCb, long long ldc, int M, int N, int K,
                                                           float alpha, float beta) {
    __shared__ float As[GK][64 + 4];
    __shared__ float Bs[GK][64 + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const long long b = blockIdx.z;
    const long long a_off = b * sAb, b_off = b * sBb, c_off = b * sCb;
    // pick the thread->element order that walks the unit-stride axis of each operand
    const bool a_k_fast = (sAk == 1), b_n_fast = (sBn == 1);
    const int ty = tid / 16, tx = tid % 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += GK) {
        for (int e = tid; e < 64 * GK; e += 256) {
            int m, kk;
            if (a_k_fast) { m = e / GK; kk = e % GK; } else { kk = e / 64; m = e % 64; }
            int gm = m0 + m, gk = k0 + kk;
            As[kk][m] = (gm < M && gk < K) ? ld_dt(A, a_off + gm * sAm + gk * sAk, dtA) : 0.f;
        }
        for (int e = tid; e < 64 * GK; e += 256) {
            int nn, kk;
            if (b_n_fast) { kk = e / 64; nn = e % 64; } else { nn = e / GK; kk = e % GK; }
            int gn = n0 + nn, gk = k0 + kk;
            Bs[kk][nn] = (gn < N && gk < K) ? ld_dt(B, b_off + gk * sBk + gn * sBn, dtB) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            long long idx = c_off + gm * ldc + gn;
            float v = alpha * acc[i][j];
            if (beta != 0.f) v += beta * ld_dt(C, idx, dtC);
            st_dt(C, idx, dtC, v);
        }
    }
}

__global__ void __launch_bounds__(256) bias_act_kernel(const float* __restrict__ x, const float* __restrict__ bias, float* __restrict__ y, long long total,
                                                       int c, float slope) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        float v = x[i] + (bias ? bias[i % c] : 0.f);
        y[i] = lrelu_f(v, slope);
    }
}

// one warp per row (cols <= a few hundred here); rows are contiguous
__global__ void __launch_bounds__(256) softmax_rows_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int cols) {
    long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + row * cols;
    float* yr = y + row * cols;
    float mx = -INFINITY;
    for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, xr[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < cols; j += 32) sum += expf(xr[j] - mx);
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    for (int j = lane; j < cols; j += 32) yr[j] = expf(xr[j] - mx) * inv;
}

// gx = y * (gy - sum(gy*y))
__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, float* __restrict__ gx,
                                                               long long rows, int cols) {
    long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* g = gy + row * cols;
    const float* p = y + row * cols;
    float d = 0.f;
    for (int j = lane; j < cols; j += 32) d += g[j] * p[j];
    d = warp_sum(d);
    for (int j = lane; j < cols; j += 32) gx[row * cols + j] = p[j] * (g[j] - d);
}

// d/dy of [y*(gy - sum gy*y)] contracted with ggx:  g_y = ggx*gy - ggx*sum(gy*y) - gy*sum(ggx*y)
__global__ void __launch_bounds__(256) softmax_rows_bwd_bwd_kernel(const float* __restrict__ ggx, const float* __restrict__ gy, const float* __restrict__ y,
                                                                   float* __restrict__ g_y, long long rows, int cols) {
    long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* q = ggx + row * cols;
    const float* g = gy + row * cols;
    const float* p = y + row * cols;
    float d1 = 0.f, d2 = 0.f;
    for (int j = lane; j < cols; j += 32) { d1 += g[j] * p[j]; d2 += q[j] * p[j]; }
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    for (int j = lane; j < cols; j += 32) g_y[row * cols + j] = q[j] * g[j] - q[j] * d1 - g[j] * d2;
}

}  // namespace gim

using namespace gim;

extern "C" {

int gim_gemm_strided(const void* A, int dtA, long long sAb, long long sAm, long long sAk, const void* B, int dtB, long long sBb, long long sBk,
                     long long sBn, void* C, int dtC, long long sCb, long long ldc, int m, int n, int k, int batch, float alpha, float beta,
                     gim_stream_t s) {
    if (m <= 0 || n <= 0 || batch <= 0) return GIM_OK;
    GIM_REQUIRE(k >= 0 && batch <= 65535 && (n + 63) / 64 <= 65535, "gemm: bad shape");
    GIM_REQUIRE((dtA == GIM_F32 || dtA == GIM_BF16) && (dtB == GIM_F32 || dtB == GIM_BF16) && (dtC == GIM_F32 || dtC == GIM_BF16), "gemm: bad dtype");
    dim3 grid((m + 63) / 64, (n + 63) / 64, batch);
    gemm_strided_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(A, dtA, sAb, sAm, sAk, B, dtB, sBb, sBk, sBn, C, dtC, sCb, ldc, m, n, k, alpha, beta);
    return check_launch("gemm_strided");
}
int gim_bias_act_fwd(const float* x, const float* bias, float* y, long long rows, int c, float slope, gim_stream_t s) {
    long long total = rows * c;
    if (total <= 0) return GIM_OK;
    bias_act_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>(x, bias, y, total, c, slope);
    return check_launch("bias_act");
}
int gim_softmax_rows_fwd(const float* x, float* y, long long rows, int cols, gim_stream_t s) {
    if (rows <= 0 || cols <= 0) return GIM_OK;
    softmax_rows_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)s>>>(x, y, rows, cols);
    return check_launch("softmax_rows_fwd");
}
int gim_softmax_rows_bwd(const float* gy, const float* y, float* gx, long long rows, int cols, gim_stream_t s) {
    if (rows <= 0 || cols <= 0) return GIM_OK;
    softmax_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)s>>>(gy, y, gx, rows, cols);
    return check_launch("softmax_rows_bwd");
}
int gim_softmax_rows_bwd_bwd(const float* ggx, const float* gy, const float* y, float* g_y, long long rows, int cols, gim_stream_t s) {
    if (rows <= 0 || cols <= 0) return GIM_OK;
    softmax_rows_bwd_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)s>>>(ggx, gy, y, g_y, rows, cols);
    return check_launch("softmax_rows_bwd_bwd");
}

}  // extern "C"
