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


// ----------------------------------------------------------------------------------------------------------------
// The same strided batched GEMM on the tensor cores for the bf16 path (attention bmm pair and its gradients: thousands of small
// matrices, N_tok <= 256).  The attention logits and their softmax gradients are cancellation-prone, so the operands are NOT simply
// rounded: each fp32 value is split hi + lo into two bf16 numbers while it is staged in shared memory (k-contiguous rows, so every
// transpose combination feeds ldmatrix the same way) and the product is a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with fp32 accumulation
// (mma.sync.m16n8k16) -- ~2^-16 relative, i.e. fp32-grade results at tensor-core speed.  These GEMMs are HBM-bound at a few GFLOP
// each, so the legacy warp-level MMA (x3) is ample; tcgen05 stays reserved for the convolutions.
// ----------------------------------------------------------------------------------------------------------------
constexpr int TBM = 64, TBN = 64, TBK = 32, TLD = TBK + 8;       // 80-byte rows: conflict-free ldmatrix

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const bf16* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}


__device__ __forceinline__ void split_store(bf16* hi, bf16* lo, float v) {
    const bf16 h = __float2bfloat16_rn(v);
    *hi = h;
    *lo = __float2bfloat16_rn(v - __bfloat162float(h));
}

// One ROWS x TBK operand tile -> smem [row][k] as hi + lo bf16.  fp32 operands whose unit-stride axis is 16-byte aligned are read
// with float4 loads (along k, or along the row axis for transposed operands); anything else element by element.
template <int ROWS>
__device__ __forceinline__ void stage_split_tile(bf16 (*hi)[TLD], bf16 (*lo)[TLD], const void* __restrict__ P, int dt, long long off, long long s_row,
                                                 long long s_k, int row0, int rows_total, int k0, int K, int tid) {
    const bool k_fast = (s_k == 1);
    const bool vec = dt == GIM_F32 && (((uintptr_t)P) & 15) == 0 && (off & 3) == 0 && (k_fast ? (s_row & 3) == 0 : (s_row == 1 && (s_k & 3) == 0));
    if (vec) {
        const float* p = (const float*)P + off;
        if (k_fast) {
            for (int e = tid; e < ROWS * (TBK / 4); e += 128) {
                const int r = e / (TBK / 4), kq = (e % (TBK / 4)) * 4;
                const int gr = row0 + r, gk = k0 + kq;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (gr < rows_total) {
                    if (gk + 3 < K) {
                        const float4 t = *reinterpret_cast<const float4*>(p + gr * s_row + gk);
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) if (gk + i < K) v[i] = p[gr * s_row + gk + i];
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) split_store(&hi[r][kq + i], &lo[r][kq + i], v[i]);
            }
        } else {
            for (int e = tid; e < (ROWS / 4) * TBK; e += 128) {
                const int kk = e / (ROWS / 4), rq = (e % (ROWS / 4)) * 4;
                const int gr = row0 + rq, gk = k0 + kk;
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                if (gk < K) {
                    if (gr + 3 < rows_total) {
                        const float4 t = *reinterpret_cast<const float4*>(p + gk * s_k + gr);
                        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) if (gr + i < rows_total) v[i] = p[gk * s_k + gr + i];
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) split_store(&hi[rq + i][kk], &lo[rq + i][kk], v[i]);
            }
        }
        return;
    }
    for (int e = tid; e < ROWS * TBK; e += 128) {
        int r, kk;
        if (k_fast) { r = e / TBK; kk = e % TBK; } else { kk = e / ROWS; r = e % ROWS; }
        const int gr = row0 + r, gk = k0 + kk;
        split_store(&hi[r][kk], &lo[r][kk], (gr < rows_total && gk < K) ? ld_dt(P, off + gr * s_row + gk * s_k, dt) : 0.f);
    }
}

__global__ void __launch_bounds__(128) gemm_strided_mma_kernel(const void* __restrict__ A, int dtA, long long sAb, long long sAm, long long sAk,
                                                               const void* __restrict__ B, int dtB, long long sBb, long long sBk, long long sBn,
                                                               void* __restrict__ C, int dtC, long long sCb, long long ldc, int M, int N, int K,
                                                               float alpha, float beta, int m_tiles, int k_per_split) {
    __shared__ __align__(16) bf16 As[TBM][TLD], Al[TBM][TLD];     // hi / lo parts
    __shared__ __align__(16) bf16 Bs[TBN][TLD], Bl[TBN][TLD];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // blockIdx.x = m tile + m_tiles * k split: weight-gradient shapes (tiny M x N, K = tens of thousands of rows) are split along K
    // over the chip and combined with fp32 atomics into a zeroed C
    const int split = blockIdx.x / m_tiles;
    const int m0 = (blockIdx.x - split * m_tiles) * TBM, n0 = blockIdx.y * TBN;
    const long long b = blockIdx.z;
    const long long a_off = b * sAb, b_off = b * sBb, c_off = b * sCb;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int k_begin = split * k_per_split, k_end = min(K, k_begin + k_per_split);
    const bool atomic_out = k_per_split < K;
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += TBK) {
        // stage the two tiles as [row][k] hi/lo bf16, walking the unit-stride axis of each operand with consecutive threads
        stage_split_tile<TBM>(As, Al, A, dtA, a_off, sAm, sAk, m0, M, k0, k_end, tid);
        stage_split_tile<TBN>(Bs, Bl, B, dtB, b_off, sBn, sBk, n0, N, k0, k_end, tid);
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < TBK; ks += 16) {
            uint32_t af[2][4], al[2][4], bfr[2][4], bl[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {     // A 16x16: matrices (rows 0-7,k 0-7) (rows 8-15,k 0-7) (rows 0-7,k 8-15) (rows 8-15,k 8-15)
                ldmatrix_x4(af[i], &As[wm + i * 16 + (lane & 15)][ks + (lane >> 4) * 8]);
                ldmatrix_x4(al[i], &Al[wm + i * 16 + (lane & 15)][ks + (lane >> 4) * 8]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {     // B, two n8 tiles per x4: (n 0-7,k 0-7) (n 0-7,k 8-15) (n 8-15,k 0-7) (n 8-15,k 8-15)
                ldmatrix_x4(bfr[j], &Bs[wn + j * 16 + (lane & 7) + ((lane >> 4) << 3)][ks + ((lane >> 3) & 1) * 8]);
                ldmatrix_x4(bl[j], &Bl[wn + j * 16 + (lane & 7) + ((lane >> 4) << 3)][ks + ((lane >> 3) & 1) * 8]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int jj = j >> 1, q = (j & 1) * 2;
                    mma_bf16_16816(acc[i][j], al[i], bfr[jj][q], bfr[jj][q + 1]);      // small terms first
                    mma_bf16_16816(acc[i][j], af[i], bl[jj][q], bl[jj][q + 1]);
                    mma_bf16_16816(acc[i][j], af[i], bfr[jj][q], bfr[jj][q + 1]);
                }
        }
        __syncthreads();
    }
    const bool c_vec = !atomic_out && dtC == GIM_F32 && beta == 0.f && (((uintptr_t)C) & 7) == 0 && ((c_off | ldc) & 1) == 0;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int gm = m0 + wm + i * 16 + (lane >> 2) + hrow * 8;
                const int gn = n0 + wn + j * 8 + (lane & 3) * 2;
                if (gm >= M) continue;
                const long long idx = c_off + gm * ldc + gn;
                if (c_vec && gn + 1 < N) {
                    *reinterpret_cast<float2*>((float*)C + idx) = make_float2(alpha * acc[i][j][hrow * 2], alpha * acc[i][j][hrow * 2 + 1]);
                } else {
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (gn + q < N) {
                            float v = alpha * acc[i][j][hrow * 2 + q];
                            if (atomic_out) { atomicAdd((float*)C + idx + q, v); continue; }
                            if (beta != 0.f) v += beta * ld_dt(C, idx + q, dtC);
                            st_dt(C, idx + q, dtC, v);
                        }
                }
            }
}

__global__ void __launch_bounds__(256) zero_strided_kernel(float* __restrict__ C, long long sCb, long long ldc, int M, int N, long long total) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int n = (int)(i % N);
        const long long r = i / N;
        C[(r / M) * sCb + (r % M) * ldc + n] = 0.f;
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
int gim_gemm_strided_bf16(const void* A, int dtA, long long sAb, long long sAm, long long sAk, const void* B, int dtB, long long sBb, long long sBk,
                          long long sBn, void* C, int dtC, long long sCb, long long ldc, int m, int n, int k, int batch, float alpha, float beta,
                          gim_stream_t s) {
    if (m <= 0 || n <= 0 || batch <= 0) return GIM_OK;
    GIM_REQUIRE(k >= 0 && batch <= 65535 && (n + TBN - 1) / TBN <= 65535, "gemm_bf16: bad shape");
    GIM_REQUIRE((dtA == GIM_F32 || dtA == GIM_BF16) && (dtB == GIM_F32 || dtB == GIM_BF16) && (dtC == GIM_F32 || dtC == GIM_BF16), "gemm_bf16: bad dtype");
    const int m_tiles = (m + TBM - 1) / TBM, n_tiles = (n + TBN - 1) / TBN;
    const long long ctas = (long long)m_tiles * n_tiles * batch;
    int splits = 1;
    if (dtC == GIM_F32 && beta == 0.f && k >= 8 * TBK && ctas < num_sms() && !deterministic()) {       // few output tiles, long K: split K over the chip
        long long want = (2LL * num_sms() + ctas - 1) / ctas, most = k / (4 * TBK);
        splits = (int)(want < most ? want : most);
        if (splits < 1) splits = 1;
    }
    int k_per_split = ((k + splits - 1) / splits + TBK - 1) / TBK * TBK;
    if (k_per_split < TBK) k_per_split = TBK;
    splits = k > 0 ? (k + k_per_split - 1) / k_per_split : 1;
    if (splits > 1) {
        const long long total = (long long)batch * m * n;
        zero_strided_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>((float*)C, sCb, ldc, m, n, total);
        int rc = check_launch("gemm_zero");
        if (rc != GIM_OK) return rc;
    } else {
        k_per_split = k > 0 ? k : 1;          // single pass over K (also covers k == 0: C = beta * C)
    }
    dim3 grid(m_tiles * splits, n_tiles, batch);
    gemm_strided_mma_kernel<<<grid, 128, 0, (cudaStream_t)s>>>(A, dtA, sAb, sAm, sAk, B, dtB, sBb, sBk, sBn, C, dtC, sCb, ldc, m, n, k, alpha, beta, m_tiles,
                                                               k_per_split);
    return check_launch("gemm_strided_mma");
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
