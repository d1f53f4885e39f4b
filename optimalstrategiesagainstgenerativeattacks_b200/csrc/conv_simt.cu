// CUDA-core (FFMA, fp32 accumulate) implicit-GEMM convolution: the fp32 parity path (rel 1e-4 gate) and the path for the
// skinny layers (Cin <= 6 or Cout <= 3) that are HBM-bound and do not map onto 128xN tensor-core tiles.
// The tensor-core path (conv_tc.cu, tcgen05 + TMEM + TMA) takes the dense bf16 layers.
#include "common.cuh"

namespace gim {

constexpr int BK = 16;

// y[pix][co] = bias[co] + sum_kf A(pix,kf) * w[tap(kf)][co][ci(kf)],  kf = tap*cin + ci flattened so tiny-Cin layers waste nothing.
template <typename T, typename TO, int BM, int BN>
__global__ void __launch_bounds__(256) conv_fwd_simt_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                                                            TO* __restrict__ y, int n, int h, int wd, int cin, int cout, int ks) {
    static_assert((BM / 4) * (BN / 4) == 256, "tile must map onto 256 threads of 4x4 outputs");
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    __shared__ int pix_h[BM], pix_w[BM];
    __shared__ long long pix_base[BM];

    const int tid = threadIdx.x;
    const int pad = (ks - 1) / 2;
    const long long npix = (long long)n * h * wd;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int ktot = ks * ks * cin;

    for (int m = tid; m < BM; m += 256) {
        long long g = m0 + m;
        if (g < npix) {
            int ww = (int)(g % wd);
            long long t = g / wd;
            int hh = (int)(t % h);
            long long img = t / h;
            pix_h[m] = hh;
            pix_w[m] = ww;
            pix_base[m] = img * (long long)h * wd;
        } else {
            pix_h[m] = -100000;
            pix_w[m] = 0;
            pix_base[m] = 0;
        }
    }
    __syncthreads();

    const int ty = tid / (BN / 4), tx = tid % (BN / 4);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < ktot; k0 += BK) {
        for (int e = tid; e < BM * BK; e += 256) {
            int m = e / BK, kk = e % BK;
            int kf = k0 + kk;
            float v = 0.f;
            if (kf < ktot) {
                int tap = kf / cin, ci = kf - tap * cin;
                int r = tap / ks, s = tap - r * ks;
                int hh = pix_h[m] + r - pad, ww = pix_w[m] + s - pad;
                if (hh >= 0 && hh < h && ww >= 0 && ww < wd) v = to_f<T>(x[(pix_base[m] + (long long)hh * wd + ww) * cin + ci]);
            }
            As[kk][m] = v;
        }
        for (int e = tid; e < BN * BK; e += 256) {
            int c = e / BK, kk = e % BK;
            int kf = k0 + kk, co = n0 + c;
            float v = 0.f;
            if (kf < ktot && co < cout) {
                int tap = kf / cin, ci = kf - tap * cin;
                v = to_f<T>(w[((long long)tap * cout + co) * cin + ci]);
            }
            Bs[kk][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long g = m0 + ty * 4 + i;
        if (g >= npix) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co < cout) y[g * cout + co] = from_f<TO>(acc[i][j] + (bias ? bias[co] : 0.f));
        }
    }
}

// gw[tap][co][ci] += sum_{pix in split} gy[pix][co] * x[pix + tap][ci];  N index nf = tap*cin + ci flattened.
template <typename T, int BM, int BN>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ gy, float* __restrict__ gw, int n, int h,
                                                              int wd, int cin, int cout, int ks, long long pix_per_split) {
    static_assert((BM / 4) * (BN / 4) == 256, "tile must map onto 256 threads of 4x4 outputs");
    __shared__ float As[BK][BM + 4];   // gy^T : [pixel][co]
    __shared__ float Bs[BK][BN + 4];   // shifted x : [pixel][nf]
    __shared__ int nf_dr[BN], nf_ds[BN], nf_ci[BN];

    const int tid = threadIdx.x;
    const int pad = (ks - 1) / 2;
    const long long npix = (long long)n * h * wd;
    const int co0 = blockIdx.x * BM;
    const int nf0 = blockIdx.y * BN;
    const int ntot = ks * ks * cin;
    long long p_begin = (long long)blockIdx.z * pix_per_split;
    long long p_end = p_begin + pix_per_split;
    if (p_end > npix) p_end = npix;

    for (int j = tid; j < BN; j += 256) {
        int nf = nf0 + j;
        if (nf < ntot) {
            int tap = nf / cin;
            nf_ci[j] = nf - tap * cin;
            nf_dr[j] = tap / ks - pad;
            nf_ds[j] = tap % ks - pad;
        } else {
            nf_ci[j] = -1;
            nf_dr[j] = 0;
            nf_ds[j] = 0;
        }
    }
    __syncthreads();

    const int ty = tid / (BN / 4), tx = tid % (BN / 4);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
        for (int e = tid; e < BM * BK; e += 256) {
            int kk = e / BM, m = e % BM;          // consecutive threads -> consecutive channels of one pixel (coalesced)
            long long g = p0 + kk;
            int co = co0 + m;
            As[kk][m] = (g < p_end && co < cout) ? to_f<T>(gy[g * cout + co]) : 0.f;
        }
        for (int e = tid; e < BN * BK; e += 256) {
            int kk = e / BN, j = e % BN;
            long long g = p0 + kk;
            float v = 0.f;
            if (g < p_end && nf_ci[j] >= 0) {
                int ww = (int)(g % wd);
                long long t = g / wd;
                int hh = (int)(t % h);
                long long img = t / h;
                hh += nf_dr[j];
                ww += nf_ds[j];
                if (hh >= 0 && hh < h && ww >= 0 && ww < wd) v = to_f<T>(x[((img * h + hh) * (long long)wd + ww) * cin + nf_ci[j]]);
            }
            Bs[kk][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int co = co0 + ty * 4 + i;
        if (co >= cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int nf = nf0 + tx * 4 + j;
            if (nf < ntot) {
                int tap = nf / cin, ci = nf - tap * cin;
                atomicAdd(&gw[((long long)tap * cout + co) * cin + ci], acc[i][j]);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) weight_cast_kernel(const float* __restrict__ w, T* __restrict__ out, long long total) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) out[i] = from_f<T>(w[i]);
}

// out[T-1-t][ci][co] = w[t][co][ci]
template <typename T>
__global__ void __launch_bounds__(256) weight_flip_kernel(const float* __restrict__ w, T* __restrict__ out, int taps, int cout, int cin) {
    long long total = (long long)taps * cout * cin;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int co = (int)(i % cout);
        long long r = i / cout;
        int ci = (int)(r % cin);
        int tf = (int)(r / cin);
        out[i] = from_f<T>(w[((long long)(taps - 1 - tf) * cout + co) * cin + ci]);
    }
}

// out[c] += sum over this CTA's rows; grid.x row chunks, grid.y channel chunks of 32; block (32, 8)
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long rows, int c, long long rows_per_cta) {
    __shared__ float sh[8][33];
    int ch = blockIdx.y * 32 + threadIdx.x;
    long long r0 = (long long)blockIdx.x * rows_per_cta;
    long long r1 = r0 + rows_per_cta;
    if (r1 > rows) r1 = rows;
    float s = 0.f;
    if (ch < c) for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += to_f<T>(x[r * c + ch]);
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += sh[j][threadIdx.x];
        atomicAdd(&out[ch], t);
    }
}

// bf16 operand copy and column sums of an fp32 [rows][c] gradient in ONE pass (Conv2d backward needs both: the operand of the
// input-/weight-gradient convolutions and the bias gradient).  Thread = 8 adjacent channels (two 16-byte loads, one 16-byte store) of
// a fixed channel group; rows advance by blockDim/(c/8) per step; per-CTA partial sums meet in shared memory, one atomic per channel
// and CTA.  c/8 must divide 256.
__global__ void __launch_bounds__(256) cast_colsum_kernel(const float* __restrict__ x, bf16* __restrict__ y, float* __restrict__ sums, long long rows, int c) {
    __shared__ float red[2048];                       // [row lane][c] for c <= 2048 / (256 / (c/8)) ... sized for the worst case c = 2048
    const int groups = c >> 3, cg = threadIdx.x % groups, rl = threadIdx.x / groups, rlanes = blockDim.x / groups;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (long long r = (long long)blockIdx.x * rlanes + rl; r < rows; r += (long long)gridDim.x * rlanes) {
        const float4* src = reinterpret_cast<const float4*>(x + r * c + cg * 8);
        const float4 a = __ldg(src), b = __ldg(src + 1);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w), p2 = __floats2bfloat162_rn(b.x, b.y),
                       p3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        *reinterpret_cast<uint4*>(y + r * c + cg * 8) = o;
    }
    // threads with the same channel group: thread (rl, cg) holds red[rl * c + cg * 8 + j]; rlanes * c == 2048
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rl * c + cg * 8 + j] = acc[j];
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        float t = 0.f;
        for (int q = 0; q < rlanes; ++q) t += red[q * c + ch];
        atomicAdd(&sums[ch], t);
    }
}

template <typename T, typename TO>
static int conv_fwd_simt(const void* x, const void* w, const float* bias, void* y, int n, int h, int wd, int cin, int cout, int ks, cudaStream_t st) {
    long long npix = (long long)n * h * wd;
    if (cout <= 16) {
        dim3 grid((unsigned)((npix + 255) / 256), (cout + 15) / 16);
        conv_fwd_simt_kernel<T, TO, 256, 16><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, bias, (TO*)y, n, h, wd, cin, cout, ks);
    } else {
        dim3 grid((unsigned)((npix + 63) / 64), (cout + 63) / 64);
        conv_fwd_simt_kernel<T, TO, 64, 64><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, bias, (TO*)y, n, h, wd, cin, cout, ks);
    }
    return check_launch("conv_fwd_simt");
}

template <typename T>
static int conv_wgrad_simt(const void* x, const void* gy, float* gw, int n, int h, int wd, int cin, int cout, int ks, cudaStream_t st) {
    long long npix = (long long)n * h * wd;
    int ntot = ks * ks * cin;
    bool skinny = ntot <= 16;
    int bm = skinny ? 256 : 64, bn = skinny ? 16 : 64;
    int tiles = ((cout + bm - 1) / bm) * ((ntot + bn - 1) / bn);
    long long want = ((long long)num_sms() * 4 + tiles - 1) / tiles;
    long long max_split = (npix + 4 * BK - 1) / (4 * BK);
    if (want > max_split) want = max_split;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    if (deterministic()) want = 1;
    long long per = (npix + want - 1) / want;
    per = (per + BK - 1) / BK * BK;
    int splits = (int)((npix + per - 1) / per);
    dim3 grid((cout + bm - 1) / bm, (ntot + bn - 1) / bn, splits);
    if (skinny) conv_wgrad_simt_kernel<T, 256, 16><<<grid, 256, 0, st>>>((const T*)x, (const T*)gy, gw, n, h, wd, cin, cout, ks, per);
    else conv_wgrad_simt_kernel<T, 64, 64><<<grid, 256, 0, st>>>((const T*)x, (const T*)gy, gw, n, h, wd, cin, cout, ks, per);
    return check_launch("conv_wgrad_simt");
}

// implemented in conv_tc.cu
int conv_fwd_tc_ex(const void* x, const void* w, const float* bias, void* y, int n, int h, int wd, int cin, int cout, int ks, int out_f32,
                   int epi, float slope, const void* mask_ref, const float* addend, cudaStream_t st);
int conv_wgrad_tc(const void* x, const void* gy, float* gw, int n, int h, int wd, int cin, int cout, int ks, cudaStream_t st);
bool conv_tc_supported(int n, int h, int wd, int cin, int cout, int ks, int dtype);
bool wgrad_tc_supported(int n, int h, int wd, int cin, int cout, int ks, int dtype);
int conv_fwd_tc_plan(int n, int h, int wd, int cin, int cout, int ks, int out_f32, int epi, int* out20);
int conv_wgrad_tc_plan(int n, int h, int wd, int cin, int cout, int ks, int* out16);

}  // namespace gim

using namespace gim;

extern "C" {

int gim_conv2d_tc_supported(int n, int h, int w, int cin, int cout, int ksize, int dtype) {
    return conv_tc_supported(n, h, w, cin, cout, ksize, dtype) ? 1 : 0;
}

int gim_conv2d_wgrad_tc_supported(int n, int h, int w, int cin, int cout, int ksize, int dtype) {
    return wgrad_tc_supported(n, h, w, cin, cout, ksize, dtype) ? 1 : 0;
}

int gim_conv2d_fwd_plan(int n, int h, int w, int cin, int cout, int ksize, int out_dtype, int epilogue, int* plan20) {
    return conv_fwd_tc_plan(n, h, w, cin, cout, ksize, out_dtype == GIM_F32 ? 1 : 0, epilogue, plan20);
}

int gim_conv2d_wgrad_plan(int n, int h, int w, int cin, int cout, int ksize, int* plan16) { return conv_wgrad_tc_plan(n, h, w, cin, cout, ksize, plan16); }

int gim_conv2d_fwd(const void* x, const void* w, const float* bias, void* y, int n, int h, int wd, int cin, int cout, int ksize, int dtype,
                   int out_dtype, int algo, gim_stream_t s) {
    GIM_REQUIRE(n > 0 && h > 0 && wd > 0 && cin > 0 && cout > 0, "conv2d_fwd: empty shape");
    GIM_REQUIRE(ksize >= 1 && (ksize & 1), "conv2d_fwd: kernel size must be odd ('same' padding)");
    GIM_REQUIRE((long long)n * h * wd / 64 + 1 < 2147483647LL, "conv2d_fwd: too many pixels");
    cudaStream_t st = (cudaStream_t)s;
    bool tc_ok = conv_tc_supported(n, h, wd, cin, cout, ksize, dtype);
    if (algo == GIM_ALGO_TCGEN05 && !tc_ok) return fail(GIM_E_UNSUPPORTED, "conv2d_fwd: shape/dtype not supported by the tcgen05 path");
    GIM_REQUIRE(out_dtype == GIM_F32 || out_dtype == dtype, "conv2d_fwd: output must be fp32 or the operand dtype");
    if ((algo == GIM_ALGO_TCGEN05) || (algo == GIM_ALGO_AUTO && tc_ok))
        return conv_fwd_tc_ex(x, w, bias, y, n, h, wd, cin, cout, ksize, out_dtype == GIM_F32 ? 1 : 0, 0, 0.f, nullptr, nullptr, st);
    if (dtype == GIM_F32) return conv_fwd_simt<float, float>(x, w, bias, y, n, h, wd, cin, cout, ksize, st);
    if (dtype == GIM_BF16 && out_dtype == GIM_F32) return conv_fwd_simt<bf16, float>(x, w, bias, y, n, h, wd, cin, cout, ksize, st);
    if (dtype == GIM_BF16) return conv_fwd_simt<bf16, bf16>(x, w, bias, y, n, h, wd, cin, cout, ksize, st);
    return fail(GIM_E_ARG, "conv2d_fwd: bad dtype");
}

int gim_conv2d_fwd_fused(const void* x, const void* w, const float* bias, void* y, const void* mask_ref, const float* addend, int n, int h, int wd,
                         int cin, int cout, int ksize, int out_dtype, int epilogue, float slope, gim_stream_t s) {
    GIM_REQUIRE(n > 0 && h > 0 && wd > 0 && cin > 0 && cout > 0, "conv2d_fwd_fused: empty shape");
    GIM_REQUIRE(ksize >= 1 && (ksize & 1), "conv2d_fwd_fused: kernel size must be odd ('same' padding)");
    GIM_REQUIRE(out_dtype == GIM_F32 || out_dtype == GIM_BF16, "conv2d_fwd_fused: bad output dtype");
    GIM_REQUIRE((epilogue & ~63) == 0, "conv2d_fwd_fused: unknown epilogue bits");
    if (!conv_tc_supported(n, h, wd, cin, cout, ksize, GIM_BF16)) return fail(GIM_E_UNSUPPORTED, "conv2d_fwd_fused: shape not supported by the tcgen05 path");
    return conv_fwd_tc_ex(x, w, bias, y, n, h, wd, cin, cout, ksize, out_dtype == GIM_F32 ? 1 : 0, epilogue, slope, mask_ref, addend, (cudaStream_t)s);
}

int gim_conv2d_wgrad(const void* x, const void* gy, float* gw, int n, int h, int wd, int cin, int cout, int ksize, int dtype, int algo, gim_stream_t s) {
    GIM_REQUIRE(n > 0 && h > 0 && wd > 0 && cin > 0 && cout > 0, "conv2d_wgrad: empty shape");
    GIM_REQUIRE(ksize >= 1 && (ksize & 1), "conv2d_wgrad: kernel size must be odd");
    cudaStream_t st = (cudaStream_t)s;
    bool tc_ok = wgrad_tc_supported(n, h, wd, cin, cout, ksize, dtype);
    if (algo == GIM_ALGO_TCGEN05 && !tc_ok) return fail(GIM_E_UNSUPPORTED, "conv2d_wgrad: shape/dtype not supported by the tcgen05 path");
    if (cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)ksize * ksize * cout * cin, st) != cudaSuccess) return fail(GIM_E_CUDA, "wgrad memset");
    if ((algo == GIM_ALGO_TCGEN05) || (algo == GIM_ALGO_AUTO && tc_ok)) return conv_wgrad_tc(x, gy, gw, n, h, wd, cin, cout, ksize, st);
    GIM_DISPATCH_DTYPE(dtype, return conv_wgrad_simt<T>(x, gy, gw, n, h, wd, cin, cout, ksize, st));
}

int gim_weight_cast(const float* w, void* out, int taps, int cout, int cin, int dtype, gim_stream_t s) {
    long long total = (long long)taps * cout * cin;
    if (total <= 0) return GIM_OK;
    GIM_DISPATCH_DTYPE(dtype, (weight_cast_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>(w, (T*)out, total)));
    return check_launch("weight_cast");
}
int gim_weight_flip(const float* w, void* out, int taps, int cout, int cin, int dtype, gim_stream_t s) {
    long long total = (long long)taps * cout * cin;
    if (total <= 0) return GIM_OK;
    GIM_DISPATCH_DTYPE(dtype, (weight_flip_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>(w, (T*)out, taps, cout, cin)));
    return check_launch("weight_flip");
}
static int colsum_launch(const void* x, float* out, long long rows, int c, int dtype, bool accumulate, gim_stream_t s);

int gim_colsum(const void* x, float* out, long long rows, int c, int dtype, gim_stream_t s) { return colsum_launch(x, out, rows, c, dtype, false, s); }
int gim_colsum_acc(const void* x, float* out, long long rows, int c, int dtype, gim_stream_t s) { return colsum_launch(x, out, rows, c, dtype, true, s); }

int gim_cast_colsum(const float* x, void* y_bf16, float* sums, long long rows, int c, int accumulate, gim_stream_t s) {
    if (c <= 0) return GIM_OK;
    GIM_REQUIRE(c % 8 == 0 && c <= 2048 && 256 % (c / 8) == 0, "cast_colsum: channels must be 8 * 2^k <= 2048");
    GIM_REQUIRE((((uintptr_t)x | (uintptr_t)y_bf16) & 15) == 0, "cast_colsum: 16-byte alignment");
    if (!accumulate && cudaMemsetAsync(sums, 0, sizeof(float) * (size_t)c, (cudaStream_t)s) != cudaSuccess) return fail(GIM_E_CUDA, "cast_colsum memset");
    if (rows <= 0) return GIM_OK;
    const int rlanes = 256 / (c / 8);
    long long want = (rows + rlanes * 4 - 1) / (rlanes * 4);                  // >= 4 row steps per CTA
    if (want > 2LL * num_sms()) want = 2LL * num_sms();
    if (want < 1 || deterministic()) want = 1;
    cast_colsum_kernel<<<(unsigned)want, 256, 0, (cudaStream_t)s>>>(x, (bf16*)y_bf16, sums, rows, c);
    return check_launch("cast_colsum");
}

static int colsum_launch(const void* x, float* out, long long rows, int c, int dtype, bool accumulate, gim_stream_t s) {
    if (c <= 0) return GIM_OK;
    if (!accumulate && cudaMemsetAsync(out, 0, sizeof(float) * (size_t)c, (cudaStream_t)s) != cudaSuccess) return fail(GIM_E_CUDA, "colsum memset");
    if (rows <= 0) return GIM_OK;
    int cchunks = (c + 31) / 32;
    long long want = ((long long)num_sms() * 4 + cchunks - 1) / cchunks;
    long long maxc = (rows + 63) / 64;
    if (want > maxc) want = maxc;
    if (want < 1 || deterministic()) want = 1;
    long long per = (rows + want - 1) / want;
    dim3 grid((unsigned)((rows + per - 1) / per), cchunks), block(32, 8);
    GIM_DISPATCH_DTYPE(dtype, (colsum_kernel<T><<<grid, block, 0, (cudaStream_t)s>>>((const T*)x, out, rows, c, per)));
    return check_launch("colsum");
}

}  // extern "C"
