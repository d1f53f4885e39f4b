// Pointwise, resampling and layout kernels (HBM-bound: 16-byte vector accesses, grid-stride, grids sized from the SM count).
#include "common.cuh"

namespace gim {

thread_local char g_err[512] = "";
long long g_launches = 0;
int g_deterministic = 0;

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ void ld16(const T* p, float (&f)[16 / sizeof(T)]) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < 16 / (int)sizeof(T); ++i) f[i] = to_f<T>(e[i]);
}
template <typename T>
__device__ __forceinline__ void st16(T* p, const float (&f)[16 / sizeof(T)]) {
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int i = 0; i < 16 / (int)sizeof(T); ++i) e[i] = from_f<T>(f[i]);
    *reinterpret_cast<uint4*>(p) = raw;
}

// generic 1/2/3-input elementwise map with 16B vectors when all pointers are 16B aligned
template <typename T, int NIN, typename F>
__global__ void __launch_bounds__(256) map_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c,
                                                  T* __restrict__ out, long long n, bool vec, F f) {
    constexpr int V = Vec16<T>::N;
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    long long nv = vec ? n / V : 0;
    for (long long i = tid; i < nv; i += stride) {
        float fa[V], fb[V], fc[V], fo[V];
        ld16<T>(a + i * V, fa);
        if (NIN > 1) ld16<T>(b + i * V, fb);
        if (NIN > 2) ld16<T>(c + i * V, fc);
#pragma unroll
        for (int j = 0; j < V; ++j) fo[j] = f(fa[j], NIN > 1 ? fb[j] : 0.f, NIN > 2 ? fc[j] : 0.f);
        st16<T>(out + i * V, fo);
    }
    for (long long i = nv * V + tid; i < n; i += stride) {
        float x = to_f<T>(a[i]);
        float y = NIN > 1 ? to_f<T>(b[i]) : 0.f;
        float z = NIN > 2 ? to_f<T>(c[i]) : 0.f;
        out[i] = from_f<T>(f(x, y, z));
    }
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <typename T, int NIN, typename F>
static int launch_map(const void* a, const void* b, const void* c, void* out, long long n, cudaStream_t st, F f, const char* name) {
    if (n <= 0) return GIM_OK;
    bool vec = aligned16(a) && aligned16(out) && (NIN < 2 || aligned16(b)) && (NIN < 3 || aligned16(c));
    int grid = ew_grid(n, 256, 2 * Vec16<T>::N);
    map_kernel<T, NIN, F><<<grid, 256, 0, st>>>((const T*)a, (const T*)b, (const T*)c, (T*)out, n, vec, f);
    return check_launch(name);
}

struct LreluF { float s; __device__ float operator()(float x, float, float) const { return x > 0.f ? x : x * s; } };
struct LreluB { float s; __device__ float operator()(float g, float x, float) const { return x > 0.f ? g : g * s; } };
struct TanhF { __device__ float operator()(float x, float, float) const { return tanhf(x); } };
struct TanhB { __device__ float operator()(float g, float y, float) const { return g * (1.f - y * y); } };
struct Axpby { float a, b; __device__ float operator()(float x, float y, float) const { return a * x + b * y; } };
struct Ax { float a; __device__ float operator()(float x, float, float) const { return a * x; } };

template <typename T>
__global__ void __launch_bounds__(256) scale_dev_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ out, long long n) {
    float sc = *s;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = from_f<T>(sc * to_f<T>(x[i]));
}

template <typename T>
__global__ void __launch_bounds__(256) dot_kernel(const T* __restrict__ x, const T* __restrict__ y, float* __restrict__ out, long long n) {
    __shared__ float sh[33];
    float acc = 0.f;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += to_f<T>(x[i]) * to_f<T>(y[i]);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

// y[n,ho,wo,c] = scale * sum_{dy,dx in 2x2} (a + b)[n,2ho+dy,2wo+dx,c]
// vector variants: one thread handles 16 bytes of adjacent channels (c % V == 0, 16-byte aligned bases)
template <typename T>
__global__ void __launch_bounds__(256) pool2_vec_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y,
                                                        int n, int h, int w, int c, float scale) {
    constexpr int V = Vec16<T>::N;
    int ho = h / 2, wo = w / 2, cv = c / V;
    long long total = (long long)n * ho * wo * cv;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % cv) * V;
        long long p = i / cv;
        int x = (int)(p % wo); p /= wo;
        int yy = (int)(p % ho);
        long long img = p / ho;
        long long base = ((img * h + 2 * yy) * w + 2 * x) * (long long)c + ch;
        long long rs = (long long)w * c;
        float acc[V], t[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        const long long offs[4] = {0, (long long)c, rs, rs + c};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ld16<T>(a + base + offs[q], t);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] += t[j];
            if (b) {
                ld16<T>(b + base + offs[q], t);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] += t[j];
            }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] *= scale;
        st16<T>(y + (((img * ho + yy) * wo + x) * (long long)c + ch), acc);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) unpool2_vec_kernel(const T* __restrict__ gy, T* __restrict__ gx, int n, int h, int w, int c, float scale) {
    constexpr int V = Vec16<T>::N;
    int ho = h / 2, wo = w / 2, cv = c / V;
    long long total = (long long)n * h * w * cv;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % cv) * V;
        long long p = i / cv;
        int x = (int)(p % w); p /= w;
        int yy = (int)(p % h);
        long long img = p / h;
        float v[V];
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] = 0.f;
        if ((yy >> 1) < ho && (x >> 1) < wo) {
            ld16<T>(gy + ((img * ho + (yy >> 1)) * wo + (x >> 1)) * (long long)c + ch, v);
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] *= scale;
        }
        st16<T>(gx + ((img * h + yy) * (long long)w + x) * c + ch, v);
    }
}


// AvgPool2d(2) of (a + b) with up to three outputs from one pass: fp32 y, bf16(y) and bf16(LeakyReLU(y)) -- the operands the next
// ResBlockDown's 1x1 and 3x3 convolutions consume, so no separate cast / activation kernels touch HBM.  c % 4 == 0.
__global__ void __launch_bounds__(256) pool2_multi_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y32,
                                                          bf16* __restrict__ yb, bf16* __restrict__ yl, int n, int h, int w, int c, float scale,
                                                          float slope) {
    int ho = h / 2, wo = w / 2, cv = c / 4;
    long long total = (long long)n * ho * wo * cv;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % cv) * 4;
        long long p = i / cv;
        int x = (int)(p % wo); p /= wo;
        int yy = (int)(p % ho);
        long long img = p / ho;
        long long base = ((img * h + 2 * yy) * w + 2 * x) * (long long)c + ch;
        long long rs = (long long)w * c;
        const long long offs[4] = {0, (long long)c, rs, rs + c};
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float4 t = *reinterpret_cast<const float4*>(a + base + offs[q]);
            acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
            if (b) {
                float4 u = *reinterpret_cast<const float4*>(b + base + offs[q]);
                acc[0] += u.x; acc[1] += u.y; acc[2] += u.z; acc[3] += u.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] *= scale;
        long long o = ((img * ho + yy) * wo + x) * (long long)c + ch;
        if (y32) *reinterpret_cast<float4*>(y32 + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (yb) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[0], acc[1]), p1 = __floats2bfloat162_rn(acc[2], acc[3]);
            uint2 r; r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(yb + o) = r;
        }
        if (yl) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(lrelu_f(acc[0], slope), lrelu_f(acc[1], slope));
            __nv_bfloat162 p1 = __floats2bfloat162_rn(lrelu_f(acc[2], slope), lrelu_f(acc[3], slope));
            uint2 r; r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(yl + o) = r;
        }
    }
}

// out[n,h,w,c] (bf16) = scale * g[n,h/2,w/2,c] (0 outside for odd h/w): the AvgPool backward written directly as the bf16 operand
// of the input- and weight-gradient convolutions.  c % 8 == 0.
__global__ void __launch_bounds__(256) unpool2_cast_kernel(const float* __restrict__ g, bf16* __restrict__ out, int n, int h, int w, int c, float scale) {
    int ho = h / 2, wo = w / 2, cv = c / 8;
    long long total = (long long)n * h * w * cv;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % cv) * 8;
        long long p = i / cv;
        int x = (int)(p % w); p /= w;
        int yy = (int)(p % h);
        long long img = p / h;
        uint4 r = make_uint4(0u, 0u, 0u, 0u);
        if ((yy >> 1) < ho && (x >> 1) < wo) {
            const float* src = g + ((img * ho + (yy >> 1)) * wo + (x >> 1)) * (long long)c + ch;
            float4 t0 = *reinterpret_cast<const float4*>(src), t1 = *reinterpret_cast<const float4*>(src + 4);
            __nv_bfloat162 p0 = __floats2bfloat162_rn(scale * t0.x, scale * t0.y), p1 = __floats2bfloat162_rn(scale * t0.z, scale * t0.w);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(scale * t1.x, scale * t1.y), p3 = __floats2bfloat162_rn(scale * t1.z, scale * t1.w);
            r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
            r.z = *reinterpret_cast<uint32_t*>(&p2); r.w = *reinterpret_cast<uint32_t*>(&p3);
        }
        *reinterpret_cast<uint4*>(out + ((img * h + yy) * (long long)w + x) * c + ch) = r;
    }
}
// even h and w: one thread per SOURCE chunk of 8 channels -- read once (32 B), rounded once, stored to the four positions of its 2x2 window
__global__ void __launch_bounds__(256) unpool2_cast_even_kernel(const float* __restrict__ g, bf16* __restrict__ out, int n, int ho, int wo, int c, float scale) {
    const int cv = c / 8, w = 2 * wo;
    long long total = (long long)n * ho * wo * cv;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ch = (int)(i % cv) * 8;
        long long p = i / cv;
        const int x = (int)(p % wo); p /= wo;
        const int yy = (int)(p % ho);
        const long long img = p / ho;
        const float* src = g + i * 8;
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(src)), t1 = __ldg(reinterpret_cast<const float4*>(src + 4));
        __nv_bfloat162 p0 = __floats2bfloat162_rn(scale * t0.x, scale * t0.y), p1 = __floats2bfloat162_rn(scale * t0.z, scale * t0.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(scale * t1.x, scale * t1.y), p3 = __floats2bfloat162_rn(scale * t1.z, scale * t1.w);
        uint4 r;
        r.x = *reinterpret_cast<uint32_t*>(&p0); r.y = *reinterpret_cast<uint32_t*>(&p1);
        r.z = *reinterpret_cast<uint32_t*>(&p2); r.w = *reinterpret_cast<uint32_t*>(&p3);
        bf16* o = out + ((img * (2 * ho) + 2 * yy) * (long long)w + 2 * x) * c + ch;
        *reinterpret_cast<uint4*>(o) = r;
        *reinterpret_cast<uint4*>(o + c) = r;
        *reinterpret_cast<uint4*>(o + (long long)w * c) = r;
        *reinterpret_cast<uint4*>(o + (long long)w * c + c) = r;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) pool2_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y,
                                                    int n, int h, int w, int c, float scale) {
    int ho = h / 2, wo = w / 2;
    long long total = (long long)n * ho * wo * c;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % c);
        long long p = i / c;
        int x = (int)(p % wo); p /= wo;
        int yy = (int)(p % ho);
        long long img = p / ho;
        long long base = ((img * h + 2 * yy) * w + 2 * x) * (long long)c + ch;
        long long rs = (long long)w * c;
        float s = to_f<T>(a[base]) + to_f<T>(a[base + c]) + to_f<T>(a[base + rs]) + to_f<T>(a[base + rs + c]);
        if (b) s += to_f<T>(b[base]) + to_f<T>(b[base + c]) + to_f<T>(b[base + rs]) + to_f<T>(b[base + rs + c]);
        y[i] = from_f<T>(scale * s);
    }
}

// gx[n,h,w,c] = scale * gy[n,h/2,w/2,c] or 0 outside
template <typename T>
__global__ void __launch_bounds__(256) unpool2_kernel(const T* __restrict__ gy, T* __restrict__ gx, int n, int h, int w, int c, float scale) {
    int ho = h / 2, wo = w / 2;
    long long total = (long long)n * h * w * c;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % c);
        long long p = i / c;
        int x = (int)(p % w); p /= w;
        int yy = (int)(p % h);
        long long img = p / h;
        float v = 0.f;
        if ((yy >> 1) < ho && (x >> 1) < wo) v = scale * to_f<T>(gy[((img * ho + (yy >> 1)) * wo + (x >> 1)) * (long long)c + ch]);
        gx[i] = from_f<T>(v);
    }
}

// NCHW fp32 -> NHWC T via a 32x32 shared-memory transpose of the [C][HW] plane of each image
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int c, int hw) {
    __shared__ float tile[32][33];
    long long img = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* xi = x + img * (long long)c * hw;
    T* yi = y + img * (long long)c * hw;
    for (int j = threadIdx.y; j < 32; j += 8) {
        int cc = c0 + j, p = p0 + threadIdx.x;
        tile[j][threadIdx.x] = (cc < c && p < hw) ? xi[(long long)cc * hw + p] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        int p = p0 + j, cc = c0 + threadIdx.x;
        if (p < hw && cc < c) yi[(long long)p * c + cc] = from_f<T>(tile[threadIdx.x][j]);
    }
}
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int c, int hw) {
    __shared__ float tile[32][33];
    long long img = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const T* xi = x + img * (long long)c * hw;
    float* yi = y + img * (long long)c * hw;
    for (int j = threadIdx.y; j < 32; j += 8) {
        int p = p0 + j, cc = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (p < hw && cc < c) ? to_f<T>(xi[(long long)p * c + cc]) : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        int cc = c0 + j, p = p0 + threadIdx.x;
        if (cc < c && p < hw) yi[(long long)cc * hw + p] = tile[threadIdx.x][j];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) copy_cols_kernel(const T* __restrict__ src, int src_ld, int src_off, T* __restrict__ dst, int dst_ld,
                                                        int dst_off, long long rows, int c) {
    long long total = rows * c;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        long long r = i / c;
        int j = (int)(i % c);
        dst[r * dst_ld + dst_off + j] = src[r * src_ld + src_off + j];
    }
}

// out[pix][t*c + ch] = x[pix + sign * tap_offset(t)][ch] (zero outside the image, zero for the padded tail j >= taps*c).
// Skinny-channel tensors (images, last layers) are unrolled over the filter taps so that their convolutions become dense
// 1x1 tensor-core GEMMs with K (or N) = taps * c.
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ x, T* __restrict__ out, int n, int h, int w, int c, int ks, int sign, int kc) {
    const int pad = (ks - 1) / 2, taps = ks * ks;
    long long total = (long long)n * h * w * kc;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int j = (int)(i % kc);
        long long pix = i / kc;
        float v = 0.f;
        if (j < taps * c) {
            int t = j / c, ch = j - t * c;
            int pw = (int)(pix % w);
            long long r = pix / w;
            int ph = (int)(r % h);
            long long img = r / h;
            int hh = ph + sign * (t / ks - pad), ww = pw + sign * (t % ks - pad);
            if (hh >= 0 && hh < h && ww >= 0 && ww < w) v = to_f<T>(x[((img * h + hh) * (long long)w + ww) * c + ch]);
        }
        out[i] = from_f<T>(v);
    }
}


// bf16 variant writing 16 bytes (8 unrolled elements) per thread: the pixel decode and the store are amortised over 8 gathers
__global__ void __launch_bounds__(256) im2col_vec8_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int n, int h, int w, int c, int ks, int sign,
                                                          int kc) {
    const int pad = (ks - 1) / 2, valid = ks * ks * c, kv = kc / 8;
    long long total = (long long)n * h * w * kv;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int j0 = (int)(i % kv) * 8;
        const long long pix = i / kv;
        const int pw = (int)(pix % w);
        const long long r = pix / w;
        const int ph = (int)(r % h);
        const long long img = r / h;
        const bf16* xi = x + img * (long long)h * w * c;
        __align__(16) bf16 v[8];
        int t = j0 / c, ch = j0 - t * c;
        int tr = t / ks, tq = t - tr * ks;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            bf16 val = __float2bfloat16_rn(0.f);
            if (j0 + e < valid) {
                const int hh = ph + sign * (tr - pad), ww = pw + sign * (tq - pad);
                if (hh >= 0 && hh < h && ww >= 0 && ww < w) val = xi[((long long)hh * w + ww) * c + ch];
            }
            v[e] = val;
            if (++ch == c) { ch = 0; if (++tq == ks) { tq = 0; ++tr; } }
        }
        *reinterpret_cast<uint4*>(out + pix * kc + j0) = *reinterpret_cast<const uint4*>(v);
    }
}

// y[pix][ch] = bias[ch] + sum_t z[pix + tap_offset(t)][t*c + ch]   (adjoint of im2col with sign = +1); z rows have length ld
__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ y, int n, int h,
                                                     int w, int c, int ks, int ld) {
    const int pad = (ks - 1) / 2, taps = ks * ks;
    long long total = (long long)n * h * w * c;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % c);
        long long pix = i / c;
        int pw = (int)(pix % w);
        long long r = pix / w;
        int ph = (int)(r % h);
        long long img = r / h;
        float acc = bias ? bias[ch] : 0.f;
        for (int t = 0; t < taps; ++t) {
            int hh = ph + t / ks - pad, ww = pw + t % ks - pad;
            if (hh >= 0 && hh < h && ww >= 0 && ww < w) acc += z[((img * h + hh) * (long long)w + ww) * ld + t * c + ch];
        }
        y[i] = acc;
    }
}


// First ResBlockDown of an encoder (image input: 1-4 channels): both input-side convolutions in one HBM-bound pass.
//   t[pix][co]   = bf16(LeakyReLU(b_r1[co] + sum_{tap,c} bf16(LeakyReLU(x))[pix+tap][c] * bf16(w_r1[tap][co][c])))     (k x k, 'same')
//   res[pp][co]  = b_l1[co] + sum_c bf16(AvgPool2(x))[pp][c] * bf16(w_l1[co][c])                                        (1x1 at the pooled resolution)
// With K = taps*c <= 36 these are outer products, not GEMMs: the tensor-core route needs an im2col pass and runs at 3 % of peak, bound by
// its epilogue; here the only traffic is the two outputs.  Operands are rounded to bf16 exactly as on the tensor-core path.
// One thread = one pixel x 8 output channels (a 16-byte store); weights live in shared memory as [tap*c][co].
__global__ void __launch_bounds__(256) first_block_kernel(const float* __restrict__ x, const float* __restrict__ w_r1, const float* __restrict__ b_r1,
                                                          const float* __restrict__ w_l1, const float* __restrict__ b_l1, bf16* __restrict__ t_out,
                                                          float* __restrict__ res_out, int n, int h, int w, int c, int co, int ks, float slope) {
    extern __shared__ float wsm[];                 // [taps*c][co] (r1), then [c][co] (l1), then biases
    const int taps = ks * ks, pad = (ks - 1) / 2, kk = taps * c;
    float* w1s = wsm;
    float* wls = w1s + kk * co;
    float* b1s = wls + c * co;
    float* bls = b1s + co;
    // layout [tap*c + ch][half][group][4]: output channel o = group*8 + half*4 + j.  A warp's two 16-byte loads per (tap, ch) then touch
    // consecutive 16-byte chunks (one per channel group): bank-conflict free, the second pixel of the warp is a broadcast
    for (int i = threadIdx.x; i < kk * co; i += blockDim.x) {
        const int k = i / co, o = i - k * co;       // k = tap*c + ch
        const int tap = k / c, ch = k - tap * c;
        const int grp = o >> 3, half = (o >> 2) & 1, j = o & 3;
        w1s[(k * 2 + half) * (co / 2) + grp * 4 + j] = __bfloat162float(__float2bfloat16_rn(w_r1[((long long)tap * co + o) * c + ch]));
    }
    for (int i = threadIdx.x; i < c * co; i += blockDim.x) {
        const int ch = i / co, o = i - ch * co;
        wls[i] = __bfloat162float(__float2bfloat16_rn(w_l1[(long long)o * c + ch]));
    }
    for (int i = threadIdx.x; i < co; i += blockDim.x) { b1s[i] = b_r1[i]; bls[i] = b_l1[i]; }
    __syncthreads();
    const int cg = co / 8;                          // channel groups of 8
    // 32-bit index arithmetic (the host checks n*h*w*co/8 < 2^31): 64-bit divisions per element would dominate this kernel
    const unsigned total = (unsigned)n * h * w * cg;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int g = (int)(i % (unsigned)cg);
        const unsigned pix = i / (unsigned)cg;
        const int pw = (int)(pix % (unsigned)w);
        const unsigned r = pix / (unsigned)w;
        const int ph = (int)(r % (unsigned)h);
        const unsigned img = r / (unsigned)h;
        const float* xi = x + (size_t)img * h * w * c;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = b1s[g * 8 + j];
        int tap = 0;
        for (int dr = -pad; dr <= pad; ++dr) {
            const int hh = ph + dr;
            for (int dq = -pad; dq <= pad; ++dq, ++tap) {
                const int ww = pw + dq;
                if (hh < 0 || hh >= h || ww < 0 || ww >= w) continue;
                const float* xq = xi + (hh * w + ww) * c;
                for (int ch = 0; ch < c; ++ch) {
                    const float xv = __bfloat162float(__float2bfloat16_rn(lrelu_f(__ldg(xq + ch), slope)));
                    const float* wrow = w1s + (tap * c + ch) * co + g * 4;
                    const float4 w0 = *reinterpret_cast<const float4*>(wrow), w1v = *reinterpret_cast<const float4*>(wrow + co / 2);
                    acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
                    acc[4] = fmaf(xv, w1v.x, acc[4]); acc[5] = fmaf(xv, w1v.y, acc[5]); acc[6] = fmaf(xv, w1v.z, acc[6]); acc[7] = fmaf(xv, w1v.w, acc[7]);
                }
            }
        }
        uint4 o;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(lrelu_f(acc[0], slope), lrelu_f(acc[1], slope));
        __nv_bfloat162 p1 = __floats2bfloat162_rn(lrelu_f(acc[2], slope), lrelu_f(acc[3], slope));
        __nv_bfloat162 p2 = __floats2bfloat162_rn(lrelu_f(acc[4], slope), lrelu_f(acc[5], slope));
        __nv_bfloat162 p3 = __floats2bfloat162_rn(lrelu_f(acc[6], slope), lrelu_f(acc[7], slope));
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        *reinterpret_cast<uint4*>(t_out + (size_t)pix * co + g * 8) = o;
        // the thread of the top-left pixel of each 2x2 window also produces the pooled residual for its 8 channels
        if (!(ph & 1) && !(pw & 1) && ph + 1 < h && pw + 1 < w) {
            float rs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) rs[j] = bls[g * 8 + j];
            for (int ch = 0; ch < c; ++ch) {
                const float* q = xi + (ph * w + pw) * c + ch;
                const float xp = __bfloat162float(__float2bfloat16_rn(0.25f * (__ldg(q) + __ldg(q + c) + __ldg(q + w * c) + __ldg(q + w * c + c))));
#pragma unroll
                for (int j = 0; j < 8; ++j) rs[j] = fmaf(xp, wls[ch * co + g * 8 + j], rs[j]);
            }
            float* dst = res_out + (((size_t)img * (h / 2) + ph / 2) * (w / 2) + pw / 2) * co + g * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(rs[0], rs[1], rs[2], rs[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(rs[4], rs[5], rs[6], rs[7]);
        }
    }
}

// Specialised variant (compile-time channel count and filter size, w % 4 == 0): one thread = 4 consecutive pixels of a row x 8 output
// channels.  The 3 x 6 input neighbourhood is loaded (and activated / rounded) once for the four pixels and every weight read from
// shared memory feeds four FMAs, so the kernel stays below the issue limit and runs at the speed of its two output streams.
template <int C, int KS>
__global__ void __launch_bounds__(256) first_block_quad_kernel(const float* __restrict__ x, const float* __restrict__ w_r1, const float* __restrict__ b_r1,
                                                               const float* __restrict__ w_l1, const float* __restrict__ b_l1, bf16* __restrict__ t_out,
                                                               float* __restrict__ res_out, int n, int h, int w, int co, float slope) {
    extern __shared__ float wsm[];
    constexpr int TAPS = KS * KS, PAD = (KS - 1) / 2, KK = TAPS * C, PX = 4, NX = PX + KS - 1;
    float* w1s = wsm;                               // [tap*C + ch][half][group][4]
    float* wls = w1s + KK * co;                     // [ch][co]
    float* b1s = wls + C * co;
    float* bls = b1s + co;
    for (int i = threadIdx.x; i < KK * co; i += blockDim.x) {
        const int k = i / co, o = i - k * co;
        const int tap = k / C, ch = k - tap * C;
        const int grp = o >> 3, half = (o >> 2) & 1, j = o & 3;
        w1s[(k * 2 + half) * (co / 2) + grp * 4 + j] = __bfloat162float(__float2bfloat16_rn(w_r1[((long long)tap * co + o) * C + ch]));
    }
    for (int i = threadIdx.x; i < C * co; i += blockDim.x) {
        const int ch = i / co, o = i - ch * co;
        wls[i] = __bfloat162float(__float2bfloat16_rn(w_l1[(long long)o * C + ch]));
    }
    for (int i = threadIdx.x; i < co; i += blockDim.x) { b1s[i] = b_r1[i]; bls[i] = b_l1[i]; }
    __syncthreads();
    const unsigned cg = co / 8, wq = w / PX;
    const unsigned total = (unsigned)n * h * wq * cg;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int g = (int)(i % cg);
        unsigned r = i / cg;
        const int pw0 = (int)(r % wq) * PX; r /= wq;
        const int ph = (int)(r % (unsigned)h);
        const unsigned img = r / (unsigned)h;
        const float* xi = x + (size_t)img * h * w * C;
        float xa[KS][NX][C];                        // bf16(LeakyReLU(x)) of the neighbourhood, zero outside the image
        float xr[2][PX][C];                         // raw x of rows ph, ph+1 (for the pooled residual)
#pragma unroll
        for (int dr = 0; dr < KS; ++dr) {
            const int hh = ph + dr - PAD;
#pragma unroll
            for (int dq = 0; dq < NX; ++dq) {
                const int ww = pw0 + dq - PAD;
                const bool in = hh >= 0 && hh < h && ww >= 0 && ww < w;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float v = in ? __ldg(xi + (hh * w + ww) * C + ch) : 0.f;
                    xa[dr][dq][ch] = __bfloat162float(__float2bfloat16_rn(lrelu_f(v, slope)));
                    if (dr >= PAD && dr <= PAD + 1 && dq >= PAD && dq < PAD + PX) xr[dr - PAD][dq - PAD][ch] = v;
                }
            }
        }
        float acc[PX][8];
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p][j] = b1s[g * 8 + j];
#pragma unroll
        for (int dr = 0; dr < KS; ++dr)
#pragma unroll
            for (int dq = 0; dq < KS; ++dq)
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float* wrow = w1s + ((dr * KS + dq) * C + ch) * co + g * 4;
                    const float4 w0 = *reinterpret_cast<const float4*>(wrow), w1v = *reinterpret_cast<const float4*>(wrow + co / 2);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1v.x, w1v.y, w1v.z, w1v.w};
#pragma unroll
                    for (int p = 0; p < PX; ++p)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(xa[dr][p + dq][ch], wv[j], acc[p][j]);
                }
        bf16* trow = t_out + (((size_t)img * h + ph) * w + pw0) * co + g * 8;
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            uint4 o;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(lrelu_f(acc[p][0], slope), lrelu_f(acc[p][1], slope));
            __nv_bfloat162 p1 = __floats2bfloat162_rn(lrelu_f(acc[p][2], slope), lrelu_f(acc[p][3], slope));
            __nv_bfloat162 p2 = __floats2bfloat162_rn(lrelu_f(acc[p][4], slope), lrelu_f(acc[p][5], slope));
            __nv_bfloat162 p3 = __floats2bfloat162_rn(lrelu_f(acc[p][6], slope), lrelu_f(acc[p][7], slope));
            o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
            o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(trow + (size_t)p * co) = o;
        }
        if (!(ph & 1)) {                            // even rows also produce the two pooled residual pixels under this quad
#pragma unroll
            for (int pp = 0; pp < PX / 2; ++pp) {
                float rs[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) rs[j] = bls[g * 8 + j];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float xp = __bfloat162float(__float2bfloat16_rn(0.25f * (xr[0][2 * pp][ch] + xr[0][2 * pp + 1][ch] + xr[1][2 * pp][ch] + xr[1][2 * pp + 1][ch])));
#pragma unroll
                    for (int j = 0; j < 8; ++j) rs[j] = fmaf(xp, wls[ch * co + g * 8 + j], rs[j]);
                }
                float* dst = res_out + (((size_t)img * (h / 2) + ph / 2) * (w / 2) + pw0 / 2 + pp) * co + g * 8;
                *reinterpret_cast<float4*>(dst) = make_float4(rs[0], rs[1], rs[2], rs[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(rs[4], rs[5], rs[6], rs[7]);
            }
        }
    }
}

// Weight gradients of the same two image-side convolutions in one pass over the gradients:
//   gw_r1[tap][co][c] = sum_pix gt[pix][co] * bf16(LeakyReLU(x))[pix+tap][c]          (gt: bf16 masked gradient of the k x k conv output)
//   gw_l1[co][c]      = sum_pp bf16(gy)[pp][co] * bf16(AvgPool2(x))[pp][c]           (gy: fp32 gradient of the pooled block output)
// Persistent CTAs walk bands of kWgRows image rows.  The band's gradient rows are staged in shared memory with 16-byte coalesced loads,
// the activated input band (with halo) next to them; thread = (output channel, pixel group) keeps its 9*C + C partial sums in
// registers across all bands: per pixel one gradient read from smem feeds 9*C FMAs whose other operand is a broadcast.  One
// shared-memory reduction and one round of fp32 atomics per CTA at the end (outputs zeroed by the caller).
constexpr int kWgRows = 4;
template <int C>
__global__ void __launch_bounds__(256) first_block_wgrad_kernel(const float* __restrict__ x, const bf16* __restrict__ gt, const float* __restrict__ gy,
                                                                float* __restrict__ gw_r1, float* __restrict__ gw_l1, int n, int h, int w, int co,
                                                                float slope) {
    constexpr int KS = 3, TAPS = 9, KK = TAPS * C;
    extern __shared__ __align__(16) unsigned char smem_wg[];
    const int band_px = kWgRows * w;
    bf16* gs = reinterpret_cast<bf16*>(smem_wg);                                  // [band_px][co]
    float* xs = reinterpret_cast<float*>(smem_wg + (size_t)band_px * co * 2);     // [(kWgRows+2)][w+2][C] activated, zero outside the image
    float* xps = xs + (kWgRows + 2) * (w + 2) * C;                                // [kWgRows/2][w/2][C] bf16(AvgPool2(x))
    float* red = xps + (kWgRows / 2) * (w / 2) * C;                               // [KK + C][co]
    const int groups = blockDim.x / co;              // pixel groups (host guarantees blockDim.x % co == 0)
    const int o = threadIdx.x % co, grp = threadIdx.x / co;
    float acc[KK], accl[C];
#pragma unroll
    for (int k = 0; k < KK; ++k) acc[k] = 0.f;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) accl[ch] = 0.f;
    const int bands_per_img = h / kWgRows, total_bands = n * bands_per_img;
    for (int band = blockIdx.x; band < total_bands; band += gridDim.x) {
        const int img = band / bands_per_img, r0 = (band - img * bands_per_img) * kWgRows;
        const float* xi = x + (size_t)img * h * w * C;
        __syncthreads();                             // previous band fully consumed
        {   // gradient rows: contiguous band_px*co bf16 in global memory
            const uint4* src = reinterpret_cast<const uint4*>(gt + ((size_t)img * h + r0) * w * co);
            uint4* dst = reinterpret_cast<uint4*>(gs);
            for (int i = threadIdx.x; i < band_px * co / 8; i += blockDim.x) dst[i] = src[i];
        }
        for (int i = threadIdx.x; i < (kWgRows + 2) * (w + 2) * C; i += blockDim.x) {
            const int ch = i % C, q = (i / C) % (w + 2), r = i / (C * (w + 2));
            const int hh = r0 + r - 1, ww = q - 1;
            const float v = (hh >= 0 && hh < h && ww >= 0 && ww < w) ? __ldg(xi + (hh * w + ww) * C + ch) : 0.f;
            xs[i] = __bfloat162float(__float2bfloat16_rn(lrelu_f(v, slope)));
        }
        for (int i = threadIdx.x; i < (kWgRows / 2) * (w / 2) * C; i += blockDim.x) {
            const int ch = i % C, q = (i / C) % (w / 2), r = i / (C * (w / 2));
            const float* p = xi + ((r0 + 2 * r) * w + 2 * q) * C + ch;
            xps[i] = __bfloat162float(__float2bfloat16_rn(0.25f * (__ldg(p) + __ldg(p + C) + __ldg(p + w * C) + __ldg(p + w * C + C))));
        }
        __syncthreads();
        // each pixel group walks its own column range of every band row with a sliding 3x3 window: per pixel one gradient read, three
        // window reads (broadcasts) and 9*C FMAs; no divisions in the loop
        const int q_lo = grp * w / groups, q_hi = (grp + 1) * w / groups;
        for (int r = 0; r < kWgRows; ++r) {
            float win[KS][KS][C];
#pragma unroll
            for (int dr = 0; dr < KS; ++dr)
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    win[dr][1][ch] = xs[((r + dr) * (w + 2) + q_lo) * C + ch];
                    win[dr][2][ch] = xs[((r + dr) * (w + 2) + q_lo + 1) * C + ch];
                }
            const bf16* grow_s = gs + (size_t)(r * w) * co + o;
            for (int q = q_lo; q < q_hi; ++q) {
#pragma unroll
                for (int dr = 0; dr < KS; ++dr)
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        win[dr][0][ch] = win[dr][1][ch];
                        win[dr][1][ch] = win[dr][2][ch];
                        win[dr][2][ch] = xs[((r + dr) * (w + 2) + q + 2) * C + ch];
                    }
                const float g = __bfloat162float(grow_s[q * co]);
#pragma unroll
                for (int dr = 0; dr < KS; ++dr)
#pragma unroll
                    for (int dq = 0; dq < KS; ++dq)
#pragma unroll
                        for (int ch = 0; ch < C; ++ch) acc[(dr * KS + dq) * C + ch] = fmaf(g, win[dr][dq][ch], acc[(dr * KS + dq) * C + ch]);
            }
        }
        const float* gyb = gy + (((size_t)img * (h / 2) + r0 / 2) * (w / 2)) * co + o;
        for (int pp = grp; pp < (kWgRows / 2) * (w / 2); pp += groups) {
            const float gl = __bfloat162float(__float2bfloat16_rn(__ldg(gyb + (size_t)pp * co)));
#pragma unroll
            for (int ch = 0; ch < C; ++ch) accl[ch] = fmaf(gl, xps[pp * C + ch], accl[ch]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (KK + C) * co; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KK; ++k) atomicAdd(&red[k * co + o], acc[k]);
#pragma unroll
    for (int ch = 0; ch < C; ++ch) atomicAdd(&red[(KK + ch) * co + o], accl[ch]);
    __syncthreads();
    for (int q = threadIdx.x; q < KK * co; q += blockDim.x) {
        const int k = q / co, oo = q - k * co;      // k = tap*C + ch  ->  gw_r1[tap][oo][ch]
        const int tap = k / C, ch = k - tap * C;
        atomicAdd(&gw_r1[((size_t)tap * co + oo) * C + ch], red[q]);
    }
    for (int q = threadIdx.x; q < C * co; q += blockDim.x) {
        const int ch = q / co, oo = q - ch * co;
        atomicAdd(&gw_l1[(size_t)oo * C + ch], red[KK * co + q]);
    }
}

// Tensor-core variant of the same pass (cout <= 128): per band the weight gradient of the k x k conv is the GEMM
//   gw_r1[co][tap*C+ch] += sum_{pix in band} gt[pix][co] * patch[pix][tap*C+ch]      (M = co, N = 9C padded to 16 / 32, K = pixels)
// on mma.sync.m16n8k16 (bf16 x bf16 -> fp32): the gradient band is staged with cp.async exactly as it lies in memory ([pix][co]) and the
// im2col patch tile [pix][N] is built in shared memory from the activated input band; both operands have the contraction index (pixel)
// as their ROW index, so both fragments come from ldmatrix.trans.  Row pitches are padded by 16 bytes (conflict-free ldmatrix).  Warp w
// owns output channels 16w..16w+15 and keeps its N/8 x 4 accumulators across all bands of the persistent CTA; the 1x1 residual gradient
// stays on the FFMA path (one FMA per pooled gradient element).  The pass is bound by reading gt (2 B/element) and gy (4 B per pooled
// element) once.
// im2col patch tile of a band: ps[pix][tap*C+ch] = bf16(xs[r+dr][q+dq][ch]) (zero beyond 9C), one 16-byte store per 8 columns; the tap
// arithmetic is resolved at compile time (the chunk index is matched against an unrolled loop), one runtime division per item
template <int C, int NP, int PP>
__device__ __forceinline__ void build_patch_tile(bf16* ps, const float* xs, int band_px, int w) {
    constexpr int KK = 9 * C, CH = NP / 8;
    for (int item = threadIdx.x; item < band_px * CH; item += blockDim.x) {
        const int pix = item / CH, chunk = item - pix * CH;
        const int r = pix / w, q = pix - r * w;
        const float* base = xs + (r * (w + 2) + q) * C;
        const int rowp = (w + 2) * C;
        uint32_t pk[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int cc = 0; cc < CH; ++cc) {
            if (chunk == cc) {
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2) {
                    float v[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int nn = cc * 8 + j2 * 2 + e;
                        const int tap = nn / C, ch = nn - tap * C, dr = tap / 3, dq = tap - dr * 3;
                        v[e] = nn < KK ? base[dr * rowp + dq * C + ch] : 0.f;
                    }
                    const __nv_bfloat162 hv = __floats2bfloat162_rn(v[0], v[1]);
                    pk[j2] = *reinterpret_cast<const uint32_t*>(&hv);
                }
            }
        }
        *reinterpret_cast<uint4*>(ps + (size_t)pix * PP + chunk * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_16816_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}

template <int C>
__global__ void __launch_bounds__(256) first_block_wgrad_mma_kernel(const float* __restrict__ x, const bf16* __restrict__ gt, const float* __restrict__ gy,
                                                                    float* __restrict__ partial, int replicas, int n, int h, int w, int co,
                                                                    float slope) {
    constexpr int KS = 3, TAPS = 9, KK = TAPS * C, NP = (KK + 15) / 16 * 16, NT = NP / 8, PP = NP + 8;
    // every CTA adds its sums into one of `replicas` zeroed copies of [gw_r1 (9*co*C) | gw_l1 (co*C)]: hundreds of CTAs adding into the
    // same 5 KB would serialise in two L2 slices; first_block_wgrad_reduce_kernel folds the copies afterwards
    float* gw_r1 = partial + (size_t)(blockIdx.x % replicas) * (KK + C) * co;
    float* gw_l1 = gw_r1 + (size_t)KK * co;
    extern __shared__ __align__(16) unsigned char smem_wg[];
    const int band_px = kWgRows * w, gp = co + 8;                                  // gp, PP: padded row pitches in bf16 elements
    bf16* gs = reinterpret_cast<bf16*>(smem_wg);                                   // [band_px][gp]   gradient band, as in memory
    bf16* ps = gs + (size_t)band_px * gp;                                          // [band_px][PP]   im2col patches of the activated input
    float* xs = reinterpret_cast<float*>(ps + (size_t)band_px * PP);               // [(kWgRows+2)][w+2][C] activated, zero outside the image
    float* xps = xs + (kWgRows + 2) * (w + 2) * C;                                 // [kWgRows/2][w/2][C] bf16(AvgPool2(x))
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = warp * 16;                        // this warp's output channels (idle when m0 >= co)
    float acc[NT][4], accl[4][C];                    // accl: residual gradient of channels l_c4..l_c4+3 over this thread's pooled pixels
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) accl[j][ch] = 0.f;
    const int chunks_per_row = co / 8;               // 16-byte chunks of a gradient row; the host guarantees 256 % chunks_per_row == 0
    const int g_r = threadIdx.x / chunks_per_row, g_c8 = (threadIdx.x - g_r * chunks_per_row) * 8, g_rstep = blockDim.x / chunks_per_row;
    const int l_c4 = (threadIdx.x % (co / 4)) * 4, l_grp = threadIdx.x / (co / 4), l_groups = blockDim.x / (co / 4);
    const int bands_per_img = h / kWgRows, total_bands = n * bands_per_img;
    for (int band = blockIdx.x; band < total_bands; band += gridDim.x) {
        const int img = band / bands_per_img, r0 = (band - img * bands_per_img) * kWgRows;
        const float* xi = x + (size_t)img * h * w * C;
        __syncthreads();                             // previous band fully consumed
        {
            const bf16* src = gt + ((size_t)img * h + r0) * w * co;
            for (int r = g_r; r < band_px; r += g_rstep) cp_async_16(gs + (size_t)r * gp + g_c8, src + (size_t)r * co + g_c8);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int i = threadIdx.x; i < (kWgRows + 2) * (w + 2) * C; i += blockDim.x) {
            const int ch = i % C, q = (i / C) % (w + 2), r = i / (C * (w + 2));
            const int hh = r0 + r - 1, ww = q - 1;
            const float v = (hh >= 0 && hh < h && ww >= 0 && ww < w) ? __ldg(xi + (hh * w + ww) * C + ch) : 0.f;
            xs[i] = lrelu_f(v, slope);
        }
        for (int i = threadIdx.x; i < (kWgRows / 2) * (w / 2) * C; i += blockDim.x) {
            const int ch = i % C, q = (i / C) % (w / 2), r = i / (C * (w / 2));
            const float* p = xi + ((r0 + 2 * r) * w + 2 * q) * C + ch;
            xps[i] = __bfloat162float(__float2bfloat16_rn(0.25f * (__ldg(p) + __ldg(p + C) + __ldg(p + w * C) + __ldg(p + w * C + C))));
        }
        __syncthreads();
        build_patch_tile<C, NP, PP>(ps, xs, band_px, w);
        // 1x1 residual gradient (FFMA): thread = (output channel o, pixel group)
        {
            const float* gyb = gy + (((size_t)img * (h / 2) + r0 / 2) * (w / 2)) * co + l_c4;
#pragma unroll 4
            for (int pp = l_grp; pp < (kWgRows / 2) * (w / 2); pp += l_groups) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(gyb + (size_t)pp * co));
                const float gl[4] = {__bfloat162float(__float2bfloat16_rn(g4.x)), __bfloat162float(__float2bfloat16_rn(g4.y)),
                                     __bfloat162float(__float2bfloat16_rn(g4.z)), __bfloat162float(__float2bfloat16_rn(g4.w))};
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float xp = xps[pp * C + ch];
#pragma unroll
                    for (int j = 0; j < 4; ++j) accl[j][ch] = fmaf(gl[j], xp, accl[j][ch]);
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (m0 < co) {
            // ldmatrix.trans row addresses: A matrices (co 0-7 | 8-15) x (pix 0-7 | 8-15); B matrices (pix 0-7 | 8-15) x (n-tile pair)
            const bf16* a_ptr = gs + (size_t)((lane & 7) + ((lane >> 4) << 3)) * gp + m0 + (((lane >> 3) & 1) << 3);
            const bf16* b_ptr = ps + (size_t)((lane & 7) + (((lane >> 3) & 1) << 3)) * PP + ((lane >> 4) << 3);
            for (int k0 = 0; k0 < band_px; k0 += 16) {
                uint32_t a[4];
                ldmatrix_x4_trans(a, a_ptr + (size_t)k0 * gp);
#pragma unroll
                for (int t2 = 0; t2 < NT / 2; ++t2) {
                    uint32_t b[4];
                    ldmatrix_x4_trans(b, b_ptr + (size_t)k0 * PP + t2 * 16);
                    mma_16816_bf16(acc[2 * t2], a, b[0], b[1]);
                    mma_16816_bf16(acc[2 * t2 + 1], a, b[2], b[3]);
                }
            }
        }
    }
    if (m0 < co) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int oo = m0 + (lane >> 2) + ((e >> 1) << 3), nn = t * 8 + ((lane & 3) << 1) + (e & 1);      // C fragment: rows g / g+8, cols 2q / 2q+1
                if (nn < KK) {
                    const int tap = nn / C, ch = nn - tap * C;
                    atomicAdd(&gw_r1[((size_t)tap * co + oo) * C + ch], acc[t][e]);
                }
            }
    }
    __syncthreads();                                 // all MMA reads of the last band are done: the patch tile is free
    float* lred = reinterpret_cast<float*>(ps);      // [co][C]
    for (int i = threadIdx.x; i < co * C; i += blockDim.x) lred[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) atomicAdd(&lred[(l_c4 + j) * C + ch], accl[j][ch]);
    __syncthreads();
    for (int i = threadIdx.x; i < co * C; i += blockDim.x) atomicAdd(&gw_l1[i], lred[i]);
}

// out[i] = sum over replicas of partial[r][i]; the first n1 values go to gw_r1, the rest to gw_l1
__global__ void __launch_bounds__(256) first_block_wgrad_reduce_kernel(const float* __restrict__ partial, int replicas, int n1, int n2, float* __restrict__ gw_r1,
                                                                       float* __restrict__ gw_l1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, total = n1 + n2;
    if (i >= total) return;
    float sum = 0.f;
    for (int r = 0; r < replicas; ++r) sum += partial[(size_t)r * total + i];
    if (i < n1) gw_r1[i] = sum; else gw_l1[i - n1] = sum;
}

// Tensor-core variant of the first-block forward (cout 64 / 128): per band of kWgRows image rows the k x k conv is the GEMM
//   acc[pix][co] = sum_k patch[pix][k] * W[co][k]        (M = pixels, N = co, K = 9C padded to 16 / 32)
// on mma.sync.m16n8k16: the im2col patch tile [pix][K] is built in shared memory from the activated band, the weights are kept
// transposed [co][K] in shared memory for the lifetime of the persistent CTA (both K-contiguous: plain ldmatrix).  Warp = 16 pixels x all
// output channels; epilogue bias + LeakyReLU -> bf16 pairs -> per-warp staging tile -> 16-byte coalesced stores (the 16 x co tile is one
// contiguous 4 KB run of the output).  The pooled 1x1 residual (an outer product) stays on FFMA with float4 stores.
__device__ __forceinline__ void ldmatrix_x4_plain(uint32_t (&r)[4], const void* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

template <int C, int CO>
__global__ void __launch_bounds__(256) first_block_mma_kernel(const float* __restrict__ x, const float* __restrict__ w_r1, const float* __restrict__ b_r1,
                                                              const float* __restrict__ w_l1, const float* __restrict__ b_l1, bf16* __restrict__ t_out,
                                                              float* __restrict__ res_out, int n, int h, int w, float slope) {
    constexpr int KS = 3, TAPS = 9, KK = TAPS * C, NP = (KK + 15) / 16 * 16, KSTEPS = NP / 16, PP = NP + 8, NT = CO / 8, SP = CO + 8;
    extern __shared__ __align__(16) unsigned char smem_fb[];
    const int band_px = kWgRows * w;
    bf16* wT = reinterpret_cast<bf16*>(smem_fb);                                   // [CO][PP]        bf16(W)[co][tap*C+ch], zero padded
    bf16* ps = wT + CO * PP;                                                       // [band_px][PP]   im2col patches of bf16(LeakyReLU(x))
    bf16* stage = ps + (size_t)band_px * PP;                                       // [8 warps][16][SP]
    float* xs = reinterpret_cast<float*>(stage + 8 * 16 * SP);                     // [(kWgRows+2)][w+2][C]
    float* xps = xs + (kWgRows + 2) * (w + 2) * C;                                 // [kWgRows/2][w/2][C] bf16(AvgPool2(x))
    float* wls = xps + (kWgRows / 2) * (w / 2) * C;                                // [C][CO]
    float* b1s = wls + C * CO;
    float* bls = b1s + CO;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < CO * PP; i += blockDim.x) {
        const int o = i / PP, k = i - o * PP;
        float v = 0.f;
        if (k < KK) {
            const int tap = k / C, ch = k - tap * C;
            v = w_r1[((size_t)tap * CO + o) * C + ch];
        }
        wT[i] = __float2bfloat16_rn(v);
    }
    for (int i = threadIdx.x; i < C * CO; i += blockDim.x) {
        const int ch = i / CO, o = i - ch * CO;
        wls[i] = __bfloat162float(__float2bfloat16_rn(w_l1[(size_t)o * C + ch]));
    }
    for (int i = threadIdx.x; i < CO; i += blockDim.x) { b1s[i] = b_r1[i]; bls[i] = b_l1[i]; }
    bf16* my_stage = stage + (size_t)warp * 16 * SP;
    const int bands_per_img = h / kWgRows, total_bands = n * bands_per_img;
    for (int band = blockIdx.x; band < total_bands; band += gridDim.x) {
        const int img = band / bands_per_img, r0 = (band - img * bands_per_img) * kWgRows;
        const float* xi = x + (size_t)img * h * w * C;
        __syncthreads();                             // previous band fully consumed (and, first time, the weights are in place)
        for (int i = threadIdx.x; i < (kWgRows + 2) * (w + 2) * C; i += blockDim.x) {
            const int ch = i % C, q = (i / C) % (w + 2), r = i / (C * (w + 2));
            const int hh = r0 + r - 1, ww = q - 1;
            const float v = (hh >= 0 && hh < h && ww >= 0 && ww < w) ? __ldg(xi + (hh * w + ww) * C + ch) : 0.f;
            xs[i] = lrelu_f(v, slope);
        }
        for (int i = threadIdx.x; i < (kWgRows / 2) * (w / 2) * C; i += blockDim.x) {
            const int ch = i % C, q = (i / C) % (w / 2), r = i / (C * (w / 2));
            const float* p = xi + ((r0 + 2 * r) * w + 2 * q) * C + ch;
            xps[i] = __bfloat162float(__float2bfloat16_rn(0.25f * (__ldg(p) + __ldg(p + C) + __ldg(p + w * C) + __ldg(p + w * C + C))));
        }
        __syncthreads();
        build_patch_tile<C, NP, PP>(ps, xs, band_px, w);
        // pooled 1x1 residual: the band's two pooled rows are one contiguous [w][CO] run of the output
        {
            float* rdst = res_out + (((size_t)img * (h / 2) + r0 / 2) * (w / 2)) * CO;
            for (int i = threadIdx.x; i < (kWgRows / 2) * (w / 2) * (CO / 4); i += blockDim.x) {
                const int pp = i / (CO / 4), c4 = (i - pp * (CO / 4)) * 4;
                float4 o = *reinterpret_cast<const float4*>(bls + c4);
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float xp = xps[pp * C + ch];
                    const float4 wv = *reinterpret_cast<const float4*>(wls + ch * CO + c4);
                    o.x = fmaf(xp, wv.x, o.x); o.y = fmaf(xp, wv.y, o.y); o.z = fmaf(xp, wv.z, o.z); o.w = fmaf(xp, wv.w, o.w);
                }
                *reinterpret_cast<float4*>(rdst + (size_t)pp * CO + c4) = o;
            }
        }
        __syncthreads();
        bf16* tdst = t_out + ((size_t)img * h + r0) * w * CO;
        for (int mt = warp; mt < band_px / 16; mt += 8) {
            uint32_t a[KSTEPS][4];
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)     // A 16x16: (rows 0-7,k 0-7) (rows 8-15,k 0-7) (rows 0-7,k 8-15) (rows 8-15,k 8-15)
                ldmatrix_x4_plain(a[ks], ps + (size_t)(mt * 16 + (lane & 15)) * PP + ks * 16 + ((lane >> 4) << 3));
            float acc[NT][4];                        // start from the bias (columns 8t + 2q, 8t + 2q + 1 of both row halves)
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const float2 bb = *reinterpret_cast<const float2*>(b1s + t * 8 + ((lane & 3) << 1));
                acc[t][0] = bb.x; acc[t][1] = bb.y; acc[t][2] = bb.x; acc[t][3] = bb.y;
            }
#pragma unroll
            for (int t2 = 0; t2 < NT / 2; ++t2)
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {   // B, two n8 tiles per x4: (n 0-7,k 0-7) (n 0-7,k 8-15) (n 8-15,k 0-7) (n 8-15,k 8-15)
                    uint32_t b[4];
                    ldmatrix_x4_plain(b, wT + (size_t)(t2 * 16 + (lane & 7) + ((lane >> 4) << 3)) * PP + ks * 16 + (((lane >> 3) & 1) << 3));
                    mma_16816_bf16(acc[2 * t2], a[ks], b[0], b[1]);
                    mma_16816_bf16(acc[2 * t2 + 1], a[ks], b[2], b[3]);
                }
            __syncwarp();                            // the previous tile's staging reads are done
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                const int col = t * 8 + ((lane & 3) << 1);
                // LeakyReLU(v) = max(v, slope * v) for 0 <= slope <= 1 (checked by the host)
                const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(acc[t][0], slope * acc[t][0]), fmaxf(acc[t][1], slope * acc[t][1]));
                const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaxf(acc[t][2], slope * acc[t][2]), fmaxf(acc[t][3], slope * acc[t][3]));
                *reinterpret_cast<__nv_bfloat162*>(my_stage + (size_t)(lane >> 2) * SP + col) = lo;
                *reinterpret_cast<__nv_bfloat162*>(my_stage + (size_t)((lane >> 2) + 8) * SP + col) = hi;
            }
            __syncwarp();
            constexpr int CH = CO / 8;               // 16-byte chunks per pixel row
            bf16* tile = tdst + (size_t)mt * 16 * CO;
#pragma unroll
            for (int i = lane; i < 16 * CH; i += 32) {
                const int r = i / CH, c8 = (i - r * CH) * 8;
                *reinterpret_cast<uint4*>(tile + (size_t)r * CO + c8) = *reinterpret_cast<const uint4*>(my_stage + (size_t)r * SP + c8);
            }
        }
    }
}

// The conv-operand producer: out = f(x) in the operand dtype, f = identity / LeakyReLU / nearest-upsample x2 (out is [n,2h,2w,c]).
// One pass (4 B read, 2 B written per element on the bf16 path) instead of activation kernel + cast kernel.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) operand_prepare_kernel(const TI* __restrict__ x, TO* __restrict__ out, int n, int h, int w, int c, int mode,
                                                              float slope, bool vec) {
    const int V = vec ? 8 : 1;
    const int cv = c / V;
    const int oh = mode == 2 ? 2 * h : h, ow = mode == 2 ? 2 * w : w;
    long long total = (long long)n * oh * ow * cv;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int ch = (int)(i % cv) * V;
        long long p = i / cv;
        long long src;
        if (mode == 2) {
            int xo = (int)(p % ow); long long r = p / ow;
            int yo = (int)(r % oh);
            long long img = r / oh;
            src = ((img * h + (yo >> 1)) * (long long)w + (xo >> 1)) * c + ch;
        } else {
            src = p * c + ch;
        }
        long long dst = p * c + ch;
        if (vec) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; j += 16 / (int)sizeof(TI)) ld16<TI>(x + src + j, *reinterpret_cast<float(*)[16 / sizeof(TI)]>(&f[j]));
            if (mode == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = lrelu_f(f[j], slope);
            }
#pragma unroll
            for (int j = 0; j < 8; j += 16 / (int)sizeof(TO)) st16<TO>(out + dst + j, *reinterpret_cast<float(*)[16 / sizeof(TO)]>(&f[j]));
        } else {
            float v = to_f<TI>(x[src]);
            out[dst] = from_f<TO>(mode == 1 ? lrelu_f(v, slope) : v);
        }
    }
}

// gx = g * (ref > 0 ? 1 : slope) with the mask taken from a tensor of another dtype (the saved bf16 operand)
template <typename TR>
__global__ void __launch_bounds__(256) lrelu_bwd_ref_kernel(const float* __restrict__ g, const TR* __restrict__ ref, float* __restrict__ gx, long long n,
                                                            float slope) {
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) gx[i] = to_f<TR>(ref[i]) > 0.f ? g[i] : g[i] * slope;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n, bool vec) {
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    long long nv = vec ? n / 8 : 0;                     // 8 elements per thread: 32 B (fp32) / 16 B (bf16) accesses
    for (long long i = tid; i < nv; i += stride) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; j += 16 / (int)sizeof(TI)) ld16<TI>(x + i * 8 + j, *reinterpret_cast<float(*)[16 / sizeof(TI)]>(&f[j]));
#pragma unroll
        for (int j = 0; j < 8; j += 16 / (int)sizeof(TO)) st16<TO>(y + i * 8 + j, *reinterpret_cast<float(*)[16 / sizeof(TO)]>(&f[j]));
    }
    for (long long i = nv * 8 + tid; i < n; i += stride) y[i] = from_f<TO>(to_f<TI>(x[i]));
}

}  // namespace gim

using namespace gim;

extern "C" {

int gim_version(void) { return 100; }
const char* gim_last_error(void) { return g_err; }
long long gim_launch_count(int reset) {
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}
int gim_set_deterministic(int on) {
    int old = g_deterministic;
    g_deterministic = on ? 1 : 0;
    return old;
}

int gim_lrelu_fwd(const void* x, void* y, long long n, float slope, int dtype, gim_stream_t s) {
    GIM_DISPATCH_DTYPE(dtype, return (launch_map<T, 1>(x, nullptr, nullptr, y, n, (cudaStream_t)s, LreluF{slope}, "lrelu_fwd")));
}
int gim_lrelu_bwd(const void* gy, const void* x, void* gx, long long n, float slope, int dtype, gim_stream_t s) {
    GIM_DISPATCH_DTYPE(dtype, return (launch_map<T, 2>(gy, x, nullptr, gx, n, (cudaStream_t)s, LreluB{slope}, "lrelu_bwd")));
}
int gim_tanh_fwd(const void* x, void* y, long long n, int dtype, gim_stream_t s) {
    GIM_DISPATCH_DTYPE(dtype, return (launch_map<T, 1>(x, nullptr, nullptr, y, n, (cudaStream_t)s, TanhF{}, "tanh_fwd")));
}
int gim_tanh_bwd(const void* gy, const void* y, void* gx, long long n, int dtype, gim_stream_t s) {
    GIM_DISPATCH_DTYPE(dtype, return (launch_map<T, 2>(gy, y, nullptr, gx, n, (cudaStream_t)s, TanhB{}, "tanh_bwd")));
}
int gim_axpby(const void* x, const void* y, void* out, long long n, float alpha, float beta, int dtype, gim_stream_t s) {
    if (y) {
        GIM_DISPATCH_DTYPE(dtype, return (launch_map<T, 2>(x, y, nullptr, out, n, (cudaStream_t)s, Axpby{alpha, beta}, "axpby")));
    }
    GIM_DISPATCH_DTYPE(dtype, return (launch_map<T, 1>(x, nullptr, nullptr, out, n, (cudaStream_t)s, Ax{alpha}, "ax")));
}
int gim_scale_dev(const void* x, const float* scalar, void* out, long long n, int dtype, gim_stream_t s) {
    if (n <= 0) return GIM_OK;
    GIM_DISPATCH_DTYPE(dtype, (scale_dev_kernel<T><<<ew_grid(n, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, scalar, (T*)out, n)));
    return check_launch("scale_dev");
}
int gim_dot(const void* x, const void* y, float* out, long long n, int dtype, gim_stream_t s) {
    if (cudaMemsetAsync(out, 0, sizeof(float), (cudaStream_t)s) != cudaSuccess) return fail(GIM_E_CUDA, "dot memset");
    if (n <= 0) return GIM_OK;
    int grid = ew_grid(n, 256, 16);
    if (grid > 2 * num_sms()) grid = 2 * num_sms();
    if (deterministic()) grid = 1;
    GIM_DISPATCH_DTYPE(dtype, (dot_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)y, out, n)));
    return check_launch("dot");
}
int gim_pool2_sum(const void* a, const void* b, void* y, int n, int h, int wd, int c, float scale, int dtype, gim_stream_t s) {
    long long total = (long long)n * (h / 2) * (wd / 2) * c;
    if (total <= 0) return GIM_OK;
    bool vec = aligned16(a) && aligned16(y) && (!b || aligned16(b));
    if (dtype == GIM_F32 && vec && c % 4 == 0)
        pool2_vec_kernel<float><<<ew_grid(total / 4, 256, 1), 256, 0, (cudaStream_t)s>>>((const float*)a, (const float*)b, (float*)y, n, h, wd, c, scale);
    else if (dtype == GIM_BF16 && vec && c % 8 == 0)
        pool2_vec_kernel<bf16><<<ew_grid(total / 8, 256, 1), 256, 0, (cudaStream_t)s>>>((const bf16*)a, (const bf16*)b, (bf16*)y, n, h, wd, c, scale);
    else
        GIM_DISPATCH_DTYPE(dtype, (pool2_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>((const T*)a, (const T*)b, (T*)y, n, h, wd, c, scale)));
    return check_launch("pool2_sum");
}
int gim_unpool2_bcast(const void* gy, void* gx, int n, int h, int wd, int c, float scale, int dtype, gim_stream_t s) {
    long long total = (long long)n * h * wd * c;
    if (total <= 0) return GIM_OK;
    bool vec = aligned16(gy) && aligned16(gx);
    if (dtype == GIM_F32 && vec && c % 4 == 0)
        unpool2_vec_kernel<float><<<ew_grid(total / 4, 256, 1), 256, 0, (cudaStream_t)s>>>((const float*)gy, (float*)gx, n, h, wd, c, scale);
    else if (dtype == GIM_BF16 && vec && c % 8 == 0)
        unpool2_vec_kernel<bf16><<<ew_grid(total / 8, 256, 1), 256, 0, (cudaStream_t)s>>>((const bf16*)gy, (bf16*)gx, n, h, wd, c, scale);
    else
        GIM_DISPATCH_DTYPE(dtype, (unpool2_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>((const T*)gy, (T*)gx, n, h, wd, c, scale)));
    return check_launch("unpool2_bcast");
}
int gim_pool2_multi(const float* a, const float* b, float* y32, void* y_bf16, void* y_lrelu_bf16, int n, int h, int wd, int c, float scale, float slope,
                    gim_stream_t s) {
    long long total = (long long)n * (h / 2) * (wd / 2) * c;
    if (total <= 0) return GIM_OK;
    GIM_REQUIRE(c % 4 == 0 && aligned16(a) && (!b || aligned16(b)) && (!y32 || aligned16(y32)) && (!y_bf16 || aligned16(y_bf16)) &&
                    (!y_lrelu_bf16 || aligned16(y_lrelu_bf16)), "pool2_multi: needs c % 4 == 0 and 16-byte aligned tensors");
    pool2_multi_kernel<<<ew_grid(total / 4, 256, 1), 256, 0, (cudaStream_t)s>>>(a, b, y32, (bf16*)y_bf16, (bf16*)y_lrelu_bf16, n, h, wd, c, scale, slope);
    return check_launch("pool2_multi");
}
int gim_unpool2_cast(const float* gy, void* gx_bf16, int n, int h, int wd, int c, float scale, gim_stream_t s) {
    long long total = (long long)n * h * wd * c;
    if (total <= 0) return GIM_OK;
    GIM_REQUIRE(c % 8 == 0 && aligned16(gy) && aligned16(gx_bf16), "unpool2_cast: needs c % 8 == 0 and 16-byte aligned tensors");
    if (!(h & 1) && !(wd & 1)) unpool2_cast_even_kernel<<<ew_grid(total / 32, 256, 1), 256, 0, (cudaStream_t)s>>>(gy, (bf16*)gx_bf16, n, h / 2, wd / 2, c, scale);
    else unpool2_cast_kernel<<<ew_grid(total / 8, 256, 1), 256, 0, (cudaStream_t)s>>>(gy, (bf16*)gx_bf16, n, h, wd, c, scale);
    return check_launch("unpool2_cast");
}
int gim_first_block_fwd(const float* x, const float* w_r1, const float* b_r1, const float* w_l1, const float* b_l1, void* t_bf16, float* res_pooled,
                        int n, int h, int wd, int c, int cout, int ksize, float slope, gim_stream_t s) {
    GIM_REQUIRE(n > 0 && h > 1 && wd > 1 && c > 0 && cout > 0, "first_block_fwd: empty shape");
    GIM_REQUIRE(ksize >= 1 && (ksize & 1) && c * ksize * ksize <= 64 && cout % 8 == 0 && !(h & 1) && !(wd & 1), "first_block_fwd: unsupported shape");
    GIM_REQUIRE(aligned16(t_bf16) && aligned16(res_pooled), "first_block_fwd: outputs must be 16-byte aligned");
    GIM_REQUIRE((long long)n * h * wd * (cout / 8) < 2147483647LL, "first_block_fwd: too many elements for 32-bit indexing");
    const size_t smem = sizeof(float) * ((size_t)ksize * ksize * c * cout + (size_t)c * cout + 2 * (size_t)cout);
    GIM_REQUIRE(smem <= 96 * 1024, "first_block_fwd: weights do not fit in shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(first_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess)
            return fail(GIM_E_CUDA, "first_block_fwd: cannot raise dynamic shared memory limit");
        attr_set = true;
    }
    const long long total = (long long)n * h * wd * (cout / 8);
    static const bool fwd_simt_only = getenv("GIM_FB_FWD_SIMT") != nullptr;
    if (!fwd_simt_only && ksize == 3 && (c == 1 || c == 3) && (cout == 128 || cout == 64) && wd % 4 == 0 && h % kWgRows == 0 && slope >= 0.f && slope <= 1.f) {     // tensor-core pass
        const int np = (9 * c + 15) / 16 * 16;
        const size_t smem_mma = 2 * ((size_t)cout * (np + 8) + (size_t)kWgRows * wd * (np + 8) + (size_t)8 * 16 * (cout + 8)) +
                                sizeof(float) * ((size_t)(kWgRows + 2) * (wd + 2) * c + (size_t)(kWgRows / 2) * (wd / 2) * c + (size_t)c * cout + 2 * (size_t)cout);
        GIM_REQUIRE(smem_mma <= 200 * 1024, "first_block_fwd: band does not fit in shared memory");
        const void* fn = c == 1 ? (cout == 128 ? (const void*)first_block_mma_kernel<1, 128> : (const void*)first_block_mma_kernel<1, 64>)
                                : (cout == 128 ? (const void*)first_block_mma_kernel<3, 128> : (const void*)first_block_mma_kernel<3, 64>);
        static bool mma_attr = false;
        if (!mma_attr) {
            if (cudaFuncSetAttribute(first_block_mma_kernel<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
                cudaFuncSetAttribute(first_block_mma_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
                cudaFuncSetAttribute(first_block_mma_kernel<3, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
                cudaFuncSetAttribute(first_block_mma_kernel<3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
                return fail(GIM_E_CUDA, "first_block_fwd: cannot raise dynamic shared memory limit");
            mma_attr = true;
        }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, smem_mma) != cudaSuccess || per_sm < 1) per_sm = 1;
        const int bands = n * (h / kWgRows);
        int grid = per_sm * num_sms();
        if (grid > bands) grid = bands;
        void* args[] = {(void*)&x, (void*)&w_r1, (void*)&b_r1, (void*)&w_l1, (void*)&b_l1, (void*)&t_bf16, (void*)&res_pooled, (void*)&n, (void*)&h, (void*)&wd, (void*)&slope};
        if (cudaLaunchKernel(fn, dim3(grid), dim3(256), args, smem_mma, (cudaStream_t)s) != cudaSuccess) return fail(GIM_E_CUDA, "first_block_mma launch failed");
        return check_launch("first_block_mma");
    }
    if (ksize == 3 && (c == 1 || c == 3) && wd % 4 == 0) {      // the two image formats of the GIM configs: grey and RGB, 3x3
        static bool quad_attr = false;
        if (!quad_attr) {
            if (cudaFuncSetAttribute(first_block_quad_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess ||
                cudaFuncSetAttribute(first_block_quad_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess)
                return fail(GIM_E_CUDA, "first_block_fwd: cannot raise dynamic shared memory limit");
            quad_attr = true;
        }
        int gq = ew_grid(total / 4, 256, 2);
        if (gq > 8 * num_sms()) gq = 8 * num_sms();
        if (c == 1) first_block_quad_kernel<1, 3><<<gq, 256, smem, (cudaStream_t)s>>>(x, w_r1, b_r1, w_l1, b_l1, (bf16*)t_bf16, res_pooled, n, h, wd, cout, slope);
        else first_block_quad_kernel<3, 3><<<gq, 256, smem, (cudaStream_t)s>>>(x, w_r1, b_r1, w_l1, b_l1, (bf16*)t_bf16, res_pooled, n, h, wd, cout, slope);
        return check_launch("first_block_quad");
    }
    int grid = ew_grid(total, 256, 4);
    if (grid > 8 * num_sms()) grid = 8 * num_sms();            // every CTA stages the weights once: keep them few and long-lived
    first_block_kernel<<<grid, 256, smem, (cudaStream_t)s>>>(x, w_r1, b_r1, w_l1, b_l1, (bf16*)t_bf16, res_pooled, n, h, wd, c, cout, ksize, slope);
    return check_launch("first_block_fwd");
}
int gim_first_block_wgrad(const float* x, const void* gt_bf16, const float* gy_pooled, float* gw_r1, float* gw_l1, float* scratch, long long scratch_floats,
                          int n, int h, int wd, int c, int cout, int ksize, float slope, gim_stream_t s) {
    GIM_REQUIRE(n > 0 && h >= kWgRows && wd > 1 && h % kWgRows == 0 && !(wd & 1), "first_block_wgrad: h must be a multiple of 4 and w even");
    GIM_REQUIRE(ksize == 3 && (c == 1 || c == 3) && cout % 8 == 0 && cout <= 256 && 256 % cout == 0, "first_block_wgrad: unsupported shape");
    GIM_REQUIRE(aligned16(gt_bf16), "first_block_wgrad: gradient tensor must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)s;
    static const bool simt_only = getenv("GIM_FB_WGRAD_SIMT") != nullptr;
    const long long out_floats = 10LL * cout * c;
    const bool mma = !simt_only && cout % 16 == 0 && cout <= 128 && wd % 4 == 0 && scratch && scratch_floats >= out_floats;
    if (!mma && (cudaMemsetAsync(gw_r1, 0, sizeof(float) * 9 * (size_t)cout * c, st) != cudaSuccess ||
                 cudaMemsetAsync(gw_l1, 0, sizeof(float) * (size_t)cout * c, st) != cudaSuccess))
        return fail(GIM_E_CUDA, "first_block_wgrad memset");
    if (mma) {                                        // tensor-core pass (band_px = 4 * wd is a multiple of 16)
        int replicas = (int)(scratch_floats / out_floats);
        if (replicas > 32) replicas = 32;
        if (cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)replicas * out_floats, st) != cudaSuccess) return fail(GIM_E_CUDA, "first_block_wgrad memset");
        const int np = (9 * c + 15) / 16 * 16;
        const size_t smem_mma = (size_t)kWgRows * wd * ((cout + 8) + (np + 8)) * 2 + sizeof(float) * ((size_t)(kWgRows + 2) * (wd + 2) * c + (size_t)(kWgRows / 2) * (wd / 2) * c);
        GIM_REQUIRE(smem_mma <= 200 * 1024, "first_block_wgrad: band does not fit in shared memory");
        static bool mma_attr = false;
        if (!mma_attr) {
            if (cudaFuncSetAttribute(first_block_wgrad_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
                cudaFuncSetAttribute(first_block_wgrad_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
                return fail(GIM_E_CUDA, "first_block_wgrad: cannot raise dynamic shared memory limit");
            mma_attr = true;
        }
        const int bands = n * (h / kWgRows);
        int per_sm = 0;                               // persistent grid = exactly the resident CTAs (registers and smem both count)
        cudaError_t oe = c == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, first_block_wgrad_mma_kernel<1>, 256, smem_mma)
                                : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, first_block_wgrad_mma_kernel<3>, 256, smem_mma);
        if (oe != cudaSuccess || per_sm < 1) per_sm = 1;
        int grid = per_sm * num_sms();
        if (grid > bands) grid = bands;
        if (deterministic()) { grid = 1; replicas = 1; }
        if (c == 1) first_block_wgrad_mma_kernel<1><<<grid, 256, smem_mma, st>>>(x, (const bf16*)gt_bf16, gy_pooled, scratch, replicas, n, h, wd, cout, slope);
        else first_block_wgrad_mma_kernel<3><<<grid, 256, smem_mma, st>>>(x, (const bf16*)gt_bf16, gy_pooled, scratch, replicas, n, h, wd, cout, slope);
        int rc = check_launch("first_block_wgrad_mma");
        if (rc != GIM_OK) return rc;
        first_block_wgrad_reduce_kernel<<<(unsigned)((out_floats + 255) / 256), 256, 0, st>>>(scratch, replicas, 9 * cout * c, cout * c, gw_r1, gw_l1);
        return check_launch("first_block_wgrad_reduce");
    }
    const size_t smem = (size_t)kWgRows * wd * cout * 2 + sizeof(float) * ((size_t)(kWgRows + 2) * (wd + 2) * c + (size_t)(kWgRows / 2) * (wd / 2) * c + (size_t)(9 * c + c) * cout);
    GIM_REQUIRE(smem <= 200 * 1024, "first_block_wgrad: band does not fit in shared memory");
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(first_block_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(first_block_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return fail(GIM_E_CUDA, "first_block_wgrad: cannot raise dynamic shared memory limit");
        attr_set = true;
    }
    const int bands = n * (h / kWgRows);
    const int per_sm = smem <= 48 * 1024 ? 4 : (smem <= 100 * 1024 ? 2 : 1);
    int grid = per_sm * num_sms();
    if (grid > bands) grid = bands;
    if (deterministic()) grid = 1;
    if (c == 1) first_block_wgrad_kernel<1><<<grid, 256, smem, st>>>(x, (const bf16*)gt_bf16, gy_pooled, gw_r1, gw_l1, n, h, wd, cout, slope);
    else first_block_wgrad_kernel<3><<<grid, 256, smem, st>>>(x, (const bf16*)gt_bf16, gy_pooled, gw_r1, gw_l1, n, h, wd, cout, slope);
    return check_launch("first_block_wgrad");
}
int gim_nchw_to_nhwc(const float* x, void* y, int n, int c, int h, int wd, int dtype, gim_stream_t s) {
    int hw = h * wd;
    if (n <= 0 || c <= 0 || hw <= 0) return GIM_OK;
    GIM_REQUIRE(n <= 65535 && (c + 31) / 32 <= 65535, "nchw_to_nhwc: batch/channels too large for one launch");
    dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
    GIM_DISPATCH_DTYPE(dtype, (nchw_to_nhwc_kernel<T><<<grid, block, 0, (cudaStream_t)s>>>(x, (T*)y, c, hw)));
    return check_launch("nchw_to_nhwc");
}
int gim_nhwc_to_nchw(const void* x, float* y, int n, int c, int h, int wd, int dtype, gim_stream_t s) {
    int hw = h * wd;
    if (n <= 0 || c <= 0 || hw <= 0) return GIM_OK;
    GIM_REQUIRE(n <= 65535 && (c + 31) / 32 <= 65535, "nhwc_to_nchw: batch/channels too large for one launch");
    dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
    GIM_DISPATCH_DTYPE(dtype, (nhwc_to_nchw_kernel<T><<<grid, block, 0, (cudaStream_t)s>>>((const T*)x, y, c, hw)));
    return check_launch("nhwc_to_nchw");
}
int gim_copy_cols(const void* src, int src_ld, int src_off, void* dst, int dst_ld, int dst_off, long long rows, int c, int dtype, gim_stream_t s) {
    long long total = rows * c;
    if (total <= 0) return GIM_OK;
    GIM_DISPATCH_DTYPE(dtype, (copy_cols_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>((const T*)src, src_ld, src_off, (T*)dst, dst_ld, dst_off, rows, c)));
    return check_launch("copy_cols");
}
int gim_operand_prepare(const void* x, int dtype_in, void* out, int dtype_out, int n, int h, int wd, int c, int mode, float slope, gim_stream_t s) {
    long long total = (long long)n * h * wd * c * (mode == 2 ? 4 : 1);
    if (total <= 0) return GIM_OK;
    GIM_REQUIRE(mode >= 0 && mode <= 2, "operand_prepare: bad mode");
    bool vec = aligned16(x) && aligned16(out) && c % 8 == 0;
    int grid = ew_grid(vec ? total / 8 : total, 256, 2);
    cudaStream_t st = (cudaStream_t)s;
    if (dtype_in == GIM_F32 && dtype_out == GIM_BF16) operand_prepare_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)x, (bf16*)out, n, h, wd, c, mode, slope, vec);
    else if (dtype_in == GIM_F32 && dtype_out == GIM_F32) operand_prepare_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (float*)out, n, h, wd, c, mode, slope, vec);
    else if (dtype_in == GIM_BF16 && dtype_out == GIM_BF16) operand_prepare_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)out, n, h, wd, c, mode, slope, vec);
    else return fail(GIM_E_ARG, "operand_prepare: unsupported dtype pair");
    return check_launch("operand_prepare");
}
int gim_lrelu_bwd_ref(const float* g, const void* ref, int ref_dtype, float* gx, long long n, float slope, gim_stream_t s) {
    if (n <= 0) return GIM_OK;
    int grid = ew_grid(n, 256, 4);
    if (ref_dtype == GIM_BF16) lrelu_bwd_ref_kernel<bf16><<<grid, 256, 0, (cudaStream_t)s>>>(g, (const bf16*)ref, gx, n, slope);
    else if (ref_dtype == GIM_F32) lrelu_bwd_ref_kernel<float><<<grid, 256, 0, (cudaStream_t)s>>>(g, (const float*)ref, gx, n, slope);
    else return fail(GIM_E_ARG, "lrelu_bwd_ref: bad dtype");
    return check_launch("lrelu_bwd_ref");
}
int gim_im2col(const void* x, void* out, int n, int h, int wd, int c, int ksize, int sign, int kc, int dtype, gim_stream_t s) {
    long long total = (long long)n * h * wd * kc;
    if (total <= 0) return GIM_OK;
    GIM_REQUIRE(ksize >= 1 && (ksize & 1) && kc >= ksize * ksize * c && (sign == 1 || sign == -1), "im2col: bad arguments");
    if (dtype == GIM_BF16 && kc % 8 == 0 && aligned16(out)) {
        im2col_vec8_kernel<<<ew_grid(total / 8, 256, 1), 256, 0, (cudaStream_t)s>>>((const bf16*)x, (bf16*)out, n, h, wd, c, ksize, sign, kc);
        return check_launch("im2col_vec8");
    }
    GIM_DISPATCH_DTYPE(dtype, (im2col_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)out, n, h, wd, c, ksize, sign, kc)));
    return check_launch("im2col");
}
int gim_col2im(const float* z, const float* bias, float* y, int n, int h, int wd, int c, int ksize, int ld, gim_stream_t s) {
    long long total = (long long)n * h * wd * c;
    if (total <= 0) return GIM_OK;
    GIM_REQUIRE(ksize >= 1 && (ksize & 1) && ld >= ksize * ksize * c, "col2im: bad arguments");
    col2im_kernel<<<ew_grid(total, 256, 1), 256, 0, (cudaStream_t)s>>>(z, bias, y, n, h, wd, c, ksize, ld);
    return check_launch("col2im");
}
int gim_cast(const void* x, int dtype_in, void* y, int dtype_out, long long n, gim_stream_t s) {
    if (n <= 0) return GIM_OK;
    int grid = ew_grid(n, 256, 16);
    cudaStream_t st = (cudaStream_t)s;
    bool vec = aligned16(x) && aligned16(y);
    if (dtype_in == GIM_F32 && dtype_out == GIM_BF16) cast_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)x, (bf16*)y, n, vec);
    else if (dtype_in == GIM_BF16 && dtype_out == GIM_F32) cast_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)x, (float*)y, n, vec);
    else if (dtype_in == GIM_F32 && dtype_out == GIM_F32) cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, n, vec);
    else if (dtype_in == GIM_BF16 && dtype_out == GIM_BF16) cast_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, n, vec);
    else return fail(GIM_E_ARG, "cast: bad dtype");
    return check_launch("cast");
}

}  // extern "C"
