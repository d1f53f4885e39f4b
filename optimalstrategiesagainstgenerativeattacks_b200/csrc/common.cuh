// Shared helpers for the libgim_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/gim_b200.h"

namespace gim {

extern thread_local char g_err[512];
extern long long g_launches;
extern int g_deterministic;      // gim_set_deterministic(): every reduction output is owned by ONE CTA (no split-K / multi-CTA atomics)
inline bool deterministic() { return g_deterministic != 0; }

inline int fail(int code, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    ++g_launches;
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return GIM_E_CUDA;
    }
    return GIM_OK;
}

#define GIM_REQUIRE(cond, msg) do { if (!(cond)) return gim::fail(GIM_E_ARG, msg); } while (0)

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// runtime-dtype element access (for the mixed-dtype GEMM only)
__device__ __forceinline__ float ld_dt(const void* p, long long i, int dt) {
    return dt == GIM_BF16 ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void st_dt(void* p, long long i, int dt, float v) {
    if (dt == GIM_BF16) ((bf16*)p)[i] = __float2bfloat16_rn(v); else ((float*)p)[i] = v;
}

__device__ __forceinline__ float lrelu_f(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result broadcast to all threads; blockDim.x multiple of 32, <= 1024
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 33 floats */) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = lane < nw ? sh[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// grid for a grid-stride elementwise kernel: enough CTAs to fill the chip a few times, not more than needed
inline int ew_grid(long long n, int threads, int per_thread = 4) {
    long long need = (n + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
    long long cap = (long long)num_sms() * 16;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

#define GIM_DISPATCH_DTYPE(dtype, ...)                                            \
    do {                                                                          \
        if ((dtype) == GIM_F32) { typedef float T; __VA_ARGS__; }                 \
        else if ((dtype) == GIM_BF16) { typedef gim::bf16 T; __VA_ARGS__; }       \
        else return gim::fail(GIM_E_ARG, "bad dtype");                            \
    } while (0)

}  // namespace gim
