// Probe: how does tcgen05.mma address a K-major SWIZZLE_128B operand whose descriptor start address is NOT 1024-byte aligned and
// whose 8-row groups are NOT 1024 bytes apart?  (Needed for the halo-staged convolution: one TMA box {64 ch, bw+2, bh+2} per K block,
// the nine filter taps addressed by descriptor start offsets r*(bw+2)+q rows, SBO = (bw+2)*128 B.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/halo_probe tools/halo_probe.cu && tools/halo_probe
//
// Shared memory holds a 256-row x 64-column bf16 matrix in the layout TMA writes (row R at base + R*128, 16-byte chunk j stored at
// chunk j ^ (R & 7), base 1024-aligned).  Pass 1: A[R][c] = R, pass 2: A[R][c] = c.  B = 64x64 identity, so D[m][n] = A[row(m)][n]:
// pass 1 shows WHICH row every M index read, pass 2 whether the chunk un-swizzling matched.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
           ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(uint32_t n, uint32_t m) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

struct Cfg { int row_off, sbo_rows, base_off, mode; };

__global__ void __launch_bounds__(128, 1) probe_kernel(float* out, Cfg c) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                         // 384 rows x 128 B
    uint8_t* sb = smem + 384 * 128;             // 64 rows x 128 B (identity)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 384 * 64; i += 128) {
        const int R = i / 64, col = i % 64;
        const float v = c.mode == 0 ? (float)(R & 255) : (float)col;
        const int j = col / 8, e = col % 8;
        *reinterpret_cast<__nv_bfloat16*>(sa + R * 128 + ((j ^ (R & 7)) << 4) + e * 2) = __float2bfloat16(v);
    }
    for (int i = tid; i < 64 * 64; i += 128) {
        const int R = i / 64, col = i % 64;
        const int j = col / 8, e = col % 8;
        *reinterpret_cast<__nv_bfloat16*>(sb + R * 128 + ((j ^ (R & 7)) << 4) + e * 2) = __float2bfloat16(R == col ? 1.f : 0.f);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(64, 128);
        for (int k = 0; k < 4; ++k) {
            const uint64_t da = make_desc(smem_u32(sa) + c.row_off * 128 + k * 32, 16, c.sbo_rows * 128, c.base_off);
            const uint64_t db = make_desc(smem_u32(sb) + k * 32, 16, 1024, 0);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db),
                "r"(idesc), "r"(k ? 1u : 0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                       "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[m * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    float* d;
    cudaMalloc(&d, 128 * 64 * 4);
    float* h = (float*)malloc(128 * 64 * 4);
    float* h2 = (float*)malloc(128 * 64 * 4);
    const size_t smem = 448 * 128 + 2048;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int cfgs[][3] = {{0, 8, 0}, {8, 8, 0}, {1, 8, 0}, {1, 8, 1}, {3, 8, 0}, {3, 8, 3}, {0, 10, 0}, {11, 10, 0}, {11, 10, 3}, {22, 10, 0}, {22, 10, 6},
                           {0, 18, 0}, {19, 18, 0}, {19, 18, 3}, {5, 16, 0}, {5, 16, 5}};
    for (auto& cf : cfgs) {
        Cfg c{cf[0], cf[1], cf[2], 0};
        probe_kernel<<<1, 128, smem>>>(d, c);
        cudaMemcpy(h, d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        c.mode = 1;
        probe_kernel<<<1, 128, smem>>>(d, c);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h2, d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int rows_ok = 0, cols_ok = 0, row_uniform = 0;
        for (int m = 0; m < 128; ++m) {
            const int want = (cf[0] + (m / 8) * cf[1] + (m % 8)) & 255;
            bool uni = true, ok = true, cok = true;
            for (int n = 0; n < 64; ++n) {
                if (h[m * 64 + n] != h[m * 64]) uni = false;
                if (h[m * 64 + n] != (float)want) ok = false;
                if (h2[m * 64 + n] != (float)n) cok = false;
            }
            rows_ok += ok; cols_ok += cok; row_uniform += uni;
        }
        printf("row_off=%2d sbo_rows=%2d base_off=%d : rows as linear model %3d/128, row-uniform %3d/128, columns in order %3d/128  (%s)\n", cf[0], cf[1], cf[2],
               rows_ok, row_uniform, cols_ok, cudaGetErrorString(e));
        if (rows_ok != 128 || cols_ok != 128) {
            printf("   m: row read (pass 1, col 0 | col 8 | col 56), first cols of pass 2\n");
            for (int m = 0; m < 24; ++m)
                printf("   m=%3d want %3d : %5.0f %5.0f %5.0f | %3.0f %3.0f %3.0f %3.0f\n", m, (cf[0] + (m / 8) * cf[1] + (m % 8)) & 255, h[m * 64], h[m * 64 + 8], h[m * 64 + 56],
                       h2[m * 64], h2[m * 64 + 8], h2[m * 64 + 16], h2[m * 64 + 56]);
        }
    }
    return 0;
}
