// Tensor-core implicit-GEMM convolution for sm_100a: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) fed by TMA.
//
//   forward / input-gradient:  Y[pix][co] = sum_{tap,ci} X[pix + tap][ci] * W[tap][co][ci]
//     A operand = a 128-pixel tile of the NHWC activation, fetched per filter tap as ONE 4-D TMA box
//                 {64 ch, bw, bh, bn} at the tap-shifted coordinates -- the zero 'same' padding is TMA out-of-bounds fill,
//                 so there is no im2col buffer and no halo logic.  128 rows x 128 B, SWIZZLE_128B, K-major.
//     B operand = W[tap][n0..n0+BN][c0..c0+64], a 2-D TMA box, K-major.
//     D         = 128 x BN fp32 accumulator in tensor memory; epilogue warps read it back with tcgen05.ld, add the bias and
//                 write NHWC bf16 or fp32.
//   Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5 = epilogue
//   (warp w owns TMEM lanes 32*(w%4)..+31).  A ring of kStages smem slots with full/empty mbarriers decouples TMA from MMA.
//
//   weight-gradient:  dW[tap][co][ci] = sum_pix dY[pix][co] * X[pix + tap][ci]   (see the second half of this file)
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace gim {

// ----------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();      // a lost arrival must fail loudly, not hang the GPU
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)map), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one lane of a converged warp (the compiler keeps the surrounding warp-uniform values in uniform registers: an `if (lane == 0)`
// role loop instead makes every descriptor a vector register that has to be moved with R2UR before each UTCHMMA)
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
// arrive on an mbarrier when all previously issued MMAs have completed (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 16-byte vector reduction into global memory (sm_90+): one instruction instead of four scalar atomics
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ---- cta_group::2 (two CTAs of a cluster share one MMA: M = 2 x 128 rows, each CTA holds its rows of A, half the rows of B and its
// half of the accumulator).  Loads of either CTA signal the LEADER's (rank 0) barrier; commits are multicast to both CTAs. ----
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;           // clears the CTA-rank bit of a shared::cluster address -> same offset in CTA 0
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();                                        // role loops leave the lanes of a warp apart; .aligned needs them together
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair when all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
// arrive on the barrier at this offset in the leader CTA (rank 0) of the cluster
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0u));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// shared-memory matrix descriptor (sm_100 format, version 1), SWIZZLE_128B.
//   K-major operand : rows of 128 B (64 bf16 along K); 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major operand: 64 MN elements contiguous (128 B) per K index, 8 K-rows = one 1024 B atom (SBO), next 64-wide MN block at LBO.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return make_desc(saddr, lbo_bytes, sbo_bytes, 2);      // 2 = SWIZZLE_128B, 6 = SWIZZLE_32B
}
// instruction descriptor, kind::f16: D fp32, A/B bf16, M=128, N=n; a_mn / b_mn = 1 for MN-major operands
__host__ __device__ constexpr uint32_t make_idesc(uint32_t n, uint32_t a_mn, uint32_t b_mn, uint32_t m = 128u) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------------------------
// forward / input-gradient kernel
// ----------------------------------------------------------------------------------------------------------------
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                      // bf16 elements = one 128-byte swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KB

struct ConvTcParams {
    int n, h, w, cin, cout, ks;
    int bw, bh, bn;                              // pixel-tile box, bw*bh*bn == 128
    int tiles_w, tiles_h, tiles_n;
    int block_n;                                 // N tile (output channels per CTA)
    int stages;
    int out_f32;                                 // 1: y is fp32, 0: bf16
    int tma_store;                               // 1: epilogue stages the tile in smem (128B-swizzled) and writes it with TMA
    int block_k;                                 // K elements per stage: 64 (128-byte rows, SWIZZLE_128B) or 16 (32-byte rows, SWIZZLE_32B)
    int epi;                                     // fused epilogue bits (persistent kernel): kEpiLrelu, kEpiMask
    int m_sub;                                   // pixel tiles per macro tile (persistent kernel): 1 or 2
    int k_chains;                                // 2: one pixel tile, but its K steps alternate between two accumulators (summed by the epilogue)
    int n_staging;                               // output staging boxes of the epilogue ring (persistent kernel): 2..4
    int pair;                                    // 1: cta_group::2 kernel variant (two-CTA clusters)
    int debug;                                   // GIM_CONV_DEBUG bits (profiling experiments only): 1 no TMA store, 2 no proxy fence, 4 no smem staging
    // halo staging (persistent kernel, k > 1): ONE activation box {64 ch, bw+k-1, bn, bh+k-1} per 64-channel block serves all k*k
    // filter taps -- tap (r, q) is the same smem box read through a descriptor whose start address is advanced by
    // (r*bn*(bw+k-1) + q) rows of 128 B and whose 8-row groups are (bw+k-1) rows apart (bw == 8; tools/halo_probe.cu shows that
    // tcgen05.mma un-swizzles by absolute smem address bits, so neither needs 1024-byte alignment).  Pixel rows of the tile are
    // ordered (h, image, w): with two 8x8 images per tile the sixteen 8-pixel groups are still an arithmetic progression.
    int halo;                                    // 1: halo staging; tensor maps are then laid out {c, w, n, h}
    int halo_w;                                  // bw + k - 1: rows of 128 B between consecutive 8-pixel groups
    int halo_bytes;                              // one halo box, rounded up to 1024 B
    int halo_tx;                                 // exact bytes of one halo box (mbarrier transaction count)
    int a_stages, b_stages;                      // depths of the activation (per channel block) and weight (per tap) rings
    float slope;
};

template <int kDummy>
__global__ void __launch_bounds__(192, 2) conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                                                             const __grid_constant__ CUtensorMap map_y, const float* __restrict__ bias,
                                                             void* __restrict__ y, const ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);       // SWIZZLE_128B needs 1024-byte alignment
    const int a_bytes = kBlockM * p.block_k * 2;
    const int stage_bytes = (a_bytes + p.block_n * p.block_k * 2 + 1023) & ~1023;
    uint64_t* full_bar = (uint64_t*)(smem + p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tmem_full_bar = empty_bar + p.stages;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // tile coordinates
    int t = blockIdx.x;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    const int tn = t;
    const int w0 = tw * p.bw, h0 = th * p.bh, img0 = tn * p.bn;
    const int n0 = blockIdx.y * p.block_n;
    const int pad = (p.ks - 1) / 2;
    const int kc_per_tap = (p.cin + p.block_k - 1) / p.block_k;  // a ragged last K block is zero-filled by TMA (OOB channels)
    const int num_kb = p.ks * p.ks * kc_per_tap;
    const uint32_t tmem_cols = p.block_n < 32 ? 32u : (uint32_t)p.block_n;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t round = kb / p.stages;
                mbar_wait(&empty_bar[s], (round & 1) ^ 1);
                const int tap = kb / kc_per_tap, kc = kb - tap * kc_per_tap;
                const int r = tap / p.ks, q = tap - r * p.ks;
                uint8_t* sa = smem + s * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                mbar_expect_tx(&full_bar[s], (uint32_t)(a_bytes + p.block_n * p.block_k * 2));
                tma_load_4d(sa, &map_x, &full_bar[s], kc * p.block_k, w0 + q - pad, h0 + r - pad, img0);
                tma_load_2d(sb, &map_w, &full_bar[s], kc * p.block_k, tap * p.cout + n0);
            }
        }
    } else if (warp == 1) {
        // the whole warp walks the loop (uniform control flow keeps the descriptors in uniform registers); one elected lane issues
        const uint32_t idesc = make_idesc((uint32_t)p.block_n, 0, 0);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % p.stages;
            const uint32_t round = kb / p.stages;
            mbar_wait(&full_bar[s], round & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * stage_bytes);
            const uint32_t sb = sa + a_bytes;
            if (elect_one()) {
                if (p.block_k == 64) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = make_desc_sw128(sa + k * 32, 16, 1024);
                        const uint64_t db = make_desc_sw128(sb + k * 32, 16, 1024);
                        umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                } else {                                  // 16 K elements: rows of 32 B, 8-row groups 256 B apart, SWIZZLE_32B
                    umma_bf16(tmem_base, make_desc(sa, 16, 256, 6), make_desc(sb, 16, 256, 6), idesc, kb != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);          // frees the smem slot once these MMAs have read it
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(tmem_full_bar);              // accumulator complete
        __syncwarp();
    } else {
        // ---- epilogue: TMEM -> registers -> (+bias) -> global NHWC ----
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int quarter = warp & 3;                // TMEM lane quarter this warp may access
        const int m = quarter * 32 + lane;           // tile row = pixel within the box
        const int lw = m % p.bw, lh = (m / p.bw) % p.bh, ln = m / (p.bw * p.bh);
        const int ww = w0 + lw, hh = h0 + lh, img = img0 + ln;
        const bool valid = ww < p.w && hh < p.h && img < p.n;
        const long long pix = ((long long)img * p.h + hh) * p.w + ww;
        if (p.tma_store) {
            // Stage the tile in the (now idle) pipeline buffers as 128-byte-swizzled boxes of 128 rows x 128 B and let TMA write
            // them: coalesced full-line stores, image-border clipping by the tensor map, no per-lane address math.
            const int cols_per_box = p.out_f32 ? 32 : 64;
            for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                uint32_t v[32];
                {
                    uint32_t lo[16], hi[16];
                    tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, lo);
                    tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c0 + 16), hi);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) { v[j] = lo[j]; v[16 + j] = hi[j]; }
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + ((bias && n0 + c0 + j < p.cout) ? __ldg(&bias[n0 + c0 + j]) : 0.f);
                if (p.out_f32) {
                    uint8_t* box = smem + (c0 / 32) * (kBlockM * 128) + m * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(box + ((j ^ (m & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                } else {
                    uint8_t* box = smem + (c0 / 64) * (kBlockM * 128) + m * 128;
                    const int jb = (c0 % 64) / 8;            // first 16-byte chunk of this 32-column group inside the 64-column box
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        __nv_bfloat162 b0 = __floats2bfloat162_rn(f[8 * j + 0], f[8 * j + 1]);
                        __nv_bfloat162 b1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
                        __nv_bfloat162 b3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
                        o.x = *reinterpret_cast<uint32_t*>(&b0);
                        o.y = *reinterpret_cast<uint32_t*>(&b1);
                        o.z = *reinterpret_cast<uint32_t*>(&b2);
                        o.w = *reinterpret_cast<uint32_t*>(&b3);
                        *reinterpret_cast<uint4*>(box + (((jb + j) ^ (m & 7)) << 4)) = o;
                    }
                }
            }
            fence_proxy_async();                               // generic-proxy smem writes -> visible to the TMA (async proxy)
            asm volatile("bar.sync 1, 128;" ::: "memory");     // the four epilogue warps only
            if (warp == 2 && lane == 0) {
                const int nbox = (p.block_n + cols_per_box - 1) / cols_per_box;
                for (int b = 0; b < nbox; ++b)
                    if (n0 + b * cols_per_box < p.cout) tma_store_4d(&map_y, smem + b * (kBlockM * 128), n0 + b * cols_per_box, w0, h0, img0);
                tma_store_commit_wait();                       // smem must stay valid until TMA has read it
            }
        } else {
        for (int c0 = 0; c0 < p.block_n; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
            if (valid && n0 + c0 < p.cout) {
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + (bias ? __ldg(&bias[n0 + c0 + j]) : 0.f);
                if (p.out_f32) {
                    float4* dst = reinterpret_cast<float4*>((float*)y + pix * p.cout + n0 + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                } else {
                    uint4* dst = reinterpret_cast<uint4*>((bf16*)y + pix * p.cout + n0 + c0);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint4 o;
                        __nv_bfloat162 b0 = __floats2bfloat162_rn(f[8 * j + 0], f[8 * j + 1]);
                        __nv_bfloat162 b1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
                        __nv_bfloat162 b3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
                        o.x = *reinterpret_cast<uint32_t*>(&b0);
                        o.y = *reinterpret_cast<uint32_t*>(&b1);
                        o.z = *reinterpret_cast<uint32_t*>(&b2);
                        o.w = *reinterpret_cast<uint32_t*>(&b3);
                        dst[j] = o;
                    }
                }
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ----------------------------------------------------------------------------------------------------------------
// persistent forward / input-gradient kernel (v2)
//   One CTA per SM walks the output tiles (static round-robin).  The accumulator is double-buffered in tensor memory
//   (2 x block_n columns), so the epilogue of tile i (TMEM -> registers -> swizzled smem -> TMA store, in 16 KB chunks through two
//   rotating staging buffers) runs while the MMA warp already accumulates tile i+1; the TMA/MMA smem ring never drains between
//   tiles.  Fused epilogues: + bias, LeakyReLU, LeakyReLU-backward mask taken from a saved bf16 operand, fp32 or bf16 output.
// ----------------------------------------------------------------------------------------------------------------
enum { kEpiLrelu = 1, kEpiMask = 2, kEpiAdd = 4, kEpiPool = 8, kEpiAddUp = 16, kEpiUnit = 32 };   // kEpiUnit: scale 1 (not 1/4) in the pool / add-up epilogues
constexpr int kBoxBytes = kBlockM * 128;                   // one staged output box: 128 rows x 128 B

// A macro tile = m_sub (1 or 2) consecutive 128-pixel tiles x one block_n-wide channel tile.  With m_sub = 2 the two pixel tiles
// share every weight box and their MMAs alternate between two independent accumulators (measured: a single dependent
// accumulation chain of N=128 MMAs tops out near 1.0 PFLOP/s, two interleaved chains or N=256 reach 1.3).
template <int kPair>
__global__ void __launch_bounds__(384, 1) conv_fwd_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                                                              const __grid_constant__ CUtensorMap map_y, const float* __restrict__ bias,
                                                              const bf16* __restrict__ mask_ref, const float* addend, void* __restrict__ y,
                                                              const ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // kPair: the CTA pair computes (2 pixel tiles) x (block_n = 256 channels); this CTA stages its own pixel tile and 128 of the 256
    // weight rows, and owns the accumulator rows of its pixel tile.
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const int b_rows = kPair ? p.block_n / 2 : p.block_n;        // weight rows staged by this CTA
    const int a_bytes = kBlockM * p.block_k * 2;
    const int b_off = p.m_sub * a_bytes;
    const int stage_bytes = (b_off + b_rows * p.block_k * 2 + 1023) & ~1023;
    // halo mode: two rings -- activation halos (one per 64-channel block, shared by all taps), then weight boxes (one per tap)
    const int a_stage_bytes = p.m_sub * p.halo_bytes;
    const int b_bytes = b_rows * p.block_k * 2;
    uint8_t* ring_b = smem + p.a_stages * a_stage_bytes;
    uint8_t* staging = p.halo ? ring_b + p.b_stages * b_bytes : smem + p.stages * stage_bytes;
    float* bias_all = (float*)(staging + 2 * p.n_staging * kBoxBytes);        // 256 floats per epilogue group
    uint64_t* full_bar = (uint64_t*)(bias_all + 512);                          // halo mode: weight ring (b_stages)
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* a_full_bar = empty_bar + 8;                                      // halo mode: activation ring (a_stages <= 4)
    uint64_t* a_empty_bar = a_full_bar + 4;
    uint64_t* tmem_full_bar = a_empty_bar + 4;                                 // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;                              // [2]
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty_bar + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;      // shfl: provably warp-uniform role dispatch
    const int pad = (p.ks - 1) / 2;
    const int kc_per_tap = (p.cin + p.block_k - 1) / p.block_k;
    const int num_kb = p.ks * p.ks * kc_per_tap;
    const int n_tiles = (p.cout + p.block_n - 1) / p.block_n;
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int m_per_tile = (kPair ? 2 : 1) * p.m_sub;              // pixel tiles per (pair-)tile
    const int total_tiles = ((m_tiles + m_per_tile - 1) / m_per_tile) * n_tiles;
    const int tile0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const uint32_t buf_cols = (uint32_t)(p.m_sub * p.k_chains * p.block_n);
    const uint32_t tmem_cols = 2u * buf_cols < 32u ? 32u : 2u * buf_cols;    // power of two by construction

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_y);
        // one ring, two producers (A: warp 0, B: warp 2) arm the same barrier; halo mode: one producer per ring
        for (int s = 0; s < 8; ++s) { mbar_init(&full_bar[s], p.halo ? 1 : 2); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(&a_full_bar[s], 1); mbar_init(&a_empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], kPair ? 16 : 8); }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (kPair) tmem_alloc_2sm(tmem_slot, tmem_cols); else tmem_alloc(tmem_slot, tmem_cols);
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();      // the peer must not signal barriers that are not initialised yet
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 || warp == 2) {
        // two TMA producers share the ring: warp 0 streams the activation boxes (A), warp 2 the weight boxes (B); each arms the
        // full barrier with its own byte count
        if (lane == 0 && p.halo) {
            // halo mode.  warp 0: one halo box per (tile, 64-channel block); warp 2: one weight box per (tile, channel block, tap)
            const bool is_a = warp == 0;
            int s = 0;
            uint32_t ph = 1;
            const uint32_t tx_bytes = (is_a ? (uint32_t)(p.m_sub * p.halo_tx) : (uint32_t)b_bytes) * (kPair ? 2u : 1u);
            const int taps = p.ks * p.ks;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int mt = tile / n_tiles;
                const int n0 = (tile - mt * n_tiles) * p.block_n;
                if (is_a) {
                    int w0[2], h0[2], img0[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        int t = (kPair ? mt * 2 + (int)rank : mt) * p.m_sub + j;
                        const int tw = t % p.tiles_w; t /= p.tiles_w;
                        const int th = t % p.tiles_h; t /= p.tiles_h;
                        w0[j] = tw * p.bw - pad; h0[j] = th * p.bh - pad; img0[j] = t * p.bn;
                    }
                    for (int kc = 0; kc < p.cin; kc += kBlockK) {
                        mbar_wait(&a_empty_bar[s], ph);
                        uint8_t* sa = smem + s * a_stage_bytes;
                        if (kPair) {
                            if (rank == 0) mbar_expect_tx(&a_full_bar[s], tx_bytes);
                            tma_load_4d_2sm(sa, &map_x, &a_full_bar[s], kc, w0[0], img0[0], h0[0]);
                            if (p.m_sub == 2) tma_load_4d_2sm(sa + p.halo_bytes, &map_x, &a_full_bar[s], kc, w0[1], img0[1], h0[1]);
                        } else {
                            mbar_expect_tx(&a_full_bar[s], tx_bytes);
                            tma_load_4d(sa, &map_x, &a_full_bar[s], kc, w0[0], img0[0], h0[0]);
                            if (p.m_sub == 2) tma_load_4d(sa + p.halo_bytes, &map_x, &a_full_bar[s], kc, w0[1], img0[1], h0[1]);
                        }
                        if (++s == p.a_stages) { s = 0; ph ^= 1; }
                    }
                } else {
                    const int row0 = n0 + (int)rank * b_rows;
                    for (int kc = 0; kc < p.cin; kc += kBlockK) {
                        for (int tap = 0; tap < taps; ++tap) {
                            mbar_wait(&empty_bar[s], ph);
                            uint8_t* sb = ring_b + s * b_bytes;
                            if (kPair) {
                                if (rank == 0) mbar_expect_tx(&full_bar[s], tx_bytes);
                                tma_load_2d_2sm(sb, &map_w, &full_bar[s], kc, tap * p.cout + row0);
                            } else {
                                mbar_expect_tx(&full_bar[s], tx_bytes);
                                tma_load_2d(sb, &map_w, &full_bar[s], kc, tap * p.cout + row0);
                            }
                            if (++s == p.b_stages) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
        } else if (lane == 0) {
            const bool is_a = warp == 0;
            int s = 0;
            uint32_t ph = 1;                                  // parity to wait for on empty[s]: the first pass over the ring is free
            // pair mode: the leader arms its barrier for both CTAs' bytes; the peer's loads complete on the leader's barrier
            const uint32_t tx_bytes = (is_a ? (uint32_t)b_off : (uint32_t)(b_rows * p.block_k * 2)) * (kPair ? 2u : 1u);
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int mt = tile / n_tiles;
                const int n0 = (tile - mt * n_tiles) * p.block_n;
                int w0[2], h0[2], img0[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    int t = (kPair ? mt * 2 + (int)rank : mt) * p.m_sub + j;
                    const int tw = t % p.tiles_w; t /= p.tiles_w;
                    const int th = t % p.tiles_h; t /= p.tiles_h;
                    w0[j] = tw * p.bw - pad; h0[j] = th * p.bh - pad; img0[j] = t * p.bn;      // beyond the last tile: img0 >= n, TMA zero-fills
                }
                int tap_row = n0 + (int)rank * b_rows, kc = 0, r = 0, q = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[s], ph);
                    uint8_t* sa = smem + s * stage_bytes;
                    if (kPair) {
                        if (rank == 0) mbar_expect_tx(&full_bar[s], tx_bytes);
                        if (is_a) {
                            tma_load_4d_2sm(sa, &map_x, &full_bar[s], kc, w0[0] + q, h0[0] + r, img0[0]);
                            if (p.m_sub == 2) tma_load_4d_2sm(sa + a_bytes, &map_x, &full_bar[s], kc, w0[1] + q, h0[1] + r, img0[1]);
                        } else {
                            tma_load_2d_2sm(sa + b_off, &map_w, &full_bar[s], kc, tap_row);
                        }
                    } else {
                        mbar_expect_tx(&full_bar[s], tx_bytes);
                        if (is_a) {
                            tma_load_4d(sa, &map_x, &full_bar[s], kc, w0[0] + q, h0[0] + r, img0[0]);
                            if (p.m_sub == 2) tma_load_4d(sa + a_bytes, &map_x, &full_bar[s], kc, w0[1] + q, h0[1] + r, img0[1]);
                        } else {
                            tma_load_2d(sa + b_off, &map_w, &full_bar[s], kc, tap_row);
                        }
                    }
                    kc += p.block_k;
                    if (kc >= p.cin) { kc = 0; tap_row += p.cout; if (++q == p.ks) { q = 0; ++r; } }
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {                                 // pair mode: the leader CTA issues the MMAs of both CTAs
            // the WHOLE warp walks the loops (uniform control flow, descriptors in uniform registers); one elected lane issues
            const uint32_t idesc = make_idesc((uint32_t)p.block_n, 0, 0, kPair ? 256u : 128u);
            // descriptors differ between stages only in the 14-bit start-address field: build them once
            const bool k64 = p.block_k == 64;
            const uint32_t s0 = smem_u32(smem);
            const uint64_t desc_a0 = k64 ? make_desc_sw128(s0, 16, 1024) : make_desc(s0, 16, 256, 6);
            const uint64_t desc_b0 = k64 ? make_desc_sw128(s0 + b_off, 16, 1024) : make_desc(s0 + b_off, 16, 256, 6);
            const uint64_t stage_step = (uint64_t)(stage_bytes >> 4), a_step = (uint64_t)(a_bytes >> 4);
            const bool two = p.m_sub == 2;
            const bool chains = p.k_chains == 2;                 // K steps alternate between two accumulators (m_sub == 1, 64-wide K blocks)
            if (p.halo) {
                // halo mode: per 64-channel block one activation halo (a ring), per tap one weight box (b ring); the tap only moves the
                // start address of the A descriptor inside the halo: (r * bn * halo_w + q) rows of 128 B; 8-pixel groups halo_w rows apart
                const uint64_t desc_ha0 = make_desc_sw128(s0, 16, (uint32_t)p.halo_w * 128u);
                const uint64_t desc_hb0 = make_desc_sw128(smem_u32(ring_b), 16, 1024);
                const uint64_t a_stage_step = (uint64_t)(a_stage_bytes >> 4), halo_step = (uint64_t)(p.halo_bytes >> 4), b_step = (uint64_t)(b_bytes >> 4);
                const int row_r = p.bn * p.halo_w;
                int sa = 0, sb = 0;
                uint32_t pha = 0, phb = 0, i = 0;
                for (int tile = tile0; tile < total_tiles; tile += tile_step, ++i) {
                    const uint32_t buf = i & 1;
                    mbar_wait(&tmem_empty_bar[buf], ((i >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * buf_cols;
                    const uint32_t tmem_d1 = tmem_d + (uint32_t)p.block_n;
                    uint32_t acc = 0;
                    for (int kc = 0; kc < p.cin; kc += kBlockK) {
                        mbar_wait(&a_full_bar[sa], pha);
                        tc_fence_after();
                        const uint64_t da_stage = desc_ha0 + (uint64_t)sa * a_stage_step;
                        for (int r = 0; r < p.ks; ++r) {
                            for (int q = 0; q < p.ks; ++q) {
                                mbar_wait(&full_bar[sb], phb);
                                tc_fence_after();
                                const uint64_t da = da_stage + (uint64_t)((r * row_r + q) * 8);
                                const uint64_t db = desc_hb0 + (uint64_t)sb * b_step;
                                if (elect_one()) {
                                    if (kPair) {
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            if (chains) umma_bf16_2sm((k & 1) ? tmem_d1 : tmem_d, da + 2 * k, db + 2 * k, idesc, k > 1 ? 1u : acc);
                                            else umma_bf16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, k ? 1u : acc);
                                            if (two) umma_bf16_2sm(tmem_d1, da + halo_step + 2 * k, db + 2 * k, idesc, k ? 1u : acc);
                                        }
                                        umma_commit_2sm(&empty_bar[sb]);
                                    } else {
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            if (chains) umma_bf16((k & 1) ? tmem_d1 : tmem_d, da + 2 * k, db + 2 * k, idesc, k > 1 ? 1u : acc);
                                            else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, k ? 1u : acc);
                                            if (two) umma_bf16(tmem_d1, da + halo_step + 2 * k, db + 2 * k, idesc, k ? 1u : acc);
                                        }
                                        umma_commit(&empty_bar[sb]);
                                    }
                                }
                                __syncwarp();
                                acc = 1u;
                                if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                            }
                        }
                        if (elect_one()) {                                    // all taps have read this halo
                            if (kPair) umma_commit_2sm(&a_empty_bar[sa]); else umma_commit(&a_empty_bar[sa]);
                        }
                        __syncwarp();
                        if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
                    }
                    if (elect_one()) {
                        if (kPair) umma_commit_2sm(&tmem_full_bar[buf]); else umma_commit(&tmem_full_bar[buf]);
                    }
                    __syncwarp();
                }
            } else {
                int s = 0;
                uint32_t ph = 0, i = 0;
                uint64_t da = desc_a0, db = desc_b0;
                for (int tile = tile0; tile < total_tiles; tile += tile_step, ++i) {
                    const uint32_t buf = i & 1;
                    mbar_wait(&tmem_empty_bar[buf], ((i >> 1) & 1) ^ 1);          // epilogue has drained this accumulator pair
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * buf_cols;
                    const uint32_t tmem_d1 = tmem_d + (uint32_t)p.block_n;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&full_bar[s], ph);
                        tc_fence_after();
                        const uint32_t acc = kb != 0 ? 1u : 0u;
                        if (!elect_one()) {
                        } else if (kPair) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (chains) umma_bf16_2sm((k & 1) ? tmem_d1 : tmem_d, da + 2 * k, db + 2 * k, idesc, k > 1 ? 1u : acc);
                                else umma_bf16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, k ? 1u : acc);
                                if (two) umma_bf16_2sm(tmem_d1, da + a_step + 2 * k, db + 2 * k, idesc, k ? 1u : acc);
                            }
                            umma_commit_2sm(&empty_bar[s]);                        // frees this smem slot in both CTAs
                        } else {
                            if (k64) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {                      // +32 B along K inside the 128-byte swizzle row
                                    if (chains) umma_bf16((k & 1) ? tmem_d1 : tmem_d, da + 2 * k, db + 2 * k, idesc, k > 1 ? 1u : acc);
                                    else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, k ? 1u : acc);
                                    if (two) umma_bf16(tmem_d1, da + a_step + 2 * k, db + 2 * k, idesc, k ? 1u : acc);
                                }
                            } else {
                                umma_bf16(tmem_d, da, db, idesc, acc);
                                if (two) umma_bf16(tmem_d1, da + a_step, db, idesc, acc);
                            }
                            umma_commit(&empty_bar[s]);
                        }
                        __syncwarp();
                        da += stage_step; db += stage_step;
                        if (++s == p.stages) { s = 0; ph ^= 1; da = desc_a0; db = desc_b0; }
                    }
                    if (elect_one()) {
                        if (kPair) umma_commit_2sm(&tmem_full_bar[buf]); else umma_commit(&tmem_full_bar[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ---- epilogue: two groups of four warps (warps 4-7 and 8-11; warp & 3 = TMEM lane quarter) take alternate 16 KB chunks of
        // every tile: TMEM -> registers -> fused pointwise -> swizzled smem -> TMA store.  A lone warp per scheduler is latency bound
        // (tcgen05.ld, barrier, dependent ALU chains); the second group doubles the drain rate of the accumulators.
        const int quarter = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int m = quarter * 32 + lane;
        const int et = threadIdx.x - 128 - grp * 128;    // 0..127 inside the group
        float* bias_sm = bias_all + grp * 256;
        staging += grp * p.n_staging * kBoxBytes;
        const int bar_id = 1 + grp;
        uint32_t chunk_ctr = 0;                          // running chunk index (identical in both groups): parity selects the owner
        int bias_n0 = -1;
        // pixel of tile row m: (image, h, w) order, or (h, image, w) in halo mode (the tensor maps are then {c, w, n, h})
        const int lw = m % p.bw;
        const int lh = p.halo ? m / (p.bw * p.bn) : (m / p.bw) % p.bh;
        const int ln = p.halo ? (m / p.bw) % p.bn : m / (p.bw * p.bh);
        const int lane_dh = p.halo ? p.bw * p.bn : p.bw;      // lane distance of the pixel one row below
        const bool store_thread = (et == 0);
        const float aux_scale = (p.epi & kEpiUnit) ? 1.f : 0.25f;      // nearest-upsample / its backward (sum pooling) vs AvgPool and its backward
        const int cols_per_chunk = p.out_f32 ? 32 : 64;
        uint32_t i = 0;
        int sbuf = 0;                                    // staging box of the current chunk (ring of p.n_staging)
        for (int tile = tile0; tile < total_tiles; tile += tile_step, ++i) {
            const int mt = tile / n_tiles;
            const int n0 = (tile - mt * n_tiles) * p.block_n;
            const uint32_t buf = i & 1;
            if (n0 != bias_n0) {                         // bias tile -> smem, only when the channel tile changes
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                for (int c = et; c < p.block_n; c += 128) bias_sm[c] = (bias && n0 + c < p.cout) ? __ldg(&bias[n0 + c]) : 0.f;
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                bias_n0 = n0;
            }
            mbar_wait(&tmem_full_bar[buf], (i >> 1) & 1);
            tc_fence_after();
            for (int sub = 0; sub < p.m_sub; ++sub) {
                int t = (kPair ? mt * 2 + (int)rank : mt) * p.m_sub + sub;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h; t /= p.tiles_h;
                const int w0 = tw * p.bw, h0 = th * p.bh, img0 = t * p.bn;
                const int ww = w0 + lw, hh = h0 + lh, img = img0 + ln;
                const bool valid = ww < p.w && hh < p.h && img < p.n;
                const long long pix = ((long long)img * p.h + hh) * p.w + ww;
                const uint32_t tmem_acc = tmem_base + buf * buf_cols + (uint32_t)(sub * p.block_n) + ((uint32_t)(quarter * 32) << 16);
                for (int c0 = 0; c0 < p.block_n; c0 += cols_per_chunk) {
                    if (((chunk_ctr++) & 1) != (uint32_t)grp) continue;
                    uint8_t* box = staging + sbuf * kBoxBytes + m * 128;
#pragma unroll 1
                    for (int h32 = 0; h32 < cols_per_chunk; h32 += 32) {
                        const int cb = c0 + h32;
                        float f[32];
                        {
                            uint32_t lo[16], hi[16];
                            tmem_ld16(tmem_acc + (uint32_t)cb, lo);
                            tmem_ld16(tmem_acc + (uint32_t)(cb + 16), hi);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) { f[j] = __uint_as_float(lo[j]); f[16 + j] = __uint_as_float(hi[j]); }
                            if (p.k_chains == 2) {                   // + the second accumulation chain of the same pixel tile
                                tmem_ld16(tmem_acc + (uint32_t)(p.block_n + cb), lo);
                                tmem_ld16(tmem_acc + (uint32_t)(p.block_n + cb + 16), hi);
                                tmem_ld_wait();
#pragma unroll
                                for (int j = 0; j < 16; ++j) { f[j] += __uint_as_float(lo[j]); f[16 + j] += __uint_as_float(hi[j]); }
                            }
                        }
                        if (p.epi & kEpiPool) {
                            // AvgPool2d(2) inside the epilogue: the pixel box is at most 16 wide, so the 2x2 window of a pixel lives in
                            // lanes m^1 (w) and m^bw (h) of the same warp.  After the two butterflies each of the four lanes of a window
                            // holds the window sum of all 32 columns and writes its own quarter (8 columns) of the pooled row.
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                f[j] += __shfl_xor_sync(0xffffffffu, f[j], 1);
                                f[j] += __shfl_xor_sync(0xffffffffu, f[j], lane_dh);
                            }
                            const int sub4 = (lw & 1) | ((lh & 1) << 1);
                            float o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float v01 = (sub4 & 1) ? f[8 + e] : f[e], v23 = (sub4 & 1) ? f[24 + e] : f[16 + e];
                                o[e] = aux_scale * ((sub4 & 2) ? v23 : v01) + bias_sm[cb + 8 * sub4 + e];
                            }
                            if ((p.epi & kEpiAdd) && valid && n0 + cb < p.cout) {
                                const long long ppix = ((long long)img * (p.h >> 1) + (hh >> 1)) * (p.w >> 1) + (ww >> 1);
                                const float4* ad = reinterpret_cast<const float4*>(addend + ppix * p.cout + n0 + cb + 8 * sub4);
                                const float4 t0 = ad[0], t1 = ad[1];
                                o[0] += t0.x; o[1] += t0.y; o[2] += t0.z; o[3] += t0.w;
                                o[4] += t1.x; o[5] += t1.y; o[6] += t1.z; o[7] += t1.w;
                            }
                            const int pm = p.halo ? (((lh >> 1) * p.bn + ln) * (p.bw >> 1)) + (lw >> 1)
                                                  : ((ln * (p.bh >> 1) + (lh >> 1)) * (p.bw >> 1)) + (lw >> 1);       // pooled row inside the box: 0..31
                            uint8_t* prow = staging + sbuf * kBoxBytes + pm * 128;
                            *reinterpret_cast<float4*>(prow + (((2 * sub4) ^ (pm & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<float4*>(prow + (((2 * sub4 + 1) ^ (pm & 7)) << 4)) = make_float4(o[4], o[5], o[6], o[7]);
                            break;                                   // fp32 output: one 32-column group per chunk
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] += bias_sm[cb + j];
                        if (p.epi & kEpiLrelu) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = lrelu_f(f[j], p.slope);
                        }
                        if (p.epi & kEpiMask) {
                            if (valid && n0 + cb < p.cout) {         // cout % 32 == 0 is required for the mask epilogue
                                const uint4* mr = reinterpret_cast<const uint4*>(mask_ref + pix * p.cout + n0 + cb);
#pragma unroll
                                for (int v4 = 0; v4 < 4; ++v4) {
                                    const uint4 raw = __ldg(mr + v4);
                                    const bf16* e = reinterpret_cast<const bf16*>(&raw);
#pragma unroll
                                    for (int j = 0; j < 8; ++j)
                                        if (!(__bfloat162float(e[j]) > 0.f)) f[8 * v4 + j] *= p.slope;
                                }
                            }
                        }
                        if (p.epi & kEpiAddUp) {                  // + 0.25 * addend[n, h/2, w/2, :]: the AvgPool backward of a half-resolution gradient
                            if (valid && n0 + cb < p.cout) {
                                const long long ppix = ((long long)img * (p.h >> 1) + (hh >> 1)) * (p.w >> 1) + (ww >> 1);
                                const float4* ad = reinterpret_cast<const float4*>(addend + ppix * p.cout + n0 + cb);
#pragma unroll
                                for (int v4 = 0; v4 < 8; ++v4) {
                                    const float4 t4 = __ldg(ad + v4);
                                    f[4 * v4] += aux_scale * t4.x; f[4 * v4 + 1] += aux_scale * t4.y; f[4 * v4 + 2] += aux_scale * t4.z; f[4 * v4 + 3] += aux_scale * t4.w;
                                }
                            }
                        }
                        if (p.epi & kEpiAdd) {
                            if (valid && n0 + cb < p.cout) {
                                const float4* ad = reinterpret_cast<const float4*>(addend + pix * p.cout + n0 + cb);
#pragma unroll
                                for (int v4 = 0; v4 < 8; ++v4) {
                                    const float4 t4 = ad[v4];          // plain load: the addend may alias the output of this launch
                                    f[4 * v4] += t4.x; f[4 * v4 + 1] += t4.y; f[4 * v4 + 2] += t4.z; f[4 * v4 + 3] += t4.w;
                                }
                            }
                        }
                        if (p.debug & 4) {
                            if (f[0] == 123.456f) box[0] = 1;
                        } else if (p.out_f32) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                *reinterpret_cast<float4*>(box + ((j ^ (m & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        } else {
                            const int jb = h32 / 8;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 o;
                                __nv_bfloat162 b0 = __floats2bfloat162_rn(f[8 * j + 0], f[8 * j + 1]);
                                __nv_bfloat162 b1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
                                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
                                __nv_bfloat162 b3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
                                o.x = *reinterpret_cast<uint32_t*>(&b0);
                                o.y = *reinterpret_cast<uint32_t*>(&b1);
                                o.z = *reinterpret_cast<uint32_t*>(&b2);
                                o.w = *reinterpret_cast<uint32_t*>(&b3);
                                *reinterpret_cast<uint4*>(box + (((jb + j) ^ (m & 7)) << 4)) = o;
                            }
                        }
                        if (cb + 32 >= p.block_n) break;             // block_n == 32 with bf16 output: half a box
                    }
                    if (!(p.debug & 2)) fence_proxy_async();
                    if (store_thread) {
                        // With NB staging boxes the box written next was last read by the store issued NB-1 chunks ago: all but the
                        // newest NB-2 stores must have left smem before anyone passes the barrier below.
                        if (p.n_staging == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        else if (p.n_staging == 3) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    if (store_thread && !(p.debug & 1)) {
                        if (n0 + c0 < p.cout) {
                            const int sh = (p.epi & kEpiPool) ? 1 : 0;
                            if (p.halo) tma_store_4d(&map_y, staging + sbuf * kBoxBytes, n0 + c0, w0 >> sh, img0, h0 >> sh);
                            else tma_store_4d(&map_y, staging + sbuf * kBoxBytes, n0 + c0, w0 >> sh, h0 >> sh, img0);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (++sbuf == p.n_staging) sbuf = 0;
                }
            }
            // every TMEM read of this macro tile by this warp has completed (wait::ld): hand the accumulator buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (kPair) mbar_arrive_leader(&tmem_empty_bar[buf]);
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[buf])) : "memory");
            }
        }
        if (store_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();      // nobody may still address the peer's barriers / TMEM / smem
    if (warp == 1) {
        tc_fence_after();
        if (kPair) tmem_dealloc_2sm(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ----------------------------------------------------------------------------------------------------------------
// host side: TMA descriptors
// ----------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

static void pixel_box(int h, int w, int& bw, int& bh, int& bn) {
    bw = next_pow2(w);
    if (bw > kBlockM) bw = kBlockM;
    bh = next_pow2(h);
    if (bh > kBlockM / bw) bh = kBlockM / bw;
    bn = kBlockM / (bw * bh);
}

// 4-D map over an NHWC bf16 tensor, box {64 ch, bw, bh, bn}, 128B swizzle, zero OOB fill
// nh_order: dimensions {c, w, n, h} (image index before the row index) -- the halo-mode tile layout
static bool make_act_map(CUtensorMap* map, const void* x, int n, int h, int w, int c, int bw, int bh, int bn, int box_k = kBlockK, bool nh_order = false) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_k, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    if (nh_order) {
        dims[2] = (cuuint64_t)n; dims[3] = (cuuint64_t)h;
        strides[1] = (cuuint64_t)h * w * c * 2; strides[2] = (cuuint64_t)w * c * 2;
        box[2] = (cuuint32_t)bn; box[3] = (cuuint32_t)bh;
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               box_k == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// 4-D map over the NHWC output (fp32: box of 32 columns, bf16: 64 columns = 128 B), 128B swizzle; TMA clips at the tensor border
static bool make_out_map(CUtensorMap* map, void* y, int n, int h, int w, int c, int bw, int bh, int bn, int out_f32, bool nh_order = false) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    const cuuint64_t es = out_f32 ? 4 : 2;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c * es, (cuuint64_t)w * c * es, (cuuint64_t)h * w * c * es};
    cuuint32_t box[4] = {(cuuint32_t)(out_f32 ? 32 : 64), (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    if (nh_order) {
        dims[2] = (cuuint64_t)n; dims[3] = (cuuint64_t)h;
        strides[1] = (cuuint64_t)h * w * c * es; strides[2] = (cuuint64_t)w * c * es;
        box[2] = (cuuint32_t)bn; box[3] = (cuuint32_t)bh;
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, y, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// 2-D map over a row-major bf16 matrix [rows][cols], box {64 cols, box_rows}
static bool make_mat_map(CUtensorMap* map, const void* m, long long rows, int cols, int box_rows, int box_k = kBlockK) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(m), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               box_k == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int pick_block_n(int cout) {
    if (cout % 8 != 0) return 0;           // 16-byte rows of the packed weight / TMA-store strides
    cout = (cout + 15) & ~15;              // a ragged last N tile reads rows beyond the tap (next tap's rows, or TMA zero fill past the
    if (cout >= 128) return 128;           // end of the matrix); those columns are never stored (the output map clips them)
    if (cout > 64) return 128;
    if (cout > 32) return 64;
    return cout > 16 ? 32 : 16;
}

bool conv_tc_supported(int n, int h, int wd, int cin, int cout, int ks, int dtype) {
    if (dtype != GIM_BF16) return false;
    if (cin % 8 != 0) return false;        // TMA needs 16-byte global strides
    if (pick_block_n(cout) == 0) return false;
    if (cout % 16 != 0 && cout < 24) return false;      // the narrow-tile kernel stores whole 16-column groups
    if (ks < 1 || !(ks & 1) || ks > 15) return false;
    return n > 0 && h > 0 && wd > 0;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// N tile of the persistent kernel: 256 halves the A-operand traffic per FLOP, 128 quantises better on small layers
static int pick_block_n2(int cout, long long m_tiles) {
    static const int forced = env_int("GIM_CONV_BN", 0);
    int base = pick_block_n(cout);
    if (base < 128) return base;
    if (forced == 128 || forced == 256) return (forced == 256 && cout % 256 != 0) ? 128 : forced;
    // Measured with the warp-uniform issue loop (tools/conv_bench.py, 640 images): two interleaved M=256 x N=128 accumulation chains
    // (pair mode, two pixel tiles per CTA) beat one M=256 x N=256 chain on every layer large enough to give each CTA two pixel tiles --
    // 256->256 @16x16 1322 -> 1550, 512->512 @8x8 1270 -> 1556, 512->256 @8x8 991 -> 1243 TFLOP/s; a single N=128 chain is the slowest
    // choice (945-988), so small layers keep the 256-wide tile
    static const int prefer_two_chains = env_int("GIM_CONV_TWO_CHAINS", 1);
    if (prefer_two_chains && cout % 128 == 0 && m_tiles >= 2 * (long long)num_sms()) return 128;
    return cout % 256 == 0 ? 256 : 128;
}

// Dry run of the launcher (gim_conv2d_fwd_plan): when set, conv_fwd_tc_ex fills this record with the configuration it WOULD launch and
// returns before it touches the driver (no tensor maps, no attributes, no launch) -- so the tile picker can be swept over shapes on a
// machine without a GPU (tests/test_host_cpu.py checks shared-memory, TMEM and ring invariants for every layer family).
static thread_local int* g_plan = nullptr;
static void fill_plan(const ConvTcParams& p, int v2, long long grid_x, long long grid_y, int threads, size_t smem, int tmem_cols) {
    int* o = g_plan;
    o[0] = v2; o[1] = p.halo; o[2] = p.block_n; o[3] = p.m_sub; o[4] = p.pair; o[5] = p.stages; o[6] = p.a_stages; o[7] = p.b_stages;
    o[8] = p.k_chains; o[9] = p.bw; o[10] = p.bh; o[11] = p.bn; o[12] = (int)grid_x; o[13] = (int)grid_y; o[14] = threads;
    o[15] = (int)smem; o[16] = tmem_cols; o[17] = p.block_k; o[18] = p.tiles_w * p.tiles_h * p.tiles_n; o[19] = p.halo ? p.halo_bytes : 0;
}

int conv_fwd_tc_ex(const void* x, const void* w, const float* bias, void* y, int n, int h, int wd, int cin, int cout, int ks, int out_f32,
                   int epi, float slope, const void* mask_ref, const float* addend, cudaStream_t st) {
    static const int use_v1 = env_int("GIM_CONV_V1", 0);
    ConvTcParams p;
    p.n = n; p.h = h; p.w = wd; p.cin = cin; p.cout = cout; p.ks = ks;
    pixel_box(h, wd, p.bw, p.bh, p.bn);
    if (epi & kEpiPool) {
        if ((h & 1) || (wd & 1) || !out_f32 || (epi & (kEpiLrelu | kEpiMask)) || cout % 32 != 0)
            return fail(GIM_E_ARG, "conv_fwd_tc: the pooling epilogue needs even h and w, fp32 output, cout % 32 == 0 and no activation epilogue");
        p.bw = next_pow2(wd) < 16 ? next_pow2(wd) : 16;            // both pooling partners inside one warp
        p.bh = next_pow2(h) < kBlockM / p.bw ? next_pow2(h) : kBlockM / p.bw;
        p.bn = kBlockM / (p.bw * p.bh);
    }
    p.block_k = cin <= 16 ? 16 : kBlockK;          // skinny inputs (padded images): 32-byte K rows, one MMA per filter tap
    // halo staging (see ConvTcParams): 3x3 filters on maps of at least 8x8 pixels; tiles are 8 pixels wide so that every 8-row group of
    // the A operand is one run of 8 consecutive halo rows
    static const int halo_mode = env_int("GIM_CONV_HALO", 1), force_v2 = env_int("GIM_CONV_V2", 0);
    // launch-latency-bound problems (e.g. Linear layers) stay on the non-persistent kernel
    const long long m_tiles0 = (long long)((wd + p.bw - 1) / p.bw) * ((h + p.bh - 1) / p.bh) * ((n + p.bn - 1) / p.bn);
    const bool tiny = m_tiles0 * ((cout + 127) / 128) <= num_sms() / 2 && (epi & ~kEpiUnit) == 0;
    const bool v2 = !use_v1 && pick_block_n(cout) >= 32 && (!tiny || force_v2);
    p.halo = (halo_mode && v2 && ks == 3 && p.block_k == kBlockK && h >= 8 && wd >= 8) ? 1 : 0;
    if (p.halo) {
        p.bw = 8;
        p.bh = next_pow2(h) < 16 ? next_pow2(h) : 16;
        p.bn = kBlockM / (p.bw * p.bh);
    }
    p.halo_w = p.bw + ks - 1;
    p.halo_tx = 128 * p.halo_w * (p.bh + ks - 1) * p.bn;
    p.halo_bytes = (p.halo_tx + 1023) & ~1023;
    p.a_stages = p.b_stages = 0;
    p.tiles_w = (wd + p.bw - 1) / p.bw;
    p.tiles_h = (h + p.bh - 1) / p.bh;
    p.tiles_n = (n + p.bn - 1) / p.bn;
    p.out_f32 = out_f32;
    p.epi = epi;
    p.slope = slope;
    const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
    if (m_tiles > 2147483647LL / 64) return fail(GIM_E_ARG, "conv_fwd_tc: too many tiles");
    if ((epi & kEpiMask) && (!mask_ref || cout % 32 != 0)) return fail(GIM_E_ARG, "conv_fwd_tc: the mask epilogue needs a reference tensor and cout % 32 == 0");
    if ((epi & (kEpiAdd | kEpiAddUp)) && (!addend || cout % 32 != 0)) return fail(GIM_E_ARG, "conv_fwd_tc: the add epilogue needs an addend tensor and cout % 32 == 0");
    if ((epi & kEpiAddUp) && ((epi & (kEpiAdd | kEpiPool)) || (h & 1) || (wd & 1))) return fail(GIM_E_ARG, "conv_fwd_tc: add-upsampled needs even h, w and excludes add / pool");
    CUtensorMap map_x, map_w, map_y;
    if (!v2 && (epi & ~kEpiUnit) != 0) return fail(GIM_E_UNSUPPORTED, "conv_fwd_tc: fused epilogues need cout >= 32");
    p.block_n = v2 ? pick_block_n2(cout, m_tiles) : pick_block_n(cout);
    if (!v2) {
        // launch-latency-bound problems (Linear layers: a handful of 128-row tiles, long K): narrower N tiles put more CTAs -- and so more
        // TMA streams -- on the chip; each CTA still walks the whole K range, which is what bounds the launch
        static const int tiny_split = env_int("GIM_CONV_TINY_BN", 1);
        while (tiny_split && p.block_n > 32 && cout % (p.block_n / 2) == 0 && m_tiles * ((cout + p.block_n - 1) / p.block_n) < num_sms()) p.block_n /= 2;
    }
    static const int force_msub = env_int("GIM_CONV_MSUB", 0), pair_mode = env_int("GIM_CONV_PAIR", 2);
    p.m_sub = (v2 && p.block_n <= 128 && m_tiles >= 2 * (long long)num_sms()) ? 2 : 1;
    if (force_msub == 1 || (force_msub == 2 && v2 && p.block_n <= 128)) p.m_sub = force_msub;
    // small layers (one pixel tile per CTA): no second tile to interleave with, so the K steps themselves alternate between two
    // accumulators that the epilogue adds; needs the 128-wide tile (2 buffers x 2 chains x 128 columns = all of TMEM)
    static const int k_chains_mode = env_int("GIM_CONV_KCHAINS", 0);      // measured: no gain on the 4x4 / 2x2 layers (they are bound by wave quantisation, not by the chain): opt-in
    p.k_chains = 1;
    if (k_chains_mode && v2 && p.m_sub == 1 && p.block_k == kBlockK && cout % 128 == 0 && cout >= 128) {
        p.block_n = 128;
        p.k_chains = 2;
    }
    // cta_group::2: pairs of CTAs share one 256 x 256 MMA tile, each staging half of the weight tile.  pair_mode 2 (default since round 2)
    // also pairs the 128-column layers (two pixel tiles per CTA, M = 256 per MMA, half of the 128-row weight tile per CTA): measured
    // +7 % on the 9x9 128->128 layer, +1 % on the 3x3 one, +0.6 % on the O step (tools/conv_bench.py, bench.py A/B)
    p.pair = (v2 && pair_mode && p.block_k == 64 && m_tiles >= 2 &&
              (p.block_n == 256 || (p.block_n == 128 && (p.m_sub == 2 || p.k_chains == 2) && cout % 128 == 0 && pair_mode > 1))) ? 1 : 0;
    const int stage_bytes = ((v2 ? p.m_sub : 1) * kBlockM * p.block_k * 2 + (p.pair ? p.block_n / 2 : p.block_n) * p.block_k * 2 + 1023) & ~1023;
    if (!g_plan) {
    if (epi & kEpiPool) {
        if (!make_out_map(&map_y, y, n, h / 2, wd / 2, cout, p.bw / 2, p.bh / 2, p.bn, 1, p.halo)) return fail(GIM_E_CUDA, "conv_fwd_tc: cuTensorMapEncodeTiled(pooled y) failed");
    } else if (!make_out_map(&map_y, y, n, h, wd, cout, p.bw, p.bh, p.bn, out_f32, p.halo)) return fail(GIM_E_CUDA, "conv_fwd_tc: cuTensorMapEncodeTiled(y) failed");
    if (p.halo) {
        if (!make_act_map(&map_x, x, n, h, wd, cin, p.halo_w, p.bh + ks - 1, p.bn, p.block_k, true)) return fail(GIM_E_CUDA, "conv_fwd_tc: cuTensorMapEncodeTiled(x halo) failed");
    } else if (!make_act_map(&map_x, x, n, h, wd, cin, p.bw, p.bh, p.bn, p.block_k)) return fail(GIM_E_CUDA, "conv_fwd_tc: cuTensorMapEncodeTiled(x) failed");
    if (!make_mat_map(&map_w, w, (long long)ks * ks * cout, cin, p.pair ? p.block_n / 2 : p.block_n, p.block_k))
        return fail(GIM_E_CUDA, "conv_fwd_tc: cuTensorMapEncodeTiled(w) failed");
    }
    if (v2) {
        static const int env_nb = env_int("GIM_CONV_NSTAGING", 0), env_stages = env_int("GIM_CONV_STAGES", 0);
        p.debug = env_int("GIM_CONV_DEBUG", 0);
        p.n_staging = env_nb >= 2 && env_nb <= 4 ? env_nb : 2;
        const int fixed = 2 * p.n_staging * kBoxBytes + 512 * (int)sizeof(float) + 64 * (int)sizeof(uint64_t) + 1024;
        int stages = (227 * 1024 - fixed) / stage_bytes;
        if (stages > 8) stages = 8;
        if (env_stages >= 2 && env_stages < stages) stages = env_stages;
        p.stages = stages;
        p.tma_store = 1;
        size_t smem = (size_t)stages * stage_bytes + fixed;
        if (p.halo) {
            // two halos in flight per pixel tile (each lasts 9 taps); the rest of the shared memory is the weight ring
            static const int env_as = env_int("GIM_CONV_ASTAGES", 0);
            const int budget = 227 * 1024 - fixed, b_bytes = (p.pair ? p.block_n / 2 : p.block_n) * kBlockK * 2;
            p.a_stages = env_as >= 1 && env_as <= 4 ? env_as : 2;
            int bs = (budget - p.a_stages * p.m_sub * p.halo_bytes) / b_bytes;
            if (bs < 3) { p.a_stages = 1; bs = (budget - p.m_sub * p.halo_bytes) / b_bytes; }
            if (bs < 2) return fail(GIM_E_UNSUPPORTED, "conv_fwd_tc: halo staging does not fit in shared memory");
            p.b_stages = bs > 8 ? 8 : bs;
            if (env_stages >= 2 && env_stages < p.b_stages) p.b_stages = env_stages;
            smem = (size_t)p.a_stages * p.m_sub * p.halo_bytes + (size_t)p.b_stages * b_bytes + fixed;
        }
        if (g_plan) {                                // dry run: report the launch and stop
            const uint32_t buf_cols = (uint32_t)(p.m_sub * p.k_chains * p.block_n);
            const int tmem_cols = (int)(2u * buf_cols < 32u ? 32u : 2u * buf_cols);
            long long gx;
            if (p.pair) {
                const long long pairs = ((m_tiles + 2 * p.m_sub - 1) / (2 * p.m_sub)) * (cout / p.block_n), max_pairs = num_sms() / 2;
                gx = 2 * (pairs < max_pairs ? pairs : max_pairs);
            } else {
                const long long total = ((m_tiles + p.m_sub - 1) / p.m_sub) * ((cout + p.block_n - 1) / p.block_n);
                gx = total < num_sms() ? total : num_sms();
            }
            fill_plan(p, 1, gx, 1, 384, smem, tmem_cols);
            return GIM_OK;
        }
        static bool attr_set2 = false;
        if (!attr_set2) {
            if (cudaFuncSetAttribute(conv_fwd_tc2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
                return fail(GIM_E_CUDA, "conv_fwd_tc: cannot raise dynamic shared memory limit");
            attr_set2 = true;
        }
        if (p.pair) {
            static bool attr_set3 = false;
            if (!attr_set3) {
                if (cudaFuncSetAttribute(conv_fwd_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
                    return fail(GIM_E_CUDA, "conv_fwd_tc: cannot raise dynamic shared memory limit");
                attr_set3 = true;
            }
            const long long pairs = ((m_tiles + 2 * p.m_sub - 1) / (2 * p.m_sub)) * (cout / p.block_n);
            const long long max_pairs = num_sms() / 2;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(2 * (pairs < max_pairs ? pairs : max_pairs)));
            cfg.blockDim = dim3(384);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (cudaLaunchKernelEx(&cfg, conv_fwd_tc2_kernel<1>, map_x, map_w, map_y, bias, (const bf16*)mask_ref, addend, y, p) != cudaSuccess)
                return fail(GIM_E_CUDA, "conv_fwd_tc: cluster launch failed");
            return check_launch("conv_fwd_tc2_pair");
        }
        const long long total = ((m_tiles + p.m_sub - 1) / p.m_sub) * ((cout + p.block_n - 1) / p.block_n);
        const int grid = (int)(total < num_sms() ? total : num_sms());
        conv_fwd_tc2_kernel<0><<<grid, 384, smem, st>>>(map_x, map_w, map_y, bias, (const bf16*)mask_ref, addend, y, p);
        return check_launch("conv_fwd_tc2");
    }
    // two CTAs per SM (<= ~110 KB each): one CTA's epilogue overlaps the other's MMA main loop
    int stages = (104 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    p.stages = stages;
    // staged TMA-store epilogue when a whole 32-column group exists and the tile fits in the idle pipeline buffers
    p.tma_store = (p.block_n >= 32 && (size_t)kBlockM * p.block_n * (out_f32 ? 4 : 2) <= (size_t)stages * stage_bytes) ? 1 : 0;
    const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;
    if (g_plan) {
        fill_plan(p, 0, m_tiles, (cout + p.block_n - 1) / p.block_n, 192, smem, p.block_n < 32 ? 32 : p.block_n);
        return GIM_OK;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return fail(GIM_E_CUDA, "conv_fwd_tc: cannot raise dynamic shared memory limit");
        attr_set = true;
    }
    dim3 grid((unsigned)m_tiles, (cout + p.block_n - 1) / p.block_n);
    conv_fwd_tc_kernel<0><<<grid, 192, smem, st>>>(map_x, map_w, map_y, bias, y, p);
    return check_launch("conv_fwd_tc");
}


int conv_fwd_tc_plan(int n, int h, int wd, int cin, int cout, int ks, int out_f32, int epi, int* out20) {
    if (!out20) return fail(GIM_E_ARG, "conv_fwd_plan: null output");
    if (!conv_tc_supported(n, h, wd, cin, cout, ks, GIM_BF16)) return fail(GIM_E_UNSUPPORTED, "conv_fwd_plan: shape not eligible for the tensor-core kernels");
    for (int i = 0; i < 20; ++i) out20[i] = 0;
    g_plan = out20;
    void* dummy = reinterpret_cast<void*>(static_cast<uintptr_t>(256));      // never dereferenced in a dry run
    const int rc = conv_fwd_tc_ex(dummy, dummy, nullptr, dummy, n, h, wd, cin, cout, ks, out_f32, epi, 0.2f, dummy, (const float*)dummy, nullptr);
    g_plan = nullptr;
    return rc;
}

// ----------------------------------------------------------------------------------------------------------------
// weight gradient:  dW[tap][co][ci] = sum_pix dY[pix][co] * X[pix + tap][ci]
//   GEMM per filter tap with M = co (128), N = ci (64 or 128), K = pixels.  Both operands are "MN-major": a TMA box
//   {64 ch, bw, bh, bn} lands as 128 pixel rows of 128 B -- exactly the canonical SWIZZLE_128B MN-major layout with the pixel
//   index as K (8-pixel groups 1024 B apart = SBO; the next 64-channel block one box further = LBO).  dY at the tile
//   position, X at the tap-shifted position (zero padding and ragged tiles = TMA OOB fill; channels beyond Cout/Cin also
//   read as zeros, so any channel count that is a multiple of 8 works).  Split-K over pixel tiles across the grid; each CTA
//   accumulates its range in TMEM and adds its partial tile into the fp32 gradient with red.global.add.
// ----------------------------------------------------------------------------------------------------------------
struct WgradTcParams {
    int n, h, w, cin, cout, ks;
    int bw, bh, bn, tiles_w, tiles_h, tiles_n;
    int block_n;                 // ci per CTA: 64 or 128
    int co_tiles, ci_tiles;
    int stages;
    int tiles_per_split;         // pixel tiles per CTA
    int total_tiles;
};

// kPair = 1: cta_group::2 -- the two CTAs of a cluster own 128 output channels each (M = 256) and each stages half of the input-
// channel tile (N/2 columns of the MN-major B operand); loads of both complete on the leader's barrier, commits are multicast.
template <int kPair>
__global__ void __launch_bounds__(192, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_x,
                                                               float* __restrict__ gw, const WgradTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const int a_bytes = 2 * kATileBytes;                               // 128 co = two 64-channel boxes
    const int b_cols = kPair ? p.block_n / 2 : p.block_n;              // input channels staged by this CTA
    const int b_bytes = (b_cols / 64) * kATileBytes;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* full_bar = (uint64_t*)(smem + p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tmem_full_bar = empty_bar + p.stages;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    int t = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tap = t % (p.ks * p.ks); t /= (p.ks * p.ks);
    const int co_groups = kPair ? p.co_tiles / 2 : p.co_tiles;
    const int co_t = t % co_groups;
    const int ci_t = t / co_groups;
    const int co0 = (kPair ? co_t * 2 + (int)rank : co_t) * 128, ci0 = ci_t * p.block_n;
    const int pad = (p.ks - 1) / 2;
    const int dr = tap / p.ks - pad, dq = tap % p.ks - pad;
    const int tile_begin = blockIdx.y * p.tiles_per_split;
    int tile_end = tile_begin + p.tiles_per_split;
    if (tile_end > p.total_tiles) tile_end = p.total_tiles;
    const int num_kb = tile_end - tile_begin;                          // >= 1 by construction of the grid
    const int n_chains = 2;                                            // independent accumulation chains, summed in the epilogue (4 measured: no gain, the 128-wide tile is bound by shared-memory operand bandwidth)
    const uint32_t tmem_cols = (uint32_t)(n_chains * p.block_n);       // 256 or 512 columns

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_gy);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (kPair) tmem_alloc_2sm(tmem_slot, tmem_cols); else tmem_alloc(tmem_slot, tmem_cols);
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 1;
            int tw = tile_begin % p.tiles_w, th = (tile_begin / p.tiles_w) % p.tiles_h, tn = tile_begin / (p.tiles_w * p.tiles_h);
            const int nb = b_cols / 64;
            const int cib = ci0 + (int)rank * b_cols;                   // first input channel staged by this CTA
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[s], ph);
                const int w0 = tw * p.bw, h0 = th * p.bh, img0 = tn * p.bn;
                uint8_t* sa = smem + s * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                if (kPair) {
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2u * (uint32_t)stage_bytes);      // both CTAs' bytes land on the leader's barrier
                    tma_load_4d_2sm(sa, &map_gy, &full_bar[s], co0, w0, h0, img0);
                    tma_load_4d_2sm(sa + kATileBytes, &map_gy, &full_bar[s], co0 + 64, w0, h0, img0);
                    for (int j = 0; j < nb; ++j)
                        tma_load_4d_2sm(sb + j * kATileBytes, &map_x, &full_bar[s], cib + 64 * j, w0 + dq, h0 + dr, img0);
                } else {
                    mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                    tma_load_4d(sa, &map_gy, &full_bar[s], co0, w0, h0, img0);
                    tma_load_4d(sa + kATileBytes, &map_gy, &full_bar[s], co0 + 64, w0, h0, img0);
                    for (int j = 0; j < nb; ++j)
                        tma_load_4d(sb + j * kATileBytes, &map_x, &full_bar[s], cib + 64 * j, w0 + dq, h0 + dr, img0);
                }
                if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } }
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {                                 // whole warp, uniform control flow; one elected lane issues
            const uint32_t idesc = make_idesc((uint32_t)p.block_n, 1, 1, kPair ? 256u : 128u);
            const uint32_t s0 = smem_u32(smem);
            const uint64_t desc_a0 = make_desc_sw128(s0, kATileBytes, 1024), desc_b0 = make_desc_sw128(s0 + a_bytes, kATileBytes, 1024);
            const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);
            const uint32_t chain_mask = (uint32_t)(n_chains - 1);           // K step k accumulates into chain k & mask
            int s = 0;
            uint32_t ph = 0;
            uint64_t da = desc_a0, db = desc_b0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t acc = kb != 0 ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kBlockM / 16; ++k) {               // 128 pixels per stage = 8 MMAs of K = 16, 2 KB apart
                        const uint32_t tmem_c = tmem_base + ((uint32_t)k & chain_mask) * (uint32_t)p.block_n;
                        if (kPair) umma_bf16_2sm(tmem_c, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, k < n_chains ? acc : 1u);
                        else umma_bf16(tmem_c, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, k < n_chains ? acc : 1u);
                    }
                    if (kPair) umma_commit_2sm(&empty_bar[s]); else umma_commit(&empty_bar[s]);
                }
                __syncwarp();
                da += stage_step; db += stage_step;
                if (++s == p.stages) { s = 0; ph ^= 1; da = desc_a0; db = desc_b0; }
            }
            if (elect_one()) {
                if (kPair) umma_commit_2sm(tmem_full_bar); else umma_commit(tmem_full_bar);
            }
            __syncwarp();
        }
    } else {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int quarter = warp & 3;
        const int co = co0 + quarter * 32 + lane;
        float* dst_row = gw + ((long long)tap * p.cout + co) * p.cin + ci0;
        for (int c0 = 0; c0 < p.block_n; c0 += 16) {
            uint32_t v[16], v1[16];
            float f[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(p.block_n + c0), v1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + __uint_as_float(v1[j]);
            if (n_chains == 4) {
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(2 * p.block_n + c0), v);
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(3 * p.block_n + c0), v1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] += __uint_as_float(v[j]) + __uint_as_float(v1[j]);
            }
            if (co < p.cout) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)                      // cin % 8 == 0: whole 4-column groups are in or out
                    if (ci0 + c0 + j < p.cin) red_add_v4(dst_row + c0 + j, f[j], f[j + 1], f[j + 2], f[j + 3]);
            }
        }
    }
    tc_fence_before();
    if (kPair) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (kPair) tmem_dealloc_2sm(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
    }
}


// ----------------------------------------------------------------------------------------------------------------
// weight gradient with the gradient tile shared between filter taps (k > 1).
//   One CTA owns up to three taps of one filter row: per 64-pixel K block it fetches the dY tile ONCE and one tap-shifted X tile
//   per tap, and issues the taps' MMAs into separate TMEM accumulators (independent chains interleave on the tensor pipe).  TMA
//   traffic per FLOP drops from 2 boxes per tap to (1 + T)/T -- the kernels are bound by the ~13 TB/s of L2->SM traffic, not by HBM.
// ----------------------------------------------------------------------------------------------------------------
constexpr int kWgPix = 64;                       // pixels per K block
constexpr int kWgBox = kWgPix * 128;             // one {64 channels, 64 pixels} TMA box: 8 KB

struct WgradTc3Params {
    int n, h, w, cin, cout, ks;
    int bw, bh, bn, tiles_w, tiles_h, tiles_n;   // 64-pixel box
    int block_n;                                 // ci per CTA: 64 or 128
    int co_tiles, ci_tiles;
    int stages, tiles_per_split, total_tiles;
    int tgroup, groups_per_row;                  // taps per CTA (<= 3), tap groups per filter row
};

template <int kDummy>
__global__ void __launch_bounds__(192, 1) conv_wgrad_tc3_kernel(const __grid_constant__ CUtensorMap map_gy, const __grid_constant__ CUtensorMap map_x,
                                                                float* __restrict__ gw, const WgradTc3Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = 2 * kWgBox;                                    // 128 co = two 64-channel boxes
    const int b_tap_bytes = (p.block_n / 64) * kWgBox;
    const int stage_bytes = a_bytes + p.tgroup * b_tap_bytes;
    uint64_t* full_bar = (uint64_t*)(smem + p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tmem_full_bar = empty_bar + p.stages;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int groups = p.ks * p.groups_per_row;
    int t = blockIdx.x;
    const int g = t % groups; t /= groups;
    const int co_t = t % p.co_tiles;
    const int ci_t = t / p.co_tiles;
    const int co0 = co_t * 128, ci0 = ci_t * p.block_n;
    const int pad = (p.ks - 1) / 2;
    const int r = g / p.groups_per_row, q0 = (g - r * p.groups_per_row) * p.tgroup;
    const int nt = min(p.tgroup, p.ks - q0);                           // taps handled here: (r, q0 .. q0+nt-1)
    const int tile_begin = blockIdx.y * p.tiles_per_split;
    const int tile_end = min(p.total_tiles, tile_begin + p.tiles_per_split);
    const int num_kb = tile_end - tile_begin;                          // >= 1 by construction of the grid
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(p.tgroup * p.block_n)) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_gy);
        tma_prefetch_desc(&map_x);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 1;
            int tw = tile_begin % p.tiles_w, th = (tile_begin / p.tiles_w) % p.tiles_h, tn = tile_begin / (p.tiles_w * p.tiles_h);
            const int nb = p.block_n / 64;
            const uint32_t tx = (uint32_t)(a_bytes + nt * b_tap_bytes);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[s], ph);
                const int w0 = tw * p.bw, h0 = th * p.bh, img0 = tn * p.bn;
                uint8_t* sa = smem + s * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                mbar_expect_tx(&full_bar[s], tx);
                tma_load_4d(sa, &map_gy, &full_bar[s], co0, w0, h0, img0);
                tma_load_4d(sa + kWgBox, &map_gy, &full_bar[s], co0 + 64, w0, h0, img0);
                for (int tp = 0; tp < nt; ++tp)
                    for (int j = 0; j < nb; ++j)
                        tma_load_4d(sb + tp * b_tap_bytes + j * kWgBox, &map_x, &full_bar[s], ci0 + 64 * j, w0 + q0 + tp - pad, h0 + r - pad, img0);
                if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } }
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        {                                                // whole warp, uniform control flow; one elected lane issues
            const uint32_t idesc = make_idesc((uint32_t)p.block_n, 1, 1);
            const uint32_t s0 = smem_u32(smem);
            const uint64_t desc_a0 = make_desc_sw128(s0, kWgBox, 1024), desc_b0 = make_desc_sw128(s0 + a_bytes, kWgBox, 1024);
            const uint64_t stage_step = (uint64_t)(stage_bytes >> 4), tap_step = (uint64_t)(b_tap_bytes >> 4);
            int s = 0;
            uint32_t ph = 0;
            uint64_t da = desc_a0, db = desc_b0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t acc = kb != 0 ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kWgPix / 16; ++k)                  // 64 pixels per stage = 4 K steps of 16, 2 KB apart
                        for (int tp = 0; tp < nt; ++tp)
                            umma_bf16(tmem_base + (uint32_t)(tp * p.block_n), da + (uint64_t)(k * 128), db + tp * tap_step + (uint64_t)(k * 128), idesc,
                                      k ? 1u : acc);
                    umma_commit(&empty_bar[s]);
                }
                __syncwarp();
                da += stage_step; db += stage_step;
                if (++s == p.stages) { s = 0; ph ^= 1; da = desc_a0; db = desc_b0; }
            }
            if (elect_one()) umma_commit(tmem_full_bar);
            __syncwarp();
        }
    } else {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int quarter = warp & 3;
        const int co = co0 + quarter * 32 + lane;
        for (int tp = 0; tp < nt; ++tp) {
            const int tap = r * p.ks + q0 + tp;
            float* dst_row = gw + ((long long)tap * p.cout + co) * p.cin + ci0;
            for (int c0 = 0; c0 < p.block_n; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tp * p.block_n + c0), v);
                tmem_ld_wait();
                if (co < p.cout) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        if (ci0 + c0 + j < p.cin)
                            red_add_v4(dst_row + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// Split-K factor of the weight-gradient grids.  Every CTA owns an SM (~200 KB of shared memory), so a grid of G CTAs runs in
// ceil(G / SMs) waves of equal-length CTAs; the old rule "enough splits for ~2 waves, rounded UP" produced 297, 306 and 360 CTAs on the hot
// layers (gim_conv2d_wgrad_plan) -- a third wave for a handful of CTAs.  Pick the split count (at most two waves) that minimises
// waves x pixel tiles per CTA instead.  GIM_WGRAD_SPLIT_MODE=0 restores the old rule.
static long long pick_wgrad_splits(long long out_tiles, long long total, long long max_split) {
    static const int mode = env_int("GIM_WGRAD_SPLIT_MODE", 1);
    const long long sms = num_sms();
    long long want = (sms * 2 + out_tiles - 1) / out_tiles;             // old rule: ~2 waves, rounded up
    if (mode) {
        long long cap = (sms * 2) / out_tiles;                          // at most two waves
        if (cap < 1) cap = 1;
        if (cap > max_split) cap = max_split;
        if (cap < 1) cap = 1;
        auto cost_of = [&](long long s) {
            const long long per = (total + s - 1) / s, eff = (total + per - 1) / per;
            return ((out_tiles * eff + sms - 1) / sms) * per;
        };
        long long best_cost = cost_of(1);
        for (long long s = 2; s <= cap; ++s) best_cost = cost_of(s) < best_cost ? cost_of(s) : best_cost;
        long long best = 1;                                             // the FINEST split within 2 % of the optimum: shorter CTAs interleave
        for (long long s = 1; s <= cap; ++s)                            // better with the kernels of the other graph branches
            if (cost_of(s) * 50 <= best_cost * 51) best = s;
        want = best;
    }
    if (want > max_split) want = max_split;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    return want;
}

// Dry run of the weight-gradient launchers (gim_conv2d_wgrad_plan), see g_plan: [0] kernel (1 plain, 2 cta_group::2, 3 tap groups),
// [1] input-channel tile, [2] stages, [3] grid x, [4] grid y (split-K), [5] threads, [6] shared memory, [7] TMEM columns, [8] pixel tiles,
// [9] pixel tiles per split, [10] taps per CTA, [11..13] pixel box.
static thread_local int* g_wplan = nullptr;

static int conv_wgrad_tc3(const void* x, const void* gy, float* gw, int n, int h, int wd, int cin, int cout, int ks, cudaStream_t st) {
    WgradTc3Params p;
    p.n = n; p.h = h; p.w = wd; p.cin = cin; p.cout = cout; p.ks = ks;
    p.bw = next_pow2(wd) < kWgPix ? next_pow2(wd) : kWgPix;
    p.bh = next_pow2(h) < kWgPix / p.bw ? next_pow2(h) : kWgPix / p.bw;
    p.bn = kWgPix / (p.bw * p.bh);
    p.tiles_w = (wd + p.bw - 1) / p.bw;
    p.tiles_h = (h + p.bh - 1) / p.bh;
    p.tiles_n = (n + p.bn - 1) / p.bn;
    long long total = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
    if (total > 2147483647LL) return fail(GIM_E_ARG, "conv_wgrad_tc: too many tiles");
    p.total_tiles = (int)total;
    p.block_n = cin > 64 ? 128 : 64;
    p.co_tiles = (cout + 127) / 128;
    p.ci_tiles = (cin + p.block_n - 1) / p.block_n;
    p.tgroup = 3;
    p.groups_per_row = (ks + p.tgroup - 1) / p.tgroup;
    const int stage_bytes = 2 * kWgBox + p.tgroup * (p.block_n / 64) * kWgBox;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > 8) p.stages = 8;
    const long long out_tiles = (long long)ks * p.groups_per_row * p.co_tiles * p.ci_tiles;
    long long max_split = (total + 7) / 8;                                         // at least ~8 K blocks per CTA
    if (max_split < 1) max_split = 1;
    long long want = pick_wgrad_splits(out_tiles, total, max_split);
    if (deterministic()) want = 1;
    p.tiles_per_split = (int)((total + want - 1) / want);
    const int splits = (int)((total + p.tiles_per_split - 1) / p.tiles_per_split);
    CUtensorMap map_gy, map_x;
    const size_t smem = (size_t)p.stages * stage_bytes + (2 * p.stages + 1) * sizeof(uint64_t) + 16 + 1024;
    if (g_wplan) {
        int* o = g_wplan;
        uint32_t tc = 32;
        while (tc < (uint32_t)(p.tgroup * p.block_n)) tc <<= 1;
        o[0] = 3; o[1] = p.block_n; o[2] = p.stages; o[3] = (int)out_tiles; o[4] = splits; o[5] = 192; o[6] = (int)smem; o[7] = (int)tc;
        o[8] = p.total_tiles; o[9] = p.tiles_per_split; o[10] = p.tgroup; o[11] = p.bw; o[12] = p.bh; o[13] = p.bn;
        return GIM_OK;
    }
    if (!make_act_map(&map_gy, gy, n, h, wd, cout, p.bw, p.bh, p.bn)) return fail(GIM_E_CUDA, "conv_wgrad_tc: cuTensorMapEncodeTiled(gy) failed");
    if (!make_act_map(&map_x, x, n, h, wd, cin, p.bw, p.bh, p.bn)) return fail(GIM_E_CUDA, "conv_wgrad_tc: cuTensorMapEncodeTiled(x) failed");
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_wgrad_tc3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return fail(GIM_E_CUDA, "conv_wgrad_tc: cannot raise dynamic shared memory limit");
        attr_set = true;
    }
    if (out_tiles > 2147483647LL) return fail(GIM_E_ARG, "conv_wgrad_tc: too many output tiles");
    dim3 grid((unsigned)out_tiles, splits);
    conv_wgrad_tc3_kernel<0><<<grid, 192, smem, st>>>(map_gy, map_x, gw, p);
    return check_launch("conv_wgrad_tc3");
}

bool wgrad_tc_supported(int n, int h, int wd, int cin, int cout, int ks, int dtype) {
    if (dtype != GIM_BF16) return false;
    if (cin % 8 != 0 || cout % 8 != 0) return false;
    if (ks < 1 || !(ks & 1) || ks > 15) return false;
    return n > 0 && h > 0 && wd > 0;
}

int conv_wgrad_tc(const void* x, const void* gy, float* gw, int n, int h, int wd, int cin, int cout, int ks, cudaStream_t st) {
    // Round 1 (tools/conv_bench.py, 640 images, both with vector reductions): sharing the dY tile between three taps only won on the
    // 128-channel 3x3 layer (997 vs 921 TFLOP/s) and lost elsewhere (785 vs 943, 821 vs 1197, 9x9: 838 vs 1055).
    // Round 2, with the warp-uniform issue loop: the <= 128-channel 3x3 layers are bound by L2->SM traffic (ncu: 3.0 GB of TMA loads per
    // launch, 14.4 TB/s), there the shared dY tile wins: 935 -> 1110 TFLOP/s on 128->128 @32x32; the 9x9 layer stays (1059 vs 1018).
    static const int tap_groups = env_int("GIM_WGRAD_TAPGROUP", 2);      // 0: never, 1: every filter, 2: 3x3 layers with <= 128 output channels (also 256->128 @16x16: 874 -> 931; loses with cout >= 256, where the pair kernel runs)
    if (ks > 1 && (tap_groups == 1 || (tap_groups == 2 && ks == 3 && cout <= 128))) return conv_wgrad_tc3(x, gy, gw, n, h, wd, cin, cout, ks, st);
    WgradTcParams p;
    p.n = n; p.h = h; p.w = wd; p.cin = cin; p.cout = cout; p.ks = ks;
    pixel_box(h, wd, p.bw, p.bh, p.bn);
    p.tiles_w = (wd + p.bw - 1) / p.bw;
    p.tiles_h = (h + p.bh - 1) / p.bh;
    p.tiles_n = (n + p.bn - 1) / p.bn;
    long long total = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
    if (total > 2147483647LL) return fail(GIM_E_ARG, "conv_wgrad_tc: too many tiles");
    p.total_tiles = (int)total;
    static const int wg_pair_mode = env_int("GIM_WGRAD_PAIR", 1);
    // cta_group::2: 256 output channels per CTA pair, each CTA stages half of a 128- or 256-wide input-channel tile
    // (measured, tools/conv_bench.py: +23 % on 256->256 @16x16, +5..6 % on the 8x8 layers, but -10..28 % on 1x1 and 4x4 layers, whose few
    // large CTAs quantise badly -- so only with a filter and at least 256 pixel tiles)
    const bool pair = wg_pair_mode && cout % 256 == 0 && cin % 128 == 0 && (wg_pair_mode > 1 || (ks > 1 && total >= 256));
    p.block_n = pair ? (cin % 256 == 0 ? 256 : 128) : (cin > 64 ? 128 : 64);
    p.co_tiles = (cout + 127) / 128;
    p.ci_tiles = (cin + p.block_n - 1) / p.block_n;
    const int stage_bytes = 2 * kATileBytes + ((pair ? p.block_n / 2 : p.block_n) / 64) * kATileBytes;
    p.stages = (200 * 1024) / stage_bytes;
    const long long out_tiles = (long long)ks * ks * (pair ? p.co_tiles / 2 : p.co_tiles) * p.ci_tiles * (pair ? 2 : 1);      // CTAs along x
    long long max_split = (total + 3) / 4;                                         // at least ~4 pixel tiles per CTA
    if (max_split < 1) max_split = 1;
    long long want = pick_wgrad_splits(out_tiles, total, max_split);
    if (deterministic()) want = 1;
    p.tiles_per_split = (int)((total + want - 1) / want);
    const int splits = (int)((total + p.tiles_per_split - 1) / p.tiles_per_split);
    CUtensorMap map_gy, map_x;
    const size_t smem = (size_t)p.stages * stage_bytes + (2 * p.stages + 1) * sizeof(uint64_t) + 16 + 1024;
    if (g_wplan) {
        int* o = g_wplan;
        o[0] = pair ? 2 : 1; o[1] = p.block_n; o[2] = p.stages; o[3] = (int)out_tiles; o[4] = splits; o[5] = 192; o[6] = (int)smem; o[7] = 2 * p.block_n;
        o[8] = p.total_tiles; o[9] = p.tiles_per_split; o[10] = 1; o[11] = p.bw; o[12] = p.bh; o[13] = p.bn;
        return GIM_OK;
    }
    if (!make_act_map(&map_gy, gy, n, h, wd, cout, p.bw, p.bh, p.bn)) return fail(GIM_E_CUDA, "conv_wgrad_tc: cuTensorMapEncodeTiled(gy) failed");
    if (!make_act_map(&map_x, x, n, h, wd, cin, p.bw, p.bh, p.bn)) return fail(GIM_E_CUDA, "conv_wgrad_tc: cuTensorMapEncodeTiled(x) failed");
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv_wgrad_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return fail(GIM_E_CUDA, "conv_wgrad_tc: cannot raise dynamic shared memory limit");
        attr_set = true;
    }
    if (out_tiles > 2147483647LL) return fail(GIM_E_ARG, "conv_wgrad_tc: too many output tiles");
    dim3 grid((unsigned)out_tiles, splits);
    if (pair) {
        static bool attr_set_pair = false;
        if (!attr_set_pair) {
            if (cudaFuncSetAttribute(conv_wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
                return fail(GIM_E_CUDA, "conv_wgrad_tc: cannot raise dynamic shared memory limit");
            attr_set_pair = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, conv_wgrad_tc_kernel<1>, map_gy, map_x, gw, p) != cudaSuccess) return fail(GIM_E_CUDA, "conv_wgrad_tc: cluster launch failed");
        return check_launch("conv_wgrad_tc_pair");
    }
    conv_wgrad_tc_kernel<0><<<grid, 192, smem, st>>>(map_gy, map_x, gw, p);
    return check_launch("conv_wgrad_tc");
}

int conv_wgrad_tc_plan(int n, int h, int wd, int cin, int cout, int ks, int* out16) {
    if (!out16) return fail(GIM_E_ARG, "conv_wgrad_plan: null output");
    if (!wgrad_tc_supported(n, h, wd, cin, cout, ks, GIM_BF16)) return fail(GIM_E_UNSUPPORTED, "conv_wgrad_plan: shape not eligible for the tensor-core kernels");
    for (int i = 0; i < 16; ++i) out16[i] = 0;
    g_wplan = out16;
    void* dummy = reinterpret_cast<void*>(static_cast<uintptr_t>(256));      // never dereferenced in a dry run
    const int rc = conv_wgrad_tc(dummy, dummy, (float*)dummy, n, h, wd, cin, cout, ks, nullptr);
    g_wplan = nullptr;
    return rc;
}

}  // namespace gim
