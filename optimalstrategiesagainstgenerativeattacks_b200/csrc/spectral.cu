// Spectral normalisation of the conv weights: one power iteration per train-mode forward call, exactly the state machine of
// torch.nn.utils.spectral_norm (old-style hook) used at model_blocks.py:492-495, 522-526, 750-751, 792-793, 836-840.
// Batch independent and tiny (<= 5.3 M floats per weight): plain coalesced GEMV kernels, fp32 throughout.
#include <stdlib.h>
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


// ----------------------------------------------------------------------------------------------------------------
// Multi-layer variants: ONE launch per phase for up to kSnChunk convolutions (blockIdx.y = layer; the layer table travels as a
// by-value kernel parameter, so the launches are CUDA-graph safe), instead of eight tiny launches per layer.  The last phase also
// writes the bf16 conv operand and its flipped/transposed twin (the dgrad operand), replacing the per-call cast / flip kernels.
// ----------------------------------------------------------------------------------------------------------------
constexpr int kSnChunk = 16;
struct SnChunk {
    gim_sn_layer l[kSnChunk];
    int n;
};

// t[j] = sum_co W[co][j] * u[co]; one CTA per 32 columns walks all rows (no atomics, no memset)
__global__ void __launch_bounds__(256) sn_wtu_multi_kernel(const __grid_constant__ SnChunk c) {
    const gim_sn_layer& L = c.l[blockIdx.y];
    const int J = L.cin * L.ksize * L.ksize;
    if ((int)blockIdx.x * 32 >= J) return;
    __shared__ float sh[8][33];
    const int j = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (j < J)
        for (int co = threadIdx.y; co < L.cout; co += 8) acc = fmaf(L.w[(long long)co * J + j], L.u[co], acc);
    sh[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && j < J) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) a += sh[q][threadIdx.x];
        L.scratch[j] = a;
    }
}
__global__ void __launch_bounds__(1024) sn_vnorm_multi_kernel(const __grid_constant__ SnChunk c, float eps, int power_iter) {
    const gim_sn_layer& L = c.l[blockIdx.x];
    const int J = L.cin * L.ksize * L.ksize;
    __shared__ float sh[33];
    float* v_used = L.aux + L.cout;
    if (power_iter) {
        const float* t = L.scratch;
        float acc = 0.f;
        for (int j = threadIdx.x; j < J; j += blockDim.x) acc += t[j] * t[j];
        const float inv = 1.f / fmaxf(sqrtf(block_sum(acc, sh)), eps);
        for (int j = threadIdx.x; j < J; j += blockDim.x) {
            const float val = t[j] * inv;
            L.v[j] = val;
            v_used[j] = val;
        }
    } else {
        for (int j = threadIdx.x; j < J; j += blockDim.x) v_used[j] = L.v[j];
    }
}
__global__ void __launch_bounds__(128) sn_wv_multi_kernel(const __grid_constant__ SnChunk c) {
    const gim_sn_layer& L = c.l[blockIdx.y];
    if ((int)blockIdx.x >= L.cout) return;
    const int J = L.cin * L.ksize * L.ksize;
    __shared__ float sh[33];
    const float* row = L.w + (long long)blockIdx.x * J;
    const float* v = L.aux + L.cout;
    float acc = 0.f;
    for (int j = threadIdx.x; j < J; j += blockDim.x) acc = fmaf(row[j], v[j], acc);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) L.scratch[J + blockIdx.x] = acc;
}
__global__ void __launch_bounds__(512) sn_unorm_multi_kernel(const __grid_constant__ SnChunk c, float eps, int power_iter) {
    const gim_sn_layer& L = c.l[blockIdx.x];
    const int J = L.cin * L.ksize * L.ksize;
    __shared__ float sh[33];
    const float* s = L.scratch + J;
    float inv = 0.f;
    if (power_iter) {
        float acc = 0.f;
        for (int i = threadIdx.x; i < L.cout; i += blockDim.x) acc += s[i] * s[i];
        inv = 1.f / fmaxf(sqrtf(block_sum(acc, sh)), eps);
    }
    float dot = 0.f;
    for (int i = threadIdx.x; i < L.cout; i += blockDim.x) {
        const float ui = power_iter ? s[i] * inv : L.u[i];
        if (power_iter) L.u[i] = ui;
        L.aux[i] = ui;
        dot += ui * s[i];
    }
    dot = block_sum(dot, sh);
    if (threadIdx.x == 0) L.aux[L.cout + J] = dot;
}
// w_sn[t][co][ci] = W[co][ci][t] / sigma (fp32), the same as bf16, and the flipped/transposed bf16 pack [T-1-t][ci][co].
// One CTA per 32(co) x 32(ci) tile: the [co][ci][taps] rows are read contiguously, transposed through shared memory in groups of
// <= 9 taps, and all three outputs are written with ci- (resp. co-) contiguous rows.
constexpr int kPackTaps = 9;
constexpr int kShPitch = 32 * kPackTaps + 1;      // tile row r holds its (ci, tap) pairs in weight order: sh[r][cil * nt + tl]; odd pitch
__global__ void __launch_bounds__(256) sn_pack_multi_kernel(const __grid_constant__ SnChunk c, unsigned mask) {
    if (!((mask >> blockIdx.y) & 1u)) return;
    const gim_sn_layer& L = c.l[blockIdx.y];
    const int taps = L.ksize * L.ksize, J = L.cin * taps;
    const int ci_tiles = (L.cin + 31) / 32, co_tiles = (L.cout + 31) / 32;
    if ((int)blockIdx.x >= ci_tiles * co_tiles) return;
    const int co0 = ((int)blockIdx.x / ci_tiles) * 32, ci0 = ((int)blockIdx.x % ci_tiles) * 32;
    __shared__ float sh[32 * kShPitch];
    const float inv = 1.f / L.aux[L.cout + J];
    bf16* wop = (bf16*)L.w_op;
    bf16* wfl = (bf16*)L.w_flip;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const int nci = min(32, L.cin - ci0);
    // gridDim.z > 1: one group of <= 9 taps per CTA (a 9x9 filter has only (cout/32)*(cin/32) tiles: too few CTAs to walk 81 taps each)
    const int t_begin = gridDim.z > 1 ? (int)blockIdx.z * kPackTaps : 0, t_end = gridDim.z > 1 ? min(taps, t_begin + kPackTaps) : taps;
    for (int t0 = t_begin; t0 < t_end; t0 += kPackTaps) {
        const int nt = min(kPackTaps, taps - t0);
        // read: for each co row, the segment [ci0 .. ci0+nci) x [t0 .. t0+nt) ; consecutive threads walk (ci, t) pairs
        for (int r = ty; r < 32; r += 8) {
            const int co = co0 + r;
            if (co < L.cout) {
                const float* row = L.w + ((long long)co * L.cin + ci0) * taps;
                if (nt == taps) {
                    for (int e = tx; e < nci * nt; e += 32) sh[r * kShPitch + e] = row[e] * inv;
                } else {
                    for (int e = tx; e < nci * nt; e += 32) {
                        const int cil = e / nt, tl = e - cil * nt;
                        sh[r * kShPitch + e] = row[cil * taps + t0 + tl] * inv;
                    }
                }
            }
        }
        __syncthreads();
        for (int tl = 0; tl < nt; ++tl) {
            const int t = t0 + tl;
            for (int r = ty; r < 32; r += 8) {
                // w_sn / w_op rows: fixed (t, co), ci contiguous
                const int co = co0 + r;
                if (co < L.cout && tx < nci) {
                    const float v = sh[r * kShPitch + tx * nt + tl];
                    const long long o = ((long long)t * L.cout + co) * L.cin + ci0 + tx;
                    L.w_sn[o] = v;
                    if (wop) wop[o] = __float2bfloat16_rn(v);
                }
                // flipped rows: fixed (T-1-t, ci), co contiguous
                const int ci = ci0 + r;
                if (wfl && ci < L.cin && co0 + tx < L.cout)
                    wfl[((long long)(taps - 1 - t) * L.cin + ci) * L.cout + co0 + tx] = __float2bfloat16_rn(sh[tx * kShPitch + r * nt + tl]);
            }
        }
        __syncthreads();
    }
}

// ---- batched backward: gW[co][ci][t] (+)= G[t][co][ci]/sigma - (sum(G.W)/sigma^2) u[co] v[ci*taps+t] for up to kSnChunk layers ----
struct SnBwdChunk {
    gim_sn_bwd_layer l[kSnChunk];
    int n;
};

// phase 1: c[layer] = sum G*W.  Same tiling as phase 2 (one CTA per 32(co) x 32(ci) tile, G transposed through shared memory so that
// both G and W are read contiguously); per-CTA partial sums combine with one fp32 atomic into the layer's scratch word (zeroed by the caller)
__global__ void __launch_bounds__(256) sn_bwd_dot_multi_kernel(const __grid_constant__ SnBwdChunk c, unsigned mask) {
    if (!((mask >> blockIdx.y) & 1u)) return;
    const gim_sn_bwd_layer& L = c.l[blockIdx.y];
    const int taps = L.ksize * L.ksize;
    const int ci_tiles = (L.cin + 31) / 32, co_tiles = (L.cout + 31) / 32;
    if ((int)blockIdx.x >= ci_tiles * co_tiles) return;
    __shared__ float sh[32 * kShPitch];
    __shared__ float red[33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float acc = 0.f;
    // normally one tile per CTA; in the deterministic mode the grid has ONE CTA per layer, which walks all tiles in order
    for (int tile = blockIdx.x; tile < ci_tiles * co_tiles; tile += gridDim.x) {
    const int co0 = (tile / ci_tiles) * 32, ci0 = (tile % ci_tiles) * 32;
    const int nci = min(32, L.cin - ci0);
    const int t_begin = gridDim.z > 1 ? (int)blockIdx.z * kPackTaps : 0, t_end = gridDim.z > 1 ? min(taps, t_begin + kPackTaps) : taps;
    for (int t0 = t_begin; t0 < t_end; t0 += kPackTaps) {
        const int nt = min(kPackTaps, taps - t0);
        for (int tl = 0; tl < nt; ++tl)
            for (int r = ty; r < 32; r += 8) {
                const int co = co0 + r;
                if (co < L.cout && tx < nci) sh[r * kShPitch + tx * nt + tl] = L.g[((long long)(t0 + tl) * L.cout + co) * L.cin + ci0 + tx];
            }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int co = co0 + r;
            if (co < L.cout) {
                const float* row = L.w + ((long long)co * L.cin + ci0) * taps;
                if (nt == taps) {
                    for (int e = tx; e < nci * nt; e += 32) acc = fmaf(sh[r * kShPitch + e], row[e], acc);
                } else {
                    for (int e = tx; e < nci * nt; e += 32) {
                        const int cil = e / nt, tl = e - cil * nt;
                        acc = fmaf(sh[r * kShPitch + e], row[cil * taps + t0 + tl], acc);
                    }
                }
            }
        }
        __syncthreads();
    }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0 && acc != 0.f) atomicAdd(L.scratch, acc);
}
// phase 2: one CTA per 32(co) x 32(ci) tile, G transposed through shared memory so that both G reads and gW writes are contiguous
__global__ void __launch_bounds__(256) sn_bwd_apply_multi_kernel(const __grid_constant__ SnBwdChunk c, unsigned mask) {
    if (!((mask >> blockIdx.y) & 1u)) return;
    const gim_sn_bwd_layer& L = c.l[blockIdx.y];
    const int taps = L.ksize * L.ksize;
    const int ci_tiles = (L.cin + 31) / 32, co_tiles = (L.cout + 31) / 32;
    if ((int)blockIdx.x >= ci_tiles * co_tiles) return;
    const int co0 = ((int)blockIdx.x / ci_tiles) * 32, ci0 = ((int)blockIdx.x % ci_tiles) * 32;
    __shared__ float sh[32 * kShPitch];
    const float inv = 1.f / *L.sigma;
    const float k = (*L.scratch) * inv * inv;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int nci = min(32, L.cin - ci0);
    const int t_begin = gridDim.z > 1 ? (int)blockIdx.z * kPackTaps : 0, t_end = gridDim.z > 1 ? min(taps, t_begin + kPackTaps) : taps;
    for (int t0 = t_begin; t0 < t_end; t0 += kPackTaps) {
        const int nt = min(kPackTaps, taps - t0);
        for (int tl = 0; tl < nt; ++tl)
            for (int r = ty; r < 32; r += 8) {
                const int co = co0 + r;
                if (co < L.cout && tx < nci) sh[r * kShPitch + tx * nt + tl] = L.g[((long long)(t0 + tl) * L.cout + co) * L.cin + ci0 + tx];
            }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int co = co0 + r;
            if (co < L.cout) {
                float* row = L.grad + ((long long)co * L.cin + ci0) * taps;
                const float uco = L.u[co];
                const float* vrow = L.v + (long long)ci0 * taps;
                for (int e = tx; e < nci * nt; e += 32) {
                    int off = e;
                    if (nt != taps) {
                        const int cil = e / nt;
                        off = cil * taps + t0 + (e - cil * nt);
                    }
                    const float val = sh[r * kShPitch + e] * inv - k * uco * vrow[off];
                    row[off] = L.accumulate ? row[off] + val : val;
                }
            }
        }
        __syncthreads();
    }
}


// ---- 16-byte variants of the three tile kernels (cin % 4 == 0 and all taps of the filter in one pass, i.e. k <= 3) ----
// The scalar kernels above keep 4-byte loads with ~4 in flight per warp: measured 0.28-0.31 of the HBM roofline.  Here every global
// access is a float4 (bf16: 8 bytes) and each thread has up to nine independent loads in flight.  Tile = 32(co) x 32(ci) x taps, held
// in shared memory in weight order sh[co][ci * taps + t].
//   packed side  G / w_sn [t][co][ci] : a row of 32 ci = 8 float4; thread (r = tid >> 3, q = tid & 7) walks the taps
//   weight side  W / grad [co][ci][t] : a row of 32 ci x taps contiguous floats = 8 * taps float4
__device__ __forceinline__ void sn_tile_load_packed(const float* __restrict__ g, float* sh, int cout, int cin, int taps, int co0, int ci0, int nci) {
    const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
    const int co = co0 + r;
    if (co < cout && 4 * q < nci) {
        float4 v[kPackTaps];
#pragma unroll
        for (int t = 0; t < kPackTaps; ++t)
            if (t < taps) v[t] = __ldg(reinterpret_cast<const float4*>(g + ((long long)t * cout + co) * cin + ci0 + 4 * q));
#pragma unroll
        for (int t = 0; t < kPackTaps; ++t)
            if (t < taps) {
                float* d = sh + r * kShPitch + (4 * q) * taps + t;
                d[0] = v[t].x; d[taps] = v[t].y; d[2 * taps] = v[t].z; d[3 * taps] = v[t].w;
            }
    }
}

__global__ void __launch_bounds__(256) sn_pack_multi_vec_kernel(const __grid_constant__ SnChunk c, unsigned mask) {
    if (!((mask >> blockIdx.y) & 1u)) return;
    const gim_sn_layer& L = c.l[blockIdx.y];
    const int taps = L.ksize * L.ksize, J = L.cin * taps;
    const int ci_tiles = (L.cin + 31) / 32, co_tiles = (L.cout + 31) / 32;
    if ((int)blockIdx.x >= ci_tiles * co_tiles) return;
    const int co0 = ((int)blockIdx.x / ci_tiles) * 32, ci0 = ((int)blockIdx.x % ci_tiles) * 32;
    __shared__ float sh[32 * kShPitch];
    const float inv = 1.f / L.aux[L.cout + J];
    bf16* wop = (bf16*)L.w_op;
    bf16* wfl = (bf16*)L.w_flip;
    const int nci = min(32, L.cin - ci0), nco = min(32, L.cout - co0);
    const int row4 = nci * taps / 4;                               // float4 per weight row segment
    // weight rows -> shared memory (scaled)
    for (int f = threadIdx.x; f < nco * row4; f += 256) {
        const int r = f / row4, c4 = f - r * row4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(L.w + ((long long)(co0 + r) * L.cin + ci0) * taps) + c4);
        float* d = sh + r * kShPitch + 4 * c4;
        d[0] = v.x * inv; d[1] = v.y * inv; d[2] = v.z * inv; d[3] = v.w * inv;
    }
    __syncthreads();
    // packed rows (fixed tap and co, ci contiguous): fp32 + bf16
    {
        const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
        if (r < nco && 4 * q < nci) {
#pragma unroll
            for (int t = 0; t < kPackTaps; ++t)
                if (t < taps) {
                    const float* sp = sh + r * kShPitch + (4 * q) * taps + t;
                    const float4 v = make_float4(sp[0], sp[taps], sp[2 * taps], sp[3 * taps]);
                    const long long o = ((long long)t * L.cout + co0 + r) * L.cin + ci0 + 4 * q;
                    *reinterpret_cast<float4*>(L.w_sn + o) = v;
                    if (wop) {
                        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
                        *reinterpret_cast<uint2*>(wop + o) = pk;
                    }
                }
        }
    }
    // flipped rows (fixed T-1-t and ci, co contiguous): bf16; thread (ci row = tid >> 3, q) writes 4 consecutive co
    if (wfl) {
        const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
        if (r < nci && 4 * q < nco) {
            const bool full = (L.cout % 4 == 0);
#pragma unroll
            for (int t = 0; t < kPackTaps; ++t)
                if (t < taps) {
                    const float* sp = sh + (4 * q) * kShPitch + r * taps + t;
                    const long long o = ((long long)(taps - 1 - t) * L.cin + ci0 + r) * L.cout + co0 + 4 * q;
                    if (full) {
                        __nv_bfloat162 lo = __floats2bfloat162_rn(sp[0], sp[kShPitch]), hi = __floats2bfloat162_rn(sp[2 * kShPitch], sp[3 * kShPitch]);
                        uint2 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&lo); pk.y = *reinterpret_cast<uint32_t*>(&hi);
                        *reinterpret_cast<uint2*>(wfl + o) = pk;
                    } else {
                        for (int j = 0; j < 4 && 4 * q + j < nco; ++j) wfl[o + j] = __float2bfloat16_rn(sp[j * kShPitch]);
                    }
                }
        }
    }
}

__global__ void __launch_bounds__(256) sn_bwd_dot_multi_vec_kernel(const __grid_constant__ SnBwdChunk c, unsigned mask) {
    if (!((mask >> blockIdx.y) & 1u)) return;
    const gim_sn_bwd_layer& L = c.l[blockIdx.y];
    const int taps = L.ksize * L.ksize;
    const int ci_tiles = (L.cin + 31) / 32, co_tiles = (L.cout + 31) / 32;
    if ((int)blockIdx.x >= ci_tiles * co_tiles) return;
    __shared__ float sh[32 * kShPitch];
    __shared__ float red[33];
    float acc = 0.f;
    // normally one tile per CTA; in the deterministic mode the grid has ONE CTA per layer, which walks all tiles in order
    for (int tile = blockIdx.x; tile < ci_tiles * co_tiles; tile += gridDim.x) {
        const int co0 = (tile / ci_tiles) * 32, ci0 = (tile % ci_tiles) * 32;
        const int nci = min(32, L.cin - ci0), nco = min(32, L.cout - co0);
        const int row4 = nci * taps / 4;
        sn_tile_load_packed(L.g, sh, L.cout, L.cin, taps, co0, ci0, nci);
        __syncthreads();
        for (int f = threadIdx.x; f < nco * row4; f += 256) {
            const int r = f / row4, c4 = f - r * row4;
            const float4 w = __ldg(reinterpret_cast<const float4*>(L.w + ((long long)(co0 + r) * L.cin + ci0) * taps) + c4);
            const float* gp = sh + r * kShPitch + 4 * c4;
            acc = fmaf(gp[0], w.x, acc); acc = fmaf(gp[1], w.y, acc); acc = fmaf(gp[2], w.z, acc); acc = fmaf(gp[3], w.w, acc);
        }
        __syncthreads();
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0 && acc != 0.f) atomicAdd(L.scratch, acc);
}

__global__ void __launch_bounds__(256) sn_bwd_apply_multi_vec_kernel(const __grid_constant__ SnBwdChunk c, unsigned mask) {
    if (!((mask >> blockIdx.y) & 1u)) return;
    const gim_sn_bwd_layer& L = c.l[blockIdx.y];
    const int taps = L.ksize * L.ksize;
    const int ci_tiles = (L.cin + 31) / 32, co_tiles = (L.cout + 31) / 32;
    if ((int)blockIdx.x >= ci_tiles * co_tiles) return;
    const int co0 = ((int)blockIdx.x / ci_tiles) * 32, ci0 = ((int)blockIdx.x % ci_tiles) * 32;
    __shared__ float sh[32 * kShPitch];
    const float inv = 1.f / *L.sigma;
    const float k = (*L.scratch) * inv * inv;
    const int nci = min(32, L.cin - ci0), nco = min(32, L.cout - co0);
    const int row4 = nci * taps / 4;
    sn_tile_load_packed(L.g, sh, L.cout, L.cin, taps, co0, ci0, nci);
    __syncthreads();
    const float4* vrow = reinterpret_cast<const float4*>(L.v + (long long)ci0 * taps);
    for (int f = threadIdx.x; f < nco * row4; f += 256) {
        const int r = f / row4, c4 = f - r * row4;
        float4* dst = reinterpret_cast<float4*>(L.grad + ((long long)(co0 + r) * L.cin + ci0) * taps) + c4;
        const float ku = k * L.u[co0 + r];
        const float4 vv = __ldg(vrow + c4);
        const float* gp = sh + r * kShPitch + 4 * c4;
        float4 o = make_float4(gp[0] * inv - ku * vv.x, gp[1] * inv - ku * vv.y, gp[2] * inv - ku * vv.z, gp[3] * inv - ku * vv.w);
        if (L.accumulate) {
            const float4 old = *dst;
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *dst = o;
    }
}

}  // namespace gim

using namespace gim;

static bool sn_vec_enabled() {
    static const bool on = !(getenv("GIM_SN_SCALAR") && atoi(getenv("GIM_SN_SCALAR")));
    return on;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }      // NULL counts as aligned (optional outputs)

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
        if (splits < 1 || deterministic()) splits = 1;
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
    int grid = deterministic() ? 1 : ew_grid(total, 256, 8);
    sn_bwd_dot_kernel<<<grid, 256, 0, st>>>(g_w_sn, weight_orig, scratch, cout, cin, taps);
    int rc = check_launch("sn_bwd_dot");
    if (rc != GIM_OK) return rc;
    sn_bwd_apply_kernel<<<ew_grid(total, 256), 256, 0, st>>>(g_w_sn, u_used, v_used, sigma, scratch, g_weight_orig, cout, cin, taps);
    return check_launch("sn_bwd_apply");
}

int gim_sn_forward_multi(const gim_sn_layer* layers, int n_layers, int power_iter, float eps, gim_stream_t s) {
    GIM_REQUIRE(n_layers >= 0 && (n_layers == 0 || layers), "sn_forward_multi: bad arguments");
    cudaStream_t st = (cudaStream_t)s;
    for (int base = 0; base < n_layers; base += kSnChunk) {
        SnChunk c;
        c.n = n_layers - base < kSnChunk ? n_layers - base : kSnChunk;
        int max_j = 0, max_co = 0;
        for (int i = 0; i < c.n; ++i) {
            c.l[i] = layers[base + i];
            GIM_REQUIRE(c.l[i].cout > 0 && c.l[i].cin > 0 && c.l[i].ksize > 0 && c.l[i].w && c.l[i].u && c.l[i].v && c.l[i].w_sn && c.l[i].aux && c.l[i].scratch,
                        "sn_forward_multi: bad layer");
            int j = c.l[i].cin * c.l[i].ksize * c.l[i].ksize;
            if (j > max_j) max_j = j;
            if (c.l[i].cout > max_co) max_co = c.l[i].cout;
        }
        for (int i = c.n; i < kSnChunk; ++i) c.l[i] = c.l[0];
        int rc;
        if (power_iter) {
            sn_wtu_multi_kernel<<<dim3((max_j + 31) / 32, c.n), dim3(32, 8), 0, st>>>(c);
            if ((rc = check_launch("sn_wtu_multi")) != GIM_OK) return rc;
        }
        sn_vnorm_multi_kernel<<<c.n, 1024, 0, st>>>(c, eps, power_iter);
        if ((rc = check_launch("sn_vnorm_multi")) != GIM_OK) return rc;
        sn_wv_multi_kernel<<<dim3(max_co, c.n), 128, 0, st>>>(c);
        if ((rc = check_launch("sn_wv_multi")) != GIM_OK) return rc;
        sn_unorm_multi_kernel<<<c.n, 512, 0, st>>>(c, eps, power_iter);
        if ((rc = check_launch("sn_unorm_multi")) != GIM_OK) return rc;
        int max_tiles = 0;
        for (int i = 0; i < c.n; ++i) {
            int tl = ((c.l[i].cin + 31) / 32) * ((c.l[i].cout + 31) / 32);
            if (tl > max_tiles) max_tiles = tl;
        }
        // 16-byte kernels for the layers with cin % 4 == 0, all taps in one pass (k <= 3) and 16-byte aligned tensors; scalar kernels for the rest
        unsigned vmask = 0;
        for (int i = 0; i < c.n && sn_vec_enabled(); ++i) {
            const gim_sn_layer& L = c.l[i];
            if (L.cin % 4 == 0 && L.ksize * L.ksize <= kPackTaps && aligned16(L.w) && aligned16(L.w_sn) && aligned16(L.w_op) && aligned16(L.w_flip)) vmask |= 1u << i;
        }
        const unsigned all = c.n >= 32 ? 0xffffffffu : ((1u << c.n) - 1u);
        if (vmask) sn_pack_multi_vec_kernel<<<dim3(max_tiles, c.n), 256, 0, st>>>(c, vmask);
        int zmax = 1;                                    // tap groups of the widest scalar-path filter (81 taps -> 9 CTAs per tile)
        for (int i = 0; i < c.n; ++i)
            if (!((vmask >> i) & 1u)) zmax = max(zmax, (c.l[i].ksize * c.l[i].ksize + kPackTaps - 1) / kPackTaps);
        if (vmask != all) sn_pack_multi_kernel<<<dim3(max_tiles, c.n, zmax), 256, 0, st>>>(c, all & ~vmask);
        if ((rc = check_launch("sn_pack_multi")) != GIM_OK) return rc;
    }
    return GIM_OK;
}

int gim_sn_backward_multi(const gim_sn_bwd_layer* layers, int n_layers, gim_stream_t s) {
    GIM_REQUIRE(n_layers >= 0 && (n_layers == 0 || layers), "sn_backward_multi: bad arguments");
    cudaStream_t st = (cudaStream_t)s;
    for (int base = 0; base < n_layers; base += kSnChunk) {
        SnBwdChunk c;
        c.n = n_layers - base < kSnChunk ? n_layers - base : kSnChunk;
        int max_tiles = 0;
        for (int i = 0; i < c.n; ++i) {
            c.l[i] = layers[base + i];
            const gim_sn_bwd_layer& L = c.l[i];
            GIM_REQUIRE(L.cout > 0 && L.cin > 0 && L.ksize > 0 && L.g && L.w && L.u && L.v && L.sigma && L.grad && L.scratch, "sn_backward_multi: bad layer");
            if (cudaMemsetAsync(L.scratch, 0, sizeof(float), st) != cudaSuccess) return fail(GIM_E_CUDA, "sn_backward_multi memset");
            int tl = ((L.cin + 31) / 32) * ((L.cout + 31) / 32);
            if (tl > max_tiles) max_tiles = tl;
        }
        for (int i = c.n; i < kSnChunk; ++i) c.l[i] = c.l[0];
        unsigned vmask = 0;
        for (int i = 0; i < c.n && sn_vec_enabled(); ++i) {
            const gim_sn_bwd_layer& L = c.l[i];
            if (L.cin % 4 == 0 && L.ksize * L.ksize <= kPackTaps && aligned16(L.g) && aligned16(L.w) && aligned16(L.grad) && aligned16(L.v)) vmask |= 1u << i;
        }
        const unsigned all = c.n >= 32 ? 0xffffffffu : ((1u << c.n) - 1u);
        const int dot_x = deterministic() ? 1 : max_tiles;
        if (vmask) sn_bwd_dot_multi_vec_kernel<<<dim3(dot_x, c.n), 256, 0, st>>>(c, vmask);
        int zmax = 1;
        for (int i = 0; i < c.n; ++i)
            if (!((vmask >> i) & 1u)) zmax = max(zmax, (c.l[i].ksize * c.l[i].ksize + kPackTaps - 1) / kPackTaps);
        if (vmask != all) sn_bwd_dot_multi_kernel<<<dim3(dot_x, c.n, deterministic() ? 1 : zmax), 256, 0, st>>>(c, all & ~vmask);
        int rc = check_launch("sn_bwd_dot_multi");
        if (rc != GIM_OK) return rc;
        if (vmask) sn_bwd_apply_multi_vec_kernel<<<dim3(max_tiles, c.n), 256, 0, st>>>(c, vmask);
        if (vmask != all) sn_bwd_apply_multi_kernel<<<dim3(max_tiles, c.n, zmax), 256, 0, st>>>(c, all & ~vmask);
        if ((rc = check_launch("sn_bwd_apply_multi")) != GIM_OK) return rc;
    }
    return GIM_OK;
}

}  // extern "C"
