// Fused self-attention over an 8x8 feature map (64 positions): the SelfAttention block of the GIM pyramids
// (reference models/model_blocks.py:517-549) between its three 1x1 convolutions and its output:
//     S[j,i] = <q_j, k_i>,  A = softmax_i(S),  y[j,:] = gamma * sum_i A[j,i] v[i,:] + x[j,:]
// One CTA per image: q, k, v and the 64x64 attention map live in shared memory, so a forward pass reads q, k, v, x once and
// writes y (+ A for the backward); the backward reads dy, q, k, v, A once and writes dq, dk, dv (+ one partial of dgamma).
// The unfused route is 2 strided GEMM launches + softmax + scale + add forward and 4 GEMMs + softmax-backward backward, each
// a full pass over HBM.  Arithmetic is plain fp32 FFMA (K = 32..256 per dot product: not worth a tensor-core pipeline, and it
// keeps the fp32 parity path and the bf16 path on the same kernel).
#include "common.cuh"

namespace gim {

constexpr int kAttP = 64;            // positions
constexpr int kAttAP = kAttP + 4;    // pitch of the 64x64 maps in smem (floats)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// rows x (4*row_f4) fp32 matrix with `ld` floats between rows in global memory -> smem with `pitch` floats per row (both % 4 == 0)
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int rows, int row_f4, int pitch, int ld) {
    for (int e = threadIdx.x; e < rows * row_f4; e += blockDim.x) {
        int r = e / row_f4, c4 = e - r * row_f4;
        cp_async16(dst + r * pitch + c4 * 4, src + (size_t)r * ld + c4 * 4);
    }
}

template <int C> struct AttnSmem {
    static constexpr int D = C / 8, DP = D + 4, CP = C + 4;
    static constexpr int fwd_floats = kAttP * C + kAttP * kAttAP + 2 * kAttP * DP;
    static constexpr int bwd_floats = 2 * kAttP * CP + 2 * kAttP * kAttAP + 2 * kAttP * DP + 64;
};

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256, 2) attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                       int ld_qk, int ld_v, const float* __restrict__ x, const float* __restrict__ gamma,
                                                       float* __restrict__ attn, float* __restrict__ y) {
    constexpr int P = kAttP, AP = kAttAP, D = C / 8, DP = D + 4, CPL = C / 128;
    extern __shared__ __align__(16) float att_smem[];
    float* v_s = att_smem;             // [P][C]      values, rows contiguous
    float* at_s = v_s + P * C;         // [P][AP]     attention map transposed: at_s[i][j] = A[j][i]
    float* q_s = at_s + P * AP;        // [P][DP]
    float* k_s = q_s + P * DP;         // [P][DP]
    const int img = blockIdx.x, tid = threadIdx.x;

    stage_rows(q_s, q + (size_t)img * P * ld_qk, P, D / 4, DP, ld_qk);
    stage_rows(k_s, k + (size_t)img * P * ld_qk, P, D / 4, DP, ld_qk);
    cp_async_wait_all();
    stage_rows(v_s, v + (size_t)img * P * ld_v, P, C / 4, C, ld_v);          // lands while the scores are computed
    asm volatile("cp.async.commit_group;" ::: "memory");
    __syncthreads();

    // scores + softmax: 4 lanes per query row j, keys i = il + 4t
    {
        const int j = tid >> 2, il = tid & 3;
        float s[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) s[t] = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < D / 4; ++d4) {
            const float4 qv = *reinterpret_cast<const float4*>(q_s + j * DP + d4 * 4);
#pragma unroll
            for (int t = 0; t < 16; ++t) s[t] += dot4(qv, *reinterpret_cast<const float4*>(k_s + (il + 4 * t) * DP + d4 * 4));
        }
        float mx = s[0];
#pragma unroll
        for (int t = 1; t < 16; ++t) mx = fmaxf(mx, s[t]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            s[t] = expf(s[t] - mx);
            sum += s[t];
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float inv = 1.f / sum;
        float* arow = attn + ((size_t)img * P + j) * P;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const float a = s[t] * inv;
            arow[il + 4 * t] = a;
            at_s[(il + 4 * t) * AP + j] = a;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // y = gamma * A v + x: warp w owns query rows 8w..8w+7, lane l the channels 4l..4l+3 (+128)
    const int w = tid >> 5, l = tid & 31, j0 = 8 * w;
    float acc[8][4 * CPL];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4 * CPL; ++c) acc[r][c] = 0.f;
#pragma unroll 4
    for (int i = 0; i < P; ++i) {
        const float4 a0 = *reinterpret_cast<const float4*>(at_s + i * AP + j0);
        const float4 a1 = *reinterpret_cast<const float4*>(at_s + i * AP + j0 + 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int p = 0; p < CPL; ++p) {
            const float4 vv = *reinterpret_cast<const float4*>(v_s + i * C + p * 128 + 4 * l);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                acc[r][4 * p + 0] += a[r] * vv.x;
                acc[r][4 * p + 1] += a[r] * vv.y;
                acc[r][4 * p + 2] += a[r] * vv.z;
                acc[r][4 * p + 3] += a[r] * vv.w;
            }
        }
    }
    const float gm = __ldg(gamma);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int p = 0; p < CPL; ++p) {
            const size_t off = ((size_t)img * P + j0 + r) * C + p * 128 + 4 * l;
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + off));
            float4 o;
            o.x = gm * acc[r][4 * p + 0] + xv.x;
            o.y = gm * acc[r][4 * p + 1] + xv.y;
            o.z = gm * acc[r][4 * p + 2] + xv.z;
            o.w = gm * acc[r][4 * p + 3] + xv.w;
            *reinterpret_cast<float4*>(y + off) = o;
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward:  T = dy v^T;  dgamma = sum(A.T);  dS = A.(gamma T - rowsum(gamma T.A));  dv = gamma A^T dy;  dq = dS k;  dk = dS^T q
// (dx = dy is returned by the caller)
// ---------------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) attn_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ q, const float* __restrict__ k,
                                                       const float* __restrict__ v, int ld_qk, int ld_v, const float* __restrict__ attn,
                                                       const float* __restrict__ gamma, float* __restrict__ dq, float* __restrict__ dk,
                                                       float* __restrict__ dv, float* __restrict__ dgamma_part) {
    constexpr int P = kAttP, AP = kAttAP, D = C / 8, DP = D + 4, CP = C + 4, CPL = C / 128, NDG = D / 4;
    extern __shared__ __align__(16) float att_smem[];
    float* g_s = att_smem;             // [P][CP]   dy
    float* v_s = g_s + P * CP;         // [P][CP]
    float* a_s = v_s + P * CP;         // [P][AP]   A[j][i]
    float* t_s = a_s + P * AP;         // [P][AP]   T, then dS
    float* q_s = t_s + P * AP;         // [P][DP]
    float* k_s = q_s + P * DP;         // [P][DP]
    float* red = k_s + P * DP;         // 64 floats of reduction scratch
    const int img = blockIdx.x, tid = threadIdx.x;

    stage_rows(g_s, gy + (size_t)img * P * C, P, C / 4, CP, C);
    stage_rows(v_s, v + (size_t)img * P * ld_v, P, C / 4, CP, ld_v);
    stage_rows(a_s, attn + (size_t)img * P * P, P, P / 4, AP, P);
    stage_rows(q_s, q + (size_t)img * P * ld_qk, P, D / 4, DP, ld_qk);
    stage_rows(k_s, k + (size_t)img * P * ld_qk, P, D / 4, DP, ld_qk);
    cp_async_wait_all();
    __syncthreads();

    // ---- T[j][i] = sum_c dy[j][c] v[i][c]: four thread groups split the channel range, 8x8 outputs per thread (rows jg+8r, cols ig+8r')
    {
        const int grp = tid >> 6, tt = tid & 63, jg = tt >> 3, ig = tt & 7;
        float acc[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
        constexpr int F4_PER_GRP = C / 16;
#pragma unroll 1
        for (int c4 = grp * F4_PER_GRP; c4 < (grp + 1) * F4_PER_GRP; ++c4) {
            float4 vv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) vv[c] = *reinterpret_cast<const float4*>(v_s + (ig + 8 * c) * CP + c4 * 4);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 gv = *reinterpret_cast<const float4*>(g_s + (jg + 8 * r) * CP + c4 * 4);
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] += dot4(gv, vv[c]);
            }
        }
#pragma unroll 1
        for (int turn = 0; turn < 4; ++turn) {
            if (grp == turn) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float* p = t_s + (jg + 8 * r) * AP + ig + 8 * c;
                        *p = turn == 0 ? acc[r][c] : *p + acc[r][c];
                    }
            }
            __syncthreads();
        }
    }

    // ---- softmax backward per row j (4 lanes per row) + dgamma partial
    const float gm = __ldg(gamma);
    {
        const int j = tid >> 2, il = tid & 3;
        float a[16], t[16], dot = 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a[u] = a_s[j * AP + il + 4 * u];
            t[u] = t_s[j * AP + il + 4 * u];
            dot += a[u] * t[u];
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
#pragma unroll
        for (int u = 0; u < 16; ++u) t_s[j * AP + il + 4 * u] = gm * a[u] * (t[u] - dot);
        const float total = block_sum(il == 0 ? dot : 0.f, red);       // (contains the barriers that publish dS)
        if (tid == 0) dgamma_part[img] = total;
    }

    // ---- dv[i][c] = gamma * sum_j A[j][i] dy[j][c]: warp w owns key rows 8w..8w+7, lane l the channels 4l..4l+3 (+128)
    {
        const int w = tid >> 5, l = tid & 31, i0 = 8 * w;
        float acc[8][4 * CPL];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4 * CPL; ++c) acc[r][c] = 0.f;
#pragma unroll 4
        for (int j = 0; j < P; ++j) {
            const float4 a0 = *reinterpret_cast<const float4*>(a_s + j * AP + i0);
            const float4 a1 = *reinterpret_cast<const float4*>(a_s + j * AP + i0 + 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int p = 0; p < CPL; ++p) {
                const float4 gv = *reinterpret_cast<const float4*>(g_s + j * CP + p * 128 + 4 * l);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    acc[r][4 * p + 0] += a[r] * gv.x;
                    acc[r][4 * p + 1] += a[r] * gv.y;
                    acc[r][4 * p + 2] += a[r] * gv.z;
                    acc[r][4 * p + 3] += a[r] * gv.w;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int p = 0; p < CPL; ++p) {
                float4 o = {gm * acc[r][4 * p + 0], gm * acc[r][4 * p + 1], gm * acc[r][4 * p + 2], gm * acc[r][4 * p + 3]};
                *reinterpret_cast<float4*>(dv + ((size_t)img * P + i0 + r) * ld_v + p * 128 + 4 * l) = o;
            }
    }

    // ---- dk (warps 0-3) and dq (warps 4-7), 4x4 outputs per thread
    if (tid < 128) {
        if (tid < 16 * NDG) {
            const int ig = tid / NDG, dg = tid - ig * NDG;      // dk[4ig+r][4dg+c] = sum_j dS[j][4ig+r] q[j][4dg+c]
            float acc[4][4] = {};
#pragma unroll 4
            for (int j = 0; j < P; ++j) {
                const float4 ds = *reinterpret_cast<const float4*>(t_s + j * AP + 4 * ig);
                const float4 qv = *reinterpret_cast<const float4*>(q_s + j * DP + 4 * dg);
                const float d[4] = {ds.x, ds.y, ds.z, ds.w};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    acc[r][0] += d[r] * qv.x;
                    acc[r][1] += d[r] * qv.y;
                    acc[r][2] += d[r] * qv.z;
                    acc[r][3] += d[r] * qv.w;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                *reinterpret_cast<float4*>(dk + ((size_t)img * P + 4 * ig + r) * ld_qk + 4 * dg) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
    } else {
        const int u = tid - 128;
        if (u < 16 * NDG) {
            const int jg = u / NDG, dg = u - jg * NDG;          // dq[4jg+r][4dg+c] = sum_i dS[4jg+r][i] k[i][4dg+c]
            float acc[4][4] = {};
#pragma unroll 2
            for (int i4 = 0; i4 < P / 4; ++i4) {
                float4 kv[4];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) kv[ii] = *reinterpret_cast<const float4*>(k_s + (4 * i4 + ii) * DP + 4 * dg);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 ds = *reinterpret_cast<const float4*>(t_s + (4 * jg + r) * AP + 4 * i4);
                    acc[r][0] += ds.x * kv[0].x + ds.y * kv[1].x + ds.z * kv[2].x + ds.w * kv[3].x;
                    acc[r][1] += ds.x * kv[0].y + ds.y * kv[1].y + ds.z * kv[2].y + ds.w * kv[3].y;
                    acc[r][2] += ds.x * kv[0].z + ds.y * kv[1].z + ds.z * kv[2].z + ds.w * kv[3].z;
                    acc[r][3] += ds.x * kv[0].w + ds.y * kv[1].w + ds.z * kv[2].w + ds.w * kv[3].w;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                *reinterpret_cast<float4*>(dq + ((size_t)img * P + 4 * jg + r) * ld_qk + 4 * dg) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
    }
}

template <int C> static int launch_fwd(const float* q, const float* k, const float* v, int ld_qk, int ld_v, const float* x, const float* gamma,
                                       float* attn, float* y, int n_img, cudaStream_t st) {
    static bool configured = false;
    const int smem = AttnSmem<C>::fwd_floats * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(attn_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return fail(GIM_E_CUDA, "attention_fwd: smem opt-in failed");
        configured = true;
    }
    attn_fwd_kernel<C><<<n_img, 256, smem, st>>>(q, k, v, ld_qk, ld_v, x, gamma, attn, y);
    return check_launch("attention_fwd");
}

template <int C> static int launch_bwd(const float* gy, const float* q, const float* k, const float* v, int ld_qk, int ld_v, const float* attn,
                                       const float* gamma, float* dq, float* dk, float* dv, float* dgamma_part, int n_img, cudaStream_t st) {
    static bool configured = false;
    const int smem = AttnSmem<C>::bwd_floats * (int)sizeof(float);
    if (!configured) {
        if (cudaFuncSetAttribute(attn_bwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return fail(GIM_E_CUDA, "attention_bwd: smem opt-in failed");
        configured = true;
    }
    attn_bwd_kernel<C><<<n_img, 256, smem, st>>>(gy, q, k, v, ld_qk, ld_v, attn, gamma, dq, dk, dv, dgamma_part);
    return check_launch("attention_bwd");
}

}  // namespace gim

using namespace gim;

static bool attention_supported(int positions, int channels) { return positions == kAttP && (channels == 128 || channels == 256); }

extern "C" {

int gim_attention_fwd(const float* q, const float* k, const float* v, int ld_qk, int ld_v, const float* x, const float* gamma, float* attn, float* y,
                      int n_img, int positions, int channels, gim_stream_t s) {
    if (n_img <= 0) return GIM_OK;
    GIM_REQUIRE(attention_supported(positions, channels), "attention_fwd: fused kernel covers 64 positions and 128 / 256 channels");
    GIM_REQUIRE(ld_qk >= channels / 8 && ld_v >= channels && ld_qk % 4 == 0 && ld_v % 4 == 0, "attention_fwd: row pitches must be multiples of 4 floats");
    GIM_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)x | (uintptr_t)y | (uintptr_t)attn) & 15) == 0, "attention_fwd: 16-byte alignment");
    return channels == 128 ? launch_fwd<128>(q, k, v, ld_qk, ld_v, x, gamma, attn, y, n_img, (cudaStream_t)s)
                           : launch_fwd<256>(q, k, v, ld_qk, ld_v, x, gamma, attn, y, n_img, (cudaStream_t)s);
}

int gim_attention_bwd(const float* gy, const float* q, const float* k, const float* v, int ld_qk, int ld_v, const float* attn, const float* gamma,
                      float* dq, float* dk, float* dv, float* dgamma_part, int n_img, int positions, int channels, gim_stream_t s) {
    if (n_img <= 0) return GIM_OK;
    GIM_REQUIRE(attention_supported(positions, channels), "attention_bwd: fused kernel covers 64 positions and 128 / 256 channels");
    GIM_REQUIRE(ld_qk >= channels / 8 && ld_v >= channels && ld_qk % 4 == 0 && ld_v % 4 == 0, "attention_bwd: row pitches must be multiples of 4 floats");
    GIM_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)gy | (uintptr_t)attn | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) == 0,
                "attention_bwd: 16-byte alignment");
    return channels == 128 ? launch_bwd<128>(gy, q, k, v, ld_qk, ld_v, attn, gamma, dq, dk, dv, dgamma_part, n_img, (cudaStream_t)s)
                           : launch_bwd<256>(gy, q, k, v, ld_qk, ld_v, attn, gamma, dq, dk, dv, dgamma_part, n_img, (cudaStream_t)s);
}

}  // extern "C"
