/* gim_b200.h -- C ABI of the B200-native GIM hot path (libgim_b200.so).
 *
 * The reference (roymor1/OptimalStrategiesAgainstGenerativeAttacks) has NO native/FFI layer: its hot
 * path is eager PyTorch library calls.  Each entry point below replaces the torch/cuDNN/cuBLAS call(s)
 * that the cited reference lines issue; the Python host code in
 * optimalstrategiesagainstgenerativeattacks_b200/ binds them with ctypes and keeps
 * the reference's nn.Module / trainer / checkpoint surface.
 *
 * Conventions
 *  - plain pointers + sizes, no torch types; every pointer is DEVICE memory owned by the caller;
 *  - kernels never allocate, free or synchronise; they are enqueued on `stream` (CUDA-graph capturable);
 *  - return 0 on success, a negative GIM_E_* code on failure (never throws); gim_last_error() gives text;
 *  - activations are NHWC ("pixels x channels"), dtype GIM_F32 or GIM_BF16; statistics, weights' masters,
 *    weight gradients, feature vectors [rows, dim] are always fp32;
 *  - conv weights are "packed": [taps = k*k][Cout][Cin] (tap t = r*k + s), the K-major B operand of the
 *    implicit GEMM; convs are stride 1, zero "same" padding (k odd) -- the only kind the reference uses
 *    (model_blocks.py:492-495, 750-751, 792-793, 838-840).
 * All file:line citations are relative to the reference root.
 */
#ifndef GIM_B200_H
#define GIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gim_stream_t;             /* cudaStream_t */

enum { GIM_F32 = 0, GIM_BF16 = 1 };
enum { GIM_OK = 0, GIM_E_ARG = -1, GIM_E_CUDA = -2, GIM_E_UNSUPPORTED = -3 };
enum { GIM_ALGO_AUTO = 0, GIM_ALGO_SIMT = 1, GIM_ALGO_TCGEN05 = 2 };

int         gim_version(void);
const char* gim_last_error(void);
/* 1 if the tcgen05/TMA implicit-GEMM path can take this conv shape/dtype, else 0 */
int         gim_conv2d_tc_supported(int n, int h, int w, int cin, int cout, int ksize, int dtype);
int         gim_conv2d_wgrad_tc_supported(int n, int h, int w, int cin, int cout, int ksize, int dtype);
/* Dry run of the tensor-core forward / input-gradient launcher (bf16 operands): the configuration gim_conv2d_fwd[_fused] WOULD launch for
 * this shape, without touching the device (works on a machine with no GPU; 148 SMs assumed there).  plan20 (host, 20 ints):
 * [0] persistent kernel (1) or non-persistent (0), [1] halo staging, [2] N tile, [3] pixel tiles per CTA, [4] cta_group::2, [5] stages of the
 * single ring, [6] activation-halo ring depth, [7] weight ring depth, [8] accumulation chains per pixel tile, [9..11] pixel box w, h, images,
 * [12] grid x, [13] grid y, [14] threads per CTA, [15] dynamic shared memory bytes, [16] TMEM columns, [17] K block, [18] pixel tiles,
 * [19] bytes of one halo box.  `epilogue`: the bits of gim_conv2d_fwd_fused. */
int         gim_conv2d_fwd_plan(int n, int h, int w, int cin, int cout, int ksize, int out_dtype, int epilogue, int* plan20);
/* The same for the tensor-core weight gradient.  plan16: [0] kernel (1 plain, 2 cta_group::2, 3 shared-dY tap groups), [1] input-channel
 * tile, [2] stages, [3] grid x (output tiles), [4] grid y (split-K over pixel tiles), [5] threads, [6] shared memory bytes, [7] TMEM columns,
 * [8] pixel tiles, [9] pixel tiles per split, [10] taps per CTA, [11..13] pixel box w, h, images. */
int         gim_conv2d_wgrad_plan(int n, int h, int w, int cin, int cout, int ksize, int* plan16);
/* counts kernel launches issued through this library since the last reset (bench.py `gpu_launches`) */
long long   gim_launch_count(int reset);
/* Parity mode: when on, every reduction output (weight gradients, column sums, split-K GEMMs, scalar dots) is owned by ONE CTA, so
 * results do not depend on the arrival order of fp32 atomics (slower; for tests).  Returns the previous setting.  The reference has no
 * equivalent knob: torch's own `torch.use_deterministic_algorithms` plays this role for its cuDNN path. */
int         gim_set_deterministic(int on);

/* ---- convolution: nn.Conv2d forward / input-grad / weight-grad (model_blocks.py:497-514, 753-773, 795-865) ---- */
/* y[n,h,w,co] = bias[co] + sum_{r,s,ci} x[n,h+r-p,w+s-p,ci] * w[r*k+s][co][ci];  x,w: dtype; y: out_dtype (fp32 accumulate
 * either way: bf16 operands with an fp32 result is the mixed-precision tensor-core path); bias fp32 or NULL */
int gim_conv2d_fwd(const void* x, const void* w, const float* bias, void* y,
                   int n, int h, int wd, int cin, int cout, int ksize, int dtype, int out_dtype, int algo, gim_stream_t stream);
/* The tcgen05 convolution with a fused epilogue (bf16 operands only; the LeakyReLU of model_blocks.py:505-509 and its backward
 * folded into the producing kernel):  v = conv(x, w) + bias;
 *   epilogue & GIM_EPI_LRELU : v = LeakyReLU_slope(v)
 *   epilogue & GIM_EPI_MASK  : v = v * (mask_ref[n,h,w,co] > 0 ? 1 : slope)   (mask_ref: bf16, same shape as y; cout % 32 == 0)
 *   epilogue & GIM_EPI_ADD   : v = v + addend[n,h,w,co]                       (addend: fp32, same shape as y, may alias y; cout % 32 == 0)
 *   epilogue & GIM_EPI_POOL  : y[n,h/2,w/2,co] = AvgPool2d(2)(conv) + bias (+ addend[n,h/2,w/2,co] with GIM_EPI_ADD): the ResBlockDown tail
 *                              (model_blocks.py:510-514) without the full-resolution tensor ever reaching HBM; even h, w; fp32 y
 *   epilogue & GIM_EPI_ADDUP : v = v + 0.25 * addend[n,h/2,w/2,co]  (fp32 half-resolution addend: the AvgPool backward of a gradient that was
 *                              propagated at the pooled resolution -- a 1x1 convolution commutes with the pooling; even h, w; cout % 32 == 0)
 * y: out_dtype (fp32 or bf16). */
#define GIM_EPI_LRELU 1
#define GIM_EPI_MASK  2
#define GIM_EPI_ADD   4
#define GIM_EPI_POOL  8
#define GIM_EPI_ADDUP 16
#define GIM_EPI_UNIT  32   /* with GIM_EPI_POOL / GIM_EPI_ADDUP: scale 1 instead of 1/4 -- sum pooling (the backward of nn.Upsample(x2), model_blocks.py:740)
                            * and the addition of a nearest-upsampled half-resolution tensor (the 1x1 residual branch of the up-sampling blocks) */
int gim_conv2d_fwd_fused(const void* x, const void* w, const float* bias, void* y, const void* mask_ref, const float* addend,
                         int n, int h, int wd, int cin, int cout, int ksize, int out_dtype, int epilogue, float slope, gim_stream_t stream);
/* gw[t][co][ci] (fp32) = sum_{n,h,w} gy[n,h,w,co] * x[n,h+r-p,w+s-p,ci]  (overwrites gw) */
int gim_conv2d_wgrad(const void* x, const void* gy, float* gw,
                     int n, int h, int wd, int cin, int cout, int ksize, int dtype, int algo, gim_stream_t stream);
/* w_fp32[t][co][ci] -> out[t][co][ci] (dtype)  /  flipped+transposed out[T-1-t][ci][co] (the dgrad operand) */
int gim_weight_cast(const float* w, void* out, int taps, int cout, int cin, int dtype, gim_stream_t stream);
int gim_weight_flip(const float* w, void* out, int taps, int cout, int cin, int dtype, gim_stream_t stream);
/* out[c] (fp32) = sum_rows x[row][c]   (bias gradient; rows = n*h*w) */
int gim_colsum(const void* x, float* out, long long rows, int c, int dtype, gim_stream_t stream);
/* out[c] += sum_rows x[row][c]: accumulates straight into a parameter's .grad (no memset, no separate accumulation kernel) */
int gim_colsum_acc(const void* x, float* out, long long rows, int c, int dtype, gim_stream_t stream);
/* y = bf16(x) and sums[c] (+)= column sums of the fp32 [rows][c] matrix x in one pass (Conv2d backward: operand copy of the output
 * gradient + bias gradient); c = 8 * 2^k <= 2048; accumulate != 0 adds into sums (e.g. straight into bias.grad) */
int gim_cast_colsum(const float* x, void* y_bf16, float* sums, long long rows, int c, int accumulate, gim_stream_t stream);

/* ---- spectral norm: torch.nn.utils.spectral_norm hook, 1 power iteration (model_blocks.py:492-495 etc.) ---- */
/* weight_orig [cout][cin][k][k] fp32; u[cout], v[cin*k*k] updated in place when power_iter!=0;
 * writes w_sn (fp32 packed [taps][cout][cin]) = W/sigma, sigma (1 float), and copies u_used/v_used for backward.
 * scratch: >= (cout + cin*k*k + 8) floats. */
int gim_sn_forward(const float* weight_orig, float* u, float* v, int power_iter, float eps,
                   float* w_sn, float* sigma, float* u_used, float* v_used, float* scratch,
                   int cout, int cin, int ksize, gim_stream_t stream);
/* g_weight_orig = (G - (sum G*W/sigma) u v^T) / sigma, G = unpacked g_w_sn */
int gim_sn_backward(const float* g_w_sn, const float* weight_orig, const float* u_used, const float* v_used,
                    const float* sigma, float* g_weight_orig, float* scratch,
                    int cout, int cin, int ksize, gim_stream_t stream);
/* The same forward for a whole list of convolutions with one launch per phase (<= 16 layers per launch group).  Per layer:
 * w [co][ci][k][k] weight_orig; u [co], v [ci*k*k] (updated in place when power_iter); w_sn fp32 [k*k][co][ci];
 * w_op / w_flip: optional bf16 copies of w_sn and of its flipped+transposed pack [k*k][ci][co] (the conv / dgrad operands), NULL to
 * skip; aux fp32 [co + ci*k*k + 1] = u_used | v_used | sigma; scratch fp32 [ci*k*k + co].  `layers` is a HOST array (copied into
 * the launch parameters: CUDA-graph safe). */
typedef struct {
    const float* w;
    float* u;
    float* v;
    float* w_sn;
    void* w_op;
    void* w_flip;
    float* aux;
    float* scratch;
    int cout, cin, ksize, reserved;
} gim_sn_layer;
int gim_sn_forward_multi(const gim_sn_layer* layers, int n_layers, int power_iter, float eps, gim_stream_t stream);
/* Batched backward of the same: grad[co][ci][k][k] (+)= g[t][co][ci]/sigma - (sum(g.w)/sigma^2) u[co] v[ci*k*k+t].  g: gradient of the
 * packed weight; u, v, sigma: the values the forward used (aux); scratch: one fp32 word per layer; accumulate != 0 adds into grad
 * (the parameter's .grad), so no separate accumulation kernel runs.  `layers` is a HOST array. */
typedef struct {
    const float* g;
    const float* w;
    const float* u;
    const float* v;
    const float* sigma;
    float* grad;
    float* scratch;
    int cout, cin, ksize, accumulate;
} gim_sn_bwd_layer;
int gim_sn_backward_multi(const gim_sn_bwd_layer* layers, int n_layers, gim_stream_t stream);

/* ---- pointwise / resampling (model_blocks.py:489-490, 740, 744; gim_img_models.py:215) ---- */
int gim_lrelu_fwd(const void* x, void* y, long long n, float slope, int dtype, gim_stream_t stream);
int gim_lrelu_bwd(const void* gy, const void* x, void* gx, long long n, float slope, int dtype, gim_stream_t stream);
int gim_tanh_fwd(const void* x, void* y, long long n, int dtype, gim_stream_t stream);
int gim_tanh_bwd(const void* gy, const void* y, void* gx, long long n, int dtype, gim_stream_t stream);
/* out = alpha*x + beta*y (y may be NULL) */
int gim_axpby(const void* x, const void* y, void* out, long long n, float alpha, float beta, int dtype, gim_stream_t stream);
/* out = (*scalar)*x ; scalar is a device fp32 (SelfAttention gamma, model_blocks.py:548) */
int gim_scale_dev(const void* x, const float* scalar, void* out, long long n, int dtype, gim_stream_t stream);
/* out[0] (fp32, overwritten) = sum x*y */
int gim_dot(const void* x, const void* y, float* out, long long n, int dtype, gim_stream_t stream);
/* y[n,ho,wo,c] = scale * sum_{2x2} (a (+ b)) , ho=h/2, wo=wd/2 (floor)  -- AvgPool2d(2) (+ fused residual add) */
int gim_pool2_sum(const void* a, const void* b, void* y, int n, int h, int wd, int c, float scale, int dtype, gim_stream_t stream);
/* gx[n,h,w,c] = scale * gy[n,h/2,w/2,c] (0 where h/2>=ho or w/2>=wo) -- nearest Upsample x2 / AvgPool backward */
int gim_unpool2_bcast(const void* gy, void* gx, int n, int h, int wd, int c, float scale, int dtype, gim_stream_t stream);
/* layout + dtype conversion at the module boundary: NCHW fp32 <-> NHWC dtype */
/* fused ResBlockDown tail (model_blocks.py:510-514): y = scale * 2x2-sum(a [+ b]) written as any of fp32 y, bf16(y), bf16(LeakyReLU(y))
 * (NULL = skip); and the AvgPool backward emitted directly as the bf16 conv operand: gx = bf16(scale * gy[h/2, w/2]) */
int gim_pool2_multi(const float* a, const float* b, float* y32, void* y_bf16, void* y_lrelu_bf16,
                    int n, int h, int wd, int c, float scale, float slope, gim_stream_t stream);
int gim_unpool2_cast(const float* gy, void* gx_bf16, int n, int h, int wd, int c, float scale, gim_stream_t stream);
/* First ResBlockDown of an encoder (image input with a few channels, model_blocks.py:497-509): both input-side convolutions in one pass.
 * x fp32 NHWC [n,h,w,c]; w_r1 fp32 packed [k*k][cout][c], w_l1 fp32 packed [1][cout][c] (spectral-normalised), biases fp32;
 * t = bf16(LeakyReLU(conv_k(LeakyReLU(x)) + b_r1)) [n,h,w,cout];  res = conv_1x1(AvgPool2(x)) + b_l1 [n,h/2,w/2,cout] fp32.
 * Operands rounded to bf16 like the tensor-core path.  c*k*k <= 64, cout % 8 == 0, even h and w. */
int gim_first_block_fwd(const float* x, const float* w_r1, const float* b_r1, const float* w_l1, const float* b_l1, void* t_bf16, float* res_pooled,
                        int n, int h, int wd, int c, int cout, int ksize, float slope, gim_stream_t stream);
/* weight gradients of the same two convolutions (3x3, c = 1 or 3): gw_r1 fp32 packed [9][cout][c] from gt = bf16 masked gradient of the
 * k x k conv output [n,h,w,cout]; gw_l1 fp32 [cout][c] from gy = fp32 gradient of the pooled block output [n,h/2,w/2,cout]; both overwritten */
/* `scratch` (>= 10*cout*c floats, ideally 32 x that; may be NULL): the tensor-core pass lets each CTA add into one of several zeroed
 * copies of the result held there and folds them afterwards; without scratch the CUDA-core pass adds into gw_* directly */
int gim_first_block_wgrad(const float* x, const void* gt_bf16, const float* gy_pooled, float* gw_r1, float* gw_l1, float* scratch, long long scratch_floats,
                          int n, int h, int wd, int c, int cout, int ksize, float slope, gim_stream_t stream);
int gim_nchw_to_nhwc(const float* x, void* y, int n, int c, int h, int wd, int dtype, gim_stream_t stream);
int gim_nhwc_to_nchw(const void* x, float* y, int n, int c, int h, int wd, int dtype, gim_stream_t stream);
/* dst[row][dst_off + j] = src[row][src_off + j], j<c  (channel concat / split, gim_img_models.py:385) */
int gim_copy_cols(const void* src, int src_ld, int src_off, void* dst, int dst_ld, int dst_off,
                  long long rows, int c, int dtype, gim_stream_t stream);
int gim_cast(const void* x, int dtype_in, void* y, int dtype_out, long long n, gim_stream_t stream);
/* conv-operand producer: out (dtype_out, NHWC) = f(x): mode 0 identity (cast), 1 LeakyReLU(slope), 2 nearest-upsample x2
 * (out is [n,2h,2w,c]) -- fuses the activation / nn.Upsample in front of a conv (model_blocks.py:505-509, 761-766) with the
 * single rounding to the operand dtype */
int gim_operand_prepare(const void* x, int dtype_in, void* out, int dtype_out, int n, int h, int wd, int c, int mode, float slope, gim_stream_t stream);
/* gx = g * (ref > 0 ? 1 : slope), mask from `ref` of dtype ref_dtype (the saved operand) */
int gim_lrelu_bwd_ref(const float* g, const void* ref, int ref_dtype, float* gx, long long n, float slope, gim_stream_t stream);
/* tap unrolling of skinny-channel tensors (c in {1,2,3,6}) so their convs become dense 1x1 tensor-core GEMMs:
 * out[pix][t*c+ch] = x[pix + sign*offset(t)][ch] (zeros outside / for j >= k*k*c), rows of length kc */
int gim_im2col(const void* x, void* out, int n, int h, int wd, int c, int ksize, int sign, int kc, int dtype, gim_stream_t stream);
/* y[pix][ch] = bias[ch] + sum_t z[pix + offset(t)][t*c+ch]  (fp32; z rows of length ld) -- the adjoint gather */
int gim_col2im(const float* z, const float* bias, float* y, int n, int h, int wd, int c, int ksize, int ld, gim_stream_t stream);

/* ---- InstanceNorm2d / ada_in (model_blocks.py:611-630, 747-748; gim_img_models.py:126) ---- */
/* per (n,c): mean and M2 = sum (x-mean)^2 over the hw pixels */
int gim_norm_stats(const void* x, float* mean, float* m2, int n, int hw, int c, int dtype, gim_stream_t stream);
/* y = act(a[n,c]*(x - mean[n,c]) + b[n,c]),  act = LeakyReLU(slope) if slope != 1 */
int gim_affine_act_fwd(const void* x, const float* mean, const float* a, const float* b, void* y, int n, int hw, int c, float slope, int dtype, gim_stream_t stream);
/* s1[n,c] = sum gyh, s2[n,c] = sum gyh*(x-mean), gyh = gy * act'(y)   (y = forward output, NULL if no act) */
/* (act_a, act_b: when y is NULL and these are not, the activation mask is recomputed as sign(act_a (x - mean) + act_b) -- the forward
 *  output was never materialised, see gim_norm_act_operand) */
int gim_norm_bwd_reduce(const void* gy, const void* x, const void* y, const float* mean, const float* act_a, const float* act_b, float* s1, float* s2,
                        int n, int hw, int c, float slope, int dtype, gim_stream_t stream);
/* gx = A[n,c]*gyh + B[n,c]*(x-mean[n,c]) + C[n,c] */
int gim_norm_bwd_apply(const void* gy, const void* x, const void* y, const float* mean, const float* act_a, const float* act_b, const float* A, const float* B,
                       const float* C, void* gx, int n, int hw, int c, float slope, int dtype, gim_stream_t stream);
/* out[n,(2)h,(2)w,c] = bf16(LeakyReLU_slope(a[n,c] (x - mean[n,c]) + b[n,c])), nearest-upsampled x2 when upsample != 0: the
 * norm -> activation -> (nn.Upsample) -> conv-operand chain of ResBlockUp / AdaResBlock2 / AdaResBlockUp2 (model_blocks.py:760-768, 805-811,
 * 851-861) in one pass; x fp32 NHWC, c % 8 == 0 */
int gim_norm_act_operand(const float* x, const float* mean, const float* a, const float* b, void* out_bf16, int n, int h, int wd, int c, float slope,
                         int upsample, gim_stream_t stream);
/* coefficient kernels on [n,c] fp32 arrays.  mode 0 = InstanceNorm(weight[c],bias[c],eps: biased var),
 * mode 1 = ada_in(std_style[n,c], mean_style[n,c], eps added to the unbiased std) */
int gim_norm_coeffs(int mode, const float* mean, const float* m2, const float* p_scale, const float* p_shift,
                    float* a, float* b, int n, int hw, int c, float eps, gim_stream_t stream);
int gim_norm_bwd_coeffs(int mode, const float* m2, const float* s1, const float* s2, const float* p_scale,
                        float* A, float* B, float* C, float* g_scale, float* g_shift,
                        int n, int hw, int c, float eps, gim_stream_t stream);

/* ---- small dense algebra (nn.Linear, torch.bmm, nn.Softmax: model_blocks.py:77-94, 541-545, 786-789) ---- */
/* C[b] = alpha * A[b] x B[b] + beta * C[b]; element (m,k) of A[b] at A + b*sAb + m*sAm + k*sAk, etc. C is [m][n] with ldc */
int gim_gemm_strided(const void* A, int dtA, long long sAb, long long sAm, long long sAk,
                     const void* B, int dtB, long long sBb, long long sBk, long long sBn,
                     void* C, int dtC, long long sCb, long long ldc,
                     int m, int n, int k, int batch, float alpha, float beta, gim_stream_t stream);
/* y[row][j] = x[row][j] + bias[j] ; act LeakyReLU(slope) if slope!=1 */
/* the same GEMM on the tensor cores: each operand split hi+lo into two bf16 numbers, product = hi*hi + lo*hi + hi*lo in fp32
 * (~2^-16 relative: fp32-grade attention logits at tensor-core speed) */
int gim_gemm_strided_bf16(const void* A, int dtA, long long sAb, long long sAm, long long sAk,
                          const void* B, int dtB, long long sBb, long long sBk, long long sBn,
                          void* C, int dtC, long long sCb, long long ldc,
                          int m, int n, int k, int batch, float alpha, float beta, gim_stream_t stream);
int gim_bias_act_fwd(const float* x, const float* bias, float* y, long long rows, int c, float slope, gim_stream_t stream);
int gim_softmax_rows_fwd(const float* x, float* y, long long rows, int cols, gim_stream_t stream);
int gim_softmax_rows_bwd(const float* gy, const float* y, float* gx, long long rows, int cols, gim_stream_t stream);
/* gradient of softmax_rows_bwd(gy,y) w.r.t. y given upstream ggx (R1 double backward) */
int gim_softmax_rows_bwd_bwd(const float* ggx, const float* gy, const float* y, float* g_y, long long rows, int cols, gim_stream_t stream);

/* ---- fused SelfAttention core over an 8x8 map (model_blocks.py:517-549, between the three 1x1 convs and the block output) ----
 * q (reference conv_g), k (conv_f): 64 rows of channels/8 floats per image, `ld_qk` floats between rows; v (conv_h): 64 rows of
 * `channels` floats, `ld_v` between rows (the three may be column blocks of ONE [n_img][64][2*channels/8 + channels] projection:
 * ld_qk = ld_v = its row length); x, y [n_img][64][channels] and attn [n_img][64][64] contiguous; all fp32, 16-byte aligned:
 *   attn[n][j][i] = softmax_i(<q_j, k_i>),  y[j] = gamma[0] * sum_i attn[j][i] v[i] + x[j].
 * One CTA per image, everything staged in shared memory.  positions must be 64 and channels 128 or 256 (GIM_E_ARG otherwise:
 * the caller composes gemm + softmax for other shapes and whenever a second-order graph is needed). */
int gim_attention_fwd(const float* q, const float* k, const float* v, int ld_qk, int ld_v, const float* x, const float* gamma,
                      float* attn, float* y, int n_img, int positions, int channels, gim_stream_t stream);
/* gradients w.r.t. q, k, v (written with the same row pitches as q, k, v); dgamma_part[n_img] = per-image partial of d/dgamma
 * (the caller sums them); d/dx = gy */
int gim_attention_bwd(const float* gy, const float* q, const float* k, const float* v, int ld_qk, int ld_v, const float* attn,
                      const float* gamma, float* dq, float* dk, float* dv, float* dgamma_part, int n_img, int positions, int channels,
                      gim_stream_t stream);

/* ---- permutation-invariant set statistics over the sample axis (gim_basic_models.py:20-51, 152-172; model_blocks.py:41-48) ---- */
/* x [b][s][d] fp32.  out_sum[b*ld + j] = scale * sum_s x ; out_std = sqrt(var_unbiased + eps) (zeros if s==1); either may be NULL */
int gim_set_stats_fwd(const float* x, float* out_sum, float* out_std, int ld_out, int b, int s, int d, float scale, float eps, gim_stream_t stream);
/* gx[b][s][d] = scale*g_sum[b][j] + g_std[b][j]*(x-mean)/((s-1)*std)   (g_* rows have stride ld_g; either may be NULL) */
int gim_set_stats_bwd(const float* g_sum, const float* g_std, int ld_g, const float* x, float* gx, int b, int s, int d, float scale, float eps, gim_stream_t stream);
/* second order of the std term: given ggx -> gg_std[b][d], g_x[b][s][d] */
int gim_set_std_bwd_bwd(const float* ggx, const float* g_std, int ld_g, const float* x, float* gg_std, float* g_x, int b, int s, int d, float eps, gim_stream_t stream);
/* y[b][s][d] = x - mean_s(x) (+ add[b][d] if add != NULL)   (gim_img_models.py:378-380, gim_gaussian_models.py:84-88) */
int gim_set_center_add(const float* x, const float* add, float* y, int b, int s, int d, int center, gim_stream_t stream);
/* y[b][s][d] = alpha * x[b][s][d] + beta * add[b][d]: Gaussian episode synthesis mu + sigma * noise on the device
 * (replaces the host-side torch.normal calls of training/gim_gaussian_training.py:71-86) */
int gim_affine_rows(const float* x, const float* add, float* y, int b, int s, int d, float alpha, float beta, gim_stream_t stream);

/* ---- ImgAttention blend (model_blocks.py:596-608): per pixel s_i = <q_i, k_i> over the c channels, (a1, a2) = softmax(s1, s2),
 * out = a1 * x1 + a2 * v2; all tensors NHWC fp32 [pixels][c], att [pixels][2].  Backward: gx1 may be NULL (x1 is an input image). ---- */
int gim_img_att_blend_fwd(const float* q1, const float* k1, const float* q2, const float* k2, const float* x1, const float* v2, float* out, float* att,
                          long long pixels, int c, gim_stream_t stream);
int gim_img_att_blend_bwd(const float* g, const float* q1, const float* k1, const float* q2, const float* k2, const float* x1, const float* v2,
                          const float* att, float* gq1, float* gk1, float* gq2, float* gk2, float* gx1, float* gv2, long long pixels, int c,
                          gim_stream_t stream);

/* ---- encoder tail: AdaptiveMaxPool2d((1,1)) (gim_img_models.py:53-54) ---- */
/* y[n][c] = LeakyReLU_slope(max_p x[n][p][c]), idx = first arg-max (slope 1: plain max) -- the encoder tail in one pass */
int gim_gmax_fwd(const void* x, float* y, int32_t* idx, int n, int hw, int c, float slope, int dtype, gim_stream_t stream);
int gim_gather_idx(const void* x, const int32_t* idx, float* y, int n, int hw, int c, int dtype, gim_stream_t stream);
int gim_scatter_idx(const float* g, const int32_t* idx, void* gx, int n, int hw, int c, int dtype, gim_stream_t stream);

/* ---- losses (gim_img_trainer.py:90-94; training/utils.py:115-124) ---- */
int gim_bce_logits_fwd(const float* x, float target, float* loss, long long n, gim_stream_t stream);
int gim_bce_logits_bwd(const float* g, const float* x, float target, float* gx, long long n, gim_stream_t stream);
/* out[b] = sum_j x[b][j]^2 (fp32 out) ;  y[b][j] = alpha * s[b] * x[b][j] */
int gim_rows_sqsum(const void* x, float* out, int b, long long l, int dtype, gim_stream_t stream);
int gim_rows_scale(const void* x, const float* s, void* y, int b, long long l, float alpha, int dtype, gim_stream_t stream);

/* ---- fused multi-tensor Adam (torch.optim.Adam at gim_img_trainer.py:50-58, gim_gaussian_trainer.py:48-49) ---- */
typedef struct {
    float*       p;      /* parameter */
    const float* g;      /* gradient */
    float*       m;      /* exp_avg */
    float*       v;      /* exp_avg_sq */
    long long    numel;
    int          group;  /* index into lrs[] */
    int          pad;
} gim_adam_tensor;
/* `table` and `lrs` are device arrays; `step` is a device int64 counter incremented by the kernel (bias
 * corrections use step+1); grad_scale multiplies g first (1/world_size after an allreduce(sum)). */
int gim_adam_multi(const gim_adam_tensor* table, int n_tensors, long long max_numel, const float* lrs,
                   long long* step, float beta1, float beta2, float eps, float grad_scale, gim_stream_t stream);
/* g = 0 for every tensor of the same table in one launch (optimizer.zero_grad() with the gradient buffers kept in place) */
int gim_zero_grads_multi(const gim_adam_tensor* table, int n_tensors, long long max_numel, gim_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GIM_B200_H */
