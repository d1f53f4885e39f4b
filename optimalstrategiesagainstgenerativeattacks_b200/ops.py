"""Differentiable operators of the GIM hot path, each a thin torch.autograd.Function over the C-ABI kernels.

Design rule: the backward of every operator that sits on the authenticator path is itself written with these
operators, so the set is closed under differentiation -- that is what lets `compute_grad2` (R1 penalty,
training/utils.py:115-124 of the reference) build a second-order graph through hand-written kernels.
  conv / conv-transpose / weight-grad   are closed among themselves,
  pool2 <-> unpool2, gmax-scatter <-> gather, colsum <-> row-broadcast, set-sum <-> set-broadcast, matmul, scale/dot,
  softmax and set-std carry an explicit second-order kernel.
Operators that only occur in the attacker (InstanceNorm, ada_in, tanh, channel concat) are first order.

Activations are NHWC tensors of the active precision (`set_precision`): float32 (parity path, CUDA-core FFMA) or
bfloat16 (tcgen05 tensor-core path).  Feature vectors, statistics, weights and weight gradients are float32.
"""
import contextlib
import os
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _cabi as C

LRELU_SLOPE = 0.2
_state = {"operand_dtype": torch.float32, "conv_algo": C.ALGO_AUTO, "input_grads_only": False, "composite": False, "defer_sn": False,
          "deterministic": False}


def set_precision(name):
    """'fp32': everything float32, CUDA-core FFMA convolutions (rel 1e-4 parity path).
    'bf16': mixed precision -- every convolution / Linear operand (activation, gradient, weight) is rounded to bfloat16
    once and multiplied on the tcgen05 tensor cores with fp32 accumulation; everything else (conv results, residual
    sums, pooling, norm statistics, losses, optimizer) stays float32 (rel 2e-2 path)."""
    _state["operand_dtype"] = {"fp32": torch.float32, "bf16": torch.bfloat16}[name]
    _cast_memo.clear()


def get_precision():
    return "fp32" if _state["operand_dtype"] == torch.float32 else "bf16"


def act_dtype():
    """Storage dtype of activations between operators (always float32; bf16 only exists as conv operands)."""
    return torch.float32


def operand_dtype():
    return _state["operand_dtype"]


def set_deterministic(flag):
    """Parity mode: every reduction output is owned by one CTA (gim_set_deterministic) and the image-side weight-gradient kernel, whose
    warps meet in shared-memory atomics, is replaced by the tensor-core route -- results are then bit-reproducible run to run (eager or
    CUDA graph).  Slower; meant for tests.  -> previous setting."""
    old = _state["deterministic"]
    _state["deterministic"] = bool(flag)
    C.set_deterministic(flag)
    return old


def set_conv_algo(algo):
    _state["conv_algo"] = {"auto": C.ALGO_AUTO, "simt": C.ALGO_SIMT, "tcgen05": C.ALGO_TCGEN05}[algo]


@contextlib.contextmanager
def input_grads_only():
    """Inside, conv backward skips weight/bias gradients (used for the R1 input-gradient pass, whose weight
    gradients nobody reads)."""
    old = _state["input_grads_only"]
    _state["input_grads_only"] = True
    try:
        yield
    finally:
        _state["input_grads_only"] = old


@contextlib.contextmanager
def composite_mode():
    """Inside, the block-level fused operators are bypassed and every block runs as a composition of the elementary operators,
    whose backward is itself differentiable -- required wherever a second-order graph is built (the R1 pass, compute_grad2)."""
    old = _state["composite"]
    _state["composite"] = True
    try:
        yield
    finally:
        _state["composite"] = old


def fused_blocks_enabled():
    """Block-level fusion exists on the bf16 tensor-core path only; the fp32 path stays the plain rel-1e-4 parity path."""
    return _state["operand_dtype"] == torch.bfloat16 and _state["conv_algo"] != C.ALGO_SIMT and not _state["composite"]


# ---- two-branch stream parallelism ------------------------------------------------------------------------------------
_side = {"stream": None, "enabled": True, "used": False}


def set_stream_parallelism(flag):
    _side["enabled"] = bool(flag)


def two_streams(fn_main, fn_side):
    """Run two independent branches (the src and env encoders see the same images) on two CUDA streams: -> (fn_main(), fn_side()).
    Inside a captured CUDA graph the branches become parallel graph branches, so one branch's kernels fill the tail waves and launch
    gaps of the other.  Autograd replays each branch's backward on the stream its forward ran on.  Results are identical to running
    the branches one after the other (they share no state)."""
    if not (_side["enabled"] and torch.cuda.is_available()):
        return fn_main(), fn_side()
    cur = torch.cuda.current_stream()
    if _side["stream"] is None or _side["stream"].device != cur.device:
        _side["stream"] = torch.cuda.Stream(device=cur.device)
    side = _side["stream"]
    if side == cur:                                       # nested use from inside the side branch
        return fn_main(), fn_side()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        b = fn_side()
    a = fn_main()
    cur.wait_stream(side)
    _side["used"] = True
    for t in (b if isinstance(b, (tuple, list)) else (b,)):
        if torch.is_tensor(t):
            t.record_stream(cur)
    return a, b


def _empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


_cast_memo = {}     # id(tensor) -> (weakref, version, operand copy): conv_l1 / attention convs / dgrad+wgrad share one cast


def register_operand(t, op):
    """Tell the operand cache that `op` already is `t` rounded to the operand dtype (written by a fused producer kernel)."""
    if op is not None and op.dtype == _state["operand_dtype"]:
        if len(_cast_memo) >= 4:
            _cast_memo.clear()
        _cast_memo[id(t)] = (weakref.ref(t), t._version, op)


def _operand(t):
    """The tensor as a conv/GEMM operand of the active precision (one rounding to bf16 on the tensor-core path)."""
    od = _state["operand_dtype"]
    if t.dtype == od:
        return t
    hit = _cast_memo.get(id(t))
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    out = torch.empty(t.shape, dtype=od, device=t.device)
    C.call("gim_cast", C.ptr(t), C.dtype_code(t), C.ptr(out), C.dtype_code(out), t.numel())
    if len(_cast_memo) >= 4:
        _cast_memo.clear()
    _cast_memo[id(t)] = (weakref.ref(t), t._version, out)
    return out


# ------------------------------------------------------------------------------------------------------------
# convolution triple
# ------------------------------------------------------------------------------------------------------------
_wops = {}          # id(packed fp32 weight) -> (weakref, bf16 copy, flipped bf16 copy) written by the batched spectral-norm pass


def _weight_as(w32, dtype, flip):
    taps, co, ci = w32.shape
    if not flip and dtype == torch.float32:
        return w32
    if dtype == torch.bfloat16:
        hit = _wops.get(id(w32))
        if hit is not None and hit[0]() is w32 and hit[2 if flip else 1] is not None:
            return hit[2] if flip else hit[1]
    out = _empty((taps, ci, co) if flip else (taps, co, ci), dtype, w32)
    C.call("gim_weight_flip" if flip else "gim_weight_cast", C.ptr(w32), C.ptr(out), taps, co, ci, C.dtype_code(out))
    return out


_profile = None      # bench.py: {"kind": [ (flops, start_event, end_event), ... ]} while a profiling pass is active


def conv_profile_begin():
    global _profile
    _profile = {}


def conv_profile_end():
    """-> {kind: {"flops", "ms", "launches"}} summed over the launches recorded since conv_profile_begin()."""
    global _profile
    prof, _profile = _profile, None
    torch.cuda.synchronize()
    out = {}
    shapes = {}
    for kind, recs in prof.items():
        out[kind] = {"flops": float(sum(r[0] for r in recs)), "ms": float(sum(r[1].elapsed_time(r[2]) for r in recs)), "launches": len(recs)}
        for r in recs:
            e = shapes.setdefault((kind,) + r[3], [0.0, 0.0, 0])
            e[0] += r[0]
            e[1] += r[1].elapsed_time(r[2])
            e[2] += 1
    out["_shapes"] = shapes
    return out


def _timed_call(kind, flops, name, *args):
    if _profile is None:
        C.call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # keep the GPU busy (~25 us of spinning BEFORE the start event) while the host enqueues event, kernel, event: otherwise an idle GPU
    # records the start event at once and the host's launch latency (5-10 us of Python + ctypes) is billed to every small kernel
    torch.cuda._sleep(50000)
    e0.record()
    C.call(name, *args)
    e1.record()
    off = {"gim_conv2d_wgrad": 3, "gim_conv2d_fwd_fused": 6}.get(name, 4)
    _profile.setdefault(kind, []).append((flops, e0, e1, tuple(args[off:off + 6])))     # (n, h, w, ci, co, ks)


def _conv_kind(n, h, w, ci, co, ks, x, wgrad):
    tc = _state["conv_algo"] != C.ALGO_SIMT and (C.wgrad_tc_supported if wgrad else C.conv_tc_supported)(n, h, w, ci, co, ks, C.dtype_code(x))
    return ("tcgen05" if tc else "cuda_core") + ("_wgrad" if wgrad else "")


def _pad_last(t, cp):
    """Zero-pad the channel (last) dim of a contiguous tensor to cp (skinny layers -> tensor-core friendly widths)."""
    c = t.shape[-1]
    if c == cp:
        return t
    out = torch.zeros(t.shape[:-1] + (cp,), dtype=t.dtype, device=t.device)
    C.call("gim_copy_cols", C.ptr(t), c, 0, C.ptr(out), cp, 0, t.numel() // c, c, C.dtype_code(t))
    return out


def _narrow_last(t, c):
    cp = t.shape[-1]
    if c == cp:
        return t
    out = torch.empty(t.shape[:-1] + (c,), dtype=t.dtype, device=t.device)
    C.call("gim_copy_cols", C.ptr(t), cp, 0, C.ptr(out), c, 0, out.numel() // c, c, C.dtype_code(t))
    return out


def _round_up(v, m):
    return (v + m - 1) // m * m


def _use_tc(x):
    return x.dtype == torch.bfloat16 and _state["conv_algo"] != C.ALGO_SIMT


def _im2col(x, ks, sign, kc):
    n, h, w, c = x.shape
    out = torch.empty((n, h, w, kc), dtype=x.dtype, device=x.device)
    C.call("gim_im2col", C.ptr(x), C.ptr(out), n, h, w, c, ks, sign, kc, C.dtype_code(x))
    return out


def _conv_raw(x, w_t, bias, ks):
    """x [n,h,w,ci], w_t [taps,co,ci] (both operand dtype), bias fp32|None -> fp32 [n,h,w,co].
    On the bf16 path layers with a skinny side (image / last layers: 1, 2, 3, 6 channels) are rewritten as dense 1x1 GEMMs so
    that they, too, run on the tcgen05 kernel at full tile efficiency:
      skinny input : unroll x over the taps (im2col of a few channels is cheap)  -> K = taps*ci
      skinny output: z = x @ W^T with N = taps*co, then gather-sum the taps (col2im)."""
    n, h, w, ci = x.shape
    taps, co, _ = w_t.shape
    if _use_tc(x) and ci % 8:
        kc = _round_up(taps * ci, 8)
        w2 = torch.zeros((1, co, kc), dtype=w_t.dtype, device=w_t.device)
        w2[0, :, :taps * ci] = w_t.permute(1, 0, 2).reshape(co, taps * ci)
        return _conv_raw(_im2col(x, ks, 1, kc), w2, bias, 1)
    if _use_tc(x) and (co % 8 or (co % 16 and co < 24)):
        ld = _round_up(taps * co, 16)
        w2 = torch.zeros((1, ld, ci), dtype=w_t.dtype, device=w_t.device)
        w2[0, :taps * co] = w_t.reshape(taps * co, ci)
        z = _conv_raw(x, w2, None, 1)
        y = _empty((n, h, w, co), torch.float32, x)
        C.call("gim_col2im", C.ptr(z), C.ptr(bias), C.ptr(y), n, h, w, co, ks, ld)
        return y
    y = _empty((n, h, w, co), torch.float32, x)
    kind = _conv_kind(n, h, w, ci, co, ks, x, False) if _profile is not None else None
    _timed_call(kind, 2.0 * n * h * w * ci * co * ks * ks, "gim_conv2d_fwd", C.ptr(x), C.ptr(w_t), C.ptr(bias), C.ptr(y), n, h, w, ci, co, ks,
                C.dtype_code(x), C.F32, _state["conv_algo"])
    return y


# ---- weight gradients off the critical path ------------------------------------------------------------------------------------
# In a backward pass only the input gradients chain from layer to layer; a weight gradient is not read before the batched spectral-norm
# backward at the END of the pass (deferred_weight_grads).  Launched on a stream of their own (forked after the layer's output gradient
# exists, joined by the flush), the weight-gradient kernels become parallel branches of the captured graph and fill the tail waves of
# the persistent forward / input-gradient kernels.  Only when the consumer is known to be the deferred flush (`_wg_async_ok`) and only
# under graph capture (GIM_WGRAD_STREAMS=2 also in eager mode, =0 never).
_wg = {"streams": {}, "used": [], "enabled": os.environ.get("GIM_WGRAD_STREAMS", "1") != "0", "eager": os.environ.get("GIM_WGRAD_STREAMS", "1") == "2"}
_DEFERRED_GRAD_FNS = ("SpectralNormPreparedFnBackward", "SpectralNormFnBackward", "_MergedRowsFnBackward")


def set_wgrad_streams(flag):
    _wg["enabled"] = bool(flag)


def _wg_async_ok(w):
    """Forward-time check: the gradient of this weight tensor flows into a spectral-norm node (whose backward only queues it)."""
    fn = getattr(w, "grad_fn", None)
    return fn is not None and type(fn).__name__ in _DEFERRED_GRAD_FNS


def _wg_stream():
    cur = torch.cuda.current_stream()
    key = (cur.device.index, cur.cuda_stream)
    ws = _wg["streams"].get(key)
    if ws is None:
        ws = _wg["streams"][key] = torch.cuda.Stream(device=cur.device)
    return cur, ws


def _wg_join():
    """The current stream waits for every weight-gradient stream that has work in flight."""
    if _wg["used"]:
        cur = torch.cuda.current_stream()
        for ws in _wg["used"]:
            cur.wait_stream(ws)
        del _wg["used"][:]


def _wgrad_raw(x, g, ks, async_ok=False):
    n, h, w, ci = x.shape
    co = g.shape[3]
    taps = ks * ks
    # only while a CUDA graph is being captured: there the fork/join is a pair of graph edges; in eager mode the extra events, the
    # cross-stream allocator bookkeeping and the lost memory reuse cost more than the overlap buys (measured: 1069 -> 518 episodes/s on
    # the eager 105x105 authenticator benchmark)
    if (async_ok and _wg["enabled"] and _side["enabled"] and _state["defer_sn"] and _profile is None and not torch.is_grad_enabled()
            and _use_tc(x) and ci % 8 == 0 and co % 8 == 0 and (torch.cuda.is_current_stream_capturing() or _wg["eager"])):
        cur, ws = _wg_stream()
        ws.wait_stream(cur)                              # x and g exist on the launching stream
        with torch.cuda.stream(ws):
            gw = _empty((taps, co, ci), torch.float32, x)
            C.call("gim_conv2d_wgrad", C.ptr(x), C.ptr(g), C.ptr(gw), n, h, w, ci, co, ks, C.dtype_code(x), _state["conv_algo"])
        x.record_stream(ws)                              # their memory must outlive the side-stream kernel
        g.record_stream(ws)
        gw.record_stream(cur)                            # consumed by the flush on the main stream
        if ws not in _wg["used"]:
            _wg["used"].append(ws)
        return gw
    if _use_tc(x) and ci % 8:                       # gw[t][co][c] = sum_p g[p][co] * xcol[p][t*ci+c]
        kc = _round_up(taps * ci, 8)
        gw2 = _wgrad_raw(_im2col(x, ks, 1, kc), g, 1)
        return gw2[0, :, :taps * ci].reshape(co, taps, ci).permute(1, 0, 2).contiguous()
    if _use_tc(x) and co % 8:                       # gw[t][c][ci] = sum_q gcol[q][t*co+c] * x[q][ci],  gcol[q][t,c] = g[q - t][c]
        ld = _round_up(taps * co, 8)
        gw2 = _wgrad_raw(x, _im2col(g, ks, -1, ld), 1)
        return gw2[0, :taps * co].reshape(taps, co, ci).contiguous()
    gw = _empty((taps, co, ci), torch.float32, x)
    kind = _conv_kind(n, h, w, ci, co, ks, x, True) if _profile is not None else None
    _timed_call(kind, 2.0 * n * h * w * ci * co * ks * ks, "gim_conv2d_wgrad", C.ptr(x), C.ptr(g), C.ptr(gw), n, h, w, ci, co, ks,
                C.dtype_code(x), _state["conv_algo"])
    return gw


PRE_NONE, PRE_LRELU, PRE_UPSAMPLE = 0, 1, 2


def _prepare_operand(x, pre, slope):
    """Operand of the active precision = prologue(x): identity / LeakyReLU / nearest-upsample x2, fused with the rounding."""
    if pre == PRE_NONE:
        return _operand(x)
    n, h, w, c = x.shape
    shape = (n, 2 * h, 2 * w, c) if pre == PRE_UPSAMPLE else (n, h, w, c)
    out = torch.empty(shape, dtype=_state["operand_dtype"], device=x.device)
    C.call("gim_operand_prepare", C.ptr(x), C.dtype_code(x), C.ptr(out), C.dtype_code(out), n, h, w, c, pre, slope)
    return out


class Conv2dFn(Function):
    """y = conv(prologue(x), w) + b.  nn.Conv2d(stride 1, 'same') of model_blocks.py:492-495, optionally with the LeakyReLU
    (model_blocks.py:505-509) or the nearest nn.Upsample(x2) (:761-766, 851-858) that precedes it folded into the operand producer."""

    @staticmethod
    def forward(ctx, x, w32, bias, ks, pre=PRE_NONE, slope=LRELU_SLOPE):
        x = _c(x)
        w32 = _c(w32)
        xop = _prepare_operand(x, pre, slope)
        ctx.cfg = (ks, pre, slope, bias is not None)
        ctx.bias_param = bias
        ctx.wg_async = _wg_async_ok(w32)
        # with a prologue the (half-size) operand is what backward needs: weight-grad input and the LeakyReLU sign mask
        ctx.save_for_backward(x if pre == PRE_NONE else xop, w32)
        return _conv_raw(xop, _weight_as(w32, operand_dtype(), False), bias, ks)

    @staticmethod
    def backward(ctx, gy):
        xs, w32 = ctx.saved_tensors
        ks, pre, slope, has_bias = ctx.cfg
        gy = _c(gy)
        gx = gw = gb = None
        want_gb = has_bias and ctx.needs_input_grad[2] and not _state["input_grads_only"]
        if want_gb and not torch.is_grad_enabled():
            # first-order pass: the bias gradient comes out of the pass that rounds gy to the bf16 operand (and goes straight into bias.grad)
            gb = _bias_grad_with_operand(gy, ctx.bias_param)
            want_gb = False
        if ctx.needs_input_grad[0]:
            gx = ConvTransposeFn.apply(gy, w32, ks)
            if pre == PRE_LRELU:
                gx = LReluBwdFn.apply(gx, xs, slope)
            elif pre == PRE_UPSAMPLE:
                gx = Pool2Fn.apply(gx, None, 1.0)
        if not _state["input_grads_only"]:
            if ctx.needs_input_grad[1]:
                if ctx.wg_async and not torch.is_grad_enabled():
                    gw = _wgrad_raw(_operand(xs), _operand(gy), ks, async_ok=True)
                else:
                    gw = WgradFn.apply(xs, gy, ks)
            if want_gb:                                # a graph of the backward is being built: the differentiable operator
                gb = ColSumFn.apply(gy)
        return gx, gw, gb, None, None, None


def conv2d(x, w32, bias, ks, pre=PRE_NONE, slope=LRELU_SLOPE):
    return Conv2dFn.apply(x, w32, bias, ks, pre, slope)


class ConvTransposeFn(Function):
    """out[n,q,ci] = sum g[n,q-t,co] w[t,co,ci]  (the conv input-gradient) = conv(g, flip(w)^T)."""

    @staticmethod
    def forward(ctx, g, w32, ks):
        g = _c(g)
        w32 = _c(w32)
        ctx.ks = ks
        ctx.save_for_backward(g, w32)
        return _conv_raw(_operand(g), _weight_as(w32, operand_dtype(), True), None, ks)

    @staticmethod
    def backward(ctx, gout):
        g, w32 = ctx.saved_tensors
        gout = _c(gout)
        gg = gw = None
        if ctx.needs_input_grad[0]:
            gg = conv2d(gout, w32, None, ctx.ks)
        if ctx.needs_input_grad[1]:
            gw = WgradFn.apply(gout, g, ctx.ks)
        return gg, gw, None


class WgradFn(Function):
    """gw[t,co,ci] = sum_{n,p} g[n,p,co] x[n,p+t,ci]  (fp32 out)."""

    @staticmethod
    def forward(ctx, x, g, ks):
        x = _c(x)
        g = _c(g)
        ctx.ks = ks
        ctx.save_for_backward(x, g)
        gw = _wgrad_raw(_operand(x), _operand(g), ks)
        return gw

    @staticmethod
    def backward(ctx, gout):
        x, g = ctx.saved_tensors
        gout = _c(gout)
        gx = gg = None
        if ctx.needs_input_grad[0]:
            gx = ConvTransposeFn.apply(g, gout, ctx.ks)
        if ctx.needs_input_grad[1]:
            gg = conv2d(x, gout, None, ctx.ks)
        return gx, gg, None


class ColSumFn(Function):
    """[..., c] -> fp32 [c] (bias gradients)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        ctx.shape = x.shape
        ctx.dtype = x.dtype
        c = x.shape[-1]
        out = _empty((c,), torch.float32, x)
        C.call("gim_colsum", C.ptr(x), C.ptr(out), x.numel() // c, c, C.dtype_code(x))
        return out

    @staticmethod
    def backward(ctx, g):
        return RowBroadcastFn.apply(g, ctx.shape, ctx.dtype)


class RowBroadcastFn(Function):
    @staticmethod
    def forward(ctx, v, shape, dtype):
        out = v.to(dtype).expand(shape).contiguous()
        return out

    @staticmethod
    def backward(ctx, g):
        return ColSumFn.apply(g), None, None


# ------------------------------------------------------------------------------------------------------------
# block-level fused operators (bf16 tensor-core path, first order)
# ------------------------------------------------------------------------------------------------------------
EPI_LRELU, EPI_MASK, EPI_ADD, EPI_POOL, EPI_ADDUP, EPI_UNIT = 1, 2, 4, 8, 16, 32


class Act:
    """An activation together with the bf16 conv operands derived from it: t32 (fp32 NHWC, what autograd tracks), tb = bf16(t32),
    tl = bf16(LeakyReLU(t32)).  Produced by the fused pooling kernel so that the next block never re-reads t32."""
    __slots__ = ("t32", "tb", "tl")

    def __init__(self, t32, tb=None, tl=None):
        self.t32, self.tb, self.tl = t32, tb, tl


def _conv_tc_fused(x, w_op, bias, ks, out_dtype, epi=0, slope=LRELU_SLOPE, mask_ref=None, addend=None, out=None):
    """tcgen05 conv with a fused epilogue.  x bf16 [n,h,w,ci] (ci % 8 == 0), w_op bf16 [taps,co,ci] (co % 16 == 0)."""
    n, h, w, ci = x.shape
    taps, co, _ = w_op.shape
    y = out if out is not None else _empty((n, h, w, co), out_dtype, x)
    _timed_call("tcgen05" if _profile is not None else None, 2.0 * n * h * w * ci * co * taps, "gim_conv2d_fwd_fused", C.ptr(x), C.ptr(w_op), C.ptr(bias),
                C.ptr(y), C.ptr(mask_ref), C.ptr(addend), n, h, w, ci, co, ks, C.dtype_code(y), epi, slope)
    return y


def _skinny_in(xop, w_t, ks):
    """A conv whose input has a few channels (images) as a dense 1x1 GEMM over the tap-unrolled input: -> (xcol, w2 [1,co,kc])."""
    taps, co, ci = w_t.shape
    kc = _round_up(taps * ci, 8)
    w2 = torch.zeros((1, co, kc), dtype=w_t.dtype, device=w_t.device)
    w2[0, :, :taps * ci] = w_t.permute(1, 0, 2).reshape(co, taps * ci)
    return _im2col(xop, ks, 1, kc), w2


def _unskinny_gw(gw2, taps, co, ci):
    return gw2[0, :, :taps * ci].reshape(co, taps, ci).permute(1, 0, 2).contiguous()


def _cast_colsum_ok(g):
    c = g.shape[-1]
    return (_state["operand_dtype"] == torch.bfloat16 and _state["conv_algo"] != C.ALGO_SIMT and g.dtype == torch.float32 and c % 8 == 0 and c <= 2048
            and 256 % (c // 8) == 0 and (c // 8) & (c // 8 - 1) == 0 and g.data_ptr() % 16 == 0)


def _bias_grad_with_operand(g, param):
    """Bias gradient of a convolution whose output gradient `g` (fp32) is also about to be rounded to the bf16 operand of the input- and
    weight-gradient convolutions: one pass writes the operand copy (registered, so _operand(g) finds it) and the column sums -- straight
    into the leaf's .grad inside `deferred_weight_grads()`, like _bias_grad."""
    if not _cast_colsum_ok(g):
        return _bias_grad(g, param)
    hit = _cast_memo.get(id(g))
    if hit is not None and hit[0]() is g and hit[1] == g._version:
        return _bias_grad(g, param)                # operand copy already exists
    c = g.shape[-1]
    op = torch.empty(g.shape, dtype=torch.bfloat16, device=g.device)
    direct = (_state["defer_sn"] and param is not None and param.is_leaf and param.requires_grad and param.dtype == torch.float32 and param.is_contiguous()
              and param.grad is not None and param.grad.is_contiguous())
    out = param.grad if direct else _empty((c,), torch.float32, g)
    C.call("gim_cast_colsum", C.ptr(g), C.ptr(op), C.ptr(out), g.numel() // c, c, 1 if direct else 0)
    register_operand(g, op)
    return None if direct else out


def _bias_grad(x, param):
    """Bias gradient sum_rows x.  Inside `deferred_weight_grads()` (i.e. under loss.backward()) it is accumulated straight into the
    leaf's .grad -- no memset, no AccumulateGrad add -- and None is returned to autograd; otherwise the column sums are returned."""
    if (_state["defer_sn"] and param is not None and param.is_leaf and param.requires_grad and param.dtype == torch.float32 and param.is_contiguous()
            and param.grad is not None and param.grad.is_contiguous()):
        c = x.shape[-1]
        C.call("gim_colsum_acc", C.ptr(x), C.ptr(param.grad), x.numel() // c, c, C.dtype_code(x))
        return None
    return _colsum(x)


class CatParamsFn(Function):
    """torch.cat(tensors, 0) of parameter-like tensors.  Inside `deferred_weight_grads()` the backward adds the row blocks of the gradient
    straight into the leaves' .grad with one multi-tensor launch (instead of one AccumulateGrad add per tensor) and returns None."""

    @staticmethod
    def forward(ctx, *tensors):
        ctx.tensors = tensors
        return torch.cat(tensors, 0)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        ts = ctx.tensors
        parts = torch.split(_c(g), [t.shape[0] for t in ts], 0)
        want = [i for i, t in enumerate(ts) if ctx.needs_input_grad[i]]
        if _state["defer_sn"] and want and all(ts[i].is_leaf and ts[i].grad is not None and ts[i].grad.is_contiguous()
                                                and ts[i].grad.dtype == parts[i].dtype for i in want):
            torch._foreach_add_([ts[i].grad for i in want], [parts[i] for i in want])
            return (None,) * len(ts)
        return tuple(parts[i] if ctx.needs_input_grad[i] else None for i in range(len(ts)))


def cat_params(tensors):
    return CatParamsFn.apply(*tensors)


def _colsum(x):
    c = x.shape[-1]
    out = _empty((c,), torch.float32, x)
    C.call("gim_colsum", C.ptr(x), C.ptr(out), x.numel() // c, c, C.dtype_code(x))
    return out


class ResBlockDownFn(Function):
    """AvgPool2(conv1x1(x)) + AvgPool2(conv_k(lrelu(conv_k(lrelu(x)))))  (reference model_blocks.py:486-514) as ONE autograd node:
    * the 1x1 residual convolution commutes with the pooling, AvgPool2(conv1x1(x)) == conv1x1(AvgPool2(x)), so it runs at the pooled
      resolution on both passes (4x fewer FLOPs and bytes; the same identity the up-sampling blocks use);
    * the activation between the two k x k convs only exists as the bf16 LeakyReLU'd operand written by the first conv's epilogue;
    * AvgPool + residual add happen inside the second conv's epilogue, the pooled output is written once;
    * in the backward the AvgPool gradient, the LeakyReLU masks and the residual-gradient sum are folded into the producing kernels'
      epilogues.
    First order only (see composite_mode).  Odd spatial sizes use the unfused pooling kernels."""

    @staticmethod
    def forward(ctx, x32, xb, xl, wl, bl, w1, b1, w2, b2, ks, slope, want_ops):
        od = torch.bfloat16
        x32 = _c(x32)
        n, h, w, ci = x32.shape
        skinny = ci % 8 != 0
        even = h % 2 == 0 and w % 2 == 0
        wl, w1, w2 = _c(wl), _c(w1), _c(w2)
        co = w2.shape[1]
        lazy = skinny and even and ci * ks * ks <= 64 and co % 32 == 0
        if lazy:
            # image-side block: both input convolutions are K <= 64 outer products -> one HBM-bound kernel, no im2col, no operand tensors
            y32 = _empty((n, h // 2, w // 2, co), torch.float32, x32)
            tl = _empty((n, h, w, co), od, x32)
            C.call("gim_first_block_fwd", C.ptr(x32), C.ptr(w1), C.ptr(b1), C.ptr(wl), C.ptr(bl), C.ptr(tl), C.ptr(y32), n, h, w, ci, co, ks, slope)
            _conv_tc_fused(tl, _weight_as(w2, od, False), b2, ks, torch.float32, EPI_POOL | EPI_ADD, addend=y32, out=y32)
            yb = yl = None
            if want_ops:
                yb = torch.empty_like(y32, dtype=od)
                yl = torch.empty_like(y32, dtype=od)
                C.call("gim_cast", C.ptr(y32), C.F32, C.ptr(yb), C.BF16, y32.numel())
                C.call("gim_operand_prepare", C.ptr(y32), C.F32, C.ptr(yl), C.BF16, n, h // 2, w // 2, co, PRE_LRELU, slope)
            ctx.cfg = (ks, slope, skinny, even, (n, h, w, ci, co), want_ops)
            ctx.wg_async = (False, False, _wg_async_ok(w2))
            ctx.lazy = True
            ctx.save_for_backward(x32, None, None, tl, wl, w1, w2)
            ctx.biases = (bl, b1, b2)
            if want_ops:
                ctx.mark_non_differentiable(yb, yl)
                return y32, yb, yl
            return y32, None, None
        ctx.lazy = False
        if xl is None:
            xl = _prepare_operand(x32, PRE_LRELU, slope)
        wl_t, w1_t, w2_t = _weight_as(wl, od, False), _weight_as(w1, od, False), _weight_as(w2, od, False)
        if skinny:
            xr, w1_e = _skinny_in(xl, w1_t, ks)
            k1 = 1
        else:
            xr, w1_e, k1 = xl, w1_t, ks
        od_out = (n, h // 2, w // 2, co)
        if even:
            # residual branch at the pooled resolution: xp = bf16(AvgPool2(x))
            if ci % 4 == 0:
                xp = _empty((n, h // 2, w // 2, ci), od, x32)
                C.call("gim_pool2_multi", C.ptr(x32), None, None, C.ptr(xp), None, n, h, w, ci, 0.25, slope)
            else:
                xp32 = _empty((n, h // 2, w // 2, ci), torch.float32, x32)
                C.call("gim_pool2_sum", C.ptr(x32), None, C.ptr(xp32), n, h, w, ci, 0.25, C.F32)
                xp = _operand(xp32)
            xa, wl_e = _skinny_in(xp, wl_t, 1) if skinny else (xp, wl_t)
            y32 = _conv_tc_fused(xa, wl_e, bl, 1, torch.float32)
            tl = _conv_tc_fused(xr, w1_e, b1, k1, od, EPI_LRELU, slope)
            _conv_tc_fused(tl, w2_t, b2, ks, torch.float32, EPI_POOL | EPI_ADD, addend=y32, out=y32)      # AvgPool + residual in the epilogue
            yb = yl = None
            if want_ops:
                yb = torch.empty_like(y32, dtype=od)
                yl = torch.empty_like(y32, dtype=od)
                C.call("gim_cast", C.ptr(y32), C.F32, C.ptr(yb), C.BF16, y32.numel())
                C.call("gim_operand_prepare", C.ptr(y32), C.F32, C.ptr(yl), C.BF16, n, h // 2, w // 2, co, PRE_LRELU, slope)
        else:
            if xb is None:
                xb = _operand(x32)
            xa, wl_e = _skinny_in(xb, wl_t, 1) if skinny else (xb, wl_t)
            res = _conv_tc_fused(xa, wl_e, bl, 1, torch.float32)
            tl = _conv_tc_fused(xr, w1_e, b1, k1, od, EPI_LRELU, slope)
            o = _conv_tc_fused(tl, w2_t, b2, ks, torch.float32)
            y32 = _empty(od_out, torch.float32, o)
            yb = torch.empty_like(y32, dtype=od) if want_ops else None
            yl = torch.empty_like(y32, dtype=od) if want_ops else None
            C.call("gim_pool2_multi", C.ptr(res), C.ptr(o), C.ptr(y32), C.ptr(yb), C.ptr(yl), n, h, w, co, 0.25, slope)
        ctx.cfg = (ks, slope, skinny, even, (n, h, w, ci, co), want_ops)
        ctx.wg_async = (_wg_async_ok(wl), _wg_async_ok(w1), _wg_async_ok(w2))
        ctx.save_for_backward(xa, xr, xl, tl, wl, w1, w2)
        ctx.biases = (bl, b1, b2)
        if want_ops:
            ctx.mark_non_differentiable(yb, yl)
            return y32, yb, yl
        return y32, None, None

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, _gb, _gl):
        xa, xr, xl, tl, wl, w1, w2 = ctx.saved_tensors
        ks, slope, skinny, even, (n, h, w, ci, co), _ = ctx.cfg
        od = torch.bfloat16
        taps = ks * ks
        gy = _c(gy)
        fused_wgrad = False
        if ctx.lazy:
            # the forward ran the fused image-side kernel: build the tensor-core operands only if a gradient that needs them is requested
            x32 = xa
            xa = xr = xl = None
            want_any_w = (not _state["input_grads_only"]) and (ctx.needs_input_grad[3] or ctx.needs_input_grad[5])
            fused_wgrad = want_any_w and ks == 3 and ci in (1, 3) and co % 8 == 0 and 256 % co == 0 and h % 4 == 0 and not _state["deterministic"]
            if (want_any_w and not fused_wgrad) or ctx.needs_input_grad[0]:
                xl = _prepare_operand(x32, PRE_LRELU, slope)
            if want_any_w and not fused_wgrad:
                xp32 = _empty((n, h // 2, w // 2, ci), torch.float32, x32)
                C.call("gim_pool2_sum", C.ptr(x32), None, C.ptr(xp32), n, h, w, ci, 0.25, C.F32)
                xa = _im2col(_operand(xp32), 1, 1, _round_up(ci, 8))
                xr = _im2col(xl, ks, 1, _round_up(taps * ci, 8))
        g = _empty((n, h, w, co), od, gy)                      # AvgPool backward, written once as the bf16 operand
        C.call("gim_unpool2_cast", C.ptr(gy), C.ptr(g), n, h, w, co, 0.25)
        need_gl = ctx.needs_input_grad[0] or ((not _state["input_grads_only"]) and ctx.needs_input_grad[3] and not fused_wgrad)
        # gradient of the residual branch: at the pooled resolution when it ran there
        gl = (_operand(gy) if even else g) if need_gl else None
        gt = _conv_tc_fused(g, _weight_as(w2, od, True), None, ks, od, EPI_MASK, slope, mask_ref=tl)      # d/d(conv_r1 output), masked
        gwl = gbl = gw1 = gb1 = gw2 = gb2 = None
        want_w = not _state["input_grads_only"]
        if fused_wgrad:
            # both image-side weight gradients in one pass over the gradients (K <= 27 outer products: no im2col, no tensor-core launch)
            gw1 = _empty((taps, co, ci), torch.float32, gy)
            gwl = _empty((1, co, ci), torch.float32, gy)
            scratch = _empty((32 * 10 * co * ci,), torch.float32, gy)        # per-CTA-group copies of the two results (see gim_b200.h)
            C.call("gim_first_block_wgrad", C.ptr(x32), C.ptr(gt), C.ptr(gy), C.ptr(gw1), C.ptr(gwl), C.ptr(scratch), scratch.numel(), n, h, w, ci, co, ks,
                   slope)
            if not ctx.needs_input_grad[3]:
                gwl = None
            if not ctx.needs_input_grad[5]:
                gw1 = None
        if want_w and ctx.needs_input_grad[3] and not fused_wgrad:
            gwl = _wgrad_raw(xa, gl, 1, async_ok=ctx.wg_async[0] and not skinny)
            if skinny:
                gwl = _unskinny_gw(gwl, 1, co, ci)
        if want_w and ctx.needs_input_grad[7]:
            gw2 = _wgrad_raw(tl, g, ks, async_ok=ctx.wg_async[2])
        if want_w and ctx.needs_input_grad[5] and not fused_wgrad:
            gw1 = _wgrad_raw(xr, gt, 1 if skinny else ks, async_ok=ctx.wg_async[1] and not skinny)
            if skinny:
                gw1 = _unskinny_gw(gw1, taps, co, ci)
        bl_p, b1_p, b2_p = ctx.biases
        if want_w and ctx.needs_input_grad[4]:
            gbl = _bias_grad(gy, bl_p)                             # both biases see the same gradient; sum(unpool(gy)/4) == sum(gy), in fp32
        if want_w and ctx.needs_input_grad[8]:
            gb2 = _bias_grad(gy, b2_p)
        if want_w and ctx.needs_input_grad[6]:
            gb1 = _bias_grad(gt, b1_p)
        gx = None
        if ctx.needs_input_grad[0]:
            if not skinny and ci % 32 == 0:
                gx1 = _conv_tc_fused(gl, _weight_as(wl, od, True), None, 1, torch.float32)          # pooled resolution when `even`
                if even:
                    gx = _conv_tc_fused(gt, _weight_as(w1, od, True), None, ks, torch.float32, EPI_MASK | EPI_ADDUP, slope, mask_ref=xl, addend=gx1)
                else:
                    gx = _conv_tc_fused(gt, _weight_as(w1, od, True), None, ks, torch.float32, EPI_MASK | EPI_ADD, slope, mask_ref=xl, addend=gx1, out=gx1)
            else:                                                  # image-side block: few channels, elementary kernels
                gx = _conv_raw(gl, _weight_as(wl, od, True), None, 1)
                if even:
                    up = _empty((n, h, w, ci), torch.float32, gx)
                    C.call("gim_unpool2_bcast", C.ptr(gx), C.ptr(up), n, h, w, ci, 0.25, C.F32)
                    gx = up
                g2 = _conv_raw(gt, _weight_as(w1, od, True), None, ks)
                gm = torch.empty_like(g2)
                C.call("gim_lrelu_bwd_ref", C.ptr(g2), C.ptr(xl), C.dtype_code(xl), C.ptr(gm), g2.numel(), slope)
                C.call("gim_axpby", C.ptr(gx), C.ptr(gm), C.ptr(gx), gx.numel(), 1.0, 1.0, C.F32)
        return gx, None, None, gwl, gbl, gw1, gb1, gw2, gb2, None, None, None


def res_block_down(x, wl, bl, w1, b1, w2, b2, ks, slope=LRELU_SLOPE, want_ops=True):
    """x: Act or fp32 NHWC tensor -> Act."""
    if not isinstance(x, Act):
        x = Act(x)
    y32, yb, yl = ResBlockDownFn.apply(x.t32, x.tb, x.tl, wl, bl, w1, b1, w2, b2, ks, slope, want_ops)
    return Act(y32, yb, yl)


# ------------------------------------------------------------------------------------------------------------
# spectral norm
# ------------------------------------------------------------------------------------------------------------
class SpectralNormFn(Function):
    """weight_orig [co,ci,k,k] -> W/sigma packed fp32 [k*k,co,ci]; one power iteration in place on (u, v) when training.
    torch.nn.utils.spectral_norm semantics (n_power_iterations=1, eps=1e-12, dim=0), u/v constants in the backward."""

    @staticmethod
    def forward(ctx, weight_orig, uv, training, eps):
        u, v = uv                                   # buffers, passed in a tuple so autograd ignores them
        co, ci, k, _ = weight_orig.shape
        w = _c(weight_orig)
        w_sn = _empty((k * k, co, ci), torch.float32, w)
        aux = _empty((co + ci * k * k + 1,), torch.float32, w)      # u_used | v_used | sigma
        scratch = _empty((co + ci * k * k + 8,), torch.float32, w)
        u_used, v_used, sigma = aux[:co], aux[co:co + ci * k * k], aux[co + ci * k * k:]
        C.call("gim_sn_forward", C.ptr(w), C.ptr(u), C.ptr(v), 1 if training else 0, eps, C.ptr(w_sn), sigma.data_ptr(),
               u_used.data_ptr(), v_used.data_ptr(), C.ptr(scratch), co, ci, k)
        ctx.save_for_backward(w, aux)
        ctx.dims = (co, ci, k)
        ctx.param = weight_orig
        return w_sn

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        w, aux = ctx.saved_tensors
        co, ci, k = ctx.dims
        g = _c(g)
        if _defer_sn_backward(ctx, g, w, aux):
            return None, None, None, None
        _wg_join()
        gw = torch.empty_like(w)
        scratch = _empty((8,), torch.float32, w)
        j = ci * k * k
        C.call("gim_sn_backward", C.ptr(g), C.ptr(w), aux[:co].data_ptr(), aux[co:co + j].data_ptr(), aux[co + j:].data_ptr(),
               C.ptr(gw), C.ptr(scratch), co, ci, k)
        return gw, None, None, None


# ---- deferred, batched spectral-norm backward ---------------------------------------------------------------------
_sn_pending = []


@contextlib.contextmanager
def deferred_weight_grads():
    """Wrap `loss.backward()`: inside, the spectral-norm backward of every convolution is not launched layer by layer (2 kernels + a
    memset + an AccumulateGrad add each) but collected and, when the backward pass ends, run for all layers at once
    (gim_sn_backward_multi), accumulating straight into `weight_orig.grad`.  Only valid for `.backward()` (which accumulates into
    .grad); `torch.autograd.grad` w.r.t. a weight_orig must not be used inside."""
    old = _state["defer_sn"]
    _state["defer_sn"] = True
    try:
        yield
    finally:
        _state["defer_sn"] = old
        _flush_sn_backward()
        if _side["used"] and _side["stream"] is not None:
            # kernels of the side branch's backward wrote into .grad directly (bias sums): the optimizer on this stream must see them
            torch.cuda.current_stream().wait_stream(_side["stream"])
        _side["used"] = False          # (the next forward that forks sets it again; a model without branches never touches the side stream)


def _defer_sn_backward(ctx, g, w, aux):
    prm = getattr(ctx, "param", None)
    if not _state["defer_sn"] or prm is None or not prm.is_leaf or not prm.requires_grad or not prm.is_contiguous():
        return False
    if not _sn_pending:
        torch.autograd.Variable._execution_engine.queue_callback(_flush_sn_backward)
    _sn_pending.append((prm, g, aux, ctx.dims))
    return True


def _flush_sn_backward():
    import ctypes
    pending, seen_round = list(_sn_pending), {}
    del _sn_pending[:]
    _wg_join()                       # weight gradients computed on their own streams
    if not pending:
        return
    rounds = []                      # two gradients of the same parameter (D's encoders run three times per step) go to separate launches
    for item in pending:
        r = seen_round.get(id(item[0]), 0)
        seen_round[id(item[0])] = r + 1
        while len(rounds) <= r:
            rounds.append([])
        rounds[r].append(item)
    if _side["used"] and _side["stream"] is not None:
        # gradients produced by the side branch's backward: order this stream after it, and keep their memory until the flush is done
        cur = torch.cuda.current_stream()
        cur.wait_stream(_side["stream"])
        for _, g, aux, _ in pending:
            g.record_stream(cur)
            aux.record_stream(cur)
    with torch.no_grad():
        for items in rounds:
            scratch = torch.empty(len(items), dtype=torch.float32, device=items[0][1].device)
            table = (C.SnBwdLayer * len(items))()
            for i, (prm, g, aux, (co, ci, k)) in enumerate(items):
                acc = 1
                if prm.grad is None:
                    prm.grad = torch.empty_like(prm)
                    acc = 0
                j = ci * k * k
                e = table[i]
                e.g, e.w, e.u, e.v, e.sigma = g.data_ptr(), prm.data_ptr(), aux.data_ptr(), aux[co:].data_ptr(), aux[co + j:].data_ptr()
                e.grad, e.scratch = prm.grad.data_ptr(), scratch[i:].data_ptr()
                e.cout, e.cin, e.ksize, e.accumulate = co, ci, k, acc
            C.call("gim_sn_backward_multi", ctypes.cast(table, ctypes.c_void_p), len(items))


def sn_prepare(convs, training, eps):
    """Spectral normalisation of a whole list of SNConv2d modules with one launch per phase (gim_sn_forward_multi) instead of eight
    launches per convolution: power iteration on (u, v) in place when training, W/sigma packed [k*k, co, ci] in fp32 plus -- on the
    bf16 path -- the bf16 operand and its flipped/transposed twin for the input-gradient convolution.  Each module keeps its result
    until its next effective_weight() call, which wraps it in the per-layer autograd node."""
    if not convs:
        return
    dev = convs[0].weight_orig.device
    bf = _state["operand_dtype"] == torch.bfloat16 and _state["conv_algo"] != C.ALGO_SIMT
    # buffer layout: every packed W/sigma first, in module order (so the weights of consecutive layers are adjacent -- the attention
    # block reads its three 1x1 projections as ONE [1, 2c/8 + c, c] weight, see merged_sn_weight), then the per-layer vectors
    dims = [(m.out_channels, m.in_channels, m.kernel_size) for m in convs]
    tots = [co * ci * k * k for co, ci, k in dims]
    n32 = sum(_round_up(t, 4) for t in tots)
    n_op = sum(_round_up(t, 8) for t in tots)
    nbf = n_op
    o_w = o_b = 0
    plan = []
    for m, (co, ci, k), tot in zip(convs, dims, tots):
        j = ci * k * k
        o_sn, o_aux, o_scr = o_w, n32, n32 + _round_up(co + j + 1, 4)
        n32 = o_scr + _round_up(j + co, 4)
        o_op, o_fl = o_b, nbf
        nbf = o_fl + _round_up(tot, 8)
        o_w += _round_up(tot, 4)
        o_b += _round_up(tot, 8)
        plan.append((m, co, ci, k, tot, j, o_sn, o_aux, o_scr, o_op, o_fl))
    f32 = torch.empty(n32, dtype=torch.float32, device=dev)
    b16 = torch.empty(nbf, dtype=torch.bfloat16, device=dev) if bf else None
    table = (C.SnLayer * len(plan))()
    for i, (m, co, ci, k, tot, j, o_sn, o_aux, o_scr, o_op, o_fl) in enumerate(plan):
        w = m.weight_orig
        if not w.is_contiguous():
            raise RuntimeError("sn_prepare: weight_orig must be contiguous")
        w_sn = f32[o_sn:o_sn + tot].view(k * k, co, ci)
        aux = f32[o_aux:o_aux + co + j + 1]
        w_op = b16[o_op:o_op + tot].view(k * k, co, ci) if bf else None
        w_fl = b16[o_fl:o_fl + tot].view(k * k, ci, co) if bf else None
        e = table[i]
        e.w, e.u, e.v = w.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr()
        e.w_sn, e.aux, e.scratch = w_sn.data_ptr(), aux.data_ptr(), f32[o_scr:].data_ptr()
        e.w_op = w_op.data_ptr() if bf else None
        e.w_flip = w_fl.data_ptr() if bf else None
        e.cout, e.cin, e.ksize, e.reserved = co, ci, k, 0
        m._prepared = (w_sn, aux, w_op, w_fl, w._version, (f32, o_sn, b16, o_op))
    import ctypes
    C.call("gim_sn_forward_multi", ctypes.cast(table, ctypes.c_void_p), len(plan), 1 if training else 0, eps)


class SpectralNormPreparedFn(Function):
    """The per-layer autograd node over a weight that sn_prepare already normalised (same backward as SpectralNormFn)."""

    @staticmethod
    def forward(ctx, weight_orig, prepared):
        w_sn, aux = prepared[0], prepared[1]
        co, ci, k, _ = weight_orig.shape
        ctx.save_for_backward(weight_orig, aux)
        ctx.dims = (co, ci, k)
        ctx.param = weight_orig
        return w_sn

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return SpectralNormFn.backward(ctx, g)[0], None


def sn_prepared_weight(weight_orig, prepared):
    w = SpectralNormPreparedFn.apply(weight_orig, prepared)
    if prepared[2] is not None:
        if len(_wops) > 256:
            for key in [k for k, v in _wops.items() if v[0]() is None]:
                del _wops[key]
        _wops[id(w)] = (weakref.ref(w), prepared[2], prepared[3])
    return w


class _MergedRowsFn(Function):
    """[1, co_1 + co_2 + .., ci] weight over 1x1 weights that already lie one after the other in the batched spectral-norm buffer:
    no copy forward, row blocks of the gradient backward."""

    @staticmethod
    def forward(ctx, holder, *ws):
        f32, off = holder
        ctx.rows = [w.shape[1] for w in ws]
        ci = ws[0].shape[2]
        return f32[off:off + sum(ctx.rows) * ci].view(1, sum(ctx.rows), ci)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return (None,) + tuple(torch.split(_c(g), ctx.rows, 1))


def merged_sn_weight(convs):
    """Packed fp32 weight [1, sum(cout), cin] of several spectral-normalised 1x1 convolutions over the same input, so that they run as
    ONE convolution (and one input-gradient / weight-gradient launch).  After a batched sn_prepare the per-layer results are adjacent in
    one buffer and the merged weight (and its bf16 operand) is just a view; otherwise the rows are concatenated."""
    preps = [m._prepared for m in convs]
    ws = [m.effective_weight() for m in convs]
    if any(w.shape[0] != 1 for w in ws):
        raise RuntimeError("merged_sn_weight: 1x1 convolutions only")
    ok = all(p is not None and len(p) > 5 and p[0] is not None for p in preps)
    if ok:
        f32, off0, b16, op0 = preps[0][5]
        off, op = off0, op0
        for p, w in zip(preps, ws):
            ok = ok and p[5][0] is f32 and p[5][1] == off and (b16 is None or p[5][3] == op) and p[0].data_ptr() == w.data_ptr()
            off += w.numel()
            op += w.numel()
    if not ok:
        return torch.cat(ws, dim=1)
    merged = _MergedRowsFn.apply((f32, off0), *ws)
    if b16 is not None:
        _wops[id(merged)] = (weakref.ref(merged), b16[op0:op0 + merged.numel()].view(merged.shape), None)
    return merged


class AttentionPackedFn(Function):
    """AttentionCoreFn over ONE projection tensor kqv [n, 64, 2d + c] = [keys | queries | values] (d = c/8): the kernels read the three
    column blocks in place and the backward writes their gradients into one tensor of the same layout."""

    @staticmethod
    def forward(ctx, kqv, x, gamma):
        kqv, x = _c(kqv), _c(x)
        n, p, c = x.shape
        d, ld = c // 8, kqv.shape[2]
        attn = _empty((n, p, p), torch.float32, x)
        y = torch.empty_like(x)
        base = C.ptr(kqv)
        C.call("gim_attention_fwd", base + 4 * d, base, base + 8 * d, ld, ld, C.ptr(x), C.ptr(gamma), C.ptr(attn), C.ptr(y), n, p, c)
        ctx.save_for_backward(kqv, attn, gamma)
        ctx.dims = (n, p, c)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        kqv, attn, gamma = ctx.saved_tensors
        n, p, c = ctx.dims
        d, ld = c // 8, kqv.shape[2]
        gy = _c(gy)
        dkqv = torch.empty_like(kqv)
        part = _empty((n,), torch.float32, kqv)
        base, dbase = C.ptr(kqv), C.ptr(dkqv)
        C.call("gim_attention_bwd", C.ptr(gy), base + 4 * d, base, base + 8 * d, ld, ld, C.ptr(attn), C.ptr(gamma),
               dbase + 4 * d, dbase, dbase + 8 * d, C.ptr(part), n, p, c)
        dgamma = part.sum().reshape(gamma.shape) if ctx.needs_input_grad[2] else None
        return dkqv, gy.view(n, p, c), dgamma


# ------------------------------------------------------------------------------------------------------------
# pointwise / resampling / layout
# ------------------------------------------------------------------------------------------------------------
class LReluFn(Function):
    @staticmethod
    def forward(ctx, x, slope):
        x = _c(x)
        y = torch.empty_like(x)
        C.call("gim_lrelu_fwd", C.ptr(x), C.ptr(y), x.numel(), slope, C.dtype_code(x))
        ctx.slope = slope
        ctx.save_for_backward(y)          # sign(y) == sign(x): keep the tensor the consumer keeps anyway
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        return LReluBwdFn.apply(gy, y.detach(), ctx.slope), None


class LReluBwdFn(Function):
    """gx = g * (ref > 0 ? 1 : slope); piecewise linear => zero second derivative w.r.t. ref (ref may be the saved bf16 operand)."""

    @staticmethod
    def forward(ctx, g, ref, slope):
        g = _c(g)
        gx = torch.empty_like(g)
        if g.dtype == torch.float32 and ref.dtype != g.dtype:
            C.call("gim_lrelu_bwd_ref", C.ptr(g), C.ptr(ref), C.dtype_code(ref), C.ptr(gx), g.numel(), slope)
        else:
            C.call("gim_lrelu_bwd", C.ptr(g), C.ptr(ref), C.ptr(gx), g.numel(), slope, C.dtype_code(g))
        ctx.slope = slope
        ctx.save_for_backward(ref)
        return gx

    @staticmethod
    def backward(ctx, gg):
        (ref,) = ctx.saved_tensors
        return LReluBwdFn.apply(gg, ref, ctx.slope), None, None


def lrelu(x, slope=LRELU_SLOPE):
    return LReluFn.apply(x, slope)


class TanhFn(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        C.call("gim_tanh_fwd", C.ptr(x), C.ptr(y), x.numel(), C.dtype_code(x))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        gy = _c(gy)
        gx = torch.empty_like(gy)
        C.call("gim_tanh_bwd", C.ptr(gy), C.ptr(y), C.ptr(gx), gy.numel(), C.dtype_code(gy))
        return gx


class AddFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        a = _c(a)
        b = _c(b)
        out = torch.empty_like(a)
        C.call("gim_axpby", C.ptr(a), C.ptr(b), C.ptr(out), a.numel(), 1.0, 1.0, C.dtype_code(a))
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


class ScaleDevFn(Function):
    """y = s * x with s a 1-element fp32 device tensor (SelfAttention.gamma, model_blocks.py:548)."""

    @staticmethod
    def forward(ctx, x, s):
        x = _c(x)
        y = torch.empty_like(x)
        C.call("gim_scale_dev", C.ptr(x), C.ptr(s), C.ptr(y), x.numel(), C.dtype_code(x))
        ctx.save_for_backward(x, s)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, s = ctx.saved_tensors
        gx = ScaleDevFn.apply(gy, s) if ctx.needs_input_grad[0] else None
        gs = DotFn.apply(gy, x).reshape(s.shape) if ctx.needs_input_grad[1] else None
        return gx, gs


class DotFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        a = _c(a)
        b = _c(b)
        out = _empty((1,), torch.float32, a)
        C.call("gim_dot", C.ptr(a), C.ptr(b), C.ptr(out), a.numel(), C.dtype_code(a))
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = _c(g.reshape(1).float())
        ga = ScaleDevFn.apply(b, g) if ctx.needs_input_grad[0] else None
        gb = ScaleDevFn.apply(a, g) if ctx.needs_input_grad[1] else None
        return ga, gb


class Pool2Fn(Function):
    """y = scale * sum over 2x2 windows of (a [+ b]); AvgPool2d(2) with the residual add folded in (model_blocks.py:497-514)."""

    @staticmethod
    def forward(ctx, a, b, scale):
        a = _c(a)
        n, h, w, c = a.shape
        if b is not None:
            b = _c(b)
        y = _empty((n, h // 2, w // 2, c), a.dtype, a)
        C.call("gim_pool2_sum", C.ptr(a), C.ptr(b), C.ptr(y), n, h, w, c, scale, C.dtype_code(a))
        ctx.hw = (h, w)
        ctx.scale = scale
        ctx.two = b is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        g = Unpool2Fn.apply(gy, ctx.hw[0], ctx.hw[1], ctx.scale)
        return g, (g if ctx.two else None), None


class Unpool2Fn(Function):
    """gx[n,h,w,c] = scale * g[n,h//2,w//2,c]: nearest Upsample(x2) (scale 1) and the AvgPool backward (scale 1/4)."""

    @staticmethod
    def forward(ctx, g, h, w, scale):
        g = _c(g)
        n, _, _, c = g.shape
        gx = _empty((n, h, w, c), g.dtype, g)
        C.call("gim_unpool2_bcast", C.ptr(g), C.ptr(gx), n, h, w, c, scale, C.dtype_code(g))
        ctx.scale = scale
        return gx

    @staticmethod
    def backward(ctx, gg):
        return Pool2Fn.apply(gg, None, ctx.scale), None, None, None


def avg_pool2_add(a, b=None):
    return Pool2Fn.apply(a, b, 0.25)


def upsample2(x):
    return Unpool2Fn.apply(x, 2 * x.shape[1], 2 * x.shape[2], 1.0)


class ToNHWCFn(Function):
    """NCHW fp32 (the reference's tensor layout) -> NHWC activation dtype."""

    @staticmethod
    def forward(ctx, x, dtype):
        x = _c(x.float())
        n, c, h, w = x.shape
        y = _empty((n, h, w, c), dtype, x)
        C.call("gim_nchw_to_nhwc", C.ptr(x), C.ptr(y), n, c, h, w, C.dtype_code(y))
        return y

    @staticmethod
    def backward(ctx, g):
        return FromNHWCFn.apply(g), None


class FromNHWCFn(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        n, h, w, c = x.shape
        ctx.dtype = x.dtype
        y = _empty((n, c, h, w), torch.float32, x)
        C.call("gim_nhwc_to_nchw", C.ptr(x), C.ptr(y), n, c, h, w, C.dtype_code(x))
        return y

    @staticmethod
    def backward(ctx, g):
        return ToNHWCFn.apply(g, ctx.dtype)


def to_nhwc(x):
    return ToNHWCFn.apply(x, torch.float32)


def from_nhwc(x):
    return FromNHWCFn.apply(x)


class CatChannelsFn(Function):
    """torch.cat((a, b), dim=channel) on NHWC tensors (gim_img_models.py:385)."""

    @staticmethod
    def forward(ctx, a, b):
        a = _c(a)
        b = _c(b)
        ca, cb = a.shape[-1], b.shape[-1]
        out = _empty(a.shape[:-1] + (ca + cb,), a.dtype, a)
        rows = a.numel() // ca
        C.call("gim_copy_cols", C.ptr(a), ca, 0, C.ptr(out), ca + cb, 0, rows, ca, C.dtype_code(a))
        C.call("gim_copy_cols", C.ptr(b), cb, 0, C.ptr(out), ca + cb, ca, rows, cb, C.dtype_code(a))
        ctx.split = (ca, cb)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _c(g)
        ca, cb = ctx.split
        rows = g.numel() // (ca + cb)
        ga = _empty(g.shape[:-1] + (ca,), g.dtype, g)
        gb = _empty(g.shape[:-1] + (cb,), g.dtype, g)
        C.call("gim_copy_cols", C.ptr(g), ca + cb, 0, C.ptr(ga), ca, 0, rows, ca, C.dtype_code(g))
        C.call("gim_copy_cols", C.ptr(g), ca + cb, ca, C.ptr(gb), cb, 0, rows, cb, C.dtype_code(g))
        return ga, gb


class ImgAttBlendFn(Function):
    """ImgAttention's blend (reference model_blocks.py:596-608) on NHWC fp32 tensors [n, h, w, c]: per pixel the two channel dot products
    <q1, k1>, <q2, k2>, their 2-way softmax and out = a1 * x1 + a2 * v2 -- one kernel forward, one backward (first order: attacker only)."""

    @staticmethod
    def forward(ctx, q1, k1, q2, k2, x1, v2):
        q1, k1, q2, k2, x1, v2 = (_c(t) for t in (q1, k1, q2, k2, x1, v2))
        c = x1.shape[-1]
        pixels = x1.numel() // c
        out = torch.empty_like(x1)
        att = _empty(x1.shape[:-1] + (2,), torch.float32, x1)
        C.call("gim_img_att_blend_fwd", C.ptr(q1), C.ptr(k1), C.ptr(q2), C.ptr(k2), C.ptr(x1), C.ptr(v2), C.ptr(out), C.ptr(att), pixels, c)
        ctx.save_for_backward(q1, k1, q2, k2, x1, v2, att)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        q1, k1, q2, k2, x1, v2, att = ctx.saved_tensors
        g = _c(g)
        c = x1.shape[-1]
        gq1, gk1, gq2, gk2, gv2 = (torch.empty_like(x1) for _ in range(5))
        gx1 = torch.empty_like(x1) if ctx.needs_input_grad[4] else None
        C.call("gim_img_att_blend_bwd", C.ptr(g), C.ptr(q1), C.ptr(k1), C.ptr(q2), C.ptr(k2), C.ptr(x1), C.ptr(v2), C.ptr(att),
               C.ptr(gq1), C.ptr(gk1), C.ptr(gq2), C.ptr(gk2), C.ptr(gx1), C.ptr(gv2), x1.numel() // c, c)
        return gq1, gk1, gq2, gk2, gx1, gv2


# ------------------------------------------------------------------------------------------------------------
# InstanceNorm2d / ada_in (attacker only: first order)
# ------------------------------------------------------------------------------------------------------------
class _NormFn(Function):
    """mode 0: InstanceNorm2d(affine, eps in the rsqrt, biased var) -- model_blocks.py:747-748, gim_img_models.py:126.
    mode 1: ada_in (unbiased std + eps) -- model_blocks.py:611-630.  Optional fused LeakyReLU(slope)."""

    @staticmethod
    def forward(ctx, x, p_scale, p_shift, mode, eps, slope):
        x = _c(x)
        n, h, w, c = x.shape
        hw = h * w
        p_scale = _c(p_scale.float())
        p_shift = _c(p_shift.float())
        st = _empty((4, n, c), torch.float32, x)          # mean | m2 | a | b
        C.call("gim_norm_stats", C.ptr(x), st[0].data_ptr(), st[1].data_ptr(), n, hw, c, C.dtype_code(x))
        C.call("gim_norm_coeffs", mode, st[0].data_ptr(), st[1].data_ptr(), C.ptr(p_scale), C.ptr(p_shift),
               st[2].data_ptr(), st[3].data_ptr(), n, hw, c, eps)
        y = torch.empty_like(x)
        C.call("gim_affine_act_fwd", C.ptr(x), st[0].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), C.ptr(y), n, hw, c, slope, C.dtype_code(x))
        ctx.cfg = (mode, eps, slope)
        ctx.save_for_backward(x, y, st, p_scale)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, y, st, p_scale = ctx.saved_tensors
        mode, eps, slope = ctx.cfg
        gy = _c(gy)
        n, h, w, c = x.shape
        hw = h * w
        yref = C.ptr(y) if slope != 1.0 else None
        red = _empty((5, n, c), torch.float32, x)         # s1 | s2 | A | B | C
        C.call("gim_norm_bwd_reduce", C.ptr(gy), C.ptr(x), yref, st[0].data_ptr(), None, None, red[0].data_ptr(), red[1].data_ptr(),
               n, hw, c, slope, C.dtype_code(x))
        if mode == 0:
            g_scale = _empty((c,), torch.float32, x)
            g_shift = _empty((c,), torch.float32, x)
        else:
            g_scale = _empty((n, c), torch.float32, x)
            g_shift = _empty((n, c), torch.float32, x)
        C.call("gim_norm_bwd_coeffs", mode, st[1].data_ptr(), red[0].data_ptr(), red[1].data_ptr(), C.ptr(p_scale),
               red[2].data_ptr(), red[3].data_ptr(), red[4].data_ptr(), C.ptr(g_scale), C.ptr(g_shift), n, hw, c, eps)
        gx = torch.empty_like(x)
        C.call("gim_norm_bwd_apply", C.ptr(gy), C.ptr(x), yref, st[0].data_ptr(), None, None, red[2].data_ptr(), red[3].data_ptr(),
               red[4].data_ptr(), C.ptr(gx), n, hw, c, slope, C.dtype_code(x))
        return gx, g_scale, g_shift, None, None, None


class NormConvFn(Function):
    """conv_k(prologue(norm(x))) as ONE autograd node on the bf16 tensor-core path: InstanceNorm2d (mode 0) or ada_in (mode 1), LeakyReLU,
    optional nearest x2 upsample and the convolution that consumes them (reference model_blocks.py:760-768 ResBlockUp, :805-811
    AdaResBlock2, :851-861 AdaResBlockUp2).  Forward: statistics -> coefficients -> ONE pass that writes the (upsampled) bf16 operand --
    the fp32 normalised activation is never materialised -- -> tcgen05 convolution, optionally with the block's residual branch added
    in its epilogue (`addend`: the half-resolution 1x1 branch, nearest-upsampled on the fly).  Backward: the upsample's sum-pooling is
    folded into the input-gradient convolution's epilogue, the activation mask is recomputed from the coefficients.  First order only."""

    @staticmethod
    def forward(ctx, x, p_scale, p_shift, w32, bias, addend, mode, eps, slope, upsample, ks):
        x = _c(x)
        w32 = _c(w32)
        n, h, w, c = x.shape
        hw = h * w
        p_scale, p_shift = _c(p_scale.float()), _c(p_shift.float())
        st = _empty((4, n, c), torch.float32, x)          # mean | m2 | a | b
        C.call("gim_norm_stats", C.ptr(x), st[0].data_ptr(), st[1].data_ptr(), n, hw, c, C.F32)
        C.call("gim_norm_coeffs", mode, st[0].data_ptr(), st[1].data_ptr(), C.ptr(p_scale), C.ptr(p_shift), st[2].data_ptr(), st[3].data_ptr(), n, hw, c, eps)
        oh, ow = (2 * h, 2 * w) if upsample else (h, w)
        xop = _empty((n, oh, ow, c), torch.bfloat16, x)
        C.call("gim_norm_act_operand", C.ptr(x), st[0].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), C.ptr(xop), n, h, w, c, slope, 1 if upsample else 0)
        w_op = _weight_as(w32, torch.bfloat16, False)
        co = w32.shape[1]
        if addend is not None:                            # + nearest-upsampled half-resolution residual branch, inside the epilogue
            y = _conv_tc_fused(xop, w_op, bias, ks, torch.float32, EPI_ADDUP | EPI_UNIT, addend=_c(addend))
        elif co % 32 == 0:
            y = _conv_tc_fused(xop, w_op, bias, ks, torch.float32)
        else:
            y = _conv_raw(xop, w_op, bias, ks)
        ctx.cfg = (mode, eps, slope, upsample, ks, bias is not None, addend is not None)
        ctx.bias_param = bias
        ctx.wg_async = _wg_async_ok(w32)
        ctx.save_for_backward(x, st, p_scale, xop, w32)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, st, p_scale, xop, w32 = ctx.saved_tensors
        mode, eps, slope, upsample, ks, has_bias, has_add = ctx.cfg
        n, h, w, c = x.shape
        hw = h * w
        gy = _c(gy)
        gb = None
        if has_bias and ctx.needs_input_grad[4]:
            gb = _bias_grad_with_operand(gy, ctx.bias_param)     # bf16 operand copy of gy + bias gradient in one pass
        gop = _operand(gy)
        gw = _wgrad_raw(xop, gop, ks, async_ok=ctx.wg_async) if ctx.needs_input_grad[3] else None
        g_add = None
        if has_add and ctx.needs_input_grad[5]:                  # d/d(half-resolution residual) = 2x2 sums of gy
            g_add = _empty((n, gy.shape[1] // 2, gy.shape[2] // 2, gy.shape[3]), torch.float32, gy)
            C.call("gim_pool2_sum", C.ptr(gy), None, C.ptr(g_add), n, gy.shape[1], gy.shape[2], gy.shape[3], 1.0, C.F32)
        gx = g_scale = g_shift = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            w_fl = _weight_as(w32, torch.bfloat16, True)
            if upsample and c % 32 == 0:
                g = _conv_tc_fused(gop, w_fl, None, ks, torch.float32, EPI_POOL | EPI_UNIT, out=_empty((n, h, w, c), torch.float32, x))      # input gradient, 2x2-summed in the epilogue
            else:
                g = _conv_tc_fused(gop, w_fl, None, ks, torch.float32) if c % 32 == 0 else _conv_raw(gop, w_fl, None, ks)
                if upsample:
                    g2 = _empty((n, h, w, c), torch.float32, g)
                    C.call("gim_pool2_sum", C.ptr(g), None, C.ptr(g2), n, 2 * h, 2 * w, c, 1.0, C.F32)
                    g = g2
            red = _empty((5, n, c), torch.float32, x)         # s1 | s2 | A | B | C
            act_a, act_b = (st[2].data_ptr(), st[3].data_ptr()) if slope != 1.0 else (None, None)
            C.call("gim_norm_bwd_reduce", C.ptr(g), C.ptr(x), None, st[0].data_ptr(), act_a, act_b, red[0].data_ptr(), red[1].data_ptr(), n, hw, c, slope, C.F32)
            if mode == 0:
                g_scale, g_shift = _empty((c,), torch.float32, x), _empty((c,), torch.float32, x)
            else:
                g_scale, g_shift = _empty((n, c), torch.float32, x), _empty((n, c), torch.float32, x)
            C.call("gim_norm_bwd_coeffs", mode, st[1].data_ptr(), red[0].data_ptr(), red[1].data_ptr(), C.ptr(p_scale),
                   red[2].data_ptr(), red[3].data_ptr(), red[4].data_ptr(), C.ptr(g_scale), C.ptr(g_shift), n, hw, c, eps)
            gx = torch.empty_like(x)
            C.call("gim_norm_bwd_apply", C.ptr(g), C.ptr(x), None, st[0].data_ptr(), act_a, act_b, red[2].data_ptr(), red[3].data_ptr(),
                   red[4].data_ptr(), C.ptr(gx), n, hw, c, slope, C.F32)
        return gx, g_scale, g_shift, gw, gb, g_add, None, None, None, None, None


def norm_conv_ok(x, w32):
    """The fused node needs the tensor-core path, fp32 NHWC input with c % 8 == 0 and a tensor-core friendly output width."""
    return fused_blocks_enabled() and x.dtype == torch.float32 and x.shape[-1] % 8 == 0 and w32.shape[1] % 16 == 0 and x.shape[0] <= 65535


def norm_conv(x, p_scale, p_shift, w32, bias, ks, mode, eps=1e-5, slope=LRELU_SLOPE, upsample=False, addend=None):
    """mode 0: InstanceNorm2d(affine) with (weight, bias) = (p_scale, p_shift); mode 1: ada_in with (std_style, mean_style)."""
    return NormConvFn.apply(x, p_scale, p_shift, w32, bias, addend, mode, eps, slope, upsample, ks)


def instance_norm(x, weight, bias, eps=1e-5, slope=1.0):
    return _NormFn.apply(x, weight, bias, 0, eps, slope)


def ada_in(x, mean_style, std_style, eps=1e-5, slope=1.0):
    return _NormFn.apply(x, std_style, mean_style, 1, eps, slope)


# ------------------------------------------------------------------------------------------------------------
# small dense algebra
# ------------------------------------------------------------------------------------------------------------
def _mm_raw(a, b, ta, tb, out_dtype):
    a = _c(a)
    b = _c(b)
    batched = a.dim() == 3
    if not batched:
        a3, b3 = a.unsqueeze(0), b.unsqueeze(0)
    else:
        a3, b3 = a, b
    bt = a3.shape[0]
    m, k = (a3.shape[2], a3.shape[1]) if ta else (a3.shape[1], a3.shape[2])
    k2, n = (b3.shape[2], b3.shape[1]) if tb else (b3.shape[1], b3.shape[2])
    if k != k2 or b3.shape[0] != bt:
        raise RuntimeError("matmul shape mismatch %s %s" % (tuple(a.shape), tuple(b.shape)))
    out = _empty((bt, m, n), out_dtype, a)
    lda, ldb = a3.shape[2], b3.shape[2]
    sam, sak = (1, lda) if ta else (lda, 1)
    sbk, sbn = (1, ldb) if tb else (ldb, 1)
    tc = _state["operand_dtype"] == torch.bfloat16 and _state["conv_algo"] != C.ALGO_SIMT
    C.call("gim_gemm_strided_bf16" if tc else "gim_gemm_strided", C.ptr(a3), C.dtype_code(a3), a3.shape[1] * a3.shape[2], sam, sak,
           C.ptr(b3), C.dtype_code(b3), b3.shape[1] * b3.shape[2], sbk, sbn,
           C.ptr(out), C.dtype_code(out), m * n, n, m, n, k, bt, 1.0, 0.0)
    return out if batched else out[0]


class MatMulFn(Function):
    """C = op(A) @ op(B), 2-D or batched 3-D, mixed fp32/bf16 operands, fp32 accumulation."""

    @staticmethod
    def forward(ctx, a, b, ta, tb, out_dtype):
        ctx.cfg = (ta, tb)
        ctx.save_for_backward(a, b)
        return _mm_raw(a, b, ta, tb, out_dtype)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        ta, tb = ctx.cfg
        ga = gb = None
        if ctx.needs_input_grad[0]:
            if not ta:
                ga = MatMulFn.apply(g, b, False, not tb, a.dtype)
            else:
                ga = MatMulFn.apply(b, g, tb, True, a.dtype)
        if ctx.needs_input_grad[1]:
            if not tb:
                gb = MatMulFn.apply(a, g, not ta, False, b.dtype)
            else:
                gb = MatMulFn.apply(g, a, True, ta, b.dtype)
        return ga, gb, None, None, None


def matmul(a, b, ta=False, tb=False, out_dtype=torch.float32):
    return MatMulFn.apply(a, b, ta, tb, out_dtype)


class BiasActFn(Function):
    """y = LeakyReLU_slope(x + bias) on fp32 [rows, c] (slope 1 = no activation)."""

    @staticmethod
    def forward(ctx, x, bias, slope):
        x = _c(x)
        y = torch.empty_like(x)
        c = x.shape[-1]
        C.call("gim_bias_act_fwd", C.ptr(x), C.ptr(bias), C.ptr(y), x.numel() // c, c, slope)
        ctx.slope = slope
        ctx.has_bias = bias is not None
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        g = gy if ctx.slope == 1.0 else LReluBwdFn.apply(gy, y.detach(), ctx.slope)
        gb = ColSumFn.apply(g) if (ctx.has_bias and ctx.needs_input_grad[1]) else None
        return g, gb, None


def linear(x, weight, bias, slope=1.0):
    """nn.Linear (+ optional LeakyReLU) on the last dim; x fp32 [..., in].  On the bf16 path a Linear is a 1x1 convolution over
    `rows` one-pixel images, so it runs on the same tcgen05 implicit-GEMM kernels (forward, input- and weight-gradient)."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    rows, k = x2.shape
    n = weight.shape[0]
    if _state["operand_dtype"] == torch.bfloat16 and k % 8 == 0 and n % 8 == 0 and rows >= 16:
        # (the same route -- hence the same single bf16 rounding of x, W and the incoming gradient -- whichever conv algorithm is
        # selected: with set_conv_algo("simt") the CUDA-core kernels consume the identical operands, which is what lets
        # tests/test_parity_gpu.py compare the tcgen05 path element by element)
        y = conv2d(x2.reshape(rows, 1, 1, k), weight.reshape(1, n, k), bias, 1).reshape(rows, n)
        if slope != 1.0:
            y = LReluFn.apply(y, slope)
    else:
        y = matmul(x2, weight, False, True)
        y = BiasActFn.apply(y, bias, slope)
    return y.reshape(shp[:-1] + (n,))


class SoftmaxRowsFn(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        cols = x.shape[-1]
        C.call("gim_softmax_rows_fwd", C.ptr(x), C.ptr(y), x.numel() // cols, cols)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        return SoftmaxBwdFn.apply(gy, y)


class SoftmaxBwdFn(Function):
    """gx = y * (g - sum(g*y)); symmetric in g, explicit kernel for d/dy."""

    @staticmethod
    def forward(ctx, g, y):
        g = _c(g)
        gx = torch.empty_like(g)
        cols = g.shape[-1]
        C.call("gim_softmax_rows_bwd", C.ptr(g), C.ptr(y), C.ptr(gx), g.numel() // cols, cols)
        ctx.save_for_backward(g, y)
        return gx

    @staticmethod
    def backward(ctx, gg):
        g, y = ctx.saved_tensors
        gg = _c(gg)
        d_g = SoftmaxBwdFn.apply(gg, y) if ctx.needs_input_grad[0] else None
        d_y = None
        if ctx.needs_input_grad[1]:
            d_y = _SoftmaxBwdBwdY.apply(gg, g, y)
        return d_g, d_y


class _SoftmaxBwdBwdY(Function):
    @staticmethod
    def forward(ctx, gg, g, y):
        out = torch.empty_like(y)
        cols = y.shape[-1]
        C.call("gim_softmax_rows_bwd_bwd", C.ptr(_c(gg)), C.ptr(_c(g)), C.ptr(y), C.ptr(out), y.numel() // cols, cols)
        return out

    @staticmethod
    def backward(ctx, *a):
        raise RuntimeError("third-order derivative through softmax is not implemented (not needed by the GIM path)")


# ------------------------------------------------------------------------------------------------------------
# fused SelfAttention core (one CTA per image; first order only)
# ------------------------------------------------------------------------------------------------------------
ATTN_POSITIONS, ATTN_CHANNELS = 64, (128, 256)


def attention_fused_ok(positions, channels, *tensors):
    """The fused kernels cover the 8x8 maps of the 32x32 pyramids (64 positions, 128 / 256 channels), fp32, first-order graphs."""
    return (not _state["composite"] and positions == ATTN_POSITIONS and channels in ATTN_CHANNELS
            and all(t.dtype == torch.float32 for t in tensors))


class AttentionCoreFn(Function):
    """y = gamma * softmax_i(<q_j, k_i>) v + x for [n, 64, c/8] queries / keys and [n, 64, c] values / input
    (reference model_blocks.py:536-548); one kernel forward, one backward."""

    @staticmethod
    def forward(ctx, q, k, v, x, gamma):
        q, k, v, x = _c(q), _c(k), _c(v), _c(x)
        n, p, c = v.shape
        attn = _empty((n, p, p), torch.float32, v)
        y = torch.empty_like(x)
        C.call("gim_attention_fwd", C.ptr(q), C.ptr(k), C.ptr(v), q.shape[2], c, C.ptr(x), C.ptr(gamma), C.ptr(attn), C.ptr(y), n, p, c)
        ctx.save_for_backward(q, k, v, attn, gamma)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        q, k, v, attn, gamma = ctx.saved_tensors
        gy = _c(gy)
        n, p, c = v.shape
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        part = _empty((n,), torch.float32, v)
        C.call("gim_attention_bwd", C.ptr(gy), C.ptr(q), C.ptr(k), C.ptr(v), q.shape[2], c, C.ptr(attn), C.ptr(gamma), C.ptr(dq), C.ptr(dk), C.ptr(dv),
               C.ptr(part), n, p, c)
        dgamma = part.sum().reshape(gamma.shape) if ctx.needs_input_grad[4] else None
        return dq, dk, dv, gy.view_as(v), dgamma


# ------------------------------------------------------------------------------------------------------------
# set statistics over the sample axis
# ------------------------------------------------------------------------------------------------------------
class SetSumFn(Function):
    """[b, s, d] -> scale * sum_s  (mean when scale = 1/s): GIMMeanStat gim_basic_models.py:28-34."""

    @staticmethod
    def forward(ctx, x, scale):
        x = _c(x)
        b, s, d = x.shape
        out = _empty((b, d), torch.float32, x)
        C.call("gim_set_stats_fwd", C.ptr(x), C.ptr(out), None, d, b, s, d, scale, 0.0)
        ctx.s = s
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, g):
        return SetBroadcastFn.apply(g, ctx.s, ctx.scale), None


class SetBroadcastFn(Function):
    @staticmethod
    def forward(ctx, g, s, scale):
        g = _c(g)
        b, d = g.shape
        dummy = g                                            # x is only read when g_std is given
        out = _empty((b, s, d), torch.float32, g)
        C.call("gim_set_stats_bwd", C.ptr(g), None, d, C.ptr(dummy), C.ptr(out), b, s, d, scale, 0.0)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, gg):
        return SetSumFn.apply(gg, ctx.scale), None, None


class SetStdFn(Function):
    """custom_std model_blocks.py:41-48: sqrt(var_unbiased + 1e-8), zeros if the set has one element."""

    @staticmethod
    def forward(ctx, x, eps):
        x = _c(x)
        b, s, d = x.shape
        out = _empty((b, d), torch.float32, x)
        C.call("gim_set_stats_fwd", C.ptr(x), None, C.ptr(out), d, b, s, d, 1.0, eps)
        ctx.eps = eps
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return SetStdBwdFn.apply(g, x, ctx.eps), None


class SetStdBwdFn(Function):
    @staticmethod
    def forward(ctx, g, x, eps):
        g = _c(g)
        b, s, d = x.shape
        gx = torch.empty_like(x)
        C.call("gim_set_stats_bwd", None, C.ptr(g), d, C.ptr(x), C.ptr(gx), b, s, d, 1.0, eps)
        ctx.eps = eps
        ctx.save_for_backward(g, x)
        return gx

    @staticmethod
    @once_differentiable
    def backward(ctx, ggx):
        g, x = ctx.saved_tensors
        b, s, d = x.shape
        ggx = _c(ggx)
        gg_std = torch.empty_like(g)
        g_x = torch.empty_like(x)
        C.call("gim_set_std_bwd_bwd", C.ptr(ggx), C.ptr(g), d, C.ptr(x), C.ptr(gg_std), C.ptr(g_x), b, s, d, ctx.eps)
        return gg_std, g_x, None


class SetMeanStdFn(Function):
    """cat(mean_s x, custom_std_s x) -> [b, 2d] in ONE pass over x (GIMMeanStdStat gim_basic_models.py:71-89 and the first two thirds of
    GIMMeanStdFcStat :152-172): the statistics kernel emits both and writes them side by side, so x is read once and there is no concat.
    Twice differentiable (R1): the backward is itself an operator with an explicit second-order kernel."""

    @staticmethod
    def forward(ctx, x, eps):
        x = _c(x)
        b, s, d = x.shape
        out = _empty((b, 2 * d), torch.float32, x)
        C.call("gim_set_stats_fwd", C.ptr(x), C.ptr(out), C.ptr(out) + 4 * d, 2 * d, b, s, d, 1.0 / s, eps)
        ctx.eps = eps
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return SetMeanStdBwdFn.apply(g, x, ctx.eps), None


class SetMeanStdBwdFn(Function):
    @staticmethod
    def forward(ctx, g, x, eps):
        g = _c(g)
        b, s, d = x.shape
        gx = torch.empty_like(x)
        C.call("gim_set_stats_bwd", C.ptr(g), C.ptr(g) + 4 * d, 2 * d, C.ptr(x), C.ptr(gx), b, s, d, 1.0 / s, eps)
        ctx.eps = eps
        ctx.save_for_backward(g, x)
        return gx

    @staticmethod
    @once_differentiable
    def backward(ctx, ggx):
        g, x = ctx.saved_tensors
        b, s, d = x.shape
        ggx = _c(ggx)
        gg = _empty((b, 2 * d), torch.float32, x)
        gg_std = _empty((b, d), torch.float32, x)
        g_x = torch.empty_like(x)
        C.call("gim_set_stats_fwd", C.ptr(ggx), C.ptr(gg), None, 2 * d, b, s, d, 1.0 / s, 0.0)             # d/d g_mean: the mean of ggx
        C.call("gim_set_std_bwd_bwd", C.ptr(ggx), C.ptr(g) + 4 * d, 2 * d, C.ptr(x), C.ptr(gg_std), C.ptr(g_x), b, s, d, ctx.eps)
        gg[:, d:] = gg_std
        return gg, g_x, None


def set_mean_std(x, eps=1e-8):
    return SetMeanStdFn.apply(x, eps)


def gaussian_episodes(batch, sizes, d, prior_sigma, src_sigma, device):
    """Device-side synthesis of Gaussian GIM episodes with the reference's distributions (training/gim_gaussian_training.py:71-86):
    mu ~ N(0, prior_sigma^2 I) per episode, every sample of the episode ~ N(mu, src_sigma^2 I).  -> (mu [b, d], [x_i [b, s_i, d] for s_i
    in sizes]).  Standard normals come from the CUDA generator (seeded by torch.manual_seed; CUDA-graph safe), one kernel per sample set
    shifts and scales them.  The reference draws on the host and copies 65 M floats per iteration at d = 1000."""
    mu_raw = torch.randn((batch, d), device=device)
    out = []
    for s in sizes:
        noise = torch.randn((batch, s, d), device=device)
        C.call("gim_affine_rows", C.ptr(noise), C.ptr(mu_raw), C.ptr(noise), batch, s, d, float(src_sigma), float(prior_sigma))
        out.append(noise)
    return mu_raw * prior_sigma, out


def set_mean(x):
    return SetSumFn.apply(x, 1.0 / x.shape[1])


def set_std(x, eps=1e-8):
    return SetStdFn.apply(x, eps)


class SetCenterAddFn(Function):
    """y = x - mean_s(x) (if center) + add[:, None]  (gim_img_models.py:378-380; gim_gaussian_models.py:84-88)."""

    @staticmethod
    def forward(ctx, x, add, center):
        x = _c(x)
        b, s, d = x.shape
        y = torch.empty_like(x)
        C.call("gim_set_center_add", C.ptr(x), C.ptr(_c(add)) if add is not None else None, C.ptr(y), b, s, d, 1 if center else 0)
        ctx.center = center
        ctx.has_add = add is not None
        return y

    @staticmethod
    def backward(ctx, g):
        gx = None
        if ctx.needs_input_grad[0]:
            gx = SetCenterAddFn.apply(g, None, True) if ctx.center else g
        ga = SetSumFn.apply(g, 1.0) if (ctx.has_add and ctx.needs_input_grad[1]) else None
        return gx, ga, None


# ------------------------------------------------------------------------------------------------------------
# encoder tail, losses
# ------------------------------------------------------------------------------------------------------------
class GlobalMaxFn(Function):
    """AdaptiveMaxPool2d((1,1)) + flatten (+ the encoder's output LeakyReLU when slope != 1) -- gim_img_models.py:53-56 -- in one pass:
    NHWC activation -> fp32 [n, c].  The backward is built from differentiable operators (mask, scatter), so it stays twice differentiable."""

    @staticmethod
    def forward(ctx, x, slope=1.0):
        x = _c(x)
        n, h, w, c = x.shape
        y = _empty((n, c), torch.float32, x)
        idx = _empty((n, c), torch.int32, x)
        C.call("gim_gmax_fwd", C.ptr(x), C.ptr(y), C.ptr(idx), n, h * w, c, float(slope), C.dtype_code(x))
        ctx.shape = x.shape
        ctx.dtype = x.dtype
        ctx.idx = idx
        ctx.slope = float(slope)
        if slope != 1.0:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        if ctx.slope != 1.0:
            (y,) = ctx.saved_tensors
            g = LReluBwdFn.apply(g, y.detach(), ctx.slope)
        return ScatterIdxFn.apply(g, ctx.idx, ctx.shape, ctx.dtype), None


class ScatterIdxFn(Function):
    @staticmethod
    def forward(ctx, g, idx, shape, dtype):
        g = _c(g)
        n, h, w, c = shape
        gx = _empty(shape, dtype, g)
        C.call("gim_scatter_idx", C.ptr(g), C.ptr(idx), C.ptr(gx), n, h * w, c, C.dtype_code(gx))
        ctx.idx = idx
        return gx

    @staticmethod
    def backward(ctx, gg):
        return GatherIdxFn.apply(gg, ctx.idx), None, None, None


class GatherIdxFn(Function):
    @staticmethod
    def forward(ctx, x, idx):
        x = _c(x)
        n, h, w, c = x.shape
        y = _empty((n, c), torch.float32, x)
        C.call("gim_gather_idx", C.ptr(x), C.ptr(idx), C.ptr(y), n, h * w, c, C.dtype_code(x))
        ctx.idx = idx
        ctx.shape = x.shape
        ctx.dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, g):
        return ScatterIdxFn.apply(g, ctx.idx, ctx.shape, ctx.dtype), None


class BCEWithLogitsFn(Function):
    """F.binary_cross_entropy_with_logits(x, full(target), reduction='none') -- gan_loss gim_img_trainer.py:90-94."""

    @staticmethod
    def forward(ctx, x, target):
        x = _c(x)
        loss = torch.empty_like(x)
        C.call("gim_bce_logits_fwd", C.ptr(x), float(target), C.ptr(loss), x.numel())
        ctx.target = float(target)
        ctx.save_for_backward(x)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _c(g)
        gx = torch.empty_like(x)
        C.call("gim_bce_logits_bwd", C.ptr(g), C.ptr(x), ctx.target, C.ptr(gx), x.numel())
        return gx, None


class RowsSqSumFn(Function):
    """[b, ...] -> fp32 [b] sum of squares per episode (compute_grad2 training/utils.py:122)."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        b = x.shape[0]
        out = _empty((b,), torch.float32, x)
        C.call("gim_rows_sqsum", C.ptr(x), C.ptr(out), b, x.numel() // b, C.dtype_code(x))
        ctx.save_for_backward(x)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = _c(g.float())
        b = x.shape[0]
        gx = torch.empty_like(x)
        C.call("gim_rows_scale", C.ptr(x), C.ptr(g), C.ptr(gx), b, x.numel() // b, 2.0, C.dtype_code(x))
        return gx
