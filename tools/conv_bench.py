"""Micro-benchmark of the tcgen05 convolution kernels on the hot layer shapes of the O / V workloads (through the C ABI).

  python tools/conv_bench.py [--images 640] [--reps 20] [--only fwd|wgrad] [--shapes 0,3]

Prints TFLOP/s per shape (CUDA events around `reps` back-to-back launches, inputs rotated through a pool larger than L2) and
checks each result once against a bf16-operand torch reference.  Also the command profiled by ncu (profiles/).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from optimalstrategiesagainstgenerativeattacks_b200 import _cabi as C

# (h, w, cin, cout, k) -- images = --images (B*5 per encoder pass at B=128)
SHAPES = [
    (32, 32, 128, 128, 3),
    (16, 16, 256, 256, 3),
    (8, 8, 512, 512, 3),
    (32, 32, 128, 128, 9),
    (16, 16, 128, 256, 3),
    (8, 8, 256, 512, 3),
    (4, 4, 512, 512, 3),
    (64, 64, 64, 64, 3),
    (16, 16, 128, 256, 1),
    (1, 1, 512, 512, 1),
    (32, 32, 8, 128, 1),
    (8, 8, 256, 256, 1),
    (1, 1, 1536, 1024, 1),
    (1, 1, 1024, 1024, 1),
    (8, 8, 128, 256, 1),
    (4, 4, 256, 512, 1),
    (8, 8, 512, 256, 3),
    (16, 16, 256, 128, 3),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=640)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--shapes", default="")
    ap.add_argument("--out-bf16", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    sel = [int(s) for s in a.shapes.split(",")] if a.shapes else range(len(SHAPES))
    for si in sel:
        h, w, ci, co, k = SHAPES[si]
        n = a.images if h * w < 4096 else max(8, a.images // 4)
        taps = k * k
        flops = 2.0 * n * h * w * ci * co * taps
        act_bytes = n * h * w * (ci * 2 + co * 4)
        pool = max(2, int(300e6 // act_bytes) + 1)
        xs = [torch.randn((n, h, w, ci), device=dev).to(torch.bfloat16) for _ in range(pool)]
        gs = [torch.randn((n, h, w, co), device=dev).to(torch.bfloat16) for _ in range(pool)]
        wt = (torch.randn((taps, co, ci), device=dev) / (ci * taps) ** 0.5).to(torch.bfloat16)
        bias = torch.randn((co,), device=dev)
        od = torch.bfloat16 if a.out_bf16 else torch.float32
        ys = [torch.empty((n, h, w, co), device=dev, dtype=od) for _ in range(pool)]
        gw = torch.empty((taps, co, ci), device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        line = "shape n=%d %dx%d ci=%d co=%d k=%d :" % (n, h, w, ci, co, k)
        if a.only in ("", "fwd"):
            def fwd(i):
                C.call("gim_conv2d_fwd", C.ptr(xs[i % pool]), C.ptr(wt), C.ptr(bias), C.ptr(ys[i % pool]), n, h, w, ci, co, k, C.BF16,
                       C.dtype_code(ys[0]), C.ALGO_TCGEN05)
            for i in range(3):
                fwd(i)
            torch.cuda.synchronize()
            e0.record()
            for i in range(a.reps):
                fwd(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            line += "  fwd %8.1f us %7.1f TFLOP/s" % (ms * 1e3, flops / ms / 1e9)
            if not a.no_check:
                nn_ = min(n, 8)
                ref = torch.nn.functional.conv2d(xs[0][:nn_].float().permute(0, 3, 1, 2),
                                                 wt.float().reshape(k, k, co, ci).permute(2, 3, 0, 1), bias, padding=(k - 1) // 2).permute(0, 2, 3, 1)
                fwd(0)
                err = float((ys[0][:nn_].float() - ref).norm() / ref.norm())
                line += " (rel %.1e)" % err
                assert err < (1e-2 if a.out_bf16 else 2e-3), err
        if a.only in ("", "fused") and co % 32 == 0:
            # fused epilogues: bf16(LeakyReLU(conv + b)) and conv * lrelu-mask(ref) -> bf16 / fp32
            yb = torch.empty((n, h, w, co), device=dev, dtype=torch.bfloat16)
            yf = torch.empty((n, h, w, co), device=dev, dtype=torch.float32)
            ref_t = torch.randn((n, h, w, co), device=dev).to(torch.bfloat16)

            def fused(i):
                C.call("gim_conv2d_fwd_fused", C.ptr(xs[i % pool]), C.ptr(wt), C.ptr(bias), C.ptr(yb), None, None, n, h, w, ci, co, k, C.BF16, 1, 0.2)
            for i in range(3):
                fused(i)
            torch.cuda.synchronize()
            e0.record()
            for i in range(a.reps):
                fused(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            line += "  fwd+lrelu->bf16 %8.1f us %7.1f TFLOP/s" % (ms * 1e3, flops / ms / 1e9)
            if not a.no_check:
                nn_ = min(n, 8)
                ref = torch.nn.functional.conv2d(xs[0][:nn_].float().permute(0, 3, 1, 2),
                                                 wt.float().reshape(k, k, co, ci).permute(2, 3, 0, 1), bias, padding=(k - 1) // 2).permute(0, 2, 3, 1)
                fused(0)
                want = torch.nn.functional.leaky_relu(ref, 0.2)
                err = float((yb[:nn_].float() - want).norm() / want.norm())
                C.call("gim_conv2d_fwd_fused", C.ptr(xs[0]), C.ptr(wt), None, C.ptr(yf), C.ptr(ref_t), None, n, h, w, ci, co, k, C.F32, 2, 0.2)
                want2 = (ref - bias) * torch.where(ref_t[:nn_].float() > 0, 1.0, 0.2)
                err2 = float((yf[:nn_] - want2).norm() / want2.norm())
                C.call("gim_conv2d_fwd_fused", C.ptr(xs[0]), C.ptr(wt), None, C.ptr(yb), C.ptr(ref_t), None, n, h, w, ci, co, k, C.BF16, 2, 0.2)
                err3 = float((yb[:nn_].float() - want2).norm() / want2.norm())
                line += " (rel %.1e mask %.1e %.1e)" % (err, err2, err3)
                assert err < 1e-2 and err2 < 2e-3 and err3 < 1e-2, (err, err2, err3)
            del yb, yf, ref_t
        if a.only in ("", "wgrad"):
            def wg(i):
                C.call("gim_conv2d_wgrad", C.ptr(xs[i % pool]), C.ptr(gs[i % pool]), C.ptr(gw), n, h, w, ci, co, k, C.BF16, C.ALGO_TCGEN05)
            for i in range(3):
                wg(i)
            torch.cuda.synchronize()
            e0.record()
            for i in range(a.reps):
                wg(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            line += "  wgrad %8.1f us %7.1f TFLOP/s" % (ms * 1e3, flops / ms / 1e9)
            if not a.no_check:
                nn_ = min(n, 8)
                xx = xs[0][:nn_].float().permute(0, 3, 1, 2).contiguous().requires_grad_(False)
                wref = torch.zeros((co, ci, k, k), device=dev, requires_grad=True)
                yref = torch.nn.functional.conv2d(xx, wref, None, padding=(k - 1) // 2)
                (gref,) = torch.autograd.grad(yref, wref, gs[0][:nn_].float().permute(0, 3, 1, 2))
                gw2 = torch.empty_like(gw)
                C.call("gim_conv2d_wgrad", C.ptr(xs[0][:nn_].contiguous()), C.ptr(gs[0][:nn_].contiguous()), C.ptr(gw2), nn_, h, w, ci, co, k, C.BF16,
                       C.ALGO_TCGEN05)
                ref = gref.permute(2, 3, 0, 1).reshape(taps, co, ci)
                err = float((gw2 - ref).norm() / ref.norm())
                line += " (rel %.1e)" % err
                assert err < 2e-3, err
        print(line, flush=True)
        del xs, gs, ys


if __name__ == "__main__":
    main()
