"""Times the SelfAttention core (8x8 map) on the fused kernels against the composed gemm + softmax route:
   python tools/attention_bench.py [--images 640] [--channels 256]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimalstrategiesagainstgenerativeattacks_b200 import ops  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=640)
    ap.add_argument("--channels", type=int, default=256)
    args = ap.parse_args()
    ops.set_precision("bf16")
    n, p, c = args.images, 64, args.channels
    q, k = (torch.randn(n, p, c // 8, device="cuda").requires_grad_() for _ in range(2))
    v, x = (torch.randn(n, p, c, device="cuda").requires_grad_() for _ in range(2))
    gm = torch.full((1,), 0.3, device="cuda").requires_grad_()
    gy = torch.randn(n, p, c, device="cuda")

    def fused():
        return ops.AttentionCoreFn.apply(q, k, v, x, gm)

    def composed():
        a = ops.SoftmaxRowsFn.apply(ops.matmul(q, k, False, True, torch.float32))
        return ops.AddFn.apply(ops.ScaleDevFn.apply(ops.matmul(a, v, False, False, torch.float32), gm), x)

    for name, fn in (("fused", fused), ("composed", composed)):
        with torch.no_grad():
            t_f = timed(fn)
        t_fb = timed(lambda: torch.autograd.grad(fn(), (q, k, v, x, gm), gy))
        hbm_f = n * p * (2 * c // 8 + 3 * c + p) * 4 / 1e6
        print("%-9s images %d channels %d: forward %.1f us (algorithmic %.0f MB -> %.2f TB/s), forward+backward %.1f us" % (
            name, n, c, t_f, hbm_f, hbm_f / t_f, t_fb), flush=True)


if __name__ == "__main__":
    main()
