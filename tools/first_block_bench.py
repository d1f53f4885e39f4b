"""Times the fused image-side kernels of the first ResBlockDown alone: python tools/first_block_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from optimalstrategiesagainstgenerativeattacks_b200 import _cabi as C

dev = torch.device("cuda", 0)
n, h, w, c, co, k = 640, 32, 32, 1, 128, 3
x = torch.randn((n, h, w, c), device=dev)
w1 = torch.randn((k * k, co, c), device=dev)
wl = torch.randn((1, co, c), device=dev)
b1, bl = torch.randn(co, device=dev), torch.randn(co, device=dev)
t = torch.empty((n, h, w, co), device=dev, dtype=torch.bfloat16)
r = torch.empty((n, h // 2, w // 2, co), device=dev)
gt = torch.randn((n, h, w, co), device=dev).bfloat16()
gy = torch.randn((n, h // 2, w // 2, co), device=dev)
scr = torch.empty(32 * 10 * co * c, device=dev)
gw1, gwl = torch.empty((k * k, co, c), device=dev), torch.empty((1, co, c), device=dev)
fwd = lambda: C.call("gim_first_block_fwd", C.ptr(x), C.ptr(w1), C.ptr(b1), C.ptr(wl), C.ptr(bl), C.ptr(t), C.ptr(r), n, h, w, c, co, k, 0.2)
wg = lambda: C.call("gim_first_block_wgrad", C.ptr(x), C.ptr(gt), C.ptr(gy), C.ptr(gw1), C.ptr(gwl), C.ptr(scr), scr.numel(), n, h, w, c, co, k, 0.2)
for name, f, nbytes in (("fwd", fwd, t.numel() * 2 + r.numel() * 4), ("wgrad", wg, gt.numel() * 2 + gy.numel() * 4)):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print("first_block_%s n=%d %dx%d c=%d co=%d k=%d: %.1f us, %.2f TB/s" % (name, n, h, w, c, co, k, us, nbytes / us / 1e6))
