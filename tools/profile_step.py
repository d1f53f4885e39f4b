"""Kernel-level time breakdown of one training iteration (torch.profiler/CUPTI) -> gpurun_out/step_profile.txt.
Usage: python tools/profile_step.py [--batch 128] [--workload O]"""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import optimalstrategiesagainstgenerativeattacks_b200 as gim
from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
from optimalstrategiesagainstgenerativeattacks_b200.gim_img_trainer import GIMImgTrainer
from optimalstrategiesagainstgenerativeattacks_b200.training_steps import au_train_step, im_train_step
from optimalstrategiesagainstgenerativeattacks_b200.utils import DataParallelMock

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--workload", default="O")
ap.add_argument("--precision", default="bf16")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_profile.txt"))
a = ap.parse_args()
size, ch, reg, au_lr, im_lr, map_lr, gflop, desc = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
gim.set_precision(a.precision)
torch.manual_seed(1)
au, im = M.get_au(size, ch, 512).to(dev), M.get_im(size, ch, 512).to(dev)
tr = DataParallelMock(GIMImgTrainer(tempfile.mkdtemp(), 5, 5, 5, au, im, au_lr, im_lr, map_lr, reg_param=reg))
leaked, real, si = bench.synth_batch(a.batch, ch, size, 1234, dev)


def it():
    tr.module.do_global_step()
    tr.module.update_learning_rate()
    _, fake, _ = im_train_step(tr, leaked, si)
    au_train_step(tr, real, fake, si)


for _ in range(3):
    it()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
it()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    it()
    torch.cuda.synchronize()
os.makedirs(os.path.dirname(a.out), exist_ok=True)
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
kern = [e for e in ev if any(t in e.key for t in ("gim::", "at::", "Memset", "Memcpy", "nccl"))]
ksum, kcount = sum(e.device_time_total for e in kern) / 1e3, sum(e.count for e in kern)
ev.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ev)
with open(a.out, "w") as f:
    f.write("workload %s batch %d precision %s: eager wall %.1f ms/iteration; kernels only: %.1f ms over %d launches\n" % (a.workload, a.batch, a.precision, wall * 1e3, ksum, kcount))
    for e in ev[:70]:
        f.write("%8.2f ms %5.1f%% %6d  %s\n" % (e.device_time_total / 1e3, 100 * e.device_time_total / tot, e.count, e.key[:110]))
print(open(a.out).read())
