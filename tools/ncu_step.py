"""One training iteration inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`).
Usage: python tools/ncu_step.py [--batch 32] [--workload O]"""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import optimalstrategiesagainstgenerativeattacks_b200 as gim
from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
from optimalstrategiesagainstgenerativeattacks_b200.gim_img_trainer import GIMImgTrainer
from optimalstrategiesagainstgenerativeattacks_b200.training_steps import au_train_step, im_train_step
from optimalstrategiesagainstgenerativeattacks_b200.utils import DataParallelMock

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--workload", default="O")
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
size, ch, reg, au_lr, im_lr, map_lr, gflop, desc = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
gim.set_precision("bf16")
torch.manual_seed(1)
au, im = M.get_au(size, ch, 512).to(dev), M.get_im(size, ch, 512).to(dev)
tr = DataParallelMock(GIMImgTrainer(tempfile.mkdtemp(), 5, 5, 5, au, im, au_lr, im_lr, map_lr, reg_param=reg))
leaked, real, si = bench.synth_batch(a.batch, ch, size, 1234, dev)


def it():
    tr.module.do_global_step()
    tr.module.update_learning_rate()
    _, fake, _ = im_train_step(tr, leaked, si)
    au_train_step(tr, real, fake, si)


for _ in range(a.warmup):
    it()
torch.cuda.synchronize()
torch.cuda.profiler.start()
it()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
