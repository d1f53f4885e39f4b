"""Secondary throughput numbers of BASELINE.json's configs that are not the headline metric (one JSON line each):

  gauss10 / gauss1000 : Gaussian GIM training, d = 10 / 1000, n=5 m=1 k=10, B episodes per iteration sampled ON THE DEVICE
                        (reference training/gim_gaussian_training.py:71-86 samples on the CPU) -- configs[0] and configs[3]
  eval                : authentication eval (GIM / replay / random-source attackers), inference only -- configs[4]

  python tools/bench_extra.py [--what gauss10,gauss1000,eval] [--batch 4096] [--steps 20]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import optimalstrategiesagainstgenerativeattacks_b200 as gim
from optimalstrategiesagainstgenerativeattacks_b200 import authentication_eval as AE
from optimalstrategiesagainstgenerativeattacks_b200 import gim_gaussian_models as GM
from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
from optimalstrategiesagainstgenerativeattacks_b200 import img_datasets as D
from optimalstrategiesagainstgenerativeattacks_b200.cuda_graph import GraphedIteration
from optimalstrategiesagainstgenerativeattacks_b200.gim_gaussian_trainer import GIMGaussianTrainer
from optimalstrategiesagainstgenerativeattacks_b200.utils import DataParallelMock


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def gauss(d, batch, steps, precision):
    dev = torch.device("cuda", 0)
    gim.set_precision(precision)
    torch.manual_seed(1)
    m, n, k, prior_sigma, src_sigma = 1, 5, 10, 10.0, 1.0
    au, im = GM.get_au(d).to(dev), GM.get_im(d).to(dev)
    tr = DataParallelMock(GIMGaussianTrainer(tempfile.mkdtemp(), m, n, k, au, im, 1e-4, 1e-4, reg_param=0.0))

    def sample():
        mu = torch.randn((batch, 1, d), device=dev) * prior_sigma
        return tuple(mu + torch.randn((batch, s, d), device=dev) * src_sigma for s in (m, n, k))     # leaked, real, si

    g = GraphedIteration(tr, *sample(), warmup=3)
    ms = timed(lambda: g(*sample()), steps)
    return {"metric": "Gaussian GIM train episodes/sec", "value": batch / (ms * 1e-3), "unit": "episodes/s", "ms_per_step": ms,
            "iterations_per_s": 1e3 / ms, "dtype": precision,
            "config": {"workload": "Gaussian GIM d=%d n=5 m=1 k=10, B=%d, device-side sampling, whole iteration in one CUDA graph" % (d, batch)}}


def evaluation(batch, steps, precision):
    dev = torch.device("cuda", 0)
    gim.set_precision(precision)
    torch.manual_seed(1)
    au, im = M.get_au(64, 3, 512).to(dev), M.get_im(64, 3, 512).to(dev)
    ds = D.ResidentGIMDataSet(D.synthetic_classes(64, 20, 3, 64, device=dev), m=5, n=5, k=5, example_cnt_per_class=64, device=dev, seed=1)
    authenticator = AE.Authenticator(AE.get_au_function(au))
    agents = {"gim": AE.Impersonator(AE.get_im_function(im, {"remove_noise_mean": True})), "replay": AE.Impersonator(AE.replay_impersonator),
              "rnd_src": AE.Impersonator(lambda leaked_sample, n: AE.rand_source_impersonator(leaked_sample, n, ds))}
    out = {}
    for name, agent in agents.items():
        b = ds.batch(range(batch))

        def step():
            authenticator.act(test_sample=b["real_sample"], si_sample=b["si_sample"])
            fake = agent.act(leaked_sample=b["leaked_sample"], n=5)
            authenticator.act(test_sample=fake, si_sample=b["si_sample"])
        ms = timed(step, steps)
        out[name] = batch / (ms * 1e-3)
    t0 = time.perf_counter()
    acc, acc_fake, acc_real, auc = AE.eval_authenticator_and_impersonator(dev, ds, batch, 0, authenticator, agents["gim"])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    return {"metric": "authentication eval episodes/sec (inference only)", "value": out["gim"], "unit": "episodes/s", "dtype": precision,
            "per_attacker_episodes_per_s": out, "full_loop_episodes_per_s": len(ds) / wall, "auc_random_init": auc,
            "config": {"workload": "VoxCeleb2-shaped 3x64x64 m=n=k=5, %d episodes, batch %d, device-resident episode source" % (len(ds), batch)}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="gauss10,gauss1000,eval")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--eval-batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    for w in a.what.split(","):
        if w == "gauss10":
            print(json.dumps(gauss(10, a.batch, a.steps, a.precision)), flush=True)
        elif w == "gauss1000":
            print(json.dumps(gauss(1000, a.batch, a.steps, a.precision)), flush=True)
        elif w == "eval":
            print(json.dumps(evaluation(a.eval_batch, max(3, a.steps // 4), a.precision)), flush=True)
