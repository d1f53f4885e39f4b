#!/usr/bin/env python
"""bench.py -- GIM train episodes/sec (fwd+bwd G+D) on N B200s of one node (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload O|V] [--batch B] [--precision bf16|fp32]
                  [--no-secondary] [--no-cpu-baseline] [--no-gpu-baseline] [--check]

A "step" is one training iteration = one attacker (G) step + one authenticator (D) step over B episodes per GPU through the
kept trainer API (GIMImgTrainer + im_train_step / au_train_step).  Default workload = BASELINE.json configs[1]: synthetic
Omniglot-shaped episodes, m=n=k=5, at the reference's own Omniglot resolution 1x32x32 (the reference Impersonator cannot run
at 105x105), reg 0, Adam(0, 0.99), lrs 1e-6/1e-5/1e-7, bf16 tensor-core path.

Rank 0 prints ONE JSON line on stdout (everything else goes to stderr):
  value            whole-job episodes/s, inputs resident in HBM
  e2e              the same through the public API with pinned-host inputs copied in and the losses read back every step
  roofline         the dominant kernel (tcgen05 implicit-GEMM conv forward/input-gradient) against the measured bf16 peak
  secondary        BASELINE configs[2] (VoxCeleb2-shaped 3x64x64, R1 reg=10): weak scaling at 128 episodes/GPU and strong scaling
                   at the reference's global batch of 128 episodes (128/N per GPU) -- SURVEY.md section 8d config 3 --, same timing rules
  other_configs    BASELINE configs[0]/[3] (Gaussian GIM d = 10 / 1000, batch sweep, device-side episode synthesis) at every N, and the
                   authenticator-only forward+backward at 1x105x105 (N = 1)
  cpu_baseline     the reference's own trainer (oracle/_ref, kind "reference"; else the oracle port, kind "port") on this box's
                   host cores: B=8, 2 warm-up + 5 timed iterations; N=1 only
  gpu_eager_baseline  the same unmodified reference code with device='cuda' (eager PyTorch/cuDNN, default fp32 flags) at the
                   GPU arm's batch: the GPU bar of BASELINE.md section 5.4; N=1 only
`--impl reference` times the reference CPU leg alone (same config/metric/unit).  `--check` (N >= 2) verifies the data-parallel
path instead of timing it: sharded gradients == global-batch gradients, replicas stay bit-identical.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (img_size, channels, reg_param, au_lr, im_lr, mapper_lr, algorithmic GFLOP/episode/iteration, description).
    # The GFLOP figure is SURVEY.md section 8d's *reduced* one (281 O / 494 V-R1 instead of 305.83 / 529.05): the G-step does not compute
    # the authenticator's weight gradients, which the reference computes and discards.
    "O": (32, 1, 0.0, 1e-6, 1e-5, 1e-7, 281.0, "GIM Omniglot-shaped 1x32x32 m=n=k=5 reg=0 (BASELINE configs[1] at the reference's Omniglot resolution)"),
    "V": (64, 3, 10.0, 1e-4, 1e-4, 1e-6, 494.0, "GIM VoxCeleb2-shaped 3x64x64 m=n=k=5 R1 reg=10 (BASELINE configs[2])"),
}
DEFAULT_BATCH = {"O": 128, "V": 128}
M_, N_, K_ = 5, 5, 5
STYLE = 512
METRIC = "GIM train episodes/sec (fwd+bwd G+D)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="O", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU per step (default 128)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--check", action="store_true", help="verify the data-parallel path (run under torchrun with N >= 2)")
    return ap.parse_args()


def config_of(workload, batch, world):
    """The static description of the workload: identical in the GPU arm and in the reference arm."""
    return {"workload": WORKLOADS[workload][7], "episodes_per_gpu_per_step": batch, "m": M_, "n": N_, "k": K_, "style_dim": STYLE,
            "parallelism": "dp%d" % world, "algorithmic_gflop_per_episode": WORKLOADS[workload][6],
            "cache": "working set (activations of %d images/step) >> 126 MB L2; 2 rotating input batches" % (batch * 45)}


def synth_batch(b, ch, size, seed, device, pin=False):
    """U(-1, 1) images, generator-seeded."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = []
    for s in (M_, N_, K_):
        t = torch.rand((b, s, ch, size, size), generator=g) * 2 - 1
        out.append(t.pin_memory() if pin else t.to(device))
    return out        # leaked, real, si


# ----------------------------------------------------------------------------------------------------------------------
# reference legs: the UNMODIFIED reference (oracle/_ref/reference.zip, packed by oracle/build_ref.py) or, when it is absent,
# the oracle port of the same training iteration.  Test infrastructure used as the *measured baseline* only.
# ----------------------------------------------------------------------------------------------------------------------
def reference_iteration_factory(workload, batch, device):
    """-> (iteration(), kind).  kind 'reference': training/gim_img_trainer.py + gim_img_training.py:157-183 of the reference itself."""
    import torch
    from oracle import ref_shim
    size, ch, reg, au_lr, im_lr, map_lr = WORKLOADS[workload][:6]
    leaked, real, si = synth_batch(batch, ch, size, 1234, device)
    if ref_shim.available() is not None:
        ref_shim.install()
        import models.gim_img_models as RM
        import training.gim_img_training as RL
        from training.gim_img_trainer import GIMImgTrainer as RefTrainer
        from training.utils import DataParallelMock as RefMock
        torch.manual_seed(1)
        au, im = RM.get_au(size, ch, STYLE).to(device), RM.get_im(size, ch, STYLE).to(device)
        tr = RefMock(RefTrainer(tempfile.mkdtemp(prefix="gim_ref_"), M_, N_, K_, au, im, au_lr, im_lr, map_lr, reg_param=reg).to(device))

        def iteration():
            tr.module.do_global_step()
            tr.module.update_learning_rate()
            _, fake, _ = RL.im_train_step(tr, leaked, si)
            o = RL.au_train_step(tr, real, fake, si)
            return o[0]
        return iteration, "reference"

    from oracle import gim_oracle as O
    from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
    torch.manual_seed(1)
    au, im = M.get_au(size, ch, STYLE), M.get_im(size, ch, STYLE)       # same initialisation as the GPU arm
    pa = {k: v.detach().clone().to(device) for k, v in au.state_dict().items()}
    pi = {k: v.detach().clone().to(device) for k, v in im.state_dict().items()}
    a_names = [k for k, _ in au.named_parameters()]
    i_names = [k for k, _ in im.named_parameters()]
    for n_ in a_names:
        pa[n_].requires_grad_()
    for n_ in i_names:
        pi[n_].requires_grad_()
    st_a = {"step": 0, "m": [torch.zeros_like(pa[x]) for x in a_names], "v": [torch.zeros_like(pa[x]) for x in a_names]}
    st_i = {"step": 0, "m": [torch.zeros_like(pi[x]) for x in i_names], "v": [torch.zeros_like(pi[x]) for x in i_names]}

    def iteration():
        for v in list(pa.values()) + list(pi.values()):
            v.grad = None
        z = torch.randn((batch, N_, STYLE), device=device)
        fake = O.impersonator(pi, leaked, N_, z)
        O.gan_loss(O.authenticator(pa, fake, si), 1.0).mean().backward()
        O.adam_step([pi[x] for x in i_names], [pi[x].grad for x in i_names], st_i, im_lr, 0.0, 0.99)
        for v in pa.values():
            v.grad = None
        out = O.img_authenticator_forward(pa, fake.detach(), real.clone(), si.clone(), reg)
        out[0].mean().backward()
        O.adam_step([pa[x] for x in a_names], [pa[x].grad for x in a_names], st_a, au_lr, 0.0, 0.99)
        return out[0].mean().detach()
    return iteration, "port"


def time_reference(workload, batch, steps, warmup, device="cpu"):
    """-> (baseline dict, seconds per iteration)."""
    import torch
    on_gpu = str(device).startswith("cuda")
    cores = os.cpu_count() or 1
    if not on_gpu:
        torch.set_num_threads(cores)
    it, kind = reference_iteration_factory(workload, batch, device)
    sync = torch.cuda.synchronize if on_gpu else (lambda: None)
    for _ in range(warmup):
        it()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        it()
    sync()
    dt = time.perf_counter() - t0
    what = ("the reference's own GIMImgTrainer + im_train_step/au_train_step (oracle/_ref/reference.zip, unmodified)" if kind == "reference"
            else "oracle/gim_oracle.py port of the iteration")
    if on_gpu:
        sample = "%d iterations of %d episodes, %d warm-up: %s on cuda, eager PyTorch/cuDNN, default flags (fp32 storage, cudnn.allow_tf32=%s, matmul tf32=%s)" % (
            steps, batch, warmup, what, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        return {"value": batch * steps / dt, "unit": "episodes/s", "kind": kind, "sample": sample, "ms_per_step": dt / steps * 1e3,
                "episodes_per_step": batch}, dt / steps
    sample = "%d iterations of %d episodes, %d warm-up: %s, torch CPU fp32, all host threads" % (steps, batch, warmup, what)
    return {"value": batch * steps / dt, "unit": "episodes/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample}, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb, sec = time_reference(args.workload, args.cpu_batch, steps, warmup)
    B = args.batch or DEFAULT_BATCH[args.workload]
    return {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "episodes/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_of(args.workload, B, max(1, args.gpus)),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py --impl ours needs a CUDA device: the GIM hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K steps bracketed by barrier + synchronize on both sides, CUDA events, max over ranks -> milliseconds."""
        import torch
        import torch.distributed as dist
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())


def measure_workload(args, D, workload, B, want_roofline, sample_clocks):
    """Build the trainer for one workload, capture the whole iteration as a CUDA graph, time `value` and `e2e`."""
    import torch
    import optimalstrategiesagainstgenerativeattacks_b200 as gim
    from optimalstrategiesagainstgenerativeattacks_b200 import _cabi, ddp, ops
    from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
    from optimalstrategiesagainstgenerativeattacks_b200.gim_img_trainer import GIMImgTrainer
    from optimalstrategiesagainstgenerativeattacks_b200.training_steps import au_train_step, finish_deferred_steps, im_train_step
    from optimalstrategiesagainstgenerativeattacks_b200.utils import DataParallelMock

    world, rank, dev = D.world, D.rank, D.dev
    size, ch, reg, au_lr, im_lr, map_lr, gflop_ep, desc = WORKLOADS[workload]
    gim.set_precision(args.precision)
    torch.manual_seed(1)                          # train_gim_on_imgs.py:6
    au, im = M.get_au(size, ch, STYLE).to(dev), M.get_im(size, ch, STYLE).to(dev)
    outdir = tempfile.mkdtemp(prefix="gim_bench_")
    trainer = DataParallelMock(GIMImgTrainer(outdir, M_, N_, K_, au, im, au_lr, im_lr, map_lr, reg_param=reg))
    if world > 1:
        ddp.attach(trainer.module.authenticator_opt)
        ddp.attach(trainer.module.impersonator_opt, defer=not os.environ.get("GIM_DDP_NO_OVERLAP"))     # G's all-reduce under the D-step
    torch.manual_seed(1000 + rank)                # per-rank noise stream for z

    n_pool = 2
    dev_pool = [synth_batch(B, ch, size, 1234 + 97 * rank + i, dev) for i in range(n_pool)]
    host_pool = [synth_batch(B, ch, size, 1234 + 97 * rank + i, dev, pin=True) for i in range(n_pool)]
    h2d_bytes = sum(t.numel() * 4 for t in host_pool[0])

    def iteration(leaked, real, si):
        trainer.module.do_global_step()
        trainer.module.update_learning_rate()
        im_loss, fake, _ = im_train_step(trainer, leaked, si)
        o = au_train_step(trainer, real, fake, si)
        finish_deferred_steps(trainer)
        return im_loss, o[0]

    graphed = None
    if not args.no_graph:
        from optimalstrategiesagainstgenerativeattacks_b200.cuda_graph import GraphedIteration
        graphed = GraphedIteration(trainer, *dev_pool[0], warmup=3)      # whole iteration = one CUDA graph

    def run(leaked, real, si):
        if graphed is not None:
            out = graphed(leaked, real, si)
            return out[0], out[1]
        return iteration(leaked, real, si)

    def step_resident(s):
        run(*dev_pool[s % n_pool])

    d2h = torch.empty(2, dtype=torch.float32).pin_memory()

    def step_e2e(s):
        if graphed is not None:
            im_loss, au_loss = run(*host_pool[s % n_pool])               # pinned host -> static device inputs, then replay
        else:
            leaked, real, si = (t.to(dev, non_blocking=True) for t in host_pool[s % n_pool])
            im_loss, au_loss = iteration(leaked, real, si)
        d2h.copy_(torch.stack((im_loss, au_loss)), non_blocking=False)      # the step's result is read on the host

    for s in range(max(3, args.warmup)):
        step_resident(s)
    sampler = ClockSampler(D.local) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    _cabi.launch_count(reset=True)
    ms_total = D.timed(step_resident, args.steps)
    launches = graphed.launches_per_replay * args.steps if graphed is not None else _cabi.launch_count()
    clocks = sampler.stop() if sampler else None
    step_e2e(0)
    ms_e2e = D.timed(step_e2e, args.steps)
    res = {"workload": workload, "desc": desc, "B": B, "ms_total": ms_total, "ms_e2e": ms_e2e, "launches": launches, "clocks": clocks,
           "h2d_bytes": h2d_bytes, "graph": graphed is not None, "gflop_ep": gflop_ep,
           "eps": B * world * args.steps / (ms_total * 1e-3), "eps_e2e": B * world * args.steps / (ms_e2e * 1e-3)}

    if want_roofline:
        # dominant-kernel roofline: one extra iteration with CUDA events around every tensor-core conv launch
        ops.set_stream_parallelism(False)         # one stream: a bracketed launch must not share the GPU with the other encoder branch
        ops.conv_profile_begin()
        iteration(*dev_pool[0])                   # eager, so that events can bracket each launch
        torch.cuda.synchronize()
        prof = ops.conv_profile_end()
        ops.set_stream_parallelism(True)
        shapes = prof.pop("_shapes", {})
        if rank == 0 and os.environ.get("GIM_PROFILE_SHAPES"):
            for key, (fl, ms, cnt) in sorted(shapes.items(), key=lambda kv: -kv[1][1])[:48]:
                print("shape[%s] %-14s n=%-6d h=%-3d w=%-3d ci=%-4d co=%-4d k=%d  x%-3d %8.3f ms %7.1f TFLOP/s" % ((workload,) + key + (cnt, ms, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0)),
                      file=sys.stderr)
        res["prof"] = prof
    # release this workload's memory before the next one is built
    del graphed, trainer, au, im, dev_pool
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def measure_gaussian(args, D, d, batch):
    """BASELINE configs[0] / configs[3]: Gaussian GIM training (m=1, n=5, k=10, prior sigma 10, source sigma 1, lr 1e-4), episodes
    synthesised on the device, the whole iteration (synthesis + G-step + D-step + both Adam updates) one CUDA-graph replay."""
    import torch
    import optimalstrategiesagainstgenerativeattacks_b200 as gim
    from optimalstrategiesagainstgenerativeattacks_b200 import ddp, ops
    from optimalstrategiesagainstgenerativeattacks_b200 import gim_gaussian_models as GM
    from optimalstrategiesagainstgenerativeattacks_b200.gim_gaussian_trainer import GIMGaussianTrainer
    from optimalstrategiesagainstgenerativeattacks_b200.gim_gaussian_training import _GraphedGaussianIteration
    from optimalstrategiesagainstgenerativeattacks_b200.utils import DataParallelMock
    m, n, k = 1, 5, 10
    gim.set_precision(args.precision)
    torch.manual_seed(1)
    au, im = GM.get_au(d).to(D.dev), GM.get_im(d).to(D.dev)
    tr = DataParallelMock(GIMGaussianTrainer(tempfile.mkdtemp(prefix="gim_bench_"), m, n, k, au, im, 1e-4, 1e-4, reg_param=0.0))
    if D.world > 1:
        ddp.attach(tr.module.authenticator_opt)
        ddp.attach(tr.module.impersonator_opt, defer=True)
    torch.manual_seed(1000 + D.rank)

    def sample():
        mu, (real, leaked, si) = ops.gaussian_episodes(batch, (n, m, k), d, 10.0, 1.0, D.dev)
        return mu, (leaked, real, si)
    graphed = _GraphedGaussianIteration(tr, sample, warmup=2)
    for _ in range(3):
        graphed()
    steps = max(args.steps, 20)
    ms = D.timed(lambda s: graphed(), steps)
    out = {"workload": "Gaussian GIM d=%d m=1 n=5 k=10" % d, "episodes_per_gpu_per_step": batch, "value": batch * D.world * steps / (ms * 1e-3), "unit": "episodes/s",
           "ms_per_step": ms / steps, "steps": steps, "statistics_bytes_read_per_D_forward": batch * (n + k) * d * 4}
    del graphed, tr, au, im
    torch.cuda.empty_cache()
    return out


def measure_authenticator_105(args, D, batch=32):
    """SURVEY.md section 8d config 2's second number: authenticator-only forward + backward at the native Omniglot resolution 1x105x105
    (the reference Impersonator cannot run there), n = k = 5, 282.5 GFLOP per episode.  Odd stage sizes (105, 52, 26, 13, 6)."""
    import torch
    import optimalstrategiesagainstgenerativeattacks_b200 as gim
    from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
    from optimalstrategiesagainstgenerativeattacks_b200 import ops
    gim.set_precision(args.precision)
    torch.manual_seed(1)
    au = M.get_au(105, 1, STYLE).to(D.dev).train()
    g = torch.Generator().manual_seed(99)
    test, si = ((torch.rand((batch, 5, 1, 105, 105), generator=g) * 2 - 1).to(D.dev) for _ in range(2))

    def step(_):
        au.zero_grad(set_to_none=True)
        loss = ops.BCEWithLogitsFn.apply(au(test, si), 1.0).mean()
        with ops.deferred_weight_grads():
            loss.backward()
    for s in range(3):
        step(s)
    steps = min(args.steps, 10)
    ms = D.timed(step, steps)
    eps = batch * D.world * steps / (ms * 1e-3)
    out = {"workload": "GIMFaceAuthenticator forward+backward, 1x105x105, n=k=5 (eager launches, no graph)", "episodes_per_gpu_per_step": batch, "value": eps,
           "unit": "episodes/s", "ms_per_step": ms / steps, "algorithmic_gflop_per_episode": 282.5, "model_tflops_per_gpu": 282.5e-3 * eps / D.world}
    del au
    torch.cuda.empty_cache()
    return out


def block_of(res, world, steps, warmup):
    return {"workload": res["desc"], "value": res["eps"], "unit": "episodes/s", "ms_per_step": res["ms_total"] / steps, "episodes_per_gpu_per_step": res["B"],
            "global_batch": res["B"] * world, "steps": steps, "warmup": warmup, "cuda_graph": res["graph"],
            "e2e": {"value": res["eps_e2e"], "unit": "episodes/s", "h2d_bytes_per_step": res["h2d_bytes"], "d2h_bytes_per_step": 8,
                    "ms_per_step": res["ms_e2e"] / steps},
            "gpu_launches": res["launches"], "algorithmic_gflop_per_episode": res["gflop_ep"],
            "whole_step_model_tflops_per_gpu": res["gflop_ep"] * 1e-3 * res["eps"] / world}


def run_ours(args):
    import torch
    from optimalstrategiesagainstgenerativeattacks_b200 import _cabi
    D = Dist()
    world, rank = D.world, D.rank
    _cabi.lib()                                   # fail loudly now if the extension is missing
    warm = max(3, args.warmup)
    B = args.batch or DEFAULT_BATCH[args.workload]
    main = measure_workload(args, D, args.workload, B, want_roofline=True, sample_clocks=True)

    secondary = None
    if not args.no_secondary and args.workload == "O" and args.precision == "bf16":
        secondary = {}
        weak = measure_workload(args, D, "V", DEFAULT_BATCH["V"], want_roofline=False, sample_clocks=False)
        secondary.update(block_of(weak, world, args.steps, warm))
        secondary["scaling"] = "weak"
        if world == 1:
            secondary["strong_scaling_global_batch_128"] = dict(secondary, scaling="strong")      # at N = 1 the two rows coincide
        elif 128 % world == 0:
            strong = measure_workload(args, D, "V", 128 // world, want_roofline=False, sample_clocks=False)
            secondary["strong_scaling_global_batch_128"] = block_of(strong, world, args.steps, warm)
            secondary["strong_scaling_global_batch_128"]["scaling"] = "strong"

    extra = None
    if not args.no_secondary and args.workload == "O" and args.precision == "bf16":
        extra = {"gaussian": [measure_gaussian(args, D, d, b) for d, b in ((10, 4096), (1000, 4096), (1000, 16384), (1000, 65536))]}
        if world == 1:
            extra["authenticator_105x105"] = measure_authenticator_105(args, D)

    if rank != 0:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "bf16_tflops_sustained of MEASURED_PEAKS.json" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (profiles/, tools/ncu_summarize.py): per launch,
    # for the layer shape that takes most of its time in this workload (128->128 3x3 @32x32 over B*5 = 640 images)
    traffic, traffic_note = None, None
    try:
        for fn in ("ncu_conv_shapes_r02.json", "ncu_conv_shapes_r01.json"):
            path = os.path.join(ROOT, "profiles", fn)
            if os.path.exists(path):
                top = json.load(open(path))["fwd_32x32_128_128_k3"]
                if args.workload == "O" and B == 128:
                    traffic = top["dram_bytes_read"] + top["dram_bytes_write"]
                    traffic_note = ("%s: per launch of the 128->128 3x3 @32x32 layer (640 images): algorithmic %.1f MB, ncu tensor-pipe active %.1f %%, "
                                    "%.0f us under ncu" % (fn, top["algorithmic_bytes"] / 1e6, top["tensor_pipe_active_pct"], top["duration_us"]))
                break
    except Exception:
        pass
    prof = main["prof"]
    tc = prof.get("tcgen05", {"flops": 0.0, "ms": 0.0, "launches": 0})
    achieved = tc["flops"] / (tc["ms"] * 1e-3) / 1e12 if tc["ms"] > 0 else 0.0
    total_conv_ms = sum(v["ms"] for v in prof.values())
    ms_step = main["ms_total"] / args.steps
    line = {
        "metric": METRIC, "value": main["eps"], "unit": "episodes/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": config_of(args.workload, B, world),
        "detail": {"cuda_graph": main["graph"], "executed_conv_gflop_per_episode": sum(v["flops"] for v in prof.values()) / 1e9 / B,
                   "whole_step_model_tflops_per_gpu": main["gflop_ep"] * 1e-3 * main["eps"] / world,
                   "whole_step_frac_of_sustained_peak": main["gflop_ep"] * 1e-3 * main["eps"] / world / peak_tf},
        "e2e": {"value": main["eps_e2e"], "unit": "episodes/s", "h2d_bytes_per_step": main["h2d_bytes"], "d2h_bytes_per_step": 8,
                "ms_per_step": main["ms_e2e"] / args.steps},
        "gpu_launches": main["launches"],
        "clocks": main["clocks"],
        "roofline": {"bound": "tensor", "kernel": "conv_fwd_tc2_kernel / conv_fwd_tc_kernel (tcgen05 implicit GEMM: conv forward + input-gradient)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic, "traffic_note": traffic_note,
                     "peak_source": peak_src, "launches_per_step": tc["launches"], "kernel_ms_per_step": tc["ms"],
                     "share_of_step": tc["ms"] / ms_step if ms_step else None,
                     "other_conv_kernels_ms_per_step": {k: v["ms"] for k, v in prof.items() if k != "tcgen05"},
                     "all_conv_ms_per_step": total_conv_ms},
    }
    if secondary is not None:
        line["secondary"] = secondary
    if extra is not None:
        line["other_configs"] = extra
    if world == 1:                                        # baselines are reported on rank 0 at N=1 only
        if not args.no_gpu_baseline:
            try:
                gb, _ = time_reference(args.workload, B, 5, 2, device=D.dev)
                gb["speedup_ours_over_gpu_eager"] = main["eps"] / gb["value"] if gb["value"] else None
                line["gpu_eager_baseline"] = gb
            except Exception as e:                        # the baseline must never take the bench line down
                line["gpu_eager_baseline"] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
            torch.cuda.empty_cache()
        if not args.no_cpu_baseline:
            cb, _ = time_reference(args.workload, args.cpu_batch, 5, 2)
            line["cpu_baseline"] = cb
    return line


def run_check(args):
    """Data-parallel correctness instead of timing (launch under torchrun, N >= 2), fp32 parity path:
      1. the all-reduced mean of the per-rank gradients over B/N episodes each == the gradient of the global batch of B episodes,
         computed on every rank alone (G-step and D-step): per tensor, median rel <= 1e-5 and max rel <= 1e-4 (fp32 sums over different
         groupings of the same terms).  The InstanceNorm biases get small random values first: at the reference's initialisation (bias 0)
         the EnvDecoder's maps are spatially constant, every InstanceNorm has variance 0 and multiplies the gradient by rsqrt(eps) = 316, and
         the bias gradients in front of the norms (exactly zero in exact arithmetic) become 316^k-amplified round-off in ANY implementation;
      2. after 3 graph-replayed training iterations every rank holds bit-identical parameters (checksums all-gathered)."""
    import torch
    import torch.distributed as dist
    import optimalstrategiesagainstgenerativeattacks_b200 as gim
    from optimalstrategiesagainstgenerativeattacks_b200 import _cabi, ddp
    from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
    from optimalstrategiesagainstgenerativeattacks_b200.cuda_graph import GraphedIteration
    from optimalstrategiesagainstgenerativeattacks_b200.gim_img_trainer import GIMImgTrainer
    from optimalstrategiesagainstgenerativeattacks_b200.utils import DataParallelMock
    D = Dist()
    world, rank, dev = D.world, D.rank, D.dev
    if world < 2:
        raise RuntimeError("--check needs torchrun with at least 2 ranks")
    size, ch, reg, au_lr, im_lr, map_lr = WORKLOADS[args.workload][:6]
    B = args.batch or 4 * world
    lo, hi = ddp.shard_range(B, rank, world)
    gim.set_precision("fp32")
    gim.set_deterministic(True)
    leaked, real, si = synth_batch(B, ch, size, 4321, dev)
    z = torch.randn((B, N_, STYLE), generator=torch.Generator().manual_seed(7)).to(dev)
    real_randn = torch.randn
    worst = {}

    def grads_of(shard):
        torch.manual_seed(1)
        au, im = M.get_au(size, ch, STYLE).to(dev), M.get_im(size, ch, STYLE).to(dev)
        with torch.no_grad():
            gen = torch.Generator().manual_seed(11)
            for n_, p in im.named_parameters():
                if n_.endswith(("in1.bias", "in2.bias")) or (".in_layers." in n_ and n_.endswith(".bias")):
                    p.copy_((0.1 * torch.randn(p.shape, generator=gen)).to(dev))
        tr = GIMImgTrainer(tempfile.mkdtemp(prefix="gim_check_"), M_, N_, K_, au, im, au_lr, im_lr, map_lr, reg_param=reg)
        sl = slice(lo, hi) if shard else slice(0, B)
        torch.randn = lambda *a, **k: z[sl].clone()
        try:
            loss, fake, _ = tr.impersonator_forward(leaked[sl], si[sl])
        finally:
            torch.randn = real_randn
        loss.mean().backward()
        g_im = [p.grad.clone() for p in im.parameters() if p.grad is not None]
        out = tr.authenticator_forward(fake.detach(), real[sl].clone(), si[sl].clone())
        au.zero_grad()
        out[0].mean().backward()
        g_au = [p.grad.clone() for p in au.parameters() if p.grad is not None]
        return g_im, g_au

    full = grads_of(False)
    mine = grads_of(True)
    for name, gs_full, gs_mine in (("impersonator", full[0], mine[0]), ("authenticator", full[1], mine[1])):
        flat = torch.cat([g.flatten() for g in gs_mine])
        dist.all_reduce(flat)
        flat /= world
        off, errs = 0, []
        scale = max(float(g.norm()) for g in gs_full)
        for g in gs_full:
            if float(g.norm()) >= 1e-6 * scale:           # (identically-zero true gradients -- biases in front of norms -- carry no signal)
                errs.append(float((flat[off:off + g.numel()] - g.flatten()).norm()) / float(g.norm()))
            off += g.numel()
        errs.sort()
        worst[name] = {"median": errs[len(errs) // 2], "max": errs[-1], "tensors": len(errs)}
    # replicas stay identical through the path bench.py times (graph + NCCL all-reduce inside it)
    gim.set_precision(args.precision)
    gim.set_deterministic(False)
    torch.manual_seed(1)
    au, im = M.get_au(size, ch, STYLE).to(dev), M.get_im(size, ch, STYLE).to(dev)
    tr = DataParallelMock(GIMImgTrainer(tempfile.mkdtemp(prefix="gim_check_"), M_, N_, K_, au, im, 1e-3, 1e-3, 1e-4, reg_param=reg))
    ddp.attach(tr.module.authenticator_opt)
    ddp.attach(tr.module.impersonator_opt, defer=True)
    torch.manual_seed(1000 + rank)
    g = GraphedIteration(tr, leaked[lo:hi], real[lo:hi], si[lo:hi], warmup=3)
    for _ in range(3):
        g()
    torch.cuda.synchronize()
    sums = torch.tensor([float(sum(p.double().sum() for p in net.parameters())) for net in (au, im)], dtype=torch.float64, device=dev)
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    identical = all(torch.equal(gathered[0], t) for t in gathered)
    ok = identical and all(v["median"] <= 1e-5 and v["max"] <= 1e-4 for v in worst.values())
    if rank != 0:
        return None
    return {"check": "ok" if ok else "FAILED", "n_gpus": world, "global_batch": B, "workload": WORKLOADS[args.workload][7],
            "sharded_vs_global_batch_gradient_rel_err": worst, "tolerance": {"median": 1e-5, "max": 1e-4},
            "replica_parameter_checksums_identical_after_3_graph_steps": identical,
            "checksums": [t.tolist() for t in gathered]}


def _finish(world):
    """Leave a multi-rank run without tearing NCCL down: destroying a communicator that CUDA graphs still reference can block
    forever, and nothing after the JSON line needs it.  Ranks meet at a last barrier, flush, and exit."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    args = parse()
    # a run that stops making progress must fail loudly with the stack of every thread instead of holding the GPU box until its limit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("GIM_BENCH_WATCHDOG_S", "900")), exit=True)
    real_stdout = sys.stdout
    with contextlib.redirect_stdout(sys.stderr):      # trainers print parameter counts etc.: stdout carries the JSON line only
        if args.impl == "reference":
            line = run_reference(args)
        elif args.check:
            line = run_check(args)
        else:
            line = run_ours(args)
    if line is not None:
        print(json.dumps(line), file=real_stdout, flush=True)
    _finish(int(os.environ.get("WORLD_SIZE", "1")) if args.impl != "reference" else 1)


if __name__ == "__main__":
    main()
