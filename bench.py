#!/usr/bin/env python
"""bench.py -- GIM train episodes/sec (fwd+bwd G+D) on N B200s of one node (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload O|V] [--batch B] [--precision bf16|fp32]

A "step" is one training iteration = one attacker (G) step + one authenticator (D) step over B episodes per GPU through the
kept trainer API (GIMImgTrainer + im_train_step / au_train_step).  Default workload = BASELINE.json configs[1]: synthetic
Omniglot-shaped episodes, m=n=k=5, at the reference's own Omniglot resolution 1x32x32 (the reference Impersonator cannot run
at 105x105), reg 0, Adam(0, 0.99), lrs 1e-6/1e-5/1e-7, bf16 tensor-core path.

Rank 0 prints ONE JSON line (see the contract in the task statement): `value` = whole-job episodes/s with inputs resident in
HBM; `e2e` = the same through the public API with pinned-host inputs copied in and the losses read back every step;
`roofline` = the dominant kernel (tcgen05 implicit-GEMM conv) against the measured bf16 peak; `cpu_baseline` = the CPU oracle
port of the same step timed on this box's host cores on a bounded sample.  `--impl reference` times that CPU port alone.
"""
import argparse
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
M_, N_, K_ = 5, 5, 5
STYLE = 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="O", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU per step (default 128 for O, 32 for V)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    return ap.parse_args()


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
# CPU leg: the oracle port of the same training iteration (test infrastructure used as the *measured baseline* only)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_iteration_factory(workload, batch):
    import torch
    from oracle import gim_oracle as O
    from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
    size, ch, reg, au_lr, im_lr, map_lr = WORKLOADS[workload][:6]
    torch.manual_seed(1)
    au, im = M.get_au(size, ch, STYLE), M.get_im(size, ch, STYLE)       # same initialisation as the GPU arm (CPU tensors)
    pa = {k: v.detach().clone() for k, v in au.state_dict().items()}
    pi = {k: v.detach().clone() for k, v in im.state_dict().items()}
    a_names = [k for k, _ in au.named_parameters()]
    i_names = [k for k, _ in im.named_parameters()]
    for n_ in a_names:
        pa[n_].requires_grad_()
    for n_ in i_names:
        pi[n_].requires_grad_()
    st_a = {"step": 0, "m": [torch.zeros_like(pa[x]) for x in a_names], "v": [torch.zeros_like(pa[x]) for x in a_names]}
    st_i = {"step": 0, "m": [torch.zeros_like(pi[x]) for x in i_names], "v": [torch.zeros_like(pi[x]) for x in i_names]}
    leaked, real, si = synth_batch(batch, ch, size, 1234, "cpu")

    def iteration():
        for v in list(pa.values()) + list(pi.values()):
            v.grad = None
        z = torch.randn((batch, N_, STYLE))
        fake = O.impersonator(pi, leaked, N_, z)
        O.gan_loss(O.authenticator(pa, fake, si), 1.0).mean().backward()
        O.adam_step([pi[x] for x in i_names], [pi[x].grad for x in i_names], st_i, im_lr, 0.0, 0.99)
        for v in pa.values():
            v.grad = None
        out = O.img_authenticator_forward(pa, fake.detach(), real.clone(), si.clone(), reg)
        out[0].mean().backward()
        O.adam_step([pa[x] for x in a_names], [pa[x].grad for x in a_names], st_a, au_lr, 0.0, 0.99)
        return float(out[0].mean())
    return iteration


def time_cpu(workload, batch, steps, warmup):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    it = cpu_iteration_factory(workload, batch)
    for _ in range(warmup):
        it()
    t0 = time.perf_counter()
    for _ in range(steps):
        it()
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "unit": "episodes/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d iterations of %d episodes (oracle/gim_oracle.py, torch CPU fp32, all host threads), %d warm-up" % (steps, batch, warmup)}, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 2))
    cb, ms = time_cpu(args.workload, args.cpu_batch, steps, warmup)
    line = {
        "impl": "reference", "metric": "GIM train episodes/sec (fwd+bwd G+D)", "value": cb["value"], "unit": "episodes/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][7], "episodes_per_step": args.cpu_batch, "m": M_, "n": N_, "k": K_, "device": "host CPU"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
def run_ours(args):
    import torch
    import torch.distributed as dist
    import optimalstrategiesagainstgenerativeattacks_b200 as gim
    from optimalstrategiesagainstgenerativeattacks_b200 import _cabi, ddp, ops
    from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
    from optimalstrategiesagainstgenerativeattacks_b200.gim_img_trainer import GIMImgTrainer
    from optimalstrategiesagainstgenerativeattacks_b200.training_steps import au_train_step, im_train_step
    from optimalstrategiesagainstgenerativeattacks_b200.utils import DataParallelMock

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device: the GIM hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()                                   # fail loudly now if the extension is missing

    size, ch, reg, au_lr, im_lr, map_lr, gflop_ep, desc = WORKLOADS[args.workload]
    B = args.batch or (128 if args.workload == "O" else 32)
    gim.set_precision(args.precision)
    torch.manual_seed(1)                          # train_gim_on_imgs.py:6
    au, im = M.get_au(size, ch, STYLE).to(dev), M.get_im(size, ch, STYLE).to(dev)
    outdir = tempfile.mkdtemp(prefix="gim_bench_")
    trainer = DataParallelMock(GIMImgTrainer(outdir, M_, N_, K_, au, im, au_lr, im_lr, map_lr, reg_param=reg))
    if world > 1:
        ddp.attach(trainer.module.authenticator_opt)
        ddp.attach(trainer.module.impersonator_opt)
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
        return im_loss, o[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

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
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _cabi.launch_count(reset=True)
    ms_total = timed(step_resident, args.steps)
    launches = graphed.launches_per_replay * args.steps if graphed is not None else _cabi.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    step_e2e(0)
    ms_e2e = timed(step_e2e, args.steps)

    # dominant-kernel roofline: one extra iteration with CUDA events around every tensor-core conv launch
    ops.set_stream_parallelism(False)             # one stream: a bracketed launch must not share the GPU with the other encoder branch
    ops.conv_profile_begin()
    iteration(*dev_pool[0])                       # eager, so that events can bracket each launch
    torch.cuda.synchronize()
    prof = ops.conv_profile_end()
    ops.set_stream_parallelism(True)
    shapes = prof.pop("_shapes", {})
    if rank == 0 and os.environ.get("GIM_PROFILE_SHAPES"):
        for key, (fl, ms, cnt) in sorted(shapes.items(), key=lambda kv: -kv[1][1])[:40]:
            print("shape %-14s n=%-6d h=%-3d w=%-3d ci=%-4d co=%-4d k=%d  x%-3d %8.3f ms %7.1f TFLOP/s" % (key + (cnt, ms, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0)),
                  file=sys.stderr)

    if rank != 0:
        _finish(world, dev)
        return
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
        ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_conv_shapes_r01.json")))
        top = ncu["fwd_32x32_128_128_k3"]
        if args.workload == "O" and B == 128:
            traffic = top["dram_bytes_read"] + top["dram_bytes_write"]
            traffic_note = ("per launch of the 128->128 3x3 @32x32 layer (640 images): algorithmic %.1f MB, ncu tensor-pipe active %.1f %%, "
                            "%.0f us under ncu" % (top["algorithmic_bytes"] / 1e6, top["tensor_pipe_active_pct"], top["duration_us"]))
    except Exception:
        pass
    tc = prof.get("tcgen05", {"flops": 0.0, "ms": 0.0, "launches": 0})
    achieved = tc["flops"] / (tc["ms"] * 1e-3) / 1e12 if tc["ms"] > 0 else 0.0
    total_conv_ms = sum(v["ms"] for v in prof.values())
    eps_total = B * world * args.steps / (ms_total * 1e-3)
    eps_e2e = B * world * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": "GIM train episodes/sec (fwd+bwd G+D)", "value": eps_total, "unit": "episodes/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": desc, "episodes_per_gpu_per_step": B, "m": M_, "n": N_, "k": K_, "style_dim": STYLE, "parallelism": "dp%d" % world, "cuda_graph": graphed is not None,
                   "cache": "working set (activations of %d images/step) >> 126 MB L2; %d rotating input batches" % (B * 45, n_pool),
                   "algorithmic_gflop_per_episode": gflop_ep,
                   "executed_conv_gflop_per_episode": sum(v["flops"] for v in prof.values()) / 1e9 / B,
                   "whole_step_model_tflops": gflop_ep * 1e-3 * eps_total / world},
        "e2e": {"value": eps_e2e, "unit": "episodes/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_fwd_tc2_kernel / conv_fwd_tc_kernel (tcgen05 implicit GEMM: conv forward + input-gradient)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic, "traffic_note": traffic_note,
                     "peak_source": peak_src, "launches_per_step": tc["launches"], "kernel_ms_per_step": tc["ms"],
                     "share_of_step": tc["ms"] / (ms_total / args.steps) if ms_total else None,
                     "other_conv_kernels_ms_per_step": {k: v["ms"] for k, v in prof.items() if k != "tcgen05"},
                     "all_conv_ms_per_step": total_conv_ms},
    }
    if not args.no_cpu_baseline and world == 1:          # reported on rank 0 at N=1 only
        torch.cuda.empty_cache()
        cb, _ = time_cpu(args.workload, args.cpu_batch, 2, 1)
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    _finish(world, dev)


def _finish(world, dev):
    """Leave a multi-rank run without tearing NCCL down: destroying a communicator that CUDA graphs still reference can block
    forever, and nothing after the JSON line needs it.  Ranks meet at a last barrier, flush, and exit."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize(dev)
    dist.barrier()
    torch.cuda.synchronize(dev)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
