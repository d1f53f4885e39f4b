"""Whole-iteration CUDA graph: one attacker step + one authenticator step (forward, backward, fused Adam, spectral-norm power
iterations, noise sampling) captured once and replayed, so the ~1.7 k kernel launches of an iteration cost no host time.

Every kernel behind the C ABI is capture-safe (no allocation, no synchronisation, stream-ordered memsets; TMA descriptors are
by-value kernel parameters), FusedAdam keeps its step counter and learning rates on the device, and gradients are zeroed in
place, so the eager trainer API (`im_train_step` / `au_train_step`, reference training/gim_img_training.py:157-183) is captured
unchanged.  Host-side bookkeeping (global step, LR scheduler) runs before each replay.

Warm-up is side-effect free: the iterations that build optimizer state, pointer tables and kernel attributes before the capture
run on a snapshot -- parameters, spectral-norm u/v, Adam moments and step counts, LR schedulers, the global step and the CUDA RNG
state are put back afterwards (in place: the addresses baked into the graph do not change) -- so the first replay is training
iteration 0 of the reference loop, not iteration 4.
"""
import copy

import torch

from . import _cabi
from .training_steps import au_train_step, finish_deferred_steps, im_train_step


class _TrainingState:
    """In-place snapshot / restore of everything a training iteration mutates."""

    def __init__(self, module):
        self.module = module
        self.tensors = [(t, t.detach().clone()) for t in list(module.parameters()) + list(module.buffers())]
        self.opts = []
        for opt in (module.authenticator_opt, module.impersonator_opt):
            saved = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in st.items()} for p, st in opt.state.items()}
            self.opts.append((opt, saved, getattr(opt, "_steps_done", None), [dict((k, v) for k, v in g.items() if k != "params") for g in opt.param_groups]))
        self.scheds = [(s, copy.deepcopy(s.state_dict())) for s in (getattr(module, "au_scheduler", None), getattr(module, "im_scheduler", None)) if s is not None]
        self.global_step = module.get_global_step()
        self.rng = torch.cuda.get_rng_state()
        self.cpu_rng = torch.get_rng_state()

    def restore(self):
        m = self.module
        with torch.no_grad():
            for t, saved in self.tensors:
                t.copy_(saved)
            for opt, saved, steps_done, groups in self.opts:
                for p, st in opt.state.items():
                    old = saved.get(id(p))
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            if old is not None and k in old:
                                v.copy_(old[k])
                            else:
                                v.zero_()              # state created by the warm-up: back to "never stepped"
                        elif old is not None and k in old:
                            st[k] = old[k]
                for g, old_g in zip(opt.param_groups, groups):
                    g.update(old_g)
                if hasattr(opt, "_steps_done"):
                    first = steps_done if steps_done is not None else max([int(float(s.get("step", 0))) for s in saved.values()] or [0])
                    opt._steps_done = first
                    if getattr(opt, "_step_dev", None) is not None:
                        opt._step_dev.fill_(first)
                    opt._lrs_host = None               # force the next sync_lrs() to upload the restored learning rates
        for s, st in self.scheds:
            s.load_state_dict(st)
        m._global_step.set(self.global_step)
        torch.cuda.set_rng_state(self.rng)
        torch.set_rng_state(self.cpu_rng)


class GraphedIteration:
    def __init__(self, trainer, leaked, real, si, warmup=3):
        """`trainer`: DataParallelMock(GIMImgTrainer | GIMGaussianTrainer); leaked/real/si: example CUDA batches (shapes are baked in)."""
        self.trainer = trainer
        self.static_in = [t.clone() for t in (leaked, real, si)]
        self.graph = None
        self.static_out = None
        self.launches_per_replay = 0
        self._capture(warmup)

    def _host_bookkeeping(self):
        m = self.trainer.module
        m.do_global_step()
        if hasattr(m, "update_learning_rate"):
            m.update_learning_rate()
        for opt in (m.authenticator_opt, m.impersonator_opt):
            opt.sync_lrs()

    def _body(self):
        leaked, real, si = self.static_in
        im_loss, fake, _ = im_train_step(self.trainer, leaked, si)
        out = au_train_step(self.trainer, real, fake, si)
        finish_deferred_steps(self.trainer)              # data parallel: G's all-reduce ran underneath the D-step; its Adam update goes here
        return (im_loss,) + tuple(out[:6])

    def _capture(self, warmup):
        m = self.trainer.module
        state = _TrainingState(m)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):              # builds optimizer state / pointer tables, sets kernel attributes
                self._host_bookkeeping()
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        state.restore()
        for opt in (m.authenticator_opt, m.impersonator_opt):
            opt.sync_lrs()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = _cabi.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_out = self._body()
        self.launches_per_replay = _cabi.launch_count() - before      # kernels recorded in the graph (capture does not execute)
        del state

    def __call__(self, leaked=None, real=None, si=None):
        """Run one iteration; new batches are copied into the static input buffers (pass None to reuse them).
        Returns (im_loss, au_loss, loss_on_real, loss_on_fake, reg, out_on_real, out_on_fake) as device scalars."""
        with torch.no_grad():                             # R1 marks the static `real` buffer requires_grad in place
            for dst, src in zip(self.static_in, (leaked, real, si)):
                if src is not None:
                    dst.copy_(src, non_blocking=True)
        self._host_bookkeeping()
        self.graph.replay()
        for opt in (self.trainer.module.authenticator_opt, self.trainer.module.impersonator_opt):
            opt.note_graph_steps(1)
        return self.static_out
