"""Whole-iteration CUDA graph: one attacker step + one authenticator step (forward, backward, fused Adam, spectral-norm power
iterations, noise sampling) captured once and replayed, so the ~14 k kernel launches of an iteration cost no host time.

Every kernel behind the C ABI is capture-safe (no allocation, no synchronisation, stream-ordered memsets; TMA descriptors are
by-value kernel parameters), FusedAdam keeps its step counter and learning rates on the device, and gradients are zeroed in
place, so the eager trainer API (`im_train_step` / `au_train_step`, reference training/gim_img_training.py:157-183) is captured
unchanged.  Host-side bookkeeping (global step, LR scheduler) runs before each replay.
"""
import torch

from . import _cabi
from .training_steps import au_train_step, im_train_step


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
        return (im_loss,) + tuple(out[:6])

    def _capture(self, warmup):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # builds optimizer state / pointer tables, sets kernel attributes
                self._host_bookkeeping()
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._host_bookkeeping()
        self.graph = torch.cuda.CUDAGraph()
        before = _cabi.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_out = self._body()
        self.launches_per_replay = _cabi.launch_count() - before      # kernels recorded in the graph (capture does not execute)

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
