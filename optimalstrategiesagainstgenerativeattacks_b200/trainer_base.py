"""Shared machinery of the two GIM trainers (image and Gaussian).

The reference has two near-identical classes (training/gim_img_trainer.py, training/gim_gaussian_trainer.py); here the parts that
do not depend on the networks live once: the game sizes, mode dispatch, the per-episode BCE loss on the kernels, the adversarial
loss assembly (real / fake / R1), checkpoint registration and the global step.  Subclasses provide `_authenticator_outputs`
(how D is evaluated on real and fake test samples) and their optimizers.  Public names, arguments and return tuples are the
reference's: `forward(mode, **kwargs)`, `authenticator_forward`, `impersonator_forward`,
`impersonator_sample`, `gan_loss`, `save`, `resume_from_ckpt`, `get_global_step`, `do_global_step`, `global_step`,
members `authenticator`, `impersonator`, `authenticator_opt`, `impersonator_opt`, `checkpoint_io`.
"""
import contextlib
import os

import torch
import torch.nn as nn

from . import ops
from .checkpoints import CheckpointIO
from .utils import GlobalStep, compute_grad2, num_parameters


@contextlib.contextmanager
def frozen(module):
    """Inside, the module's parameters do not require grad (graphs recorded inside treat them as constants)."""
    touched = [p for p in module.parameters() if p.requires_grad]
    for p in touched:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in touched:
            p.requires_grad_(True)


class GIMTrainerBase(nn.Module):
    CHECKPOINT_DIR = "ckpts"
    MODES = ("authenticator_forward", "impersonator_forward", "impersonator_sample")

    def __init__(self, outdir, m, n, k, authenticator, impersonator, reg_param, remove_noise_mean):
        super().__init__()
        self.m, self.n, self.k = m, n, k
        self.authenticator, self.impersonator = authenticator, impersonator
        self.reg_param, self.remove_noise_mean = reg_param, remove_noise_mean
        self._global_step = GlobalStep()
        self.checkpoint_dir = os.path.join(outdir, self.CHECKPOINT_DIR)
        for role, net in (("Authenticator", authenticator), ("impersonator", impersonator)):
            print("{} has {} parameters".format(role, num_parameters(net.parameters())))

    def _register_checkpointables(self):
        """Call once the optimizers exist: same key set as the reference's checkpoints."""
        self.checkpoint_io = CheckpointIO(checkpoint_dir=self.checkpoint_dir)
        self.checkpoint_io.register_modules(authenticator=self.authenticator, impersonator=self.impersonator,
                                            authenticator_opt=self.authenticator_opt, impersonator_opt=self.impersonator_opt,
                                            global_step=self._global_step)

    # ---- dispatch (nn.DataParallel-style single entry point of the reference) ----
    def forward(self, mode, **kwargs):
        if mode not in self.MODES:
            raise ValueError("unsupported mode")
        return getattr(self, mode)(**kwargs)

    # ---- losses ----
    def gan_loss(self, dis_out, target, reduce=False):
        """binary_cross_entropy_with_logits against a constant target, one value per episode (mean if `reduce`)."""
        per_episode = ops.BCEWithLogitsFn.apply(dis_out, float(target))
        return per_episode.mean() if reduce else per_episode.squeeze()

    def _authenticator_outputs(self, fake_sample, real_sample, si_sample, second_order):
        """-> (logits on real, logits on fake); `second_order`: the real/si branch will be differentiated twice (R1)."""
        raise NotImplementedError

    def authenticator_forward(self, fake_sample, real_sample, si_sample, grad=True):
        """D's loss: BCE(real -> 1) + BCE(fake -> 0) + reg_param * R1(real, si).  Returns the reference's 9-tuple
        (loss, loss_on_real, loss_on_fake, reg, out_on_real, out_on_fake, pred_on_real, pred_on_fake, fake_sample)."""
        r1 = self.reg_param > 0
        if r1:                                           # the reference marks both inputs even when grad=False
            real_sample.requires_grad_()
            si_sample.requires_grad_()
        out_on_real, out_on_fake = self._authenticator_outputs(fake_sample, real_sample, si_sample, second_order=r1 and grad)
        loss_on_real = self.gan_loss(dis_out=out_on_real, target=1.)
        loss_on_fake = self.gan_loss(dis_out=out_on_fake, target=0.)
        reg = self.reg_param * compute_grad2(out_on_real, (real_sample, si_sample)) if (r1 and grad) else torch.zeros_like(loss_on_real)
        with torch.no_grad():
            pred_on_real, pred_on_fake = out_on_real.detach() >= 0, out_on_fake.detach() >= 0
        return (loss_on_real + loss_on_fake + reg, loss_on_real.detach(), loss_on_fake.detach(), reg, out_on_real.detach(), out_on_fake.detach(),
                pred_on_real, pred_on_fake, fake_sample.detach())

    def _attack(self, leaked_sample):
        return self.impersonator(leaked_sample=leaked_sample, n=self.n, remove_noise_mean=self.remove_noise_mean)

    def _judge_attack(self, fake_sample, si_sample):
        return self.authenticator(test_sample=fake_sample, si_sample=si_sample)

    def impersonator_forward(self, leaked_sample, si_sample):
        """G's loss: BCE(D(fake, si) -> 1).  Returns (loss per episode, fake_sample, D's logits)."""
        fake_sample = self._attack(leaked_sample)
        auth_out = self._judge_attack(fake_sample, si_sample)
        return self.gan_loss(dis_out=auth_out, target=1.), fake_sample, auth_out

    def impersonator_sample(self, leaked_sample):
        with torch.no_grad():
            return self._attack(leaked_sample)

    # ---- bookkeeping ----
    def get_global_step(self):
        return self._global_step.get()

    def do_global_step(self):
        return self._global_step.step()

    @property
    def global_step(self):
        return self.get_global_step()

    def resume_from_ckpt(self, ckpt_path):
        self.checkpoint_io.load(ckpt_path)
        print('Resuming training from iteration {}'.format(self.get_global_step()))

    def _save(self, last_epoch):
        print("\nSaving checkpoint...\n")
        step = self.get_global_step()
        self.checkpoint_io.save(global_step=step, last_epoch=last_epoch, filename="model_{:08}.pt".format(step))
