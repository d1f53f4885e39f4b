"""GIMImgTrainer -- the reference's training/gim_img_trainer.py:21-183 surface (constructor arguments, modes, return tuples, optimizer
layout, LR schedule, checkpoint registration) on the shared trainer base; losses and optimizers run on libgim_b200 kernels."""
import contextlib

import torch.optim as optim

from . import ops
from .fused_adam import FusedAdam
from .trainer_base import GIMTrainerBase, frozen

_frozen = frozen        # (kept under its old private name for callers inside the package)


class GIMImgTrainer(GIMTrainerBase):
    # G's parameter groups in the reference's order (:50-57); the noise mapper is the only one with its own learning rate
    G_GROUPS = ("src_encoder", "env_encoder", "env_decoder", "img2img", "img_att", "env_noise_mapper")

    def __init__(self, outdir, m, n, k, authenticator, impersonator, au_lr, im_lr, env_noise_mapping_lr,
                 beta1=0., beta2=0.99, lr_milestones=(), lr_gamma=0.3, reg_param=10., remove_noise_mean=True):
        super().__init__(outdir, m, n, k, authenticator, impersonator, reg_param, remove_noise_mean)
        betas = (beta1, beta2)
        self.authenticator_opt = FusedAdam(self.authenticator.parameters(), lr=au_lr, betas=betas)
        groups = [{'params': getattr(self.impersonator, name).parameters(), 'lr': env_noise_mapping_lr if name == "env_noise_mapper" else im_lr}
                  for name in self.G_GROUPS]
        self.impersonator_opt = FusedAdam(groups, lr=im_lr, betas=betas)
        self.au_scheduler = self.get_lr_scheduler(optimizer=self.authenticator_opt, milestones=lr_milestones, gamma=lr_gamma)
        self.im_scheduler = self.get_lr_scheduler(optimizer=self.impersonator_opt, milestones=lr_milestones, gamma=lr_gamma)
        self._register_checkpointables()

    # ---- networks ----
    def _authenticator_outputs(self, fake_sample, real_sample, si_sample, second_order):
        """Reference :96-142.  Per encoder the call order is the reference's (si, real, fake -- it fixes the spectral-norm power-iteration
        sequence); the src and env encoders are independent networks, so they run as two stream branches.  The R1 penalty differentiates
        the real/si branch twice: that branch uses the elementary (twice differentiable) operators, the fake branch the fused blocks."""
        au = self.authenticator

        def branch(encode):
            with (ops.composite_mode() if second_order else contextlib.nullcontext()):
                e_si, e_real = encode(si_sample), encode(real_sample)
            return e_si, e_real, encode(fake_sample)

        (si_src, real_src, fake_src), (si_env, real_env, fake_env) = ops.two_streams(lambda: branch(au.src_encode_sample),
                                                                                    lambda: branch(au.env_encode_sample))
        out_on_real = au.dis(test_src=real_src, test_env=real_env, si_src=si_src, si_env=si_env)
        out_on_fake = au.dis(test_src=fake_src, test_env=fake_env, si_src=si_src, si_env=si_env)
        return out_on_real, out_on_fake

    def _judge_attack(self, fake_sample, si_sample):
        # The reference lets the G-step's backward fill the authenticator's .grad and then discards it (authenticator_opt.zero_grad() at
        # training/gim_img_training.py:172 precedes its only use): D's weights are constants here, so neither its weight gradients nor
        # the whole backward of the si branch are computed.
        with frozen(self.authenticator):
            return self.authenticator(test_sample=fake_sample, si_sample=si_sample)

    # ---- schedule / bookkeeping ----
    def get_lr_scheduler(self, optimizer, milestones, gamma):
        return optim.lr_scheduler.MultiStepLR(optimizer=optimizer, milestones=milestones, gamma=gamma, last_epoch=self.global_step)

    def update_learning_rate(self):
        for sched in (self.au_scheduler, self.im_scheduler):
            if sched is not None:
                sched.step()

    @property
    def au_lr(self):
        return self.au_scheduler.get_last_lr()[0]

    @property
    def im_lr(self):
        return self.im_scheduler.get_last_lr()[0]

    @property
    def im_noise_mapping_lr(self):
        return self.im_scheduler.get_last_lr()[-1]

    def save(self, epoch):
        self._save(last_epoch=epoch)
