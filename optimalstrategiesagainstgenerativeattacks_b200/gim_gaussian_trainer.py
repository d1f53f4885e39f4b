"""GIMGaussianTrainer -- the reference's training/gim_gaussian_trainer.py:21-139 surface (plain Adam(lr) with default betas, no LR
schedule, `save()` without an epoch) on the shared trainer base; losses and optimizers run on libgim_b200 kernels."""
from .fused_adam import FusedAdam
from .trainer_base import GIMTrainerBase


class GIMGaussianTrainer(GIMTrainerBase):
    def __init__(self, outdir, m, n, k, authenticator, impersonator, au_lr, im_lr, reg_param=0., remove_noise_mean=True):
        super().__init__(outdir, m, n, k, authenticator, impersonator, reg_param, remove_noise_mean)
        self.authenticator_opt = FusedAdam(self.authenticator.parameters(), lr=au_lr)
        self.impersonator_opt = FusedAdam(self.impersonator.parameters(), lr=im_lr)
        self._register_checkpointables()

    def _authenticator_outputs(self, fake_sample, real_sample, si_sample, second_order):
        # reference :84-110: D(real, si) first, then D(fake, si); the Gaussian D has no block-level fused operators to bypass
        out_on_real = self.authenticator(test_sample=real_sample, si_sample=si_sample)
        out_on_fake = self.authenticator(test_sample=fake_sample, si_sample=si_sample)
        return out_on_real, out_on_fake

    def save(self):
        self._save(last_epoch=1)
