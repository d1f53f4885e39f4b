"""GIMImgTrainer -- same constructor, modes, return tuples, optimizer layout and checkpoint registration as the reference's
training/gim_img_trainer.py; losses and optimizers run on libgim_b200 kernels."""
import contextlib
import os

import torch
import torch.nn as nn
import torch.optim as optim

from . import ops
from .checkpoints import CheckpointIO
from .fused_adam import FusedAdam
from .utils import GlobalStep, compute_grad2, num_parameters


@contextlib.contextmanager
def _frozen(module):
    """Inside, the module's parameters do not require grad (graphs recorded inside treat them as constants)."""
    params = [p for p in module.parameters() if p.requires_grad]
    for p in params:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p in params:
            p.requires_grad_(True)


class GIMImgTrainer(nn.Module):
    CHECKPOINT_DIR = "ckpts"

    def __init__(self, outdir, m, n, k, authenticator, impersonator, au_lr, im_lr, env_noise_mapping_lr,
                 beta1=0., beta2=0.99, lr_milestones=(), lr_gamma=0.3, reg_param=10., remove_noise_mean=True):
        super().__init__()
        self.m = m
        self.n = n
        self.k = k
        self.authenticator = authenticator
        self.impersonator = impersonator
        self._global_step = GlobalStep()
        self.reg_param = reg_param
        self.remove_noise_mean = remove_noise_mean

        # reference :50-58 -- one Adam for D, one Adam with six groups for G (the noise mapper has its own lr)
        self.authenticator_opt = FusedAdam(self.authenticator.parameters(), lr=au_lr, betas=(beta1, beta2))
        self.impersonator_opt = FusedAdam([
            {'params': self.impersonator.src_encoder.parameters(), 'lr': im_lr},
            {'params': self.impersonator.env_encoder.parameters(), 'lr': im_lr},
            {'params': self.impersonator.env_decoder.parameters(), 'lr': im_lr},
            {'params': self.impersonator.img2img.parameters(), 'lr': im_lr},
            {'params': self.impersonator.img_att.parameters(), 'lr': im_lr},
            {'params': self.impersonator.env_noise_mapper.parameters(), 'lr': env_noise_mapping_lr}
        ], lr=im_lr, betas=(beta1, beta2))
        self.au_scheduler = self.get_lr_scheduler(optimizer=self.authenticator_opt, milestones=lr_milestones, gamma=lr_gamma)
        self.im_scheduler = self.get_lr_scheduler(optimizer=self.impersonator_opt, milestones=lr_milestones, gamma=lr_gamma)

        print("Authenticator has {} parameters".format(num_parameters(self.authenticator.parameters())))
        print("impersonator has {} parameters".format(num_parameters(self.impersonator.parameters())))

        self.checkpoint_dir = os.path.join(outdir, self.CHECKPOINT_DIR)
        self.checkpoint_io = CheckpointIO(checkpoint_dir=self.checkpoint_dir)
        self.checkpoint_io.register_modules(
            authenticator=self.authenticator, impersonator=self.impersonator,
            authenticator_opt=self.authenticator_opt, impersonator_opt=self.impersonator_opt,
            global_step=self._global_step)

    def forward(self, mode, **kwargs):
        if mode == "authenticator_forward":
            return self.authenticator_forward(**kwargs)
        elif mode == "impersonator_forward":
            return self.impersonator_forward(**kwargs)
        elif mode == "impersonator_sample":
            return self.impersonator_sample(**kwargs)
        raise ValueError("unsupported mode")

    def gan_loss(self, dis_out, target, reduce=False):
        """Per-episode BCE-with-logits (reference :90-94)."""
        loss = ops.BCEWithLogitsFn.apply(dis_out, float(target))
        return loss.mean() if reduce else loss.squeeze()

    def authenticator_forward(self, fake_sample, real_sample, si_sample, grad=True):
        """Reference :96-142 (encode order si, real, fake; R1 on the real branch)."""
        if self.reg_param > 0:
            real_sample.requires_grad_()
            si_sample.requires_grad_()
        au = self.authenticator
        # the R1 penalty differentiates the real/si branch twice: run it with the elementary (twice differentiable) operators
        # (per encoder the call order is the reference's: si, real, fake; the src and env encoders are independent -> two streams)
        def branch(encode):
            with (ops.composite_mode() if (grad and self.reg_param > 0) else contextlib.nullcontext()):
                e_si = encode(si_sample)
                e_real = encode(real_sample)
            return e_si, e_real, encode(fake_sample)
        (au_si_src, au_real_src, au_fake_src), (au_si_env, au_real_env, au_fake_env) = ops.two_streams(
            lambda: branch(au.src_encode_sample), lambda: branch(au.env_encode_sample))

        out_on_real = au.dis(test_src=au_real_src, test_env=au_real_env, si_src=au_si_src, si_env=au_si_env)
        loss_on_real = self.gan_loss(dis_out=out_on_real, target=1.)
        if grad and self.reg_param > 0:
            reg = self.reg_param * compute_grad2(out_on_real, (real_sample, si_sample))
        else:
            reg = torch.zeros_like(loss_on_real)

        out_on_fake = au.dis(test_src=au_fake_src, test_env=au_fake_env, si_src=au_si_src, si_env=au_si_env)
        loss_on_fake = self.gan_loss(dis_out=out_on_fake, target=0.)

        with torch.no_grad():
            pred_on_real = torch.ge(out_on_real.detach(), 0)
            pred_on_fake = torch.ge(out_on_fake.detach(), 0)

        loss = loss_on_real + loss_on_fake + reg
        return (loss, loss_on_real.detach(), loss_on_fake.detach(), reg, out_on_real.detach(), out_on_fake.detach(),
                pred_on_real.detach(), pred_on_fake.detach(), fake_sample.detach())

    def impersonator_forward(self, leaked_sample, si_sample):
        fake_sample = self.impersonator(leaked_sample=leaked_sample, n=self.n, remove_noise_mean=self.remove_noise_mean)
        # The reference lets this backward fill the authenticator's .grad and then discards it (authenticator_opt.zero_grad() at
        # training/gim_img_training.py:172 precedes its only use): treat D's weights as constants here, so neither its weight
        # gradients nor the whole backward of the si branch are computed.
        with _frozen(self.authenticator):
            auth_out = self.authenticator(test_sample=fake_sample, si_sample=si_sample)
        loss = self.gan_loss(dis_out=auth_out, target=1.)
        return loss, fake_sample, auth_out

    def impersonator_sample(self, leaked_sample):
        with torch.no_grad():
            return self.impersonator(leaked_sample=leaked_sample, n=self.n, remove_noise_mean=self.remove_noise_mean)

    def resume_from_ckpt(self, ckpt_path):
        self.checkpoint_io.load(ckpt_path)
        print('Resuming training from iteration {}'.format(self.get_global_step()))

    def save(self, epoch):
        print("\nSaving checkpoint...\n")
        self.checkpoint_io.save(global_step=self.get_global_step(), last_epoch=epoch,
                                filename="model_{:08}.pt".format(self.get_global_step()))

    def get_lr_scheduler(self, optimizer, milestones, gamma):
        return optim.lr_scheduler.MultiStepLR(optimizer=optimizer, milestones=milestones, gamma=gamma, last_epoch=self.global_step)

    def update_learning_rate(self):
        if self.au_scheduler is not None:
            self.au_scheduler.step()
        if self.im_scheduler is not None:
            self.im_scheduler.step()

    def get_global_step(self):
        return self._global_step.get()

    def do_global_step(self):
        return self._global_step.step()

    @property
    def au_lr(self):
        return self.au_scheduler.get_last_lr()[0]

    @property
    def im_lr(self):
        return self.im_scheduler.get_last_lr()[0]

    @property
    def im_noise_mapping_lr(self):
        return self.im_scheduler.get_last_lr()[-1]

    @property
    def global_step(self):
        return self.get_global_step()
