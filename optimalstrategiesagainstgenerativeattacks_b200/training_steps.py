"""The per-iteration step functions of the reference's training loops (training/gim_img_training.py:76-95, 157-183;
training/gim_gaussian_training.py:21-47), callable with either trainer."""
import torch

from . import ops


def im_train_step(trainer, leaked_sample, si_sample):
    trainer.module.impersonator.train()
    trainer.module.impersonator_opt.zero_grad()
    loss, fake_sample, au_out = trainer.forward(mode='impersonator_forward', leaked_sample=leaked_sample, si_sample=si_sample)
    loss = loss.mean()
    with ops.deferred_weight_grads():          # one batched spectral-norm backward for all convolutions at the end of the pass
        loss.backward()
    trainer.module.impersonator_opt.step()
    return loss.detach(), fake_sample.detach(), au_out.detach()


def au_train_step(trainer, real_sample, fake_sample, si_sample):
    trainer.module.authenticator.train()
    trainer.module.authenticator_opt.zero_grad()
    (loss, loss_on_real, loss_on_fake, reg, out_on_real, out_on_fake, pred_on_real, pred_on_fake, fake_sample) = trainer.forward(
        mode='authenticator_forward', fake_sample=fake_sample, real_sample=real_sample, si_sample=si_sample)
    loss = loss.mean()
    with ops.deferred_weight_grads():
        loss.backward()
    trainer.module.authenticator_opt.step()
    return (loss.detach(), loss_on_real.detach().mean(), loss_on_fake.detach().mean(), reg.detach().mean(),
            out_on_real.detach().mean(), out_on_fake.detach().mean(),
            pred_on_real.detach(), pred_on_fake.detach(), fake_sample.detach())


def finish_deferred_steps(trainer):
    """Data parallel with overlapped communication (ddp.attach(..., defer=True)): apply the optimizer steps whose gradient all-reduce was
    launched asynchronously.  Called once at the end of an iteration; a no-op otherwise."""
    for opt in (trainer.module.impersonator_opt, trainer.module.authenticator_opt):
        flush = getattr(opt, "flush", None)
        if flush is not None:
            flush()


def im_eval_step(trainer, leaked_sample, si_sample):
    trainer.module.impersonator.eval()
    with torch.no_grad():
        loss, fake_sample, au_out = trainer.forward(mode='impersonator_forward', leaked_sample=leaked_sample, si_sample=si_sample)
        loss = loss.mean()
    return loss.detach(), fake_sample.detach(), au_out.detach()


def au_eval_step(trainer, real_sample, fake_sample, si_sample):
    trainer.module.authenticator.eval()
    with torch.no_grad():
        (loss, loss_on_real, loss_on_fake, reg, out_on_real, out_on_fake, pred_on_real, pred_on_fake, fake_sample) = trainer.forward(
            mode='authenticator_forward', fake_sample=fake_sample, real_sample=real_sample, si_sample=si_sample, grad=False)
        loss = loss.mean()
    return (loss.detach(), loss_on_real.detach().mean(), loss_on_fake.detach().mean(), reg.detach().mean(),
            out_on_real.detach().mean(), out_on_fake.detach().mean(),
            pred_on_real.detach(), pred_on_fake.detach(), fake_sample.detach())
