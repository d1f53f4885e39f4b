"""The Gaussian-GIM training loop -- mirrors the reference's training/gim_gaussian_training.py (`im_train_step` :21-30,
`au_train_step` :33-47, `train` :50-151, `train_gim_gaussian` :154-232) on the B200 path.

Kept: function names and arguments, the per-iteration order (global step, episode synthesis, G-step, D-step), every logged
category / key with its global step, the `save_stats_every` distance statistics, the `save_every` checkpoint cadence, the
KeyboardInterrupt / PermissionError checkpoint.
Changed for a GPU that runs > 10^6 episodes/s:
  * episodes are synthesised ON THE DEVICE (`ops.gaussian_episodes`: mu ~ N(0, prior_sigma^2), x | mu ~ N(mu, src_sigma^2), the
    reference's distributions; the reference draws 65 M normals per iteration on the host at d = 1000 and copies them);
  * the whole iteration -- synthesis, G-step, D-step, both Adam updates -- is ONE CUDA-graph replay (`use_cuda_graph=True`);
  * the ten scalars the reference reads back with `.item()` every iteration are kept in a device-side ring and delivered to the
    logger every `log_every` iterations with ONE device-to-host copy -- same keys, same global steps, same values;
  * several GPUs = one process per GPU under torchrun (`ddp.attach`), not nn.DataParallel.
`logger` is anything with `add_scalar(category=, k=, v=, global_step=)`; `ScalarLog` keeps the values in memory.
"""
import os

import torch
import torch.distributed as dist

from . import ddp, ops
from . import model_blocks as mb
from .gim_gaussian_trainer import GIMGaussianTrainer
from .gim_img_training import ScalarLog, _save_rank0
from .training_steps import au_train_step, finish_deferred_steps, im_train_step   # noqa: F401  (the step functions are re-exported under the reference's names)
from .utils import DataParallelMock, get_device

# (category, key) of the per-iteration scalars, in the order the reference logs them (:93-112)
SCALARS = (('train losses', 'im loss'), ('train losses', 'au loss'), ('train losses', 'au loss on real'), ('train losses', 'au loss on fake'),
           ('train losses', 'au reg'), ('train au out', 'au out on real'), ('train au out', 'au out on fake'),
           ('train accuracy', 'au acc'), ('train accuracy', 'au acc on real'), ('train accuracy', 'au acc on fake'))


class _DeferredScalars:
    """Device-side ring of the per-iteration scalars; `flush` hands them to the logger with one copy."""

    def __init__(self, logger, capacity, device):
        self.logger = logger
        self.buf = torch.zeros((capacity, len(SCALARS)), dtype=torch.float32, device=device)
        self.steps = []

    def put(self, global_step, values):
        self.buf[len(self.steps)] = values
        self.steps.append(global_step)
        if len(self.steps) == self.buf.shape[0]:
            self.flush()

    def flush(self):
        if not self.steps:
            return
        host = self.buf[:len(self.steps)].cpu()
        for row, gs in zip(host.tolist(), self.steps):
            for (cat, key), v in zip(SCALARS, row):
                self.logger.add_scalar(category=cat, k=key, v=v, global_step=gs)
        self.steps = []


def _iteration_scalars(im_loss, au_out):
    """The ten per-iteration scalars as one device vector (reference :88-112)."""
    au_loss, au_loss_on_real, au_loss_on_fake, au_reg, au_out_on_real, au_out_on_fake, au_pred_on_real, au_pred_on_fake = au_out[:8]
    acc_real = au_pred_on_real.to(torch.float).mean()
    acc_fake = torch.eq(au_pred_on_fake, 0).to(torch.float).mean()
    return torch.stack([t.reshape(()).float() for t in (im_loss, au_loss, au_loss_on_real, au_loss_on_fake, au_reg, au_out_on_real, au_out_on_fake,
                                                         0.5 * (acc_real + acc_fake), acc_real, acc_fake)])


class _GraphedGaussianIteration:
    """Episode synthesis + G-step + D-step (+ the scalar vector) captured as one CUDA graph; warm-up is side-effect free."""

    def __init__(self, trainer, sample, warmup=2):
        from .cuda_graph import _TrainingState
        self.trainer, self.sample = trainer, sample
        m = trainer.module
        state = _TrainingState(m)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        state.restore()
        for opt in (m.authenticator_opt, m.impersonator_opt):
            opt.sync_lrs()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._body()

    def _body(self):
        mu, (leaked, real, si) = self.sample()
        im_loss, fake, _ = im_train_step(self.trainer, leaked, si)
        au_out = au_train_step(self.trainer, real, fake, si)
        finish_deferred_steps(self.trainer)
        return _iteration_scalars(im_loss, au_out), mu, leaked, real, au_out[8]

    def __call__(self):
        self.graph.replay()
        for opt in (self.trainer.module.authenticator_opt, self.trainer.module.impersonator_opt):
            opt.note_graph_steps(1)
        return self.out


def train(device, trainer, logger, n_iters, batch_size, src_dim, src_sigma, prior_sigma, save_stats_every, save_every, use_cuda_graph=True,
          log_every=100):
    """Reference :50-151.  `batch_size` is this process's batch."""
    m_, n_, k_ = trainer.module.m, trainer.module.n, trainer.module.k

    def sample():                                            # real, leaked, si of one batch of episodes (:71-86)
        mu, (real, leaked, si) = ops.gaussian_episodes(batch_size, (n_, m_, k_), src_dim, prior_sigma, src_sigma, device)
        return mu, (leaked, real, si)

    scalars = _DeferredScalars(logger, max(1, log_every), device)
    graphed = _GraphedGaussianIteration(trainer, sample) if (use_cuda_graph and n_iters > 0) else None      # (warm-up leaves no trace)
    try:
        for _ in range(n_iters):
            trainer.module.do_global_step()
            global_step = trainer.module.get_global_step()
            if graphed is not None:
                vec, mu, leaked_sample, real_sample, fake_sample = graphed()
            else:
                mu, (leaked_sample, real_sample, si_sample) = sample()
                im_loss, fake_sample, _ = im_train_step(trainer=trainer, leaked_sample=leaked_sample, si_sample=si_sample)
                au_out = au_train_step(trainer=trainer, real_sample=real_sample, fake_sample=fake_sample, si_sample=si_sample)
                finish_deferred_steps(trainer)
                vec, fake_sample = _iteration_scalars(im_loss, au_out), au_out[8]
            scalars.put(global_step, vec)

            if global_step % save_stats_every == 0:          # :115-147 (five host reads, at the reference's cadence)
                with torch.no_grad():
                    sigma = torch.full_like(mu, float(src_sigma))
                    l1 = lambda a, b: (a - b).abs().mean().item()
                    fake_mean, real_mean = fake_sample.mean(dim=1), real_sample.mean(dim=1)
                    logger.add_scalar(category='im distances', k='l1_dist_from_leaked_sample_mean', v=l1(fake_mean, leaked_sample.mean(dim=1)), global_step=global_step)
                    logger.add_scalar(category='im distances', k='l1_dist_from_gt_sample_mean', v=l1(fake_mean, mu), global_step=global_step)
                    logger.add_scalar(category='im distances', k='l1_dist_from_gt_std', v=l1(mb.custom_std(fake_sample), sigma), global_step=global_step)
                    logger.add_scalar(category='real distances', k='l1_dist_from_gt_sample_mean', v=l1(real_mean, mu), global_step=global_step)
                    logger.add_scalar(category='real distances', k='l1_dist_from_gt_std', v=l1(mb.custom_std(real_sample), sigma), global_step=global_step)
            if global_step % save_every == 0:
                scalars.flush()
                _save_rank0(trainer, None)
    finally:
        scalars.flush()


def train_gim_gaussian(device_name, device_ids, outdir, authenticator, impersonator, m, n, k, src_dim, src_sigma, prior_sigma, reg_param, remove_noise_mean,
                       au_lr, im_lr, resume_from_ckpt, n_iters, batch_size, save_every, save_stats_every, logger=None, use_cuda_graph=True, log_every=100):
    """Reference :154-232.  Returns (trainer, logger)."""
    device = get_device(device_type=device_name, device_ids=device_ids)
    assert batch_size % max(1, len(device_ids)) == 0
    logger = logger if logger is not None else ScalarLog()
    authenticator, impersonator = authenticator.to(device), impersonator.to(device)
    trainer = GIMGaussianTrainer(outdir=outdir, m=m, n=n, k=k, authenticator=authenticator, impersonator=impersonator, au_lr=au_lr, im_lr=im_lr,
                                 reg_param=reg_param, remove_noise_mean=remove_noise_mean).to(device)
    if resume_from_ckpt:
        trainer.resume_from_ckpt(ckpt_path=resume_from_ckpt)
        trainer.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        ddp.broadcast_module_state(trainer)
        ddp.attach(trainer.authenticator_opt)
        ddp.attach(trainer.impersonator_opt, defer=True)
    trainer = DataParallelMock(trainer)
    os.makedirs(outdir, exist_ok=True)
    try:
        train(device=device, trainer=trainer, logger=logger, n_iters=n_iters, batch_size=batch_size, src_dim=src_dim, src_sigma=src_sigma,
              prior_sigma=prior_sigma, save_stats_every=save_stats_every, save_every=save_every, use_cuda_graph=use_cuda_graph, log_every=log_every)
    except (KeyboardInterrupt, PermissionError) as e:
        print("\n%s\nSaving checkpoint...\n" % type(e).__name__)
        _save_rank0(trainer, None)
    return trainer, logger
