"""The image-GIM training loop -- mirrors the reference's training/gim_img_training.py (`eval_step` :96-154, `train_epoch`
:186-357, `train_gim_imgs` :360-445) on the B200 path.

Kept: function names, arguments, the per-iteration order (global step, LR schedule, G-step or G-eval, D-step), the logged
categories / keys, `save_every` / `eval_every` / `tb_log_every` / `tb_log_enc_every` cadence, checkpoint files.
Changed for a GPU that runs > 2000 episodes/s:
  * batches come from a device-resident `ResidentGIMDataSet.iter_batches` when the dataset offers it (else a DataLoader);
  * nothing is read back per iteration: the log buffers hold device scalars and are reduced on the device every `tb_log_every`
    iterations (the reference calls `.item()` a dozen times per iteration in its Gaussian loop and per log interval here);
  * with `use_cuda_graph=True` (and n_au_steps == 1) the whole G+D iteration is one CUDA-graph replay (`cuda_graph.GraphedIteration`);
  * several GPUs = one process per GPU under torchrun (`ddp.attach`), not nn.DataParallel: `device_ids` lists the local device only;
  * the image grids of `sample_and_save_imgs` (:34-73) go to `logger.add_imgs` (kept signature); `ScalarLog` tiles them into one
    grid and, given an `img_dir`, writes `<img_dir>/<category>/<k>/<step>.png` like the reference's Logger -- with a 30-line PNG
    encoder instead of torchvision / PIL, from one device-to-host copy per grid.
`logger` is anything with `add_scalar(category=, k=, v=, global_step=)` (+ `add_imgs` for the grids); `ScalarLog` below keeps the
values in memory.
"""
import itertools
import os

import torch
import torch.distributed as dist

from . import ddp
from . import model_blocks as mb
from .gim_img_trainer import GIMImgTrainer
from .training_steps import au_eval_step, au_train_step, finish_deferred_steps, im_eval_step, im_train_step
from .utils import DataParallelMock, get_device


def make_grid(imgs, nrow=5, padding=2):
    """[n, C, H, W] in [0, 1] -> one [C, rows*(H+pad)+pad, cols*(W+pad)+pad] tile image (torchvision.utils.make_grid's layout, which
    the reference's Logger.add_imgs uses, training/logger.py:43-52)."""
    imgs = imgs.detach().float().cpu()
    n, c, h, w = imgs.shape
    cols = min(nrow, n)
    rows = (n + cols - 1) // cols
    grid = torch.zeros((c, rows * (h + padding) + padding, cols * (w + padding) + padding))
    for i in range(n):
        r, q = divmod(i, cols)
        y0, x0 = padding + r * (h + padding), padding + q * (w + padding)
        grid[:, y0:y0 + h, x0:x0 + w] = imgs[i]
    return grid


def write_png(path, img):
    """Minimal PNG encoder (8-bit gray or RGB, no dependencies): img [C, H, W] float in [0, 1], C in (1, 3)."""
    import struct
    import zlib
    c, h, w = img.shape
    if c not in (1, 3):
        raise ValueError("write_png: 1 or 3 channels")
    data = (img.clamp(0, 1) * 255.0 + 0.5).to(torch.uint8).permute(1, 2, 0).contiguous().numpy()
    raw = b"".join(b"\x00" + data[y].tobytes() for y in range(h))

    def chunk(tag, payload):
        body = tag + payload
        return struct.pack(">I", len(payload)) + body + struct.pack(">I", zlib.crc32(body) & 0xffffffff)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0 if c == 1 else 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


class ScalarLog:
    """Minimal stand-in for the reference's Logger (training/logger.py): records add_scalar / add_imgs calls; with `img_dir` the image
    grids are also written as PNG files in the reference's directory layout."""

    def __init__(self, img_dir=None):
        self.scalars = {}
        self.images = {}
        self.img_dir = img_dir

    def add_scalar(self, category, k, v, global_step):
        self.scalars.setdefault((category, k), []).append((int(global_step), float(v)))

    def add_imgs(self, imgs, category, k, global_step, nrow=5):
        grid = make_grid(imgs, nrow=nrow)
        self.images[(category, k)] = (int(global_step), grid)
        if self.img_dir is not None:
            outdir = os.path.join(self.img_dir, category, k)
            os.makedirs(outdir, exist_ok=True)
            write_png(os.path.join(outdir, '%08d.png' % global_step), grid)


def save_imgs(logger, img_sample, category, k, global_step):
    """Reference :21-31: the first episode's images, [-1, 1] -> [0, 1], as one grid."""
    imgs_for_save = (img_sample[0].clamp(-1, 1) + 1) / 2.0
    logger.add_imgs(imgs=imgs_for_save, category=category, k=k, global_step=global_step)


def sample_and_save_imgs(device, logger, trainer, ds, ds_prefix, indices, dbg=False):
    """Reference :34-73: for the listed episodes, the leaked sample and what the attacker makes of it (plus real / si when dbg)."""
    if ds is None or not hasattr(logger, "add_imgs"):
        return
    with torch.no_grad():
        global_step = trainer.module.get_global_step()
        for idx in indices:
            data = ds[idx]
            category = "{} imgs_{:04}".format(ds_prefix, idx)
            leaked_sample = data["leaked_sample"].unsqueeze(0).to(device)
            fake_sample = trainer.forward(mode='impersonator_sample', leaked_sample=leaked_sample)
            save_imgs(logger=logger, img_sample=leaked_sample, category=category, k="leaked", global_step=global_step)
            save_imgs(logger=logger, img_sample=fake_sample, category=category, k="impersonator", global_step=global_step)
            if dbg:
                save_imgs(logger=logger, img_sample=data["real_sample"].unsqueeze(0).to(device), category=category, k="real", global_step=global_step)
                save_imgs(logger=logger, img_sample=data["si_sample"].unsqueeze(0).to(device), category=category, k="si", global_step=global_step)


def _batches(ds, batch_size, shuffle, num_workers, drop_last=True):
    if hasattr(ds, "iter_batches"):
        n = len(ds) // batch_size if drop_last else (len(ds) + batch_size - 1) // batch_size
        return ds.iter_batches(batch_size, shuffle=shuffle, drop_last=drop_last), n
    from torch.utils.data import DataLoader
    loader = DataLoader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, drop_last=drop_last)
    return loader, len(loader)


def _mean(buf):
    return torch.stack([t.reshape(()) for t in buf]).mean().item()


def eval_step(device, trainer, ds, logger, batch_size):
    """Reference :96-154: G-eval + D-eval over the validation split, nine scalars logged."""
    stats = {k: [] for k in ("au_loss", "au_loss_on_real", "au_loss_on_fake", "au_out_on_real", "au_out_on_fake", "au_acc", "au_acc_on_real",
                             "au_acc_on_fake", "im_loss")}
    global_step = trainer.module.get_global_step()
    batches, n_iters = _batches(ds, batch_size, False, 0)
    for data_batch in itertools.islice(batches, n_iters):
        real_sample, leaked_sample, si_sample = (data_batch[k].to(device) for k in ("real_sample", "leaked_sample", "si_sample"))
        im_loss, fake_sample, _ = im_eval_step(trainer=trainer, leaked_sample=leaked_sample, si_sample=si_sample)
        (au_loss, au_loss_on_real, au_loss_on_fake, _reg, au_out_on_real, au_out_on_fake, au_pred_on_real, au_pred_on_fake, fake_sample) = au_eval_step(
            trainer=trainer, real_sample=real_sample, fake_sample=fake_sample, si_sample=si_sample)
        acc_real = au_pred_on_real.to(torch.float).mean()
        acc_fake = torch.eq(au_pred_on_fake, 0).to(torch.float).mean()
        for k, v in (("au_loss", au_loss), ("au_loss_on_real", au_loss_on_real), ("au_loss_on_fake", au_loss_on_fake), ("au_out_on_real", au_out_on_real),
                     ("au_out_on_fake", au_out_on_fake), ("au_acc", 0.5 * (acc_real + acc_fake)), ("au_acc_on_real", acc_real), ("au_acc_on_fake", acc_fake),
                     ("im_loss", im_loss)):
            stats[k].append(v)
    if not stats["au_loss"]:
        return
    for cat, key, name in (("eval losses", "dis loss", "au_loss"), ("eval losses", "dis loss on real", "au_loss_on_real"),
                           ("eval losses", "dis loss on fake", "au_loss_on_fake"), ("eval au out", "au out on real", "au_out_on_real"),
                           ("eval au out", "au out on fake", "au_out_on_fake"), ("eval accuracy", "dis acc", "au_acc"),
                           ("eval accuracy", "dis acc on real", "au_acc_on_real"), ("eval accuracy", "dis acc on fake", "au_acc_on_fake"),
                           ("eval losses", "gen loss", "im_loss")):
        logger.add_scalar(category=cat, k=key, v=_mean(stats[name]), global_step=global_step)


def _log_encodings(trainer, logger, real_sample, si_sample, fake_sample, global_step):
    """Reference :300-340: feature statistics of D's two encoders on the current batch."""
    au = trainer.module.authenticator
    with torch.no_grad():
        enc = {}
        for which, encode in (("src", au.src_encode_sample), ("env", au.env_encode_sample)):
            enc[which] = {"real": encode(real_sample), "si": encode(si_sample), "fake": encode(fake_sample)}
        for which in ("src", "env"):
            e = enc[which]
            logger.add_scalar(category='train-au_%s_mean' % which, k='abs[real-si]', v=torch.abs(e["real"].mean(1) - e["si"].mean(1)).mean().item(),
                              global_step=global_step)
            logger.add_scalar(category='train-au_%s_mean' % which, k='abs[fake-si]', v=torch.abs(e["fake"].mean(1) - e["si"].mean(1)).mean().item(),
                              global_step=global_step)
            for name in ("real", "si", "fake"):
                logger.add_scalar(category='train-au_%s_std' % which, k=name, v=mb.custom_std(e[name]).mean().item(), global_step=global_step)


def train_epoch(device, logger, epoch, trainer, train_ds, val_ds, train_batch_size, val_batch_size, num_workers, save_every, eval_every, save_imgs_every,
                train_eval_indices, val_eval_indices, tb_log_every=100, tb_log_enc_every=500, n_au_steps=1, dbg=False, use_cuda_graph=False):
    """Reference :186-357.  One pass over `train_ds`."""
    names = ("au_loss", "au_loss_on_real", "au_loss_on_fake", "au_reg", "au_out_on_real", "au_out_on_fake", "im_loss")
    buf = {k: [] for k in names}
    pred_real, pred_fake = [], []
    batches, n_batches = _batches(train_ds, train_batch_size, True, num_workers)
    num_iters = min(50, n_batches) if dbg else n_batches
    graphed = None
    for data_batch in itertools.islice(batches, num_iters):
        real_sample, leaked_sample, si_sample = (data_batch[k].to(device) for k in ("real_sample", "leaked_sample", "si_sample"))
        if use_cuda_graph and n_au_steps == 1 and trainer.module.reg_param == 0:
            # whole iteration as one graph replay (host bookkeeping -- global step, LR schedule -- happens inside GraphedIteration)
            if graphed is None:
                from .cuda_graph import GraphedIteration
                graphed = getattr(trainer, "_graphed_iteration", None) or GraphedIteration(trainer, leaked_sample, real_sample, si_sample)
                trainer._graphed_iteration = graphed
            out = graphed(leaked_sample, real_sample, si_sample)
            global_step = trainer.module.global_step
            im_loss, au_loss, au_loss_on_real, au_loss_on_fake, au_reg, au_out_on_real, au_out_on_fake = (t.clone() for t in out)
            au_pred_on_real = au_pred_on_fake = None
            fake_sample = None
        else:
            trainer.module.do_global_step()
            trainer.module.update_learning_rate()
            global_step = trainer.module.global_step
            if (global_step + 1) % n_au_steps == 0:
                im_loss, fake_sample, _ = im_train_step(trainer=trainer, leaked_sample=leaked_sample, si_sample=si_sample)
            else:
                im_loss, fake_sample, _ = im_eval_step(trainer=trainer, leaked_sample=leaked_sample, si_sample=si_sample)
            (au_loss, au_loss_on_real, au_loss_on_fake, au_reg, au_out_on_real, au_out_on_fake, au_pred_on_real, au_pred_on_fake, fake_sample) = au_train_step(
                trainer=trainer, real_sample=real_sample, fake_sample=fake_sample, si_sample=si_sample)
            finish_deferred_steps(trainer)
        for k, v in zip(names, (au_loss, au_loss_on_real, au_loss_on_fake, au_reg, au_out_on_real, au_out_on_fake, im_loss)):
            buf[k].append(v)
        if au_pred_on_real is not None:
            pred_real.append(au_pred_on_real.view(-1))
            pred_fake.append(au_pred_on_fake.view(-1))

        if global_step % tb_log_every == 0:
            m = trainer.module
            logger.add_scalar(category='lr', k='au', v=m.au_lr, global_step=global_step)
            logger.add_scalar(category='lr', k='im', v=m.im_lr, global_step=global_step)
            logger.add_scalar(category='lr', k='im_lm', v=m.im_noise_mapping_lr, global_step=global_step)
            for cat, key, name in (('train_losses', 'dis_loss', "au_loss"), ('train_losses', 'dis_loss_on_real', "au_loss_on_real"),
                                   ('train_losses', 'dis_loss_on_fake', "au_loss_on_fake"), ('train_losses', 'dis_reg', "au_reg"),
                                   ('train_au_out', 'au_out_on_real', "au_out_on_real"), ('train_au_out', 'au_out_on_fake', "au_out_on_fake"),
                                   ('train losses', 'gen loss', "im_loss")):
                logger.add_scalar(category=cat, k=key, v=_mean(buf[name]), global_step=global_step)
            if pred_real:
                acc_real = torch.cat(pred_real).to(torch.float).mean()
                acc_fake = torch.eq(torch.cat(pred_fake), 0).to(torch.float).mean()
                logger.add_scalar(category='train_accuracy', k='dis_acc', v=(0.5 * (acc_real + acc_fake)).item(), global_step=global_step)
                logger.add_scalar(category='train_accuracy', k='dis_acc_on_real', v=acc_real.item(), global_step=global_step)
                logger.add_scalar(category='train_accuracy', k='dis_acc_on_fake', v=acc_fake.item(), global_step=global_step)
            buf = {k: [] for k in names}
            pred_real, pred_fake = [], []

        if fake_sample is not None and global_step % tb_log_enc_every == 0:
            _log_encodings(trainer, logger, real_sample, si_sample, fake_sample, global_step)
        if global_step % save_every == 0:
            _save_rank0(trainer, epoch)
        if global_step % save_imgs_every == 0:
            sample_and_save_imgs(device=device, logger=logger, trainer=trainer, ds=train_ds, ds_prefix='train', indices=train_eval_indices, dbg=dbg)
            sample_and_save_imgs(device=device, logger=logger, trainer=trainer, ds=val_ds, ds_prefix='val', indices=val_eval_indices, dbg=dbg)
        if global_step % eval_every == 0 and val_ds is not None:
            eval_step(device=device, trainer=trainer, ds=val_ds, logger=logger, batch_size=val_batch_size)


def train_gim_imgs(device_name, device_ids, outdir, train_ds, val_ds, authenticator, impersonator, m, n, k, reg_param, remove_noise_mean, au_lr, im_lr,
                   beta1, beta2, env_noise_mapping_lr, lr_gamma, milestones, resume_from_ckpt, n_epochs, batch_size, num_workers, save_every, eval_every,
                   save_imgs_every, train_eval_indices, val_eval_indices, n_au_steps=1, dbg=False, logger=None, use_cuda_graph=False):
    """Reference :360-445.  `batch_size` is the per-process batch; under torchrun each rank trains on its own episodes and the
    gradients are averaged (ddp.attach), which reproduces nn.DataParallel's mean over the global batch."""
    device = get_device(device_type=device_name, device_ids=device_ids)
    logger = logger if logger is not None else ScalarLog()
    authenticator, impersonator = authenticator.to(device), impersonator.to(device)
    trainer = GIMImgTrainer(outdir=outdir, m=m, n=n, k=k, authenticator=authenticator, impersonator=impersonator, au_lr=au_lr, im_lr=im_lr,
                            env_noise_mapping_lr=env_noise_mapping_lr, beta1=beta1, beta2=beta2, lr_milestones=milestones, lr_gamma=lr_gamma,
                            reg_param=reg_param, remove_noise_mean=remove_noise_mean).to(device)
    if resume_from_ckpt:
        trainer.resume_from_ckpt(ckpt_path=resume_from_ckpt)
        trainer.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        ddp.broadcast_module_state(trainer)            # replicas start from rank 0's parameters and spectral-norm u/v
        ddp.attach(trainer.authenticator_opt)
        ddp.attach(trainer.impersonator_opt, defer=True)      # G's all-reduce overlaps the D-step (flushed at the end of the iteration)
    trainer = DataParallelMock(trainer)
    os.makedirs(outdir, exist_ok=True)
    for ep in range(n_epochs):
        try:
            train_epoch(device=device, logger=logger, epoch=ep, trainer=trainer, train_ds=train_ds, val_ds=val_ds, train_batch_size=batch_size,
                        val_batch_size=batch_size, num_workers=num_workers, save_every=save_every, eval_every=eval_every, save_imgs_every=save_imgs_every,
                        train_eval_indices=train_eval_indices, val_eval_indices=val_eval_indices, n_au_steps=n_au_steps, dbg=dbg,
                        use_cuda_graph=use_cuda_graph)
        except KeyboardInterrupt:
            _save_rank0(trainer, ep)
            raise
    _save_rank0(trainer, n_epochs - 1)
    return trainer, logger


def _save_rank0(trainer, epoch):
    """Checkpoints are written by rank 0 only (replicas are identical); the other ranks wait until the file is complete."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if not multi or dist.get_rank() == 0:
        if epoch is None:
            trainer.module.save()                      # GIMGaussianTrainer.save() takes no epoch (reference gim_gaussian_trainer.py:125)
        else:
            trainer.module.save(epoch=epoch)
    if multi:
        dist.barrier()
