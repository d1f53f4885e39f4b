"""Generate (in float64) tests/golden/*.npz + schemas.json by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Every case draws weights with oracle/fill.py (numpy RandomState, key order) and inputs with `seeded`,
so a fixture stores only seeds-by-convention and the reference's OUTPUTS.  The noise z that
GIMFaceImpersonator.forward draws with torch.randn (gim_img_models.py:374) is injected by patching
torch.randn for the duration of the call.
"""
import contextlib
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim                      # noqa: E402
from oracle.fill import fill_state_dict, schema_of, seeded    # noqa: E402

ref_shim.install()
import models.gim_img_models as ref_img           # noqa: E402
import models.gim_gaussian_models as ref_gauss    # noqa: E402
from training.gim_img_trainer import GIMImgTrainer              # noqa: E402
from training.gim_gaussian_trainer import GIMGaussianTrainer    # noqa: E402
from training.utils import DataParallelMock       # noqa: E402
import training.gim_img_training as ref_img_loop  # noqa: E402
import training.gim_gaussian_training as ref_gauss_loop   # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


@contextlib.contextmanager
def inject_randn(zs):
    """Make torch.randn return the queued tensors (one per call)."""
    real = torch.randn
    queue = list(zs)

    def fake(*a, **k):
        return queue.pop(0).clone()
    torch.randn = fake
    try:
        yield
    finally:
        torch.randn = real


def load(module, seed):
    """Deterministic fp32-representable weights, then the whole module in float64: the vectors are the reference's
    algorithm evaluated (almost) exactly, so fp32 implementations can be judged against the truth."""
    sd = fill_state_dict(schema_of(module), seed)
    module.load_state_dict(sd)
    return module.double()


_seeded32 = seeded


def seeded(*a, **k):
    return _seeded32(*a, **k).double()


def grad_summary(module):
    rows = []
    for _, prm in module.named_parameters():
        if prm.grad is None:
            rows.append([np.nan, np.nan])
        else:
            g = prm.grad.double()
            rows.append([g.sum().item(), g.norm().item()])
    return np.asarray(rows, dtype=np.float64)


N_SAMPLES = 128          # gradient elements stored per parameter tensor
FULL_LIMIT = 20000       # parameters up to this size may be stored in full (see grad_full)


def sample_index(numel, j):
    """The element positions stored for the j-th parameter tensor (shared with tests/conftest.py: `grad_sample_index`)."""
    if numel <= N_SAMPLES:
        return np.arange(numel)
    return np.sort(np.random.RandomState(7919 + j).choice(numel, N_SAMPLES, replace=False))


def grad_samples(module):
    """[n_params, N_SAMPLES] float64: actual gradient ELEMENTS at seeded positions of every parameter (NaN rows: no gradient; short
    tensors are NaN-padded).  Unlike (sum, norm) summaries these see a permuted / transposed gradient."""
    rows = np.full((len(list(module.parameters())), N_SAMPLES), np.nan)
    for j, (_, prm) in enumerate(module.named_parameters()):
        if prm.grad is None:
            continue
        g = prm.grad.detach().double().reshape(-1).numpy()
        idx = sample_index(g.size, j)
        rows[j, :idx.size] = g[idx]
    return rows


def grad_full(module, prefix, count=10):
    """{prefix + name: float32 gradient} for `count` whole tensors spread over the sub-networks: the largest that fit FULL_LIMIT."""
    named = [(n, p) for n, p in module.named_parameters() if p.grad is not None and 256 <= p.numel() <= FULL_LIMIT]
    by_top = {}
    for n, p in named:
        by_top.setdefault(n.split(".")[0], []).append((n, p))
    picked, tops = [], sorted(by_top)
    for lst in by_top.values():
        lst.sort(key=lambda np_: -np_[1].numel())
    while len(picked) < count and any(by_top.values()):
        for t in tops:
            if by_top[t] and len(picked) < count:
                picked.append(by_top[t].pop(0))
    return {prefix + n: p.grad.detach().float().numpy() for n, p in picked}


def param_summary(module):
    return np.asarray([[v.double().sum().item(), v.double().norm().item()] for v in module.state_dict().values()],
                      dtype=np.float64)


def npy(t):
    return t.detach().cpu().numpy()


def case_au(size, ch, sd_dim, b, n, k, seed, with_grads=True):
    au = load(ref_img.get_au(size, ch, sd_dim), seed)
    au.train()
    test = seeded((b, n, ch, size, size), seed + 1, 0.5, 1.0).requires_grad_(with_grads)
    si = seeded((b, k, ch, size, size), seed + 2, 0.5, 1.0).requires_grad_(with_grads)
    out = au(test, si)
    res = {"out": npy(out)}
    if with_grads:
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, torch.ones_like(out), reduction="none").mean()
        loss.backward()
        res.update(grad_full(au, "gfull."))
        res.update(loss=npy(loss), grads=grad_summary(au), gsamp=grad_samples(au), g_test=npy(test.grad), g_si=npy(si.grad),
                   g_mlp_last=npy(au.dis.mlp.model[4].weight.grad),
                   u_after=npy(au.src_encoder.down_blocks[0].conv_r1.weight_u),
                   v_after=npy(au.src_encoder.down_blocks[0].conv_r1.weight_v))
    au.eval()
    with torch.no_grad():
        res["out_eval"] = npy(au(test.detach(), si.detach()))
    return res


def case_im(size, ch, sd_dim, b, m, n, seed, with_grads=True, use_img_att=False):
    im = load(ref_img.get_im(size, ch, sd_dim, use_img_att=use_img_att), seed)
    im.train()
    leaked = seeded((b, m, ch, size, size), seed + 1, 0.5, 1.0)
    z = seeded((b, n, sd_dim), seed + 3)
    with inject_randn([z]):
        fake = im(leaked, n, True)
    res = {"fake": npy(fake)}
    if with_grads:
        probe = seeded(tuple(fake.shape), seed + 4)
        (fake * probe).sum().backward()
        res["grads"] = grad_summary(im)
        res["gsamp"] = grad_samples(im)
        res.update(grad_full(im, "gfull."))
        res["g_noise_last"] = npy(im.env_noise_mapper.model[6].weight.grad)
    return res


def case_img_steps(size, ch, sd_dim, b, m, n, k, reg, iters, seed, lrs=(1e-3, 1e-3, 1e-4)):
    au = load(ref_img.get_au(size, ch, sd_dim), seed)
    im = load(ref_img.get_im(size, ch, sd_dim), seed + 10)
    tr = DataParallelMock(GIMImgTrainer("/tmp/gim_golden_out", m, n, k, au, im, lrs[0], lrs[1], lrs[2], reg_param=reg))
    rec = {key: [] for key in ("im_loss", "au_loss", "loss_real", "loss_fake", "reg", "out_real", "out_fake")}
    for it in range(iters):
        leaked = seeded((b, m, ch, size, size), seed + 100 * it + 1, 0.5, 1.0)
        real = seeded((b, n, ch, size, size), seed + 100 * it + 2, 0.5, 1.0)
        si = seeded((b, k, ch, size, size), seed + 100 * it + 3, 0.5, 1.0)
        z = seeded((b, n, sd_dim), seed + 100 * it + 4)
        tr.module.do_global_step()
        tr.module.update_learning_rate()
        with inject_randn([z]):
            im_loss, fake, _ = ref_img_loop.im_train_step(tr, leaked, si)
        o = ref_img_loop.au_train_step(tr, real, fake, si)
        rec["im_loss"].append(im_loss.item())
        for key, val in zip(("au_loss", "loss_real", "loss_fake", "reg", "out_real", "out_fake"), o[:6]):
            rec[key].append(val.item())
        if it == 0:
            rec["fake0"] = npy(fake)
    res = {key: np.asarray(v) for key, v in rec.items()}
    res["au_params"] = param_summary(au)
    res["im_params"] = param_summary(im)
    return res


def case_step_grads(size, ch, sd_dim, b, m, n, k, reg, seed, lrs=(1e-4, 1e-4, 1e-6)):
    """ONE full training iteration (G-step, D-step; R1 when reg > 0) of the reference trainer at full network width, batch b:
    losses, the generated images and the gradients both optimizers consumed (element samples of every tensor + a dozen whole tensors)."""
    au = load(ref_img.get_au(size, ch, sd_dim), seed)
    im = load(ref_img.get_im(size, ch, sd_dim), seed + 10)
    tr = DataParallelMock(GIMImgTrainer("/tmp/gim_golden_out", m, n, k, au, im, lrs[0], lrs[1], lrs[2], reg_param=reg))
    leaked = seeded((b, m, ch, size, size), seed + 1, 0.5, 1.0)
    real = seeded((b, n, ch, size, size), seed + 2, 0.5, 1.0)
    si = seeded((b, k, ch, size, size), seed + 3, 0.5, 1.0)
    z = seeded((b, n, sd_dim), seed + 4)
    tr.module.do_global_step()
    tr.module.update_learning_rate()
    with inject_randn([z]):
        im_loss, fake, au_out = ref_img_loop.im_train_step(tr, leaked, si)
    res = {"im_loss": npy(im_loss), "fake": npy(fake).astype(np.float32), "g_au_out": npy(au_out),
           "im_grads": grad_summary(im), "im_gsamp": grad_samples(im)}
    res.update(grad_full(im, "im_gfull."))
    o = ref_img_loop.au_train_step(tr, real, fake, si)
    res.update(au_loss=npy(o[0]), loss_real=npy(o[1]), loss_fake=npy(o[2]), reg=npy(o[3]), out_real=npy(o[4]), out_fake=npy(o[5]),
               au_grads=grad_summary(au), au_gsamp=grad_samples(au))
    res.update(grad_full(au, "au_gfull."))
    return res


def state_samples(module):
    """[n_entries, N_SAMPLES] float64 element samples of every state-dict tensor (same positions rule as grad_samples)."""
    vals = list(module.state_dict().values())
    rows = np.full((len(vals), N_SAMPLES), np.nan)
    for j, v in enumerate(vals):
        a = v.detach().double().reshape(-1).numpy()
        idx = sample_index(a.size, j)
        rows[j, :idx.size] = a[idx]
    return rows


def case_gauss(d, b, m, n, k, iters, seed, reg=0.0, compact=False):
    au = load(ref_gauss.get_au(d), seed)
    im = load(ref_gauss.get_im(d), seed + 10)
    real = seeded((b, n, d), seed + 1)
    si = seeded((b, k, d), seed + 2)
    leaked = seeded((b, m, d), seed + 5)
    z = seeded((b, n, d), seed + 3)
    res = {}
    real_g = real.clone().requires_grad_()
    out = au(real_g, si)
    out.sum().backward()
    res.update(au_out=npy(out), au_g_real=npy(real_g.grad), au_grads=grad_summary(au), au_gsamp=grad_samples(au))
    au.zero_grad()
    with inject_randn([z]):
        fake = im(leaked, n, True)
    res["fake"] = npy(fake)
    tr = DataParallelMock(GIMGaussianTrainer("/tmp/gim_golden_out", m, n, k, au, im, 1e-2, 1e-2, reg_param=reg))
    rec = {key: [] for key in ("im_loss", "au_loss", "reg")}
    for it in range(iters):
        real = seeded((b, n, d), seed + 100 * it + 1)
        si = seeded((b, k, d), seed + 100 * it + 2)
        leaked = seeded((b, m, d), seed + 100 * it + 5)
        z = seeded((b, n, d), seed + 100 * it + 3)
        tr.module.do_global_step()
        with inject_randn([z]):
            im_loss, fake, _ = ref_gauss_loop.im_train_step(tr, leaked, si)
        o = ref_gauss_loop.au_train_step(tr, real, fake, si)
        rec["im_loss"].append(im_loss.item())
        rec["au_loss"].append(o[0].item())
        rec["reg"].append(o[3].item())
    res.update({key: np.asarray(v) for key, v in rec.items()})
    if compact:          # wide models: element samples instead of whole tensors
        res["au_final_samp"], res["im_final_samp"] = state_samples(au), state_samples(im)
        res["au_g_real"], res["fake"] = res["au_g_real"].astype(np.float32), res["fake"].astype(np.float32)
        return res
    for key, v in au.state_dict().items():
        res["au_final." + key] = npy(v)
    for key, v in im.state_dict().items():
        res["im_final." + key] = npy(v)
    return res


def case_sn_steps(seed):
    """weight_u / weight_v / effective weight after 3 train-mode calls of one spectral-normed conv."""
    conv = torch.nn.utils.spectral_norm(torch.nn.Conv2d(6, 8, 3, padding=1))
    sd = fill_state_dict(schema_of(conv), seed)
    conv.load_state_dict(sd)
    conv.double().train()
    x = seeded((2, 6, 5, 5), seed + 1)
    outs = []
    for _ in range(3):
        outs.append(npy(conv(x)))
    return {"y": np.stack(outs), "u": npy(conv.weight_u), "v": npy(conv.weight_v)}


def case_episodes():
    """img_datasets.py:79-85 / 160-169 with an explicitly seeded `random`."""
    rows = []
    for seed, index, per_cls, n_imgs, m, n, k in [(0, 7, 4, 20, 1, 5, 5), (1, 123, 10, 37, 5, 5, 5), (2, 0, 1, 15, 5, 5, 5)]:
        random.seed(seed)
        cls = index // per_cls
        idx = random.sample(list(range(n_imgs)), m + n + k)
        rows.append({"seed": seed, "index": index, "per_cls": per_cls, "n_imgs": n_imgs, "m": m, "n": n, "k": k,
                     "cls": cls, "leaked": idx[:m], "real": idx[m:m + n], "si": idx[m + n:]})
    return rows


def optimizer_groups(im):
    return [len(list(mod.parameters())) for mod in
            (im.src_encoder, im.env_encoder, im.env_decoder, im.img2img, im.img_att, im.env_noise_mapper)]


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    schemas = {}
    for name, (size, ch, sd_dim) in {"s16": (16, 3, 64), "O": (32, 1, 512), "V": (64, 3, 512)}.items():
        au = ref_img.get_au(size, ch, sd_dim)
        im = ref_img.get_im(size, ch, sd_dim)
        schemas[name] = {
            "cfg": [size, ch, sd_dim],
            "au": schema_of(au), "im": schema_of(im),
            "au_params": [k for k, _ in au.named_parameters()],
            "im_params": [k for k, _ in im.named_parameters()],
            "im_groups": optimizer_groups(im),
        }
    for d in (10, 1000):
        schemas["gauss%d" % d] = {"au": schema_of(ref_gauss.get_au(d)), "im": schema_of(ref_gauss.get_im(d))}
    # init parity: the product must reproduce the reference's own initialisation under the same seed
    torch.manual_seed(1)
    au = ref_img.get_au(16, 3, 64)
    im = ref_img.get_im(16, 3, 64)
    schemas["init_seed1_s16"] = {"au": param_summary(au).tolist(), "im": param_summary(im).tolist()}
    schemas["episodes"] = case_episodes()
    with open(os.path.join(GOLDEN, "schemas.json"), "w") as f:
        json.dump(schemas, f)

    cases = {
        "au_s16": lambda: case_au(16, 3, 64, 2, 3, 2, 11),
        "im_s16": lambda: case_im(16, 3, 64, 2, 2, 3, 21),
        "im_s16_att": lambda: case_im(16, 3, 64, 2, 2, 3, 221, use_img_att=True),
        "steps_s16_r1": lambda: case_img_steps(16, 3, 64, 2, 2, 3, 2, 10.0, 2, 31),
        "steps_s16_noreg": lambda: case_img_steps(16, 3, 64, 2, 2, 3, 2, 0.0, 2, 41),
        "au_O": lambda: case_au(32, 1, 512, 1, 2, 2, 51, with_grads=False),
        "im_O": lambda: case_im(32, 1, 512, 1, 1, 1, 61, with_grads=False),
        "au_V": lambda: case_au(64, 3, 512, 1, 1, 1, 71, with_grads=False),
        "im_V": lambda: case_im(64, 3, 512, 1, 1, 1, 81, with_grads=False),
        "step_O": lambda: case_step_grads(32, 1, 512, 2, 2, 2, 2, 0.0, 151),
        "step_V": lambda: case_step_grads(64, 3, 512, 2, 1, 2, 1, 10.0, 171),
        "gauss_d1000": lambda: case_gauss(1000, 8, 1, 5, 10, 2, 191, compact=True),
        "gauss_d10": lambda: case_gauss(10, 16, 1, 5, 10, 3, 91),
        "gauss_d10_r1": lambda: case_gauss(10, 16, 2, 3, 4, 2, 95, reg=1.0),
        "sn_steps": lambda: case_sn_steps(5),
    }
    only = sys.argv[1:]
    for name, fn in cases.items():
        if only and name not in only:
            continue
        res = fn()
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **res)
        print(name, {k: getattr(v, "shape", None) for k, v in res.items()})


if __name__ == "__main__":
    main()
