"""CPU: pin oracle/gim_oracle.py against vectors produced by the UNMODIFIED reference (oracle/make_golden.py).

The golden vectors are the reference evaluated in float64.  The oracle evaluated in float64 must agree to
round-off (TOL); that pins the restatement's semantics.  (The fp32 noise floor of the same algorithm is measured
in tests/test_parity_gpu.py, where it sets the tolerance for ill-conditioned quantities.)
"""
import numpy as np
import torch

from conftest import load_golden, rel_err, sample_errs, sample_rows
from oracle import gim_oracle as O
from oracle.fill import fill_state_dict
from oracle.fill import seeded as _seeded32

TOL = 1e-9
GTOL = 1e-7


def seeded(*a, **k):
    return _seeded32(*a, **k).double()


def params_of(schema, seed, grad=True):
    p = {k: v.double() for k, v in fill_state_dict([(k, s) for k, s in schema], seed).items()}
    if grad:
        for k, v in p.items():
            if not k.endswith(("weight_u", "weight_v")):
                v.requires_grad_()
    return p


def grad_rows(p, names):
    rows = []
    for n in names:
        g = p[n].grad
        rows.append([np.nan, np.nan] if g is None else [g.double().sum().item(), g.double().norm().item()])
    return np.asarray(rows)


def check_rows(got, want, tol):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    nan = np.isnan(want[:, 1])
    assert (np.isnan(got[:, 1]) == nan).all()
    scale = np.maximum(want[~nan, 1], 1e-6 * np.nanmax(want[:, 1]))
    assert (np.abs(got[~nan, 1] - want[~nan, 1]) / scale).max() < tol
    assert (np.abs(got[~nan, 0] - want[~nan, 0]) / (scale * 30)).max() < tol


def check_samples(p, names, gold, samp_key, full_prefix, tol=GTOL):
    """Element-level pin: the sampled gradient elements of EVERY tensor and the whole tensors stored in the golden."""
    err = sample_errs(sample_rows([p[n].grad for n in names]), gold[samp_key])
    assert np.nanmax(err) < tol, [(names[j], err[j]) for j in np.argsort(-np.nan_to_num(err))[:4]]
    full = [k for k in gold if k.startswith(full_prefix)]
    assert len(full) >= 8
    for k in full:
        assert rel_err(p[k[len(full_prefix):]].grad, gold[k]) < 1e-6, k        # stored as float32


def test_spectral_norm_steps():
    g = load_golden("sn_steps")
    p = fill_state_dict([("c.bias", [8]), ("c.weight_orig", [8, 6, 3, 3]), ("c.weight_u", [8]), ("c.weight_v", [54])], 5)
    p = {k: v.double() for k, v in p.items()}
    x = seeded((2, 6, 5, 5), 6)
    for i in range(3):
        y = O.sn_conv(p, "c", x, 1, training=True)
        assert rel_err(y, g["y"][i]) < TOL
    assert rel_err(p["c.weight_u"], g["u"]) < TOL and rel_err(p["c.weight_v"], g["v"]) < TOL


def test_authenticator_small(schemas):
    g = load_golden("au_s16")
    s = schemas["s16"]
    p = params_of(s["au"], 11)
    test = seeded((2, 3, 3, 16, 16), 12, 0.5, 1.0).requires_grad_()
    si = seeded((2, 2, 3, 16, 16), 13, 0.5, 1.0).requires_grad_()
    out = O.authenticator(p, test, si, training=True)
    assert rel_err(out, g["out"]) < TOL
    loss = O.gan_loss(out, 1.0).mean()
    loss.backward()
    assert rel_err(loss, g["loss"]) < TOL
    assert rel_err(test.grad, g["g_test"]) < GTOL and rel_err(si.grad, g["g_si"]) < GTOL
    check_rows(grad_rows(p, s["au_params"]), g["grads"], GTOL)
    check_samples(p, s["au_params"], g, "gsamp", "gfull.")
    assert rel_err(p["dis.mlp.model.4.weight"].grad, g["g_mlp_last"]) < GTOL
    assert rel_err(p["src_encoder.down_blocks.0.conv_r1.weight_u"], g["u_after"]) < TOL
    with torch.no_grad():
        out_eval = O.authenticator(p, test, si, training=False)
    assert rel_err(out_eval, g["out_eval"]) < TOL


def test_impersonator_small(schemas):
    g = load_golden("im_s16")
    s = schemas["s16"]
    p = params_of(s["im"], 21)
    leaked = seeded((2, 2, 3, 16, 16), 22, 0.5, 1.0)
    z = seeded((2, 3, 64), 24)
    fake = O.impersonator(p, leaked, 3, z)
    assert rel_err(fake, g["fake"]) < TOL
    (fake * seeded(tuple(fake.shape), 25)).sum().backward()
    check_rows(grad_rows(p, s["im_params"]), g["grads"], GTOL)
    check_samples(p, s["im_params"], g, "gsamp", "gfull.")
    assert rel_err(p["env_noise_mapper.model.6.weight"].grad, g["g_noise_last"]) < GTOL


def _run_steps(schemas, name, reg, seed, iters=2, b=2, m=2, n=3, k=2, lrs=(1e-3, 1e-3, 1e-4)):
    g = load_golden(name)
    s = schemas["s16"]
    pa = params_of(s["au"], seed)
    pi = params_of(s["im"], seed + 10)
    a_names, i_names = s["au_params"], s["im_params"]
    # optimizer groups gim_img_trainer.py:51-58: the noise mapper has its own lr
    i_lr = [lrs[2] if n_.startswith("env_noise_mapper.") else lrs[1] for n_ in i_names]
    st_a = {"step": 0, "m": [torch.zeros_like(pa[n_]) for n_ in a_names], "v": [torch.zeros_like(pa[n_]) for n_ in a_names]}
    st_i = [{"step": 0, "m": [torch.zeros_like(pi[n_])], "v": [torch.zeros_like(pi[n_])]} for n_ in i_names]
    rec = {key: [] for key in ("im_loss", "au_loss", "loss_real", "loss_fake", "reg", "out_real", "out_fake")}
    for it in range(iters):
        leaked = seeded((b, m, 3, 16, 16), seed + 100 * it + 1, 0.5, 1.0)
        real = seeded((b, n, 3, 16, 16), seed + 100 * it + 2, 0.5, 1.0)
        si = seeded((b, k, 3, 16, 16), seed + 100 * it + 3, 0.5, 1.0)
        z = seeded((b, n, 64), seed + 100 * it + 4)
        # G step (gim_img_training.py:157-166)
        for v in list(pa.values()) + list(pi.values()):
            v.grad = None
        fake = O.impersonator(pi, leaked, n, z)
        loss = O.gan_loss(O.authenticator(pa, fake, si), 1.0).mean()
        loss.backward()
        rec["im_loss"].append(loss.item())
        for j, n_ in enumerate(i_names):
            O.adam_step([pi[n_]], [pi[n_].grad], st_i[j], i_lr[j], 0.0, 0.99)
        if it == 0:
            assert rel_err(fake, g["fake0"]) < TOL
        # D step (gim_img_training.py:169-183)
        for v in pa.values():
            v.grad = None
        o = O.img_authenticator_forward(pa, fake.detach(), real, si, reg)
        o[0].mean().backward()
        O.adam_step([pa[n_] for n_ in a_names], [pa[n_].grad for n_ in a_names], st_a, lrs[0], 0.0, 0.99)
        for key, val in zip(("au_loss", "loss_real", "loss_fake", "reg", "out_real", "out_fake"), o):
            rec[key].append(val.mean().item())
    for key, v in rec.items():
        assert rel_err(v, g[key]) < GTOL, key
    au_rows = np.asarray([[v.double().sum().item(), v.double().norm().item()] for v in pa.values()])
    im_rows = np.asarray([[v.double().sum().item(), v.double().norm().item()] for v in pi.values()])
    assert np.abs(au_rows[:, 1] - g["au_params"][:, 1]).max() / np.abs(g["au_params"][:, 1]).max() < GTOL
    assert np.abs(im_rows[:, 1] - g["im_params"][:, 1]).max() / np.abs(g["im_params"][:, 1]).max() < GTOL


def test_training_steps_r1(schemas):
    _run_steps(schemas, "steps_s16_r1", 10.0, 31)


def test_training_steps_noreg(schemas):
    _run_steps(schemas, "steps_s16_noreg", 0.0, 41)


def test_full_size_forward(schemas):
    for name, seed, (size, ch), (b, n, k) in (("O", 51, (32, 1), (1, 2, 2)), ("V", 71, (64, 3), (1, 1, 1))):
        g = load_golden("au_" + name)
        p = params_of(schemas[name]["au"], seed, grad=False)
        test = seeded((b, n, ch, size, size), seed + 1, 0.5, 1.0)
        si = seeded((b, k, ch, size, size), seed + 2, 0.5, 1.0)
        with torch.no_grad():
            assert rel_err(O.authenticator(p, test, si, True), g["out"]) < TOL
            assert rel_err(O.authenticator(p, test, si, False), g["out_eval"]) < TOL
    for name, seed, (size, ch) in (("O", 61, (32, 1)), ("V", 81, (64, 3))):
        g = load_golden("im_" + name)
        p = params_of(schemas[name]["im"], seed, grad=False)
        leaked = seeded((1, 1, ch, size, size), seed + 1, 0.5, 1.0)
        z = seeded((1, 1, 512), seed + 3)
        with torch.no_grad():
            assert rel_err(O.impersonator(p, leaked, 1, z), g["fake"]) < GTOL


def test_gaussian(schemas):
    for name, seed, (m, n, k), reg, iters in (("gauss_d10", 91, (1, 5, 10), 0.0, 3), ("gauss_d10_r1", 95, (2, 3, 4), 1.0, 2)):
        g = load_golden(name)
        d, b = 10, 16
        pa = params_of(schemas["gauss10"]["au"], seed)
        pi = params_of(schemas["gauss10"]["im"], seed + 10)
        real = seeded((b, n, d), seed + 1).requires_grad_()
        si = seeded((b, k, d), seed + 2)
        out = O.gaussian_authenticator(pa, real, si)
        out.sum().backward()
        assert rel_err(out, g["au_out"]) < TOL and rel_err(real.grad, g["au_g_real"]) < GTOL
        fake = O.gaussian_impersonator(pi, seeded((b, m, d), seed + 5), n, seeded((b, n, d), seed + 3))
        assert rel_err(fake, g["fake"]) < TOL
        a_names = list(pa.keys())
        i_names = list(pi.keys())
        st_a = {"step": 0, "m": [torch.zeros_like(pa[x]) for x in a_names], "v": [torch.zeros_like(pa[x]) for x in a_names]}
        st_i = {"step": 0, "m": [torch.zeros_like(pi[x]) for x in i_names], "v": [torch.zeros_like(pi[x]) for x in i_names]}
        for it in range(iters):
            real = seeded((b, n, d), seed + 100 * it + 1)
            si = seeded((b, k, d), seed + 100 * it + 2)
            leaked = seeded((b, m, d), seed + 100 * it + 5)
            z = seeded((b, n, d), seed + 100 * it + 3)
            for v in list(pa.values()) + list(pi.values()):
                v.grad = None
            fake = O.gaussian_impersonator(pi, leaked, n, z)
            loss = O.gan_loss(O.gaussian_authenticator(pa, fake, si), 1.0).mean()
            loss.backward()
            assert abs(loss.item() - g["im_loss"][it]) < GTOL * max(1, abs(g["im_loss"][it]))
            O.adam_step([pi[x] for x in i_names], [pi[x].grad for x in i_names], st_i, 1e-2, 0.9, 0.999)
            for v in pa.values():
                v.grad = None
            o = O.gaussian_authenticator_forward(pa, fake.detach(), real, si, reg)
            o[0].mean().backward()
            assert abs(o[0].mean().item() - g["au_loss"][it]) < GTOL * max(1, abs(g["au_loss"][it]))
            O.adam_step([pa[x] for x in a_names], [pa[x].grad for x in a_names], st_a, 1e-2, 0.9, 0.999)
        for x in a_names:
            assert rel_err(pa[x], g["au_final." + x]) < GTOL, x
        for x in i_names:
            assert rel_err(pi[x], g["im_final." + x]) < GTOL, x


def test_full_width_training_step_gradients(schemas):
    """One G-step + D-step of the reference trainer at full width (Omniglot-shaped, batch 2): losses, generated images and the
    gradients both optimizers consume, element by element.  (The VoxCeleb2-shaped twin with R1, step_V, is checked on the GPU box.)"""
    g = load_golden("step_O")
    s = schemas["O"]
    seed, b, m, n, k, size, ch = 151, 2, 2, 2, 2, 32, 1
    pa, pi = params_of(s["au"], seed), params_of(s["im"], seed + 10)
    leaked = seeded((b, m, ch, size, size), seed + 1, 0.5, 1.0)
    real = seeded((b, n, ch, size, size), seed + 2, 0.5, 1.0)
    si = seeded((b, k, ch, size, size), seed + 3, 0.5, 1.0)
    z = seeded((b, n, 512), seed + 4)
    fake = O.impersonator(pi, leaked, n, z)
    loss = O.gan_loss(O.authenticator(pa, fake, si), 1.0).mean()
    loss.backward()
    assert rel_err(fake, g["fake"]) < 1e-6 and rel_err(loss, g["im_loss"]) < GTOL
    check_rows(grad_rows(pi, s["im_params"]), g["im_grads"], 10 * GTOL)
    check_samples(pi, s["im_params"], g, "im_gsamp", "im_gfull.", 10 * GTOL)
    for v in pa.values():
        v.grad = None
    o = O.img_authenticator_forward(pa, fake.detach(), real, si, 0.0)
    o[0].mean().backward()
    assert rel_err(o[0].mean(), g["au_loss"]) < GTOL
    check_samples(pa, s["au_params"], g, "au_gsamp", "au_gfull.", 10 * GTOL)


def test_gaussian_d1000(schemas):
    """BASELINE configs[3] width: forward, input gradient, two training iterations (sampled post-Adam parameters)."""
    g = load_golden("gauss_d1000")
    d, b, (m, n, k), seed = 1000, 8, (1, 5, 10), 191
    pa, pi = params_of(schemas["gauss1000"]["au"], seed), params_of(schemas["gauss1000"]["im"], seed + 10)
    real = seeded((b, n, d), seed + 1).requires_grad_()
    si = seeded((b, k, d), seed + 2)
    out = O.gaussian_authenticator(pa, real, si)
    out.sum().backward()
    assert rel_err(out, g["au_out"]) < TOL and rel_err(real.grad, g["au_g_real"]) < 1e-6
    a_names, i_names = list(pa.keys()), list(pi.keys())
    err = sample_errs(sample_rows([pa[x].grad for x in a_names]), g["au_gsamp"])
    assert np.nanmax(err) < GTOL
    st_a = {"step": 0, "m": [torch.zeros_like(pa[x]) for x in a_names], "v": [torch.zeros_like(pa[x]) for x in a_names]}
    st_i = {"step": 0, "m": [torch.zeros_like(pi[x]) for x in i_names], "v": [torch.zeros_like(pi[x]) for x in i_names]}
    for it in range(2):
        real = seeded((b, n, d), seed + 100 * it + 1)
        si = seeded((b, k, d), seed + 100 * it + 2)
        leaked = seeded((b, m, d), seed + 100 * it + 5)
        z = seeded((b, n, d), seed + 100 * it + 3)
        for v in list(pa.values()) + list(pi.values()):
            v.grad = None
        fake = O.gaussian_impersonator(pi, leaked, n, z)
        loss = O.gan_loss(O.gaussian_authenticator(pa, fake, si), 1.0).mean()
        loss.backward()
        assert abs(loss.item() - g["im_loss"][it]) < GTOL * max(1, abs(g["im_loss"][it]))
        O.adam_step([pi[x] for x in i_names], [pi[x].grad for x in i_names], st_i, 1e-2, 0.9, 0.999)
        for v in pa.values():
            v.grad = None
        o = O.gaussian_authenticator_forward(pa, fake.detach(), real, si, 0.0)
        o[0].mean().backward()
        assert abs(o[0].mean().item() - g["au_loss"][it]) < GTOL * max(1, abs(g["au_loss"][it]))
        O.adam_step([pa[x] for x in a_names], [pa[x].grad for x in a_names], st_a, 1e-2, 0.9, 0.999)
    assert np.nanmax(sample_errs(sample_rows([pa[x] for x in a_names]), g["au_final_samp"], 0.0)) < 1e-6
    assert np.nanmax(sample_errs(sample_rows([pi[x] for x in i_names]), g["im_final_samp"], 0.0)) < 1e-6


def test_episode_indices(schemas):
    for row in schemas["episodes"]:
        rng = O.make_rng(row["seed"])
        cls, leaked, real, si = O.episode_indices(row["index"], row["per_cls"], row["n_imgs"], row["m"], row["n"], row["k"], rng)
        assert (cls, leaked, real, si) == (row["cls"], row["leaked"], row["real"], row["si"])
