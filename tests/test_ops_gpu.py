"""GPU: every C-ABI operator (forward, backward and -- on the authenticator path -- double backward) against a float64
torch-CPU statement of the same op.  fp32 gate 1e-4 (north_star), bf16 gate 2e-2."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


@pytest.fixture(autouse=True)
def _cuda_only():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from optimalstrategiesagainstgenerativeattacks_b200 import ops
    ops.set_precision("fp32")
    ops.set_conv_algo("auto")
    yield
    ops.set_precision("fp32")
    ops.set_conv_algo("auto")


def ops_mod():
    from optimalstrategiesagainstgenerativeattacks_b200 import ops
    return ops


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale)


def to_dev_nhwc(x64, dtype):
    """NCHW float64 CPU -> NHWC device tensor (leaf)."""
    return x64.permute(0, 2, 3, 1).contiguous().to("cuda", dtype).requires_grad_()


def nchw(t):
    return t.detach().double().cpu().permute(0, 3, 1, 2)


def pack_w(w64):
    """OIHW -> packed fp32 [k*k, co, ci]."""
    co, ci, k, _ = w64.shape
    return w64.permute(2, 3, 0, 1).reshape(k * k, co, ci).contiguous().to("cuda", torch.float32).requires_grad_()


def unpack_w(gw, k):
    taps, co, ci = gw.shape
    return gw.detach().double().cpu().reshape(k, k, co, ci).permute(2, 3, 0, 1)


CONV_SHAPES = [  # n, ci, co, k, h, w
    (2, 3, 64, 3, 16, 16), (3, 64, 64, 3, 8, 8), (2, 6, 64, 9, 16, 16), (2, 64, 3, 9, 16, 16), (2, 64, 16, 1, 7, 5),
    (1, 1, 128, 3, 32, 32), (5, 128, 128, 3, 4, 4), (2, 64, 128, 3, 13, 13), (130, 32, 8, 1, 1, 1), (3, 64, 40, 3, 9, 7), (200, 72, 1000, 1, 1, 1),
    (4, 64, 256, 3, 16, 16), (9, 128, 512, 1, 8, 8),          # 256-wide N tiles: the cta_group::2 (two-CTA cluster) kernel variant
]


@pytest.mark.parametrize("prec,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_fwd_bwd_double_bwd(shape, prec, tol):
    ops = ops_mod()
    ops.set_precision(prec)
    dt = ops.act_dtype()
    n, ci, co, k, h, w = shape
    x64 = rnd(n, ci, h, w, seed=1).requires_grad_()
    w64 = (rnd(co, ci, k, k, seed=2) / np.sqrt(ci * k * k)).requires_grad_()
    b64 = rnd(co, seed=3, scale=0.1).requires_grad_()
    probe = rnd(n, co, h, w, seed=4)
    y64 = F.conv2d(x64, w64, b64, padding=(k - 1) // 2)
    # first order + an R1-style second-order term: || d(sum(y*probe))/dx ||^2
    (gx64,) = torch.autograd.grad((y64 * probe).sum(), x64, create_graph=True)
    total64 = (y64 * probe).sum() + 0.5 * gx64.pow(2).sum()
    gX, gW, gB = torch.autograd.grad(total64, (x64, w64, b64))

    x = to_dev_nhwc(x64.detach(), dt)
    wp = pack_w(w64.detach())
    b = b64.detach().to("cuda", torch.float32).requires_grad_()
    pr = probe.permute(0, 2, 3, 1).contiguous().to("cuda", dt)
    y = ops.conv2d(x, wp, b, k)
    assert rel_err(nchw(y), y64) < tol
    s = ops.DotFn.apply(y, pr)
    (gx,) = torch.autograd.grad(s.sum(), x, create_graph=True)
    assert rel_err(nchw(gx), gx64) < tol
    total = s.sum() + 0.5 * ops.RowsSqSumFn.apply(gx.reshape(1, -1)).sum()
    dX, dW, dB = torch.autograd.grad(total, (x, wp, b))
    assert rel_err(nchw(dX), gX) < tol
    assert rel_err(unpack_w(dW, k), gW) < tol
    assert rel_err(dB, gB) < tol


def test_spectral_norm_state_machine_and_grad():
    ops = ops_mod()
    from oracle import gim_oracle as O
    for (co, ci, k) in [(8, 6, 3), (64, 128, 1), (128, 128, 9), (3, 64, 9)]:
        p = {"c.weight_orig": rnd(co, ci, k, k, seed=5).requires_grad_(), "c.weight_u": F.normalize(rnd(co, seed=6), dim=0),
             "c.weight_v": F.normalize(rnd(ci * k * k, seed=7), dim=0)}
        w = p["c.weight_orig"].detach().to("cuda", torch.float32).requires_grad_()
        u = p["c.weight_u"].to("cuda", torch.float32).clone()
        v = p["c.weight_v"].to("cuda", torch.float32).clone()
        probe = rnd(k * k, co, ci, seed=8)
        for step, training in enumerate([True, True, False, True]):
            ref = O.sn_weight(p, "c", training)                      # OIHW
            ref_packed = ref.permute(2, 3, 0, 1).reshape(k * k, co, ci)
            got = ops.SpectralNormFn.apply(w, (u, v), training, 1e-12)
            assert rel_err(got, ref_packed) < 1e-5, (co, ci, k, step)
            assert rel_err(u, p["c.weight_u"]) < 1e-5 and rel_err(v, p["c.weight_v"]) < 1e-5
        (g_ref,) = torch.autograd.grad((ref_packed * probe).sum(), p["c.weight_orig"])
        (g_got,) = torch.autograd.grad((got * probe.to("cuda", torch.float32)).sum(), w)
        assert rel_err(g_got, g_ref) < 1e-4


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-6), ("bf16", 1e-2)])
def test_pointwise_and_resampling(prec, tol):
    ops = ops_mod()
    ops.set_precision(prec)
    dt = ops.act_dtype()
    x64 = rnd(3, 5, 13, 9, seed=1).requires_grad_()
    r64 = rnd(3, 5, 13, 9, seed=2).requires_grad_()
    x, r = to_dev_nhwc(x64.detach(), dt), to_dev_nhwc(r64.detach(), dt)
    xq, rq = nchw(x).requires_grad_(), nchw(r).requires_grad_()       # the rounded inputs the device actually sees
    # leaky relu + avg-pool-of-sum (odd sizes floor) + nearest upsample + tanh, one chain
    ref = torch.tanh(F.interpolate(F.avg_pool2d(F.leaky_relu(xq, 0.2) + rq, 2), scale_factor=2, mode="nearest"))
    got = ops.TanhFn.apply(ops.upsample2(ops.avg_pool2_add(ops.lrelu(x), r)))
    assert got.shape == (3, 12, 8, 5)
    assert rel_err(nchw(got), ref) < max(tol, 1e-6)
    probe = rnd(*ref.shape, seed=3)
    gr = torch.autograd.grad((ref * probe).sum(), (xq, rq))
    gg = torch.autograd.grad(ops.DotFn.apply(got, probe.permute(0, 2, 3, 1).contiguous().to("cuda", dt)).sum(), (x, r))
    assert rel_err(nchw(gg[0]), gr[0]) < max(tol, 1e-5) and rel_err(nchw(gg[1]), gr[1]) < max(tol, 1e-5)
    # layout round trip
    img = rnd(4, 3, 10, 7, seed=4).float().cuda()
    assert rel_err(ops.from_nhwc(ops.to_nhwc(img)), img) < (1e-7 if prec == "fp32" else 4e-3)
    # channel concat
    a, b = rnd(2, 3, 4, 4, seed=5), rnd(2, 6, 4, 4, seed=6)
    cat = ops.CatChannelsFn.apply(to_dev_nhwc(a, dt), to_dev_nhwc(b, dt))
    assert rel_err(nchw(cat), torch.cat((a, b), 1)) < (1e-7 if prec == "fp32" else 4e-3)


@pytest.mark.parametrize("prec,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
@pytest.mark.parametrize("shape", [(3, 16, 8, 8), (2, 64, 1, 1), (2, 70, 16, 16), (4, 512, 2, 2)])
def test_instance_norm_and_ada_in(shape, prec, tol):
    from oracle import gim_oracle as O
    ops = ops_mod()
    ops.set_precision(prec)
    dt = ops.act_dtype()
    n, c, h, w = shape
    x = to_dev_nhwc(rnd(n, c, h, w, seed=1) + 0.5, dt)
    xq = nchw(x).requires_grad_()
    probe = rnd(n, c, h, w, seed=9)
    pr = probe.permute(0, 2, 3, 1).contiguous().to("cuda", dt)
    wt, bs = (1 + 0.1 * rnd(c, seed=2)).requires_grad_(), (0.1 * rnd(c, seed=3)).requires_grad_()
    ref = O.lrelu(O.instance_norm(xq, wt, bs))
    wd, bd = wt.detach().float().cuda().requires_grad_(), bs.detach().float().cuda().requires_grad_()
    got = ops.instance_norm(x, wd, bd, 1e-5, 0.2)
    assert rel_err(nchw(got), ref) < tol
    gr = torch.autograd.grad((ref * probe).sum(), (xq, wt, bs))
    gg = torch.autograd.grad(ops.DotFn.apply(got, pr).sum(), (x, wd, bd))
    if h * w > 1:
        assert rel_err(nchw(gg[0]), gr[0]) < tol * 5
    else:
        assert float(gg[0].abs().max()) == 0.0
    assert rel_err(gg[1], gr[1]) < tol * 5 + 1e-6 and rel_err(gg[2], gr[2]) < tol * 5
    if h * w > 1:
        ms, ss = rnd(n, c, seed=4).requires_grad_(), (1 + 0.2 * rnd(n, c, seed=5)).requires_grad_()
        xq2 = nchw(x).requires_grad_()
        ref = O.ada_in(xq2, ms, ss)
        msd, ssd = ms.detach().float().cuda().requires_grad_(), ss.detach().float().cuda().requires_grad_()
        x2 = x.detach().clone().requires_grad_()
        got = ops.ada_in(x2, msd, ssd, 1e-5)
        assert rel_err(nchw(got), ref) < tol
        gr = torch.autograd.grad((ref * probe).sum(), (xq2, ms, ss))
        gg = torch.autograd.grad(ops.DotFn.apply(got, pr).sum(), (x2, msd, ssd))
        assert rel_err(nchw(gg[0]), gr[0]) < tol * 5
        assert rel_err(gg[1], gr[1]) < tol * 5 and rel_err(gg[2], gr[2]) < tol * 5


def test_matmul_linear_softmax_with_double_backward():
    ops = ops_mod()
    for ta in (False, True):
        for tb in (False, True):
            a64 = rnd(3, *((17, 70) if ta else (70, 17)), seed=1).requires_grad_()
            b64 = rnd(3, *((33, 17) if tb else (17, 33)), seed=2).requires_grad_()
            ref = torch.matmul(a64.transpose(1, 2) if ta else a64, b64.transpose(1, 2) if tb else b64)
            a, b = a64.detach().float().cuda().requires_grad_(), b64.detach().float().cuda().requires_grad_()
            got = ops.matmul(a, b, ta, tb)
            assert rel_err(got, ref) < 1e-5
            probe = rnd(*ref.shape, seed=3)
            gr = torch.autograd.grad((ref * probe).sum(), (a64, b64))
            gg = torch.autograd.grad((got * probe.float().cuda()).sum(), (a, b))
            assert rel_err(gg[0], gr[0]) < 1e-5 and rel_err(gg[1], gr[1]) < 1e-5
    # attention-shaped chain with second order: softmax(g f^T) h
    f64, g64, h64 = (rnd(2, 20, 8, seed=s).requires_grad_() for s in (4, 5, 6))
    def chain(f, g, h, mm, sm):
        return mm(sm(mm(g, f, False, True)), h, False, False)
    ref = chain(f64, g64, h64, lambda a, b, ta, tb: torch.matmul(a, b.transpose(1, 2) if tb else b), lambda t: torch.softmax(t, -1))
    f, g, h = (t.detach().float().cuda().requires_grad_() for t in (f64, g64, h64))
    got = chain(f, g, h, ops.matmul, ops.SoftmaxRowsFn.apply)
    assert rel_err(got, ref) < 1e-5
    probe = rnd(*ref.shape, seed=7)
    g1r = torch.autograd.grad((ref * probe).sum(), (f64, g64), create_graph=True)
    g2r = torch.autograd.grad(sum(t.pow(2).sum() for t in g1r), (f64, g64, h64))
    g1 = torch.autograd.grad((got * probe.float().cuda()).sum(), (f, g), create_graph=True)
    g2 = torch.autograd.grad(sum(ops.RowsSqSumFn.apply(t.reshape(1, -1)).sum() for t in g1), (f, g, h))
    for a_, b_ in zip(g1 + g2, g1r + g2r):
        assert rel_err(a_, b_) < 1e-4
    # linear + fused bias / leaky relu, second order through the mask
    x64, w64, b64 = rnd(9, 40, seed=8).requires_grad_(), rnd(24, 40, seed=9).requires_grad_(), rnd(24, seed=10).requires_grad_()
    ref = F.leaky_relu(F.linear(x64, w64, b64), 0.2)
    x, w, b = (t.detach().float().cuda().requires_grad_() for t in (x64, w64, b64))
    got = ops.linear(x, w, b, 0.2)
    assert rel_err(got, ref) < 1e-5
    (gxr,) = torch.autograd.grad(ref.sum(), x64, create_graph=True)
    gr = torch.autograd.grad(gxr.pow(2).sum(), (w64,))
    (gx,) = torch.autograd.grad(got.sum(), x, create_graph=True)
    gg = torch.autograd.grad(ops.RowsSqSumFn.apply(gx.reshape(1, -1)).sum(), (w,))
    assert rel_err(gx, gxr) < 1e-5 and rel_err(gg[0], gr[0]) < 1e-4


def test_set_statistics_with_double_backward():
    from oracle import gim_oracle as O
    ops = ops_mod()
    for (b, s, d) in [(4, 5, 64), (3, 1, 10), (7, 10, 33)]:
        x64 = rnd(b, s, d, seed=1).requires_grad_()
        x = x64.detach().float().cuda().requires_grad_()
        ref = torch.cat((x64.mean(1), O.custom_std(x64)), -1)
        got = torch.cat((ops.set_mean(x), ops.set_std(x)), -1)
        assert rel_err(got, ref) < 1e-5 or (s == 1 and float(got[:, d:].abs().max()) == 0)
        xf = x64.detach().float().cuda().requires_grad_()
        fused = ops.set_mean_std(xf)                   # one pass, mean | std side by side
        assert torch.equal(fused, got.detach())
        if s == 1:
            continue
        pf, qf = rnd(b, 2 * d, seed=2).float().cuda(), rnd(b, s, d, seed=3).float().cuda()
        (f1,) = torch.autograd.grad((fused * pf).sum(), xf, create_graph=True)
        (f2,) = torch.autograd.grad(ops.DotFn.apply(f1, qf).sum(), xf)
        probe = rnd(b, 2 * d, seed=2)
        q = rnd(b, s, d, seed=3)          # (sum g1^2 is ~independent of x for the std term, so contract with a random tensor)
        (g1r,) = torch.autograd.grad((ref * probe).sum(), x64, create_graph=True)
        (g2r,) = torch.autograd.grad((g1r * q).sum(), x64)
        (g1,) = torch.autograd.grad((got * probe.float().cuda()).sum(), x, create_graph=True)
        (g2,) = torch.autograd.grad(ops.DotFn.apply(g1, q.float().cuda()).sum(), x)
        assert rel_err(g1, g1r) < 1e-5 and rel_err(g2, g2r) < 1e-4
        assert rel_err(f1, g1r) < 1e-5 and rel_err(f2, g2r) < 1e-4
    w64, add64 = rnd(3, 4, 6, seed=3).requires_grad_(), rnd(3, 6, seed=4).requires_grad_()
    ref = w64 - w64.mean(1, keepdim=True) + add64.unsqueeze(1)
    w, add = w64.detach().float().cuda().requires_grad_(), add64.detach().float().cuda().requires_grad_()
    got = ops.SetCenterAddFn.apply(w, add, True)
    probe = rnd(3, 4, 6, seed=5)
    gr = torch.autograd.grad((ref * probe).sum(), (w64, add64))
    gg = torch.autograd.grad((got * probe.float().cuda()).sum(), (w, add))
    assert rel_err(got, ref) < 1e-6 and rel_err(gg[0], gr[0]) < 1e-6 and rel_err(gg[1], gr[1]) < 1e-6


def test_gaussian_episode_sampler_distribution():
    """Device-side episode synthesis (reference training/gim_gaussian_training.py:71-86): mu ~ N(0, prior^2), x | mu ~ N(mu, src^2)."""
    ops = ops_mod()
    torch.manual_seed(3)
    mu, (leaked, real, si) = ops.gaussian_episodes(4096, (1, 5, 10), 16, 10.0, 1.0, torch.device("cuda"))
    assert leaked.shape == (4096, 1, 16) and real.shape == (4096, 5, 16) and si.shape == (4096, 10, 16) and mu.shape == (4096, 16)
    assert abs(float(mu.std()) - 10.0) < 0.2 and abs(float(mu.mean())) < 0.2
    for x in (leaked, real, si):
        r = x - mu.unsqueeze(1)
        assert abs(float(r.std()) - 1.0) < 0.02 and abs(float(r.mean())) < 0.02
    # samples of one episode share mu, different sets are independent draws
    assert abs(float(((real.mean(1) - mu) * (si.mean(1) - mu)).mean())) < 0.01
    torch.manual_seed(3)
    mu2, _ = ops.gaussian_episodes(4096, (1, 5, 10), 16, 10.0, 1.0, torch.device("cuda"))
    assert torch.equal(mu, mu2)                        # seeded by torch.manual_seed


def test_global_max_bce_rows():
    ops = ops_mod()
    x64 = rnd(3, 40, 6, 5, seed=1).requires_grad_()
    x = to_dev_nhwc(x64.detach(), torch.float32)
    ref = F.leaky_relu(torch.amax(x64, dim=(2, 3)), 0.2)
    got = ops.lrelu(ops.GlobalMaxFn.apply(x))
    assert rel_err(got, ref) < 1e-6
    probe = rnd(3, 40, seed=2)
    # the one-pass encoder tail (max + LeakyReLU in the same kernel) gives the same values, first and second order
    xf = to_dev_nhwc(x64.detach(), torch.float32)
    fused = ops.GlobalMaxFn.apply(xf, 0.2)
    assert torch.equal(fused, got.detach())
    (gf,) = torch.autograd.grad((fused * probe.float().cuda()).sum(), xf, create_graph=True)
    (gr,) = torch.autograd.grad((ref * probe).sum(), x64)
    (gg,) = torch.autograd.grad((got * probe.float().cuda()).sum(), x, create_graph=True)
    assert rel_err(nchw(gg), gr) < 1e-6 and rel_err(nchw(gf), gr) < 1e-6
    # second order: scatter's backward is a gather of the same indices
    x64b = x64.detach().clone().requires_grad_()
    xb = to_dev_nhwc(x64b.detach(), torch.float32)
    (g1r,) = torch.autograd.grad(F.leaky_relu(torch.amax(x64b, dim=(2, 3)), 0.2).pow(2).sum(), x64b, create_graph=True)
    (g2r,) = torch.autograd.grad((g1r * x64b).sum(), x64b)
    (g1,) = torch.autograd.grad(ops.GlobalMaxFn.apply(xb, 0.2).pow(2).sum(), xb, create_graph=True)
    (g2,) = torch.autograd.grad(ops.DotFn.apply(g1, xb).sum(), xb)
    assert rel_err(nchw(g1), g1r) < 1e-6 and rel_err(nchw(g2), g2r) < 1e-6
    z64 = rnd(11, 1, seed=3).requires_grad_()
    z = z64.detach().float().cuda().requires_grad_()
    for target in (0.0, 1.0):
        ref = F.binary_cross_entropy_with_logits(z64, torch.full_like(z64, target), reduction="none")
        got = ops.BCEWithLogitsFn.apply(z, target)
        assert rel_err(got, ref) < 1e-6
        assert rel_err(torch.autograd.grad(got.sum(), z)[0], torch.autograd.grad(ref.sum(), z64)[0]) < 1e-6


def test_fused_adam_matches_torch_adam():
    from optimalstrategiesagainstgenerativeattacks_b200.fused_adam import FusedAdam
    shapes = [(7,), (64, 3, 3, 3), (513,), (128, 130), (1,)]
    ref_p = [torch.nn.Parameter(rnd(*s, seed=i).float()) for i, s in enumerate(shapes)]
    dev_p = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref_p]
    ref = torch.optim.Adam([{"params": ref_p[:3], "lr": 1e-2}, {"params": ref_p[3:], "lr": 3e-3}], betas=(0.0, 0.99))
    opt = FusedAdam([{"params": dev_p[:3], "lr": 1e-2}, {"params": dev_p[3:], "lr": 3e-3}], betas=(0.0, 0.99))
    for it in range(5):
        for i, (a, b) in enumerate(zip(ref_p, dev_p)):
            if i == 4:
                continue                      # never receives a gradient: must be skipped, no state
            g = rnd(*a.shape, seed=100 * it + i).float()
            a.grad = g.clone()
            b.grad = g.clone().cuda() if b.grad is None else b.grad.copy_(g.cuda())
        ref.step()
        opt.step()
    for a, b in zip(ref_p, dev_p):
        assert rel_err(b, a) < 2e-6
    sd = opt.state_dict()
    assert set(sd["state"].keys()) == {0, 1, 2, 3} and float(sd["state"][0]["step"]) == 5.0
    assert rel_err(sd["state"][1]["exp_avg_sq"], ref.state_dict()["state"][1]["exp_avg_sq"]) < 1e-6
    # round trip through torch's own optimizer state format
    opt2 = FusedAdam([{"params": dev_p[:3], "lr": 1e-2}, {"params": dev_p[3:], "lr": 3e-3}], betas=(0.0, 0.99))
    opt2.load_state_dict(sd)
    for i, (a, b) in enumerate(zip(ref_p, dev_p)):
        if i != 4:
            g = rnd(*a.shape, seed=999 + i).float()
            a.grad = g.clone()
            b.grad.copy_(g.cuda())
    ref.step()
    opt2.step()
    for a, b in zip(ref_p, dev_p):
        assert rel_err(b, a) < 2e-6


TC_SHAPES = [  # n, ci, co, k, h, w  -- every tensor-core-eligible layer family of the O (32x32x1) and V (64x64x3) networks
    (3, 64, 64, 3, 16, 16), (2, 128, 128, 3, 32, 32), (5, 256, 256, 3, 16, 16), (3, 512, 512, 3, 8, 8), (9, 512, 512, 3, 4, 4),
    (2, 128, 128, 9, 32, 32), (2, 64, 64, 9, 64, 64), (3, 128, 256, 3, 16, 16), (4, 256, 128, 1, 16, 16), (2, 128, 16, 1, 16, 16),
    (2, 256, 32, 1, 8, 8), (130, 512, 512, 3, 1, 1), (33, 512, 512, 3, 2, 2), (1, 64, 128, 3, 52, 52), (2, 128, 256, 3, 13, 13),
    (1, 64, 64, 3, 105, 105), (2, 16, 128, 1, 16, 16), (3, 32, 256, 1, 8, 8), (2, 64, 64, 1, 4, 4),
]


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_tcgen05_matches_cuda_core(shape):
    """The tcgen05/TMEM/TMA implicit GEMM against the CUDA-core kernel on identical bf16 operands (forward and dgrad form)."""
    ops = ops_mod()
    from optimalstrategiesagainstgenerativeattacks_b200 import _cabi as C
    ops.set_precision("bf16")
    n, ci, co, k, h, w = shape
    assert C.conv_tc_supported(n, h, w, ci, co, k, C.BF16)
    x = (rnd(n, h, w, ci, seed=1)).to("cuda", torch.bfloat16)
    wp = (rnd(k * k, co, ci, seed=2) / np.sqrt(ci * k * k)).to("cuda", torch.float32)
    b = rnd(co, seed=3).to("cuda", torch.float32)
    gy = (rnd(n, h, w, co, seed=4)).to("cuda", torch.bfloat16)
    outs, wg = {}, {}
    for algo in ("simt", "tcgen05"):
        ops.set_conv_algo(algo)
        with torch.no_grad():
            outs[algo] = ops.conv2d(x, wp, b, k).float()
            wg[algo] = ops.WgradFn.apply(x, gy, k)
        torch.cuda.synchronize()
    assert rel_err(outs["tcgen05"], outs["simt"]) < 4e-3
    assert float((outs["tcgen05"] - outs["simt"]).abs().max()) < 0.05 * float(outs["simt"].abs().max())
    assert C.wgrad_tc_supported(n, h, w, ci, co, k, C.BF16)
    assert rel_err(wg["tcgen05"], wg["simt"]) < 1e-4          # same bf16 operands, fp32 accumulation in both


@pytest.mark.parametrize("prec,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_conv_with_fused_prologue(prec, tol):
    """conv(lrelu(x)) and conv(upsample2(x)) with the prologue folded into the operand producer, first and second order."""
    ops = ops_mod()
    ops.set_precision(prec)
    n, ci, co, k, h, w = 3, 64, 32, 3, 6, 10
    x64 = rnd(n, ci, h, w, seed=1).requires_grad_()
    w64 = (rnd(co, ci, k, k, seed=2) / np.sqrt(ci * k * k)).requires_grad_()
    b64 = rnd(co, seed=3, scale=0.1).requires_grad_()
    for pre, fn in ((ops.PRE_LRELU, lambda t: F.leaky_relu(t, 0.2)), (ops.PRE_UPSAMPLE, lambda t: F.interpolate(t, scale_factor=2, mode="nearest"))):
        y64 = F.conv2d(fn(x64), w64, b64, padding=1)
        probe = rnd(*y64.shape, seed=4)
        (gx64,) = torch.autograd.grad((y64 * probe).sum(), x64, create_graph=True)
        gX, gW, gB = torch.autograd.grad((y64 * probe).sum() + 0.5 * gx64.pow(2).sum(), (x64, w64, b64))
        x = to_dev_nhwc(x64.detach(), torch.float32)
        wp = pack_w(w64.detach())
        b = b64.detach().to("cuda", torch.float32).requires_grad_()
        pr = probe.permute(0, 2, 3, 1).contiguous().to("cuda", torch.float32)
        y = ops.conv2d(x, wp, b, k, pre, 0.2)
        assert rel_err(nchw(y), y64) < tol
        s = ops.DotFn.apply(y, pr)
        (gx,) = torch.autograd.grad(s.sum(), x, create_graph=True)
        assert rel_err(nchw(gx), gx64) < tol
        dX, dW, dB = torch.autograd.grad(s.sum() + 0.5 * ops.RowsSqSumFn.apply(gx.reshape(1, -1)).sum(), (x, wp, b))
        assert rel_err(nchw(dX), gX) < tol and rel_err(unpack_w(dW, k), gW) < tol and rel_err(dB, gB) < tol


@pytest.mark.parametrize("shape", [(3, 64, 128, 3, 16, 16), (2, 1, 128, 3, 32, 32), (2, 2, 64, 9, 16, 16), (4, 256, 512, 3, 8, 8), (2, 32, 64, 3, 13, 9),
                                   (2, 3, 64, 3, 16, 16), (1, 2, 32, 3, 6, 10)])
def test_res_block_down_fused(shape):
    """The block-level fused ResBlockDown (bf16 path: epilogue-fused LeakyReLU / masks / residual add, multi-output pooling) against
    the float64 statement of model_blocks.py:486-514, and against the composition of elementary operators it replaces."""
    ops = ops_mod()
    ops.set_precision("bf16")
    n, ci, co, k, h, w = shape
    x64 = rnd(n, ci, h, w, seed=1).requires_grad_()
    ws = [(rnd(co, ci, 1, 1, seed=2) / np.sqrt(ci)), (rnd(co, ci, k, k, seed=3) / np.sqrt(ci * k * k)), (rnd(co, co, k, k, seed=4) / np.sqrt(co * k * k))]
    bs = [rnd(co, seed=5 + i, scale=0.1) for i in range(3)]
    for t in ws + bs:
        t.requires_grad_()
    pad = (k - 1) // 2
    res = F.conv2d(x64, ws[0], bs[0])
    out = F.conv2d(F.leaky_relu(F.conv2d(F.leaky_relu(x64, 0.2), ws[1], bs[1], padding=pad), 0.2), ws[2], bs[2], padding=pad)
    y64 = F.avg_pool2d(res, 2) + F.avg_pool2d(out, 2)
    probe = rnd(*y64.shape, seed=9)
    ref = torch.autograd.grad((y64 * probe).sum(), [x64] + ws + bs)
    pr = probe.permute(0, 2, 3, 1).contiguous().to("cuda", torch.float32)

    def run(fused):
        x = to_dev_nhwc(x64.detach(), torch.float32)
        wp = [pack_w(t.detach()) for t in ws]
        bp = [t.detach().to("cuda", torch.float32).requires_grad_() for t in bs]
        if fused:
            y = ops.res_block_down(x, wp[0], bp[0], wp[1], bp[1], wp[2], bp[2], k, 0.2, True)
            assert y.tb.dtype == torch.bfloat16 and torch.equal(y.tb, y.t32.bfloat16())
            assert torch.equal(y.tl, F.leaky_relu(y.t32, 0.2).bfloat16())
            y = y.t32
        else:
            r = ops.conv2d(x, wp[0], bp[0], 1)
            o = ops.conv2d(ops.conv2d(x, wp[1], bp[1], k, ops.PRE_LRELU, 0.2), wp[2], bp[2], k, ops.PRE_LRELU, 0.2)
            y = ops.avg_pool2_add(r, o)
        g = torch.autograd.grad(ops.DotFn.apply(y, pr).sum(), [x] + wp + bp)
        return y, g

    yf, gf = run(True)
    yc, gc = run(False)
    assert rel_err(nchw(yf), y64) < BF16_TOL
    # same arithmetic except that the fused residual branch rounds AvgPool(x) instead of x to bf16 (the 1x1 conv runs after the pooling)
    assert rel_err(yf, yc) < 3e-3
    names = ["x", "w_l1", "w_r1", "w_r2", "b_l1", "b_r1", "b_r2"]
    for i, nm in enumerate(names):
        a, c, r = gf[i], gc[i], ref[i]
        if nm == "x":
            a, c = nchw(a), nchw(c)
        elif nm.startswith("w"):
            kk = 1 if nm == "w_l1" else k
            a, c = unpack_w(a, kk), unpack_w(c, kk)
        assert rel_err(a, c) < 5e-3, nm                   # fused vs composite differ only by where bf16 roundings of gradients fall
        # gradients pass through two LeakyReLU masks decided on bf16-rounded activations: both bf16 evaluations sit ~1-2 % from float64
        assert rel_err(a, r) < 2.5 * BF16_TOL, nm


def test_matmul_tensor_core_path():
    """bf16 path: the strided batched GEMM on mma.sync with hi+lo split operands (fp32-grade) for every transpose combination and
    ragged tile edges, including the attention shapes of the O / V nets."""
    ops = ops_mod()
    ops.set_precision("bf16")
    r16 = lambda t: t.float().double()
    for (bt, m, n, k) in ((3, 70, 33, 17), (2, 64, 64, 32), (5, 64, 256, 64), (2, 256, 16, 256), (1, 130, 100, 200), (1, 10, 40, 5000), (2, 20, 10, 3333)):
        for ta in (False, True):
            for tb in (False, True):
                a64 = rnd(bt, *((k, m) if ta else (m, k)), seed=1)
                b64 = rnd(bt, *((n, k) if tb else (k, n)), seed=2)
                ref = torch.matmul(r16(a64).transpose(1, 2) if ta else r16(a64), r16(b64).transpose(1, 2) if tb else r16(b64))
                a, b = a64.float().cuda().requires_grad_(), b64.float().cuda().requires_grad_()
                got = ops.matmul(a, b, ta, tb)
                assert rel_err(got, ref) < 3e-5, (bt, m, n, k, ta, tb)
                probe = rnd(*ref.shape, seed=3)
                gr = torch.autograd.grad((torch.matmul(a64.requires_grad_().transpose(1, 2) if ta else a64.requires_grad_(),
                                                       b64.requires_grad_().transpose(1, 2) if tb else b64.requires_grad_()) * probe).sum(), (a64, b64))
                gg = torch.autograd.grad((got * probe.float().cuda()).sum(), (a, b))
                assert rel_err(gg[0], gr[0]) < 3e-5 and rel_err(gg[1], gr[1]) < 3e-5


def test_full_size_conv_kernels():
    """The tcgen05 kernels at the O workload's full per-pass size (640 images = B 128 x 5 samples): forward (incl. the cta_group::2 variant
    and the fused LeakyReLU / mask epilogues) and weight gradient on the hot layer shapes, each checked on an image subset against torch
    with the same bf16 operands (tools/conv_bench.py asserts rel < 2e-3 fp32-out / 1e-2 bf16-out)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "conv_bench.py"), "--reps", "1", "--shapes", "0,1,2,6,8"], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("shape n=640") == 5 and "rel" in res.stdout


FULL_SHAPES = [  # n, ci, co, k, h, w -- enough tiles that every SM pair of the persistent kernel walks several rounds, ragged last tiles
    (640, 128, 128, 3, 32, 32),      # halo staging, pair mode, two pixel tiles per CTA (two accumulation chains)
    (333, 256, 256, 3, 16, 16),      # odd image count: zero-filled tail tiles
    (321, 512, 512, 3, 8, 8),        # two 8x8 images per tile ((h, image, w) pixel order), odd count
    (61, 64, 64, 3, 64, 64),         # 64-wide N tile, single-CTA variant
    (200, 128, 256, 3, 13, 13),      # odd map: ragged 8-pixel-wide tiles
    (150, 64, 128, 3, 26, 26),
    (640, 512, 512, 3, 4, 4),        # below 8x8: per-tap boxes (no halo), 256-wide tile
    (640, 256, 128, 3, 16, 16),      # weight gradient through the shared-dY tap-group kernel (cout <= 128)
]


@pytest.mark.parametrize("shape", FULL_SHAPES)
def test_conv_tcgen05_full_size_every_image(shape):
    """Round-2 kernels at full size, EVERY image checked (the benchmark-sized launches walk many tiles per persistent CTA; a wrong tile
    coordinate, halo offset or ring phase would corrupt only some of them): forward against cuDNN on the same bf16 operands, per image;
    weight gradient against the CUDA-core kernel."""
    ops = ops_mod()
    from optimalstrategiesagainstgenerativeattacks_b200 import _cabi as C
    ops.set_precision("bf16")
    ops.set_conv_algo("tcgen05")
    n, ci, co, k, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((n, h, w, ci), device="cuda", generator=g).to(torch.bfloat16)
    wp = (torch.randn((k * k, co, ci), device="cuda", generator=g) / np.sqrt(ci * k * k)).to(torch.bfloat16)
    b = torch.randn((co,), device="cuda", generator=g)
    y = torch.empty((n, h, w, co), device="cuda", dtype=torch.float32)
    C.call("gim_conv2d_fwd", C.ptr(x), C.ptr(wp), C.ptr(b), C.ptr(y), n, h, w, ci, co, k, C.BF16, C.F32, C.ALGO_TCGEN05)
    torch.cuda.synchronize()
    w_oihw = wp.float().reshape(k, k, co, ci).permute(2, 3, 0, 1).contiguous()
    worst = 0.0
    for i0 in range(0, n, 64):                       # fp32 reference in slabs (no TF32)
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            ref = torch.nn.functional.conv2d(x[i0:i0 + 64].float().permute(0, 3, 1, 2), w_oihw, b, padding=(k - 1) // 2).permute(0, 2, 3, 1)
        finally:
            torch.backends.cudnn.allow_tf32 = old
        d = (y[i0:i0 + 64] - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)
        worst = max(worst, float(d.max()))
    assert worst < 1e-4, worst                       # same operands, fp32 accumulation on both sides: only the summation order differs
    gy = torch.randn((n, h, w, co), device="cuda", generator=g).to(torch.bfloat16)
    gw = {}
    for algo, code in (("tc", C.ALGO_TCGEN05), ("simt", C.ALGO_SIMT)):
        out = torch.empty((k * k, co, ci), device="cuda", dtype=torch.float32)
        C.call("gim_conv2d_wgrad", C.ptr(x), C.ptr(gy), C.ptr(out), n, h, w, ci, co, k, C.BF16, code)
        gw[algo] = out
    torch.cuda.synchronize()
    assert rel_err(gw["tc"], gw["simt"]) < 1e-4


@pytest.mark.parametrize("channels,n_img", [(128, 3), (256, 5), (256, 160)])
def test_fused_attention_core(channels, n_img):
    """gim_attention_fwd / _bwd (one CTA per image, 64 positions) against float64 torch and against the composed gemm + softmax route."""
    ops = ops_mod()
    p, d = 64, channels // 8
    q64, k64 = (rnd(n_img, p, d, seed=s, scale=0.7).requires_grad_() for s in (1, 2))
    v64, x64 = (rnd(n_img, p, channels, seed=s).requires_grad_() for s in (3, 4))
    gm64 = torch.tensor([0.37], dtype=torch.float64, requires_grad=True)
    probe = rnd(n_img, p, channels, seed=5)

    def reference(q, k, v, x, gm):
        return gm * torch.matmul(torch.softmax(torch.matmul(q, k.transpose(1, 2)), -1), v) + x

    ref = reference(q64, k64, v64, x64, gm64)
    gref = torch.autograd.grad((ref * probe).sum(), (q64, k64, v64, x64, gm64))
    leaves = [t.detach().float().cuda().requires_grad_() for t in (q64, k64, v64, x64, gm64)]
    assert ops.attention_fused_ok(p, channels, *leaves[:4])
    got = ops.AttentionCoreFn.apply(*leaves)
    assert rel_err(got, ref) < 1e-5
    ggot = torch.autograd.grad((got * probe.float().cuda()).sum(), leaves)
    for a_, b_ in zip(ggot, gref):
        assert rel_err(a_, b_) < 2e-5
    with ops.composite_mode():                  # second-order graphs take the composed route
        assert not ops.attention_fused_ok(p, channels, *leaves[:4])


def test_self_attention_module_routes_agree():
    """SelfAttention at the O-config shape (8x8 map, 256 channels): fused route == composed route (fp32), gradients included."""
    from optimalstrategiesagainstgenerativeattacks_b200 import model_blocks as mb
    ops = ops_mod()
    torch.manual_seed(0)
    att = mb.SelfAttention(256).cuda().eval()
    with torch.no_grad():
        att.gamma.fill_(0.5)
    x = torch.randn(4, 8, 8, 256, device="cuda").requires_grad_()
    probe = torch.randn(4, 8, 8, 256, device="cuda")
    params = [x, att.gamma, att.conv_f.weight_orig, att.conv_h.weight_orig]

    def run():
        y = att(x)
        return y, torch.autograd.grad((y * probe).sum(), params)

    y_f, g_f = run()
    with ops.composite_mode():
        y_c, g_c = run()
    assert rel_err(y_f, y_c) < 1e-5
    for a_, b_ in zip(g_f, g_c):
        assert rel_err(a_, b_) < 1e-4


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", BF16_TOL)])
def test_encoder_merged_attention_projections(precision, tol):
    """Full-size O-config encoder (1x32x32 -> 512; attention on the 8x8 x 256 map): the default route -- batched spectral norm, the three
    attention projections as one merged 1x1 convolution, fused attention kernels, direct .grad accumulation -- against the composed
    route (`composite_mode`: three convolutions, gemm + softmax), outputs and parameter gradients."""
    from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as gm
    ops = ops_mod()
    ops.set_precision(precision)
    torch.manual_seed(1)
    enc = gm.Encoder(img_size=32, img_channels=1, style_dim=512).cuda().train()
    with torch.no_grad():
        enc.att.gamma.fill_(0.7)
        warm = torch.randn(4, 1, 32, 32, device="cuda")
        for _ in range(6):                      # let the power iterations settle, then freeze u / v (eval)
            enc(warm)
    enc.eval()
    x = torch.randn(6, 1, 32, 32, device="cuda")
    probe = torch.randn(6, 512, device="cuda")
    names = ["att.gamma", "att.conv_f.weight_orig", "att.conv_g.bias", "att.conv_h.weight_orig", "att.conv_h.bias", "down_blocks.1.conv_r1.weight_orig",
             "down_blocks.2.conv_r2.bias"]
    params = dict(enc.named_parameters())

    def run(composed):
        for p_ in enc.parameters():
            p_.grad = torch.zeros_like(p_)
        with (ops.composite_mode() if composed else torch.enable_grad()):
            y = enc(x)
            with ops.deferred_weight_grads():
                (y * probe).sum().backward()
        return y.detach().clone(), [params[n].grad.detach().clone() for n in names]

    y_f, g_f = run(False)
    y_c, g_c = run(True)
    assert rel_err(y_f, y_c) < tol
    for name, a_, b_ in zip(names, g_f, g_c):
        assert float(b_.abs().max()) > 0, name
        assert rel_err(a_, b_) < (tol if precision == "fp32" else 2.5 * tol), name


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,ci,co,k,h,w,upsample,with_add", [(3, 64, 64, 3, 8, 8, True, False), (2, 128, 64, 3, 4, 4, False, True), (5, 512, 512, 3, 1, 1, True, False),
                                                              (2, 64, 32, 9, 8, 8, False, True), (3, 32, 48, 3, 6, 10, False, False), (2, 64, 128, 3, 2, 2, True, False)])
def test_norm_conv_fused(mode, n, ci, co, k, h, w, upsample, with_add):
    """The fused norm -> LeakyReLU -> (nearest x2) -> conv (+ upsampled half-resolution residual) node of the attacker's blocks against the float64
    statement (reference model_blocks.py:760-768, 805-811, 851-861) and against the composition of elementary operators it replaces."""
    ops = ops_mod()
    ops.set_precision("bf16")
    if h * w == 1 and mode == 1:
        pytest.skip("ada_in needs more than one pixel")
    x64 = rnd(n, ci, h, w, seed=1).requires_grad_()
    w64 = (rnd(co, ci, k, k, seed=2) / np.sqrt(ci * k * k)).requires_grad_()
    b64 = rnd(co, seed=3, scale=0.1).requires_grad_()
    if mode == 0:
        sc64, sh64 = (1.0 + 0.1 * rnd(ci, seed=4)).requires_grad_(), rnd(ci, seed=5, scale=0.1).requires_grad_()
    else:
        sc64, sh64 = (1.0 + 0.1 * rnd(n, ci, seed=4)).requires_grad_(), rnd(n, ci, seed=5, scale=0.1).requires_grad_()
    oh, ow = (2 * h, 2 * w) if upsample else (h, w)
    add64 = rnd(n, co, oh // 2, ow // 2, seed=6).requires_grad_() if with_add else None
    # float64 statement
    mu = x64.mean(dim=(2, 3), keepdim=True)
    if mode == 0:
        xn = (x64 - mu) / torch.sqrt(((x64 - mu) ** 2).mean(dim=(2, 3), keepdim=True) + 1e-5) * sc64.view(1, -1, 1, 1) + sh64.view(1, -1, 1, 1)
    else:
        sd = torch.sqrt(((x64 - mu) ** 2).sum(dim=(2, 3), keepdim=True) / (h * w - 1)) + 1e-5
        xn = sc64.view(n, ci, 1, 1) * (x64 - mu) / sd + sh64.view(n, ci, 1, 1)
    t = F.leaky_relu(xn, 0.2)
    if upsample:
        t = F.interpolate(t, scale_factor=2, mode="nearest")
    y64 = F.conv2d(t, w64, b64, padding=(k - 1) // 2)
    if with_add:
        y64 = y64 + F.interpolate(add64, scale_factor=2, mode="nearest")
    probe = rnd(*y64.shape, seed=9)
    wrt = [x64, sc64, sh64, w64, b64] + ([add64] if with_add else [])
    ref = torch.autograd.grad((y64 * probe).sum(), wrt)
    pr = probe.permute(0, 2, 3, 1).contiguous().to("cuda", torch.float32)

    def run(fused):
        x = to_dev_nhwc(x64.detach(), torch.float32)
        wp = pack_w(w64.detach())
        b = b64.detach().to("cuda", torch.float32).requires_grad_()
        sc, sh = (v.detach().to("cuda", torch.float32).requires_grad_() for v in (sc64, sh64))
        add = to_dev_nhwc(add64.detach(), torch.float32) if with_add else None
        if fused:
            assert ops.norm_conv_ok(x, wp)
            y = ops.norm_conv(x, sc, sh, wp, b, k, mode, 1e-5, 0.2, upsample=upsample, addend=add)
        else:
            y = ops.instance_norm(x, sc, sh, 1e-5, 0.2) if mode == 0 else ops.ada_in(x, sh, sc, 1e-5, 0.2)
            y = ops.conv2d(y, wp, b, k, ops.PRE_UPSAMPLE if upsample else ops.PRE_NONE, 0.2)
            if with_add:
                y = ops.AddFn.apply(y, ops.upsample2(add))
        g = torch.autograd.grad(ops.DotFn.apply(y, pr).sum(), [x, sc, sh, wp, b] + ([add] if with_add else []))
        return y, g

    yf, gf = run(True)
    yc, gc = run(False)
    assert rel_err(nchw(yf), y64) < BF16_TOL
    assert rel_err(yf, yc) < 1e-5                      # identical operands, fp32 accumulation in both
    for i, nm in enumerate(["x", "scale", "shift", "w", "b"] + (["addend"] if with_add else [])):
        a, c, r = gf[i], gc[i], ref[i]
        if nm == "w":
            a, c = unpack_w(a, k), unpack_w(c, k)
        elif a.dim() == 4:
            a, c = nchw(a), nchw(c)
        if float(r.norm()) < 1e-9:                     # (the bias in front of nothing normalising it is fine; degenerate 1x1 cases give zeros)
            continue
        assert rel_err(a, c) < 5e-3, (nm, rel_err(a, c))
        assert rel_err(a, r) < (5e-2 if h * w <= 4 else BF16_TOL), (nm, rel_err(a, r))
