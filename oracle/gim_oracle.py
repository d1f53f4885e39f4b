"""CPU oracle for the GIM hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU fp32 *restatement* of the reference's algorithm
for the path named in BASELINE.json `north_star`.  It is the checker the parity
tests compare the CUDA path against.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import it; nothing in
`optimalstrategiesagainstgenerativeattacks_b200/` does (the product path fails
loudly when the CUDA library is missing instead of falling back here).

Parity pin: the reference ships no tests or golden vectors, so
the restatement is pinned against outputs of the reference itself:
`oracle/make_golden.py` imports the unmodified reference from /root/reference
(through `oracle/ref_shim.py`), runs it on seeded inputs and writes
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below
against those vectors.

Style: purely functional over a flat ``{state_dict key: tensor}`` dict (the
reference's checkpoint schema), NCHW fp32, autograd by torch.
All `file:line` citations are relative to /root/reference.
"""
import math
import random

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2          # nn.LeakyReLU(0.2): models/model_blocks.py:489, 744; gim_img_models.py:36
SN_EPS = 1e-12             # torch.nn.utils.spectral_norm default eps (model_blocks.py:492-495)
IN_EPS = 1e-5              # nn.InstanceNorm2d default eps (model_blocks.py:747-748)
ADAIN_EPS = 1e-5           # ada_in(..., eps=1e-5) model_blocks.py:611
STD_EPS = 1e-8             # custom_std model_blocks.py:45


def lrelu(x):
    return torch.where(x > 0, x, x * LRELU_SLOPE)


# ---------------------------------------------------------------------------------------------
# Optional emulation of the build's bf16 tensor-core numerics (NOT part of the reference): every conv / Linear operand --
# activation, weight and the incoming gradient in the backward -- is rounded to bfloat16 once, products accumulate in fp32,
# everything else stays fp32.  tests/test_parity_gpu.py uses it to separate "implementation error" (CUDA path vs this
# emulation: tight) from "arithmetic error" (this emulation vs the float64 reference: what bf16 operands cost).
# ---------------------------------------------------------------------------------------------
OPERAND_BF16 = False


def set_operand_rounding(flag):
    global OPERAND_BF16
    OPERAND_BF16 = bool(flag)


def _r(t):
    return t.bfloat16().to(t.dtype)


class _ConvBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, padding):
        xr, wr = _r(x), _r(w)
        ctx.save_for_backward(xr, wr)
        ctx.padding = padding
        ctx.has_bias = b is not None
        return F.conv2d(xr, wr, b, stride=1, padding=padding)

    @staticmethod
    def backward(ctx, gy):
        xr, wr = ctx.saved_tensors
        gb = gy.sum(dim=(0, 2, 3)) if ctx.has_bias else None
        gyr = _r(gy)
        gx = torch.nn.grad.conv2d_input(xr.shape, wr, gyr, stride=1, padding=ctx.padding)
        gw = torch.nn.grad.conv2d_weight(xr, wr.shape, gyr, stride=1, padding=ctx.padding)
        return gx, gw, gb, None


class _LinearBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        xr, wr = _r(x), _r(w)
        ctx.save_for_backward(xr, wr)
        return F.linear(xr, wr, b)

    @staticmethod
    def backward(ctx, gy):
        xr, wr = ctx.saved_tensors
        gyr = _r(gy)
        g2 = gyr.reshape(-1, gyr.shape[-1])
        return gyr @ wr, g2.t() @ xr.reshape(-1, xr.shape[-1]), gy.reshape(-1, gy.shape[-1]).sum(0)


def conv2d_op(x, w, b, padding):
    if OPERAND_BF16:
        return _ConvBF16.apply(x, w, b, padding)
    return F.conv2d(x, w, b, stride=1, padding=padding)


# ---------------------------------------------------------------------------------------------
# spectral norm (torch.nn.utils.spectral_norm, old-style hook; call sites model_blocks.py:492-495,
# 522-526, 750-751, 792-793, 836-840).  One power iteration per forward call in train mode,
# in place on weight_u / weight_v; sigma = u^T W v with u, v treated as constants.
# ---------------------------------------------------------------------------------------------
def sn_weight(p, prefix, training=True):
    w = p[prefix + ".weight_orig"]
    u = p[prefix + ".weight_u"]
    v = p[prefix + ".weight_v"]
    w_mat = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v_new = F.normalize(torch.mv(w_mat.t(), u), dim=0, eps=SN_EPS)
            u_new = F.normalize(torch.mv(w_mat, v_new), dim=0, eps=SN_EPS)
            u.copy_(u_new)
            v.copy_(v_new)
    u_c = u.detach().clone()
    v_c = v.detach().clone()
    sigma = torch.dot(u_c, torch.mv(w_mat, v_c))
    return w / sigma


def sn_conv(p, prefix, x, padding, training=True):
    w = sn_weight(p, prefix, training)
    return conv2d_op(x, w, p[prefix + ".bias"], padding)


def avg_pool2(x):
    return F.avg_pool2d(x, 2)                                # nn.AvgPool2d(2) model_blocks.py:490


def upsample2(x):
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)   # nn.Upsample(scale_factor=2) nearest, :740


def instance_norm(x, weight, bias):
    """nn.InstanceNorm2d(affine=True), torch-1.2 semantics incl. the 1x1 map: biased variance."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(2, 3), keepdim=True)
    y = (x - mean) / torch.sqrt(var + IN_EPS)
    return y * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


def ada_in(x, mean_style, std_style):
    """model_blocks.py:611-630: unbiased std, eps added to the std."""
    b, c, h, w = x.shape
    f = x.reshape(b, c, -1)
    n = f.shape[2]
    mean = f.mean(dim=2, keepdim=True)
    var = ((f - mean) ** 2).sum(dim=2, keepdim=True) / (n - 1)
    std = torch.sqrt(var) + ADAIN_EPS
    out = std_style.view(b, c, 1) * (f - mean) / std + mean_style.view(b, c, 1)
    return out.reshape(b, c, h, w)


def linear(p, prefix, x):
    w, b = p[prefix + ".weight"], p[prefix + ".bias"]
    rows = x.numel() // x.shape[-1]
    if OPERAND_BF16 and w.shape[1] % 8 == 0 and w.shape[0] % 16 == 0 and rows >= 16:     # the layers the build runs on tensor cores
        return _LinearBF16.apply(x, w, b)
    return F.linear(x, w, b)


def mlp(p, prefix, x, n_layers):
    """mb.MLP model_blocks.py:77-94: Linear, LeakyReLU(0.2), ..., Linear at indices 0,2,4,..."""
    for i in range(n_layers):
        x = linear(p, "%s.model.%d" % (prefix, 2 * i), x)
        if i < n_layers - 1:
            x = lrelu(x)
    return x


# ---------------------------------------------------------------------------------------------
# blocks (models/model_blocks.py)
# ---------------------------------------------------------------------------------------------
def res_block_down(p, prefix, x, ksize=3, training=True):
    """ResBlockDown model_blocks.py:497-514."""
    pad = (ksize - 1) // 2
    if OPERAND_BF16 and x.shape[-1] % 2 == 0 and x.shape[-2] % 2 == 0:
        # emulation of the build's bf16 numerics only: its fused block runs the 1x1 residual conv AFTER the pooling (the same value in
        # exact arithmetic: a 1x1 convolution commutes with AvgPool), so the bf16 operand rounding falls on AvgPool(x) instead of x
        left = sn_conv(p, prefix + ".conv_l1", avg_pool2(x), 0, training)
    else:
        left = avg_pool2(sn_conv(p, prefix + ".conv_l1", x, 0, training))
    out = sn_conv(p, prefix + ".conv_r1", lrelu(x), pad, training)
    out = sn_conv(p, prefix + ".conv_r2", lrelu(out), pad, training)
    return left + avg_pool2(out)


def self_attention(p, prefix, x, training=True):
    """SelfAttention model_blocks.py:530-549 (softmax over dim -2: columns sum to one)."""
    b, c, h, w = x.shape
    f = sn_conv(p, prefix + ".conv_f", x, 0, training).reshape(b, -1, h * w).transpose(1, 2)
    g = sn_conv(p, prefix + ".conv_g", x, 0, training).reshape(b, -1, h * w)
    hp = sn_conv(p, prefix + ".conv_h", x, 0, training).reshape(b, -1, h * w)
    att = torch.softmax(torch.bmm(f, g), dim=-2)
    out = torch.bmm(hp, att).reshape(b, c, h, w)
    return p[prefix + ".gamma"] * out + x


def res_block_up(p, prefix, x, training=True):
    """ResBlockUp model_blocks.py:753-773."""
    left = sn_conv(p, prefix + ".conv_l1", upsample2(x), 0, training)
    out = instance_norm(x, p[prefix + ".in1.weight"], p[prefix + ".in1.bias"])
    out = sn_conv(p, prefix + ".conv_r1", upsample2(lrelu(out)), 1, training)
    out = instance_norm(out, p[prefix + ".in2.weight"], p[prefix + ".in2.bias"])
    out = sn_conv(p, prefix + ".conv_r2", lrelu(out), 1, training)
    return out + left


def ada_res_block2(p, prefix, x, style, training=True):
    """AdaResBlock2 model_blocks.py:795-814."""
    m1 = linear(p, prefix + ".lin1_mean", style)
    s1 = linear(p, prefix + ".lin1_std", style)
    m2 = linear(p, prefix + ".lin2_mean", style)
    s2 = linear(p, prefix + ".lin2_std", style)
    out = sn_conv(p, prefix + ".conv1", x, 1, training)
    out = lrelu(ada_in(out, m1, s1))
    out = sn_conv(p, prefix + ".conv2", out, 1, training)
    return ada_in(out, m2, s2) + x


def ada_res_block_up2(p, prefix, x, style, ksize=3, training=True):
    """AdaResBlockUp2 model_blocks.py:842-865."""
    pad = (ksize - 1) // 2
    m1 = linear(p, prefix + ".lin1_mean", style)
    s1 = linear(p, prefix + ".lin1_std", style)
    m2 = linear(p, prefix + ".lin2_mean", style)
    s2 = linear(p, prefix + ".lin2_std", style)
    left = sn_conv(p, prefix + ".conv_l1", upsample2(x), 0, training)
    out = upsample2(lrelu(ada_in(x, m1, s1)))
    out = sn_conv(p, prefix + ".conv_r1", out, pad, training)
    out = lrelu(ada_in(out, m2, s2))
    out = sn_conv(p, prefix + ".conv_r2", out, pad, training)
    return out + left


# ---------------------------------------------------------------------------------------------
# image networks (models/gim_img_models.py)
# ---------------------------------------------------------------------------------------------
def n_down_blocks(img_size):
    return int(math.log2(img_size)) - 2                                 # gim_img_models.py:30, 111, 175


# Test-only hook (NOT part of the reference): callable(prefix, x[N,C,H,W]) -> [N,C] used instead of the global max, so that a test can
# record one evaluation's arg-max routing and replay it in another (tests/test_parity_width_gpu.py isolates what arg-max flips cost).
GMAX_HOOK = None


def encoder(p, prefix, x, img_size, training=True):
    """Encoder gim_img_models.py:43-57 -> [N, style_dim]."""
    nb = n_down_blocks(img_size)
    att_loc = int(math.ceil(nb / 2))
    for i in range(nb):
        if i == att_loc:
            x = self_attention(p, prefix + ".att", x, training)
        x = res_block_down(p, "%s.down_blocks.%d" % (prefix, i), x, 3, training)
    x = torch.amax(x, dim=(2, 3)) if GMAX_HOOK is None else GMAX_HOOK(prefix, x)
    return lrelu(x)


def env_decoder(p, prefix, x, img_size, training=True):
    """EnvDecoder gim_img_models.py:85-95: [N, style_dim] -> [N, C, S, S], no output nonlinearity."""
    nb = int(math.log2(img_size))
    att_loc = int(math.ceil(nb / 2))
    x = x.reshape(x.shape[0], x.shape[1], 1, 1)
    for i in range(nb):
        if i == att_loc:
            x = self_attention(p, prefix + ".att", x, training)
        x = res_block_up(p, "%s.up_blocks.%d" % (prefix, i), x, training)
    return x


def img2img(p, prefix, x, style, img_size, n_res=5, training=True):
    """AdaInImage2Image gim_img_models.py:248-257 = down (129-139), 5 AdaIN res (160-162), AdaIN up + tanh (211-215)."""
    nb = n_down_blocks(img_size)
    att_loc = int(math.ceil(nb / 2))
    d = prefix + ".down_block"
    for i in range(nb):
        if i == att_loc:
            x = self_attention(p, d + ".att", x, training)
        x = res_block_down(p, "%s.down_blocks.%d" % (d, i), x, 9 if i == 0 else 3, training)
        x = instance_norm(x, p["%s.in_layers.%d.weight" % (d, i)], p["%s.in_layers.%d.bias" % (d, i)])
    for i in range(n_res):
        x = ada_res_block2(p, "%s.adain_res_block.res_blocks.%d" % (prefix, i), x, style, training)
    u = prefix + ".adain_up_block"
    for i in range(nb):
        if i == att_loc:
            x = self_attention(p, u + ".att", x, training)
        x = ada_res_block_up2(p, "%s.up_blocks.%d" % (u, i), x, style, 9 if i == nb - 1 else 3, training)
    return torch.tanh(x)


def custom_std(x):
    """model_blocks.py:41-48."""
    if x.shape[1] > 1:
        return torch.sqrt(x.var(1) + STD_EPS)
    return torch.zeros((x.shape[0],) + tuple(x.shape[2:]), dtype=x.dtype)


def mean_std_fc_stat(p, prefix, x):
    """GIMMeanStdFcStat gim_basic_models.py:163-172 with GIMFCStat :125-127 (MLP of 4 Linear layers)."""
    fc = mlp(p, prefix + ".fc.stat", x, 4).mean(1)
    return torch.cat((x.mean(1), custom_std(x), fc), dim=-1)


def face_dis(p, prefix, test_src, test_env, si_src, si_env):
    """GIMFaceDis gim_img_models.py:279-299."""
    x = torch.cat((test_src.mean(1), si_src.mean(1),
                   mean_std_fc_stat(p, prefix + ".stat", test_env),
                   mean_std_fc_stat(p, prefix + ".stat", si_env)), dim=-1)
    return mlp(p, prefix + ".mlp", x, 3)


def encode_sample(p, prefix, sample, training=True):
    """src/env_encode_sample gim_img_models.py:328-340."""
    b, s = sample.shape[:2]
    x = encoder(p, prefix, sample.reshape((b * s,) + tuple(sample.shape[2:])), sample.shape[-1], training)
    return x.reshape(b, s, -1)


def authenticator(p, test_sample, si_sample, training=True):
    """GIMFaceAuthenticator.forward gim_img_models.py:313-326 (call order: src test, src si, env test, env si)."""
    test_src = encode_sample(p, "src_encoder", test_sample, training)
    si_src = encode_sample(p, "src_encoder", si_sample, training)
    test_env = encode_sample(p, "env_encoder", test_sample, training)
    si_env = encode_sample(p, "env_encoder", si_sample, training)
    return face_dis(p, "dis", test_src, test_env, si_src, si_env)


def impersonator(p, leaked, n, z, remove_noise_mean=True, n_noise_layers=4, training=True):
    """GIMFaceImpersonator.forward gim_img_models.py:364-397 with the noise z passed in (use_img_att=False)."""
    b, m, c, s, _ = leaked.shape
    expanded = leaked[:, 0].unsqueeze(1).expand(-1, n, -1, -1, -1)
    src = encode_sample(p, "src_encoder", leaked, training).mean(1)
    env = encode_sample(p, "env_encoder", leaked, training).mean(1)
    w = mlp(p, "env_noise_mapper", z, n_noise_layers)
    if remove_noise_mean:
        w = w - w.mean(1, keepdim=True)
    noisy_env = env.unsqueeze(1) + w
    env_img = env_decoder(p, "env_decoder", noisy_env.reshape(b * n, -1), s, training).reshape(b, n, c, s, s)
    x = torch.cat((env_img, expanded), dim=2).reshape(b * n, 2 * c, s, s)
    style = src.unsqueeze(1).expand(-1, n, -1).reshape(b * n, -1)
    return img2img(p, "img2img", x, style, s, 5, training).reshape(b, n, c, s, s)


# ---------------------------------------------------------------------------------------------
# Gaussian networks (models/gim_gaussian_models.py)
# ---------------------------------------------------------------------------------------------
def mean_std_stat(x):
    """GIMMeanStdStat gim_basic_models.py:81-89."""
    return torch.cat((x.mean(1), custom_std(x)), dim=-1)


def gaussian_authenticator(p, test_sample, si_sample):
    """GIMGaussianDis gim_gaussian_models.py:31-41: MLP(4d -> d -> 2d -> 1)."""
    x = torch.cat((mean_std_stat(test_sample), mean_std_stat(si_sample)), dim=-1)
    return mlp(p, "dis.mlp", x, 3)


def gaussian_impersonator(p, leaked, n, z, remove_noise_mean=True):
    """GIMGaussianImpersonator gim_gaussian_models.py:75-89 (out_mlp is never used)."""
    src = leaked.mean(1)
    w = mlp(p, "env_noise_mapper", z, 1)
    if remove_noise_mean:
        w = w - w.mean(1, keepdim=True)
    return w + src.unsqueeze(1)


# ---------------------------------------------------------------------------------------------
# trainer math (training/gim_img_trainer.py, training/gim_gaussian_trainer.py, training/utils.py)
# ---------------------------------------------------------------------------------------------
def gan_loss(dis_out, target):
    """gan_loss gim_img_trainer.py:90-94: per-sample BCE-with-logits, squeezed."""
    t = torch.full_like(dis_out, target)
    return F.binary_cross_entropy_with_logits(dis_out, t, reduction="none").squeeze()


def compute_grad2(out, x_in):
    """compute_grad2 training/utils.py:115-124 (R1: per-episode squared input-gradient norm)."""
    b = x_in[0].shape[0]
    grads = torch.autograd.grad(out.sum(), x_in, create_graph=True, retain_graph=True)
    return sum(g.pow(2).reshape(b, -1).sum(1) for g in grads)


def img_authenticator_forward(p, fake, real, si, reg_param, training=True):
    """GIMImgTrainer.authenticator_forward gim_img_trainer.py:96-142 (encode order: si, real, fake)."""
    if reg_param > 0:
        real.requires_grad_()
        si.requires_grad_()
    si_src = encode_sample(p, "src_encoder", si, training)
    si_env = encode_sample(p, "env_encoder", si, training)
    real_src = encode_sample(p, "src_encoder", real, training)
    real_env = encode_sample(p, "env_encoder", real, training)
    fake_src = encode_sample(p, "src_encoder", fake, training)
    fake_env = encode_sample(p, "env_encoder", fake, training)
    out_real = face_dis(p, "dis", real_src, real_env, si_src, si_env)
    loss_real = gan_loss(out_real, 1.0)
    reg = reg_param * compute_grad2(out_real, (real, si)) if reg_param > 0 else torch.zeros_like(loss_real)
    out_fake = face_dis(p, "dis", fake_src, fake_env, si_src, si_env)
    loss_fake = gan_loss(out_fake, 0.0)
    return loss_real + loss_fake + reg, loss_real, loss_fake, reg, out_real, out_fake


def gaussian_authenticator_forward(p, fake, real, si, reg_param=0.0):
    """GIMGaussianTrainer.authenticator_forward gim_gaussian_trainer.py:84-110."""
    if reg_param > 0:
        real.requires_grad_()
        si.requires_grad_()
    out_real = gaussian_authenticator(p, real, si)
    loss_real = gan_loss(out_real, 1.0)
    reg = reg_param * compute_grad2(out_real, (real, si)) if reg_param > 0 else torch.zeros_like(loss_real)
    out_fake = gaussian_authenticator(p, fake, si)
    loss_fake = gan_loss(out_fake, 0.0)
    return loss_real + loss_fake + reg, loss_real, loss_fake, reg, out_real, out_fake


def adam_step(params, grads, state, lr, beta1, beta2, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) as used at gim_img_trainer.py:50-58,
    gim_gaussian_trainer.py:48-49.  `state` = {"step": int, "m": [...], "v": [...]}; params with grad None are skipped."""
    state["step"] += 1
    t = state["step"]
    bc1 = 1.0 - beta1 ** t
    bc2 = 1.0 - beta2 ** t
    with torch.no_grad():
        for i, (w, g) in enumerate(zip(params, grads)):
            if g is None:
                continue
            state["m"][i].mul_(beta1).add_(g, alpha=1.0 - beta1)
            state["v"][i].mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
            denom = (state["v"][i].sqrt() / math.sqrt(bc2)).add_(eps)
            w.addcdiv_(state["m"][i], denom, value=-lr / bc1)


# ---------------------------------------------------------------------------------------------
# episode index algebra (data_handling/img_datasets.py:68-103, 153-187) -- integer only, bit exact
# ---------------------------------------------------------------------------------------------
def episode_indices(index, example_cnt_per_class, n_imgs_in_class, m, n, k, rng):
    """cls = index // example_cnt_per_class; idx = rng.sample(range(n_imgs), m+n+k);
    leaked = idx[:m], real = idx[m:m+n], si = idx[m+n:]  (img_datasets.py:79-103)."""
    cls = index // example_cnt_per_class
    idx = rng.sample(list(range(n_imgs_in_class)), m + n + k)
    return cls, idx[:m], idx[m:m + n], idx[m + n:]


def make_rng(seed):
    return random.Random(seed)
