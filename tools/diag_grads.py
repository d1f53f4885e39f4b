"""Gradient parity diagnosis at full width (development tool; uses the oracle = test infrastructure).
Gradients of one G-step + D-step from four evaluations on identical weights / inputs / noise:
  A ours bf16 tcgen05 (fused)   B ours bf16 CUDA cores   E oracle fp32 + bf16 operand rounding (device)   R oracle float64 (device)
optionally with the encoders' arg-max routing of R forced into all of them (--force-routing).
python tools/diag_grads.py [--weights fill|init|trained] [--workload O|V] [--batch 4] [--force-routing]"""
import argparse
import contextlib
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import optimalstrategiesagainstgenerativeattacks_b200 as gim
from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
from optimalstrategiesagainstgenerativeattacks_b200 import ops
from optimalstrategiesagainstgenerativeattacks_b200.gim_img_trainer import GIMImgTrainer
from optimalstrategiesagainstgenerativeattacks_b200.training_steps import au_train_step, im_train_step
from optimalstrategiesagainstgenerativeattacks_b200.utils import DataParallelMock
from oracle import gim_oracle as O
from oracle.fill import fill_state_dict, schema_of

ap = argparse.ArgumentParser()
ap.add_argument("--weights", default="fill")
ap.add_argument("--workload", default="O")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--force-routing", action="store_true")
ap.add_argument("--single-term", action="store_true", help="D-step loss = BCE(D(real, si) -> 1) only (no real/fake cancellation)")
ap.add_argument("--in-bias", type=float, default=None, help="with --weights init: set every InstanceNorm bias to N(0, s) (non-degenerate decoder)")
a = ap.parse_args()
size, ch, reg = (32, 1, 0.0) if a.workload == "O" else (64, 3, 10.0)
B, n = a.batch, 5
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

torch.manual_seed(1)
au0, im0 = M.get_au(size, ch, 512), M.get_im(size, ch, 512)
if a.weights == "fill":
    au0.load_state_dict(fill_state_dict(schema_of(au0), 151))
    im0.load_state_dict(fill_state_dict(schema_of(im0), 161))
else:
    with torch.no_grad():
        for net in (au0, im0):
            for n_, p in net.named_parameters():
                if n_.endswith("gamma"):
                    p.fill_(0.5)
                if a.in_bias is not None and (".in1.bias" in n_ or ".in2.bias" in n_ or ".in_layers." in n_ and n_.endswith("bias")):
                    p.normal_(0.0, a.in_bias)
sd_a = {k: v.detach().clone() for k, v in au0.state_dict().items()}
sd_i = {k: v.detach().clone() for k, v in im0.state_dict().items()}
a_names, i_names = [k for k, _ in au0.named_parameters()], [k for k, _ in im0.named_parameters()]
gen = torch.Generator().manual_seed(77)
leaked, real, si = ((torch.rand((B, 5, ch, size, size), generator=gen) * 2 - 1).to(dev) for _ in range(3))
z = torch.randn((B, n, 512), generator=gen).to(dev)
routing = {}          # (prefix, call index) -> idx [N, C] recorded from R


@contextlib.contextmanager
def inject(zz):
    real_randn = torch.randn
    torch.randn = lambda *x, **k: zz.clone()
    try:
        yield
    finally:
        torch.randn = real_randn


def ours(algo):
    gim.set_precision("bf16")
    gim.set_conv_algo(algo)
    au, im = M.get_au(size, ch, 512), M.get_im(size, ch, 512)
    au.load_state_dict(sd_a)
    im.load_state_dict(sd_i)
    au, im = au.to(dev), im.to(dev)
    counts = {}
    orig_fwd = M.Encoder.forward
    names = {id(au.src_encoder): "au.src_encoder", id(au.env_encoder): "au.env_encoder", id(im.src_encoder): "im.src_encoder", id(im.env_encoder): "im.env_encoder"}

    def forced(self, x):
        x = M._as_nhwc(x)
        from optimalstrategiesagainstgenerativeattacks_b200 import model_blocks as mb
        mb.sn_prepare_module(self, skip=() if self.att_loc < self.n_down_blocks else (self.att,))
        for i, block in enumerate(self.down_blocks):
            if i == self.att_loc:
                x = self.att(x)
            x = block(x, want_ops=i + 1 < self.n_down_blocks)
        if isinstance(x, ops.Act):
            x = x.t32
        key = names[id(self)]
        j = counts.get(key, 0)
        counts[key] = j + 1
        idx = routing[(key, j)].to(torch.int32).contiguous()
        x = ops.GatherIdxFn.apply(x, idx)
        return ops.lrelu(x)
    if a.force_routing:
        M.Encoder.forward = forced
    try:
        tr = DataParallelMock(GIMImgTrainer(tempfile.mkdtemp(), 5, 5, 5, au, im, 1e-6, 1e-6, 1e-7, reg_param=reg))
        tr.module.do_global_step()
        tr.module.update_learning_rate()
        with inject(z):
            im_loss, fake, _ = im_train_step(tr, leaked, si)
        g_im = [None if p.grad is None else p.grad.detach().double().clone() for p in im.parameters()]
        if a.single_term:
            au.zero_grad()
            out = au(real.clone(), si.clone())
            ops.BCEWithLogitsFn.apply(out, 1.0).mean().backward()
            o = (ops.BCEWithLogitsFn.apply(out, 1.0).mean().detach(),)
        else:
            o = au_train_step(tr, real.clone(), fake, si.clone())
        g_au = [None if p.grad is None else p.grad.detach().double().clone() for p in au.parameters()]
    finally:
        M.Encoder.forward = orig_fwd
    return dict(fake=fake.double(), im=g_im, au=g_au, im_loss=float(im_loss), au_loss=float(o[0]))


def oracle(dtype, rounding, record=False):
    pa = {k: v.to(dev, dtype).clone() for k, v in sd_a.items()}
    pi = {k: v.to(dev, dtype).clone() for k, v in sd_i.items()}
    for n_ in a_names:
        pa[n_].requires_grad_()
    for n_ in i_names:
        pi[n_].requires_grad_()
    counts = {}
    which = {"net": "im"}

    def hook(prefix, x):
        key = "%s.%s" % (which["net"], prefix)
        j = counts.get(key, 0)
        counts[key] = j + 1
        flat = x.flatten(2)
        if record:
            routing[(key, j)] = flat.argmax(-1)
        return flat.gather(2, routing[(key, j)].unsqueeze(-1)).squeeze(-1)
    O.set_operand_rounding(rounding)
    O.GMAX_HOOK = hook if (a.force_routing or record) else None
    try:
        # G-step: attacker forward (its own encoders), then the authenticator on (fake, si)
        class P(dict):
            pass
        which["net"] = "im"
        fake = None

        def impersonator_then_au():
            nonlocal fake
            which["net"] = "im"
            b, m = leaked.shape[:2]
            lk, zz = leaked.to(dtype), z.to(dtype)
            src = O.encode_sample(pi, "src_encoder", lk).mean(1)
            env = O.encode_sample(pi, "env_encoder", lk).mean(1)
            w = O.mlp(pi, "env_noise_mapper", zz, 4)
            w = w - w.mean(1, keepdim=True)
            noisy = env.unsqueeze(1) + w
            env_img = O.env_decoder(pi, "env_decoder", noisy.reshape(b * n, -1), size).reshape(b, n, ch, size, size)
            expanded = lk[:, 0].unsqueeze(1).expand(-1, n, -1, -1, -1)
            x = torch.cat((env_img, expanded), dim=2).reshape(b * n, 2 * ch, size, size)
            style = src.unsqueeze(1).expand(-1, n, -1).reshape(b * n, -1)
            fake = O.img2img(pi, "img2img", x, style, size, 5).reshape(b, n, ch, size, size)
            which["net"] = "au"
            return O.gan_loss(O.authenticator(pa, fake, si.to(dtype)), 1.0).mean()
        l_im = impersonator_then_au()
        l_im.backward()
        g_im = [None if pi[n_].grad is None else pi[n_].grad.double().clone() for n_ in i_names]
        for v in pa.values():
            v.grad = None
        which["net"] = "au"
        if a.single_term:
            out = (O.gan_loss(O.authenticator(pa, real.to(dtype), si.to(dtype)), 1.0),)
        else:
            out = O.img_authenticator_forward(pa, fake.detach(), real.to(dtype).clone(), si.to(dtype).clone(), reg)
        out[0].mean().backward()
        g_au = [None if pa[n_].grad is None else pa[n_].grad.double().clone() for n_ in a_names]
    finally:
        O.set_operand_rounding(False)
        O.GMAX_HOOK = None
    return dict(fake=fake.detach().double(), im=g_im, au=g_au, im_loss=float(l_im), au_loss=float(out[0].mean()))


def errs(x, y):
    top = max(float(t.norm()) for t in y if t is not None)
    out = []
    for gx, gy in zip(x, y):
        if gy is None or float(gy.norm()) < 1e-7 * top:
            out.append(np.nan)
        else:
            out.append(float((gx - gy).norm() / gy.norm()))
    return np.asarray(out)


R = oracle(torch.float64, False, record=True)
E = oracle(torch.float32, True)
F32 = oracle(torch.float32, False)
A = ours("tcgen05")
Bq = ours("simt")
print("weights=%s workload=%s batch=%d force_routing=%s in_bias=%s" % (a.weights, a.workload, B, a.force_routing, a.in_bias))
print("losses: R %.6f/%.6f  E %.6f/%.6f  A %.6f/%.6f  B %.6f/%.6f" % (R["im_loss"], R["au_loss"], E["im_loss"], E["au_loss"], A["im_loss"], A["au_loss"], Bq["im_loss"], Bq["au_loss"]))
print("fake: A-B %.2e  A-E %.2e  A-R %.2e  E-R %.2e  fp32-R %.2e" % tuple(float((x["fake"] - y["fake"]).norm() / y["fake"].norm()) for x, y in ((A, Bq), (A, E), (A, R), (E, R), (F32, R))))
for which, names in (("au", a_names), ("im", i_names)):
    for tag, x, y in (("A-B", A, Bq), ("A-E", A, E), ("B-E", Bq, E), ("A-R", A, R), ("E-R", E, R), ("fp32-R", F32, R)):
        e = errs(x[which], y[which])
        q = np.nanquantile(e, [0.5, 0.9, 1.0])
        j = int(np.nanargmax(e))
        print("%s grads %-6s median %.2e  q90 %.2e  max %.2e (%s)  frac<=2e-2 %.2f" % (which, tag, q[0], q[1], q[2], names[j], float(np.nanmean(e <= 2e-2))))
