"""Stage-by-stage parity diagnosis at full width (development tool; uses the oracle, so it is test infrastructure):
four evaluations of the same attacker forward on the same weights / inputs --
  A  ours, bf16 tensor-core path (fused blocks)      B  ours, bf16 CUDA-core path (elementary kernels)
  E  oracle with bf16 operand rounding (fp32, device) R  oracle in float64 (device)
and the relative distances between them per stage.  python tools/diag_parity.py [--weights init|fill] [--workload O|V] [--batch 4]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import optimalstrategiesagainstgenerativeattacks_b200 as gim
from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
from optimalstrategiesagainstgenerativeattacks_b200 import ops
from oracle import gim_oracle as O
from oracle.fill import fill_state_dict, schema_of

ap = argparse.ArgumentParser()
ap.add_argument("--weights", default="init")
ap.add_argument("--workload", default="O")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--gamma", type=float, default=0.5)
a = ap.parse_args()
size, ch = (32, 1) if a.workload == "O" else (64, 3)
B, n = a.batch, 5
dev = torch.device("cuda", 0)


def rel(x, y):
    x, y = x.double().flatten(), y.double().flatten()
    return float((x - y).norm() / y.norm())


torch.manual_seed(1)
im = M.get_im(size, ch, 512)
if a.weights == "fill":
    im.load_state_dict(fill_state_dict(schema_of(im), 161))
else:
    with torch.no_grad():
        for n_, p in im.named_parameters():
            if n_.endswith("gamma"):
                p.fill_(a.gamma)
sd0 = {k: v.detach().clone() for k, v in im.state_dict().items()}
gen = torch.Generator().manual_seed(77)
leaked = (torch.rand((B, 5, ch, size, size), generator=gen) * 2 - 1).to(dev)
z = torch.randn((B, n, 512), generator=gen).to(dev)


def ours(algo):
    gim.set_precision("bf16")
    gim.set_conv_algo(algo)
    net = M.get_im(size, ch, 512)
    net.load_state_dict(sd0)
    net = net.to(dev).train()
    out = {}
    with torch.no_grad():
        out["src_enc"] = net.src_encode_sample(leaked)
        out["env_enc"] = net.env_encode_sample(leaked)
        out["w"] = net.env_noise_mapper(z)
        noisy = ops.SetCenterAddFn.apply(out["w"], ops.set_mean(out["env_enc"]), True)
        out["env_img"] = net.env_decoder(noisy.view(B * n, 512))
        expanded = leaked[:, 0].unsqueeze(1).expand(-1, n, -1, -1, -1).reshape(B * n, ch, size, size)
        x = torch.cat((out["env_img"], expanded), 1)
        style = ops.set_mean(out["src_enc"]).unsqueeze(1).expand(-1, n, -1).reshape(B * n, 512)
        # img2img stages
        xx = M._as_nhwc(x)
        skip = [m.att for m, nb in ((net.img2img.down_block, net.img2img.down_block.n_down_blocks), (net.img2img.adain_up_block, net.img2img.adain_up_block.n_up_blocks)) if not m.att_loc < nb]
        from optimalstrategiesagainstgenerativeattacks_b200 import model_blocks as mb
        mb.sn_prepare_module(net.img2img, skip=skip)
        blocks = list(net.img2img.adain_res_block.res_blocks) + list(net.img2img.adain_up_block.up_blocks)
        styles = mb.batched_style_projections(blocks, style)
        n_res = len(net.img2img.adain_res_block.res_blocks)
        d = net.img2img.down_block(xx)
        out["i2i_down"] = ops.from_nhwc(d)
        r = net.img2img.adain_res_block(x=d, style=style, styles=styles[:n_res])
        out["i2i_res"] = ops.from_nhwc(r)
        out["fake"] = ops.from_nhwc(net.img2img.adain_up_block(x=r, style=style, styles=styles[n_res:]))
    return {k: v.float().clone() for k, v in out.items()}


def oracle(dtype, rounding):
    p = {k: v.to(dev, dtype).clone() for k, v in sd0.items()}
    O.set_operand_rounding(rounding)
    out = {}
    try:
        with torch.no_grad():
            lk, zz = leaked.to(dtype), z.to(dtype)
            out["src_enc"] = O.encode_sample(p, "src_encoder", lk)
            out["env_enc"] = O.encode_sample(p, "env_encoder", lk)
            out["w"] = O.mlp(p, "env_noise_mapper", zz, 4)
            w = out["w"] - out["w"].mean(1, keepdim=True)
            noisy = out["env_enc"].mean(1).unsqueeze(1) + w
            out["env_img"] = O.env_decoder(p, "env_decoder", noisy.reshape(B * n, -1), size)
            expanded = lk[:, 0].unsqueeze(1).expand(-1, n, -1, -1, -1).reshape(B * n, ch, size, size)
            x = torch.cat((out["env_img"], expanded), 1)
            style = out["src_enc"].mean(1).unsqueeze(1).expand(-1, n, -1).reshape(B * n, -1)
            nb = O.n_down_blocks(size)
            att_loc = -(-nb // 2)
            dd = "img2img.down_block"
            for i in range(nb):
                if i == att_loc:
                    x = O.self_attention(p, dd + ".att", x)
                x = O.res_block_down(p, "%s.down_blocks.%d" % (dd, i), x, 9 if i == 0 else 3)
                x = O.instance_norm(x, p["%s.in_layers.%d.weight" % (dd, i)], p["%s.in_layers.%d.bias" % (dd, i)])
            out["i2i_down"] = x
            for i in range(5):
                x = O.ada_res_block2(p, "img2img.adain_res_block.res_blocks.%d" % i, x, style)
            out["i2i_res"] = x
            u = "img2img.adain_up_block"
            for i in range(nb):
                if i == att_loc:
                    x = O.self_attention(p, u + ".att", x)
                x = O.ada_res_block_up2(p, "%s.up_blocks.%d" % (u, i), x, style, 9 if i == nb - 1 else 3)
            out["fake"] = torch.tanh(x)
    finally:
        O.set_operand_rounding(False)
    return out


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
A, Bq, E, R = ours("tcgen05"), ours("simt"), oracle(torch.float32, True), oracle(torch.float64, False)
F32 = oracle(torch.float32, False)
print("%-10s %10s %10s %10s %10s %10s %10s %10s" % ("stage", "A vs B", "A vs E", "B vs E", "A vs R", "B vs R", "E vs R", "fp32 vs R"))
for k in ("src_enc", "env_enc", "w", "env_img", "i2i_down", "i2i_res", "fake"):
    print("%-10s %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e %10.2e" % (k, rel(A[k], Bq[k]), rel(A[k], E[k]), rel(Bq[k], E[k]), rel(A[k], R[k]), rel(Bq[k], R[k]),
                                                                    rel(E[k], R[k]), rel(F32[k], R[k])))
