"""Single-GPU check of data-parallel algebra: gradients of the global batch vs the mean of the gradients of its two halves (fp32 path)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import optimalstrategiesagainstgenerativeattacks_b200 as gim
from optimalstrategiesagainstgenerativeattacks_b200 import gim_img_models as M
from optimalstrategiesagainstgenerativeattacks_b200.gim_img_trainer import GIMImgTrainer

dev = torch.device("cuda", 0)
size, ch, B = 32, 1, 8
gim.set_precision("fp32")
gim.set_deterministic(True)
leaked, real, si = bench.synth_batch(B, ch, size, 4321, dev)
z = torch.randn((B, 5, 512), generator=torch.Generator().manual_seed(7)).to(dev)
real_randn = torch.randn


def grads_of(sl):
    torch.manual_seed(1)
    au, im = M.get_au(size, ch, 512).to(dev), M.get_im(size, ch, 512).to(dev)
    tr = GIMImgTrainer(tempfile.mkdtemp(), 5, 5, 5, au, im, 1e-6, 1e-5, 1e-7, reg_param=0.0)
    torch.randn = lambda *a, **k: z[sl].clone()
    try:
        loss, fake, _ = tr.impersonator_forward(leaked[sl], si[sl])
    finally:
        torch.randn = real_randn
    loss.mean().backward()
    return [(n, p.grad.clone()) for n, p in im.named_parameters() if p.grad is not None], fake.detach(), loss.detach()


full, fake_f, loss_f = grads_of(slice(0, B))
a, fake_a, loss_a = grads_of(slice(0, B // 2))
b, fake_b, loss_b = grads_of(slice(B // 2, B))
print("fake equal per episode:", float((torch.cat((fake_a, fake_b)) - fake_f).abs().max()), "loss", float((torch.cat((loss_a, loss_b)) - loss_f).abs().max()))
rows = []
for (n, gf), (_, ga), (_, gb) in zip(full, a, b):
    m = 0.5 * (ga + gb)
    rows.append((float((m - gf).norm() / max(float(gf.norm()), 1e-30)), n, float(gf.norm())))
rows.sort(reverse=True)
for r in rows[:12]:
    print("%.3e  %-60s |g| %.3e" % r)
print("median", sorted(r[0] for r in rows)[len(rows) // 2])
