"""Gaussian GIM networks -- mirrors the reference's models/gim_gaussian_models.py (GIMGaussianDis :17-41,
GIMGaussianAuthenticator :47-60, GIMGaussianImpersonator :66-89, get_im :95-99, get_au :102-107)."""
import torch
import torch.nn as nn

from . import model_blocks as mb
from . import ops
from .gim_basic_models import GIMMeanStdStat


class GIMGaussianDis(nn.Module):
    def __init__(self, src_dim, stat):
        super().__init__()
        self.src_dim = src_dim
        self.stat = stat
        self.n_stats = stat.n_stats
        self.mlp = mb.MLP((self.n_stats * src_dim * 2, src_dim, 2 * src_dim, 1))
        self.mlp.apply(mb.weights_init('kaiming'))

    def forward(self, test_sample, si_sample):
        """test [batch, n, d], si [batch, k, d] -> [batch, 1]"""
        x = torch.cat((self.stat(test_sample), self.stat(si_sample)), dim=-1)
        return self.mlp(x)


class GIMGaussianAuthenticator(nn.Module):
    def __init__(self, dis):
        super().__init__()
        self.dis = dis

    def forward(self, test_sample, si_sample):
        return self.dis(test_sample=test_sample, si_sample=si_sample)


class GIMGaussianImpersonator(nn.Module):
    def __init__(self, src_dim, env_noise_mapper):
        super().__init__()
        self.src_dim = src_dim
        self.env_noise_mapper = env_noise_mapper
        self.out_mlp = mb.MLP((2 * src_dim, 2 * src_dim, src_dim))      # constructed and check-pointed, never used (:73)

    def forward(self, leaked_sample, n, remove_noise_mean=True):
        batch_size, m, src_dim = leaked_sample.size()
        src = ops.set_mean(leaked_sample)
        z = torch.randn((batch_size, n, self.src_dim), device=leaked_sample.device)
        w = self.env_noise_mapper(z)
        return ops.SetCenterAddFn.apply(w, src, bool(remove_noise_mean))


def get_im(src_dim):
    env_noise_mapper = mb.MLP([src_dim, src_dim])
    return GIMGaussianImpersonator(src_dim=src_dim, env_noise_mapper=env_noise_mapper)


def get_au(src_dim):
    stat = GIMMeanStdStat()
    dis = GIMGaussianDis(src_dim=src_dim, stat=stat)
    return GIMGaussianAuthenticator(dis=dis)
