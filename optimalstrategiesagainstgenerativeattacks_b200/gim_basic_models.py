"""Permutation-invariant sample statistics -- mirrors the reference's models/gim_basic_models.py (the used subset:
GIMMeanStat :20-34, GIMStdStat :37-51, GIMMeanStdStat :71-89, GIMFCStat :113-127, GIMMeanStdFcStat :152-172)."""
import torch
import torch.nn as nn

from . import model_blocks as mb
from . import ops


class GIMMeanStat(nn.Module):
    def __init__(self):
        super().__init__()
        self.n_stats = 1

    def forward(self, x):
        """[batch, sample_size, latent] -> [batch, latent]"""
        return ops.set_mean(x)


class GIMStdStat(nn.Module):
    def __init__(self):
        super().__init__()
        self.n_stats = 1

    def forward(self, x):
        return mb.custom_std(x)


class GIMMeanStdStat(nn.Module):
    def __init__(self):
        super().__init__()
        self.n_stats = 2
        self.sample_mean = GIMMeanStat()
        self.sample_std = GIMStdStat()

    def forward(self, x):
        return ops.set_mean_std(x)                      # mean | std side by side from one pass over x


class GIMFCStat(nn.Module):
    def __init__(self, style_dim, n_stats=1, hidden_layers=()):
        super().__init__()
        self.style_dim = style_dim
        self.n_stats = n_stats
        self.fc_layer_dims = [style_dim] + [*hidden_layers] + [n_stats * style_dim]
        self.stat = mb.MLP(self.fc_layer_dims)
        self.sample_mean = GIMMeanStat()

    def forward(self, x):
        return self.sample_mean(self.stat(x))


class GIMMeanStdFcStat(nn.Module):
    def __init__(self, style_dim, fc_n_stats, fc_hidden_layers):
        super().__init__()
        self.n_stats = 2 + fc_n_stats
        self.sample_mean = GIMMeanStat()
        self.sample_std = GIMStdStat()
        self.fc = GIMFCStat(style_dim=style_dim, n_stats=fc_n_stats, hidden_layers=fc_hidden_layers)

    def forward(self, x):
        return torch.cat((ops.set_mean_std(x), self.fc(x)), dim=-1)
