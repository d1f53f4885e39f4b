"""Image GIM networks -- same classes, constructor arguments, attribute names, forward signatures and state-dict schema as
the reference's models/gim_img_models.py.  Module inputs/outputs are the reference's NCHW fp32 tensors; inside, activations
are NHWC in the active precision and every op is a libgim_b200 kernel."""
import math

import torch
import torch.nn as nn

from . import model_blocks as mb
from . import ops
from .gim_basic_models import GIMMeanStdFcStat


def _as_nhwc(x):
    """Accept the reference's NCHW fp32 image tensors at the module boundary; internal tensors are tagged NHWC."""
    if isinstance(x, NHWC):
        return x.t
    return ops.to_nhwc(x)


class NHWC:
    """Marks an activation that is already in the internal NHWC layout (used between the kept-name modules)."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t


def _down_plan(img_size, img_channels, style_dim, min_n_channels):
    """Channel plan of a down-sampling pyramid that ends at 4x4 (reference gim_img_models.py:27-33, 108-113):
    -> (number of blocks, floor of the channel count, [c_in, c_1, ..., c_n], index of the block the attention precedes)."""
    n_blocks = int(math.log2(img_size)) - 2
    floor = int(max(min_n_channels, style_dim / (2 ** (n_blocks - 1))))
    widths = [img_channels] + [min(style_dim, int(floor * (2 ** i))) for i in range(n_blocks)]
    return n_blocks, floor, widths, int(math.ceil(n_blocks / 2))


def _up_plan(n_blocks, img_channels, style_dim, floor):
    """Mirror image for the up-sampling pyramids (reference :70-74, 151-156): widest first, image channels last."""
    widths = [min(style_dim, int(floor * (2 ** i))) for i in range(n_blocks)][::-1] + [img_channels]
    return widths, int(math.ceil(n_blocks / 2))


class Encoder(nn.Module):
    """Reference gim_img_models.py:19-57: [N, C, S, S] -> [N, style_dim]: ResBlockDown pyramid to 4x4 (self-attention before block
    `att_loc`), global max, LeakyReLU."""

    def __init__(self, img_size, img_channels, style_dim=512, min_n_channels=64, use_out_lrelu=True):
        super().__init__()
        self.img_size, self.img_channels, self.style_dim, self.use_out_lrelu = img_size, img_channels, style_dim, use_out_lrelu
        self.n_down_blocks, self.min_n_channels, self.channel_sizes, self.att_loc = _down_plan(img_size, img_channels, style_dim, min_n_channels)
        self.down_blocks = nn.ModuleList(mb.ResBlockDown(ci, co) for ci, co in zip(self.channel_sizes[:-1], self.channel_sizes[1:]))
        self.att = mb.SelfAttention(self.channel_sizes[self.att_loc])

    def forward(self, x):
        x = _as_nhwc(x)
        mb.sn_prepare_module(self, skip=() if self.att_loc < self.n_down_blocks else (self.att,))
        for i, block in enumerate(self.down_blocks):
            if i == self.att_loc:
                x = self.att(x)
            x = block(x, want_ops=i + 1 < self.n_down_blocks)       # the next consumer reads the bf16 operands
        if isinstance(x, ops.Act):
            x = x.t32
        return ops.GlobalMaxFn.apply(x, 0.2 if self.use_out_lrelu else 1.0)      # global max + LeakyReLU in one pass


class EnvDecoder(nn.Module):
    """Reference gim_img_models.py:63-95: [N, style_dim] -> [N, C, S, S]: ResBlockUp pyramid from 1x1 (no output nonlinearity)."""

    def __init__(self, img_size, img_channels, style_dim=512, min_n_channels=64):
        super().__init__()
        self.img_size, self.img_channels, self.style_dim, self.min_n_channels = img_size, img_channels, style_dim, min_n_channels
        self.n_up_blocks = int(math.log2(img_size))
        self.channel_sizes, self.att_loc = _up_plan(self.n_up_blocks, img_channels, style_dim, min_n_channels)
        self.up_blocks = nn.ModuleList(mb.ResBlockUp(ci, co) for ci, co in zip(self.channel_sizes[:-1], self.channel_sizes[1:]))
        self.att = mb.SelfAttention(self.channel_sizes[self.att_loc])

    def forward(self, x, nhwc_out=False):
        n, c = x.shape
        x = x.reshape(n, 1, 1, c)                          # a [N, C] feature vector is a 1x1 NHWC activation
        if x.dtype != ops.act_dtype():
            x = ops.to_nhwc(x.reshape(n, c, 1, 1))
        mb.sn_prepare_module(self, skip=() if self.att_loc < self.n_up_blocks else (self.att,))
        for i, block in enumerate(self.up_blocks):
            if i == self.att_loc:
                x = self.att(x)
            x = block(x)
        return NHWC(x) if nhwc_out else ops.from_nhwc(x)


class Img2ImgDownModule(nn.Module):
    """Reference gim_img_models.py:101-139: ResBlockDown (9x9 in the first block) + affine InstanceNorm per level."""

    def __init__(self, img_size, img_channels, style_dim=512, min_n_channels=64):
        super().__init__()
        self.img_size, self.img_channels, self.style_dim = img_size, img_channels, style_dim
        self.n_down_blocks, self.min_n_channels, self.channel_sizes, self.att_loc = _down_plan(img_size, img_channels, style_dim, min_n_channels)
        self.down_blocks, self.in_layers = nn.ModuleList(), nn.ModuleList()
        for i, (ci, co) in enumerate(zip(self.channel_sizes[:-1], self.channel_sizes[1:])):
            wide = {"conv_size": 9, "padding_size": 4} if i == 0 else {}
            self.down_blocks.append(mb.ResBlockDown(ci, co, **wide))
            self.in_layers.append(mb.InstanceNormAffine(co))
        self.att = mb.SelfAttention(self.channel_sizes[self.att_loc])

    def forward(self, x):
        for i, (block, norm) in enumerate(zip(self.down_blocks, self.in_layers)):
            if i == self.att_loc:
                x = self.att(x)
            x = block(x)
            x = norm(x.t32 if isinstance(x, ops.Act) else x)
        return x


class Img2ImgAdaInResModule(nn.Module):
    """Reference gim_img_models.py:142-162: `n_blocks` AdaResBlock2 at the bottleneck resolution."""

    def __init__(self, style_dim=512, n_blocks=5):
        super().__init__()
        self.style_dim, self.n_blocks = style_dim, n_blocks
        self.res_blocks = nn.ModuleList(mb.AdaResBlock2(channels=style_dim, style_dim=style_dim) for _ in range(n_blocks))

    def forward(self, x, style, styles=None):
        for i, block in enumerate(self.res_blocks):
            x = block(x=x, style=style, styles=None if styles is None else styles[i])
        return x


class Img2ImgAdaInUpModule(nn.Module):
    """Reference gim_img_models.py:165-215: AdaResBlockUp2 pyramid back to the image (9x9 in the last block), tanh."""

    def __init__(self, img_size, img_channels, style_dim=512, min_n_channels=64):
        super().__init__()
        self.img_size, self.img_channels, self.style_dim = img_size, img_channels, style_dim
        self.n_up_blocks = int(math.log2(img_size)) - 2
        self.min_n_channels = int(max(min_n_channels, style_dim / (2 ** (self.n_up_blocks - 1))))
        self.channel_sizes, self.att_loc = _up_plan(self.n_up_blocks, img_channels, style_dim, self.min_n_channels)
        self.up_blocks = nn.ModuleList()
        for i, (ci, co) in enumerate(zip(self.channel_sizes[:-1], self.channel_sizes[1:])):
            k = 9 if i == self.n_up_blocks - 1 else 3
            self.up_blocks.append(mb.AdaResBlockUp2(in_channels=ci, out_channels=co, style_dim=style_dim, conv_size=k, padding_size=(k - 1) // 2))
        self.att = mb.SelfAttention(self.channel_sizes[self.att_loc])

    def forward(self, x, style, styles=None):
        for i, block in enumerate(self.up_blocks):
            if i == self.att_loc:
                x = self.att(x)
            x = block(x=x, style=style, styles=None if styles is None else styles[i])
        return ops.TanhFn.apply(x)


class AdaInImage2Image(nn.Module):
    """Reference gim_img_models.py:218-257: [N, in_channels, S, S], style [N, style_dim] -> [N, out_channels, S, S]."""

    def __init__(self, img_size, in_channels, out_channels, style_dim, n_adain_res_blocks=5, min_n_channels=64):
        super().__init__()
        self.img_size, self.in_channels, self.out_channels, self.style_dim = img_size, in_channels, out_channels, style_dim
        self.n_adain_res_blocks, self.min_n_channels = n_adain_res_blocks, min_n_channels
        self.down_block = Img2ImgDownModule(img_size=img_size, img_channels=in_channels, style_dim=style_dim, min_n_channels=min_n_channels)
        self.adain_res_block = Img2ImgAdaInResModule(style_dim=style_dim, n_blocks=n_adain_res_blocks)
        self.adain_up_block = Img2ImgAdaInUpModule(img_size=img_size, img_channels=out_channels, style_dim=style_dim, min_n_channels=min_n_channels)

    def forward(self, x, style):
        x = _as_nhwc(x)
        skip = [m.att for m, n_blocks in ((self.down_block, self.down_block.n_down_blocks), (self.adain_up_block, self.adain_up_block.n_up_blocks))
                if not m.att_loc < n_blocks]
        mb.sn_prepare_module(self, skip=skip)
        # the 4 x (n_res + n_up) style Linears share their input: one GEMM over the concatenated weights
        blocks = list(self.adain_res_block.res_blocks) + list(self.adain_up_block.up_blocks)
        styles = mb.batched_style_projections(blocks, style)
        n_res = len(self.adain_res_block.res_blocks)
        x = self.down_block(x)
        x = self.adain_res_block(x=x, style=style, styles=styles[:n_res])
        x = self.adain_up_block(x=x, style=style, styles=styles[n_res:])
        return ops.from_nhwc(x)


class GIMFaceDis(nn.Module):
    """Reference gim_img_models.py:263-299."""

    def __init__(self, src_dim, env_dim, stat):
        super().__init__()
        self.src_dim = src_dim
        self.env_dim = env_dim
        self.stat = stat
        self.n_stats = stat.n_stats
        mlp_input_dim = 2 * (self.n_stats * env_dim + src_dim)
        self.mlp = mb.MLP((mlp_input_dim, env_dim + src_dim, 2 * (env_dim + src_dim), 1))
        self.mlp.apply(mb.weights_init('kaiming'))

    def forward(self, test_src, test_env, si_src, si_env):
        test_src_mean = ops.set_mean(test_src)
        si_src_mean = ops.set_mean(si_src)
        test_env_stat = self.stat(test_env)
        si_env_stat = self.stat(si_env)
        x = torch.cat((test_src_mean, si_src_mean, test_env_stat, si_env_stat), dim=-1)
        return self.mlp(x)


class _EncodeMixin:
    def _encode(self, encoder, sample):
        batch_size, sample_size = sample.size(0), sample.size(1)
        x = encoder(sample.reshape(batch_size * sample_size, *sample.size()[2:]))
        return x.view(batch_size, sample_size, *x.size()[1:])

    def src_encode_sample(self, sample):
        return self._encode(self.src_encoder, sample)

    def env_encode_sample(self, sample):
        return self._encode(self.env_encoder, sample)


class GIMFaceAuthenticator(nn.Module, _EncodeMixin):
    """Reference gim_img_models.py:304-340."""

    def __init__(self, src_encoder, env_encoder, dis):
        super().__init__()
        self.src_encoder = src_encoder
        self.env_encoder = env_encoder
        self.dis = dis

    def forward(self, test_sample, si_sample):
        # same call order per encoder as the reference (src: test, si; env: test, si); the two encoders are independent -> two streams
        (test_src, si_src), (test_env, si_env) = ops.two_streams(
            lambda: (self.src_encode_sample(test_sample), self.src_encode_sample(si_sample)),
            lambda: (self.env_encode_sample(test_sample), self.env_encode_sample(si_sample)))
        return self.dis(test_src=test_src, test_env=test_env, si_src=si_src, si_env=si_env)


class GIMFaceImpersonator(nn.Module, _EncodeMixin):
    """Reference gim_img_models.py:346-423."""

    def __init__(self, src_encoder, env_encoder, env_decoder, img2img, env_noise_mapper, use_img_att=False):
        super().__init__()
        self.src_encoder = src_encoder
        self.env_encoder = env_encoder
        self.env_decoder = env_decoder
        self.img2img = img2img
        self.env_noise_mapper = env_noise_mapper
        self.style_dim = src_encoder.style_dim
        assert src_encoder.style_dim == env_encoder.style_dim == env_decoder.style_dim == img2img.style_dim
        self.use_img_att = use_img_att
        self.img_att = mb.ImgAttention(img1_channels=self.src_encoder.img_channels, img2_channels=self.img2img.out_channels)

    def forward(self, leaked_sample, n, remove_noise_mean=True):
        batch_size, m, img_channels, img_size, _ = leaked_sample.size()
        expanded_img = leaked_sample[:, 0].unsqueeze(1).expand(-1, n, -1, -1, -1)

        src, env = ops.two_streams(lambda: ops.set_mean(self.src_encode_sample(leaked_sample)),
                                   lambda: ops.set_mean(self.env_encode_sample(leaked_sample)))

        z = torch.randn((batch_size, n, self.style_dim), device=leaked_sample.device)
        w = self.env_noise_mapper(z)
        noisy_env = ops.SetCenterAddFn.apply(w, env, bool(remove_noise_mean))

        env_img = self.env_decoder(noisy_env.view(batch_size * n, self.style_dim), nhwc_out=True).t
        exp_nhwc = ops.to_nhwc(expanded_img.reshape(batch_size * n, img_channels, img_size, img_size))
        x = self.generate_img(env_img=NHWC(ops.CatChannelsFn.apply(env_img, exp_nhwc)), src=src, n=n)

        if self.use_img_att:                               # reference :392-396 (non-default)
            x = ops.from_nhwc(self.img_att(x1=exp_nhwc, x2=ops.to_nhwc(x.reshape(batch_size * n, *x.size()[2:]))))
            x = x.view(batch_size, n, *x.size()[1:])
        return x

    def generate_img(self, env_img, src, n=None):
        if isinstance(env_img, NHWC):
            batch_size = src.size(0)
            x = env_img
        else:
            batch_size, n = env_img.size(0), env_img.size(1)
            x = env_img.contiguous().view(batch_size * n, *env_img.size()[2:])
        style = src.unsqueeze(1).expand(-1, n, -1).contiguous().view(batch_size * n, self.style_dim)
        gen_img = self.img2img(x=x, style=style)
        return gen_img.view(batch_size, n, *gen_img.size()[1:])


def get_im(img_size, img_channels, style_dim, use_img_att=False, num_env_noise_layers=4):
    """Reference gim_img_models.py:429-449."""
    src_encoder = Encoder(img_size=img_size, img_channels=img_channels, style_dim=style_dim)
    env_encoder = Encoder(img_size=img_size, img_channels=img_channels, style_dim=style_dim)
    decoder = EnvDecoder(img_size=img_size, img_channels=img_channels, style_dim=style_dim)
    img2img = AdaInImage2Image(img_size=img_size, in_channels=2 * img_channels, out_channels=img_channels, style_dim=style_dim)
    env_noise_mapper = mb.MLP([style_dim for _ in range(num_env_noise_layers + 1)])
    return GIMFaceImpersonator(src_encoder=src_encoder, env_encoder=env_encoder, env_decoder=decoder, img2img=img2img,
                               env_noise_mapper=env_noise_mapper, use_img_att=use_img_att)


def get_au(img_size, img_channels, style_dim):
    """Reference gim_img_models.py:452-463."""
    stat = GIMMeanStdFcStat(style_dim=style_dim, fc_n_stats=2, fc_hidden_layers=(style_dim * 2, style_dim * 3, style_dim * 2))
    dis = GIMFaceDis(src_dim=style_dim, env_dim=style_dim, stat=stat)
    src_encoder = Encoder(img_size=img_size, img_channels=img_channels, style_dim=style_dim)
    env_encoder = Encoder(img_size=img_size, img_channels=img_channels, style_dim=style_dim)
    return GIMFaceAuthenticator(src_encoder=src_encoder, env_encoder=env_encoder, dis=dis)
