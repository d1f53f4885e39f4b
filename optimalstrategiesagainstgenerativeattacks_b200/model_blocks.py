"""Building blocks of the GIM image networks -- same constructors, attribute names and state-dict schema as the
reference's models/model_blocks.py, every forward running on the libgim_b200 kernels (NHWC activations inside).

Only the blocks the entry points instantiate are provided: weights_init, custom_std, MLP, ResBlockDown,
SelfAttention, ImgAttConvBlock / ImgAttention (constructed for checkpoint parity), ada_in, ResBlockUp, AdaResBlock2,
AdaResBlockUp2.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

SN_EPS = 1e-12


def weights_init(init_type='kaiming'):
    """Reference model_blocks.py:18-38 (applied to the discriminator MLPs with 'kaiming')."""
    def init_fun(m):
        if isinstance(m, (nn.Linear, SNConv2d)) and hasattr(m, 'weight'):
            if init_type == 'gaussian':
                nn.init.normal_(m.weight.data, 0.0, 0.02)
            elif init_type == 'xavier':
                nn.init.xavier_normal_(m.weight.data, gain=math.sqrt(2))
            elif init_type == 'kaiming':
                nn.init.kaiming_normal_(m.weight.data, a=0.2)
            elif init_type == 'orthogonal':
                nn.init.orthogonal_(m.weight.data, gain=math.sqrt(2))
            elif init_type == 'default':
                pass
            else:
                assert 0, "Unsupported initialization: {}".format(init_type)
            if hasattr(m, 'bias') and m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)
    return init_fun


def custom_std(x):
    """Reference model_blocks.py:41-48 on [batch, sample, latent] fp32."""
    return ops.set_std(x)


def ada_in(feature, mean_style, std_style, eps=1e-5):
    """Reference model_blocks.py:611-630; `feature` is an NHWC activation here, styles [batch, channels(,1)]."""
    n, _, _, c = feature.shape
    return ops.ada_in(feature, mean_style.reshape(n, c), std_style.reshape(n, c), eps)


class Linear(nn.Linear):
    """nn.Linear with the forward on the C-ABI GEMM (+ fused bias / LeakyReLU)."""

    def forward(self, x, slope=1.0):
        return ops.linear(x, self.weight, self.bias, slope)


class MLP(nn.Module):
    """Reference model_blocks.py:77-94: Linear / LeakyReLU(0.2) alternating, keys `model.{0,2,4,..}.{weight,bias}`."""

    def __init__(self, layer_dims):
        super().__init__()
        assert len(layer_dims) >= 2
        layers = []
        inp_dim = layer_dims[0]
        for out_dim in layer_dims[1:-1]:
            layers.append(Linear(inp_dim, out_dim))
            layers.append(nn.LeakyReLU(0.2))
            inp_dim = out_dim
        layers.append(Linear(inp_dim, layer_dims[-1]))
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        mods = list(self.model)
        i = 0
        while i < len(mods):
            fused = i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU)
            x = mods[i](x, mods[i + 1].negative_slope if fused else 1.0)
            i += 2 if fused else 1
        return x


def _sn_state_dict_hook(module, state_dict, prefix, local_metadata):
    """torch.nn.utils.spectral_norm's SpectralNormStateDictHook: the entry its load pre-hook reads to tell a current-format
    (weight_orig / weight_u / weight_v) state dict from a legacy one, so files written here load in the reference's modules."""
    local_metadata.setdefault('spectral_norm', {})['weight.version'] = 1


class SNConv2d(nn.Module):
    """nn.utils.spectral_norm(nn.Conv2d(cin, cout, k, padding=(k-1)//2)) -- parameters `bias`, `weight_orig`, buffers
    `weight_u`, `weight_v`, initialised with the same RNG draws as torch (Conv2d.reset_parameters, then u, v)."""

    def __init__(self, in_channels, out_channels, kernel_size, padding=0):
        super().__init__()
        assert padding == (kernel_size - 1) // 2, "the GIM path only uses 'same' convolutions"
        conv = nn.Conv2d(in_channels, out_channels, kernel_size, padding=padding)
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.bias = nn.Parameter(conv.bias.data)
        self.weight_orig = nn.Parameter(conv.weight.data)
        w = conv.weight.data
        height, width = w.shape[0], w[0].numel()
        u = F.normalize(w.new_empty(height).normal_(0, 1), dim=0, eps=SN_EPS)
        v = F.normalize(w.new_empty(width).normal_(0, 1), dim=0, eps=SN_EPS)
        self.register_buffer('weight_u', u)
        self.register_buffer('weight_v', v)
        self._prepared = None
        self._register_state_dict_hook(_sn_state_dict_hook)

    def effective_weight(self):
        """Packed fp32 [k*k, cout, cin] weight W/sigma; runs the power iteration when self.training (unless sn_prepare_module already
        did it for this call as part of a batched pass)."""
        prep, self._prepared = self._prepared, None
        if prep is not None and prep[4] == self.weight_orig._version:       # (a result left over from an aborted forward is dropped)
            return ops.sn_prepared_weight(self.weight_orig, prep)
        return ops.SpectralNormFn.apply(self.weight_orig, (self.weight_u, self.weight_v), self.training, SN_EPS)

    def forward(self, x, pre=ops.PRE_NONE, slope=0.2):
        """`pre`: operator folded in front of the conv (ops.PRE_LRELU: LeakyReLU(slope); ops.PRE_UPSAMPLE: nearest x2)."""
        return ops.Conv2dFn.apply(x, self.effective_weight(), self.bias, self.kernel_size, pre, slope)


def sn_prepare_module(module, skip=()):
    """Batched spectral norm for every SNConv2d under `module` that its forward is about to call exactly once (reference semantics:
    one power iteration per forward call of each conv, torch.nn.utils.spectral_norm).  `skip`: sub-modules whose convs do not run."""
    skipped = set()
    for sm in skip:
        skipped.update(id(m) for m in sm.modules())
    convs = [m for m in module.modules() if isinstance(m, SNConv2d) and id(m) not in skipped]
    if convs:
        ops.sn_prepare(convs, module.training, SN_EPS)



class InstanceNormAffine(nn.Module):
    """nn.InstanceNorm2d(channels, affine=True) (no running stats): keys `weight`, `bias`."""

    def __init__(self, channels, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))

    def forward(self, x, slope=1.0):
        return ops.instance_norm(x, self.weight, self.bias, self.eps, slope)


class ResBlockDown(nn.Module):
    """Reference model_blocks.py:486-514.  AvgPool(left) + AvgPool(right) is one fused pool-of-sum kernel."""

    def __init__(self, in_channel, out_channel, conv_size=3, padding_size=1):
        super().__init__()
        self.lrelu = nn.LeakyReLU(0.2)
        self.avg_pool2d = nn.AvgPool2d(2)
        self.conv_l1 = SNConv2d(in_channel, out_channel, 1)
        self.conv_r1 = SNConv2d(in_channel, out_channel, conv_size, padding=padding_size)
        self.conv_r2 = SNConv2d(out_channel, out_channel, conv_size, padding=padding_size)

    def forward(self, x, want_ops=False):
        """x: fp32 NHWC tensor or ops.Act.  On the bf16 tensor-core path the whole block is one fused autograd node and the result is
        an ops.Act (fp32 output + the bf16 operands of the next block when `want_ops`); otherwise a plain tensor."""
        if ops.fused_blocks_enabled() and self.conv_r2.out_channels % 32 == 0:
            return ops.res_block_down(x, self.conv_l1.effective_weight(), self.conv_l1.bias, self.conv_r1.effective_weight(), self.conv_r1.bias,
                                      self.conv_r2.effective_weight(), self.conv_r2.bias, self.conv_r1.kernel_size, 0.2, want_ops)
        if isinstance(x, ops.Act):
            x = x.t32
        out = self.conv_r1(x, ops.PRE_LRELU)          # conv(lrelu(x)): the activation is fused into the operand producer
        out = self.conv_r2(out, ops.PRE_LRELU)
        if ops.get_precision() == "bf16" and x.shape[1] % 2 == 0 and x.shape[2] % 2 == 0:
            # mixed-precision path: the 1x1 residual conv commutes with the pooling and runs at the pooled resolution, exactly like the
            # fused block (ops.ResBlockDownFn) -- same operands, same single bf16 rounding of AvgPool2(x); 4x fewer FLOPs
            out_res = self.conv_l1(ops.avg_pool2_add(x))
            return ops.AddFn.apply(ops.avg_pool2_add(out), out_res)
        out_res = self.conv_l1(x)
        return ops.avg_pool2_add(out_res, out)


class SelfAttention(nn.Module):
    """Reference model_blocks.py:517-549 (softmax over the key axis, gamma * out + x)."""

    def __init__(self, in_channel):
        super().__init__()
        self.conv_f = SNConv2d(in_channel, in_channel // 8, 1)
        self.conv_g = SNConv2d(in_channel, in_channel // 8, 1)
        self.conv_h = SNConv2d(in_channel, in_channel, 1)
        self.softmax = nn.Softmax(-2)
        self.gamma = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        if isinstance(x, ops.Act):
            ops.register_operand(x.t32, x.tb)             # the three 1x1 convs read the bf16 copy the pooling kernel already wrote
            x = x.t32
        n, h, w, c = x.shape
        if ops.attention_fused_ok(h * w, c, x):
            # 8x8 maps: the three projections are ONE 1x1 convolution [keys | queries | values] (x is read once; one input-gradient and
            # one weight-gradient launch instead of three) feeding the fused attention kernels
            convs = (self.conv_f, self.conv_g, self.conv_h)
            kqv = ops.conv2d(x, ops.merged_sn_weight(convs), ops.cat_params([m.bias for m in convs]), 1)
            return ops.AttentionPackedFn.apply(kqv.reshape(n, h * w, -1), x.reshape(n, h * w, c), self.gamma).reshape(n, h, w, c)
        f = self.conv_f(x).reshape(n, h * w, -1)          # keys    [n, N, c/8]
        g = self.conv_g(x).reshape(n, h * w, -1)          # queries [n, N, c/8]
        hp = self.conv_h(x).reshape(n, h * w, c)          # values  [n, N, c]
        p = ops.matmul(g, f, False, True, torch.float32)  # p[j, i] = <g_j, f_i> = attention_map[i, j]
        a = ops.SoftmaxRowsFn.apply(p)                    # softmax over i  (reference: dim=-2 of [i, j])
        out = ops.matmul(a, hp, False, False, x.dtype).reshape(n, h, w, c)
        return ops.AddFn.apply(ops.ScaleDevFn.apply(out, self.gamma), x)


class ImgAttConvBlock(nn.Module):
    """Reference model_blocks.py:551-580 -- constructed for checkpoint parity; only runs when use_img_att=True."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lrelu = nn.LeakyReLU(0.2)
        self.conv_l1 = SNConv2d(in_channels, out_channels, 1)
        self.conv_r1 = SNConv2d(in_channels, out_channels, 9, padding=4)
        self.conv_r2 = SNConv2d(out_channels, out_channels, 3, padding=1)

    def forward(self, x):
        out = self.conv_r1(x, ops.PRE_LRELU)
        out = self.conv_r2(out, ops.PRE_LRELU)
        return ops.AddFn.apply(self.conv_l1(x), out)


class ImgAttention(nn.Module):
    """Reference model_blocks.py:583-608 (use_img_att=True; the CLI default is False): five ImgAttConvBlocks produce two queries, two
    keys and a value image, a per-pixel 2-way softmax over the channel dot products blends the leaked image x1 with the value image."""

    def __init__(self, img1_channels, img2_channels):
        super().__init__()
        self.q1conv = ImgAttConvBlock(img1_channels + img2_channels, img1_channels)
        self.q2conv = ImgAttConvBlock(img1_channels + img2_channels, img1_channels)
        self.k1conv = ImgAttConvBlock(img1_channels, img1_channels)
        self.k2conv = ImgAttConvBlock(img2_channels, img1_channels)
        self.v2conv = ImgAttConvBlock(img2_channels, img1_channels)

    def forward(self, x1, x2):
        """x1, x2: NHWC fp32 activations [n, h, w, c1], [n, h, w, c2] -> [n, h, w, c1]."""
        sn_prepare_module(self)
        x = ops.CatChannelsFn.apply(x1, x2)
        q1, q2 = self.q1conv(x), self.q2conv(x)
        k1, k2, v2 = self.k1conv(x1), self.k2conv(x2), self.v2conv(x2)
        return ops.ImgAttBlendFn.apply(q1, k1, q2, k2, x1, v2)


class ResBlockUp(nn.Module):
    """Reference model_blocks.py:733-773.  conv_l1 is 1x1, so conv(up(x)) == up(conv(x)) pixel for pixel: the conv runs at
    the low resolution (4x fewer FLOPs)."""

    def __init__(self, in_channel, out_channel, out_size=None, scale=2, conv_size=3, padding_size=1, use_norm=True):
        super().__init__()
        assert out_size is None and scale == 2
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = nn.Upsample(size=out_size, scale_factor=scale)
        self.lrelu = nn.LeakyReLU(0.2)
        self.conv_l1 = SNConv2d(in_channel, out_channel, 1)
        self.in1 = InstanceNormAffine(in_channel)
        self.in2 = InstanceNormAffine(out_channel)
        self.conv_r1 = SNConv2d(in_channel, out_channel, conv_size, padding=padding_size)
        self.conv_r2 = SNConv2d(out_channel, out_channel, conv_size, padding=padding_size)

    def forward(self, x):
        w1 = self.conv_r1.effective_weight()
        if ops.norm_conv_ok(x, w1):
            # bf16 tensor-core path: norm -> LeakyReLU -> (upsample) -> conv as fused nodes (the normalised fp32 activations never exist);
            # the half-resolution 1x1 residual branch is added inside the second convolution's epilogue
            res = self.conv_l1(x)
            out = ops.norm_conv(x, self.in1.weight, self.in1.bias, w1, self.conv_r1.bias, self.conv_r1.kernel_size, 0, self.in1.eps, 0.2, upsample=True)
            w2 = self.conv_r2.effective_weight()
            fuse_add = w2.shape[1] % 32 == 0
            out = ops.norm_conv(out, self.in2.weight, self.in2.bias, w2, self.conv_r2.bias, self.conv_r2.kernel_size, 0, self.in2.eps, 0.2,
                                addend=res if fuse_add else None)
            return out if fuse_add else ops.AddFn.apply(out, ops.upsample2(res))
        out_res = ops.upsample2(self.conv_l1(x))
        out = self.in1(x, 0.2)
        out = ops.Conv2dFn.apply(out, w1, self.conv_r1.bias, self.conv_r1.kernel_size, ops.PRE_UPSAMPLE, 0.2)     # conv(upsample(.)): the 4x tensor only exists as the operand
        out = self.in2(out, 0.2)
        out = self.conv_r2(out)
        return ops.AddFn.apply(out, out_res)


class AdaResBlock2(nn.Module):
    """Reference model_blocks.py:776-814."""

    def __init__(self, channels, style_dim):
        super().__init__()
        self.style_dim = style_dim
        self.channels = channels
        self.lrelu = nn.LeakyReLU(0.2)
        self.lin1_mean = Linear(style_dim, channels)
        self.lin1_std = Linear(style_dim, channels)
        self.lin2_mean = Linear(style_dim, channels)
        self.lin2_std = Linear(style_dim, channels)
        self.conv1 = SNConv2d(channels, channels, 3, padding=1)
        self.conv2 = SNConv2d(channels, channels, 3, padding=1)

    def style_linears(self):
        return [self.lin1_mean, self.lin1_std, self.lin2_mean, self.lin2_std]

    def forward(self, x, style, styles=None):
        """`styles`: the four style projections if the caller already computed them (one batched GEMM for all blocks)."""
        mean_st1, std_st1, mean_st2, std_st2 = styles if styles is not None else [lin(style) for lin in self.style_linears()]
        out = self.conv1(x)
        w2 = self.conv2.effective_weight()
        n, c = out.shape[0], out.shape[-1]
        if ops.norm_conv_ok(out, w2):                 # ada_in -> LeakyReLU -> conv2 as one fused node (bf16 tensor-core path)
            out = ops.norm_conv(out, std_st1.reshape(n, c), mean_st1.reshape(n, c), w2, self.conv2.bias, self.conv2.kernel_size, 1, 1e-5, 0.2)
        else:
            out = ops.ada_in(out, mean_st1, std_st1, 1e-5, 0.2)
            out = ops.Conv2dFn.apply(out, w2, self.conv2.bias, self.conv2.kernel_size, ops.PRE_NONE, 0.2)
        out = ops.ada_in(out, mean_st2, std_st2, 1e-5)
        return ops.AddFn.apply(out, x)


class AdaResBlockUp2(nn.Module):
    """Reference model_blocks.py:817-865 (same 1x1-conv/upsample commutation as ResBlockUp)."""

    def __init__(self, in_channels, out_channels, style_dim, out_size=None, scale=2, conv_size=3, padding_size=1):
        super().__init__()
        assert out_size is None and scale == 2
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.style_dim = style_dim
        self.upsample = nn.Upsample(size=out_size, scale_factor=scale)
        self.lrelu = nn.LeakyReLU(0.2)
        self.lin1_mean = Linear(style_dim, in_channels)
        self.lin1_std = Linear(style_dim, in_channels)
        self.lin2_mean = Linear(style_dim, out_channels)
        self.lin2_std = Linear(style_dim, out_channels)
        self.conv_l1 = SNConv2d(in_channels, out_channels, 1)
        self.conv_r1 = SNConv2d(in_channels, out_channels, conv_size, padding=padding_size)
        self.conv_r2 = SNConv2d(out_channels, out_channels, conv_size, padding=padding_size)

    def style_linears(self):
        return [self.lin1_mean, self.lin1_std, self.lin2_mean, self.lin2_std]

    def forward(self, x, style, styles=None):
        mean_st1, std_st1, mean_st2, std_st2 = styles if styles is not None else [lin(style) for lin in self.style_linears()]
        w1 = self.conv_r1.effective_weight()
        n = x.shape[0]
        if ops.norm_conv_ok(x, w1):                   # same fusion as ResBlockUp with ada_in statistics
            res = self.conv_l1(x)
            out = ops.norm_conv(x, std_st1.reshape(n, -1), mean_st1.reshape(n, -1), w1, self.conv_r1.bias, self.conv_r1.kernel_size, 1, 1e-5, 0.2, upsample=True)
            w2 = self.conv_r2.effective_weight()
            fuse_add = w2.shape[1] % 32 == 0
            out = ops.norm_conv(out, std_st2.reshape(n, -1), mean_st2.reshape(n, -1), w2, self.conv_r2.bias, self.conv_r2.kernel_size, 1, 1e-5, 0.2,
                                addend=res if fuse_add else None)
            return out if fuse_add else ops.AddFn.apply(out, ops.upsample2(res))
        out_res = ops.upsample2(self.conv_l1(x))
        out = ops.ada_in(x, mean_st1, std_st1, 1e-5, 0.2)
        out = ops.Conv2dFn.apply(out, w1, self.conv_r1.bias, self.conv_r1.kernel_size, ops.PRE_UPSAMPLE, 0.2)
        out = ops.ada_in(out, mean_st2, std_st2, 1e-5, 0.2)
        out = self.conv_r2(out)
        return ops.AddFn.apply(out, out_res)


def batched_style_projections(blocks, style):
    """All `lin{1,2}_{mean,std}` projections of a list of AdaIN blocks (reference model_blocks.py:795-805, 836-846: four nn.Linear per
    block, all applied to the same style vectors) as ONE GEMM over the concatenated weights instead of 4 x len(blocks) tiny ones; the
    per-layer results (and, through autograd, the per-layer gradients) are slices of it.  -> list of 4-tuples, one per block."""
    lins = [lin for b in blocks for lin in b.style_linears()]
    sizes = [lin.out_features for lin in lins]
    total = sum(sizes)
    pad = (-total) % 8                                    # the tensor-core GEMM wants a multiple of 8 output features
    ws, bs = [lin.weight for lin in lins], [lin.bias for lin in lins]
    if pad:
        ws.append(style.new_zeros((pad, style.shape[-1])))
        bs.append(style.new_zeros((pad,)))
    y = ops.linear(style, ops.cat_params(ws), ops.cat_params(bs))
    outs = torch.split(y, sizes + ([pad] if pad else []), dim=-1)      # (split, not slices: its backward is one concatenation)
    return [tuple(outs[4 * i:4 * i + 4]) for i in range(len(blocks))]
