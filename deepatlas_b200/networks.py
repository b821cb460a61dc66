"""Host-side mirror of the reference's network registry for the volumetric hot path.

Same class names, constructor signatures, ``forward`` contracts, ``weights_init()`` and -- because
checkpoints are the on-disk contract (models/base.py:98-108, strict=True) -- the same parameter /
buffer names and shapes as

  * ``UNet_generator`` / ``UNet_light``   lib/network_factory/unets.py:182-280, __init__.py:12-15
  * ``UNet``                              lib/network_factory/unets.py:70-179
  * ``VoxelMorphCVPR2018``                lib/network_factory/voxel_morph.py:18-101

``torch.nn.Conv3d`` / ``ConvTranspose3d`` / ``BatchNorm3d`` objects are used purely as parameter
containers (so names, shapes and the xavier RNG stream match the reference); their own ``forward`` is
never called -- every FLOP runs in ``libdeepatlas_b200.so`` through :mod:`deepatlas_b200.ops`.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from . import ops

_ACT = {"ReLU": (nn.ReLU, 0.0), "LeakyReLU": (nn.LeakyReLU, 0.01)}


_BN_DEFER = None   # list of pending running-statistics updates while a deferred_bn_updates context is open


class deferred_bn_updates:
    """Inside this context the BatchNorm layers of this module normalise with batch statistics as usual but leave
    ``running_mean`` / ``running_var`` / ``num_batches_tracked`` alone; ``apply()`` performs those updates afterwards, in
    call order, with a handful of multi-tensor launches.  For a pass that runs on a side stream next to another pass of
    the SAME network (the joint step's two segmentation passes): the buffers then see the two updates in sequence, as
    in the reference's two sequential calls, instead of racing."""

    def __init__(self):
        self.items = []

    def __enter__(self):
        global _BN_DEFER
        if _BN_DEFER is not None:
            raise RuntimeError("deferred_bn_updates does not nest")
        _BN_DEFER = self.items
        return self

    def __exit__(self, *exc):
        global _BN_DEFER
        _BN_DEFER = None

    def apply(self):
        items, self.items = self.items, []
        if not items:
            return
        mom = items[0][0].momentum if items[0][0].momentum is not None else 0.1
        eps = items[0][0].eps
        uniform = all((bn.momentum if bn.momentum is not None else 0.1) == mom and bn.eps == eps for bn, *_ in items)
        with torch.no_grad():
            if uniform:
                var = torch._foreach_pow([i for _, _, i, _ in items], -2.0)          # var + eps
                torch._foreach_sub_(var, eps)
                torch._foreach_mul_(var, [float(M) / float(M - 1) if M > 1 else 1.0 for *_, M in items])   # unbiased
                torch._foreach_lerp_([bn.running_mean for bn, *_ in items], [m for _, m, _, _ in items], mom)
                torch._foreach_lerp_([bn.running_var for bn, *_ in items], var, mom)
                nbt = [bn.num_batches_tracked for bn, *_ in items if bn.num_batches_tracked is not None]
                if nbt:
                    torch._foreach_add_(nbt, 1)
            else:
                for bn, mean, invstd, M in items:
                    m_ = bn.momentum if bn.momentum is not None else 0.1
                    var = (invstd.pow(-2) - bn.eps) * (float(M) / float(M - 1) if M > 1 else 1.0)
                    bn.running_mean.lerp_(mean, m_)
                    bn.running_var.lerp_(var, m_)
                    if bn.num_batches_tracked is not None:
                        bn.num_batches_tracked.add_(1)


def _bn_apply(bn: nn.BatchNorm3d, y, slope):
    training = bn.training or bn.running_mean is None
    if _BN_DEFER is not None and bn.training and bn.running_mean is not None:
        out, mean, invstd = ops.bn_act(y, bn.weight, bn.bias, None, None, training=True,
                                       momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps, slope=slope,
                                       return_stats=True)
        _BN_DEFER.append((bn, mean, invstd, y.numel() // y.shape[1]))
        return out
    if bn.training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return ops.bn_act(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, training=training,
                      momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps, slope=slope)


class _Block(nn.Sequential):
    """conv | deconv -> [BN] -> activation, with the reference's child names."""

    def __init__(self, children: OrderedDict, slope: float, kind: str):
        super().__init__(children)
        self._slope, self._kind = slope, kind  # plain attributes: no extra state_dict keys

    def forward(self, x, x2=None):  # noqa: D401 - x2 is the un-materialised skip concatenation
        mods = list(self.children())
        op = mods[0]
        bn = mods[1] if isinstance(mods[1], nn.BatchNorm3d) else None
        fused = None if bn is not None else self._slope
        if self._kind == "conv":
            y = ops.conv3d(x, op.weight, op.bias, x2=x2, transposed=False, stride=op.stride[0],
                           pad=op.padding[0], slope=fused)
        elif self._kind == "convT3":
            y = ops.conv3d(x, op.weight, op.bias, x2=x2, transposed=True, stride=1, pad=1, slope=fused)
        else:  # deconv k2 s2
            if x2 is not None:
                raise RuntimeError("deconv block takes a single source")
            y = ops.deconv_k2s2(x, op.weight, op.bias)
            if fused is not None:
                y = _LeakyFunction.apply(y, fused)
        if bn is not None:
            y = _bn_apply(bn, y, self._slope)
        return y


class _LeakyFunction(torch.autograd.Function):
    """Stand-alone activation for the (rare) deconv-without-BN configuration: y = leaky(x)."""

    @staticmethod
    def forward(ctx, x, slope):
        # identity batch-norm (mean 0, invstd 1) + activation through the same kernel
        C = x.shape[1]
        zero = torch.zeros(C, device=x.device)
        one = torch.ones(C, device=x.device)
        y = torch.empty_like(x)
        ops._lib.call("da_bn_act_fwd", ops._p(x), ops._p(zero), ops._p(one), None, None, x.shape[0], C,
                      x[0, 0].numel(), 1, float(slope), ops._p(y), ops._stream())
        ctx.save_for_backward(y)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        ops._lib.call("da_act_bwd", ops._p(dy), ops._p(y), float(ctx.slope), dy.numel(), ops._p(dx), ops._stream())
        return dx, None


def convBlock(in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=True, batchnorm=False,
              act="ReLU", named=True):
    """Mirror of unets.convBlock (lib/network_factory/unets.py:24-39)."""
    act_cls, slope = _ACT[act]
    names = ("conv", "BN", "nonlinear") if named else ("0", "1", "2")
    ch = OrderedDict()
    ch[names[0]] = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=bias)
    if batchnorm:
        ch[names[1]] = nn.BatchNorm3d(out_channels)
        ch[names[2]] = act_cls()
    else:
        ch[names[1]] = act_cls()
    return _Block(ch, slope, "conv")


def deconvBlock(in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, bias=True,
                batchnorm=False, act="ReLU", named=True):
    """Mirror of unets.deconvBlock (lib/network_factory/unets.py:42-58)."""
    act_cls, slope = _ACT[act]
    if (kernel_size, stride, padding, output_padding) == (2, 2, 0, 0):
        kind = "deconv"
    elif (kernel_size, stride, padding, output_padding) == (3, 1, 1, 0):
        kind = "convT3"
    else:
        raise NotImplementedError(
            f"deepatlas_b200: ConvTranspose3d k{kernel_size} s{stride} p{padding} is not on the hot path")
    names = ("deconv", "BN", "nonlinear") if named else ("0", "1", "2")
    ch = OrderedDict()
    ch[names[0]] = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                      output_padding=output_padding, bias=bias)
    if batchnorm:
        ch[names[1]] = nn.BatchNorm3d(out_channels)
        ch[names[2]] = act_cls()
    else:
        ch[names[1]] = act_cls()
    return _Block(ch, slope, kind)


def init_conv_weights(m):
    """unets.init_conv_weights (lib/network_factory/unets.py:61-67)."""
    if m.__class__.__name__.find("Conv") != -1:
        if m.weight is not None:
            nn.init.xavier_normal_(m.weight.data)
        if m.bias is not None:
            m.bias.data.zero_()


class _Conv1x1(nn.Conv3d):
    """The 1x1 output head (unets.py:250 / :98): an nn.Conv3d for its parameters, our kernel for its math."""

    def forward(self, x):
        return ops.conv3d(x, self.weight, self.bias, stride=1, pad=0)


class _MaxPool2(nn.MaxPool3d):
    def forward(self, x):
        return ops.maxpool2(x)


class _ConvK2S2(nn.Conv3d):
    """The strided down-sampler of maxpool=False (unets.py:231): Conv3d kernel 2, stride 2, no padding."""

    def forward(self, x):
        return ops.conv_k2s2(x, self.weight, self.bias)


class _UpsampleTrilinear2(nn.Upsample):
    """nn.Upsample(scale_factor=2, mode='trilinear') of upsample=True (unets.py:236)."""

    def forward(self, x):
        return ops.upsample_trilinear2(x)


def UNet_generator(encoders, decoders, act="ReLU", upsample=False, maxpool=True, res=False):
    """Class factory mirroring lib/network_factory/unets.py:182-280, all variants: ``upsample=True`` (trilinear x2
    instead of the k2 s2 deconvolution), ``maxpool=False`` (k2 s2 convolution instead of max-pooling) and
    ``res=True`` (``enc(x) + x`` / ``dec(cat) + x`` with torch's channel broadcasting)."""

    class UNetTemplate(nn.Module):
        def __init__(self, in_channel, n_classes, bias=False, BN=False):
            super().__init__()
            self.in_channel, self.n_classes = in_channel, n_classes
            self.levels = len(encoders)
            self.encoders = nn.ModuleList()
            self.decoders = nn.ModuleList()
            self.down_samplers = nn.ModuleList()
            self.up_samplers = nn.ModuleList()
            self.maxpool, self.upsample, self.res = maxpool, upsample, res
            enc = None
            for i, enc in enumerate(encoders):
                if i == 0:
                    enc = (in_channel,) + tuple(enc)
                self.encoders.append(nn.Sequential(*[convBlock(enc[k], enc[k + 1], bias=bias, batchnorm=BN, act=act)
                                                     for k in range(len(enc) - 1)]))
                if i < len(encoders) - 1:
                    self.down_samplers.append(_MaxPool2(2) if self.maxpool else
                                              _ConvK2S2(enc[-1], encoders[i + 1][0], kernel_size=2, stride=2, padding=0,
                                                        bias=bias))
            n_inner = len(enc) - 1  # the reference re-uses the last encoder tuple's length (unets.py:247)
            for i, dec in enumerate(decoders):
                if self.upsample:
                    self.up_samplers.append(_UpsampleTrilinear2(scale_factor=2, mode="trilinear"))
                else:
                    self.up_samplers.append(deconvBlock(encoders[-1][-1] if i == 0 else decoders[i - 1][-1], dec[0],
                                                        kernel_size=2, stride=2, bias=bias, batchnorm=BN, act=act))
                dec = (encoders[-(i + 2)][-1] + dec[0],) + tuple(dec[1:])
                blocks = [convBlock(dec[k], dec[k + 1], kernel_size=3, stride=1, padding=1, bias=bias, batchnorm=BN,
                                    act=act) for k in range(n_inner)]
                if i == len(decoders) - 1:
                    blocks.append(_Conv1x1(dec[-1], n_classes, kernel_size=1, stride=1, padding=0, bias=bias))
                self.decoders.add_module("decBlock{}".format(i), nn.Sequential(*blocks))

        def weights_init(self):
            self.apply(init_conv_weights)

        @property
        def head(self):
            """The 1x1x1 class head (unets.py:250): the last module of the last decoder block."""
            return list(self.decoders.children())[-1][-1]   # (the blocks are registered by name, not by index)

        def forward(self, x):
            return self._run(x, False)

        def forward_features(self, x):
            """Everything of ``forward`` up to (not including) the class head: the head's input, for callers that fuse
            head + softmax + Dice (``DiceLossMultiClass.forward_head``).  ``head(forward_features(x)) == forward(x)``."""
            if self.res:
                raise NotImplementedError("deepatlas_b200: res=True adds the decoder input to the logits (unets.py:275); "
                                          "the head is not separable")
            return self._run(x, True)

        def _run(self, x, stop_before_head):
            skips = []
            for i, enc in enumerate(self.encoders):
                y = x
                for blk in enc:
                    y = blk(y)
                x = ops.add(y, x) if self.res else y                      # unets.py:264
                if i < self.levels - 1:
                    skips.append(x)
                    x = self.down_samplers[i](x)
            for j, dec in enumerate(self.decoders):
                x = self.up_samplers[j](x)
                skip = skips.pop()
                y = x
                last = j == len(self.decoders) - 1
                for k, blk in enumerate(dec):
                    if stop_before_head and last and k == len(dec) - 1:
                        return y
                    y = blk(y, skip) if k == 0 else blk(y)  # first conv reads cat(x, skip) as two sources
                x = ops.add(y, x) if self.res else y                      # unets.py:275
            return x

    return UNetTemplate


UNet_light = UNet_generator(encoders=[(8, 16), (16, 16, 32), (32, 32, 64), (64, 64, 64)],
                            decoders=[(64, 64, 64), (64, 32, 32), (32, 16, 16)],
                            act="LeakyReLU", maxpool=True, upsample=False, res=False)
UNet_light.__name__ = UNet_light.__qualname__ = "UNet_light"


class UNet(nn.Module):
    """Mirror of lib/network_factory/unets.py:70-179 (32-channel base; decoders are ConvTranspose3d)."""

    def __init__(self, in_channel, n_classes, bias=False, BN=False):
        super().__init__()
        self.in_channel, self.n_classes = in_channel, n_classes
        e = lambda i, o: convBlock(i, o, bias=bias, batchnorm=BN, act="ReLU", named=False)  # noqa: E731
        d = lambda i, o, k, s, p=0: deconvBlock(i, o, k, stride=s, padding=p, bias=bias, batchnorm=BN,  # noqa: E731
                                                act="ReLU", named=False)
        self.ec0, self.ec1, self.ec2, self.ec3 = e(in_channel, 32), e(32, 64), e(64, 64), e(64, 128)
        self.ec4, self.ec5, self.ec6, self.ec7 = e(128, 128), e(128, 256), e(256, 256), e(256, 512)
        self.pool0, self.pool1, self.pool2 = _MaxPool2(2), _MaxPool2(2), _MaxPool2(2)
        self.dc9 = d(512, 512, 2, 2)
        self.dc8 = d(256 + 512, 256, 3, 1, 1)
        self.dc7 = d(256, 256, 3, 1, 1)
        self.dc6 = d(256, 256, 2, 2)
        self.dc5 = d(128 + 256, 128, 3, 1, 1)
        self.dc4 = d(128, 128, 3, 1, 1)
        self.dc3 = d(128, 128, 2, 2)
        self.dc2 = d(64 + 128, 64, 3, 1, 1)
        self.dc1 = d(64, 64, 3, 1, 1)
        self.dc0 = _Conv1x1(64, n_classes, kernel_size=1, stride=1, padding=0, bias=bias)

    def weights_init(self):
        self.apply(init_conv_weights)

    def forward(self, x):
        syn0 = self.ec1(self.ec0(x))
        syn1 = self.ec3(self.ec2(self.pool0(syn0)))
        syn2 = self.ec5(self.ec4(self.pool1(syn1)))
        e7 = self.ec7(self.ec6(self.pool2(syn2)))
        d7 = self.dc7(self.dc8(self.dc9(e7), syn2))
        d4 = self.dc4(self.dc5(self.dc6(d7), syn1))
        d1 = self.dc1(self.dc2(self.dc3(d4), syn0))
        return self.dc0(d1)


class convBlockVM(nn.Module):
    """Mirror of modules.convBlock (lib/network_factory/modules.py:28-62): attributes ``conv`` / ``bn``;
    conv -> [bn] -> ReLU (activation class instantiated per call in the reference) -> optional x += x."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False, batchnorm=False,
                 act=nn.ReLU, residual=False):
        super().__init__()
        if kernel_size != 3 or padding != 1 or stride not in (1, 2):
            raise NotImplementedError("deepatlas_b200: modules.convBlock is built for k3 p1 stride 1/2")
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=bias)
        self.bn = nn.BatchNorm3d(out_channels) if batchnorm else None
        if isinstance(act, str):
            act = _ACT[act][0]
        self.nonlinear = act
        self.residual = residual
        self._slope = None if act is None else (0.01 if act is nn.LeakyReLU else 0.0)

    def forward(self, x, x2=None):
        fused = self._slope if self.bn is None else None
        y = ops.conv3d(x, self.conv.weight, self.conv.bias, x2=x2, stride=self.conv.stride[0], pad=1, slope=fused)
        if self.bn is not None:
            y = _bn_apply(self.bn, y, self._slope)
        if self.residual:
            y = y + y  # the reference's `x += x` (modules.py:59-60), kept
        return y


class deconvBlockVM(nn.Module):
    """Mirror of modules.deconvBlock (lib/network_factory/modules.py:65-86): attributes ``deconv`` / ``bn``;
    ConvTranspose3d (k2 s2, or k3 s1 p1) -> [bn] -> activation -> optional ``x += input``."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, bias=False,
                 batchnorm=False, residual=False, act=nn.ReLU):
        super().__init__()
        if (kernel_size, stride, padding, output_padding) == (2, 2, 0, 0):
            self._kind = "deconv"
        elif (kernel_size, stride, padding, output_padding) == (3, 1, 1, 0):
            self._kind = "convT3"
        else:
            raise NotImplementedError(
                f"deepatlas_b200: ConvTranspose3d k{kernel_size} s{stride} p{padding} is not on the hot path")
        self.deconv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                         output_padding=output_padding, bias=bias)
        self.bn = nn.BatchNorm3d(out_channels) if batchnorm else None
        if isinstance(act, str):
            act = _ACT[act][0]
        self.nonlinear = act
        self.residual = residual
        self._slope = 0.01 if act is nn.LeakyReLU else 0.0

    def forward(self, input):
        fused = self._slope if self.bn is None else None
        if self._kind == "convT3":
            x = ops.conv3d(input, self.deconv.weight, self.deconv.bias, transposed=True, stride=1, pad=1, slope=fused)
        else:
            x = ops.deconv_k2s2(input, self.deconv.weight, self.deconv.bias)
            if fused is not None:
                x = _LeakyFunction.apply(x, fused)
        if self.bn is not None:
            x = _bn_apply(self.bn, x, self._slope)
        if self.residual:
            x = ops.add(x, input)
        return x


class VoxelMorphCVPR2018(nn.Module):
    """Mirror of lib/network_factory/voxel_morph.py:18-101.  ``forward(source, target)`` returns
    ``(disp_field, warped_source, deform_field)``; the identity grid is folded into the warp kernel."""

    def __init__(self, input_channel=2, output_channel=3, enc_filters=(16, 32, 32, 32, 32),
                 dec_filters=(32, 32, 32, 8, 8)):
        super().__init__()
        self.input_channel, self.output_channel = input_channel, output_channel
        self.enc_filters, self.dec_filters = enc_filters, dec_filters
        self.encoders = nn.ModuleList()
        self.decoders = nn.ModuleList()
        self.upsampling = nn.Upsample(scale_factor=2, mode="trilinear")  # unused in the reference forward too
        for i in range(len(enc_filters)):
            self.encoders.append(convBlockVM(input_channel if i == 0 else enc_filters[i - 1], enc_filters[i],
                                             stride=1 if i == 0 else 2, bias=True))
        for i in range(len(dec_filters)):
            if i == 0:
                cin = enc_filters[-1]
            elif i < 4:
                cin = dec_filters[i - 1] + enc_filters[4 - i]
            else:
                cin = dec_filters[i - 1]
            self.decoders.append(convBlockVM(cin, dec_filters[i], stride=1, bias=True))
        self.flow = nn.Conv3d(dec_filters[-1] + enc_filters[0], output_channel, kernel_size=3, stride=1, padding=1,
                              bias=True)
        self.id_transform = None

    def weights_init(self):
        self.apply(init_conv_weights)

    def forward(self, source, target):
        if self.input_channel == source.shape[1] + target.shape[1]:
            x1 = self.encoders[0](source, target)  # cat(source, target) read as two sources
        else:
            raise ValueError("VoxelMorphCVPR2018: source/target channels do not add up to input_channel")
        x2 = self.encoders[1](x1)
        x3 = self.encoders[2](x2)
        x4 = self.encoders[3](x3)
        x5 = self.encoders[4](x4)
        up = ops.upsample_nearest
        d1 = self.decoders[0](up(x5, x4.shape[2:]))
        d2 = self.decoders[1](up(d1, x3.shape[2:]), up(x4, x3.shape[2:]))
        d3 = self.decoders[2](up(d2, x2.shape[2:]), up(x3, x2.shape[2:]))
        d4 = self.decoders[3](d3, x2)
        d5 = self.decoders[4](up(d4, x1.shape[2:]))
        disp_field = ops.conv3d(d5, self.flow.weight, self.flow.bias, x2=x1, stride=1, pad=1)
        warped_source, deform_field = ops.warp3d(source, disp_field, add_identity=True, want_phi=True)
        return disp_field, warped_source, deform_field


network_dic = {
    "voxel_morph_cvpr": VoxelMorphCVPR2018,
    "UNet": UNet,
    "UNet_light": UNet_light,
}


def get_available_networks():
    return tuple(network_dic.keys())


def get_network(network_name):
    """lib/network_factory/__init__.py:19-23: KeyError listing the available names on a bad name."""
    if network_name in get_available_networks():
        return network_dic[network_name]
    raise KeyError("Network \"{}\" is not avaiable!\n Choose from: {}".format(network_name, get_available_networks()))
