"""3D-UNet -- mirror of the reference's ``components/unet3d.py`` (itself vendored from wolny/pytorch-3dunet): same
public names, constructor signatures, module tree and therefore the same ``state_dict`` keys
(``encoders.{i}.basic_module.SingleConv{1,2}.{groupnorm,conv}.*``, ``decoders.{i}...``, ``final_conv.*``).

Execution is different: ``Abstract3DUNet.forward`` keeps activations channels-last (NDHWC) end to end and runs every
``SingleConv`` of order 'gcr' as  GroupNorm-statistics kernel -> one implicit-GEMM convolution kernel that applies the
GroupNorm affine while loading its operand and ReLU in its epilogue (``gnb_groupnorm_stats`` + ``gnb_conv3d_k3``);
max-pool, nearest-upsample + concat and the final 1x1x1 convolution are single kernels as well.  The result is
returned as a logical NCDHW tensor in ``channels_last_3d`` memory format (no transposes at the boundary).

Only the configuration the GarmentNets pipeline instantiates is executable (``basic_module=DoubleConv``, layer order
'gcr', nearest upsampling, ref networks/conv_implicit_wnf.py:109-113); the other symbols of the reference file
exist for API compatibility and refuse to run instead of falling back to another backend.
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from .. import ops


# tensor-core (TMA + tcgen05) convolutions; False selects the fp32 FFMA kernel everywhere (used by the parity tests)
USE_TENSOR_CORES = True
FUSE_UPSAMPLE_CONCAT = True   # decoders: GroupNorm passes read (skip, low-res) directly instead of a materialised upsample + concat
USE_STACKED_DX = True    # Cout in {32, 64} layers: stacked-kw tensor-core kernel (gnb_conv3d_tc_dx)


def number_of_features_per_level(init_channel_number, num_levels):
    return [init_channel_number * 2 ** k for k in range(num_levels)]


def conv3d(in_channels, out_channels, kernel_size, bias, padding=1):
    return nn.Conv3d(in_channels, out_channels, kernel_size, padding=padding, bias=bias)


def create_conv(in_channels, out_channels, kernel_size, order, num_groups, padding=1):
    """(name, module) pairs of one conv layer; ``order`` is a string over g(roupnorm) b(atchnorm) c(onv) r(elu)
    l(eaky relu) e(lu).  A conv gets a bias only when no norm layer is present (ref components/unet3d.py:50-53)."""
    assert 'c' in order, "Conv layer MUST be present"
    assert order[0] not in 'rle', 'Non-linearity cannot be the first operation in the layer'
    conv_at = order.index('c')
    has_norm = any(ch in order for ch in 'gb')
    out = []
    for pos, ch in enumerate(order):
        norm_channels = in_channels if pos < conv_at else out_channels
        if ch == 'c':
            out.append(('conv', conv3d(in_channels, out_channels, kernel_size, not has_norm, padding=padding)))
        elif ch == 'g':
            groups = num_groups if norm_channels >= num_groups else 1
            assert norm_channels % groups == 0, (
                f'Expected number of channels in input to be divisible by num_groups. '
                f'num_channels={norm_channels}, num_groups={groups}')
            out.append(('groupnorm', nn.GroupNorm(num_groups=groups, num_channels=norm_channels)))
        elif ch == 'b':
            out.append(('batchnorm', nn.BatchNorm3d(norm_channels)))
        elif ch == 'r':
            out.append(('ReLU', nn.ReLU(inplace=True)))
        elif ch == 'l':
            out.append(('LeakyReLU', nn.LeakyReLU(negative_slope=0.1, inplace=True)))
        elif ch == 'e':
            out.append(('ELU', nn.ELU(inplace=True)))
        else:
            raise ValueError(f"Unsupported layer type '{ch}'. MUST be one of ['b', 'g', 'r', 'l', 'e', 'c']")
    return out


def _unsupported(what):
    raise NotImplementedError(f"{what} is not on the GarmentNets inference hot path; garmentnets_b200 implements "
                              "DoubleConv with layer order 'gcr' and has no fallback backend")


class SingleConv(nn.Sequential):
    def __init__(self, in_channels, out_channels, kernel_size=3, order='crg', num_groups=8, padding=1):
        super().__init__()
        for name, module in create_conv(in_channels, out_channels, kernel_size, order, num_groups, padding=padding):
            self.add_module(name, module)
        self.order = order
        self._fusable = order in ('gcr', 'gc') and kernel_size == 3 and padding == 1

    def packed_weight(self) -> torch.Tensor:
        """Conv weight [Cout,Cin,3,3,3] re-laid as [27, Cin, Cout] (tap-major, Cout contiguous); cached."""
        w = self.conv.weight
        key = (w._version, w.data_ptr())
        cached = getattr(self, '_gnb_wt', None)
        if cached is None or cached[0] != key:
            with torch.no_grad():
                wt = w.permute(2, 3, 4, 1, 0).reshape(27, w.shape[1], w.shape[0]).contiguous()
            cached = (key, wt)
            self._gnb_wt = cached
        return cached[1]

    def forward_ndhwc(self, x: torch.Tensor) -> torch.Tensor:
        if not self._fusable or self.conv.bias is not None:
            _unsupported(f"SingleConv(order='{self.order}')")
        gn = self.groupnorm
        scale, shift = ops.groupnorm_stats(x, gn.num_groups, gn.eps, gn.weight, gn.bias)
        B, D, H, W, Cin = x.shape
        Cout = self.conv.out_channels
        if USE_TENSOR_CORES and ops.conv3d_tc_supported(B, D, H, W, Cin, Cout):
            # TMA + tcgen05 implicit GEMM on the normalised activation written once as fp16 hi + lo
            xh, xl = ops.gn_apply_split(x, scale, shift)
            if USE_STACKED_DX and ops.conv3d_tc_dx_supported(B, D, H, W, Cin, Cout):
                # narrow layer: the three kw taps share one activation box (shift applied to the output)
                return ops.conv3d_tc_dx(xh, xl, Cin, self.packed_weight_tc(dx=True), Cout, relu='r' in self.order)
            return ops.conv3d_tc(xh, xl, Cin, self.packed_weight_tc(), Cout, relu='r' in self.order)
        if USE_TENSOR_CORES and Cout % 128 == 0 and Cout > 128 and ops.conv3d_tc_supported(B, D, H, W, Cin, 128):
            # wide layer (the 4^3 level's 128 -> 256): Cout / 128 launches of the 128-column tensor-core kernel on slices of the
            # weight (the fp32 SIMT fallback took 0.31 ms for 5 GFLOP); the slices are concatenated along the channel axis
            xh, xl = ops.gn_apply_split(x, scale, shift)
            parts = [ops.conv3d_tc(xh, xl, Cin, w, 128, relu='r' in self.order) for w in self.packed_weight_tc_slices(128)]
            return torch.cat(parts, dim=-1)
        return ops.conv3d_k3(x, self.packed_weight(), scale, shift, relu='r' in self.order)

    def packed_weight_tc_slices(self, width: int):
        w = self.conv.weight
        key = (w._version, w.data_ptr(), width, ops.conv_tc_cross_precision())
        cached = getattr(self, '_gnb_wt_tc_slices', None)
        if cached is None or cached[0] != key:
            cached = (key, [ops.conv3d_tc_pack_weights(w[o:o + width].contiguous()) for o in range(0, w.shape[0], width)])
            self._gnb_wt_tc_slices = cached
        return cached[1]

    def forward_ndhwc_cat(self, skip: torch.Tensor, x_low: torch.Tensor) -> torch.Tensor:
        """``forward_ndhwc(cat((skip, nearest_upsample_2x(x_low)), channel))`` -- the decoder's joining (ref
        components/unet3d.py:291,325-330) -- without writing the concatenated tensor: the GroupNorm statistics and the
        normalise-and-split pass read the two sources directly."""
        B, D, H, W, Cs = skip.shape
        Cx = x_low.shape[-1]
        Cin, Cout = Cs + Cx, self.conv.out_channels
        gn = self.groupnorm
        ok = (self._fusable and self.conv.bias is None and USE_TENSOR_CORES and FUSE_UPSAMPLE_CONCAT and Cs % 4 == 0 and Cx % 4 == 0
              and (Cin // gn.num_groups) % 4 == 0 and tuple(x_low.shape[1:4]) == (D // 2, H // 2, W // 2) and D % 2 == 0 and H % 2 == 0
              and W % 2 == 0 and skip.is_contiguous() and x_low.is_contiguous() and ops.conv3d_tc_supported(B, D, H, W, Cin, Cout))
        if not ok:
            return self.forward_ndhwc(ops.upsample_concat(skip, x_low))
        scale, shift = ops.groupnorm_stats_cat(skip, x_low, gn.num_groups, gn.eps, gn.weight, gn.bias)
        xh, xl = ops.gn_apply_split_cat(skip, x_low, scale, shift)
        if USE_STACKED_DX and ops.conv3d_tc_dx_supported(B, D, H, W, Cin, Cout):
            return ops.conv3d_tc_dx(xh, xl, Cin, self.packed_weight_tc(dx=True), Cout, relu='r' in self.order)
        return ops.conv3d_tc(xh, xl, Cin, self.packed_weight_tc(), Cout, relu='r' in self.order)

    def packed_weight_tc(self, dx: bool = False) -> torch.Tensor:
        w = self.conv.weight
        key = (w._version, w.data_ptr(), ops.conv_tc_cross_precision())   # the packed image depends on the cross-term format
        attr = '_gnb_wt_tc_dx' if dx else '_gnb_wt_tc'
        cached = getattr(self, attr, None)
        if cached is None or cached[0] != key:
            cached = (key, ops.conv3d_tc_dx_pack_weights(w) if dx else ops.conv3d_tc_pack_weights(w))
            setattr(self, attr, cached)
        return cached[1]

    def forward(self, x):
        out = self.forward_ndhwc(ops.to_channels_last(x))
        return out.permute(0, 4, 1, 2, 3)


class DoubleConv(nn.Sequential):
    """Two SingleConvs.  Encoder: c1 = max(out//2, in) channels in between; decoder: in->out, out->out
    (ref components/unet3d.py:122-137)."""

    def __init__(self, in_channels, out_channels, encoder, kernel_size=3, order='crg', num_groups=8):
        super().__init__()
        mid = max(out_channels // 2, in_channels) if encoder else out_channels
        self.add_module('SingleConv1', SingleConv(in_channels, mid, kernel_size, order, num_groups))
        self.add_module('SingleConv2', SingleConv(mid, out_channels, kernel_size, order, num_groups))

    def forward_ndhwc(self, x):
        return self.SingleConv2.forward_ndhwc(self.SingleConv1.forward_ndhwc(x))

    def forward(self, x):
        return self.forward_ndhwc(ops.to_channels_last(x)).permute(0, 4, 1, 2, 3)


class ExtResNetBlock(nn.Module):
    """Residual block of the reference file (:147-192); never instantiated by GarmentNets.  Constructible, not runnable."""

    def __init__(self, in_channels, out_channels, kernel_size=3, order='cge', num_groups=8, **kwargs):
        super().__init__()
        self.conv1 = SingleConv(in_channels, out_channels, kernel_size=kernel_size, order=order, num_groups=num_groups)
        self.conv2 = SingleConv(out_channels, out_channels, kernel_size=kernel_size, order=order, num_groups=num_groups)
        n_order = ''.join(ch for ch in order if ch not in 'rel')
        self.conv3 = SingleConv(out_channels, out_channels, kernel_size=kernel_size, order=n_order, num_groups=num_groups)
        if 'l' in order:
            self.non_linearity = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        elif 'e' in order:
            self.non_linearity = nn.ELU(inplace=True)
        else:
            self.non_linearity = nn.ReLU(inplace=True)

    def forward(self, x):
        _unsupported("ExtResNetBlock")


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, conv_kernel_size=3, apply_pooling=True, pool_kernel_size=(2, 2, 2),
                 pool_type='max', basic_module=DoubleConv, conv_layer_order='crg', num_groups=8):
        super().__init__()
        assert pool_type in ['max', 'avg']
        if not apply_pooling:
            self.pooling = None
        elif pool_type == 'max':
            self.pooling = nn.MaxPool3d(kernel_size=pool_kernel_size)
        else:
            self.pooling = nn.AvgPool3d(kernel_size=pool_kernel_size)
        self.basic_module = basic_module(in_channels, out_channels, encoder=True, kernel_size=conv_kernel_size,
                                         order=conv_layer_order, num_groups=num_groups)

    def forward_ndhwc(self, x):
        if self.pooling is not None:
            k = self.pooling.kernel_size
            if not isinstance(self.pooling, nn.MaxPool3d) or tuple(k if isinstance(k, (tuple, list)) else (k,) * 3) != (2, 2, 2):
                _unsupported("pooling other than MaxPool3d(2)")
            x = ops.maxpool3d_2(x)
        return self.basic_module.forward_ndhwc(x)

    def forward(self, x):
        return self.forward_ndhwc(ops.to_channels_last(x)).permute(0, 4, 1, 2, 3)


class Upsampling(nn.Module):
    def __init__(self, transposed_conv, in_channels=None, out_channels=None, kernel_size=3, scale_factor=(2, 2, 2),
                 mode='nearest'):
        super().__init__()
        self.transposed_conv = transposed_conv
        self.mode = mode
        if transposed_conv:
            self.upsample = nn.ConvTranspose3d(in_channels, out_channels, kernel_size=kernel_size, stride=scale_factor,
                                               padding=1)
        else:
            self.upsample = partial(self._interpolate, mode=mode)

    @staticmethod
    def _interpolate(x, size, mode):
        _unsupported("stand-alone Upsampling (it is fused with the skip concat)")

    def forward(self, encoder_features, x):
        return self.upsample(x, encoder_features.size()[2:])


class Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, scale_factor=(2, 2, 2), basic_module=DoubleConv,
                 conv_layer_order='crg', num_groups=8, mode='nearest'):
        super().__init__()
        concat = basic_module == DoubleConv
        self.upsampling = Upsampling(transposed_conv=not concat, in_channels=in_channels, out_channels=out_channels,
                                     kernel_size=kernel_size, scale_factor=scale_factor, mode=mode)
        self.joining = partial(self._joining, concat=concat)
        if not concat:
            in_channels = out_channels
        self.basic_module = basic_module(in_channels, out_channels, encoder=False, kernel_size=kernel_size,
                                         order=conv_layer_order, num_groups=num_groups)
        self._concat = concat

    @staticmethod
    def _joining(encoder_features, x, concat):
        _unsupported("stand-alone joining (it is fused with the upsampling)")

    def forward_ndhwc(self, encoder_features, x):
        if not self._concat or self.upsampling.mode != 'nearest':
            _unsupported("Decoder with transposed-conv / summation joining")
        # nearest upsample to the skip's size + cat((encoder_features, x), channel): read in place by the first SingleConv's
        # GroupNorm passes (forward_ndhwc_cat); one materialising kernel (gnb_upsample_concat) otherwise
        bm = self.basic_module
        if isinstance(bm, DoubleConv):
            return bm.SingleConv2.forward_ndhwc(bm.SingleConv1.forward_ndhwc_cat(encoder_features, x))
        return bm.forward_ndhwc(ops.upsample_concat(encoder_features, x))

    def forward(self, encoder_features, x):
        y = self.forward_ndhwc(ops.to_channels_last(encoder_features), ops.to_channels_last(x))
        return y.permute(0, 4, 1, 2, 3)


class FinalConv(nn.Sequential):
    def __init__(self, in_channels, out_channels, kernel_size=3, order='crg', num_groups=8):
        super().__init__()
        self.add_module('SingleConv', SingleConv(in_channels, in_channels, kernel_size, order, num_groups))
        self.add_module('final_conv', nn.Conv3d(in_channels, out_channels, 1))

    def forward(self, x):
        _unsupported("FinalConv")


class Abstract3DUNet(nn.Module):
    def __init__(self, in_channels, out_channels, final_sigmoid, basic_module, f_maps=64, layer_order='gcr',
                 num_groups=8, num_levels=4, is_segmentation=False, testing=False, **kwargs):
        super().__init__()
        self.testing = testing
        if isinstance(f_maps, int):
            f_maps = number_of_features_per_level(f_maps, num_levels=num_levels)
        f_maps = list(f_maps)
        self.encoders = nn.ModuleList([
            Encoder(in_channels if i == 0 else f_maps[i - 1], f, apply_pooling=i > 0, basic_module=basic_module,
                    conv_layer_order=layer_order, num_groups=num_groups)
            for i, f in enumerate(f_maps)])
        rev = f_maps[::-1]
        self.decoders = nn.ModuleList([
            Decoder(rev[i] + rev[i + 1] if basic_module == DoubleConv else rev[i], rev[i + 1], basic_module=basic_module,
                    conv_layer_order=layer_order, num_groups=num_groups)
            for i in range(len(rev) - 1)])
        self.final_conv = nn.Conv3d(f_maps[0], out_channels, 1)
        if is_segmentation:
            self.final_activation = nn.Sigmoid() if final_sigmoid else nn.Softmax(dim=1)
        else:
            self.final_activation = None

    def forward_ndhwc(self, x: torch.Tensor, apply_final: bool = True) -> torch.Tensor:
        """``apply_final=False`` stops before ``final_conv`` (the fast tier folds that 1x1x1 convolution into the
        implicit decoders' first Linear: two affine maps in a row are one)."""
        skips = []
        for enc in self.encoders:
            x = enc.forward_ndhwc(x)
            skips.append(x)
        for dec, skip in zip(self.decoders, reversed(skips[:-1])):
            x = dec.forward_ndhwc(skip, x)
        if not apply_final:
            return x
        # 1x1x1 convolution with bias == per-voxel linear layer on channels-last data
        B, D, H, W, C = x.shape
        w = self.final_conv.weight.view(self.final_conv.out_channels, C)
        y = ops.linear(x.view(-1, C), w, self.final_conv.bias, relu=False)
        return y.view(B, D, H, W, -1)

    def forward(self, x):
        if self.training:
            _unsupported("training-mode forward")
        if self.testing and self.final_activation is not None:
            _unsupported("final_activation")
        y = self.forward_ndhwc(ops.to_channels_last(x))
        return y.permute(0, 4, 1, 2, 3)  # logical NCDHW, channels_last_3d strides


class UNet3D(Abstract3DUNet):
    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order='gcr', num_groups=8,
                 num_levels=4, is_segmentation=True, **kwargs):
        super().__init__(in_channels=in_channels, out_channels=out_channels, final_sigmoid=final_sigmoid,
                         basic_module=DoubleConv, f_maps=f_maps, layer_order=layer_order, num_groups=num_groups,
                         num_levels=num_levels, is_segmentation=is_segmentation, **kwargs)


class ResidualUNet3D(Abstract3DUNet):
    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order='gcr', num_groups=8,
                 num_levels=5, is_segmentation=True, **kwargs):
        super().__init__(in_channels=in_channels, out_channels=out_channels, final_sigmoid=final_sigmoid,
                         basic_module=ExtResNetBlock, f_maps=f_maps, layer_order=layer_order, num_groups=num_groups,
                         num_levels=num_levels, is_segmentation=is_segmentation, **kwargs)
