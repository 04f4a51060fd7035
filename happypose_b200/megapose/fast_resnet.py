"""Inference executor for the (unchanged) torch ResNet backbones of the coarse / refiner networks.

The networks stay the reference torch modules (BASELINE.json: "The ResNet refiner and coarse networks remain the
reference torch modules, run in bf16"): parameters, state_dict names and eval-mode semantics are untouched.  What this
file changes is HOW an eval-mode forward is issued on the GPU:

  * every BatchNorm2d is folded into the preceding convolution (w' = w * g / sqrt(v + eps), b' = beta - mean * g /
    sqrt(v + eps), computed once in float32) -- eval-mode BN is an affine map, so this is the same function;
  * conv + bias + ReLU and conv + bias + residual-add + ReLU are issued as ONE cuDNN call each
    (torch.cudnn_convolution_relu / torch.cudnn_convolution_add_relu), channels_last, in the module's compute dtype;
  * the 7x7 / stride 2 stem convolution is issued as the equivalent 4x4 / stride 1 convolution over the 2x2
    space-to-depth of the zero-padded input (4x deeper reduction per tap: several times faster on the tensor cores than
    a strided 9- or 27-channel stem); the space-to-depth bf16 tensor is written by libhpb200 straight from the float32
    network input (ops.pack_input_s2d_bf16), which also replaces the dtype / layout conversion pass;
  * the stem's 3x3 / stride 2 max-pool runs as libhpb200's bf16 NHWC streaming kernel (ops.maxpool3x3s2_bf16).

For a 576-hypothesis coarse batch this removes ~70 batch-norm / ReLU / add kernels per forward (about a third of the
network's device time).  The plain module path is kept for float32 parity runs and for CPU execution.
"""
from __future__ import annotations

import os

from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn


def fold_conv_bn(conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(weight, bias) in float32 of conv followed by eval-mode bn."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if bn is None:
        return w, b
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return w * scale.view(-1, 1, 1, 1), (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()


def _has_fused_ops(device: torch.device) -> bool:
    return device.type == "cuda" and hasattr(torch, "cudnn_convolution_relu") and hasattr(torch, "cudnn_convolution_add_relu")


class _Conv:
    """One folded convolution with its execution recipe."""

    def __init__(self, conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d], dtype, pad_in_to: int = 0):
        w, b = fold_conv_bn(conv, bn)
        if pad_in_to and w.shape[1] % pad_in_to:
            extra = pad_in_to - w.shape[1] % pad_in_to
            w = F.pad(w, (0, 0, 0, 0, 0, extra))
        self.c_in = w.shape[1]
        self.weight = w.to(dtype).contiguous(memory_format=torch.channels_last)
        self.bias = b.to(dtype).contiguous()
        self.stride, self.padding, self.dilation, self.groups = conv.stride, conv.padding, conv.dilation, conv.groups
        self.tc_ctx = None  # set by FoldedResNet for the convolutions libhpb200's tensor-core kernel serves
        self.bias_f32 = b.float().contiguous()

    def plain(self, x):
        return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)

    def relu(self, x, fused: bool):
        if fused:
            return torch.cudnn_convolution_relu(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        return F.relu_(self.plain(x))

    def add_relu(self, x, z, fused: bool):
        if self.tc_ctx is not None and x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last) \
                and z.is_contiguous(memory_format=torch.channels_last):
            # layer1's relu(conv2 + b + identity) on libhpb200's tcgen05 kernel: cuDNN only has a generic implicit-GEMM tile for
            # the conv+add+relu form of this shape (0.77 PFLOP/s; its plain conv+relu gets a weight-stationary kernel at 1.27)
            from .. import ops

            y = ops.conv3x3_bias_relu_bf16(self.tc_ctx, x, self.weight, self.bias_f32, z, self.weight_tc)
            if y is not None:
                return y
        if fused:
            return torch.cudnn_convolution_add_relu(x, self.weight, z, 1.0, self.bias, self.stride, self.padding, self.dilation, self.groups)
        return F.relu_(self.plain(x).add_(z))


def s2d_weight(w: torch.Tensor, c_padded: int) -> torch.Tensor:
    """[O,C,7,7] stem weight -> [O,c_padded,4,4] weight of the equivalent stride-1 convolution over the space-to-depth
    input: w'[o, (r*2+s)*Cs + c, a, b] = w[o, c, 2a+r, 2b+s] (tap index 7 = zero), Cs = c_padded / 4 channels reserved per
    sub-pixel (r,s), channels c >= C of every block zero."""
    O, C = w.shape[:2]
    assert c_padded % 4 == 0 and c_padded // 4 >= C
    Cs = c_padded // 4
    w8 = F.pad(w, (0, 1, 0, 1))                                  # [O,C,8,8]
    w8 = w8.reshape(O, C, 4, 2, 4, 2).permute(0, 3, 5, 1, 2, 4)    # [O,r,s,C,a,b]
    w8 = F.pad(w8, (0, 0, 0, 0, 0, Cs - C))                      # [O,r,s,Cs,a,b]
    return w8.reshape(O, c_padded, 4, 4)


def s2d_reference(x: torch.Tensor, c_padded: int) -> torch.Tensor:
    """Plain-torch statement of ops.pack_input_s2d_bf16 (used by the tests): z[n,(r*2+s)*Cs+c,I,J] = xpad[n,c,2I+r,2J+s]."""
    n, C, H, W = x.shape
    assert c_padded % 4 == 0 and c_padded // 4 >= C
    Cs = c_padded // 4
    xp = F.pad(x, (3, 3, 3, 3))
    z = xp.reshape(n, C, H // 2 + 3, 2, W // 2 + 3, 2).permute(0, 3, 5, 1, 2, 4)   # [n,r,s,C,I,J]
    z = F.pad(z, (0, 0, 0, 0, 0, Cs - C))
    return z.reshape(n, c_padded, H // 2 + 3, W // 2 + 3)


class FoldedResNet:
    """Eval-mode forward of a torchvision-style ResNet (BasicBlock) with BN folded and fused cuDNN epilogues.
    Returns what backbone(x) returns: [b, num_classes] after avgpool + fc.
    `ctx` (a happypose_b200 Context) enables the libhpb200 kernels for the stem (space-to-depth packing, max-pool)."""

    def __init__(self, net: nn.Module, dtype: torch.dtype, ctx=None):
        assert not net.training, "folding batch-norm needs eval mode"
        self.dtype = dtype
        self.ctx = ctx
        dev = net.conv1.weight.device
        self.fused = _has_fused_ops(dev)
        self.stem = _Conv(net.conv1, net.bn1, dtype, pad_in_to=8 if dev.type == "cuda" else 0)
        self.stem_s2d = None
        c1 = net.conv1
        if (ctx is not None and dtype == torch.bfloat16 and c1.kernel_size == (7, 7) and c1.stride == (2, 2)
                and c1.padding == (3, 3) and c1.dilation == (1, 1) and c1.groups == 1 and c1.in_channels <= 64):
            w, b = fold_conv_bn(c1, net.bn1)
            # cuDNN's bf16 tensor-core kernels want the reduction channels in multiples of 64: measured at b=576 the
            # 4x4 stem takes 3.7 ms with 40 channels, 2.2 ms with 48 and 1.6 ms with 64 (scripts/stem_bench.py)
            self.s2d_channels = (4 * c1.in_channels + 63) // 64 * 64
            s2d = _Conv.__new__(_Conv)
            s2d.c_in = self.s2d_channels
            s2d.weight = s2d_weight(w, self.s2d_channels).to(dtype).contiguous(memory_format=torch.channels_last)
            s2d.bias = b.to(dtype).contiguous()
            s2d.stride, s2d.padding, s2d.dilation, s2d.groups = (1, 1), (0, 0), (1, 1), 1
            s2d.tc_ctx = None
            self.stem_s2d = s2d
            # libhpb200's tcgen05 implicit GEMM serves the 64 -> 64 channel stem (hpb_stem_tc.cu): float32 folded bias, and the
            # mask of non-zero 16-channel weight slices (49 of 64 for a 7x7 kernel) so the empty ones are not multiplied
            # HPB200_TC_STEM=0 keeps the stem on cuDNN (A/B measurements)
            self.tc_stem = self.s2d_channels == 64 and s2d.weight.shape[0] == 64 and os.environ.get("HPB200_TC_STEM", "1") != "0"
            if self.tc_stem:
                from .. import ops

                s2d.bias_f32 = b.float().contiguous()
                s2d.k_slice_mask = ops.stem_k_slice_mask(s2d.weight)
        mp = net.maxpool
        self.fast_pool = (ctx is not None and dtype == torch.bfloat16 and isinstance(mp, nn.MaxPool2d)
                          and mp.kernel_size in (3, (3, 3)) and mp.stride in (2, (2, 2)) and mp.padding in (1, (1, 1))
                          and mp.dilation in (1, (1, 1)) and not mp.ceil_mode)
        self.maxpool = net.maxpool
        self.blocks: List[Tuple[_Conv, _Conv, Optional[_Conv]]] = []
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            for blk in layer:
                assert type(blk).__name__ == "BasicBlock", "only BasicBlock ResNets (ResNet-18/34) are folded"
                down = None
                c2 = _Conv(blk.conv2, blk.bn2, dtype)
                if blk.downsample is not None:
                    # out = relu(conv2(.) + b2 + down(x) + b_down): the projection's (folded) bias is moved into conv2's
                    # bias (summed in float32), so the 1x1 projection is a bias-free convolution -- no separate
                    # elementwise bias-add pass over the projected feature map
                    down = _Conv(blk.downsample[0], blk.downsample[1], dtype)
                    _, b2 = fold_conv_bn(blk.conv2, blk.bn2)
                    _, bd = fold_conv_bn(blk.downsample[0], blk.downsample[1])
                    c2.bias = (b2 + bd).to(dtype).contiguous()
                    down.bias = None
                    c2.bias_f32 = (b2 + bd).float().contiguous()
                # layer1 (64 -> 64, 3x3 / stride 1 / pad 1, no projection): the block's second convolution on the library's
                # tensor-core kernel.  HPB200_TC_LAYER1=0 keeps cuDNN (A/B measurements).
                cv = blk.conv2
                if (ctx is not None and dtype == torch.bfloat16 and down is None and cv.in_channels == 64 and cv.out_channels == 64
                        and cv.kernel_size == (3, 3) and cv.stride == (1, 1) and cv.padding == (1, 1) and cv.dilation == (1, 1)
                        and cv.groups == 1 and os.environ.get("HPB200_TC_LAYER1", "1") != "0"):
                    from .. import ops

                    c2.tc_ctx = ctx
                    c2.weight_tc = ops.conv3x3_residual_weight(c2.weight)
                self.blocks.append((_Conv(blk.conv1, blk.bn1, dtype), c2, down))
        # the 512-d feature (average pool + fc) is computed in float32: it feeds the pose / logit heads directly
        self.fc_w = net.fc.weight.detach().float()
        self.fc_b = net.fc.bias.detach().float()
        # snapshot signature: a later load_state_dict / fine-tuning step / .train() invalidates the fold (PosePredictor
        # checks `stale()` on every eager forward and re-folds)
        self._tracked = list(net.parameters()) + list(net.buffers())
        self._net = net
        self._signature = self._current_signature()

    def _current_signature(self):
        return (self._net.training, self._tracked[0].device, sum(t._version for t in self._tracked))

    def stale(self) -> bool:
        """True when the wrapped module changed since the fold (in-place weight updates, .train(), device move)."""
        return self._current_signature() != self._signature

    @property
    def in_channels(self) -> int:
        """Channels the stem expects (the module's input channels rounded up to a multiple of 8 on CUDA)."""
        return self.stem.c_in

    def _stem(self, x: torch.Tensor, fused: bool) -> torch.Tensor:
        if (self.stem_s2d is not None and x.dtype == torch.float32 and x.is_contiguous() and x.shape[2] % 2 == 0
                and x.shape[3] % 2 == 0):
            from .. import ops

            z = ops.pack_input_s2d_bf16(self.ctx, x, self.s2d_channels)  # fp32 planar -> bf16 NHWC space-to-depth
            return self._stem_s2d(z, fused)
        if x.shape[1] < self.stem.c_in:  # zero channels meet zero weights
            if self.ctx is not None and self.dtype == torch.bfloat16 and x.dtype == torch.float32 and x.is_contiguous():
                from .. import ops

                x = ops.pack_input_bf16(self.ctx, x, self.stem.c_in)
            else:
                x = F.pad(x, (0, 0, 0, 0, 0, self.stem.c_in - x.shape[1]))
        x = x.to(dtype=self.dtype, memory_format=torch.channels_last)
        return self.stem.relu(x, fused)

    def _stem_s2d(self, z: torch.Tensor, fused: bool) -> torch.Tensor:
        """relu(conv4x4(z) + b) of the space-to-depth input: the library's tensor-core kernel when it serves the shape, else cuDNN."""
        if getattr(self, "tc_stem", False) and z.dtype == torch.bfloat16 and z.is_contiguous(memory_format=torch.channels_last):
            from .. import ops

            y = ops.stem_conv4x4_relu_bf16(self.ctx, z, self.stem_s2d.weight, self.stem_s2d.bias_f32, self.stem_s2d.k_slice_mask)
            if y is not None:
                return y
        return self.stem_s2d.relu(z, fused)

    @property
    def accepts_s2d(self) -> bool:
        """True when the stem runs as the 4x4 convolution over the space-to-depth input, i.e. when a caller may hand in
        that tensor directly (ops.render_s2d_bf16) instead of the float32 planar network input."""
        return self.stem_s2d is not None

    def __call__(self, x: torch.Tensor, packed_s2d: bool = False) -> torch.Tensor:
        """x: the float32 planar network input, or (packed_s2d=True) its bf16 space-to-depth form [b,s2d_channels,H/2+3,W/2+3]."""
        fused = self.fused
        stem = self._stem_s2d if packed_s2d else self._stem
        try:
            y = stem(x, fused)
        except torch.cuda.OutOfMemoryError:
            raise
        except RuntimeError as exc:
            # only "this cuDNN build has no fused conv+bias+relu kernel for the dtype / shape" selects the plain
            # conv + relu_ recipe; anything else (launch failures, bad shapes) is a genuine error
            msg = str(exc).lower()
            if not fused or not any(t in msg for t in ("cudnn", "unsupported", "not supported", "no kernel", "unable to find")):
                raise
            self.fused = fused = False
            y = stem(x, fused)
        x = y
        if self.fast_pool and x.is_contiguous(memory_format=torch.channels_last) and x.shape[1] % 8 == 0:
            from .. import ops

            x = ops.maxpool3x3s2_bf16(self.ctx, x)
        else:
            x = self.maxpool(x)
        for c1, c2, down in self.blocks:
            identity = x if down is None else down.plain(x)
            x = c2.add_relu(c1.relu(x, fused), identity, fused)
        x = x.mean(dim=(2, 3), dtype=torch.float32)  # float32 accumulate and result: no bf16 rounding of the feature
        return F.linear(x, self.fc_w, self.fc_b)


def try_fold(backbone: nn.Module, dtype: torch.dtype, ctx=None) -> Optional[FoldedResNet]:
    """FoldedResNet for torchvision-style BasicBlock ResNets in eval mode, else None (the module is run as is)."""
    needed = ("conv1", "bn1", "maxpool", "layer1", "layer2", "layer3", "layer4", "fc")
    if backbone.training or not all(hasattr(backbone, n) for n in needed):
        return None
    try:
        return FoldedResNet(backbone, dtype, ctx)
    except AssertionError:
        return None
