"""CPU: folding batch-norm into the convolutions (fast_resnet.FoldedResNet) is the same eval-mode function as the module."""
import torch

from happypose_b200.megapose.backbones import make_backbone
from happypose_b200.megapose.fast_resnet import fold_conv_bn, s2d_reference, s2d_weight, try_fold


def _randomise_bn(net, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def test_fold_conv_bn_is_exact_affine():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(5, 7, 3, padding=1, bias=False)
    bn = torch.nn.BatchNorm2d(7).eval()
    _randomise_bn(bn)
    w, b = fold_conv_bn(conv, bn)
    x = torch.randn(2, 5, 9, 11)
    with torch.no_grad():
        ref = bn(conv(x))
        got = torch.nn.functional.conv2d(x, w, b, padding=1)
    assert torch.allclose(ref, got, atol=1e-5)


def test_folded_resnet34_matches_module_fp32():
    torch.manual_seed(1)
    for n_in in (9, 27):
        net = make_backbone("vanilla_resnet34", n_in).eval()
        _randomise_bn(net, seed=n_in)
        folded = try_fold(net, torch.float32)
        assert folded is not None and folded.in_channels == n_in  # no channel padding on the CPU
        x = torch.randn(2, n_in, 64, 96)
        with torch.no_grad():
            ref = net(x)
            got = folded(x)
        assert got.shape == ref.shape == (2, 512)
        assert torch.allclose(ref, got, rtol=1e-3, atol=1e-3 * ref.abs().max().item())


def test_try_fold_declines_training_mode_and_other_architectures():
    net = make_backbone("vanilla_resnet34", 9)
    assert try_fold(net.train(), torch.float32) is None
    assert try_fold(make_backbone("resnet18", 6).eval(), torch.float32) is None  # pre-activation WideResNet: run as is


def test_space_to_depth_stem_is_the_same_convolution():
    """7x7 / stride 2 / pad 3 over C channels == 4x4 / stride 1 / pad 0 over the 2x2 space-to-depth of the padded input."""
    torch.manual_seed(2)
    for C, H, W in ((9, 24, 32), (27, 16, 20), (3, 8, 8)):
        w = torch.randn(16, C, 7, 7)
        b = torch.randn(16)
        x = torch.randn(2, C, H, W)
        cz = (4 * C + 7) // 8 * 8
        ref = torch.nn.functional.conv2d(x, w, b, stride=2, padding=3)
        got = torch.nn.functional.conv2d(s2d_reference(x, cz), s2d_weight(w, cz), b)
        assert got.shape == ref.shape == (2, 16, H // 2, W // 2)
        assert torch.allclose(ref, got, atol=1e-3)


def test_stem_k_slice_mask_of_a_7x7_kernel():
    """The space-to-depth form of a 7x7 kernel leaves 15 of the 64 (tap, 16-channel slice) blocks of the 4x4x64 weight empty:
    sub-pixel column s = 1 of the taps kw = 3 and sub-pixel row r = 1 of the taps kh = 3 (the 8th row / column of the footprint).
    ops.stem_k_slice_mask must report exactly the pattern hpb_stem_tc.cu's specialised kernel skips (bit 4 * (4 kh + kw) + k,
    k = 2 r + s), and a dense weight must report all 64."""
    from happypose_b200 import ops
    from happypose_b200.megapose.fast_resnet import s2d_weight

    g = torch.Generator().manual_seed(3)
    w = s2d_weight(torch.randn(64, 9, 7, 7, generator=g), 64)
    mask = ops.stem_k_slice_mask(w)
    expect = 0
    for kh in range(4):
        for kw in range(4):
            for k in range(4):
                r, s = k >> 1, k & 1
                if not ((kw == 3 and s == 1) or (kh == 3 and r == 1)):
                    expect |= 1 << (4 * (4 * kh + kw) + k)
    assert mask == expect and bin(mask).count("1") == 49 and mask & 1
    assert ops.stem_k_slice_mask(torch.randn(64, 64, 4, 4, generator=g)) == (1 << 64) - 1
    # the weight the kernel sees is the channels_last tensor: [O][kh][kw][C] in memory, slice k = channels 16k .. 16k+15
    wl = w.contiguous(memory_format=torch.channels_last)
    flat = wl.permute(0, 2, 3, 1).reshape(64, 16, 4, 16)
    for tap in range(16):
        for k in range(4):
            assert bool((flat[:, tap, k] != 0).any()) == bool((mask >> (4 * tap + k)) & 1)


def test_conv3x3_residual_weight_layout():
    """ops.conv3x3_residual_weight: [O][10 taps][C] = the nine taps of the [O,C,3,3] weight in (kh, kw) order, then a 64 x 64
    identity (tap 9) through which hpb_conv3x3_tc_kernel<RES> adds the residual tile on the tensor core."""
    from happypose_b200 import ops

    g = torch.Generator().manual_seed(5)
    w = torch.randn(64, 64, 3, 3, generator=g).to(torch.bfloat16)
    we = ops.conv3x3_residual_weight(w.contiguous(memory_format=torch.channels_last))
    assert we.shape == (64, 10, 64) and we.dtype == torch.bfloat16 and we.is_contiguous()
    for kh in range(3):
        for kw in range(3):
            assert torch.equal(we[:, 3 * kh + kw, :], w[:, :, kh, kw])
    assert torch.equal(we[:, 9, :].float(), torch.eye(64))
    # the same bytes the plain kernel reads for taps 0..8: the channels_last weight is [O][kh][kw][C] in memory
    flat = w.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1).reshape(64, 9, 64)
    assert torch.equal(we[:, :9], flat)
