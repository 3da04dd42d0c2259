"""CostRegNet layer kernels (csrc/costreg.cu) against plain PyTorch fp32 on the CPU
(mvsnet.py:18-36: Conv3d / ConvTranspose3d + eval-mode BatchNorm3d + ReLU [+ skip]) at shapes that
reach every kernel variant: the tiled stride-1 kernel at each tile depth, the direct kernel with
and without the K split, both channel-group widths, batches, volumes that are not tile multiples."""
import importlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    importlib.import_module('3dvnet_b200.build').build()
    return importlib.import_module('3dvnet_b200.ops')


def _case(seed, n, Cin, Cout, D, H, W, transposed=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, Cin, D, H, W, generator=g)
    wshape = (Cin, Cout, 3, 3, 3) if transposed else (Cout, Cin, 3, 3, 3)
    w = torch.randn(wshape, generator=g) / (27 * Cin) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.1
    return x, w, scale, shift


@pytest.mark.parametrize('n,Cin,Cout,D,H,W,stride', [
    (1, 32, 8, 48, 56, 56, 1),     # tiled kernel, tile depth 3 or 4
    (3, 32, 8, 40, 56, 56, 1),     # tiled kernel, more CTAs than one round
    (1, 8, 8, 40, 50, 60, 1),      # tiled kernel, ragged tiles in y and x
    (1, 16, 16, 48, 28, 28, 1),    # direct, 16-wide channel group, K split
    (2, 32, 32, 24, 14, 14, 1),    # direct, K split 4+
    (1, 64, 64, 12, 7, 7, 1),      # direct, 8-wide group, K split 8
    (1, 8, 16, 96, 56, 56, 2),     # stride 2, no split
    (1, 16, 32, 48, 28, 28, 2),
    (2, 32, 64, 24, 14, 14, 2),
    (1, 24, 8, 9, 11, 13, 1),      # odd sizes, Cin not a power of two
    (1, 8, 8, 9, 11, 13, 2),
])
def test_conv3d_bn_relu(ops, n, Cin, Cout, D, H, W, stride):
    x, w, scale, shift = _case(0, n, Cin, Cout, D, H, W)
    ref = F.relu(F.conv3d(x, w, stride=stride, padding=1) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    skip = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    got = ops.conv3d_bn_relu(x.cuda(), w.cuda(), scale.cuda(), shift.cuda(), stride)
    got_skip = ops.conv3d_bn_relu(x.cuda(), w.cuda(), scale.cuda(), shift.cuda(), stride, skip.cuda())
    tol = 2e-5 * max(1.0, float(ref.abs().max()))   # fp32, summation order differs from MKL-DNN
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=0, atol=tol)
    np.testing.assert_allclose(got_skip.cpu().numpy(), (ref + skip).numpy(), rtol=0, atol=tol)


@pytest.mark.parametrize('n,D,H,W', [
    (1, 9, 11, 13),        # smaller than one patch column / z segment in every direction
    (2, 17, 20, 31),       # ragged patches and z segments, two volumes
    (1, 96, 56, 56),       # BASELINE C2: 480 work items on 148 persistent CTAs
    (1, 8, 6, 14),         # exactly one interior patch
    (1, 1, 1, 1),
])
def test_first_layer_on_tcgen05(ops, n, D, H, W):
    """csrc/conv3d_tc.cu (32 -> 8, 3xTF32) against float64, and against the CUDA-core kernel it replaces"""
    x, w, scale, shift = _case(5, n, 32, 8, D, H, W)
    ref = F.relu(F.conv3d(x.double(), w.double(), padding=1) * scale.double().view(1, -1, 1, 1, 1)
                 + shift.double().view(1, -1, 1, 1, 1))
    assert ops.conv3d_mode() == 'tc'
    got = ops.conv3d_bn_relu(x.cuda(), w.cuda(), scale.cuda(), shift.cuda(), 1)
    again = ops.conv3d_bn_relu(x.cuda(), w.cuda(), scale.cuda(), shift.cuda(), 1)
    ops.set_conv3d_mode('ffma')
    try:
        ffma = ops.conv3d_bn_relu(x.cuda(), w.cuda(), scale.cuda(), shift.cuda(), 1)
    finally:
        ops.set_conv3d_mode('tc')
    tol = 4e-6 * max(1.0, float(ref.abs().max()))   # fp32-grade: 3xTF32 products, fp32 accumulation
    np.testing.assert_allclose(got.cpu().double().numpy(), ref.numpy(), rtol=0, atol=tol)
    np.testing.assert_allclose(ffma.cpu().double().numpy(), ref.numpy(), rtol=0, atol=tol)
    assert torch.equal(got, again)                  # deterministic (fixed summation order)


@pytest.mark.parametrize('n,Cin,Cout,D,H,W', [
    (1, 64, 32, 12, 7, 7), (1, 32, 16, 24, 14, 14), (1, 16, 8, 48, 28, 28), (2, 16, 8, 5, 6, 7), (1, 24, 16, 4, 4, 4),
])
def test_deconv3d_bn_relu(ops, n, Cin, Cout, D, H, W):
    x, w, scale, shift = _case(2, n, Cin, Cout, D, H, W, transposed=True)
    ref = F.relu(F.conv_transpose3d(x, w, stride=2, padding=1, output_padding=1) * scale.view(1, -1, 1, 1, 1)
                 + shift.view(1, -1, 1, 1, 1))
    skip = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3))
    got = ops.deconv3d_bn_relu(x.cuda(), w.cuda(), scale.cuda(), shift.cuda(), skip.cuda())
    tol = 2e-5 * max(1.0, float(ref.abs().max()))
    np.testing.assert_allclose(got.cpu().numpy(), (ref + skip).numpy(), rtol=0, atol=tol)


@pytest.mark.parametrize('n,Cin,D,H,W', [(1, 8, 96, 56, 56), (2, 8, 16, 16, 24), (1, 8, 40, 9, 13), (1, 16, 7, 5, 5)])
def test_prob_softargmin(ops, n, Cin, D, H, W):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(n, Cin, D, H, W, generator=g)
    w = torch.randn(1, Cin, 3, 3, 3, generator=g) * 0.5   # sharp softmax: the depth is sensitive
    bias = 0.3
    reg = F.conv3d(x, w, bias=torch.tensor([bias]), padding=1).squeeze(1)
    d0, d1 = 0.5, 0.5 + 0.05 * (D - 1)
    depth_values = torch.linspace(d0, d1, D).view(1, D, 1, 1)
    ref_depth = (F.softmax(-reg, dim=1) * depth_values).sum(1)   # mvsnet.py:220-227
    depth, got_reg = ops.prob_softargmin(x.cuda(), w.cuda(), bias, d0, d1, want_reg=True)
    depth2, none = ops.prob_softargmin(x.cuda(), w.cuda(), bias, d0, d1, want_reg=False)
    assert none is None
    np.testing.assert_allclose(got_reg.cpu().numpy(), reg.numpy(), rtol=0, atol=2e-5 * float(reg.abs().max()))
    np.testing.assert_allclose(depth.cpu().numpy(), ref_depth.numpy(), rtol=0, atol=2e-5)
    np.testing.assert_array_equal(depth.cpu().numpy(), depth2.cpu().numpy())
