"""Anchors the MinkowskiEngine restatement (oracle/minkowski_cpu.py, parity unpinned
against ME itself) on dense torch operators: a generalised sparse convolution must equal
the dense convolution of the zero-filled grid at its active sites, and the sparse trilinear
interpolation must equal trilinear grid_sample of the zero-padded dense volume — the
formulation the reference itself spells out in HypothesisDecoder.forward_forloop
(/root/reference/mv3d/subnetworks/refinement.py:54-93). CPU only."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import minkowski_cpu as mk


def _random_cloud(seed, n=300, extent=12, batches=2):
    rng = np.random.RandomState(seed)
    c = np.concatenate([rng.randint(0, batches, (n, 1)), rng.randint(0, extent, (n, 3))], axis=1)
    c = np.unique(c, axis=0)
    return mk.CoordMap(c, 1)


def _densify(cmap, feat, extent, batches):
    s = cmap.stride
    vol = torch.zeros(batches, feat.shape[1], extent, extent, extent)
    c = cmap.coords
    vol[c[:, 0], :, c[:, 3] // s, c[:, 2] // s, c[:, 1] // s] = feat
    return vol


def _dense_weight(W):
    # W[k] with k = kx + 3 ky + 9 kz  ->  [Cout, Cin, kz, ky, kx]
    return W.view(3, 3, 3, W.shape[1], W.shape[2]).permute(4, 3, 0, 1, 2).contiguous()


def test_conv_stride1_equals_dense():
    torch.manual_seed(0)
    cm = _random_cloud(0)
    f = torch.randn(len(cm), 5)
    W = torch.randn(27, 5, 7)
    out, om = mk.conv3(f, cm, W, 1)
    dense = F.conv3d(_densify(cm, f, 12, 2), _dense_weight(W), padding=1)
    c = cm.coords
    np.testing.assert_allclose(out.numpy(), dense[c[:, 0], :, c[:, 3], c[:, 2], c[:, 1]].numpy(), atol=1e-4)


def test_conv_stride2_equals_dense_and_coords():
    torch.manual_seed(1)
    cm = _random_cloud(1)
    f = torch.randn(len(cm), 4)
    W = torch.randn(27, 4, 6)
    out, om = mk.conv3(f, cm, W, 2)
    assert om.stride == 2
    expect = np.unique(np.concatenate([cm.coords[:, :1], cm.coords[:, 1:] // 2 * 2], axis=1), axis=0)
    assert sorted(map(tuple, om.coords)) == sorted(map(tuple, expect))
    dense = F.conv3d(_densify(cm, f, 12, 2), _dense_weight(W), stride=2, padding=1)
    c = om.coords
    np.testing.assert_allclose(out.numpy(), dense[c[:, 0], :, c[:, 3] // 2, c[:, 2] // 2, c[:, 1] // 2].numpy(),
                               atol=1e-4)


def test_conv_transpose_equals_dense():
    torch.manual_seed(2)
    fine = _random_cloud(2)
    coarse = fine.strided(2)
    f = torch.randn(len(coarse), 6)
    W = torch.randn(27, 6, 3)
    out = mk.conv3_transpose(f, coarse, fine, W)
    wd = W.view(3, 3, 3, 6, 3).permute(3, 4, 0, 1, 2).contiguous()  # [Cin, Cout, kz, ky, kx]
    dense = F.conv_transpose3d(_densify(coarse, f, 6, 2), wd, stride=2, padding=1, output_padding=1)
    c = fine.coords
    np.testing.assert_allclose(out.numpy(), dense[c[:, 0], :, c[:, 3], c[:, 2], c[:, 1]].numpy(), atol=1e-4)


def test_interpolation_equals_dense_trilinear():
    torch.manual_seed(3)
    for stride in (1, 2, 4):
        cm = _random_cloud(3 + stride, n=200, extent=8, batches=1)
        cm = mk.CoordMap(np.concatenate([cm.coords[:, :1], cm.coords[:, 1:] * stride], axis=1), stride)
        f = torch.randn(len(cm), 5)
        q = torch.rand(500, 3) * (8 * stride + 2) - 1.0   # includes queries outside the occupied box
        qb = torch.cat([torch.zeros(500, 1), q], dim=1)
        out = mk.interpolate(cm, f, qb)
        # dense volume with one voxel of zero padding on each side (refinement.py:61-71)
        n = 8 + 3
        vol = torch.zeros(1, 5, n, n, n)
        c = cm.coords
        vol[0, :, c[:, 1] // stride + 1, c[:, 2] // stride + 1, c[:, 3] // stride + 1] = f.t()
        g = (q / stride + 1.0) / (n - 1) * 2 - 1
        g = g[None, None, None][..., [2, 1, 0]]   # grid_sample wants (z, y, x) for a volume indexed [x][y][z]
        ref = F.grid_sample(vol, g, 'bilinear', padding_mode='zeros', align_corners=True)[0, :, 0, 0].t()
        np.testing.assert_allclose(out.numpy(), ref.numpy(), atol=2e-5)


def test_kernel_offset_order_is_x_fastest():
    assert tuple(mk.OFFSETS[0]) == (-1, -1, -1) and tuple(mk.OFFSETS[1]) == (0, -1, -1)
    assert tuple(mk.OFFSETS[3]) == (-1, 0, -1) and tuple(mk.OFFSETS[9]) == (-1, -1, 0)
    assert tuple(mk.OFFSETS[13]) == (0, 0, 0)
