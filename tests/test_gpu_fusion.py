"""Depth-map fusion kernel (csrc/fusion.cu) against the golden output of the unmodified
reference (tests/golden/fusion.npz) and the CPU oracle. The consistency test is a threshold on fp32
re-projections whose operation order differs from torch's bmm, so a pixel within ~1e-6 of the 0.1 m
threshold (or of a nearest-neighbour rounding boundary) may flip: the vote counts must agree on at least
99.8 % of the pixels and 99.8 % of the fused positions to 1e-4 m (a flipped nearest-neighbour sample with the
same vote count moves a point by centimetres; bounded by 5 cm)."""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def fus():
    importlib.import_module('3dvnet_b200.build').build()
    return importlib.import_module('3dvnet_b200.mv3d.eval.pointcloudfusion_custom')


def _compare(pts, n_valid, valid, ref_pts, ref_n, ref_valid):
    same = (n_valid.cpu().long() == ref_n.long())
    assert same.float().mean().item() > 0.998
    assert (valid.cpu() != ref_valid).float().mean().item() < 0.002
    err = (pts.cpu() - ref_pts).abs().max(dim=-1)[0][same]
    assert (err < 1e-4).float().mean().item() > 0.998   # a flipped nearest-neighbour sample moves a point by ~cm
    assert float(err.max()) < 0.05


def test_fusion_golden(fus):
    from oracle import fusion
    g = np.load(os.path.join(HERE, 'golden', 'fusion.npz'))
    d, p, k = (torch.from_numpy(g[n]) for n in ('depth', 'poses', 'K'))
    ref_pts, ref_n, ref_valid = fusion.process_scene(d, p, k, 0.1, 3)   # pinned to the reference by the CPU suite
    pts, n_valid, valid = fus.fuse(d.to(DEV), p.to(DEV), k.to(DEV), d.shape[0], 0.1, 3)
    _compare(pts, n_valid, valid, ref_pts, ref_n, ref_valid)
    # reference-shaped entry points
    fused_pts, fused_rgb, all_valid = fus.process_scene(d, torch.from_numpy(g['images']), p, k, 0.1, 3)
    assert all_valid.shape == g['ref_valid'].shape and fused_pts.shape[1] == 3 and fused_rgb.shape == fused_pts.shape
    assert abs(fused_pts.shape[0] - g['ref_pts'].shape[0]) <= 0.002 * g['ref_valid'].size
    idx = torch.arange(d.shape[0]) != 0
    p0, rgb0, v0 = fus.process_depth(d[0], torch.from_numpy(g['images'][0]), d[idx], None, p[0], p[idx], k[0], k[idx], 0.1, 3)
    assert (v0 != g['ref_valid0']).mean() < 0.002 and p0.shape[0] == int(v0.sum())


@pytest.mark.parametrize('n,size,thresh,votes', [(3, (24, 32), 0.1, 1), (12, (60, 80), 0.05, 4), (70, (30, 40), 0.1, 3)])
def test_fusion_matches_oracle(fus, n, size, thresh, votes):
    from oracle import fusion
    synth = importlib.import_module('3dvnet_b200.synth')
    R, t, K = synth.make_cameras(n, size, seed=n)
    depth = synth.ray_box_depth(R, t, K, size, size)
    depth = depth + np.random.RandomState(n).normal(0, 0.02, depth.shape).astype(np.float32)
    poses = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    poses[:, :3, :3], poses[:, :3, 3] = R, t
    d, p, k = torch.from_numpy(depth), torch.from_numpy(poses), torch.from_numpy(K)
    ref_pts, ref_n, ref_valid = fusion.process_scene(d, p, k, thresh, votes)
    pts, n_valid, valid = fus.fuse(d.to(DEV), p.to(DEV), k.to(DEV), n, thresh, votes)
    _compare(pts, n_valid, valid, ref_pts, ref_n, ref_valid)
