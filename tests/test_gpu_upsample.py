"""PropagationNet / upsampling cascade on the GPU (csrc/upsample.cu + tcgen05 gather-GEMM, SURVEY.md
§8f.1) against the golden outputs of the unmodified reference (tests/golden/propagation.npz) and the
CPU oracle on fresh shapes. fp32 tolerance: the 3xTF32 contraction is fp32-grade, summation order
differs from MKL-DNN (1e-5 of the depth range); the nearest-upsampling indices must be exact."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = 'cuda'
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def mods():
    importlib.import_module('3dvnet_b200.build').build()
    return dict(ops=importlib.import_module('3dvnet_b200.ops'),
                up=importlib.import_module('3dvnet_b200.mv3d.subnetworks.upsampling'),
                lm=importlib.import_module('3dvnet_b200.mv3d.lightningmodel'))


def _golden():
    g = np.load(os.path.join(HERE, 'golden', 'propagation.npz'))
    t = {k: torch.from_numpy(g[k]) for k in g.files}
    p = {name: {k[len(name) + 1:]: v for k, v in t.items() if k.startswith(name + '.')} for name in ('quarter', 'half', 'full')}
    return t, p


def _net(mods, in_dim, params):
    m = mods['up'].PropagationNet(in_dim, 32)
    missing, unexpected = m.load_state_dict(params, strict=True)
    return m.to(DEV).eval()


def test_nearest_upsampling_indices_exact(mods):
    """the fused input kernel reproduces F.interpolate(mode='nearest') for non-integer ratios (56->64, 56->80)"""
    ops = mods['ops']
    for (h, w), (H, W) in (((56, 56), (64, 80)), ((8, 8), (12, 20)), ((64, 80), (128, 160)), ((7, 5), (31, 33))):
        d = torch.arange(2 * h * w, dtype=torch.float32).view(2, h, w)
        ref = F.interpolate(d.unsqueeze(1), (H, W), mode='nearest').squeeze(1)
        feats = torch.zeros(2, 3, H, W)
        x = torch.empty((2 * H * W, 32), device=DEV)
        up = torch.empty((2, H, W), device=DEV)
        ops.lib().call('dv3d_propagation_input', ops._p(feats.to(DEV)), 3, ops._p(d.to(DEV)), 2, h, w, H, W, 32, ops._p(x),
                       ops._p(up), ops._stream())
        np.testing.assert_array_equal(up.cpu().numpy(), ref.numpy())
        np.testing.assert_array_equal(x[:, 3].cpu().numpy(), ref.reshape(-1).numpy())
        assert float(x[:, 4:].abs().max()) == 0.0


def test_propagation_cascade_golden(mods):
    t, p = _golden()
    nq, nh, nf = _net(mods, 33, p['quarter']), _net(mods, 33, p['half']), _net(mods, 4, p['full'])
    with torch.no_grad():
        dq = nq.forward_from(t['feats_quarter'].to(DEV), t['depth'].to(DEV))
        dh = nh.forward_from(t['feats_half'].to(DEV), dq)
        df = nf.forward_from(t['images'].to(DEV), dh)
        # reference-shaped call: depth already upsampled by the caller (upsampling.py:23)
        dq2 = nq(t['feats_quarter'].to(DEV), t['up_quarter'].unsqueeze(1).to(DEV))
    np.testing.assert_allclose(dq.cpu().numpy(), t['ref_quarter'].numpy(), rtol=0, atol=2e-5)
    np.testing.assert_array_equal(dq2.cpu().numpy(), dq.cpu().numpy())
    np.testing.assert_allclose(dh.cpu().numpy(), t['ref_half'].numpy(), rtol=0, atol=4e-5)
    np.testing.assert_allclose(df.cpu().numpy(), t['ref_full'].numpy(), rtol=0, atol=6e-5)


@pytest.mark.parametrize('mode', ['tf32x3', 'f32'])
@pytest.mark.parametrize('n,in_dim,H,W', [(1, 33, 64, 80), (3, 4, 37, 53), (1, 4, 256, 320)])
def test_propagation_matches_oracle(mods, mode, n, in_dim, H, W):
    from oracle import upsample
    ops = mods['ops']
    ops.set_gemm_mode(mode)
    try:
        _, p = _golden()
        params = p['full'] if in_dim == 4 else p['quarter']
        net = _net(mods, in_dim, params)
        g = torch.Generator().manual_seed(H * W + n)
        feats = torch.randn(n, in_dim - 1, H, W, generator=g)
        depth = torch.rand(n, 1, H, W, generator=g) * 4 + 0.5
        with torch.no_grad():
            ref = upsample.propagation_net(feats, depth, params)
            got = net(feats.to(DEV), depth.to(DEV))
        np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=0, atol=3e-5)
    finally:
        ops.set_gemm_mode('tf32x3')


def test_pl3dvnet_upsample_cascade(mods):
    """PL3DVNet.upsample = eval-3dvnet.py:101-125 on the drop-in model"""
    from oracle import upsample
    cfg = dict(depth_start=0.5, depth_interval=0.05, n_intervals=16, size=(16, 16))
    torch.manual_seed(5)
    net = mods['lm'].PL3DVNet(cfg, cfg, 0.08, feat_dim=32, img_size=(64, 80)).eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(DEV)
    g = torch.Generator().manual_seed(6)
    depth = torch.rand(2, 16, 16, generator=g) * 4 + 0.5
    fq, fh, img = torch.randn(3, 32, 16, 20, generator=g), torch.randn(3, 32, 32, 40, generator=g), torch.randn(3, 3, 64, 80, generator=g)
    ref_idx = torch.tensor([0, 2])
    sub = lambda pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    with torch.no_grad():
        ref = upsample.upsample_cascade(depth, fq[ref_idx], fh[ref_idx], img[ref_idx], sub('refine_quarter.'),
                                        sub('refine_half.'), sub('refine_full.'))
        got = net.upsample(depth.to(DEV), ref_idx.to(DEV), fq.to(DEV), fh.to(DEV), img.to(DEV))
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=0, atol=6e-5)
