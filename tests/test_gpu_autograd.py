"""Backward of the homography warp + variance (csrc/planesweep.cu: *_bwd_kernel, exposed as torch.autograd.Function
in 3dvnet_b200/mv3d/functional.py) against torch autograd through the ORACLE's restatement of the reference's own
differentiable formulation: F.grid_sample(bilinear, align_corners=True) on a no_grad grid + two scatter means
(/root/reference/mv3d/subnetworks/mvsnet.py:187-216, /root/reference/mv3d/lightningmodel.py:147-169,190-228).
Tolerance: 2e-5 of the largest gradient entry (fp32 atomics in unspecified order vs. a sequential fp32 CPU sum)."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'
TOL = 2e-5


@pytest.fixture(scope='module')
def mods():
    importlib.import_module('3dvnet_b200.build').build()
    return dict(fn=importlib.import_module('3dvnet_b200.mv3d.functional'),
                lm=importlib.import_module('3dvnet_b200.mv3d.lightningmodel'),
                synth=importlib.import_module('3dvnet_b200.synth'))


def _edges(kind, b):
    if kind == 'regular':
        return b.ref_src_edges
    # ragged + self edge + duplicate + more sources than one staging pass (EMAX = 8)
    e = [(0, 1)] + [(3, s) for s in (1, 2, 3, 3, 4)] + [(6, s) for s in range(13) if s != 6]
    return torch.tensor(e, dtype=torch.int64).t().contiguous()


@pytest.mark.parametrize('kind', ['regular', 'irregular'])
def test_planesweep_variance_backward_matches_autograd(kind, mods, exact_warp):
    import oracle.planesweep as o
    img, plane, D = (64, 80), (16, 24), 16
    b = mods['synth'].make_batch(1, 13, img, plane, 32, 2, 2, False, 7)
    e = _edges(kind, b)
    g = torch.Generator().manual_seed(1)
    n_ref = len(torch.unique(e[0]))
    G = torch.randn(n_ref, 32, D, *plane, generator=g)
    # reference gradient: CPU autograd through grid_sample + scatter mean
    f_cpu = b.feats_quarter.clone().requires_grad_(True)
    want_var = o.planesweep_var(f_cpu, b.rotmats, b.tvecs, b.K, e, 0.5, 0.3, D, img, plane)
    (want_var * G).sum().backward()
    # ours
    f_gpu = b.feats_quarter.to(DEV).requires_grad_(True)
    got_var = mods['fn'].planesweep_variance(f_gpu, b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV), e, 0.5, 0.3, D, plane,
                                             img)
    assert torch.equal(got_var.detach().cpu().view(torch.int32), want_var.detach().view(torch.int32))  # forward: bit-exact
    (got_var * G.to(DEV)).sum().backward()
    assert f_gpu.grad.shape == f_cpu.grad.shape
    err = (f_gpu.grad.cpu() - f_cpu.grad).abs().max().item()
    assert err <= TOL * f_cpu.grad.abs().max().item(), (err, f_cpu.grad.abs().max().item())
    # images that are nobody's source get exactly zero gradient
    unused = sorted(set(range(13)) - set(e[1].tolist()))
    for i in unused:
        assert float(f_gpu.grad[i].abs().max()) == 0.0


@pytest.mark.parametrize('n_side,offset', [(0, 0.0), (3, 0.05)])
def test_point_variance_backward_matches_autograd(n_side, offset, mods):
    import oracle.pointcloud as o
    img, plane = (64, 80), (16, 16)
    b = mods['synth'].make_batch(1, 9, img, plane, 32, 2, 2, True, 5)
    e = b.ref_src_edges
    ref_idx = torch.unique(e[0])
    depth = b.depth_images.clone()
    db = b.images_batch[ref_idx]
    f_cpu = b.feats_quarter.clone().requires_grad_(True)
    if n_side == 0:
        pts_o, feat_o, _ = o.feature_rich_pointcloud(depth, db, f_cpu, b.rotmats, b.tvecs, b.K, e, img)
        feat_o = feat_o.unsqueeze(1)
    else:
        pts_o, feat_o, _ = o.hypothesis_points(depth, db, f_cpu, b.rotmats, b.tvecs, b.K, e, offset, n_side, img)
    g = torch.Generator().manual_seed(2)
    G = torch.randn(feat_o.shape, generator=g)
    (feat_o * G).sum().backward()
    f_gpu = b.feats_quarter.to(DEV).requires_grad_(True)
    pts, feat = mods['fn'].point_variance(f_gpu, b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV), e, depth.to(DEV), img,
                                          n_side, offset)
    assert not pts.requires_grad and feat.requires_grad
    assert torch.equal(feat.detach().cpu().view(torch.int32), feat_o.detach().reshape(feat.shape).view(torch.int32))
    (feat * G.to(DEV).reshape(feat.shape)).sum().backward()
    err = (f_gpu.grad.cpu() - f_cpu.grad).abs().max().item()
    assert err <= TOL * f_cpu.grad.abs().max().item(), (err, f_cpu.grad.abs().max().item())


def test_mvsnet_trains_end_to_end(mods):
    """MVSNet.forward in training mode: images -> backbone/FPN (cuDNN) -> OUR warp + variance (forward and backward
    kernels) -> CostRegNet (cuDNN autograd) -> soft-argmin; the loss reaches the backbone, and the training
    composition agrees with the inference kernels on the same weights."""
    lm, synth = mods['lm'], mods['synth']
    img, plane, D = (64, 80), (16, 16), 16
    cfg = dict(depth_start=0.5, depth_interval=0.3, n_intervals=D, size=plane)
    b = synth.make_batch(1, 5, img, plane, 32, 1, 1, False, 2, with_images=True).to(DEV)
    net = lm.PL3DVNet(cfg, cfg, 0.3, feat_dim=32, img_size=img)
    net.load_state_dict(synth.make_params(0), strict=False)
    net = net.to(DEV)
    net.eval()
    with torch.no_grad():
        d_eval, _, fq_eval, _ = net.mvsnet(b, 0.5, 0.3, D, plane)
    # gradient mode with the SAME (eval) BatchNorm statistics: cuDNN regulariser (fp32: its TF32 default would
    # differ by 1e-3) vs our inference kernels
    fq = fq_eval.clone().requires_grad_(True)
    x_var = net.mvsnet.cost_volume(fq, b, 0.5, 0.3, D, plane)
    assert x_var.requires_grad
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        d_grad, _ = net.mvsnet.cnn_3d.depth(x_var, 0.5, 0.5 + 0.3 * (D - 1))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    np.testing.assert_allclose(d_grad.detach().cpu().numpy(), d_eval.cpu().numpy(), rtol=0, atol=2e-4)
    (d_grad - b.depth_images[:d_grad.shape[0]].to(DEV)).abs().mean().backward()
    assert fq.grad is not None and torch.isfinite(fq.grad).all() and float(fq.grad.abs().max()) > 0
    # full training-mode step through the backbone
    net.train()
    depth, fh, fq2, fe = net.mvsnet(b, 0.5, 0.3, D, plane)
    loss = (depth - b.depth_images[:depth.shape[0]]).abs().mean()
    loss.backward()
    grads = [p.grad for p in net.mvsnet.feat_shrinker.parameters() if p.grad is not None]
    assert grads and all(torch.isfinite(g).all() for g in grads) and any(float(g.abs().max()) > 0 for g in grads)
    assert net.mvsnet.cnn_3d.conv0.conv.weight.grad is not None
