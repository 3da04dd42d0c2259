"""The plane-sweep kernel's opt-in tolerance mode (DV3D_WARP=fast: affine-in-depth projection, reciprocal instead of
IEEE divisions) against the CPU oracle and against the kernel's default exact mode: x_var within 3e-5 of its scale
on average, max error bounded, identical zero-padding behaviour."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def mods():
    importlib.import_module('3dvnet_b200.build').build()
    return dict(ops=importlib.import_module('3dvnet_b200.ops'), synth=importlib.import_module('3dvnet_b200.synth'))


def _both(ops, f, cams, plan, d0, dd, D, plane, img):
    out = {}
    old = ops.warp_mode()
    try:
        for m in ('exact', 'fast'):
            ops.set_warp_mode(m)
            out[m] = ops.planesweep_var(f, cams, plan, d0, dd, D, plane, img)
    finally:
        ops.set_warp_mode(old)
    return out


@pytest.mark.parametrize('img,plane,D,n_imgs,nb,na,self_edge', [((64, 80), (16, 16), 16, 5, 2, 2, False),
                                                                ((64, 80), (16, 24), 16, 7, 1, 1, True),
                                                                ((256, 320), (56, 56), 96, 8, 4, 3, False),
                                                                ((256, 320), (64, 80), 96, 5, 2, 2, True)])
def test_fast_mode_matches_oracle_within_tolerance(mods, img, plane, D, n_imgs, nb, na, self_edge):
    import oracle.planesweep as o
    ops, synth = mods['ops'], mods['synth']
    b = synth.make_batch(1, n_imgs, img, plane, 32, nb, na, self_edge, 13)
    d0, dd = 0.5, (0.3 if D == 16 else 0.05)
    want = o.planesweep_var(b.feats_quarter, b.rotmats, b.tvecs, b.K, b.ref_src_edges, d0, dd, D, img, plane)
    plan = ops.edge_plan(b.ref_src_edges, torch.device(DEV))
    cams = ops.camera_tables(b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV))
    got = _both(ops, ops.nchw_to_nhwc(b.feats_quarter.to(DEV)), cams, plan, d0, dd, D, plane, img)
    assert torch.equal(got['exact'].cpu().view(torch.int32), want.view(torch.int32))
    err = (got['fast'].cpu() - want).abs()
    scale, peak = want.abs().mean().item(), want.abs().max().item()
    assert err.mean().item() <= 3e-5 * scale, err.mean().item() / scale
    assert err.max().item() <= 5e-4 * peak, err.max().item() / peak
    # same support: voxels that project outside every source view are exactly zero in both modes
    zero = want == 0
    if zero.any():
        assert (got['fast'].cpu()[zero].abs() <= 1e-6 * peak).all()


def test_exact_is_the_default_and_fast_is_deterministic(mods):
    import os
    ops, synth = mods['ops'], mods['synth']
    if not os.environ.get('DV3D_WARP'):
        assert ops.warp_mode() == 'exact'
    img, plane, D = (64, 80), (16, 16), 16
    b = synth.make_batch(1, 5, img, plane, 32, 2, 2, False, 3)
    plan = ops.edge_plan(b.ref_src_edges, torch.device(DEV))
    cams = ops.camera_tables(b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV))
    f = ops.nchw_to_nhwc(b.feats_quarter.to(DEV))
    old = ops.warp_mode()
    ops.set_warp_mode('fast')
    try:
        a = ops.planesweep_var(f, cams, plan, 0.5, 0.3, D, plane, img)
        assert torch.equal(a, ops.planesweep_var(f, cams, plan, 0.5, 0.3, D, plane, img))
    finally:
        ops.set_warp_mode(old)
