"""Pins oracle/ against the golden vectors recorded from the UNMODIFIED reference modules
(oracle/make_golden.py). CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_inputs
from oracle import planesweep, costreg, pipeline, pointcloud, scenemodel
from oracle.voxelize import voxelize

CASES = ['c1_tiny', 'c1_selfedge_2scenes']


def _params(synth, g):
    p = synth.make_params(int(g['seed']))
    assert synth.params_checksum(p) == pytest.approx(float(g['params_checksum']), rel=1e-12), \
        'seeded weights drifted from the ones the golden file was recorded with'
    return p


@pytest.mark.parametrize('name', CASES)
def test_path_a_matches_reference(name, synth):
    g = load_golden(name)
    t, cfg, img_size = golden_inputs(g)
    p = _params(synth, g)
    depth, x_var, x_reg = pipeline.initial_depth(t['feats_quarter'], t['rotmats'], t['tvecs'], t['K'],
                                                 t['ref_src_edges'], cfg, img_size, p, return_all=True)
    np.testing.assert_allclose(x_var.numpy(), g['ref_x_var'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(x_reg.numpy(), g['ref_x_reg'], rtol=0, atol=2e-5)
    np.testing.assert_allclose(depth.numpy(), g['ref_depth_init'], rtol=0, atol=2e-5)


def test_scalar_sampler_matches_grid_sample():
    g = load_golden('c1_tiny')
    t, cfg, img_size = golden_inputs(g)
    v = planesweep.planesweep_var_numpy(t['feats_quarter'], t['rotmats'], t['tvecs'], t['K'], t['ref_src_edges'],
                                        cfg['depth_start'], cfg['depth_interval'], cfg['n_intervals'], img_size,
                                        cfg['size'])
    np.testing.assert_allclose(v.numpy(), g['ref_x_var'], rtol=0, atol=2e-6)


def test_voxelize_bit_exact_adversarial():
    g = load_golden('voxelize_adversarial')
    a_pts, a_idx, a_batch, a_edges = voxelize(g['pts'], g['batch'], float(g['edge_len']))
    assert a_idx.dtype == np.int32 and a_batch.dtype == np.int64 and a_edges.dtype == np.int64
    np.testing.assert_array_equal(a_idx, g['ref_anchor_idx3d'])
    np.testing.assert_array_equal(a_batch, g['ref_anchor_batch'])
    np.testing.assert_array_equal(a_edges, g['ref_anchor_pts_edges'])
    np.testing.assert_array_equal(a_pts.view(np.int32), g['ref_anchor_pts'].view(np.int32))


@pytest.mark.parametrize('name', CASES)
def test_path_b_matches_reference(name, synth):
    g = load_golden(name)
    t, cfg, img_size = golden_inputs(g)
    p = _params(synth, g)
    edge_len = float(g['edge_len'])
    depth = torch.from_numpy(g['ref_depth_init']).clone()
    ref_idx = torch.unique(t['ref_src_edges'][0])
    depth_batch = t['images_batch'][ref_idx]
    args = (t['feats_quarter'], t['rotmats'], t['tvecs'], t['K'], t['ref_src_edges'])

    xs, mid = pipeline.model_scene(depth, depth_batch, *args, edge_len, img_size, p, return_all=True)
    np.testing.assert_allclose(mid['pts'].numpy(), g['ref_pts'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(mid['pts_feat'].numpy(), g['ref_pts_feat'], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(mid['pts_batch'].numpy(), g['ref_pts_batch'])
    np.testing.assert_array_equal(mid['anchor_idx3d'].numpy(), g['ref_anchor_idx3d'])
    np.testing.assert_array_equal(mid['anchor_batch'].numpy(), g['ref_anchor_batch'])
    np.testing.assert_array_equal(mid['anchor_pts_edges'].numpy(), g['ref_anchor_pts_edges'])
    np.testing.assert_array_equal(mid['anchor_pts'].numpy(), g['ref_anchor_pts'])
    np.testing.assert_allclose(mid['pointnet'].numpy(), g['ref_pointnet'], rtol=0, atol=1e-5)
    for li, lv in enumerate(xs):
        np.testing.assert_array_equal(lv['idx'].numpy(), g['ref_xs%d_idx' % li])
        np.testing.assert_array_equal(lv['batch'].numpy(), g['ref_xs%d_batch' % li])
        np.testing.assert_allclose(lv['pts'].numpy(), g['ref_xs%d_pts' % li], rtol=0, atol=1e-6)
        np.testing.assert_allclose(lv['feats'].numpy(), g['ref_xs%d_feats' % li], rtol=0, atol=5e-5)

    off = pipeline.run_pointflow(xs, depth, depth_batch, *args, float(g['offsets'][0][0]), 3, img_size, p)
    np.testing.assert_allclose(off.numpy(), g['ref_offset0'], rtol=0, atol=1e-5)


@pytest.mark.parametrize('name', CASES)
def test_refinement_schedule_matches_reference(name, synth):
    g = load_golden(name)
    t, cfg, img_size = golden_inputs(g)
    p = _params(synth, g)
    ref_idx = torch.unique(t['ref_src_edges'][0])
    out = pipeline.refine(torch.from_numpy(g['ref_depth_init']), t['images_batch'][ref_idx], t['feats_quarter'],
                          t['rotmats'], t['tvecs'], t['K'], t['ref_src_edges'], float(g['edge_len']), img_size, p,
                          offsets_list=g['offsets'].tolist())
    np.testing.assert_allclose(out.numpy(), g['ref_depth_final'], rtol=0, atol=5e-5)


def test_propagation_oracle_matches_reference_golden():
    """oracle/upsample.py against outputs of the unmodified reference PropagationNet cascade"""
    import os
    import numpy as np
    import torch
    from oracle import upsample
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'propagation.npz'))
    t = {k: torch.from_numpy(g[k]) for k in g.files}
    p = {name: {k[len(name) + 1:]: v for k, v in t.items() if k.startswith(name + '.')} for name in ('quarter', 'half', 'full')}
    with torch.no_grad():
        dq = upsample.propagation_net(t['feats_quarter'], torch.nn.functional.interpolate(
            t['depth'].unsqueeze(1), t['feats_quarter'].shape[-2:], mode='nearest'), p['quarter'])
        full = upsample.upsample_cascade(t['depth'], t['feats_quarter'], t['feats_half'], t['images'], p['quarter'],
                                         p['half'], p['full'])
    np.testing.assert_array_equal(dq.numpy(), g['ref_quarter'])
    np.testing.assert_array_equal(full.numpy(), g['ref_full'])


def test_fusion_oracle_matches_reference_golden():
    """oracle/fusion.py against the unmodified reference pointcloudfusion_custom.process_scene (CPU-patched .cuda())"""
    import os
    import numpy as np
    import torch
    from oracle import fusion
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'fusion.npz'))
    d, p, k = (torch.from_numpy(g[n]) for n in ('depth', 'poses', 'K'))
    pts, n_valid, valid = fusion.process_scene(d, p, k, 0.1, 3)
    np.testing.assert_array_equal(valid.numpy(), g['ref_valid'])
    fused = pts[valid.view(valid.shape[0], -1)].numpy()
    np.testing.assert_allclose(fused, g['ref_pts'], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(valid[0].numpy(), g['ref_valid0'])


def test_irregular_edge_lists_match_reference(synth):
    """ragged / shuffled / self-only / duplicated / 12-source edge lists: the oracle against the unmodified
    reference (oracle/make_golden.py::case_irregular_edges). tests/test_gpu_parity.py compares the CUDA kernels
    with the oracle bit for bit on edge lists of the same kinds."""
    g = load_golden('c1_irregular_edges')
    t, cfg, img_size = golden_inputs(g)
    p = _params(synth, g)
    e = t['ref_src_edges']
    assert e.shape == (2, 20) and torch.unique(e[0]).tolist() == g['ref_ref_idx'].tolist() == [0, 3, 6, 9, 11]
    depth, x_var, _ = pipeline.initial_depth(t['feats_quarter'], t['rotmats'], t['tvecs'], t['K'], e, cfg, img_size, p,
                                             return_all=True)
    np.testing.assert_allclose(x_var.numpy(), g['ref_x_var'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(depth.numpy(), g['ref_depth_init'], rtol=0, atol=2e-5)
    assert float(np.abs(g['ref_x_var'][3]).max()) == 0.0          # reference 9 has only its self-edge
    ref_depth = torch.from_numpy(g['ref_depth_init'])
    depth_batch = t['images_batch'][torch.unique(e[0])]
    pts, feat, batch = pointcloud.feature_rich_pointcloud(ref_depth, depth_batch, t['feats_quarter'], t['rotmats'],
                                                          t['tvecs'], t['K'], e, img_size)
    np.testing.assert_array_equal(pts.numpy().view(np.int32), g['ref_pts'].view(np.int32))
    np.testing.assert_allclose(feat.numpy(), g['ref_pts_feat'], rtol=0, atol=1e-6)
    np.testing.assert_array_equal(batch.numpy(), g['ref_pts_batch'])


def test_plane_crop_of_the_oracle_equals_the_full_volume(synth):
    """oracle.planesweep_var(planes=slice) - used by the C5 GPU crop test - is a bitwise sub-volume of the full one"""
    import oracle.planesweep as o
    b = synth.make_batch(1, 5, (64, 80), (16, 16), 32, 2, 2, False, 4)
    args = (b.feats_quarter, b.rotmats, b.tvecs, b.K, b.ref_src_edges, 0.5, 0.3, 16, (64, 80), (16, 16))
    full = o.planesweep_var(*args)
    for lo, hi in ((0, 8), (5, 9), (8, 16)):
        crop = o.planesweep_var(*args, planes=slice(lo, hi))
        assert torch.equal(full[:, :, lo:hi].contiguous().view(torch.int32), crop.view(torch.int32))
