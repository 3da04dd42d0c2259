"""GPU parity at the FULL sizes of BASELINE.json's configs, against the CPU oracle (not against the CUDA path
itself): the C2 step over 12 seeds - free-running AND teacher-forced stage by stage -, the reference's own default
configuration (8 cm voxels mv3d/config.py:22, 5-view windows that contain the reference itself
mv3d/dsets/dataset.py:133-137), a C3-shaped scene of 8 reference views, and a depth-plane crop of the C5
(512x640, D=192, 10 src) variance slab.

Pass criteria (north_star: depth within 1e-3 abs-rel, voxel indices bit-exact):

* TEACHER-FORCED (every stage started from the oracle's input of that stage - the map stage -> stage is continuous
  except at the voxelisation, so these bounds are tight and hold for EVERY seed):
    x_var bit-exact (default mode) and within 3e-5 (mean, of its scale) in the kernel's opt-in fast mode |
    initial depth <= 2e-5 abs-rel, <= 2e-4 m max | voxel tables bit-exact |
    sparse features <= 5e-4 of the level's max | PointFlow offsets <= 5e-5 m max per pass.
* FREE-RUNNING (2 x (scene model + 3 PointFlow passes) on its own depth): the schedule is NOT a continuous map - a
  1e-6 depth difference can move a point across a voxel face, which changes that voxel's PointNet input by up to
  the cell size. The reference algorithm does this to itself: the oracle against the oracle with its initial depth
  perturbed by 2e-6 (one CostRegNet rounding difference) differs by up to 3.7e-3 abs-rel on these seeds
  (profiles/r1_05_parity_seeds.md). Criterion per seed:
    median per-pixel relative error <= 2.5e-4 (half of the pixels within a quarter of the tolerance even when a
    voxel flipped in the first scene model and shifted the whole second iteration; measured 1e-7 .. 8e-5), AND
    abs-rel <= 1e-3, OR the seed is in CHAOTIC_SEEDS and abs-rel <= that seed's own oracle-vs-perturbed-oracle
    abs-rel (measured inside the test, asserted to be > 1e-3 - i.e. the reference would fail its own tolerance).
"""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'

IMG, PLANE, D = (256, 320), (56, 56), 96
CFG = dict(depth_start=0.5, depth_interval=0.05, n_intervals=D, size=PLANE)
OFFSETS = [[0.05, 0.05, 0.025], [0.05, 0.05, 0.025]]  # eval-3dvnet.py:23
# seeds whose ORACLE-vs-perturbed-ORACLE abs-rel exceeds 1e-3 (profiles/r1_05_parity_seeds.md, re-measured below):
# only these may exceed 1e-3 free-running, and only up to that self-sensitivity
CHAOTIC_SEEDS = (1, 2, 5, 6)

TF_XVAR_MEAN, TF_XVAR_MAX = 3e-5, 1e-3   # fast-mode x_var: mean error / mean |x_var|, max error / max |x_var|
TF_DEPTH0_ABSREL, TF_DEPTH0_MAX = 2e-5, 2e-4
TF_FEAT_REL = 5e-4
TF_OFFSET_MAX = 5e-5
FREE_MEDIAN_REL = 2.5e-4
FREE_ABSREL = 1e-3


class NS(object):
    pass


@pytest.fixture(scope='module')
def mods():
    importlib.import_module('3dvnet_b200.build').build()
    torch.set_num_threads(os.cpu_count() or 1)
    return dict(lm=importlib.import_module('3dvnet_b200.mv3d.lightningmodel'),
                ops=importlib.import_module('3dvnet_b200.ops'),
                synth=importlib.import_module('3dvnet_b200.synth'))


def make_net(mods, cfg, edge_len, img_size, params):
    net = mods['lm'].PL3DVNet(cfg, cfg, edge_len, feat_dim=32, img_size=img_size)
    net.load_state_dict(params, strict=False)
    return net.to(DEV).eval()


def oracle_trajectory(b, params, cfg, edge_len, img_size, offsets_list, d0=None):
    """the oracle's schedule with every stage's input and output kept"""
    from oracle import pipeline
    out = NS()
    args = (b.feats_quarter, b.rotmats, b.tvecs, b.K, b.ref_src_edges)
    with torch.no_grad():
        if d0 is None:
            out.d0, out.x_var, _ = pipeline.initial_depth(*args, cfg, img_size, params, return_all=True)
        else:
            out.d0 = d0
        ref_idx = torch.unique(b.ref_src_edges[0])
        out.db = b.images_batch[ref_idx]
        depth = out.d0.clone()
        out.outer = []
        for offsets in offsets_list:
            xs = pipeline.model_scene(depth, out.db, *args, edge_len, img_size, params)
            o = NS()
            o.depth_in, o.xs, o.passes = depth.clone(), xs, []
            for offset in offsets:
                off = pipeline.run_pointflow(xs, depth, out.db, *args, offset, 3, img_size, params)
                o.passes.append((depth.clone(), offset, off))
                depth = depth + off
            out.outer.append(o)
        out.final = depth
    return out


def abs_rel(got, ref):
    return (torch.abs(got - ref) / (ref + 1e-7)).mean().item()


def check_teacher_forced(net, b, tr, cfg):
    """every stage of the CUDA path on the ORACLE's input of that stage"""
    g = NS()
    g.feats_quarter, g.rotmats, g.tvecs, g.K = (b.feats_quarter.to(DEV), b.rotmats.to(DEV), b.tvecs.to(DEV),
                                                b.K.to(DEV))
    g.ref_src_edges = b.ref_src_edges
    gargs = (g.feats_quarter, g.rotmats, g.tvecs, g.K, g.ref_src_edges)
    db = tr.db.to(DEV)
    ops = importlib.import_module('3dvnet_b200.ops')
    with torch.no_grad():
        cv = lambda: net.mvsnet.cost_volume(g.feats_quarter, g, cfg['depth_start'], cfg['depth_interval'],
                                            cfg['n_intervals'], cfg['size'])
        mode = ops.warp_mode()
        ops.set_warp_mode('fast')       # the opt-in tolerance mode of the warp kernel (DV3D_WARP=fast)
        try:
            x_fast = cv()
        finally:
            ops.set_warp_mode(mode)
        err = (x_fast.cpu() - tr.x_var).abs()
        scale = tr.x_var.abs().mean().item()
        assert err.mean().item() <= TF_XVAR_MEAN * scale and err.max().item() <= TF_XVAR_MAX * tr.x_var.abs().max().item(), \
            (err.mean().item() / scale, err.max().item() / tr.x_var.abs().max().item())
        del x_fast
        x_var = cv()                    # default mode: the reference's operation chain, bit for bit
        if mode == 'exact':
            assert torch.equal(x_var.cpu().view(torch.int32), tr.x_var.view(torch.int32)), 'x_var is not bit-exact'
        d_end = cfg['depth_start'] + cfg['depth_interval'] * (cfg['n_intervals'] - 1)
        d0, _ = net.mvsnet.cnn_3d.depth(x_var, cfg['depth_start'], d_end)
        d0 = d0.cpu()
        assert abs_rel(d0, tr.d0) <= TF_DEPTH0_ABSREL, abs_rel(d0, tr.d0)
        assert (d0 - tr.d0).abs().max().item() <= TF_DEPTH0_MAX, (d0 - tr.d0).abs().max().item()
        for o in tr.outer:
            xs = net.model_scene(o.depth_in.to(DEV), db, *gargs)
            assert len(xs) == len(o.xs)
            for lg, lo in zip(xs, o.xs):
                assert torch.equal(lg['idx'].cpu(), lo['idx']), 'voxel indices differ'
                assert torch.equal(lg['batch'].cpu(), lo['batch'])
                err = (lg['feats'].cpu() - lo['feats']).abs().max().item() / lo['feats'].abs().max().item()
                assert err <= TF_FEAT_REL, err
            for depth_in, offset, off_o in o.passes:
                off = net.run_pointflow(xs, depth_in.to(DEV), db, *gargs, offset, 3).cpu()
                assert (off - off_o).abs().max().item() <= TF_OFFSET_MAX, (off - off_o).abs().max().item()


def free_running(net, b, cfg, offsets_list):
    with torch.no_grad():
        return net.hot_path(b.feats_quarter.to(DEV), b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV), b.ref_src_edges,
                            b.images_batch.to(DEV), cfg, offsets_list).cpu()


@pytest.mark.parametrize('seed', range(12))
def test_c2_step_matches_oracle(seed, mods):
    """BASELINE configs[1] (1 ref + 7 src, 256x320, D=96, 4 cm voxels), the bench workload, seeds 0..11"""
    from oracle import pipeline
    synth = mods['synth']
    b = synth.make_batch(1, 8, IMG, PLANE, 32, 4, 3, False, seed)
    params = synth.make_params(0)
    net = make_net(mods, CFG, 0.04, IMG, params)
    tr = oracle_trajectory(b, params, CFG, 0.04, IMG, OFFSETS)
    check_teacher_forced(net, b, tr, CFG)
    got = free_running(net, b, CFG, OFFSETS)
    rel = abs_rel(got, tr.final)
    med = torch.median((torch.abs(got - tr.final) / (tr.final + 1e-7)).flatten()).item()
    assert med <= FREE_MEDIAN_REL, (seed, med)
    if rel > FREE_ABSREL:
        assert seed in CHAOTIC_SEEDS, 'seed %d: abs-rel %.3e > 1e-3 and the seed is not a chaotic one' % (seed, rel)
        # how far does the reference algorithm move under an fp32-rounding-sized change of its own input?
        gen = torch.Generator().manual_seed(100 + seed)
        d0p = tr.d0 * (1.0 + 2e-6 * torch.randn(tr.d0.shape, generator=gen))
        with torch.no_grad():
            refp = pipeline.refine(d0p, tr.db, b.feats_quarter, b.rotmats, b.tvecs, b.K, b.ref_src_edges, 0.04, IMG,
                                   params, offsets_list=OFFSETS)
        sens = abs_rel(refp, tr.final)
        assert sens > FREE_ABSREL, 'seed %d is listed as chaotic but the oracle moves by only %.3e' % (seed, sens)
        assert rel <= sens, 'seed %d: abs-rel %.3e exceeds the oracle\'s own sensitivity %.3e' % (seed, rel, sens)


def test_chaotic_seeds_are_chaotic_in_the_oracle(mods):
    """the justification of CHAOTIC_SEEDS, checked on one of them: CPU oracle vs CPU oracle, 2e-6 apart at the
    input, disagree by more than the north-star tolerance - no implementation can meet 1e-3 there free-running"""
    from oracle import pipeline
    synth = mods['synth']
    seed = 6
    b = synth.make_batch(1, 8, IMG, PLANE, 32, 4, 3, False, seed)
    params = synth.make_params(0)
    args = (b.feats_quarter, b.rotmats, b.tvecs, b.K, b.ref_src_edges)
    with torch.no_grad():
        d0 = pipeline.initial_depth(*args, CFG, IMG, params)
        db = b.images_batch[torch.unique(b.ref_src_edges[0])]
        ref = pipeline.refine(d0, db, *args, 0.04, IMG, params, offsets_list=OFFSETS)
        gen = torch.Generator().manual_seed(100 + seed)
        refp = pipeline.refine(d0 * (1.0 + 2e-6 * torch.randn(d0.shape, generator=gen)), db, *args, 0.04, IMG, params,
                               offsets_list=OFFSETS)
    assert abs_rel(refp, ref) > FREE_ABSREL


@pytest.mark.parametrize('plane,n_ref', [((56, 56), 7), ((64, 80), 2)])
def test_reference_default_configuration(plane, n_ref, mods):
    """the reference's own defaults at full size: GRID_EDGE_LEN = 0.08 (mv3d/config.py:22), N_REF_IMGS = 7 with
    N_SRC_ON_EITHER_SIDE = 2 -> 5-edge windows that contain the reference itself (dataset.py:129-137),
    DEPTH_TEST plane 56x56 (config.py:27-32) and the full quarter-resolution plane 64x80"""
    synth = mods['synth']
    cfg = dict(CFG, size=plane)
    b = synth.make_batch(1, n_ref + 4, IMG, plane, 32, 2, 2, True, 21)
    assert b.ref_src_edges.shape[1] == 5 * n_ref and (b.ref_src_edges[0] == b.ref_src_edges[1]).sum() == n_ref
    params = synth.make_params(0)
    net = make_net(mods, cfg, 0.08, IMG, params)
    tr = oracle_trajectory(b, params, cfg, 0.08, IMG, OFFSETS)
    check_teacher_forced(net, b, tr, cfg)
    got = free_running(net, b, cfg, OFFSETS)
    rel = abs_rel(got, tr.final)
    med = torch.median((torch.abs(got - tr.final) / (tr.final + 1e-7)).flatten()).item()
    assert med <= FREE_MEDIAN_REL and rel <= FREE_ABSREL, (rel, med)


def test_c3_scene_of_8_reference_views(mods):
    """BASELINE configs[2] per GPU: one scene, 8 reference views + 7 halo keyframes, 7 sources each"""
    synth = mods['synth']
    b = synth.make_batch(1, 15, IMG, PLANE, 32, 4, 3, False, 3)
    assert len(torch.unique(b.ref_src_edges[0])) == 8
    params = synth.make_params(0)
    net = make_net(mods, CFG, 0.04, IMG, params)
    tr = oracle_trajectory(b, params, CFG, 0.04, IMG, OFFSETS)
    check_teacher_forced(net, b, tr, CFG)
    got = free_running(net, b, CFG, OFFSETS)
    med = torch.median((torch.abs(got - tr.final) / (tr.final + 1e-7)).flatten()).item()
    assert med <= FREE_MEDIAN_REL, med
    # 8 views per scene: 25 088 points, far more voxel-face crossings than one view - the free-running criterion
    # is the oracle's own sensitivity (measured here), bounded below by the north-star tolerance
    from oracle import pipeline
    gen = torch.Generator().manual_seed(103)
    with torch.no_grad():
        refp = pipeline.refine(tr.d0 * (1.0 + 2e-6 * torch.randn(tr.d0.shape, generator=gen)), tr.db, b.feats_quarter,
                               b.rotmats, b.tvecs, b.K, b.ref_src_edges, 0.04, IMG, params, offsets_list=OFFSETS)
    assert abs_rel(got, tr.final) <= max(FREE_ABSREL, abs_rel(refp, tr.final))


def test_c5_variance_slab_crop_is_bit_exact(mods, exact_warp):
    """BASELINE configs[4] (512x640, D=192, 10 src): depth planes 88..95 and 184..191 of the CUDA slab against the
    oracle evaluated on exactly those planes of the full 192-plane linspace"""
    import oracle.planesweep as ops_a
    ops, synth = mods['ops'], mods['synth']
    img, Dn, plane, n_src = (512, 640), 192, (112, 112), 10
    b = synth.make_batch(1, 1 + n_src, img, plane, 32, 5, 5, False, 0)
    plan = ops.edge_plan(b.ref_src_edges, torch.device(DEV))
    cams = ops.camera_tables(b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV))
    got = ops.planesweep_var(ops.nchw_to_nhwc(b.feats_quarter.to(DEV)), cams, plan, 0.5, 0.05, Dn, plane, img)
    for lo in (88, 184):
        want = ops_a.planesweep_var(b.feats_quarter, b.rotmats, b.tvecs, b.K, b.ref_src_edges, 0.5, 0.05, Dn, img, plane,
                                    planes=slice(lo, lo + 8))
        assert want.shape == (1, 32, 8) + plane
        assert torch.equal(got[:, :, lo:lo + 8].cpu().view(torch.int32), want.view(torch.int32))
