"""Diagnostic (GPU box): where does the CUDA path drift from the CPU oracle on the bench
workload? Teacher-forced per-stage errors and free-running accumulated error.
    python tools/diag_parity.py [seed]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pipeline  # noqa: E402

import bench  # noqa: E402


def stats(name, got, ref):
    d = (got - ref).abs().flatten()
    rel = (d / (ref.flatten().abs() + 1e-7))
    q = torch.quantile(d, torch.tensor([0.5, 0.9, 0.99, 0.999]))
    print('%-34s mean|d| %.3e  absrel %.3e  q50 %.1e q90 %.1e q99 %.1e q999 %.1e max %.1e  frac>1e-3 %.4f  '
          'bit-equal %.6f' % (name, d.mean(), rel.mean(), q[0], q[1], q[2], q[3], d.max(),
                              (d > 1e-3).float().mean(), (d == 0).float().mean()))


def main():
    torch.set_num_threads(os.cpu_count())
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    synth = importlib.import_module('3dvnet_b200.synth')
    arg = sys.argv[1] if len(sys.argv) > 1 else '0'
    if arg.isdigit():
        b, params = bench.synth_inputs(int(arg), 1)
        cfg, img, edge_len, offsets_list = bench.DEPTH_CFG, bench.IMG_SIZE, bench.EDGE_LEN, bench.OFFSETS_LIST
    else:  # a golden case name, e.g. c1_selfedge_2scenes
        import numpy as np
        g = np.load(os.path.join(ROOT, 'tests', 'golden', arg + '.npz'))
        b = type('B', (), {k: torch.from_numpy(g[k]) for k in ('feats_quarter', 'rotmats', 'tvecs', 'K',
                                                                'ref_src_edges', 'images_batch')})
        cfg = dict(depth_start=float(g['depth_start']), depth_interval=float(g['depth_interval']),
                   n_intervals=int(g['D']), size=tuple(int(v) for v in g['plane']))
        img = tuple(int(v) for v in g['img_size'])
        edge_len, offsets_list = float(g['edge_len']), g['offsets'].tolist()
        params = synth.make_params(int(g['seed']))
    bench.EDGE_LEN, bench.OFFSETS_LIST = edge_len, offsets_list
    net = lm.PL3DVNet(cfg, cfg, bench.EDGE_LEN, feat_dim=32, img_size=img)
    net.load_state_dict(params, strict=False)
    net = net.cuda().eval()
    dev = 'cuda'
    fq, R, t, K, e = b.feats_quarter, b.rotmats, b.tvecs, b.K, b.ref_src_edges
    gfq, gR, gt, gK = fq.to(dev), R.to(dev), t.to(dev), K.to(dev)
    ref_idx = torch.unique(e[0])
    db = b.images_batch[ref_idx]
    gdb = db.to(dev)
    with torch.no_grad():
        d_o, xv_o, xr_o = pipeline.initial_depth(fq, R, t, K, e, cfg, img, params, return_all=True)
        batch = type('B', (), dict(rotmats=gR, tvecs=gt, K=gK, ref_src_edges=e))
        xv = net.mvsnet.cost_volume(gfq, batch, cfg['depth_start'], cfg['depth_interval'], cfg['n_intervals'],
                                    cfg['size'])
        stats('x_var', xv.cpu(), xv_o)
        d_end = cfg['depth_start'] + cfg['depth_interval'] * (cfg['n_intervals'] - 1)
        d_g, xr = net.mvsnet.cnn_3d.depth(xv, cfg['depth_start'], d_end, want_reg=True)
        stats('x_reg (own x_var)', xr.cpu(), xr_o)
        stats('depth0 (own x_var)', d_g.cpu(), d_o)
        d_g2, xr2 = net.mvsnet.cnn_3d.depth(xv_o.to(dev), cfg['depth_start'], d_end, want_reg=True)
        stats('x_reg (oracle x_var)', xr2.cpu(), xr_o)
        stats('depth0 (oracle x_var)', d_g2.cpu(), d_o)

        # free-running both; teacher-forced GPU stage on the oracle's depth
        depth_o = d_o.clone()
        depth_g = d_g.clone()
        for it, offsets in enumerate(bench.OFFSETS_LIST):
            xs_o, mid = pipeline.model_scene(depth_o, db, fq, R, t, K, e, bench.EDGE_LEN, img, params,
                                             return_all=True)
            xs_g = net.model_scene(depth_g, gdb, gfq, gR, gt, gK, e)
            pts_tf, feat_tf, _ = net.construct_feature_rich_pointcloud(depth_o.to(dev), gdb, gfq, gR, gt, gK, e)
            stats('  TF pts', pts_tf.cpu(), mid['pts'])
            stats('  TF pts_feat', feat_tf.cpu(), mid['pts_feat'])
            xs_tf = net.model_scene(depth_o.to(dev), gdb, gfq, gR, gt, gK, e)
            print('iter %d: voxels oracle %d / gpu free %d / gpu teacher-forced %d' % (
                it, mid['anchor_pts'].shape[0], xs_g[-1]['feats'].shape[0], xs_tf[-1]['feats'].shape[0]))
            if xs_tf[-1]['feats'].shape[0] == xs_o[-1]['feats'].shape[0]:
                for li in range(3):
                    stats('  TF xs[%d].feats' % li, xs_tf[li]['feats'].cpu(), xs_o[li]['feats'])
            for offset in offsets:
                off_o = pipeline.run_pointflow(xs_o, depth_o, db, fq, R, t, K, e, offset, 3, img, params)
                off_tf = net.run_pointflow(xs_tf, depth_o.to(dev), gdb, gfq, gR, gt, gK, e, offset, 3)
                stats('  TF offset (%.3f)' % offset, off_tf.cpu(), off_o)
                off_g = net.run_pointflow(xs_g, depth_g, gdb, gfq, gR, gt, gK, e, offset, 3)
                depth_o += off_o
                depth_g += off_g
                stats('  free depth after (%.3f)' % offset, depth_g.cpu(), depth_o)


if __name__ == '__main__':
    main()
