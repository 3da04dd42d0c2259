"""In-situ (warm, back-to-back) device time of every stage of the C2 hot path, with CUDA events:
    python tools/stage_times.py [--iters 20] [--refs 1]
Complements the ncu launch lists (cold-cache, serialised): shows how much of a step is launch
gaps / host synchronisation rather than kernel time."""
import argparse
import importlib
import os
import sys
import time
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--refs', type=int, default=1)
    args = ap.parse_args()
    ops = importlib.import_module('3dvnet_b200.ops')
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    dev = torch.device('cuda', 0)
    b, params = bench.synth_inputs(0, args.refs)
    net = lm.PL3DVNet(bench.DEPTH_CFG, bench.DEPTH_CFG, bench.EDGE_LEN, feat_dim=32, img_size=bench.IMG_SIZE)
    net.load_state_dict(params, strict=False)
    net = net.to(dev).eval()
    d = {k: getattr(b, k).to(dev) for k in ('feats_quarter', 'rotmats', 'tvecs', 'K', 'images_batch')}
    cfg = bench.DEPTH_CFG
    plan = ops.edge_plan(b.ref_src_edges, dev)
    batch = Namespace(rotmats=d['rotmats'], tvecs=d['tvecs'], K=d['K'], ref_src_edges=plan)
    nhwc = net._nhwc.get(d['feats_quarter'])
    depth_batch = d['images_batch'][plan.ref_idx]

    def mvs():
        return net.mvsnet.depth_from_features(d['feats_quarter'], batch, cfg['depth_start'], cfg['depth_interval'],
                                              cfg['n_intervals'], cfg['size'], feats_nhwc=nhwc, plan=plan)

    def timeit(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        t0 = time.perf_counter()
        s.record()
        for _ in range(n):
            fn()
        e.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n, 1e3 * t_host / n, (ops.launch_count() - l0) / n

    with torch.no_grad():
        depth = mvs()
        xs = net.model_scene(depth, depth_batch, d['feats_quarter'], d['rotmats'], d['tvecs'], d['K'], plan)
        g = torch.Generator().manual_seed(1)
        n_imgs = d['feats_quarter'].shape[0]
        feats_half = torch.randn(n_imgs, 32, bench.IMG_SIZE[0] // 2, bench.IMG_SIZE[1] // 2, generator=g).to(dev)
        images = torch.randn(n_imgs, 3, *bench.IMG_SIZE, generator=g).to(dev)
        stages = [
            ('mvs (plane sweep + CostRegNet + soft-argmin)', mvs),
            ('cost volume only', lambda: net.mvsnet.cost_volume(d['feats_quarter'], batch, cfg['depth_start'],
                                                                cfg['depth_interval'], cfg['n_intervals'], cfg['size'],
                                                                nhwc, plan)),
            ('model_scene', lambda: net.model_scene(depth, depth_batch, d['feats_quarter'], d['rotmats'], d['tvecs'],
                                                    d['K'], plan)),
            ('run_pointflow', lambda: net.run_pointflow(xs, depth, depth_batch, d['feats_quarter'], d['rotmats'],
                                                        d['tvecs'], d['K'], plan, 0.05, 3)),
            ('upsample: 3 PropagationNets, 56x56 -> 256x320 (after the path, SURVEY 8f.1)',
             lambda: net.upsample(depth, plan.ref_idx, d['feats_quarter'], feats_half, images)),
            ('hot_path (whole step)', lambda: net.hot_path(d['feats_quarter'], d['rotmats'], d['tvecs'], d['K'],
                                                           b.ref_src_edges, d['images_batch'], cfg, bench.OFFSETS_LIST)),
        ]
        print('| stage | device ms | host ms to enqueue | dv3d launches |')
        print('|---|---:|---:|---:|')
        for name, fn in stages:
            dev_ms, host_ms, launches = timeit(fn, args.iters)
            print('| %s | %.3f | %.3f | %.0f |' % (name, dev_ms, host_ms, launches))


if __name__ == '__main__':
    main()
