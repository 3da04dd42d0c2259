"""Profiling driver: runs the C2 hot path a few times and brackets ONE region of the last
pass with cudaProfilerStart/Stop, so that `ncu --profile-from-start off` captures only it.

    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/x \
        python tools/prof_step.py --region scene

regions: mvs (plane sweep + CostRegNet + soft-argmin), scene (model_scene), flow (one run_pointflow)
No L2-flush buffer is allocated: ncu saves/restores device memory around every replay pass.
"""
import argparse
import importlib
import os
import sys
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload constants + synthetic inputs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--region', default='scene', choices=['mvs', 'scene', 'flow', 'all'])
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
    prof = torch.cuda.profiler

    def region(name):
        class R(object):
            def __enter__(self_):
                if last and args.region in (name, 'all'):
                    torch.cuda.synchronize()
                    prof.start()

            def __exit__(self_, *a):
                if last and args.region in (name, 'all'):
                    torch.cuda.synchronize()
                    prof.stop()
        return R()

    with torch.no_grad():
        for it in range(3):
            last = it == 2
            plan = ops.edge_plan(b.ref_src_edges, dev)
            batch = Namespace(rotmats=d['rotmats'], tvecs=d['tvecs'], K=d['K'], ref_src_edges=plan)
            nhwc = net._nhwc.get(d['feats_quarter'])
            with region('mvs'):
                depth = net.mvsnet.depth_from_features(d['feats_quarter'], batch, cfg['depth_start'],
                                                       cfg['depth_interval'], cfg['n_intervals'], cfg['size'],
                                                       feats_nhwc=nhwc, plan=plan)
            depth_batch = d['images_batch'][plan.ref_idx]
            with region('scene'):
                xs = net.model_scene(depth, depth_batch, d['feats_quarter'], d['rotmats'], d['tvecs'], d['K'], plan)
            with region('flow'):
                off = net.run_pointflow(xs, depth, depth_batch, d['feats_quarter'], d['rotmats'], d['tvecs'], d['K'],
                                        plan, 0.05, 3)
            depth = depth + off
    torch.cuda.synchronize()
    print('ok', float(depth.mean()))


if __name__ == '__main__':
    main()
