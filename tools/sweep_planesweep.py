"""HBM-roofline sweep of the fused plane-sweep warp + variance kernel (BASELINE.json configs[4] and
the other configs of SURVEY.md §8d): algorithmic bytes (every source feature map read once + the
[C,D,h,w] slab written once) / measured kernel time, against the measured HBM peak.
    python tools/sweep_planesweep.py [--iters 20]
Timing: CUDA events around the single kernel launch, L2 flushed (256 MiB memset) before every launch.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # name, image size, D, plane, n_src, n_ref
    ('C1 64x80 D=16 plane 16x16, 2 src', (64, 80), 16, (16, 16), 2, 1),
    ('C2 256x320 D=96 plane 56x56, 7 src', (256, 320), 96, (56, 56), 7, 1),
    ('C2 256x320 D=96 plane 64x80, 7 src', (256, 320), 96, (64, 80), 7, 1),
    ('C2 x 8 ref views (plane 56x56)', (256, 320), 96, (56, 56), 7, 8),
    ('C5 512x640 D=192 plane 112x112, 10 src', (512, 640), 192, (112, 112), 10, 1),
    ('C5 512x640 D=192 plane 128x160, 10 src', (512, 640), 192, (128, 160), 10, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    ops = importlib.import_module('3dvnet_b200.ops')
    synth = importlib.import_module('3dvnet_b200.synth')
    dev = torch.device('cuda', 0)
    peaks = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peak = float(json.load(open(peaks))['hbm_gbs']) if os.path.exists(peaks) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print('| case | mode | voxels | edges | algorithmic MB | kernel us | GB/s | frac of %.0f GB/s |' % peak)
    print('|---|---|---:|---:|---:|---:|---:|---:|')
    for name, img, D, plane, n_src, n_ref in [c for c in CASES for _ in (0, 1)][::2]:
        b = synth.make_batch(1, n_ref + n_src, img, plane, 32, n_src - n_src // 2, n_src // 2, False, 0)
        plan = ops.edge_plan(b.ref_src_edges, dev)
        nhwc = ops.nchw_to_nhwc(b.feats_quarter.to(dev))
        cams = ops.camera_tables(b.rotmats.to(dev), b.tvecs.to(dev), b.K.to(dev))
        out = torch.empty((plan.n_ref, 32, D) + plane, dtype=torch.float32, device=dev)
        Hf, Wf = b.feats_quarter.shape[-2:]
        algo = plan.n_edges * 32 * Hf * Wf * 4 + out.numel() * 4
        for mode in ('exact', 'fast'):
            ops.set_warp_mode(mode)
            ts = []
            for i in range(args.iters + 3):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                ops.planesweep_var(nhwc, cams, plan, 0.5, 0.05, D, plane, img, out=out)
                e.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(s.elapsed_time(e) * 1e3)
            us = float(np.median(ts))
            gbs = algo / us / 1e3
            print('| %s | %s | %d | %d | %.1f | %.1f | %.0f | %.3f |' % (name, mode, plan.n_ref * D * plane[0] * plane[1],
                                                                       plan.n_edges, algo / 1e6, us, gbs, gbs / peak))


if __name__ == '__main__':
    main()
