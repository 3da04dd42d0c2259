"""Device time of the three parts of PL3DVNet.full_pass at BASELINE C2 (8 images of 256x320, 1 reference view): 2D
backbone + FPN (CUDA graph of torchvision / cuDNN kernels), the hot path (native engine), nearest upsampling + three
PropagationNets.  CUDA events around each part, L2 not flushed (the parts run back to back as in the pipeline).
    python tools/full_pass_stages.py [--iters 20]"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    sys.argv = sys.argv[:1]
    import bench
    ops = importlib.import_module('3dvnet_b200.ops')
    synth = importlib.import_module('3dvnet_b200.synth')
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    dev = torch.device('cuda', 0)
    b = synth.make_batch(1, 1 + bench.N_SRC, bench.IMG_SIZE, bench.PLANE, 32, 4, 3, False, 0)
    net = lm.PL3DVNet(bench.DEPTH_CFG, bench.DEPTH_CFG, bench.EDGE_LEN, feat_dim=32, img_size=bench.IMG_SIZE)
    net.load_state_dict(synth.make_params(0), strict=False)
    net = net.to(dev).eval()
    images = torch.rand(1 + bench.N_SRC, 3, *bench.IMG_SIZE, device=dev)
    R, t, K, ib = b.rotmats.to(dev), b.tvecs.to(dev), b.K.to(dev), b.images_batch.to(dev)
    plan = ops.edge_plan(b.ref_src_edges, dev)
    names = ('backbone + FPN', 'hot path (engine)', 'upsampling + 3 PropagationNets')
    ts = {n: [] for n in names}
    with torch.no_grad():
        for i in range(args.iters + 3):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            fh, fq = net._backbone(images)
            ev[1].record()
            depth = net.hot_path(fq, R, t, K, plan, ib, bench.DEPTH_CFG, bench.OFFSETS_LIST)
            ev[2].record()
            final = net.upsample(depth, plan.ref_idx, fq, fh, images)
            ev[3].record()
            torch.cuda.synchronize()
            if i >= 3:
                for k, n in enumerate(names):
                    ts[n].append(ev[k].elapsed_time(ev[k + 1]))
    for n in names:
        print('| %-32s | %.3f ms |' % (n, float(np.median(ts[n]))))
    print('| %-32s | %.3f ms |' % ('sum', sum(float(np.median(ts[n])) for n in names)))


if __name__ == '__main__':
    main()
