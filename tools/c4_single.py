"""The 64-view scene of BASELINE configs[3] through PL3DVNet.hot_path on one GPU: two warm passes, then ONE pass between
cudaProfilerStart / Stop, for an ncu launch list of the scene-scale kernels:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/c4_single.py [--refs 64]
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--refs', type=int, default=64)
    args = ap.parse_args()
    sys.argv = sys.argv[:1]
    import bench
    synth = importlib.import_module('3dvnet_b200.synth')
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    dev = torch.device('cuda', 0)
    b = synth.make_batch(1, args.refs + bench.N_SRC, bench.IMG_SIZE, bench.PLANE, 32, 4, 3, False, 0)
    net = lm.PL3DVNet(bench.DEPTH_CFG, bench.DEPTH_CFG, bench.EDGE_LEN, feat_dim=32, img_size=bench.IMG_SIZE)
    net.load_state_dict(synth.make_params(0), strict=False)
    net = net.to(dev).eval()
    fq, R, t, K = b.feats_quarter.to(dev), b.rotmats.to(dev), b.tvecs.to(dev), b.K.to(dev)
    e, ib = b.ref_src_edges, b.images_batch.to(dev)
    with torch.no_grad():
        for _ in range(2):
            net.hot_path(fq, R, t, K, e, ib, bench.DEPTH_CFG, bench.OFFSETS_LIST)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        net.hot_path(fq, R, t, K, e, ib, bench.DEPTH_CFG, bench.OFFSETS_LIST)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()


if __name__ == '__main__':
    main()
