"""Pair-major vs output-stationary sparse convolution at the sizes of the 64-view scene (BASELINE configs[3]) and of one
rank's share at 8 GPUs: time of ops.sparse_conv (GEMM + reduce, L2 flushed) for both variants, and what the heuristic
picks.  DV3D_PAIR_WS=0 times the general persistent kernel instead of the weight-stationary pair kernel.
    python tools/sparse_conv_paths.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ops = importlib.import_module('3dvnet_b200.ops')
    g = torch.Generator().manual_seed(0)
    dev = 'cuda'
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, it=5):
        ts = []
        for i in range(it + 2):
            flush.zero_()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            z.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(z) * 1e3)
        return float(np.median(ts))

    print('| voxels | C | density | pair tiles | dense tile-chunks | heuristic | pair-major us | output-stationary us | max diff |')
    print('|---:|---:|---:|---:|---:|---|---:|---:|---:|')
    for n, C, density in ((151635, 128, 0.43), (193602, 64, 0.11), (54803, 128, 0.6), (19000, 128, 0.43), (24000, 64, 0.11),
                          (6850, 128, 0.6), (13700, 128, 0.6), (3045, 128, 0.10), (2408, 128, 0.38)):
        nbr = (torch.arange(n).view(-1, 1) + torch.randint(-3000, 3000, (n, 27), generator=g)).clamp_(0, n - 1)
        nbr[torch.rand(n, 27, generator=g) >= density] = -1
        km = ops.KernelMap(nbr.int().to(dev)).build_plan()
        ops.finish_plans([km])
        feat, W = torch.randn(n, C, generator=g).to(dev), torch.randn(27, C, C, generator=g).to(dev)
        pw = ops.pack_weights(W.reshape(-1, C).contiguous())
        gw, gb = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
        ws = ops.sparse_conv_workspace(128, dev)
        fn = lambda: ops.sparse_conv(feat, km, W, gw, gb, None, True, packed=pw, workspace=ws)
        auto = km.use_pairs
        km.use_pairs = True
        tp, yp = timed(fn), fn().clone()
        km.use_pairs = False
        td, yd = timed(fn), fn().clone()
        print('| %d | %d | %.2f | %d | %d | %s | %.0f | %.0f | %.1e |' % (n, C, density, km.n_tiles, 27 * ((n + 127) // 128),
                                                                       'pairs' if auto else 'output-stationary', tp, td,
                                                                       float((yp - yd).abs().max())))


if __name__ == '__main__':
    main()
