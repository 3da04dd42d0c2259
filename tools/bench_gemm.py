"""In-situ (warm, back-to-back) timing of the tcgen05 gather-GEMM on the three shapes that carry
the C2 step: the first decoder Conv1d, a sparse convolution at C2 voxel counts (K-split), a
PointNet linear.  Compares 3xTF32 with plain TF32 (1 MMA instead of 3, same operand traffic from
global memory): the ratio tells MMA / shared-memory bound from latency / launch bound.
    python tools/bench_gemm.py [--iters 50]
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
    ap.add_argument('--iters', type=int, default=50)
    args = ap.parse_args()
    ops = importlib.import_module('3dvnet_b200.ops')
    dev = 'cuda'
    g = torch.Generator().manual_seed(0)

    def rnd(*shape):
        return torch.randn(*shape, generator=g).to(dev)

    def timeit(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return 1e3 * s.elapsed_time(e) / args.iters

    rows = []
    for mode in ('tf32x3', 'tf32'):
        ops.set_gemm_mode(mode)
        # decoder Conv1d 0 / 1
        for Cin in (352, 128):
            n_pts = 3136
            x = rnd(n_pts, 8, Cin)
            w = rnd(3, Cin, 128)
            sc, sh = rnd(128), rnd(128)
            packed = ops.pack_weights(w.reshape(-1, 128).contiguous())
            out = torch.empty((n_pts, 8, 128), device=dev)
            us = timeit(lambda: ops.conv1d_bn_relu(x, w, sc, sh, out=out, packed=packed))
            rows.append((mode, 'conv1d M=25088 K=%d N=128' % (3 * Cin), us, 2.0 * 25088 * 3 * Cin * 128 / us / 1e6))
        # sparse conv at C2 sizes
        for n, Cin, Cout, density in ((3000, 128, 128, 0.15), (3000, 128, 128, 1.0), (3100, 64, 64, 0.15),
                                      (50000, 128, 128, 0.3)):
            nbr = torch.randint(0, n, (n, 27), generator=g)
            nbr[torch.rand(n, 27, generator=g) >= density] = -1
            nbr = nbr.int().to(dev)
            feat = rnd(n, Cin)
            W = rnd(27, Cin, Cout)
            gw, gb = rnd(Cout), rnd(Cout)
            packed = ops.pack_weights(W.reshape(-1, Cout).contiguous())
            ws = ops.sparse_conv_workspace(Cout, dev)
            us = timeit(lambda: ops.sparse_conv(feat, nbr, W, gw, gb, feat if Cin == Cout else None, True, packed=packed,
                                                workspace=ws))
            rows.append((mode, 'sparse conv n=%d %d->%d density %.2f (split)' % (n, Cin, Cout, density), us,
                         2.0 * n * 27 * Cin * Cout / us / 1e6))
            us = timeit(lambda: ops.sparse_conv(feat, nbr, W, gw, gb, feat if Cin == Cout else None, True, packed=packed))
            rows.append((mode, 'sparse conv n=%d %d->%d density %.2f (no split)' % (n, Cin, Cout, density), us,
                         2.0 * n * 27 * Cin * Cout / us / 1e6))
        # PointNet linear
        x = rnd(3136, 128)
        wk = rnd(128, 128)
        b = rnd(128)
        packed = ops.pack_weights(wk)
        us = timeit(lambda: ops.linear(x, wk, b, True, packed=packed))
        rows.append((mode, 'linear M=3136 K=128 N=128', us, 2.0 * 3136 * 128 * 128 / us / 1e6))
    # (the pipeline-stage ablation switches of round 1 were removed from the product kernel; the per-phase cost of a
    # tile is now read from in-kernel clock stamps: tools/gemm_phases.py, tools/decoder_phases.py)
    print('| mode | shape | us / launch | dense-equivalent TFLOP/s |')
    print('|---|---|---:|---:|')
    for r in rows:
        print('| %s | %s | %.1f | %.1f |' % r)


if __name__ == '__main__':
    main()
