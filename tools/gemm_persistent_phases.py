"""Steady-state phases of the persistent tcgen05 gather-GEMM on a large pair-major sparse convolution (the level-1
layers of the 64-view scene: ~150 k voxels, 128 channels, ~12 neighbours per voxel): clock64 stamps of every CTA's 8th
and 9th tile (dv3d_gemm_set_timing_buffer).
    python tools/gemm_persistent_phases.py [n_voxels] [density]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NAMES = ['split + store of the tile\'s chunks (waits for free stages)', 'request of the next tile (row numbers, first chunks)',
         'wait for the MMAs', 'TMEM -> smem tile (general persistent kernel only)', 'epilogue + global stores']


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 151635
    density = float(sys.argv[2]) if len(sys.argv) > 2 else 0.43
    ops = importlib.import_module('3dvnet_b200.ops')
    g = torch.Generator().manual_seed(0)
    dev = 'cuda'
    C = 128
    # neighbours near the row itself, as in a voxel-sorted level
    nbr = (torch.arange(n).view(-1, 1) + torch.randint(-3000, 3000, (n, 27), generator=g)).clamp_(0, n - 1)
    nbr[torch.rand(n, 27, generator=g) >= density] = -1
    km = ops.KernelMap(nbr.int().to(dev)).build_plan()
    ops.finish_plans([km])
    feat, W = torch.randn(n, C, generator=g).to(dev), torch.randn(27, C, C, generator=g).to(dev)
    pw = ops.pack_weights(W.reshape(-1, C).contiguous())
    gw, gb = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
    fn = lambda: ops.sparse_conv(feat, km, W, gw, gb, None, True, packed=pw)
    buf = torch.zeros(148 * 12, dtype=torch.int64, device=dev)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ops.lib().call('dv3d_gemm_set_timing_buffer', buf.data_ptr())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    ops.lib().call('dv3d_gemm_set_timing_buffer', None)
    t = buf.cpu().numpy().reshape(148, 12).astype(np.float64)
    t = t[t[:, 11] > 0]
    print('n = %d, %d pair tiles (%.1f per CTA), pairs path %s; GEMM + reduce %.1f us' % (n, km.n_tiles, km.n_tiles / 148.0, km.use_pairs,
                                                                                           s.elapsed_time(e) * 1e3))
    if t.shape[0] == 0:
        print('(no stamps: the launch took the weight-stationary pair kernel; DV3D_PAIR_WS=0 selects the general persistent kernel)')
        return
    d = np.diff(t[:, :6], axis=1)
    for i, nm in enumerate(NAMES):
        print('| %-60s | %7.0f cycles | %5.2f us |' % (nm, d[:, i].mean(), d[:, i].mean() / 1965.0))
    per = (t[:, 6] - t[:, 0]).mean()
    print('| %-60s | %7.0f cycles | %5.2f us |' % ('tile period (start of tile 8 -> start of tile 9)', per, per / 1965.0))


if __name__ == '__main__':
    main()
