"""Where a CTA of the tcgen05 gather-GEMM spends its time on the small problems of the scene model (clock64 stamps,
dv3d_gemm_set_timing_buffer): a PointNet linear (3136 rows, K = 128 / 256) and pair-major sparse convolutions.
    python tools/gemm_phases.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NAMES = ['prologue (mbarrier init, TMEM alloc)', 'wait for the previous kernel (PDL)', 'row table -> smem',
         'gather + split + store (this thread)', 'wait for the MMAs', 'TMEM -> smem tile', 'epilogue + stores']


def report(ops, label, fn, n_ctas):
    buf = torch.zeros(n_ctas * 12, dtype=torch.int64, device='cuda')
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
    full = buf.cpu().numpy().reshape(n_ctas, 12).astype(np.float64)
    full = full[full[:, 0] > 0]
    t = full[:, :8]
    d = np.diff(t, axis=1)
    print('\n%s: %d CTAs, launch %.1f us (events, incl. launch latency)' % (label, len(t), s.elapsed_time(e) * 1e3))
    for i, n in enumerate(NAMES):
        print('| %-40s | %7.0f cycles | %5.2f us |' % (n, d[:, i].mean(), d[:, i].mean() / 1965.0))
    print('|   epilogue detail: params loaded +%.0f, two rows done +%.0f, all rows +%.0f cycles after the tile sync' % (
        (full[:, 8] - full[:, 6]).mean(), (full[:, 9] - full[:, 6]).mean(), (full[:, 7] - full[:, 6]).mean()))
    tot = t[:, 7] - t[:, 0]
    print('| %-40s | %7.0f cycles | %5.2f us |' % ('CTA total', tot.mean(), tot.mean() / 1965.0))


def main():
    ops = importlib.import_module('3dvnet_b200.ops')
    g = torch.Generator().manual_seed(0)
    dev = 'cuda'
    rnd = lambda *s: torch.randn(*s, generator=g).to(dev)
    N = 3136
    x, w, b = rnd(N, 128), rnd(128, 128), rnd(128)
    packed = ops.pack_weights(w)
    report(ops, 'PointNet linear M=3136 K=128 N=128', lambda: ops.linear(x, w, b, True, packed=packed), 25)
    pool, seg = rnd(N, 128), torch.arange(N, dtype=torch.int32, device=dev)
    w2 = rnd(256, 128)
    p2 = ops.pack_weights(w2)
    report(ops, 'PointNet linear on [x | pool[seg]] K=256', lambda: ops.linear(x, w2, b, True, pool=pool, seg=seg, packed=p2), 25)
    for n, C, density in ((3100, 64, 0.045), (3000, 128, 0.10), (2400, 128, 0.38)):
        nbr = torch.randint(0, n, (n, 27), generator=g)
        nbr[torch.rand(n, 27, generator=g) >= density] = -1
        km = ops.KernelMap(nbr.int().to(dev)).build_plan()
        ops.finish_plans([km])
        feat, W = rnd(n, C), rnd(27, C, C)
        pw = ops.pack_weights(W.reshape(-1, C).contiguous())
        gw, gb = rnd(C), rnd(C)
        report(ops, 'pair-major sparse conv n=%d C=%d density %.3f (%d tiles, pairs path %s)' % (n, C, density, km.n_tiles, km.use_pairs),
               lambda: ops.sparse_conv(feat, km, W, gw, gb, None, True, packed=pw), max(km.n_tiles, 1))


if __name__ == '__main__':
    main()
