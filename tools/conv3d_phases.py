"""Where an output plane of the tcgen05 first CostRegNet layer spends its time: clock64 stamps written by
conv3d_c32_c8_tc_kernel (dv3d_conv3d_set_timing_buffer) for the first work item of every CTA at C2 size.
    python tools/conv3d_phases.py [n_ref]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    ops = importlib.import_module('3dvnet_b200.ops')
    x = torch.randn(n, 32, 96, 56, 56, device='cuda')
    w = torch.randn(8, 32, 3, 3, 3, device='cuda') / 30
    sc, sh = torch.ones(8, device='cuda'), torch.zeros(8, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    buf = torch.zeros(148 * 64, dtype=torch.int64, device='cuda')
    for mode in ('ffma', 'tc'):
        ops.set_conv3d_mode(mode)
        ts = []
        for i in range(8):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            ops.conv3d_bn_relu(x, w, sc, sh, 1)
            e.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(s.elapsed_time(e) * 1e3)
        print('%s: %.1f us per launch (L2 flushed), %d ref view(s)' % (mode, float(np.median(ts)), n))
    ops.lib().call('dv3d_conv3d_set_timing_buffer', buf.data_ptr())
    ops.conv3d_bn_relu(x, w, sc, sh, 1)
    torch.cuda.synchronize()
    ops.lib().call('dv3d_conv3d_set_timing_buffer', None)
    t = buf.cpu().numpy().reshape(148, 64).astype(np.float64)
    t = t[t[:, 0] > 0]
    names = ['mma: inputs + TMEM ready', 'mma: issued', 'epi: accumulators complete', 'epi: TMEM drained into T',
             'epi: gathered + stored']
    print('%d CTAs; stamps relative to the CTA start, SM cycles, mean over CTAs' % t.shape[0])
    print('| plane | ' + ' | '.join(names) + ' |')
    nz = max(zi + 1 for zi in range(8) if (t[:, 5 + zi * 5] > 0).all())   # planes per work item (at most 8 are stamped)
    for zi in range(nz):
        print('| %d | ' % zi + ' | '.join('%7.0f' % (t[:, 1 + zi * 5 + k] - t[:, 0]).mean() for k in range(5)) + ' |')
    per = (t[:, 5 + (nz - 1) * 5] - t[:, 5 + 2 * 5]).mean() / (nz - 3)
    print('steady state: %.0f cycles per output plane (%.2f us)' % (per, per / 1965.0))


if __name__ == '__main__':
    main()
