"""Where a tile of the fused decoder spends its time: clock64 stamps written by decoder_fused_kernel
(dv3d_decoder_set_timing_buffer) for every tile of one launch at C2 size.
    python tools/decoder_phases.py [n_pts]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n_pts = int(sys.argv[1]) if len(sys.argv) > 1 else 3136
    ops = importlib.import_module('3dvnet_b200.ops')
    ref = importlib.import_module('3dvnet_b200.mv3d.subnetworks.refinement')
    dec = ref.HypothesisDecoder(352, 128).cuda().eval()
    x = torch.randn(n_pts, 8, 352, device='cuda')
    tiles = (n_pts * 8 + 127) // 128
    buf = torch.zeros(tiles * 8, dtype=torch.int64, device='cuda')
    for _ in range(3):
        dec.run(x, 7, 0.05)
    ops.lib().call('dv3d_decoder_set_timing_buffer', buf.data_ptr())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    dec.run(x, 7, 0.05)
    e.record()
    torch.cuda.synchronize()
    ops.lib().call('dv3d_decoder_set_timing_buffer', None)
    t = buf.cpu().numpy().reshape(tiles, 8).astype(np.float64)
    names = ['entry -> deps ready (prologue, TMEM alloc)', 'layer 1 main loop', 'epilogue 1', 'layer 2 main loop', 'epilogue 2',
             'layer 3 main loop', 'epilogue 3 + head']
    d = np.diff(t, axis=1)
    print('%d tiles, launch %.1f us; per-tile phases in SM cycles (mean / min / max) and us at 1.965 GHz' % (tiles, s.elapsed_time(e) * 1e3))
    for i, n in enumerate(names):
        print('| %-45s | %8.0f | %8.0f | %8.0f | %6.2f us |' % (n, d[:, i].mean(), d[:, i].min(), d[:, i].max(), d[:, i].mean() / 1965.0))
    tot = t[:, 7] - t[:, 0]
    print('| %-45s | %8.0f | %8.0f | %8.0f | %6.2f us |' % ('tile total', tot.mean(), tot.min(), tot.max(), tot.mean() / 1965.0))


if __name__ == '__main__':
    main()
