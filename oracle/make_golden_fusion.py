"""Generate tests/golden/fusion.npz by running the UNMODIFIED reference depth fusion
(/root/reference/mv3d/eval/pointcloudfusion_custom.py:process_scene) on the CPU: the reference
hard-codes `.cuda()`, which is patched to the identity for this run (no GPU in the build
container); nothing else is changed.  Build container only:

    python oracle/make_golden_fusion.py
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.modules.setdefault('tqdm', types.SimpleNamespace(tqdm=lambda x, **k: x))
torch.Tensor.cuda = lambda self, *a, **k: self
spec = importlib.util.spec_from_file_location('ref_fusion', '/root/reference/mv3d/eval/pointcloudfusion_custom.py')
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
synth = importlib.import_module('3dvnet_b200.synth')

if __name__ == '__main__':
    n, size = 6, (48, 64)
    R, t, K = synth.make_cameras(n, size, seed=7)
    depth = synth.ray_box_depth(R, t, K, size, size)           # multi-view consistent box-room depths at image resolution
    rng = np.random.RandomState(0)
    depth = depth + rng.normal(0, 0.02, depth.shape).astype(np.float32)   # noise around the 0.1 threshold scale
    depth[0, :6, :6] = 0.0                                      # invalid depths
    poses = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    poses[:, :3, :3], poses[:, :3, 3] = R, t
    images = rng.rand(n, size[0], size[1], 3).astype(np.float32)
    td, tp, tk, ti = (torch.from_numpy(a) for a in (depth, poses, K, images))
    pts, rgb, valid = ref.process_scene(td.clone(), ti, tp, tk, 0.1, 3)
    # per-image intermediates of process_depth for image 0 (its own function, same patch)
    idx = torch.arange(n) != 0
    p0, rgb0, v0 = ref.process_depth(td[0].clone(), ti[0], td[idx].clone(), ti[idx], tp[0], tp[idx], tk[0], tk[idx], 0.1, 3)
    out = os.path.join(ROOT, 'tests', 'golden', 'fusion.npz')
    np.savez_compressed(out, depth=depth, poses=poses, K=K, images=images, ref_pts=pts, ref_rgb=rgb, ref_valid=valid,
                        ref_pts0=p0, ref_valid0=v0)
    print('wrote', out, os.path.getsize(out), 'fused points', pts.shape[0], 'valid fraction %.3f' % valid.mean())
