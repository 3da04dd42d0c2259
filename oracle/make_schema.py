"""Record the state_dict schema (key -> shape) of the UNMODIFIED reference PL3DVNet into
tests/golden/state_dict_schema.json, so that checkpoint interchange
is tested without /root/reference at test time.  Build container only:

    python oracle/make_schema.py

MinkowskiEngine layer shapes come from the CPU shim (oracle/shims): [27,Cin,Cout] kernels for
3x3x3 convolutions, [Cin,Cout] for 1x1 - the published ME 0.5 layout, unpinned.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, 'shims'), '/root/reference', ROOT]

import torchvision  # noqa: E402

_orig = torchvision.models.mnasnet1_0
torchvision.models.mnasnet1_0 = lambda pretrained=False, **kw: _orig(weights=None)

from mv3d.lightningmodel import PL3DVNet  # noqa: E402  (the reference)

if __name__ == '__main__':
    cfg = dict(depth_start=0.5, depth_interval=0.05, n_intervals=96, size=(56, 56))
    torch.manual_seed(0)
    net = PL3DVNet(cfg, cfg, 0.08, feat_dim=32, img_size=(256, 320))   # mv3d/config.py:22-42 defaults
    schema = {k: list(v.shape) for k, v in net.state_dict().items()}
    out = os.path.join(ROOT, 'tests', 'golden', 'state_dict_schema.json')
    json.dump(schema, open(out, 'w'), indent=0, sort_keys=True)
    print(len(schema), 'entries ->', out)
