"""Generate tests/golden/propagation.npz by running the UNMODIFIED reference PropagationNet
(/root/reference/mv3d/subnetworks/upsampling.py, pure torch) and the upsampling cascade of
/root/reference/mv3d/eval-3dvnet.py:101-125 on seeded inputs.  Build container only:

    python oracle/make_golden_upsample.py
"""
import importlib.util
import os

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('ref_upsampling', '/root/reference/mv3d/subnetworks/upsampling.py')
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def seeded_net(in_dim, seed):
    torch.manual_seed(seed)
    m = ref.PropagationNet(in_dim, 32).eval()
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():   # non-trivial BatchNorm statistics, sharper logits
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.2)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.weight.data.copy_(torch.randn(mod.num_features, generator=g) * 0.3 + 1.0)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.3)
    m.conv4[1].weight.data.mul_(4.0)
    return m


if __name__ == '__main__':
    g = torch.Generator().manual_seed(0)
    save = {}
    nets = {'quarter': seeded_net(33, 10), 'half': seeded_net(33, 20), 'full': seeded_net(4, 30)}
    for name, m in nets.items():
        for k, v in m.state_dict().items():
            save['%s.%s' % (name, k)] = v.numpy()
    n = 2
    depth = torch.rand(n, 8, 8, generator=g) * 4 + 0.5                 # plane-resolution depth
    fq = torch.randn(n, 32, 12, 20, generator=g)                        # "quarter" (non-integer ratios 8->12, 8->20)
    fh = torch.randn(n, 32, 24, 40, generator=g)
    img = torch.randn(n, 3, 48, 80, generator=g)
    with torch.no_grad():
        d = F.interpolate(depth.unsqueeze(1), fq.shape[-2:], mode='nearest')        # eval-3dvnet.py:103
        save['up_quarter'] = d.squeeze(1).numpy()
        dq = nets['quarter'](fq, d)
        d = F.interpolate(dq.unsqueeze(1), fh.shape[-2:], mode='nearest')
        dh = nets['half'](fh, d)
        d = F.interpolate(dh.unsqueeze(1), img.shape[-2:], mode='nearest')
        df = nets['full'](img, d)
    save.update(depth=depth.numpy(), feats_quarter=fq.numpy(), feats_half=fh.numpy(), images=img.numpy(),
                ref_quarter=dq.numpy(), ref_half=dh.numpy(), ref_full=df.numpy())
    out = os.path.join(ROOT, 'tests', 'golden', 'propagation.npz')
    np.savez_compressed(out, **save)
    print('wrote', out, os.path.getsize(out), 'bytes', 'depth range', float(df.min()), float(df.max()))
