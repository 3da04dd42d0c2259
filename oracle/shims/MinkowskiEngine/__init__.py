"""CPU stand-in for the MinkowskiEngine 0.5.0 API surface the reference uses
(scenemodeling.py:10-12,27-28,36-39,100-104,160-162,181-188,194,206,213-216;
refinement.py:26,39), built on oracle/minkowski_cpu.py. TEST INFRASTRUCTURE ONLY — it lets
oracle/make_golden.py run the reference's unmodified SparseUNet / HypothesisDecoder."""
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import minkowski_cpu as mk  # noqa: E402


class CoordinateManager(object):
    def __init__(self):
        self.maps = {}


class SparseTensor(object):
    def __init__(self, features, coordinates=None, coordinate_map_key=None, coordinate_manager=None):
        if coordinates is not None:
            c = coordinates.detach().cpu().numpy().astype(np.int64)
            order = np.argsort(mk.encode(c), kind='stable')
            cmap = mk.CoordMap(c[order], 1, presorted=True)
            assert (np.diff(cmap.keys) > 0).all(), 'shim supports unique coordinates only'
            self.coordinate_manager = CoordinateManager()
            self.coordinate_manager.maps[1] = cmap
            self.coordinate_map_key = 1
            self._F = features[torch.from_numpy(order)]
        else:
            self.coordinate_manager = coordinate_manager
            self.coordinate_map_key = coordinate_map_key
            self._F = features

    @property
    def cmap(self):
        return self.coordinate_manager.maps[self.coordinate_map_key]

    @property
    def F(self):
        return self._F

    @property
    def C(self):
        return torch.from_numpy(self.cmap.coords).int()

    @property
    def tensor_stride(self):
        return [self.cmap.stride] * 3

    def __iadd__(self, other):
        assert other.coordinate_map_key == self.coordinate_map_key
        self._F = self._F + other.F
        return self


class MinkowskiConvolution(nn.Module):
    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, dimension=None):
        super().__init__()
        assert dimension == 3 and dilation == 1 and kernel_size in (1, 3) and not bias
        self.kernel_size, self.stride = kernel_size, stride
        kv = kernel_size ** 3
        shape = (kv, in_channels, out_channels) if kv > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(*shape))
        n = (out_channels if self.transposed else in_channels) * kv
        self.kernel.data.uniform_(-1.0 / math.sqrt(n), 1.0 / math.sqrt(n))

    def forward(self, x):
        if self.kernel_size == 1:
            return SparseTensor(x.F @ self.kernel, coordinate_map_key=x.coordinate_map_key,
                                coordinate_manager=x.coordinate_manager)
        cm = x.coordinate_manager
        if not self.transposed:
            feat, out_map = mk.conv3(x.F, x.cmap, self.kernel, self.stride)
            key = out_map.stride
            if key not in cm.maps:
                cm.maps[key] = out_map
            return SparseTensor(feat, coordinate_map_key=key, coordinate_manager=cm)
        key = x.cmap.stride // self.stride
        assert key in cm.maps, 'transposed convolution expects the finer map to exist'
        feat = mk.conv3_transpose(x.F, x.cmap, cm.maps[key], self.kernel)
        return SparseTensor(feat, coordinate_map_key=key, coordinate_manager=cm)


class MinkowskiConvolutionTranspose(MinkowskiConvolution):
    transposed = True


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum)

    def forward(self, x):
        return SparseTensor(self.bn(x.F), coordinate_map_key=x.coordinate_map_key,
                            coordinate_manager=x.coordinate_manager)


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return SparseTensor(torch.relu(x.F), coordinate_map_key=x.coordinate_map_key,
                            coordinate_manager=x.coordinate_manager)


class MinkowskiInterpolation(nn.Module):
    def forward(self, x, tfield):
        return mk.interpolate(x.cmap, x.F, tfield.detach())


def cat(*tensors):
    key = tensors[0].coordinate_map_key
    assert all(t.coordinate_map_key == key for t in tensors)
    return SparseTensor(torch.cat([t.F for t in tensors], dim=1), coordinate_map_key=key,
                        coordinate_manager=tensors[0].coordinate_manager)
