"""Coarse-to-fine depth upsampling — drop-in for /root/reference/mv3d/subnetworks/upsampling.py:14-36
(same constructor, forward signature and state_dict keys). The step immediately after the hot path
(SURVEY.md §8f.1): the four Conv2d+BN+ReLU layers run as 9-slice gather-GEMMs on the tcgen05
kernel, the nearest upsampling in front (eval-3dvnet.py:101-125) and the softmax-weighted 3x3
gather behind them as two small kernels (csrc/upsample.cu)."""
import torch.nn as nn

from ... import ops
from .._pack import PackCache, fold_bn, require_eval


def _conv_bn_relu(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class PropagationNet(nn.Module):
    def __init__(self, in_dim=4, h_dim=32):
        super().__init__()
        if h_dim > 64:
            raise NotImplementedError('the propagation kernels keep activations in 64-column rows (h_dim <= 64)')
        self.conv1 = _conv_bn_relu(in_dim, h_dim)
        self.conv2 = _conv_bn_relu(h_dim, h_dim)
        self.conv3 = _conv_bn_relu(h_dim, h_dim)
        self.conv4 = _conv_bn_relu(h_dim, 9)
        self.unfold = nn.Unfold(kernel_size=3, stride=1, padding=0)  # kept for module-tree parity; unused
        self.in_dim = in_dim
        self._pack = PackCache()

    def _layers(self):
        def build():
            out = []
            for seq in (self.conv1, self.conv2, self.conv3, self.conv4):
                w = seq[0].weight.detach().float()                      # [Cout, Cin, 3, 3]
                cout, cin = w.shape[:2]
                cin_pad = (cin + 31) // 32 * 32
                w_kn = w.new_zeros(9, cin_pad, 64)                      # row t*Cin_pad + ci, t = ky*3 + kx
                w_kn[:, :cin, :cout] = w.permute(2, 3, 1, 0).reshape(9, cin, cout)
                w_kn = w_kn.reshape(9 * cin_pad, 64).contiguous()
                scale, shift = w.new_zeros(64), w.new_zeros(64)
                s, b = fold_bn(seq[1])
                scale[:cout], shift[:cout] = s, b
                out.append((w_kn, ops.pack_weights(w_kn), scale, shift))
            return out
        tensors = [p for p in self.parameters()] + [b for b in self.buffers()]
        return self._pack.get(tensors, build, ops.gemm_mode())

    def forward(self, features, depth):
        """features [b,in_dim-1,h,w], depth [b,1,h,w] (already at the feature resolution) -> [b,h,w]"""
        require_eval(self)
        return ops.propagation_net(features.detach().float().contiguous(), depth.detach().float()[:, 0].contiguous(),
                                   self._layers())

    def forward_from(self, features, depth_lo):
        """fused F.interpolate(depth_lo, features.shape[-2:], mode='nearest') + forward: depth_lo [b,h,w]"""
        require_eval(self)
        return ops.propagation_net(features.detach().float().contiguous(), depth_lo.detach().float().contiguous(),
                                   self._layers())
