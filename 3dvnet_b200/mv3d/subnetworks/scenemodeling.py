"""Scene modelling — drop-in for /root/reference/mv3d/subnetworks/scenemodeling.py: per-voxel
PointNet and the sparse 3D-UNet, with MinkowskiEngine and torch_scatter replaced by the
voxel hash / kernel-map / gather-GEMM kernels of lib3dvnet_b200 (csrc/sparse.cu, gemm.cu,
pointnet.cu). Module tree and parameter names follow the reference so that its checkpoints
load by name (MinkowskiEngine kernels are [27,Cin,Cout] / [Cin,Cout])."""
import math

import torch
import torch.nn as nn

from ... import ops
from .._pack import PackCache, require_eval


class SparseConvolution(nn.Module):
    """Parameter holder with MinkowskiConvolution's ``kernel`` tensor (bias=False)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, transposed=False):
        super().__init__()
        self.kernel_size, self.stride, self.transposed = kernel_size, stride, transposed
        kv = kernel_size ** 3
        shape = (kv, in_channels, out_channels) if kv > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(*shape))
        n = (out_channels if transposed else in_channels) * kv
        nn.init.uniform_(self.kernel, -1.0 / math.sqrt(n), 1.0 / math.sqrt(n))
        self._pack = PackCache()

    def weights(self):
        """(kernel fp32, packed tensor-core image or None) for ops.sparse_conv / concat_linear"""
        k = self.kernel.detach()
        packed = self._pack.get([self.kernel], lambda: ops.pack_weights(k.float().reshape(-1, k.shape[-1]).contiguous()),
                                ops.gemm_mode())
        return k, packed


class MinkowskiGroupNorm(nn.Module):
    """scenemodeling.py:78-113: torch GroupNorm over the [N,C] feature rows, i.e. per voxel."""

    def __init__(self, num_groups, num_channels, eps=1e-5, affine=True):
        super().__init__()
        if num_channels // num_groups != 16 or eps != 1e-5:
            raise NotImplementedError('fused GroupNorm supports 16 channels per group, eps 1e-5')
        self.gn = nn.GroupNorm(num_groups, num_channels, eps=eps, affine=affine)


class ReLU(nn.Module):
    """placeholder keeping the reference's Sequential indices (ME.MinkowskiReLU has no state)"""


class SparseResidual3d(nn.Module):
    """relu(x + GN2(conv2(relu(GN1(conv1(x)))))) (scenemodeling.py:16-44)."""

    def __init__(self, feat_dim, norm='gn', num_groups=None):
        super().__init__()
        assert norm == 'gn'
        self.n1 = MinkowskiGroupNorm(num_groups, feat_dim)
        self.n2 = MinkowskiGroupNorm(num_groups, feat_dim)
        nn.init.constant_(self.n2.gn.weight, 0)
        self.conv1 = SparseConvolution(feat_dim, feat_dim)
        self.conv2 = SparseConvolution(feat_dim, feat_dim)

    def forward(self, x, nbr, ws=None):
        w1, p1 = self.conv1.weights()
        w2, p2 = self.conv2.weights()
        h = ops.sparse_conv(x, nbr, w1, self.n1.gn.weight.detach(), self.n1.gn.bias.detach(), None, True, packed=p1,
                            workspace=ws)
        return ops.sparse_conv(h, nbr, w2, self.n2.gn.weight.detach(), self.n2.gn.bias.detach(), x, True, packed=p2,
                               workspace=ws)


class PointNet(nn.Module):
    """scenemodeling.py:116-144."""

    def __init__(self, hidden_dim, out_dim, in_dim=3):
        super().__init__()
        self.fc_pos = nn.Linear(in_dim, hidden_dim)
        self.fc1 = nn.Linear(hidden_dim, hidden_dim)
        self.fc2 = nn.Linear(2 * hidden_dim, hidden_dim)
        self.fc3 = nn.Linear(2 * hidden_dim, hidden_dim)
        self.fc4 = nn.Linear(2 * hidden_dim, hidden_dim)
        self.fc_out = nn.Linear(hidden_dim, out_dim)
        self.in_dim = in_dim
        self.in_pad = (in_dim + 31) // 32 * 32  # K extent of the tensor-core GEMM: multiples of 32
        self._pack = PackCache()

    def _weights(self):
        def build():
            out = {}
            for name in ('fc_pos', 'fc1', 'fc2', 'fc3', 'fc4', 'fc_out'):
                fc = getattr(self, name)
                w = fc.weight.detach().float().t().contiguous()  # [K, Cout]
                if name == 'fc_pos' and w.shape[0] != self.in_pad:
                    w = torch.cat((w, w.new_zeros(self.in_pad - w.shape[0], w.shape[1])), 0).contiguous()
                out[name] = (w, fc.bias.detach().float().contiguous(), ops.pack_weights(w))
            return out
        return self._pack.get([p for p in self.parameters()], build, ops.gemm_mode())

    def forward_padded(self, x_pad, seg, n_idx):
        """x_pad [N, in_pad] (zero-padded input rows), seg [N] int32"""
        w = self._weights()
        def fc(name, x, relu, **kw):
            wt, b, packed = w[name]
            return ops.linear(x, wt, b, relu_input=relu, packed=packed, **kw)

        x = fc('fc_pos', x_pad, False)
        x = fc('fc1', x, True)
        for name in ('fc2', 'fc3', 'fc4'):
            pool = ops.segment_max(x, seg, n_idx)
            x = fc(name, x, True, pool=pool, seg=seg)
        pool = ops.segment_max(x, seg, n_idx)
        return fc('fc_out', pool, True)

    def forward(self, pts, idx, n_idx):
        """pts [N,in_dim], idx [N] int64 -> [n_idx,out_dim]"""
        x = pts.float()
        if x.shape[1] != self.in_pad:
            x = torch.cat((x, x.new_zeros(x.shape[0], self.in_pad - x.shape[1])), 1)
        return self.forward_padded(x.contiguous(), idx.int().contiguous(), n_idx)


class SparseScene(object):
    """Coordinate levels + kernel maps of one voxelised scene (what ME keeps in its
    coordinate manager). Built once per model_scene call, shared by all layers."""

    def __init__(self, idx, batch, n_levels, dims=None):
        dev = idx.device
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        idx = idx.int().contiguous()
        batch = batch.long().contiguous()
        if dims is None:  # bound of the index range; one sync
            mx = torch.cat((idx.max(dim=0)[0].long(), batch.max().view(1))).cpu()
            dims = (int(mx[0]) + 1, int(mx[1]) + 1, int(mx[2]) + 1)
            n_batch = int(mx[3]) + 1
        else:
            dims, n_batch = dims[:3], dims[3]
        self.n_batch = n_batch
        self.levels = [ops.SparseLevel(ops.make_coords(idx, batch), 1, self.err)]
        for _ in range(1, n_levels):
            self.levels.append(ops.coarsen(self.levels[-1], dims, n_batch, self.err))
        # K-split workspace of the tensor-core sparse convolution, shared by all layers
        self.ws = ops.sparse_conv_workspace(128, dev)
        # every kernel map of the U-Net and its pair-major plan, then one sync for the counts
        n = len(self.levels)
        maps = [self.same(l) for l in range(n)] + [self.down(l) for l in range(n - 1)] + [self.up(l) for l in range(n - 1)]
        ops.build_plans(maps)
        ops.finish_plans(maps)

    def same(self, l):       # k3 s1 on level l
        lv = self.levels[l]
        return lv.kernel_map(lv, lv.stride)

    def down(self, l):       # k3 s2: level l -> l+1 (rows of level l+1)
        return self.levels[l + 1].kernel_map(self.levels[l], self.levels[l].stride)

    def up(self, l):         # transposed k3 s2: level l+1 -> l (rows of level l)
        return self.levels[l].kernel_map(self.levels[l + 1], -self.levels[l].stride)


class SparseUNet(nn.Module):
    """scenemodeling.py:147-237."""

    def __init__(self, dims=(64, 128, 128), n_groups=(4, 8, 8), n_res=(1, 2, 3)):
        super().__init__()
        self.n_levels = len(dims)
        self.res_down = nn.ModuleList(
            nn.Sequential(*[SparseResidual3d(dims[i], 'gn', n_groups[i]) for _ in range(n)])
            for i, n in enumerate(n_res))
        self.down = nn.ModuleList(
            nn.Sequential(SparseConvolution(dims[i - 1], dims[i], 3, 2), MinkowskiGroupNorm(n_groups[i], dims[i]),
                          ReLU()) for i in range(1, len(dims)))
        n_res, dims, n_groups = n_res[::-1], dims[::-1], n_groups[::-1]
        self.res_up = nn.ModuleList(
            nn.Sequential(*[SparseResidual3d(dims[i + 1], 'gn', n_groups[i + 1]) for _ in range(n)])
            for i, n in enumerate(n_res[1:]))
        self.up = nn.ModuleList()
        self.feat_adj = nn.ModuleList()
        for i in range(1, len(dims)):
            self.up.append(nn.Sequential(SparseConvolution(dims[i - 1], dims[i], 3, 2, transposed=True),
                                         MinkowskiGroupNorm(n_groups[i], dims[i]), ReLU()))
            self.feat_adj.append(nn.Sequential(SparseConvolution(2 * dims[i], dims[i], 1, 1),
                                               MinkowskiGroupNorm(n_groups[i], dims[i]), ReLU()))

    def forward(self, F, pts, idx, batch, res, scene=None):
        """F [Nv,dims[0]], pts [Nv,3] voxel centres, idx [Nv,3] int32, batch [Nv] int64 ->
        list (coarse -> fine) of dicts feats / pts / res / batch / idx / stride / sparse."""
        require_eval(self)
        if scene is None:
            scene = SparseScene(idx, batch, self.n_levels)
        nl = self.n_levels
        x = F.float().contiguous()
        ws = scene.ws
        for blk in self.res_down[0]:
            x = blk(x, scene.same(0), ws)
        xs = [x]
        for i in range(1, nl):
            conv, gn = self.down[i - 1][0], self.down[i - 1][1].gn
            wk, wp = conv.weights()
            x = ops.sparse_conv(x, scene.down(i - 1), wk, gn.weight.detach(), gn.bias.detach(), None, True, packed=wp,
                                workspace=ws)
            for blk in self.res_down[i]:
                x = blk(x, scene.same(i), ws)
            xs.append(x)
        out = [(xs[-1], nl - 1)]
        for i in range(nl - 1):
            l = nl - 2 - i  # target (finer) level
            conv, gn = self.up[i][0], self.up[i][1].gn
            wk, wp = conv.weights()
            up = ops.sparse_conv(x, scene.up(l), wk, gn.weight.detach(), gn.bias.detach(), None, True, packed=wp,
                                 workspace=ws)
            adj, gn = self.feat_adj[i][0], self.feat_adj[i][1].gn
            wk, wp = adj.weights()
            x = ops.concat_linear_gn_relu(up, xs[l], wk, gn.weight.detach(), gn.bias.detach(), packed=wp)
            for blk in self.res_up[i]:
                x = blk(x, scene.same(l), ws)
            out.append((x, l))

        origin = ops.batch_origin(pts.float().contiguous(), idx.int().contiguous(), batch.long().contiguous(),
                                  scene.n_batch, res)
        info = []
        for feats, l in out:
            lv = scene.levels[l]
            x_pts, x_idx, x_batch = ops.level_points(lv, origin, res)
            info.append({'feats': feats, 'pts': x_pts, 'res': lv.stride * res, 'batch': x_batch, 'idx': x_idx,
                         'stride': lv.stride, 'sparse': lv, 'origin': origin})
        return info
