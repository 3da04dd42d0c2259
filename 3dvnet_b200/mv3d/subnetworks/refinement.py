"""PointFlow hypothesis decoder — drop-in for /root/reference/mv3d/subnetworks/refinement.py.
Trilinear sparse interpolation (csrc/sparse.cu) writes straight into the decoder operand
[n_pts, 8, 352]; the Conv1d stack runs as row-shifted gather-GEMMs (csrc/gemm.cu) and the last
convolution + softmax (+ expected offset) as one kernel (csrc/decoder.cu)."""
import torch
import torch.nn as nn

from ... import ops
from .._pack import PackCache, fold_bn, require_eval


def conv1d_bn_relu(in_channels, out_channels, kernel_size=3, stride=1, padding=1):
    return nn.Sequential(nn.Conv1d(in_channels, out_channels, kernel_size, stride, padding, bias=False),
                         nn.BatchNorm1d(out_channels), nn.ReLU(inplace=True))


class HypothesisDecoder(nn.Module):
    def __init__(self, in_dim=128 + 128 + 64, h_dim=256, kernel_size=3, padding=1):
        super().__init__()
        if kernel_size != 3 or padding != 1:
            raise NotImplementedError('the decoder kernels implement kernel_size=3, padding=1 (HYP_KSIZE/HYP_PAD)')
        self.in_dim = in_dim
        self.net = nn.Sequential(conv1d_bn_relu(in_dim, h_dim, kernel_size, 1, padding),
                                 conv1d_bn_relu(h_dim, h_dim, kernel_size, 1, padding),
                                 conv1d_bn_relu(h_dim, h_dim, kernel_size, 1, padding),
                                 nn.Conv1d(h_dim, 1, kernel_size, 1, padding))
        self._pack = PackCache()
        self._operand = None
        self._ws = None
        self._fused = None

    def _weights(self):
        def build():
            layers, fused = [], []
            for i in range(3):
                w = self.net[i][0].weight.detach().float().permute(2, 1, 0).contiguous()  # [3, Cin, Cout]
                layers.append((w,) + fold_bn(self.net[i][1]) + (ops.pack_weights(w.reshape(-1, w.shape[2])),))
                fused.append(ops.decoder_pack_weights(w))
            head = (self.net[3].weight.detach().float().contiguous(), float(self.net[3].bias.detach().cpu()))
            # the one-kernel decoder (csrc/decoder_fused.cu) applies when every layer could be packed for it
            self._fused = fused if all(f is not None for f in fused) else None
            return layers, head
        return self._pack.get([p for p in self.parameters()] + [b for b in self.buffers()], build, ops.gemm_mode())

    def operand(self, n_pts, device):
        """[n_pts, 8, in_dim] buffer whose padding row stays zero (reused between calls)."""
        op = self._operand
        if op is None or op.shape[0] < n_pts or op.device != device:
            op = self._operand = torch.zeros((n_pts, ops.ROWS_PER_POINT, self.in_dim), dtype=torch.float32,
                                             device=device)
        return op[:n_pts]

    def fill_levels(self, xs, pts, pts_batch, operand):
        """interpolated level features at channel offsets [fine | ... | coarse | var] (refinement.py:29-41)"""
        n_hyp = pts.shape[1]
        off = 0
        for x in xs[::-1]:  # xs is coarse -> fine; the finest level comes first in the operand
            ops.sparse_interp(pts, pts_batch, n_hyp, x['sparse'], x['origin'], x['res'], x['feats'], operand, off)
            off += x['feats'].shape[1]
        return off

    def run(self, operand, n_hyp, offset=None, want_prob=True):
        if n_hyp != ops.ROWS_PER_POINT - 1:
            # the Conv1d gather-GEMM zeroes only row 7 of the 8-row operand: fewer hypotheses would read stale rows
            raise NotImplementedError('HypothesisDecoder: the kernels implement 7 hypotheses (n=3, '
                                      'lightningmodel.py:187), got %d' % n_hyp)
        layers, head = self._weights()
        if self._fused is not None:
            return ops.decoder_fused(operand, layers[0][0].shape[1], self._fused, [l[1] for l in layers],
                                     [l[2] for l in layers], head[0], head[1], 0.0 if offset is None else offset,
                                     want_prob)
        x = operand
        if self._ws is None or self._ws.device != operand.device:
            self._ws = ops.sparse_conv_workspace(128, operand.device)   # zeroed once; launches leave it zero
        for w, scale, shift, packed in layers:
            x = ops.conv1d_bn_relu(x, w, scale, shift, packed=packed, workspace=self._ws)
        return ops.decoder_head(x, n_hyp, head[0], head[1], 0.0 if offset is None else offset, want_prob)

    def forward(self, xs, pts, pts_feat, pts_batch):
        """xs from SparseUNet; pts [Np,n_hyp,3]; pts_feat [Np,n_hyp,C]; pts_batch [Np] -> softmax [Np,n_hyp]"""
        require_eval(self)
        n_pts, n_hyp = pts.shape[:2]
        operand = self.operand(n_pts, pts.device)
        off = self.fill_levels(xs, pts.float().contiguous(), pts_batch.long().contiguous(), operand)
        operand[:, :n_hyp, off:off + pts_feat.shape[2]] = pts_feat
        _, prob = self.run(operand, n_hyp)
        return prob
