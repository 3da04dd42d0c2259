"""ORACLE (test infrastructure, never the product path): CPU restatement of the
MinkowskiEngine 0.5.0 operations the reference calls.

MinkowskiEngine is a third-party dependency that is NOT under /root/reference (installed
from git master, "only tested with 0.5.0", /root/reference/README.md:28-32) and cannot be
installed here, so its published algorithm is restated from SURVEY.md Appendix A.4:
**parity unpinned** against ME itself.  What pins it instead:

* the reference's own call sites (scenemodeling.py:10-12,27-28,36-39,160-162,181-188,194,
  206,213-216; refinement.py:26,39), which this module's API mirrors through
  oracle/shims/MinkowskiEngine;
* the reference's dense restatement of the interpolation,
  HypothesisDecoder.forward_forloop (refinement.py:46-97), checked in
  tests/test_oracle_minkowski.py;
* dense torch conv3d / conv_transpose3d on densified grids (same test file), which is what
  a generalised sparse convolution is defined to equal on its active sites.

Conventions (A.4): coordinates are int [N,4] = [batch,x,y,z]; kernel offsets enumerate
x fastest, k = (dx+1) + 3(dy+1) + 9(dz+1); weights W[k] are [Cin,Cout]; no bias;
strided maps are floor(c / s) * s; the transposed convolution reuses the existing finer
map and the forward strided kernel map with in/out swapped; interpolation takes the 8
corners lower + {0,ts}^3 with weights prod(1 - |q - c| / ts), missing voxels contribute 0.
Row order of every coordinate map here is ascending (batch, z, y, x) — the order of the
reference's voxel ids (utils.py:45-48) — which ME never exposes to the reference.
"""
import numpy as np
import torch

_BIAS = 1 << 15
_BITS = 16

OFFSETS = np.array([[dx, dy, dz] for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)], dtype=np.int64)


def encode(coords):
    """[N,4] int (b,x,y,z) -> int64 key ordered by (b, z, y, x)."""
    c = np.asarray(coords, dtype=np.int64)
    return (((c[:, 0] << _BITS | (c[:, 3] + _BIAS)) << _BITS | (c[:, 2] + _BIAS)) << _BITS) | (c[:, 1] + _BIAS)


class CoordMap(object):
    """Unique coordinates at one tensor stride, rows in ascending key order."""

    def __init__(self, coords, stride, presorted=False):
        coords = np.asarray(coords, dtype=np.int64)
        keys = encode(coords)
        if not presorted:
            order = np.argsort(keys, kind='stable')
            keys = keys[order]
            coords = coords[order]
            assert keys.size < 2 or (np.diff(keys) > 0).all(), 'coordinates must be unique'
        self.coords = coords
        self.keys = keys
        self.stride = int(stride)

    def __len__(self):
        return self.coords.shape[0]

    def lookup(self, query):
        """row index of every query coordinate, -1 where absent."""
        q = encode(query)
        pos = np.searchsorted(self.keys, q)
        pos = np.minimum(pos, max(len(self.keys) - 1, 0))
        hit = (self.keys[pos] == q) if len(self.keys) else np.zeros(q.shape, bool)
        return np.where(hit, pos, -1)

    def strided(self, s=2):
        """Output map of a stride-s convolution: unique(floor(c / (s ts)) * (s ts))."""
        ns = self.stride * s
        c = self.coords.copy()
        c[:, 1:] = np.floor_divide(c[:, 1:], ns) * ns
        keys = encode(c)
        _, first = np.unique(keys, return_index=True)
        return CoordMap(c[first], ns)


def kernel_map(in_map, out_map, step):
    """For every kernel offset k the (in_row, out_row) pairs with
    in_coord = out_coord + OFFSETS[k] * step."""
    pairs = []
    for k in range(27):
        q = out_map.coords.copy()
        q[:, 1:] += OFFSETS[k] * step
        idx = in_map.lookup(q)
        out_rows = np.nonzero(idx >= 0)[0]
        pairs.append((idx[out_rows], out_rows))
    return pairs


def sparse_conv(feat, weight, pairs, n_out):
    """out[o] = sum_k sum_{(i,o) in pairs[k]} feat[i] @ W[k]  (gather, GEMM, scatter-add)."""
    out = torch.zeros((n_out, weight.shape[2]), dtype=feat.dtype)
    for k, (i_rows, o_rows) in enumerate(pairs):
        if len(i_rows):
            out.index_add_(0, torch.from_numpy(o_rows), feat[torch.from_numpy(i_rows)] @ weight[k])
    return out


def conv3(feat, cmap, weight, stride=1):
    """MinkowskiConvolution(kernel_size=3, stride in {1,2}, dimension=3, bias=False)."""
    if stride == 1:
        out_map = cmap
    else:
        out_map = cmap.strided(stride)
    pairs = kernel_map(cmap, out_map, cmap.stride)
    return sparse_conv(feat, weight, pairs, len(out_map)), out_map


def conv3_transpose(feat, coarse_map, fine_map, weight):
    """MinkowskiConvolutionTranspose(kernel_size=3, stride=2): output lives on the existing
    finer map; pairs are the forward (fine -> coarse) map with in/out swapped."""
    fwd = kernel_map(fine_map, coarse_map, fine_map.stride)  # (fine_row, coarse_row)
    pairs = [(c_rows, f_rows) for (f_rows, c_rows) in fwd]
    return sparse_conv(feat, weight, pairs, len(fine_map))


def interpolate(cmap, feat, q):
    """MinkowskiInterpolation: q is float [Nq,4] = [batch, x, y, z] in base-voxel units."""
    q = torch.as_tensor(q)
    ts = float(cmap.stride)
    b = q[:, 0].round().long().numpy()
    qs = q[:, 1:]
    lower = torch.floor(qs / ts) * ts
    out = torch.zeros((q.shape[0], feat.shape[1]), dtype=feat.dtype)
    lower_i = lower.long().numpy()
    for corner in range(8):
        off = np.array([(corner >> 0) & 1, (corner >> 1) & 1, (corner >> 2) & 1], dtype=np.int64) * cmap.stride
        nb = lower_i + off
        w = torch.ones(q.shape[0], dtype=feat.dtype)
        for d in range(3):
            w = w * (1 - (qs[:, d] - torch.from_numpy(nb[:, d]).to(feat.dtype)).abs() / ts)
        rows = cmap.lookup(np.concatenate([b[:, None], nb], axis=1))
        hit = np.nonzero(rows >= 0)[0]
        if len(hit):
            hit_t = torch.from_numpy(hit)
            out.index_add_(0, hit_t, feat[torch.from_numpy(rows[hit])] * w[hit_t, None])
    return out
