"""ORACLE (test infrastructure): plane-sweep warp + variance, restated from
/root/reference/mv3d/utils.py:86-108 and /root/reference/mv3d/subnetworks/mvsnet.py:176-216.
"""
import numpy as np
import torch
import torch.nn.functional as F


def lattice(img_size, plane_size):
    """u_j, v_i of the plane lattice (utils.py:92-93): fp32 linspace over the FULL image."""
    H, W = img_size
    h, w = plane_size
    return (np.linspace(0, W - 1, w, dtype=np.float32), np.linspace(0, H - 1, h, dtype=np.float32))


def plane_sweep_points(depth_start, depth_interval, n_planes, R, t, K, img_size, plane_size, planes=None):
    """World coordinates of every frustum voxel of every image, [n,3,D*h*w], flattened
    [d][i][j] (utils.py:86-108). The pixel*depth product is formed in float64 and rounded to
    float32 once (utils.py:96-100); the two matrix products are float32.
    `planes` (a slice; test crops of volumes too large for the CPU) keeps only those depth planes of the
    full n_planes linspace - every retained voxel is computed exactly as in the full volume."""
    u, v = lattice(img_size, plane_size)
    z = np.linspace(depth_start, depth_start + (n_planes - 1) * depth_interval, n_planes, dtype=np.float32)
    if planes is not None:
        z = z[planes]
    uu, vv = np.meshgrid(u, v)  # [h,w]
    pix = np.stack([uu.astype(np.float64), vv.astype(np.float64), np.ones_like(uu, dtype=np.float64)])  # [3,h,w]
    pts = pix[:, None] * z.astype(np.float64)[None, :, None, None]  # [3,D,h,w]
    pts = torch.from_numpy(pts).float().reshape(3, -1)
    n = R.shape[0]
    pts = pts.unsqueeze(0).expand(n, 3, pts.shape[1]).to(R.dtype)
    cam = torch.bmm(torch.inverse(K), pts)
    return torch.bmm(R.transpose(2, 1), cam - t[..., None])


def projection_matrices(R, t, K):
    """P = K [R|t] with the full-resolution K (mvsnet.py:196-197)."""
    return torch.bmm(K, torch.cat((R, t[..., None]), dim=2))


def project_to_grid(P_edge, pts_edge, img_size):
    """q = P [X;1]; z = |q_z| + 1e-8; normalise by the FULL image size (mvsnet.py:199-206).
    pts_edge [E,3,N] -> grid [E,N,1,2] in [-1,1]."""
    H, W = img_size
    ones = torch.ones((pts_edge.shape[0], 1, pts_edge.shape[2]), dtype=pts_edge.dtype)
    q = torch.bmm(P_edge, torch.cat((pts_edge, ones), dim=1))
    zb = q[:, 2].abs() + 1e-8
    xy = q[:, :2] / zb[:, None]
    grid = xy.transpose(2, 1).reshape(xy.shape[0], -1, 1, 2).clone()
    grid[..., 0] = (grid[..., 0] / float(W - 1)) * 2 - 1.0
    grid[..., 1] = (grid[..., 1] / float(H - 1)) * 2 - 1.0
    return grid


def group_variance(x, gather_idx, n_ref):
    """mean and mean-of-squares over the edges of each reference, divisor = edge count
    (torch_scatter 'mean'), var = E[x^2] - E[x]^2 (mvsnet.py:214-216)."""
    cnt = torch.zeros(n_ref, dtype=x.dtype).index_add_(0, gather_idx, torch.ones(gather_idx.shape, dtype=x.dtype))
    cnt = cnt.clamp_(min=1).view(-1, *([1] * (x.dim() - 1)))
    shape = (n_ref,) + tuple(x.shape[1:])
    avg = torch.zeros(shape, dtype=x.dtype).index_add_(0, gather_idx, x) / cnt
    avg_sq = torch.zeros(shape, dtype=x.dtype).index_add_(0, gather_idx, x ** 2) / cnt
    return avg_sq - avg ** 2


def planesweep_var(feats_quarter, rotmats, tvecs, K, ref_src_edges, depth_start, depth_interval, n_planes,
                   img_size, plane_size, planes=None):
    """x_var [n_ref,C,D,h,w] exactly as MVSNet.forward builds it (mvsnet.py:179,187-216)."""
    ref_idx, gather_idx = torch.unique(ref_src_edges[0], return_inverse=True)
    pts = plane_sweep_points(depth_start, depth_interval, n_planes, rotmats, tvecs, K, img_size, plane_size, planes)
    if planes is not None:
        n_planes = len(range(*planes.indices(n_planes)))
    P = projection_matrices(rotmats, tvecs, K)
    grid = project_to_grid(P[ref_src_edges[1]], pts[ref_src_edges[0]], img_size)
    x = F.grid_sample(feats_quarter[ref_src_edges[1]], grid, mode='bilinear', align_corners=True)
    x = x.squeeze(3).view(-1, feats_quarter.shape[1], n_planes, *plane_size)
    return group_variance(x, gather_idx, len(ref_idx))


# --------------------------------------------------------------------------------------
# Independent scalar restatement of the sampler (pure numpy), used on small cases to pin
# the semantics of F.grid_sample(bilinear, zeros, align_corners=True) that the kernel
# reimplements: ix = ((g+1)/2) (Wf-1); the four taps nw/ne/sw/se are weighted by the
# opposite-corner areas and each is dropped independently when out of bounds.
def bilinear_zeros_numpy(feat, gx, gy):
    C, Hf, Wf = feat.shape
    gx = np.asarray(gx, dtype=np.float32)
    gy = np.asarray(gy, dtype=np.float32)
    ix = ((gx + np.float32(1)) / np.float32(2)) * np.float32(Wf - 1)
    iy = ((gy + np.float32(1)) / np.float32(2)) * np.float32(Hf - 1)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    out = np.zeros((C, gx.shape[0]), dtype=np.float32)
    for (dx, dy) in ((0, 0), (1, 0), (0, 1), (1, 1)):
        xt = x0 + dx
        yt = y0 + dy
        wx = (x0 + 1 - ix) if dx == 0 else (ix - x0)
        wy = (y0 + 1 - iy) if dy == 0 else (iy - y0)
        ok = (xt >= 0) & (xt <= Wf - 1) & (yt >= 0) & (yt <= Hf - 1)
        xi = np.where(ok, xt, 0).astype(np.int64)
        yi = np.where(ok, yt, 0).astype(np.int64)
        val = feat[:, yi, xi] * ok[None].astype(np.float32)
        out += val * (wx * wy).astype(np.float32)[None]
    return out


def planesweep_var_numpy(feats_quarter, rotmats, tvecs, K, ref_src_edges, depth_start, depth_interval, n_planes,
                         img_size, plane_size):
    """Same quantity through the scalar sampler and a plain per-edge loop (small cases)."""
    fq = feats_quarter.numpy()
    ref_idx, gather_idx = torch.unique(ref_src_edges[0], return_inverse=True)
    pts = plane_sweep_points(depth_start, depth_interval, n_planes, rotmats, tvecs, K, img_size, plane_size)
    P = projection_matrices(rotmats, tvecs, K)
    grid = project_to_grid(P[ref_src_edges[1]], pts[ref_src_edges[0]], img_size).numpy()
    n_ref = len(ref_idx)
    C = fq.shape[1]
    N = grid.shape[1]
    s = np.zeros((n_ref, C, N), dtype=np.float32)
    s2 = np.zeros((n_ref, C, N), dtype=np.float32)
    cnt = np.zeros(n_ref, dtype=np.float32)
    for e in range(ref_src_edges.shape[1]):
        x = bilinear_zeros_numpy(fq[int(ref_src_edges[1, e])], grid[e, :, 0, 0], grid[e, :, 0, 1])
        r = int(gather_idx[e])
        s[r] += x
        s2[r] += x * x
        cnt[r] += 1
    cnt = np.maximum(cnt, 1)[:, None, None]
    avg = s / cnt
    var = s2 / cnt - avg * avg
    return torch.from_numpy(var.reshape(n_ref, C, n_planes, *plane_size))
