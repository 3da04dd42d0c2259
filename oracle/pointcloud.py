"""ORACLE (test infrastructure): point-level back-projection / re-projection / variance,
restated from /root/reference/mv3d/lightningmodel.py:132-174 (feature-rich point cloud)
and :187-235 (the 2n+1 PointFlow hypotheses per pixel)."""
import numpy as np
import torch
import torch.nn.functional as F

from .planesweep import lattice, projection_matrices, project_to_grid, group_variance


def image_points(n, img_size, plane_size):
    """[u; v; 1] on the plane lattice, x fastest (utils.py:67-83)."""
    u, v = lattice(img_size, plane_size)
    uu, vv = np.meshgrid(u, v)
    pts = np.stack((uu.reshape(-1), vv.reshape(-1), np.ones(uu.size, dtype=np.float32)))
    return torch.from_numpy(pts)[None].repeat(n, 1, 1)


def backproject(depth_flat, pts_img, R_T, K_inv, t):
    """X = R^T (K^-1 [u d, v d, d] - t) (lightningmodel.py:142-144)."""
    return torch.bmm(R_T, torch.bmm(K_inv, pts_img * depth_flat) - t.unsqueeze(-1))


def sample_variance(pts, img_feats, rotmats, tvecs, K, ref_src_edges, gather_idx, img_size):
    """Project pts [n_ref,3,N] into every edge's source image, bilinear fetch, variance over
    the edges of each reference (lightningmodel.py:147-169)."""
    P = projection_matrices(rotmats, tvecs, K)
    grid = project_to_grid(P[ref_src_edges[1]], pts[gather_idx], img_size)
    x = F.grid_sample(img_feats[ref_src_edges[1]], grid, mode='bilinear', align_corners=True).squeeze(3)
    return group_variance(x, gather_idx, pts.shape[0])


def feature_rich_pointcloud(depth, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges, img_size):
    """-> pts [n_ref*P,3], pts_feat [n_ref*P,C], pts_batch [n_ref*P] (lightningmodel.py:132-174)."""
    ref_idx, gather_idx = torch.unique(ref_src_edges[0], return_inverse=True)
    n = depth.shape[0]
    K_inv = torch.inverse(K[ref_idx])
    R_T = rotmats[ref_idx].transpose(2, 1)
    pts_img = image_points(n, img_size, depth.shape[1:]).type_as(depth)
    pts = backproject(depth.reshape(n, 1, -1), pts_img, R_T, K_inv, tvecs[ref_idx])
    x_var = sample_variance(pts, img_feats, rotmats, tvecs, K, ref_src_edges, gather_idx, img_size)
    P = depth.shape[1] * depth.shape[2]
    return (pts.transpose(2, 1).reshape(-1, 3), x_var.transpose(2, 1).reshape(-1, img_feats.shape[1]),
            depth_batch.unsqueeze(1).expand(n, P).reshape(-1))


def hypothesis_points(depth, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges, offset, n_side, img_size):
    """-> pts_hyp [n_ref*P, 2n+1, 3], pts_feat [n_ref*P, 2n+1, C], pts_batch [n_ref*P]
    (lightningmodel.py:187-235): hypothesis i is depth + i*offset, i = -n..n."""
    n = depth.shape[0]
    ref_idx, gather_idx = torch.unique(ref_src_edges[0], return_inverse=True)
    K_inv = torch.inverse(K[ref_idx])
    R_T = rotmats[ref_idx].transpose(2, 1)
    pts_img = image_points(n, img_size, depth.shape[1:]).type_as(depth)
    P = pts_img.shape[2]
    n_hyp = 2 * n_side + 1
    hyp = torch.empty((n, 3, n_hyp, P), dtype=torch.float32)
    for i in range(-n_side, n_side + 1):
        hyp[:, :, i + n_side] = backproject(depth.reshape(n, 1, -1) + i * offset, pts_img, R_T, K_inv, tvecs[ref_idx])
    x_var = sample_variance(hyp.view(n, 3, n_hyp * P), img_feats, rotmats, tvecs, K, ref_src_edges, gather_idx,
                            img_size)
    C = img_feats.shape[1]
    pts_feat = x_var.view(n, C, n_hyp, P).permute(0, 3, 2, 1).reshape(P * n, n_hyp, C)
    pts_hyp = hyp.permute(0, 3, 2, 1).reshape(P * n, n_hyp, 3)
    pts_batch = depth_batch.unsqueeze(1).expand(n, P).reshape(-1)
    return pts_hyp, pts_feat, pts_batch
