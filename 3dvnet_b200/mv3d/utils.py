"""Geometry helpers — drop-in for the hot-path functions of /root/reference/mv3d/utils.py
(voxelize :38-64, slice_edges :32-35, plane-sweep / image point builders :67-108,
freeze_batchnorm). voxelize runs on the B200 (csrc/voxelize.cu) and is bit-exact."""
import numpy as np
import torch

from .. import ops


def freeze_batchnorm(module):
    if isinstance(module, torch.nn.modules.batchnorm._BatchNorm):
        module.eval()
        for p in module.parameters():
            p.requires_grad = False


def slice_edges(edges, index_start, index_end, slice_dim=0):
    """edges whose ``slice_dim`` row lies in [index_start, index_end) (utils.py:32-35)."""
    row = edges[slice_dim]
    return edges[:, (row >= index_start) & (row < index_end)]


def voxelize(pts, pts_batch, edge_len, return_aux=False):
    """-> anchor_pts [Nv,3] f32, anchor_idx3d [Nv,3] i32, anchor_batch [Nv] i64,
    anchor_pts_edges [2,N] i64 (row 0 = voxel of each point, row 1 = arange(N))."""
    a_pts, a_idx, a_batch, p_anchor, grid = ops.voxelize(pts.contiguous(), pts_batch.contiguous(), edge_len)
    edges = torch.stack((p_anchor.long(), torch.arange(pts.shape[0], dtype=torch.long, device=pts.device)), dim=0)
    if return_aux:
        return a_pts, a_idx, a_batch, edges, p_anchor, grid
    return a_pts, a_idx, a_batch, edges


def _lattice(img_size, plane_size):
    return (np.linspace(0, img_size[1] - 1, plane_size[1], dtype=np.float32),
            np.linspace(0, img_size[0] - 1, plane_size[0], dtype=np.float32))


def build_img_pts(img_size=(240, 320), plane_size=(56, 56)):
    """[u; v; 1] on the plane lattice, x fastest: numpy [3, h*w] (utils.py:67-77)."""
    u, v = _lattice(img_size, plane_size)
    uu, vv = np.meshgrid(u, v)
    return np.stack((uu.reshape(-1), vv.reshape(-1), np.ones(uu.size, dtype=np.float32)))


def batched_build_img_pts_tensor(n_batch, img_size=(240, 320), plane_size=(56, 56)):
    return torch.from_numpy(build_img_pts(img_size, plane_size))[None].repeat(n_batch, 1, 1)


def batched_build_plane_sweep_volume_tensor(depth_start, depth_interval, n_planes, R, t, K, img_size=(240, 320),
                                            plane_size=(56, 56)):
    """World points of every frustum voxel, [n,3,D*h*w] flattened [d][y][x] (utils.py:86-108).
    API compatibility only: the cost-volume kernel derives these coordinates in registers."""
    u, v = _lattice(img_size, plane_size)
    z = np.linspace(depth_start, depth_start + (n_planes - 1) * depth_interval, n_planes, dtype=np.float32)
    uu, vv = np.meshgrid(u, v)
    pix = np.stack([uu.astype(np.float64), vv.astype(np.float64), np.ones(uu.shape)])
    pts = torch.from_numpy(pix[:, None] * z.astype(np.float64)[None, :, None, None]).float().reshape(1, 3, -1)
    pts = pts.to(R.device).expand(R.shape[0], 3, -1)
    return torch.bmm(R.transpose(2, 1), torch.bmm(torch.inverse(K), pts) - t[..., None])
