"""ORACLE (test infrastructure): voxelisation of the back-projected points, restated from
/root/reference/mv3d/utils.py:38-64 together with the third-party kernel it calls:
torch_geometric 1.6.3 `voxel_grid` -> torch_cluster 1.5.8 `grid_cluster` (absent here;
algorithm per SURVEY.md A.3). All index arithmetic is done exactly as there: fp32
subtract, fp32 true divide, truncation; int64 ids. Outputs must be bit-exact."""
import numpy as np


def voxel_ids(pts, batch, edge_len):
    """1-D voxel id of every point incl. its batch slot (utils.py:39-45 + grid_cluster)."""
    pts = np.asarray(pts, dtype=np.float32)
    batch = np.asarray(batch, dtype=np.int64)
    e = np.float32(edge_len)
    bmin = pts.min(axis=0)
    bmax = pts.max(axis=0)
    # grid_cluster: cells per dimension = trunc((end - start) / size) + 1, batch dim has size 1
    n_cells = ((bmax - bmin) / e).astype(np.int64) + 1
    cell = ((pts - bmin[None]) / e).astype(np.int64)  # fp32 subtract, fp32 divide, truncate
    ids = cell[:, 0] + cell[:, 1] * n_cells[0] + cell[:, 2] * (n_cells[0] * n_cells[1])
    ids = ids + batch.astype(np.float32).astype(np.int64) * (n_cells[0] * n_cells[1] * n_cells[2])
    return ids, bmin, bmax


def voxelize(pts, batch, edge_len):
    """-> anchor_pts [Nv,3] f32, anchor_idx3d [Nv,3] i32, anchor_batch [Nv] i64,
    anchor_pts_edges [2,N] i64 (utils.py:38-64).

    The reference decodes the ids with grid_size = ceil((max-min)/edge) (utils.py:41-42,
    53-57) although grid_cluster strides them with trunc(.)+1; both are reproduced as
    written (they differ only when an extent is an exact multiple of the edge)."""
    pts = np.asarray(pts, dtype=np.float32)
    batch = np.asarray(batch, dtype=np.int64)
    e = np.float32(edge_len)
    ids, bmin, bmax = voxel_ids(pts, batch, edge_len)
    grid_size = np.ceil((bmax - bmin) / e).astype(np.int64)
    max_grid_idx = grid_size[0] * grid_size[1] * grid_size[2]

    anchor_idx, inv = np.unique(ids, return_inverse=True)  # ascending ids, like torch.unique
    n_anchor = anchor_idx.shape[0]
    anchor_batch = np.full(n_anchor, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(anchor_batch, inv, batch)

    a = anchor_idx - anchor_batch * max_grid_idx
    gxy = grid_size[0] * grid_size[1]
    idx3d = np.zeros((n_anchor, 3), dtype=np.int32)
    idx3d[:, 2] = (a // gxy).astype(np.int32)
    rem = a - idx3d[:, 2].astype(np.int64) * gxy
    idx3d[:, 1] = (rem // grid_size[0]).astype(np.int32)
    idx3d[:, 0] = (rem % grid_size[0]).astype(np.int32)
    # int32 * python float -> float32 tensor in torch; the sum order is (idx*e + bmin) + e/2
    anchor_pts = (idx3d.astype(np.float32) * e + bmin[None]) + np.float32(edge_len / 2.0)

    n_batch = int(anchor_batch.max()) + 1
    min_idx = np.zeros((n_batch, 3), dtype=np.int32)  # scatter-min leaves empty slots at 0
    filled = np.zeros(n_batch, dtype=bool)
    for b in range(n_batch):
        m = anchor_batch == b
        if m.any():
            min_idx[b] = idx3d[m].min(axis=0)
            filled[b] = True
    idx3d = idx3d - min_idx[anchor_batch]
    edges = np.stack([inv.astype(np.int64), np.arange(pts.shape[0], dtype=np.int64)])
    return anchor_pts.astype(np.float32), idx3d, anchor_batch, edges
