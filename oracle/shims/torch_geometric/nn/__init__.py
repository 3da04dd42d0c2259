"""torch_geometric 1.6.3 `voxel_grid` -> torch_cluster 1.5.8 `grid_cluster`, restated
(call site /root/reference/mv3d/utils.py:45; semantics SURVEY.md A.3): the batch index is
appended as a fourth coordinate with cell size 1; per dimension the cell is
trunc((pos - start) / size) computed in the dtype of `pos` (fp32: subtract, then true
divide), the per-dimension cell count is trunc((end - start) / size) + 1, and the flat id
runs x fastest, batch slowest."""
import torch


def voxel_grid(pos, batch, size, start=None, end=None):
    pos = pos.unsqueeze(-1) if pos.dim() == 1 else pos
    dim = pos.size(1)
    size = size.tolist() if torch.is_tensor(size) else size
    start = start.tolist() if torch.is_tensor(start) else start
    end = end.tolist() if torch.is_tensor(end) else end
    size = [size] * dim if not isinstance(size, (list, tuple)) else list(size)
    pos = torch.cat([pos, batch.unsqueeze(-1).type_as(pos)], dim=-1)
    size = size + [1]
    start = list(start) + [0]
    end = list(end) + [batch.max().item()]
    size_t = torch.tensor(size, dtype=pos.dtype, device=pos.device)
    start_t = torch.tensor(start, dtype=pos.dtype, device=pos.device)
    end_t = torch.tensor(end, dtype=pos.dtype, device=pos.device)
    p = pos - start_t.unsqueeze(0)
    num_voxels = (end_t - start_t).true_divide(size_t).to(torch.long) + 1
    num_voxels = num_voxels.cumprod(0)
    num_voxels = torch.cat([torch.ones(1, dtype=torch.long), num_voxels], 0)[:size_t.numel()]
    out = p.true_divide(size_t.view(1, -1)).to(torch.long)
    out = out * num_voxels.view(1, -1)
    return out.sum(1)
