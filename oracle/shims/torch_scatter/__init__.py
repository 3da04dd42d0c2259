"""Shim of torch_scatter 2.0.5 `scatter` for the call sites the reference uses
(mvsnet.py:214-215, lightningmodel.py:167-168,227-228, utils.py:50,61,
scenemodeling.py:129-141, refinement.py:33). Semantics: SURVEY.md A.2."""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand_as(src)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    if dim < 0:
        dim = src.dim() + dim
    idx = _broadcast(index, src, dim)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    if reduce in ('sum', 'add'):
        return torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, idx, src)
    if reduce == 'mean':
        s = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add_(dim, idx, src)
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).scatter_add_(
            0, index, torch.ones(index.shape, dtype=src.dtype, device=src.device)).clamp_(min=1)
        cnt = _broadcast(cnt, s, dim)
        if src.is_floating_point():
            return s.true_divide_(cnt)
        return s.div_(cnt, rounding_mode='floor')
    if reduce in ('min', 'max'):
        res = torch.zeros(shape, dtype=src.dtype, device=src.device)
        return res.scatter_reduce_(dim, idx, src, 'amin' if reduce == 'min' else 'amax', include_self=False)
    raise ValueError(reduce)
