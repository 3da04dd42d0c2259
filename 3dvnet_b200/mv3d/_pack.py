"""Derived, kernel-friendly copies of module parameters (folded BatchNorm, transposed
weights), cached until a parameter is modified in place or re-assigned."""
import torch


class PackCache(object):
    def __init__(self):
        self._key = None
        self._val = None

    def get(self, tensors, build, extra=None):
        key = tuple((t.data_ptr(), t._version, t.device) for t in tensors) + (extra,)
        if key != self._key:
            with torch.no_grad():
                self._val = build()
            self._key = key
        return self._val


def fold_bn(bn):
    """eval-mode BatchNorm as y = x * scale + shift."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias - bn.running_mean * scale
    return scale.float().contiguous(), shift.float().contiguous()


def require_eval(module):
    if module.training:
        raise NotImplementedError(
            '%s: this module has a forward (inference) kernel only - call .eval(). The differentiable pieces are '
            'MVSNet (warp + variance as torch.autograd.Function, mv3d/functional.py, regulariser through cuDNN) and '
            'the point-level variance features' % type(module).__name__)
