"""Minimal `Data` so /root/reference/mv3d/dsets/batch.py imports."""
import torch


class Data(object):
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __inc__(self, key, value):
        return 0

    def __cat_dim__(self, key, value):
        return 0

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Dataset(object):
    pass
