"""pytorch_lightning 1.1.2 stub: LightningModule = nn.Module + save_hyperparameters()
(captures the caller's constructor arguments into self.hparams) + no-op log."""
import inspect
import torch


class _HParams(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class LightningModule(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.current_epoch = 0

    def save_hyperparameters(self):
        frame = inspect.currentframe().f_back
        args = inspect.getargvalues(frame)
        hp = _HParams()
        for name in args.args:
            if name != 'self':
                hp[name] = args.locals[name]
        object.__setattr__(self, '_hparams', hp)

    @property
    def hparams(self):
        return self._hparams

    @property
    def device(self):
        return next(self.parameters()).device

    def log(self, *a, **k):
        pass


class Trainer(object):
    def __init__(self, *a, **k):
        raise RuntimeError('stub')
