import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def synth():
    return importlib.import_module('3dvnet_b200.synth')


@pytest.fixture(scope='session')
def pkg():
    return importlib.import_module('3dvnet_b200')


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: g[k] for k in g.files}


def golden_inputs(g):
    """torch views of a pipeline golden file's inputs + its depth config."""
    t = {k: torch.from_numpy(g[k]) for k in ('feats_quarter', 'rotmats', 'tvecs', 'K', 'ref_src_edges',
                                              'images_batch')}
    cfg = dict(depth_start=float(g['depth_start']), depth_interval=float(g['depth_interval']),
               n_intervals=int(g['D']), size=tuple(int(v) for v in g['plane']))
    img_size = tuple(int(v) for v in g['img_size'])
    return t, cfg, img_size


@pytest.fixture
def exact_warp():
    """the plane-sweep kernel's verification mode (bit-exact x_var) for the duration of a test"""
    ops = importlib.import_module('3dvnet_b200.ops')
    old = ops.warp_mode()
    ops.set_warp_mode('exact')
    yield
    ops.set_warp_mode(old)
