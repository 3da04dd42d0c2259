"""Checkpoint interchange: the drop-in PL3DVNet must expose exactly
the state_dict keys and shapes of the reference's PL3DVNet, so that the published checkpoint
(README.md:91, eval-3dvnet.py:134) loads by name.  The schema fixture was recorded from the
unmodified reference modules by oracle/make_schema.py."""
import importlib
import json
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _net():
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    cfg = dict(depth_start=0.5, depth_interval=0.05, n_intervals=96, size=(56, 56))
    return lm.PL3DVNet(cfg, cfg, 0.08, feat_dim=32, img_size=(256, 320))


def test_state_dict_matches_reference_schema():
    schema = json.load(open(os.path.join(HERE, 'golden', 'state_dict_schema.json')))
    mine = {k: list(v.shape) for k, v in _net().state_dict().items()}
    assert sorted(mine) == sorted(schema)
    assert {k: mine[k] for k in schema} == schema


def test_lightning_checkpoint_round_trip(tmp_path):
    """a Lightning-style checkpoint {'state_dict', 'hyper_parameters'} loads through load_from_checkpoint"""
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    net = _net()
    path = os.path.join(str(tmp_path), 'epoch=0-step=1.ckpt')
    hp = dict(depth_train=net.depth_train, depth_test=net.depth_test, edge_len=0.08, feat_dim=32, img_size=(256, 320))
    torch.save({'state_dict': net.state_dict(), 'hyper_parameters': hp}, path)
    back = lm.PL3DVNet.load_from_checkpoint(path)
    assert back.hparams.edge_len == 0.08 and back.edge_len == 0.08
    for k, v in net.state_dict().items():
        assert torch.equal(v, back.state_dict()[k]), k
