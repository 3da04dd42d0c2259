"""Multi-GPU path on real devices (skipped below 2 GPUs): a scene whose reference views are
sharded over 2 ranks, one NCCL all-gather of the point rows, must give every rank exactly the
sparse feature levels of the single-GPU model_scene."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    synth = importlib.import_module('3dvnet_b200.synth')
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    par = importlib.import_module('3dvnet_b200.parallel')
    img, plane = (64, 80), (16, 16)
    cfg = dict(depth_start=0.5, depth_interval=0.3, n_intervals=16, size=plane)
    b = synth.make_batch(1, 11, img, plane, 32, 2, 2, True, 3)       # 7 reference views
    net = lm.PL3DVNet(cfg, cfg, 0.3, feat_dim=32, img_size=img)
    net.load_state_dict(synth.make_params(0), strict=False)
    net = net.to(dev).eval()
    fq, R, t, K = b.feats_quarter.to(dev), b.rotmats.to(dev), b.tvecs.to(dev), b.K.to(dev)
    e = b.ref_src_edges
    ref_idx = torch.unique(e[0])
    depth = b.depth_images.to(dev)
    with torch.no_grad():
        start, end = par.shard_range(len(ref_idx), world, rank)
        xs = par.model_scene_sharded(net, depth[start:end].contiguous(), b.images_batch, fq, R, t, K, e)
        xs_ref = net.model_scene(depth, b.images_batch.to(dev)[ref_idx.to(dev)], fq, R, t, K, e)
    same = all(torch.equal(a['feats'], r['feats']) and torch.equal(a['idx'], r['idx']) for a, r in zip(xs, xs_ref))
    np.save(os.path.join(out_dir, 'r%d.npy' % rank), np.array([same, xs[-1]['feats'].shape[0]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_model_scene_sharded_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    importlib.import_module('3dvnet_b200.build').build()
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        same, n = np.load(tmp_path / ('r%d.npy' % r))
        assert same == 1 and n > 0
