"""SURVEY.md section 8d's full pipeline (PL3DVNet.full_pass: backbone + FPN -> hot path -> three PropagationNets,
/root/reference/mv3d/eval-3dvnet.py:58-125) against the CPU: the backbone hand-off (section 8f.2: channels-last cuDNN
feature maps consumed by the warp kernels without a transposition pass) and the composition of the three stages."""
import copy
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def setup():
    importlib.import_module('3dvnet_b200.build').build()
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    synth = importlib.import_module('3dvnet_b200.synth')
    img, plane, D = (64, 80), (16, 16), 16
    cfg = dict(depth_start=0.5, depth_interval=0.3, n_intervals=D, size=plane)
    b = synth.make_batch(1, 5, img, plane, 32, 1, 1, False, 4, with_images=True)
    params = synth.make_params(0)
    torch.manual_seed(7)
    net = lm.PL3DVNet(cfg, cfg, 0.3, feat_dim=32, img_size=img)
    net.load_state_dict(params, strict=False)
    net_cpu = copy.deepcopy(net).eval()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False     # fp32 convolutions on both sides
    yield dict(net=net.to(DEV).eval(), net_cpu=net_cpu, b=b, cfg=cfg, img=img, params=params)
    torch.backends.cudnn.allow_tf32 = tf32


def _run_gpu(s, offsets):
    b = s['b']
    return s['net'].full_pass(b.images.to(DEV), b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV), b.ref_src_edges,
                              b.images_batch.to(DEV), offsets, s['cfg'])


def test_backbone_handoff_matches_cpu(setup):
    """the channels-last cuDNN backbone + FPN produce the CPU modules' feature maps (fp32 rounding), and the NHWC
    storage they arrive in is consumed as is: hot_path on them == hot_path on a contiguous NCHW copy, bit for bit"""
    s = setup
    out = _run_gpu(s, [])
    with torch.no_grad():
        fh, fq, _, _, _ = s['net_cpu'].mvsnet.feat_shrinker(*s['net_cpu'].mvsnet.feat_extractor(s['b'].images))
    for got, want in ((out['feats_quarter'], fq), (out['feats_half'], fh)):
        assert got.shape == want.shape
        err = (got.cpu() - want).abs().max().item() / want.abs().max().item()
        assert err <= 2e-5, err
    fq_gpu = out['feats_quarter']
    assert fq_gpu.is_contiguous(memory_format=torch.channels_last)
    b = s['b']
    args = (b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV), b.ref_src_edges, b.images_batch.to(DEV), s['cfg'], [[0.05]])
    with torch.no_grad():
        d_cl = s['net'].hot_path(fq_gpu, *args)
        d_nchw = s['net'].hot_path(fq_gpu.contiguous(), *args)
    assert torch.equal(d_cl, d_nchw)


def test_full_pass_matches_cpu_pipeline(setup):
    from oracle import pipeline, upsample
    s = setup
    b, net_cpu = s['b'], s['net_cpu']
    offsets = [[0.05, 0.025]]
    out = _run_gpu(s, offsets)
    with torch.no_grad():
        fh, fq, _, _, _ = net_cpu.mvsnet.feat_shrinker(*net_cpu.mvsnet.feat_extractor(b.images))
        ref_idx = torch.unique(b.ref_src_edges[0])
        sd = net_cpu.state_dict()
        pq, ph, pf = (pipeline.sub(sd, k) for k in ('refine_quarter.', 'refine_half.', 'refine_full.'))
        # stage by stage on the GPU's own hand-offs (continuous maps): initial depth, then the cascade
        d0_cpu = pipeline.initial_depth(out['feats_quarter'].cpu().contiguous(), b.rotmats, b.tvecs, b.K, b.ref_src_edges,
                                        s['cfg'], s['img'], s['params'])
        d0_gpu = _run_gpu(s, [])['ref'].cpu()
        np.testing.assert_allclose(d0_gpu.numpy(), d0_cpu.numpy(), rtol=0, atol=1e-4)
        final_tf = upsample.upsample_cascade(out['ref'].cpu(), out['feats_quarter'].cpu()[ref_idx],
                                             out['feats_half'].cpu()[ref_idx], b.images[ref_idx], pq, ph, pf)
        np.testing.assert_allclose(out['final'].cpu().numpy(), final_tf.numpy(), rtol=0, atol=2e-4)
        # free-running, everything on the CPU from the images: the bulk of the pixels agrees
        depth = pipeline.refine(pipeline.initial_depth(fq, b.rotmats, b.tvecs, b.K, b.ref_src_edges, s['cfg'], s['img'],
                                                       s['params']),
                                b.images_batch[ref_idx], fq, b.rotmats, b.tvecs, b.K, b.ref_src_edges, 0.3, s['img'],
                                s['params'], offsets_list=offsets)
        final = upsample.upsample_cascade(depth, fq[ref_idx], fh[ref_idx], b.images[ref_idx], pq, ph, pf)
    assert out['final'].shape == final.shape == (len(ref_idx),) + s['img']
    rel = (out['final'].cpu() - final).abs() / (final.abs() + 1e-7)
    assert torch.median(rel).item() <= 1e-4 and rel.mean().item() <= 2e-3, (torch.median(rel).item(), rel.mean().item())


def test_process_scene_plugin_equals_full_pass(setup):
    """the reference's pred_func plugin (eval-3dvnet.py:26-127): same signature / return, same depth as full_pass
    whatever the backbone chunking"""
    from argparse import Namespace
    ev = importlib.import_module('3dvnet_b200.mv3d.eval_3dvnet')
    s = setup
    b = s['b']
    offsets = [[0.05, 0.025]]
    want = _run_gpu(s, offsets)['final'].cpu().numpy()
    batch = Namespace(images=b.images, rotmats=b.rotmats, tvecs=b.tvecs, K=b.K, ref_src_edges=b.ref_src_edges)
    old = ev.BACKBONE_BATCH
    try:
        for chunk in (64, 2):
            ev.BACKBONE_BATCH = chunk
            got, a, c = ev.process_scene(batch, None, None, s['net'], depth_config=s['cfg'], offsets_list=offsets)
            assert a is None and c is None and got.shape == want.shape
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-5)
    finally:
        ev.BACKBONE_BATCH = old
