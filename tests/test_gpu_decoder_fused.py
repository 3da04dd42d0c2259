"""The one-kernel PointFlow decoder (csrc/decoder_fused.cu) against the reference module definition evaluated in
float64 on the CPU (refinement.py:17-25,42-44: three Conv1d(k=3,pad=1)+BN+ReLU, Conv1d(128->1), softmax over the
hypotheses) and against the per-layer gather-GEMM path it replaces."""
import copy
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def mods():
    importlib.import_module('3dvnet_b200.build').build()
    return dict(ops=importlib.import_module('3dvnet_b200.ops'),
                ref=importlib.import_module('3dvnet_b200.mv3d.subnetworks.refinement'))


def _decoder(mods, in_dim, seed):
    g = torch.Generator().manual_seed(seed)
    dec = mods['ref'].HypothesisDecoder(in_dim, 128, 3, 1)
    with torch.no_grad():
        for p in dec.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (1.5 / np.sqrt(max(1, p[0].numel()))))
        for m in dec.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_var.copy_(torch.rand(m.bias.shape, generator=g) + 0.5)
        dec.net[3].weight.mul_(4.0)   # a sharp softmax is sensitive to every layer
    return dec.eval()


def _reference(dec, x7, offset):
    """float64 CPU evaluation of the reference module: x7 [n_pts, 7, C] -> prob [n_pts,7], expected offset"""
    net = copy.deepcopy(dec.net).double()
    with torch.no_grad():
        logits = net(x7.double().transpose(1, 2)).squeeze(1)
        prob = torch.softmax(logits, dim=1)
    vals = torch.linspace(-3 * offset, 3 * offset, 7).double()
    return prob, (prob * vals[None]).sum(1)


@pytest.mark.parametrize('n_pts,in_dim,mode', [(3136, 352, 'tf32x3'), (1, 352, 'tf32x3'), (37, 352, 'tf32x3'),
                                              (500, 64, 'tf32x3'), (2048, 352, 'tf32')])
def test_fused_decoder_matches_float64_reference(mods, n_pts, in_dim, mode):
    ops = mods['ops']
    old = ops.gemm_mode()
    ops.set_gemm_mode(mode)
    try:
        dec = _decoder(mods, in_dim, 1).to(DEV)
        g = torch.Generator().manual_seed(n_pts)
        x = torch.randn(n_pts, 8, in_dim, generator=g)
        x[:, 7] = 1e3   # the padding row must be IGNORED by the fused kernel, whatever it holds
        prob_ref, off_ref = _reference(dec.cpu(), x[:, :7], 0.05)
        dec = dec.to(DEV)
        off, prob = dec.run(x.to(DEV), 7, 0.05, want_prob=True)
        assert dec._fused is not None, 'the fused kernel was not selected'
        tol = 2e-5 if mode == 'tf32x3' else 1e-2
        np.testing.assert_allclose(prob.cpu().double().numpy(), prob_ref.numpy(), rtol=0, atol=tol)
        np.testing.assert_allclose(off.cpu().double().numpy(), off_ref.numpy(), rtol=0, atol=tol * 0.15)
        np.testing.assert_allclose(prob.sum(1).cpu().numpy(), 1.0, atol=1e-5)
        again, _ = dec.run(x.to(DEV), 7, 0.05, want_prob=False)
        assert torch.equal(again, off), 'the fused decoder is not deterministic'
    finally:
        ops.set_gemm_mode(old)


def test_fused_decoder_agrees_with_per_layer_path_and_accumulates_depth(mods):
    ops = mods['ops']
    dec = _decoder(mods, 352, 2).to(DEV)
    g = torch.Generator().manual_seed(5)
    n_pts = 777
    x = torch.randn(n_pts, 8, 352, generator=g).to(DEV)
    x[:, 7] = 0
    off_f, prob_f = dec.run(x, 7, 0.025, want_prob=True)
    fused, dec._fused = dec._fused, None       # per-layer gather-GEMMs + decoder_head_kernel
    try:
        layers, head = dec._weights()
        dec._fused = None
        off_l, prob_l = dec.run(x, 7, 0.025, want_prob=True)
    finally:
        dec._fused = fused
    np.testing.assert_allclose(prob_f.cpu().numpy(), prob_l.cpu().numpy(), rtol=0, atol=2e-5)
    np.testing.assert_allclose(off_f.cpu().numpy(), off_l.cpu().numpy(), rtol=0, atol=2e-6)
    # depth += offset inside the kernel == torch's fp32 add of the returned offset
    layers, head = dec._weights()
    depth = torch.rand(n_pts, generator=g).to(DEV) + 1.0
    want = depth + off_f
    ops.decoder_fused(x, 352, dec._fused, [l[1] for l in layers], [l[2] for l in layers], head[0], head[1], 0.025,
                      depth_accum=depth)
    assert torch.equal(depth, want)
