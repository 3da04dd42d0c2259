"""The persistent variant of the tcgen05 gather-GEMM (single-slice launches: Linear, pair-major sparse convolution,
PropagationNet rows; csrc/gemm_tc.cu: gather_gemm_tc_persistent_kernel) must give BIT-IDENTICAL results to the
one-tile-per-CTA kernel - same arithmetic, different scheduling - for one tile per CTA, many tiles per CTA, ragged last
tiles, both tile widths and both precisions."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def ops():
    importlib.import_module('3dvnet_b200.build').build()
    o = importlib.import_module('3dvnet_b200.ops')
    yield o
    o.lib().call('dv3d_set_gemm_persistent', 0)


def both(ops, fn):
    out = []
    for mode in (-1, 1):
        ops.lib().call('dv3d_set_gemm_persistent', mode)
        out.append(fn().clone())
    ops.lib().call('dv3d_set_gemm_persistent', 0)
    return out


@pytest.mark.parametrize('M,K,N,relu_in,mode', [(3136, 128, 128, True, 'tf32x3'), (50001, 128, 128, False, 'tf32x3'),
                                               (40000, 256, 128, True, 'tf32x3'), (33000, 64, 64, True, 'tf32x3'),
                                               (1, 128, 64, False, 'tf32x3'), (25000, 128, 128, True, 'tf32')])
def test_linear_identical(ops, M, K, N, relu_in, mode):
    old = ops.gemm_mode()
    ops.set_gemm_mode(mode)
    try:
        g = torch.Generator().manual_seed(M)
        x, w, b = (torch.randn(M, K, generator=g).to(DEV), torch.randn(K, N, generator=g).to(DEV),
                   torch.randn(N, generator=g).to(DEV))
        packed = ops.pack_weights(w)
        a, p = both(ops, lambda: ops.linear(x, w, b, relu_in, packed=packed))
        assert torch.equal(a, p)
        ref = (torch.relu(x) if relu_in else x).double() @ w.double() + b.double()
        tol = 2e-5 if mode == 'tf32x3' else 5e-3
        assert (p.double() - ref).abs().max().item() <= tol * ref.abs().max().item()
    finally:
        ops.set_gemm_mode(old)


def test_linear_with_pooled_second_input_is_not_eligible_but_still_right(ops):
    """two slices ([x | pool[seg]]) take the regular kernel in every mode"""
    g = torch.Generator().manual_seed(3)
    M = 20000
    x, pool = torch.randn(M, 128, generator=g).to(DEV), torch.randn(777, 128, generator=g).to(DEV)
    seg = torch.randint(0, 777, (M,), generator=g).int().to(DEV)
    w, b = torch.randn(256, 128, generator=g).to(DEV), torch.randn(128, generator=g).to(DEV)
    packed = ops.pack_weights(w)
    a, p = both(ops, lambda: ops.linear(x, w, b, True, pool=pool, seg=seg, packed=packed))
    assert torch.equal(a, p)


@pytest.mark.parametrize('n,C,density', [(3000, 64, 0.05), (30000, 128, 0.3), (20000, 64, 0.2)])
def test_pair_major_sparse_conv_identical(ops, n, C, density):
    g = torch.Generator().manual_seed(n)
    nbr = torch.randint(0, n, (n, 27), generator=g)
    nbr[torch.rand(n, 27, generator=g) >= density] = -1
    km = ops.KernelMap(nbr.int().to(DEV)).build_plan()
    ops.finish_plans([km])
    assert km.use_pairs
    feat, W = torch.randn(n, C, generator=g).to(DEV), (torch.randn(27, C, C, generator=g) * 0.1).to(DEV)
    pw = ops.pack_weights(W.reshape(-1, C).contiguous())
    gw, gb = torch.rand(C, generator=g).to(DEV) + 0.5, torch.randn(C, generator=g).to(DEV)
    a, p = both(ops, lambda: ops.sparse_conv(feat, km, W, gw, gb, feat, True, packed=pw))
    assert torch.equal(a, p)
    assert torch.isfinite(p).all()


def test_automatic_mode_uses_it_only_beyond_148_tiles(ops):
    """the default policy keeps single-wave launches on the regular kernel; both give the same bits anyway"""
    g = torch.Generator().manual_seed(9)
    w, b = torch.randn(128, 128, generator=g).to(DEV), torch.randn(128, generator=g).to(DEV)
    packed = ops.pack_weights(w)
    for M in (148 * 128, 148 * 128 + 1, 400 * 128 - 5):
        x = torch.randn(M, 128, generator=g).to(DEV)
        ops.lib().call('dv3d_set_gemm_persistent', 0)
        auto = ops.linear(x, w, b, True, packed=packed)
        ops.lib().call('dv3d_set_gemm_persistent', -1)
        never = ops.linear(x, w, b, True, packed=packed)
        ops.lib().call('dv3d_set_gemm_persistent', 0)
        assert torch.equal(auto, never)


@pytest.mark.parametrize('n,Cin,Cout,density,mode', [(9000, 128, 128, 0.4, 'tf32x3'), (30000, 64, 64, 0.1, 'tf32x3'),
                                                    (7000, 64, 128, 0.3, 'tf32x3'), (7000, 128, 64, 0.3, 'tf32x3'),
                                                    (300, 128, 128, 0.5, 'tf32x3'), (9000, 128, 128, 0.4, 'tf32')])
def test_weight_stationary_pair_gemm_identical(ops, n, Cin, Cout, density, mode):
    """the pair-major sparse convolution through pair_gemm_ws_kernel (contiguous tile ranges, resident weight block)
    and through the general gather-GEMM kernels: same arithmetic, bit-identical rows; all four (Cin, Cout) variants,
    many offset changes per CTA (n large) and one tile per CTA (n small)"""
    old = ops.gemm_mode()
    ops.set_gemm_mode(mode)
    try:
        g = torch.Generator().manual_seed(n + Cin)
        nbr = (torch.arange(n).view(-1, 1) + torch.randint(-500, 500, (n, 27), generator=g)).clamp_(0, n - 1)
        nbr[torch.rand(n, 27, generator=g) >= density] = -1
        km = ops.KernelMap(nbr.int().to(DEV)).build_plan()
        ops.finish_plans([km])
        km.use_pairs = True
        feat, W = torch.randn(n, Cin, generator=g).to(DEV), (torch.randn(27, Cin, Cout, generator=g) / (27 * Cin) ** 0.5).to(DEV)
        pw = ops.pack_weights(W.reshape(-1, Cout).contiguous())
        gw, gb = torch.randn(Cout, generator=g).to(DEV), torch.randn(Cout, generator=g).to(DEV)
        res = torch.randn(n, Cout, generator=g).to(DEV) if Cin == Cout else None
        out = []
        for ws in (1, 0):
            ops.lib().call('dv3d_set_pair_gemm_mode', ws)
            out.append(ops.sparse_conv(feat, km, W, gw, gb, res, True, packed=pw).clone())
        ops.lib().call('dv3d_set_pair_gemm_mode', 1)
        assert torch.equal(out[0], out[1])
        # against float64: conv -> GroupNorm(16 channels per group) -> + residual -> ReLU
        valid = (nbr >= 0).to(DEV)
        gathered = feat.double()[nbr.clamp(min=0).to(DEV)] * valid.unsqueeze(-1)                     # [n, 27, Cin]
        y = torch.einsum('nkc,kcd->nd', gathered, W.double())
        yg = y.view(n, Cout // 16, 16)
        yg = (yg - yg.mean(-1, keepdim=True)) / torch.sqrt(yg.var(-1, unbiased=False, keepdim=True) + 1e-5)
        ref = yg.reshape(n, Cout) * gw.double() + gb.double()
        if res is not None:
            ref = ref + res.double()
        ref = torch.relu(ref)
        tol = 1e-4 if mode == 'tf32x3' else 3e-2
        assert (out[0].double() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
    finally:
        ops.lib().call('dv3d_set_pair_gemm_mode', 1)
        ops.set_gemm_mode(old)
