"""GPU parity of the tcgen05 gather-GEMM (csrc/gemm_tc.cu) through the C ABI: every slice
kind it serves (plain / gathered Linear, 27-slice sparse convolution with GroupNorm +
residual, row-shifted Conv1d, concat + 1x1) against a float64 torch restatement of the same
contraction, and against the fp32 CUDA-core kernel. 3xTF32 must be fp32-grade (the tolerance
is that of an fp32 GEMM); plain TF32 is checked at its own (10-bit mantissa) tolerance."""
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
    o.set_gemm_mode('tf32x3')


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def group_norm_rows(y, w, b):
    n, c = y.shape
    g = y.view(n, c // 16, 16)
    mean = g.mean(2, keepdim=True)
    var = ((g - mean) ** 2).mean(2, keepdim=True)
    return ((g - mean) / torch.sqrt(var + 1e-5)).view(n, c) * w + b


TOL = {'tf32x3': 2e-5, 'tf32': 4e-3, 'f32': 2e-5}


def check(got, ref, mode):
    ref = ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item() / scale
    assert err < TOL[mode], 'mode %s: max err / max|ref| = %.3e' % (mode, err)


@pytest.mark.parametrize('mode', ['tf32x3', 'tf32', 'f32'])
@pytest.mark.parametrize('M,K,N', [(1000, 64, 128), (128, 128, 64), (25088, 128, 128), (77, 32, 64), (4097, 352, 128)])
def test_linear(ops, mode, M, K, N):
    ops.set_gemm_mode(mode)
    x, w, b = rnd(M, K, seed=1), rnd(K, N, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    y = ops.linear(x, w, b, relu_input=True, packed=ops.pack_weights(w))
    check(y, torch.relu(x.double()) @ w.double() + b.double(), mode)


@pytest.mark.parametrize('mode', ['tf32x3', 'f32'])
def test_linear_gathered_second_slice(ops, mode):
    ops.set_gemm_mode(mode)
    M, n_seg = 5000, 613
    x, pool = rnd(M, 128, seed=4), rnd(n_seg, 128, seed=5)
    seg = torch.randint(0, n_seg, (M,), generator=torch.Generator().manual_seed(6)).int().to(DEV)
    w, b = rnd(256, 128, seed=7, scale=1 / 16.0), rnd(128, seed=8)
    y = ops.linear(x, w, b, relu_input=False, pool=pool, seg=seg, packed=ops.pack_weights(w))
    ref = torch.cat((x, pool[seg.long()]), 1).double() @ w.double() + b.double()
    check(y, ref, mode)


def sparse_ref(feat, nbr, W, gw, gb, residual, relu):
    n_out = nbr.shape[0]
    acc = torch.zeros(n_out, W.shape[2], dtype=torch.float64, device=feat.device)
    for k in range(27):
        idx = nbr[:, k].long()
        live = idx >= 0
        acc[live] += feat[idx[live]].double() @ W[k].double()
    y = acc
    if gw is not None:
        y = group_norm_rows(y, gw.double(), gb.double())
    if residual is not None:
        y = y + residual.double()
    return torch.relu(y) if relu else y


@pytest.mark.parametrize('mode', ['tf32x3', 'tf32', 'f32'])
@pytest.mark.parametrize('n_in,n_out,Cin,Cout,density', [(3134, 3134, 128, 128, 0.3), (900, 3000, 128, 64, 0.2),
                                                         (3000, 700, 64, 128, 0.5), (200, 200, 64, 64, 0.0),
                                                         (50000, 50000, 128, 128, 0.33)])
def test_sparse_conv(ops, mode, n_in, n_out, Cin, Cout, density):
    ops.set_gemm_mode(mode)
    g = torch.Generator().manual_seed(n_in + n_out)
    nbr = torch.randint(0, n_in, (n_out, 27), generator=g)
    nbr[torch.rand(n_out, 27, generator=g) >= density] = -1
    if n_out >= 256:
        nbr[128:256, :13] = -1  # a tile with dead slices
    nbr = nbr.int().to(DEV)
    feat = rnd(n_in, Cin, seed=11)
    W = rnd(27, Cin, Cout, seed=12, scale=(9 * Cin) ** -0.5)
    gw, gb = rnd(Cout, seed=13) * 0.3 + 1.0, rnd(Cout, seed=14) * 0.1
    residual = rnd(n_out, Cout, seed=15) if Cin == Cout and n_in == n_out else None
    packed = ops.pack_weights(W.reshape(-1, Cout).contiguous())
    y = ops.sparse_conv(feat, nbr, W, gw, gb, residual, True, packed=packed)
    ref = sparse_ref(feat, nbr, W, gw, gb, residual, True)
    check(y, ref, mode)
    # K-split over CTAs (small levels): same result, bit-identical from launch to launch
    ws = ops.sparse_conv_workspace(Cout, DEV)
    if ws is not None:
        y_s = ops.sparse_conv(feat, nbr, W, gw, gb, residual, True, packed=packed, workspace=ws)
        check(y_s, ref, mode)
        y_s2 = ops.sparse_conv(feat, nbr, W, gw, gb, residual, True, packed=packed, workspace=ws)
        assert torch.equal(y_s, y_s2)
        assert int(ws[:4096].max()) == 0  # counters left zero
    # no normalisation, no residual, no ReLU: the raw contraction
    y2 = ops.sparse_conv(feat, nbr, W, None, None, None, False, packed=packed)
    check(y2, sparse_ref(feat, nbr, W, None, None, None, False), mode)


@pytest.mark.parametrize('mode', ['tf32x3', 'f32'])
@pytest.mark.parametrize('n_pts,Cin', [(3136, 352), (100, 128), (3136 * 4, 128), (2400, 128), (16 * 150, 352)])
def test_conv1d_rows(ops, mode, n_pts, Cin):
    ops.set_gemm_mode(mode)
    x = rnd(n_pts, 8, Cin, seed=21)
    x[:, 7] = 0
    w = rnd(3, Cin, 128, seed=22, scale=(3 * Cin) ** -0.5)
    scale, shift = rnd(128, seed=23) * 0.2 + 1.0, rnd(128, seed=24) * 0.1
    y = ops.conv1d_bn_relu(x, w, scale, shift, packed=ops.pack_weights(w.reshape(-1, 128).contiguous()))
    xr = x[:, :7].double().permute(0, 2, 1)  # [n, Cin, 7]
    wt = w.double().permute(2, 1, 0)          # [Cout, Cin, 3]
    ref = torch.nn.functional.conv1d(xr, wt, padding=1) * scale.double()[None, :, None] + shift.double()[None, :, None]
    ref = torch.relu(ref).permute(0, 2, 1)
    check(y[:, :7], ref, mode)
    assert (y[:, 7] == 0).all()
    # thin last round of tiles launched separately with its taps split over CTAs (C2: 196 = 148 + 48 tiles)
    ws = ops.sparse_conv_workspace(128, DEV)
    if ws is not None:
        y_s = ops.conv1d_bn_relu(x, w, scale, shift, packed=ops.pack_weights(w.reshape(-1, 128).contiguous()), workspace=ws)
        check(y_s[:, :7], ref, mode)
        assert (y_s[:, 7] == 0).all()
        y_s2 = ops.conv1d_bn_relu(x, w, scale, shift, packed=ops.pack_weights(w.reshape(-1, 128).contiguous()), workspace=ws)
        assert torch.equal(y_s, y_s2)
        assert int(ws[:4096].max()) == 0  # counters left zero


@pytest.mark.parametrize('mode', ['tf32x3', 'f32'])
def test_concat_linear(ops, mode):
    ops.set_gemm_mode(mode)
    n = 2999
    a, b = rnd(n, 128, seed=31), rnd(n, 128, seed=32)
    W = rnd(256, 128, seed=33, scale=1 / 16.0)
    gw, gb = rnd(128, seed=34) * 0.3 + 1.0, rnd(128, seed=35) * 0.1
    y = ops.concat_linear_gn_relu(a, b, W, gw, gb, packed=ops.pack_weights(W))
    ref = torch.relu(group_norm_rows(torch.cat((a, b), 1).double() @ W.double(), gw.double(), gb.double()))
    check(y, ref, mode)


def test_tc_matches_f32_kernel_bitwise_shape_and_close(ops):
    """same inputs through both kernels: agreement at fp32 rounding level"""
    M = 10000
    x, w, b = rnd(M, 128, seed=41), rnd(128, 128, seed=42, scale=1 / 11.0), rnd(128, seed=43)
    ops.set_gemm_mode('tf32x3')
    y_tc = ops.linear(x, w, b, relu_input=False, packed=ops.pack_weights(w))
    ops.set_gemm_mode('f32')
    y_f32 = ops.linear(x, w, b, relu_input=False, packed=None)
    assert (y_tc - y_f32).abs().max().item() < 2e-5 * y_f32.abs().max().item()


@pytest.mark.parametrize('mode', ['tf32x3', 'tf32'])
@pytest.mark.parametrize('n_in,n_out,Cin,Cout,density', [(3134, 3134, 64, 64, 0.045), (3045, 3045, 128, 128, 0.10),
                                                         (2408, 2408, 128, 128, 0.38), (900, 3000, 128, 64, 0.2),
                                                         (3000, 700, 64, 128, 0.3), (100, 100, 64, 64, 0.0),
                                                         (40000, 40000, 128, 128, 0.05)])
def test_sparse_conv_pair_major(ops, mode, n_in, n_out, Cin, Cout, density):
    """pair-major sparse convolution (csrc/sparse_pairs.cu: plan -> gather-GEMM on existing pairs
    -> ordered reduce + epilogue) against the float64 restatement and the output-stationary kernel"""
    ops.set_gemm_mode(mode)
    g = torch.Generator().manual_seed(n_in + 7 * n_out)
    nbr = torch.randint(0, n_in, (n_out, 27), generator=g)
    nbr[torch.rand(n_out, 27, generator=g) >= density] = -1
    if n_out >= 512:
        nbr[300:420] = -1          # rows with no neighbour at all
        nbr[:, 5] = -1             # an offset with no pair at all
    nbr = nbr.int().to(DEV)
    feat = rnd(n_in, Cin, seed=31)
    W = rnd(27, Cin, Cout, seed=32, scale=(9 * Cin) ** -0.5)
    gw, gb = rnd(Cout, seed=33) * 0.3 + 1.0, rnd(Cout, seed=34) * 0.1
    residual = rnd(n_out, Cout, seed=35) if Cin == Cout and n_in == n_out else None
    packed = ops.pack_weights(W.reshape(-1, Cout).contiguous())
    km = ops.KernelMap(nbr).build_plan()
    ops.finish_plans([km])
    live = (nbr >= 0)
    assert km.n_pairs == int(live.sum())
    assert km.n_tiles == int(((live.sum(0) + 127) // 128).sum())
    km.use_pairs = True   # force the variant under test whatever the library would choose
    y = ops.sparse_conv(feat, km, W, gw, gb, residual, True, packed=packed)
    ref = sparse_ref(feat, nbr, W, gw, gb, residual, True)
    check(y, ref, mode)
    y_again = ops.sparse_conv(feat, km, W, gw, gb, residual, True, packed=packed)
    assert torch.equal(y, y_again)     # fixed summation order: bit-reproducible
    y_dense = ops.sparse_conv(feat, nbr, W, gw, gb, residual, True, packed=packed)
    check(y, y_dense.double(), mode)
    # raw contraction, no epilogue
    y2 = ops.sparse_conv(feat, km, W, None, None, None, False, packed=packed)
    check(y2, sparse_ref(feat, nbr, W, None, None, None, False), mode)


def test_pair_major_choice_is_a_pure_function_of_counts(ops):
    L = importlib.import_module('3dvnet_b200._lib').lib()
    f = L.raw('dv3d_sparse_conv_prefers_pairs')
    assert f(3134, 51) == 1 and f(2408, 203) == 1     # C2 levels 0 and 2 (SURVEY sizes)
    assert f(9765, 1459) == 0                          # dense level of an 8-view scene
    assert f(100, 0) == 0
