"""Size-independent properties of the CUDA path at the FULL sizes of BASELINE.json's configs, where the
CPU oracle would take minutes: exact power-of-two scaling and permutation behaviour of the warp +
variance kernel at C5 (512x640, D=192, 10 src), sortedness / idempotence of the voxeliser on 2 M points,
pair-plan invariants of a large kernel map, determinism of the whole pass at C2."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def mods():
    importlib.import_module('3dvnet_b200.build').build()
    return dict(ops=importlib.import_module('3dvnet_b200.ops'), synth=importlib.import_module('3dvnet_b200.synth'),
                utils=importlib.import_module('3dvnet_b200.mv3d.utils'),
                lm=importlib.import_module('3dvnet_b200.mv3d.lightningmodel'))


def test_planesweep_c5_scaling_and_nonnegativity(mods):
    """var(2^k x) = 4^k var(x) bit for bit (every operation of the chain commutes with a power-of-two scale),
    var >= -eps, and the kernel is deterministic - at BASELINE configs[4] size (308 MB slab)"""
    ops, synth = mods['ops'], mods['synth']
    img, D, plane, n_src = (512, 640), 192, (112, 112), 10
    b = synth.make_batch(1, 1 + n_src, img, plane, 32, 5, 5, False, 0)
    plan = ops.edge_plan(b.ref_src_edges, torch.device(DEV))
    cams = ops.camera_tables(b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV))
    f = b.feats_quarter.to(DEV)
    v1 = ops.planesweep_var(ops.nchw_to_nhwc(f), cams, plan, 0.5, 0.05, D, plane, img)
    v1b = ops.planesweep_var(ops.nchw_to_nhwc(f), cams, plan, 0.5, 0.05, D, plane, img)
    v8 = ops.planesweep_var(ops.nchw_to_nhwc(f * 8.0), cams, plan, 0.5, 0.05, D, plane, img)
    assert v1.shape == (1, 32, D) + plane
    assert torch.equal(v1, v1b)
    assert torch.equal(v8, v1 * 64.0)
    assert float(v1.min()) > -1e-5 and torch.isfinite(v1).all()
    assert float(v1.max()) > 0.1   # the views really disagree somewhere


def test_voxelize_2m_points_sorted_and_complete(mods):
    rng = np.random.RandomState(5)
    n, edge = 2_000_000, 0.04
    # extents that are NOT exact multiples of the edge: on an exact multiple the reference decodes the ids with
    # ceil(extent/edge) cells although they were built with trunc(.)+1 (utils.py:41 vs grid_cluster) - that case is
    # pinned bit for bit by tests/golden/voxelize_adversarial.npz, it is not a property of a voxelisation
    pts = torch.from_numpy((rng.uniform(-4, 4, size=(n, 3)) * [0.99125, 0.79125, 0.33125]).astype(np.float32)).to(DEV)   # extents 198.25 / 158.25 / 66.25 edges
    batch = torch.from_numpy(np.sort(rng.randint(0, 3, size=n)).astype(np.int64)).to(DEV)
    a_pts, a_idx, a_batch, edges = mods['utils'].voxelize(pts, batch, edge)
    nv = a_pts.shape[0]
    assert 0 < nv <= n and edges.shape == (2, n)
    # every point lies in the cube of its anchor; every anchor is used; anchors of a point share its batch
    assert ((pts - a_pts[edges[0]]).abs() <= edge / 2 + 1e-4).all()
    assert torch.equal(torch.unique(edges[0]), torch.arange(nv, device=DEV))
    assert torch.equal(a_batch[edges[0]], batch)
    assert torch.equal(edges[1], torch.arange(n, device=DEV))
    # sortedness: ascending (batch, z, y, x) = ascending voxel id (utils.py:45-48), strictly (unique)
    lo = a_idx.min(0)[0]
    assert int(lo.min()) >= 0
    key = ((a_batch * 4096 + a_idx[:, 2].long()) * 4096 + a_idx[:, 1].long()) * 4096 + a_idx[:, 0].long()
    # idx3d is shifted by the per-batch minimum (utils.py:61-62): order inside a batch is what the id order implies
    for b in range(3):
        kb = key[a_batch == b]
        assert (kb[1:] > kb[:-1]).all()
    # completeness: as many anchors as distinct (batch, cell) triples, computed independently
    # (a tensor divisor: torch's CUDA division by a python scalar multiplies by the reciprocal)
    cell = ((pts - pts.min(0)[0]) / torch.full_like(pts, edge)).long()
    cid = ((batch * 4096 + cell[:, 2]) * 4096 + cell[:, 1]) * 4096 + cell[:, 0]
    assert torch.unique(cid).numel() == nv


def test_pair_plan_invariants_large(mods):
    ops = mods['ops']
    g = torch.Generator().manual_seed(9)
    n = 300_000
    nbr = torch.randint(0, n, (n, 27), generator=g)
    nbr[torch.rand(n, 27, generator=g) >= 0.2] = -1
    nbr = nbr.int().to(DEV)
    km = ops.KernelMap(nbr).build_plan()
    ops.finish_plans([km])
    live = nbr >= 0
    counts = live.sum(0)
    assert km.n_pairs == int(live.sum()) and km.n_tiles == int(((counts + 127) // 128).sum())
    plan = km.plan.view(torch.int32)
    hdr = plan[:64].cpu()
    assert hdr[:27].tolist() == counts.cpu().tolist()
    cap = (n * 27 + 127) // 128 + 27
    tile_k = plan[64:64 + km.n_tiles].cpu()
    assert (tile_k[1:] >= tile_k[:-1]).all()                          # tiles grouped by offset
    off = 64 + (cap + 63) // 64 * 64
    pair_in = plan[off:off + km.n_tiles * 128]
    pair_slot = plan[off + cap * 128:off + cap * 128 + n * 27].view(n, 27)
    assert int((pair_in >= 0).sum()) == km.n_pairs
    # slot(m, k) points at the pair's input row; slots of an offset ascend with the output row
    m_idx, k_idx = torch.nonzero(live, as_tuple=True)
    slots = pair_slot[m_idx, k_idx].long()
    assert torch.equal(pair_in[slots], nbr[m_idx, k_idx])
    assert int(pair_slot[~live].max()) == -1
    for k in (0, 13, 26):
        s = pair_slot[:, k][live[:, k]]
        assert (s[1:] > s[:-1]).all()


def test_hot_path_c2_is_deterministic(mods):
    bench = importlib.import_module('bench')
    b, params = bench.synth_inputs(2, 1)
    net = mods['lm'].PL3DVNet(bench.DEPTH_CFG, bench.DEPTH_CFG, bench.EDGE_LEN, feat_dim=32, img_size=bench.IMG_SIZE)
    net.load_state_dict(params, strict=False)
    net = net.to(DEV).eval()
    args = (b.feats_quarter.to(DEV), b.rotmats.to(DEV), b.tvecs.to(DEV), b.K.to(DEV), b.ref_src_edges,
            b.images_batch.to(DEV), bench.DEPTH_CFG, bench.OFFSETS_LIST)
    outs = [net.hot_path(*args).cpu() for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert torch.isfinite(outs[0]).all() and float(outs[0].min()) > 0.2 and float(outs[0].max()) < 6.0


def test_batched_levels_equal_single_level_calls(mods):
    """dv3d_coarsen_enqueue_batch / dv3d_hash_build_batch (what the engine uses) give exactly the levels,
    and tables answering exactly the lookups, of the one-level entry points (what the composed path uses)."""
    import ctypes
    ops = mods['ops']
    L = ops.lib()
    g = torch.Generator().manual_seed(21)
    dims, n_batch = (97, 83, 41), 3
    cells = torch.unique(torch.randint(0, dims[0] * dims[1] * dims[2] * n_batch, (60_000,), generator=g))
    b = cells // (dims[0] * dims[1] * dims[2])
    r = cells % (dims[0] * dims[1] * dims[2])
    coords = torch.stack([b, r % dims[0], (r // dims[0]) % dims[1], r // (dims[0] * dims[1])], 1).int().to(DEV)
    n = coords.shape[0]
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    fine = ops.SparseLevel(coords, 1, err)
    single = [ops.coarsen(fine, dims, n_batch, err)]
    single.append(ops.coarsen(single[0], dims, n_batch, err))

    strides = (ctypes.c_int * 2)(2, 4)
    wsb = [(L.raw('dv3d_coarsen_workspace_bytes')(dims[0], dims[1], dims[2], n_batch, s) + 255) // 256 * 256 for s in (2, 4)]
    span = torch.empty(sum(wsb), dtype=torch.uint8, device=DEV)
    ws = (ctypes.c_void_p * 2)(span.data_ptr(), span.data_ptr() + wsb[0])
    wsbytes = (ctypes.c_size_t * 2)(*wsb)
    outs = [torch.full((n, 4), -7, dtype=torch.int32, device=DEV) for _ in range(2)]
    outp = (ctypes.c_void_p * 2)(*[o.data_ptr() for o in outs])
    stream = torch.cuda.current_stream().cuda_stream
    L.call('dv3d_coarsen_enqueue_batch', coords.data_ptr(), n, strides, 2, dims[0], dims[1], dims[2], n_batch, ws, wsbytes,
           n, outp, stream)
    counts = []
    for l in range(2):
        n_out = ctypes.c_longlong(0)
        L.call('dv3d_coarsen_finish', ws[l], 2 << l, dims[0], dims[1], dims[2], n_batch, n, ctypes.byref(n_out), stream)
        counts.append(n_out.value)
    for l in range(2):
        assert counts[l] == single[l].n
        assert torch.equal(outs[l][:counts[l]], single[l].coords)

    # batched tables: the kernel maps built from them equal those of the single-table build
    tb = [L.raw('dv3d_hash_bytes')(c) for c in counts]
    tables = [torch.empty(t, dtype=torch.uint8, device=DEV) for t in tb]
    L.call('dv3d_hash_build_batch', (ctypes.c_void_p * 2)(*[outs[l].data_ptr() for l in range(2)]),
           (ctypes.c_longlong * 2)(*counts), (ctypes.c_void_p * 2)(*[t.data_ptr() for t in tables]),
           (ctypes.c_size_t * 2)(*tb), 2, err.data_ptr(), stream)
    for l in range(2):
        want = single[l].kernel_map(single[l], single[l].stride).nbr
        got = torch.empty_like(want)
        L.call('dv3d_kernel_map', single[l].coords.data_ptr(), counts[l], tables[l].data_ptr(), tb[l], single[l].stride,
               got.data_ptr(), stream)
        assert torch.equal(got, want)
    assert int(err.item()) == 0
    with pytest.raises(ops.Dv3dError):
        L.call('dv3d_coarsen_enqueue_batch', coords.data_ptr(), n, strides, 5, dims[0], dims[1], dims[2], n_batch, ws,
               wsbytes, n, outp, stream)
