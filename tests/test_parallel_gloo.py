"""Host-side logic of the multi-GPU path (3dvnet_b200/parallel.py) with the gloo backend on
CPU, world_size 2 and 3: shard plan, edge slicing, and the single variable-length all-gather
whose result must equal the single-process point rows bit for bit."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_ref, P, C, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    par = importlib.import_module('3dvnet_b200.parallel')
    g = torch.Generator().manual_seed(0)
    rows_all = torch.randn(n_ref * P, C, generator=g)          # what one process would hold
    start, end = par.shard_range(n_ref, world, rank)
    mine = rows_all[start * P:end * P].clone()
    counts = [c * P for c in par.shard_counts(n_ref, world)]
    got = par.all_gather_rows(mine, counts)
    ok = torch.equal(got, rows_all)
    # mismatching plan must raise, not hang
    raised = False
    try:
        par.all_gather_rows(mine[:-1] if mine.shape[0] else mine.new_zeros(1, C), counts)
    except RuntimeError:
        raised = True
    np.save(os.path.join(out_dir, 'r%d.npy' % rank), np.array([ok, raised, start, end]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,n_ref', [(2, 8), (2, 7), (3, 4), (2, 1)])
def test_all_gather_rows_matches_single_process(world, n_ref, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_ref, 12, 35, str(tmp_path)), nprocs=world, join=True)
    covered = []
    for r in range(world):
        ok, raised, start, end = np.load(tmp_path / ('r%d.npy' % r))
        assert ok == 1 and raised == 1
        covered += list(range(int(start), int(end)))
    assert covered == list(range(n_ref))          # contiguous, disjoint, complete


def test_shard_plan_and_local_edges():
    par = importlib.import_module('3dvnet_b200.parallel')
    synth = importlib.import_module('3dvnet_b200.synth')
    for n, w in ((64, 8), (7, 2), (3, 4), (0, 2)):
        c = par.shard_counts(n, w)
        assert sum(c) == n and max(c) - min(c) <= 1
    e = torch.from_numpy(synth.make_edges(12, 2, 2, include_self=True))
    e = e[:, torch.randperm(e.shape[1], generator=torch.Generator().manual_seed(1))]
    ref_idx, gather = torch.unique(e[0], return_inverse=True)
    seen = 0
    for rank in range(3):
        s, t = par.shard_range(len(ref_idx), 3, rank)
        le, lref = par.local_edges(e, s, t)
        assert lref.tolist() == ref_idx[s:t].tolist()
        assert set(le[0].tolist()) <= set(lref.tolist())
        # original relative order of the kept edges
        keep = (gather >= s) & (gather < t)
        assert torch.equal(le, e[:, keep])
        seen += le.shape[1]
    assert seen == e.shape[1]


def test_native_shard_plans_tile_the_reference_views():
    """ShardPlan (the per-rank input of dv3d_hot_path_sharded): the ranks' ranges tile the sorted reference views,
    each rank's CSR plan holds exactly its edges, more ranks than views leaves the tail ranks empty"""
    par = importlib.import_module('3dvnet_b200.parallel')
    synth = importlib.import_module('3dvnet_b200.synth')
    e = torch.from_numpy(synth.make_edges(7, 2, 2, include_self=True))
    ib = torch.zeros(int(e.max()) + 1, dtype=torch.long)
    n_ref = len(torch.unique(e[0]))
    for world in (2, 3, 8, n_ref + 2):
        at, edges = 0, 0
        for rank in range(world):
            sp = par.ShardPlan(e, ib, 'cpu', world, rank)
            assert sp.n_ref == n_ref and sp.start == at and sp.depth_batch_all.shape[0] == n_ref
            at = sp.end
            if sp.end == sp.start:
                assert sp.plan is None
            else:
                assert sp.plan.n_ref == sp.end - sp.start
                edges += sp.plan.n_edges
        assert at == n_ref and edges == e.shape[1]


def test_row_ranges_partition_every_level():
    par = importlib.import_module('3dvnet_b200.parallel')
    for n in (0, 1, 7, 8, 9, 1000, 200_704):
        for world in (2, 3, 8):
            got = [par.row_range(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))           # contiguous, in rank order
            assert max(b - a for a, b in got) == (n + world - 1) // world


def test_symm_heap_blocks_are_recycled_at_identical_offsets():
    """the bump allocator of the symmetric heap must hand out the same offsets on every rank for the same
    sequence of rows()/release() calls (pure host logic; the device side is covered on 2 GPUs)"""
    par = importlib.import_module('3dvnet_b200.parallel')

    def run():
        h = par.SymmHeap.__new__(par.SymmHeap)
        h.nbytes = 1 << 20
        h.buf = torch.zeros(h.nbytes, dtype=torch.uint8)
        h.reset()
        base = h.buf.data_ptr()
        a, b = h.rows(100, 64), h.rows(100, 64)
        c = h.rows(50, 128)
        h.release(a)
        d = h.rows(100, 64)                       # reuses a
        e = h.rows(100, 64)                       # fresh
        offs = [t.data_ptr() - base for t in (a, b, c, d, e)]
        with pytest.raises(RuntimeError):
            h.rows(1 << 20, 64)
        return offs

    o1, o2 = run(), run()
    assert o1 == o2
    assert o1[0] == 256 and o1[3] == o1[0] and len(set(o1[:3] + o1[4:])) == 4
    assert all(o % 256 == 0 for o in o1)


def test_sharded_unet_schedule_is_symmetric_and_never_reuses_a_block_too_early(monkeypatch):
    """Host-side model of `sparse_unet_sharded`: with the CUDA ops replaced by recorders, every rank (also one
    whose row range is empty) must issue the SAME sequence of heap allocations / releases / barriers - that is what
    keeps the offsets symmetric - and a recycled block may only be written after a barrier that follows the last
    read of its previous contents (the peers write into it from their side)."""
    par = importlib.import_module('3dvnet_b200.parallel')
    ops = importlib.import_module('3dvnet_b200.ops')
    sm = importlib.import_module('3dvnet_b200.mv3d.subnetworks.scenemodeling')
    monkeypatch.setattr(sm.SparseConvolution, 'weights', lambda self: (self.kernel.detach(), None))
    unet = sm.SparseUNet().eval()
    n_rows = [9, 5, 1]                                          # level sizes; the coarsest has ONE row

    class Level(object):
        def __init__(self, n, stride):
            self.n, self.stride = n, stride

    class Heap(object):
        def __init__(self, log):
            self.log, self.free, self.next_id, self.ids = log, {}, 0, {}

        def rows(self, n, C):
            pool = self.free.get((n, C))
            if pool:
                t = pool.pop()
                self.log.append(('reuse', self.ids[id(t)]))
                return t
            t = torch.zeros(n, C)
            self.ids[id(t)] = self.next_id
            self.keep = getattr(self, 'keep', []) + [t]
            self.log.append(('alloc', self.next_id, n, C))
            self.next_id += 1
            return t

        def release(self, t):
            self.log.append(('release', self.ids[id(t)]))
            self.free.setdefault(tuple(t.shape), []).append(t)

        def barrier(self):
            self.log.append(('barrier',))

    def run(world, rank):
        log = []
        heap = Heap(log)

        def base_id(t):                      # a row slice shares storage with its heap block
            for blk in heap.keep:
                if t.untyped_storage().data_ptr() == blk.untyped_storage().data_ptr():
                    return heap.ids[id(blk)]
            return None                      # not a heap block (PointNet features, local `up`)

        def sparse_conv(feat, km, W, gw, gb, residual, relu, packed=None, workspace=None, out=None):
            reads = [base_id(feat)] + ([base_id(residual)] if residual is not None else [])
            if out is None:
                out = torch.zeros(km, W.shape[-1])
            log.append(('conv', tuple(r for r in reads if r is not None), base_id(out)))
            return out

        def concat_linear(a, b, W, gw, gb, packed=None, out=None):
            log.append(('conv', tuple(r for r in (base_id(a), base_id(b)) if r is not None), base_id(out)))
            return out

        monkeypatch.setattr(ops, 'sparse_conv', sparse_conv)
        monkeypatch.setattr(ops, 'concat_linear_gn_relu', concat_linear)
        monkeypatch.setattr(ops, 'batch_origin', lambda *a: torch.zeros(1, 3))
        monkeypatch.setattr(ops, 'level_points', lambda lv, o, r: (torch.zeros(lv.n, 3),) * 3)
        scene = type('S', (), {})()
        scene.levels = [Level(n, 1 << l) for l, n in enumerate(n_rows)]
        scene.range = [par.row_range(n, world, rank) for n in n_rows]
        scene.n_batch, scene.ws = 1, None
        rows = lambda l: scene.range[l][1] - scene.range[l][0]
        scene.same = [rows(l) or None for l in range(3)]        # the recorder only needs the local row count
        scene.down = [rows(l + 1) or None for l in range(2)]
        scene.up = [rows(l) or None for l in range(2)]
        out = par.sparse_unet_sharded(unet, torch.zeros(n_rows[0], 64), torch.zeros(n_rows[0], 3),
                                      torch.zeros(n_rows[0], 3, dtype=torch.int32),
                                      torch.zeros(n_rows[0], dtype=torch.int64), 0.04, scene, heap)
        assert [o['feats'].shape[0] for o in out] == n_rows[::-1]
        return log

    logs = {(w, r): run(w, r) for w, r in ((2, 0), (2, 1), (4, 0), (4, 3))}
    assert par.row_range(1, 4, 3) == (1, 1)                      # rank 3 of 4 owns no row of the coarsest level
    heap_events = lambda log: [e for e in log if e[0] != 'conv']
    for key in ((2, 1), (4, 0), (4, 3)):
        assert heap_events(logs[key]) == heap_events(logs[(2, 0)]), key
    assert sum(e[0] == 'barrier' for e in logs[(2, 0)]) == 22     # 18 residual convs + 2 strided + 2 feature-adjust
    assert any(e[0] == 'reuse' for e in logs[(2, 0)])
    # recycling discipline on the rank that computes every layer
    log = logs[(2, 0)]
    last_read, barrier_since_read, awaiting_write = {}, {}, set()
    for ev in log:
        if ev[0] == 'barrier':
            for b in barrier_since_read:
                barrier_since_read[b] = True
        elif ev[0] == 'conv':
            for b in ev[1]:
                assert b not in awaiting_write, 'block %d read between its reuse and its first write' % b
                barrier_since_read[b] = False
            if ev[2] is not None and ev[2] in awaiting_write:
                assert barrier_since_read.get(ev[2], True), 'block %d overwritten before a barrier' % ev[2]
                awaiting_write.discard(ev[2])
        elif ev[0] == 'reuse':
            awaiting_write.add(ev[1])
