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
