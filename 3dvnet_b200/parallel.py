"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

The reference is single-GPU (`mv3d/config.py:3-5`). Two ways of spreading the path exist here:

* scenes / reference views are independent for the cost volume, the point-level warp and the
  PointFlow decoder: every rank takes a contiguous range of reference views and nothing is
  exchanged (`shard_range`, `local_edges`);
* one scene spanning GPUs: `utils.voxelize` needs the bounding box and the unique voxel set of
  ALL points (`mv3d/utils.py:39-48`) and the PointNet max-pools over all points of a voxel
  (`mv3d/subnetworks/scenemodeling.py:129`), so each rank back-projects its own reference views
  and ONE all-gather of the `[N_g, 3 + C]` fp32 point rows follows (`all_gather_rows`; the row
  counts follow from the shard plan, no size exchange). Ranks hold contiguous ranges, so the
  gathered array is bit-identical to the single-GPU point cloud and every rank builds the same
  voxel tables; the (small) scene model then runs redundantly on every rank.

Works with the `nccl` backend on CUDA tensors and with `gloo` on CPU tensors (the tests run
world_size 2 on CPU).
"""
import torch
import torch.distributed as dist


def shard_range(n, world, rank):
    """contiguous, balanced range [start, end) of `n` items for `rank` of `world`"""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_counts(n, world):
    return [shard_range(n, world, r)[1] - shard_range(n, world, r)[0] for r in range(world)]


def local_edges(ref_src_edges, start, end):
    """edges whose reference is one of the unique references [start, end) (in ascending image
    order, mvsnet.py:179), original relative order kept. Replaces utils.slice_edges
    (`mv3d/utils.py:32-35`) for the sharded drivers. -> ([2, E_local], ref_idx_local)"""
    ref_idx, gather = torch.unique(ref_src_edges[0], return_inverse=True)
    keep = (gather >= start) & (gather < end)
    return ref_src_edges[:, keep], ref_idx[start:end]


def all_gather_rows(x, counts, group=None):
    """x [counts[rank], C] on every rank -> [sum(counts), C] in rank order, with ONE collective
    (all_gather_into_tensor on rows padded to max(counts))."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(counts) != world or x.shape[0] != counts[rank]:
        raise RuntimeError('all_gather_rows: counts %s do not match world %d / local rows %d'
                           % (counts, world, x.shape[0]))
    m = max(counts)
    if m == 0:
        return x.new_empty((0,) + tuple(x.shape[1:]))
    if x.shape[0] == m:
        padded = x.contiguous()
    else:
        padded = x.new_zeros((m,) + tuple(x.shape[1:]))
        padded[:x.shape[0]] = x
    out = x.new_empty((world * m,) + tuple(x.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    if all(c == m for c in counts):
        return out
    return torch.cat([out[r * m:r * m + c] for r, c in enumerate(counts)], dim=0)


def model_scene_sharded(net, depth_local, images_batch, img_feats, rotmats, tvecs, K, ref_src_edges, group=None,
                        heap=None):
    """`PL3DVNet.model_scene` (lightningmodel.py:176-185) for one batch of scenes whose reference
    views are sharded over the ranks of `group`: `depth_local` [n_local,h,w] holds this rank's
    contiguous range of the references. Returns the same `xs` on every rank. With a `SymmHeap` the sparse
    U-Net is sharded by voxel rows as well (`sparse_unet_sharded`); without, every rank runs it in full."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ref_idx = torch.unique(ref_src_edges[0])
    n_ref = ref_idx.shape[0]
    start, end = shard_range(n_ref, world, rank)
    if depth_local.shape[0] != end - start:
        raise RuntimeError('model_scene_sharded: rank %d expects %d reference views, got %d'
                           % (rank, end - start, depth_local.shape[0]))
    edges_local, ref_local = local_edges(ref_src_edges, start, end)
    P = depth_local.shape[1] * depth_local.shape[2]
    dev = depth_local.device
    depth_batch_all = images_batch.to(dev)[ref_idx.to(dev)]
    if end > start:
        pts, feat, _ = net.construct_feature_rich_pointcloud(depth_local, depth_batch_all[start:end], img_feats,
                                                             rotmats, tvecs, K, edges_local)
        rows = torch.cat((pts, feat), dim=1)
    else:
        rows = torch.empty((0, 3 + img_feats.shape[1]), dtype=torch.float32, device=dev)
    counts = [c * P for c in shard_counts(n_ref, world)]
    rows = all_gather_rows(rows, counts, group)
    pts_all, feat_all = rows[:, :3].contiguous(), rows[:, 3:].contiguous()
    pts_batch = depth_batch_all.unsqueeze(1).expand(n_ref, P).reshape(-1).contiguous()
    if heap is not None:
        return scene_from_points_sharded(net, pts_all, feat_all, pts_batch, heap)
    return net.scene_from_points(pts_all, feat_all, pts_batch)


def hot_path_sharded(net, feats_quarter, rotmats, tvecs, K, ref_src_edges, images_batch, depth_config, offsets_list,
                     group=None, heap=None):
    """The whole hot path (plane sweep -> CostRegNet -> soft-argmin -> offsets_list x (scene model + PointFlow
    passes), eval-3dvnet.py:60-99) for scenes whose reference views are sharded over the ranks of `group`
    (BASELINE config C4). Every rank holds all feature maps and cameras (they are inputs, broadcast once by the
    caller); it computes the cost volumes, depths and PointFlow passes of ITS contiguous range of reference views,
    and the only exchange is the one all-gather of point rows inside each `model_scene_sharded`.
    Returns this rank's depth maps [n_local, h, w] and its range (start, end) of the sorted reference views."""
    from .mv3d.lightningmodel import Namespace
    from . import ops
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ref_idx = torch.unique(ref_src_edges[0])
    start, end = shard_range(ref_idx.shape[0], world, rank)
    edges_local, ref_local = local_edges(ref_src_edges, start, end)
    dev = feats_quarter.device
    h, w = depth_config['size']
    if end == start:
        depth = torch.empty((0, h, w), dtype=torch.float32, device=dev)
    else:
        plan = ops.edge_plan(edges_local, dev)
        batch = Namespace(rotmats=rotmats, tvecs=tvecs, K=K, ref_src_edges=plan)
        depth = net.mvsnet.depth_from_features(feats_quarter, batch, depth_config['depth_start'],
                                               depth_config['depth_interval'], depth_config['n_intervals'],
                                               depth_config['size']).clone()
    depth_batch = images_batch.to(dev)[ref_local.to(dev)]
    for offsets in offsets_list:
        xs = model_scene_sharded(net, depth, images_batch, feats_quarter, rotmats, tvecs, K, ref_src_edges, group, heap)
        if end > start:
            for offset in offsets:
                depth += net.run_pointflow(xs, depth, depth_batch, feats_quarter, rotmats, tvecs, K, edges_local,
                                           offset, 3)
    return depth, (start, end)


class ShardPlan(object):
    """What does not change between steps of `hot_path_sharded_native` for one edge list: this rank's range of the
    sorted reference views, the CSR edge plan of that range (resident on the device) and the view -> scene map."""

    def __init__(self, ref_src_edges, images_batch, device, world, rank):
        from . import ops
        ref_idx = torch.unique(ref_src_edges[0])
        self.n_ref = int(ref_idx.shape[0])
        self.start, self.end = shard_range(self.n_ref, world, rank)
        edges_local, _ = local_edges(ref_src_edges, self.start, self.end)
        self.plan = ops.edge_plan(edges_local, device) if self.end > self.start else None
        self.depth_batch_all = images_batch.to(device)[ref_idx.to(device)].long().contiguous()


def hot_path_sharded_native(net, feats_quarter, rotmats, tvecs, K, ref_src_edges, images_batch, depth_config,
                            offsets_list, heap, plan=None, return_init=False):
    """`hot_path_sharded` from ONE native call per rank (csrc/engine.cu dv3d_hot_path_sharded): no Python between
    the kernels, the point rows travel as peer copies into the symmetric heap instead of an NCCL all-gather, and
    the sparse U-Net runs row-sharded with epilogue stores into peer memory. `heap` is this group's `SymmHeap`
    (at least `native_heap_bytes` large); `plan` a `ShardPlan` to reuse across steps.
    Returns this rank's depth maps [n_local, h, w] and its range (start, end) of the sorted reference views."""
    from . import ops
    from .mv3d._pack import require_eval
    require_eval(net)
    _require_peer_store_gemm(ops)
    dev = feats_quarter.device
    if plan is None:
        plan = ShardPlan(ref_src_edges, images_batch, dev, heap.world, heap.rank)
    fq = feats_quarter.detach().float()
    if fq.dim() == 4 and fq.is_contiguous(memory_format=torch.channels_last) and not fq.is_contiguous():
        nhwc = fq.permute(0, 2, 3, 1)
    else:
        nhwc = ops.nchw_to_nhwc(fq.contiguous())
    out = ops.hot_path_engine_sharded(net.engine_params(), nhwc, rotmats.float().contiguous(), tvecs.float().contiguous(),
                                      K.float().contiguous(), plan.plan, plan.start, plan.depth_batch_all, depth_config,
                                      net.hparams.img_size, net.edge_len, offsets_list, heap, want_init=return_init)
    return out, (plan.start, plan.end)


def native_heap_bytes(net, n_ref_total, size):
    """size of the `SymmHeap` `hot_path_sharded_native` needs for `n_ref_total` reference views of `size` (h, w)"""
    import ctypes
    from . import ops
    return int(ops.lib().raw('dv3d_hot_path_sharded_heap_bytes')(ctypes.byref(net.engine_params()), int(n_ref_total),
                                                                  int(size[0]), int(size[1])))


# ----------------------------------------------------------------------------- sparse U-Net sharded by voxel rows
def _require_peer_store_gemm(ops):
    """Only the tcgen05 gather-GEMM and the pair-reduce kernel store their epilogue into the peers' copies of a
    symmetric buffer (csrc/gemm_tc.cu, sparse_pairs.cu -> symm_attach). The fp32 CUDA-core verification kernel
    (csrc/gemm.cu) writes locally only: with it every rank would read stale rows of the ranges other ranks own."""
    if ops.gemm_mode() == 'f32':
        raise RuntimeError("the row-sharded sparse U-Net needs a tensor-core GEMM mode ('tf32x3' or 'tf32'): the "
                           "'f32' verification kernel has no peer-store epilogue (DV3D_GEMM / ops.set_gemm_mode)")


class _RawCuda(object):
    """a raw device allocation seen by torch through the CUDA array interface"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}


class SymmHeap(object):
    """One buffer of `nbytes` per rank plus mappings of every peer's copy (CUDA IPC: dv3d_symm_alloc /
    dv3d_symm_open, handles exchanged with one all_gather_object), registered with the library so that
    sparse-convolution epilogues writing into the local copy also store into the peers' (csrc/symm.cu). A bump allocator hands out the same offsets on every rank as long as the
    ranks allocate the same sizes in the same order (they do: the coordinate levels are identical on all ranks).
    The first 256 bytes hold the barrier flags."""

    def __init__(self, nbytes, group=None):
        import ctypes
        from . import ops
        self.ops, self.group = ops, group
        _require_peer_store_gemm(ops)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 2 <= self.world <= 8:
            raise RuntimeError('SymmHeap: 2..8 ranks, got %d' % self.world)
        L = ops.lib()
        dev = torch.cuda.current_device()
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        L.call('dv3d_symm_alloc', self.nbytes, ctypes.byref(ptr), handle)
        self.ptr = ptr.value
        self.buf = torch.as_tensor(_RawCuda(self.ptr, self.nbytes), device=torch.device('cuda', dev))
        assert self.buf.data_ptr() == self.ptr
        self.err = torch.zeros(1, dtype=torch.int32, device=self.buf.device)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group)
        self.peers = []                                   # mapped peer blocks, ascending rank order, this rank left out
        for r, h in enumerate(handles):
            if r != self.rank:
                p = ctypes.c_void_p()
                L.call('dv3d_symm_open', h, ctypes.byref(p))
                self.peers.append(p.value)
        self._peer_ptrs = (ctypes.c_void_p * len(self.peers))(*self.peers)
        L.call('dv3d_symm_register', self.ptr, self.nbytes, self._peer_ptrs, len(self.peers))
        self.epoch = 0
        self.reset()
        dist.barrier(group)                                # every copy is zeroed, mapped and registered before any peer store

    def reset(self):
        self.off = 256
        self.free = {}

    def rows(self, n, C):
        """[n, C] fp32 at the same offset of every rank's buffer (a released block of that shape is reused)"""
        pool = self.free.get((n, C))
        if pool:
            return pool.pop()
        nbytes = (n * C * 4 + 255) // 256 * 256
        if self.off + nbytes > self.nbytes:
            raise RuntimeError('SymmHeap: %d bytes exhausted allocating [%d, %d] at offset %d'
                               % (self.nbytes, n, C, self.off))
        t = self.buf[self.off:self.off + n * C * 4].view(torch.float32).view(n, C)
        self.off += nbytes
        return t

    def release(self, t):
        """Hands a block back. Call it after the last layer reading `t` AND that layer's barrier were enqueued:
        the next layer - the first that can be given the block as its output, on any rank - starts behind that
        barrier, when no rank reads the old contents any more."""
        self.free.setdefault(tuple(t.shape), []).append(t)

    def barrier(self):
        """cross-GPU barrier on the current stream: the peers' stores of the layer before it are visible after it"""
        self.epoch += 1
        self.ops.lib().call('dv3d_symm_barrier', self.ptr, self._peer_ptrs, len(self.peers), self.rank,
                            self.epoch, self.err.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def check(self):
        if int(self.err.item()) != 0:
            raise RuntimeError('SymmHeap: a peer did not reach a barrier within the timeout')

    def close(self):
        """collective: no rank may still store into a block another rank frees"""
        if self.ptr is None:
            return
        L = self.ops.lib()
        torch.cuda.synchronize()
        dist.barrier(self.group)
        L.call('dv3d_symm_unregister', self.ptr)
        for p in self.peers:
            L.call('dv3d_symm_close', p)
        dist.barrier(self.group)
        self.buf = None
        L.call('dv3d_symm_free', self.ptr)
        self.ptr = None


def row_range(n, world, rank):
    """rows [rank*m, min((rank+1)*m, n)) with m = ceil(n / world): equal blocks, the tail ranks may be short or empty"""
    m = (n + world - 1) // world
    return min(rank * m, n), min((rank + 1) * m, n)


class ShardedScene(object):
    """Coordinate levels and hash tables of the whole scene (identical on every rank) with kernel maps and
    pair-major plans for THIS rank's rows of every level: rows [rank*m, min((rank+1)*m, n)), m = ceil(n / world)."""

    def __init__(self, idx, batch, n_levels, dims, world, rank):
        import ctypes
        from . import ops
        dev = idx.device
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.n_batch = dims[3]
        self.levels = [ops.SparseLevel(ops.make_coords(idx.int().contiguous(), batch.long().contiguous()), 1, self.err)]
        for _ in range(1, n_levels):
            self.levels.append(ops.coarsen(self.levels[-1], dims[:3], dims[3], self.err))
        self.range = [row_range(lv.n, world, rank) for lv in self.levels]
        self.ws = ops.sparse_conv_workspace(128, dev)
        L = ops.lib()

        def kmap(out_l, in_l, step):
            r0, r1 = self.range[out_l]
            if r1 == r0:
                return None
            nbr = torch.empty((r1 - r0, 27), dtype=torch.int32, device=dev)
            L.call('dv3d_kernel_map', self.levels[out_l].coords[r0:r1].data_ptr(), r1 - r0, self.levels[in_l].table.data_ptr(),
                   self.levels[in_l].table_bytes, step, nbr.data_ptr(), torch.cuda.current_stream().cuda_stream)
            return ops.KernelMap(nbr)

        n = n_levels
        self.same = [kmap(l, l, self.levels[l].stride) for l in range(n)]
        self.down = [kmap(l + 1, l, self.levels[l].stride) for l in range(n - 1)]
        self.up = [kmap(l, l + 1, -self.levels[l].stride) for l in range(n - 1)]
        maps = [m for m in self.same + self.down + self.up if m is not None]
        ops.build_plans(maps)
        ops.finish_plans(maps)


def sparse_unet_sharded(unet, F, pts, idx, batch, res, scene, heap):
    """`SparseUNet.forward` (scenemodeling.py:191-237) with every layer computed for this rank's voxel rows only.
    A layer whose output is gathered through a kernel map by the next one writes into a symmetric buffer (the
    epilogue stores the rows into every rank's copy) and is followed by one cross-GPU barrier; row-local
    consumers (the transposed convolution feeding the 1x1 'feature adjust') stay local. Returns the same list of
    level dicts as the single-GPU module, with full feature tensors on every rank."""
    from . import ops
    _require_peer_store_gemm(ops)
    nl = unet.n_levels

    def conv(x_full, km, mod, gn, level, residual=None):
        n_full, C = scene.levels[level].n, mod.kernel.shape[-1]
        y = heap.rows(n_full, C)
        r0, r1 = scene.range[level]
        if r1 > r0:
            wk, wp = mod.weights()
            ops.sparse_conv(x_full, km, wk, gn.weight.detach(), gn.bias.detach(),
                            None if residual is None else residual[r0:r1], True, packed=wp, workspace=scene.ws, out=y[r0:r1])
        heap.barrier()
        return y

    def res_chain(blocks, x, level, free_x):
        """residual blocks on one level; intermediate tensors go back to the heap, the last output stays"""
        for blk in blocks:
            h = conv(x, scene.same[level], blk.conv1, blk.n1.gn, level)
            y = conv(h, scene.same[level], blk.conv2, blk.n2.gn, level, residual=x)
            heap.release(h)
            if free_x:
                heap.release(x)
            x, free_x = y, True
        return x

    x = res_chain(unet.res_down[0], F.float().contiguous(), 0, False)   # F is not a heap block
    xs = [x]
    for i in range(1, nl):
        x = conv(x, scene.down[i - 1], unet.down[i - 1][0], unet.down[i - 1][1].gn, i)
        x = res_chain(unet.res_down[i], x, i, True)
        xs.append(x)
    out = [(xs[-1], nl - 1)]
    for i in range(nl - 1):
        l = nl - 2 - i
        r0, r1 = scene.range[l]
        adj, gn_adj = unet.feat_adj[i][0], unet.feat_adj[i][1].gn
        y = heap.rows(scene.levels[l].n, adj.kernel.shape[-1])
        if r1 > r0:
            mod, gn = unet.up[i][0], unet.up[i][1].gn
            wk, wp = mod.weights()
            up = ops.sparse_conv(x, scene.up[l], wk, gn.weight.detach(), gn.bias.detach(), None, True, packed=wp,
                                 workspace=scene.ws)                      # this rank's rows only, consumed row-locally
            wk, wp = adj.weights()
            ops.concat_linear_gn_relu(up, xs[l][r0:r1], wk, gn_adj.weight.detach(), gn_adj.bias.detach(), packed=wp,
                                      out=y[r0:r1])
        heap.barrier()
        x = res_chain(unet.res_up[i], y, l, True)
        out.append((x, l))

    origin = ops.batch_origin(pts.float().contiguous(), idx.int().contiguous(), batch.long().contiguous(),
                              scene.n_batch, res)
    info = []
    for feats, l in out:
        lv = scene.levels[l]
        x_pts, x_idx, x_batch = ops.level_points(lv, origin, res)
        # out of the heap: the next scene-model call reuses it
        info.append({'feats': feats.clone(), 'pts': x_pts, 'res': lv.stride * res, 'batch': x_batch, 'idx': x_idx,
                     'stride': lv.stride, 'sparse': lv, 'origin': origin})
    return info


def scene_from_points_sharded(net, pts, pts_feat, pts_batch, heap):
    """`PL3DVNet.scene_from_points` with the sparse U-Net row-sharded over the ranks of the heap's group
    (voxelisation and PointNet stay redundant: 15 % of the scene model at 64 reference views)."""
    from . import ops
    from .mv3d._pack import require_eval
    require_eval(net)
    a_pts, a_idx, a_batch, seg, grid = ops.voxelize(pts, pts_batch, net.edge_len)
    x = ops.pointnet_input(pts, pts_feat, a_pts, seg, net.pointnet.in_pad)
    x = net.pointnet.forward_padded(x, seg, a_pts.shape[0])
    dims = (int(grid.n_cells[0]), int(grid.n_cells[1]), int(grid.n_cells[2]), int(grid.n_batch))
    scene = ShardedScene(a_idx, a_batch, net.sparse_conv.n_levels, dims, heap.world, heap.rank)
    heap.reset()
    return sparse_unet_sharded(net.sparse_conv, x, a_pts, a_idx, a_batch, net.edge_len, scene, heap)
