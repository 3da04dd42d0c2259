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


def model_scene_sharded(net, depth_local, images_batch, img_feats, rotmats, tvecs, K, ref_src_edges, group=None):
    """`PL3DVNet.model_scene` (lightningmodel.py:176-185) for one batch of scenes whose reference
    views are sharded over the ranks of `group`: `depth_local` [n_local,h,w] holds this rank's
    contiguous range of the references. Returns the same `xs` on every rank."""
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
    return net.scene_from_points(pts_all, feat_all, pts_batch)
