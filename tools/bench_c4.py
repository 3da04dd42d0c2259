"""BASELINE config C4: ONE scene of 64 reference views (71 keyframes, 256x320, D=96, 7 src, 4 cm voxels) whose
reference views are sharded over the ranks; the only exchange is one NCCL all-gather of the [N_g, 35] fp32 point
rows per scene-model call (3dvnet_b200/parallel.py). STRONG scaling: the scene is fixed, ranks split it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_c4.py [--refs 64] [--steps 5] [--warmup 3]

Prints one JSON line on rank 0 (supplementary to bench.py's C2 line; same timing rules: warm-up, L2 flush, CUDA
events on the launching stream, barrier on both sides, max over ranks)."""
import argparse
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--refs', type=int, default=64)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--symm', type=int, default=1, help='1: sparse U-Net sharded by voxel rows over a symmetric heap (N>1)')
    ap.add_argument('--heap-mb', type=int, default=2048)
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29511')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    if local == 0:
        importlib.import_module('3dvnet_b200.build').build()
    dist.barrier()
    synth = importlib.import_module('3dvnet_b200.synth')
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    par = importlib.import_module('3dvnet_b200.parallel')
    ops = importlib.import_module('3dvnet_b200.ops')

    img, plane = (256, 320), (56, 56)
    cfg = dict(depth_start=0.5, depth_interval=0.05, n_intervals=96, size=plane)
    offsets = [[0.05, 0.05, 0.025]] * 2
    b = synth.make_batch(1, args.refs + 7, img, plane, 32, 4, 3, False, 0)   # identical on every rank (seeded)
    net = lm.PL3DVNet(cfg, cfg, 0.04, feat_dim=32, img_size=img)
    net.load_state_dict(synth.make_params(0), strict=False)
    net = net.to(dev).eval()
    fq, R, t, K = b.feats_quarter.to(dev), b.rotmats.to(dev), b.tvecs.to(dev), b.K.to(dev)
    e, ib = b.ref_src_edges, b.images_batch.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    heap = par.SymmHeap(args.heap_mb << 20) if (world > 1 and args.symm) else None

    def step():
        return par.hot_path_sharded(net, fq, R, t, K, e, ib, cfg, offsets, heap=heap)

    with torch.no_grad():
        for _ in range(args.warmup):
            depth, rng = step()
        torch.cuda.synchronize()
        dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        n0 = ops.launch_count()
        for a, z in ev:
            flush.zero_()
            a.record()
            depth, rng = step()
            z.record()
        torch.cuda.synchronize()
        dist.barrier()
    ms = torch.tensor([sum(a.elapsed_time(z) for a, z in ev)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = ops.launch_count() - n0
    # the depth maps of the sharded run, gathered for a checksum every N must reproduce within the parity tolerance
    counts = par.shard_counts(args.refs, world)
    full = par.all_gather_rows(depth.reshape(depth.shape[0], -1), counts)
    if heap is not None:
        heap.check()
    if rank == 0:
        sec = float(ms.item()) / 1e3
        print(json.dumps({
            'metric': 'ref-views/sec, one 64-view scene sharded over ranks (BASELINE C4)', 'unit': 'ref-views/s',
            'value': args.steps * args.refs / sec, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * sec / args.steps, 'scaling': 'strong', 'higher_is_better': True, 'dtype': 'f32',
            'data': 'synthetic', 'gpu_launches': launches,
            'config': {'workload': 'C4: 1 scene, %d refs + 7 halo keyframes, 256x320, D=96, 7 src, 4cm voxels, '
                                   '2x(scene model + 3 PointFlow)' % args.refs,
                       'collective': 'one all_gather_into_tensor of [N_g,35] fp32 rows per scene-model call (2 per step)',
                       'sparse_unet': ('sharded by voxel rows; epilogues store into every rank\'s symmetric buffer over '
                                       'NVLink, one flag barrier per layer') if heap is not None else 'redundant on every rank',
                       'all_gather_bytes_per_call': args.refs * plane[0] * plane[1] * 35 * 4,
                       'l2': 'flushed before every step', 'timing': 'CUDA events, max over ranks'},
            'depth_mean': float(full.double().mean().item()), 'depth_checksum': float(full.double().abs().sum().item()),
        }))
    if heap is not None:
        heap.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
