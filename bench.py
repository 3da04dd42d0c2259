"""Benchmark of the 3DVNet hot path (BASELINE.json): ref-views/sec at 256x320, D=96, 7 source
views, synthetic ScanNet-shape batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of BASELINE.json configs[1]: 1 reference
+ 7 source views (8 images, quarter-resolution features 32x64x80), plane-sweep cost volume ->
CostRegNet -> soft-argmin, then 2 x (scene model at 4 cm voxels + 3 PointFlow passes).
N > 1: one process per GPU (torchrun), every rank runs its own scenes (no collective on the
data path; weak scaling), time = max over ranks.

`--impl reference` times the reference algorithm's CPU restatement (oracle/, the only other
place this file may execute it) on the host cores for the same metric and config.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

IMG_SIZE, PLANE, D, N_SRC = (256, 320), (56, 56), 96, 7
DEPTH_CFG = dict(depth_start=0.5, depth_interval=0.05, n_intervals=D, size=PLANE)
EDGE_LEN = 0.04
OFFSETS_LIST = [[0.05, 0.05, 0.025], [0.05, 0.05, 0.025]]  # eval-3dvnet.py:23
METRIC = 'ref-views/sec at 256x320, D=96, 7 src'
WORKLOAD = 'C2: 1 ref + 7 src, 256x320, D=96, plane 56x56, C=32, 4cm voxels, 2x(scene model + 3 PointFlow)'
# algorithmic bytes of the warp+variance kernel per reference view:
# every source feature map read once + the [C,D,h,w] slab written once
ALGO_BYTES = N_SRC * 32 * 64 * 80 * 4 + 32 * D * PLANE[0] * PLANE[1] * 4


def synth_inputs(seed, refs_per_step):
    synth = importlib.import_module('3dvnet_b200.synth')
    n_imgs = refs_per_step + N_SRC
    return synth.make_batch(1, n_imgs, IMG_SIZE, PLANE, 32, 4, 3, False, seed), synth.make_params(0)


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples of one GPU during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap'}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.02)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def cpu_step(b, params):
    """The reference algorithm's CPU restatement of one step (oracle/pipeline.py)."""
    from oracle import pipeline
    with torch.no_grad():
        return pipeline.hot_path(b.feats_quarter, b.rotmats, b.tvecs, b.K, b.ref_src_edges, b.images_batch,
                                 DEPTH_CFG, EDGE_LEN, IMG_SIZE, params)


def cpu_full_step(net_cpu, b, images, params):
    """SURVEY.md section 8d's full pipeline on the CPU: the backbone / FPN modules themselves (torchvision, fp32)
    -> the oracle's hot path -> the oracle's upsampling cascade (oracle/upsample.py)"""
    from oracle import pipeline, upsample
    with torch.no_grad():
        fh, fq, _, _, _ = net_cpu.mvsnet.feat_shrinker(*net_cpu.mvsnet.feat_extractor(images))
        depth = pipeline.hot_path(fq, b.rotmats, b.tvecs, b.K, b.ref_src_edges, b.images_batch, DEPTH_CFG, EDGE_LEN,
                                  IMG_SIZE, params)
        ref_idx = torch.unique(b.ref_src_edges[0])
        sd = net_cpu.state_dict()
        return depth, upsample.upsample_cascade(depth, fq[ref_idx], fh[ref_idx], images[ref_idx],
                                                pipeline.sub(sd, 'refine_quarter.'), pipeline.sub(sd, 'refine_half.'),
                                                pipeline.sub(sd, 'refine_full.')), fq


def full_model(params, device):
    """PL3DVNet with seeded random backbone / PropagationNet weights (no checkpoint exists offline) + the
    synthetic hot-path weights; identical on every call"""
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    torch.manual_seed(1234)
    net = lm.PL3DVNet(DEPTH_CFG, DEPTH_CFG, EDGE_LEN, feat_dim=32, img_size=IMG_SIZE)
    net.load_state_dict(params, strict=False)
    return net.to(device).eval()


def synth_images(seed, n_imgs):
    return torch.randn(n_imgs, 3, *IMG_SIZE, generator=torch.Generator().manual_seed(9000 + seed))


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b, params = synth_inputs(0, args.refs_per_step)
    for _ in range(args.warmup):
        cpu_step(b, params)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(b, params)
    dt = time.perf_counter() - t0
    value = args.steps * args.refs_per_step / dt
    # supplementary: SURVEY section 8d's FULL pipeline (backbone + FPN -> hot path -> 3 PropagationNets) on the CPU
    net_cpu = full_model(params, 'cpu')
    images = synth_images(0, args.refs_per_step + N_SRC)
    cpu_full_step(net_cpu, b, images, params)
    t0 = time.perf_counter()
    n_full = max(1, min(args.steps, 3))
    for _ in range(n_full):
        cpu_full_step(net_cpu, b, images, params)
    dt_full = (time.perf_counter() - t0) / n_full
    sample = '%d step(s) of the full workload (each %d ref view(s)), torch CPU threads=%d' % (
        args.steps, args.refs_per_step, cores)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'ref-views/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'refs_per_step': args.refs_per_step,
                   'imgs_per_step': args.refs_per_step + N_SRC,
                   'parallelism': 'rank 0 only: the reference is single-process (mv3d/config.py:3-5)',
                   'reference_arm': 'port (oracle/pipeline.py, the CPU restatement; the reference itself needs '
                                    'torch_scatter / torch_geometric / MinkowskiEngine, absent here)',
                   'l2': 'n/a (CPU)', 'timing': 'host wall clock around the timed steps'},
        'cpu_baseline': {'value': value, 'unit': 'ref-views/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'ref-views/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'e2e_full': {'value': args.refs_per_step / dt_full, 'unit': 'ref-views/s', 'ms_per_step': 1e3 * dt_full,
                     'pipeline': 'images -> MnasNet + FPN (torch CPU) -> oracle hot path -> oracle PropagationNet x3 -> '
                                 'full-resolution depth (eval-3dvnet.py:58-125)', 'steps': n_full},
        'gpu_launches': 0}), flush=True)


def run_c4(args, rank, world, dev, dist, ops, lm):
    """BASELINE configs[3] inside the N > 1 line: ONE scene of 64 reference views whose views are sharded over the
    ranks (3dvnet_b200/parallel.py): one NCCL all-gather of the [N_g, 35] fp32 point rows per scene-model call, the
    sparse U-Net sharded by voxel rows with epilogue stores into every rank's symmetric buffer over NVLink
    (csrc/symm.cu). STRONG scaling. Checked here, on the driver's hardware, against the single-GPU pass of the same
    scene computed on rank 0: depth abs-rel and bit-equality of the voxel indices of every sparse level."""
    par = importlib.import_module('3dvnet_b200.parallel')
    synth = importlib.import_module('3dvnet_b200.synth')
    refs, steps, warm = args.c4_refs, max(2, min(args.steps, 4)), 2
    b = synth.make_batch(1, refs + N_SRC, IMG_SIZE, PLANE, 32, 4, 3, False, 0)   # identical on every rank (seeded)
    net = lm.PL3DVNet(DEPTH_CFG, DEPTH_CFG, EDGE_LEN, feat_dim=32, img_size=IMG_SIZE)
    net.load_state_dict(synth.make_params(0), strict=False)
    net = net.to(dev).eval()
    fq, R, t, K = b.feats_quarter.to(dev), b.rotmats.to(dev), b.tvecs.to(dev), b.K.to(dev)
    e, ib = b.ref_src_edges, b.images_batch.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    heap = par.SymmHeap(max(2048 << 20, par.native_heap_bytes(net, refs, PLANE)))
    splan = par.ShardPlan(e, ib, dev, world, rank)

    def step():   # one native call per rank (csrc/engine.cu dv3d_hot_path_sharded)
        return par.hot_path_sharded_native(net, fq, R, t, K, e, ib, DEPTH_CFG, OFFSETS_LIST, heap, plan=splan)

    def step_composed():   # round 1's path: the same schedule composed op by op from Python, NCCL all-gather
        return par.hot_path_sharded(net, fq, R, t, K, e, ib, DEPTH_CFG, OFFSETS_LIST, heap=heap)

    def timed(fn):
        for _ in range(warm):
            out = fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        n0, ep0 = ops.launch_count(), heap.epoch
        for a, z in ev:
            flush.zero_()
            a.record()
            out = fn()
            z.record()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        launches, barriers = ops.launch_count() - n0, (heap.epoch - ep0) // steps
        heap.check()
        ms = torch.tensor([sum(a.elapsed_time(z) for a, z in ev)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return out, ms, launches, barriers

    with torch.no_grad():
        (depth_c, _), ms_c, _, _ = timed(step_composed)
        os.environ['DV3D_SHARD_BALANCE'] = '1'      # coarse levels cut by work (csrc/sparse.cu) instead of equal row counts
        (depth_b, _), ms_b, _, _ = timed(step)
        os.environ['DV3D_SHARD_BALANCE'] = '0'      # the default: equal row counts on every level, as the composed path shards
        os.environ['DV3D_SHARD_NEIGHBOUR_WAIT'] = '0'   # every barrier waits for every rank
        (depth_f, _), ms_f, _, _ = timed(step)
        os.environ['DV3D_SHARD_NEIGHBOUR_WAIT'] = '1'   # the default: a layer's barrier waits for the ranks that send rows
        (depth, rng), ms, launches, barriers = timed(step)
        native_equals_composed = bool(torch.equal(depth, depth_c))
        full_equals_neighbour = bool(torch.equal(depth, depth_f))
        # where the step goes: one more (untimed) step with the engine's stage events, on every rank
        ops.engine_profile(True)
        dist.barrier()
        step()
        torch.cuda.synchronize()
        acc = {}
        for sid, v in ops.engine_profile_read():
            acc[ops.STAGE_NAMES[sid]] = acc.get(ops.STAGE_NAMES[sid], 0.0) + v
        ops.engine_profile(False)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {k: round(v, 3) for k, v in acc.items()})
        stage_ms = per_rank[0]
        stage_minmax = {k: [min(r.get(k, 0.0) for r in per_rank), max(r.get(k, 0.0) for r in per_rank)] for k in per_rank[0]}
        counts = par.shard_counts(refs, world)
        full = par.all_gather_rows(depth.reshape(depth.shape[0], -1).contiguous(), counts)

        # parity on this hardware: the single-GPU pass of the same scene (rank 0), its initial depth broadcast so
        # that every rank voxelises the same points in the sharded scene model
        h, w = PLANE
        d_init = torch.empty((refs, h, w), dtype=torch.float32, device=dev)
        abs_rel, ms_one, stage_one = None, None, None
        if rank == 0:
            d_one, d0 = net.hot_path(fq, R, t, K, e, ib, DEPTH_CFG, OFFSETS_LIST, return_init=True)
            d_init.copy_(d0)
            abs_rel = float((torch.abs(full.view_as(d_one) - d_one) / (d_one + 1e-7)).mean().item())
            # the same scene on ONE GPU (rank 0, the other ranks idle), same timing rules: the strong-scaling base
            ev1 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            torch.cuda.synchronize()
            for a, z in ev1:
                flush.zero_()
                a.record()
                net.hot_path(fq, R, t, K, e, ib, DEPTH_CFG, OFFSETS_LIST)
                z.record()
            torch.cuda.synchronize()
            ms_one = sum(a.elapsed_time(z) for a, z in ev1) / steps
            ops.engine_profile(True)
            net.hot_path(fq, R, t, K, e, ib, DEPTH_CFG, OFFSETS_LIST)
            torch.cuda.synchronize()
            acc = {}
            for sid, v in ops.engine_profile_read():
                acc[ops.STAGE_NAMES[sid]] = acc.get(ops.STAGE_NAMES[sid], 0.0) + v
            ops.engine_profile(False)
            stage_one = {k: round(v, 3) for k, v in acc.items()}
        dist.broadcast(d_init, 0)
        start, end = par.shard_range(refs, world, rank)
        xs_sh = par.model_scene_sharded(net, d_init[start:end].contiguous(), ib, fq, R, t, K, e, heap=heap)
        heap.check()
        idx_equal, feat_err = None, None
        if rank == 0:
            ref_idx = torch.unique(e[0]).to(dev)
            xs_one = net.model_scene(d_init, ib[ref_idx], fq, R, t, K, e)
            idx_equal = all(torch.equal(a['idx'], r['idx']) and torch.equal(a['batch'], r['batch'])
                            for a, r in zip(xs_sh, xs_one))
            feat_err = max(float((a['feats'] - r['feats']).abs().max() / r['feats'].abs().max())
                           for a, r in zip(xs_sh, xs_one)) if idx_equal else None
    heap.close()
    sec = float(ms.item()) / 1e3
    return {'workload': 'C4: 1 scene, %d refs + 7 halo keyframes, 256x320, D=96, 7 src, 4cm voxels, 2x(scene model + '
                        '3 PointFlow), reference views sharded over %d ranks' % (refs, world),
            'value': steps * refs / sec, 'unit': 'ref-views/s', 'scaling': 'strong', 'steps': steps, 'warmup': warm,
            'ms_per_step': 1e3 * sec / steps,
            'single_gpu': None if ms_one is None else {
                'ms_per_step': ms_one, 'value': refs / (ms_one * 1e-3), 'stage_ms': stage_one,
                'note': 'the same 64-view scene through PL3DVNet.hot_path on rank 0 alone (engine, one native call)'},
            'speedup_vs_single_gpu': None if ms_one is None else ms_one / (1e3 * sec / steps),
            'abs_rel_vs_single_gpu': abs_rel, 'voxel_idx_equal': idx_equal,
            'sparse_feat_max_rel_err_vs_single_gpu': feat_err,
            'composed_from_python': {'ms_per_step': float(ms_c.item()) / steps,
                                     'note': 'the same schedule op by op from Python with an NCCL all-gather (round 1)',
                                     'depth_bit_identical_to_native': native_equals_composed},
            'native_work_balanced_rows_ms_per_step': float(ms_b.item()) / steps,
            'native_full_barriers': {'ms_per_step': float(ms_f.item()) / steps, 'depth_bit_identical': full_equals_neighbour,
                                     'note': 'every layer barrier waits for all ranks instead of the ranks that send rows'},
            'stage_ms_rank0': stage_ms, 'stage_ms_min_max_over_ranks': stage_minmax,
            'stage_note': "'barrier' is part of 'unet'; 'levels' runs on a side stream beside 'pointnet'",
            'exchange_bytes': 2 * refs * PLANE[0] * PLANE[1] * 35 * 4, 'exchanges': 2, 'barriers': barriers,
            'gpu_launches_per_rank': launches // steps,
            'collective': 'point rows [N_local,3]+[N_local,32] fp32 copied into every peer heap (NVLink peer copies) + one '
                          'flag barrier, once per scene-model call; sparse U-Net layers exchange rows by epilogue stores '
                          'into peer memory + one flag barrier per layer; no NCCL call inside the step',
            'timing': 'CUDA events on the launching stream, L2 flushed per step, barrier both sides, max over ranks'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--refs-per-step', type=int, default=1,
                    help='reference views per step (BASELINE configs[1] = 1; larger values are a sweep, not the metric)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--c4-refs', type=int, default=64,
                    help='N > 1 only: reference views of the single sharded scene of the supplementary "c4" block '
                         '(BASELINE configs[3]; 0 = skip)')
    ap.add_argument('--streams', type=int, default=4,
                    help='supplementary measurement: this many steps in flight on separate CUDA streams, one host '
                         'thread each (0 = skip); the headline value / e2e stay single-stream')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — the product path has no CPU fallback '
                         '(use --impl reference for the CPU baseline)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    # the in-tree library: rank 0 (re)builds it if needed, the others wait and only load it
    if local == 0:
        importlib.import_module('3dvnet_b200.build').build()
    if dist is not None:
        dist.barrier()
    ops = importlib.import_module('3dvnet_b200.ops')
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    warm = max(args.warmup, 3)

    # every rank works on its own scenes (seed = rank): no data-path collective
    b, params = synth_inputs(rank, args.refs_per_step)
    net = lm.PL3DVNet(DEPTH_CFG, DEPTH_CFG, EDGE_LEN, feat_dim=32, img_size=IMG_SIZE)
    net.load_state_dict(params, strict=False)
    net = net.to(dev).eval()

    host = {k: getattr(b, k).pin_memory() for k in ('feats_quarter', 'rotmats', 'tvecs', 'K', 'images_batch')}
    edges = b.ref_src_edges  # host int64 [2,E]
    resident = {k: v.to(dev) for k, v in host.items()}
    n_ref = args.refs_per_step
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # resident arm: EVERYTHING the step reads is already in HBM, including the CSR form of the edge list
    plan_resident = ops.edge_plan(edges, dev)

    def step_resident():
        return net.hot_path(resident['feats_quarter'], resident['rotmats'], resident['tvecs'], resident['K'],
                            plan_resident, resident['images_batch'], DEPTH_CFG, OFFSETS_LIST)

    # End to end: every step uploads its inputs from pinned host memory and reads its depth map back, all inside
    # the timed region, as a two-deep pipeline (what a server does): step i+1's inputs travel on a copy stream while
    # step i computes, and step i's result is read back asynchronously - the host blocks on it one step later, so
    # the host-side work of a step (edge CSR build, launches) overlaps the tail of the previous step on the GPU.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    out_hosts = [torch.empty((n_ref,) + PLANE, dtype=torch.float32).pin_memory() for _ in range(2)]
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]     # the step that read bufs[slot] has finished with it
    read_back = [torch.cuda.Event(), torch.cuda.Event()]    # out_hosts[slot] holds its step's depth map
    e2e_state = {'i': 0, 'primed': False}

    def upload(slot):
        copy_stream.wait_event(consumed[slot])
        with torch.cuda.stream(copy_stream):
            for k, v in host.items():
                bufs[slot][k].copy_(v, non_blocking=True)
            uploaded[slot].record(copy_stream)

    for ev_ in consumed + read_back:
        ev_.record()

    def step_e2e():
        i = e2e_state['i']
        slot = i % 2
        if not e2e_state['primed']:   # first step of a timed / warm-up sequence uploads its own inputs
            upload(slot)
            e2e_state['primed'] = True
        torch.cuda.current_stream().wait_event(uploaded[slot])
        upload((i + 1) % 2)
        d = bufs[slot]
        read_back[slot].synchronize()  # the result of step i-2 has been consumed by the host: its buffer is free
        # a FRESH host edge tensor every step: the CSR build on the host and its upload (ops.EdgePlan) are paid
        # inside the timed region, like the other inputs (ops.edge_plan caches on tensor identity)
        depth = net.hot_path(d['feats_quarter'], d['rotmats'], d['tvecs'], d['K'], edges.clone(), d['images_batch'],
                             DEPTH_CFG, OFFSETS_LIST)
        consumed[slot].record()
        out_hosts[slot].copy_(depth, non_blocking=True)
        read_back[slot].record()
        e2e_state['i'] = i + 1
        return out_hosts[slot]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """per-step CUDA events on the launching stream; L2 flushed before every step, outside the events"""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s, e in ev:
            flush.zero_()
            s.record()
            fn()
            e.record()
        barrier()
        return sum(s.elapsed_time(e) for s, e in ev)  # ms

    for _ in range(warm):
        step_resident()
    torch.cuda.synchronize()

    # `value`: the timed region proper, nothing but the steps (no stage events inside the engine)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ops.launch_count()
    ms = timed(step_resident, args.steps)
    launches = ops.launch_count() - launches0
    # kernel / stage timing: a SECOND pass of the same K steps under the same conditions (L2 flushed per step)
    # with the engine recording CUDA events on the launching stream around its stages (dv3d_engine_profile)
    ops.engine_profile(True)
    ms_profiled = timed(step_resident, args.steps)
    recs = ops.engine_profile_read()
    ops.engine_profile(False)
    sampler.stop_flag = True
    sampler.join()
    stage_ms = {}
    for sid, t in recs:
        stage_ms.setdefault(sid, []).append(t)
    k_ms = float(np.mean(stage_ms[ops.STAGE_PLANESWEEP]))
    g_ms = float(np.mean(stage_ms[ops.STAGE_DEC_GEMM0]))
    stages_per_step = {ops.STAGE_NAMES[sid]: float(np.sum(v)) / args.steps for sid, v in sorted(stage_ms.items())}

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    copy_stream.synchronize()

    # Supplementary: SURVEY section 8d's FULL pipeline, end to end through the public API (PL3DVNet.full_pass):
    # images + cameras uploaded from pinned host memory, backbone + FPN (cuDNN, channels-last, fp32 storage, TF32
    # convolutions allowed - PyTorch's default, what the reference runs with), the hot path (engine), the three
    # PropagationNets, full-resolution depth read back.  Two steps in flight like `e2e`: uploads on the copy stream,
    # results awaited one step later.
    full = None
    if world == 1:
        net_full = full_model(params, dev)
        img_host = synth_images(rank, n_ref + N_SRC).pin_memory()
        img_devs = [torch.empty_like(img_host, device=dev) for _ in range(2)]
        full_hosts = [torch.empty((n_ref,) + IMG_SIZE, dtype=torch.float32).pin_memory() for _ in range(2)]
        smalls = [{k: torch.empty_like(v, device=dev) for k, v in host.items() if k != 'feats_quarter'} for _ in range(2)]
        up2, cons2, rb2 = ([torch.cuda.Event(), torch.cuda.Event()] for _ in range(3))
        st2 = {'i': 0, 'primed': False}
        for ev_ in cons2 + rb2:
            ev_.record()

        def upload_full(slot):
            copy_stream.wait_event(cons2[slot])
            with torch.cuda.stream(copy_stream):
                img_devs[slot].copy_(img_host, non_blocking=True)
                for k in smalls[slot]:
                    smalls[slot][k].copy_(host[k], non_blocking=True)
                up2[slot].record(copy_stream)

        def step_full():
            i = st2['i']
            slot = i % 2
            if not st2['primed']:
                upload_full(slot)
                st2['primed'] = True
            torch.cuda.current_stream().wait_event(up2[slot])
            upload_full((i + 1) % 2)
            rb2[slot].synchronize()
            sm = smalls[slot]
            out = net_full.full_pass(img_devs[slot], sm['rotmats'], sm['tvecs'], sm['K'], edges.clone(), sm['images_batch'],
                                     OFFSETS_LIST)
            cons2[slot].record()
            full_hosts[slot].copy_(out['final'], non_blocking=True)
            rb2[slot].record()
            st2['i'] = i + 1
            return out

        for _ in range(3):
            step_full()
        ms_full = timed(step_full, args.steps)
        copy_stream.synchronize()
        full = {'value': args.steps * n_ref / (ms_full * 1e-3), 'unit': 'ref-views/s', 'ms_per_step': ms_full / args.steps,
                'h2d_bytes_per_step': int(img_host.numel() * 4 + sum(host[k].numel() * host[k].element_size() for k in smalls[0])
                                          + edges.numel() * 4 + 64),
                'd2h_bytes_per_step': int(full_hosts[0].numel() * 4),
                'pipeline': 'images -> MnasNet + FPN (torchvision / cuDNN, channels-last, fp32 storage, TF32 convolutions '
                            'allowed as by PyTorch default) -> dv3d_hot_path -> PropagationNet x3 (tcgen05 gather-GEMM) -> '
                            'full-resolution depth (eval-3dvnet.py:58-125); two steps in flight: uploads on a copy stream, '
                            'results awaited one step later'}
        del net_full

    # Supplementary: S independent steps in flight (one host thread + CUDA stream each). A single C2
    # step is ~240 short dependent kernels that leave most SMs idle; a server fills the GPU this way.
    ms_multi = None
    if args.streams > 1:
        import threading
        S = args.streams
        streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
        per = max(1, args.steps // S)

        def worker(i, n):
            torch.cuda.set_device(dev)
            with torch.cuda.stream(streams[i]):
                for _ in range(n):
                    step_resident()

        def run_all(n):
            th = [threading.Thread(target=worker, args=(i, n)) for i in range(S)]
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()

        run_all(2)  # warm-up: arenas and side streams of every stream
        barrier()
        start = torch.cuda.Event(enable_timing=True)
        start.record()
        for st_ in streams:
            st_.wait_event(start)
        run_all(per)
        ends = []
        for st_ in streams:
            e_ = torch.cuda.Event(enable_timing=True)
            e_.record(st_)
            ends.append(e_)
        barrier()
        ms_multi = max(start.elapsed_time(e_) for e_ in ends)
        units_multi = per * S * n_ref

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cnt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        launches = int(cnt.item())
    ms, ms_e2e = float(t[0]), float(t[1])
    units = args.steps * n_ref * world

    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        peak, peak_src = float(pk['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        tpeak, tpeak_src = float(pk.get('bf16_tflops_sustained', pk['bf16_tflops'])), \
            'measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        tpeak, tpeak_src = 1400.0, 'fallback (B200_PROFILING.md, sustained)'
    achieved = ALGO_BYTES * n_ref / (k_ms * 1e-3) / 1e9
    traffic = {}
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    # the dominant kernel: the one-kernel PointFlow decoder (csrc/decoder_fused.cu, tcgen05), one launch per
    # PointFlow pass. Useful FLOPs 2 M (K N) summed over its layers, with M = n_ref * 3136 points * 7 hypothesis
    # rows (the 8th row of a point is zero padding and is NOT counted): Conv1d k=3 352->128, 128->128, 128->128,
    # 128->1 (refinement.py:17-25)
    gemm_flops = 2.0 * (n_ref * PLANE[0] * PLANE[1] * 7) * (3 * (352 + 128 + 128) * 128 + 3 * 128)
    g_achieved = gemm_flops / (g_ms * 1e-3) / 1e12

    line = {
        'metric': METRIC, 'value': units / (ms * 1e-3), 'unit': 'ref-views/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': warm, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'refs_per_step': n_ref, 'imgs_per_step': n_ref + N_SRC,
                   'parallelism': 'scenes sharded over %d rank(s), no collective' % world,
                   'l2': 'flushed (256 MiB memset) before every timed step, outside the per-step events',
                   'timing': 'sum of per-step CUDA events on the launching stream, max over ranks; stage / kernel '
                             'times come from a second pass of the same steps with engine stage events on',
                   'reference_arm': 'port (oracle/pipeline.py, the CPU restatement; the reference itself needs '
                                    'torch_scatter / torch_geometric / MinkowskiEngine, absent here)'},
        'e2e': {'value': units / (ms_e2e * 1e-3), 'unit': 'ref-views/s',
                'h2d_bytes_per_step': int(sum(v.numel() * v.element_size() for v in host.values())
                                          + edges.numel() * 4 + 64),
                'd2h_bytes_per_step': int(out_hosts[0].numel() * 4), 'ms_per_step': ms_e2e / args.steps,
                'pipeline': 'two steps in flight: uploads on a copy stream, results read back asynchronously and '
                            'awaited one step later; every upload and read-back of the K steps is inside the timed region'},
        'e2e_full': full,
        'gpu_launches': launches,
        'roofline': {'kernel': 'decoder_fused_kernel (whole PointFlow decoder of %d points x 7 hypotheses in one tcgen05 '
                               'launch: 3 x Conv1d+BN+ReLU, head, softmax, depth update)' % (n_ref * PLANE[0] * PLANE[1]),
                     'bound': 'tensor', 'achieved': g_achieved, 'peak': tpeak, 'unit': 'TFLOP/s',
                     'frac': g_achieved / tpeak, 'traffic': traffic.get('decoder_fused_kernel'),
                     'peak_source': tpeak_src, 'algorithmic_flops_per_launch': gemm_flops, 'kernel_ms': g_ms,
                     'note': 'useful fp32-grade FLOPs; the kernel issues 3 TF32 MMAs per useful one (3xTF32 split) and '
                             'TF32 runs at half the bf16 rate, so 1/6 of the bf16 peak is its ceiling; at 1 ref view the '
                             '196 row tiles take 2 rounds on 148 SMs (1.32 rounds of work)'},
        'roofline_warp': {'kernel': 'planesweep_var_kernel', 'bound': 'hbm', 'achieved': achieved, 'peak': peak,
                          'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic.get('planesweep_var_kernel'),
                          'peak_source': peak_src, 'algorithmic_bytes_per_launch': ALGO_BYTES * n_ref,
                          'kernel_ms': k_ms},
        'stages_ms_per_step': stages_per_step, 'ms_per_step_with_stage_events': ms_profiled / args.steps,
        'multi_stream': None if ms_multi is None else {
            'streams': args.streams, 'steps': per * args.streams, 'value': units_multi * world / (ms_multi * 1e-3),
            'unit': 'ref-views/s', 'note': 'supplementary, rank 0 only: independent steps in flight on separate CUDA '
                                           'streams (one host thread each), inputs resident, no L2 flush (the '
                                           'concurrent arenas exceed L2); value / e2e above are single-stream'},
        'clocks': sampler.summary(),
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_step(b, params)  # warm-up (lazy init)
        n_cpu = 3
        t0 = time.perf_counter()
        for _ in range(n_cpu):
            ref = cpu_step(b, params)
        dt = time.perf_counter() - t0
        got = step_resident().cpu()
        rel = (torch.abs(got - ref) / (ref + 1e-7)).mean().item()
        line['cpu_baseline'] = {'value': n_cpu * n_ref / dt, 'unit': 'ref-views/s', 'cores': cores, 'kind': 'port',
                                'sample': '%d full steps of the same workload through oracle/pipeline.py '
                                          '(torch CPU, %d threads)' % (n_cpu, cores)}
        line['abs_rel_vs_oracle'] = rel
    if dist is not None and args.c4_refs > 0:
        # free the weak-scaling buffers first: the 64-view scene needs ~6 GB per rank
        del bufs, resident, flush
        ops._arena.clear()
        torch.cuda.empty_cache()
        try:
            line['c4'] = run_c4(args, rank, world, dev, dist, ops, lm)
        except Exception as exc:  # the headline line must survive a failure of the supplementary block
            line['c4'] = {'error': '%s: %s' % (type(exc).__name__, exc)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
