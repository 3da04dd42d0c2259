"""How the sparse U-Net layers of the 64-view scene (BASELINE configs[3]) scale when a rank computes only 1/w of the
rows of a level: per layer type and level, kernel time for the rows of rank r of w (w = 1, 2, 4, 8) on ONE GPU, no
peers.  This is the compute side of the row-sharded U-Net (csrc/engine.cu model_scene); the difference to the
measured sharded step is exchange + barriers + imbalance.
    python tools/unet_shard_scaling.py [--refs 64]
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--refs', type=int, default=64)
    ap.add_argument('--iters', type=int, default=5)
    args = ap.parse_args()
    sys.argv = sys.argv[:1]
    import bench
    ops = importlib.import_module('3dvnet_b200.ops')
    synth = importlib.import_module('3dvnet_b200.synth')
    lm = importlib.import_module('3dvnet_b200.mv3d.lightningmodel')
    par = importlib.import_module('3dvnet_b200.parallel')
    dev = torch.device('cuda', 0)
    b = synth.make_batch(1, args.refs + bench.N_SRC, bench.IMG_SIZE, bench.PLANE, 32, 4, 3, False, 0)
    net = lm.PL3DVNet(bench.DEPTH_CFG, bench.DEPTH_CFG, bench.EDGE_LEN, feat_dim=32, img_size=bench.IMG_SIZE)
    net.load_state_dict(synth.make_params(0), strict=False)
    net = net.to(dev).eval()
    fq, R, t, K = b.feats_quarter.to(dev), b.rotmats.to(dev), b.tvecs.to(dev), b.K.to(dev)
    e, ib = b.ref_src_edges, b.images_batch.to(dev)
    with torch.no_grad():
        _, d0 = net.hot_path(fq, R, t, K, e, ib, bench.DEPTH_CFG, bench.OFFSETS_LIST, return_init=True)
        ref_idx = torch.unique(e[0]).to(dev)
        pts, feat, _ = net.construct_feature_rich_pointcloud(d0, ib[ref_idx], fq, R, t, K, e)
        pts_batch = ib[ref_idx].unsqueeze(1).expand(args.refs, d0.shape[1] * d0.shape[2]).reshape(-1).contiguous()
        a_pts, a_idx, a_batch, seg, grid = ops.voxelize(pts, pts_batch, net.edge_len)
        dims = (int(grid.n_cells[0]), int(grid.n_cells[1]), int(grid.n_cells[2]), int(grid.n_batch))
        unet = net.sparse_conv
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed(fn):
            ts = []
            for i in range(args.iters + 2):
                flush.zero_()
                a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                z.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ts.append(a.elapsed_time(z) * 1e3)
            return float(np.median(ts))

        print('grid cells', dims)
        rows = []
        for world in (1, 2, 4, 8):
            for rank in sorted({0, world // 2, world - 1}):
                sc = par.ShardedScene(a_idx, a_batch, unet.n_levels, dims, world, rank)
                n = [lv.n for lv in sc.levels]
                if world == 1:
                    print('voxels per level', n)
                    print('pair-major tiles (128 pairs each): same', [m.n_tiles for m in sc.same], 'down', [m.n_tiles for m in sc.down],
                          'up', [m.n_tiles for m in sc.up], 'pairs path', [m.use_pairs for m in sc.same + sc.down + sc.up])
                x = [torch.randn(n[0], 64, device=dev), torch.randn(n[1], 128, device=dev), torch.randn(n[2], 128, device=dev)]

                def conv(xin, km, mod, gn, level, residual=None):
                    r0, r1 = sc.range[level]
                    y = torch.empty((n[level], mod.kernel.shape[-1]), device=dev)
                    wk, wp = mod.weights()
                    return lambda: ops.sparse_conv(xin, km, wk, gn.weight.detach(), gn.bias.detach(),
                                                   None if residual is None else residual[r0:r1], True, packed=wp,
                                                   workspace=sc.ws, out=y[r0:r1])
                layers = [
                    ('same L0 64->64', conv(x[0], sc.same[0], unet.res_down[0][0].conv1, unet.res_down[0][0].n1.gn, 0)),
                    ('down L0->L1', conv(x[0], sc.down[0], unet.down[0][0], unet.down[0][1].gn, 1)),
                    ('same L1 128->128', conv(x[1], sc.same[1], unet.res_down[1][0].conv1, unet.res_down[1][0].n1.gn, 1)),
                    ('down L1->L2', conv(x[1], sc.down[1], unet.down[1][0], unet.down[1][1].gn, 2)),
                    ('same L2 128->128', conv(x[2], sc.same[2], unet.res_down[2][0].conv1, unet.res_down[2][0].n1.gn, 2)),
                    ('up L2->L1', conv(x[2], sc.up[1], unet.up[0][0], unet.up[0][1].gn, 1)),
                    ('up L1->L0', conv(x[1], sc.up[0], unet.up[1][0], unet.up[1][1].gn, 0)),
                ]
                for name, fn in layers:
                    rows.append((world, rank, name, timed(fn)))
        names = [r[2] for r in rows if r[0] == 1]
        print('| layer | ' + ' | '.join('w=%d r=%d' % (w, r) for w, r, nm, _ in rows if nm == names[0]) + ' |')
        for nm in names:
            print('| %s | ' % nm + ' | '.join('%.0f' % us for w, r, n2, us in rows if n2 == nm) + ' |')
        # weights of the layer types in one U-Net pass: L0 same x4, L1 same x8, L2 same x6
        for world in (1, 2, 4, 8):
            tot = 0.0
            for nm, k in (('same L0 64->64', 4), ('down L0->L1', 1), ('same L1 128->128', 8), ('down L1->L2', 1),
                          ('same L2 128->128', 6), ('up L2->L1', 1), ('up L1->L0', 1)):
                tot += k * max(us for w, r, n2, us in rows if w == world and n2 == nm)
            print('world %d: sum over the layers of one U-Net pass (slowest rank of those measured): %.2f ms' % (world, tot / 1e3))


if __name__ == '__main__':
    main()
