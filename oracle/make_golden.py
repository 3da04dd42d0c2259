"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference is pure Python; its third-party imports that cannot be installed offline are
satisfied by oracle/shims (torch_scatter, torch_geometric.voxel_grid, pytorch_lightning,
open3d stubs, and the CPU MinkowskiEngine restatement). The 2D backbone is replaced by a
pass-through that returns seeded synthetic feature maps, so that everything recorded here
is produced by the reference's own hot-path code: mvsnet.py:179-229,
lightningmodel.py:124-242, utils.py:38-108, scenemodeling.py, refinement.py.

Weights are NOT stored: they are regenerated from `synth.make_params(seed)`; each file
carries `params_checksum` so RNG drift is detected instead of silently failing parity.
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
sys.path[:0] = [os.path.join(HERE, 'shims'), REF, ROOT]

import torchvision  # noqa: E402

_orig_mnas = torchvision.models.mnasnet1_0
torchvision.models.mnasnet1_0 = lambda pretrained=False, **kw: _orig_mnas(weights=None)

synth = importlib.import_module('3dvnet_b200.synth')
from mv3d.lightningmodel import PL3DVNet  # noqa: E402  (the reference)
from mv3d import utils as ref_utils  # noqa: E402
from mv3d.eval.metricfunctions import calc_2d_depth_metrics  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


class _PassExtractor(nn.Module):
    def forward(self, images):
        return (images,)


class _FixedShrinker(nn.Module):
    def __init__(self):
        super().__init__()
        self.feats = None

    def forward(self, images):
        fq = self.feats
        return None, fq, None, None, None


def build_reference_net(img_size, depth_cfg, edge_len, seed):
    torch.manual_seed(seed)
    net = PL3DVNet(depth_cfg, depth_cfg, edge_len, feat_dim=32, img_size=img_size)
    params = synth.make_params(seed)
    sd = net.state_dict()
    hot = [k for k in sd if k.startswith(('mvsnet.cnn_3d.', 'pointnet.', 'sparse_conv.', 'decoder.'))]
    assert sorted(hot) == sorted(params.keys()), (set(hot) ^ set(params.keys()))
    for k in hot:
        assert tuple(sd[k].shape) == tuple(params[k].shape), (k, sd[k].shape, params[k].shape)
    net.load_state_dict(params, strict=False)
    net.mvsnet.feat_extractor = _PassExtractor()
    net.mvsnet.feat_shrinker = _FixedShrinker()
    net.eval()
    return net, params


def run_reference(batch, img_size, depth_cfg, edge_len, seed, offsets_list):
    net, params = build_reference_net(img_size, depth_cfg, edge_len, seed)
    net.mvsnet.feat_shrinker.feats = batch.feats_quarter
    batch.images = torch.zeros(batch.feats_quarter.shape[0], 3, *img_size)
    cap = {}
    h = net.mvsnet.cnn_3d.register_forward_hook(lambda m, i, o: cap.update(x_var=i[0].detach(), x_reg=o.detach()))
    out = {}
    with torch.no_grad():
        depth, depth_batch, _, fq, _, ref_idx = net.make_initial_depth_predictions(batch, depth_cfg)
        h.remove()
        out.update(x_var=cap['x_var'], x_reg=cap['x_reg'].squeeze(1), depth_init=depth.clone(), ref_idx=ref_idx)
        args = (batch.feats_quarter, batch.rotmats, batch.tvecs, batch.K, batch.ref_src_edges)
        # intermediates of the first scene-modelling pass
        pts, pts_feat, pts_batch = net.construct_feature_rich_pointcloud(depth, depth_batch, *args)
        a_pts, a_idx, a_batch, a_edges = ref_utils.voxelize(pts, pts_batch, net.edge_len)
        out.update(pts=pts, pts_feat=pts_feat, pts_batch=pts_batch, anchor_pts=a_pts, anchor_idx3d=a_idx,
                   anchor_batch=a_batch, anchor_pts_edges=a_edges)
        x = torch.cat((pts[a_edges[1]] - a_pts[a_edges[0]], pts_feat[a_edges[1]]), dim=1)
        out['pointnet'] = net.pointnet(x, a_edges[0], a_pts.shape[0])
        first = True
        for offsets in offsets_list:
            xs = net.model_scene(depth, depth_batch, *args)
            if first:
                for li, lv in enumerate(xs):
                    out['xs%d_feats' % li] = lv['feats']
                    out['xs%d_pts' % li] = lv['pts']
                    out['xs%d_idx' % li] = lv['idx']
                    out['xs%d_batch' % li] = lv['batch']
            for offset in offsets:
                off = net.run_pointflow(xs, depth, depth_batch, *args, offset, 3)
                if first:
                    out['offset0'] = off.clone()
                    first = False
                depth += off
        out['depth_final'] = depth
    return out, params


def case_pipeline(name, n_scenes, n_imgs, img_size, plane, D, interval, edge_len, n_before, n_after, include_self,
                  seed, offsets_list):
    depth_cfg = dict(depth_start=0.5, depth_interval=interval, n_intervals=D, size=plane)
    b = synth.make_batch(n_scenes, n_imgs, img_size, plane, 32, n_before, n_after, include_self, seed)
    out, params = run_reference(b, img_size, depth_cfg, edge_len, seed, offsets_list)
    m = calc_2d_depth_metrics(out['depth_final'], out['depth_init'])
    print(name, 'n_ref', out['depth_init'].shape[0], 'voxels', out['anchor_pts'].shape[0],
          'levels', [out['xs%d_feats' % i].shape for i in range(3)],
          'depth_init std %.4f' % out['depth_init'].std().item(),
          'refine abs-rel vs init %.4f' % m['abs_rel'].item())
    save = dict(
        feats_quarter=b.feats_quarter.numpy(), rotmats=b.rotmats.numpy(), tvecs=b.tvecs.numpy(), K=b.K.numpy(),
        ref_src_edges=b.ref_src_edges.numpy(), images_batch=b.images_batch.numpy(),
        img_size=np.array(img_size), plane=np.array(plane), D=np.array(D), depth_start=np.array(0.5),
        depth_interval=np.array(interval), edge_len=np.array(edge_len), seed=np.array(seed),
        offsets=np.array(offsets_list, dtype=np.float64), params_checksum=np.array(synth.params_checksum(params)))
    for k, v in out.items():
        save['ref_' + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **save)


def case_irregular_edges(name, seed):
    """Path A + the feature-rich point cloud of the reference on edge lists its collation can produce but the
    pipeline cases do not: ragged (1 / 3 / 12 sources per reference), shuffled, non-consecutive references, a
    reference whose only edge is a self-edge, a duplicated edge."""
    img_size, plane, D = (64, 80), (16, 16), 16
    depth_cfg = dict(depth_start=0.5, depth_interval=0.3, n_intervals=D, size=plane)
    b = synth.make_batch(1, 13, img_size, plane, 32, 2, 2, False, seed)
    g = torch.Generator().manual_seed(seed)
    edges = [(0, 1)] + [(3, s) for s in (1, 2, 4)] + [(6, s) for s in range(13) if s != 6] + [(9, 9)] + \
            [(11, 10), (11, 10), (11, 12)]
    e = torch.tensor(edges, dtype=torch.int64).t().contiguous()
    b.ref_src_edges = e[:, torch.randperm(e.shape[1], generator=g)].contiguous()
    net, params = build_reference_net(img_size, depth_cfg, 0.3, seed)
    net.mvsnet.feat_shrinker.feats = b.feats_quarter
    b.images = torch.zeros(b.feats_quarter.shape[0], 3, *img_size)
    cap = {}
    h = net.mvsnet.cnn_3d.register_forward_hook(lambda m, i, o: cap.update(x_var=i[0].detach()))
    with torch.no_grad():
        depth, depth_batch, _, _, _, ref_idx = net.make_initial_depth_predictions(b, depth_cfg)
        h.remove()
        pts, pts_feat, pts_batch = net.construct_feature_rich_pointcloud(depth, depth_batch, b.feats_quarter, b.rotmats,
                                                                         b.tvecs, b.K, b.ref_src_edges)
    print(name, 'edges', tuple(b.ref_src_edges.shape), 'refs', ref_idx.tolist(), 'x_var', tuple(cap['x_var'].shape))
    np.savez_compressed(
        os.path.join(GOLD, name + '.npz'), feats_quarter=b.feats_quarter.numpy(), rotmats=b.rotmats.numpy(),
        tvecs=b.tvecs.numpy(), K=b.K.numpy(), ref_src_edges=b.ref_src_edges.numpy(), images_batch=b.images_batch.numpy(),
        img_size=np.array(img_size), plane=np.array(plane), D=np.array(D), depth_start=np.array(0.5),
        depth_interval=np.array(0.3), edge_len=np.array(0.3), seed=np.array(seed),
        params_checksum=np.array(synth.params_checksum(params)), ref_x_var=cap['x_var'].numpy(),
        ref_depth_init=depth.numpy(), ref_ref_idx=ref_idx.numpy(), ref_pts=pts.numpy(), ref_pts_feat=pts_feat.numpy(),
        ref_pts_batch=pts_batch.numpy())


def case_voxelize(name, seed):
    """utils.voxelize on adversarial point sets: points exactly on cell boundaries, an
    extent that is an exact multiple of the edge (ceil vs trunc+1), several batches."""
    rng = np.random.RandomState(seed)
    e = 0.04
    pts = rng.uniform(-1.0, 1.5, size=(4000, 3)).astype(np.float32)
    pts[:200] = (np.round(pts[:200] / e) * e).astype(np.float32)        # on boundaries
    pts[0] = (0.0, 0.0, 0.0)
    pts[1] = (np.float32(e) * 50, np.float32(e) * 25, np.float32(e) * 10)  # exact-multiple extents
    pts[2:200] = np.clip(pts[2:200], 0, None)
    pts = np.clip(pts, 0.0, [np.float32(e) * 50, np.float32(e) * 25, np.float32(e) * 10]).astype(np.float32)
    batch = np.sort(rng.randint(0, 3, size=pts.shape[0])).astype(np.int64)
    a_pts, a_idx, a_batch, a_edges = ref_utils.voxelize(torch.from_numpy(pts), torch.from_numpy(batch), e)
    print(name, 'anchors', a_pts.shape[0])
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), pts=pts, batch=batch, edge_len=np.array(e),
                        ref_anchor_pts=a_pts.numpy(), ref_anchor_idx3d=a_idx.numpy(),
                        ref_anchor_batch=a_batch.numpy(), ref_anchor_pts_edges=a_edges.numpy())


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    # C1 of BASELINE.json: 1 ref + 2 src, 64x80, D=16, plane 16x16 (multiple-of-8 rule)
    case_pipeline('c1_tiny', 1, 3, (64, 80), (16, 16), 16, 0.3, 0.3, 1, 1, False, 0, [[0.05, 0.05, 0.025]])
    # reference eval topology (window incl. the reference itself), two collated scenes, ragged planes
    case_pipeline('c1_selfedge_2scenes', 2, 6, (64, 80), (16, 24), 16, 0.3, 0.2, 2, 2, True, 1,
                  [[0.05, 0.05, 0.025], [0.05, 0.05, 0.025]])
    case_voxelize('voxelize_adversarial', 3)
    case_irregular_edges('c1_irregular_edges', 4)
