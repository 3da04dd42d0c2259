"""ORACLE (test infrastructure): the hot path end to end on the CPU, restated from
/root/reference/mv3d/lightningmodel.py:124-130,176-242 and the refinement schedule of
/root/reference/mv3d/eval-3dvnet.py:23,65-99 (2 outer iterations x offsets
[0.05, 0.05, 0.025], `depth += offset`). Starts from quarter-resolution feature maps: the
2D backbone/FPN that produces them is outside the hot path."""
import torch

from . import planesweep, costreg, pointcloud, scenemodel
from .voxelize import voxelize

OFFSETS_LIST = [[0.05, 0.05, 0.025], [0.05, 0.05, 0.025]]


def sub(params, prefix):
    return {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}


def initial_depth(feats_quarter, rotmats, tvecs, K, ref_src_edges, depth_cfg, img_size, params, return_all=False):
    """MVSNet.forward after the 2D feature extraction (mvsnet.py:187-229)."""
    d0, dd, D, size = depth_cfg['depth_start'], depth_cfg['depth_interval'], depth_cfg['n_intervals'], depth_cfg['size']
    x_var = planesweep.planesweep_var(feats_quarter, rotmats, tvecs, K, ref_src_edges, d0, dd, D, img_size, size)
    x_reg = costreg.costregnet(x_var, sub(params, 'mvsnet.cnn_3d.')).squeeze(1)
    depth = costreg.soft_argmin(x_reg, d0, dd, D)
    if return_all:
        return depth, x_var, x_reg
    return depth


def model_scene(depth, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges, edge_len, img_size, params,
                return_all=False):
    """lightningmodel.py:176-185."""
    pts, pts_feat, pts_batch = pointcloud.feature_rich_pointcloud(depth, depth_batch, img_feats, rotmats, tvecs, K,
                                                                  ref_src_edges, img_size)
    a_pts, a_idx, a_batch, a_edges = voxelize(pts.numpy(), pts_batch.numpy(), edge_len)
    a_pts, a_idx, a_batch, a_edges = (torch.from_numpy(a_pts), torch.from_numpy(a_idx), torch.from_numpy(a_batch),
                                      torch.from_numpy(a_edges))
    x = torch.cat((pts[a_edges[1]] - a_pts[a_edges[0]], pts_feat[a_edges[1]]), dim=1)
    x = scenemodel.pointnet(x, a_edges[0], a_pts.shape[0], sub(params, 'pointnet.'))
    xs = scenemodel.sparse_unet(x, a_pts, a_idx, a_batch, edge_len, sub(params, 'sparse_conv.'))
    if return_all:
        return xs, dict(pts=pts, pts_feat=pts_feat, pts_batch=pts_batch, anchor_pts=a_pts, anchor_idx3d=a_idx,
                        anchor_batch=a_batch, anchor_pts_edges=a_edges, pointnet=x)
    return xs


def run_pointflow(xs, depth, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges, offset, n_side, img_size,
                  params, return_all=False):
    """lightningmodel.py:187-242 -> expected offset [n_ref,h,w]."""
    pts_hyp, pts_feat, pts_batch = pointcloud.hypothesis_points(depth, depth_batch, img_feats, rotmats, tvecs, K,
                                                                ref_src_edges, offset, n_side, img_size)
    prob = scenemodel.hypothesis_decoder(xs, pts_hyp, pts_feat, pts_batch, sub(params, 'decoder.'))
    vals = torch.linspace(-n_side * offset, n_side * offset, 2 * n_side + 1).type_as(prob).unsqueeze(0)
    out = torch.sum(vals * prob, dim=1).view(depth.shape)
    if return_all:
        return out, prob
    return out


def refine(depth, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges, edge_len, img_size, params,
           offsets_list=OFFSETS_LIST):
    """eval-3dvnet.py:73-99 without the python chunking (whole scene at once)."""
    depth = depth.clone()
    for offsets in offsets_list:
        xs = model_scene(depth, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges, edge_len, img_size, params)
        for offset in offsets:
            depth += run_pointflow(xs, depth, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges, offset, 3,
                                   img_size, params)
    return depth


def hot_path(feats_quarter, rotmats, tvecs, K, ref_src_edges, images_batch, depth_cfg, edge_len, img_size, params):
    """Initial depth + volumetric refinement: one bench "step" (BASELINE.json configs[1])."""
    depth = initial_depth(feats_quarter, rotmats, tvecs, K, ref_src_edges, depth_cfg, img_size, params)
    ref_idx = torch.unique(ref_src_edges[0])
    depth_batch = images_batch[ref_idx]
    return refine(depth, depth_batch, feats_quarter, rotmats, tvecs, K, ref_src_edges, edge_len, img_size, params)
