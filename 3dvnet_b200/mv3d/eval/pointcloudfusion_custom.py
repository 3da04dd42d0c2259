"""Depth-map fusion — drop-in for /root/reference/mv3d/eval/pointcloudfusion_custom.py
(process_depth :10-96, process_scene :98-116), the consumer of the path's depth maps
(SURVEY.md §8f.4). One kernel launch fuses every reference image of a scene against all the
others (csrc/fusion.cu); the reference loops over images in Python and materialises
[n_src,3,h*w] tensors per image."""
import ctypes

import numpy as np
import torch

from ... import ops


def fuse(depths, poses, K, n_ref, z_thresh, n_consistent_thresh):
    """depths [n,h,w], poses [n,4,4] world->camera, K [n,3,3] (CUDA fp32); references are images
    0..n_ref-1, sources of a reference are all the other images.
    -> pts_avg [n_ref,h*w,3] f32, n_valid [n_ref,h*w] i32, valid [n_ref,h,w] bool"""
    ops._chk(depths, torch.float32, 'depths', 3), ops._chk(poses, torch.float32, 'poses', 3), ops._chk(K, torch.float32, 'K', 3)
    n, h, w = depths.shape
    dev = depths.device
    L = ops.lib()
    ws_bytes = L.raw('dv3d_depth_fusion_workspace_bytes')(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    pts = torch.empty((n_ref, h * w, 3), dtype=torch.float32, device=dev)
    n_valid = torch.empty((n_ref, h * w), dtype=torch.int32, device=dev)
    valid = torch.empty((n_ref, h * w), dtype=torch.uint8, device=dev)
    L.call('dv3d_depth_fusion', ops._p(depths), ops._p(poses), ops._p(K), n, n_ref, h, w, float(z_thresh),
           int(n_consistent_thresh), ops._p(ws), ws_bytes, ops._p(pts), ops._p(n_valid), ops._p(valid), ops._stream())
    return pts, n_valid, valid.view(n_ref, h, w).bool()


def process_depth(ref_depth, ref_image, src_depths, src_images, ref_P, src_Ps, ref_K, src_Ks, z_thresh=0.1,
                  n_consistent_thresh=3):
    """pointcloudfusion_custom.py:10-96 -> (pts_filtered [m,3], rgb_filtered [m,3], valid [h,w]) numpy"""
    dev = torch.device('cuda')
    depths = torch.cat((ref_depth.unsqueeze(0), src_depths), 0).float().to(dev).contiguous()
    poses = torch.cat((ref_P.unsqueeze(0), src_Ps), 0).float().to(dev).contiguous()
    K = torch.cat((ref_K.unsqueeze(0), src_Ks), 0).float().to(dev).contiguous()
    pts, _, valid = fuse(depths, poses, K, 1, z_thresh, n_consistent_thresh)
    v = valid[0]
    pts_filtered = pts[0][v.view(-1)].cpu().numpy()
    rgb_filtered = ref_image.to(dev)[v].view(-1, 3).cpu().numpy()
    return pts_filtered, rgb_filtered, v.cpu().numpy()


def process_scene(depth_preds, images, poses, K, z_thresh, n_consistent_thresh):
    """pointcloudfusion_custom.py:98-116 -> (fused_pts [m,3], fused_rgb [m,3], all_valid [n,h,w]) numpy"""
    dev = torch.device('cuda')
    d = depth_preds.float().to(dev).contiguous()
    pts, _, valid = fuse(d, poses.float().to(dev).contiguous(), K.float().to(dev).contiguous(), d.shape[0], z_thresh,
                         n_consistent_thresh)
    flat = valid.view(valid.shape[0], -1)
    fused_pts = pts[flat].cpu().numpy()
    fused_rgb = images.to(dev)[valid].view(-1, 3).cpu().numpy()
    return fused_pts, fused_rgb, valid.cpu().numpy()
