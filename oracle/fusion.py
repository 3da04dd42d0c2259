"""ORACLE (test infrastructure): depth-map fusion on the CPU, restated from
/root/reference/mv3d/eval/pointcloudfusion_custom.py:10-116 (the .cuda() transfers and the
IMG_BATCH chunking removed; same arithmetic, fp32)."""
import numpy as np
import torch
import torch.nn.functional as F


def process_depth(ref_depth, src_depths, ref_P, src_Ps, ref_K, src_Ks, z_thresh=0.1, n_consistent_thresh=3):
    """-> pts_avg [h*w,3], n_valid [h*w] int64, valid [h,w] bool  (pointcloudfusion_custom.py:10-96)"""
    n_src = src_depths.shape[0]
    h, w = int(ref_depth.shape[0]), int(ref_depth.shape[1])
    n_pts = h * w
    ref_K_inv, src_Ks_inv, ref_P_inv = torch.inverse(ref_K), torch.inverse(src_Ks), torch.inverse(ref_P)
    xx, yy = np.meshgrid(np.linspace(0, w - 1, w), np.linspace(0, h - 1, h))
    pix = torch.from_numpy(np.stack((xx, yy, np.ones_like(xx)), axis=0)).float()
    pts = ref_P_inv[:3, :3] @ (ref_K_inv @ (pix * ref_depth.unsqueeze(0)).view(3, n_pts)) + ref_P_inv[:3, 3, None]   # :31-32
    reproj = torch.bmm(src_Ps[:, :3, :3], pts.unsqueeze(0).repeat(n_src, 1, 1)) + src_Ps[:, :3, 3, None]            # :46-47
    reproj = torch.bmm(src_Ks, reproj)
    z = reproj[:, 2]
    reproj = reproj / z.unsqueeze(1)
    valid_z = z > 1e-4
    valid_x = (reproj[:, 0] >= 0.) & (reproj[:, 0] <= float(w - 1))
    valid_y = (reproj[:, 1] >= 0.) & (reproj[:, 1] <= float(h - 1))
    grid = torch.clone(reproj[:, :2]).transpose(2, 1).view(n_src, n_pts, 1, 2)
    grid[..., 0] = (grid[..., 0] / float(w - 1)) * 2 - 1.0
    grid[..., 1] = (grid[..., 1] / float(h - 1)) * 2 - 1.0
    z_sample = F.grid_sample(src_depths.unsqueeze(1), grid, mode='nearest', align_corners=True, padding_mode='zeros')
    z_sample = z_sample.squeeze(1).squeeze(-1)
    valid_per_src = (torch.abs(z - z_sample) < z_thresh) & valid_x & valid_y & valid_z                             # :63-66
    n_valid = torch.sum(valid_per_src.int(), dim=0)
    pts_sample = torch.bmm(src_Ks_inv, reproj * z_sample.unsqueeze(1))                                              # :70-73
    pts_sample = torch.bmm(src_Ps[:, :3, :3].transpose(2, 1), pts_sample - src_Ps[:, :3, 3, None])
    valid = n_valid >= n_consistent_thresh
    pts_avg = pts.clone()
    for i in range(n_src):                                                                                          # :82-88
        ps = pts_sample[i]
        bad = torch.isnan(ps)
        ps = torch.where(bad, torch.zeros_like(ps), ps)
        vi = valid_per_src[i] & ~torch.any(bad, dim=0)
        pts_avg = pts_avg + ps * vi.float().unsqueeze(0)
    pts_avg = pts_avg / (n_valid + 1).float().unsqueeze(0)
    return pts_avg.transpose(1, 0), n_valid, valid.view(h, w)


def process_scene(depth_preds, poses, K, z_thresh, n_consistent_thresh):
    """every image against all the others (pointcloudfusion_custom.py:98-116) -> stacked per-image results"""
    n = depth_preds.shape[0]
    idx = torch.arange(n)
    out = [process_depth(depth_preds[i], depth_preds[idx != i], poses[i], poses[idx != i], K[i], K[idx != i], z_thresh,
                         n_consistent_thresh) for i in range(n)]
    return torch.stack([o[0] for o in out]), torch.stack([o[1] for o in out]), torch.stack([o[2] for o in out])
