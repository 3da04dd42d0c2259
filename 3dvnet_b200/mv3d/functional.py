"""torch.autograd.Function drop-ins for the differentiable part of the two hot paths: the homography warp +
variance aggregation, volume level (/root/reference/mv3d/subnetworks/mvsnet.py:187-216) and point level
(/root/reference/mv3d/lightningmodel.py:147-169, :190-228).

The reference builds its sampling grid under ``torch.no_grad()`` and trains through ``F.grid_sample`` and the
two ``torch_scatter`` means, i.e. the gradient flows to the SOURCE FEATURE MAPS only (not to cameras, not to the
depth the points were back-projected with). The Functions below do the same: forward = the fused CUDA kernels of
csrc/planesweep.cu (no x_vox, no grid), backward = csrc/planesweep.cu:planesweep_var_bwd_kernel /
points_var_bwd_kernel, which recompute the samples from the saved inputs and scatter
``g * (2/n) (x_e - mean) * w_tap`` into an NHWC gradient map with 16-byte vector atomics."""
import torch

from .. import ops


def _to_nhwc(feats):
    f = feats.detach().float()
    if f.dim() == 4 and f.is_contiguous(memory_format=torch.channels_last) and not f.is_contiguous():
        return f.permute(0, 2, 3, 1)
    return ops.nchw_to_nhwc(f.contiguous())


class PlaneSweepVariance(torch.autograd.Function):
    """x_var [n_ref,C,D,h,w] = variance over the source views of the features warped into the reference frustum.
    Differentiable w.r.t. ``feats`` [n_imgs,C,Hf,Wf]."""

    @staticmethod
    def forward(ctx, feats, rotmats, tvecs, K, plan, depth_start, depth_interval, n_planes, plane_size, img_size):
        nhwc = _to_nhwc(feats)
        cams = ops.camera_tables(rotmats.detach().float().contiguous(), tvecs.detach().float().contiguous(),
                                 K.detach().float().contiguous())
        ctx.save_for_backward(nhwc, cams)
        ctx.plan, ctx.args = plan, (float(depth_start), float(depth_interval), int(n_planes), tuple(plane_size),
                                    tuple(img_size))
        ctx.in_dtype = feats.dtype
        return ops.planesweep_var(nhwc, cams, plan, depth_start, depth_interval, n_planes, tuple(plane_size),
                                  tuple(img_size))

    @staticmethod
    def backward(ctx, grad_out):
        nhwc, cams = ctx.saved_tensors
        d0, dd, D, plane, img = ctx.args
        g = ops.planesweep_var_backward(nhwc, cams, ctx.plan, d0, dd, D, plane, img, grad_out.float().contiguous())
        # NHWC storage viewed as [n,C,Hf,Wf]: a channels-last gradient, what a channels-last backbone wants anyway
        return (g.permute(0, 3, 1, 2).to(ctx.in_dtype),) + (None,) * 9


class PointVariance(torch.autograd.Function):
    """World points of every pixel hypothesis (no gradient, as in the reference) and their variance features
    [n_ref*P, n_hyp, C], differentiable w.r.t. ``feats``."""

    @staticmethod
    def forward(ctx, feats, rotmats, tvecs, K, plan, depth, img_size, n_side, offset):
        nhwc = _to_nhwc(feats)
        cams = ops.camera_tables(rotmats.detach().float().contiguous(), tvecs.detach().float().contiguous(),
                                 K.detach().float().contiguous())
        depth = depth.detach().float().contiguous()
        pts, feat = ops.points_var(nhwc, cams, plan, depth, tuple(img_size), int(n_side), float(offset))
        ctx.save_for_backward(nhwc, cams, depth)
        ctx.plan, ctx.args, ctx.in_dtype = plan, (tuple(img_size), int(n_side), float(offset)), feats.dtype
        ctx.mark_non_differentiable(pts)
        return pts, feat

    @staticmethod
    def backward(ctx, _grad_pts, grad_feat):
        nhwc, cams, depth = ctx.saved_tensors
        img, n_side, offset = ctx.args
        g = ops.points_var_backward(nhwc, cams, ctx.plan, depth, img, n_side, offset, grad_feat.float().contiguous())
        return (g.permute(0, 3, 1, 2).to(ctx.in_dtype),) + (None,) * 8


def planesweep_variance(feats, rotmats, tvecs, K, ref_src_edges, depth_start, depth_interval, n_planes, plane_size,
                        img_size):
    plan = ops.edge_plan(ref_src_edges, feats.device)
    return PlaneSweepVariance.apply(feats, rotmats, tvecs, K, plan, depth_start, depth_interval, n_planes, plane_size,
                                    img_size)


def point_variance(feats, rotmats, tvecs, K, ref_src_edges, depth, img_size, n_side=0, offset=0.0):
    plan = ops.edge_plan(ref_src_edges, feats.device)
    return PointVariance.apply(feats, rotmats, tvecs, K, plan, depth, img_size, n_side, offset)
