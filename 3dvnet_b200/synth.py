"""Seeded synthetic ScanNet-shape inputs and reference-schema weights.

There is no dataset and no checkpoint in this environment, so tests, the golden-vector
script and bench.py all draw their inputs from here:

* cameras on a smooth circular trajectory inside a 5 x 4 x 2.7 m box room, consecutive
  keyframes ~0.1 apart in the reference's pose distance
  (sqrt(|t|^2 + 2/3 tr(I-R)), /root/reference/mv3d/dsets/frameselector.py:123),
  stored the way the reference stores them: world->camera ``rotmats``, ``tvecs = -R C``
  (/root/reference/mv3d/dsets/dataset.py:214-215), one full-resolution ``K`` per image
  (dataset.py:216);
* depth maps by ray/box intersection on the reference's plane lattice
  ``u = linspace(0, W-1, w)``, ``v = linspace(0, H-1, h)`` (mv3d/utils.py:67-83);
* ``ref_src_edges`` in the reference's layout: row 0 = reference image index, row 1 =
  source image index, collated across scenes by adding the running image count
  (mv3d/dsets/batch.py:19-29);
* weights under the reference's ``state_dict`` names, "sensitised"
  (non-trivial BN running statistics, non-zero GroupNorm gains on the residual branches,
  sharpened last convolutions) so that parity checks are sensitive to every layer.

Everything is numpy / CPU torch and deterministic for a given seed.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

ROOM = (5.0, 4.0, 2.7)


# ----------------------------------------------------------------------------- cameras
def make_intrinsics(img_size):
    """fx = fy = 0.9 W, principal point at the image centre."""
    H, W = img_size
    f = 0.9 * W
    return np.array([[f, 0, (W - 1) / 2.0], [0, f, (H - 1) / 2.0], [0, 0, 1]], dtype=np.float64)


def make_cameras(n_imgs, img_size=(256, 320), seed=0, room=ROOM):
    """Circular outward-looking trajectory. Returns float32 rotmats [n,3,3] (world->cam),
    tvecs [n,3], K [n,3,3]."""
    rng = np.random.RandomState(1000 + seed)
    radius = rng.uniform(0.8, 1.1)
    height = rng.uniform(1.2, 1.6)
    pitch = math.radians(rng.uniform(-8.0, 8.0))
    phi0 = rng.uniform(0, 2 * math.pi)
    step = 0.08 / radius  # ~0.08 m of arc per keyframe
    cx, cy = room[0] / 2.0, room[1] / 2.0
    rot = np.zeros((n_imgs, 3, 3))
    tv = np.zeros((n_imgs, 3))
    for k in range(n_imgs):
        phi = phi0 + k * step
        yaw = phi + math.radians(rng.uniform(-1.0, 1.0))
        th = pitch + math.radians(rng.uniform(-0.5, 0.5))
        C = np.array([cx + radius * math.cos(phi), cy + radius * math.sin(phi),
                      height + rng.uniform(-0.01, 0.01)])
        fwd = np.array([math.cos(yaw) * math.cos(th), math.sin(yaw) * math.cos(th), -math.sin(th)])
        right = np.array([math.sin(yaw), -math.cos(yaw), 0.0])
        down = np.cross(fwd, right)
        R_wc = np.stack([right, down, fwd], axis=1)  # camera axes as columns (cam->world)
        R = R_wc.T
        rot[k] = R
        tv[k] = -R @ C
    K = np.repeat(make_intrinsics(img_size)[None], n_imgs, axis=0)
    return rot.astype(np.float32), tv.astype(np.float32), K.astype(np.float32)


def ray_box_depth(rotmats, tvecs, K, img_size, plane_size, room=ROOM, clip=(0.5, 5.25)):
    """z-depth of the room's walls seen from every camera on the plane lattice."""
    H, W = img_size
    h, w = plane_size
    u = np.linspace(0, W - 1, w)
    v = np.linspace(0, H - 1, h)
    uu, vv = np.meshgrid(u, v)
    pix = np.stack([uu.ravel(), vv.ravel(), np.ones(h * w)])  # [3, P]
    n = rotmats.shape[0]
    out = np.zeros((n, h, w), dtype=np.float32)
    lo = np.zeros(3)
    hi = np.array(room)
    for i in range(n):
        R = rotmats[i].astype(np.float64)
        C = -R.T @ tvecs[i].astype(np.float64)
        d_cam = np.linalg.inv(K[i].astype(np.float64)) @ pix  # z component == 1
        d_w = R.T @ d_cam
        with np.errstate(divide='ignore', invalid='ignore'):
            t_lo = (lo[:, None] - C[:, None]) / d_w
            t_hi = (hi[:, None] - C[:, None]) / d_w
        t_exit = np.where(d_w > 0, t_hi, np.where(d_w < 0, t_lo, np.inf)).min(axis=0)
        out[i] = np.clip(t_exit, clip[0], clip[1]).reshape(h, w).astype(np.float32)
    return out


def make_edges(n_imgs, n_before=4, n_after=3, include_self=False):
    """Every image with a full neighbourhood is a reference; its sources are the trajectory
    neighbours p-n_before..p-1, p+1..p+n_after. ``include_self`` gives the
    reference's own eval topology, a window that contains the reference itself
    (/root/reference/mv3d/dsets/dataset.py:133-137)."""
    ref, src = [], []
    for p in range(n_before, n_imgs - n_after):
        for q in range(p - n_before, p + n_after + 1):
            if q == p and not include_self:
                continue
            ref.append(p)
            src.append(q)
    return np.array([ref, src], dtype=np.int64)


class SynthBatch:
    """Duck-typed stand-in for the reference's PyG ``Batch`` (mv3d/dsets/batch.py:6-16):
    ``images``/``feats_quarter``, ``rotmats``, ``tvecs``, ``K``, ``ref_src_edges``,
    ``images_batch`` plus the synthetic ground-truth ``depth_images``."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


def make_batch(n_scenes=1, n_imgs=8, img_size=(256, 320), plane_size=(56, 56), feat_dim=32,
               n_before=4, n_after=3, include_self=False, seed=0, with_images=False):
    """``n_scenes`` independent scenes of ``n_imgs`` keyframes each, collated the way PyG
    collates the reference's Batch. Features are N(0,1) at quarter resolution."""
    g = torch.Generator().manual_seed(seed)
    H, W = img_size
    rot, tv, Ks, edges, ib, depth = [], [], [], [], [], []
    for s in range(n_scenes):
        r, t, k = make_cameras(n_imgs, img_size, seed=seed * 131 + s)
        e = make_edges(n_imgs, n_before, n_after, include_self)
        refs = np.unique(e[0])
        depth.append(ray_box_depth(r[refs], t[refs], k[refs], img_size, plane_size))
        edges.append(e + s * n_imgs)
        rot.append(r)
        tv.append(t)
        Ks.append(k)
        ib.append(np.full(n_imgs, s, dtype=np.int64))
    n_total = n_scenes * n_imgs
    b = SynthBatch(
        rotmats=torch.from_numpy(np.concatenate(rot)),
        tvecs=torch.from_numpy(np.concatenate(tv)),
        K=torch.from_numpy(np.concatenate(Ks)),
        ref_src_edges=torch.from_numpy(np.concatenate(edges, axis=1)),
        images_batch=torch.from_numpy(np.concatenate(ib)),
        depth_images=torch.from_numpy(np.concatenate(depth)),
        feats_quarter=torch.randn(n_total, feat_dim, H // 4, W // 4, generator=g),
        img_size=img_size,
    )
    if with_images:
        b.images = torch.randn(n_total, 3, H, W, generator=g)
    return b


# ----------------------------------------------------------------------------- weights
def _bn(prefix, c, g, out):
    out[prefix + '.weight'] = torch.rand(c, generator=g) + 0.5
    out[prefix + '.bias'] = 0.1 * torch.randn(c, generator=g)
    out[prefix + '.running_mean'] = 0.1 * torch.randn(c, generator=g)
    out[prefix + '.running_var'] = torch.rand(c, generator=g) + 0.5
    out[prefix + '.num_batches_tracked'] = torch.tensor(1, dtype=torch.long)


def _uniform(shape, fan_in, g, gain=1.0):
    bound = gain * math.sqrt(3.0 / fan_in)
    return (torch.rand(*shape, generator=g) * 2 - 1) * bound


def make_costreg_params(seed=0, in_ch=32, base=8, sharpen=4.0):
    """CostRegNet under the reference's names (mv3d/subnetworks/mvsnet.py:133-152)."""
    g = torch.Generator().manual_seed(7000 + seed)
    p = OrderedDict()
    b = base
    convs = [(in_ch, b), (b, 2 * b), (2 * b, 2 * b), (2 * b, 4 * b), (4 * b, 4 * b), (4 * b, 8 * b), (8 * b, 8 * b)]
    for i, (ci, co) in enumerate(convs):
        p['conv%d.conv.weight' % i] = _uniform((co, ci, 3, 3, 3), ci * 27, g, gain=1.6)
        _bn('conv%d.bn' % i, co, g, p)
    for i, (ci, co) in zip((7, 8, 9), [(8 * b, 4 * b), (4 * b, 2 * b), (2 * b, b)]):
        # ConvTranspose3d weight layout is [Cin, Cout, 3, 3, 3]
        p['conv%d.deconv.weight' % i] = _uniform((ci, co, 3, 3, 3), ci * 27 / 8.0, g, gain=1.6)
        _bn('conv%d.bn' % i, co, g, p)
    p['prob.weight'] = sharpen * _uniform((1, b, 3, 3, 3), b * 27, g)
    p['prob.bias'] = 0.1 * torch.randn(1, generator=g)
    return p


def make_pointnet_params(seed=0, hidden=128, out_dim=64, in_dim=35):
    """PointNet (mv3d/subnetworks/scenemodeling.py:116-125)."""
    g = torch.Generator().manual_seed(7100 + seed)
    p = OrderedDict()
    for name, (co, ci) in (('fc_pos', (hidden, in_dim)), ('fc1', (hidden, hidden)), ('fc2', (hidden, 2 * hidden)),
                           ('fc3', (hidden, 2 * hidden)), ('fc4', (hidden, 2 * hidden)), ('fc_out', (out_dim, hidden))):
        p[name + '.weight'] = _uniform((co, ci), ci, g, gain=1.4)
        p[name + '.bias'] = 0.1 * torch.randn(co, generator=g)
    return p


def _gn(prefix, c, g, out):
    out[prefix + '.gn.weight'] = torch.rand(c, generator=g) + 0.5
    out[prefix + '.gn.bias'] = 0.1 * torch.randn(c, generator=g)


def make_sparse_unet_params(seed=0, dims=(64, 128, 128), n_res=(1, 2, 3)):
    """SparseUNet (mv3d/subnetworks/scenemodeling.py:147-189); MinkowskiEngine kernels are
    [27, Cin, Cout] (k=3) or [Cin, Cout] (k=1), no biases."""
    g = torch.Generator().manual_seed(7200 + seed)
    p = OrderedDict()

    def res(prefix, c):
        p[prefix + '.n1.gn.weight'] = torch.rand(c, generator=g) + 0.5
        p[prefix + '.n1.gn.bias'] = 0.1 * torch.randn(c, generator=g)
        p[prefix + '.n2.gn.weight'] = 0.5 * (torch.rand(c, generator=g) + 0.5)  # reference zero-inits this
        p[prefix + '.n2.gn.bias'] = 0.1 * torch.randn(c, generator=g)
        p[prefix + '.conv1.kernel'] = _uniform((27, c, c), 9 * c, g, gain=1.4)
        p[prefix + '.conv2.kernel'] = _uniform((27, c, c), 9 * c, g, gain=1.4)

    for i, n in enumerate(n_res):
        for l in range(n):
            res('res_down.%d.%d' % (i, l), dims[i])
    for i in range(1, len(dims)):
        p['down.%d.0.kernel' % (i - 1)] = _uniform((27, dims[i - 1], dims[i]), 9 * dims[i - 1], g, gain=1.4)
        _gn('down.%d.1' % (i - 1), dims[i], g, p)
    rdims = dims[::-1]
    rres = n_res[::-1]
    for i, n in enumerate(rres[1:]):
        for l in range(n):
            res('res_up.%d.%d' % (i, l), rdims[i + 1])
    for i in range(1, len(rdims)):
        p['up.%d.0.kernel' % (i - 1)] = _uniform((27, rdims[i - 1], rdims[i]), 4 * rdims[i - 1], g, gain=1.4)
        _gn('up.%d.1' % (i - 1), rdims[i], g, p)
        p['feat_adj.%d.0.kernel' % (i - 1)] = _uniform((2 * rdims[i], rdims[i]), 2 * rdims[i], g, gain=1.4)
        _gn('feat_adj.%d.1' % (i - 1), rdims[i], g, p)
    return p


def make_decoder_params(seed=0, in_dim=352, h_dim=128, sharpen=6.0):
    """HypothesisDecoder (mv3d/subnetworks/refinement.py:17-25)."""
    g = torch.Generator().manual_seed(7300 + seed)
    p = OrderedDict()
    for i, ci in enumerate((in_dim, h_dim, h_dim)):
        p['net.%d.0.weight' % i] = _uniform((h_dim, ci, 3), ci * 3, g, gain=1.4)
        _bn('net.%d.1' % i, h_dim, g, p)
    p['net.3.weight'] = sharpen * _uniform((1, h_dim, 3), h_dim * 3, g)
    p['net.3.bias'] = 0.1 * torch.randn(1, generator=g)
    return p


def make_params(seed=0, feat_dim=32):
    """All hot-path weights keyed like PL3DVNet's state_dict."""
    out = OrderedDict()
    for prefix, d in (('mvsnet.cnn_3d.', make_costreg_params(seed, feat_dim, 8)),
                      ('pointnet.', make_pointnet_params(seed, 4 * feat_dim, 2 * feat_dim, feat_dim + 3)),
                      ('sparse_conv.', make_sparse_unet_params(seed, (2 * feat_dim, 128, 128))),
                      ('decoder.', make_decoder_params(seed, 128 + 128 + 3 * feat_dim, 128))):
        for k, v in d.items():
            out[prefix + k] = v
    return out


def params_checksum(params):
    """Order-sensitive float64 checksum, stored in the golden files to detect RNG drift."""
    acc = 0.0
    for i, (k, v) in enumerate(params.items()):
        acc += (i + 1) * float(v.double().sum()) + float(v.double().abs().sum())
    return acc
