"""torch-tensor front end of the C ABI (include/dv3d.h): checks dtypes / layouts, allocates
outputs with torch (device memory + streams are torch's job here, nothing else) and
enqueues the CUDA kernels of lib3dvnet_b200.so on the current stream.

No function in this module has a CPU or eager-PyTorch fallback: tensors must live on a CUDA
device and the shared library must load, otherwise the call raises.
"""
import ctypes
import os

import numpy as np
import torch

from ._lib import lib, VoxelGrid, Dv3dError, DV3D_ENOSPC  # noqa: F401

ROWS_PER_POINT = 8  # decoder operand: 7 hypotheses + 1 zero row (csrc/decoder.cu)

# Which kernel runs the dense contractions (csrc/gemm.cu / gemm_tc.cu):
#   'tf32x3' tcgen05 tensor cores, 3xTF32 split accumulation (fp32-grade)  [default]
#   'tf32'   tcgen05 tensor cores, plain TF32
#   'f32'    fp32 CUDA-core kernel (verification path)
GEMM_MODES = ('tf32x3', 'tf32', 'f32')
_gemm_mode = None


def set_gemm_mode(mode):
    global _gemm_mode
    if mode not in GEMM_MODES:
        raise ValueError('gemm mode must be one of %s, got %r' % (GEMM_MODES, mode))
    if mode != 'f32':
        lib().call('dv3d_set_gemm_precision', 1 if mode == 'tf32x3' else 2)
    _gemm_mode = mode


def gemm_mode():
    if _gemm_mode is None:
        set_gemm_mode(os.environ.get('DV3D_GEMM', 'tf32x3'))
    return _gemm_mode


# Arithmetic of the plane-sweep warp + variance kernel (csrc/planesweep.cu):
#   'exact'  the reference's fp32 operation chain, x_var bit-identical to the CPU PyTorch path (default)
#   'fast'   affine-in-depth projection + reciprocal, x_var within ~2e-5 of its scale, 12 % faster (DV3D_WARP=fast)
WARP_MODES = ('exact', 'fast')


CONV3D_MODES = ('tc', 'ffma')


def set_conv3d_mode(mode):
    """'tc': the first CostRegNet layer (32 -> 8) on tcgen05 with 3xTF32 arithmetic (csrc/conv3d_tc.cu, default);
    'ffma': every layer on the fp32 CUDA-core kernels (DV3D_CONV3D=ffma)"""
    if mode not in CONV3D_MODES:
        raise ValueError('conv3d mode must be one of %s, got %r' % (CONV3D_MODES, mode))
    lib().call('dv3d_set_conv3d_mode', CONV3D_MODES.index(mode))


def conv3d_mode():
    return CONV3D_MODES[int(lib().raw('dv3d_get_conv3d_mode')())]


def set_warp_mode(mode):
    if mode not in WARP_MODES:
        raise ValueError('warp mode must be one of %s, got %r' % (WARP_MODES, mode))
    lib().call('dv3d_set_warp_mode', WARP_MODES.index(mode))


def warp_mode():
    return WARP_MODES[int(lib().raw('dv3d_get_warp_mode')())]


def pack_weights(w_kn):
    """[K, N] fp32 (K % 32 == 0, N in {64,128}) -> the tcgen05 kernel's packed image, or None
    in 'f32' mode."""
    if gemm_mode() == 'f32':
        return None
    _chk(w_kn, torch.float32, 'weight', 2)
    K, N = w_kn.shape
    nbytes = lib().raw('dv3d_gemm_pack_bytes')(K, N)
    if nbytes == 0:
        raise RuntimeError('pack_weights: K must be a positive multiple of 32, got K=%d N=%d' % (K, N))
    out = torch.empty(nbytes, dtype=torch.uint8, device=w_kn.device)
    lib().call('dv3d_gemm_pack_weights', _p(w_kn), K, N, _p(out), _stream())
    return out


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype, name, ndim=None):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError('%s must be a CUDA tensor (3dvnet_b200 has no CPU path)' % name)
    if t.dtype != dtype:
        raise RuntimeError('%s must be %s, got %s' % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError('%s must be contiguous' % name)
    if ndim is not None and t.dim() != ndim:
        raise RuntimeError('%s must have %d dims, got shape %s' % (name, ndim, tuple(t.shape)))
    return t


def launch_count():
    return int(lib().raw('dv3d_launch_count')())


# ----------------------------------------------------------------------------- edges
class EdgePlan(object):
    """CSR view of ``ref_src_edges`` [2,E] (row 0 = reference image, row 1 = source image):
    references are ``unique(row 0)`` in ascending order (mvsnet.py:179), the edges of every
    reference are kept in their original relative order. Replaces utils.slice_edges'
    O(E*range) compare and the gather_idx of mvsnet.py:179 / lightningmodel.py:134."""

    def __init__(self, ref_src_edges, device):
        e = ref_src_edges.detach().cpu().numpy() if torch.is_tensor(ref_src_edges) else np.asarray(ref_src_edges)
        if e.ndim != 2 or e.shape[0] != 2:
            raise RuntimeError('ref_src_edges must be [2, n_edges], got %s' % (e.shape,))
        ref_idx, gather = np.unique(e[0], return_inverse=True)
        order = np.argsort(gather, kind='stable')
        counts = np.bincount(gather, minlength=len(ref_idx))
        rowptr = np.zeros(len(ref_idx) + 1, dtype=np.int32)
        np.cumsum(counts, out=rowptr[1:])
        self.n_ref = int(len(ref_idx))
        self.n_edges = int(e.shape[1])
        packed = np.concatenate([rowptr, e[1][order].astype(np.int32), e[0][order].astype(np.int32),
                                 ref_idx.astype(np.int32)])
        dev = torch.from_numpy(packed).to(device, non_blocking=True)
        n, E = self.n_ref, self.n_edges
        self.rowptr, self.edge_src, self.edge_ref, self.ref_img = dev[:n + 1], dev[n + 1:n + 1 + E], \
            dev[n + 1 + E:n + 1 + 2 * E], dev[n + 1 + 2 * E:]
        self.ref_idx = torch.from_numpy(ref_idx.astype(np.int64)).to(device, non_blocking=True)
        self.device = device


_plan_cache = {}


def edge_plan(ref_src_edges, device):
    """Cached EdgePlan for an edge tensor (keyed on storage + version)."""
    if isinstance(ref_src_edges, EdgePlan):
        return ref_src_edges
    key = (ref_src_edges.data_ptr(), ref_src_edges._version, tuple(ref_src_edges.shape), str(device))
    hit = _plan_cache.get(key)
    if hit is None:
        if len(_plan_cache) > 64:
            _plan_cache.clear()
        # the entry keeps the edge tensor alive: its address cannot be handed to another tensor while cached
        hit = _plan_cache[key] = (EdgePlan(ref_src_edges, device), ref_src_edges)
    return hit[0]


# ----------------------------------------------------------------------------- path A
def nchw_to_nhwc(x):
    _chk(x, torch.float32, 'features', 4)
    n, C, H, W = x.shape
    out = torch.empty((n, H, W, C), dtype=torch.float32, device=x.device)
    lib().call('dv3d_nchw_to_nhwc', _p(x), _p(out), n, C, H * W, _stream())
    return out


def camera_tables(rotmats, tvecs, K):
    """[n_imgs,36] = Kinv | P | R | t per image, in the reference's own fp32 arithmetic"""
    _chk(rotmats, torch.float32, 'rotmats', 3), _chk(tvecs, torch.float32, 'tvecs', 2), _chk(K, torch.float32, 'K', 3)
    n = rotmats.shape[0]
    out = torch.empty((n, 36), dtype=torch.float32, device=rotmats.device)
    lib().call('dv3d_camera_tables', _p(rotmats), _p(tvecs), _p(K), n, _p(out), _stream())
    return out


def planesweep_var(feats_nhwc, cams, plan, depth_start, depth_interval, n_planes, plane_size, img_size, out=None):
    """x_var [n_ref,C,D,h,w] (mvsnet.py:187-216)."""
    _chk(feats_nhwc, torch.float32, 'feats_nhwc', 4)
    n_imgs, Hf, Wf, C = feats_nhwc.shape
    h, w = plane_size
    if out is None:
        out = torch.empty((plan.n_ref, C, n_planes, h, w), dtype=torch.float32, device=feats_nhwc.device)
    lib().call('dv3d_planesweep_var', _p(feats_nhwc), n_imgs, C, Hf, Wf, _p(cams), _p(plan.ref_img), _p(plan.rowptr),
               _p(plan.edge_src), plan.n_ref, float(depth_start), float(depth_interval), int(n_planes), h, w,
               img_size[0], img_size[1], _p(out), _stream())
    return out


def points_var(feats_nhwc, cams, plan, depth, img_size, n_side, offset, feat_out=None, feat_off=0):
    """World points of every pixel hypothesis + their variance feature
    (lightningmodel.py:132-174 with n_side=0, :187-235 with n_side=3).
    -> pts [n_ref*P, n_hyp, 3], feat [n_ref*P, rows, ld] (rows = n_hyp or the padded 8)."""
    _chk(depth, torch.float32, 'depth', 3)
    n_imgs, Hf, Wf, C = feats_nhwc.shape
    n_ref, h, w = depth.shape
    n_hyp = 2 * n_side + 1
    Np = n_ref * h * w
    pts = torch.empty((Np, n_hyp, 3), dtype=torch.float32, device=depth.device)
    if feat_out is None:
        feat_out = torch.empty((Np, n_hyp, C), dtype=torch.float32, device=depth.device)
    rows, ld = feat_out.shape[1], feat_out.shape[2]
    lib().call('dv3d_points_var', _p(feats_nhwc), n_imgs, C, Hf, Wf, _p(cams), _p(plan.ref_img), _p(plan.rowptr),
               _p(plan.edge_src), _p(depth), n_ref, h, w, img_size[0], img_size[1], n_side, float(offset), _p(pts),
               _p(feat_out), rows, ld, feat_off, _stream())
    return pts, feat_out


def planesweep_var_backward(feats_nhwc, cams, plan, depth_start, depth_interval, n_planes, plane_size, img_size, grad_x_var):
    """gradient of planesweep_var w.r.t. feats_nhwc (mvsnet.py:209-216 trained through grid_sample + scatter mean)"""
    _chk(feats_nhwc, torch.float32, 'feats_nhwc', 4), _chk(grad_x_var, torch.float32, 'grad_x_var', 5)
    n_imgs, Hf, Wf, C = feats_nhwc.shape
    h, w = plane_size
    if tuple(grad_x_var.shape) != (plan.n_ref, C, n_planes, h, w):
        raise RuntimeError('grad_x_var must be %s, got %s' % ((plan.n_ref, C, n_planes, h, w), tuple(grad_x_var.shape)))
    grad = torch.zeros_like(feats_nhwc, memory_format=torch.contiguous_format)
    lib().call('dv3d_planesweep_var_backward', _p(feats_nhwc), n_imgs, C, Hf, Wf, _p(cams), _p(plan.ref_img), _p(plan.rowptr),
               _p(plan.edge_src), plan.n_ref, float(depth_start), float(depth_interval), int(n_planes), h, w, img_size[0],
               img_size[1], _p(grad_x_var), _p(grad), _stream())
    return grad


def points_var_backward(feats_nhwc, cams, plan, depth, img_size, n_side, offset, grad_feat, feat_off=0):
    """gradient of points_var's variance features [n_pts, rows, ld] (channels feat_off..) w.r.t. feats_nhwc"""
    _chk(feats_nhwc, torch.float32, 'feats_nhwc', 4), _chk(depth, torch.float32, 'depth', 3)
    _chk(grad_feat, torch.float32, 'grad_feat', 3)
    n_imgs, Hf, Wf, C = feats_nhwc.shape
    n_ref, h, w = depth.shape
    rows, ld = grad_feat.shape[1], grad_feat.shape[2]
    if grad_feat.shape[0] != n_ref * h * w:
        raise RuntimeError('grad_feat must have %d rows, got %d' % (n_ref * h * w, grad_feat.shape[0]))
    grad = torch.zeros_like(feats_nhwc, memory_format=torch.contiguous_format)
    lib().call('dv3d_points_var_backward', _p(feats_nhwc), n_imgs, C, Hf, Wf, _p(cams), _p(plan.ref_img), _p(plan.rowptr),
               _p(plan.edge_src), _p(depth), n_ref, h, w, img_size[0], img_size[1], int(n_side), float(offset),
               _p(grad_feat), rows, ld, int(feat_off), _p(grad), _stream())
    return grad


def conv3d_bn_relu(x, weight, scale, shift, stride=1, skip=None):
    _chk(x, torch.float32, 'x', 5), _chk(weight, torch.float32, 'weight', 5)
    n, Cin, D, H, W = x.shape
    Cout = weight.shape[0]
    if stride == 1:
        oshape = (n, Cout, D, H, W)
    else:
        oshape = (n, Cout, (D + 1) // 2, (H + 1) // 2, (W + 1) // 2)
    y = torch.empty(oshape, dtype=torch.float32, device=x.device)
    lib().call('dv3d_conv3d_bn_relu', _p(x), n, Cin, D, H, W, _p(weight), _p(scale), _p(shift), Cout, stride,
               _p(skip), _p(y), _stream())
    return y


def deconv3d_bn_relu(x, weight, scale, shift, skip=None):
    n, Cin, D, H, W = x.shape
    Cout = weight.shape[1]
    y = torch.empty((n, Cout, 2 * D, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
    lib().call('dv3d_deconv3d_bn_relu', _p(x), n, Cin, D, H, W, _p(weight), _p(scale), _p(shift), Cout, _p(skip),
               _p(y), _stream())
    return y


def prob_softargmin(x, weight, bias, depth_start, depth_end, want_reg=False):
    n, Cin, D, H, W = x.shape
    depth = torch.empty((n, H, W), dtype=torch.float32, device=x.device)
    reg = torch.empty((n, D, H, W), dtype=torch.float32, device=x.device) if want_reg else None
    lib().call('dv3d_prob_softargmin', _p(x), n, Cin, D, H, W, _p(weight), float(bias), float(depth_start),
               float(depth_end), _p(reg), _p(depth), _stream())
    return depth, reg


# ----------------------------------------------------------------------------- voxelise
def voxelize(pts, pts_batch, edge_len):
    """utils.py:38-64 -> anchor_pts [Nv,3] f32, anchor_idx3d [Nv,3] i32, anchor_batch [Nv] i64,
    point_anchor [N] i32 (row 0 of anchor_pts_edges), grid (VoxelGrid). Syncs twice."""
    _chk(pts, torch.float32, 'pts', 2), _chk(pts_batch, torch.int64, 'pts_batch', 1)
    N = pts.shape[0]
    if N == 0:
        raise RuntimeError('voxelize: empty point set (the reference fails on pts.min of an empty tensor)')
    dev = pts.device
    L = lib()
    grid = VoxelGrid()
    scratch = torch.empty(64, dtype=torch.uint8, device=dev)
    L.call('dv3d_voxel_grid', _p(pts), _p(pts_batch), N, float(edge_len), ctypes.byref(grid), _p(scratch), _stream())
    ws_bytes = L.raw('dv3d_voxelize_workspace_bytes')(ctypes.byref(grid), N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cap = N
    a_pts = torch.empty((cap, 3), dtype=torch.float32, device=dev)
    a_idx = torch.empty((cap, 3), dtype=torch.int32, device=dev)
    a_batch = torch.empty((cap,), dtype=torch.int64, device=dev)
    p_anchor = torch.empty((N,), dtype=torch.int32, device=dev)
    n_anchor = ctypes.c_longlong(0)
    L.call('dv3d_voxelize', _p(pts), _p(pts_batch), N, ctypes.byref(grid), _p(ws), ws_bytes, cap,
           ctypes.byref(n_anchor), _p(a_pts), _p(a_idx), _p(a_batch), _p(p_anchor), _stream())
    nv = n_anchor.value
    return a_pts[:nv], a_idx[:nv], a_batch[:nv], p_anchor, grid


# ----------------------------------------------------------------------------- PointNet
def pointnet_input(pts, pts_feat, anchor_pts, seg, out_ld=64):
    N, C = pts_feat.shape
    out = torch.empty((N, out_ld), dtype=torch.float32, device=pts.device)
    lib().call('dv3d_pointnet_input', _p(pts), _p(pts_feat), pts_feat.stride(0), _p(anchor_pts), _p(seg), N, C,
               out_ld, _p(out), _stream())
    return out


def linear(x_a, weight_kn, bias, relu_input, pool=None, seg=None, Ca=None, packed=None):
    N, lda = x_a.shape
    Ca = lda if Ca is None else Ca
    Cb = 0 if pool is None else pool.shape[1]
    Cout = weight_kn.shape[1]
    assert weight_kn.shape[0] == Ca + Cb
    y = torch.empty((N, Cout), dtype=torch.float32, device=x_a.device)
    lib().call('dv3d_linear', _p(x_a), Ca, lda, _p(pool), _p(seg), Cb, N, _p(weight_kn), _p(packed), _p(bias), Cout,
               int(relu_input), _p(y), _stream())
    return y


def segment_max(x, seg, n_seg):
    N, C = x.shape
    out = torch.empty((n_seg, C), dtype=torch.float32, device=x.device)
    lib().call('dv3d_segment_max', _p(x), _p(seg), N, C, n_seg, _p(out), _stream())
    return out


# ----------------------------------------------------------------------------- sparse levels
class SparseLevel(object):
    """One coordinate level of the sparse U-Net: coords [n,4] int32 (b,x,y,z) sorted by
    (b,z,y,x), tensor stride, hash table."""

    def __init__(self, coords, stride, err_flag):
        self.coords = coords
        self.n = coords.shape[0]
        self.stride = stride
        nbytes = lib().raw('dv3d_hash_bytes')(self.n)
        self.table = torch.empty(nbytes, dtype=torch.uint8, device=coords.device)
        self.table_bytes = nbytes
        lib().call('dv3d_hash_build', _p(coords), self.n, _p(self.table), nbytes, _p(err_flag), _stream())
        self._maps = {}

    def kernel_map(self, in_level, step):
        """KernelMap of this level's rows onto in_level at coords + offset_k * step."""
        key = (id(in_level), step)
        m = self._maps.get(key)
        if m is None:
            nbr = torch.empty((self.n, 27), dtype=torch.int32, device=self.coords.device)
            lib().call('dv3d_kernel_map', _p(self.coords), self.n, _p(in_level.table), in_level.table_bytes, step,
                       _p(nbr), _stream())
            m = self._maps[key] = KernelMap(nbr)
        return m


class KernelMap(object):
    """nbr [n_out,27] int32 (-1 = absent) plus, for the tensor-core path, the pair-major plan of
    csrc/sparse_pairs.cu and the library's choice between the two sparse-convolution variants."""

    def __init__(self, nbr):
        self.nbr = nbr
        self.n_out = nbr.shape[0]
        self.plan = None
        self.n_tiles = self.n_pairs = 0
        self.use_pairs = False
        self._P = None

    def build_plan(self):
        build_plans([self])
        return self

    def workspace(self, Cout):
        need = lib().raw('dv3d_sparse_conv_pairs_workspace_bytes')(self.n_tiles, Cout)
        if self._P is None or self._P.numel() < need:
            self._P = torch.empty(max(need, 16), dtype=torch.uint8, device=self.nbr.device)
        return self._P


def build_plans(maps):
    """enqueue the plan kernels of several KernelMaps (two launches per 16 maps, no sync);
    finish_plans() reads the counts back"""
    if gemm_mode() == 'f32':
        return
    todo = [m for m in maps if m.plan is None]
    for i in range(0, len(todo), 16):
        grp = todo[i:i + 16]
        n = len(grp)
        sizes = [lib().raw('dv3d_pair_plan_bytes')(m.n_out) for m in grp]
        for m, nb in zip(grp, sizes):
            m.plan = torch.empty(nb, dtype=torch.uint8, device=m.nbr.device)
        lib().call('dv3d_pair_plan_build', (ctypes.c_void_p * n)(*[m.nbr.data_ptr() for m in grp]),
                   (ctypes.c_longlong * n)(*[m.n_out for m in grp]),
                   (ctypes.c_void_p * n)(*[m.plan.data_ptr() for m in grp]), (ctypes.c_size_t * n)(*sizes), n, _stream())


def finish_plans(maps):
    """one sync for the tile / pair counts of several KernelMaps"""
    maps = [m for m in maps if m.plan is not None]
    if not maps:
        return
    n = len(maps)
    ptrs = (ctypes.c_void_p * n)(*[m.plan.data_ptr() for m in maps])
    tiles = (ctypes.c_longlong * n)()
    pairs = (ctypes.c_longlong * n)()
    lib().call('dv3d_pair_plan_counts', ptrs, n, tiles, pairs, _stream())
    for i, m in enumerate(maps):
        m.n_tiles, m.n_pairs = int(tiles[i]), int(pairs[i])
        m.use_pairs = bool(lib().raw('dv3d_sparse_conv_prefers_pairs')(m.n_out, m.n_tiles))


def make_coords(idx3d, batch):
    n = idx3d.shape[0]
    coords = torch.empty((n, 4), dtype=torch.int32, device=idx3d.device)
    lib().call('dv3d_make_coords', _p(idx3d), _p(batch), n, _p(coords), _stream())
    return coords


def coarsen(level, dims, n_batch, err_flag):
    """Output level of a stride-2 convolution on ``level`` (A.4). Syncs once."""
    L = lib()
    ns = level.stride * 2
    ws_bytes = L.raw('dv3d_coarsen_workspace_bytes')(dims[0], dims[1], dims[2], n_batch, ns)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=level.coords.device)
    out = torch.empty((level.n, 4), dtype=torch.int32, device=level.coords.device)
    n_out = ctypes.c_longlong(0)
    L.call('dv3d_coarsen', _p(level.coords), level.n, ns, dims[0], dims[1], dims[2], n_batch, _p(ws), ws_bytes,
           level.n, _p(out), ctypes.byref(n_out), _stream())
    return SparseLevel(out[:n_out.value], ns, err_flag)


def sparse_conv_workspace(Cout, device):
    """zeroed K-split workspace for sparse_conv (any level, up to Cout output channels)"""
    if gemm_mode() == 'f32':
        return None
    return torch.zeros(lib().raw('dv3d_sparse_conv_workspace_bytes')(Cout), dtype=torch.uint8, device=device)


def sparse_conv(feat, nbr, W, gn_weight=None, gn_bias=None, residual=None, relu=False, packed=None, workspace=None,
                out=None):
    """nbr: [n_out,27] int32 tensor, or a KernelMap (which may route to the pair-major variant).
    out: optional contiguous [n_out, Cout] destination (e.g. this rank's rows of a symmetric buffer)"""
    n_in, Cin = feat.shape
    Cout = W.shape[2]
    assert W.shape[0] == 27 and W.shape[1] == Cin
    if isinstance(nbr, KernelMap):
        km, nbr = nbr, nbr.nbr
        if km.use_pairs and packed is not None:
            out = _out_rows(out, km.n_out, Cout, feat.device)
            P = km.workspace(Cout)
            lib().call('dv3d_sparse_conv_pairs', _p(feat), n_in, Cin, _p(km.plan), km.n_tiles, km.n_out, _p(packed), Cout,
                       _p(gn_weight), _p(gn_bias), _p(residual), int(relu), _p(P), P.numel(), _p(out), _stream())
            return out
    n_out = nbr.shape[0]
    out = _out_rows(out, n_out, Cout, feat.device)
    ws_bytes = 0 if workspace is None else workspace.numel()
    lib().call('dv3d_sparse_conv', _p(feat), n_in, Cin, _p(nbr), n_out, _p(W), _p(packed), Cout, _p(gn_weight),
               _p(gn_bias), _p(residual), int(relu), _p(workspace) if ws_bytes else None, ws_bytes, _p(out), _stream())
    return out


def _out_rows(out, n, C, device):
    if out is None:
        return torch.empty((n, C), dtype=torch.float32, device=device)
    if tuple(out.shape) != (n, C) or out.dtype != torch.float32 or not out.is_contiguous():
        raise RuntimeError('out must be a contiguous float32 [%d, %d] tensor, got %s' % (n, C, tuple(out.shape)))
    return out


def concat_linear_gn_relu(a, b, W, gn_weight, gn_bias, packed=None, out=None):
    n, Ca = a.shape
    Cb = b.shape[1]
    Cout = W.shape[1]
    out = _out_rows(out, n, Cout, a.device)
    lib().call('dv3d_concat_linear_gn_relu', _p(a), Ca, _p(b), Cb, n, _p(W), _p(packed), Cout, _p(gn_weight),
               _p(gn_bias), _p(out), _stream())
    return out


def batch_origin(anchor_pts, idx3d, batch, n_batch, res):
    origin = torch.zeros((n_batch, 3), dtype=torch.float32, device=anchor_pts.device)
    lib().call('dv3d_batch_origin', _p(anchor_pts), _p(idx3d), _p(batch), anchor_pts.shape[0], float(res), _p(origin),
               _stream())
    return origin


def level_points(level, origin, res):
    n = level.n
    dev = level.coords.device
    pts = torch.empty((n, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((n, 3), dtype=torch.int64, device=dev)
    batch = torch.empty((n,), dtype=torch.int64, device=dev)
    lib().call('dv3d_level_points', _p(level.coords), n, _p(origin), float(res), _p(pts), _p(idx), _p(batch),
               _stream())
    return pts, idx, batch


# ----------------------------------------------------------------------------- decoder
def sparse_interp(pts, pts_batch, n_hyp, level, origin, res, feat, out, out_off):
    n_pts = pts.shape[0]
    lib().call('dv3d_sparse_interp', _p(pts), _p(pts_batch), n_pts, n_hyp, out.shape[1], _p(origin), float(res),
               level.stride, _p(level.table), level.table_bytes, _p(feat), feat.shape[1], _p(out), out.shape[2],
               out_off, _stream())


def conv1d_bn_relu(x, weight_tkn, scale, shift, out=None, packed=None, workspace=None):
    """x [n_pts, 8, Cin] -> [n_pts, 8, Cout] (refinement.py:8-13). workspace: sparse_conv_workspace()
    buffer that lets a thin last round of tiles split its taps over CTAs."""
    n_pts, rows, ldx = x.shape
    Cin, Cout = weight_tkn.shape[1], weight_tkn.shape[2]
    if out is None:
        out = torch.empty((n_pts, rows, Cout), dtype=torch.float32, device=x.device)
    ws_bytes = 0 if workspace is None else workspace.numel()
    lib().call('dv3d_conv1d_bn_relu', _p(x), n_pts, rows, Cin, ldx, _p(weight_tkn), _p(packed), _p(scale), _p(shift),
               Cout, _p(out), out.shape[2], _p(workspace) if ws_bytes else None, ws_bytes, _stream())
    return out


def decoder_pack_weights(weight_tkn):
    """[3, Cin, 128] fp32 -> the tap-stationary image of csrc/decoder_fused.cu, or None when the fused kernel
    does not apply ('f32' verification mode, hidden width != 128, Cin % 16 != 0)"""
    if gemm_mode() == 'f32':
        return None
    _chk(weight_tkn, torch.float32, 'weight', 3)
    _, Cin, Cout = weight_tkn.shape
    nbytes = lib().raw('dv3d_decoder_pack_bytes')(Cin)
    if nbytes == 0 or Cout != 128:
        return None
    out = torch.empty(nbytes, dtype=torch.uint8, device=weight_tkn.device)
    lib().call('dv3d_decoder_pack_weights', _p(weight_tkn), Cin, Cout, _p(out), _stream())
    return out


def decoder_fused(x, Cin, packed3, scale3, shift3, head_weight, head_bias, offset, want_prob=False, depth_accum=None):
    """the three Conv1d+BN+ReLU layers, the head, softmax and expected offset in one kernel
    (refinement.py:42-44 + lightningmodel.py:238-241). x [n_pts, 8, ld] -> (offset [n_pts], prob [n_pts,7] | None)"""
    n_pts, rows, ldx = x.shape
    off = torch.empty((n_pts,), dtype=torch.float32, device=x.device)
    prob = torch.empty((n_pts, 7), dtype=torch.float32, device=x.device) if want_prob else None
    arr = ctypes.c_void_p * 3
    lib().call('dv3d_decoder_fused', _p(x), n_pts, rows, Cin, ldx, arr(*[t.data_ptr() for t in packed3]),
               arr(*[t.data_ptr() for t in scale3]), arr(*[t.data_ptr() for t in shift3]), 128, _p(head_weight),
               float(head_bias), float(offset), 1 if gemm_mode() == 'tf32x3' else 2, _p(prob), _p(off), _p(depth_accum),
               _stream())
    return off, prob


def decoder_head(x, n_hyp, weight, bias, offset, want_prob=False):
    n_pts, rows, ldx = x.shape
    off = torch.empty((n_pts,), dtype=torch.float32, device=x.device)
    prob = torch.empty((n_pts, n_hyp), dtype=torch.float32, device=x.device) if want_prob else None
    lib().call('dv3d_decoder_head', _p(x), n_pts, n_hyp, rows, weight.shape[1], ldx, _p(weight), float(bias),
               float(offset), _p(prob), _p(off), _stream())
    return off, prob


# ----------------------------------------------------------------------------- upsampling
_grid_maps = {}


def grid_map_2d(n, H, W, device):
    """[n*H*W, 9] int32 neighbour rows of a dense image grid (-1 = zero padding), cached per shape"""
    key = (n, H, W, str(device))
    m = _grid_maps.get(key)
    if m is None:
        if len(_grid_maps) > 16:
            _grid_maps.clear()
        m = torch.empty((n * H * W, 9), dtype=torch.int32, device=device)
        lib().call('dv3d_grid_map_2d', n, H, W, _p(m), _stream())
        _grid_maps[key] = m
    return m


def propagation_net(features, depth_lo, layers):
    """PropagationNet (upsampling.py:14-36) on depth_lo [n,h,w] nearest-upsampled to the feature
    resolution (eval-3dvnet.py:101-125). features [n,C,H,W]; layers: four (W_kn [9*Cin,64], packed,
    scale [64], shift [64]) -> [n,H,W]."""
    _chk(features, torch.float32, 'features', 4), _chk(depth_lo, torch.float32, 'depth', 3)
    n, C, H, W = features.shape
    h, w = depth_lo.shape[1:]
    dev = features.device
    M = n * H * W
    ld0 = layers[0][0].shape[0] // 9
    x = torch.empty((M, ld0), dtype=torch.float32, device=dev)
    depth_up = torch.empty((n, H, W), dtype=torch.float32, device=dev)
    lib().call('dv3d_propagation_input', _p(features), C, _p(depth_lo), n, h, w, H, W, ld0, _p(x), _p(depth_up), _stream())
    nbr = grid_map_2d(n, H, W, dev)
    a = torch.empty((M, 64), dtype=torch.float32, device=dev)
    b = torch.empty((M, 64), dtype=torch.float32, device=dev)
    src, ldx = x, ld0
    for i, (W_kn, packed, scale, shift) in enumerate(layers):
        dst = a if i % 2 == 0 else b
        lib().call('dv3d_conv2d3x3_bn_relu_rows', _p(src), M, W_kn.shape[0] // 9, ldx, _p(nbr), _p(W_kn), _p(packed),
                   _p(scale), _p(shift), 64, _p(dst), 64, _stream())
        src, ldx = dst, 64
    out = torch.empty((n, H, W), dtype=torch.float32, device=dev)
    lib().call('dv3d_propagation_output', _p(src), 64, _p(depth_up), n, H, W, _p(out), _stream())
    return out


# ----------------------------------------------------------------------------- engine
class Conv3dParams(ctypes.Structure):
    """dv3d_conv3d_params_t"""
    _fields_ = [('weight', ctypes.c_void_p), ('scale', ctypes.c_void_p), ('shift', ctypes.c_void_p),
                ('Cin', ctypes.c_int), ('Cout', ctypes.c_int), ('kind', ctypes.c_int), ('reserved', ctypes.c_int)]


class DenseParams(ctypes.Structure):
    """dv3d_dense_params_t"""
    _fields_ = [('W', ctypes.c_void_p), ('Wp', ctypes.c_void_p), ('a', ctypes.c_void_p), ('b', ctypes.c_void_p),
                ('K', ctypes.c_int), ('N', ctypes.c_int)]


MAX_LEVELS, MAX_RES = 3, 4


class NetParams(ctypes.Structure):
    """dv3d_net_params_t"""
    _fields_ = [('costreg', Conv3dParams * 10), ('prob_weight', ctypes.c_void_p), ('prob_bias', ctypes.c_float),
                ('pointnet_in_pad', ctypes.c_int), ('pointnet', DenseParams * 6),
                ('n_levels', ctypes.c_int), ('n_res', ctypes.c_int * MAX_LEVELS),
                ('res_down', ((DenseParams * 2) * MAX_RES) * MAX_LEVELS),
                ('down', DenseParams * (MAX_LEVELS - 1)), ('up', DenseParams * (MAX_LEVELS - 1)),
                ('feat_adj', DenseParams * (MAX_LEVELS - 1)),
                ('res_up', ((DenseParams * 2) * MAX_RES) * (MAX_LEVELS - 1)),
                ('dec', DenseParams * 3), ('dec_head_weight', ctypes.c_void_p), ('dec_head_bias', ctypes.c_float),
                ('dec_fused', ctypes.c_void_p * 3)]


def dense_params(W, Wp, a, b):
    """W [K,N] fp32 (kept alive by the caller), Wp packed image or None, a / b per-channel vectors or None"""
    d = DenseParams()
    d.W, d.Wp = W.data_ptr(), (Wp.data_ptr() if Wp is not None else None)
    d.a, d.b = (a.data_ptr() if a is not None else None), (b.data_ptr() if b is not None else None)
    d.K, d.N = W.shape[0], W.shape[1]
    return d


_arena = {}


def hot_path_engine(net_params, feats_nhwc, rotmats, tvecs, K, plan, depth_batch, depth_cfg, img_size, edge_len,
                    offsets_list, want_init=False):
    """The whole hot path from one native call (csrc/engine.cu) -> depth [n_ref,h,w] (and the
    initial soft-argmin depth when want_init)."""
    _chk(feats_nhwc, torch.float32, 'feats_nhwc', 4), _chk(depth_batch, torch.int64, 'depth_batch', 1)
    _chk(rotmats, torch.float32, 'rotmats', 3), _chk(tvecs, torch.float32, 'tvecs', 2), _chk(K, torch.float32, 'K', 3)
    n_imgs, Hf, Wf, C = feats_nhwc.shape
    h, w = depth_cfg['size']
    D = int(depth_cfg['n_intervals'])
    dev = feats_nhwc.device
    L = lib()
    need = L.raw('dv3d_hot_path_workspace_bytes')(ctypes.byref(net_params), n_imgs, plan.n_ref, D, h, w)
    akey = (dev, torch.cuda.current_stream().cuda_stream)   # concurrent calls on different streams: one arena each
    arena = _arena.get(akey)
    if arena is None or arena.numel() < need:
        arena = _arena[akey] = torch.empty(need, dtype=torch.uint8, device=dev)
    n_outer = len(offsets_list)
    n_inner = len(offsets_list[0]) if n_outer else 0
    if any(len(o) != n_inner for o in offsets_list):
        raise RuntimeError('hot_path: every refinement iteration must have the same number of PointFlow passes')
    offs = (ctypes.c_double * max(1, n_outer * n_inner))(*[float(v) for o in offsets_list for v in o])
    depth = torch.empty((plan.n_ref, h, w), dtype=torch.float32, device=dev)
    init = torch.empty_like(depth) if want_init else None
    L.call('dv3d_hot_path', ctypes.byref(net_params), _p(feats_nhwc), n_imgs, Hf, Wf, _p(rotmats), _p(tvecs), _p(K),
           _p(plan.ref_img), _p(plan.rowptr), _p(plan.edge_src), plan.n_ref, _p(depth_batch),
           float(depth_cfg['depth_start']), float(depth_cfg['depth_interval']), D, h, w, img_size[0], img_size[1],
           float(edge_len), offs, n_outer, n_inner, _p(arena), arena.numel(), _p(init), _p(depth), _stream())
    return (depth, init) if want_init else depth


def hot_path_engine_sharded(net_params, feats_nhwc, rotmats, tvecs, K, plan_local, ref_start, depth_batch_all, depth_cfg,
                            img_size, edge_len, offsets_list, heap, want_init=False):
    """dv3d_hot_path_sharded (csrc/engine.cu): this rank's share of a scene spanning the ranks of `heap`
    (parallel.SymmHeap). `plan_local` is the EdgePlan of the rank's contiguous range of reference views (None when
    the rank has none), `ref_start` its first view among the sorted references, `depth_batch_all` [n_ref_total].
    -> depth [n_local,h,w] (and the initial soft-argmin depth when want_init)."""
    _chk(feats_nhwc, torch.float32, 'feats_nhwc', 4), _chk(depth_batch_all, torch.int64, 'depth_batch_all', 1)
    _chk(rotmats, torch.float32, 'rotmats', 3), _chk(tvecs, torch.float32, 'tvecs', 2), _chk(K, torch.float32, 'K', 3)
    n_imgs, Hf, Wf, C = feats_nhwc.shape
    h, w = depth_cfg['size']
    D = int(depth_cfg['n_intervals'])
    dev = feats_nhwc.device
    L = lib()
    n_total = int(depth_batch_all.shape[0])
    n_local = plan_local.n_ref if plan_local is not None else 0
    need = L.raw('dv3d_hot_path_workspace_bytes')(ctypes.byref(net_params), n_imgs, n_total, D, h, w)
    akey = (dev, torch.cuda.current_stream().cuda_stream)
    arena = _arena.get(akey)
    if arena is None or arena.numel() < need:
        arena = _arena[akey] = torch.empty(need, dtype=torch.uint8, device=dev)
    heap_need = L.raw('dv3d_hot_path_sharded_heap_bytes')(ctypes.byref(net_params), n_total, h, w)
    if heap.nbytes < heap_need:
        raise RuntimeError('hot_path_sharded: the symmetric heap has %d bytes, %d reference views of %dx%d need %d'
                           % (heap.nbytes, n_total, h, w, heap_need))
    n_outer = len(offsets_list)
    n_inner = len(offsets_list[0]) if n_outer else 0
    if any(len(o) != n_inner for o in offsets_list):
        raise RuntimeError('hot_path: every refinement iteration must have the same number of PointFlow passes')
    offs = (ctypes.c_double * max(1, n_outer * n_inner))(*[float(v) for o in offsets_list for v in o])
    depth = torch.empty((n_local, h, w), dtype=torch.float32, device=dev)
    init = torch.empty_like(depth) if want_init else None
    epoch = ctypes.c_int(heap.epoch)
    pl = plan_local
    try:
        L.call('dv3d_hot_path_sharded', ctypes.byref(net_params), _p(feats_nhwc), n_imgs, Hf, Wf, _p(rotmats), _p(tvecs),
               _p(K), _p(pl.ref_img) if pl else None, _p(pl.rowptr) if pl else None, _p(pl.edge_src) if pl else None,
               n_local, int(ref_start), n_total, _p(depth_batch_all), float(depth_cfg['depth_start']),
               float(depth_cfg['depth_interval']), D, h, w, img_size[0], img_size[1], float(edge_len), offs, n_outer,
               n_inner, _p(arena), arena.numel(), heap.rank, heap.world, heap.ptr, heap.nbytes, heap._peer_ptrs,
               ctypes.byref(epoch), heap.err.data_ptr(), _p(init), _p(depth), _stream())
    finally:
        heap.epoch = epoch.value
    return (depth, init) if want_init else depth


(STAGE_PLANESWEEP, STAGE_COSTREG, STAGE_SOFTARGMIN, STAGE_POINTCLOUD, STAGE_VOXELIZE, STAGE_POINTNET, STAGE_LEVELS,
 STAGE_UNET, STAGE_FLOW_WARP, STAGE_FLOW_INTERP, STAGE_DEC_GEMM0, STAGE_DEC_REST, STAGE_EXCHANGE,
 STAGE_BARRIER) = range(14)
STAGE_NAMES = ('planesweep_var', 'costreg', 'softargmin', 'pointcloud', 'voxelize', 'pointnet', 'levels', 'unet',
               'flow_warp', 'flow_interp', 'dec_gemm0', 'dec_rest', 'exchange', 'barrier')


def engine_profile(enable):
    """record CUDA events around the stages of dv3d_hot_path (clears earlier records)"""
    lib().call('dv3d_engine_profile', int(bool(enable)))


def engine_profile_read(cap=1 << 16):
    """[(stage id, ms), ...] in launch order; synchronise the stream first"""
    ids = (ctypes.c_int * cap)()
    ms = (ctypes.c_float * cap)()
    n = lib().raw('dv3d_engine_profile_read')(ids, ms, cap)
    if n < 0:
        raise Dv3dError(n, 'dv3d_engine_profile_read', lib().last_error())
    return [(ids[i], ms[i]) for i in range(n)]
