"""PL3DVNet — drop-in for /root/reference/mv3d/lightningmodel.py's inference interface:
constructor, ``hparams``, the four methods the eval driver calls
(/root/reference/mv3d/eval-3dvnet.py:58-59,75-76,88-98) and the sub-module names of the
checkpoint schema. pytorch_lightning is not required: this is a plain nn.Module.

The reference's inline back-projection / re-projection / variance code
(lightningmodel.py:132-174,187-235) is replaced by csrc/planesweep.cu:points_var_kernel."""
from argparse import Namespace

import torch
import torch.nn as nn

from .. import ops
from . import functional
from ._pack import PackCache, fold_bn, require_eval
from .subnetworks.mvsnet import MVSNet
from .subnetworks.scenemodeling import PointNet, SparseUNet, SparseScene
from .subnetworks.refinement import HypothesisDecoder
from .subnetworks.upsampling import PropagationNet


class _FeatureCache(object):
    """channels-last copies of feature tensors, keyed on storage + version. Every entry keeps its
    SOURCE tensor alive: the caching allocator cannot hand that address to the next scene's
    same-shape feature map while the entry exists, so a key can never match a different tensor."""

    def __init__(self, size=4):
        self.size, self.items = size, []

    def get(self, feats):
        key = (feats.data_ptr(), feats._version, tuple(feats.shape), feats.dtype)
        for k, src, v in self.items:
            if k == key and src is feats:
                return v
        v = ops.nchw_to_nhwc(feats.detach().float().contiguous())
        self.items = ([(key, feats, v)] + self.items)[:self.size]
        return v


class PL3DVNet(nn.Module):
    def __init__(self, depth_train, depth_test, edge_len, feat_dim=16, img_size=(256, 320), hyp_ksize=3, hyp_pad=1,
                 lr=1e-3, lr_step=100, lr_gamma=0.1, finetune=False):
        super().__init__()
        self.depth_train, self.depth_test, self.edge_len = depth_train, depth_test, edge_len
        self.feat_dim, self.img_size = feat_dim, img_size
        self.hyp_ksize, self.hyp_pad = hyp_ksize, hyp_pad
        self.lr, self.lr_step, self.lr_gamma, self.finetune = lr, lr_step, lr_gamma, finetune
        self.hparams = Namespace(depth_train=depth_train, depth_test=depth_test, edge_len=edge_len,
                                 feat_dim=feat_dim, img_size=img_size, hyp_ksize=hyp_ksize, hyp_pad=hyp_pad, lr=lr,
                                 lr_step=lr_step, lr_gamma=lr_gamma, finetune=finetune)
        if feat_dim != 32:
            raise NotImplementedError('the warp kernels are specialised for IMG_FEAT_DIM = 32 '
                                      '(/root/reference/mv3d/config.py:42), got %d' % feat_dim)
        self.mvsnet = MVSNet(feat_dim, img_size)
        self.pointnet = PointNet(4 * feat_dim, 2 * feat_dim, feat_dim + 3)
        self.sparse_conv = SparseUNet(dims=(2 * feat_dim, 128, 128), n_groups=(4, 8, 8), n_res=(1, 2, 3))
        self.decoder = HypothesisDecoder(128 + 128 + 3 * feat_dim, 128, hyp_ksize, hyp_pad)
        self.refine_quarter = PropagationNet(in_dim=feat_dim + 1, h_dim=32)
        self.refine_half = PropagationNet(in_dim=feat_dim + 1, h_dim=32)
        self.refine_full = PropagationNet(in_dim=3 + 1, h_dim=32)
        self._nhwc = _FeatureCache()
        self._engine_pack = PackCache()
        self._engine_tensors = None
        self._graphs = {}

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def load_from_checkpoint(cls, path, map_location=None, **overrides):
        """Lightning-style checkpoint: {'state_dict', 'hyper_parameters'} (lightningmodel.py:33)."""
        ckpt = torch.load(path, map_location=map_location or 'cpu', weights_only=False)
        hp = dict(ckpt.get('hyper_parameters', {}))
        hp.update(overrides)
        net = cls(**hp)
        # Lightning's loader is strict (lightningmodel.py:33 via LightningModule.load_from_checkpoint):
        # a checkpoint whose keys drift from the schema must not load with randomly initialised layers
        missing, unexpected = net.load_state_dict(ckpt['state_dict'], strict=False)
        missing = [k for k in missing if not k.endswith('num_batches_tracked')]
        if missing or unexpected:
            raise RuntimeError('load_from_checkpoint: state_dict does not match the reference schema '
                               '(missing %s, unexpected %s)' % (missing[:8], list(unexpected)[:8]))
        return net

    # ------------------------------------------------------------------ hot path A
    def make_initial_depth_predictions(self, batch, depth_config):
        """lightningmodel.py:124-130"""
        depth_pred, feats_half, feats_quarter, feats_eighth = self.mvsnet(
            batch, depth_config['depth_start'], depth_config['depth_interval'], depth_config['n_intervals'],
            depth_config['size'])
        ref_idx = ops.edge_plan(batch.ref_src_edges, depth_pred.device).ref_idx
        depth_batch = batch.images_batch.to(depth_pred.device)[ref_idx]
        return depth_pred, depth_batch, feats_half, feats_quarter, feats_eighth, ref_idx

    # ------------------------------------------------------------------ hot path B
    def _geometry(self, img_feats, rotmats, tvecs, K, ref_src_edges):
        dev = img_feats.device
        plan = ops.edge_plan(ref_src_edges, dev)
        R, t, Kc = rotmats.float().contiguous(), tvecs.float().contiguous(), K.float().contiguous()
        return plan, self._nhwc.get(img_feats), ops.camera_tables(R, t, Kc)

    def construct_feature_rich_pointcloud(self, depth_pred, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges):
        """-> pts [n_ref*P,3], pts_feat [n_ref*P,C], pts_batch [n_ref*P] (lightningmodel.py:132-174)"""
        depth = depth_pred.detach().float().contiguous()
        if torch.is_grad_enabled() and img_feats.requires_grad:
            # training: the variance features carry a gradient to the feature maps (mv3d/functional.py); the
            # points do not (the reference builds the re-projection under no_grad, lightningmodel.py:147)
            pts, feat = functional.point_variance(img_feats, rotmats, tvecs, K, ref_src_edges, depth,
                                                  self.hparams.img_size, 0, 0.0)
        else:
            plan, nhwc, cams = self._geometry(img_feats, rotmats, tvecs, K, ref_src_edges)
            pts, feat = ops.points_var(nhwc, cams, plan, depth, self.hparams.img_size, 0, 0.0)
        n, P = depth.shape[0], depth.shape[1] * depth.shape[2]
        pts_batch = depth_batch.unsqueeze(1).expand(n, P).reshape(-1)
        return pts.view(-1, 3), feat.view(-1, feat.shape[2]), pts_batch

    def model_scene(self, depth_pred, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges, return_pts=False):
        """lightningmodel.py:176-185"""
        require_eval(self)
        pts, pts_feat, pts_batch = self.construct_feature_rich_pointcloud(depth_pred, depth_batch, img_feats, rotmats,
                                                                          tvecs, K, ref_src_edges)
        xs = self.scene_from_points(pts, pts_feat, pts_batch.contiguous())
        return (xs, pts) if return_pts else xs

    def scene_from_points(self, pts, pts_feat, pts_batch):
        """voxelise -> PointNet -> sparse U-Net on a feature-rich point cloud (lightningmodel.py:180-184).
        Also the entry of the multi-GPU path after its all-gather (3dvnet_b200/parallel.py)."""
        require_eval(self)
        a_pts, a_idx, a_batch, seg, grid = ops.voxelize(pts, pts_batch, self.edge_len)
        x = ops.pointnet_input(pts, pts_feat, a_pts, seg, self.pointnet.in_pad)
        x = self.pointnet.forward_padded(x, seg, a_pts.shape[0])
        dims = (int(grid.n_cells[0]), int(grid.n_cells[1]), int(grid.n_cells[2]), int(grid.n_batch))
        scene = SparseScene(a_idx, a_batch, self.sparse_conv.n_levels, dims)
        return self.sparse_conv(x, a_pts, a_idx, a_batch, self.edge_len, scene=scene)

    def run_pointflow(self, xs, depth_pred, depth_batch, img_feats, rotmats, tvecs, K, ref_src_edges, offset, n,
                      return_prob=False):
        """expected depth offset [n_ref,h,w] over 2n+1 hypotheses (lightningmodel.py:187-242)"""
        require_eval(self)
        plan, nhwc, cams = self._geometry(img_feats, rotmats, tvecs, K, ref_src_edges)
        depth = depth_pred.detach().float().contiguous()
        n_ref, h, w = depth.shape
        n_pts, n_hyp = n_ref * h * w, 2 * n + 1
        operand = self.decoder.operand(n_pts, depth.device)
        var_off = operand.shape[2] - self.hparams.feat_dim
        pts_hyp, _ = ops.points_var(nhwc, cams, plan, depth, self.hparams.img_size, n, offset,
                                    feat_out=operand, feat_off=var_off)
        pts_batch = depth_batch.unsqueeze(1).expand(n_ref, h * w).reshape(-1).contiguous()
        got = self.decoder.fill_levels(xs, pts_hyp, pts_batch, operand)
        assert got == var_off
        off, prob = self.decoder.run(operand, n_hyp, offset, want_prob=return_prob)
        off = off.view(n_ref, h, w)
        return (off, prob) if return_prob else off

    # ------------------------------------------------------------------ full inference pass
    def refine_depth(self, depth_pred, depth_batch, feats_quarter, rotmats, tvecs, K, ref_src_edges, offsets_list):
        """eval-3dvnet.py:73-99 without the python chunking"""
        depth = depth_pred.clone()
        for offsets in offsets_list:
            xs = self.model_scene(depth, depth_batch, feats_quarter, rotmats, tvecs, K, ref_src_edges)
            for offset in offsets:
                depth += self.run_pointflow(xs, depth, depth_batch, feats_quarter, rotmats, tvecs, K, ref_src_edges,
                                            offset, 3)
        return depth

    def engine_params(self):
        """dv3d_net_params_t for csrc/engine.cu: device pointers of the parameters and of their
        kernel-friendly derived copies (folded BN, transposed / packed weights), cached until a
        parameter changes. The tensors behind the pointers are kept alive in the returned holder."""
        def build():
            keep = []

            def dense(W, Wp, a, b):
                keep.extend((W, Wp, a, b))
                return ops.dense_params(W, Wp, a, b)

            P = ops.NetParams()
            c3 = self.mvsnet.cnn_3d
            layers = [c3.conv0, c3.conv1, c3.conv2, c3.conv3, c3.conv4, c3.conv5, c3.conv6, c3.conv7, c3.conv8, c3.conv9]
            for i, m in enumerate(layers):
                deconv = hasattr(m, 'deconv')
                wt = (m.deconv if deconv else m.conv).weight.detach().float().contiguous()
                scale, shift = (t.to(wt.device) for t in fold_bn(m.bn))
                keep.extend((wt, scale, shift))
                e = P.costreg[i]
                e.weight, e.scale, e.shift = wt.data_ptr(), scale.data_ptr(), shift.data_ptr()
                e.Cin, e.Cout = (wt.shape[0], wt.shape[1]) if deconv else (wt.shape[1], wt.shape[0])
                e.kind = 2 if deconv else (1 if m.stride == 2 else 0)
            pw = c3.prob.weight.detach().float().contiguous()
            keep.append(pw)
            P.prob_weight, P.prob_bias = pw.data_ptr(), float(c3.prob.bias.detach().float().cpu())
            pn = self.pointnet._weights()
            P.pointnet_in_pad = self.pointnet.in_pad
            for i, name in enumerate(('fc_pos', 'fc1', 'fc2', 'fc3', 'fc4', 'fc_out')):
                w, b, packed = pn[name]
                P.pointnet[i] = dense(w, packed, None, b)
            u = self.sparse_conv

            def sconv(conv, gn):
                k, packed = conv.weights()
                W = k.float().reshape(-1, k.shape[-1]).contiguous()
                return dense(W, packed, gn.weight.detach().float().contiguous(), gn.bias.detach().float().contiguous())

            nl = u.n_levels
            if nl > ops.MAX_LEVELS or max(len(seq) for seq in u.res_down) > ops.MAX_RES:
                raise NotImplementedError('engine supports up to %d levels x %d residual blocks' % (ops.MAX_LEVELS, ops.MAX_RES))
            P.n_levels = nl
            for l in range(nl):
                P.n_res[l] = len(u.res_down[l])
                for b, blk in enumerate(u.res_down[l]):
                    P.res_down[l][b][0] = sconv(blk.conv1, blk.n1.gn)
                    P.res_down[l][b][1] = sconv(blk.conv2, blk.n2.gn)
            for i in range(nl - 1):
                P.down[i] = sconv(u.down[i][0], u.down[i][1].gn)
                P.up[i] = sconv(u.up[i][0], u.up[i][1].gn)
                P.feat_adj[i] = sconv(u.feat_adj[i][0], u.feat_adj[i][1].gn)
                for b, blk in enumerate(u.res_up[i]):
                    P.res_up[i][b][0] = sconv(blk.conv1, blk.n1.gn)
                    P.res_up[i][b][1] = sconv(blk.conv2, blk.n2.gn)
            layers, head = self.decoder._weights()
            for i, (w, scale, shift, packed) in enumerate(layers):
                P.dec[i] = dense(w.reshape(-1, w.shape[2]), packed, scale, shift)
            keep.append(head[0])
            P.dec_head_weight, P.dec_head_bias = head[0].data_ptr(), head[1]
            if self.decoder._fused is not None:
                keep.extend(self.decoder._fused)
                for i, f in enumerate(self.decoder._fused):
                    P.dec_fused[i] = f.data_ptr()
            return P, keep
        # invalidation key: storage + version of the hot-path parameters only (the module tree is
        # static; walking all ~560 tensors incl. the 2D backbone costs more than a PointFlow pass)
        if self._engine_tensors is None:
            mods = (self.mvsnet.cnn_3d, self.pointnet, self.sparse_conv, self.decoder)
            self._engine_tensors = [t for m in mods for t in list(m.parameters()) + list(m.buffers())]
        return self._engine_pack.get(self._engine_tensors, build, ops.gemm_mode())[0]

    def hot_path(self, feats_quarter, rotmats, tvecs, K, ref_src_edges, images_batch, depth_config, offsets_list,
                 return_init=False):
        """One pass of the hot path from quarter-resolution features: plane-sweep cost volume ->
        CostRegNet -> soft-argmin, then the volumetric refinement schedule. -> depth [n_ref,h,w].
        This is the benchmark "step" (BASELINE.json configs[1]). The whole pass is enqueued by one
        native call (csrc/engine.cu); hot_path_composed is the same pass composed op by op."""
        require_eval(self)
        dev = feats_quarter.device
        plan = ops.edge_plan(ref_src_edges, dev)
        depth_batch = images_batch[plan.ref_idx].long().contiguous()
        # the channels-last copy is made on every call (one small kernel): the storage-keyed cache
        # of the composed path must not be trusted for freshly uploaded tensors
        fq = feats_quarter.detach().float()
        if fq.dim() == 4 and fq.is_contiguous(memory_format=torch.channels_last) and not fq.is_contiguous():
            nhwc = fq.permute(0, 2, 3, 1)   # a channels-last backbone already emits the layout the warp kernel wants
        else:
            nhwc = ops.nchw_to_nhwc(fq.contiguous())
        return ops.hot_path_engine(self.engine_params(), nhwc, rotmats.float().contiguous(),
                                   tvecs.float().contiguous(), K.float().contiguous(), plan, depth_batch, depth_config,
                                   self.hparams.img_size, self.edge_len, offsets_list, want_init=return_init)

    def hot_path_composed(self, feats_quarter, rotmats, tvecs, K, ref_src_edges, images_batch, depth_config,
                          offsets_list):
        """hot_path through the reference-shaped modules, one C-ABI call per op (cross-check of the engine)."""
        require_eval(self)
        with torch.no_grad():
            dev = feats_quarter.device
            plan = ops.edge_plan(ref_src_edges, dev)
            batch = Namespace(rotmats=rotmats, tvecs=tvecs, K=K, ref_src_edges=plan)
            depth = self.mvsnet.depth_from_features(
                feats_quarter, batch, depth_config['depth_start'], depth_config['depth_interval'],
                depth_config['n_intervals'], depth_config['size'], feats_nhwc=self._nhwc.get(feats_quarter), plan=plan)
            depth_batch = images_batch[plan.ref_idx]
            return self.refine_depth(depth, depth_batch, feats_quarter, rotmats, tvecs, K, plan, offsets_list)

    def upsample(self, depth, ref_idx, feats_quarter, feats_half, images):
        """nearest upsampling + the three PropagationNets (lightningmodel.py:84-112,
        eval-3dvnet.py:101-125): [n_ref,h,w] -> [n_ref,H,W]; the nearest upsampling is fused into
        each PropagationNet's input kernel"""
        d = self.refine_quarter.forward_from(feats_quarter[ref_idx], depth)
        d = self.refine_half.forward_from(feats_half[ref_idx], d)
        return self.refine_full.forward_from(images[ref_idx], d)

    def _backbone(self, images):
        """feats_half, feats_quarter of the 2D backbone + FPN (mvsnet.py:183-185), channels-last. The ~200 cuDNN /
        elementwise launches of MnasNet are host-bound in eager mode (2.5 ms for 8 images), so they are captured once
        per input shape into a CUDA graph and replayed (DV3D_BACKBONE_GRAPH=0: eager). The returned tensors are the
        graph's static outputs: valid until the next call with the same shape."""
        import os
        m = self.mvsnet

        def run(x):
            fh, fq, _, _, _ = m.feat_shrinker(*m.feat_extractor(x))
            return fh, fq
        if os.environ.get('DV3D_BACKBONE_GRAPH', '1') == '0' or not images.is_cuda:
            return run(images.contiguous(memory_format=torch.channels_last))
        first = next(m.feat_extractor.parameters())
        key = (tuple(images.shape), images.dtype, images.device, first.data_ptr(), first._version)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) > 4:
                self._graphs.clear()
            static_in = torch.empty_like(images, memory_format=torch.channels_last)
            static_in.copy_(images)
            side = torch.cuda.Stream(device=images.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up outside capture: cuDNN algorithm selection, lazy init
                for _ in range(2):
                    run(static_in)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = run(static_in)
            g = self._graphs[key] = (graph, static_in, outs)
        graph, static_in, outs = g
        static_in.copy_(images)
        graph.replay()
        return outs

    def full_pass(self, images, rotmats, tvecs, K, ref_src_edges, images_batch, offsets_list, depth_config=None):
        """SURVEY.md section 8d's full pipeline of one batch, the loop body of process_scene
        (eval-3dvnet.py:58-125): 2D backbone + FPN (torchvision / cuDNN, channels-last so that the NHWC feature maps
        feed the warp kernels without a transposition pass) -> the hot path through the native engine (one C-ABI
        call) -> nearest upsampling + the three PropagationNets. -> dict(ref [n_ref,h,w], final [n_ref,H,W])."""
        require_eval(self)
        with torch.no_grad():
            cfg = self.hparams.depth_test if depth_config is None else depth_config
            fh, fq = self._backbone(images)
            plan = ops.edge_plan(ref_src_edges, images.device)
            depth = self.hot_path(fq, rotmats, tvecs, K, plan, images_batch, cfg, offsets_list)
            return {'ref': depth, 'final': self.upsample(depth, plan.ref_idx, fq, fh, images), 'feats_quarter': fq,
                    'feats_half': fh}

    def forward(self, batch, offsets, n_iters):
        """Inference-only counterpart of lightningmodel.py:48-122: returns the depth maps of every
        stage instead of losses (no ground truth is consumed)."""
        require_eval(self)
        with torch.no_grad():
            cfg = self.hparams.depth_test
            depth, depth_batch, fh, fq, _, ref_idx = self.make_initial_depth_predictions(batch, cfg)
            out = {'initial': depth}
            refined = self.refine_depth(depth, depth_batch, fq, batch.rotmats, batch.tvecs, batch.K,
                                        batch.ref_src_edges, [list(offsets)] * n_iters)
            out['ref'] = refined
            out['final'] = self.upsample(refined, ref_idx, fq, fh, batch.images)
            return out
