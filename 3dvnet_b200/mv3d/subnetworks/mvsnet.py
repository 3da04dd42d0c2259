"""MVSNet-style initial depth prediction — drop-in for
/root/reference/mv3d/subnetworks/mvsnet.py (class names, signatures and state_dict keys kept:
SURVEY.md Appendix C). Plane-sweep warp + variance, CostRegNet and soft-argmin run as
hand-written sm_100a kernels (csrc/planesweep.cu, csrc/costreg.cu); the 2D backbone stays a
torchvision/cuDNN module (outside the hot path)."""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

from ... import ops
from .. import functional, utils
from .._pack import PackCache, fold_bn, require_eval


def _autograd_path(module, *tensors):
    """True when the call must be differentiable (training mode, or an input that carries a gradient): the 3D
    convolutions then run through torch / cuDNN so that autograd sees them (library work, like the 2D backbone);
    the warp + variance in front of them stays ours in both directions (mv3d/functional.py). Inference never
    takes this path."""
    return module.training or (torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors))


class ConvBnRelu3d(nn.Module):
    """mvsnet.py:18-25 — parameters live in .conv / .bn like the reference."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        assert kernel_size == 3 and pad == 1 and stride in (1, 2)
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm3d(out_channels)
        self.stride = stride
        self._pack = PackCache()

    def forward(self, x, skip=None):
        if _autograd_path(self, x, skip):
            y = F.relu(self.bn(self.conv(x)))           # mvsnet.py:24-25
            return y if skip is None else skip + y      # mvsnet.py:159-161
        scale, shift = self._pack.get([self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var],
                                      lambda: fold_bn(self.bn))
        return ops.conv3d_bn_relu(x, self.conv.weight.detach(), scale, shift, self.stride, skip)


class DeconvBnRelu3d(nn.Module):
    """mvsnet.py:28-36 (stride 2, padding 1, output_padding 1)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.deconv = nn.ConvTranspose3d(in_channels, out_channels, 3, padding=1, output_padding=1, stride=2,
                                         bias=False)
        self.bn = nn.BatchNorm3d(out_channels)
        self._pack = PackCache()

    def forward(self, x, skip=None):
        if _autograd_path(self, x, skip):
            y = F.relu(self.bn(self.deconv(x)))         # mvsnet.py:35-36
            return y if skip is None else skip + y
        scale, shift = self._pack.get([self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var],
                                      lambda: fold_bn(self.bn))
        return ops.deconv3d_bn_relu(x, self.deconv.weight.detach(), scale, shift, skip)


class CostRegNet(nn.Module):
    """mvsnet.py:133-163: 3-level 3D U-Net on the variance volume."""

    def __init__(self, in_channels, base_channels):
        super().__init__()
        b = base_channels
        self.conv0 = ConvBnRelu3d(in_channels, b)
        self.conv1 = ConvBnRelu3d(b, b * 2, stride=2)
        self.conv2 = ConvBnRelu3d(b * 2, b * 2)
        self.conv3 = ConvBnRelu3d(b * 2, b * 4, stride=2)
        self.conv4 = ConvBnRelu3d(b * 4, b * 4)
        self.conv5 = ConvBnRelu3d(b * 4, b * 8, stride=2)
        self.conv6 = ConvBnRelu3d(b * 8, b * 8)
        self.conv7 = DeconvBnRelu3d(b * 8, b * 4)
        self.conv8 = DeconvBnRelu3d(b * 4, b * 2)
        self.conv9 = DeconvBnRelu3d(b * 2, b)
        self.prob = nn.Conv3d(b, 1, 3, stride=1, padding=1)
        self._pack = PackCache()

    def features(self, x):
        """everything up to (not including) the prob convolution: [n,8,D,h,w]"""
        if x.shape[2] % 8 or x.shape[3] % 8 or x.shape[4] % 8:
            raise RuntimeError('CostRegNet: D, h, w must be multiples of 8 (three stride-2 levels), got %s'
                               % (tuple(x.shape[2:]),))
        conv0 = self.conv0(x)
        conv2 = self.conv2(self.conv1(conv0))
        conv4 = self.conv4(self.conv3(conv2))
        x = self.conv6(self.conv5(conv4))
        x = self.conv7(x, skip=conv4)
        x = self.conv8(x, skip=conv2)
        return self.conv9(x, skip=conv0)

    def _prob_bias(self):
        return self._pack.get([self.prob.bias], lambda: float(self.prob.bias.detach().float().cpu()))

    def forward(self, x):
        """[n,C,D,h,w] -> x_reg [n,1,D,h,w]"""
        f = self.features(x)
        if _autograd_path(self, f):
            return self.prob(f)
        _, reg = ops.prob_softargmin(f, self.prob.weight.detach(), self._prob_bias(), 0.0, 1.0, want_reg=True)
        return reg.unsqueeze(1)

    def depth(self, x, depth_start, depth_end, want_reg=False):
        """fused prob conv + softmax(-x) + expectation (mvsnet.py:219-227)"""
        f = self.features(x)
        if _autograd_path(self, f):
            x_reg = self.prob(f).squeeze(1)
            prob = F.softmax(-x_reg, dim=1)
            vals = torch.linspace(depth_start, depth_end, x_reg.shape[1]).type_as(x_reg).view(1, -1, 1, 1)
            depth = torch.sum(vals * prob, dim=1)
            return depth, (x_reg if want_reg else None)
        return ops.prob_softargmin(f, self.prob.weight.detach(), self._prob_bias(), depth_start, depth_end, want_reg)


class FeatureExtractor(nn.Module):
    """MnasNet-1.0 trunk split at the five FPN taps (mvsnet.py:55-80). Weights are random
    here (no network access); a reference checkpoint loads by name."""

    def __init__(self):
        super().__init__()
        layers = list(torchvision.models.mnasnet1_0(weights=None).layers.children())
        self.layer1 = nn.Sequential(*layers[0:8])
        self.layer2 = nn.Sequential(*layers[8:9])
        self.layer3 = nn.Sequential(*layers[9:10])
        self.layer4 = nn.Sequential(*layers[10:12])
        self.layer5 = nn.Sequential(*layers[12:14])

    def forward(self, image):
        l1 = self.layer1(image)
        l2 = self.layer2(l1)
        l3 = self.layer3(l2)
        l4 = self.layer4(l3)
        return l1, l2, l3, l4, self.layer5(l4)

    def train(self, mode=True):
        super().train(mode)
        self.apply(utils.freeze_batchnorm)  # backbone BN always frozen (mvsnet.py:75-80)
        return self


class FeatureShrinker(nn.Module):
    """FPN to feat_dim channels at 1/2 .. 1/32 resolution (mvsnet.py:83-105)."""

    def __init__(self, feat_dim):
        super().__init__()
        self.fpn = torchvision.ops.FeaturePyramidNetwork([16, 24, 40, 96, 320], feat_dim, extra_blocks=None)

    def forward(self, *layers):
        out = self.fpn(OrderedDict(('layer%d' % (i + 1), l) for i, l in enumerate(layers)))
        return tuple(out['layer%d' % i] for i in range(1, 6))


class MVSNet(nn.Module):
    def __init__(self, feat_dim=32, img_size=(240, 320)):
        super().__init__()
        self.feat_dim = feat_dim
        self.img_size = img_size
        self.feat_extractor = FeatureExtractor()
        self.feat_shrinker = FeatureShrinker(feat_dim)
        self.cnn_3d = CostRegNet(feat_dim, 8)

    def cost_volume(self, features_quarter, batch, depth_start, depth_interval, n_planes, depth_img_size,
                    feats_nhwc=None, plan=None):
        """x_var [n_ref,C,D,h,w] (mvsnet.py:187-216) without materialising x_vox. Differentiable w.r.t. the feature
        maps (autograd.Function of mv3d/functional.py) whenever they carry a gradient."""
        plan = ops.edge_plan(batch.ref_src_edges, features_quarter.device) if plan is None else plan
        if torch.is_grad_enabled() and features_quarter.requires_grad:
            return functional.PlaneSweepVariance.apply(features_quarter, batch.rotmats, batch.tvecs, batch.K, plan,
                                                       depth_start, depth_interval, n_planes, tuple(depth_img_size),
                                                       tuple(self.img_size))
        if feats_nhwc is None:
            fq = features_quarter.detach().float()
            if fq.is_contiguous(memory_format=torch.channels_last) and not fq.is_contiguous():
                feats_nhwc = fq.permute(0, 2, 3, 1)
            else:
                feats_nhwc = ops.nchw_to_nhwc(fq.contiguous())
        cams = ops.camera_tables(batch.rotmats.float().contiguous(), batch.tvecs.float().contiguous(),
                                 batch.K.float().contiguous())
        return ops.planesweep_var(feats_nhwc, cams, plan, depth_start, depth_interval, n_planes,
                                  tuple(depth_img_size), tuple(self.img_size))

    def depth_from_features(self, features_quarter, batch, depth_start, depth_interval, n_planes, depth_img_size,
                            feats_nhwc=None, plan=None):
        x_var = self.cost_volume(features_quarter, batch, depth_start, depth_interval, n_planes, depth_img_size,
                                 feats_nhwc, plan)
        depth_end = depth_start + depth_interval * (n_planes - 1)
        depth, _ = self.cnn_3d.depth(x_var, depth_start, depth_end)
        return depth

    def forward(self, batch, depth_start, depth_interval, n_planes, depth_img_size):
        """-> depth_img [n_ref,h,w], features_half, features_quarter, features_eighth (mvsnet.py:176-229).
        Works in training mode too: warp + variance forward/backward are the CUDA kernels of csrc/planesweep.cu,
        the 2D backbone and the 3D regulariser train through torch / cuDNN."""
        # channels-last through the cuDNN backbone: the FPN then emits NHWC feature maps, which the warp
        # kernels consume without a transposition pass
        images = batch.images.contiguous(memory_format=torch.channels_last)
        fh, fq, fe, _, _ = self.feat_shrinker(*self.feat_extractor(images))
        depth = self.depth_from_features(fq, batch, depth_start, depth_interval, n_planes, depth_img_size)
        return depth, fh, fq, fe
