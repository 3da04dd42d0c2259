"""ORACLE (test infrastructure): PropagationNet and the upsampling cascade on the CPU, restated from
/root/reference/mv3d/subnetworks/upsampling.py:14-36 and /root/reference/mv3d/eval-3dvnet.py:101-125."""
import torch
import torch.nn.functional as F


def _conv_bn_relu(x, p, name):
    x = F.conv2d(x, p[name + '.0.weight'], padding=1)                                  # upsampling.py:6-11
    x = F.batch_norm(x, p[name + '.1.running_mean'], p[name + '.1.running_var'], p[name + '.1.weight'],
                     p[name + '.1.bias'], training=False, eps=1e-5)
    return F.relu(x)


def propagation_net(features, depth, p):
    """features [b,C,h,w], depth [b,1,h,w] -> [b,h,w] (upsampling.py:23-36)"""
    x = torch.cat((features, depth), dim=1)
    for name in ('conv1', 'conv2', 'conv3', 'conv4'):
        x = _conv_bn_relu(x, p, name)
    prob = F.softmax(x, dim=1)
    b, c, h, w = prob.shape
    unfold = F.unfold(F.pad(depth, (1, 1, 1, 1), mode='replicate'), kernel_size=3)     # [b, 9, h*w]
    return torch.sum(prob.view(b, c, h * w) * unfold, dim=1).view(b, h, w)


def upsample_cascade(depth, feats_quarter, feats_half, images, p_quarter, p_half, p_full):
    """eval-3dvnet.py:101-125: nearest upsampling + PropagationNet at 1/4, 1/2 and full resolution"""
    d = F.interpolate(depth.unsqueeze(1), feats_quarter.shape[-2:], mode='nearest')
    d = propagation_net(feats_quarter, d, p_quarter)
    d = F.interpolate(d.unsqueeze(1), feats_half.shape[-2:], mode='nearest')
    d = propagation_net(feats_half, d, p_half)
    d = F.interpolate(d.unsqueeze(1), images.shape[-2:], mode='nearest')
    return propagation_net(images, d, p_full)
