"""ORACLE (test infrastructure): CostRegNet + soft-argmin in eval mode, restated from
/root/reference/mv3d/subnetworks/mvsnet.py:18-36,133-163,219-227. `p` holds the
`cnn_3d.*` tensors under the reference's names. Inference only:
BatchNorm uses its running statistics (the eval driver calls net.eval(),
/root/reference/mv3d/eval/main.py:40)."""
import torch
import torch.nn.functional as F


def _bn_relu(x, p, name, eps=1e-5):
    x = F.batch_norm(x, p[name + '.running_mean'], p[name + '.running_var'], p[name + '.weight'],
                     p[name + '.bias'], training=False, eps=eps)
    return F.relu(x)


def conv_bn_relu3d(x, p, name, stride=1):
    """ConvBnRelu3d (mvsnet.py:18-25): k=3, pad=1, no bias."""
    return _bn_relu(F.conv3d(x, p[name + '.conv.weight'], None, stride, 1), p, name + '.bn')


def deconv_bn_relu3d(x, p, name):
    """DeconvBnRelu3d (mvsnet.py:28-36): k=3, stride=2, pad=1, output_padding=1, no bias."""
    return _bn_relu(F.conv_transpose3d(x, p[name + '.deconv.weight'], None, 2, 1, 1), p, name + '.bn')


def costregnet(x_var, p, return_all=False):
    """[n,32,D,h,w] -> [n,1,D,h,w] (mvsnet.py:154-163)."""
    c0 = conv_bn_relu3d(x_var, p, 'conv0')
    c2 = conv_bn_relu3d(conv_bn_relu3d(c0, p, 'conv1', 2), p, 'conv2')
    c4 = conv_bn_relu3d(conv_bn_relu3d(c2, p, 'conv3', 2), p, 'conv4')
    c6 = conv_bn_relu3d(conv_bn_relu3d(c4, p, 'conv5', 2), p, 'conv6')
    x7 = c4 + deconv_bn_relu3d(c6, p, 'conv7')
    x8 = c2 + deconv_bn_relu3d(x7, p, 'conv8')
    x9 = c0 + deconv_bn_relu3d(x8, p, 'conv9')
    out = F.conv3d(x9, p['prob.weight'], p['prob.bias'], 1, 1)
    if return_all:
        return out, dict(conv0=c0, conv2=c2, conv4=c4, conv6=c6, x7=x7, x8=x8, x9=x9)
    return out


def soft_argmin(x_reg, depth_start, depth_interval, n_planes):
    """softmax(-x_reg) over D, expectation of the plane depths (mvsnet.py:220-227)."""
    prob = F.softmax(-x_reg, dim=1)
    depth_vals = torch.linspace(depth_start, depth_start + depth_interval * (n_planes - 1), n_planes).type_as(x_reg)
    return torch.sum(depth_vals.view(1, -1, 1, 1) * prob, dim=1)
