"""Coarse-to-fine depth upsampling — same interface as
/root/reference/mv3d/subnetworks/upsampling.py:14-36. Runs after the hot path (SURVEY.md §8f
"next"); plain torch/cuDNN."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv_bn_relu(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class PropagationNet(nn.Module):
    def __init__(self, in_dim=4, h_dim=32):
        super().__init__()
        self.conv1 = _conv_bn_relu(in_dim, h_dim)
        self.conv2 = _conv_bn_relu(h_dim, h_dim)
        self.conv3 = _conv_bn_relu(h_dim, h_dim)
        self.conv4 = _conv_bn_relu(h_dim, 9)
        self.unfold = nn.Unfold(kernel_size=3, stride=1, padding=0)

    def forward(self, features, depth):
        """features [b,in_dim-1,h,w], depth [b,1,h,w] -> [b,h,w]: per-pixel softmax over the 3x3
        neighbourhood, weighted sum of the (replicate-padded) depths."""
        b, _, h, w = depth.shape
        x = self.conv4(self.conv3(self.conv2(self.conv1(torch.cat((features, depth), dim=1)))))
        prob = F.softmax(x, dim=1)
        nbrs = self.unfold(F.pad(depth, (1, 1, 1, 1), mode='replicate')).view(b, 9, h, w)
        return torch.sum(nbrs * prob, dim=1)
