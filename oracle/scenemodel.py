"""ORACLE (test infrastructure): PointNet, sparse 3D-UNet and PointFlow hypothesis decoder,
restated from /root/reference/mv3d/subnetworks/scenemodeling.py:16-44,78-237 and
/root/reference/mv3d/subnetworks/refinement.py:17-44 on top of oracle/minkowski_cpu.py
(MinkowskiEngine restatement — parity unpinned against ME, see that file's header).
`p` dictionaries use the reference's state_dict names."""
import numpy as np
import torch
import torch.nn.functional as F

from . import minkowski_cpu as mk


# --------------------------------------------------------------------------- PointNet
def segment_max(x, idx, n):
    """torch_scatter 'max' along dim 0: empty segments stay 0."""
    out = torch.zeros((n, x.shape[1]), dtype=x.dtype)
    return out.scatter_reduce_(0, idx[:, None].expand_as(x), x, 'amax', include_self=False)


def pointnet(x, idx, n_idx, p):
    """scenemodeling.py:127-144."""
    def lin(name, v):
        return F.linear(v, p[name + '.weight'], p[name + '.bias'])
    x = lin('fc1', F.relu(lin('fc_pos', x)))
    for name in ('fc2', 'fc3', 'fc4'):
        pool = segment_max(x, idx, n_idx)
        x = lin(name, F.relu(torch.cat((x, pool[idx]), dim=1)))
    return lin('fc_out', F.relu(segment_max(x, idx, n_idx)))


# --------------------------------------------------------------------------- sparse UNet
def group_norm_rows(x, p, name, n_groups, eps=1e-5):
    """torch.nn.GroupNorm applied to the [N,C] feature matrix (scenemodeling.py:91-104):
    every voxel is normalised on its own over each group of C/n_groups channels."""
    return F.group_norm(x, n_groups, p[name + '.gn.weight'], p[name + '.gn.bias'], eps)


def sparse_residual(x, cmap, p, name, n_groups):
    """relu(x + GN2(conv2(relu(GN1(conv1(x)))))) (scenemodeling.py:41-44)."""
    pairs = mk.kernel_map(cmap, cmap, cmap.stride)
    h = mk.sparse_conv(x, p[name + '.conv1.kernel'], pairs, len(cmap))
    h = F.relu(group_norm_rows(h, p, name + '.n1', n_groups))
    h = mk.sparse_conv(h, p[name + '.conv2.kernel'], pairs, len(cmap))
    h = group_norm_rows(h, p, name + '.n2', n_groups)
    return F.relu(h + x)


def sparse_unet(feat, pts, idx, batch, res, p, dims=(64, 128, 128), n_groups=(4, 8, 8), n_res=(1, 2, 3)):
    """scenemodeling.py:191-237 -> list (coarse -> fine) of dicts with feats / pts / res /
    batch / idx / stride / cmap. Rows of every level are in ascending (batch,z,y,x) order."""
    coords = np.concatenate([batch.numpy()[:, None], idx.numpy().astype(np.int64)], axis=1)
    order = np.argsort(mk.encode(coords), kind='stable')
    cmap = mk.CoordMap(coords[order], 1, presorted=True)
    x = feat[torch.from_numpy(order)]

    maps = [cmap]
    for l in range(n_res[0]):
        x = sparse_residual(x, cmap, p, 'res_down.0.%d' % l, n_groups[0])
    skips = [x]
    for i in range(1, len(dims)):
        x, cmap = mk.conv3(x, cmap, p['down.%d.0.kernel' % (i - 1)], 2)
        x = F.relu(group_norm_rows(x, p, 'down.%d.1' % (i - 1), n_groups[i]))
        for l in range(n_res[i]):
            x = sparse_residual(x, cmap, p, 'res_down.%d.%d' % (i, l), n_groups[i])
        maps.append(cmap)
        skips.append(x)

    rg = n_groups[::-1]
    rres = n_res[::-1]
    maps = maps[::-1]
    skips = skips[::-1]
    outs = [(skips[0], maps[0])]
    x = skips[0]
    for i in range(1, len(dims)):
        up = mk.conv3_transpose(x, maps[i - 1], maps[i], p['up.%d.0.kernel' % (i - 1)])
        up = F.relu(group_norm_rows(up, p, 'up.%d.1' % (i - 1), rg[i]))
        x = torch.cat((up, skips[i]), dim=1) @ p['feat_adj.%d.0.kernel' % (i - 1)]
        x = F.relu(group_norm_rows(x, p, 'feat_adj.%d.1' % (i - 1), rg[i]))
        for l in range(rres[i]):
            x = sparse_residual(x, maps[i], p, 'res_up.%d.%d' % (i - 1, l), rg[i])
        outs.append((x, maps[i]))

    # voxel positions of every level: idx * res + position of index (0,0,0) of that batch,
    # the latter recovered from the FIRST input voxel of the batch (scenemodeling.py:219-226)
    n_batches = int(batch.max()) + 1
    info = []
    for x, cm in outs:
        x_idx = torch.from_numpy(cm.coords[:, 1:])
        x_batch = torch.from_numpy(cm.coords[:, 0])
        x_pts = torch.empty((len(cm), 3), dtype=torch.float32)
        for b in range(n_batches):
            m_in = batch == b
            m_out = x_batch == b
            pts_min = pts[m_in][0] - (idx[m_in][0] * res)
            x_pts[m_out] = x_idx[m_out] * res + pts_min
        info.append(dict(feats=x, pts=x_pts, res=cm.stride * res, batch=x_batch, idx=x_idx, stride=cm.stride,
                         cmap=cm))
    return info


# --------------------------------------------------------------------------- decoder
def interpolate_levels(xs, pts, pts_feat, pts_batch):
    """Trilinear sparse features of every level prepended to the variance feature
    (refinement.py:29-41): channel order [fine | mid | coarse | var] for xs coarse->fine."""
    n_pts, n_hyp = pts.shape[:2]
    features = pts_feat
    for x in xs:
        n_b = int(x['batch'].max()) + 1
        min_pts = torch.zeros((n_b, 3), dtype=torch.float32).scatter_reduce_(
            0, x['batch'][:, None].expand(-1, 3), x['pts'], 'amin', include_self=False)
        q = pts - min_pts[pts_batch].unsqueeze(1).expand(*pts.shape)
        q = (q / x['res']) * x['stride']
        qb = torch.cat((pts_batch.view(-1, 1, 1).repeat(1, n_hyp, 1).float(), q), dim=2).view(n_pts * n_hyp, 4)
        f = mk.interpolate(x['cmap'], x['feats'], qb).view(n_pts, n_hyp, -1)
        features = torch.cat((f, features), dim=2)
    return features


def decoder_net(features, p, eps=1e-5):
    """Conv1d(k=3,pad=1)+BN+ReLU x3, Conv1d -> 1, softmax over the hypotheses
    (refinement.py:20-25,42-44). features [Np, n_hyp, Cin]."""
    x = features.transpose(2, 1)
    for i in range(3):
        x = F.conv1d(x, p['net.%d.0.weight' % i], None, 1, 1)
        x = F.batch_norm(x, p['net.%d.1.running_mean' % i], p['net.%d.1.running_var' % i],
                         p['net.%d.1.weight' % i], p['net.%d.1.bias' % i], training=False, eps=eps)
        x = F.relu(x)
    x = F.conv1d(x, p['net.3.weight'], p['net.3.bias'], 1, 1)
    return F.softmax(x.squeeze(1), dim=1)


def hypothesis_decoder(xs, pts, pts_feat, pts_batch, p):
    return decoder_net(interpolate_levels(xs, pts, pts_feat, pts_batch), p)
