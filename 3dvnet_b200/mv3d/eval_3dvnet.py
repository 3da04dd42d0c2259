"""`process_scene` - the `pred_func` plugin of the reference's evaluation harness
(/root/reference/mv3d/eval-3dvnet.py:26-127, passed to mv3d.eval.main.main at :134-150), same signature and
return value: ``process_scene(batch, scene, dset, net) -> (all_depth [n_ref, H, W] numpy, None, None)``.

The reference chunks the scene three ways to fit a 24 GB GPU (INIT_DEPTH_BATCH = 18 reference views per MVSNet call,
OFFSET_BATCH = 16 per PointFlow call, UPSAMPLE_BATCH = 100 per PropagationNet call, eval-3dvnet.py:12-14) and drives
every op from Python. None of the chunk boundaries changes a result: the 2D backbone, the cost volume, the PointFlow
passes and the PropagationNets are per-image / per-reference-view computations, only the scene model sees all views -
and it is not chunked in the reference either (eval-3dvnet.py:76). On a B200 (180 GB) the scene goes through whole:
the backbone in image chunks (its activations are the only memory that scales badly), then ONE native call for the
hot path of all reference views (csrc/engine.cu), then the three PropagationNets."""
import torch

from .. import ops

# eval-3dvnet.py:16-23
DEPTH_CONFIG = {'depth_start': 0.5, 'depth_interval': 0.05, 'n_intervals': 96, 'size': (56, 56)}
OFFSETS_LIST = [[0.05, 0.05, 0.025], [0.05, 0.05, 0.025]]
BACKBONE_BATCH = 64   # images per backbone call


def process_scene(batch, scene=None, dset=None, net=None, depth_config=None, offsets_list=None):
    """batch: the reference's Batch of ONE scene (images [n,3,H,W], rotmats, tvecs, K, ref_src_edges); `scene` and
    `dset` are accepted for signature compatibility (the reference uses dset.n_src_on_either_side only to compute
    its chunk windows)."""
    if net is None:
        raise ValueError('process_scene: net is required')
    cfg = DEPTH_CONFIG if depth_config is None else depth_config
    offsets = OFFSETS_LIST if offsets_list is None else offsets_list
    dev = net.device
    with torch.no_grad():
        images = batch.images.to(dev)
        n_imgs = images.shape[0]
        fh, fq = [], []
        for i in range(0, n_imgs, BACKBONE_BATCH):          # eval-3dvnet.py:41-63 (feature part)
            h, q = net._backbone(images[i:i + BACKBONE_BATCH])
            fh.append(h.clone())
            fq.append(q.clone())
        fh, fq = torch.cat(fh), torch.cat(fq)
        fq = fq.contiguous(memory_format=torch.channels_last)
        plan = ops.edge_plan(batch.ref_src_edges, dev)
        images_batch = getattr(batch, 'images_batch', None)
        if images_batch is None:                            # eval-3dvnet.py:55,69: one scene
            images_batch = torch.zeros(n_imgs, dtype=torch.long)
        depth = net.hot_path(fq, batch.rotmats.to(dev), batch.tvecs.to(dev), batch.K.to(dev), plan,
                             images_batch.to(dev), cfg, offsets)      # eval-3dvnet.py:58-99
        full = net.upsample(depth, plan.ref_idx, fq, fh, images)      # eval-3dvnet.py:101-125
        return full.detach().cpu().numpy(), None, None
