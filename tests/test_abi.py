"""CPU checks of the drop-in boundary: the shared library builds/loads, exports every symbol
include/dv3d.h declares, and rejects bad arguments without touching a GPU."""
import ctypes
import importlib
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def L():
    build = importlib.import_module('3dvnet_b200.build')
    build.build()
    return importlib.import_module('3dvnet_b200._lib').lib()


def test_header_parses_and_every_symbol_is_exported(L):
    names = set(L.protos)
    for must in ('dv3d_planesweep_var', 'dv3d_points_var', 'dv3d_conv3d_bn_relu', 'dv3d_deconv3d_bn_relu',
                 'dv3d_prob_softargmin', 'dv3d_voxel_grid', 'dv3d_voxelize', 'dv3d_linear', 'dv3d_segment_max',
                 'dv3d_hash_build', 'dv3d_coarsen', 'dv3d_kernel_map', 'dv3d_sparse_conv',
                 'dv3d_concat_linear_gn_relu', 'dv3d_sparse_interp', 'dv3d_conv1d_bn_relu', 'dv3d_decoder_head',
                 'dv3d_camera_tables', 'dv3d_nchw_to_nhwc', 'dv3d_launch_count',
                 'dv3d_gemm_pack_weights', 'dv3d_gemm_pack_bytes', 'dv3d_set_gemm_precision'):
        assert must in names, must
    for n in names:
        assert hasattr(L.cdll, n), 'library does not export %s' % n
    assert len(names) >= 30


def test_abi_version_and_error_reporting(L):
    assert L.cdll.dv3d_abi_version() == 4
    # argument validation happens before any CUDA call: safe without a GPU
    rc = L.cdll.dv3d_planesweep_var(None, 1, 16, 4, 4, None, None, None, None, 1, 0.5, 0.05, 8, 8, 8, 16, 16, None,
                                    None)
    assert rc == -1 and 'C must be 32' in L.last_error()
    rc = L.cdll.dv3d_sparse_conv(None, 0, 64, None, 0, None, None, 64, None, None, None, 0, None, 0, None, None)
    assert rc == -1 and 'bad arguments' in L.last_error()
    assert L.cdll.dv3d_hash_bytes(1000) == 2048 * 12
    # tensor-core weight image: big + small part of every value; K must be a multiple of 32
    assert L.cdll.dv3d_gemm_pack_bytes(27 * 128, 128) == 27 * 128 * 128 * 8
    assert L.cdll.dv3d_gemm_pack_bytes(48, 128) == 0
    assert L.cdll.dv3d_set_gemm_precision(3) == -1 and L.cdll.dv3d_get_gemm_precision() == 1


def test_struct_layout_matches_header(L):
    mod = importlib.import_module('3dvnet_b200._lib')
    assert ctypes.sizeof(mod.VoxelGrid) == 96


def test_ops_refuse_cpu_tensors():
    import torch
    ops = importlib.import_module('3dvnet_b200.ops')
    with pytest.raises(RuntimeError, match='CUDA tensor'):
        ops.nchw_to_nhwc(torch.zeros(1, 32, 4, 4))


def test_forward_only_modules_refuse_training_mode():
    """the sparse scene model and the decoder have inference kernels only: training mode must raise, not fall back
    (MVSNet is differentiable: tests/test_gpu_autograd.py)"""
    import torch
    m = importlib.import_module('3dvnet_b200.mv3d.subnetworks.refinement')
    dec = m.HypothesisDecoder(352, 128)
    assert dec.training
    with pytest.raises(NotImplementedError, match='inference'):
        dec(None, torch.zeros(4, 7, 3), torch.zeros(4, 7, 32), torch.zeros(4, dtype=torch.long))


def test_edge_plan_matches_reference_grouping():
    import numpy as np
    import torch
    ops = importlib.import_module('3dvnet_b200.ops')
    synth = importlib.import_module('3dvnet_b200.synth')
    e = torch.from_numpy(synth.make_edges(9, 2, 2, include_self=True))
    e = e[:, torch.randperm(e.shape[1], generator=torch.Generator().manual_seed(0))]  # ragged, unsorted
    e = e[:, :-3]
    plan = ops.EdgePlan(e, 'cpu')
    ref_idx, gather = torch.unique(e[0], return_inverse=True)  # mvsnet.py:179
    assert plan.ref_idx.tolist() == ref_idx.tolist()
    rp = plan.rowptr.numpy()
    for r in range(plan.n_ref):
        mine = plan.edge_src[rp[r]:rp[r + 1]].tolist()
        theirs = e[1][gather == r].tolist()
        assert mine == theirs          # same members, same relative order
        assert set(plan.edge_ref[rp[r]:rp[r + 1]].tolist()) == {int(ref_idx[r])}
    assert rp[-1] == e.shape[1]
