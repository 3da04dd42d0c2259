"""oracle/ — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package, and only as the checker or the timed CPU baseline — never as the
product path (3dvnet_b200/ raises if its CUDA library is missing; it has no CPU fallback).

Pinning status:
* path A (plane sweep, warp, variance, CostRegNet, soft-argmin), the point-level warp /
  variance, voxelize, PointNet, GroupNorm wiring, decoder Conv1d stack: PINNED against the
  unmodified reference modules imported from /root/reference in the build container
  (oracle/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
  The reference has no tests or golden vectors of its own.
* torch_scatter / torch_geometric / torch_cluster semantics: restated,
  the libraries are absent -> parity unpinned at that boundary.
* MinkowskiEngine ops: restated (A.4) -> PARITY UNPINNED against ME; anchored on the
  reference's dense `forward_forloop` and dense conv3d (tests/test_oracle_minkowski.py).
"""
