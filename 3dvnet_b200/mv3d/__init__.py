"""Host-side mirror of the reference's ``mv3d`` package for the hot path: same module paths,
class names, constructor / forward signatures and state_dict keys
(/root/reference/mv3d/{utils,lightningmodel}.py, /root/reference/mv3d/subnetworks/*.py), with
every heavy operation enqueued on the B200 through lib3dvnet_b200.so (include/dv3d.h)."""
