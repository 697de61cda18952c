"""B200-native photometric-alignment hot path of amin-abouee/semi-direct-visual-odometry.

The product is libsvo_b200.so (hand-written sm_100a CUDA behind the C ABI of include/svo_b200.h)
plus the reference-shaped C++ host classes in host/.  `capi` is the ctypes harness over the C ABI,
`synth` renders the synthetic KITTI-shaped workloads.  The directory name contains hyphens, so import
it with importlib.import_module("semi-direct-visual-odometry_b200") (the repo root on sys.path).
Nothing in this package touches oracle/ -- that is test infrastructure.
"""
from . import capi, shard, synth  # noqa: F401
from .capi import Context, MultiContext, SvoError, load  # noqa: F401
