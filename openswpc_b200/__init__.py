"""openswpc_b200 -- B200-native (sm_100a) implementation of OpenSWPC's swpc_3d time-stepping hot path.

Product code: hand-written CUDA kernels behind the C ABI of include/swpc3d_b200.h (openswpc_b200/csrc),
plus the host-side mirror of the reference's setup chain.  There is no CPU fallback.
"""
__version__ = "0.1.0"
