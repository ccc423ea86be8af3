"""B200-native burst-coast hot path (sde_burst_coast drop-in): CUDA kernels behind a C ABI.

`params` mirrors the reference's parameter layers; `swarm` binds libsbc.so through ctypes.
"""
from . import params  # noqa: F401
