"""remhos_b200: B200-native implementation of the Remhos explicit DG transport/remap RK-stage
path (CUDA kernels + C ABI in csrc/, host-side mirror of the reference's solver interfaces in
host/).  Python here is only the ctypes binding used by tests and bench.py."""
from .capi import Mesh, Halo, Context, RmhError, launch_count, LIB_PATH  # noqa: F401
