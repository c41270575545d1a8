"""mpi4dl_b200 -- B200-native spatial-parallel convolution engine behind the torchgems API.

Layout (only what the hot path needs):
    csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/spconv.h)
    libspconv.so     built in-tree by build.py (nvcc -gencode arch=compute_100a,code=sm_100a)
    _lib.py          ctypes binding (no fallback: raises when the library is missing)
    torchgems/       host-side mirror of the reference's torchgems package for this path
"""
__version__ = "0.1.0"
