"""relightable-nr on B200: hand-written sm_100a kernels behind the reference's operator API.

The CUDA library (librnr_b200.so, C ABI in include/rnr_b200.h) is mandatory: importing the host
modules works without it (so CPU-only tooling can introspect), but every operator raises if it is
missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"
