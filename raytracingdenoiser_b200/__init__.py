"""B200-native implementation of NRD's per-pixel spatio-temporal filter chains behind the NRD dispatch/resource API.

The product is the native library `libnrd_b200.so` (C-ABI in include/nrd_b200.h: the nine NRD entry points + the CUDA
executor).  `nrd` is the ctypes mirror of that interface, `scene` generates synthetic inputs, `harness` wires both to
torch device memory.  Importing `nrd` fails loudly when the library has not been built.
"""
__all__ = ["nrd", "scene", "harness", "build"]
