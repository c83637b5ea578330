"""B200-native candidate-pair similarity scoring (drop-in for the reference's similarity.py path).

Layout: ``csrc/`` CUDA kernels + C ABI (include/blp.h), ``_lib`` ctypes binding and nvcc recipe,
``graph`` the device-resident graph handle, ``similarity`` / ``util`` the reference-facing
signatures and file formats, ``synth`` seeded Yelp-shaped inputs, ``dist`` the multi-GPU pair
sharding.  Importing the package does not need a GPU; scoring does, and has no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ['_lib', 'graph', 'similarity', 'util', 'synth', 'dist']
