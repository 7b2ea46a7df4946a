"""karamelo_b200 - B200-native Material Point Method time-step engine.

The product is native: ``csrc/`` holds the sm_100a CUDA kernels behind the C ABI of
``include/kml.h`` (``libkml.so``); ``host/`` holds the C++ driver that keeps Karamelo's
input-script command surface (``libkml_host.so`` and the ``kml`` CLI).  This Python
package is plumbing only: ctypes bindings used by the tests, ``bench.py`` and the
multi-GPU launcher.  It never falls back to a CPU implementation: if the CUDA
libraries have not been built, importing :class:`Engine` raises.
"""
from .api import Engine, P, N, lib_dir, load_host_library, KmlError  # noqa: F401

__all__ = ["Engine", "P", "N", "lib_dir", "load_host_library", "KmlError"]
