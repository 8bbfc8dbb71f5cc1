"""shgo_b200 -- B200-native sampling-and-classification hot path for SHGO.

Public surface:
    shgo, SHGO          drop-in for the reference's ``shgo()`` / ``SHGO`` (shgo/_shgo.py:22,490):
                        same arguments and OptimizeResult fields; the vertex complex behind it is
                        :class:`DeviceComplex` (CUDA), local minimisation stays in scipy.
    DeviceComplex       replacement for shgo._shgo_lib._complex.Complex at the `SHGO.HC` seam.
    functors            registered objectives / constraint sets with device functors.
    hotpath             tensor-level access to the kernels (C ABI in include/shgo_b200.h).
"""
from . import functors  # noqa: F401

__all__ = ["shgo", "SHGO", "DeviceComplex", "functors", "hotpath", "install", "uninstall"]


def __getattr__(name):
    # torch / CUDA are imported lazily so that `import shgo_b200` works in tooling contexts
    if name == "hotpath":
        import importlib
        return importlib.import_module(".hotpath", __name__)
    if name in ("shgo", "SHGO", "install", "uninstall"):
        from . import api
        return getattr(api, name)
    if name == "DeviceComplex":
        from .complex import DeviceComplex
        return DeviceComplex
    raise AttributeError(name)
