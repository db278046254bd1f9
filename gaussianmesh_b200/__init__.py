"""gaussianmesh_b200 -- B200-native (sm_100a) rasterization path for mesh-bound Gaussian splatting.

The product is `diff_gaussian_rasterizater/libCudaRasterizer.so` (hand-written CUDA behind the C ABI of
include/gm_rasterizer.h and the reference's CudaRasterizer:: C++ symbols); this package is the host
side that mirrors the reference's Python interface for that path.  There is no CPU fallback: every
submodule except `build` and `synthetic` loads the library on import and raises if it is missing or
stale.  Build with `python -m gaussianmesh_b200.build`.
"""
__version__ = "0.1.0"


def __getattr__(name):
    # `build` must stay importable before the library exists, so the binding is loaded on first use.
    if name in ("version", "RasterizerError", "lib"):
        from . import _lib
        return getattr(_lib, name)
    raise AttributeError(name)
