"""Differentiable Gaussian rasterizer op -- drop-in for the reference package
`gaussian_renderer.diff_gaussian_rasterizater` (reference `__init__.py:1-328`).

Same public names and call signatures: `GaussianRasterizationSettings`, `GaussianRasterizer`,
`NewGaussianRasterizer`, `_RasterizeGaussians` / `_NewRasterizeGaussians`.  The host framework is
torch (Jittor is not installable here); the native side is libCudaRasterizer.so in this directory,
called through ctypes (`rasterize_points.py`).

Differences a caller can see, all opt-in:
  * `GaussianRasterizer(settings, arena=RenderArena(...))` renders without the reference's
    mid-frame host synchronisation (see `gaussianmesh_b200.arena`).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
from torch import nn

from . import rasterize_points
from . import rasterize_points_deformed


class GaussianRasterizationSettings(NamedTuple):
    # reference __init__.py:6-18
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _make_function(glue, name):
    class _Fn(torch.autograd.Function):
        """reference __init__.py:23-124 (`execute` -> forward, `grad` -> backward)."""

        @staticmethod
        def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    raster_settings, arena):
            args = (
                raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                raster_settings.image_height, raster_settings.image_width, sh, raster_settings.sh_degree,
                raster_settings.campos, raster_settings.prefiltered, raster_settings.debug,
            )
            if raster_settings.debug:
                try:
                    out = glue.RasterizeGaussiansCUDA(*args, arena=arena)
                except Exception as ex:
                    torch.save(args, "snapshot_fw.dump")   # reference :64-66
                    print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                    raise ex
            else:
                out = glue.RasterizeGaussiansCUDA(*args, arena=arena)
            num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = out
            ctx.raster_settings = raster_settings
            ctx.num_rendered = num_rendered
            ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
                                  geomBuffer, binningBuffer, imgBuffer)
            ctx.mark_non_differentiable(radii)
            return color, radii

        @staticmethod
        def backward(ctx, grad_out_color, _grad_radii):
            raster_settings = ctx.raster_settings
            (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh,
             geomBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
            args = (raster_settings.bg, means3D, radii, colors_precomp, scales, rotations,
                    raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                    raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                    grad_out_color, sh, raster_settings.sh_degree, raster_settings.campos, geomBuffer,
                    ctx.num_rendered, binningBuffer, imgBuffer, raster_settings.debug)
            if raster_settings.debug:
                try:
                    grads = glue.RasterizeGaussiansBackwardCUDA(*args)
                except Exception as ex:
                    torch.save(args, "snapshot_bw.dump")   # reference :105-107
                    print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                    raise ex
            else:
                grads = glue.RasterizeGaussiansBackwardCUDA(*args)
            (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
             grad_scales, grad_rotations) = grads

            def g(grad, inp):
                # absent inputs are empty sentinels (reference :158-168): no gradient for them
                return grad if inp.numel() != 0 else None

            # reference :112-122: (means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
            # cov3Ds_precomp, None)
            return (grad_means3D, grad_means2D, g(grad_sh, sh), g(grad_colors_precomp, colors_precomp),
                    grad_opacities, g(grad_scales, scales), g(grad_rotations, rotations),
                    g(grad_cov3Ds_precomp, cov3Ds_precomp), None, None)

    _Fn.__name__ = _Fn.__qualname__ = name
    return _Fn


_RasterizeGaussians = _make_function(rasterize_points, "_RasterizeGaussians")
_NewRasterizeGaussians = _make_function(rasterize_points_deformed, "_NewRasterizeGaussians")


def _make_module(fn, glue, name):
    class _Rasterizer(nn.Module):
        """reference __init__.py:126-174 (GaussianRasterizer) / :280-328 (NewGaussianRasterizer)."""

        def __init__(self, raster_settings: GaussianRasterizationSettings, arena=None):
            super().__init__()
            self.raster_settings = raster_settings
            self.arena = arena

        def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
            with torch.no_grad():
                rs = self.raster_settings
                return glue.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
            if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
                raise Exception('Please provide excatly one of either SHs or precomputed colors!')
            if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                    ((scales is not None or rotations is not None) and cov3D_precomp is not None):
                raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
            empty = means3D.new_empty(0)
            shs = empty if shs is None else shs
            colors_precomp = empty if colors_precomp is None else colors_precomp
            scales = empty if scales is None else scales
            rotations = empty if rotations is None else rotations
            cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
            return fn.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                            self.raster_settings, self.arena)

        # Jittor modules are called through `execute`; keep the alias so ported call sites work.
        execute = forward

    _Rasterizer.__name__ = _Rasterizer.__qualname__ = name
    return _Rasterizer


GaussianRasterizer = _make_module(_RasterizeGaussians, rasterize_points, "GaussianRasterizer")
NewGaussianRasterizer = _make_module(_NewRasterizeGaussians, rasterize_points_deformed, "NewGaussianRasterizer")

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "NewGaussianRasterizer",
           "_RasterizeGaussians", "_NewRasterizeGaussians", "rasterize_points", "rasterize_points_deformed"]
