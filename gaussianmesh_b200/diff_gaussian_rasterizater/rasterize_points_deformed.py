"""Twin of rasterize_points.py used by the edit-time renderer (`NewGaussianRasterizer`).

The reference's `rasterize_points_deformed.py` is a copy of `rasterize_points.py` whose only
difference is that the number of SH coefficients per Gaussian is hard-coded to 16 in the forward
(reference rasterize_points_deformed.py:157 and :230) -- needed because the edit renderer passes a
precomputed covariance together with SHs, for which the first file derives M = 0.
"""
from __future__ import annotations

import functools

from . import rasterize_points as _rp

M_DEFORMED = 16

compute_buffer_size = _rp.compute_buffer_size
mark_visible = _rp.mark_visible
RasterizeGaussiansCUDA = functools.partial(_rp.RasterizeGaussiansCUDA, _force_m=M_DEFORMED)
# the backward of the reference twin still reads M from the sh tensor (its :301)
RasterizeGaussiansBackwardCUDA = _rp.RasterizeGaussiansBackwardCUDA
