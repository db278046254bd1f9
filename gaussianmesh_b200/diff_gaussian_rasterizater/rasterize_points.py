"""FFI glue between the autograd op and libCudaRasterizer.so.

Mirrors the reference's `gaussian_renderer/diff_gaussian_rasterizater/rasterize_points.py`
(same function names, argument order and return tuples) with torch tensors + ctypes in place of
Jittor vars + `jt.code`.  Every device pointer handed to the library is a torch CUDA tensor owned by
the caller; kernels are launched on torch's current stream.

There is no CPU path: tensors must live on a CUDA device and the shared library must be present
(`gaussianmesh_b200._lib` raises at import otherwise).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .._lib import lib, check, RasterizerError, GM_ERR_BAD_ARGUMENT, GM_BACKWARD_OVERWRITE

# rasterize_points_deformed.py differs from rasterize_points.py only by forcing the number of SH
# coefficients per Gaussian to 16 (reference :157 and :230); it passes `_force_m=16`.


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor; empty tensors are the reference's "absent" sentinel -> NULL."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.numel() == 0:
        return t
    if not t.is_cuda:
        raise RasterizerError(name, GM_ERR_BAD_ARGUMENT, "tensor is not on a CUDA device (there is no CPU path)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


def compute_buffer_size(means3D: torch.Tensor, image_width: int, image_height: int) -> Tuple[int, int]:
    """Bytes of the geometry and image chunks (reference rasterize_points.py:63-86)."""
    P = means3D.shape[0]
    return int(lib.gm_required_geom(P)), int(lib.gm_required_image(image_width * image_height))


def mark_visible(means3D: torch.Tensor, viewmatrix: torch.Tensor, projmatrix: torch.Tensor) -> torch.Tensor:
    """Near-plane visibility mask (reference rasterize_points.py:40-58 -> Rasterizer::markVisible)."""
    means3D = _f32(means3D, "mark_visible")
    P = means3D.shape[0]
    present = torch.zeros(P, dtype=torch.bool, device=means3D.device)
    if P != 0:
        viewmatrix = _f32(viewmatrix, "mark_visible")
        projmatrix = _f32(projmatrix, "mark_visible")
        check(lib.gm_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), present.data_ptr(), _stream()),
              "gm_mark_visible")
    return present


def _num_coeffs(sh: torch.Tensor, cov3D_precomp: torch.Tensor, force_m: Optional[int]) -> int:
    if force_m is not None:
        return force_m
    # reference rasterize_points.py:154-158: M is taken from sh only when no precomputed covariance is
    # given (a quirk the `_deformed` variant exists to work around); kept for drop-in parity.
    if cov3D_precomp.numel() == 0 and sh.dim() >= 2:
        return int(sh.shape[1])
    return 0


def RasterizeGaussiansCUDA(
    background: torch.Tensor,
    means3D: torch.Tensor,
    colors: torch.Tensor,
    opacity: torch.Tensor,
    scales: torch.Tensor,
    rotations: torch.Tensor,
    scale_modifier: float,
    cov3D_precomp: torch.Tensor,
    viewmatrix: torch.Tensor,
    projmatrix: torch.Tensor,
    tan_fovx: float,
    tan_fovy: float,
    image_height: int,
    image_width: int,
    sh: torch.Tensor,
    degree: int,
    campos: torch.Tensor,
    prefiltered: bool,
    debug: bool,
    arena=None,
    _force_m: Optional[int] = None,
) -> Tuple[int, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Forward rasterization (reference rasterize_points.py:88-274).

    Returns (num_rendered, out_color[3,H,W], radii[P], geomBuffer, binningBuffer, imageBuffer).

    Without `arena` this is the reference's two-phase protocol: gm_forward_0, a blocking 4-byte
    read of the instance count, allocation of the binning chunk, gm_forward_1.  With a
    `gaussianmesh_b200.arena.RenderArena` the single-call, sync-free gm_forward is used over the
    arena's persistent chunks and `num_rendered` is the arena's binning capacity (the value the
    backward needs); the arena verifies the overflow flag later.
    """
    means3D = _f32(means3D, "means3D")
    P = means3D.shape[0]
    dev = means3D.device
    H, W = int(image_height), int(image_width)
    if P == 0:
        # reference :156,233: the native call is skipped and the image stays all-zero (NOT background)
        z = torch.zeros(0, dtype=torch.uint8, device=dev)
        return 0, torch.zeros(3, H, W, dtype=torch.float32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev), z, z, z

    background = _f32(background, "background")
    colors = _f32(colors, "colors")
    opacity = _f32(opacity, "opacity")
    scales = _f32(scales, "scales")
    rotations = _f32(rotations, "rotations")
    cov3D_precomp = _f32(cov3D_precomp, "cov3D_precomp")
    viewmatrix = _f32(viewmatrix, "viewmatrix")
    projmatrix = _f32(projmatrix, "projmatrix")
    sh = _f32(sh, "sh")
    campos = _f32(campos, "campos")
    M = _num_coeffs(sh, cov3D_precomp, _force_m)

    view_args = (_ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity), _ptr(scales), float(scale_modifier),
                 _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                 float(tan_fovx), float(tan_fovy))
    stream = _stream()

    if arena is not None:
        return arena.forward(P, int(degree), M, background, W, H, view_args, bool(prefiltered), bool(debug), stream)

    geom_size, img_size = compute_buffer_size(means3D, W, H)
    geomBuffer = torch.empty(geom_size, dtype=torch.uint8, device=dev)
    radii = torch.zeros(P, dtype=torch.int32, device=dev)
    num_rendered = check(
        lib.gm_forward_0(geomBuffer.data_ptr(), P, int(degree), M, _ptr(background), W, H, *view_args,
                         int(bool(prefiltered)), radii.data_ptr(), int(bool(debug)), stream),
        "gm_forward_0")
    binningBuffer = torch.empty(int(lib.gm_required_binning(num_rendered)), dtype=torch.uint8, device=dev)
    imageBuffer = torch.empty(img_size, dtype=torch.uint8, device=dev)
    out_color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
    check(
        lib.gm_forward_1(geomBuffer.data_ptr(), binningBuffer.data_ptr(), imageBuffer.data_ptr(), P, int(degree), M,
                         num_rendered, _ptr(background), W, H, *view_args, int(bool(prefiltered)),
                         out_color.data_ptr(), radii.data_ptr(), int(bool(debug)), stream),
        "gm_forward_1")
    return num_rendered, out_color, radii, geomBuffer, binningBuffer, imageBuffer


def RasterizeGaussiansBackwardCUDA(
    background: torch.Tensor,
    means3D: torch.Tensor,
    radii: torch.Tensor,
    colors: torch.Tensor,
    scales: torch.Tensor,
    rotations: torch.Tensor,
    scale_modifier: float,
    cov3D_precomp: torch.Tensor,
    viewmatrix: torch.Tensor,
    projmatrix: torch.Tensor,
    tan_fovx: float,
    tan_fovy: float,
    dL_dout_color: torch.Tensor,
    sh: torch.Tensor,
    degree: int,
    campos: torch.Tensor,
    geomBuffer: torch.Tensor,
    R: int,
    binningBuffer: torch.Tensor,
    imageBuffer: torch.Tensor,
    debug: bool,
    _force_m: Optional[int] = None,
) -> Tuple[torch.Tensor, ...]:
    """Backward rasterization (reference rasterize_points.py:276-398).

    Returns (dL_dmeans2D[P,3], dL_dcolors[P,3], dL_dopacity[P,1], dL_dmeans3D[P,3], dL_dcov3D[P,6],
    dL_dsh[P,M,3], dL_dscales[P,3], dL_drotations[P,4]).
    """
    means3D = _f32(means3D, "means3D")
    P = means3D.shape[0]
    dev = means3D.device
    sh = _f32(sh, "sh")
    # reference :301: M = sh.size(1) if sh.size(0) != 0 else 0 (the `_deformed` twin forces 16)
    M = _force_m if _force_m is not None else (int(sh.shape[1]) if sh.numel() != 0 else 0)
    Mg = int(sh.shape[1]) if sh.numel() != 0 else 0
    # One slab for all nine gradient tensors (reference :302-310 allocates nine zero tensors).  Only the four
    # atomically accumulated ones are zero-filled; the per-Gaussian rows are written in full by
    # gm_backward_ex(GM_BACKWARD_OVERWRITE).  Unused variants (e.g. dL_dsh on the colour path) are zero-filled too.
    use_sh, use_sr = colors.numel() == 0, cov3D_precomp.numel() == 0
    sizes = [3 * P, 4 * P, P, 3 * P,                                  # means2D, conic, opacity, colors (accumulated)
             3 * P, 6 * P, Mg * 3 * P, 3 * P, 4 * P]                  # means3D, cov3D, sh, scales, rotations
    offs = [0]
    for sz in sizes:
        offs.append(offs[-1] + ((sz + 31) // 32) * 32)   # keep every tensor 128-byte aligned
    slab = torch.empty(offs[-1], dtype=torch.float32, device=dev)
    slab[:offs[4]].zero_()
    part = [slab[offs[i]:offs[i] + sizes[i]] for i in range(len(sizes))]
    if not use_sh:
        part[6].zero_()
    if not use_sr:
        part[7].zero_(); part[8].zero_()
    dL_dmeans2D = part[0].view(P, 3)
    dL_dconic = part[1].view(P, 2, 2)
    dL_dopacity = part[2].view(P, 1)
    dL_dcolors = part[3].view(P, 3)
    dL_dmeans3D = part[4].view(P, 3)
    dL_dcov3D = part[5].view(P, 6)
    dL_dsh = part[6].view(P, Mg, 3)
    dL_dscales = part[7].view(P, 3)
    dL_drotations = part[8].view(P, 4)
    if P != 0:
        dL_dout_color = _f32(dL_dout_color, "dL_dout_color")
        H, W = int(dL_dout_color.shape[1]), int(dL_dout_color.shape[2])
        background = _f32(background, "background")
        colors = _f32(colors, "colors")
        scales = _f32(scales, "scales")
        rotations = _f32(rotations, "rotations")
        cov3D_precomp = _f32(cov3D_precomp, "cov3D_precomp")
        viewmatrix = _f32(viewmatrix, "viewmatrix")
        projmatrix = _f32(projmatrix, "projmatrix")
        campos = _f32(campos, "campos")
        check(
            lib.gm_backward_ex(P, int(degree), M, int(R), _ptr(background), W, H, _ptr(means3D), _ptr(sh), _ptr(colors),
                            _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
                            _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy),
                            _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                            dL_dout_color.data_ptr(), dL_dmeans2D.data_ptr(), dL_dconic.data_ptr(),
                            dL_dopacity.data_ptr(), dL_dcolors.data_ptr(), dL_dmeans3D.data_ptr(),
                            dL_dcov3D.data_ptr(), _ptr(dL_dsh) if Mg else None, dL_dscales.data_ptr(),
                            dL_drotations.data_ptr(), int(bool(debug)), GM_BACKWARD_OVERWRITE, _stream()),
            "gm_backward_ex")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations
