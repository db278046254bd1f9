"""ctypes binding of libCudaRasterizer.so (the flat C ABI of include/gm_rasterizer.h).

There is NO fallback: if the library is missing or does not export a symbol the header declares,
importing this module raises.  Build it with `python -m gaussianmesh_b200.build`.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "diff_gaussian_rasterizater" / "libCudaRasterizer.so"

GM_OK = 0
GM_ERR_CUDA = -1
GM_ERR_BAD_ARGUMENT = -2
GM_ERR_TOO_MANY_TILES = -3
GM_ERR_BINNING_OVERFLOW = -4
GM_BACKWARD_OVERWRITE = 1

_ERR_NAMES = {
    GM_ERR_CUDA: "GM_ERR_CUDA",
    GM_ERR_BAD_ARGUMENT: "GM_ERR_BAD_ARGUMENT",
    GM_ERR_TOO_MANY_TILES: "GM_ERR_TOO_MANY_TILES",
    GM_ERR_BINNING_OVERFLOW: "GM_ERR_BINNING_OVERFLOW",
}

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_z = C.c_size_t

class AdamTensor(C.Structure):
    """gm_adam_tensor of include/gm_rasterizer.h"""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_size_t), ("lr", C.c_float), ("lr_head", C.c_float), ("period", C.c_uint32),
                ("split", C.c_uint32)]


class ForwardEpilogue(C.Structure):
    """gm_forward_epilogue of include/gm_rasterizer.h"""
    _fields_ = [("target", C.c_void_p), ("target_is_u8", C.c_int), ("loss", C.c_void_p), ("dL_dimg", C.c_void_p),
                ("zero_ptr", C.c_void_p), ("zero_floats", C.c_size_t)]


class AdamSegment(C.Structure):
    """gm_adam_segment of include/gm_rasterizer.h"""
    _fields_ = [("offset", C.c_size_t), ("numel", C.c_size_t), ("lr", C.c_float), ("lr_head", C.c_float),
                ("period", C.c_uint32), ("split", C.c_uint32)]


# name -> (restype, argtypes); the order of arguments is that of include/gm_rasterizer.h
_VIEW_ARGS = [_p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _f, _f]  # means3D .. tan_fovy
SIGNATURES = {
    "gm_version": (C.c_char_p, []),
    "gm_last_error": (C.c_char_p, []),
    "gm_profile_num_stages": (_i, []),
    "gm_profile_stage_name": (C.c_char_p, [_i]),
    "gm_profile_begin": (None, []),
    "gm_profile_end": (_i, [_p, _p]),
    "gm_required_geom": (_z, [_z]),
    "gm_required_image": (_z, [_z]),
    "gm_required_binning": (_z, [_z]),
    "gm_binning_capacity": (_z, [_z]),
    "gm_mark_visible": (_i, [_i, _p, _p, _p, _p, _p]),
    "gm_forward_0": (_i, [_p, _i, _i, _i, _p, _i, _i, *_VIEW_ARGS, _i, _p, _i, _p]),
    "gm_forward_1": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _i, _i, *_VIEW_ARGS, _i, _p, _p, _i, _p]),
    "gm_forward": (_i, [_p, _p, _z, _p, _i, _i, _i, _p, _i, _i, *_VIEW_ARGS, _i, _p, _p, _i, _p, _p]),
    "gm_forward_ex": (_i, [_p, _p, _z, _p, _i, _i, _i, _p, _i, _i, *_VIEW_ARGS, _i, _p, _p, _i, _p, C.POINTER(ForwardEpilogue), _p]),
    "gm_forward_status": (_i, [_p, _p, _p, _p]),
    "gm_geom_view": (None, [_p, _z, _p]),
    "gm_backward": (_i, [_i, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _f, _f, _p,
                         _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "gm_backward_ex": (_i, [_i, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _f, _f, _p,
                            _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "gm_mesh_bind_forward": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gm_mesh_bind_backward": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _p, _p, _p,
                                   _p, _p, _p, _p, _p, _p]),
    "gm_deform_gaussians": (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "gm_sh_to_rgb_rotated": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "gm_sh_to_rgb_rotated_backward": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gm_cov3d_from_scale_rot": (_i, [_i, _p, _f, _p, _p, _p]),
    "gm_cov3d_from_scale_rot_backward": (_i, [_i, _p, _f, _p, _p, _p, _p, _p]),
    "gm_load_mesh": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "gm_l1_loss": (_i, [_z, _p, _p, _p, _p, _p]),
    "gm_l1_loss_u8": (_i, [_z, _p, _p, _p, _p, _p]),
    "gm_image_u8_to_float": (_i, [_z, _p, _p, _p]),
    "gm_photometric_scratch_bytes": (_z, [_i, _i, _i]),
    "gm_photometric_loss": (_i, [_i, _i, _i, _p, _p, _f, _p, _p, _p, _p]),
    "gm_mesh_restrict_loss": (_i, [_i, _p, _p, _p, _p, _f, _p, _p, _i, _p]),
    "gm_adam_step": (_i, [_i, C.POINTER(AdamTensor), _i, _f, _f, _f, _p]),
    "gm_adam_step_gated": (_i, [_i, C.POINTER(AdamTensor), _i, _f, _f, _f, _p, _p]),
    "gm_frame_overflow_flag": (_p, [_p]),
    "gm_densify_stats_gated": (_i, [_i, _p, _p, _p, _p, _p, _p, _p]),
    "gm_adam_shard_range": (None, [_z, _i, _i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "gm_adam_step_sharded_p2p": (_i, [_i, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i, C.POINTER(AdamSegment), _z,
                                      _p, _p, _i, _f, _f, _f, _p]),
    "gm_adam_step_sharded_mc": (_i, [_i, _i, _p, _p, _p, _i, C.POINTER(AdamSegment), _z, _p, _p, _i, _f, _f, _f, _p]),
    "gm_densify_stats": (_i, [_i, _p, _p, _p, _p, _p, _p]),
    "gm_acap_build_rings": (_i, [_i, _i, _p, _p, _p, _p, _p]),
    "gm_acap_rest": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gm_acap_get_rs": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
}


class RasterizerError(RuntimeError):
    def __init__(self, where: str, code: int, detail: str = ""):
        self.code = code
        name = _ERR_NAMES.get(code, str(code))
        super().__init__(f"{where} failed: {name}" + (f" ({detail})" if detail else ""))


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library is the product and there is no fallback. "
            "Build it with `python -m gaussianmesh_b200.build`.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, where: str) -> int:
    if rc < 0:
        detail = lib.gm_last_error().decode() if rc in (GM_ERR_CUDA, GM_ERR_BAD_ARGUMENT) else ""
        if rc == GM_ERR_BAD_ARGUMENT and "aligned" not in detail:
            detail = ""          # the library only describes alignment rejections; anything else in there is stale
        raise RasterizerError(where, rc, detail)
    return rc


def version() -> str:
    return lib.gm_version().decode()


def profile_begin() -> None:
    lib.gm_profile_begin()


def profile_end() -> dict:
    """{stage name: (total ms, launches)} for every stage launched since profile_begin()."""
    n = lib.gm_profile_num_stages()
    ms = (C.c_float * n)()
    cnt = (C.c_uint64 * n)()
    check(lib.gm_profile_end(ms, cnt), "gm_profile_end")
    return {lib.gm_profile_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}
