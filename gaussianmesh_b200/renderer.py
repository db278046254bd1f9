"""Render functions above the rasterizer op -- the callers of the hot path.

  render()              mirror of gaussian_renderer.render (reference gaussian_renderer/__init__.py:26-143)
                        for the mesh-bound model, CUDA SH / CUDA covariance branch
  MeshGaussianModel     the accessors of scene/mesh_based_gaussian_model.py:122-174 that feed render(),
                        computed by the fused bind kernel
  DeformedObject        edit-time object: SingleObjectDeform.deform_gaussian + ObjectVisualTool.render_gaussian
                        (edittool/__init__.py:103-131, 400-475)
  ViewBatchRenderer     sync-free forward rendering of a list of views on one GPU
  train_step()          config-4 step: forward + L1 + full backward through the C ABI without autograd overhead
  shard_views()         contiguous-block partition of independent views over ranks (no collective)

torch is plumbing here (device memory, streams); all arithmetic is in libCudaRasterizer.so.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import mesh_gaussians as mg
from ._lib import lib, check, GM_BACKWARD_OVERWRITE, RasterizerError, GM_ERR_BAD_ARGUMENT, ForwardEpilogue
from .arena import RenderArena
from .diff_gaussian_rasterizater import (GaussianRasterizationSettings, GaussianRasterizer, NewGaussianRasterizer)
from .cameras import DeviceCamera, upload_cameras  # noqa: F401  (re-exported)
from .synthetic import Camera  # noqa: F401
from .view_shard import shard_views, ShardContext  # noqa: F401  (re-exported)


# views: DeviceCamera / upload_cameras live in cameras.py (no dependency on the library) and are re-exported here


def make_settings(cam, bg: torch.Tensor, sh_degree: int, scaling_modifier: float = 1.0, debug: bool = False
                  ) -> GaussianRasterizationSettings:
    """reference gaussian_renderer/__init__.py:46-62"""
    return GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width),
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
        scale_modifier=scaling_modifier, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
        sh_degree=sh_degree, campos=cam.camera_center, prefiltered=False, debug=debug)


# ---------------------------------------------------------------------------------------------
# training-time model accessors + render()
# ---------------------------------------------------------------------------------------------
class MeshGaussianModel:
    """Parameters and per-Gaussian face data of MeshBasedGaussianModel, with the activations of
    scene/mesh_based_gaussian_model.py:34-43,122-174 evaluated by ONE fused kernel per frame
    (`mesh_bind`) instead of ~10 elementwise kernels; SH features are kept as a single [P,16,3]
    tensor so that get_features needs no per-frame concat (reference :167-170)."""

    def __init__(self, arrays: Dict[str, np.ndarray], device, sh_degree: int = 3, alpha_distance: float = 4.0,
                 requires_grad: bool = True):
        t = lambda k: torch.from_numpy(np.ascontiguousarray(arrays[k])).to(device)
        self._bc = t("bc_logits").requires_grad_(requires_grad)
        self._distance = t("distance").requires_grad_(requires_grad)
        self._scaling = t("log_scales").requires_grad_(requires_grad)
        self._rotation = t("rot_raw" if "rot_raw" in arrays else "rotations").requires_grad_(requires_grad)
        self._opacity = t("opacity_logit").requires_grad_(requires_grad)
        self._features = t("shs").requires_grad_(requires_grad)
        self.vertex1, self.vertex2, self.vertex3 = t("vertex1"), t("vertex2"), t("vertex3")
        self.normal, self.r = t("normal"), t("r")
        # mesh bookkeeping carried through densification (scene/mesh_based_gaussian_model.py:59-66); optional
        # ("triangles" is the synthetic generator's name, "vertex_index" the PLY loader's -- io.load_mesh_gaussian_ply)
        tri_key = "triangles" if "triangles" in arrays else ("vertex_index" if "vertex_index" in arrays else None)
        self.vertex_index = t(tri_key).long() if tri_key else None
        self.fid = t("face_id").long().view(-1, 1) if "face_id" in arrays else None
        self.v = t("mesh_vertices") if "mesh_vertices" in arrays else None
        self.alpha_distance = alpha_distance
        self.active_sh_degree = sh_degree
        self.max_sh_degree = 3
        self.screenspace_points = torch.zeros(self._bc.shape[0], 3, device=device, requires_grad=requires_grad)
        self._cache = None

    def parameters(self) -> List[torch.Tensor]:
        return [self._bc, self._distance, self._scaling, self._rotation, self._opacity, self._features]

    def to_arrays(self) -> Dict[str, np.ndarray]:
        """The arrays io.save_mesh_gaussian_ply writes (scene/mesh_based_gaussian_model.py:305-334): parameters, per-Gaussian
        face data and the mesh bookkeeping carried through densification."""
        c = lambda x: x.detach().cpu().numpy()
        P = self._bc.shape[0]
        with torch.no_grad():
            xyz = c(self.activate()[0]) if self._bc.is_cuda else np.zeros((P, 3), np.float32)
        tri = c(self.vertex_index).astype(np.int32) if self.vertex_index is not None else np.zeros((P, 3), np.int32)
        fid = c(self.fid).astype(np.int32).reshape(P, 1) if self.fid is not None else np.zeros((P, 1), np.int32)
        return {"xyz": xyz, "normal": c(self.normal), "bc_logits": c(self._bc), "vertex1": c(self.vertex1),
                "vertex2": c(self.vertex2), "vertex3": c(self.vertex3), "distance": c(self._distance), "vertex_index": tri,
                "r": c(self.r), "face_id": fid, "shs": c(self._features), "opacity_logit": c(self._opacity),
                "log_scales": c(self._scaling), "rot_raw": c(self._rotation)}

    def activate(self):
        """(xyz, scaling, rotation, opacity) -- one kernel, one autograd node."""
        return mg.mesh_bind(self._bc, self._distance, self._scaling, self._rotation, self._opacity, self.vertex1,
                            self.vertex2, self.vertex3, self.normal, self.r, self.alpha_distance)

    @property
    def get_xyz(self): return self.activate()[0]
    @property
    def get_scaling(self): return self.activate()[1]
    @property
    def get_rotation(self): return self.activate()[2]
    @property
    def get_opacity(self): return self.activate()[3]
    @property
    def get_features(self): return self._features
    @property
    def get_number(self): return self._bc.shape[0]


class GaussianModel:
    """Accessors of the plain (background) GaussianModel, scene/gaussian_model.py:96-116, with the three
    activations in one fused kernel."""

    def __init__(self, arrays: Dict[str, np.ndarray], device, sh_degree: int = 3, requires_grad: bool = True):
        t = lambda k: torch.from_numpy(np.ascontiguousarray(arrays[k])).to(device)
        self._xyz = t("means3D").requires_grad_(requires_grad)
        self._scaling = t("log_scales").requires_grad_(requires_grad)
        self._rotation = t("rot_raw" if "rot_raw" in arrays else "rotations").requires_grad_(requires_grad)
        self._opacity = t("opacity_logit").requires_grad_(requires_grad)
        self._features = t("shs").requires_grad_(requires_grad)
        self.active_sh_degree = sh_degree
        self.max_sh_degree = 3
        self.screenspace_points = torch.zeros(self._xyz.shape[0], 3, device=device, requires_grad=requires_grad)

    def parameters(self) -> List[torch.Tensor]:
        return [self._xyz, self._scaling, self._rotation, self._opacity, self._features]

    def activate(self):
        """(xyz, scaling, rotation, opacity)"""
        s, r, o = mg.activate(self._scaling, self._rotation, self._opacity)
        return self._xyz, s, r, o

    @property
    def get_xyz(self): return self._xyz
    @property
    def get_scaling(self): return self.activate()[1]
    @property
    def get_rotation(self): return self.activate()[2]
    @property
    def get_opacity(self): return self.activate()[3]
    @property
    def get_features(self): return self._features


@dataclass
class PipelineParams:
    """arguments/__init__.py:64-69.  The two *_python flags select the reference's "Python" pipeline variants
    (gaussian_renderer/__init__.py:78-94); here they run as one kernel each instead of a Jittor op chain."""
    convert_SHs_python: bool = False
    compute_cov3D_python: bool = False
    debug: bool = False


def _fresh_screenspace_points(pc, P: int, device) -> torch.Tensor:
    """reference gaussian_renderer/__init__.py:34 + reset_viewspace_point: a zero tensor whose gradient is the 2-D mean
    gradient of THIS render.  Models that did not opt into gradients (requires_grad=False) share their buffer."""
    pts = getattr(pc, "screenspace_points", None)
    if pts is not None and not pts.requires_grad:
        return pts
    return torch.zeros(P, 3, dtype=torch.float32, device=device, requires_grad=True)


def _python_color_and_cov(pc, pipe: PipelineParams, campos, scaling_modifier: float, means3D, scales, override_color):
    """The branch logic of gaussian_renderer/__init__.py:74-96 -> (shs, colors_precomp, scales, rotations, cov3D_precomp)."""
    cov3D_precomp = None
    rotations = None
    if pipe.compute_cov3D_python:
        # pc.get_covariance(scaling_modifier): the RAW rotation parameter, normalised inside (:79)
        cov3D_precomp = mg.covariance_from_scaling_rotation(scales, scaling_modifier, pc._rotation)
        scales_out = None
    else:
        scales_out = scales
    shs = colors_precomp = None
    if override_color is None:
        if pipe.convert_SHs_python:
            colors_precomp = mg.sh_to_rgb_rotated(means3D, campos, None, pc.get_features, pc.active_sh_degree)   # :87-92
        else:
            shs = pc.get_features
    else:
        colors_precomp = override_color
    return shs, colors_precomp, scales_out, rotations, cov3D_precomp


def render(viewpoint_camera, pc: MeshGaussianModel, pipe: PipelineParams, bg_color: torch.Tensor,
           scaling_modifier: float = 1.0, override_color: Optional[torch.Tensor] = None,
           bg_gaussian: Optional["GaussianModel"] = None, arena: Optional[RenderArena] = None) -> Dict[str, torch.Tensor]:
    """reference gaussian_renderer/__init__.py:26-143, every branch: CUDA or Python SH colours, CUDA or Python
    covariance, and the optional frozen `bg_gaussian` set appended as precomputed covariance (+ precomputed colours
    when the foreground has them), :100-121 -- which, as in the reference, needs compute_cov3D_python.

    Reference quirk kept by the glue: rasterize_points.py:154-158 derives the SH row length M from the SH tensor only when
    no precomputed covariance is passed, so compute_cov3D_python WITHOUT convert_SHs_python rasterizes with M = 0 (every
    Gaussian evaluates SH row 0) -- in the reference and here alike."""
    means3D, scales, rotations, opacity = pc.activate()
    P = means3D.shape[0]
    screenspace_points = _fresh_screenspace_points(pc, P, means3D.device)
    pc.screenspace_points = screenspace_points
    shs, colors_precomp, scales_r, _, cov3D_precomp = _python_color_and_cov(
        pc, pipe, viewpoint_camera.camera_center, scaling_modifier, means3D, scales, override_color)
    rotations_r = None if pipe.compute_cov3D_python else rotations
    if bg_gaussian is not None:
        if cov3D_precomp is None:
            # the reference concatenates onto cov3D_precomp = None here and fails inside jt.concat (:117)
            raise Exception("render(bg_gaussian=...) needs pipe.compute_cov3D_python: the background set is appended "
                            "as precomputed 3D covariance")
        with torch.no_grad():
            b_xyz, b_scale, b_rot, b_opacity = bg_gaussian.activate()
            b_cov = mg.covariance_from_scaling_rotation(b_scale, 1.0, b_rot)              # :101-104
            b_shs = bg_gaussian.get_features.detach()
        screenspace_points = torch.cat([screenspace_points, torch.zeros_like(b_xyz)], dim=0)      # :35-37
        if shs is not None:
            shs = torch.cat([shs, b_shs], dim=0)                                                  # :118-119
        else:
            with torch.no_grad():
                b_colors = mg.sh_to_rgb_rotated(b_xyz, viewpoint_camera.camera_center, None, b_shs, 3)   # :109-113
            colors_precomp = torch.cat([colors_precomp, b_colors], dim=0)                          # :121
        means3D = torch.cat([means3D, b_xyz.detach()], dim=0)
        opacity = torch.cat([opacity, b_opacity], dim=0)
        cov3D_precomp = torch.cat([cov3D_precomp, b_cov], dim=0)
        screenspace_points.retain_grad()
    raster_settings = make_settings(viewpoint_camera, bg_color, pc.active_sh_degree, scaling_modifier, pipe.debug)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings, arena=arena)
    rendered_image, radii = rasterizer(means3D=means3D, means2D=screenspace_points, shs=shs,
                                       colors_precomp=colors_precomp, opacities=opacity, scales=scales_r,
                                       rotations=rotations_r, cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii, "vertex1": pc.vertex1, "vertex2": pc.vertex2, "vertex3": pc.vertex3,
            "scale": scales_r}


def bg_render(viewpoint_camera, pc: GaussianModel, pipe: PipelineParams, bg_color: torch.Tensor,
              scaling_modifier: float = 1.0, override_color: Optional[torch.Tensor] = None,
              mesh_gaussians: Optional[MeshGaussianModel] = None, arena: Optional[RenderArena] = None
              ) -> Dict[str, torch.Tensor]:
    """reference gaussian_renderer/__init__.py:146-260: the background model is trainable, the mesh-bound
    Gaussians (if given) are appended with their gradients stopped (:223-234)."""
    means3D, scales, rotations, opacity = pc.activate()
    screenspace_points = _fresh_screenspace_points(pc, means3D.shape[0], means3D.device)
    pc.screenspace_points = screenspace_points
    shs, colors_precomp, scales_r, _, cov3D_precomp = _python_color_and_cov(
        pc, pipe, viewpoint_camera.camera_center, scaling_modifier, means3D, scales, override_color)
    rotations_r = None if pipe.compute_cov3D_python else rotations
    if mesh_gaussians is not None:
        if shs is None or cov3D_precomp is not None:
            raise Exception("mesh_gaussians can only be combined with the CUDA SH / covariance branches "
                            "(the reference concatenates SHs, scales and rotations, :229-233)")
        with torch.no_grad():
            m_xyz, m_scale, m_rot, m_opacity = mesh_gaussians.activate()
            m_shs = mesh_gaussians.get_features.detach()
        screenspace_points = torch.cat([screenspace_points, torch.zeros_like(m_xyz)], dim=0)
        screenspace_points.retain_grad()
        means3D = torch.cat([means3D, m_xyz], dim=0)
        scales_r = torch.cat([scales_r, m_scale], dim=0)
        rotations_r = torch.cat([rotations_r, m_rot], dim=0)
        shs = torch.cat([shs, m_shs], dim=0)
        opacity = torch.cat([opacity, m_opacity], dim=0)
    raster_settings = make_settings(viewpoint_camera, bg_color, pc.active_sh_degree, scaling_modifier, pipe.debug)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings, arena=arena)
    rendered_image, radii = rasterizer(means3D=means3D, means2D=screenspace_points, shs=shs,
                                       colors_precomp=colors_precomp, opacities=opacity, scales=scales_r,
                                       rotations=rotations_r, cov3D_precomp=cov3D_precomp)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}


# ---------------------------------------------------------------------------------------------
# edit-time object
# ---------------------------------------------------------------------------------------------
class DeformedObject:
    """SingleObjectDeform + ObjectVisualTool.render_gaussian for one mesh-bound object.

    load:    pos [P,3], cov [P,6] or [P,3,3], opacity [P,1], shs [P,16,3], triangles [P,3] (vertex ids of each
             Gaussian's face), weights [P,3] (barycentric, edittool/__init__.py:95-99), rest vertices [Vn,3]
    deform:  once per deformed mesh (edittool/__init__.py:103-131), one kernel
    render:  per frame: rotated-direction SH colour (one kernel) -> NewGaussianRasterizer with
             colors_precomp + cov3D_precomp (edittool/__init__.py:442-472)
    """

    def __init__(self, pos, cov, opacity, shs, triangles, weights, vertex_rest, device):
        dev = torch.device(device)
        f = lambda a: torch.as_tensor(a, dtype=torch.float32).contiguous().to(dev)
        self.pos, self.cov, self.opacity, self.shs = f(pos), f(cov), f(opacity), f(shs)
        self.weights, self.vertex = f(weights), f(vertex_rest)
        self.triangles = torch.as_tensor(triangles, dtype=torch.int32).contiguous().to(dev)
        P = self.pos.shape[0]
        # identity deformation until deform() is called (edittool/__init__.py:57-60)
        self.deform_pos = self.pos
        c = self.cov
        self.deform_cov6 = c if c.dim() == 2 else torch.stack(
            [c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], dim=1).contiguous()
        self.deform_rot = torch.eye(3, device=dev).expand(P, 3, 3).contiguous()
        self.device = dev

    @classmethod
    def load_mesh(cls, pos, cov, opacity, shs, proj_pos, face_id, vertex, faces, device) -> "DeformedObject":
        """SingleObjectDeform.load_gaussian + load_mesh (edittool/__init__.py:49-101, face-id branch): `face_id` [P] is the
        face every Gaussian is bound to (the PLY's fid), `proj_pos` its projection onto that face (get_proj_xyz),
        `vertex` / `faces` the rest mesh as igl.read_triangle_mesh returns it.  The per-Gaussian vertex ids and the
        area-ratio barycentric weights are computed on the device in float64 (gm_load_mesh)."""
        dev = torch.device(device)
        proj = torch.as_tensor(proj_pos, dtype=torch.float32).contiguous().to(dev)
        triangles, weights = mg.load_mesh(vertex, faces, face_id, proj)
        obj = cls(pos, cov, opacity, shs, triangles, weights.float(), vertex, dev)
        obj.coord = weights                      # float64, like self.coord in the reference
        obj.index_tri = torch.as_tensor(face_id).reshape(-1, 1)
        return obj

    def deform(self, vertex_deformed, vertex_R, vertex_S) -> None:
        f = lambda a: torch.as_tensor(a, dtype=torch.float32).contiguous().to(self.device)
        self.deform_pos, self.deform_cov6, self.deform_rot = mg.deform_gaussians(
            self.vertex, f(vertex_deformed), f(vertex_R), f(vertex_S), self.triangles, self.weights, self.pos, self.cov)

    def render_gaussian(self, cam, bg: torch.Tensor, arena: Optional[RenderArena] = None) -> torch.Tensor:
        colors = mg.sh_to_rgb_rotated(self.deform_pos, cam.camera_center, self.deform_rot, self.shs, 3)
        rasterizer = NewGaussianRasterizer(make_settings(cam, bg, 3), arena=arena)
        means2D = torch.zeros_like(self.deform_pos)
        with torch.no_grad():
            image, _ = rasterizer(means3D=self.deform_pos, means2D=means2D, shs=None, colors_precomp=colors,
                                  opacities=self.opacity, scales=None, rotations=None, cov3D_precomp=self.deform_cov6)
        return image


class SceneRenderer:
    """SceneVisualTool (edittool/__init__.py:133-231): a static background Gaussian set plus any number of
    deformable mesh-bound objects rendered together.

    The reference re-concatenates every attribute of every set on every frame, factorises the covariances
    back into (scale, quaternion) with a batched `eigh` plus a host-side determinant sign fix, and renders
    through the scale/rotation path with the un-rotated SHs (:181-219).  Sigma = V diag(lambda) V^T is the
    covariance it started from, so this class renders the same Gaussians through the precomputed-covariance path
    directly (SHs evaluated in CUDA, M forced to 16 as in rasterize_points_deformed.py): no eigh, no host
    round trip, and the concatenation is rebuilt only when an object is added or deformed."""

    def __init__(self, device, bg_means3D=None, bg_cov=None, bg_opacity=None, bg_shs=None):
        self.device = torch.device(device)
        self.objects: List[DeformedObject] = []
        f = lambda a: None if a is None else torch.as_tensor(a, dtype=torch.float32).contiguous().to(self.device)
        self.bg = None if bg_means3D is None else (f(bg_means3D), _pack_cov(f(bg_cov)), f(bg_opacity), f(bg_shs))
        self._cat = None
        self.arena = RenderArena(self.device, strict=False)

    def add_gaussian(self, obj: "DeformedObject") -> None:
        self.objects.append(obj)
        self._cat = None

    def deform_one_gaussian(self, index: int, vertex_deformed, vertex_R, vertex_S) -> None:
        self.objects[index].deform(vertex_deformed, vertex_R, vertex_S)
        self._cat = None

    def _gather(self):
        if self._cat is None:
            parts = ([self.bg] if self.bg is not None else []) + \
                    [(o.deform_pos, o.deform_cov6, o.opacity, o.shs) for o in self.objects]
            self._cat = tuple(torch.cat([p[i] for p in parts], dim=0).contiguous() for i in range(4))
        return self._cat

    def render_gaussian(self, cam, bg: Optional[torch.Tensor] = None) -> torch.Tensor:
        means3D, cov6, opacity, shs = self._gather()
        if bg is None:
            bg = torch.ones(3, dtype=torch.float32, device=self.device)      # edittool/__init__.py:166
        rasterizer = NewGaussianRasterizer(make_settings(cam, bg, 3), arena=self.arena)
        with torch.no_grad():
            image, _ = rasterizer(means3D=means3D, means2D=torch.zeros_like(means3D), shs=shs, colors_precomp=None,
                                  opacities=opacity, scales=None, rotations=None, cov3D_precomp=cov6)
        # retire THIS frame before handing the image out (one sync per frame is fine for the viewer): a view that
        # needs more instances than the arena held rendered only part of its tiles -- grow and render it again
        if self.arena.verify():
            return self.render_gaussian(cam, bg)
        return image


def _pack_cov(c: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """strip_symmetric (edittool/general_utils.py:26-37) for [P,3,3]; [P,6] passes through."""
    if c is None or c.dim() == 2:
        return c
    return torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], dim=1).contiguous()


# ---------------------------------------------------------------------------------------------
# sync-free batch rendering of independent views
# ---------------------------------------------------------------------------------------------
class ViewBatchRenderer:
    """Forward-render many views of one static Gaussian set on one GPU with no host synchronisation
    inside the loop.  Overflowed frames (arena high-water mark too low) are re-rendered after the
    batch, so results never depend on the arena size.

    `lanes` > 1 renders that many frames concurrently, each on its own CUDA stream with its own arena: the
    latency-bound stages of one frame (preprocess, emit, sort) overlap the issue-bound blend of another."""

    def __init__(self, device, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                 cov3D_precomp=None, sh_degree: int = 3, scale_modifier: float = 1.0, force_m: Optional[int] = None,
                 lanes: int = 1):
        self.device = torch.device(device)
        c = lambda t: None if t is None else t.detach().to(self.device, torch.float32).contiguous()
        self.means3D, self.opacities, self.shs = c(means3D), c(opacities), c(shs)
        self.colors, self.scales, self.rotations, self.cov = c(colors_precomp), c(scales), c(rotations), c(cov3D_precomp)
        if (self.shs is None) == (self.colors is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((self.scales is None or self.rotations is None) == (self.cov is None)):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        self.P = self.means3D.shape[0]
        self.D = sh_degree
        self.M = force_m if force_m is not None else (self.shs.shape[1] if self.shs is not None else 0)
        self.scale_modifier = scale_modifier
        self.lanes = max(1, int(lanes))
        self.arenas = [RenderArena(self.device, strict=False) for _ in range(self.lanes)]
        self.arena = self.arenas[0]
        self._radii = [torch.empty(self.P, dtype=torch.int32, device=self.device) for _ in range(self.lanes)]
        self.radii = self._radii[0]
        self._streams = [torch.cuda.Stream(self.device) for _ in range(self.lanes)] if self.lanes > 1 else [None]
        self._count = 0

    def _view_args(self, cam) -> tuple:
        p = lambda t: None if t is None else t.data_ptr()
        return (p(self.means3D), p(self.shs), p(self.colors), p(self.opacities), p(self.scales),
                float(self.scale_modifier), p(self.rotations), p(self.cov), cam.world_view_transform.data_ptr(),
                cam.full_proj_transform.data_ptr(), cam.camera_center.data_ptr(),
                math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5))

    def reserve_for(self, cams: Sequence, bg: torch.Tensor) -> int:
        """Size the arenas for these views up front (setup, not rendering)."""
        worst = self.arenas[0].reserve_for_views(self.P, self.D, self.M, bg, cams[0].image_width, cams[0].image_height,
                                                 [self._view_args(c) for c in cams])
        for a in self.arenas[1:]:
            a._ensure(self.P, cams[0].image_width * cams[0].image_height)
            a.high_water = worst
            a._want = self.arenas[0]._want
            a.reserve(a._want)
        return worst

    # fork / join of the side streams around a batch (no-ops with one lane)
    def begin_batch(self) -> None:
        if self.lanes > 1:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            for s in self._streams:
                s.wait_event(ev)

    def end_batch(self) -> None:
        if self.lanes > 1:
            cur = torch.cuda.current_stream(self.device)
            for s in self._streams:
                ev = torch.cuda.Event()
                ev.record(s)
                cur.wait_event(ev)

    def render_into(self, cam, bg: torch.Tensor, out: torch.Tensor) -> int:
        """Enqueue one view; `out` is a [3,H,W] float32 device tensor.  Returns the lane used."""
        lane = self._count % self.lanes
        self._count += 1
        va = self._view_args(cam)
        if self.lanes == 1:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            self.arenas[0].forward(self.P, self.D, self.M, bg, cam.image_width, cam.image_height, va, False, False,
                                   stream, out_color=out, radii=self._radii[0])
        else:
            with torch.cuda.stream(self._streams[lane]):
                self.arenas[lane].forward(self.P, self.D, self.M, bg, cam.image_width, cam.image_height, va, False, False,
                                          self._streams[lane].cuda_stream, out_color=out, radii=self._radii[lane])
        return lane

    def verify(self) -> List[Tuple[int, int]]:
        """(lane, frame-in-lane) of every frame that overflowed its arena; arenas are grown."""
        return [(l, f) for l, a in enumerate(self.arenas) for f in a.verify()]

    def render_views(self, cams: Sequence, bg: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Render all `cams` (DeviceCamera) into out [n,3,H,W]; returns it.  One sync at the end."""
        n = len(cams)
        if out is None:
            out = torch.empty(n, 3, cams[0].image_height, cams[0].image_width, dtype=torch.float32, device=self.device)
        first = [a.frames for a in self.arenas]
        where = {}
        self.begin_batch()
        for i, cam in enumerate(cams):
            lane = self._count % self.lanes
            where[(lane, self.arenas[lane].frames)] = i
            self.render_into(cam, bg, out[i])
        self.end_batch()
        redo = [where[k] for k in self.verify() if k in where]
        if redo:
            self.begin_batch()
            for i in redo:              # grown arenas; re-render the frames that did not fit
                self.render_into(cams[i], bg, out[i])
            self.end_batch()
            if self.verify():
                raise RuntimeError("arena overflow persisted after growth")   # cannot happen: grown to the high-water mark
        del first
        return out


# ---------------------------------------------------------------------------------------------
# config-4 training step without autograd bookkeeping
# ---------------------------------------------------------------------------------------------
class TrainStep:
    """forward + L1(target) + full backward to (means3D, shs, opacities, scales, rotations) through the
    C ABI, reusing every buffer across steps (train_mesh_gaussian.py:85-96 with the L1 term only;
    the optimizer step is excluded).  Gradients are left in `self.grads`."""

    def __init__(self, device, means3D, opacities, shs, scales, rotations, W: int, H: int, sh_degree: int = 3):
        self.device = torch.device(device)
        c = lambda t: t.detach().to(self.device, torch.float32).contiguous()
        self.means3D, self.opacities, self.shs, self.scales, self.rotations = map(c, (means3D, opacities, shs, scales, rotations))
        self.P, self.W, self.H, self.D, self.M = self.means3D.shape[0], W, H, sh_degree, self.shs.shape[1]
        P, M = self.P, self.M
        self.arena = RenderArena(self.device, strict=False)
        self.overflowed_frames = 0          # frames whose gradients are partial (size the arena with reserve_for)
        self.image = torch.empty(3, H, W, dtype=torch.float32, device=self.device)
        self.dL_dimg = torch.empty_like(self.image)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.radii = torch.empty(P, dtype=torch.int32, device=self.device)
        # accumulated (atomically added) gradients first: only this part of the slab is zeroed per step; the
        # per-Gaussian rows behind it are fully written by gm_backward_ex(GM_BACKWARD_OVERWRITE)
        sizes = {"means2D": 3 * P, "conic": 4 * P, "opacity": P, "colors": 3 * P,
                 "means3D": 3 * P, "cov3D": 6 * P, "sh": 3 * M * P, "scales": 3 * P, "rotations": 4 * P}
        offs, total = {}, 0
        for k, s in sizes.items():
            offs[k] = total
            total += ((s + 31) // 32) * 32
        self._slab = torch.empty(total, dtype=torch.float32, device=self.device)
        self._accum = self._slab[:offs["means3D"]]
        self.grads = {k: self._slab[offs[k]:offs[k] + sizes[k]] for k in sizes}

    def _view_args(self, cam) -> tuple:
        p = lambda t: t.data_ptr()
        return (p(self.means3D), p(self.shs), None, p(self.opacities), p(self.scales), 1.0, p(self.rotations), None,
                p(cam.world_view_transform), p(cam.full_proj_transform), p(cam.camera_center),
                math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5))

    def reserve_for(self, cams: Sequence, bg: torch.Tensor) -> int:
        """Size the arena for these views up front (setup, not rendering)."""
        return self.arena.reserve_for_views(self.P, self.D, self.M, bg, self.W, self.H, [self._view_args(c) for c in cams])

    def verify(self) -> int:
        """Block until every submitted frame is done; returns how many frames overflowed the arena so far (their
        gradients cover only the tiles that fit).  0 after reserve_for() on the views being trained."""
        self.overflowed_frames += len(self.arena.verify())
        if getattr(self, "_graph", None) is not None:
            torch.cuda.synchronize(self.device)
            self.overflowed_frames += int(self._graph_overflow.item())
            self._graph_overflow.zero_()
        return self.overflowed_frames

    # ------------------------------------------------------------------------------------------
    # the same step as ONE CUDA graph launch
    # ------------------------------------------------------------------------------------------
    def capture(self, cam, bg: torch.Tensor, target: torch.Tensor) -> None:
        """Capture step() into a CUDA graph (the programmatic-launch edges between the kernels are kept as graph edges).
        The graph reads the camera from a static 35-float buffer and the target from a static buffer of `target`'s dtype
        and shape; step_graph() copies a frame's camera / target into them and replays.  Intrinsics (FoV) are kernel
        arguments and therefore fixed at capture time; the arena must already be sized (reserve_for)."""
        if self.arena.capacity == 0:
            raise RasterizerError("TrainStep.capture", GM_ERR_BAD_ARGUMENT, "size the arena first (reserve_for)")
        self._g_cam = torch.zeros(35, dtype=torch.float32, device=self.device)
        self._g_cam.copy_(torch.cat([cam.world_view_transform.reshape(-1), cam.full_proj_transform.reshape(-1),
                                     cam.camera_center.reshape(-1)]))
        self._g_target = torch.empty_like(target)
        self._g_target.copy_(target)
        self._g_fov = (float(cam.FoVx), float(cam.FoVy))
        self._g_view = DeviceCamera(cam.image_width, cam.image_height, cam.FoVx, cam.FoVy, self._g_cam[0:16].view(4, 4),
                                    self._g_cam[16:32].view(4, 4), self._g_cam[32:35])
        self._graph_overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._g_bg = bg
        self.step(self._g_view, bg, self._g_target)          # warm-up outside the capture (module loading, attributes)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._step_captured(self._g_view, bg, self._g_target)

    def _step_captured(self, cam, bg, target) -> None:
        """step() without the arena's event bookkeeping (events recorded during capture cannot be queried): fixed capacity,
        overflow accumulated on the device."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        va = self._view_args(cam)
        a = self.arena
        check(lib.gm_forward_ex(a.geom.data_ptr(), a.binning.data_ptr(), a.binning.numel(), a.image.data_ptr(), self.P, self.D,
                                self.M, bg.data_ptr(), self.W, self.H, *va, 0, self.image.data_ptr(), self.radii.data_ptr(), 0,
                                None, self._epilogue(target), stream), "gm_forward")
        g = self.grads
        check(lib.gm_backward_ex(self.P, self.D, self.M, a.capacity, bg.data_ptr(), self.W, self.H, va[0], va[1], None, va[4], 1.0,
                                 va[6], None, va[8], va[9], va[10], va[11], va[12], self.radii.data_ptr(), a.geom.data_ptr(),
                                 a.binning.data_ptr(), a.image.data_ptr(), self.dL_dimg.data_ptr(), g["means2D"].data_ptr(),
                                 g["conic"].data_ptr(), g["opacity"].data_ptr(), g["colors"].data_ptr(), g["means3D"].data_ptr(),
                                 g["cov3D"].data_ptr(), g["sh"].data_ptr(), g["scales"].data_ptr(), g["rotations"].data_ptr(), 0,
                                 GM_BACKWARD_OVERWRITE, stream), "gm_backward_ex")
        # frame header word 2 = overflow flag of this frame (gm_frame_overflow_flag)
        hdr = a.geom[:16].view(torch.int32)
        self._graph_overflow.add_(hdr[2:3])

    def step_graph(self, cam, bg: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """One training step as a single graph launch: two small device copies (camera, target) + cudaGraphLaunch."""
        if getattr(self, "_graph", None) is None:
            self.capture(cam, bg, target)
        if (float(cam.FoVx), float(cam.FoVy)) != self._g_fov or bg.data_ptr() != self._g_bg.data_ptr():
            raise RasterizerError("TrainStep.step_graph", GM_ERR_BAD_ARGUMENT,
                                  "the graph was captured for other intrinsics / another background tensor: capture() again")
        self._g_cam[0:16].copy_(cam.world_view_transform.reshape(-1), non_blocking=True)
        self._g_cam[16:32].copy_(cam.full_proj_transform.reshape(-1), non_blocking=True)
        self._g_cam[32:35].copy_(cam.camera_center.reshape(-1), non_blocking=True)
        if target.data_ptr() != self._g_target.data_ptr():
            self._g_target.copy_(target, non_blocking=True)
        self._graph.replay()
        return self.loss

    def _epilogue(self, target: torch.Tensor) -> ForwardEpilogue:
        """The L1 loss against `target` (uint8 = the 8-bit ground-truth image, value / 255 inside the kernel; float32 as it is)
        and the clearing of the accumulated gradient buffers, both folded into the blend kernel (gm_forward_ex)."""
        return ForwardEpilogue(target.data_ptr(), int(target.dtype == torch.uint8), self.loss.data_ptr(), self.dL_dimg.data_ptr(),
                               self._accum.data_ptr(), self._accum.numel())

    def step(self, cam, bg: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """Enqueue one training step; returns the (device) loss tensor.  No host synchronisation."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        va = self._view_args(cam)
        cap, _, _, geom, binning, image_state = self.arena.forward(
            self.P, self.D, self.M, bg, self.W, self.H, va, False, False, stream, out_color=self.image, radii=self.radii,
            epilogue=self._epilogue(target))
        if self.arena.overflowed:          # frames seen to have overflowed since the last step (arena already grown)
            self.overflowed_frames += len(self.arena.overflowed)
            self.arena.overflowed.clear()
        g = self.grads
        check(lib.gm_backward_ex(self.P, self.D, self.M, cap, bg.data_ptr(), self.W, self.H, va[0], va[1], None, va[4], 1.0,
                              va[6], None, va[8], va[9], va[10], va[11], va[12], self.radii.data_ptr(),
                              geom.data_ptr(), binning.data_ptr(), image_state.data_ptr(), self.dL_dimg.data_ptr(),
                              g["means2D"].data_ptr(), g["conic"].data_ptr(), g["opacity"].data_ptr(),
                              g["colors"].data_ptr(), g["means3D"].data_ptr(), g["cov3D"].data_ptr(),
                                 g["sh"].data_ptr(), g["scales"].data_ptr(), g["rotations"].data_ptr(), 0,
                                 GM_BACKWARD_OVERWRITE, stream),
              "gm_backward_ex")
        return self.loss
