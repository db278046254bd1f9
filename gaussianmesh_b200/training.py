"""The rest of one training iteration around the rasterizer op.

  l1_loss / ssim / photometric_loss   utils/loss_utils.py:17-18,36-82 and the combination of
                                      train_mesh_gaussian.py:91,94, fused with their gradient (two launches)
  mesh_restrict_loss                  utils/loss_utils.py:84-107
  get_expon_lr_func                   utils/general_utils.py:29-62 (host arithmetic)
  Adam                                the jt.nn.Adam(l, lr=0.0, eps=1e-15) of
                                      scene/mesh_based_gaussian_model.py:248-258: every group in ONE launch
  TrainingIteration                   train_mesh_gaussian.py:73-147 for a fixed set of Gaussians: learning-rate
                                      schedule, bind + render, loss, backward, densification statistics, Adam --
                                      through the C ABI on reused buffers, no autograd tape, no host round trip

torch is plumbing (device memory, streams, autograd hooks); all arithmetic is in libCudaRasterizer.so.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from ._lib import lib, check, AdamTensor, RasterizerError, GM_ERR_BAD_ARGUMENT, GM_BACKWARD_OVERWRITE, ForwardEpilogue
from .arena import RenderArena
from .mesh_gaussians import _c, _p, _stream, l1_loss  # noqa: F401  (l1_loss re-exported)


# ---------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------
def _photometric(img: torch.Tensor, gt: torch.Tensor, lambda_dssim: float, want_grad: bool):
    img, gt = _c(img), _c(gt)
    if img.dim() == 4 and img.shape[0] == 1:
        img, gt = img[0], gt[0]
    if img.dim() != 3 or img.shape != gt.shape:
        raise RasterizerError("photometric_loss", GM_ERR_BAD_ARGUMENT, "expected two [C,H,W] images of the same shape")
    Cc, H, W = img.shape
    scratch = torch.empty(lib.gm_photometric_scratch_bytes(Cc, H, W), dtype=torch.uint8, device=img.device)
    out = torch.empty(3, dtype=torch.float32, device=img.device)
    grad = torch.empty_like(img) if want_grad else None
    check(lib.gm_photometric_loss(Cc, H, W, _p(img), _p(gt), float(lambda_dssim), _p(scratch), _p(out), _p(grad),
                                  _stream()), "gm_photometric_loss")
    return out, grad


class _Photometric(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, gt, lambda_dssim):
        out, grad = _photometric(img, gt, lambda_dssim, True)
        ctx.save_for_backward(grad)
        ctx.shape = img.shape
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g, _g_parts):
        (grad,) = ctx.saved_tensors
        return (grad * g).view(ctx.shape), None, None


def photometric_loss(image: torch.Tensor, gt_image: torch.Tensor, lambda_dssim: float = 0.2, return_parts: bool = False):
    """(1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))
    (train_mesh_gaussian.py:91,94 without the mrloss term), differentiable w.r.t. `image`.
    With return_parts also returns the device tensor (loss, L1, SSIM)."""
    loss, parts = _Photometric.apply(image, gt_image, float(lambda_dssim))
    return (loss, parts) if return_parts else loss


class _SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2):
        out, grad = _photometric(img1, img2, 1.0, True)     # loss = 1 - ssim
        ctx.save_for_backward(grad)
        ctx.shape = img1.shape
        return out[2].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * (-g)).view(ctx.shape), None


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """utils/loss_utils.py:36-82; differentiable w.r.t. img1 (the rendered image).  Only the configuration the
    training loop uses (11-tap window, mean over everything) exists on the device."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("ssim: only window_size=11, size_average=True (train_mesh_gaussian.py:94) is implemented")
    return _SSIM.apply(img1, img2)


class _MeshRestrict(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scale, point1, point2, point3, weight):
        scale, point1, point2, point3 = map(_c, (scale, point1, point2, point3))
        P = scale.shape[0]
        loss = torch.empty(1, dtype=torch.float32, device=scale.device)
        grad = torch.empty_like(scale)
        check(lib.gm_mesh_restrict_loss(P, _p(scale), _p(point1), _p(point2), _p(point3), float(weight), _p(loss),
                                        _p(grad), 0, _stream()), "gm_mesh_restrict_loss")
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None


def mesh_restrict_loss(scale, point1, point2, point3, weight: float = 10) -> torch.Tensor:
    """utils/loss_utils.py:102-107: sum(clamp(max(scale, dim=1) - weight * circumradius, min=0)); differentiable w.r.t.
    `scale` (the vertices are constants of the model)."""
    return _MeshRestrict.apply(scale, point1, point2, point3, weight)


# ---------------------------------------------------------------------------------------------
# schedule + optimizer
# ---------------------------------------------------------------------------------------------
def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """utils/general_utils.py:29-62: log-linear interpolation from lr_init to lr_final with an optional
    sine-eased delay."""
    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        if lr_delay_steps > 0:
            delay_rate = lr_delay_mult + (1 - lr_delay_mult) * np.sin(0.5 * np.pi * np.clip(step / lr_delay_steps, 0, 1))
        else:
            delay_rate = 1.0
        t = np.clip(step / max_steps, 0, 1)
        log_lerp = np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t)
        return delay_rate * log_lerp
    return helper


@dataclass
class OptimizationParams:
    """arguments/__init__.py:71-91 (defaults)."""
    iterations: int = 30_000
    position_lr_init: float = 0.00016
    position_lr_final: float = 0.0000016
    position_lr_delay_mult: float = 0.01
    position_lr_max_steps: int = 30_000
    feature_lr: float = 0.0025
    opacity_lr: float = 0.05
    scaling_lr: float = 0.005
    rotation_lr: float = 0.001
    percent_dense: float = 0.01
    lambda_dssim: float = 0.2
    densification_interval: int = 200
    opacity_reset_interval: int = 3000
    densify_from_iter: int = 500
    densify_until_iter: int = 15_000
    densify_grad_threshold: float = 0.0002
    random_background: bool = True
    alpha_mrloss: float = 6


class Adam:
    """jt.nn.Adam(param_groups, lr, eps, betas) as the reference uses it
    (scene/mesh_based_gaussian_model.py:248-258): per-group learning rates, dense update of every element.
    `param_groups` is a list of {'params': [tensor, ...], 'lr': float, 'name': str}; a group may add
    'lr_head', 'period', 'split' to give elements with (index mod period) < split their own rate.
    step() updates every tensor of every group with ONE kernel launch."""

    def __init__(self, param_groups: Sequence[dict], lr: float = 0.0, eps: float = 1e-8, betas=(0.9, 0.999)):
        self.lr, self.eps, self.betas = lr, eps, betas
        self.param_groups: List[dict] = []
        self.n_step = 0
        self.state: Dict[int, Dict[str, torch.Tensor]] = {}
        for g in param_groups:
            self.add_param_group(g)

    def add_param_group(self, group: dict) -> None:
        g = dict(group)
        g["params"] = list(g["params"])
        for p in g["params"]:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RasterizerError("Adam", GM_ERR_BAD_ARGUMENT, "parameters must be contiguous float32 CUDA tensors")
            self.state[id(p)] = {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
        self.param_groups.append(g)

    def zero_grad(self) -> None:
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None

    def step(self, grads: Optional[Dict[int, torch.Tensor]] = None, skip_flag: Optional[int] = None) -> None:
        """`grads` maps id(param) -> gradient tensor; default is param.grad.  Parameters without a gradient are
        skipped (a stop_grad parameter in the reference).  `skip_flag` is the device address of a uint32 that, when
        non-zero at execution time, turns the launch into a no-op (gm_frame_overflow_flag: a frame that overflowed its
        arena must not reach the parameters)."""
        self.n_step += 1
        rows, keep = [], []
        for g in self.param_groups:
            lr = float(g.get("lr", self.lr))
            for p in g["params"]:
                gr = p.grad if grads is None else grads.get(id(p))
                if gr is None:
                    continue
                gr = _c(gr)
                if gr.numel() != p.numel():
                    raise RasterizerError("Adam", GM_ERR_BAD_ARGUMENT, "gradient / parameter size mismatch")
                keep.append(gr)
                st = self.state[id(p)]
                rows.append(AdamTensor(p.data_ptr(), gr.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                       p.numel(), lr, float(g.get("lr_head", lr)), int(g.get("period", 0)),
                                       int(g.get("split", 0))))
        if not rows:
            return
        table = (AdamTensor * len(rows))(*rows)
        check(lib.gm_adam_step_gated(len(rows), table, self.n_step, float(self.betas[0]), float(self.betas[1]),
                                     float(self.eps), skip_flag, _stream()), "gm_adam_step")


# ---------------------------------------------------------------------------------------------
# one training iteration on reused buffers
# ---------------------------------------------------------------------------------------------
class TrainingIteration:
    """train_mesh_gaussian.py:73-147 for a fixed set of mesh-bound Gaussians (no densify_and_prune, which changes
    the tensor shapes): update_learning_rate -> bind + activations -> render -> (1-l) L1 + l (1-SSIM) + mrloss ->
    backward through the rasterizer and the binding -> densification statistics -> Adam.  Twelve kernel stages per
    iteration, nothing allocated, nothing read back.  `model` is a renderer.MeshGaussianModel; its parameter
    tensors are updated in place."""

    # parameter tensors in the order of the flat parameter / gradient vectors: (model attribute, gradient name, row width)
    PARAMS = (("_features", "sh", None), ("_bc", "bc", 3), ("_distance", "distance", 1), ("_scaling", "log_scale", 3),
              ("_rotation", "rot_raw", 4), ("_opacity", "opacity_logit", 1))

    def __init__(self, model, opt: OptimizationParams, W: int, H: int, spatial_lr_scale: float = 1.0,
                 alloc=None, flat_params: bool = False):
        """`alloc(numel)` returns a float32 device vector (default torch.empty); the view-parallel trainer passes a
        symmetric-memory allocator.  The gradients of the six parameter tensors always live in ONE flat vector
        (`param_grads`, tensors at 32-float aligned offsets: `layout`); with flat_params the parameters themselves
        are moved into a vector of the same layout (`flat_parameters`) and the model's tensors become views of it."""
        self.model, self.opt, self.W, self.H = model, opt, W, H
        self.device = model._bc.device
        self.spatial_lr_scale = spatial_lr_scale
        self._alloc, self._flat = alloc, flat_params
        self.iteration = 0
        self.skipped_iterations = 0       # iterations whose frame overflowed the arena (update dropped on the device)
        self._allocate()

    def _allocate(self) -> None:
        """Buffers, gradient layout and a fresh optimizer for the model's CURRENT number of Gaussians (construction,
        and again after densify_and_prune changed it)."""
        model, opt, W, H, dev = self.model, self.opt, self.W, self.H, self.device
        alloc, flat_params, spatial_lr_scale = self._alloc, self._flat, self.spatial_lr_scale
        P, M = model._bc.shape[0], model._features.shape[1]
        self.P, self.M = P, M
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        if alloc is None:
            alloc = lambda n: torch.empty(n, dtype=torch.float32, device=dev)
        self.layout, total = {}, 0
        for attr, gname, width in self.PARAMS:
            n = getattr(model, attr).numel()
            self.layout[gname] = (total, n)
            total += ((n + 31) // 32) * 32
        self.flat_numel = total
        self.param_grads = alloc(total)
        self.param_grads.zero_()
        self.flat_parameters = None
        if flat_params:
            self.flat_parameters = alloc(total)
            self.flat_parameters.zero_()
            for attr, gname, _ in self.PARAMS:
                off, n = self.layout[gname]
                old = getattr(model, attr)
                view = self.flat_parameters[off:off + n].view(old.shape)
                view.copy_(old.detach())
                setattr(model, attr, view)
        self.xyz, self.scale, self.rot, self.opacity = f(P, 3), f(P, 3), f(P, 4), f(P, 1)
        self.image, self.dL_dimg = f(3, H, W), f(3, H, W)
        self.radii = torch.empty(P, dtype=torch.int32, device=dev)
        self.losses = torch.zeros(4, dtype=torch.float32, device=dev)     # loss w/o mrloss, L1, SSIM, mrloss
        self.scratch = torch.empty(lib.gm_photometric_scratch_bytes(3, H, W), dtype=torch.uint8, device=dev)
        self.arena = RenderArena(dev, headroom=1.5, strict=False)     # the parameters move: instance counts drift
        # gradient slab: the atomically accumulated part first (zeroed per step), overwritten rows behind it
        sizes = {"means2D": 3 * P, "conic": 4 * P, "opacity": P, "colors": 3 * P,
                 "means3D": 3 * P, "cov3D": 6 * P, "scales": 3 * P, "rotations": 4 * P}
        offs, total = {}, 0
        for k, s in sizes.items():
            offs[k] = total
            total += ((s + 31) // 32) * 32
        self._slab = torch.empty(total, dtype=torch.float32, device=dev)
        self._accum = self._slab[:offs["means3D"]]
        self.grads = {k: self._slab[offs[k]:offs[k] + sizes[k]] for k in sizes}
        for gname, (off, n) in self.layout.items():
            self.grads[gname] = self.param_grads[off:off + n]
        # densification statistics (scene/mesh_based_gaussian_model.py:241,245-246)
        self.max_radii2D = torch.zeros(P, dtype=torch.float32, device=dev)
        self.bc_gradient_accum = torch.zeros(P, 1, dtype=torch.float32, device=dev)
        self.denom = torch.zeros(P, 1, dtype=torch.float32, device=dev)
        # optimizer (scene/mesh_based_gaussian_model.py:248-262); f_dc / f_rest share one [P,16,3] tensor
        pos_lr = opt.position_lr_init * spatial_lr_scale
        self.optimizer = Adam([
            {"params": [model._bc], "lr": pos_lr, "name": "bc"},
            {"params": [model._distance], "lr": pos_lr, "name": "distance"},
            {"params": [model._features], "lr": opt.feature_lr / 20.0, "lr_head": opt.feature_lr, "period": 3 * M,
             "split": 3, "name": "features"},
            {"params": [model._opacity], "lr": opt.opacity_lr, "name": "opacity"},
            {"params": [model._scaling], "lr": opt.scaling_lr, "name": "scaling"},
            {"params": [model._rotation], "lr": opt.rotation_lr, "name": "rotation"},
        ], lr=0.0, eps=1e-15)
        self.bc_scheduler_args = get_expon_lr_func(lr_init=pos_lr, lr_final=opt.position_lr_final * spatial_lr_scale,
                                                   lr_delay_mult=opt.position_lr_delay_mult,
                                                   max_steps=opt.position_lr_max_steps)
        g = self.grads
        self._grad_of = {id(model._bc): g["bc"], id(model._distance): g["distance"], id(model._features): g["sh"],
                         id(model._opacity): g["opacity_logit"], id(model._scaling): g["log_scale"],
                         id(model._rotation): g["rot_raw"]}

    # ------------------------------------------------------------------------------------------
    # densification (scene/mesh_based_gaussian_model.py:482-585): infrequent model surgery -- every
    # densification_interval iterations -- done with tensor indexing; it is data movement, not arithmetic
    # ------------------------------------------------------------------------------------------
    def densify_and_prune(self, max_grad: float, min_opacity: float = 0.005, extent: float = 0.0,
                          max_screen_size=None, N: int = 4) -> int:
        """reference :581-585: grads = bc_gradient_accum / denom (NaN -> 0), then densify_and_split.  (The reference
        ignores min_opacity / extent / max_screen_size as well.)  Returns the number of Gaussians that were split."""
        grads = self.bc_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        return self.densify_and_split(grads, max_grad, extent, N)

    def densify_and_split(self, grads: torch.Tensor, grad_threshold: float, scene_extent: float = 0.0, N: int = 4) -> int:
        """reference :504-563 + utils/general_utils.py:133-212: every Gaussian whose mean screen-space gradient reaches
        the threshold is replaced by N children on the 4-way split of its face (N = 5 keeps a copy on the parent face):
        barycentric logits 1/3, distance 0, scale / 3.2, everything else inherited; the three edge midpoints become new
        mesh vertices; the survivors keep their Adam moments, children start at zero; statistics restart."""
        if N not in (4, 5):
            raise ValueError("N must be 4 or 5 (split_mesh_and_gaussian / split_mesh_and_gaussian_pro)")
        # with flat_params the new tensors are moved into a freshly allocated flat vector by _allocate() below (for a
        # symmetric-memory model that is a collective: every rank must densify with the same `grads`, which the merged
        # statistics of view_parallel guarantee); the sharded Adam moments of that mode are carried by
        # ViewParallelTrainer.densify_and_prune, the ones handled here are the replicated optimizer's
        m = self.model
        sel = grads.reshape(-1) >= grad_threshold
        S = int(sel.sum().item())                       # the reference reads this back too (:514)
        if S == 0:
            return 0
        keep = ~sel
        rep = lambda t: t[sel].repeat_interleave(N, dim=0)
        dev = self.device
        a, b, c = m.vertex1[sel], m.vertex2[sel], m.vertex3[sel]
        ab, ac, bc_ = (a + b) / 2, (a + c) / 2, (b + c) / 2
        # utils/general_utils.py:137-151 (+191-193 for N = 5): child k's triangle
        v1 = [a, ab, ac, ab] + ([a] if N == 5 else [])
        v2 = [ab, b, bc_, bc_] + ([b] if N == 5 else [])
        v3 = [ac, bc_, c, ac] + ([c] if N == 5 else [])
        stack = lambda parts: torch.stack(parts, dim=1).reshape(S * N, 3)
        new = {
            "_bc": torch.full((S * N, 3), 1.0 / 3.0, dtype=torch.float32, device=dev),
            "_distance": torch.zeros(S * N, 1, dtype=torch.float32, device=dev),
            "_features": rep(m._features), "_opacity": rep(m._opacity), "_rotation": rep(m._rotation),
            "_scaling": torch.log(rep(torch.exp(m._scaling)) / (4 * 0.8)),
        }
        const = {"vertex1": stack(v1), "vertex2": stack(v2), "vertex3": stack(v3), "normal": rep(m.normal), "r": rep(m.r)}
        if getattr(m, "fid", None) is not None:
            const["fid"] = rep(m.fid)
        if getattr(m, "vertex_index", None) is not None:
            v_origin_num = int(m.v.shape[0]) if getattr(m, "v", None) is not None else int(m.vertex_index.max().item()) + 1
            idx = m.vertex_index[sel].unsqueeze(1).repeat(1, N, 1)                       # [S,N,3]
            tmp = torch.arange(S * 3, device=dev, dtype=idx.dtype).view(S, 3) + v_origin_num
            idx[:, 0, 1], idx[:, 0, 2] = tmp[:, 0], tmp[:, 1]                           # utils/general_utils.py:160-168
            idx[:, 1, 0], idx[:, 1, 2] = tmp[:, 0], tmp[:, 2]
            idx[:, 2, 0], idx[:, 2, 1] = tmp[:, 1], tmp[:, 2]
            idx[:, 3, 0], idx[:, 3, 1], idx[:, 3, 2] = tmp[:, 0], tmp[:, 2], tmp[:, 1]
            const["vertex_index"] = idx.reshape(S * N, 3)
            if getattr(m, "v", None) is not None:
                m.v = torch.cat([m.v, torch.stack([ab, ac, bc_], dim=1).reshape(S * 3, 3)], dim=0)   # :153-155
        # survivors first, children behind them (concat then prune_points, :547-563)
        old_state = {attr: self.optimizer.state[id(getattr(m, attr))] for attr, _, _ in self.PARAMS}
        moments = {}
        for attr, _, _ in self.PARAMS:
            old = getattr(m, attr)
            st = old_state[attr]
            moments[attr] = (torch.cat([st["exp_avg"][keep], torch.zeros_like(new[attr])], dim=0),
                             torch.cat([st["exp_avg_sq"][keep], torch.zeros_like(new[attr])], dim=0))
            setattr(m, attr, torch.cat([old.detach()[keep], new[attr]], dim=0).contiguous())
        for name, val in const.items():
            setattr(m, name, torch.cat([getattr(m, name)[keep], val], dim=0).contiguous())
        m.screenspace_points = torch.zeros(m._bc.shape[0], 3, device=dev,
                                           requires_grad=bool(getattr(m.screenspace_points, "requires_grad", False)))
        n_step, lr_groups = self.optimizer.n_step, {g["name"]: g["lr"] for g in self.optimizer.param_groups}
        persistent = {k: getattr(self, k) for k in ("arena", "losses", "image", "dL_dimg", "scratch")}
        self._allocate()                                 # buffers for the new size, statistics restart at zero (:499-501)
        for k, v in persistent.items():
            setattr(self, k, v)
        self.optimizer.n_step = n_step
        for g in self.optimizer.param_groups:
            g["lr"] = lr_groups[g["name"]]
        for attr, _, _ in self.PARAMS:
            st = self.optimizer.state[id(getattr(m, attr))]
            st["exp_avg"], st["exp_avg_sq"] = moments[attr][0].contiguous(), moments[attr][1].contiguous()
        return S

    def reset_opacity(self) -> None:
        """reference :334-339: opacity <- inverse_sigmoid(min(sigmoid(opacity), 0.01)), its Adam moments zeroed
        (replace_tensor_to_optimizer, :411-422).  In place: the tensor keeps its address (and its place in a flat vector)."""
        m = self.model
        x = torch.minimum(torch.sigmoid(m._opacity), torch.full_like(m._opacity, 0.01))
        m._opacity.copy_(torch.log(x / (1 - x)))
        st = self.optimizer.state[id(m._opacity)]
        st["exp_avg"].zero_()
        st["exp_avg_sq"].zero_()

    def after_backward(self, white_background: bool = False) -> bool:
        """The bookkeeping of train_mesh_gaussian.py:114-131 that follows the statistics: densify every
        densification_interval iterations, reset the opacities every opacity_reset_interval.  Returns the reference's
        `update_flag` (True: this iteration's optimizer step is skipped, :139)."""
        it, opt = self.iteration, self.opt
        update_flag = False
        if it < opt.densify_until_iter:
            if it > opt.densify_from_iter and it % opt.densification_interval == 0:
                self.densify_and_prune(opt.densify_grad_threshold, 0.005, 0.0, 20 if it > opt.opacity_reset_interval else None, 5)
                update_flag = True
            if it % opt.opacity_reset_interval == 0 or (white_background and it == opt.densify_from_iter):
                self.reset_opacity()
        return update_flag

    # scene/mesh_based_gaussian_model.py:280-288
    def update_learning_rate(self, iteration: int) -> float:
        lr = float(self.bc_scheduler_args(iteration))
        for g in self.optimizer.param_groups:
            if g["name"] in ("bc", "distance"):
                g["lr"] = lr
        return lr

    def _bind(self, stream) -> None:
        m = self.model
        check(lib.gm_mesh_bind_forward(self.P, _p(m._bc), _p(m._distance), _p(m.vertex1), _p(m.vertex2), _p(m.vertex3),
                                       _p(m.normal), _p(m.r), float(m.alpha_distance), _p(m._scaling), _p(m._rotation),
                                       _p(m._opacity), _p(self.xyz), _p(self.scale), _p(self.rot), _p(self.opacity),
                                       stream), "gm_mesh_bind_forward")

    def _view_args(self, cam) -> tuple:
        p = lambda t: t.data_ptr()
        return (p(self.xyz), p(self.model._features), None, p(self.opacity), p(self.scale), 1.0, p(self.rot), None,
                p(cam.world_view_transform), p(cam.full_proj_transform), p(cam.camera_center),
                math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5))

    def reserve_for(self, cams: Sequence, bg: torch.Tensor) -> int:
        """Size the arena for these views with the current parameters (setup, not training)."""
        self._bind(torch.cuda.current_stream(self.device).cuda_stream)
        return self.arena.reserve_for_views(self.P, self.model.active_sh_degree, self.M, bg, self.W, self.H,
                                            [self._view_args(c) for c in cams])

    def adam_segments(self):
        """The flat layout as (offset, numel, lr, lr_head, period, split) rows with the CURRENT learning rates."""
        by_attr = {id(p): g for g in self.optimizer.param_groups for p in g["params"]}
        rows = []
        for attr, gname, _ in self.PARAMS:
            g = by_attr[id(getattr(self.model, attr))]
            off, n = self.layout[gname]
            lr = float(g["lr"])
            rows.append((off, n, lr, float(g.get("lr_head", lr)), int(g.get("period", 0)), int(g.get("split", 0))))
        return rows

    def step(self, cam, bg: torch.Tensor, gt_image: torch.Tensor, iteration: Optional[int] = None,
             optimizer_step: bool = True, densify: bool = False) -> torch.Tensor:
        """Enqueue one iteration; returns the device tensor (photometric loss, L1, SSIM, mrloss) -- the reference's
        `loss` is [0] + [3].  No host synchronisation.  With optimizer_step=False the iteration stops after the
        backward pass and the statistics (gradients in `param_grads`): the view-parallel trainer exchanges them.
        With densify=True the densification / opacity-reset schedule of the reference runs after the statistics (one
        host read-back on the iterations that densify, like the reference's)."""
        m, opt = self.model, self.opt
        self.iteration = self.iteration + 1 if iteration is None else iteration
        it = self.iteration
        self.update_learning_rate(it)
        if it % 1000 == 0 and m.active_sh_degree < m.max_sh_degree:          # train_mesh_gaussian.py:80-81
            m.active_sh_degree += 1
        stream = torch.cuda.current_stream(self.device).cuda_stream
        D = m.active_sh_degree
        self._bind(stream)
        va = self._view_args(cam)
        # the accumulated gradient buffers are cleared inside the blend kernel (gm_forward_ex): one launch less per iteration
        cap, _, _, geom, binning, image_state = self.arena.forward(
            self.P, D, self.M, bg, self.W, self.H, va, False, False, stream, out_color=self.image, radii=self.radii,
            epilogue=ForwardEpilogue(None, 0, None, None, self._accum.data_ptr(), self._accum.numel()))
        # An overflowed frame (tiles past the arena's capacity not rendered) must not reach the parameters: the
        # statistics and the optimizer are gated on the frame's device-side overflow word, the host sees the overflow
        # one or two iterations later (arena.poll in forward), counts the dropped iteration, and the arena has grown.
        gate = lib.gm_frame_overflow_flag(geom.data_ptr())
        if self.arena.overflowed:
            self.skipped_iterations += len(self.arena.overflowed)
            self.arena.overflowed.clear()
        check(lib.gm_photometric_loss(3, self.H, self.W, _p(self.image), _p(gt_image), float(opt.lambda_dssim),
                                      _p(self.scratch), _p(self.losses), _p(self.dL_dimg), stream), "gm_photometric_loss")
        g = self.grads
        check(lib.gm_backward_ex(self.P, D, self.M, cap, bg.data_ptr(), self.W, self.H, va[0], va[1], None, va[4], 1.0,
                                 va[6], None, va[8], va[9], va[10], va[11], va[12], self.radii.data_ptr(),
                                 geom.data_ptr(), binning.data_ptr(), image_state.data_ptr(), self.dL_dimg.data_ptr(),
                                 _p(g["means2D"]), _p(g["conic"]), _p(g["opacity"]), _p(g["colors"]), _p(g["means3D"]),
                                 _p(g["cov3D"]), _p(g["sh"]), _p(g["scales"]), _p(g["rotations"]), 0,
                                 GM_BACKWARD_OVERWRITE, stream), "gm_backward_ex")
        # train_mesh_gaussian.py:93: mrloss on the activated scale; its gradient joins the rasterizer's dL/dscale
        check(lib.gm_mesh_restrict_loss(self.P, _p(self.scale), _p(m.vertex1), _p(m.vertex2), _p(m.vertex3),
                                        float(opt.alpha_mrloss), self.losses.data_ptr() + 12, _p(g["scales"]), 1, stream),
              "gm_mesh_restrict_loss")
        check(lib.gm_mesh_bind_backward(self.P, _p(m._bc), _p(m._distance), _p(m.vertex1), _p(m.vertex2), _p(m.vertex3),
                                        _p(m.normal), _p(m.r), float(m.alpha_distance), _p(m._scaling), _p(m._rotation),
                                        _p(m._opacity), _p(g["means3D"]), _p(g["scales"]), _p(g["rotations"]),
                                        _p(g["opacity"]), _p(g["bc"]), _p(g["distance"]), _p(g["log_scale"]),
                                        _p(g["rot_raw"]), _p(g["opacity_logit"]), stream), "gm_mesh_bind_backward")
        if it < opt.densify_until_iter:                                       # train_mesh_gaussian.py:114-121
            check(lib.gm_densify_stats_gated(self.P, _p(self.radii), _p(g["means2D"]), _p(self.max_radii2D),
                                             _p(self.bc_gradient_accum), _p(self.denom), gate, stream), "gm_densify_stats")
        if densify and self.after_backward():                                  # train_mesh_gaussian.py:123-131,139
            return self.losses
        if optimizer_step and it < opt.iterations:                            # train_mesh_gaussian.py:136-147
            self.optimizer.step(self._grad_of, skip_flag=gate)
        return self.losses
