#!/bin/bash
# compute-sanitizer passes over a small forward+backward (memcheck, racecheck, synccheck, initcheck).
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import scenes
from gaussianmesh_b200.diff_gaussian_rasterizater import GaussianRasterizer
from gaussianmesh_b200.renderer import make_settings, TrainStep
from gaussianmesh_b200.mesh_gaussians import l1_loss
dev = torch.device("cuda:0")
for (P, W, H, kw) in [(3000, 200, 120, {}), (6000, 48, 48, dict(extent=0.3, log_scale_mean=math.log(0.03)))]:
    sc = scenes.free_scene(P, dev, seed=1, **kw)
    for k in ("means3D", "opacities", "shs", "scales", "rotations"):
        sc[k].requires_grad_(True)
    cam = scenes.camera(dev, W, H, index=2)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    m2d = torch.zeros(P, 3, device=dev, requires_grad=True)
    color, radii = GaussianRasterizer(make_settings(cam, bg, 3))(sc["means3D"], m2d, sc["opacities"], shs=sc["shs"],
                                                                 scales=sc["scales"], rotations=sc["rotations"])
    target = torch.rand(3, H, W, device=dev)
    l1_loss(color, target).backward()
    ts = TrainStep(dev, sc["means3D"], sc["opacities"], sc["shs"], sc["scales"], sc["rotations"], W, H)
    ts.step(cam, bg, target); ts.step(cam, bg, target)
    torch.cuda.synchronize()
    print("case", P, W, H, "ok", float(color.mean()), int((radii > 0).sum()))

# surface shell: the sub-bucket path of big_bucket_sort_pack_kernel (with exact duplicates: its block-wide fallback)
P, W, H = 40000, 160, 120
sc = scenes.free_scene(P, dev, seed=3)
g = torch.Generator(device="cpu").manual_seed(5)
d = torch.randn(P, 3, generator=g); d = d / d.norm(dim=1, keepdim=True)
shell = (1.5 + 0.002 * torch.randn(P, 1, generator=g)) * d
shell[P // 2: P // 2 + P // 4] = shell[0]
cam = scenes.camera(dev, W, H, index=1)
bg = torch.zeros(3, device=dev)
color, radii = GaussianRasterizer(make_settings(cam, bg, 3))(shell.to(dev).contiguous(), torch.zeros(P, 3, device=dev), sc["opacities"] * 0.1,
                                                             shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
torch.cuda.synchronize()
print("case shell ok", float(color.mean()), int((radii > 0).sum()))

# the training iteration: photometric loss, mesh-restrict loss, bind, densification statistics, Adam; fused exchange at world 1
from gaussianmesh_b200 import synthetic
from gaussianmesh_b200.renderer import MeshGaussianModel
from gaussianmesh_b200.training import OptimizationParams, TrainingIteration
from gaussianmesh_b200.view_parallel import ViewParallelTrainer
P, W, H = 3001, 150, 70
V, F = synthetic.icosphere(2)
arrays = synthetic.mesh_bound_scene(P, V, F, seed=0)
cam = scenes.camera(dev, W, H, index=0)
target = torch.rand(3, H, W, device=dev)
it = TrainingIteration(MeshGaussianModel(arrays, dev, requires_grad=False), OptimizationParams(alpha_mrloss=0.05), W, H)
it.step(cam, bg, target); losses = it.step(cam, bg, target)
vp = ViewParallelTrainer(MeshGaussianModel(arrays, dev, requires_grad=False), OptimizationParams(alpha_mrloss=0.05), W, H, mode="p2p")
vp.step(cam, bg, target); vp.step(cam, bg, target)
torch.cuda.synchronize()
print("case iteration ok", losses.tolist())

# round-2 kernels: Python pipeline variants (covariance / SH colours with their backward), load_mesh, uint8 target loss
from gaussianmesh_b200.renderer import PipelineParams, render, GaussianModel
from gaussianmesh_b200.mesh_gaussians import load_mesh, l1_loss as l1
pc = MeshGaussianModel(arrays, dev)
bgm = GaussianModel(synthetic.gaussian_scene(1500, seed=2, extent=3.0), dev)
for flags in ((True, False), (False, True), (True, True)):
    out = render(cam, pc, PipelineParams(convert_SHs_python=flags[1], compute_cov3D_python=flags[0]), bg,
                 bg_gaussian=bgm if flags[0] else None)
    tu8 = torch.randint(0, 256, (3, H, W), dtype=torch.uint8, device=dev)
    l1(out["render"], tu8).backward()
tri, w = load_mesh(V.astype("float64"), F, arrays["face_id"], pc.activate()[0].detach())
torch.cuda.synchronize()
print("case python pipeline ok", float(w.sum()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case " gpurun_out/sanitize_$tool.log | tail -6
done
# the selectable kernel variants (tensor-core backward, barrier-free forward ring)
for tool in memcheck racecheck; do
  GM_BLEND_BWD=${SAN_BWD:-mma} GM_BLEND_FWD=${SAN_FWD:-ring} timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitize_variants_$tool.log 2>&1
  echo "== variants $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case " gpurun_out/sanitize_variants_$tool.log | tail -6
done
