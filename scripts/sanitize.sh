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
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case " gpurun_out/sanitize_$tool.log | tail -5
done
