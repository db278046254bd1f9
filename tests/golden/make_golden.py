"""Generate the golden vectors of tests/golden/*.npz by running the UNMODIFIED reference CUDA rasterizer
(oracle/_ref, built from /root/reference by oracle/Makefile) on a B200:

    gpurun -- 'python tests/golden/make_golden.py'      # writes gpurun_out/golden/*.npz; copy them to tests/golden/

Inputs are regenerated from seeds by gaussianmesh_b200.synthetic (cases() below); each file stores the
reference's outputs (image, radii, instance count, per-Gaussian state, gradients for a seeded dL/dpixel).
The reference's own tests hold no vectors for this path, so these files are what pins the oracle.
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gaussianmesh_b200 import synthetic  # noqa: E402

CASES = {
    # name: (P, W, H, degree, variant, bg, scene kwargs, camera index of 7, want grads)
    "c1_10k_256": (10_000, 256, 256, 3, "sh", (0.0, 0.0, 0.0), {}, 2, False),
    "grad_2k_128x96": (2_000, 128, 96, 3, "sh", (0.1, 0.2, 0.3), {"log_scale_mean": math.log(0.03)}, 1, True),
    "colors_cov_3k_160x120": (3_000, 160, 120, 3, "colors+cov", (1.0, 1.0, 1.0), {"log_scale_mean": math.log(0.02)}, 3, True),
    "sh1_3k_250x130": (3_000, 250, 130, 1, "sh", (0.0, 0.5, 1.0), {"log_scale_mean": math.log(0.02)}, 5, True),
}


def case_inputs(name):
    P, W, H, degree, variant, bg, kw, cam_index, want_grads = CASES[name]
    seed = sum(map(ord, name)) % 1000
    arrays = synthetic.gaussian_scene(P, seed=seed, **kw)
    cam = synthetic.orbit_cameras(7, W, H)[cam_index]
    colors = cov = None
    if "colors" in variant:
        colors = np.random.default_rng(seed + 7).uniform(0.0, 1.0, size=(P, 3)).astype(np.float32)
    if "cov" in variant:
        cov = synthetic.packed_covariance(arrays["scales"], arrays["rotations"])
    dL = np.random.default_rng(seed + 1).uniform(-0.5, 0.5, size=(3, H, W)).astype(np.float32)
    return arrays, cam, np.asarray(bg, np.float32), degree, variant, colors, cov, dL, want_grads


def main():
    import torch
    import refcuda
    from gaussianmesh_b200.renderer import DeviceCamera
    dev = torch.device("cuda:0")
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in CASES:
        arrays, cam_h, bg, degree, variant, colors, cov, dL, want_grads = case_inputs(name)
        cam = DeviceCamera.upload(cam_h, dev)
        t = lambda a: None if a is None else torch.from_numpy(a).to(dev)
        kw = {}
        if "sh" in variant:
            kw["shs"] = t(arrays["shs"])
        else:
            kw["colors"] = t(colors)
        if "cov" in variant:
            kw["cov3D"] = t(cov)
        else:
            kw["scales"], kw["rotations"] = t(arrays["scales"]), t(arrays["rotations"])
        fr = refcuda.RefFrame(t(bg), t(arrays["means3D"]), t(arrays["opacities"]), cam.world_view_transform.contiguous(),
                              cam.full_proj_transform.contiguous(), cam.camera_center.contiguous(), cam_h.tanfovx,
                              cam_h.tanfovy, cam_h.image_height, cam_h.image_width, degree, **kw)
        rec = {"color": fr.color.cpu().numpy(), "radii": fr.radii.cpu().numpy(), "num_rendered": np.int64(fr.R)}
        gs = fr.geom_state()
        for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D", "clamped", "tiles_touched"):
            rec["state_" + k] = gs[k].cpu().numpy()
        ims = fr.image_state()
        rec["final_T"] = ims["final_T"].cpu().numpy()
        rec["n_contrib"] = ims["n_contrib"].cpu().numpy()
        if want_grads:
            g = fr.backward(t(dL))
            for k, v in g.items():
                rec["grad_" + k] = v.cpu().numpy()
        if colors is not None:
            rec["in_colors"] = colors
        if cov is not None:
            rec["in_cov3D"] = cov
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **rec)
        print(name, "R =", fr.R, "visible =", int((fr.radii > 0).sum()), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
