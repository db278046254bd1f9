"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every symbol the header
declares (no compute calls without a GPU), the reference's C++ symbols are present, the product has no CPU
path, cameras follow the reference's conventions, and the N>1 view-shard plumbing works (gloo, world_size 2)."""
import ctypes
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gaussianmesh_b200", "diff_gaussian_rasterizater", "libCudaRasterizer.so")


@pytest.fixture(scope="session")
def built_lib():
    from gaussianmesh_b200 import build
    return str(build.build())


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gm_rasterizer.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(built_lib):
    names = _declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gm_rasterizer.h but not exported"
    from gaussianmesh_b200 import _lib
    missing = [n for n in names if n not in _lib.SIGNATURES]
    assert not missing, f"no ctypes signature for {missing}"
    assert _lib.version().endswith("sm_100a")


def test_reference_cxx_symbols_are_exported(built_lib):
    """The mangled names the reference's Jittor glue links against."""
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    for sym in ["_ZN14CudaRasterizer10Rasterizer11markVisibleEiPfS1_S1_Pb",
                "_ZN14CudaRasterizer10Rasterizer9forward_0EPciiiPKfiiS3_S3_S3_S3_S3_fS3_S3_S3_S3_S3_ffbPib",
                "_ZN14CudaRasterizer10Rasterizer9forward_1EPcS1_S1_iiiiPKfiiS3_S3_S3_S3_S3_fS3_S3_S3_S3_S3_ffbPfPib",
                "_ZN14CudaRasterizer10Rasterizer8backwardEiiiiPKfiiS2_S2_S2_S2_fS2_S2_S2_S2_S2_ffPKiPcS5_S5_S2_PfS6_S6_S6_S6_S6_S6_S6_S6_b",
                "_ZN14CudaRasterizer13GeometryState9fromChunkERPcm", "_ZN14CudaRasterizer10ImageState9fromChunkERPcm",
                "_ZN14CudaRasterizer12BinningState9fromChunkERPcm"]:
        assert sym in out, sym
    assert "Rasterizer7forwardESt8function" in out


def test_library_is_sm_100a_with_bulk_copies(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass, "blend kernels must stage splat records with cp.async.bulk (SASS UBLKCP)"


def test_chunk_sizing_is_monotone_and_invertible(built_lib):
    from gaussianmesh_b200._lib import lib
    prev = 0
    for R in [0, 1, 7, 8, 1000, 123457, 6_500_000]:
        need = lib.gm_required_binning(R)
        assert need >= prev
        prev = need
        cap = lib.gm_binning_capacity(need)
        assert cap >= R and lib.gm_required_binning(cap) <= need
    assert lib.gm_required_geom(1_000_000) < 200 * 1_000_000
    assert lib.gm_required_image(1920 * 1080) <= 8 * 1920 * 1080 + 1024


def test_no_cpu_path(built_lib):
    import torch
    from gaussianmesh_b200 import RasterizerError
    from gaussianmesh_b200.diff_gaussian_rasterizater import GaussianRasterizer
    from gaussianmesh_b200.renderer import make_settings
    from gaussianmesh_b200 import synthetic
    cam = synthetic.orbit_cameras(2, 32, 32)[0]

    class C:  # host tensors
        image_width, image_height, FoVx, FoVy = 32, 32, cam.FoVx, cam.FoVy
        world_view_transform = torch.from_numpy(cam.world_view_transform)
        full_proj_transform = torch.from_numpy(cam.full_proj_transform)
        camera_center = torch.from_numpy(cam.camera_center)

    sc = {k: torch.from_numpy(v) for k, v in synthetic.gaussian_scene(8).items()}
    r = GaussianRasterizer(make_settings(C, torch.zeros(3), 3))
    with pytest.raises(RasterizerError, match="no CPU path"):
        r(sc["means3D"], torch.zeros(8, 3), sc["opacities"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    from gaussianmesh_b200.arena import RenderArena
    with pytest.raises(RasterizerError):
        RenderArena("cpu")


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gaussianmesh_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").lower() or f == "build.py" or "import oracle" not in text, f
                assert "import oracle" not in text and "from oracle" not in text and "refcuda" not in text, f


def test_camera_conventions():
    from gaussianmesh_b200 import synthetic
    cam = synthetic.orbit_cameras(8, 320, 180)[3]
    W2C = cam.world_view_transform.T                     # stored transposed (scene/cameras.py:48)
    assert np.allclose(W2C[3], [0, 0, 0, 1])
    assert np.allclose(W2C[:3, :3] @ W2C[:3, :3].T, np.eye(3), atol=1e-6)
    c = cam.camera_center
    assert np.allclose(W2C[:3, :3] @ c + W2C[:3, 3], 0, atol=1e-5)      # campos maps to the view origin
    o = W2C @ np.array([0, 0, 0, 1.0])
    assert o[2] > 0 and abs(o[0]) < 1e-5 and abs(o[1]) < 1e-5          # looks at the origin, +z forward
    p = np.array([0.3, -0.2, 0.1, 1.0]) @ cam.full_proj_transform       # row-vector convention
    v = np.array([0.3, -0.2, 0.1, 1.0]) @ cam.world_view_transform
    assert np.isclose(p[3], v[2], atol=1e-5)                            # w = view-space depth
    assert np.isclose(cam.tanfovy / cam.tanfovx, 180 / 320, atol=1e-6)
    assert cam.packed().shape == (35,)


def test_shard_views_partition():
    from gaussianmesh_b200.view_shard import shard_views
    assert [len(shard_views(100, 8, r)) for r in range(8)] == [13, 13, 13, 13, 12, 12, 12, 12]
    for n, w in [(100, 8), (200, 8), (7, 4), (3, 8), (0, 2), (1, 1)]:
        got = [i for r in range(w) for i in shard_views(n, w, r)]
        assert got == list(range(n))
    with pytest.raises(ValueError):
        shard_views(10, 2, 2)


_WORKER = r"""
import os, sys, json, hashlib
sys.path.insert(0, {root!r})
import numpy as np
from gaussianmesh_b200.view_shard import ShardContext
from gaussianmesh_b200 import synthetic
from oracle import cpu_oracle
ctx = ShardContext("gloo")
arrays = synthetic.gaussian_scene(300, seed=0, log_scale_mean=-3.0)
cams = synthetic.orbit_cameras(5, 48, 32)
mine = list(ctx.views(len(cams)))
ctx.barrier()
rec = {{}}
for i in mine:
    img = cpu_oracle.forward(arrays, cams[i], np.zeros(3, np.float32), 3)["color"]
    rec[i] = hashlib.sha1(img.tobytes()).hexdigest()
slow = ctx.max_over_ranks(10.0 + ctx.rank)
allrec = ctx.gather_objects(rec)
if ctx.rank == 0:
    merged = {{}}
    for r in allrec:
        merged.update(r)
    print(json.dumps({{"world": ctx.world, "max": slow, "views": sorted(int(k) for k in merged), "hash": merged}}))
ctx.close()
"""


def test_two_rank_view_shard_over_gloo(tmp_path):
    """world_size 2 on CPU: each rank renders its own contiguous block of views (here with the CPU oracle as the
    stand-in renderer), no data-path collective; rank 0 sees the union and the max-over-ranks reduction."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "cpu"], check=True, capture_output=True)
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["world"] == 2 and res["max"] == 11.0 and res["views"] == [0, 1, 2, 3, 4]
    # the same views rendered by one process give the same images (view sharding changes nothing)
    import hashlib
    from gaussianmesh_b200 import synthetic
    from oracle import cpu_oracle
    arrays = synthetic.gaussian_scene(300, seed=0, log_scale_mean=-3.0)
    cams = synthetic.orbit_cameras(5, 48, 32)
    for i, cam in enumerate(cams):
        img = cpu_oracle.forward(arrays, cam, np.zeros(3, np.float32), 3)["color"]
        assert hashlib.sha1(img.tobytes()).hexdigest() == res["hash"][str(i)]


def test_mesh_gaussian_ply_roundtrip_and_schema(tmp_path):
    """The PLY schema of MeshBasedGaussianModel.save_ply / load_ply (scene/mesh_based_gaussian_model.py:290-409)."""
    from gaussianmesh_b200 import io, synthetic
    V, F = synthetic.icosphere(1)
    a = synthetic.mesh_bound_scene(300, V, F, seed=3)
    rec = {"xyz": np.random.default_rng(0).normal(size=(300, 3)).astype(np.float32), "normal": a["normal"],
           "bc_logits": a["bc_logits"], "vertex1": a["vertex1"], "vertex2": a["vertex2"], "vertex3": a["vertex3"],
           "distance": a["distance"], "vertex_index": a["triangles"], "r": a["r"], "face_id": a["face_id"][:, None],
           "shs": a["shs"], "opacity_logit": a["opacity_logit"], "log_scales": a["log_scales"], "rot_raw": a["rot_raw"]}
    path = str(tmp_path / "point_cloud.ply")
    io.save_mesh_gaussian_ply(path, rec)
    header = open(path, "rb").read(4096).split(b"end_header")[0].decode()
    props = [l.split()[-1] for l in header.splitlines() if l.startswith("property")]
    assert props == io.mesh_gaussian_attributes(45) and len(props) == 24 + 3 + 45 + 1 + 3 + 4
    assert all(l.split()[1] == "float" for l in header.splitlines() if l.startswith("property"))
    assert props[:9] == ['x', 'y', 'z', 'nx', 'ny', 'nz', 'ca', 'cb', 'cc'] and props[18:24] == ['dis', 'v_index1', 'v_index2', 'v_index3', 'radius', 'face_id']
    back = io.load_mesh_gaussian_ply(path)
    for k in ("xyz", "bc_logits", "vertex1", "vertex2", "vertex3", "normal", "distance", "opacity_logit", "r", "shs",
              "log_scales", "rot_raw"):
        assert np.array_equal(back[k], rec[k].astype(np.float32)), k
    assert np.array_equal(back["vertex_index"], a["triangles"]) and np.array_equal(back["face_id"][:, 0], a["face_id"])
    # f_dc / f_rest are channel-major on disk: f_rest_0..14 are the red channel of coefficients 1..15
    v = io.read_ply_vertices(path)
    assert np.array_equal(v["f_rest_0"], a["shs"][:, 1, 0]) and np.array_equal(v["f_rest_15"], a["shs"][:, 1, 1])
    assert np.array_equal(v["f_dc_2"], a["shs"][:, 0, 2])


def test_loaded_ply_builds_a_model_that_saves_back(tmp_path, built_lib):
    """load_mesh_gaussian_ply -> MeshGaussianModel -> to_arrays -> save_mesh_gaussian_ply keeps every column, including the
    per-Gaussian vertex ids and face ids the densification bookkeeping needs (the loader names them vertex_index / face_id,
    the synthetic generator triangles / face_id: the model takes either)."""
    import torch
    from gaussianmesh_b200 import io, synthetic
    from gaussianmesh_b200.renderer import MeshGaussianModel
    V, F = synthetic.icosphere(1)
    a = synthetic.mesh_bound_scene(120, V, F, seed=5)
    rec = {"xyz": np.zeros((120, 3), np.float32), "normal": a["normal"], "bc_logits": a["bc_logits"], "vertex1": a["vertex1"],
           "vertex2": a["vertex2"], "vertex3": a["vertex3"], "distance": a["distance"], "vertex_index": a["triangles"], "r": a["r"],
           "face_id": a["face_id"][:, None], "shs": a["shs"], "opacity_logit": a["opacity_logit"], "log_scales": a["log_scales"],
           "rot_raw": a["rot_raw"]}
    path = str(tmp_path / "in.ply")
    io.save_mesh_gaussian_ply(path, rec)
    model = MeshGaussianModel(io.load_mesh_gaussian_ply(path), torch.device("cpu"), requires_grad=False)
    assert model.vertex_index is not None and torch.equal(model.vertex_index, torch.from_numpy(a["triangles"]).long())
    assert model.fid is not None and torch.equal(model.fid.view(-1), torch.from_numpy(a["face_id"]).long())
    out = str(tmp_path / "out.ply")
    io.save_mesh_gaussian_ply(out, model.to_arrays())
    back = io.load_mesh_gaussian_ply(out)
    for k in ("bc_logits", "vertex1", "vertex2", "vertex3", "normal", "distance", "opacity_logit", "r", "shs", "log_scales",
              "rot_raw", "vertex_index"):
        assert np.array_equal(back[k], rec[k].astype(back[k].dtype)), k
    assert np.array_equal(back["face_id"][:, 0], a["face_id"])


def test_acap_log_rotations_follow_the_reference_branches(built_lib):
    """GetRS(..., _R = 0): the breadth-first branch selection of the log-rotations (RefMesh::bfscorrot + logrot) against
    the reference's own golden vector test/LOGRNEW.txt, whose angles run up to 5.97 rad -- far beyond the principal range."""
    import ctypes as C
    from gaussianmesh_b200._lib import lib, check
    from gaussianmesh_b200.acap import log_rotations_bfs
    d = np.load(os.path.join(ROOT, "tests", "golden", "acap_1_to_2.npz"))
    V, F = d["V_rest"], np.ascontiguousarray(d["F"], np.int32)
    Vn, Fn = V.shape[0], F.shape[0]
    ring_off, ring = np.zeros(Vn + 1, np.int32), np.zeros(3 * Fn + Vn, np.int32)
    face_off, face_list = np.zeros(Vn + 1, np.int32), np.zeros(3 * Fn, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib.gm_acap_build_rings(Vn, Fn, p(F), p(ring_off), p(ring), p(face_off), p(face_list)), "gm_acap_build_rings")
    gold = d["logR_gold"]
    angle = np.sqrt(gold[:, 0, 1] ** 2 + gold[:, 0, 2] ** 2 + gold[:, 1, 2] ** 2)
    assert angle.max() > np.pi                                    # the golden really leaves the principal branch
    got = log_rotations_bfs(d["R_gold"], ring_off, ring)
    err = np.abs(got - gold).reshape(Vn, -1).max(axis=1)
    # two vertices sit on a branch point (exp of the stored logarithm is within rounding of a rotation by pi)
    assert (err <= 1e-6).mean() >= 0.995 and err.max() <= 0.05, (float((err <= 1e-6).mean()), float(err.max()))


def test_cameras_json_roundtrip(tmp_path):
    """camera_to_JSON (utils/camera_utils.py:64-84) <-> ObjectVisualTool.get_camera (edittool/__init__.py:547-584)."""
    import json
    from gaussianmesh_b200 import io, synthetic
    cams = synthetic.orbit_cameras(5, 640, 360)
    entries = io.cameras_to_json(cams)
    assert set(entries[0]) == {"id", "img_name", "width", "height", "position", "rotation", "fy", "fx"}
    assert np.allclose(entries[0]["position"], cams[0].camera_center, atol=1e-5)
    path = tmp_path / "cameras.json"
    path.write_text(json.dumps(entries))
    back = io.load_cameras_json(str(path))
    for a, b in zip(cams, back):
        assert (a.image_width, a.image_height) == (b.image_width, b.image_height)
        assert abs(a.FoVx - b.FoVx) < 1e-6 and abs(a.FoVy - b.FoVy) < 1e-6
        assert np.allclose(a.world_view_transform, b.world_view_transform, atol=1e-5)
        assert np.allclose(a.full_proj_transform, b.full_proj_transform, atol=1e-4)
        assert np.allclose(a.camera_center, b.camera_center, atol=1e-5)


_EXCHANGE_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch
import torch.distributed as dist
dist.init_process_group("gloo")
from gaussianmesh_b200.view_parallel import GradientExchange
ex = GradientExchange()
r = ex.rank
flat = torch.arange(10, dtype=torch.float32) * (r + 1)                 # rank 0: i, rank 1: 2 i  -> mean 1.5 i
ex.average_(flat)
max_r = torch.tensor([1.0, 5.0, 0.0, 9.0]); acc = torch.zeros(4, 1); den = torch.ones(4, 1)
inc_max = torch.tensor([0.0, 7.0, 3.0, 0.0]) if r == 0 else torch.tensor([2.0, 0.0, 4.0, 0.0])
inc_sum = torch.tensor([[0.0, 0.5, 0.25, 0.0], [0.0, 1.0, 1.0, 0.0]]) if r == 0 else torch.tensor([[1.0, 0.0, 0.75, 0.0], [1.0, 0.0, 1.0, 0.0]])
ex.merge_stats_(max_r, acc, den, inc_max, inc_sum)
out = [None] * ex.world
dist.all_gather_object(out, {{"flat": flat.tolist(), "max": max_r.tolist(), "acc": acc.view(-1).tolist(), "den": den.view(-1).tolist()}})
if r == 0:
    print(json.dumps(out))
dist.destroy_process_group()
"""


def test_gradient_exchange_two_ranks_over_gloo(tmp_path):
    """The device-agnostic half of view-parallel training (gradient average, per-view densification statistics merged
    as consecutive single-view iterations would leave them) at world_size 2 on CPU."""
    script = tmp_path / "exchange_worker.py"
    script.write_text(_EXCHANGE_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29733", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("[")][-1])
    assert len(res) == 2 and res[0] == res[1]
    assert res[0]["flat"] == [1.5 * i for i in range(10)]
    assert res[0]["max"] == [2.0, 7.0, 4.0, 9.0] and res[0]["acc"] == [1.0, 0.5, 1.0, 0.0] and res[0]["den"] == [2.0, 2.0, 3.0, 1.0]


def test_adam_shard_ranges_partition_the_flat_vector(built_lib):
    import ctypes as C
    from gaussianmesh_b200._lib import lib
    for total, world in [(1000, 3), (64, 2), (32, 4), (60_000_096, 8), (0, 2), (96, 1)]:
        at = 0
        for r in range(world):
            lo, hi = C.c_size_t(), C.c_size_t()
            lib.gm_adam_shard_range(total, world, r, C.byref(lo), C.byref(hi))
            assert lo.value == at and hi.value >= lo.value and (lo.value % 32 == 0 or lo.value == total)
            at = hi.value
        assert at == total


def test_row_interval_tile_walk_keeps_every_contributing_tile():
    """The closed-form tile walk of preprocess (csrc/preprocess.cu: per tile row, the x-interval of the threshold
    ellipse cut by the row's band) restated in numpy float32 and held against brute force: every tile that contains a
    pixel centre with q(d) <= 2 log(255 o) -- the blend kernels' alpha >= 1/255 test (forward.cu:336-345) -- must be
    kept; and the walk must not keep more than a thin rim beyond the exact set of tiles the ellipse touches."""
    f = np.float32
    rng = np.random.default_rng(21)
    T, W, H = 16, 640, 480
    kept_total = exact_total = 0
    for _ in range(400):
        # a random conic from a random 2-D covariance (as preprocess builds it) and a random centre / opacity
        s1, s2 = np.exp(rng.normal(1.2, 0.8)), np.exp(rng.normal(1.2, 0.8))
        th = rng.uniform(0, np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        cov = R @ np.diag([s1 ** 2, s2 ** 2]) @ R.T + 0.3 * np.eye(2)
        det = cov[0, 0] * cov[1, 1] - cov[0, 1] ** 2
        a, b, c = f(cov[1, 1] / det), f(-cov[0, 1] / det), f(cov[0, 0] / det)
        mx, my = f(rng.uniform(-20, W + 20)), f(rng.uniform(-20, H + 20))
        o = f(rng.uniform(0.01, 1.0))
        thr = f(max(0.0, np.log(f(255.0) * o)) + 1e-4)
        # brute force over pixel centres
        ys, xs = np.mgrid[0:H, 0:W]
        dx, dy = (xs - mx).astype(f), (ys - my).astype(f)
        q = a * dx * dx + f(2) * b * dx * dy + c * dy * dy
        hit = q <= f(2) * thr
        exact = {(int(y) // T, int(x) // T) for y, x in zip(*np.nonzero(hit))}
        # the walk (float32, same expressions as the kernel)
        lvl = f(2) * thr + f(0.05)
        ex, ey = f(1.01) * np.sqrt(lvl * f(cov[0, 0])) + f(1), f(1.01) * np.sqrt(lvl * f(cov[1, 1])) + f(1)
        tiles_x, tiles_y = (W + T - 1) // T, (H + T - 1) // T
        tx_lo, tx_hi = max(0, int(np.floor((mx - ex) / T))), min(tiles_x, int(np.floor((mx + ex) / T)) + 1)
        ty_lo, ty_hi = max(0, int(np.floor((my - ey) / T))), min(tiles_y, int(np.floor((my + ey) / T)) + 1)
        kept = set()
        if tx_lo < tx_hi and ty_lo < ty_hi:
            det_c = a * c - b * b
            bx = max(abs(f(tx_lo * T) - mx), abs(f(tx_hi * T) - mx))
            by = max(abs(f(ty_lo * T) - my), abs(f(ty_hi * T) - my))
            Lq = f(2) * thr + f(2e-3) + f(4e-6) * (a * bx * bx + f(2) * abs(b) * bx * by + c * by * by)
            inv_a = f(1) / a
            dx_ext = np.sqrt(Lq * c / det_c)
            dy_at = b / c * dx_ext
            aL = a * Lq
            for ty in range(ty_lo, ty_hi):
                y0b, y1b = f(ty * T) - my, f(min(ty * T + T - 1, H - 1)) - my
                dyM, dym = min(y1b, max(y0b, -dy_at)), min(y1b, max(y0b, dy_at))
                DM, Dm = aL - det_c * dyM * dyM, aL - det_c * dym * dym
                if DM < 0 and Dm < 0:
                    continue
                xmax = (-b * dyM + np.sqrt(max(DM, f(0)))) * inv_a
                xmin = (-b * dym - np.sqrt(max(Dm, f(0)))) * inv_a
                ta = max(tx_lo, int(np.ceil((mx + xmin - f(T - 1)) / T - 1e-3)))
                tb = min(tx_hi - 1, int(np.floor((mx + xmax) / T + 1e-3)))
                kept |= {(ty, tx) for tx in range(ta, tb + 1)}
        assert exact <= kept, (exact - kept, (a, b, c, mx, my, o))
        kept_total += len(kept)
        exact_total += len(exact)
    assert exact_total > 1000
    # tiles the continuous ellipse touches without covering a pixel centre: a thin rim only
    assert kept_total <= 1.35 * exact_total


def test_column_sum_flush_matches_the_per_pixel_sums():
    """The backward blend's flush (csrc/blend_bwd.cu, blend_backward_cols_kernel) restated in numpy float32 with the
    kernel's expression order: row partial sums over the eight pixels of a row with u = px - block centre, the six
    pixel-centred moments, the join of the two 16-pixel halves, the shift to the splat (dx = X - u) -- held against the
    sums the reference adds pixel by pixel (backward.cu:537-554: sum q dx, q dy, q dx^2, q dx dy, q dy^2 with
    dx = x_splat - pixel) in float64.  Splats inside, beside and far from the 8x4 block."""
    f = np.float32
    rng = np.random.default_rng(5)
    worst = 0.0
    for case in range(300):
        bx0, by0 = f(8 * rng.integers(0, 200)), f(4 * rng.integers(0, 250))
        cx, cy = bx0 + f(3.5), by0 + f(1.5)
        far = (0.0, 6.0, 40.0)[case % 3]
        xs = f(cx + rng.normal(0, 3.0 + far))
        ys = f(cy + rng.normal(0, 2.0 + far))
        q = (rng.normal(0, 1, 32) * np.exp(rng.normal(-3, 2, 32))).astype(f)     # G dL/dalpha, any sign, wide range
        q[rng.random(32) < 0.3] = 0                                               # pixels that skip the splat
        Sq = Su = Sv = Suu = Suv = Svv = None
        halves = []
        for half in range(2):
            acc = [f(0)] * 6
            for r in range(2):
                R0 = R1 = R2 = f(0)
                for u in range(8):
                    p = half * 16 + r * 8 + u
                    uu = f(u) - f(3.5)
                    R0 = f(R0 + q[p])
                    R1 = f(np.float32(q[p]) * uu + R1)
                    R2 = f(np.float32(q[p]) * f(uu * uu) + R2)
                vv = f(2 * half + r) - f(1.5)
                acc = [f(acc[0] + R0), f(acc[1] + R1), f(vv * R0 + acc[2]), f(acc[3] + R2), f(vv * R1 + acc[4]),
                       f(f(vv * vv) * R0 + acc[5])]
            halves.append(acc)
        Sq, Su, Sv, Suu, Suv, Svv = (f(a + b) for a, b in zip(*halves))
        X, Y = f(xs - cx), f(ys - cy)
        Sx, Sy = f(X * Sq - Su), f(Y * Sq - Sv)
        Sxx = f(X * f(X * Sq - f(2) * Su) + Suu)
        Sxy = f(X * f(Y * Sq - Sv) + f(-Y * Su + Suv))
        Syy = f(Y * f(Y * Sq - f(2) * Sv) + Svv)
        # the reference's sums, float64
        px = bx0 + (np.arange(32) % 8).astype(np.float64)
        py = by0 + (np.arange(32) // 8).astype(np.float64)
        dx, dy = float(xs) - px, float(ys) - py
        q64 = q.astype(np.float64)
        want = [q64.sum(), (q64 * dx).sum(), (q64 * dy).sum(), (q64 * dx * dx).sum(), (q64 * dx * dy).sum(), (q64 * dy * dy).sum()]
        # the scale an fp32 accumulation of the same terms is exact to
        mag = [np.abs(q64).sum(), (np.abs(q64) * np.abs(dx)).sum(), (np.abs(q64) * np.abs(dy)).sum(), (np.abs(q64) * dx * dx).sum(),
               (np.abs(q64) * np.abs(dx * dy)).sum(), (np.abs(q64) * dy * dy).sum()]
        for got, w, m in zip((Sq, Sx, Sy, Sxx, Sxy, Syy), want, mag):
            err = abs(float(got) - w) / max(m, 1e-30)
            worst = max(worst, err)
            assert err <= 5e-6, (case, float(got), w, m)      # observed worst: 5.4e-7
    assert worst > 0.0        # (the restatement is float32, not a copy of the float64 sums)


def test_cull_mask_entries_of_consecutive_tiles_never_collide():
    """csrc/state.h: the survivor masks of chunk c of tile t live at entry (start_t >> 5) + t + c of the key array.  For
    any tile layout the binning produces (starts multiples of 4, start_{t+1} >= start_t + n_t) the entries of different
    tiles are disjoint and stay below the bound cull_mask_fits() checks."""
    rng = np.random.default_rng(3)
    for trial in range(200):
        T = int(rng.integers(1, 400))
        counts = rng.integers(0, 200, size=T) * (rng.random(T) < 0.7)
        if trial % 5 == 0:
            counts[rng.integers(0, T)] = int(rng.integers(3000, 9000))
        starts, at = [], 0
        for n in counts:
            starts.append(at)
            at += (int(n) + 3) & ~3                                   # tile segments start at multiples of 4 instances
        num_rendered = at
        seen = {}
        for t, (s0, n) in enumerate(zip(starts, counts)):
            for c in range((int(n) + 31) // 32):
                e = (s0 >> 5) + t + c
                assert e not in seen, (trial, t, seen[e])
                seen[e] = t
        if seen:
            assert max(seen) + 4 < (num_rendered >> 5) + T + 8        # the backward's 4-chunk bulk copy stays inside too


def test_header_is_plain_c(tmp_path):
    """include/gm_rasterizer.h is the C-ABI contract: it must compile as C99 (no C++ in the signatures) and the two
    three structs passed by pointer must have the layout the ctypes mirror assumes."""
    import ctypes as C
    src = tmp_path / "hdr.c"
    src.write_text('#include <stdio.h>\n#include "gm_rasterizer.h"\n'
                   'int main(void) { printf("%zu %zu %zu\\n", sizeof(gm_adam_tensor), sizeof(gm_adam_segment), '
                   'sizeof(gm_forward_epilogue)); return 0; }\n')
    exe = tmp_path / "hdr"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True, capture_output=True)
    sizes = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    from gaussianmesh_b200._lib import AdamTensor, AdamSegment, ForwardEpilogue
    assert [int(x) for x in sizes] == [C.sizeof(AdamTensor), C.sizeof(AdamSegment), C.sizeof(ForwardEpilogue)]
