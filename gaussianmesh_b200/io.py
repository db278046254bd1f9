"""On-disk formats either side of the hot path, numpy only (no plyfile / Jittor):

  * the mesh-bound Gaussian PLY written by MeshBasedGaussianModel.save_ply and read by load_ply
    (scene/mesh_based_gaussian_model.py:290-409): binary little-endian, one `vertex` element, every
    property float32, in the order of construct_list_of_attributes (:290-303);
  * cameras.json as written by camera_to_JSON (utils/camera_utils.py:64-84) and read by
    ObjectVisualTool.get_camera (edittool/__init__.py:547-584).

Everything here is host-side bookkeeping; it produces the arrays `renderer.MeshGaussianModel` /
`renderer.DeformedObject` and `synthetic.Camera` are built from.
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Sequence

import numpy as np

from .synthetic import Camera, get_projection_matrix, get_world2view2

_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "int": "<i4", "int32": "<i4",
              "uint": "<u4", "uint32": "<u4", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
              "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1"}


def mesh_gaussian_attributes(num_rest: int = 45) -> List[str]:
    """construct_list_of_attributes (scene/mesh_based_gaussian_model.py:290-303) for 3 DC + num_rest SH values."""
    l = ['x', 'y', 'z', 'nx', 'ny', 'nz', 'ca', 'cb', 'cc', 'v1x', 'v1y', 'v1z', 'v2x', 'v2y', 'v2z', 'v3x', 'v3y', 'v3z',
         'dis', 'v_index1', 'v_index2', 'v_index3', 'radius', 'face_id']
    l += [f'f_dc_{i}' for i in range(3)]
    l += [f'f_rest_{i}' for i in range(num_rest)]
    l.append('opacity')
    l += [f'scale_{i}' for i in range(3)]
    l += [f'rot_{i}' for i in range(4)]
    return l


def read_ply_vertices(path: str) -> Dict[str, np.ndarray]:
    """All scalar properties of the `vertex` element of a binary-little-endian or ascii PLY."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    count = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties in the vertex element are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "binary_little_endian":
            data = np.frombuffer(f.read(count * np.dtype(props).itemsize), dtype=np.dtype(props), count=count)
        elif fmt == "ascii":
            raw = np.loadtxt(f, max_rows=count, ndmin=2)
            data = np.empty(count, dtype=np.dtype(props))
            for i, (name, _) in enumerate(props):
                data[name] = raw[:, i]
        else:
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
    return {name: np.asarray(data[name]) for name, _ in props}


def load_mesh_gaussian_ply(path: str, max_sh_degree: int = 3) -> Dict[str, np.ndarray]:
    """MeshBasedGaussianModel.load_ply (scene/mesh_based_gaussian_model.py:341-409) -> the arrays
    `renderer.MeshGaussianModel` takes.  Note the reference stores f_dc / f_rest channel-major
    ([P,3,K] flattened) and transposes to [P,K,3] on load; `shs` here is the concatenated [P,16,3].

    One deliberate difference: the reference's load_ply assigns `self._bc = jt.array(xyz)` (:393) -- the saved barycentric
    logits ca / cb / cc are read (:347-349) and dropped, so get_xyz of a reloaded model is meaningless there and the edit
    tool only ever uses get_load_xyz.  Here `bc_logits` are the saved ca / cb / cc (what save_ply wrote, :321) and the
    saved positions are returned separately as `xyz`, so a loaded model binds to its faces exactly as the trained one did.
    `vertex_index` (v_index1..3) are the per-Gaussian vertex ids MeshGaussianModel carries through densification."""
    v = read_ply_vertices(path)
    stack = lambda *names: np.stack([v[n] for n in names], axis=1).astype(np.float32)
    P = v["x"].shape[0]
    K = (max_sh_degree + 1) ** 2
    rest_names = sorted((n for n in v if n.startswith("f_rest_")), key=lambda n: int(n.split("_")[-1]))
    if len(rest_names) != 3 * K - 3:
        raise ValueError(f"{path}: expected {3 * K - 3} f_rest_* properties, found {len(rest_names)}")
    f_dc = stack("f_dc_0", "f_dc_1", "f_dc_2").reshape(P, 3, 1)
    f_rest = np.stack([v[n] for n in rest_names], axis=1).astype(np.float32).reshape(P, 3, K - 1)
    shs = np.concatenate([f_dc, f_rest], axis=2).transpose(0, 2, 1)          # [P,K,3]
    scale_names = sorted((n for n in v if n.startswith("scale_")), key=lambda n: int(n.split("_")[-1]))
    rot_names = sorted((n for n in v if n.startswith("rot")), key=lambda n: int(n.split("_")[-1]))
    return {
        "xyz": stack("x", "y", "z"),
        "bc_logits": stack("ca", "cb", "cc"),
        "vertex1": stack("v1x", "v1y", "v1z"), "vertex2": stack("v2x", "v2y", "v2z"), "vertex3": stack("v3x", "v3y", "v3z"),
        "normal": stack("nx", "ny", "nz"),
        "face_id": v["face_id"].astype(np.int32)[:, None],
        "vertex_index": np.stack([v["v_index1"], v["v_index2"], v["v_index3"]], axis=1).astype(np.int32),
        "distance": v["dis"].astype(np.float32)[:, None],
        "opacity_logit": v["opacity"].astype(np.float32)[:, None],
        "r": v["radius"].astype(np.float32)[:, None],
        "shs": np.ascontiguousarray(shs, dtype=np.float32),
        "log_scales": np.stack([v[n] for n in scale_names], axis=1).astype(np.float32),
        "rot_raw": np.stack([v[n] for n in rot_names], axis=1).astype(np.float32),
    }


def save_mesh_gaussian_ply(path: str, a: Dict[str, np.ndarray]) -> None:
    """MeshBasedGaussianModel.save_ply (scene/mesh_based_gaussian_model.py:305-334): same property names, order,
    float32 type and channel-major SH flattening.  `a` holds the keys load_mesh_gaussian_ply returns."""
    P = a["xyz"].shape[0]
    shs = np.asarray(a["shs"], dtype=np.float32)                               # [P,K,3]
    f_dc = shs[:, :1, :].transpose(0, 2, 1).reshape(P, -1)
    f_rest = shs[:, 1:, :].transpose(0, 2, 1).reshape(P, -1)
    cols = [a["xyz"], a["normal"], a["bc_logits"], a["vertex1"], a["vertex2"], a["vertex3"], a["distance"],
            a["vertex_index"], a["r"], a["face_id"], f_dc, f_rest, a["opacity_logit"], a["log_scales"], a["rot_raw"]]
    table = np.concatenate([np.asarray(c, dtype=np.float32).reshape(P, -1) for c in cols], axis=1)
    names = mesh_gaussian_attributes(f_rest.shape[1])
    assert table.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % P
    header += "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(np.ascontiguousarray(table, dtype="<f4").tobytes())


# ---------------------------------------------------------------------------------------------
def focal2fov(focal: float, pixels: float) -> float:
    return 2 * math.atan(pixels / (2 * focal))


def fov2focal(fov: float, pixels: float) -> float:
    return pixels / (2 * math.tan(fov / 2))


def camera_from_RT(R: np.ndarray, T: np.ndarray, fov_x: float, fov_y: float, width: int, height: int,
                   znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """scene/cameras.py:42-51 / edittool/camera_utils.py: R is camera-to-world, T = -R^T C."""
    wv = get_world2view2(R, T).transpose(1, 0)
    proj = get_projection_matrix(znear, zfar, fov_x, fov_y).transpose(1, 0)
    full = (wv.astype(np.float32) @ proj.astype(np.float32)).astype(np.float32)
    center = np.linalg.inv(wv)[3, :3].astype(np.float32)
    return Camera(int(width), int(height), float(fov_x), float(fov_y), np.ascontiguousarray(wv, dtype=np.float32), full, center)


def load_cameras_json(path: str) -> List[Camera]:
    """ObjectVisualTool.get_camera (edittool/__init__.py:547-584)."""
    cams = []
    with open(path) as f:
        for entry in json.load(f):
            W2C = np.zeros((4, 4))
            W2C[:3, :3] = np.array(entry["rotation"])
            W2C[:3, 3] = np.array(entry["position"])
            W2C[3, 3] = 1
            Rt = np.linalg.inv(W2C)
            T = Rt[:3, 3]
            R = Rt[:3, :3].transpose()
            w, h = entry["width"], entry["height"]
            cams.append(camera_from_RT(R, T, focal2fov(entry["fx"], w), focal2fov(entry["fy"], h), w, h))
    return cams


def cameras_to_json(cams: Sequence[Camera], names: Sequence[str] = ()) -> List[dict]:
    """camera_to_JSON (utils/camera_utils.py:64-84) for every camera: 'position' / 'rotation' are the
    camera-to-world translation and rotation."""
    out = []
    for i, c in enumerate(cams):
        C2W = np.linalg.inv(c.world_view_transform.T.astype(np.float64))
        out.append({"id": i, "img_name": names[i] if i < len(names) else f"{i:05d}", "width": c.image_width,
                    "height": c.image_height, "position": C2W[:3, 3].tolist(), "rotation": [r.tolist() for r in C2W[:3, :3]],
                    "fy": fov2focal(c.FoVy, c.image_height), "fx": fov2focal(c.FoVx, c.image_width)})
    return out
