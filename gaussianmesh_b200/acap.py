"""ACAP rotation / shear extraction on the GPU -- the drop-in for `pyACAP.pyACAP`.

    tool = pyACAP(mesh_path_or_(V, F))                    # reference: pyACAP.pyACAP(mesh_path), edittool/__init__.py:102
    R1, S1 = tool.GetRS(rest_V, deformed_V, 1, ncpu)      # reference: edittool/__init__.py:109
    cur_rot, cur_shear = R1.reshape(-1, 3, 3), S1.reshape(-1, 3, 3)

R1 / S1 are float32 CUDA tensors of shape [1, 9*Vn] (pyACAP returns 1 x 9Vn matrices); R1 holds the TRANSPOSE of
each polar rotation, exactly as the reference library returns it (FeatureVector.cpp:560-570), so
`SingleObjectDeform.deform_gaussian` / `DeformedObject.deform` can consume it unchanged.  Only `_R = 1` (rotation
matrices) is implemented -- the only mode the reference calls; `cpunum` is accepted and ignored.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple, Union

import numpy as np
import torch

from ._lib import lib, check, RasterizerError, GM_ERR_BAD_ARGUMENT


def read_obj(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """Vertices and triangle faces of a Wavefront OBJ (what igl.read_triangle_mesh / OpenMesh give the reference)."""
    V, F = [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                V.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                idx = [int(x.split("/")[0]) for x in t[1:]]
                idx = [i - 1 if i > 0 else len(V) + i for i in idx]
                for k in range(1, len(idx) - 1):           # fan-triangulate polygons
                    F.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(V, np.float64), np.asarray(F, np.int32)


class pyACAP:
    def __init__(self, mesh: Union[str, Tuple[np.ndarray, np.ndarray]], device="cuda"):
        V, F = read_obj(mesh) if isinstance(mesh, str) else mesh
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RasterizerError("pyACAP", GM_ERR_BAD_ARGUMENT, "a CUDA device is required (there is no CPU path)")
        V = np.ascontiguousarray(V, np.float64)
        F = np.ascontiguousarray(F, np.int32)
        self.Vn, self.Fn = V.shape[0], F.shape[0]
        ring_off = np.zeros(self.Vn + 1, np.int32)
        ring = np.zeros(3 * self.Fn + self.Vn, np.int32)
        face_off = np.zeros(self.Vn + 1, np.int32)
        face_list = np.zeros(max(3 * self.Fn, 1), np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib.gm_acap_build_rings(self.Vn, self.Fn, p(F), p(ring_off), p(ring), p(face_off), p(face_list)),
              "gm_acap_build_rings")
        E = int(ring_off[-1])
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.V0, self.F = d(V), d(F)
        self.ring_off, self.ring = d(ring_off), d(ring[:max(E, 1)])
        self.face_off, self.face_list = d(face_off), d(face_list)
        self.sqrt_w = torch.empty(max(E, 1), dtype=torch.float64, device=self.device)
        self.n0 = torch.empty(self.Vn, 3, dtype=torch.float64, device=self.device)
        self.ata_inv = torch.empty(self.Vn, 9, dtype=torch.float64, device=self.device)
        self._n1 = torch.empty(self.Vn, 3, dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.gm_acap_rest(self.Vn, self.V0.data_ptr(), self.F.data_ptr(), self.ring_off.data_ptr(), self.ring.data_ptr(),
                               self.face_off.data_ptr(), self.face_list.data_ptr(), self.sqrt_w.data_ptr(), self.n0.data_ptr(),
                               self.ata_inv.data_ptr(), stream), "gm_acap_rest")

    def GetRS(self, ref_array, deformed_array, _R=1, cpunum=0) -> Tuple[torch.Tensor, torch.Tensor]:
        """(R [1, 9 Vn], S [1, 9 Vn]) float32 on the device.  `ref_array` must be the mesh the object was built from
        (the reference ignores it too: mainpy.cpp:60-64 only updates the deformed copy)."""
        V1 = torch.as_tensor(deformed_array, dtype=torch.float64).contiguous().to(self.device)
        if V1.shape != (self.Vn, 3):
            raise RasterizerError("pyACAP.GetRS", GM_ERR_BAD_ARGUMENT, f"deformed vertices must be [{self.Vn}, 3]")
        R = torch.empty(1, 9 * self.Vn, dtype=torch.float32, device=self.device)
        S = torch.empty(1, 9 * self.Vn, dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.gm_acap_get_rs(self.Vn, self.V0.data_ptr(), V1.data_ptr(), self.F.data_ptr(), self.ring_off.data_ptr(),
                                 self.ring.data_ptr(), self.face_off.data_ptr(), self.face_list.data_ptr(),
                                 self.sqrt_w.data_ptr(), self.n0.data_ptr(), self.ata_inv.data_ptr(), self._n1.data_ptr(),
                                 R.data_ptr(), S.data_ptr(), stream), "gm_acap_get_rs")
        if not _R:
            # log-rotations with the branch (axis sign, multiples of 2 pi) propagated breadth-first over the mesh
            # (RefMesh::bfscorrot, FeatureVector.cpp:269-318; logrot, Align.cpp:131-276).  A sequential walk over the
            # vertices of a control mesh: host side, like the ring builder; no caller of the reference uses it.
            logr = log_rotations_bfs(R.view(self.Vn, 3, 3).cpu().numpy(), self.ring_off.cpu().numpy(), self.ring.cpu().numpy())
            R = torch.from_numpy(logr.reshape(1, -1).astype(np.float32)).to(self.device)
        return R, S


def _skew(axis: np.ndarray, angle: float) -> np.ndarray:
    """Rot::ToLogR (Align.cpp:112-122)"""
    m = np.zeros((3, 3))
    m[0, 1], m[0, 2], m[1, 2] = -axis[2], axis[1], -axis[0]
    return angle * (m - m.T)


def log_rotations_bfs(R_out: np.ndarray, ring_off: np.ndarray, ring: np.ndarray, root: int = 0) -> np.ndarray:
    """GetRS(..., _R = 0): per-vertex log-rotation matrices [Vn,3,3] as the reference returns them (the transposed
    log of the polar rotation, FeatureVector.cpp:531-545).  `R_out` [Vn,3,3] is what GetRS(_R = 1) returns (r^T).
    The principal logarithm is ambiguous by the sign of the axis and by multiples of 2 pi; the reference fixes both
    by walking the mesh breadth-first from `root` and keeping every vertex close to its BFS parent."""
    r = np.swapaxes(np.asarray(R_out, np.float64), 1, 2)
    Vn = r.shape[0]
    axis = np.zeros((Vn, 3)); theta = np.zeros(Vn); circlek = np.zeros(Vn)
    out = np.zeros((Vn, 3, 3))
    visited = np.zeros(Vn, bool)
    acos = np.arccos(np.clip((np.trace(r, axis1=1, axis2=2) - 1.0) / 2.0, -1.0, 1.0))

    def raw_axis(j, th):
        t = (r[j] - r[j].T) / (2.0 * np.sin(th))
        return np.array([t[2, 1], t[0, 2], t[1, 0]])

    for start in range(root, Vn):
        if visited[start]:
            continue
        visited[start] = True
        queue = [(start, -1)]
        head = 0
        while head < len(queue):
            j, fa = queue[head]
            head += 1
            th = float(acos[j])
            if fa < 0:                                    # logrot(r), Align.cpp:131-154
                if abs(th) <= 1e-6:
                    axis[j], theta[j], circlek[j] = 0.0, 0.0, 0.0
                else:
                    a = raw_axis(j, th)
                    axis[j], theta[j], circlek[j] = a / np.linalg.norm(a), th, 0.0
            else:                                         # logrot(r, parent), Align.cpp:180-276
                if abs(th) <= 1e-6:
                    axis[j], theta[j], circlek[j] = axis[fa], 0.0, circlek[fa]
                else:
                    a = axis[fa].copy() if abs(th - np.pi) <= 1e-6 else raw_axis(j, th)
                    if float(a @ axis[fa]) < 0.0:
                        a, th = -a, 2.0 * np.pi - th
                    axis[j], theta[j], circlek[j] = a / np.linalg.norm(a), th, circlek[fa]
                if abs(th - theta[fa]) > np.pi:
                    circlek[j] += -1.0 if theta[fa] < np.pi else 1.0
            out[j] = _skew(axis[j], circlek[j] * 2.0 * np.pi + theta[j]).T
            for e in range(int(ring_off[j]), int(ring_off[j + 1])):
                k = int(ring[e])
                if not visited[k]:
                    visited[k] = True
                    queue.append((k, j))
    return out
