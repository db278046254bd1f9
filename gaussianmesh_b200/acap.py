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
        if not _R:
            raise NotImplementedError("only _R = 1 (rotation matrices) is implemented; the reference never asks for log-rotations")
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
        return R, S
