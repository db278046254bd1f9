"""TEST INFRASTRUCTURE ONLY -- numpy float64 restatement of pyACAP.GetRS(ref, deformed, _R=1, ncpu)
(reference ACAP/pyACAPv1.zip: mainpy.cpp:60-64 -> src/FeatureVector.cpp RefMesh::RefMesh :81-173 and
RefMesh::GetRS :428-590, src/Align.cpp AffineAlign :60-100 and polarDec :31-56).

Per vertex i with one-ring neighbours j (cyclic order):
    rest:      p_k = w_k^(1/2) (x_j - x_i),  w_k = sqrt(sexp((cot a + cot b) / 2))      (:121-156, sexp :69-72)
               p_n = normal_i * (mean |x_j - x_i|) * 0.3                                  (:164, normalScale :22)
               AtA^-1 = (sum p p^T)^-1                                                    (Align.cpp:60-75)
    deformed:  v_k = w_k^(1/2) (x'_j - x'_i),  v_n from the deformed normal and lengths   (:453-465)
               T = (AtA^-1 sum_k p_k v_k^T)^T   (least-squares affine map, T p ~ v)        (Align.cpp:77-95)
               T = r s  (polar decomposition by SVD, reflection moved to the smallest singular value, Align.cpp:31-56)
    returns    R = r^T flattened row-major, S = s                                         (:560-590)

Pinned against the golden vectors that ship in the reference zip (tests/golden/acap_1_to_2.npz, see
tests/golden/make_acap_golden.py).  Vertex normals follow OpenMesh's update_normals(): unit face normals
averaged per vertex, then normalised.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

NORMAL_SCALE = 0.3
EPS = 1e-10


def one_rings(num_vertices: int, F: np.ndarray):
    """Cyclically ordered one-ring neighbours of every vertex (closed fans: a cycle; boundary fans: a chain starting
    at a boundary edge, as OpenMesh's circulators deliver them).  Returns (offsets [Vn+1], neighbours [E])."""
    nxt = [dict() for _ in range(num_vertices)]     # around v: neighbour a -> next neighbour b for each face (v, a, b)
    for a, b, c in F.tolist():
        nxt[a][b] = c
        nxt[b][c] = a
        nxt[c][a] = b
    offsets, nbrs = [0], []
    for v in range(num_vertices):
        m = nxt[v]
        if not m:
            offsets.append(len(nbrs))
            continue
        targets = set(m.values())
        starts = [a for a in m if a not in targets]          # chain heads exist only on the boundary
        a = starts[0] if starts else next(iter(m))
        ring, seen = [], set()
        while a is not None and a not in seen:
            ring.append(a)
            seen.add(a)
            a = m.get(a)
        nbrs.extend(ring)
        offsets.append(len(nbrs))
    return np.asarray(offsets, np.int64), np.asarray(nbrs, np.int64)


def vertex_normals(V: np.ndarray, F: np.ndarray) -> np.ndarray:
    fn = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    fn = fn / np.maximum(np.linalg.norm(fn, axis=1, keepdims=True), 1e-300)
    vn = np.zeros_like(V)
    for k in range(3):
        np.add.at(vn, F[:, k], fn)
    return vn / np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-300)


def _cotan(a, b):
    na, nb = np.linalg.norm(a), np.linalg.norm(b)
    if na < EPS or nb < EPS:
        return 0.0
    c = float(a @ b) / (na * nb)
    if c == 1:
        return 1.0
    return c / np.sqrt(1 - c * c)


def _sexp(x):
    return np.exp(x) if x <= 0 else 1 + x


def rest_state(V: np.ndarray, F: np.ndarray) -> Dict[str, np.ndarray]:
    """RefMesh::RefMesh (:81-173): ring structure, the fourth-root cotangent weights and AtA^-1 per vertex."""
    V = np.asarray(V, np.float64)
    Vn = V.shape[0]
    off, nbr = one_rings(Vn, F)
    sw = np.ones(nbr.shape[0])                 # sqrt(w_k): the factor applied to the edge vectors
    normals = vertex_normals(V, F)
    ata_inv = np.zeros((Vn, 3, 3))
    for i in range(Vn):
        ring = nbr[off[i]:off[i + 1]]
        n = len(ring)
        vec = []
        lens = 0.0
        for k in range(n):
            prev_, next_ = V[ring[(k + n - 1) % n]], V[ring[(k + 1) % n]]
            w1 = _cotan(V[i] - prev_, V[ring[k]] - prev_)
            w2 = _cotan(V[i] - next_, V[ring[k]] - next_)
            w = np.sqrt(_sexp(0.5 * (w1 + w2)))
            if w != w or w > 100000:
                w = 1.0
            sw[off[i] + k] = np.sqrt(w)
            q = V[ring[k]] - V[i]
            lens += np.linalg.norm(q)
            vec.append(sw[off[i] + k] * q)
        if n == 0:
            continue
        vec.append(normals[i] * (lens / n * NORMAL_SCALE))
        P = np.asarray(vec)
        ata_inv[i] = np.linalg.inv(P.T @ P)
    return {"offsets": off, "neighbours": nbr, "sqrt_w": sw, "ata_inv": ata_inv, "V": V, "F": np.asarray(F), "normals": normals}


def polar_dec(a: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Align.cpp:31-56"""
    U, sv, Vt = np.linalg.svd(a)
    r = U @ Vt
    s = Vt.T @ np.diag(sv) @ Vt
    if np.linalg.det(r) < 0:
        k = int(np.argmin(sv))
        U = U.copy()
        U[:, k] *= -1
        sv = sv.copy()
        sv[k] *= -1
        r = U @ Vt
        s = Vt.T @ np.diag(sv) @ Vt
    return r, s


def get_rs(rest: Dict[str, np.ndarray], V_deformed: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """RefMesh::GetRS(ms, _R=1, ncpu) (:428-590): returns (R [Vn,3,3] = r^T, S [Vn,3,3])."""
    V0, F, off, nbr, sw = rest["V"], rest["F"], rest["offsets"], rest["neighbours"], rest["sqrt_w"]
    Vd = np.asarray(V_deformed, np.float64)
    n0, nd = rest["normals"], vertex_normals(Vd, F)
    Vn = V0.shape[0]
    R = np.tile(np.eye(3), (Vn, 1, 1))
    S = np.tile(np.eye(3), (Vn, 1, 1))
    for i in range(Vn):
        ring = nbr[off[i]:off[i + 1]]
        n = len(ring)
        if n == 0:
            continue
        w = sw[off[i]:off[i + 1], None]
        q0, qd = V0[ring] - V0[i], Vd[ring] - Vd[i]
        p = np.vstack([w * q0, n0[i] * (np.linalg.norm(q0, axis=1).sum() / n * NORMAL_SCALE)])
        v = np.vstack([w * qd, nd[i] * (np.linalg.norm(qd, axis=1).sum() / n * NORMAL_SCALE)])
        T = (rest["ata_inv"][i] @ (p.T @ v)).T
        r, s = polar_dec(T)
        R[i], S[i] = r.T, s
    return R, S
