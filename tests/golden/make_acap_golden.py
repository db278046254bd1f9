"""Extract the ACAP golden vectors that ship inside the reference (ACAP/pyACAPv1.zip: test/1.obj -> test/2.obj,
test/LOGRNEW.txt, test/S.txt; compared -- without an assert -- by the zip's own test.py:86-130) into
tests/golden/acap_1_to_2.npz.  Run in the build container (needs /root/reference):

    python tests/golden/make_acap_golden.py

R_gold = expm of the second line of LOGRNEW.txt (test.py:88-93; the file stores the transposed log-rotation, so this
is what pyACAP.GetRS(..., _R=1) returns: the TRANSPOSE of the polar rotation), S_gold = second line of S.txt.
"""
import io
import os
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ZIP = "/root/reference/ACAP/pyACAPv1.zip"


def read_obj(text):
    V, F = [], []
    for line in text.splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            V.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            F.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.asarray(V, np.float64), np.asarray(F, np.int32)


def exp_r(logr):
    """test.py:6-16 (Rodrigues on the stored skew matrices)"""
    res = np.zeros_like(logr)
    theta = np.sqrt(logr[:, 0, 1] ** 2 + logr[:, 0, 2] ** 2 + logr[:, 1, 2] ** 2)
    for i in range(logr.shape[0]):
        if theta[i] == 0:
            res[i] = np.eye(3)
        else:
            x = logr[i] / theta[i]
            res[i] = np.eye(3) + x * np.sin(theta[i]) + x @ x * (1 - np.cos(theta[i]))
    return res


def main():
    z = zipfile.ZipFile(ZIP)
    V0, F = read_obj(z.read("test/1.obj").decode())
    V1, F1 = read_obj(z.read("test/2.obj").decode())
    assert np.array_equal(F, F1)
    logr = np.array([float(x) for x in z.read("test/LOGRNEW.txt").decode().splitlines()[1].split()]).reshape(-1, 3, 3)
    S = np.array([float(x) for x in z.read("test/S.txt").decode().splitlines()[1].split()]).reshape(-1, 3, 3)
    out = os.path.join(ROOT, "tests", "golden", "acap_1_to_2.npz")
    # logR_gold: the stored rows themselves = what GetRS(..., _R=0) returns (FeatureVector.cpp:531-557)
    np.savez_compressed(out, V_rest=V0, V_deformed=V1, F=F, R_gold=exp_r(logr), S_gold=S, logR_gold=logr)
    print(out, os.path.getsize(out), "bytes;", V0.shape, F.shape)


if __name__ == "__main__":
    main()
