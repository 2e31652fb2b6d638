"""ctypes view of oracle/liboracle_port.so — the plain-C restatement of the reference path (oracle/port).
TEST INFRASTRUCTURE: imported by tests/ only."""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(REPO, "oracle", "liboracle_port.so")
_dp = C.POINTER(C.c_double)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_SO):
            subprocess.run(["make", "-s", "-C", os.path.join(REPO, "oracle"), "port"], check=True)
        _lib = C.CDLL(PORT_SO)
        _lib.port_sbessel.restype = C.c_double
        _lib.port_sbessel.argtypes = [C.c_int, C.c_double]
        _lib.port_wigner_3j.restype = C.c_double
        _lib.port_wigner_3j.argtypes = [C.c_int] * 6
    return _lib


def dptr(a):
    return a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def sbessel(l, x):
    return lib().port_sbessel(int(l), float(x))


def mkarray(b, e, n):
    out = np.zeros(n)
    lib().port_mkarray(C.c_double(b), C.c_double(e), C.c_int(n), dptr(out))
    return out


def expand(xyz, ff, qvals, L):
    """ff[natoms][3] = vacuum, dummy, water(=h2o*sa) factors"""
    xyz, ff, qvals = _c(xyz), _c(ff), _c(qvals)
    n = len(xyz)
    coef = np.zeros((3, len(qvals), (L + 1) ** 2, 2))
    fv, fd, fw = _c(ff[:, 0]), _c(ff[:, 1]), _c(ff[:, 2])
    lib().port_expand(C.c_int(n), dptr(xyz), dptr(fv), dptr(fd), dptr(fw), dptr(qvals), C.c_int(len(qvals)), C.c_int(L), dptr(coef))
    return coef


def wigner_d(L, beta):
    out = np.zeros((L + 1, 2 * L + 1, 2 * L + 1))
    lib().port_wigner_d(C.c_int(L), C.c_double(beta), dptr(out))
    return out


def dsymb(L):
    out = np.zeros(((L + 1) ** 2, L + 1, 2 * L + 1))
    lib().port_dsymb(C.c_int(L), dptr(out))
    return out


def opt_params(eq, ei, ee, qvals, rm):
    a = np.zeros(6 * len(qvals))
    scal = np.zeros(3)
    lib().port_opt_params(dptr(_c(eq)), dptr(_c(ei)), dptr(_c(ee)), C.c_int(len(eq)), dptr(_c(qvals)), C.c_int(len(qvals)),
                          C.c_double(rm), dptr(a), dptr(scal))
    return a, scal


def fit(x, a, scal, qvals):
    x = _c(x)
    out = np.zeros((len(x), 4))
    lib().port_fit(dptr(x), C.c_int(len(x)), dptr(_c(a)), dptr(_c(scal)), dptr(_c(qvals)), C.c_int(len(qvals)), dptr(out))
    return out


def scores(index, coefA, coefB, a, scal, qvals, zvals, L, want_cross=False, init=None):
    idx = np.ascontiguousarray(index, dtype=np.int64)
    n = len(idx)
    s = np.zeros(n) if init is None else init[0].copy()
    c1 = np.zeros(n) if init is None else init[1].copy()
    c2 = np.zeros(n) if init is None else init[2].copy()
    qvals, zvals = _c(qvals), _c(zvals)
    cross = np.zeros((n, 6, len(qvals))) if want_cross else None
    lib().port_scores(dptr(s), dptr(c1), dptr(c2), idx.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_int(n), dptr(_c(coefA)),
                      dptr(_c(coefB)), dptr(_c(a)), dptr(_c(scal)), dptr(qvals), C.c_int(len(qvals)), dptr(zvals),
                      C.c_int(len(zvals)), C.c_int(L), dptr(cross) if want_cross else None)
    return (s, c1, c2, cross) if want_cross else (s, c1, c2)
