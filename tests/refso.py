"""ctypes view of oracle/_ref/libsxsref.so (the UNMODIFIED reference compiled by oracle/Makefile).

TEST INFRASTRUCTURE: only tests/, bench.py's cpu_baseline / --impl reference leg and
__graft_entry__.smoke() may import this.  Entry points are the flat-array wrappers of
oracle/ref_harness.c.
"""
import ctypes as C
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(REPO, "oracle", "_ref", "libsxsref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available():
    return os.path.exists(REF_SO)


# the same unmodified sources compiled WITH FMA contraction (oracle/Makefile, target sens): not a second oracle but
# the reference's own answer to a 1e-16-level perturbation of its arithmetic — the noise floor parity is read against
SENS_SO = os.path.join(REPO, "oracle", "_ref", "libsxsref_fma.so")

_libs = {}


def sens_available():
    return os.path.exists(SENS_SO)


def lib(path=None):
    path = path or REF_SO
    if path not in _libs:
        l = C.CDLL(path)
        l.ref_sbessel.restype = C.c_double
        l.ref_sbessel.argtypes = [C.c_int, C.c_double]
        l.ref_wigner_3j.restype = C.c_double
        l.ref_wigner_3j.argtypes = [C.c_int] * 6
        _libs[path] = l
    return _libs[path]


def dptr(a):
    return a.ctypes.data_as(_dp)


def iptr(a):
    return a.ctypes.data_as(_ip)


def _names(lst):
    arr = (C.c_char_p * len(lst))(*[s.encode() if isinstance(s, str) else s for s in lst])
    return arr


def mkarray(begin, end, n):
    out = np.zeros(n)
    lib().ref_mkarray(C.c_double(begin), C.c_double(end), C.c_int(n), dptr(out))
    return out


def load_pdb(pdb, prm, centre):
    cap = 200000
    xyz = np.zeros((cap, 3))
    rad = np.zeros(cap)
    res = C.create_string_buffer(8 * cap)
    atm = C.create_string_buffer(8 * cap)
    shift = np.zeros(3)
    n = lib().ref_load_pdb(pdb.encode(), prm.encode(), C.c_int(centre), C.c_int(cap), dptr(xyz), dptr(rad),
                           res, atm, dptr(shift))
    if n < 0:
        raise RuntimeError("ref_load_pdb failed")
    resn = [res.raw[8 * i:8 * i + 8].split(b"\0")[0].decode() for i in range(n)]
    atmn = [atm.raw[8 * i:8 * i + 8].split(b"\0")[0].decode() for i in range(n)]
    return dict(xyz=xyz[:n].copy(), radius=rad[:n].copy(), res=resn, atm=atmn, shift=shift)


def form_factors(map_path, res, atm):
    n = len(res)
    ff = np.zeros((n, 3))
    bad = lib().ref_form_factors(map_path.encode(), C.c_int(n), _names(res), _names(atm), dptr(ff))
    return ff, bad


def expand(map_path, xyz, res, atm, radius, qvals, L, sa=None, water_mode=0):
    n = len(res)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    radius = np.ascontiguousarray(radius, dtype=np.float64)
    qvals = np.ascontiguousarray(qvals, dtype=np.float64)
    coef = np.zeros((3, len(qvals), (L + 1) ** 2, 2))
    rm = C.c_double(0)
    if sa is None:
        sa_arr = np.zeros(n)
    else:
        sa_arr = np.ascontiguousarray(sa, dtype=np.float64).copy()
    lib().ref_expand(map_path.encode(), C.c_int(n), dptr(xyz), _names(res), _names(atm), dptr(radius),
                     dptr(sa_arr), C.c_int(water_mode), dptr(qvals), C.c_int(len(qvals)), C.c_int(L),
                     dptr(coef), C.byref(rm))
    return coef, rm.value, sa_arr


def profile_read(path):
    cap = 100000
    q = np.zeros(cap)
    i = np.zeros(cap)
    e = np.zeros(cap)
    n = lib().ref_profile_read(path.encode(), C.c_int(cap), dptr(q), dptr(i), dptr(e))
    if n < 0:
        raise RuntimeError("cannot read " + path)
    return q[:n].copy(), i[:n].copy(), e[:n].copy()


def opt_params(exp_q, exp_in, exp_err, qvals, rm):
    a = np.zeros(6 * len(qvals))
    scal = np.zeros(3)
    lib().ref_opt_params(dptr(np.ascontiguousarray(exp_q)), dptr(np.ascontiguousarray(exp_in)),
                         dptr(np.ascontiguousarray(exp_err)), C.c_int(len(exp_q)),
                         dptr(np.ascontiguousarray(qvals)), C.c_int(len(qvals)), C.c_double(rm), dptr(a), dptr(scal))
    return a, scal


def scores(index_list, coefA, coefB, a, scal, qvals, zvals, L, skip=1, init=None, so=None):
    idx = np.ascontiguousarray(index_list, dtype=np.int32)
    n = len(idx)
    s = np.zeros(n) if init is None else init[0].copy()
    c1 = np.zeros(n) if init is None else init[1].copy()
    c2 = np.zeros(n) if init is None else init[2].copy()
    qvals = np.ascontiguousarray(qvals, dtype=np.float64)
    zvals = np.ascontiguousarray(zvals, dtype=np.float64)
    lib(so).ref_scores(dptr(s), dptr(c1), dptr(c2), iptr(idx), C.c_int(n), dptr(np.ascontiguousarray(coefA)),
                     dptr(np.ascontiguousarray(coefB)), dptr(np.ascontiguousarray(a)), dptr(np.ascontiguousarray(scal)),
                     dptr(qvals), C.c_int(len(qvals)), dptr(zvals), C.c_int(len(zvals)), C.c_int(L), C.c_int(skip))
    return s, c1, c2


def fit(x, a, scal, qvals, rescale=True, so=None):
    """x: [npts][6][qnum] cross terms -> [npts][4] = score, c1, c2, nfg"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    npts = x.shape[0]
    out = np.zeros((npts, 4))
    qvals = np.ascontiguousarray(qvals, dtype=np.float64)
    lib(so).ref_fit(dptr(x), C.c_int(npts), dptr(np.ascontiguousarray(a)), dptr(np.ascontiguousarray(scal)),
                  dptr(qvals), C.c_int(len(qvals)), C.c_int(1 if rescale else 0), dptr(out))
    return out


def profile_from_spf(coef, L, rm, qvals, c1, c2):
    qvals = np.ascontiguousarray(qvals, dtype=np.float64)
    n = len(qvals)
    i = np.zeros(n)
    e = np.zeros(n)
    lib().ref_profile_from_spf(dptr(np.ascontiguousarray(coef)), C.c_int(n), C.c_int(L), C.c_double(rm), dptr(qvals),
                               C.c_double(c1), C.c_double(c2), dptr(i), dptr(e))
    return i, e


def fitted_profile(coef, L, a, scal, qvals):
    qvals = np.ascontiguousarray(qvals, dtype=np.float64)
    n = len(qvals)
    i = np.zeros(n)
    e = np.zeros(n)
    o = np.zeros(3)
    lib().ref_fitted_profile(dptr(np.ascontiguousarray(coef)), C.c_int(n), C.c_int(L), dptr(np.ascontiguousarray(a)),
                             dptr(np.ascontiguousarray(scal)), dptr(qvals), dptr(i), dptr(e), dptr(o))
    return i, e, o


def ft2euler(tv, rm, ref_lig):
    out = np.zeros(6)
    lib().ref_ft2euler(dptr(np.ascontiguousarray(tv, dtype=np.float64)), dptr(np.ascontiguousarray(rm, dtype=np.float64)),
                       dptr(np.ascontiguousarray(ref_lig, dtype=np.float64)), dptr(out))
    return out


def wigner_d(L, beta):
    out = np.zeros((L + 1, 2 * L + 1, 2 * L + 1))
    lib().ref_wigner_d(C.c_int(L), C.c_double(beta), dptr(out))
    return out


def sbessel(l, x):
    return lib().ref_sbessel(int(l), float(x))


def wigner_3j(j1, j2, j3, m1, m2, m3):
    return lib().ref_wigner_3j(j1, j2, j3, m1, m2, m3)
