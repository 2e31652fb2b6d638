"""ctypes binding of libfmftsaxs.so — the stub a Python host writes against include/fmftsaxs/*.h.

Three layers of the same library are reachable:
  * flat adapters of the reference-shaped API (sxs_flat.h): ``expand``, ``scores``, … — what a user of
    the reference would call (they run sxs_compute_saxs_scores etc. underneath);
  * the CUDA C-ABI (sxs_cuda.h): ``Plan`` and the ``cuda_*`` functions, plain pointers and sizes;
  * raw access via ``lib()`` for anything else.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SXS_LIB_PATH") or os.path.join(PKG, "libfmftsaxs.so")  # override: tuning builds only

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_llp = C.POINTER(C.c_longlong)

_lib = None


class BuildError(RuntimeError):
    pass


def lib():
    """Loads the shared library; no fallback — a missing build is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BuildError(
                "libfmftsaxs.so is not built (run `python -m libfmftsaxs_b200.build`); "
                "this package has no CPU or pure-Python path")
        _lib = C.CDLL(LIB_PATH)
        _lib.sxs_cuda_last_error.restype = C.c_char_p
        _lib.sxs_cuda_plan_create.restype = C.c_void_p
        _lib.sxs_cuda_plan_create.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        _lib.sxs_cuda_plan_destroy.argtypes = [C.c_void_p]
        _lib.sxs_cuda_plan_set_molecules.argtypes = [C.c_void_p, _dp, _dp]
        _lib.sxs_cuda_plan_set_experiment.argtypes = [C.c_void_p, _dp, C.c_double, C.c_double]
        _lib.sxs_cuda_plan_set_translations.argtypes = [C.c_void_p, _dp, C.c_int]
        _lib.sxs_cuda_plan_score_i32.argtypes = [C.c_void_p, _ip, C.c_longlong, C.c_int, C.c_int, _dp, _dp, _dp]
        _lib.sxs_cuda_plan_score_i64.argtypes = [C.c_void_p, _llp, C.c_longlong, C.c_int, C.c_int, _dp, _dp, _dp]
        _lib.sxs_cuda_plan_score_dev_i32.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.sxs_cuda_plan_score_dev_i64.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.sxs_cuda_plan_stats.argtypes = [C.c_void_p, _llp]
        _lib.sxs_cuda_plan_fit_evaluations.argtypes = [C.c_void_p, _llp]
        _lib.sxs_cuda_plan_set_profiling.argtypes = [C.c_void_p, C.c_int]
        _lib.sxs_cuda_plan_kernel_times.argtypes = [C.c_void_p, _dp, _llp]
        _lib.sxs_cuda_plan_cross_terms_i32.argtypes = [C.c_void_p, _ip, C.c_longlong, _dp]
        _lib.sxs_cuda_plan_scan_topk.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _llp, _dp, _dp, _dp]
        for f in (_lib.sxs_cuda_plan_set_molecules, _lib.sxs_cuda_plan_set_experiment, _lib.sxs_cuda_plan_set_translations,
                  _lib.sxs_cuda_plan_score_i32, _lib.sxs_cuda_plan_score_i64, _lib.sxs_cuda_plan_score_dev_i32,
                  _lib.sxs_cuda_plan_score_dev_i64, _lib.sxs_cuda_plan_stats, _lib.sxs_cuda_plan_fit_evaluations,
                  _lib.sxs_cuda_plan_set_profiling, _lib.sxs_cuda_plan_kernel_times, _lib.sxs_cuda_plan_cross_terms_i32,
                  _lib.sxs_cuda_plan_scan_topk):
            f.restype = C.c_int
        _lib.sxs_cuda_plan_destroy.restype = None
        _lib.sxs_sbessel.restype = C.c_double
        _lib.sxs_sbessel.argtypes = [C.c_int, C.c_double]
        _lib.sxs_wigner_3j.restype = C.c_double
        _lib.sxs_wigner_3j.argtypes = [C.c_int] * 6
    return _lib


def last_error():
    return lib().sxs_cuda_last_error().decode()


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, last_error()))


def dptr(a):
    return a.ctypes.data_as(_dp)


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


def _names(lst):
    return (C.c_char_p * len(lst))(*[s.encode() if isinstance(s, str) else s for s in lst])


def device_count():
    return lib().sxs_cuda_device_count()


def fp64_peak(device=0):
    """measured DFMA throughput in TFLOP/s"""
    v = C.c_double(0)
    _check(lib().sxs_cuda_fp64_peak(C.c_int(device), C.byref(v)), "sxs_cuda_fp64_peak")
    return v.value


def cuda_exp(x, device=0):
    """exp() as the device evaluates it inside the objective (sxs_cuda_exp_array)"""
    x = _c(x)
    y = np.zeros_like(x)
    f = lib().sxs_cuda_exp_array
    f.argtypes = [C.c_int, _dp, C.c_longlong, _dp]
    f.restype = C.c_int
    _check(f(device, dptr(x), x.size, dptr(y)), "sxs_cuda_exp_array")
    return y


def cuda_ft_rows_to_indices(rot_id, trans, rots, ref_lig, zvals, L, device=0):
    """ft rows -> flat 64-bit grid indices on the device (sxs_cuda_ft_rows_to_indices): -1 = z not on the table,
    -2 = rotation id outside the table.  rots: (nrot, 9) row-major matrices."""
    rot_id = np.ascontiguousarray(rot_id, dtype=np.int32)
    trans = _c(np.asarray(trans, dtype=np.float64).reshape(-1, 3))
    rots = _c(np.asarray(rots, dtype=np.float64).reshape(-1, 9))
    ref_lig = _c(np.asarray(ref_lig, dtype=np.float64).reshape(3))
    zvals = _c(zvals)
    out = np.zeros(len(rot_id), dtype=np.int64)
    f = lib().sxs_cuda_ft_rows_to_indices
    f.argtypes = [C.c_int, _ip, _dp, C.c_longlong, _dp, C.c_longlong, _dp, _dp, C.c_int, C.c_int, _llp]
    f.restype = C.c_int
    _check(f(device, rot_id.ctypes.data_as(_ip), dptr(trans), len(rot_id), dptr(rots), len(rots), dptr(ref_lig), dptr(zvals),
             len(zvals), int(L), out.ctypes.data_as(_llp)), "sxs_cuda_ft_rows_to_indices")
    return out


# ------------------------------------------------------------------ reference-shaped API (flat adapters)

def mkarray(begin, end, n):
    """sxs_mkarray"""
    f = lib().sxs_mkarray
    f.restype = _dp
    f.argtypes = [C.c_double, C.c_double, C.c_int]
    p = f(begin, end, n)
    out = np.array([p[i] for i in range(n)])
    lib().free(p) if hasattr(lib(), "free") else None
    return out


def load_pdb(pdb, prm, centre):
    cap = 400000
    xyz = np.zeros((cap, 3))
    rad = np.zeros(cap)
    res = C.create_string_buffer(8 * cap)
    atm = C.create_string_buffer(8 * cap)
    shift = np.zeros(3)
    n = lib().sxs_flat_load_pdb(pdb.encode(), prm.encode(), C.c_int(centre), C.c_int(cap), dptr(xyz), dptr(rad),
                                res, atm, dptr(shift))
    if n < 0:
        raise RuntimeError("sxs_flat_load_pdb failed")
    resn = [res.raw[8 * i:8 * i + 8].split(b"\0")[0].decode() for i in range(n)]
    atmn = [atm.raw[8 * i:8 * i + 8].split(b"\0")[0].decode() for i in range(n)]
    return dict(xyz=xyz[:n].copy(), radius=rad[:n].copy(), res=resn, atm=atmn, shift=shift)


def expand(map_path, xyz, res, atm, radius, qvals, L, sa=None, water_mode=0):
    """atom_grp2spf_inplace (water_mode 0/1) or atom_grp2spf (2) -> (coef[3][Q][(L+1)^2][2], rm, sa)"""
    n = len(res)
    xyz, radius, qvals = _c(xyz), _c(radius), _c(qvals)
    coef = np.zeros((3, len(qvals), (L + 1) ** 2, 2))
    rm = C.c_double(0)
    sa_arr = np.zeros(n) if sa is None else _c(sa).copy()
    lib().sxs_flat_expand(map_path.encode(), C.c_int(n), dptr(xyz), _names(res), _names(atm), dptr(radius),
                          dptr(sa_arr), C.c_int(water_mode), dptr(qvals), C.c_int(len(qvals)), C.c_int(L),
                          dptr(coef), C.byref(rm))
    return coef, rm.value, sa_arr


def profile_read(path):
    cap = 100000
    q, i, e = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    n = lib().sxs_flat_profile_read(path.encode(), C.c_int(cap), dptr(q), dptr(i), dptr(e))
    if n < 0:
        raise RuntimeError("cannot read " + path)
    return q[:n].copy(), i[:n].copy(), e[:n].copy()


def opt_params(exp_q, exp_in, exp_err, qvals, rm):
    a = np.zeros(6 * len(qvals))
    scal = np.zeros(3)
    lib().sxs_flat_opt_params(dptr(_c(exp_q)), dptr(_c(exp_in)), dptr(_c(exp_err)), C.c_int(len(exp_q)),
                              dptr(_c(qvals)), C.c_int(len(qvals)), C.c_double(rm), dptr(a), dptr(scal))
    return a, scal


def scores(index_list, coefA, coefB, a, scal, qvals, zvals, L, skip=1, init=None):
    """sxs_compute_saxs_scores (int32 indices) / sxs_compute_saxs_scores64 (int64 indices)"""
    idx = np.ascontiguousarray(index_list)
    n = len(idx)
    s = np.zeros(n) if init is None else init[0].copy()
    c1 = np.zeros(n) if init is None else init[1].copy()
    c2 = np.zeros(n) if init is None else init[2].copy()
    qvals, zvals = _c(qvals), _c(zvals)
    args_tail = (dptr(_c(coefA)), dptr(_c(coefB)), dptr(_c(a)), dptr(_c(scal)), dptr(qvals), C.c_int(len(qvals)),
                 dptr(zvals), C.c_int(len(zvals)), C.c_int(L), C.c_int(skip))
    if idx.dtype == np.int64:
        lib().sxs_flat_scores64(dptr(s), dptr(c1), dptr(c2), idx.ctypes.data_as(_llp), C.c_longlong(n), *args_tail)
    else:
        idx = idx.astype(np.int32)
        lib().sxs_flat_scores(dptr(s), dptr(c1), dptr(c2), idx.ctypes.data_as(_ip), C.c_int(n), *args_tail)
    return s, c1, c2


def profile_from_spf(coef, L, rm, qvals, c1, c2):
    qvals = _c(qvals)
    n = len(qvals)
    i, e = np.zeros(n), np.zeros(n)
    lib().sxs_flat_profile_from_spf(dptr(_c(coef)), C.c_int(n), C.c_int(L), C.c_double(rm), dptr(qvals),
                                    C.c_double(c1), C.c_double(c2), dptr(i), dptr(e))
    return i, e


def fitted_profile(coef, L, a, scal, qvals):
    qvals = _c(qvals)
    n = len(qvals)
    i, e, o = np.zeros(n), np.zeros(n), np.zeros(3)
    lib().sxs_flat_fitted_profile(dptr(_c(coef)), C.c_int(n), C.c_int(L), dptr(_c(a)), dptr(_c(scal)), dptr(qvals),
                                  dptr(i), dptr(e), dptr(o))
    return i, e, o


def ft2euler(tv, rm, ref_lig):
    out = np.zeros(6)
    lib().sxs_flat_ft2euler(dptr(_c(tv)), dptr(_c(rm)), dptr(_c(ref_lig)), dptr(out))
    return out


def euler_to_index(euler, z_index, L):
    euler = _c(euler).reshape(-1, 6)
    z_index = np.ascontiguousarray(z_index, dtype=np.int32)
    out = np.zeros(len(euler), dtype=np.int32)
    lib().sxs_flat_euler_to_index(dptr(euler), z_index.ctypes.data_as(_ip), C.c_int(len(euler)), C.c_int(L),
                                  out.ctypes.data_as(_ip))
    return out


def ft_rows_to_indices(rot_id, trans, rot_mats, ref_lig, zvals, L, nthreads=0):
    """in-memory ft rows -> (flat index int64, ft id, serial number) of the rows on the z table
    (include/fmftsaxs/index.h: sxs_ft_rows_to_indices64)"""
    rot_id = np.ascontiguousarray(rot_id, dtype=np.int32)
    trans = _c(trans).reshape(-1, 3)
    rot_mats = _c(rot_mats).reshape(-1, 9)
    zvals = _c(zvals)
    n = len(rot_id)
    index = np.zeros(n, dtype=np.int64)
    ft_id = np.zeros(n, dtype=np.int32)
    order = np.zeros(n, dtype=np.int32)
    f = lib().sxs_flat_ft_rows_to_indices
    f.restype = C.c_longlong
    kept = f(index.ctypes.data_as(_llp), ft_id.ctypes.data_as(_ip), order.ctypes.data_as(_ip), rot_id.ctypes.data_as(_ip),
             dptr(trans), C.c_longlong(n), dptr(rot_mats), C.c_int(len(rot_mats)), dptr(_c(ref_lig)), dptr(zvals),
             C.c_int(len(zvals)), C.c_int(L), C.c_int(nthreads))
    return index[:kept], ft_id[:kept], order[:kept]


def ft_file2euler_file(eu_path, ft_path, rm_path, ref_lig):
    lib().sxs_flat_ft_file2euler_file(str(eu_path).encode(), str(ft_path).encode(), str(rm_path).encode(), dptr(_c(ref_lig)))


def wigner_d(L, beta):
    out = np.zeros((L + 1, 2 * L + 1, 2 * L + 1))
    lib().sxs_flat_wigner_d(C.c_int(L), C.c_double(beta), dptr(out))
    return out


def tables(L):
    nb, N = L + 1, 2 * L + 1
    ds = np.zeros(nb * nb * nb * N)
    dw = np.zeros(nb * nb * N * N)
    tw = np.zeros(2 * N)
    lib().sxs_flat_tables(C.c_int(L), dptr(ds), dptr(dw), dptr(tw))
    return ds, dw, tw


def bessel_table(zvals, qvals, L):
    zvals, qvals = _c(zvals), _c(qvals)
    out = np.zeros(len(zvals) * len(qvals) * (2 * L + 1))
    lib().sxs_flat_bessel_table(dptr(zvals), C.c_int(len(zvals)), dptr(qvals), C.c_int(len(qvals)), C.c_int(L),
                                dptr(out))
    return out


def sbessel(l, x):
    return lib().sxs_sbessel(int(l), float(x))


def wigner_3j(j1, j2, j3, m1, m2, m3):
    return lib().sxs_wigner_3j(j1, j2, j3, m1, m2, m3)


# ------------------------------------------------------------------ CUDA C-ABI

def cuda_fit_profiles(x, a, qvals, mult, peak, rescale=True, device=0):
    """x[npts][6][qnum] cross terms -> [npts][4] = chi, c1, c2, evaluations (kernel K4 alone)"""
    x = _c(x)
    npts = x.shape[0]
    qvals = _c(qvals)
    out = np.zeros((npts, 4))
    _check(lib().sxs_cuda_fit_profiles(C.c_int(device), dptr(x), C.c_longlong(npts), dptr(_c(a)), dptr(qvals),
                                       C.c_int(len(qvals)), C.c_double(mult), C.c_double(peak),
                                       C.c_int(1 if rescale else 0), dptr(out)), "sxs_cuda_fit_profiles")
    return out


class Plan:
    """sxs_cuda_plan: device-resident scoring state for one (L, q grid) on one GPU."""

    def __init__(self, L, qvals, device=0):
        self.L, self.N, self.nb = L, 2 * L + 1, L + 1
        self.qvals = _c(qvals)
        self.device = device
        ds, dw, tw = tables(L)
        self._h = lib().sxs_cuda_plan_create(device, L, len(self.qvals), dptr(self.qvals), dptr(ds), dptr(dw), dptr(tw))
        if not self._h:
            raise RuntimeError("sxs_cuda_plan_create failed: " + last_error())
        self.znum = 0

    def close(self):
        if self._h:
            lib().sxs_cuda_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_molecules(self, coefA, coefB):
        _check(lib().sxs_cuda_plan_set_molecules(self._h, dptr(_c(coefA)), dptr(_c(coefB))), "set_molecules")

    def set_experiment(self, a, mult, peak):
        _check(lib().sxs_cuda_plan_set_experiment(self._h, dptr(_c(a)), mult, peak), "set_experiment")

    def set_translations(self, zvals):
        bes = bessel_table(zvals, self.qvals, self.L)
        self.znum = len(zvals)
        _check(lib().sxs_cuda_plan_set_translations(self._h, dptr(bes), self.znum), "set_translations")

    def score(self, index, z_lo=0, z_hi=None, init=None):
        """host buffers in, host buffers out (H2D/D2H inside)"""
        idx = np.ascontiguousarray(index)
        n = len(idx)
        z_hi = self.znum if z_hi is None else z_hi
        s = np.zeros(n) if init is None else init[0].copy()
        c1 = np.zeros(n) if init is None else init[1].copy()
        c2 = np.zeros(n) if init is None else init[2].copy()
        if idx.dtype == np.int64:
            rc = lib().sxs_cuda_plan_score_i64(self._h, idx.ctypes.data_as(_llp), n, z_lo, z_hi, dptr(s), dptr(c1), dptr(c2))
        else:
            idx = idx.astype(np.int32)
            rc = lib().sxs_cuda_plan_score_i32(self._h, idx.ctypes.data_as(_ip), n, z_lo, z_hi, dptr(s), dptr(c1), dptr(c2))
        _check(rc, "sxs_cuda_plan_score")
        return s, c1, c2

    def score_device(self, d_index_ptr, n, d_scores_ptr, d_c1_ptr, d_c2_ptr, stream_ptr=0, z_lo=0, z_hi=None, i64=False):
        """raw device pointers (e.g. torch tensor .data_ptr()); work is queued on `stream_ptr`"""
        z_hi = self.znum if z_hi is None else z_hi
        f = lib().sxs_cuda_plan_score_dev_i64 if i64 else lib().sxs_cuda_plan_score_dev_i32
        _check(f(self._h, d_index_ptr, n, z_lo, z_hi, d_scores_ptr, d_c1_ptr, d_c2_ptr, stream_ptr), "score_device")

    def scan_topk(self, k, z_lo=0, z_hi=None):
        """score every grid point of the z steps [z_lo, z_hi) and return the k best (index int64, chi, c1, c2)"""
        z_hi = self.znum if z_hi is None else z_hi
        idx = np.zeros(k, dtype=np.int64)
        s, c1, c2 = np.zeros(k), np.zeros(k), np.zeros(k)
        _check(lib().sxs_cuda_plan_scan_topk(self._h, C.c_int(z_lo), C.c_int(z_hi), C.c_int(k), idx.ctypes.data_as(_llp),
                                             dptr(s), dptr(c1), dptr(c2)), "scan_topk")
        return idx, s, c1, c2

    def cross_terms(self, index):
        idx = np.ascontiguousarray(index, dtype=np.int32)
        out = np.zeros((len(idx), 6, len(self.qvals)))
        _check(lib().sxs_cuda_plan_cross_terms_i32(self._h, idx.ctypes.data_as(_ip), len(idx), dptr(out)), "cross_terms")
        return out

    def set_profiling(self, on=True):
        _check(lib().sxs_cuda_plan_set_profiling(self._h, 1 if on else 0), "set_profiling")

    def kernel_times(self):
        """device ms and timed launches per kernel class since the last call"""
        ms = (C.c_double * 5)()
        n = (C.c_longlong * 5)()
        _check(lib().sxs_cuda_plan_kernel_times(self._h, ms, n), "kernel_times")
        names = ["sort", "translate", "cross", "fit", "scatter"]
        return {k: (ms[i], n[i]) for i, k in enumerate(names)}

    def fit_evaluations(self):
        """histogram of objective evaluations per fit of the last score call"""
        h = (C.c_longlong * 64)()
        _check(lib().sxs_cuda_plan_fit_evaluations(self._h, h), "fit_evaluations")
        return np.array(list(h))

    def stats(self):
        st = (C.c_longlong * 5)()
        lib().sxs_cuda_plan_stats(self._h, st)
        return dict(points=st[0], slabs=st[1], launches=st[2], cross_groups=st[3], groups=st[4])
