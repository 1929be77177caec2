"""ctypes binding of include/ruviii.h (libruviii_b200.so).

The shared library is the product; this module only marshals numpy buffers into its C ABI.  There is no
fallback of any kind: if the library is missing or no sm_100 GPU is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libruviii_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_SOLVER, ERR_SINGULAR, ERR_UNSUPPORTED, ERR_STATE = range(8)
MODE_STACKED, MODE_FAST = 0, 1

# every symbol include/ruviii.h declares (tests/test_cabi.py checks the header against this list and the .so)
SYMBOLS = [
    "ruviii_abi_version", "ruviii_device_count", "ruviii_create", "ruviii_nccl_unique_id", "ruviii_create_rank",
    "ruviii_destroy", "ruviii_last_error", "ruviii_normalise", "ruviii_normalise_blocks", "ruviii_plan", "ruviii_upload", "ruviii_run",
    "ruviii_download", "ruviii_download_averaged", "ruviii_download_replicate_data", "ruviii_host_alloc", "ruviii_host_free", "ruviii_barrier", "ruviii_eigenvalues", "ruviii_set_fullalpha", "ruviii_set_eta", "ruviii_prps", "ruviii_set_prps_sets", "ruviii_pca", "ruviii_gene_stats", "ruviii_gene_spearman", "ruviii_regress_out",
    "ruviii_fast_residop", "ruviii_probe", "ruviii_mat_mult", "ruviii_mat_subtract", "ruviii_mat_transpose", "ruviii_mat_inverse",
]


class Params(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("mode", C.c_int32), ("apply_log", C.c_int32),
                ("average_rep", C.c_int32), ("pseudo_count", C.c_double), ("n_alpha_rows", C.c_int32),
                ("gram_side", C.c_int32), ("want_ps_rows", C.c_int32), ("keep_resident", C.c_int32), ("reserved", C.c_int32 * 6)]


class Info(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("m", C.c_int32), ("m1", C.c_int32), ("n_ps", C.c_int32),
                ("n", C.c_int32), ("p", C.c_int32), ("active_rows", C.c_int32), ("r", C.c_int32),
                ("n_ctl", C.c_int32), ("kmax_eff", C.c_int32), ("no_replicates", C.c_int32),
                ("gram_side", C.c_int32), ("n_shards", C.c_int32), ("kernel_launches", C.c_int32),
                ("ms_log", C.c_float), ("ms_residop", C.c_float), ("ms_gram", C.c_float),
                ("ms_gram_reduce", C.c_float), ("ms_allreduce_gram", C.c_float), ("ms_eigen", C.c_float),
                ("ms_bcast", C.c_float), ("ms_project", C.c_float), ("ms_ctl", C.c_float),
                ("ms_allreduce_ctl", C.c_float), ("ms_solve", C.c_float), ("ms_update", C.c_float),
                ("ms_total", C.c_float), ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
                ("eig_top", C.c_double), ("eig_kmax", C.c_double), ("eig_next", C.c_double),
                ("eig_solver", C.c_int32), ("eig_iterations", C.c_int32), ("gram_order", C.c_int32),
                ("pipelined", C.c_int32), ("host_t_enter", C.c_double), ("host_t_done", C.c_double),
                ("ms_gather", C.c_float), ("eig_rr", C.c_int32), ("eig_filter_products", C.c_int32), ("eig_chol2", C.c_int32)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class RuviiiError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("ruviii error %d: %s" % (code, message))
        self.code = code


_lib = None


def load_library():
    """dlopen the engine.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libruviii_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C ruviii_b200/csrc`. There is no CPU or PyTorch fallback for the RUV-III hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL if False else C.RTLD_LOCAL)
    dp, ip, bp, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_void_p
    H = C.c_void_p
    lib.ruviii_abi_version.restype = C.c_int
    lib.ruviii_device_count.restype = C.c_int
    lib.ruviii_create.argtypes = [C.POINTER(H), ip, C.c_int32]
    lib.ruviii_nccl_unique_id.argtypes = [vp]
    lib.ruviii_create_rank.argtypes = [C.POINTER(H), C.c_int32, C.c_int32, C.c_int32, vp]
    lib.ruviii_destroy.argtypes = [H]
    lib.ruviii_destroy.restype = None
    lib.ruviii_last_error.argtypes = [H]
    lib.ruviii_last_error.restype = C.c_char_p
    lib.ruviii_normalise.argtypes = [H, C.POINTER(Params), vp, C.c_int64, C.c_int32, C.c_int32, vp, C.c_int64,
                                     C.c_int32, vp, vp, vp, C.c_int32, vp, vp, vp, vp, C.POINTER(Info)]
    lib.ruviii_normalise_blocks.argtypes = [H, C.POINTER(Params), vp, C.c_int64, C.c_int32, C.c_int32, vp, C.c_int64,
                                            C.c_int32, vp, vp, vp, C.c_int32, C.POINTER(C.c_void_p), vp, vp, vp, C.POINTER(Info)]
    lib.ruviii_plan.argtypes = [H, C.POINTER(Params), C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, C.c_int32]
    lib.ruviii_upload.argtypes = [H, vp, C.c_int64, vp, C.c_int64]
    lib.ruviii_run.argtypes = [H, C.POINTER(Info)]
    lib.ruviii_download.argtypes = [H, vp, vp, vp, vp]
    lib.ruviii_download_averaged.argtypes = [H, vp]
    lib.ruviii_download_replicate_data.argtypes = [H, vp, C.c_int64]
    lib.ruviii_barrier.argtypes = [H]
    lib.ruviii_host_alloc.argtypes = [C.c_size_t]
    lib.ruviii_host_alloc.restype = C.c_void_p
    lib.ruviii_host_free.argtypes = [C.c_void_p]
    lib.ruviii_host_free.restype = None
    lib.ruviii_eigenvalues.argtypes = [H, vp, C.c_int32]
    lib.ruviii_set_fullalpha.argtypes = [H, vp, C.c_int32]
    lib.ruviii_set_eta.argtypes = [H, vp, C.c_int32, C.c_int32]
    lib.ruviii_prps.argtypes = [H, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double, vp, vp, C.c_int32, vp, C.c_int64]
    lib.ruviii_set_prps_sets.argtypes = [H, vp, vp, C.c_int32]
    lib.ruviii_gene_stats.argtypes = [H, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, vp, vp, vp, vp, vp]
    lib.ruviii_gene_spearman.argtypes = [H, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]
    lib.ruviii_regress_out.argtypes = [H, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, vp, C.c_int32, vp, C.c_int64,
                                       vp, vp]
    lib.ruviii_pca.argtypes = [H, vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                               C.c_int32, vp, vp, vp, vp]
    lib.ruviii_fast_residop.argtypes = [H, vp, C.c_int32, C.c_int32, vp, vp]
    lib.ruviii_probe.argtypes = [H, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    lib.ruviii_mat_mult.argtypes = [H, vp, C.c_int32, C.c_int32, vp, C.c_int32, vp]
    lib.ruviii_mat_subtract.argtypes = [H, vp, vp, C.c_int64, vp]
    lib.ruviii_mat_transpose.argtypes = [H, vp, C.c_int32, C.c_int32, vp]
    lib.ruviii_mat_inverse.argtypes = [H, vp, C.c_int32, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("ruviii_destroy", "ruviii_last_error", "ruviii_host_alloc", "ruviii_host_free"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _rowmajor(a, name):
    """A C-contiguous float64 2-D view; an F-ordered genes x samples assay must be passed as ``assay.T``."""
    a = np.asarray(a)
    if a.dtype != np.float64 or a.ndim != 2:
        raise TypeError("%s must be a 2-D float64 array" % name)
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    return a


def make_params(mode=MODE_STACKED, apply_log=False, pseudo_count=1.0, average_rep=False, n_alpha_rows=0,
                want_ps_rows=False, gram_side=0, keep_resident=False):
    p = Params()
    p.struct_size = C.sizeof(Params)
    p.mode = mode
    p.apply_log = int(bool(apply_log))
    p.pseudo_count = float(pseudo_count)
    p.average_rep = int(bool(average_rep))
    p.n_alpha_rows = int(n_alpha_rows)
    p.want_ps_rows = int(bool(want_ps_rows))
    p.gram_side = int(gram_side)
    p.keep_resident = int(bool(keep_resident))
    return p


class PinnedArray:
    """A float64 numpy array in page-locked host memory (ruviii_host_alloc); keep the object alive while in use."""

    def __init__(self, shape):
        lib = load_library()
        self._lib = lib
        n = int(np.prod(shape))
        self._ptr = lib.ruviii_host_alloc(max(n, 1) * 8)
        if not self._ptr:
            raise RuviiiError(ERR_CUDA, "ruviii_host_alloc failed (no CUDA device?)")
        buf = (C.c_double * max(n, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "_ptr", None):
                self.array = None
                self._lib.ruviii_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


class Engine:
    """Owns one ``ruviii_handle`` (device buffers, NCCL communicators, cuSOLVER handle)."""

    def __init__(self, devices=None, *, rank=None, world=None, device=None, nccl_id=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        if rank is None:
            devices = [0] if devices is None else list(devices)
            arr = (C.c_int32 * len(devices))(*devices)
            rc = self._lib.ruviii_create(C.byref(self._h), arr, len(devices))
        else:
            buf = None if nccl_id is None else C.create_string_buffer(bytes(nccl_id), 128)
            rc = self._lib.ruviii_create_rank(C.byref(self._h), int(device), int(rank), int(world), buf)
        if rc != OK:
            raise RuviiiError(rc, self._lib.ruviii_last_error(None).decode())
        self._plan = None
        self._regressed = None

    @staticmethod
    def nccl_unique_id():
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.ruviii_nccl_unique_id(buf)
        if rc != OK:
            raise RuviiiError(rc, lib.ruviii_last_error(None).decode())
        return buf.raw

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ruviii_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != OK:
            raise RuviiiError(rc, self._lib.ruviii_last_error(self._h).decode())

    # ---- resident API -----------------------------------------------------------------------------------
    def plan(self, params, m1, n, n_ps, group_of_row, ctl, ks):
        group_of_row = np.ascontiguousarray(group_of_row, dtype=np.int32)
        ctl = np.ascontiguousarray(np.asarray(ctl).astype(np.uint8))
        ks = np.ascontiguousarray(ks, dtype=np.int32)
        if ctl.shape[0] != n:
            raise RuviiiError(ERR_INVALID, "len(ctl) != n")
        need = n_ps if params.mode == MODE_FAST else m1 + n_ps
        if group_of_row.shape[0] != need:
            raise RuviiiError(ERR_INVALID, "group_of_row has %d entries, expected %d" % (group_of_row.shape[0], need))
        self._check(self._lib.ruviii_plan(self._h, C.byref(params), m1, n, n_ps, _ptr(group_of_row), _ptr(ctl),
                                          _ptr(ks), ks.shape[0]))
        self._plan = dict(params=params, m1=m1, n=n, n_ps=n_ps, ks=[int(k) for k in ks])

    def upload(self, Y, PS):
        """PS = None: replicate.data is built on the device from Y (needs set_prps_sets with the plan's n_ps)."""
        Y = _rowmajor(Y, "Y")
        if PS is None:
            self._check(self._lib.ruviii_upload(self._h, _ptr(Y), Y.strides[0] // 8, None, 0))
            return
        PS = _rowmajor(PS, "PS")
        self._check(self._lib.ruviii_upload(self._h, _ptr(Y), Y.strides[0] // 8, _ptr(PS), PS.strides[0] // 8))

    def run(self):
        info = Info()
        self._check(self._lib.ruviii_run(self._h, C.byref(info)))
        self._info = info
        return info

    def barrier(self):
        self._check(self._lib.ruviii_barrier(self._h))

    def set_fullalpha(self, fullalpha):
        fa = _rowmajor(fullalpha, "fullalpha")
        self._check(self._lib.ruviii_set_fullalpha(self._h, _ptr(fa), fa.shape[0]))

    def set_eta(self, eta):
        """ruv::RUV1 pre-step (ruvIII.R:116): eta is q x n row-major, intercept row included by the caller; None clears."""
        if eta is None:
            self._check(self._lib.ruviii_set_eta(self._h, None, 0, 0))
            return
        e = _rowmajor(eta, "eta")
        self._check(self._lib.ruviii_set_eta(self._h, _ptr(e), e.shape[0], e.shape[1]))

    @staticmethod
    def _csr(sets):
        """List of index lists -> (start int32[n+1], rows int32[total])."""
        start = np.zeros(len(sets) + 1, dtype=np.int32)
        start[1:] = np.cumsum([len(x) for x in sets])
        rows = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.int32) for x in sets]) if len(sets) else
                                    np.zeros(0, dtype=np.int32), dtype=np.int32)
        return start, rows

    def prps(self, Y, sets, apply_log=False, pseudo_count=1.0):
        """K0: pseudo-sample j = mean of the rows ``sets[j]`` of Y (samples x genes), after log2(Y + pseudo_count) when
        ``apply_log`` (prpsForCategoricalUV.R:91,123-139).  Returns n_ps x n."""
        Y = _rowmajor(Y, "Y")
        start, rows = self._csr(sets)
        out = np.empty((len(sets), Y.shape[1]), dtype=np.float64)
        self._check(self._lib.ruviii_prps(self._h, _ptr(Y), Y.strides[0] // 8, Y.shape[0], Y.shape[1], int(bool(apply_log)),
                                          float(pseudo_count), _ptr(start), _ptr(rows), len(sets), _ptr(out), out.shape[1]))
        return out

    def pca(self, X=None, k=10, slot=-1, apply_log=False, pseudo_count=1.0, center=True, scale=False, want_v=True):
        """computePCA.R:120-134 fast path.  X: samples x genes host matrix, or None for data already on the device
        (slot >= 0: output block of the last run, -1: the uploaded assay).  Returns dict(u m1 x k, d k, v n x k, stage_ms)."""
        if X is not None:
            X = _rowmajor(X, "X")
            m1, n, ldx, ptr = X.shape[0], X.shape[1], X.strides[0] // 8, _ptr(X)
        else:
            m1, n, ldx, ptr = self._plan["m1"], self._plan["n"], 0, None
        u = np.empty((m1, k), dtype=np.float64)
        d = np.empty(k, dtype=np.float64)
        v = np.empty((k, n), dtype=np.float64) if want_v else None
        ms = (C.c_float * 6)()
        self._check(self._lib.ruviii_pca(self._h, ptr, ldx, m1, n, int(slot), int(bool(apply_log)), float(pseudo_count),
                                         int(bool(center)), int(bool(scale)), int(k), _ptr(u), _ptr(d),
                                         _ptr(v) if want_v else None, C.cast(ms, C.c_void_p)))
        return dict(u=u, d=d, v=None if v is None else v.T, stage_ms=dict(zip(
            ["upload_log", "center", "gram", "eigen", "project", "total"], [float(x) for x in ms])))

    def regress_out(self, X=None, design=None, slot=-1, apply_log=False, pseudo_count=1.0, return_residuals=False):
        """K12 + K4 + K6: residuals of lm(X ~ design) for every gene (supervisedFindNCG.R:111-133).  ``design``: samples x p model
        matrix.  The residuals stay on the device -- ``gene_stats`` / ``gene_spearman`` read them with ``slot=-2`` -- and are
        also returned (samples x genes) with ``return_residuals``.  Returns dict(rank, ms[, residuals])."""
        D = np.ascontiguousarray(design, dtype=np.float64)
        if D.ndim != 2:
            raise RuviiiError(ERR_INVALID, "design must be a samples x p matrix")
        if X is not None:
            X = _rowmajor(X, "X")
            m1, n, ldx, ptr = X.shape[0], X.shape[1], X.strides[0] // 8, _ptr(X)
        else:
            m1, n, ldx, ptr = self._plan["m1"], self._plan["n"], 0, None
        if D.shape[0] != m1:
            raise RuviiiError(ERR_INVALID, "design needs one row per sample")
        res = np.empty((m1, n), dtype=np.float64) if return_residuals else None
        rank, ms = C.c_int32(0), C.c_float(0.0)
        self._check(self._lib.ruviii_regress_out(self._h, ptr, ldx, m1, n, int(slot), int(bool(apply_log)), float(pseudo_count), _ptr(D),
                                                 D.shape[1], None if res is None else _ptr(res), n, C.cast(C.byref(rank), C.c_void_p),
                                                 C.cast(C.byref(ms), C.c_void_p)))
        self._regressed = (m1, n)
        out = dict(rank=int(rank.value), ms=float(ms.value))
        if res is not None:
            out["residuals"] = res
        return out

    def gene_stats(self, X=None, groups=None, covariate=None, slot=-1, apply_log=False, pseudo_count=1.0):
        """K11: per-gene one-way ANOVA F across ``groups`` (ids < 0 are left out) and / or Pearson correlation with
        ``covariate`` (supervisedFindNCG.R:179-191,214-229).  Returns dict(F, r, ms)."""
        if X is not None:
            X = _rowmajor(X, "X")
            m1, n, ldx, ptr = X.shape[0], X.shape[1], X.strides[0] // 8, _ptr(X)
        elif slot == -2:
            (m1, n), ldx, ptr = self._regressed, 0, None
        else:
            m1, n, ldx, ptr = self._plan["m1"], self._plan["n"], 0, None
        g = None if groups is None else np.ascontiguousarray(groups, dtype=np.int32)
        c = None if covariate is None else np.ascontiguousarray(covariate, dtype=np.float64)
        if (g is not None and g.shape[0] != m1) or (c is not None and c.shape[0] != m1):
            raise RuviiiError(ERR_INVALID, "groups / covariate need one entry per sample")
        F = np.empty(n, dtype=np.float64) if g is not None else None
        r = np.empty(n, dtype=np.float64) if c is not None else None
        ms = C.c_float(0.0)
        self._check(self._lib.ruviii_gene_stats(self._h, ptr, ldx, m1, n, int(slot), int(bool(apply_log)), float(pseudo_count),
                                                None if g is None else _ptr(g), None if c is None else _ptr(c),
                                                None if F is None else _ptr(F), None if r is None else _ptr(r),
                                                C.cast(C.byref(ms), C.c_void_p)))
        return dict(F=F, r=r, ms=float(ms.value))

    def gene_spearman(self, X=None, covariate=None, slot=-1):
        """K11b: Spearman's rho of every gene with ``covariate`` (supervisedFindNCG.R:214-229, corr.method = "spearman")."""
        if X is not None:
            X = _rowmajor(X, "X")
            m1, n, ldx, ptr = X.shape[0], X.shape[1], X.strides[0] // 8, _ptr(X)
        elif slot == -2:
            (m1, n), ldx, ptr = self._regressed, 0, None
        else:
            m1, n, ldx, ptr = self._plan["m1"], self._plan["n"], 0, None
        c = np.ascontiguousarray(covariate, dtype=np.float64)
        if c.shape[0] != m1:
            raise RuviiiError(ERR_INVALID, "covariate needs one entry per sample")
        rho = np.empty(n, dtype=np.float64)
        ms = C.c_float(0.0)
        self._check(self._lib.ruviii_gene_spearman(self._h, ptr, ldx, m1, n, int(slot), _ptr(c), _ptr(rho), C.cast(C.byref(ms), C.c_void_p)))
        return dict(rho=rho, ms=float(ms.value))

    def set_prps_sets(self, sets):
        """Fused K0: the next plans with n_ps = len(sets) build replicate.data on the device when PS is None."""
        if not sets:
            self._check(self._lib.ruviii_set_prps_sets(self._h, None, None, 0))
            return
        start, rows = self._csr(sets)
        self._check(self._lib.ruviii_set_prps_sets(self._h, _ptr(start), _ptr(rows), len(sets)))

    def eigenvalues(self, count):
        out = np.empty(count, dtype=np.float64)
        self._check(self._lib.ruviii_eigenvalues(self._h, _ptr(out), count))
        return out

    def download_replicate_data(self):
        """replicate.data as the device holds it (uploaded, or built from the assay by ``set_prps_sets``): n_ps x n."""
        pl = self._plan
        out = np.empty((pl["n_ps"], pl["n"]), dtype=np.float64)
        self._check(self._lib.ruviii_download_replicate_data(self._h, _ptr(out), out.shape[1]))
        return out

    def download_averaged(self):
        pl, info = self._plan, self._info
        out = np.empty((len(pl["ks"]), info.p, pl["n"]), dtype=np.float64)
        self._check(self._lib.ruviii_download_averaged(self._h, _ptr(out)))
        return out

    def download(self, want_fullalpha=False, want_W=False, want_ps=False, out=None):
        pl, info = self._plan, self._info
        nk = len(pl["ks"])
        newY = out if out is not None else np.empty((nk, pl["m1"], pl["n"]), dtype=np.float64)
        skip = info.no_replicates or info.kmax_eff == 0
        newY_ps = np.empty((nk, pl["n_ps"], pl["n"]), dtype=np.float64) if want_ps else None
        fa = W = None
        if want_fullalpha and not skip:
            rows = info.r if pl["params"].n_alpha_rows < 0 else max(info.kmax_eff, min(info.r, pl["params"].n_alpha_rows))
            fa = np.empty((rows, pl["n"]), dtype=np.float64)
        k_eff = [0 if skip else min(k, info.kmax_eff) for k in pl["ks"]]     # kmax_eff = min(max(k), r), or min(max(k), nrow(fullalpha))
        rows_w = pl["m1"] if pl["params"].mode == MODE_FAST else pl["m1"] + pl["n_ps"]
        if want_W and not skip:
            W = np.empty(sum(rows_w * k for k in k_eff), dtype=np.float64)
        self._check(self._lib.ruviii_download(self._h, _ptr(newY), _ptr(newY_ps), _ptr(fa), _ptr(W)))
        Ws = None
        if W is not None:
            Ws, off = [], 0
            for k in k_eff:
                Ws.append(W[off:off + rows_w * k].reshape(rows_w, k))
                off += rows_w * k
        return dict(newY=newY, newY_ps=newY_ps, fullalpha=fa, W=Ws, info=info)

    # ---- one shot ---------------------------------------------------------------------------------------
    def normalise(self, params, Y, PS, group_of_row, ctl, ks, want_fullalpha=False, want_W=False, want_ps=False,
                  out=None, ps_on_device=0):
        """Host buffers in, host buffers out, through ``ruviii_normalise`` (what the R shim calls).
        ``ps_on_device`` = n_ps > 0 with PS = None: the PRPS rows come from the sets given to ``set_prps_sets``."""
        Y = _rowmajor(Y, "Y")
        if ps_on_device:
            m1, n = Y.shape
            group_of_row = np.ascontiguousarray(group_of_row, dtype=np.int32)
            ks_arr = np.ascontiguousarray(ks, dtype=np.int32)
            self.plan(params, m1, n, int(ps_on_device), group_of_row, np.asarray(ctl), ks_arr)
            self.upload(Y, None)
            self.run()
            return self.download(want_fullalpha=want_fullalpha, want_W=want_W, want_ps=want_ps, out=out)
        PS = _rowmajor(PS, "PS") if PS is not None and np.size(PS) else np.zeros((0, Y.shape[1]))
        m1, n = Y.shape
        n_ps = PS.shape[0]
        if n_ps and PS.shape[1] != n:
            raise RuviiiError(ERR_INVALID, "replicate.data has %d columns, the assay has %d genes" % (PS.shape[1], n))
        group_of_row = np.ascontiguousarray(group_of_row, dtype=np.int32)
        ctl_u8 = np.ascontiguousarray(np.asarray(ctl).astype(np.uint8))
        if ctl_u8.shape[0] != n:
            raise RuviiiError(ERR_INVALID, "len(ctl) != n")
        need = n_ps if params.mode == MODE_FAST else m1 + n_ps
        if group_of_row.shape[0] != need:
            raise RuviiiError(ERR_INVALID, "group_of_row has %d entries, expected %d" % (group_of_row.shape[0], need))
        ks_arr = np.ascontiguousarray(ks, dtype=np.int32)
        nk = ks_arr.shape[0]
        # output sizes depend on r, which the plan computes: plan first, then size the optional outputs
        self.plan(params, m1, n, n_ps, group_of_row, ctl_u8, ks_arr)
        self.upload(Y, PS)
        self.run()
        res = self.download(want_fullalpha=want_fullalpha, want_W=want_W, want_ps=want_ps, out=out)
        if params.average_rep:
            res["averaged"] = self.download_averaged()
        return res

    def normalise_oneshot(self, params, Y, PS, group_of_row, ctl, ks, out):
        """The literal one-call C entry point (newY only); used by bench.py's end-to-end leg."""
        Y = _rowmajor(Y, "Y")
        PS = _rowmajor(PS, "PS")
        group_of_row = np.ascontiguousarray(group_of_row, dtype=np.int32)
        ctl_u8 = np.ascontiguousarray(np.asarray(ctl).astype(np.uint8))
        ks_arr = np.ascontiguousarray(ks, dtype=np.int32)
        info = Info()
        self._check(self._lib.ruviii_normalise(
            self._h, C.byref(params), _ptr(Y), Y.strides[0] // 8, Y.shape[0], Y.shape[1], _ptr(PS),
            PS.strides[0] // 8, PS.shape[0], _ptr(group_of_row), _ptr(ctl_u8), _ptr(ks_arr), ks_arr.shape[0],
            _ptr(out), None, None, None, C.byref(info)))
        self._info = info
        return info

    def normalise_blocks(self, params, Y, PS, group_of_row, ctl, ks, outs):
        """ruviii_normalise_blocks: one separately allocated m1 x n output per k (what the R shim hands over)."""
        Y = _rowmajor(Y, "Y")
        PS = _rowmajor(PS, "PS")
        group_of_row = np.ascontiguousarray(group_of_row, dtype=np.int32)
        ctl_u8 = np.ascontiguousarray(np.asarray(ctl).astype(np.uint8))
        ks_arr = np.ascontiguousarray(ks, dtype=np.int32)
        ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
        info = Info()
        self._check(self._lib.ruviii_normalise_blocks(
            self._h, C.byref(params), _ptr(Y), Y.strides[0] // 8, Y.shape[0], Y.shape[1], _ptr(PS),
            PS.strides[0] // 8, PS.shape[0], _ptr(group_of_row), _ptr(ctl_u8), _ptr(ks_arr), ks_arr.shape[0],
            ptrs, None, None, None, C.byref(info)))
        self._info = info
        return info

    def fast_residop(self, A, group_of_row):
        A = _rowmajor(A, "A")
        g = np.ascontiguousarray(group_of_row, dtype=np.int32)
        out = np.empty_like(A)
        self._check(self._lib.ruviii_fast_residop(self._h, _ptr(A), A.shape[0], A.shape[1], _ptr(g), _ptr(out)))
        return out

    # ---- the generic primitives of RcppExports.R:12-26 (R matrices are column-major: numpy arrays go in Fortran order) ----
    @staticmethod
    def _colmajor(a, name):
        a = np.asarray(a, dtype=np.float64)
        if a.ndim != 2:
            raise TypeError("%s must be a matrix" % name)
        return np.asfortranarray(a)

    def matrixMult(self, A, B):
        A, B = self._colmajor(A, "A"), self._colmajor(B, "B")
        if A.shape[1] != B.shape[0]:
            raise RuviiiError(ERR_INVALID, "non-conformable arguments")
        out = np.empty((A.shape[0], B.shape[1]), dtype=np.float64, order="F")
        self._check(self._lib.ruviii_mat_mult(self._h, _ptr(A), A.shape[0], A.shape[1], _ptr(B), B.shape[1], _ptr(out)))
        return out

    def matSubtraction(self, A, B):
        A, B = self._colmajor(A, "A"), self._colmajor(B, "B")
        if A.shape != B.shape:
            raise RuviiiError(ERR_INVALID, "non-conformable arrays")
        out = np.empty(A.shape, dtype=np.float64, order="F")
        self._check(self._lib.ruviii_mat_subtract(self._h, _ptr(A), _ptr(B), A.size, _ptr(out)))
        return out

    def matTranspose(self, A):
        A = self._colmajor(A, "A")
        out = np.empty((A.shape[1], A.shape[0]), dtype=np.float64, order="F")
        self._check(self._lib.ruviii_mat_transpose(self._h, _ptr(A), A.shape[0], A.shape[1], _ptr(out)))
        return out

    def matInverse(self, A):
        A = self._colmajor(A, "A")
        if A.shape[0] != A.shape[1]:
            raise RuviiiError(ERR_INVALID, "'a' must be a square matrix")
        out = np.empty(A.shape, dtype=np.float64, order="F")
        self._check(self._lib.ruviii_mat_inverse(self._h, _ptr(A), A.shape[0], _ptr(out)))
        return out

    def probe(self, which, size=0):
        v = C.c_double()
        self._check(self._lib.ruviii_probe(self._h, which, size, C.byref(v)))
        return v.value
