"""ctypes wrapper around the CPU oracle (oracle/saige_oracle.cpp).

TEST INFRASTRUCTURE: import only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (saigegds_b200) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libsaige_oracle.so")


class OrcParams(C.Structure):
    _fields_ = [("tol", C.c_double), ("tolPCG", C.c_double), ("seed", C.c_int), ("maxiter", C.c_int),
                ("maxiterPCG", C.c_int), ("no_iteration", C.c_int), ("nrun", C.c_int), ("num_marker", C.c_int),
                ("traceCVcutoff", C.c_double), ("ratioCVcutoff", C.c_double), ("verbose", C.c_int)]


def default_params(**kw) -> OrcParams:
    """Defaults of seqFitNullGLMM_SPA (R/saige_main.r:223-229)."""
    d = dict(tol=0.02, tolPCG=1e-5, seed=200, maxiter=20, maxiterPCG=500, no_iteration=0, nrun=30,
             num_marker=30, traceCVcutoff=0.0025, ratioCVcutoff=0.001, verbose=0)
    d.update(kw)
    return OrcParams(**d)


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(HERE, f))
                                             for f in ("saige_oracle.cpp", "saige_score_oracle.cpp")):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_last_error.restype = C.c_char_p
        for f in ("orc_num_products", "orc_num_pcg", "orc_num_pcg_iter"):
            getattr(_lib, f).restype = C.c_long
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


FAMILY = {"binomial": 0, "gaussian": 1}


class Oracle:
    def __init__(self):
        self.h = C.c_void_p(lib().orc_create())
        self.n = self.m = 0

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    def store_2b_geno(self, packed: np.ndarray, n_samp: int, num_thread: int = 1, borrow: bool = False, want_diag: bool = True):
        """packed: [M][ceil(N/4)] uint8 (one row per variant == R RawMatrix column).  borrow: keep a pointer to `packed`
        instead of copying it (what the reference does with R's matrix); want_diag=False skips the serial diag(GRM) pass."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        m, nb = packed.shape
        self.n, self.m = int(n_samp), int(m)
        lut = np.empty(4 * m)
        diag = np.empty(self.n)
        self._borrowed = packed if borrow else None
        lib().orc_store_2b_geno_ex(self.h, _p(packed, C.c_ubyte), C.c_long(self.n), C.c_long(nb), C.c_long(m),
                                   C.c_int(num_thread), _p(lut), _p(diag), C.c_int(int(borrow)), C.c_int(int(not want_diag)))
        return lut.reshape(m, 4), diag

    def store_sp_geno(self, sp_geno_list, n_samp: int, num_thread: int = 1):
        """saige_store_sp_geno: sp_geno_list = list of int32 vectors (n1, n2, n3, indices) as made by get_sparse."""
        m = len(sp_geno_list)
        self.n, self.m = int(n_samp), int(m)
        offsets = np.zeros(m + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(v) for v in sp_geno_list])
        data = np.ascontiguousarray(np.concatenate(sp_geno_list), dtype=np.int32)
        lut = np.empty(4 * m)
        diag = np.empty(self.n)
        lib().orc_store_sp_geno(self.h, _p(data, C.c_int), _p(offsets, C.c_long), C.c_long(self.n), C.c_long(m),
                                C.c_int(num_thread), _p(lut), _p(diag))
        return lut.reshape(m, 4), diag

    def allele_counts(self):
        nv = np.empty(self.m, dtype=np.int32)
        sm = np.empty(self.m, dtype=np.int32)
        lib().orc_allele_counts(self.h, _p(nv, C.c_int), _p(sm, C.c_int))
        return nv, sm

    def get_geno_ds(self, snp_idx: int):
        ds = np.empty(self.n)
        lib().orc_get_geno_ds(self.h, C.c_long(snp_idx), _p(ds))
        return ds

    def grm_mv(self, b):
        b = _f64(b)
        out = np.empty(self.n)
        lib().orc_grm_mv(self.h, _p(b), _p(out))
        return out

    def diag_sigma(self, w, tau):
        w, tau = _f64(w), _f64(tau)
        out = np.empty(self.n)
        lib().orc_diag_sigma(self.h, _p(w), _p(tau), _p(out))
        return out

    def pcg(self, w, tau, b, maxiterPCG=500, tolPCG=1e-5):
        w, tau, b = _f64(w), _f64(tau), _f64(b)
        x = np.empty(self.n)
        it = C.c_int(0)
        lib().orc_pcg(self.h, _p(w), _p(tau), _p(b), C.c_int(maxiterPCG), C.c_double(tolPCG), _p(x), C.byref(it))
        return x, it.value

    # ---- R RNG ----
    def set_seed(self, seed):
        lib().orc_set_seed(self.h, C.c_uint(seed))

    def unif_rand(self, n):
        out = np.empty(n)
        lib().orc_unif_rand(self.h, C.c_long(n), _p(out))
        return out

    def rademacher(self, n):
        out = np.empty(n)
        lib().orc_rademacher(self.h, C.c_long(n), _p(out))
        return out

    def sample_int(self, n):
        out = np.empty(n, dtype=np.int32)
        lib().orc_sample_int(self.h, C.c_int(n), _p(out, C.c_int))
        return out

    # ---- drivers ----
    def fit_AI_PCG(self, trait, fit0, X, tau, params=None):
        """saige_fit_AI_PCG_binary / _quant.  fit0: saigegds_b200.rsetup.Fit0-like."""
        params = params or default_params()
        X = np.asfortranarray(X, dtype=np.float64)
        n, p = X.shape
        y = _f64(fit0.y)
        off = None if fit0.offset is None else _f64(fit0.offset)
        eta, mu, coef, tau = _f64(fit0.linear_predictors), _f64(fit0.fitted_values), _f64(fit0.coefficients), _f64(tau)
        o = dict(coefficients=np.empty(p), tau=np.empty(2), linear_predictors=np.empty(n), fitted_values=np.empty(n),
                 residuals=np.empty(n), cov=np.empty((p, p), order="F"))
        conv = C.c_int(0)
        rc = lib().orc_fit_AI_PCG(self.h, C.c_int(1 if trait == "quantitative" else 0), C.c_int(FAMILY[fit0.family]),
                                  C.c_long(n), C.c_int(p), _p(y), _p(off) if off is not None else None, _p(eta), _p(mu),
                                  _p(coef), _p(X), _p(tau), C.byref(params), _p(o["coefficients"]), _p(o["tau"]),
                                  _p(o["linear_predictors"]), _p(o["fitted_values"]), _p(o["residuals"]), _p(o["cov"]),
                                  C.byref(conv))
        if rc != 0:
            raise RuntimeError(lib().orc_last_error(self.h).decode())
        o["converged"] = bool(conv.value)
        return o

    def calc_var_ratio(self, trait, fit0, tau, noK, marker_list, params=None, cap=4096):
        params = params or default_params()
        n = len(fit0.y)
        X1 = np.asfortranarray(noK.X1, dtype=np.float64)
        XV = np.asfortranarray(noK.XV, dtype=np.float64)
        XXVX_inv = np.asfortranarray(noK.XXVX_inv, dtype=np.float64)
        p = X1.shape[1]
        eta, mu, tau = _f64(fit0.linear_predictors), _f64(fit0.fitted_values), _f64(tau)
        ml = np.ascontiguousarray(marker_list, dtype=np.int32)
        oid = np.empty(cap, dtype=np.int32)
        cols = [np.empty(cap) for _ in range(5)]
        n_out = C.c_int(0)
        rc = lib().orc_calc_var_ratio(self.h, C.c_int(1 if trait == "quantitative" else 0), C.c_int(FAMILY[fit0.family]),
                                      C.c_long(n), C.c_int(p), _p(eta), _p(mu), _p(tau), _p(X1), _p(XV), _p(XXVX_inv),
                                      C.byref(params), _p(ml, C.c_int), C.c_long(len(ml)), C.c_int(cap),
                                      _p(oid, C.c_int), *[_p(c) for c in cols], C.byref(n_out))
        if rc != 0:
            raise RuntimeError(lib().orc_last_error(self.h).decode())
        k = n_out.value
        return dict(id=oid[:k].copy(), maf=cols[0][:k].copy(), mac=cols[1][:k].copy(), var1=cols[2][:k].copy(),
                    var2=cols[3][:k].copy(), ratio=cols[4][:k].copy())

    def GxG_snp_bin(self, fit0, tau, inter_term, noK, params=None):
        """saige_GxG_snp_bin (src/saige_fitnull.cpp:1480-1558) on the stored genotypes."""
        params = params or default_params()
        n = len(fit0.y)
        X1 = np.asfortranarray(noK.X1, dtype=np.float64)
        XV = np.asfortranarray(noK.XV, dtype=np.float64)
        XXVX_inv = np.asfortranarray(noK.XXVX_inv, dtype=np.float64)
        p = X1.shape[1]
        y, eta, mu, tau, g = _f64(fit0.y), _f64(fit0.linear_predictors), _f64(fit0.fitted_values), _f64(tau), _f64(inter_term)
        out = np.empty(7)
        rc = lib().orc_GxG_snp_bin(self.h, C.c_int(FAMILY[fit0.family]), C.c_long(n), C.c_int(p), _p(y), _p(eta), _p(mu), _p(tau),
                                   _p(g), _p(X1), _p(XV), _p(XXVX_inv), C.byref(params), _p(out))
        if rc != 0:
            raise RuntimeError(lib().orc_last_error(self.h).decode())
        return dict(beta=out[0], SE=out[1], n_nonzero=int(out[2]), pval=out[3], p_norm=out[4], converged=bool(out[5]),
                    tau_G=out[6])

    @property
    def num_products(self):
        return lib().orc_num_products(self.h)


def init_nullmod(trait, y, mu, noK_X1, noK_XV, noK_XXVX_inv, noK_V, tau):
    """The derived arrays of .init_nullmod (R/assoc_single.r:17-67), all samples (ii = 1..n).  Matrices are K x n in
    R's column-major order, i.e. C-contiguous [n][K] here."""
    y, mu = _f64(y), _f64(mu)
    X1 = _f64(noK_X1)                                   # n x K
    XXVX_inv = _f64(noK_XXVX_inv)                       # n x K
    V = _f64(noK_V)
    m = dict(trait=trait, tau=_f64(tau), y=y, mu=mu, y_mu=y - mu, mu2=mu * (1 - mu),
             t_XXVX_inv=np.ascontiguousarray(XXVX_inv),                       # [n][K] == K x n column-major
             XV=np.ascontiguousarray(_f64(noK_XV).T),                         # noK_XV is K x n -> [n][K]
             t_XVX_inv_XV=np.ascontiguousarray(XXVX_inv * V[:, None]),
             t_X=np.ascontiguousarray(X1))
    if trait == "binary":
        m["XVX"] = np.ascontiguousarray(X1.T @ (X1 * m["mu2"][:, None]))
    else:
        m["XVX"] = np.ascontiguousarray(X1.T @ X1)
    m["S_a"] = _f64((X1 * m["y_mu"][:, None]).sum(axis=0))
    return m


def score_test(model, dosage, var_ratio, maf=float("nan"), mac=10.0, missing=0.1, spa_pval=0.05):
    """seqAssocGLMM_SPA's per-variant test (src/saige_main.cpp:188-407) on dosage [n_var][n] (NaN = missing).
    Returns dict of arrays (AF, mac, num, beta, SE, pval, p_norm, converged) and the `valid` mask."""
    ds = np.array(dosage, dtype=np.float64, order="C", copy=True)
    n_var, n = ds.shape
    K = model["t_X"].shape[1]
    out = np.empty((n_var, 8))
    valid = np.empty(n_var, dtype=np.int32)
    lib().orc_score_test(C.c_int(0 if model["trait"] == "binary" else 1), C.c_long(n), C.c_int(K), _p(model["tau"]),
                         _p(model["y"]), _p(model["mu"]), _p(model["y_mu"]), _p(model["mu2"]), _p(model["t_XXVX_inv"]),
                         _p(model["XV"]), _p(model["t_XVX_inv_XV"]), _p(model["XVX"]), _p(model["t_X"]), _p(model["S_a"]),
                         C.c_double(var_ratio), C.c_double(maf), C.c_double(mac), C.c_double(missing), C.c_double(spa_pval),
                         C.c_long(n_var), _p(ds), _p(out), _p(valid, C.c_int))
    names = ["AF", "mac", "num", "beta", "SE", "pval", "p_norm", "converged"]
    res = {k: out[:, i].copy() for i, k in enumerate(names)}
    res["valid"] = valid.astype(bool)
    return res


def get_sparse(geno, n_samp=None):
    """saige_get_sparse (src/saige_fitnull.cpp:252-320) on one variant: uint8 codes, int32 genotypes or float64 dosages."""
    geno = np.ascontiguousarray(geno)
    n = len(geno) if n_samp is None else int(n_samp)
    if n > len(geno):
        raise ValueError("No enough genotypes.")
    if geno.dtype == np.uint8:
        t = 0
    elif geno.dtype == np.int32:
        t = 1
    elif geno.dtype == np.float64:
        t = 2
    else:
        raise ValueError("Invalid data type.")
    out = np.empty(n + 4, dtype=np.int32)
    lib().orc_get_sparse.restype = C.c_long
    k = lib().orc_get_sparse(geno.ctypes.data_as(C.c_void_p), C.c_int(t), C.c_long(n), _p(out, C.c_int))
    return out[:k].copy()


def qnorm(p):
    lib().orc_qnorm.restype = C.c_double
    return lib().orc_qnorm(C.c_double(p))


def max_threads():
    return lib().orc_max_threads()


def synth_geno(n_samp: int, n_var: int, var_offset: int = 0, seed: int = 200, missing_rate: float = 0.005,
               num_thread: int | None = None) -> np.ndarray:
    """The synthetic genotypes of SURVEY.md 8(d) on the CPU: bit-identical to the device generator
    (Context.store_synthetic / synth_to_host), so both arms of bench.py consume the same bytes."""
    nb = (n_samp + 3) // 4
    out = np.empty((n_var, nb), dtype=np.uint8)
    lib().orc_synth_geno(C.c_long(n_samp), C.c_long(n_var), C.c_long(var_offset), C.c_ulonglong(seed), C.c_double(missing_rate),
                         _p(out, C.c_ubyte), C.c_int(num_thread or max_threads()))
    return out
