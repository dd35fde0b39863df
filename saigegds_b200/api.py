"""Host-side mirror of the reference interface for the null-model hot path.

Names, argument meaning and error behaviour follow the reference:

  native entry points (src/saige_fitnull.cpp)         here
  -------------------------------------------         ---------------------------------
  saige_store_2b_geno        (:159)                   Context.saige_store_2b_geno
  saige_init_sparse / saige_get_sparse (:244, :252)   saige_get_sparse
  saige_store_sp_geno        (:324)                   Context.saige_store_sp_geno
  get_crossprod_b_grm        (:436)                   Context.get_crossprod_b_grm
  get_diag_sigma / PCG_diag_sigma (:543, :582)        Context.get_diag_sigma / PCG_diag_sigma
  saige_fit_AI_PCG_binary / _quant (:949, :1103)      Context.saige_fit_AI_PCG_binary / _quant
  saige_calc_var_ratio_binary / _quant (:1255, :1366) Context.saige_calc_var_ratio_binary / _quant
  saige_GxG_snp_bin          (:1480)                  Context.saige_GxG_snp_bin
  R driver seqFitNullGLMM_SPA (R/saige_main.r:223)    seqFitNullGLMM_SPA

Everything numeric runs in libsaigegds_b200.so on the GPU; this module only marshals numpy arrays
into the C-ABI.  torch is used for plumbing only (device selection, torch.distributed rendezvous).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from . import rsetup


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a, order="C"):
    return np.require(a, dtype=np.float64, requirements=["F" if order == "F" else "C", "A"])


def make_param(tol=0.02, tolPCG=1e-5, seed=200, maxiter=20, maxiterPCG=500, no_iteration=False, nrun=30,
               num_marker=30, traceCVcutoff=0.0025, ratioCVcutoff=0.001, verbose=False, indent="", **_ignored) -> L.Param:
    """The `param` list of R/saige_main.r:442-453 (num.thread has no meaning on the GPU and is ignored)."""
    return L.Param(tol, tolPCG, int(seed), int(maxiter), int(maxiterPCG), int(bool(no_iteration)), int(nrun),
                   int(num_marker), traceCVcutoff, ratioCVcutoff, int(bool(verbose)), indent.encode())


_GENO_TYPE = {np.dtype(np.uint8): 0, np.dtype(np.int32): 1, np.dtype(np.float64): 2}


def saige_get_sparse(geno: np.ndarray, num_samp: int | None = None) -> np.ndarray:
    """saige_get_sparse (src/saige_fitnull.cpp:252-320; saige_init_sparse's buffer is allocated here): one variant's
    genotypes -- uint8 codes, int32 genotypes or float64 dosages, anything outside 0/1/2 = missing -- as the int32 vector
    (n1, n2, n3, indices of 1s, of 2s, of missing), counted on the minor allele.  Host-only library call."""
    g = np.ascontiguousarray(geno)
    if g.dtype not in _GENO_TYPE:
        raise L.InvalidArgument(L.SGB_ERR_INVALID, "Invalid data type.")
    n = len(g) if num_samp is None else int(num_samp)
    if n > len(g):
        raise L.InvalidArgument(L.SGB_ERR_INVALID, "No enough genotypes.")
    out = np.empty(n + 3, dtype=np.int32)
    k = C.c_int64(0)
    L.check(L.lib().sgb_get_sparse(g.ctypes.data_as(C.c_void_p), C.c_int(_GENO_TYPE[g.dtype]), C.c_int64(n),
                                   _p(out, C.c_int32), C.byref(k)))
    return out[:k.value].copy()


def _flatten_sparse(sp_geno_list):
    offsets = np.zeros(len(sp_geno_list) + 1, dtype=np.int64)
    if len(sp_geno_list):
        offsets[1:] = np.cumsum([len(v) for v in sp_geno_list])
        data = np.ascontiguousarray(np.concatenate(sp_geno_list), dtype=np.int32)
    else:
        data = np.zeros(1, dtype=np.int32)
    return data, offsets


def sparse_to_packed(sp_geno_list, num_samp: int) -> np.ndarray:
    """The host packing step of saige_store_sp_geno on its own: list of sparse vectors -> uint8 [M][ceil(N/4)]."""
    data, offsets = _flatten_sparse(sp_geno_list)
    out = np.empty((len(sp_geno_list), (int(num_samp) + 3) // 4), dtype=np.uint8)
    L.check(L.lib().sgb_sparse_to_packed(_p(data, C.c_int32), _p(offsets, C.c_int64), C.c_int64(num_samp),
                                         C.c_int64(len(sp_geno_list)), _p(out, C.c_ubyte)))
    return out


class Context:
    """One GPU, one shard of variants: the state kept in file-scope statics at saige_fitnull.cpp:122-131."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        L.check(L.lib().sgb_ctx_create(C.byref(self._h), C.c_int(device)))
        self.device = device
        self.n_samp = self.n_var = self.n_var_total = 0
        self.var_offset = 0
        self.rank, self.world = 0, 1

    def close(self):
        if getattr(self, "_h", None):
            for ptr in getattr(self, "_pinned", []):
                L.lib().sgb_free_host(self._h, ptr)
            self._pinned = []
            L.lib().sgb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration ----
    def set_print(self, fn):
        """Route the library's verbose output (Rprintf in the reference) to `fn(str)`; None restores stdout."""
        if fn is None:
            self._print_cb = None
            L.check(L.lib().sgb_set_callbacks(self._h, None, None, None))
            return
        self._print_cb = C.CFUNCTYPE(None, C.c_char_p)(lambda s: fn(s.decode("utf-8", "replace")))
        L.check(L.lib().sgb_set_callbacks(self._h, self._print_cb, None, None))

    def set_kernel(self, name: str):
        L.check(L.lib().sgb_set_kernel(self._h, C.c_int(L.KERNEL[name])))

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        L.check(L.lib().sgb_comm_init(self._h, buf, C.c_int(rank), C.c_int(world)))
        self.rank, self.world = rank, world

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_ubyte * 128)()
        L.check(L.lib().sgb_comm_unique_id(buf))
        return bytes(buf)

    # ---- saige_store_2b_geno ----
    def saige_store_2b_geno(self, rawgeno: np.ndarray, num_samp: int, n_variant_total: int | None = None,
                            variant_offset: int = 0):
        """rawgeno: uint8 [n_variant_local][ceil(num_samp/4)] -- the transpose of R's RawMatrix, i.e. the same
        bytes in memory.  Returns (buf_std_geno [M][4], buf_diag_grm [N]) as the reference fills them."""
        g = np.require(rawgeno, dtype=np.uint8, requirements=["C", "A"])
        if g.ndim != 2:
            raise L.InvalidArgument(L.SGB_ERR_INVALID, "rawgeno must be a 2-d uint8 array")
        m, nb = g.shape
        total = m if n_variant_total is None else int(n_variant_total)
        lut = np.empty((m, 4))
        diag = np.empty(int(num_samp))
        L.check(L.lib().sgb_store_2b_geno(self._h, _p(g, C.c_ubyte), C.c_int64(num_samp), C.c_int64(nb), C.c_int64(m),
                                          C.c_int64(total), C.c_int64(variant_offset), _p(lut), _p(diag)))
        self.n_samp, self.n_var, self.n_var_total, self.var_offset = int(num_samp), m, total, variant_offset
        return lut, diag

    def saige_store_sp_geno(self, sp_geno_list, num_samp: int, n_variant_total: int | None = None,
                            variant_offset: int = 0):
        """saige_store_sp_geno (:324-388): sp_geno_list = this rank's list of `saige_get_sparse` vectors.  Returns
        (buf_std_geno [M][4] in the reference's sparse layout, buf_diag_grm [N])."""
        m = len(sp_geno_list)
        data, offsets = _flatten_sparse(sp_geno_list)
        total = m if n_variant_total is None else int(n_variant_total)
        lut = np.empty((m, 4))
        diag = np.empty(int(num_samp))
        L.check(L.lib().sgb_store_sp_geno(self._h, _p(data, C.c_int32), _p(offsets, C.c_int64), C.c_int64(num_samp),
                                          C.c_int64(m), C.c_int64(total), C.c_int64(variant_offset), _p(lut), _p(diag)))
        self.n_samp, self.n_var, self.n_var_total, self.var_offset = int(num_samp), m, total, variant_offset
        return lut, diag

    def store_gds_geno(self, allele_bits: np.ndarray, n_samp_file: int, n_variant_file: int, sample_sel=None,
                       maf=float("nan"), missing_rate=float("nan")):
        """Genotypes straight from the bytes of a GDS `genotype/data` node (bit2 allele pairs): 2-bit packing, allele
        counts and the MAF / missing-rate filter of R/saige_main.r:319 run on the device (replaces
        SeqArray:::.seqGet2bGeno, :420).  Returns dict(variant_sel, n_valid_alleles, n_alt_alleles, lut, diag)."""
        bits = np.require(allele_bits, dtype=np.uint8, requirements=["C", "A"]).reshape(-1)
        if bits.size < (int(n_samp_file) * int(n_variant_file) + 1) // 2:
            raise L.InvalidArgument(L.SGB_ERR_INVALID, "allele_bits is shorter than n_samp_file * n_variant_file nibbles")
        sel = None if sample_sel is None else np.require(sample_sel, dtype=np.int32, requirements=["C"])
        n = int(n_samp_file) if sel is None else len(sel)
        keep = np.zeros(n_variant_file, dtype=np.int32)
        nv = np.zeros(n_variant_file, dtype=np.int32)
        na = np.zeros(n_variant_file, dtype=np.int32)
        lut = np.empty((n_variant_file, 4))
        diag = np.empty(n)
        m = C.c_int64(0)
        L.check(L.lib().sgb_store_gds_geno(self._h, _p(bits, C.c_ubyte), C.c_int64(n_samp_file), C.c_int64(n_variant_file),
                                           None if sel is None else _p(sel, C.c_int32), C.c_int64(n), C.c_double(maf),
                                           C.c_double(missing_rate), _p(keep, C.c_int32), C.byref(m), _p(nv, C.c_int32),
                                           _p(na, C.c_int32), _p(lut), _p(diag)))
        self.n_samp, self.n_var, self.n_var_total, self.var_offset = n, int(m.value), int(m.value), 0
        return dict(variant_sel=keep.astype(bool), n_valid_alleles=nv, n_alt_alleles=na, lut=lut[:m.value].copy(), diag=diag)

    def store_synthetic(self, n_samp: int, n_variant_local: int, n_variant_total: int | None = None,
                        variant_offset: int = 0, seed: int = 200, missing_rate: float = 0.005, want_outputs=False):
        """Generate the SURVEY section 8(d) synthetic genotypes directly in HBM and store them."""
        total = n_variant_local if n_variant_total is None else int(n_variant_total)
        ptr = C.c_void_p()
        L.check(L.lib().sgb_synth_geno_device(self._h, C.c_int64(n_samp), C.c_int64(n_variant_local),
                                              C.c_int64(variant_offset), C.c_uint64(seed), C.c_double(missing_rate),
                                              C.byref(ptr)))
        nb = (n_samp + 3) // 4
        lut = np.empty((n_variant_local, 4)) if want_outputs else None
        diag = np.empty(n_samp) if want_outputs else None
        L.check(L.lib().sgb_store_2b_geno_device(self._h, ptr, C.c_int(1), C.c_int64(n_samp), C.c_int64(nb),
                                                 C.c_int64(n_variant_local), C.c_int64(total), C.c_int64(variant_offset),
                                                 _p(lut) if want_outputs else None, _p(diag) if want_outputs else None))
        self.n_samp, self.n_var, self.n_var_total, self.var_offset = n_samp, n_variant_local, total, variant_offset
        return lut, diag

    def synth_to_host(self, n_samp: int, n_variant_local: int, variant_offset: int = 0, seed: int = 200,
                      missing_rate: float = 0.005) -> np.ndarray:
        """The same synthetic bytes as store_synthetic, copied to the host (so the oracle can consume them)."""
        ptr = C.c_void_p()
        L.check(L.lib().sgb_synth_geno_device(self._h, C.c_int64(n_samp), C.c_int64(n_variant_local),
                                              C.c_int64(variant_offset), C.c_uint64(seed), C.c_double(missing_rate),
                                              C.byref(ptr)))
        nb = (n_samp + 3) // 4
        out = np.empty((n_variant_local, nb), dtype=np.uint8)
        L.check(L.lib().sgb_copy_from_device(self._h, _p(out, C.c_ubyte), ptr, C.c_int64(out.nbytes)))
        L.check(L.lib().sgb_free_device(self._h, ptr))
        return out

    def allele_counts(self):
        nv = np.empty(self.n_var, dtype=np.int32)
        sm = np.empty(self.n_var, dtype=np.int32)
        L.check(L.lib().sgb_allele_counts(self._h, _p(nv, C.c_int32), _p(sm, C.c_int32)))
        return nv, sm

    def get_geno_ds(self, snp_idx: int) -> np.ndarray:
        ds = np.empty(self.n_samp)
        L.check(L.lib().sgb_get_geno_ds(self._h, C.c_int64(snp_idx), _p(ds)))
        return ds

    # ---- products and solves ----
    def get_crossprod_b_grm(self, b: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """(1/M) G G' b for one vector [N] or k vectors [N, k] (host in, host out).  `out` (same shape, float64, Fortran
        order) is filled in place when given -- with `pinned_empty` buffers for both, the copies run at PCIe speed."""
        b = np.asarray(b, dtype=np.float64)
        one = b.ndim == 1
        bb = _f64(b.reshape(self.n_samp, -1), "F")
        if out is not None:
            oo = out.reshape(self.n_samp, -1)
            if oo.dtype != np.float64 or not oo.flags.f_contiguous or oo.shape != bb.shape:
                raise L.InvalidArgument(L.SGB_ERR_INVALID, "out must be float64, Fortran-contiguous and shaped like b")
            L.check(L.lib().sgb_grm_mv(self._h, _p(bb), _p(oo), C.c_int(bb.shape[1])))
            return out
        oo = np.empty_like(bb, order="F")
        L.check(L.lib().sgb_grm_mv(self._h, _p(bb), _p(oo), C.c_int(bb.shape[1])))
        return oo[:, 0].copy() if one else oo

    def pinned_empty(self, shape) -> np.ndarray:
        """A float64 Fortran-order array in page-locked host memory (sgb_malloc_host); freed with the context."""
        n = int(np.prod(shape))
        ptr = C.c_void_p()
        L.check(L.lib().sgb_malloc_host(self._h, C.c_int64(8 * max(n, 1)), C.byref(ptr)))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (C.c_double * max(n, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape, order="F")

    def get_diag_sigma(self, w, tau) -> np.ndarray:
        w, tau = _f64(w), _f64(tau)
        out = np.empty(self.n_samp)
        L.check(L.lib().sgb_diag_sigma(self._h, _p(w), _p(tau), _p(out)))
        return out

    def PCG_diag_sigma(self, w, tau, b, maxiterPCG=500, tolPCG=1e-5):
        """Returns (x, iterations); b may be [N] or [N, k] (k independent solves in lock-step)."""
        w, tau = _f64(w), _f64(tau)
        b = np.asarray(b, dtype=np.float64)
        one = b.ndim == 1
        bb = _f64(b.reshape(self.n_samp, -1), "F")
        k = bb.shape[1]
        x = np.empty_like(bb, order="F")
        it = np.zeros(k, dtype=np.int32)
        L.check(L.lib().sgb_pcg(self._h, _p(w), _p(tau), _p(bb), C.c_int(k), C.c_int(maxiterPCG), C.c_double(tolPCG),
                                _p(x), _p(it, C.c_int)))
        return (x[:, 0].copy(), int(it[0])) if one else (x, it)

    # ---- fits ----
    def _fit(self, fn, fit0: rsetup.Fit0, X, tau, param):
        X = _f64(X, "F")
        n, p = X.shape
        y, eta, mu, coef = _f64(fit0.y), _f64(fit0.linear_predictors), _f64(fit0.fitted_values), _f64(fit0.coefficients)
        off = None if fit0.offset is None else _f64(fit0.offset)
        f = L.Fit0(n, p, _p(y), _p(off) if off is not None else None, _p(eta), _p(mu), _p(coef), L.FAMILY[fit0.family])
        o = dict(coefficients=np.empty(p), linear_predictors=np.empty(n), fitted_values=np.empty(n), residuals=np.empty(n),
                 cov=np.empty((p, p), order="F"))
        g = L.Glmm(_p(o["coefficients"]), (C.c_double * 2)(), _p(o["linear_predictors"]), _p(o["fitted_values"]),
                   _p(o["residuals"]), _p(o["cov"]), 0)
        tau = _f64(tau)
        L.check(fn(self._h, C.byref(f), _p(X), _p(tau), C.byref(param), C.byref(g)))
        o["tau"] = np.array([g.tau[0], g.tau[1]])
        o["converged"] = bool(g.converged)
        return o

    def saige_fit_AI_PCG_binary(self, fit0, X, tau, param=None):
        return self._fit(L.lib().sgb_fit_AI_PCG_binary, fit0, X, tau, param or make_param())

    def saige_fit_AI_PCG_quant(self, fit0, X, tau, param=None):
        return self._fit(L.lib().sgb_fit_AI_PCG_quant, fit0, X, tau, param or make_param())

    def _var_ratio(self, fn, fit0, glmm_tau, noK: rsetup.ObjNoK, param, marker_list, cap=4096):
        n = len(fit0.y)
        X1, XV, XXVX_inv = _f64(noK.X1, "F"), _f64(noK.XV, "F"), _f64(noK.XXVX_inv, "F")
        p = X1.shape[1]
        y, eta, mu, coef = _f64(fit0.y), _f64(fit0.linear_predictors), _f64(fit0.fitted_values), _f64(fit0.coefficients)
        f = L.Fit0(n, p, _p(y), None, _p(eta), _p(mu), _p(coef), L.FAMILY[fit0.family])
        nk = L.NoK(p, _p(X1), _p(XV), _p(XXVX_inv))
        ml = np.require(marker_list, dtype=np.int32, requirements=["C"])
        oid = np.empty(cap, dtype=np.int32)
        cols = [np.empty(cap) for _ in range(5)]
        vr = L.VarRatio(cap, 0, _p(oid, C.c_int), *[_p(c) for c in cols])
        tau = _f64(glmm_tau)
        L.check(fn(self._h, C.byref(f), _p(tau), C.byref(nk), C.byref(param), _p(ml, C.c_int32), C.c_int64(len(ml)),
                   C.byref(vr)))
        k = vr.n
        return dict(id=oid[:k].copy(), maf=cols[0][:k].copy(), mac=cols[1][:k].copy(), var1=cols[2][:k].copy(),
                    var2=cols[3][:k].copy(), ratio=cols[4][:k].copy())

    def saige_calc_var_ratio_binary(self, fit0, glmm, noK, param, marker_list):
        return self._var_ratio(L.lib().sgb_calc_var_ratio_binary, fit0, glmm["tau"], noK, param or make_param(), marker_list)

    def saige_calc_var_ratio_quant(self, fit0, glmm, noK, param, marker_list):
        return self._var_ratio(L.lib().sgb_calc_var_ratio_quant, fit0, glmm["tau"], noK, param or make_param(), marker_list)

    def saige_GxG_snp_bin(self, fit0, glmm, inter_term, noK: rsetup.ObjNoK, param=None, verbose=False) -> dict:
        """saige_GxG_snp_bin (src/saige_fitnull.cpp:1480-1558): score test + full saddle-point approximation of one
        interaction term under the fitted mixed model.  Returns the columns of the reference's one-row data.frame."""
        n = len(fit0.y)
        X1, XV, XXVX_inv = _f64(noK.X1, "F"), _f64(noK.XV, "F"), _f64(noK.XXVX_inv, "F")
        p = X1.shape[1]
        y, eta, mu, coef = _f64(fit0.y), _f64(fit0.linear_predictors), _f64(fit0.fitted_values), _f64(fit0.coefficients)
        f = L.Fit0(n, p, _p(y), None, _p(eta), _p(mu), _p(coef), L.FAMILY[fit0.family])
        nk = L.NoK(p, _p(X1), _p(XV), _p(XXVX_inv))
        g = _f64(inter_term)
        if g.shape != (n,):
            raise L.InvalidArgument(L.SGB_ERR_INVALID, "inter_term must have one value per sample")
        tau = _f64(glmm["tau"])
        out = L.GxG()
        L.check(L.lib().sgb_GxG_snp_bin(self._h, C.byref(f), _p(tau), _p(g), C.byref(nk), C.byref(param or make_param()),
                                        C.c_int(int(bool(verbose))), C.byref(out)))
        return {"beta": out.beta, "SE": out.SE, "n_nonzero": int(out.n_nonzero), "pval": out.pval, "p.norm": out.p_norm,
                "converged": bool(out.converged), "tau_G": out.tau_G}

    # ---- R RNG ----
    def set_seed(self, seed: int):
        L.check(L.lib().sgb_r_set_seed(self._h, C.c_uint32(seed)))

    def runif(self, n: int) -> np.ndarray:
        out = np.empty(n)
        L.check(L.lib().sgb_r_unif_rand(self._h, C.c_int64(n), _p(out)))
        return out

    def sample_int(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int32)
        L.check(L.lib().sgb_r_sample_int(self._h, C.c_int32(n), _p(out, C.c_int32)))
        return out

    # ---- instrumentation / device buffers (bench.py) ----
    def stats(self) -> dict:
        s = L.Stats()
        L.check(L.lib().sgb_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def reset_stats(self):
        L.check(L.lib().sgb_reset_stats(self._h))

    def set_profiling(self, on: bool):
        L.check(L.lib().sgb_set_profiling(self._h, C.c_int(int(on))))

    def kernel_times(self) -> dict:
        """{kernel name: (total ms, launches)} accumulated since set_profiling(True)."""
        buf = C.create_string_buffer(1 << 16)
        L.check(L.lib().sgb_kernel_times(self._h, buf, C.c_int64(len(buf))))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, cnt = line.rsplit(" ", 2)
            out[name] = (float(ms), int(cnt))
        return out

    def device_vector(self, host: np.ndarray) -> "DeviceArray":
        host = _f64(host, "F")
        d = DeviceArray(self, host.nbytes)
        L.check(L.lib().sgb_copy_to_device(self._h, d.ptr, _p(host), C.c_int64(host.nbytes)))
        return d

    def device_empty(self, nbytes: int) -> "DeviceArray":
        return DeviceArray(self, nbytes)

    def time_products_device(self, b: "DeviceArray", out: "DeviceArray", k: int, reps: int) -> float:
        ms = C.c_float(0)
        L.check(L.lib().sgb_time_products_device(self._h, b.ptr, out.ptr, C.c_int(k), C.c_int(reps), C.byref(ms)))
        return float(ms.value)

    def grm_mv_device(self, b: "DeviceArray", out: "DeviceArray", k: int = 1):
        L.check(L.lib().sgb_grm_mv_device(self._h, b.ptr, out.ptr, C.c_int(k)))


class DeviceArray:
    def __init__(self, ctx: Context, nbytes: int):
        self.ctx, self.nbytes = ctx, nbytes
        self.ptr = C.c_void_p()
        L.check(L.lib().sgb_malloc_device(ctx._h, C.c_int64(nbytes), C.byref(self.ptr)))

    def to_host(self, shape, dtype=np.float64, order="F") -> np.ndarray:
        out = np.empty(shape, dtype=dtype, order=order)
        L.check(L.lib().sgb_copy_from_device(self.ctx._h, out.ctypes.data_as(C.c_void_p), self.ptr, C.c_int64(out.nbytes)))
        return out

    def free(self):
        if self.ptr:
            L.lib().sgb_free_device(self.ctx._h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


_default_ctx: Context | None = None


def default_context() -> Context:
    """Process-global context: the analogue of the reference's file-scope state (not re-entrant either)."""
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


@dataclass
class NullModel:
    """The `ClassSAIGE_NullModel` list assembled at R/saige_main.r:516-628."""
    coefficients: np.ndarray
    tau: np.ndarray
    linear_predictors: np.ndarray
    fitted_values: np.ndarray
    residuals: np.ndarray
    cov: np.ndarray
    converged: bool
    obj_noK: rsetup.ObjNoK
    var_ratio: dict
    trait_type: str
    sample_id: np.ndarray | None = None
    variant_id: np.ndarray | None = None


def _gds_front_end(ctx, path, data: dict, used: list, sample_col: str, maf, missing_rate, max_num_snp, variant_id, seed, verbose, threads):
    """The part of seqFitNullGLMM_SPA that talks to the GDS file (R/saige_main.r:273-333), without SeqArray: rows of `data` with a
    missing value are dropped, the rest are matched to the file's samples and put in file order, the variants are filtered on the
    selected samples (maf / missing.rate, or an explicit variant.id list), at most max.num.snp of them are kept (R's
    `set.seed(seed); sample(which(v), max.num.snp)`), and the genotypes are stored in `ctx`.  Returns (data in file order, variant ids)."""
    from . import gds as G
    if sample_col in used:
        raise ValueError("'%s' should not be in the formula." % sample_col)
    if sample_col not in data:
        raise ValueError("'%s' should be one of the columns in 'data'." % sample_col)
    ids = np.asarray(data[sample_col])
    if len(set(ids.tolist())) != len(ids):
        raise ValueError("'%s' in data should be unique." % sample_col)
    ok = np.ones(len(ids), dtype=bool)
    for k in used:                                                          # na.omit(data[, c(sample.col, vars)])
        ok &= ~np.isnan(np.asarray(data[k], dtype=np.float64))
    g = G.read_gds_genotypes(str(path), threads=threads)
    row_of = {sid: i for i, sid in enumerate(ids.astype(g.sample_id.dtype).tolist()) if ok[i]}
    file_sel = np.array([i for i, sid in enumerate(g.sample_id.tolist()) if sid in row_of], dtype=np.int32)
    if len(file_sel) == 0:
        raise ValueError("No common sample.id between 'data' and the GDS file.")
    rows = np.array([row_of[sid] for sid in g.sample_id[file_sel].tolist()], dtype=np.int64)   # data <- data[match(sid, ...), ]
    sub = {k: np.asarray(v)[rows] for k, v in data.items()}
    sel = None if len(file_sel) == g.n_sample else file_sel

    def subset_bits(mask):
        """allele bits of the variants in `mask` (the node is [variant][sample][2] x 2 bits)"""
        per = g.n_sample * 4
        if per % 8 == 0:
            return np.ascontiguousarray(g.allele_bits[:g.n_variant * (per // 8)].reshape(g.n_variant, per // 8)[mask]).reshape(-1)
        nib = np.stack([g.allele_bits & 15, g.allele_bits >> 4], axis=1).reshape(-1)[:g.n_variant * g.n_sample]
        nib = nib.reshape(g.n_variant, g.n_sample)[mask].reshape(-1)
        if nib.size % 2:
            nib = np.append(nib, np.uint8(0))
        return (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)

    vid = np.asarray(g.variant_id)
    if variant_id is None:
        if verbose:
            print("Filtering variants:")
        r = ctx.store_gds_geno(g.allele_bits, g.n_sample, g.n_variant, sample_sel=sel, maf=maf, missing_rate=missing_rate)
        keep = r["variant_sel"].copy()
    else:
        keep = np.isin(vid, np.asarray(variant_id))
        ctx.store_gds_geno(subset_bits(keep), g.n_sample, int(keep.sum()), sample_sel=sel)
    n_kept = int(keep.sum())
    if max_num_snp > 0 and n_kept > max_num_snp:
        ctx.set_seed(seed)
        pick = np.sort(np.flatnonzero(keep)[ctx.sample_int(n_kept)[:int(max_num_snp)] - 1])
        keep = np.zeros_like(keep)
        keep[pick] = True
        ctx.store_gds_geno(subset_bits(keep), g.n_sample, int(keep.sum()), sample_sel=sel)
    if verbose:
        print("    # of samples: %d\n    # of variants: %d%s" % (len(file_sel), int(keep.sum()),
              " (randomly selected from %d)" % n_kept if n_kept > keep.sum() else ""))
    return sub, vid[keep]


def seqFitNullGLMM_SPA(formula: str, data: dict, packed_geno, trait_type="binary", sample_id=None,
                       variant_id=None, inv_norm=True, X_transform=True, tol=0.02, maxiter=20, nrun=30, tolPCG=1e-5,
                       maxiterPCG=500, num_marker=30, tau_init=(0, 0), traceCVcutoff=0.0025, ratioCVcutoff=0.001,
                       seed=200, verbose=False, ctx: Context | None = None, sample_col="sample.id", maf=0.005,
                       missing_rate=0.01, max_num_snp=1000000, threads: int = 8) -> NullModel:
    """Mirror of seqFitNullGLMM_SPA (R/saige_main.r:223-654).

    `packed_geno` is the path of a SeqArray GDS file -- read without SeqArray (gds.py); `data[sample_col]`, `maf`, `missing_rate`,
    `max_num_snp` and `variant_id` then select samples and variants like the reference (:273-333) and the model carries the
    file's sample and variant ids -- or genotypes already in memory, replacing the loading at :388-421: a uint8 array
    [n_variant][ceil(n_samp/4)] in the 2-bit format (geno.sparse=FALSE, SeqArray:::.seqGet2bGeno) or a list of
    `saige_get_sparse` vectors (geno.sparse=TRUE, the reference's default); sample / variant filtering is then the caller's job.
    Linearly dependent covariate columns are dropped before the QR transform like the reference does (:362-376); the returned
    coefficients then belong to the kept columns.  The random marker order of the variance-ratio step mirrors R's
    `set.seed(seed, sample.kind = "Rounding")`, the setting the reference's golden fixtures were produced with -- under R >= 3.6's
    default sample.kind ("Rejection") an R session draws a different marker order.
    """
    if trait_type not in ("binary", "quantitative"):
        raise ValueError("Invalid 'trait.type'.")
    ctx = ctx or default_context()
    phenovar, terms, intercept = rsetup.parse_formula(formula)
    from_gds = isinstance(packed_geno, (str, os.PathLike))
    if from_gds:
        data, gds_variant_id = _gds_front_end(ctx, packed_geno, data, [phenovar] + list(terms), sample_col, maf, missing_rate,
                                              max_num_snp, variant_id, seed, verbose, threads)
        sample_id, variant_id = data[sample_col], gds_variant_id
    y = np.asarray(data[phenovar], dtype=np.float64)
    n = len(y)
    X = rsetup.model_matrix(data, terms, intercept)
    if X.shape[1] <= 1:
        X_transform = False
    X_qrr = None
    if X_transform:
        keep = rsetup.independent_columns(X)                               # :362-376: covariates with NA coefficients are dropped
        if len(keep) < X.shape[1]:
            if verbose:                                                     # same text as :369-373
                names = (["(Intercept)"] if intercept else []) + list(terms)
                drop = [names[j] for j in range(X.shape[1]) if j not in set(keep.tolist())]
                print("    exclude %d covariates (%s) to avoid multi collinearity." % (len(drop), ", ".join(drop)))
            X = X[:, keep]
        X, X_qrr = rsetup.qr_transform(X)                                   # :378-380
    if from_gds:
        pass                                                                # stored by _gds_front_end
    elif isinstance(packed_geno, np.ndarray) and packed_geno.ndim == 2:
        ctx.saige_store_2b_geno(packed_geno, n)                             # :437
    else:
        ctx.saige_store_sp_geno(packed_geno, n)                             # :434
    n_var = ctx.n_var_total
    param = make_param(tol=tol, tolPCG=tolPCG, seed=seed, maxiter=maxiter, maxiterPCG=maxiterPCG, nrun=nrun,
                       num_marker=num_marker, traceCVcutoff=traceCVcutoff, ratioCVcutoff=ratioCVcutoff, verbose=verbose)
    if trait_type == "binary":
        if len(np.unique(y)) != 2:
            raise ValueError("The outcome variable has more than 2 categories!")
        fit0 = rsetup.glm_binomial(X, y)                                    # :480
        noK = rsetup.null_model_binary(X, fit0)                             # :488
        tau = rsetup.initial_tau_binary(tau_init)                           # :491-497
        glmm = ctx.saige_fit_AI_PCG_binary(fit0, X, tau, param)             # :501
        ctx.set_seed(seed)                                                  # :509
        vr = ctx.saige_calc_var_ratio_binary(fit0, glmm, noK, param, ctx.sample_int(n_var))
    else:
        if inv_norm:                                                        # :536-548
            f = rsetup.glm_gaussian(X, y)
            y = rsetup.rank_norm(f.residuals) * rsetup.sd(f.residuals)
        fit0 = rsetup.glm_gaussian(X, y)                                    # :551
        noK = rsetup.null_model_quant(X, fit0)                              # :560-570
        tau = rsetup.initial_tau_quant(fit0, tau_init)                      # :573-583
        glmm = ctx.saige_fit_AI_PCG_quant(fit0, noK.X1, tau, param)         # :586
        ctx.set_seed(seed)                                                  # :594
        vr = ctx.saige_calc_var_ratio_quant(fit0, glmm, noK, param, ctx.sample_int(n_var))
    order = np.argsort(vr["id"], kind="stable")                             # :512-513
    vr = {k: v[order] for k, v in vr.items()}
    if variant_id is not None:
        vr["id"] = np.asarray(variant_id)[vr["id"] - 1]
    coef = glmm["coefficients"]
    if X_transform:
        coef = np.linalg.solve(X_qrr, coef * np.sqrt(n))                    # :620
    return NullModel(coefficients=coef, tau=glmm["tau"], linear_predictors=glmm["linear_predictors"],
                     fitted_values=glmm["fitted_values"], residuals=glmm["residuals"], cov=glmm["cov"],
                     converged=glmm["converged"], obj_noK=noK, var_ratio=vr, trait_type=trait_type,
                     sample_id=None if sample_id is None else np.asarray(sample_id),
                     variant_id=None if variant_id is None else np.asarray(variant_id))
