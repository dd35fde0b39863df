"""Host-side mirror of the single-variant association scan.

  reference                                                     here
  ---------------------------------------------------------     -------------------------------
  .init_nullmod              (R/assoc_single.r:17-67)           init_nullmod
  saige_score_test_init      (src/saige_main.cpp:101-155)       ScoreTest.__init__
  saige_score_test_bin/quant (src/saige_main.cpp:188-407)       ScoreTest.test / ScoreTest.test_stored
  seqAssocGLMM_SPA           (R/assoc_single.r:92-334)          seqAssocGLMM_SPA

The reference calls the native test once per variant from seqApply; here a batch of variants goes to the GPU in one call
(libsaigegds_b200.so, csrc/score.cu).  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib as L
from .api import Context, NullModel, _f64, _p, default_context

COLUMNS = ("AF.alt", "mac", "num", "beta", "SE", "pval", "p.norm", "converged")


def init_nullmod(modobj: NullModel, ii=None, maf=float("nan"), mac=10.0, missing=0.1, spa_pval=0.05,
                 var_ratio=float("nan")) -> dict:
    """.init_nullmod (R/assoc_single.r:17-67): the derived model arrays.  `ii` selects / reorders samples (1-based in R,
    0-based here; None = all).  K x n matrices of the reference are returned as C-contiguous [n][K] (the same memory)."""
    if not np.isfinite(var_ratio):
        r = np.asarray(modobj.var_ratio["ratio"], dtype=np.float64)             # seqAssocGLMM_SPA, :157-158
        r = r[~np.isnan(r)]
        var_ratio = float(np.mean(r)) if len(r) else float("nan")
    if not np.isfinite(var_ratio):
        raise ValueError("Invalid variance ratio in the SAIGE model.")
    if modobj.trait_type not in ("binary", "quantitative"):
        raise ValueError("Invalid 'modobj$trait.type': %s." % modobj.trait_type)
    noK = modobj.obj_noK
    y, mu = _f64(noK.y), _f64(modobj.fitted_values)
    ii = np.arange(len(y)) if ii is None else np.asarray(ii, dtype=np.int64)
    X1 = _f64(noK.X1)[ii]
    XXVX_inv = _f64(noK.XXVX_inv)[ii]
    V = _f64(noK.V)[ii]
    m = dict(maf=maf, mac=mac, missing=missing, spa_pval=spa_pval, trait=modobj.trait_type, tau=_f64(modobj.tau),
             y=np.ascontiguousarray(y[ii]), mu=np.ascontiguousarray(mu[ii]), var_ratio=float(var_ratio))
    m["y_mu"] = m["y"] - m["mu"]
    m["mu2"] = m["mu"] * (1 - m["mu"])
    m["t_XXVX_inv"] = np.ascontiguousarray(XXVX_inv)
    m["XV"] = np.ascontiguousarray(_f64(noK.XV)[:, ii].T)
    m["t_XVX_inv_XV"] = np.ascontiguousarray(XXVX_inv * V[:, None])
    m["t_X"] = np.ascontiguousarray(X1)
    if modobj.trait_type == "binary":
        m["XVX"] = np.ascontiguousarray(X1.T @ (X1 * m["mu2"][:, None]))
    else:
        m["XVX"] = np.ascontiguousarray(X1.T @ X1)
    m["S_a"] = _f64((X1 * m["y_mu"][:, None]).sum(axis=0))
    return m


class ScoreTest:
    """saige_score_test_init + saige_score_test_bin / _quant on one GPU context."""

    def __init__(self, mobj: dict, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self.mobj = mobj
        self.n, self.K = mobj["t_X"].shape
        keep = {k: _f64(mobj[k]) for k in ("tau", "y", "mu", "y_mu", "mu2", "t_XXVX_inv", "XV", "t_XVX_inv_XV", "t_X", "XVX", "S_a")}
        sm = L.ScoreModel(0 if mobj["trait"] == "binary" else 1, self.n, self.K, *[_p(keep[k]) for k in
                          ("tau", "y", "mu", "y_mu", "mu2", "t_XXVX_inv", "XV", "t_XVX_inv_XV", "t_X", "XVX", "S_a")],
                          float(mobj["var_ratio"]))
        L.check(L.lib().sgb_score_test_init(self.ctx._h, C.byref(sm), C.c_double(mobj["maf"]), C.c_double(mobj["mac"]),
                                            C.c_double(mobj["missing"]), C.c_double(mobj["spa_pval"])))

    def set_path(self, name: str):
        """'tensor' (default for packed genotypes: class sums on the tcgen05 tensor cores + a per-variant finishing kernel),
        'tiled' (shared-memory-tiled CUDA-core kernel for every variant; always used for dosages) -- either way the saddle-point
        candidates go through the per-variant kernel -- or 'per_variant' (every variant through the per-variant kernel)."""
        L.check(L.lib().sgb_score_test_set_path(self.ctx._h, C.c_int({"tiled": 0, "per_variant": 1, "tensor": 2}[name])))

    @staticmethod
    def _result(out, valid):
        r = {k: out[:, i].copy() for i, k in enumerate(COLUMNS)}
        r["valid"] = valid.astype(bool)
        return r

    def test(self, geno: np.ndarray) -> dict:
        """geno: uint8 [n_var][ceil(n/4)] 2-bit codes, or float64 [n_var][n] dosages (NaN = missing).  Returns the columns
        of the reference's result vector per variant plus `valid` (False where the reference returns NULL)."""
        if geno.ndim != 2:
            raise L.InvalidArgument(L.SGB_ERR_INVALID, "Input dosage should be a matrix.")
        nv = geno.shape[0]
        out = np.empty((nv, 8))
        valid = np.empty(nv, dtype=np.int32)
        if geno.dtype == np.uint8:
            g = np.require(geno, requirements=["C", "A"])
            L.check(L.lib().sgb_score_test_packed(self.ctx._h, _p(g, C.c_ubyte), C.c_int64(g.shape[1]), C.c_int64(nv),
                                                  _p(out), _p(valid, C.c_int32)))
        elif geno.dtype == np.float64:
            if geno.shape[1] != self.n:
                raise L.InvalidArgument(L.SGB_ERR_INVALID, "Invalid dimension of dosages: %dx%d." % geno.shape)
            g = np.require(geno, requirements=["C", "A"])
            L.check(L.lib().sgb_score_test_dosage(self.ctx._h, _p(g), C.c_int64(nv), _p(out), _p(valid, C.c_int32)))
        else:
            raise L.InvalidArgument(L.SGB_ERR_INVALID, "Invalid type of dosages.")
        return self._result(out, valid)

    def test_stored(self, first: int, n_variant: int):
        """The same test on variants [first, first + n_variant) of the genotype matrix the context already stores.
        Returns (result, kernel milliseconds)."""
        out = np.empty((n_variant, 8))
        valid = np.empty(n_variant, dtype=np.int32)
        ms = C.c_float(0)
        L.check(L.lib().sgb_score_test_stored(self.ctx._h, C.c_int64(first), C.c_int64(n_variant), _p(out),
                                              _p(valid, C.c_int32), C.byref(ms)))
        return self._result(out, valid), float(ms.value)


def _scan_gds(gdsfile, modobj: NullModel, maf, mac, missing, spa_pval, var_ratio, batch_variants, kernel_path, ctx, threads):
    """seqAssocGLMM_SPA on a GDS file (R/assoc_single.r:116-155, 233-300): the file is read without SeqArray (gds.py), the samples
    of the model are selected in file order (`seqSetFilter(sample.id=)`), the model rows are reordered to that order
    (`ii <- match(sid, modobj$sample.id)`), the genotype node is turned into 2-bit dosage rows on the device (they replace the
    genotypes the context stores) and scanned from HBM in batches."""
    from . import gds as G
    if modobj.sample_id is None:
        raise ValueError("The model has no sample IDs to match against the GDS file.")
    g = G.read_gds_genotypes(str(gdsfile), threads=threads)
    model_ids = np.asarray(modobj.sample_id).astype(g.sample_id.dtype)
    pos = {sid: i for i, sid in enumerate(model_ids.tolist())}
    in_model = np.array([s in pos for s in g.sample_id.tolist()])
    sel = np.flatnonzero(in_model).astype(np.int32)
    if len(sel) != len(model_ids):
        raise ValueError("Some of sample IDs are not available in the GDS file.")
    ii = np.array([pos[s] for s in g.sample_id[sel].tolist()], dtype=np.int64)
    if g.n_variant <= 0:
        raise ValueError("No variant in the genotypic data set!")
    mobj = init_nullmod(modobj, ii, maf, mac, missing, spa_pval, var_ratio)
    ctx = ctx or default_context()
    r = ctx.store_gds_geno(g.allele_bits, g.n_sample, g.n_variant, sample_sel=None if len(sel) == g.n_sample else sel)
    st = ScoreTest(mobj, ctx)
    if kernel_path is not None:
        st.set_path(kernel_path)
    n_stored = int(np.sum(r["variant_sel"]))
    parts = [st.test_stored(a, min(batch_variants, n_stored - a))[0] for a in range(0, n_stored, batch_variants)]
    res = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    return res, np.asarray(g.variant_id)[r["variant_sel"]]


def seqAssocGLMM_SPA(geno, modobj: NullModel, maf=float("nan"), mac=10.0, missing=0.1, spa_pval=0.05,
                     var_ratio=float("nan"), variant_id=None, sample_index=None, batch_bytes=1 << 30,
                     kernel_path: str | None = None, ctx: Context | None = None, threads: int = 8) -> dict:
    """Mirror of seqAssocGLMM_SPA (R/assoc_single.r:92-334).  `geno` is the path of a SeqArray GDS file (read without SeqArray, the
    model's `sample_id` selects and orders the samples as the reference does) or genotypes already in memory (2-bit packed uint8
    [n_var][ceil(n/4)] or float64 dosages [n_var][n]; `sample_index` then reorders the model).  Returns the data.frame columns id,
    AF.alt, mac, num, beta, SE, pval (+ p.norm, converged for binary traits) of the variants that pass the filters."""
    if isinstance(geno, (str, os.PathLike)):
        res, ids_all = _scan_gds(geno, modobj, maf, mac, missing, spa_pval, var_ratio, 1 << 16, kernel_path, ctx, threads)
        return _answer(res, ids_all if variant_id is None else np.asarray(variant_id), modobj)
    if geno.shape[0] <= 0:
        raise ValueError("No variant in the genotypic data set!")
    mobj = init_nullmod(modobj, sample_index, maf, mac, missing, spa_pval, var_ratio)
    st = ScoreTest(mobj, ctx)
    if kernel_path is not None:
        st.set_path(kernel_path)
    step = max(1, int(batch_bytes // max(1, geno.shape[1] * geno.itemsize)))
    parts = [st.test(geno[a:a + step]) for a in range(0, geno.shape[0], step)]
    res = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    return _answer(res, np.arange(1, geno.shape[0] + 1) if variant_id is None else np.asarray(variant_id), modobj)


def _answer(res: dict, ids, modobj: NullModel) -> dict:
    keep = res.pop("valid")
    ans = {"id": ids[keep]}
    for k in COLUMNS:
        ans[k] = res[k][keep]
    ans["num"] = ans["num"].astype(np.int64)
    if modobj.trait_type == "binary":
        ans["converged"] = ans["converged"] == 1
    else:
        del ans["p.norm"], ans["converged"]
    return ans
