"""Single-variant score test + saddle-point approximation (src/saige_main.cpp:188-407, src/SPATest.cpp).

CPU part: the device body (saigegds_b200/csrc/score_body.h) is compiled for one host thread by tests/native/
score_body_check.cpp and compared with the reference's golden p-values and with the oracle -- the arithmetic is checked
before it reaches a GPU.  GPU part (-m gpu): the CUDA kernel through the C-ABI against the same goldens and the oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import saigegds_b200 as sg
from saigegds_b200 import rsetup
from conftest import ROOT, dosage_all

NAMES = ["AF.alt", "mac", "num", "beta", "SE", "pval", "p.norm", "converged"]
ORC = {"AF.alt": "AF", "p.norm": "p_norm"}


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if len(a) else 0.0


def golden_modobj(fx, trait):
    md = fx.model if trait == "binary" else fx.model_quant
    noK = rsetup.ObjNoK(y=md["noK_y"], mu=md["noK_mu"], res=md["noK_res"], V=md["noK_V"], X1=md["noK_X1"], XV=md["noK_XV"],
                        XXVX_inv=md["noK_XXVX_inv"])
    return sg.NullModel(coefficients=md["coefficients"], tau=md["tau"], linear_predictors=md["linear_predictors"],
                        fitted_values=md["fitted_values"], residuals=md["residuals"], cov=md["cov"], converged=True,
                        obj_noK=noK, var_ratio={"ratio": md["vr_ratio"]}, trait_type=trait)


def oracle_model(fx, trait):
    from oracle import oracle as orc
    md = fx.model if trait == "binary" else fx.model_quant
    return orc.init_nullmod(trait, md["noK_y"], md["fitted_values"], md["noK_X1"], md["noK_XV"], md["noK_XXVX_inv"],
                            md["noK_V"], md["tau"]), float(np.mean(md["vr_ratio"]))


def random_dosages(rng, n, n_var, integer):
    """Rare to common variants, coded allele major or minor, 0-30 % missing, optionally real-valued dosages."""
    af = np.concatenate([rng.uniform(0.002, 0.05, n_var // 2), rng.uniform(0.05, 0.98, n_var - n_var // 2)])
    d = rng.binomial(2, af[:, None], size=(n_var, n)).astype(np.float64)
    if not integer:
        d = np.clip(d + rng.uniform(-0.2, 0.2, size=d.shape) * (rng.random(d.shape) < 0.3), 0, 2)
    miss = rng.choice([0.0, 0.01, 0.08, 0.3], size=n_var)
    d[rng.random(d.shape) < miss[:, None]] = np.nan
    d[0] = 0.0            # monomorphic -> filtered (maf == 0)
    d[1] = np.nan         # nothing called -> filtered (num == 0)
    return d


def pack(d):
    c = np.where(np.isnan(d), 3, d).astype(np.uint8)
    n_var, n = c.shape
    nb = (n + 3) // 4
    full = np.zeros((n_var, nb * 4), dtype=np.uint8)      # pad samples: code 0, never read
    full[:, :n] = c
    q = full.reshape(n_var, nb, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)


def compare(got, ref, tol, exact_counts=True):
    v = ref["valid"]
    assert np.array_equal(got["valid"], v)
    for k in NAMES:
        a, b = got[k][v], ref[ORC.get(k, k)][v]
        if k in ("num", "converged") or (exact_counts and k in ("AF.alt", "mac")):
            assert np.array_equal(a, b), k
        else:
            assert relmax(a, b) < tol, (k, relmax(a, b))
    assert np.all(np.isnan(got["pval"][~v]))


# ------------------------------------------------------------------ CPU: the device body on one host thread
@pytest.fixture(scope="module")
def body(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("native") / "score_body_check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "native", "score_body_check.cpp")])
    lib = C.CDLL(so)

    def run(m, var_ratio, geno, maf=float("nan"), mac=10.0, missing=0.1, spa_pval=0.05):
        P = lambda a, t=C.c_double: a.ctypes.data_as(C.POINTER(t))     # noqa: E731
        geno = np.ascontiguousarray(geno)
        n_var = geno.shape[0]
        n, K = m["t_X"].shape
        out = np.empty((n_var, 8))
        valid = np.empty(n_var, dtype=np.int32)
        packed = geno.dtype == np.uint8
        rc = lib.score_body_check(
            C.c_int(0 if m["trait"] == "binary" else 1), C.c_long(n), C.c_int(K), C.c_double(m["tau"][0]), P(m["mu"]),
            P(m["y_mu"]), P(m["mu2"]), P(m["t_XVX_inv_XV"]), P(m["XVX"]), P(m["t_X"]), P(m["S_a"]), C.c_double(var_ratio),
            C.c_double(maf), C.c_double(mac), C.c_double(missing), C.c_double(spa_pval), C.c_long(n_var),
            None if packed else P(geno), P(geno, C.c_ubyte) if packed else None, C.c_long(geno.shape[1]), P(out),
            P(valid, C.c_int))
        assert rc == 0
        r = {k: out[:, i].copy() for i, k in enumerate(NAMES)}
        r["valid"] = valid.astype(bool)
        return r
    run.set_dual = lambda on: lib.score_body_set_dual(C.c_int(int(on)))
    return run


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
@pytest.mark.parametrize("source", ["packed", "dosage"])
def test_device_body_reproduces_golden_pvalues(body, fx, trait, source):
    """All 10,000 rows of saige_pval*.rds, including the 434 saddle-point-adjusted ones (test_SAIGE.R:79-106)."""
    m, vr = oracle_model(fx, trait)
    pv = fx.pval if trait == "binary" else fx.pval_quant
    r = body(m, vr, fx.packed_all if source == "packed" else dosage_all(fx), mac=4.0)
    ids = pv["id"] - 1
    assert r["valid"][ids].all() and r["valid"].sum() == len(ids)
    assert np.array_equal(r["AF.alt"][ids], pv["AF_alt"]) and np.array_equal(r["mac"][ids], pv["mac"])
    assert np.array_equal(r["num"][ids].astype(np.int64), pv["num"].astype(np.int64))
    assert relmax(r["pval"][ids], pv["pval"]) < 1e-10
    assert relmax(r["beta"][ids], pv["beta"]) < 1e-9 and relmax(r["SE"][ids], pv["SE"]) < 1e-9
    if trait == "binary":
        assert relmax(r["p.norm"][ids], pv["p_norm"]) < 1e-10
        assert np.array_equal(r["converged"][ids].astype(int), pv["converged"])
        # rows the saddle-point step really moved (inside the cutoff it returns the normal p-value again, :311-313)
        moved = np.abs(pv["pval"] - pv["p_norm"]) > 1e-9 * pv["p_norm"]
        assert np.array_equal(np.abs(r["pval"][ids] - r["p.norm"][ids]) > 1e-9 * r["p.norm"][ids], moved) and moved.sum() > 300


def test_both_roots_in_one_pass_equals_two_newton_iterations_bit_for_bit(body, fx):
    """saddle_prob_dual (the candidate kernel of the tensor scan: one pass over the (g, mu) pairs serves the current Newton step of
    both roots) against saddle_prob (SPATest.cpp's order) on one host thread: every root takes the same sequence of steps, so all
    10,000 golden rows -- 434 of them saddle-point adjusted -- and a low-threshold run with many more candidates agree bit for bit."""
    for kw in (dict(mac=4.0), dict(mac=2.0, spa_pval=0.5, missing=0.3)):
        body.set_dual(False)
        a = body(oracle_model(fx, "binary")[0], oracle_model(fx, "binary")[1], fx.packed_all[:4000], **kw)
        body.set_dual(True)
        try:
            b = body(oracle_model(fx, "binary")[0], oracle_model(fx, "binary")[1], fx.packed_all[:4000], **kw)
        finally:
            body.set_dual(False)
        assert np.array_equal(a["valid"], b["valid"])
        assert all(np.array_equal(a[k], b[k], equal_nan=True) for k in NAMES)
        assert np.sum(a["pval"][a["valid"]] != a["p.norm"][a["valid"]]) > 100


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
@pytest.mark.parametrize("integer", [True, False])
def test_device_body_matches_oracle_with_missing_and_filters(body, fx, trait, integer):
    from oracle import oracle as orc
    m, vr = oracle_model(fx, trait)
    d = random_dosages(np.random.default_rng(31 + integer), fx.n_samp, 600, integer)
    kw = dict(maf=0.001, mac=3.0, missing=0.25, spa_pval=0.2)
    ref = orc.score_test(m, d, vr, **kw)
    assert 0 < ref["valid"].sum() < len(d)
    compare(body(m, vr, d, **kw), ref, 1e-9, exact_counts=integer)
    if integer:
        compare(body(m, vr, pack(d), **kw), ref, 1e-9)


def test_init_nullmod_matches_the_reference_arrays(fx):
    for trait in ("binary", "quantitative"):
        want, vr = oracle_model(fx, trait)
        got = sg.init_nullmod(golden_modobj(fx, trait))
        assert got["var_ratio"] == vr
        for k in ("tau", "y", "mu", "y_mu", "mu2", "t_XXVX_inv", "XV", "t_XVX_inv_XV", "XVX", "t_X", "S_a"):
            assert np.array_equal(got[k], want[k]), k
    sub = sg.init_nullmod(golden_modobj(fx, "binary"), ii=[5, 2, 9])
    assert sub["t_X"].shape == (3, 3) and sub["y"].tolist() == fx.model["noK_y"][[5, 2, 9]].tolist()
    bad = golden_modobj(fx, "binary")
    bad.var_ratio = {"ratio": np.array([np.nan])}
    with pytest.raises(ValueError, match="Invalid variance ratio"):
        sg.init_nullmod(bad)


# ------------------------------------------------------------------ GPU: the CUDA kernel through the C-ABI
@pytest.mark.gpu
@pytest.mark.parametrize("trait", ["binary", "quantitative"])
@pytest.mark.parametrize("path", ["tensor", "tiled", "per_variant"])
def test_gpu_score_test_reproduces_golden_pvalues(gpu, fx, trait, path):
    pv = fx.pval if trait == "binary" else fx.pval_quant
    ans = sg.seqAssocGLMM_SPA(fx.packed_all, golden_modobj(fx, trait), mac=4, ctx=gpu, kernel_path=path)
    assert np.array_equal(ans["id"], pv["id"])
    assert np.array_equal(ans["AF.alt"], pv["AF_alt"]) and np.array_equal(ans["mac"], pv["mac"])
    assert np.array_equal(ans["num"], pv["num"].astype(np.int64))
    assert relmax(ans["pval"], pv["pval"]) < 1e-9
    assert relmax(ans["beta"], pv["beta"]) < 1e-8 and relmax(ans["SE"], pv["SE"]) < 1e-8
    if trait == "binary":
        assert relmax(ans["p.norm"], pv["p_norm"]) < 1e-9
        assert np.array_equal(ans["converged"].astype(int), pv["converged"])
        moved = np.abs(pv["pval"] - pv["p_norm"]) > 1e-9 * pv["p_norm"]
        assert np.array_equal(np.abs(ans["pval"] - ans["p.norm"]) > 1e-9 * ans["p.norm"], moved) and moved.sum() > 300
    else:
        assert "p.norm" not in ans
    # real-valued dosage input, in two batches: same numbers
    ans2 = sg.seqAssocGLMM_SPA(dosage_all(fx), golden_modobj(fx, trait), mac=4, ctx=gpu, batch_bytes=40 << 20,
                               kernel_path=path)
    assert np.array_equal(ans2["id"], ans["id"]) and relmax(ans2["pval"], ans["pval"]) < 1e-12
    if path == "tensor":
        # dosages go through the FP64 tiled kernel, packed codes through the fixed-point class sums: equal to the quantisation of the
        # model columns (2^-46 of a column's largest element), i.e. relative to the scale of beta, not to a beta that is nearly zero
        assert np.max(np.abs(ans2["beta"] - ans["beta"])) < 1e-12 * np.max(np.abs(ans["beta"]))
        assert relmax(ans2["beta"], ans["beta"]) < 1e-8
    else:
        assert relmax(ans2["beta"], ans["beta"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_gpu_score_test_matches_oracle_with_missing_and_filters(gpu, fx, trait):
    from oracle import oracle as orc
    m, vr = oracle_model(fx, trait)
    st = sg.ScoreTest(sg.init_nullmod(golden_modobj(fx, trait), maf=0.001, mac=3.0, missing=0.25, spa_pval=0.2), gpu)
    kw = dict(maf=0.001, mac=3.0, missing=0.25, spa_pval=0.2)
    for path in ("tensor", "tiled", "per_variant"):
        st.set_path(path)
        for integer in (True, False):
            d = random_dosages(np.random.default_rng(31 + integer), fx.n_samp, 600, integer)
            ref = orc.score_test(m, d, vr, **kw)
            compare(st.test(d), ref, 1e-8, exact_counts=integer)
            if integer:
                compare(st.test(pack(d)), ref, 1e-8)
    # bit-reproducible: fixed reduction order, no atomics on the data path
    a, b = st.test(d), st.test(d)
    assert all(np.array_equal(a[k], b[k], equal_nan=True) for k in NAMES)


@pytest.mark.gpu
def test_gpu_score_test_on_the_stored_matrix_and_errors(gpu, fx):
    st = sg.ScoreTest(sg.init_nullmod(golden_modobj(fx, "binary"), mac=4.0), gpu)
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    host = st.test(fx.packed[100:1100])
    stored, ms = st.test_stored(100, 1000)
    assert ms > 0 and all(np.array_equal(host[k], stored[k], equal_nan=True) for k in NAMES + ["valid"])
    with pytest.raises(sg.InvalidArgument):
        st.test_stored(len(fx.packed) - 5, 10)
    with pytest.raises(sg.InvalidArgument, match="Invalid type of dosages"):
        st.test(np.zeros((2, fx.n_samp), dtype=np.float32))
    with pytest.raises(sg.InvalidArgument, match="Invalid dimension"):
        st.test(np.zeros((2, fx.n_samp + 1)))
    with pytest.raises(sg.InvalidArgument):
        st.test(np.zeros((2, 7), dtype=np.uint8))


@pytest.mark.gpu
def test_gpu_fit_then_gpu_scan_matches_golden_pvalues(gpu, fx):
    """test.saige_pval end to end on the GPU: null model through the sparse entry, then the association scan; the
    reference's own tolerance is 1e-7 absolute on p-values (test_SAIGE.R:98-105), ours 1e-6 relative."""
    from oracle import oracle as orc
    sp = [orc.get_sparse(c) for c in
          np.stack([(fx.packed >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(len(fx.packed), -1)[:, :fx.n_samp].astype(np.uint8)]
    mod = sg.seqFitNullGLMM_SPA("y ~ x1 + x2", fx.pheno, sp, trait_type="binary", variant_id=fx.variant_id, ctx=gpu)
    ans = sg.seqAssocGLMM_SPA(fx.packed_all, mod, mac=4, ctx=gpu)
    pv = fx.pval
    assert np.array_equal(ans["id"], pv["id"])
    assert relmax(ans["pval"], pv["pval"]) < 1e-6
    assert relmax(ans["beta"], pv["beta"]) < 1e-6


def synthetic_model(rng, n, K, trait):
    """A self-consistent null model with K fixed-effect columns (what .init_nullmod would hand over)."""
    from oracle import oracle as orc
    X = np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(K - 1)])
    beta = np.concatenate([[-1.0], rng.normal(0, 0.15, K - 1)])
    if trait == "binary":
        mu = 1 / (1 + np.exp(-(X @ beta)))
        y = (rng.random(n) < mu).astype(np.float64)
        V = mu * (1 - mu) * rng.uniform(0.9, 1.1, n)          # the glm's V differs slightly from the mixed model's mu(1-mu)
    else:
        mu = X @ beta
        y = mu + rng.standard_normal(n)
        V = np.ones(n)
    XVX_inv = np.linalg.inv(X.T @ (X * V[:, None]))
    return orc.init_nullmod(trait, y, mu, X, (X * V[:, None]).T.copy(), X @ XVX_inv, V, np.array([1.0, 0.4]))


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
@pytest.mark.parametrize("K", [1, 8, 17, 32])
def test_device_body_matches_oracle_for_other_covariate_counts(body, trait, K):
    """The fixture has K = 3; the single-pass algebra must hold for any K (1 = intercept only, > 16 = the guarded kernels)."""
    from oracle import oracle as orc
    rng = np.random.default_rng(100 + K)
    n = 700
    m = synthetic_model(rng, n, K, trait)
    d = random_dosages(rng, n, 200, integer=True)
    kw = dict(maf=0.002, mac=2.0, missing=0.35, spa_pval=0.3)
    ref = orc.score_test(m, d, 0.9, **kw)
    assert ref["valid"].sum() > 100
    compare(body(m, 0.9, d, **kw), ref, 1e-8)
    compare(body(m, 0.9, pack(d), **kw), ref, 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("trait", ["binary", "quantitative"])
@pytest.mark.parametrize("K", [1, 8, 17, 32])
def test_gpu_score_test_for_other_covariate_counts(gpu, trait, K):
    """The CUDA kernels for covariate counts other than the fixture's K = 3: exact-K instantiations (1, 8) and the guarded
    K > 16 variant with 128-sample tiles (17, 32), both kernel paths, dosage and packed input, against the oracle."""
    from oracle import oracle as orc
    rng = np.random.default_rng(100 + K)
    n = 700
    m = synthetic_model(rng, n, K, trait)
    d = random_dosages(rng, n, 200, integer=True)
    kw = dict(maf=0.002, mac=2.0, missing=0.35, spa_pval=0.3)
    ref = orc.score_test(m, d, 0.9, **kw)
    assert ref["valid"].sum() > 100
    st = sg.ScoreTest(dict(m, var_ratio=0.9, **kw), gpu)
    for path in ("tensor", "tiled", "per_variant"):
        st.set_path(path)
        compare(st.test(d), ref, 1e-8)
        compare(st.test(pack(d)), ref, 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_gpu_tensor_scan_equals_tiled_scan_over_many_boxes(gpu, trait):
    """The tcgen05 class-sum GEMMs (contraction split over CTAs, several 128-row blocks, a ragged last box and a ragged last row block)
    against the CUDA-core tiled kernel on the same packed block: n = 20,011 samples (40 TMA boxes of 512), 1,003 variants, K = 10."""
    rng = np.random.default_rng(77)
    n, K, n_var = 20011, 10, 1003
    m = synthetic_model(rng, n, K, trait)
    d = random_dosages(rng, n, n_var, integer=True)
    kw = dict(maf=0.0005, mac=2.0, missing=0.35, spa_pval=0.3)
    st = sg.ScoreTest(dict(m, var_ratio=0.9, **kw), gpu)
    p = pack(d)
    st.set_path("tensor")
    a = st.test(p)
    st.set_path("tiled")
    b = st.test(p)
    assert np.array_equal(a["valid"], b["valid"]) and a["valid"].sum() > 900
    v = a["valid"]
    for k in ("AF.alt", "mac", "num", "converged"):
        assert np.array_equal(a[k][v], b[k][v]), k
    for k in ("beta", "SE", "pval", "p.norm"):
        assert relmax(a[k][v], b[k][v]) < 1e-10, (k, relmax(a[k][v], b[k][v]))
    st.set_path("tensor")
    a2 = st.test(p)
    assert all(np.array_equal(a[k], a2[k], equal_nan=True) for k in NAMES)     # integer limbs: bit-reproducible
    # a variant's row does not depend on the batch it travels in (one TMA box of rows mostly out of bounds, odd offsets)
    for lo, hi in ((0, 1), (5, 8), (130, 387)):
        sub = st.test(p[lo:hi])
        assert all(np.array_equal(sub[k], a[k][lo:hi], equal_nan=True) for k in NAMES + ["valid"]), (lo, hi)
