"""GPU parity tests (pytest -m gpu): the CUDA library, called through its C-ABI, against the CPU oracle,
the committed goldens of the reference, and size-independent properties at larger shapes.

Tolerances (north star): integer work bit-exact; per-product vectors <= 1e-10 relative to the output's
infinity norm; tau, variance ratio and fitted values <= 1e-6 relative (observed far tighter).
"""
import numpy as np
import pytest

from conftest import random_packed

pytestmark = pytest.mark.gpu
PROD_TOL = 1e-10
FIT_TOL = 1e-6


def relinf(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def gstore(gpu, fx):
    lut, diag = gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    return dict(lut=lut, diag=diag)


# ---------------------------------------------------------------- store: counts, LUT, diag, decode
def test_allele_counts_bit_exact(gpu, gstore, oracle):
    nv, sm = gpu.allele_counts()
    onv, osm = oracle.allele_counts()
    assert np.array_equal(nv, onv) and np.array_equal(sm, osm)


def test_lut_bit_exact_and_diag(gpu, gstore, oracle):
    assert np.array_equal(gstore["lut"], oracle.lut)          # same IEEE operations on identical integers
    assert relinf(gstore["diag"], oracle.diag) < 1e-13


def test_decode_bit_exact(gpu, gstore, oracle):
    for j in (0, 17, 9975):
        a, b = gpu.get_geno_ds(j), oracle.get_geno_ds(j)
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))


@pytest.mark.parametrize("n_samp,n_var,missing", [(1, 3, 0.0), (7, 5, 0.3), (1001, 257, 0.05), (4099, 130, 0.02),
                                                  (2048, 128, 0.0), (6150, 33, 0.5)])
def test_ragged_shapes_missing_and_pad_codes(gpu, n_samp, n_var, missing):
    """N % 4 != 0 with arbitrary pad codes, missing genotypes, monomorphic and all-missing variants."""
    from oracle.oracle import Oracle
    rng = np.random.default_rng(n_samp * 1000 + n_var)
    packed = random_packed(rng, n_samp, n_var, missing)
    if n_var >= 5:
        packed[1, :] = 0x00        # monomorphic: af = 0 -> all-zero LUT (saige_fitnull.cpp:195-196)
        packed[2, :] = 0xFF        # no valid call at all
        packed[3, :] = 0xAA        # af = 1
    o = Oracle()
    olut, odiag = o.store_2b_geno(packed, n_samp)
    lut, diag = gpu.saige_store_2b_geno(packed, n_samp)
    nv, sm = gpu.allele_counts()
    onv, osm = o.allele_counts()
    assert np.array_equal(nv, onv) and np.array_equal(sm, osm)
    assert np.array_equal(lut, olut)
    assert relinf(diag, odiag) < 1e-12 or np.max(np.abs(odiag)) == 0
    for j in range(min(n_var, 4)):
        a, b = gpu.get_geno_ds(j), o.get_geno_ds(j)
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
    b = rng.standard_normal(n_samp)
    want = o.grm_mv(b)
    for kernel in ("simt", "imma", "imma2"):
        gpu.set_kernel(kernel)
        try:
            got = gpu.get_crossprod_b_grm(b)
        finally:
            gpu.set_kernel("auto")
        assert np.max(np.abs(got - want)) <= PROD_TOL * max(np.max(np.abs(want)), 1e-300) + 1e-300, kernel


# ---------------------------------------------------------------- product
def test_product_matches_oracle_on_fixture(gpu, gstore, oracle, fx):
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    rng = np.random.default_rng(1)
    for scale in (1.0, 1e-6, 1e8):
        b = rng.standard_normal(fx.n_samp) * scale
        assert relinf(gpu.get_crossprod_b_grm(b), oracle.grm_mv(b)) < PROD_TOL
    b = np.zeros(fx.n_samp)
    assert np.all(gpu.get_crossprod_b_grm(b) == 0)
    u = rng.integers(0, 2, fx.n_samp) * 2.0 - 1            # Rademacher vectors of the trace estimator
    assert relinf(gpu.get_crossprod_b_grm(u), oracle.grm_mv(u)) < PROD_TOL
    spike = np.zeros(fx.n_samp); spike[3] = 1e12; spike[500] = 1e-12   # wide dynamic range
    assert relinf(gpu.get_crossprod_b_grm(spike), oracle.grm_mv(spike)) < PROD_TOL


@pytest.mark.parametrize("kernel", ["simt", "imma", "imma2"])
def test_both_kernels_match_oracle_on_fixture(gpu, fx, oracle, kernel):
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    gpu.set_kernel(kernel)
    try:
        rng = np.random.default_rng(21)
        for scale in (1.0, 1e-9, 1e9):
            b = rng.standard_normal(fx.n_samp) * scale
            assert relinf(gpu.get_crossprod_b_grm(b), oracle.grm_mv(b)) < PROD_TOL
        b = rng.standard_normal(fx.n_samp) * 10.0 ** rng.uniform(-12, 12, fx.n_samp)    # 24 decades of dynamic range
        assert relinf(gpu.get_crossprod_b_grm(b), oracle.grm_mv(b)) < PROD_TOL
    finally:
        gpu.set_kernel("auto")


def test_imma_is_bit_reproducible(gpu, fx):
    """Exact integer arithmetic: repeated products are bit-identical."""
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    gpu.set_kernel("imma")
    try:
        b = np.random.default_rng(22).standard_normal(fx.n_samp)
        a1, a2 = gpu.get_crossprod_b_grm(b), gpu.get_crossprod_b_grm(b)
        assert np.array_equal(a1, a2)
    finally:
        gpu.set_kernel("auto")


def test_product_multi_rhs_equals_single(gpu, fx, oracle):
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    rng = np.random.default_rng(2)
    B = rng.standard_normal((fx.n_samp, 5))
    out = gpu.get_crossprod_b_grm(B)
    for k in range(5):
        assert relinf(out[:, k], oracle.grm_mv(B[:, k])) < PROD_TOL


@pytest.mark.parametrize("kernel", ["simt", "imma", "imma2"])
def test_product_properties_at_scale(gpu, kernel):
    """N=50K, M=4K synthetic (too big for the scalar oracle in seconds): linearity, symmetry, PSD, and a
    column-sampled check against the definition."""
    n, m = 50000, 4000
    gpu.set_kernel(kernel)
    try:
        lut, diag = gpu.store_synthetic(n, m, seed=7, missing_rate=0.01, want_outputs=True)
        rng = np.random.default_rng(3)
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        Ax, Ay = gpu.get_crossprod_b_grm(x), gpu.get_crossprod_b_grm(y)
        Axy = gpu.get_crossprod_b_grm(2.5 * x - 0.5 * y)
        assert relinf(Axy, 2.5 * Ax - 0.5 * Ay) < 1e-11                      # linearity
        assert abs(y @ Ax - x @ Ay) / abs(y @ Ax) < 1e-10                    # symmetry
        assert x @ Ax > 0                                                    # positive semi-definite
        # definition on decoded columns: x'Ax = (1/M) sum_j (g_j'x)^2; e_i'A e_i = diag_i
        tot = 0.0
        for j in range(0, m, 400):
            ds = gpu.get_geno_ds(j)
            code = np.where(np.isnan(ds), 3, ds).astype(np.int64)
            tot += (lut[j][code] @ x) ** 2
        e = np.zeros(n); e[12345] = 1.0
        assert abs(gpu.get_crossprod_b_grm(e)[12345] - diag[12345]) / diag[12345] < 1e-11
        assert np.isfinite(tot)
        if kernel != "simt":                                                 # the kernels agree at scale
            gpu.set_kernel("simt")
            assert relinf(Ax, gpu.get_crossprod_b_grm(x)) < PROD_TOL
    finally:
        gpu.set_kernel("auto")


@pytest.mark.parametrize("n,m,label", [(430000, 100000, "C3"), (430000, 300000, "C5")])
def test_full_size_products(gpu, n, m, label):
    """BASELINE.json's full shapes (C3: 430K x 100K, 10.75 GB packed; C5: 430K x 300K, 32 GB): far beyond the scalar oracle, so
    the three independent kernels (fused single pass -- all 140 CTAs and every tile --, two-pass tensor core, FP64 CUDA core)
    are checked against each other and through size-independent properties: linearity, symmetry, positive definiteness,
    e_i'A e_i = diag(GRM)_i, bit reproducibility, multi-RHS == single RHS."""
    try:
        lut, diag = gpu.store_synthetic(n, m, seed=200, missing_rate=0.005, want_outputs=True)
        rng = np.random.default_rng(5)
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        gpu.set_kernel("imma")                                                # fused single pass (forced)
        Ax, Ay = gpu.get_crossprod_b_grm(x), gpu.get_crossprod_b_grm(y)
        assert np.array_equal(Ax, gpu.get_crossprod_b_grm(x))                 # bit reproducible
        assert relinf(gpu.get_crossprod_b_grm(2.5 * x - 0.5 * y), 2.5 * Ax - 0.5 * Ay) < 1e-11
        assert abs(y @ Ax - x @ Ay) / abs(y @ Ax) < 1e-10
        assert x @ Ax > 0
        e = np.zeros(n); e[123457] = 1.0
        assert abs(gpu.get_crossprod_b_grm(e)[123457] - diag[123457]) / diag[123457] < 1e-11
        B = np.stack([x, y, e], axis=1)
        out = gpu.get_crossprod_b_grm(B)
        assert np.array_equal(out[:, 0], Ax) and np.array_equal(out[:, 1], Ay)
        gpu.set_kernel("imma2")                                               # two HBM passes
        assert relinf(gpu.get_crossprod_b_grm(x), Ax) < PROD_TOL
        if label == "C3":
            gpu.set_kernel("simt")                                            # FP64 CUDA cores, no quantisation anywhere
            assert relinf(gpu.get_crossprod_b_grm(x), Ax) < PROD_TOL
        gpu.set_kernel("auto")                                                # nearly full slices: auto == fused here
        assert np.array_equal(gpu.get_crossprod_b_grm(x), Ax)
    finally:
        gpu.set_kernel("auto")
        gpu.store_synthetic(1024, 64)                                         # release the big shard


def test_synthetic_generator_is_shard_invariant(gpu):
    full = gpu.synth_to_host(1003, 64, 0, seed=11)
    part = gpu.synth_to_host(1003, 20, 30, seed=11)
    assert np.array_equal(full[30:50], part)
    codes = np.stack([(full >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(64, -1)
    assert np.all(codes[:, 1003:] == 3)                      # pad codes are 3
    assert 0.0 < np.mean(codes[:, :1003] == 3) < 0.02        # ~0.5 % missing


# ---------------------------------------------------------------- diag sigma, PCG
def test_diag_sigma_and_pcg_match_oracle(gpu, fx, oracle, setup_binary):
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    mu = setup_binary["fit0"].fitted_values
    w = mu * (1 - mu)
    tau = np.array([1.0, 0.5])
    assert relinf(gpu.get_diag_sigma(w, tau), oracle.diag_sigma(w, tau)) < 1e-13
    B = np.column_stack([setup_binary["X"], np.random.default_rng(4).standard_normal((fx.n_samp, 2))])
    X, iters = gpu.PCG_diag_sigma(w, tau, B)
    for k in range(B.shape[1]):
        xo, ito = oracle.pcg(w, tau, B[:, k])
        assert iters[k] == ito                               # same iterate sequence
        assert relinf(X[:, k], xo) < 1e-10
    x0, it0 = gpu.PCG_diag_sigma(w, np.array([1.0, 0.0]), B[:, 0])   # tau[1] == 0 skips the GRM product (:568)
    xo, ito = oracle.pcg(w, [1.0, 0.0], B[:, 0])
    assert it0 == ito and relinf(x0, xo) < 1e-12


# ---------------------------------------------------------------- full fits against the reference's goldens
def test_binary_fit_matches_reference_golden(gpu, fx, setup_binary):
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    s, g = setup_binary, fx.model
    gpu.reset_stats()
    r = gpu.saige_fit_AI_PCG_binary(s["fit0"], s["X"], s["tau"])
    assert r["tau"][0] == 1.0 and abs(r["tau"][1] - g["tau"][1]) / g["tau"][1] < FIT_TOL
    coef = np.linalg.solve(s["R"], r["coefficients"] * np.sqrt(fx.n_samp))
    assert relinf(coef, g["coefficients"]) < FIT_TOL
    assert relinf(r["cov"], g["cov"]) < FIT_TOL
    assert relinf(r["fitted_values"], g["fitted_values"]) < FIT_TOL
    assert relinf(r["linear_predictors"], g["linear_predictors"]) < FIT_TOL
    assert relinf(r["residuals"], g["residuals"]) < FIT_TOL
    assert r["converged"]
    st = gpu.stats()
    # 898 products in the reference; 180 of them (GRM u_i, identical in every AI step) are cached here
    assert 0 < st["n_products"] <= 898


def test_quant_fit_matches_reference_golden(gpu, fx, setup_quant):
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    s, g = setup_quant, fx.model_quant
    r = gpu.saige_fit_AI_PCG_quant(s["fit0"], s["noK"].X1, s["tau"])
    assert abs(r["tau"][0] - g["tau"][0]) / g["tau"][0] < FIT_TOL and r["tau"][1] == 0
    assert relinf(r["cov"], g["cov"]) < FIT_TOL
    assert relinf(r["fitted_values"], g["fitted_values"]) < FIT_TOL
    assert r["converged"]


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_var_ratio_matches_reference_golden(gpu, fx, setup_binary, setup_quant, trait):
    import saigegds_b200 as sg
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    s, g = (setup_binary, fx.model) if trait == "binary" else (setup_quant, fx.model_quant)
    gpu.set_seed(200)
    ml = gpu.sample_int(len(fx.packed))
    fn = gpu.saige_calc_var_ratio_binary if trait == "binary" else gpu.saige_calc_var_ratio_quant
    vr = fn(s["fit0"], {"tau": g["tau"]}, s["noK"], sg.make_param(), ml)
    order = np.argsort(vr["id"])
    assert np.array_equal(fx.variant_id[vr["id"][order] - 1], g["vr_id"])           # same 30 markers
    assert np.array_equal(vr["mac"][order], g["vr_mac"]) and np.array_equal(vr["maf"][order], g["vr_maf"])
    for k in ("var1", "var2", "ratio"):
        assert relinf(vr[k][order], g["vr_" + k]) < FIT_TOL, k


@pytest.mark.parametrize("trait", ["binary", "quantitative"])
def test_pvalues_downstream_of_the_gpu_fit(gpu, fx, oracle, setup_binary, setup_quant, trait):
    """north_star: downstream seqAssocGLMM_SPA p-values <= 1e-6 relative.  The null model fitted on the GPU and the one
    fitted by the CPU oracle go through the same score test + SPA (the oracle's restatement, pinned against
    saige_pval*.rds in test_oracle_golden.py); all 10,000 variants, 434 of them saddle-point adjusted.  The reference's
    own golden p-values are reproduced to 1e-6 as well (its test uses 1e-7 on a bit-identical model)."""
    from conftest import dosage_all
    from oracle import oracle as orc
    import saigegds_b200 as sg
    s, g, pv = (setup_binary, fx.model, fx.pval) if trait == "binary" else (setup_quant, fx.model_quant, fx.pval_quant)
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    if trait == "binary":
        rg = gpu.saige_fit_AI_PCG_binary(s["fit0"], s["X"], s["tau"])
        ro = oracle.fit_AI_PCG("binary", s["fit0"], s["X"], s["tau"])
        vfn = gpu.saige_calc_var_ratio_binary
    else:
        rg = gpu.saige_fit_AI_PCG_quant(s["fit0"], s["noK"].X1, s["tau"])
        ro = oracle.fit_AI_PCG("quantitative", s["fit0"], s["X"], s["tau"])
        vfn = gpu.saige_calc_var_ratio_quant
    gpu.set_seed(200)
    ml = gpu.sample_int(len(fx.packed))
    vr_g = float(np.mean(vfn(s["fit0"], {"tau": rg["tau"]}, s["noK"], sg.make_param(), ml)["ratio"]))
    oracle.set_seed(200)
    vr_o = float(np.mean(oracle.calc_var_ratio(trait, s["fit0"], ro["tau"], s["noK"], oracle.sample_int(len(fx.packed)))["ratio"]))
    assert abs(vr_g - vr_o) / vr_o < FIT_TOL
    ds = dosage_all(fx)
    noK = s["noK"]
    res = {}
    for name, fit, vr in (("gpu", rg, vr_g), ("oracle", ro, vr_o)):
        m = orc.init_nullmod(trait, noK.y, fit["fitted_values"], noK.X1, noK.XV, noK.XXVX_inv, noK.V, fit["tau"])
        res[name] = orc.score_test(m, ds, vr, mac=4.0)
    a, b = res["gpu"], res["oracle"]
    assert np.array_equal(a["valid"], b["valid"]) and a["valid"].sum() == 10000
    assert np.array_equal(a["converged"], b["converged"])
    for k in ("pval", "beta", "SE"):
        assert np.max(np.abs(a[k] - b[k]) / np.abs(b[k])) < FIT_TOL, k
    ids = pv["id"] - 1
    assert np.max(np.abs(a["pval"][ids] - pv["pval"]) / pv["pval"]) < FIT_TOL


def test_driver_end_to_end_like_the_reference_test(fx):
    """test.saige_fit_null_model (inst/unitTests/test_SAIGE.R:44-76), tolerance 1e-6 instead of 1e-4."""
    import saigegds_b200 as sg
    for trait, col, g in (("binary", "y", fx.model), ("quantitative", "yy", fx.model_quant)):
        data = dict(fx.pheno)
        glmm = sg.seqFitNullGLMM_SPA("%s ~ x1 + x2" % col, data, fx.packed, trait_type=trait, variant_id=fx.variant_id)
        assert np.allclose(glmm.tau, g["tau"], rtol=FIT_TOL, atol=0)
        assert np.allclose(glmm.coefficients, g["coefficients"], rtol=1e-5, atol=1e-9)
        assert np.array_equal(glmm.var_ratio["id"], g["vr_id"])
        assert np.allclose(glmm.var_ratio["ratio"], g["vr_ratio"], rtol=FIT_TOL)
        assert np.allclose(glmm.fitted_values, g["fitted_values"], rtol=FIT_TOL, atol=1e-12)


def test_fit_matches_oracle_with_missing_data(gpu):
    """A fit on synthetic data with missing genotypes and N % 4 != 0 (never exercised by the reference's tests)."""
    from oracle.oracle import Oracle, default_params
    import saigegds_b200 as sg
    from saigegds_b200 import rsetup
    rng = np.random.default_rng(10)
    n, m = 603, 1500
    packed = random_packed(rng, n, m, missing=0.03, maf_lo=0.05)
    x1 = rng.standard_normal(n)
    y = (rng.random(n) < 1 / (1 + np.exp(-(-1 + 0.5 * x1)))).astype(np.float64)
    X, _ = rsetup.qr_transform(rsetup.model_matrix({"x1": x1}, ["x1"]))
    fit0 = rsetup.glm_binomial(X, y)
    o = Oracle(); o.store_2b_geno(packed, n)
    ro = o.fit_AI_PCG("binary", fit0, X, [1.0, 0.5], default_params(nrun=10, maxiter=6))
    gpu.saige_store_2b_geno(packed, n)
    rg = gpu.saige_fit_AI_PCG_binary(fit0, X, [1.0, 0.5], sg.make_param(nrun=10, maxiter=6))
    assert np.allclose(rg["tau"], ro["tau"], rtol=FIT_TOL, atol=1e-12)
    assert relinf(rg["fitted_values"], ro["fitted_values"]) < FIT_TOL
    assert relinf(rg["cov"], ro["cov"]) < FIT_TOL


# ---------------------------------------------------------------- RNG and error behaviour
def test_r_rng_matches_oracle(gpu, oracle):
    gpu.set_seed(42); oracle.set_seed(42)
    assert np.array_equal(gpu.runif(1000), oracle.unif_rand(1000))
    gpu.set_seed(200); oracle.set_seed(200)
    assert np.array_equal(gpu.sample_int(9976), oracle.sample_int(9976))


def test_error_behaviour(gpu, fx):
    import saigegds_b200 as sg
    c = sg.Context(0)
    with pytest.raises(sg.SgbError):                       # product before store
        c.n_samp = 4
        c.get_crossprod_b_grm(np.zeros(4))
    with pytest.raises(sg.InvalidArgument):                # wrong bytes-per-variant
        c.saige_store_2b_geno(np.zeros((3, 5), dtype=np.uint8), 100)
    with pytest.raises(sg.InvalidArgument):                # unknown kernel id
        from saigegds_b200 import _lib
        _lib.check(_lib.lib().sgb_set_kernel(c._h, 99))
    with pytest.raises(sg.InvalidArgument):                # fit with the wrong number of samples
        c.saige_store_2b_geno(fx.packed[:10], fx.n_samp)
        from saigegds_b200 import rsetup
        X = np.ones((5, 1))
        c.saige_fit_AI_PCG_binary(rsetup.Fit0(np.zeros(5), np.zeros(1), np.zeros(5), np.full(5, .5), "binomial"), X, [1, .5])
    c.close()


def test_multi_gpu_sharding_matches_single_gpu():
    """Variant-sharded product (single-RHS and batched) / PCG / fit / variance ratio over NCCL on all GPUs of the box (2, 4 or 8;
    skipped on a 1-GPU box).  tests/multi_gpu_check.py can also be launched by hand; its log is kept under profiles/."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(torch.cuda.device_count()), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode()[-3000:]


def test_verbose_log_has_the_reference_format(gpu, fx, setup_binary):
    """verbose=TRUE output: the text Rprintf produces in saige_fit_AI_PCG_binary (saige_fitnull.cpp:1027-1032,
    print_vec :934-945, "%0.7g") with the tau trajectory of the reference's fit of its own fixture."""
    import saigegds_b200 as sg
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    lines = []
    gpu.set_print(lines.append)
    try:
        gpu.saige_fit_AI_PCG_binary(setup_binary["fit0"], setup_binary["X"], setup_binary["tau"], sg.make_param(verbose=True))
    finally:
        gpu.set_print(None)
    text = "".join(lines)
    want_tau = ["0.4994116", "0.3287896", "0.2817812", "0.3211452", "0.3361534"]
    for it, tv in enumerate(want_tau, start=1):
        assert ("Iteration %d:\n    tau: (1, %s)\n    fixed coeff: (" % (it, tv)) in text, (it, text[:2000])
    assert "Final tau: (1, 0.3322063)\n    fixed coeff: (" in text
    import re
    assert re.search(r"Initial variance component estimates, tau:\n    Sigma_E: 1, Sigma_G: 0\.5\n", text)
    coef = re.findall(r"fixed coeff: \(([^)]*)\)", text)
    assert all(len(c.split(", ")) == 3 for c in coef)
