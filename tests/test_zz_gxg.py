"""saige_GxG_snp_bin (src/saige_fitnull.cpp:1480-1558), the native routine under seqGLMM_GxG_spa.

The reference ships no golden fixture for it.  CPU: the oracle's restatement is tied to what IS pinned -- for an interaction
term equal to a variance-ratio marker its var1 must equal the golden var1 x MAC, and its full saddle-point p-value must
agree with the partially-normal one of the score test when every sample is kept.  GPU: the library against the oracle.
"""
import numpy as np
import pytest


def marker_term(fx, oracle, k):
    j = int(np.where(fx.variant_id == fx.model["vr_id"][k])[0][0])
    ds = oracle.get_geno_ds(j)
    ac = np.nansum(ds)
    if ac / (2 * len(ds)) > 0.5:
        ds, ac = 2 - ds, 2 * len(ds) - ac
    return ds, ac


def test_oracle_gxg_is_tied_to_the_golden_variance_ratio(oracle, fx, setup_binary):
    s, g = setup_binary, fx.model
    for k in (0, 7, 29):
        ds, ac = marker_term(fx, oracle, k)
        r = oracle.GxG_snp_bin(s["fit0"], g["tau"], ds, s["noK"])
        noK, y, mu = s["noK"], s["fit0"].y, s["fit0"].fitted_values
        G = ds - noK.XXVX_inv @ (noK.XV @ ds)
        S = np.sum((y - mu) * G)
        assert ac == g["vr_mac"][k] and r["n_nonzero"] == int(np.sum(ds != 0))
        assert abs(S / r["beta"] / ac - g["vr_var1"][k]) < 1e-10 * g["vr_var1"][k]       # var1 of :1532 == golden var1 * MAC
        var2 = np.sum(mu * (1 - mu) * G * G)
        z2 = (S / np.sqrt(S / r["beta"])) ** 2                                           # Tstat^2 / var1
        from math import erfc, sqrt
        assert abs(r["p_norm"] - erfc(sqrt(z2 / 2))) < 1e-12 and var2 > 0
        assert r["converged"] and r["tau_G"] == g["tau"][1] and 0 < r["pval"] <= 1


def test_oracle_saddle_prob_full_equals_fast_with_all_samples(fx):
    """Saddle_Prob (SPATest.cpp:232-296) == Saddle_Prob_Fast (:298-374) when no sample is summarised by the normal part."""
    import ctypes as C
    from oracle import oracle as orc
    lib = orc.lib()
    lib.orc_saddle_prob.restype = C.c_double
    rng = np.random.default_rng(4)
    n = 500
    mu = rng.uniform(0.02, 0.6, n)
    g = rng.normal(0, 0.3, n) * (rng.random(n) < 0.4)
    m1, var1 = float(np.sum(mu * g)), float(np.sum(mu * (1 - mu) * g * g))
    for z in (0.5, 2.5, 4.0, -3.0, 7.0):
        q = m1 + z * np.sqrt(var1)
        conv, pna = C.c_int(0), C.c_double(0)
        p = lib.orc_saddle_prob(C.c_double(q), C.c_double(m1), C.c_double(var1), C.c_long(n), orc._p(mu), orc._p(g),
                                C.c_double(2.0), C.byref(conv), C.byref(pna))
        assert 0 < p <= 1 and conv.value == 1
        if abs(z) < 2:
            assert p == pna.value
        else:
            assert p != pna.value and abs(np.log(p / pna.value)) < 3


@pytest.mark.gpu
def test_gpu_gxg_matches_oracle(gpu, oracle, fx, setup_binary):
    s, g = setup_binary, fx.model
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    glmm = {"tau": g["tau"]}
    rng = np.random.default_rng(8)
    terms = [marker_term(fx, oracle, 3)[0]]
    a, b = marker_term(fx, oracle, 5)[0], marker_term(fx, oracle, 11)[0]
    terms.append(a * b + (rng.random(fx.n_samp) < 0.03))                # a sparse SNP x SNP product
    terms.append(fx.pheno["x1"] * marker_term(fx, oracle, 20)[0])       # SNP x covariate, real-valued
    y = s["fit0"].y
    terms.append(np.where(y > 0, 1.0, 0.0) * (rng.random(fx.n_samp) < 0.2))   # strongly associated: saddle-point branch
    for t in terms:
        want = oracle.GxG_snp_bin(s["fit0"], g["tau"], t, s["noK"])
        got = gpu.saige_GxG_snp_bin(s["fit0"], glmm, t, s["noK"])
        assert got["n_nonzero"] == want["n_nonzero"] and got["converged"] == want["converged"]
        assert got["tau_G"] == want["tau_G"]
        for k, ko in (("beta", "beta"), ("SE", "SE"), ("pval", "pval"), ("p.norm", "p_norm")):
            assert abs(got[k] - want[ko]) <= 1e-6 * abs(want[ko]), (k, got[k], want[ko])
    assert want["pval"] != want["p_norm"]                                # the last term went through the saddle point
    import saigegds_b200 as sg
    with pytest.raises(sg.InvalidArgument):
        gpu.saige_GxG_snp_bin(s["fit0"], glmm, terms[0][:-1], s["noK"])
