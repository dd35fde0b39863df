"""Sparse-genotype entry points: saige_get_sparse / saige_store_sp_geno, the reference's default path.

CPU part: the oracle's restatement of the sparse store and product (saige_fitnull.cpp:233-388, 445-476) is pinned on the
golden model -- the reference's own fixtures were produced through this path (geno.sparse=TRUE is the default,
R/saige_main.r:228) -- and the library's host-side packing is bit-exact against it.  GPU part: the stored state equals the
oracle's sparse state.
"""
import numpy as np
import pytest

import saigegds_b200 as sg
from conftest import random_packed


def unpack(packed, n):
    c = np.stack([(packed >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(packed.shape[0], -1)
    return np.ascontiguousarray(c[:, :n].astype(np.uint8))


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def sp_fixture(fx):
    from oracle import oracle as orc
    codes = unpack(fx.packed, fx.n_samp)
    return [orc.get_sparse(c) for c in codes]


@pytest.fixture(scope="module")
def oracle_sp(fx, sp_fixture):
    from oracle.oracle import Oracle, build
    build()
    o = Oracle()
    o.lut, o.diag = o.store_sp_geno(sp_fixture, fx.n_samp, num_thread=1)
    return o


# ------------------------------------------------------------------ CPU: oracle pinned, host packing bit-exact
def test_oracle_sparse_store_equals_dense_store(oracle, oracle_sp, fx):
    nv, sm = oracle_sp.allele_counts()
    nvd, smd = oracle.allele_counts()
    flipped = smd > nvd
    assert np.array_equal(nv, nvd) and np.array_equal(sm, np.where(flipped, 2 * nvd - smd, smd))
    # sparse table: entry 0 as the dense one, entries 1..3 relative to it (:358)
    keep = ~flipped
    assert np.array_equal(oracle_sp.lut[keep, 0], oracle.lut[keep, 0])
    for k in (1, 2, 3):
        assert np.array_equal(oracle_sp.lut[keep, k], oracle.lut[keep, k] - oracle.lut[keep, 0])
    assert rel(oracle_sp.diag, oracle.diag) < 1e-12
    b = np.random.default_rng(3).standard_normal(fx.n_samp)
    assert rel(oracle_sp.grm_mv(b), oracle.grm_mv(b)) < 1e-12        # the GRM does not depend on the coded allele


def test_oracle_sparse_fit_matches_golden(oracle_sp, setup_binary, fx):
    s = setup_binary
    n0 = oracle_sp.num_products
    r = oracle_sp.fit_AI_PCG("binary", s["fit0"], s["X"], s["tau"])
    assert abs(r["tau"][1] - 0.33220628660290813) / 0.33220628660290813 < 1e-10
    assert rel(r["fitted_values"], fx.model["fitted_values"]) < 1e-10
    assert oracle_sp.num_products - n0 == 898


def test_oracle_sparse_var_ratio_matches_golden(oracle_sp, setup_binary, fx):
    oracle_sp.set_seed(200)
    ml = oracle_sp.sample_int(len(fx.packed))
    vr = oracle_sp.calc_var_ratio("binary", setup_binary["fit0"], fx.model["tau"], setup_binary["noK"], ml)
    order = np.argsort(vr["id"])
    assert np.array_equal(fx.variant_id[vr["id"][order] - 1], fx.model["vr_id"])
    for k in ("maf", "mac", "var1", "var2", "ratio"):
        assert rel(vr[k][order], fx.model["vr_" + k]) < 1e-10, k


@pytest.mark.parametrize("dtype", [np.uint8, np.int32, np.float64])
def test_get_sparse_matches_oracle(dtype):
    """saige_get_sparse on the three SEXP types, with missing values, flips and out-of-range codes."""
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    for n, maf in ((1, 0.3), (7, 0.9), (1000, 0.02), (1001, 0.8), (4099, 0.5)):
        g = rng.binomial(2, maf, size=n).astype(np.float64)
        g[rng.random(n) < 0.05] = np.nan
        if dtype == np.float64:
            x = g + rng.uniform(-0.3, 0.3, size=n)          # dosages round to the nearest code
            x[rng.random(n) < 0.02] = 7.5                   # out of range -> missing
        elif dtype == np.int32:
            x = np.where(np.isnan(g), np.iinfo(np.int32).min, g).astype(np.int32)     # NA_integer_
            x[rng.random(n) < 0.02] = -1
        else:
            x = np.where(np.isnan(g), 3, g).astype(np.uint8)
            x[rng.random(n) < 0.02] = 255
        got = sg.saige_get_sparse(x)
        want = orc.get_sparse(x)
        assert got.dtype == np.int32 and np.array_equal(got, want)
        assert len(got) == 3 + got[:3].sum()


def test_get_sparse_errors():
    with pytest.raises(sg.InvalidArgument, match="Invalid data type"):
        sg.saige_get_sparse(np.zeros(4, dtype=np.float32))
    with pytest.raises(sg.InvalidArgument, match="No enough genotypes"):
        sg.saige_get_sparse(np.zeros(4, dtype=np.uint8), 5)
    assert list(sg.saige_get_sparse(np.zeros(0, dtype=np.uint8))) == [0, 0, 0]


def test_sparse_to_packed_is_bit_exact(fx, sp_fixture):
    """Packing the index lists gives back the fixture's own 2-bit matrix (with the minor allele counted)."""
    got = sg.sparse_to_packed(sp_fixture, fx.n_samp)
    codes = unpack(fx.packed, fx.n_samp)
    flip = codes.astype(np.int64).sum(axis=1) > fx.n_samp          # no missing calls in the fixture
    want = np.where(flip[:, None], 2 - codes, codes).astype(np.uint8)
    assert np.array_equal(unpack(got, fx.n_samp), want)
    # ragged sample counts: pad samples of the last byte must read as missing
    rng = np.random.default_rng(5)
    for n in (1, 5, 6, 7, 1023):
        p = random_packed(rng, n, 37, missing=0.1)
        c = unpack(p, n)
        sp = [sg.saige_get_sparse(r) for r in c]
        q = sg.sparse_to_packed(sp, n)
        full = np.stack([(q >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(len(q), -1)
        assert np.all(full[:, n:] == 3)
        nv = (c < 3).sum(axis=1)
        sm = np.where(c < 3, c, 0).sum(axis=1)
        want = np.where((sm > nv)[:, None] & (c < 3), 2 - c, c)
        assert np.array_equal(full[:, :n], want)


def test_sparse_to_packed_rejects_malformed_vectors():
    good = np.array([1, 0, 1, 2, 3], dtype=np.int32)
    assert sg.sparse_to_packed([good], 4).tolist() == [[0b11010000]]
    for bad in (np.array([1, 0, 1, 2], dtype=np.int32),            # counts exceed the length
                np.array([1, 0, 0, 4], dtype=np.int32),            # index == n_samp
                np.array([0, 1, 0, -1], dtype=np.int32),
                np.array([0, 0], dtype=np.int32)):
        with pytest.raises(sg.InvalidArgument):
            sg.sparse_to_packed([good, bad], 4)


# ------------------------------------------------------------------ GPU: the stored state equals the oracle's sparse state
@pytest.mark.gpu
def test_store_sp_geno_matches_oracle(gpu, oracle_sp, sp_fixture, fx):
    lut, diag = gpu.saige_store_sp_geno(sp_fixture, fx.n_samp)
    assert np.array_equal(lut, oracle_sp.lut)                      # r_buf_geno, sparse layout, bit-exact
    assert rel(diag, oracle_sp.diag) < 1e-12
    nv, sm = gpu.allele_counts()
    onv, osm = oracle_sp.allele_counts()
    assert np.array_equal(nv, onv) and np.array_equal(sm, osm)
    for j in (0, 17, 9975):
        assert np.array_equal(gpu.get_geno_ds(j), oracle_sp.get_geno_ds(j), equal_nan=True)
    b = np.random.default_rng(9).standard_normal(fx.n_samp)
    assert rel(gpu.get_crossprod_b_grm(b), oracle_sp.grm_mv(b)) < 1e-10


@pytest.mark.gpu
def test_store_sp_geno_ragged_with_missing(gpu):
    """N not a multiple of 4, 3 % missing calls, several host slabs' worth of flips."""
    from oracle import oracle as orc
    rng = np.random.default_rng(21)
    n, m = 1237, 1500
    codes = unpack(random_packed(rng, n, m, missing=0.03), n)
    codes[::3] = np.where(codes[::3] < 3, 2 - codes[::3], 3)       # every third variant: coded allele is the major one
    sp = [sg.saige_get_sparse(c) for c in codes]
    o = orc.Oracle()
    olut, odiag = o.store_sp_geno(sp, n)
    lut, diag = gpu.saige_store_sp_geno(sp, n)
    assert np.array_equal(lut, olut)
    assert rel(diag, odiag) < 1e-12
    b = rng.standard_normal(n)
    assert rel(gpu.get_crossprod_b_grm(b), o.grm_mv(b)) < 1e-10
    for j in (0, 3, m - 1):
        assert np.array_equal(gpu.get_geno_ds(j), o.get_geno_ds(j), equal_nan=True)


@pytest.mark.gpu
def test_null_model_through_the_sparse_entry_matches_golden(gpu, sp_fixture, fx):
    """seqFitNullGLMM_SPA with geno.sparse=TRUE (the reference's default call, test_SAIGE.R:69): golden tau and ratios."""
    mod = sg.seqFitNullGLMM_SPA("y ~ x1 + x2", fx.pheno, sp_fixture, trait_type="binary", variant_id=fx.variant_id,
                                ctx=gpu)
    g = fx.model
    assert abs(mod.tau[1] - g["tau"][1]) / g["tau"][1] < 1e-6
    assert rel(mod.coefficients, g["coefficients"]) < 1e-6
    assert np.array_equal(mod.var_ratio["id"], g["vr_id"])
    assert rel(mod.var_ratio["ratio"], g["vr_ratio"]) < 1e-6
