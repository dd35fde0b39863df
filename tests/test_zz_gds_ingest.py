"""Genotype ingestion from the bytes of a GDS `genotype/data` node (device half): bit2 allele pairs ->
2-bit dosage rows, integer allele counts, MAF / missing-rate filter (R/saige_main.r:319, :420) on the GPU.

The fixture's own GDS genotypes are rebuilt as allele pairs from the committed 2-bit matrix (tests/golden/make_golden.py
decoded them from inst/extdata/grm1k_10k_snp.gds the other way round), so the check is: same kept variants as the golden
model (9,976 with MAF >= 0.005), same bytes, same look-up table as the oracle.
"""
import numpy as np
import pytest

import saigegds_b200 as sg


def to_allele_bits(dosage, rng, partial_missing=False):
    """dosage [M][n] in {0,1,2,3} -> GDS bit2 stream of allele pairs (nibble per sample, no row padding) + the pairs."""
    m, n = dosage.shape
    a0 = np.zeros((m, n), dtype=np.uint8)
    a1 = np.zeros((m, n), dtype=np.uint8)
    het = dosage == 1
    first = rng.random((m, n)) < 0.5
    a0[het & first] = 1
    a1[het & ~first] = 1
    hom = dosage == 2
    a0[hom] = 1
    a1[hom] = np.where(rng.random((m, n)) < 0.1, 2, 1)[hom]            # a second alternative allele now and then
    mis = dosage == 3
    a0[mis] = 3
    a1[mis] = 3
    if partial_missing:                                                # one allele called, the other not: dosage missing
        half = mis & (rng.random((m, n)) < 0.5)
        a1[half] = rng.integers(0, 2, size=(m, n), dtype=np.uint8)[half]
    nib = (a0 | (a1 << 2)).reshape(-1)
    if nib.size % 2:
        nib = np.append(nib, np.uint8(0))
    return (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8), a0, a1


def unpack(packed, n):
    c = np.stack([(packed >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(packed.shape[0], -1)
    return c[:, :n]


@pytest.mark.gpu
def test_gds_ingest_reproduces_the_fixture_store(gpu, oracle, fx):
    rng = np.random.default_rng(12)
    dosage = unpack(fx.packed_all, fx.n_samp)
    bits, _, _ = to_allele_bits(dosage, rng)
    r = gpu.store_gds_geno(bits, fx.n_samp, len(dosage), maf=0.005)
    assert np.array_equal(r["variant_sel"], fx.keep) and gpu.n_var == 9976          # R/saige_main.r:319 on the fixture
    assert np.all(r["n_valid_alleles"] == 2 * fx.n_samp)
    assert np.array_equal(r["n_alt_alleles"], dosage.astype(np.int64).sum(axis=1))
    assert np.array_equal(r["lut"], oracle.lut)                                     # same bytes stored -> same table
    assert np.max(np.abs(r["diag"] - oracle.diag)) < 1e-12
    for j in (0, 4999, 9975):
        assert np.array_equal(gpu.get_geno_ds(j), oracle.get_geno_ds(j), equal_nan=True)
    b = rng.standard_normal(fx.n_samp)
    got, want = gpu.get_crossprod_b_grm(b), oracle.grm_mv(b)
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-10


@pytest.mark.gpu
def test_gds_ingest_ragged_subset_missing(gpu):
    """Odd sample count in the file (variants start mid-byte), a sample subset, missing and half-missing genotypes."""
    rng = np.random.default_rng(13)
    n_file, m = 1237, 301
    maf = rng.uniform(0.001, 0.5, m)
    dosage = rng.binomial(2, maf[:, None], size=(m, n_file)).astype(np.uint8)
    dosage[rng.random(dosage.shape) < rng.choice([0.0, 0.02, 0.2], size=m)[:, None]] = 3
    bits, a0, a1 = to_allele_bits(dosage, rng, partial_missing=True)
    sel = np.sort(rng.choice(n_file, size=1001, replace=False)).astype(np.int32)
    r = gpu.store_gds_geno(bits, n_file, m, sample_sel=sel, maf=0.01, missing_rate=0.1)
    s0, s1 = a0[:, sel], a1[:, sel]
    valid = (s0 != 3).sum(axis=1) + (s1 != 3).sum(axis=1)
    alt = ((s0 == 1) | (s0 == 2)).sum(axis=1) + ((s1 == 1) | (s1 == 2)).sum(axis=1)
    assert np.array_equal(r["n_valid_alleles"], valid) and np.array_equal(r["n_alt_alleles"], alt)
    with np.errstate(invalid="ignore", divide="ignore"):
        af = alt / valid
    keep = (np.minimum(af, 1 - af) >= 0.01) & (1 - valid / (2.0 * len(sel)) <= 0.1)
    assert np.array_equal(r["variant_sel"], keep) and 0 < keep.sum() < m
    want = np.where((s0 == 3) | (s1 == 3), np.nan, (s0 != 0).astype(np.float64) + (s1 != 0))[keep]
    for j in (0, 1, int(keep.sum()) - 1):
        assert np.array_equal(gpu.get_geno_ds(j), want[j], equal_nan=True)
    with pytest.raises(sg.InvalidArgument):
        gpu.store_gds_geno(bits, n_file, m, sample_sel=np.array([0, n_file], dtype=np.int32))
    with pytest.raises(sg.InvalidArgument):
        gpu.store_gds_geno(bits[:100], n_file, m)
