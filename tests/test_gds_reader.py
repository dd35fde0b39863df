"""Host half of the GDS ingestion (saigegds_b200/gds.py): the reference's own file
inst/extdata/grm1k_10k_snp.gds, read without SeqArray / gdsfmt, must give the genotypes the committed fixture holds (which
tests/golden/make_golden.py decoded independently) and, through the device half, the golden variant selection.

The file lives in the reference tree, which exists in the build container only: these tests skip elsewhere.  A synthetic
container (same framing: CoreArray magic, nodes as runs of xz streams) covers the reader on every machine.
"""
import lzma
import os

import numpy as np
import pytest

from saigegds_b200 import gds

REF_GDS = os.path.join(os.environ.get("SAIGE_REFERENCE", "/root/reference"), "inst", "extdata", "grm1k_10k_snp.gds")
needs_ref = pytest.mark.skipif(not os.path.exists(REF_GDS), reason="reference tree not present on this machine")


def bits_from_packed(packed, n):
    """2-bit dosage rows -> GDS bit2 allele pairs (dosage 1 = (1, 0), 2 = (1, 1)): the convention the fixture file uses."""
    c = np.stack([(packed >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(packed.shape[0], -1)[:, :n]
    a0 = (c >= 1).astype(np.uint8)
    a1 = (c >= 2).astype(np.uint8)
    nib = (a0 | (a1 << 2)).reshape(-1)
    return (nib[0::2] | (nib[1::2] << 4)).astype(np.uint8)


@needs_ref
def test_reads_the_reference_gds_file(fx):
    g = gds.read_gds_genotypes(REF_GDS)
    assert (g.n_sample, g.n_variant) == (1000, 10000)
    assert g.sample_id[0] == "s1" and g.sample_id[-1] == "s1000" and len(g.sample_id) == 1000
    assert np.array_equal(g.variant_id, np.arange(1, 10001))
    assert len(g.stream_offsets) == 3 and g.allele_bits.size == 5_000_000
    # allele pairs -> alt-allele dosage == the committed fixture (decoded independently by make_golden.py)
    v = np.stack([(g.allele_bits >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(10000, 1000, 2)
    dosage = (v != 0).sum(axis=2).astype(np.uint8)
    want = np.stack([(fx.packed_all >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(10000, -1)[:, :1000]
    assert np.array_equal(dosage, want)


def fake_gds(tmp_path, n_sample, n_variant, rng, string_ids=True):
    """A file with the framing the reader relies on: magic, then nodes as runs of xz streams with other bytes in between."""
    codes = rng.integers(0, 4, size=(n_variant, n_sample, 2), dtype=np.uint8)
    flat = codes.reshape(-1)
    if flat.size % 4:
        flat = np.append(flat, np.zeros(4 - flat.size % 4, dtype=np.uint8))
    q = flat.reshape(-1, 4)
    bits = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8).tobytes()
    if string_ids:
        sid = b"".join(bytes([len(s)]) + s for s in (("id%d" % i).encode() for i in range(n_sample)))
    else:
        sid = (np.arange(n_sample, dtype="<i4") + 1000).tobytes()
    vid = np.arange(1, n_variant + 1, dtype="<i4").tobytes()
    cut = max(65536, (len(bits) // 3 // 4096) * 4096)
    parts = [bits[i:i + cut] for i in range(0, len(bits), cut)]
    blob = gds.GDS_MAGIC + b"\x00\x01junk" + lzma.compress(sid) + b"\x07\x07" + lzma.compress(vid) + b"pad" + \
        lzma.compress(np.arange(n_variant, dtype="<i4").tobytes()) + b"".join(lzma.compress(p) for p in parts) + b"tail" + lzma.compress(b"\x01" * n_variant)
    path = tmp_path / "fake.gds"
    path.write_bytes(blob)
    return str(path), np.frombuffer(bits, dtype=np.uint8), len(parts)


@pytest.mark.parametrize("n_sample,n_variant", [(640, 900), (1237, 301)])
def test_reader_on_a_synthetic_container(tmp_path, n_sample, n_variant):
    rng = np.random.default_rng(n_sample)
    path, bits, nparts = fake_gds(tmp_path, n_sample, n_variant, rng)
    g = gds.read_gds_genotypes(path)
    assert (g.n_sample, g.n_variant) == (n_sample, n_variant)
    assert g.sample_id[3] == "id3" and len(g.stream_offsets) == nparts
    assert np.array_equal(g.allele_bits, bits)
    with pytest.raises(gds.GdsFormatError):
        gds.read_gds_genotypes(path, n_sample=n_sample + 1)
    bad = tmp_path / "not.gds"
    bad.write_bytes(b"hello")
    with pytest.raises(gds.GdsFormatError):
        gds.read_gds_genotypes(str(bad))


@needs_ref
@pytest.mark.gpu
def test_gds_file_to_stored_genotypes(gpu, fx, oracle):
    """File -> device store with the MAF filter of R/saige_main.r:319: the golden 9,976-variant selection, the oracle's table."""
    g, r = gds.store_from_gds(gpu, REF_GDS, maf=0.005)
    assert np.array_equal(r["variant_sel"], fx.keep) and gpu.n_var == 9976
    assert np.array_equal(r["lut"], oracle.lut)
    sub = [s for s in g.sample_id[::2]]
    g2, r2 = gds.store_from_gds(gpu, REF_GDS, sample_id=sub, maf=0.005)
    assert gpu.n_samp == 500 and np.array_equal(r2["sample_sel"], np.arange(0, 1000, 2))


def gds_from_packed(tmp_path, packed, n_sample, name="fixture.gds"):
    """The fixture's genotypes in the framing the reader relies on (sample ids s1.., variant ids 1.., bit2 allele pairs)."""
    bits = bits_from_packed(packed, n_sample).tobytes()
    sid = b"".join(bytes([len(s)]) + s for s in (("s%d" % (i + 1)).encode() for i in range(n_sample)))
    vid = np.arange(1, packed.shape[0] + 1, dtype="<i4").tobytes()
    cut = max(65536, (len(bits) // 3 // 4096) * 4096)
    blob = gds.GDS_MAGIC + b"\x00\x01hdr" + lzma.compress(sid) + b"\x07" + lzma.compress(vid) + b"pad" + \
        b"".join(lzma.compress(bits[i:i + cut]) for i in range(0, len(bits), cut)) + b"tail"
    path = tmp_path / name
    path.write_bytes(blob)
    return str(path)


@pytest.mark.gpu
def test_association_scan_straight_from_a_gds_file(gpu, fx, tmp_path):
    """seqAssocGLMM_SPA(gdsfile, modobj) (R/assoc_single.r:92-334) without SeqArray: file -> xz streams -> device dosage rows -> scan.
    The model's samples are given in another order than the file's (the reference reorders the model rows with
    `match(sid, modobj$sample.id)`, :141-145) and the golden p-values of all 10,000 variants must come out."""
    import saigegds_b200 as sg
    from test_score_test import golden_modobj, relmax
    path = gds_from_packed(tmp_path, fx.packed_all, fx.n_samp)
    perm = np.random.default_rng(5).permutation(fx.n_samp)
    mod = golden_modobj(fx, "binary")
    ids = np.array(["s%d" % (i + 1) for i in range(fx.n_samp)])
    noK = mod.obj_noK
    mod.sample_id = ids[perm]
    mod.fitted_values, mod.linear_predictors, mod.residuals = mod.fitted_values[perm], mod.linear_predictors[perm], mod.residuals[perm]
    noK.y, noK.mu, noK.res, noK.V = noK.y[perm], noK.mu[perm], noK.res[perm], noK.V[perm]
    noK.X1, noK.XXVX_inv, noK.XV = noK.X1[perm], noK.XXVX_inv[perm], noK.XV[:, perm]
    ans = sg.seqAssocGLMM_SPA(path, mod, mac=4, ctx=gpu)
    pv = fx.pval
    assert np.array_equal(ans["id"], pv["id"]) and np.array_equal(ans["mac"], pv["mac"])
    assert relmax(ans["pval"], pv["pval"]) < 1e-9 and relmax(ans["p.norm"], pv["p_norm"]) < 1e-9
    assert np.array_equal(ans["converged"].astype(int), pv["converged"])
    mod.sample_id = np.append(ids[perm][:-1], "nobody")
    with pytest.raises(ValueError, match="not available in the GDS file"):
        sg.seqAssocGLMM_SPA(path, mod, mac=4, ctx=gpu)


@pytest.mark.gpu
def test_null_model_fit_and_scan_from_a_gds_file_like_the_reference_workflow(gpu, fx, tmp_path):
    """The reference's own workflow (inst/unitTests/test_SAIGE.R: seqFitNullGLMM_SPA(y ~ x1 + x2, pheno, gdsfile) then
    seqAssocGLMM_SPA(gdsfile, glmm, mac = 4)) on a GDS-framed file of the fixture, through the Python mirror with no SeqArray:
    the phenotype table is given in another row order with two extra rows (an unknown sample, a missing covariate), the MAF filter
    runs on the device -> the golden 9,976 variants, tau, variance ratio, and the golden p-values of the scan."""
    import saigegds_b200 as sg
    from test_score_test import relmax
    path = gds_from_packed(tmp_path, fx.packed_all, fx.n_samp)
    n = fx.n_samp
    perm = np.random.default_rng(11).permutation(n)
    ids = np.array(["s%d" % (i + 1) for i in range(n)])
    data = {"sample.id": np.append(ids[perm], ["ghost", "s0"]),
            "y": np.append(fx.pheno["y"][perm], [1.0, 0.0]),
            "x1": np.append(fx.pheno["x1"][perm], [0.3, np.nan]),
            "x2": np.append(fx.pheno["x2"][perm], [1.0, 0.0])}
    glmm = sg.seqFitNullGLMM_SPA("y ~ x1 + x2", data, path, trait_type="binary", ctx=gpu)
    g = fx.model
    assert gpu.n_samp == n and gpu.n_var == 9976
    assert np.array_equal(glmm.sample_id, ids) and np.array_equal(glmm.variant_id, fx.variant_id)
    assert abs(glmm.tau[1] - g["tau"][1]) / g["tau"][1] < 1e-6
    assert np.max(np.abs(glmm.coefficients - g["coefficients"]) / np.abs(g["coefficients"])) < 1e-6
    assert np.array_equal(glmm.var_ratio["id"], g["vr_id"]) and np.allclose(glmm.var_ratio["ratio"], g["vr_ratio"], rtol=1e-6)
    ans = sg.seqAssocGLMM_SPA(path, glmm, mac=4, ctx=gpu)
    pv = fx.pval
    assert np.array_equal(ans["id"], pv["id"]) and relmax(ans["pval"], pv["pval"]) < 1e-6
    # an explicit variant list instead of the filters, and the sample column checks
    some = fx.variant_id[::7]
    glmm2 = sg.seqFitNullGLMM_SPA("y ~ x1 + x2", data, path, trait_type="binary", ctx=gpu, variant_id=some, num_marker=5)
    assert gpu.n_var == len(some) and np.array_equal(glmm2.variant_id, some)
    with pytest.raises(ValueError, match="should be unique"):
        sg.seqFitNullGLMM_SPA("y ~ x1 + x2", dict(data, **{"sample.id": np.append(ids[perm], ["s1", "s2"])}), path, ctx=gpu)
    with pytest.raises(ValueError, match="No common sample.id"):
        sg.seqFitNullGLMM_SPA("y ~ x1 + x2", dict(data, **{"sample.id": np.array(["q%d" % i for i in range(n + 2)])}), path, ctx=gpu)
