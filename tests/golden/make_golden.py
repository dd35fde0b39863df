#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference's own test data.

Run in the build container only (it reads /root/reference, which does not exist on the
GPU box):  python tests/golden/make_golden.py

Outputs (committed, small):
  tests/golden/grm1k_10k.npz        packed 2-bit genotypes (all 10,000 variants; `keep` marks the
                                    9,976 with MAF >= 0.005) of inst/extdata/grm1k_10k_snp.gds in the
                                    layout consumed by saige_store_2b_geno
                                    (src/saige_fitnull.cpp:159-230, 410-425) + pheno.txt.gz columns
  tests/golden/saige_model.npz      inst/unitTests/saige_model.rds        (binary null model)
  tests/golden/saige_model_quant.npz inst/unitTests/saige_model_quant.rds (quantitative null model)
  tests/golden/saige_pval.npz / saige_pval_quant.npz   numeric columns of the p-value goldens

The GDS container is read without gdsfmt: the genotype node is a bit2 array
[variant=10000][sample=1000][ploidy=2] stored as consecutive xz streams.
"""
import gzip
import lzma
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from rds_reader import read_rds  # noqa: E402

REF = os.environ.get("SAIGE_REFERENCE", "/root/reference")
XZ_MAGIC = b"\xfd7zXZ\x00"


def xz_streams(buf):
    """Yield (offset, decompressed bytes) of every xz stream in a byte string."""
    pos = 0
    while True:
        pos = buf.find(XZ_MAGIC, pos)
        if pos < 0:
            return
        d = lzma.LZMADecompressor()
        try:
            out = d.decompress(buf[pos:])
            if d.eof:
                yield pos, out
        except lzma.LZMAError:
            pass
        pos += 1


def read_genotypes(path, n_var=10000, n_samp=1000, want_raw=False):
    buf = open(path, "rb").read()
    streams = list(xz_streams(buf))
    want = n_var * n_samp * 2 // 4
    # genotype/data = the run of consecutive streams whose sizes add up to `want`
    for i in range(len(streams)):
        tot, j = 0, i
        while j < len(streams) and tot < want:
            tot += len(streams[j][1])
            j += 1
        if tot == want and j - i >= 1 and len(streams[i][1]) >= 65536:
            raw = b"".join(s[1] for s in streams[i:j])
            offs = [s[0] for s in streams[i:j]]
            break
    else:
        raise RuntimeError("genotype streams not found")
    b = np.frombuffer(raw, dtype=np.uint8)
    vals = np.stack([(b >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)
    alleles = vals.reshape(n_var, n_samp, 2)
    if want_raw:
        return alleles, offs, b
    return alleles, offs


def pack_2bit(dosage):
    """dosage: [M][N] uint8 in {0,1,2,3}; returns [M][ceil(N/4)] bytes, sample 4j+k in bits 2k..2k+1."""
    m, n = dosage.shape
    nb = (n + 3) // 4
    pad = np.full((m, nb * 4), 3, dtype=np.uint8)   # padding = missing code
    pad[:, :n] = dosage
    q = pad.reshape(m, nb, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)


def model_to_dict(m):
    out = {}
    for k in ("coefficients", "tau", "linear.predictors", "fitted.values", "residuals"):
        out[k.replace(".", "_")] = np.asarray(m[k].value, dtype=np.float64)
    out["cov"] = m["cov"].as_array()
    out["converged"] = np.asarray(m["converged"].value)
    out["variant_id"] = np.asarray(m["variant.id"].value, dtype=np.int32)
    nok = m["obj.noK"]
    for k in ("y", "mu", "res", "V"):
        out["noK_" + k] = np.asarray(nok[k].value, dtype=np.float64)
    for k in ("X1", "XV", "XXVX_inv"):
        out["noK_" + k] = nok[k].as_array()
    vr = m["var.ratio"]
    for k in vr.keys():
        out["vr_" + k] = np.asarray(vr[k].value)
    return out


def pval_to_dict(p):
    out = {}
    for k in p.keys():
        v = p[k].value
        if isinstance(v, np.ndarray):
            out[k.replace(".", "_")] = v
    return out


def main():
    ext = os.path.join(REF, "inst", "extdata")
    ut = os.path.join(REF, "inst", "unitTests")
    alleles, offs = read_genotypes(os.path.join(ext, "grm1k_10k_snp.gds"))
    print("genotype xz streams at file offsets", offs)
    assert alleles.max() <= 1, "fixture has no missing / multi-allelic calls"
    dosage = (alleles != 0).sum(axis=2).astype(np.uint8)        # alt-allele dosage, [10000][1000]
    af = dosage.sum(axis=1) / (2.0 * dosage.shape[1])
    keep = np.minimum(af, 1 - af) >= 0.005                       # R/saige_main.r:319 (maf=0.005)
    variant_id = (np.nonzero(keep)[0] + 1).astype(np.int32)

    with gzip.open(os.path.join(ext, "pheno.txt.gz"), "rt") as f:
        hdr = f.readline().split()
        rows = [ln.split() for ln in f if ln.strip()]
    cols = {h: [r[i] for r in rows] for i, h in enumerate(hdr)}

    pv = read_rds(os.path.join(ut, "saige_pval.rds"))
    af_alt = np.asarray(pv["AF.alt"].value)
    assert np.max(np.abs(af_alt - af)) == 0.0, "decode disagrees with golden AF.alt"
    mod = read_rds(os.path.join(ut, "saige_model.rds"))
    assert np.array_equal(np.asarray(mod["variant.id"].value), variant_id)

    np.savez_compressed(
        os.path.join(HERE, "grm1k_10k.npz"),
        packed_all=pack_2bit(dosage), n_samp=np.int64(dosage.shape[1]), variant_id=variant_id,
        af_alt_all=af, keep=keep,
        sample_id=np.array(cols["sample.id"]),
        y=np.array(cols["y"], dtype=np.float64), yy=np.array(cols["yy"], dtype=np.float64),
        x1=np.array(cols["x1"], dtype=np.float64), x2=np.array(cols["x2"], dtype=np.float64))
    np.savez_compressed(os.path.join(HERE, "saige_model.npz"), **model_to_dict(mod))
    np.savez_compressed(os.path.join(HERE, "saige_model_quant.npz"),
                        **model_to_dict(read_rds(os.path.join(ut, "saige_model_quant.rds"))))
    np.savez_compressed(os.path.join(HERE, "saige_pval.npz"), **pval_to_dict(pv))
    np.savez_compressed(os.path.join(HERE, "saige_pval_quant.npz"),
                        **pval_to_dict(read_rds(os.path.join(ut, "saige_pval_quant.rds"))))
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print("%-28s %8d bytes" % (fn, os.path.getsize(os.path.join(HERE, fn))))


if __name__ == "__main__":
    main()
