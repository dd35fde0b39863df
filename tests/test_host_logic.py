"""CPU tests of the host-side logic: R set-up restatement, C-ABI surface, sharding (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def test_rsetup_reproduces_golden_obj_noK(fx, setup_binary, setup_quant):
    for s, g in ((setup_binary, fx.model), (setup_quant, fx.model_quant)):
        noK = s["noK"]
        for k in ("y", "mu", "res", "V", "X1", "XV", "XXVX_inv"):
            assert rel(getattr(noK, k), g["noK_" + k]) < 1e-10, k


def test_initial_tau(setup_binary, setup_quant):
    assert list(setup_binary["tau"]) == [1.0, 0.5]
    t = setup_quant["tau"]
    assert t[0] == t[1] and t[0] > 0


def test_parse_formula():
    from saigegds_b200.rsetup import parse_formula
    assert parse_formula("y ~ x1 + x2") == ("y", ["x1", "x2"], True)
    assert parse_formula("y ~ x_0 + x_1 -1") == ("y", ["x_0", "x_1"], False)


def test_shard_range_covers_everything():
    from saigegds_b200 import shard_range
    for m, w in ((10, 3), (100000, 8), (7, 8), (9976, 2)):
        r = [shard_range(m, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == m
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_library_exports_every_declared_symbol():
    """The shared library must export exactly what include/saigegds_b200.h declares."""
    from saigegds_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "saigegds_b200.h")).read()
    declared = set(re.findall(r"\b(sgb_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import saigegds_b200 as sg
    with pytest.raises(sg.SgbError) as e:
        sg.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "saigegds_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src and "saige_oracle" not in src, fn


_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from saigegds_b200 import shard_range
from saigegds_b200.dist import allreduce_sum_numpy, broadcast_bytes
from oracle.oracle import Oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
d = np.load(os.path.join({root!r}, "tests", "golden", "grm1k_10k.npz"))
packed = d["packed_all"][d["keep"]][:2001]; n = int(d["n_samp"]); m = len(packed)
uid = broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128, 0)
assert uid == bytes(range(128))
a, b = shard_range(m, rank, 2)
o = Oracle(); o.store_2b_geno(packed[a:b], n)
vec = np.random.default_rng(5).standard_normal(n)
part = o.grm_mv(vec) * ((b - a) / m)          # local (1/M_local) -> contribution to (1/M_total)
full = allreduce_sum_numpy(part)
ref = Oracle(); ref.store_2b_geno(packed, n)
want = ref.grm_mv(vec)
err = np.max(np.abs(full - want)) / np.max(np.abs(want))
assert err < 1e-12, err
print("rank", rank, "ok", err)
"""


def test_sharded_product_equals_full_product_gloo(tmp_path):
    """world_size-2 gloo run: variant shards + sum all-reduce reproduce the unsharded product."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_fixture_gds_bytes_are_what_store_gds_geno_takes(fx):
    """The decompressed `genotype/data` bytes of the reference's own GDS file, read the way sgb_store_gds_geno reads them
    (one nibble per sample, allele 1 in bits 0-1, allele 2 in bits 2-3, no row padding; csrc/store.cu gds_to_dosage_kernel),
    give the committed 2-bit matrix and the golden variant filter.  Needs /root/reference (build container only)."""
    path = "/root/reference/inst/extdata/grm1k_10k_snp.gds"
    if not os.path.exists(path):
        pytest.skip("the reference tree is not present on this machine")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import read_genotypes
    _, _, raw = read_genotypes(path, want_raw=True)
    n, m = fx.n_samp, len(fx.packed_all)
    assert raw.size == (n * m + 1) // 2
    nib = np.stack([raw & 15, raw >> 4], axis=1).reshape(-1)[:n * m].reshape(m, n)
    a0, a1 = nib & 3, nib >> 2
    code = np.where((a0 == 3) | (a1 == 3), 3, (a0 != 0).astype(np.uint8) + (a1 != 0)).astype(np.uint8)
    want = np.stack([(fx.packed_all >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(m, -1)[:, :n]
    assert np.array_equal(code, want)
    valid = (a0 != 3).sum(axis=1) + (a1 != 3).sum(axis=1)
    alt = ((a0 == 1) | (a0 == 2)).sum(axis=1) + ((a1 == 1) | (a1 == 2)).sum(axis=1)
    af = alt / valid
    assert np.array_equal(np.minimum(af, 1 - af) >= 0.005, fx.keep)          # seqSetFilterCond(maf=0.005), R/saige_main.r:319


def test_oracle_synthetic_generator_is_shard_invariant_and_calibrated():
    """The CPU restatement of the device generator (bench.py's CPU arm): a shard equals the rows of the full matrix, pad codes
    are 3, ~0.5 % missing, allele frequencies spread over (0.005, 0.5)."""
    from oracle import oracle as orc
    orc.build()
    full = orc.synth_geno(1003, 64, 0, 11, 0.005, 2)
    part = orc.synth_geno(1003, 20, 30, 11, 0.005, 1)
    assert np.array_equal(full[30:50], part)
    codes = np.stack([(full >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(64, -1)
    assert np.all(codes[:, 1003:] == 3)
    assert 0.0 < np.mean(codes[:, :1003] == 3) < 0.02
    big = orc.synth_geno(20000, 200, 0, 200, 0.005, 2)
    c = np.stack([(big >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(200, -1)[:, :20000]
    af = np.where(c == 3, 0, c).sum(1) / (2.0 * (c != 3).sum(1))
    assert 0.004 < af.min() < 0.05 and 0.4 < af.max() < 0.51


def test_collinear_covariates_are_dropped_like_lm_does():
    """R/saige_main.r:362-376: columns with NA coefficients in lm(y ~ X - 1) (dependent on earlier columns) are excluded."""
    from saigegds_b200 import rsetup
    rng = np.random.default_rng(0)
    X = np.column_stack([np.ones(50), rng.standard_normal(50), rng.standard_normal(50)])
    X = np.column_stack([X, 2 * X[:, 1] - X[:, 0], rng.standard_normal(50), np.zeros(50), X[:, 2] * (1 + 1e-12)])
    assert rsetup.independent_columns(X).tolist() == [0, 1, 2, 4]
    assert rsetup.independent_columns(X[:, :3]).tolist() == [0, 1, 2]


def test_hot_kernels_hold_the_instructions_the_design_claims():
    """SASS of the in-tree library (cuobjdump, no GPU needed): tcgen05 / TMA / tensor-memory instructions in the batched kernel,
    mma.sync + TMA in the fused single-RHS kernel, mma.sync + cp.async in the small-K GEMM, system-scope loads in the peer-memory
    all-reduce -- the mnemonics profiles/r02_sass_opcounts.txt lists."""
    import shutil
    if shutil.which("cuobjdump") is None or shutil.which("cu++filt") is None:
        pytest.skip("no cuobjdump in this environment")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_opcounts
    counts, _ = sass_opcounts.kernel_counts()

    def of(fragment):
        hits = [c for name, c in counts.items() if fragment in name]
        assert hits, fragment
        return hits

    for c in of("umma_pair_kernel"):
        assert c["UTCIMMA.2CTA"] >= 1 and c["UTCBAR.2CTA.MULTICAST"] >= 1 and c["LDTM"] >= 1 and c["STTM"] >= 1 and c["UTMALDG"] >= 1
    for c in of("umma_gemm_kernel"):
        assert c["UTCIMMA"] + c["UTCIMMA.2CTA"] >= 1 and c["UTMALDG"] >= 1 and c["LDTM"] >= 1
    for c in of("imma_fused_kernel"):
        assert c["IMMA"] >= 48 and c["UTMALDG"] >= 1 and c["LDSM"] >= 12 and c["USETMAXREG"] == 2
    for c in of("imma_small_gemm_kernel"):
        assert c["IMMA"] >= 32 and c["LDGSTS"] >= 1
    for c in of("peer_allreduce_kernel"):
        assert c["LDG.SYS"] >= 1 and c["STG.SYS"] >= 1
