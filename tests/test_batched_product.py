"""GPU parity tests of the batched K-right-hand-side GRM product (csrc/grm_umma.cuh: tcgen05.mma kind::i8, accumulators in
tensor memory) and of the full-width single-RHS kernels against the CPU oracle.

Reference call sites of the batched product: the 30 trace solves (saige_fitnull.cpp:646-654), the (1+p) solves of get_coeff_w
(:744-752), the variance-ratio markers (:1321); each column must equal get_crossprod_b_grm (:435-536) of that column.
Tolerance (north star): <= 1e-10 relative to the output's infinity norm.
"""
import numpy as np
import pytest

from conftest import random_packed

pytestmark = pytest.mark.gpu
PROD_TOL = 1e-10


def relinf(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_batched_matches_oracle_on_fixture(gpu, fx, oracle):
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    rng = np.random.default_rng(31)
    for k in (2, 5, 30, 33):                               # 33 columns = two passes
        B = rng.standard_normal((fx.n_samp, k))
        B[:, 0] *= 1e-7                                    # every column has its own fixed-point exponent
        B[:, -1] *= 1e9
        if k == 2:
            gpu.set_kernel("umma")                         # kernel "auto" takes the batched path from three columns on
        out = gpu.get_crossprod_b_grm(B)
        gpu.set_kernel("auto")
        for c in range(k):
            assert relinf(out[:, c], oracle.grm_mv(B[:, c])) < PROD_TOL, (k, c)
    gpu.set_kernel("umma")                                 # the batched kernels for a single column
    try:
        b = rng.standard_normal(fx.n_samp)
        assert relinf(gpu.get_crossprod_b_grm(b), oracle.grm_mv(b)) < PROD_TOL
        u = rng.integers(0, 2, fx.n_samp) * 2.0 - 1        # Rademacher vector of the trace estimator
        assert relinf(gpu.get_crossprod_b_grm(u), oracle.grm_mv(u)) < PROD_TOL
        wide = rng.standard_normal(fx.n_samp) * 10.0 ** rng.uniform(-12, 12, fx.n_samp)    # 24 decades
        assert relinf(gpu.get_crossprod_b_grm(wide), oracle.grm_mv(wide)) < PROD_TOL
        assert np.all(gpu.get_crossprod_b_grm(np.zeros(fx.n_samp)) == 0)
    finally:
        gpu.set_kernel("auto")


def test_batched_columns_do_not_depend_on_the_batch(gpu, fx):
    """A column's result is bit-identical whatever else is in the batch (exact integer arithmetic, per-column exponents):
    the PCG iterates of a column cannot depend on which other columns are still active."""
    gpu.saige_store_2b_geno(fx.packed, fx.n_samp)
    rng = np.random.default_rng(32)
    B = rng.standard_normal((fx.n_samp, 30))
    full = gpu.get_crossprod_b_grm(B)
    part = gpu.get_crossprod_b_grm(B[:, 7:12])
    assert np.array_equal(full[:, 7:12], part)
    assert np.array_equal(full, gpu.get_crossprod_b_grm(B))            # and bit-reproducible
    # two to four columns: the GEMMs run on mma.sync (imma_small_gemm_kernel<NT>) and produce the same limbs as tcgen05
    for k0, k in ((0, 2), (5, 3), (20, 4)):
        assert np.array_equal(full[:, k0:k0 + k], gpu.get_crossprod_b_grm(B[:, k0:k0 + k])), k
    gpu.set_kernel("umma")
    try:
        assert np.array_equal(full[:, 3], gpu.get_crossprod_b_grm(B[:, 3]))
    finally:
        gpu.set_kernel("auto")


@pytest.mark.parametrize("n_samp,n_var,missing", [(1, 3, 0.0), (7, 5, 0.3), (1001, 257, 0.05), (4099, 130, 0.02),
                                                  (2048, 128, 0.0), (6150, 33, 0.5), (513, 1025, 0.01)])
def test_batched_ragged_shapes(gpu, n_samp, n_var, missing):
    """N % 4 != 0, M % 4 != 0 (the sample-major copy pads variants), arbitrary pad codes, monomorphic / all-missing variants."""
    from oracle.oracle import Oracle
    rng = np.random.default_rng(n_samp * 1000 + n_var + 1)
    packed = random_packed(rng, n_samp, n_var, missing)
    if n_var >= 5:
        packed[1, :] = 0x00
        packed[2, :] = 0xFF
        packed[3, :] = 0xAA
    o = Oracle()
    o.store_2b_geno(packed, n_samp)
    gpu.saige_store_2b_geno(packed, n_samp)
    B = rng.standard_normal((n_samp, 6))
    got = gpu.get_crossprod_b_grm(B)                       # six columns: tcgen05
    for c in range(6):
        want = o.grm_mv(B[:, c])
        assert np.max(np.abs(got[:, c] - want)) <= PROD_TOL * max(np.max(np.abs(want)), 1e-300) + 1e-300
    # two, three and four columns: mma.sync GEMMs, same limbs
    assert np.array_equal(gpu.get_crossprod_b_grm(B[:, :3]), got[:, :3])
    assert np.array_equal(gpu.get_crossprod_b_grm(B[:, 2:6]), got[:, 2:6])
    assert np.array_equal(gpu.get_crossprod_b_grm(B[:, 4:6]), got[:, 4:6])


def test_batched_equals_single_rhs_kernels_at_scale(gpu):
    """N = 50K, M = 4K synthetic, 1 % missing: batched (tcgen05) vs two-pass IMMA vs FP64 CUDA cores, column by column."""
    n, m = 50000, 4000
    try:
        gpu.store_synthetic(n, m, seed=7, missing_rate=0.01)
        rng = np.random.default_rng(33)
        B = rng.standard_normal((n, 6))
        got = gpu.get_crossprod_b_grm(B)
        for name in ("imma2", "simt"):
            gpu.set_kernel(name)
            ref = gpu.get_crossprod_b_grm(B)               # column loop over the single-RHS kernel
            for c in range(6):
                assert relinf(got[:, c], ref[:, c]) < 1e-11, (name, c)
    finally:
        gpu.set_kernel("auto")


def test_pcg_iterates_do_not_depend_on_the_product_kernel(gpu):
    """SURVEY.md H4 at C2 width (N = 50K): the digit-sliced products (batched tcgen05, two-pass IMMA, fused) leave the PCG
    iteration counts of every column equal to those of the FP64 CUDA-core kernel, solutions within 1e-9."""
    n, m = 50000, 8192
    try:
        gpu.store_synthetic(n, m, seed=9, missing_rate=0.005)
        rng = np.random.default_rng(34)
        w = rng.uniform(0.05, 0.25, n)
        tau = np.array([1.0, 0.4])
        B = np.column_stack([rng.standard_normal(n) for _ in range(5)] + [rng.integers(0, 2, n) * 2.0 - 1])
        res = {}
        for name in ("simt", "auto", "imma2", "imma"):
            gpu.set_kernel(name)
            res[name] = gpu.PCG_diag_sigma(w, tau, B)
        xs, its = res["simt"]
        assert its.min() >= 2
        for name in ("auto", "imma2", "imma"):
            x, it = res[name]
            assert np.array_equal(it, its), name
            assert relinf(x, xs) < 1e-9, name
    finally:
        gpu.set_kernel("auto")


@pytest.mark.parametrize("m", [4096])
def test_full_width_kernels_match_oracle(gpu, m):
    """N = 430,000 (BASELINE configs 3-5 width: the fused kernel on all its CTAs with full 3,072-sample slices, 140 arrivals per
    limb, the a-priori e bound) x M = 4,096 variants against the CPU oracle on the same bytes (synth_to_host)."""
    from oracle.oracle import Oracle, max_threads
    n = 430000
    try:
        host = gpu.synth_to_host(n, m, 0, seed=200, missing_rate=0.005)
        o = Oracle()
        o.store_2b_geno(host, n, num_thread=max_threads())
        gpu.store_synthetic(n, m, seed=200, missing_rate=0.005)
        rng = np.random.default_rng(35)
        b = rng.standard_normal(n)
        want = o.grm_mv(b)
        for name in ("imma", "imma2", "umma", "auto"):
            gpu.set_kernel(name)
            assert relinf(gpu.get_crossprod_b_grm(b), want) < PROD_TOL, name
        gpu.set_kernel("auto")
        B = np.column_stack([b, rng.standard_normal(n) * 1e-5, rng.integers(0, 2, n) * 2.0 - 1])
        out = gpu.get_crossprod_b_grm(B)                   # batched
        assert relinf(out[:, 0], want) < PROD_TOL
        for c in (1, 2):
            assert relinf(out[:, c], o.grm_mv(B[:, c])) < PROD_TOL, c
        # three columns run the integer GEMMs on mma.sync (imma_small_gemm_kernel), six on tcgen05: both against the oracle at this
        # width, and the shared columns bit for bit (same limbs)
        B6 = np.column_stack([B, rng.standard_normal((n, 3))])
        out6 = gpu.get_crossprod_b_grm(B6)
        assert np.array_equal(out6[:, :3], out)
        assert relinf(out6[:, 0], want) < PROD_TOL
        assert relinf(out6[:, 5], o.grm_mv(B6[:, 5])) < PROD_TOL
        assert np.array_equal(gpu.get_crossprod_b_grm(B6[:, 4:6]), out6[:, 4:6])      # two columns: two n-tiles
    finally:
        gpu.set_kernel("auto")
        gpu.store_synthetic(1024, 64)


def test_full_size_batched_product(gpu):
    """BASELINE config 3 shape (N = 430K, M = 100K) with K = 30 columns: the batched product against the fused single-RHS
    kernel on sampled columns, bit reproducibility, linearity across columns."""
    n, m, k = 430000, 100000, 30
    try:
        gpu.store_synthetic(n, m, seed=200, missing_rate=0.005)
        rng = np.random.default_rng(36)
        B = rng.standard_normal((n, k))
        B[:, 2] = 2.5 * B[:, 0] - 0.5 * B[:, 1]
        out = gpu.get_crossprod_b_grm(B)
        assert relinf(out[:, 2], 2.5 * out[:, 0] - 0.5 * out[:, 1]) < 1e-11
        assert abs(B[:, 1] @ out[:, 0] - B[:, 0] @ out[:, 1]) / abs(B[:, 1] @ out[:, 0]) < 1e-10      # symmetry
        gpu.set_kernel("imma")
        for c in (0, 17, 29):
            assert relinf(out[:, c], gpu.get_crossprod_b_grm(B[:, c])) < 1e-11, c
        gpu.set_kernel("auto")
        assert np.array_equal(out, gpu.get_crossprod_b_grm(B))
    finally:
        gpu.set_kernel("auto")
        gpu.store_synthetic(1024, 64)


def test_synthetic_generator_matches_the_oracle_generator(gpu):
    """bench.py feeds the CPU arm with the oracle's restatement of the device generator: the bytes must be identical."""
    from oracle import oracle as orc
    for n, m, off, seed in ((1003, 64, 0, 11), (4099, 130, 977, 200), (50000, 96, 12345, 200)):
        dev = gpu.synth_to_host(n, m, off, seed=seed, missing_rate=0.005)
        cpu = orc.synth_geno(n, m, off, seed, 0.005)
        assert np.array_equal(dev, cpu), (n, m, off, seed)
