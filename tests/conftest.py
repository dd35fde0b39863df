import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Fixture:
    """The reference's own test inputs (inst/extdata/grm1k_10k_snp.gds + pheno.txt.gz) and goldens."""

    def __init__(self):
        d = np.load(os.path.join(GOLDEN, "grm1k_10k.npz"))
        self.n_samp = int(d["n_samp"])
        self.packed_all = d["packed_all"]
        self.keep = d["keep"]
        self.packed = np.ascontiguousarray(self.packed_all[self.keep])   # 9,976 variants with MAF >= 0.005
        self.variant_id = d["variant_id"]
        self.af_alt_all = d["af_alt_all"]
        self.pheno = {k: d[k] for k in ("y", "yy", "x1", "x2")}
        self.model = dict(np.load(os.path.join(GOLDEN, "saige_model.npz")))
        self.model_quant = dict(np.load(os.path.join(GOLDEN, "saige_model_quant.npz")))
        self.pval = dict(np.load(os.path.join(GOLDEN, "saige_pval.npz")))
        self.pval_quant = dict(np.load(os.path.join(GOLDEN, "saige_pval_quant.npz")))


@pytest.fixture(scope="session")
def fx():
    return Fixture()


@pytest.fixture(scope="session")
def setup_binary(fx):
    """R-side set-up of the binary fit (R/saige_main.r:356-387, 480-497)."""
    from saigegds_b200 import rsetup
    X0 = rsetup.model_matrix(fx.pheno, ["x1", "x2"])
    X, R = rsetup.qr_transform(X0)
    fit0 = rsetup.glm_binomial(X, fx.pheno["y"])
    return dict(X=X, R=R, fit0=fit0, noK=rsetup.null_model_binary(X, fit0), tau=rsetup.initial_tau_binary())


@pytest.fixture(scope="session")
def setup_quant(fx):
    from saigegds_b200 import rsetup
    X0 = rsetup.model_matrix(fx.pheno, ["x1", "x2"])
    X, R = rsetup.qr_transform(X0)
    f = rsetup.glm_gaussian(X, fx.pheno["yy"])
    y = rsetup.rank_norm(f.residuals) * rsetup.sd(f.residuals)
    fit0 = rsetup.glm_gaussian(X, y)
    return dict(X=X, R=R, fit0=fit0, noK=rsetup.null_model_quant(X, fit0), tau=rsetup.initial_tau_quant(fit0))


@pytest.fixture(scope="session")
def oracle(fx):
    from oracle.oracle import Oracle, build
    build()
    o = Oracle()
    o.lut, o.diag = o.store_2b_geno(fx.packed, fx.n_samp, num_thread=1)
    return o


@pytest.fixture(scope="session")
def gpu():
    """One library context on cuda:0; fails (not skips) if the CUDA extension or device is missing."""
    import saigegds_b200 as sg
    return sg.Context(0)


def dosage_all(fx):
    """Alt-allele dosages [10000][n] of every variant of the fixture (NaN = missing), as seqAssocGLMM_SPA reads them."""
    p = fx.packed_all
    d = np.stack([(p >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(p.shape[0], -1)[:, :fx.n_samp].astype(np.float64)
    d[d == 3] = np.nan
    return d


def random_packed(rng, n_samp, n_var, missing=0.02, maf_lo=0.01):
    """Random 2-bit packed matrix [n_var][ceil(n/4)] with missing codes; pad bits random (raw, unsanitised)."""
    maf = rng.uniform(maf_lo, 0.5, size=n_var)
    g = rng.binomial(2, maf[:, None], size=(n_var, n_samp)).astype(np.uint8)
    g[rng.random((n_var, n_samp)) < missing] = 3
    nb = (n_samp + 3) // 4
    full = rng.integers(0, 4, size=(n_var, nb * 4), dtype=np.uint8)   # pad codes deliberately arbitrary
    full[:, :n_samp] = g
    q = full.reshape(n_var, nb, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)
