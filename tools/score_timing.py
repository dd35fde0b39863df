"""Timing of the score-test kernel on synthetic data of the headline shape (N = 430K samples, K = 10 covariates): variants
per second from the matrix already resident in HBM (CUDA events around the kernel) and end to end from packed host
batches.  Usage: python tools/score_timing.py [n_samp] [n_var] [K]  ->  one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saigegds_b200 as sg  # noqa: E402
from saigegds_b200 import rsetup  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 430000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
rng = np.random.default_rng(7)
X = np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(K - 1)])
beta = np.concatenate([[-2.0], rng.normal(0, 0.2, K - 1)])
mu = 1 / (1 + np.exp(-(X @ beta)))
y = (rng.random(n) < mu).astype(np.float64)
V = mu * (1 - mu)
XVX_inv = np.linalg.inv(X.T @ (X * V[:, None]))
noK = rsetup.ObjNoK(y=y, mu=mu, res=y - mu, V=V, X1=X, XV=(X * V[:, None]).T.copy(), XXVX_inv=X @ XVX_inv)
mod = sg.NullModel(coefficients=beta, tau=np.array([1.0, 0.3]), linear_predictors=X @ beta, fitted_values=mu, residuals=y - mu,
                   cov=XVX_inv, converged=True, obj_noK=noK, var_ratio={"ratio": np.array([1.0])}, trait_type="binary")
ctx = sg.Context(0)
ctx.store_synthetic(n, m)
st = sg.ScoreTest(sg.init_nullmod(mod), ctx)
st.set_path("per_variant")
st.test_stored(0, min(m, 256))                      # warm-up
ref, ms_pv = st.test_stored(0, m)
st.set_path("tiled")
st.test_stored(0, min(m, 256))
res, ms = st.test_stored(0, m)
res2, ms2 = st.test_stored(0, m)
dev = {k: float(np.nanmax(np.abs(res[k] - ref[k]) / np.maximum(np.abs(ref[k]), 1e-300))) for k in ("pval", "beta", "p.norm")}
host = ctx.synth_to_host(n, m)
t0 = time.perf_counter()
res3 = st.test(host)
e2e = time.perf_counter() - t0
assert all(np.array_equal(res[k], res3[k], equal_nan=True) for k in res)
pv = res["pval"][res["valid"]]
ctx.set_profiling(True)
st.test_stored(0, m)
kt = ctx.kernel_times()
ctx.set_profiling(False)
line = {"n_samp": n, "n_variant": m, "K": K, "kernel_ms": [ms, ms2], "per_variant_path_ms": ms_pv, "kernels": kt,
        "tiled_vs_per_variant_rel": dev, "variants_per_s_resident": m / (min(ms, ms2) * 1e-3),
        "variants_per_s_host_packed": m / e2e, "packed_GBps_resident": (n / 4) * m / (min(ms, ms2) * 1e-3) / 1e9,
        "valid": int(res["valid"].sum()), "spa_adjusted": int(np.sum(res["pval"] != res["p.norm"]) - np.sum(~res["valid"])),
        "p_below_0.05": float(np.mean(pv < 0.05)), "converged": float(np.mean(res["converged"][res["valid"]]))}
print(json.dumps(line))
