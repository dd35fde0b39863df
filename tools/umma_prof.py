"""Cycle breakdown of the batched GEMM kernel's issuer thread and expander warps (SGB_UMMA_PROF=1 build path)."""
import os, sys
import numpy as np
os.environ["SGB_UMMA_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saigegds_b200 as sg
N, M = 430000, 100000
ctx = sg.Context(0)
ctx.store_synthetic(N, M, M, 0, seed=200, missing_rate=0.005)
rng = np.random.default_rng(1)
for k in (2, 30):
    B = np.asfortranarray(rng.standard_normal((N, k)))
    d_b = ctx.device_vector(B.reshape(-1, order="F"))
    d_out = ctx.device_empty(8 * N * k)
    print("k =", k, flush=True)
    for _ in range(2):
        ctx.grm_mv_device(d_b, d_out, k)
    d_b.free(); d_out.free()
