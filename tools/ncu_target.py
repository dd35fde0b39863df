"""Short target for ncu: a few GRM products at the BASELINE shape.  python tools/ncu_target.py [single|batched] [M] [K]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saigegds_b200 as sg

mode = sys.argv[1] if len(sys.argv) > 1 else "batched"
N, M = 430000, int(sys.argv[2]) if len(sys.argv) > 2 else 100000
K = int(sys.argv[3]) if len(sys.argv) > 3 else (30 if mode == "batched" else 1)
ctx = sg.Context(0)
ctx.store_synthetic(N, M, M, 0, seed=200, missing_rate=0.005)
B = np.asfortranarray(np.random.default_rng(1).standard_normal((N, K)))
d_b = ctx.device_vector(B.reshape(-1, order="F"))
d_out = ctx.device_empty(8 * N * K)
for _ in range(3):
    ctx.grm_mv_device(d_b, d_out, K)
ctx.synchronize() if hasattr(ctx, "synchronize") else None
print("done", mode, N, M, K)
