"""Device-resident timing of k-column GRM products (batched tcgen05 path vs the single-RHS kernels), with the per-kernel
event times of one profiled call.  Usage: python tools/multicol_timing.py [N M] [k ...]   -> one JSON line per k on stdout."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saigegds_b200 as sg

args = [int(a) for a in sys.argv[1:]]
N, M = (args[0], args[1]) if len(args) >= 2 else (430000, 100000)
ks = args[2:] if len(args) > 2 else [1, 2, 4, 8, 16, 30]
ctx = sg.Context(0)
ctx.store_synthetic(N, M, M, 0, seed=200, missing_rate=0.005)
rng = np.random.default_rng(1)
for k in ks:
    B = np.asfortranarray(rng.standard_normal((N, k)))
    d_b = ctx.device_vector(B.reshape(-1, order="F"))
    d_out = ctx.device_empty(8 * N * k)
    line = {"N": N, "M": M, "k": k}
    for name in ("auto", "imma"):                       # auto: batched from k >= 2; imma: column loop over the fused kernel
        if name == "imma" and k > 4:
            continue
        ctx.set_kernel(name)
        for _ in range(3):
            ctx.grm_mv_device(d_b, d_out, k)
        reps = 10 if k <= 8 else 5
        ms = ctx.time_products_device(d_b, d_out, k, reps) / reps
        line[name] = {"ms_per_call": ms, "ms_per_column": ms / k}
        if name == "auto":
            ctx.set_profiling(True)
            ctx.grm_mv_device(d_b, d_out, k)
            line["kernels_ms"] = {kk: round(v[0], 4) for kk, v in ctx.kernel_times().items()}
            ctx.set_profiling(False)
    ctx.set_kernel("auto")
    print(json.dumps(line), flush=True)
    d_b.free(); d_out.free()
