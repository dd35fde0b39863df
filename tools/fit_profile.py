"""Per-kernel event times of one null-model fit (profiling mode serialises the stream: shares, not absolutes)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import saigegds_b200 as sg
from saigegds_b200 import rsetup
import bench
n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (430000, 100000)
trait = sys.argv[3] if len(sys.argv) > 3 else "binary"
ctx = sg.Context(0)
ctx.store_synthetic(n, m, seed=200, missing_rate=0.005)
ph = bench.synth_phenotype(ctx, n, m)
X, _ = rsetup.qr_transform(rsetup.model_matrix(ph, ["x1", "x2"]))
param = sg.make_param()
if trait == "binary":
    fit0 = rsetup.glm_binomial(X, ph["y"])
else:
    f = rsetup.glm_gaussian(X, ph["yy"])
    fit0 = rsetup.glm_gaussian(X, rsetup.rank_norm(f.residuals) * rsetup.sd(f.residuals))
def run():
    t0 = time.perf_counter()
    if trait == "binary":
        g = ctx.saige_fit_AI_PCG_binary(fit0, X, rsetup.initial_tau_binary(), param)
    else:
        g = ctx.saige_fit_AI_PCG_quant(fit0, X, rsetup.initial_tau_quant(fit0), param)
    return time.perf_counter() - t0, g
t_first, g = run()
t_plain, g = run()
ctx.set_profiling(True)
t_prof, _ = run()
kt = ctx.kernel_times()
ctx.set_profiling(False)
tot = sum(v[0] for v in kt.values())
rows = sorted(kt.items(), key=lambda kv: -kv[1][0])
print(json.dumps({"n": n, "m": m, "trait": trait, "first_fit_s": t_first, "fit_s": t_plain, "fit_profiled_s": t_prof, "kernel_ms_total": tot, "tau": list(map(float, g["tau"])),
                  "kernels": {k: {"ms": round(v[0], 2), "launches": v[1]} for k, v in rows}}, indent=1))
