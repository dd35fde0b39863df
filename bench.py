#!/usr/bin/env python
"""bench.py -- GRM-vector products per second at N=430K x M=100K (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (CUDA, sm_100a)
  python bench.py --impl reference [...]                         # the reference algorithm on the host cores

One "step" = one product out = (1/M) G_std G_std' b over the whole packed genotype matrix
(get_crossprod_b_grm, saige_fitnull.cpp:435-536) for one right-hand side.  For N > 1 GPUs the fixed
N x M matrix is split by variant block across ranks (strong scaling) and every step ends with one
sum all-reduce of the N-vector.

value  : device-resident throughput (b and out already in HBM), CUDA events, max over ranks.
e2e    : the same step through the C-ABI entry point sgb_grm_mv with pinned HOST buffers (host->device copy of b,
         device->host copy of the result inside the timed region) -- the call an R user's .Call makes.
roofline: algorithmic bytes = ceil(N/4)*M_local packed bytes per product over the measured
         product time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
cpu_baseline: the CPU oracle's product (same algorithm and threading as the reference) on all host cores,
         on a bounded variant sample, extrapolated linearly in M to the full shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# NCCL_DEBUG=VERSION makes NCCL print a banner on stdout; rank 0 must print exactly one JSON line.
if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
    os.environ["NCCL_DEBUG"] = "WARN"

N_SAMP = int(os.environ.get("SGB_BENCH_N", 430000))
N_VAR = int(os.environ.get("SGB_BENCH_M", 100000))
MISSING = 0.005
METRIC = "grm_vector_products_per_s"
UNIT = "products/s"


def workload_name():
    return "synthetic N=%d M=%d (UKBB-scale) single-RHS GRM product, maf~U(0.005,0.5), %.1f%% missing" % (
        N_SAMP, N_VAR, 100 * MISSING)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ CPU arm
def numpy_packed_sample(n_samp, n_var, seed=200):
    """Same distribution as the device generator (store.cu synth_kernel); bytes differ, timing does not."""
    rng = np.random.default_rng(seed)
    nb = (n_samp + 3) // 4
    pool_n = min(n_var, 64)          # 64 distinct variants, tiled: the loop's cost does not depend on the codes
    pool = np.empty((pool_n, nb), dtype=np.uint8)
    for j in range(pool_n):
        maf = rng.uniform(0.005, 0.5)
        g = rng.binomial(2, maf, size=nb * 4).astype(np.uint8)
        g[rng.random(nb * 4) < MISSING] = 3
        g[n_samp:] = 3
        q = g.reshape(nb, 4)
        pool[j] = q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)
    return np.ascontiguousarray(pool[np.arange(n_var) % pool_n])


def cpu_product_rate(steps, warmup, m_sample=None):
    """Time the oracle's get_crossprod_b_grm on all host cores over a variant sample; extrapolate to N_VAR."""
    from oracle.oracle import Oracle, build, max_threads
    build()
    cores = max_threads()
    if m_sample is None:
        # ~0.9 G genotypes/s/core: aim at ~1 s per sampled product
        m_sample = int(max(64, min(N_VAR, 1.0 * 0.6e9 * cores / N_SAMP)))
    packed = numpy_packed_sample(N_SAMP, m_sample)
    o = Oracle()
    o.store_2b_geno(packed, N_SAMP, num_thread=cores)
    b = np.random.default_rng(1).standard_normal(N_SAMP)
    for _ in range(warmup):
        o.grm_mv(b)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.grm_mv(b)
    dt = (time.perf_counter() - t0) / steps
    full = dt * (N_VAR / m_sample)
    return dict(value=1.0 / full, unit=UNIT, cores=cores, kind="port",
                sample="%d of %d variants x %d samples, %d timed products of %.3f s each, scaled linearly in M"
                       % (m_sample, N_VAR, N_SAMP, steps, dt)), full


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    cb, full = cpu_product_rate(steps, max(1, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": workload_name(), "n_samp": N_SAMP, "n_var": N_VAR},
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import saigegds_b200 as sg
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = sg.Context(local_rank)
    if world > 1:
        sg.init_comm_from_torch(ctx)
    a, b_end = sg.shard_range(N_VAR, rank, world)
    m_local = b_end - a
    ctx.store_synthetic(N_SAMP, m_local, N_VAR, a, seed=200, missing_rate=MISSING)
    if args.kernel:
        ctx.set_kernel(args.kernel)

    def barrier():
        # torch's NCCL communicator and the library's own one must never have kernels in flight at the same time
        # (two communicators progressing in different orders on different ranks can deadlock): drain both sides.
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def note(msg):
        if os.environ.get("SGB_BENCH_VERBOSE"):
            print("[bench rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    note("stored %d variants" % m_local)
    rng = np.random.default_rng(1)
    b_host = rng.standard_normal(N_SAMP)
    d_b = ctx.device_vector(b_host)
    d_out = ctx.device_empty(8 * N_SAMP)

    # ---- device-resident throughput (value) ----
    # nvidia-smi takes a few hundred ms to deliver its first sample: start it before the warm-up and keep the GPU busy with
    # the same product until samples arrive, so that the clocks line describes the load of the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        ctx.grm_mv_device(d_b, d_out, 1)
    if world == 1:
        t_wait = time.perf_counter()
        while len(sampler.lines) < 2 and time.perf_counter() - t_wait < 3.0:
            ctx.grm_mv_device(d_b, d_out, 1)
    else:
        for _ in range(150):                 # every rank must issue the same number of products (each ends in a collective)
            ctx.grm_mv_device(d_b, d_out, 1)
    if rank == 0:
        sampler.lines.clear()
    note("warm-up done")
    ctx.reset_stats()
    barrier()
    ms = ctx.time_products_device(d_b, d_out, 1, args.steps)       # CUDA events on the launching stream, synced both sides
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    note("timed region done: %.3f ms" % ms)
    ms = max_over_ranks(ms)
    st = ctx.stats()
    launches = int(st["n_kernel_launches"])
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step

    # ---- end to end through the C-ABI with host buffers (e2e) ----
    # the step's input and result live in page-locked host memory (sgb_malloc_host); each call copies b host->device,
    # runs the product and copies the result device->host before it returns
    b_pin = ctx.pinned_empty(N_SAMP)
    out_pin = ctx.pinned_empty(N_SAMP)
    b_pin[:] = b_host
    for _ in range(2):
        ctx.get_crossprod_b_grm(b_pin, out=out_pin)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.get_crossprod_b_grm(b_pin, out=out_pin)
    barrier()
    out_host = np.array(out_pin)
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    note("e2e done")

    # ---- per-kernel event timing (separate, untimed pass) ----
    ctx.set_profiling(True)
    for _ in range(3):
        ctx.grm_mv_device(d_b, d_out, 1)
    ktimes = ctx.kernel_times()
    ctx.set_profiling(False)
    note("profiling pass done")

    # orderly teardown on every rank: library communicator first, then torch's process group
    d_b.free(); d_out.free()
    ctx.close()
    if dist is not None:
        barrier()
        dist.destroy_process_group()
    note("teardown done")
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    alg_bytes = ((N_SAMP + 3) // 4) * m_local
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    kern = {k: {"ms_per_launch": v[0] / v[1], "launches_per_product": v[1] / 3.0,
                "gbs_on_packed_bytes": alg_bytes / (v[0] / v[1] * 1e-3) / 1e9} for k, v in ktimes.items()}
    # dominant kernel = the one with the largest share of the step (event-timed, serialised pass above)
    dom = max(kern, key=lambda k: kern[k]["ms_per_launch"] * kern[k]["launches_per_product"])
    dom_gbs = kern[dom]["gbs_on_packed_bytes"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_ncu_fused_summary.json")
    if os.path.exists(tp) and world == 1 and N_SAMP == 430000 and N_VAR == 100000:
        tj = json.load(open(tp))
        if tj.get("kernel") == dom:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "dominant kernel: algorithmic bytes = ceil(N/4)*M_local packed bytes (read once) / its event-timed launch; "
                        "traffic = dram read+write bytes of one launch from the committed ncu --set full capture",
                "whole_product": {"achieved": achieved, "frac": achieved / peak,
                                  "note": "all kernels of one step (the headline `value`) against the same single-pass byte count"},
                "kernels": kern}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(), "n_samp": N_SAMP, "n_var": N_VAR, "n_var_per_gpu": m_local,
                   "parallelism": "variant-sharded x%d + sum all-reduce of the N-vector" % world,
                   "l2": "inputs (%.2f GB packed per GPU) exceed the 126 MB L2; no flush needed" % (alg_bytes / 1e9),
                   "kernel": args.kernel or "auto"},
        "clocks": clocks,
        "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * N_SAMP, "d2h_bytes_per_step": 8 * N_SAMP,
                "ms_per_step": e2e_s * 1e3, "note": "sgb_grm_mv with pinned host b/out (sgb_malloc_host); genotypes resident as in the reference"},
        "gpu_launches": launches,
        "roofline": roofline,
        "result_checksum": float(np.sum(out_host)),
    }
    if world == 1 and not args.no_cpu:
        cb, _ = cpu_product_rate(3, 1)
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ null-model fit (secondary metric)
def synth_phenotype(ctx, n, m, seed=7, n_causal=1000):
    """x1 ~ N(0,1), x2 ~ Bernoulli(0.5), g = sqrt(0.3) * standardised sum of causal columns."""
    rng = np.random.default_rng(seed)
    x1, x2 = rng.standard_normal(n), (rng.random(n) < 0.5).astype(np.float64)
    g = np.zeros(n)
    for j in rng.choice(m, size=min(n_causal, m), replace=False):
        ds = ctx.get_geno_ds(int(j))
        ds = np.where(np.isnan(ds), np.nanmean(ds), ds)
        sd = ds.std()
        if sd > 0:
            g += (ds - ds.mean()) / sd * rng.standard_normal()
    g = np.sqrt(0.3) * (g - g.mean()) / g.std()
    y = (rng.random(n) < 1 / (1 + np.exp(-(-2 + 0.5 * x1 + 0.5 * x2 + g)))).astype(np.float64)
    yy = x1 + x2 + g + rng.standard_normal(n)
    return dict(x1=x1, x2=x2, y=y, yy=yy)


def run_fit(args):
    """Null-model fit wall time (saige_fit_AI_PCG_* + saige_calc_var_ratio_*) on synthetic data, one GPU."""
    import saigegds_b200 as sg
    from saigegds_b200 import rsetup
    n, m = args.fit_n, args.fit_m
    ctx = sg.Context(0)
    t0 = time.perf_counter()
    ctx.store_synthetic(n, m, seed=200, missing_rate=MISSING)
    t_store = time.perf_counter() - t0
    ph = synth_phenotype(ctx, n, m)
    X, _ = rsetup.qr_transform(rsetup.model_matrix(ph, ["x1", "x2"]))
    param = sg.make_param(verbose=bool(os.environ.get("SGB_BENCH_VERBOSE")))
    out = {}
    for trait in args.fit_traits.split(","):
        ctx.reset_stats()
        # host-side restatement of the R set-up (glm start values, SPAtest null object): numpy on the box's shared cores,
        # reported separately from the native fit
        t0 = time.perf_counter()
        if trait == "binary":
            fit0 = rsetup.glm_binomial(X, ph["y"])
            noK = rsetup.null_model_binary(X, fit0)
        else:
            f = rsetup.glm_gaussian(X, ph["yy"])
            fit0 = rsetup.glm_gaussian(X, rsetup.rank_norm(f.residuals) * rsetup.sd(f.residuals))
            noK = rsetup.null_model_quant(X, fit0)
        t_setup = time.perf_counter() - t0
        t0 = time.perf_counter()
        if trait == "binary":
            glmm = ctx.saige_fit_AI_PCG_binary(fit0, X, rsetup.initial_tau_binary(), param)
        else:
            glmm = ctx.saige_fit_AI_PCG_quant(fit0, noK.X1, rsetup.initial_tau_quant(fit0), param)
        t_fit = time.perf_counter() - t0
        st_fit = ctx.stats()
        t0 = time.perf_counter()
        ctx.set_seed(200)
        fn = ctx.saige_calc_var_ratio_binary if trait == "binary" else ctx.saige_calc_var_ratio_quant
        vr = fn(fit0, glmm, noK, param, ctx.sample_int(m))
        t_vr = time.perf_counter() - t0
        st = ctx.stats()
        out[trait] = {"fit_s": t_fit, "host_setup_s": t_setup, "var_ratio_s": t_vr, "tau": [float(x) for x in glmm["tau"]],
                      "converged": glmm["converged"], "products_fit": int(st_fit["n_products"]),
                      "products_total": int(st["n_products"]), "pcg_solves": int(st["n_pcg_solves"]),
                      "pcg_iterations": int(st["n_pcg_iterations"]), "var_ratio_mean": float(np.mean(vr["ratio"])),
                      "n_markers": int(len(vr["ratio"]))}
    line = {"metric": "null_model_fit_wall_s", "unit": "s", "n_gpus": 1, "higher_is_better": False, "dtype": "f64",
            "data": "synthetic", "config": {"workload": "synthetic N=%d M=%d null fit + variance ratio" % (n, m)},
            "store_s": t_store, "traits": out}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ association scan
ASSOC_K = 10           # fixed-effect columns (intercept + 9 covariates)


def assoc_null_model(n, K=ASSOC_K, seed=7):
    """A self-consistent binary null model on synthetic covariates (what seqFitNullGLMM_SPA would hand over); the phenotype is
    drawn under the null, so ~5 % of the variants take the saddle-point branch."""
    import saigegds_b200 as sg
    from saigegds_b200 import rsetup
    rng = np.random.default_rng(seed)
    X = np.column_stack([np.ones(n)] + [rng.standard_normal(n) for _ in range(K - 1)])
    beta = np.concatenate([[-2.0], rng.normal(0, 0.2, K - 1)])
    mu = 1 / (1 + np.exp(-(X @ beta)))
    y = (rng.random(n) < mu).astype(np.float64)
    V = mu * (1 - mu)
    XVX_inv = np.linalg.inv(X.T @ (X * V[:, None]))
    noK = rsetup.ObjNoK(y=y, mu=mu, res=y - mu, V=V, X1=X, XV=(X * V[:, None]).T.copy(), XXVX_inv=X @ XVX_inv)
    return sg.NullModel(coefficients=beta, tau=np.array([1.0, 0.3]), linear_predictors=X @ beta, fitted_values=mu,
                        residuals=y - mu, cov=XVX_inv, converged=True, obj_noK=noK, var_ratio={"ratio": np.array([1.0])},
                        trait_type="binary")


def cpu_assoc_rate(n, m_sample):
    """The oracle's score test + SPA (= the reference's algorithm, src/saige_main.cpp:288-407) on all host cores."""
    from oracle import oracle as orc
    orc.build()
    mod = assoc_null_model(n)
    noK = mod.obj_noK
    m = orc.init_nullmod("binary", noK.y, mod.fitted_values, noK.X1, noK.XV, noK.XXVX_inv, noK.V, mod.tau)
    packed = numpy_packed_sample(n, m_sample)
    d = np.stack([(packed >> s) & 3 for s in (0, 2, 4, 6)], axis=2).reshape(m_sample, -1)[:, :n].astype(np.float64)
    d[d == 3] = np.nan
    t0 = time.perf_counter()
    orc.score_test(m, d, 1.0)
    dt = time.perf_counter() - t0
    return dict(value=m_sample / dt, unit="variants/s", cores=orc.max_threads(), kind="port",
                sample="%d variants x %d samples, K=%d, one timed pass of %.2f s (includes one copy of the dosage block)"
                       % (m_sample, n, ASSOC_K, dt))


def run_assoc(args):
    """Single-variant score test + SPA scan (seqAssocGLMM_SPA's inner loop) at N = 430K, K = 10, one GPU: variants per second
    with the genotypes resident in HBM (`value`) and from packed host batches (`e2e`)."""
    import saigegds_b200 as sg
    n, m = N_SAMP, args.assoc_m
    config = {"workload": "synthetic N=%d samples, %d variants, K=%d covariates, binary trait under the null, "
                          "maf~U(0.005,0.5), 0.5%% missing: score test + SPA per variant" % (n, m, ASSOC_K),
              "n_samp": n, "n_var": m, "K": ASSOC_K}
    if args.impl == "reference":
        cb = cpu_assoc_rate(n, args.assoc_cpu_m)
        print(json.dumps({"impl": "reference", "metric": "association_variants_per_s", "value": cb["value"],
                          "unit": "variants/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "higher_is_better": True,
                          "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb, "gpu_launches": 0,
                          "e2e": {"value": cb["value"], "unit": "variants/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}), flush=True)
        return
    ctx = sg.Context(0)
    ctx.store_synthetic(n, m, seed=200, missing_rate=MISSING)
    st = sg.ScoreTest(sg.init_nullmod(assoc_null_model(n)), ctx)
    for _ in range(max(1, args.warmup)):
        st.test_stored(0, min(m, 512))
    mon = ClockSampler(0)
    mon.start()
    ctx.reset_stats()
    times = []
    for _ in range(max(1, min(args.steps, 20))):
        res, ms = st.test_stored(0, m)
        times.append(ms)
    launches = ctx.stats()["n_kernel_launches"]
    host = ctx.synth_to_host(n, m)
    t0 = time.perf_counter()
    res_h = st.test(host)
    e2e_s = time.perf_counter() - t0
    clocks = mon.stop()
    assert all(np.array_equal(res[k], res_h[k], equal_nan=True) for k in res)
    ctx.set_profiling(True)
    st.test_stored(0, m)
    kt = ctx.kernel_times()
    ctx.set_profiling(False)
    ms = float(np.mean(times))
    # the kernel that bounds the scan reads n (2K + 2) doubles of model values per variant from shared memory
    smem_bytes = float(n) * (2 * ASSOC_K + 2) * 8 * m
    smem_peak = 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9          # GB/s: 128 B/clk/SM
    tiled_ms = kt.get("score_tiled_kernel", (ms, 1))[0]
    line = {"metric": "association_variants_per_s", "value": m / (ms * 1e-3), "unit": "variants/s", "n_gpus": 1,
            "steps": len(times), "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": m / e2e_s, "unit": "variants/s", "h2d_bytes_per_step": int(host.nbytes),
                    "d2h_bytes_per_step": int(m * 8 * 8 + m * 4),
                    "note": "ScoreTest.test on packed host batches (sgb_score_test_packed): copy in, both kernels, copy out"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "shared-memory", "kernel": "score_tiled_kernel", "achieved": smem_bytes / (tiled_ms * 1e-3) / 1e9,
                         "peak": smem_peak, "unit": "GB/s", "frac": smem_bytes / (tiled_ms * 1e-3) / 1e9 / smem_peak,
                         "traffic": None, "peak_source": "148 SMs x 128 B/clk x SM clock under load",
                         "note": "algorithmic shared-memory bytes = n (2K+2) 8 B per variant; the packed genotypes "
                                 "(n/4 B per variant from HBM) are 0.1 % of that",
                         "kernels": {k: {"ms": v[0], "launches": v[1]} for k, v in kt.items()}},
            "spa_adjusted": int(np.sum(res["pval"][res["valid"]] != res["p.norm"][res["valid"]])),
            "valid": int(res["valid"].sum())}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_assoc_rate(n, args.assoc_cpu_m)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default=None, choices=[None, "auto", "simt", "imma", "imma2"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--mode", default="product", choices=["product", "fit", "assoc"],
                    help="product: the headline metric; fit: null-model fit wall time; assoc: score test + SPA scan "
                         "(secondary metrics)")
    ap.add_argument("--assoc-m", type=int, default=9472, help="variants in the association-scan benchmark")
    ap.add_argument("--assoc-cpu-m", type=int, default=192, help="variants the CPU arm of --mode assoc times")
    ap.add_argument("--fit-n", type=int, default=50000)
    ap.add_argument("--fit-m", type=int, default=100000)
    ap.add_argument("--fit-traits", default="binary,quantitative")
    args = ap.parse_args()
    if args.mode == "assoc":
        run_assoc(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.mode == "fit":
        run_fit(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
