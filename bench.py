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
cpu_baseline: the CPU oracle's product (same algorithm and threading as the reference) on all host cores, on the first
         variants of the SAME bytes the GPU stores (oracle generator == device generator), scaled linearly in M.
batched: the same product for K = 30 right-hand sides at once (the trace step of the fit) through the tcgen05 path.
fit    : wall time of the C3 binary null-model fit + variance ratio on the stored shard(s) (second half of the metric).

  python bench.py --mode batched   # C5: N=430K, M=300K, K=30, tensor roofline
  python bench.py --mode fit       # fit wall times only (also under torch.distributed.run)
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
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_oracle(cores):
    """torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm sets the thread count itself, before libgomp loads."""
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from oracle import oracle as orc
    orc.build()
    return orc


def numpy_packed_sample(n_samp, n_var, seed=200):
    """Used by --mode assoc only: same distribution as the device generator, 64 distinct variants tiled."""
    rng = np.random.default_rng(seed)
    nb = (n_samp + 3) // 4
    pool_n = min(n_var, 64)
    pool = np.empty((pool_n, nb), dtype=np.uint8)
    for j in range(pool_n):
        maf = rng.uniform(0.005, 0.5)
        g = rng.binomial(2, maf, size=nb * 4).astype(np.uint8)
        g[rng.random(nb * 4) < MISSING] = 3
        g[n_samp:] = 3
        q = g.reshape(nb, 4)
        pool[j] = q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)
    return np.ascontiguousarray(pool[np.arange(n_var) % pool_n])


def cpu_product_rate(steps, warmup, budget_s, b_host=None):
    """Time the oracle's get_crossprod_b_grm on all host cores.  The genotypes are the first m_s variants of the very matrix the
    GPU arm stores (the oracle's generator is bit-identical to the device one, tests/test_batched_product.py); m_s = all N_VAR
    variants when (warmup + steps) products fit the time budget, otherwise the largest prefix that does (said in `sample`)."""
    cores = host_cores()
    orc = load_oracle(cores)
    b = np.random.default_rng(1).standard_normal(N_SAMP) if b_host is None else b_host
    # calibration on 1,024 variants (stored separately; ~0.3 s)
    o = orc.Oracle()
    cal = orc.synth_geno(N_SAMP, min(4096, N_VAR), 0, 200, MISSING, cores)
    o.store_2b_geno(cal, N_SAMP, num_thread=cores, borrow=True, want_diag=False)
    o.grm_mv(b)
    t0 = time.perf_counter()
    o.grm_mv(b)
    per_var = (time.perf_counter() - t0) / len(cal)
    gen_per_var = 4.0e-4 * 8 / max(cores, 1) * (N_SAMP / 430000.0)                  # ~1.9 G genotypes/s on 8 cores
    m_s = int(min(N_VAR, max(1024, budget_s / ((steps + warmup) * per_var + gen_per_var))))
    t0 = time.perf_counter()
    packed = orc.synth_geno(N_SAMP, m_s, 0, 200, MISSING, cores) if m_s != len(cal) else cal
    t_gen = time.perf_counter() - t0
    o.store_2b_geno(packed, N_SAMP, num_thread=cores, borrow=True, want_diag=False)
    for _ in range(warmup):
        o.grm_mv(b)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = o.grm_mv(b)
    dt = (time.perf_counter() - t0) / steps
    full = dt * (N_VAR / m_s)
    what = "all %d variants (full product measured, no extrapolation)" % N_VAR if m_s == N_VAR else \
        "the first %d of %d variants, scaled linearly in M" % (m_s, N_VAR)
    return dict(value=1.0 / full, unit=UNIT, cores=cores, kind="port",
                sample="%s x %d samples, same bytes and same b as the GPU arm; %d warm-up + %d timed products of %.3f s each "
                       "(matrix generated on the host in %.1f s, not timed)" % (what, N_SAMP, warmup, steps, dt, t_gen),
                b_dot_out_sample=float(b @ out) * (m_s / N_VAR), m_sample=m_s), full


def product_config(world, kernel):
    m_local = sg_shard(N_VAR, 0, world)
    return {"workload": workload_name(), "n_samp": N_SAMP, "n_var": N_VAR, "n_var_per_gpu": m_local,
            "parallelism": "variant-sharded x%d + sum all-reduce of the N-vector" % world,
            "l2": "inputs (%.2f GB packed per GPU) exceed the 126 MB L2; no flush needed" % (((N_SAMP + 3) // 4) * m_local / 1e9),
            "kernel": kernel or "auto"}


def sg_shard(m, rank, world):
    """Variants of rank `rank` (same split as saigegds_b200.shard_range, restated so the CPU arm does not import the package)."""
    base, rem = divmod(int(m), int(world))
    return base + (1 if rank < rem else 0)


ARITHMETIC = ("FP64 in / FP64 out; the two matrix-vector halves are exact integer arithmetic on int8 tensor cores: b and e are "
              "quantised to 56-bit fixed point relative to their largest element (8 signed base-128 digits, mma.sync u8 x s8 -> s32) "
              "in the single-RHS kernels, 46-bit (6 signed base-256 digits, tcgen05.mma kind::i8) in the batched path; "
              "dot / e / h and all scalings in FP64; <= 1e-10 of ||out||_inf vs the FP64 oracle (tests)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb, full = cpu_product_rate(steps, warmup, float(os.environ.get("SGB_BENCH_CPU_BUDGET_S", 300)))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": product_config(max(1, args.gpus), args.kernel),
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if cb["m_sample"] == N_VAR:
        # same bytes, same b: this equals the GPU arm's result_invariants.b_dot_out up to rounding
        line["result_invariants"] = {"b_dot_out": cb["b_dot_out_sample"]}
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
        while len(sampler.lines) < 6 and time.perf_counter() - t_wait < 3.0:
            ctx.grm_mv_device(d_b, d_out, 1)
    else:
        for _ in range(600):                 # every rank must issue the same number of products (each ends in a collective)
            ctx.grm_mv_device(d_b, d_out, 1)
    # (the samples of the warm-up stay in: the timed region of 20 products lasts ~70 ms, one nvidia-smi period; the warm-up
    #  runs the same product back to back, so the clocks line describes the same load)
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

    # ---- the same product for K = 30 right-hand sides at once (trace step of the fit; saige_fitnull.cpp:646-654) ----
    batched = None
    if not args.no_batched:
        kb = 30
        Bh = np.asfortranarray(np.random.default_rng(2).standard_normal((N_SAMP, kb)))
        Bh[:, 0] = b_host
        d_B = ctx.device_vector(Bh.reshape(-1, order="F"))
        d_O = ctx.device_empty(8 * N_SAMP * kb)
        for _ in range(2):
            ctx.grm_mv_device(d_B, d_O, kb)
        barrier()
        reps = 5
        msb = max_over_ranks(ctx.time_products_device(d_B, d_O, kb, reps)) / reps
        barrier()
        col0 = ctx.get_crossprod_b_grm(Bh[:, :3])[:, 0]            # k = 3: batched path, through the host entry point
        batched = {"k": kb, "ms_per_call": msb, "products_per_s": kb * 1e3 / msb,
                   "col0_vs_single_rhs_relinf": float(np.max(np.abs(col0 - out_host)) / np.max(np.abs(out_host))),
                   "note": "tcgen05.mma kind::i8, accumulators in TMEM; one pass over both orientations of the packed matrix "
                           "serves all K columns (csrc/grm_umma.cuh); device-resident, all-reduce included"}
        d_B.free(); d_O.free()
        note("batched done")

    # ---- null-model fit on the stored shard(s): second half of the metric ----
    fit = None
    if not args.no_fit:
        fit = fit_on_stored(ctx, N_SAMP, N_VAR, a, m_local, rank, world, dist, ["binary"])["binary"]
        note("fit done")

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
                "gbs_on_packed_bytes": alg_bytes / (v[0] / v[1] * 1e-3) / 1e9} for k, v in ktimes.items() if v[0] > 0}
    # dominant kernel = the one with the largest share of the step (event-timed, serialised pass above)
    dom = max(kern, key=lambda k: kern[k]["ms_per_launch"] * kern[k]["launches_per_product"])
    dom_gbs = kern[dom]["gbs_on_packed_bytes"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_ncu_fused_summary.json")
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "r01_ncu_fused_summary.json")
    if os.path.exists(tp) and world == 1 and N_SAMP == 430000 and N_VAR == 100000:
        tj = json.load(open(tp))
        if tj.get("kernel") == dom:
            traffic = tj.get("dram_bytes_per_launch")
    # the three floors of the single-RHS kernel: HBM stream, tensor pipe (16 int8 MACs per genotype = 8 digit
    # planes x 2 phases, legacy mma.sync IMMA measured at 2,048 MAC/clk/SM = 595 TMAC/s at 1,965 MHz) and the half-rate integer
    # pipe that forms the operands (LOP3: 7 per 16 genotypes in phase A + B, 64 lanes/clk/SM)
    geno = float(N_SAMP) * m_local
    floors = {"hbm_ms": alg_bytes / (peak * 1e9) * 1e3, "imma_pipe_ms": 16.0 * geno / (2048.0 * 148 * 1.965e9) * 1e3,
              "int_pipe_ms": (7.0 / 16.0) * geno / (64.0 * 148 * 1.965e9) * 1e3,
              "note": "HBM: packed bytes / measured copy bandwidth; tensor: 16 MAC per genotype / measured mma.sync IMMA rate "
                      "(profiles/r01_microbench_b200.txt); integer: operand-forming LOP3s at 64 lanes/clk/SM"}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
                "traffic": traffic, "peak_source": peak_src, "floors": floors,
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
        "config": product_config(world, args.kernel),
        "arithmetic": ARITHMETIC,
        "clocks": clocks,
        "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * N_SAMP, "d2h_bytes_per_step": 8 * N_SAMP,
                "ms_per_step": e2e_s * 1e3, "note": "sgb_grm_mv with pinned host b/out (sgb_malloc_host); genotypes resident as in the reference"},
        "gpu_launches": launches,
        "roofline": roofline,
        # b'(GRM b) = (1/M) sum_j (g_j'b)^2 > 0: a checksum that does not cancel, identical for every number of GPUs up to rounding
        "result_invariants": {"b_dot_out": float(b_host @ out_host), "out_l2": float(np.linalg.norm(out_host)),
                              "out_absmax": float(np.max(np.abs(out_host)))},
    }
    if batched is not None:
        line["batched"] = batched
    if fit is not None:
        line["fit"] = fit
    if world == 1 and not args.no_cpu:
        cb, _ = cpu_product_rate(2, 1, float(os.environ.get("SGB_BENCH_CPU_BUDGET_S", 25)), b_host)
        # the same invariant on the CPU: b'(GRM_s b) over the sampled variants, against the GPU's value restricted to them is not
        # available, so report it for the record only
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ null-model fit (second half of the metric)
def synth_phenotype(ctx, n, m_total, var_offset=0, m_local=None, dist=None, seed=7, n_causal=1000, h2=None, intercept=None):
    """x1 ~ N(0,1), x2 ~ Bernoulli(0.5), g = sqrt(0.3) * standardised sum of causal columns.  With the
    variants sharded over ranks every rank adds the causal columns it owns and the partial sums are all-reduced, so all ranks
    hold the same phenotype."""
    m_local = m_total if m_local is None else m_local
    # SURVEY 8(d) prescribes g = sqrt(0.3) x standardised score and intercept -2 (the C3 settings).  Below ~100K samples the binary
    # fit of that phenotype ends at Sigma_G = 0 (the GRM product is then skipped, :568, and the benchmark would measure nothing), so
    # smaller shapes use a stronger genetic component and a more balanced outcome; both are reported with the fit.
    if h2 is None:
        h2 = float(os.environ.get("SGB_BENCH_H2", 0.3 if n >= 200000 else 1.0))
    if intercept is None:
        intercept = float(os.environ.get("SGB_BENCH_INTERCEPT", -2.0 if n >= 200000 else -1.0))
    rng = np.random.default_rng(seed)
    x1, x2 = rng.standard_normal(n), (rng.random(n) < 0.5).astype(np.float64)
    causal = rng.choice(m_total, size=min(n_causal, m_total), replace=False)
    effect = rng.standard_normal(len(causal))
    g = np.zeros(n)
    for j, beta in zip(causal, effect):
        if var_offset <= j < var_offset + m_local:
            ds = ctx.get_geno_ds(int(j - var_offset))
            ds = np.where(np.isnan(ds), np.nanmean(ds), ds)
            sd = ds.std()
            if sd > 0:
                g += (ds - ds.mean()) / sd * beta
    if dist is not None:
        import torch
        t = torch.from_numpy(g).cuda()
        dist.all_reduce(t)
        torch.cuda.synchronize()
        g = t.cpu().numpy()
    g = np.sqrt(h2) * (g - g.mean()) / g.std()
    u, e = rng.random(n), rng.standard_normal(n)
    y = (u < 1 / (1 + np.exp(-(intercept + 0.5 * x1 + 0.5 * x2 + g)))).astype(np.float64)
    yy = x1 + x2 + g + e
    return dict(x1=x1, x2=x2, y=y, yy=yy, h2=h2, intercept=intercept)


def fit_on_stored(ctx, n, m_total, var_offset, m_local, rank, world, dist, traits):
    """saige_fit_AI_PCG_* + saige_calc_var_ratio_* (saige_fitnull.cpp:949-1474) on the genotypes already stored in `ctx`
    (every rank its shard); wall-clock seconds, max over ranks."""
    import saigegds_b200 as sg
    from saigegds_b200 import rsetup

    def wall_max(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
        return float(t.item())

    ph = synth_phenotype(ctx, n, m_total, var_offset, m_local, dist)
    X, _ = rsetup.qr_transform(rsetup.model_matrix(ph, ["x1", "x2"]))
    param = sg.make_param(verbose=bool(os.environ.get("SGB_BENCH_VERBOSE")) and rank == 0)
    out = {}
    for trait in traits:
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
        res = None
        for rep in range(2):          # first pass: lazy one-time set-up (sample-major copy, work buffers); second pass: reported
            ctx.reset_stats()
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
            vr = fn(fit0, glmm, noK, param, ctx.sample_int(m_total))
            t_vr = time.perf_counter() - t0
            st = ctx.stats()
            cur = {"fit_s": wall_max(t_fit), "var_ratio_s": wall_max(t_vr)}
            if rep == 0:
                first = cur
            res = cur
        out[trait] = {"fit_s": res["fit_s"], "var_ratio_s": res["var_ratio_s"], "first_call_fit_s": first["fit_s"],
                      "host_setup_s": t_setup, "tau": [float(x) for x in glmm["tau"]], "converged": bool(glmm["converged"]),
                      "products_fit": int(st_fit["n_products"]), "products_total": int(st["n_products"]),
                      "pcg_solves": int(st["n_pcg_solves"]), "pcg_iterations": int(st["n_pcg_iterations"]),
                      "host_syncs_fit": int(st_fit["n_host_syncs"]), "host_wait_s_fit": float(st_fit["host_wait_s"]),
                      "var_ratio_mean": float(np.mean(vr["ratio"])), "n_markers": int(len(vr["ratio"])),
                      "workload": "synthetic N=%d M=%d %s trait, y ~ x1 + x2 + g, var(g) = %.2g from 1,000 causal variants, intercept %.1f; "
                                  "seqFitNullGLMM_SPA defaults (nrun 30, tol 0.02, tolPCG 1e-5)" % (n, m_total, trait, ph["h2"], ph["intercept"]),
                      "note": "fit_s = saige_fit_AI_PCG_%s, var_ratio_s = saige_calc_var_ratio_%s; second call on the stored "
                              "genotypes (first_call_fit_s includes building the sample-major copy once)" % (trait, trait)}
    return out


def cpu_fit_times(n, m, traits, cores):
    """The oracle's full fit (= the reference's algorithm, all host cores) on a shape it finishes in seconds, for the record."""
    orc = load_oracle(cores)
    from saigegds_b200 import rsetup
    packed = orc.synth_geno(n, m, 0, 200, MISSING, cores)
    o = orc.Oracle()
    t0 = time.perf_counter()
    o.store_2b_geno(packed, n, num_thread=cores)
    t_store = time.perf_counter() - t0

    class _Ctx:          # the phenotype generator only needs get_geno_ds
        def get_geno_ds(self, j):
            return o.get_geno_ds(j)
    ph = synth_phenotype(_Ctx(), n, m)
    X, _ = rsetup.qr_transform(rsetup.model_matrix(ph, ["x1", "x2"]))
    out = {"n": n, "m": m, "cores": cores, "store_s": t_store}
    for trait in traits:
        if trait == "binary":
            fit0 = rsetup.glm_binomial(X, ph["y"])
            tau0 = rsetup.initial_tau_binary()
        else:
            f = rsetup.glm_gaussian(X, ph["yy"])
            fit0 = rsetup.glm_gaussian(X, rsetup.rank_norm(f.residuals) * rsetup.sd(f.residuals))
            tau0 = rsetup.initial_tau_quant(fit0)
        t0 = time.perf_counter()
        r = o.fit_AI_PCG(trait if trait == "binary" else "quantitative", fit0, X, tau0)
        out[trait] = {"fit_s": time.perf_counter() - t0, "tau": [float(x) for x in r["tau"]], "products": int(o.num_products)}
    return out


def run_fit(args):
    """Null-model fit wall time on synthetic data at N GPUs (variants sharded; launch with torch.distributed.run for N > 1)."""
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
    n, m = args.fit_n, args.fit_m
    ctx = sg.Context(local_rank)
    if world > 1:
        sg.init_comm_from_torch(ctx)
    a, b_end = sg.shard_range(m, rank, world)
    t0 = time.perf_counter()
    ctx.store_synthetic(n, b_end - a, m, a, seed=200, missing_rate=MISSING)
    t_store = time.perf_counter() - t0
    out = fit_on_stored(ctx, n, m, a, b_end - a, rank, world, dist, args.fit_traits.split(","))
    ctx.close()
    if dist is not None:
        import torch
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {"metric": "null_model_fit_wall_s", "value": out[args.fit_traits.split(",")[0]]["fit_s"], "unit": "s", "n_gpus": world,
            "higher_is_better": False, "scaling": "strong", "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic N=%d M=%d null fit + variance ratio" % (n, m),
                       "parallelism": "variant-sharded x%d" % world},
            "store_s": t_store, "traits": out}
    if args.cpu_fit and world == 1:
        cn, cm = [int(x) for x in args.cpu_fit.split("x")]
        line["cpu_fit"] = cpu_fit_times(cn, cm, args.fit_traits.split(","), host_cores())
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ batched K-RHS product (BASELINE config 5)
def run_batched(args):
    """K right-hand sides per step through the batched tcgen05 path (the trace estimation of the fit: 30 Rademacher vectors per
    PCG iteration, saige_fitnull.cpp:646-654) at N = 430K, M = 300K (C5).  value = columns (products) per second."""
    global N_VAR
    N_VAR = int(os.environ.get("SGB_BENCH_M5", 300000))
    K = int(os.environ.get("SGB_BENCH_K", 30))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "synthetic N=%d M=%d, K=%d right-hand sides per step (batched multi-RHS GRM product, trace estimation), "
                          "maf~U(0.005,0.5), %.1f%% missing" % (N_SAMP, N_VAR, K, 100 * MISSING),
              "n_samp": N_SAMP, "n_var": N_VAR, "k": K, "n_var_per_gpu": sg_shard(N_VAR, 0, world),
              "parallelism": "variant-sharded x%d + sum all-reduce of the N x K block" % world,
              "l2": "inputs (%.1f GB packed per GPU, both orientations) exceed the 126 MB L2; no flush needed"
                    % (2 * ((N_SAMP + 3) // 4) * sg_shard(N_VAR, 0, world) / 1e9)}
    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        cb, full = cpu_product_rate(max(1, min(steps, 3)), min(warmup, 1), float(os.environ.get("SGB_BENCH_CPU_BUDGET_S", 120)))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                          "steps": steps, "warmup": warmup, "ms_per_step": full * K * 1e3, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "gpu_launches": 0, "note": "the reference has no batched product: K columns cost K single products",
                          "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
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

    def barrier():
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

    rng = np.random.default_rng(1)
    B = np.asfortranarray(rng.integers(0, 2, (N_SAMP, K)) * 2.0 - 1.0)          # Rademacher columns, as in get_trace
    d_B = ctx.device_vector(B.reshape(-1, order="F"))
    d_O = ctx.device_empty(8 * N_SAMP * K)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    warm = max(3, args.warmup)
    for _ in range(warm):
        ctx.grm_mv_device(d_B, d_O, K)
    ctx.reset_stats()
    barrier()
    steps = max(1, args.steps)
    ms = max_over_ranks(ctx.time_products_device(d_B, d_O, K, steps)) / steps
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = int(ctx.stats()["n_kernel_launches"])
    # e2e: the C-ABI entry point with pinned host blocks (N x K in, N x K out)
    b_pin, o_pin = ctx.pinned_empty((N_SAMP, K)), ctx.pinned_empty((N_SAMP, K))
    b_pin[:] = B
    ctx.get_crossprod_b_grm(b_pin, out=o_pin)
    barrier()
    t0 = time.perf_counter()
    e_steps = max(1, min(steps, 5))
    for _ in range(e_steps):
        ctx.get_crossprod_b_grm(b_pin, out=o_pin)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e_steps)
    out_host = np.array(o_pin)
    # single-RHS kernel on two of the columns: the batched result must agree with it
    ctx.set_kernel("imma")
    chk = max(float(np.max(np.abs(ctx.get_crossprod_b_grm(B[:, c]) - out_host[:, c])) / np.max(np.abs(out_host[:, c]))) for c in (0, K - 1))
    ctx.set_kernel("auto")
    ctx.set_profiling(True)
    ctx.grm_mv_device(d_B, d_O, K)
    kt = ctx.kernel_times()
    ctx.set_profiling(False)
    d_B.free(); d_O.free()
    ctx.close()
    if dist is not None:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    gemm_ms = sum(v[0] / v[1] for k, v in kt.items() if k.startswith("umma_gemm_kernel"))
    macs = 2.0 * float(N_SAMP) * m_local * 6 * K                    # both phases, 6 digit columns per right-hand side
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    bf16 = float(pk.get("bf16_tflops", 1590.0))
    probe_tops = 4179.0          # tools/umma_probe.cu on this pool's B200: back-to-back kind::i8 MMAs, A from TMEM, N = 240
    peak_tops = max(2.0 * bf16, probe_tops)
    ach_tops = 2.0 * macs / (gemm_ms * 1e-3) / 1e12
    hbm, hbm_src = measured_peaks()
    line = {"metric": METRIC, "value": K * 1e3 / ms, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8 digits / f64",
            "arithmetic": ARITHMETIC, "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * N_SAMP * K, "d2h_bytes_per_step": 8 * N_SAMP * K,
                    "ms_per_step": e2e_s * 1e3, "note": "sgb_grm_mv with K pinned host columns in and out"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "umma_gemm_kernel (phase A + phase B)", "achieved": ach_tops, "peak": peak_tops,
                         "unit": "TOP/s", "frac": ach_tops / peak_tops, "traffic": None,
                         "peak_source": "measured int8 tensor rate of this part: 4,179 TOP/s for back-to-back tcgen05.mma kind::i8 (tools/umma_probe.cu, "
                                        "profiles/r02_umma_probe_b200.txt); MEASURED_PEAKS.json has no int8 figure -- 2 x its bf16_tflops would be "
                                        "%.0f TOP/s, which this kernel exceeds; nominal 4,500" % (2.0 * bf16),
                         "executed_over_algorithmic": (((6 * K + 15) // 16) * 16) / (6.0 * K),
                         "algorithmic_macs_per_step": macs,
                         "note": "algorithmic int8 MACs = 2 phases x N x M_local x 6 digit columns x K, over the event-timed duration of the "
                                 "two GEMM launches (serialised profiling pass)",
                         "hbm": {"packed_bytes_both_orientations": 2 * ((N_SAMP + 3) // 4) * m_local,
                                 "achieved_gbs": 2 * ((N_SAMP + 3) // 4) * m_local / (gemm_ms * 1e-3) / 1e9, "peak_gbs": hbm, "peak_source": hbm_src},
                         "kernels": {k: {"ms_per_launch": v[0] / v[1], "launches": v[1]} for k, v in kt.items() if v[0] > 0}},
            "batched_vs_single_rhs_relinf": chk,
            "result_invariants": {"b_dot_out_col0": float(B[:, 0] @ out_host[:, 0]), "out_l2": float(np.linalg.norm(out_host))}}
    if world == 1 and not args.no_cpu:
        cb, _ = cpu_product_rate(2, 1, float(os.environ.get("SGB_BENCH_CPU_BUDGET_S", 25)))
        line["cpu_baseline"] = cb
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
    st.test_stored(0, m)          # the workspaces grow to the batch size here, not inside the timed steps
    mon = ClockSampler(0)
    mon.start()
    ctx.reset_stats()
    times = []
    for _ in range(max(1, min(args.steps, 20))):
        res, ms = st.test_stored(0, m)
        times.append(ms)
    launches = ctx.stats()["n_kernel_launches"]
    host0 = ctx.synth_to_host(n, m)
    # the batch in page-locked host memory (sgb_malloc_host), as a pipeline that reads genotype blocks for the GPU would hold it
    pin = ctx.pinned_empty((host0.size + 7) // 8)
    host = pin.view(np.uint8)[:host0.size].reshape(host0.shape)
    host[...] = host0
    del host0
    st.test(host[:512])
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
    # dense part of the scan: three integer GEMMs (bit planes of the codes x digit planes of the 2K + 4 model columns) on tcgen05
    ncols = 2 * ASSOC_K + 4
    gemm_key = next((k for k in kt if k.startswith("umma_pair_kernel")), None)
    shares = {k: v[0] / max(1e-9, sum(x[0] for x in kt.values())) for k, v in kt.items()}
    if gemm_key is not None:
        gemm_ms = kt[gemm_key][0] / max(1, kt[gemm_key][1])
        macs = 3.0 * float(n) * m * 6 * ncols
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak_tops = max(2.0 * float(pk.get("bf16_tflops", 1590.0)), 4179.0)
        ach = 2.0 * macs / (gemm_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "umma_pair_kernel (class sums: low / high / low & high bit planes of the 2-bit codes x "
                                             "6 digit planes of %d model columns)" % ncols,
                "achieved": ach, "peak": peak_tops, "unit": "TOP/s", "frac": ach / peak_tops, "traffic": None,
                "peak_source": "int8 tcgen05 rate measured with tools/umma_probe.cu on this pool's B200 (4,179 TOP/s), or 2 x bf16 of "
                               "MEASURED_PEAKS.json if larger",
                "note": "algorithmic MACs = 3 planes x n x variants x 6 digits x (2K + 4) columns; a pair MMA costs ~96 clk whatever "
                        "N <= 192, so %d digit columns of 192 leave the pipe under-used by construction. The scan as a whole is bound by "
                        "spa_candidate_kernel (saddle-point candidates, FP64 transcendentals): see `shares`" % (6 * ncols)}
    else:
        # CUDA-core tiled kernel: reads n (2K + 2) doubles of model values per variant from shared memory
        smem_bytes = float(n) * (2 * ASSOC_K + 2) * 8 * m
        smem_peak = 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9          # GB/s: 128 B/clk/SM
        tiled_ms = kt.get("score_tiled_kernel", (ms, 1))[0]
        roof = {"bound": "shared-memory", "kernel": "score_tiled_kernel", "achieved": smem_bytes / (tiled_ms * 1e-3) / 1e9,
                "peak": smem_peak, "unit": "GB/s", "frac": smem_bytes / (tiled_ms * 1e-3) / 1e9 / smem_peak, "traffic": None,
                "peak_source": "148 SMs x 128 B/clk x SM clock under load",
                "note": "algorithmic shared-memory bytes = n (2K+2) 8 B per variant; the packed genotypes (n/4 B per variant from "
                        "HBM) are 0.1 % of that"}
    roof["kernels"] = {k: {"ms": v[0], "launches": v[1]} for k, v in kt.items()}
    roof["shares"] = shares
    line = {"metric": "association_variants_per_s", "value": m / (ms * 1e-3), "unit": "variants/s", "n_gpus": 1,
            "steps": len(times), "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": m / e2e_s, "unit": "variants/s", "h2d_bytes_per_step": int(host.nbytes),
                    "d2h_bytes_per_step": int(m * 8 * 8 + m * 4),
                    "note": "ScoreTest.test on a packed host batch in page-locked memory (sgb_score_test_packed): copy in, all kernels, copy out"},
            "gpu_launches": int(launches), "score_path": os.environ.get("SGB_SCORE_PATH", "tensor"),
            "roofline": roof,
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
    ap.add_argument("--kernel", default=None, choices=[None, "auto", "simt", "imma", "imma2", "umma"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--mode", default="product", choices=["product", "batched", "fit", "assoc"],
                    help="product: the headline metric; fit: null-model fit wall time; assoc: score test + SPA scan "
                         "(secondary metrics)")
    ap.add_argument("--assoc-m", type=int, default=9472, help="variants in the association-scan benchmark")
    ap.add_argument("--assoc-cpu-m", type=int, default=192, help="variants the CPU arm of --mode assoc times")
    ap.add_argument("--fit-n", type=int, default=50000)
    ap.add_argument("--fit-m", type=int, default=100000)
    ap.add_argument("--fit-traits", default="binary,quantitative")
    ap.add_argument("--cpu-fit", default="", help="--mode fit: also time the CPU oracle's fit at NxM (e.g. 1000x9976 or 50000x10000)")
    ap.add_argument("--no-batched", action="store_true", help="product mode: skip the K = 30 batched leg")
    ap.add_argument("--no-fit", action="store_true", help="product mode: skip the null-model fit leg")
    args = ap.parse_args()
    if args.mode == "assoc":
        run_assoc(args)
    elif args.mode == "batched":
        run_batched(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.mode == "fit":
        run_fit(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
