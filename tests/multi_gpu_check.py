"""Multi-GPU parity check, run under torchrun (one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/multi_gpu_check.py

Every rank stores its variant shard of the reference fixture; the sharded product, the PCG solve and the whole
binary null-model fit must reproduce the single-GPU results / the reference's golden tau.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import saigegds_b200 as sg
    from saigegds_b200 import rsetup
    from conftest import Fixture
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fx = Fixture()
    n, m = fx.n_samp, len(fx.packed)
    ctx = sg.Context(local)
    sg.init_comm_from_torch(ctx)
    a, b = sg.shard_range(m, rank, world)
    lut, diag = ctx.saige_store_2b_geno(fx.packed[a:b], n, n_variant_total=m, variant_offset=a)
    # single-GPU reference on every rank (same device, second context)
    ref = sg.Context(local)
    rlut, rdiag = ref.saige_store_2b_geno(fx.packed, n)
    assert np.array_equal(lut, rlut[a:b])
    assert np.max(np.abs(diag - rdiag)) / np.max(np.abs(rdiag)) < 1e-13
    vec = np.random.default_rng(3).standard_normal(n)
    for kern in ("simt", "imma"):
        ctx.set_kernel(kern); ref.set_kernel(kern)
        got, want = ctx.get_crossprod_b_grm(vec), ref.get_crossprod_b_grm(vec)
        err = np.max(np.abs(got - want)) / np.max(np.abs(want))
        assert err < 1e-12, (kern, err)
    ctx.set_kernel("auto"); ref.set_kernel("auto")
    # batched tcgen05 path on the shards (k >= 2) against the single-GPU single-RHS kernel, and N-invariance of b'(GRM b)
    B = np.random.default_rng(4).standard_normal((n, 6))
    gotB = ctx.get_crossprod_b_grm(B)
    ref.set_kernel("imma2")
    for c in range(6):
        wantc = ref.get_crossprod_b_grm(B[:, c])
        errc = np.max(np.abs(gotB[:, c] - wantc)) / np.max(np.abs(wantc))
        assert errc < 1e-11, ("batched", c, errc)
    ref.set_kernel("auto")
    inv = float(vec @ ctx.get_crossprod_b_grm(vec)), float(vec @ ref.get_crossprod_b_grm(vec))
    assert abs(inv[0] - inv[1]) / inv[1] < 1e-12, inv
    w = np.full(n, 0.2)
    x, it = ctx.PCG_diag_sigma(w, np.array([1.0, 0.5]), vec)
    xr, itr = ref.PCG_diag_sigma(w, np.array([1.0, 0.5]), vec)
    assert it == itr and np.max(np.abs(x - xr)) / np.max(np.abs(xr)) < 1e-10
    # the whole fit + variance ratio, against the reference's golden model
    X, R = rsetup.qr_transform(rsetup.model_matrix(fx.pheno, ["x1", "x2"]))
    fit0 = rsetup.glm_binomial(X, fx.pheno["y"])
    glmm = ctx.saige_fit_AI_PCG_binary(fit0, X, rsetup.initial_tau_binary())
    g = fx.model
    assert abs(glmm["tau"][1] - g["tau"][1]) / g["tau"][1] < 1e-6, glmm["tau"]
    ctx.set_seed(200)
    vr = ctx.saige_calc_var_ratio_binary(fit0, {"tau": g["tau"]}, rsetup.null_model_binary(X, fit0), sg.make_param(),
                                         ctx.sample_int(m))
    order = np.argsort(vr["id"])
    assert np.array_equal(fx.variant_id[vr["id"][order] - 1], g["vr_id"])
    assert np.allclose(vr["ratio"][order], g["vr_ratio"], rtol=1e-6)
    dist.barrier()
    if rank == 0:
        print("multi-GPU check ok: world=%d tau=%r b'Ab sharded=%.15g single=%.15g (sharded product, batched product, PCG, fit, "
              "variance ratio vs single GPU / reference golden)" % (world, glmm["tau"], inv[0], inv[1]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
