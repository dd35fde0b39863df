import sys, numpy as np
sys.path.insert(0, '.')
import saigegds_b200 as sg
ctx = sg.Context(0)
N, M = 430000, 100000
ctx.store_synthetic(N, M, M, 0, seed=200, missing_rate=0.005)
rng = np.random.default_rng(1)
for k in (1, 4, 30):
    B = np.asfortranarray(rng.standard_normal((N, k)))
    d_b = ctx.device_vector(B.reshape(-1, order="F"))
    d_out = ctx.device_empty(8 * N * k)
    for _ in range(3): ctx.grm_mv_device(d_b, d_out, k)
    ms = ctx.time_products_device(d_b, d_out, k, 10)
    print("k=%d: %.3f ms per call, %.3f ms per column" % (k, ms / 10, ms / 10 / k))
    d_b.free(); d_out.free()
