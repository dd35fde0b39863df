# Fused-kernel tuning sweep: lag of phase B behind phase A (env SGB_FUSED_LAG)
mkdir -p gpurun_out
for lag in ${LAGS:-4 3 5 6 2}; do
  SGB_FUSED_LAG=$lag timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/sw.json 2> gpurun_out/sw.err || { echo "lag=$lag FAILED"; tail -3 gpurun_out/sw.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/sw.json')); k=d['roofline'].pop('kernels'); print('lag=$lag', round(d['value'],1), round(d['ms_per_step'],3), {a[:14]:round(b['ms_per_launch'],3) for a,b in k.items() if b['ms_per_launch']>0.05})"
done
