# Fused-kernel tuning sweep; each config is lag:poll_ns (SGB_FUSED_LAG, SGB_FUSED_POLL_NS)
mkdir -p gpurun_out
for cfg in ${CFGS:-6:100 5:100 7:100}; do
  IFS=: read lag poll <<< "$cfg"
  SGB_FUSED_LAG=$lag SGB_FUSED_POLL_NS=$poll timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/sw.json 2> gpurun_out/sw.err || { echo "cfg=$cfg FAILED"; tail -3 gpurun_out/sw.err; continue; }
  python -c "
import json; d=json.load(open('gpurun_out/sw.json')); k=d['roofline'].pop('kernels'); print('cfg=$cfg', round(d['value'],1), round(d['ms_per_step'],3), {a[:14]:round(b['ms_per_launch'],3) for a,b in k.items() if b['ms_per_launch']>0.05})"
done
