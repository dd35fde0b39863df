mkdir -p gpurun_out
for cfg in "1 1 4" "1 1 3" "0 2 4" "0 1 4" "1 2 4"; do
  set -- $cfg
  SGB_SPARSE_FORK=$1 SGB_SPARSE_GRID_MULT=$2 SGB_DOTS_STAGES=$3 timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/sw.json 2> gpurun_out/sw.err
  python -c "
import json; d=json.load(open('gpurun_out/sw.json')); k=d['roofline'].pop('kernels'); print('fork=$1 mult=$2 stages=$3', round(d['value'],1), round(d['ms_per_step'],3), {a[:12]:round(b['ms_per_launch'],3) for a,b in k.items() if b['ms_per_launch']>0.3})"
done
