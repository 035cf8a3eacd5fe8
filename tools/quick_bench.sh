#!/bin/bash
# usage: tools/quick_bench.sh <tag>   -- GPU tests + two bench lines (smoothing off / on), summarised
tag=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/t_$tag.log 2>&1; tail -4 gpurun_out/t_$tag.log
for m in "--no-smoothing" ""; do
  n=$(echo "$m" | tr -d " -"); f=gpurun_out/b_${tag}_${n:-smooth}.log
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $m > $f 2>&1
  tail -1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$m |', 'ms/step %.3f'%d['ms_per_step'], {k:round(v,3) for k,v in d['roofline']['stage_ms'].items() if v}, 'frac %.3f'%d['roofline']['frac'], 'e2e %.0fM pts/s'%(d['e2e']['value']/1e6), 'e2e ms %.1f'%d['e2e']['ms_per_step'], 'value %.1fG'%(d['value']/1e9))
except Exception as e: print('fail', e); print(open('$f').read()[-2000:])
"
done
