#!/bin/bash
# usage: tools/env_sweep.sh VAR v1 v2 ...  -- bench (smoothing on) with VAR set to each value
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | V="$var=$v" python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print(os.environ['V'], 'ms/step %.3f'%d['ms_per_step'], {k:round(x,3) for k,x in d['roofline']['stage_ms'].items() if x})"
done
