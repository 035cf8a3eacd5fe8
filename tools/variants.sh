#!/bin/bash
# usage: tools/variants.sh [bench flags --] v1 v2 ...  -- bench with tmc2-rs_b200/variant_<v>.so; results are NOT checked
flags=""
if [ "$1" == "--no-smoothing" ]; then flags="--no-smoothing"; shift; fi
for v in "$@"; do
  TMC2_LIB=$PWD/tmc2-rs_b200/variant_$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $flags 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$v $flags', 'ms/step %.3f'%d['ms_per_step'], {k:round(x,3) for k,x in d['roofline']['stage_ms'].items() if x})
except Exception as e: print('$v fail', e)"
done
