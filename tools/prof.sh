#!/bin/bash
# usage: tools/prof.sh <tag> [bench flags]   -- ncu --set full capture of the emit kernel (one launch) + launch list
tag=$1; shift
ncu --set full --clock-control none --import-source on -k regex:emit_kernel -s 3 -c 1 -o gpurun_out/prof_${tag} -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_${tag}.log 2>&1
tail -2 gpurun_out/ncu_${tag}.log | cut -c1-300
