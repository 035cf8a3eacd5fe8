#!/bin/bash
# usage (on the GPU box): tools/round_artifacts.sh r01   -- bench line, ncu launch list, ncu --set full of the emit kernel
tag=${1:-r01}
python bench.py > gpurun_out/bench_${tag}.log 2>&1; tail -1 gpurun_out/bench_${tag}.log > gpurun_out/bench_${tag}.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:emit_kernel -s 3 -c 1 -o gpurun_out/${tag}_emit_full -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"smooth_probe|smooth_apply|smooth_clear|count_kernel" -s 4 -c 4 -o gpurun_out/${tag}_others_full -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2_${tag}.log 2>&1
tail -1 gpurun_out/bench_${tag}.json | cut -c1-400
