#!/bin/bash
# usage (on the GPU box): tools/round_artifacts.sh r02   -- bench lines, ncu launch list, ncu --set full of the emit kernel (both
# instantiations) and of the other kernels.  Everything lands in gpurun_out/; the summaries are made here with tools/ncu_summary.py.
tag=${1:-r02}
timeout 300 python bench.py > gpurun_out/bench_${tag}.log 2>&1; tail -1 gpurun_out/bench_${tag}.log > gpurun_out/${tag}_bench_line.json
timeout 200 python bench.py --no-smoothing --no-cpu-baseline --quick > gpurun_out/bench_${tag}_nosmoothing.log 2>&1; tail -1 gpurun_out/bench_${tag}_nosmoothing.log > gpurun_out/${tag}_bench_line_nosmoothing.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.log 2>&1; tail -1 gpurun_out/bench_${tag}_ref.log > gpurun_out/${tag}_bench_line_reference.json
for c in c1 c3 c4; do timeout 200 python bench.py --config $c --steps 5 --no-cpu-baseline --quick 2>/dev/null | tail -1; done > gpurun_out/${tag}_other_configs.jsonl
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/plain_${tag}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:emit_kernel -s 3 -c 1 -o gpurun_out/${tag}_emit_full -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_full_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:emit_kernel -s 3 -c 1 -o gpurun_out/${tag}_emit_nosmoothing_full -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick --no-smoothing > gpurun_out/ncu_full_ns_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"smooth_probe|smooth_apply|smooth_clear|count_kernel" -s 4 -c 4 -o gpurun_out/${tag}_others_full -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > gpurun_out/ncu_full2_${tag}.log 2>&1
cut -c1-300 gpurun_out/${tag}_bench_line.json
