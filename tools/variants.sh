for v in libtmc2gpu variant_w4c3 variant_w8c4 variant_w4c4 variant_w2c3; do
  for m in "--no-smoothing" "--no-smoothing --two-pass" ""; do
    TMC2_LIB=$PWD/tmc2-rs_b200/$v.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $m 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v','$m', 'ms/step %.3f'%d['ms_per_step'], 'unpack %.3f'%d['roofline']['stage_ms']['unpack'], 'filt %.3f'%d['roofline']['stage_ms']['geometry_smoothing'],'frac %.3f'%d['roofline']['frac'], 'e2e %.0fM'%(d['e2e']['value']/1e6))"
  done
done
