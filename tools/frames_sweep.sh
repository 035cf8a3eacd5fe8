#!/bin/bash
# per-frame stage times for different GOF lengths (a short GOF stays L2-resident across steps)
for f in 2 4 8 32; do
  python bench.py --frames $f --no-smoothing --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | F=$f python -c "
import json,sys,os
f=int(os.environ['F']); d=json.loads(sys.stdin.read()); st=d['roofline']['stage_ms']
print('frames %d unpack/frame %.2f us count/frame %.2f us b2p/frame %.2f us total/frame %.2f us'%(f, 1e3*st['unpack']/f, 1e3*st['count_scan']/f, 1e3*st['block_to_patch']/f, 1e3*d['ms_per_step']/f))"
done
