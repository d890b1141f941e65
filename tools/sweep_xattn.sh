#!/bin/bash
# Sweep the cross-attention ring depth / key splits on the C3 shape; each combination is one short bench.py run
# (the kernel's per-launch time comes from bench.py's profiled step).  Output: gpurun_out/xattn_sweep.txt
mkdir -p gpurun_out
: > gpurun_out/xattn_sweep.txt
for st in 3 4 6; do
  for sp in 1 2 3; do
    WB200_XATTN_STAGES=$st WB200_XATTN_SPLITS=$sp timeout 300 python bench.py --steps 1 --warmup 1 \
      --decode-steps 24 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readlines()[-1]); r=l['roofline']
print('stages=$st splits=$sp avg_launch_us=%.1f GB/s=%.0f frac=%.3f' % (r['avg_launch_ms']*1e3, r['achieved'], r['frac']))" >> gpurun_out/xattn_sweep.txt
  done
done
cat gpurun_out/xattn_sweep.txt
