#!/bin/bash
# A/B programmatic dependent launch and the self-attention ring configuration on the C3 decode.
mkdir -p gpurun_out
: > gpurun_out/selfattn_sweep2.txt
run() {
  WB200_PDL=$1 WB200_SA_CFG=$2 timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --breakdown --breakdown-ids 2 2>gpurun_out/sweep_err.txt | python -c "
import json,sys
l=json.loads(sys.stdin.readlines()[-1]); b=l['breakdown']['self_attention']
print('pdl=$1 cfg=$2 self_attn_total_ms=%.1f avg_us=%.1f  RTFx=%.1f e2e=%.1f' % (b['ms'], 1e3*b['ms']/b['launches'], l['value'], l['e2e']['value']))" >> gpurun_out/selfattn_sweep2.txt 2>&1
}
run 0 2,4,4
run 1 2,4,4
run 1 2,3,4
run 1 2,5,4
run 1 1,6,4
run 1 1,8,4
run 1 2,4,2
run 1 2,4,5
cat gpurun_out/selfattn_sweep2.txt
tail -5 gpurun_out/sweep_err.txt
