#!/bin/bash
# First GPU call of the next round: validate and A/B the switches that were built in round 1 but never measured
# (profiles/r1_summary.md section 5).  Every run is under `timeout`: WB200_GEMM_EARLY_B has never executed on
# hardware and a pipeline bug there would hang rather than fail.
#   gpurun --timeout 1500 -- 'bash tools/ab_switches.sh'
mkdir -p gpurun_out
out=gpurun_out/ab_switches.txt
: > $out
bench() {   # label, env assignments..., -- extra bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 240 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --breakdown --breakdown-ids 1,2,3 "$@" \
    2>gpurun_out/ab_err.txt | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.readlines()[-1]); b=l['breakdown']
    print('$label: RTFx=%.1f e2e=%.1f  cross=%.1fus self=%.1fus gemm=%.1fus' % (l['value'], l['e2e']['value'],
          1e3*b['cross_attention']['ms']/b['cross_attention']['launches'], 1e3*b['self_attention']['ms']/b['self_attention']['launches'],
          1e3*b['gemm']['ms']/b['gemm']['launches']))
except Exception as e:
    print('$label: FAILED', e)" >> $out
}
# 1. correctness of each switch on the small models (the staged xfail tests + the decode parity tests under the switch)
timeout 600 python -m pytest tests/test_zz_kv_layout_gpu.py tests/test_zz_transcribe_batch_gpu.py tests/test_zz_decode_options_gpu.py -q -rxX 2>&1 | tail -15 >> $out
WB200_KV_HEAD_MAJOR=1 timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "teacher_forced or end_to_end or batched_beam or alignment" 2>&1 | tail -3 >> $out
WB200_GEMM_EARLY_B=1 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_primitives_gpu.py -q -x -k "teacher_forced or end_to_end or batched_beam or linear" 2>&1 | tail -3 >> $out
# 2. throughput on C3
bench baseline --
bench head_major WB200_KV_HEAD_MAJOR=1 --
bench early_b WB200_GEMM_EARLY_B=1 --
bench streams2 -- --decode-streams 2
bench streams4 -- --decode-streams 4
bench all WB200_KV_HEAD_MAJOR=1 WB200_GEMM_EARLY_B=1 -- --decode-streams 2
cat $out
