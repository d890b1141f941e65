#!/bin/bash
# Round-end evidence run on one B200: GPU test suite, bench line (with per-class breakdown), ncu launch list and
# one full-set capture of the self-attention kernel.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r1_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r1_gpu_tests_final.log
python bench.py --steps 3 --warmup 3 --breakdown > gpurun_out/r1_bench_c3_final.json 2> gpurun_out/bench_final.err
python -c "
import json
l=json.loads(open('gpurun_out/r1_bench_c3_final.json').read().strip().splitlines()[-1])
print('RTFx', l['value'], 'e2e', l['e2e']['value'], 'ms', l['ms_per_step'], 'roofline', l['roofline']['frac'], l['roofline']['avg_launch_ms'], 'cpu', l['cpu_baseline']['value'], l['clocks'])
for k,v in l['breakdown'].items(): print(k, v)
"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r1_launches_v4.csv python tools/profile_step.py --decode-steps 6 > gpurun_out/ncu_list.log 2>&1
gzip -f gpurun_out/r1_launches_v4.csv; ls -la gpurun_out/r1_launches_v4.csv.gz
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:self_attention -s 100 -c 2 \
  -o gpurun_out/r1_self_attn_v4 -f python tools/profile_step.py --decode-steps 120 > gpurun_out/ncu_sa.log 2>&1
ls -la gpurun_out/r1_self_attn_v4.ncu-rep
