cd $GRAFT_REPO_ROOT
# 1. the kernels new in this step first: few-rows fused layer, beam-window self attention v2, cluster top-K
timeout 900 python -m pytest tests/test_fused_layer_gpu.py tests/test_xattn_tma_gpu.py tests/test_zz_kv_layout_gpu.py -x -q -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_run12_tests_a.log; cat gpurun_out/r2_run12_tests_a.log
if grep -q "passed" gpurun_out/r2_run12_tests_a.log && ! grep -q "failed\|error" gpurun_out/r2_run12_tests_a.log; then
  timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_fused_layer_gpu.py --deselect tests/test_xattn_tma_gpu.py --deselect tests/test_zz_kv_layout_gpu.py 2>&1 | tail -8 > gpurun_out/r2_run12_tests_b.log; cat gpurun_out/r2_run12_tests_b.log
fi
for sa in 1 0; do
  WB200_SATTN_TMA=$sa timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown --breakdown-ids 2 2> gpurun_out/r2_run12_bench_$sa.err > gpurun_out/r2_run12_bench_$sa.json
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_run12_bench_$sa.json').read().strip().splitlines()[-1]); b=l['breakdown']
    print('sattn_tma=$sa RTFx=%.1f e2e=%.1f ms/step=%.1f launches=%d decode_step=%.3fms hbm_frac=%.3f xattn_frac=%.3f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline_decode_step']['ms_per_decode_step'], l['roofline_decode_step']['frac'], l['roofline']['frac']), {k:(round(v['ms'],1), v['launches']) for k,v in b.items()}, l['phases_ms_per_step'])
except Exception as e:
    print('sattn=$sa FAILED', e); print(open('gpurun_out/r2_run12_bench_$sa.err').read()[-1500:])
PY
done
# 2. c4 (turbo, one audio, greedy): few-rows form on / off
for rows in 1 0; do
  WB200_FUSED_ROWS=$rows timeout 500 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline 2> gpurun_out/r2_run12_c4_rows$rows.err > gpurun_out/r2_run12_c4_rows$rows.json
  tail -c 1800 gpurun_out/r2_run12_c4_rows$rows.json; echo; tail -3 gpurun_out/r2_run12_c4_rows$rows.err
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launches_turbo_b1_rows.csv python tools/profile_step.py --model turbo --batch 1 --beam 1 --dtype fp16 --decode-steps 12 > gpurun_out/ncu_turbo_b1_rows.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_turbo_b1_rows.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
t=collections.Counter(); n=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    k=r[ki].split('(')[0].replace('void ','').replace('wb::','')[:46]; t[k]+=v; n[k]+=1
for k,v in t.most_common(16): print('%-48s %5d launches %9.1f us total %7.2f us each'%(k,n[k],v,v/n[k]))
PY
