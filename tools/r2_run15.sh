cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_fused_layer_gpu.py -x -q -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_run15_tests_a.log; grep -c "few-rows" gpurun_out/r2_run15_tests_a.log; tail -4 gpurun_out/r2_run15_tests_a.log | cut -c1-220
if grep -q "passed" gpurun_out/r2_run15_tests_a.log && ! grep -q "failed\|error" gpurun_out/r2_run15_tests_a.log; then
  timeout 1800 python -m pytest tests/test_model_gpu.py tests/test_large_dims_gpu.py tests/test_zz_decode_options_gpu.py tests/test_zz_transcribe_batch_gpu.py tests/test_timing_gpu.py -x -q -s 2>&1 | grep -i "few-rows\|passed\|failed\|error\|assert" | tail -14 > gpurun_out/r2_run15_tests_b.log; cat gpurun_out/r2_run15_tests_b.log | cut -c1-220
fi
timeout 500 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline 2> gpurun_out/r2_run15_c4.err > gpurun_out/r2_run15_c4.json
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_run15_c4.json').read().strip().splitlines()[-1])
    print('c4 RTFx=%.1f e2e=%.1f launches=%d us/iter=%.1f frac=%.3f batch16=%.1f' % (l['value'], l['e2e']['value'], l['gpu_launches'], l['roofline']['us_per_iteration'], l['roofline']['frac'], l['transcribe_batch']['value']), l['phases_ms_per_step'])
except Exception as e:
    print('c4 FAILED', e); print(open('gpurun_out/r2_run15_c4.err').read()[-1500:])
PY
timeout 400 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2> gpurun_out/r2_run15_c2.err > gpurun_out/r2_run15_c2.json
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_run15_c2.json').read().strip().splitlines()[-1])
    print('c2 RTFx=%.1f e2e=%.1f ms/step=%.1f launches=%d decode_step=%.3fms frac=%.3f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline_decode_step']['ms_per_decode_step'], l['roofline_decode_step']['frac']), l['phases_ms_per_step'])
except Exception as e:
    print('c2 FAILED', e); print(open('gpurun_out/r2_run15_c2.err').read()[-1500:])
PY
lst() {  # name, model, batch, dtype
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_$1.csv python tools/profile_step.py --model $2 --batch $3 --beam 1 --dtype $4 --decode-steps 12 > gpurun_out/ncu_$1.log 2>&1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_$1.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
t=collections.Counter(); n=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    k=r[ki].split('(')[0].replace('void ','').replace('wb::','')[:46]; t[k]+=v; n[k]+=1
print('== $1')
for k,v in t.most_common(7): print('%-48s %5d launches %9.1f us total %7.2f us each'%(k,n[k],v,v/n[k]))
PY
}
lst turbo_b1_stack2 turbo 1 fp16
lst base_en_b32_stack2 base.en 32 fp16
# source-level counters of the one-launch stack (turbo, one audio)
timeout 420 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dec_rows -s 8 -c 1 \
  -o gpurun_out/r2_dec_rows_stack_turbo_b1 -f python tools/profile_step.py --model turbo --batch 1 --beam 1 --dtype fp16 --decode-steps 12 > gpurun_out/ncu_dec_rows.log 2>&1
tail -2 gpurun_out/ncu_dec_rows.log
