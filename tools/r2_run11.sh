cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_xattn_tma_gpu.py tests/test_zz_kv_layout_gpu.py -x -q -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r2_run11_tests_a.log; cat gpurun_out/r2_run11_tests_a.log
if grep -q "passed" gpurun_out/r2_run11_tests_a.log && ! grep -q "failed\|error" gpurun_out/r2_run11_tests_a.log; then
  timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_fused_layer_gpu.py tests/test_large_dims_gpu.py tests/test_zz_decode_options_gpu.py tests/test_zz_transcribe_batch_gpu.py -x -q 2>&1 | tail -6 > gpurun_out/r2_run11_tests_b.log; cat gpurun_out/r2_run11_tests_b.log
fi
for sa in 1 0; do
  WB200_SATTN_TMA=$sa timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown --breakdown-ids 1,2,6,8 2> gpurun_out/r2_run11_bench_$sa.err > gpurun_out/r2_run11_bench_$sa.json
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_run11_bench_$sa.json').read().strip().splitlines()[-1]); b=l['breakdown']
    print('sattn_tma=$sa RTFx=%.1f e2e=%.1f ms/step=%.1f launches=%d decode_step=%.3fms hbm_frac=%.3f xattn_frac=%.3f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline_decode_step']['ms_per_decode_step'], l['roofline_decode_step']['frac'], l['roofline']['frac']), {k:(round(v['ms'],1), v['launches']) for k,v in b.items()}, l['phases_ms_per_step'])
except Exception as e:
    print('sattn=$sa FAILED', e); print(open('gpurun_out/r2_run11_bench_$sa.err').read()[-1500:])
PY
done
# where does a batch-1 decoder iteration (c4) go?  per-kernel device time of a short turbo greedy decode
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launches_turbo_b1.csv python tools/profile_step.py --model turbo --batch 1 --beam 1 --dtype fp16 --decode-steps 12 > gpurun_out/ncu_turbo_b1.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_turbo_b1.csv')) if len(r)>5]
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
