cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -k "selection or sampling or free_running or end_to_end or batched" 2>&1 | tail -6 > gpurun_out/r2_run10_tests.log; cat gpurun_out/r2_run10_tests.log
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown --breakdown-ids 1,2,6,8 2> gpurun_out/r2_run10_bench.err > gpurun_out/r2_run10_bench.json
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_run10_bench.json').read().strip().splitlines()[-1]); b=l['breakdown']
    print('RTFx=%.1f e2e=%.1f ms/step=%.1f launches=%d decode_step=%.3fms hbm_frac=%.3f xattn_frac=%.3f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline_decode_step']['ms_per_decode_step'], l['roofline_decode_step']['frac'], l['roofline']['frac']), {k:(round(v['ms'],1), v['launches']) for k,v in b.items()}, l['phases_ms_per_step'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2_run10_bench.err').read()[-1500:])
PY
