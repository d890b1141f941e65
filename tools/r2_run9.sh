cd $GRAFT_REPO_ROOT
timeout 120 tools/bin/bench_dec_layer 320 1280 200 > gpurun_out/r2_dec_layer_trace_v6.txt 2>&1; cat gpurun_out/r2_dec_layer_trace_v6.txt
timeout 600 python -m pytest tests/test_fused_layer_gpu.py -x -q -s 2>&1 | tail -14 > gpurun_out/r2_run9_tests_a.log; cat gpurun_out/r2_run9_tests_a.log
if grep -q "passed" gpurun_out/r2_run9_tests_a.log && ! grep -q "failed\|error" gpurun_out/r2_run9_tests_a.log; then
  timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_large_dims_gpu.py -x -q -k "selection or teacher or end_to_end or batched or baseline_dims" 2>&1 | tail -5 > gpurun_out/r2_run9_tests_b.log; cat gpurun_out/r2_run9_tests_b.log
fi
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown --breakdown-ids 1,2,8 2> gpurun_out/r2_run9_bench.err > gpurun_out/r2_run9_bench.json
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_run9_bench.json').read().strip().splitlines()[-1]); b=l['breakdown']
    print('RTFx=%.1f e2e=%.1f ms/step=%.1f launches=%d decode_step=%.3fms hbm_frac=%.3f xattn_frac=%.3f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline_decode_step']['ms_per_decode_step'], l['roofline_decode_step']['frac'], l['roofline']['frac']), {k:(round(v['ms'],1), v['launches']) for k,v in b.items()}, l['phases_ms_per_step'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/r2_run9_bench.err').read()[-1500:])
PY
