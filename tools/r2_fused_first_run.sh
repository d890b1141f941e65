cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_fused_layer_gpu.py -x -q -s 2>&1 | tail -40 > gpurun_out/r2_fused_tests.log
tail -25 gpurun_out/r2_fused_tests.log
if grep -q "passed" gpurun_out/r2_fused_tests.log && ! grep -q "failed" gpurun_out/r2_fused_tests.log; then
  timeout 900 python -m pytest tests/test_model_gpu.py tests/test_large_dims_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r2_fused_model_tests.log
  tail -8 gpurun_out/r2_fused_model_tests.log
  for f in 1 0; do
    WB200_FUSED_LAYER=$f timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --breakdown --breakdown-ids 1,2,3,5,8 2> gpurun_out/r2_bench_fused$f.err > gpurun_out/r2_bench_fused$f.json
    python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_fused$f.json').read().strip().splitlines()[-1]); b=l['breakdown']
    print('fused=$f RTFx=%.1f e2e=%.1f ms/step=%.1f launches=%d' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches']), {k:(round(v['ms'],1), v['launches']) for k,v in b.items()})
except Exception as e:
    print('fused=$f FAILED', e); print(open('gpurun_out/r2_bench_fused$f.err').read()[-1500:])
PY
  done
fi
