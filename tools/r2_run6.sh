cd $GRAFT_REPO_ROOT
timeout 120 tools/bin/bench_dec_layer 320 1280 200 > gpurun_out/r2_dec_layer_trace_v3.txt 2>&1; cat gpurun_out/r2_dec_layer_trace_v3.txt
timeout 600 python -m pytest tests/test_fused_layer_gpu.py tests/test_xattn_tma_gpu.py tests/test_zz_kv_layout_gpu.py -x -q -s 2>&1 | tail -25 > gpurun_out/r2_run6_tests_a.log; cat gpurun_out/r2_run6_tests_a.log
if grep -q "passed" gpurun_out/r2_run6_tests_a.log && ! grep -q "failed\|error" gpurun_out/r2_run6_tests_a.log; then
  timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_large_dims_gpu.py tests/test_zz_decode_options_gpu.py tests/test_zz_transcribe_batch_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r2_run6_tests_b.log; cat gpurun_out/r2_run6_tests_b.log
fi
for cfg in "1 1" "1 0" "0 1"; do
  set -- $cfg
  WB200_FUSED_LAYER=$1 WB200_XATTN_TMA=$2 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown --breakdown-ids 1,2,3,5,8 2> gpurun_out/r2_run6_bench_$1$2.err > gpurun_out/r2_run6_bench_$1$2.json
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_run6_bench_$1$2.json').read().strip().splitlines()[-1]); b=l['breakdown']
    print('fused=$1 xattn_tma=$2 RTFx=%.1f e2e=%.1f ms/step=%.1f launches=%d decode_step=%.3fms hbm_frac=%.3f xattn_frac=%.3f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline_decode_step']['ms_per_decode_step'], l['roofline_decode_step']['frac'], l['roofline']['frac']), {k:(round(v['ms'],1), v['launches']) for k,v in b.items()})
except Exception as e:
    print('fused=$1 xattn=$2 FAILED', e); print(open('gpurun_out/r2_run6_bench_$1$2.err').read()[-1500:])
PY
done
