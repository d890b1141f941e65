cd $GRAFT_REPO_ROOT
# few-rows form generalised to <= 32 rows; self-attention A/B after removing the division from the gather kernel
timeout 900 python -m pytest tests/test_fused_layer_gpu.py "tests/test_model_gpu.py::test_batched_beam_equals_per_audio" tests/test_zz_kv_layout_gpu.py -x -q -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_run13_tests_a.log; cat gpurun_out/r2_run13_tests_a.log | cut -c1-200
if grep -q "passed" gpurun_out/r2_run13_tests_a.log && ! grep -q "failed\|error" gpurun_out/r2_run13_tests_a.log; then
  timeout 1800 python -m pytest tests/test_model_gpu.py tests/test_large_dims_gpu.py tests/test_zz_decode_options_gpu.py tests/test_zz_transcribe_batch_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r2_run13_tests_b.log; cat gpurun_out/r2_run13_tests_b.log
fi
for sa in 0 1; do
  WB200_SATTN_TMA=$sa timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --breakdown --breakdown-ids 2 2> gpurun_out/r2_run13_bench_$sa.err > gpurun_out/r2_run13_bench_$sa.json
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_run13_bench_$sa.json').read().strip().splitlines()[-1]); b=l['breakdown']
    print('sattn_tma=$sa RTFx=%.1f e2e=%.1f ms/step=%.1f launches=%d decode_step=%.3fms hbm_frac=%.3f xattn_frac=%.3f' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches'], l['roofline_decode_step']['ms_per_decode_step'], l['roofline_decode_step']['frac'], l['roofline']['frac']), {k:(round(v['ms'],1), v['launches']) for k,v in b.items()}, l['phases_ms_per_step'])
except Exception as e:
    print('sattn=$sa FAILED', e); print(open('gpurun_out/r2_run13_bench_$sa.err').read()[-1500:])
PY
done
# c2 (base.en greedy, 32 rows): few-rows form on / off
for rows in 1 0; do
  WB200_FUSED_ROWS=$rows timeout 400 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2> gpurun_out/r2_run13_c2_rows$rows.err > gpurun_out/r2_run13_c2_rows$rows.json
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_run13_c2_rows$rows.json').read().strip().splitlines()[-1])
    print('c2 rows=$rows RTFx=%.1f e2e=%.1f ms/step=%.1f launches=%d' % (l['value'], l['e2e']['value'], l['ms_per_step'], l['gpu_launches']), l.get('roofline_decode_step'), l['phases_ms_per_step'])
except Exception as e:
    print('c2 rows=$rows FAILED', e); print(open('gpurun_out/r2_run13_c2_rows$rows.err').read()[-1500:])
PY
done
# where does the beam-window self-attention kernel stall?  one launch at the mean history length, source-level counters
timeout 420 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:self_attention_tma -s 3520 -c 1 \
  -o gpurun_out/r2_self_attn_tma_midL -f python tools/profile_step.py --decode-steps 116 > gpurun_out/ncu_sattn.log 2>&1
tail -3 gpurun_out/ncu_sattn.log
ls -la gpurun_out/*.ncu-rep | tail -3
