cd $GRAFT_REPO_ROOT
timeout 120 tools/bin/bench_dec_layer 320 1280 200 > gpurun_out/r2_dec_layer_trace.txt 2>&1; cat gpurun_out/r2_dec_layer_trace.txt
timeout 300 python -m pytest tests/test_fused_layer_gpu.py -x -q -s 2>&1 | tail -12
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_c3_newbench.json 2> gpurun_out/r2_bench_c3_newbench.err; tail -c 3000 gpurun_out/r2_bench_c3_newbench.json; tail -5 gpurun_out/r2_bench_c3_newbench.err
