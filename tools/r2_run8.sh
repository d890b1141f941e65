cd $GRAFT_REPO_ROOT
timeout 120 tools/bin/bench_dec_layer 320 1280 200 > gpurun_out/r2_dec_layer_trace_v5.txt 2>&1; head -9 gpurun_out/r2_dec_layer_trace_v5.txt
timeout 900 python bench.py --config c4 --steps 2 --warmup 2 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; tail -c 2500 gpurun_out/r2_bench_c4.json; tail -3 gpurun_out/r2_bench_c4.err
bash tools/r2_evidence.sh 2>&1 | tail -30
