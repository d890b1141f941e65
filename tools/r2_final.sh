#!/bin/bash
# Round-2 closing run on one B200: the full GPU suite, smoke(), the three single-GPU bench lines, the launch list of the
# headline step and --set full captures of the kernels new in this round's second half.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_final_tests.log; cat gpurun_out/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -3 gpurun_out/r2_final_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 2> gpurun_out/r2_final_bench_c3.err > gpurun_out/r2_final_bench_c3.json; tail -c 600 gpurun_out/r2_final_bench_c3.json; echo
timeout 400 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2> gpurun_out/r2_final_bench_c2.err > gpurun_out/r2_final_bench_c2.json; tail -c 400 gpurun_out/r2_final_bench_c2.json; echo
timeout 500 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline 2> gpurun_out/r2_final_bench_c4.err > gpurun_out/r2_final_bench_c4.json; tail -c 400 gpurun_out/r2_final_bench_c4.json; echo
timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_final_launches.csv python tools/profile_step.py --decode-steps 6 > gpurun_out/ncu_final_launches.log 2>&1
gzip -f gpurun_out/r2_final_launches.csv
cap() {  # name, kernel regex, skip, count, extra args of profile_step
  timeout 420 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 \
    -o gpurun_out/r2_$1 -f python tools/profile_step.py $5 > gpurun_out/ncu_$1.log 2>&1
  if [ -f gpurun_out/r2_$1.ncu-rep ]; then
    ncu -i gpurun_out/r2_$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_select.py > gpurun_out/r2_$1_ncu_full_selected.csv
    echo "$1: $(wc -l < gpurun_out/r2_$1_ncu_full_selected.csv) lines"
  else
    echo "$1: capture failed"; tail -5 gpurun_out/ncu_$1.log
  fi
}
cap select_c3 'filter_topk|beam_update' 4 2 "--decode-steps 4"
cap select_cluster_b1 'filter_topk' 4 1 "--model turbo --batch 1 --beam 1 --dtype fp16 --decode-steps 8"
cap dec_rows_base_en_b32 'dec_rows' 4 1 "--model base.en --batch 32 --beam 1 --dtype fp16 --decode-steps 8"
