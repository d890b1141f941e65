#!/bin/bash
# Round-2 evidence run on one B200: ncu launch list of one short step + `--set full` captures of the decode kernels
# at the headline shape (large-v3, batch 64, beam 5, bf16) and of the small kernels, reduced to CSV selections.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SEL='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__bytes_read.sum,|dram__bytes_write.sum,|dram__cycles_active|gpu__dram_throughput|sm__pipe_tensor|sm__inst_executed_pipe_tensor|sm__warps_active|launch__registers_per_thread|launch__grid_size|launch__block_size|lts__t_sector_hit_rate|sm__throughput|l1tex__data_pipe|smsp__cycles_active|launch__occupancy|lts__throughput|sm__cycles_elapsed.avg '
cap() {  # name, kernel regex, skip, count, decode steps
  timeout 420 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 \
    -o gpurun_out/r2_$1 -f python tools/profile_step.py --decode-steps $5 > gpurun_out/ncu_$1.log 2>&1
  if [ -f gpurun_out/r2_$1.ncu-rep ]; then
    ncu -i gpurun_out/r2_$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_select.py > gpurun_out/r2_$1_ncu_full_selected.csv
    echo "$1: $(wc -l < gpurun_out/r2_$1_ncu_full_selected.csv) lines"
  else
    echo "$1: capture failed"; tail -5 gpurun_out/ncu_$1.log
  fi
}
timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launches.csv python tools/profile_step.py --decode-steps 6 > gpurun_out/ncu_list.log 2>&1
gzip -f gpurun_out/r2_launches.csv; ls -la gpurun_out/r2_launches.csv.gz
cap dec_layer dec_layer_kernel 20 4 8
cap cross_attn_tma cross_attention_tma 40 2 8
cap self_attn_midL self_attention_kernel 3600 2 120
cap select 'filter_topk|beam_update|range_softmax|greedy_update' 4 4 8
cap frontend 'log_mel|layernorm_kernel|embed_kernel|transpose_to16' 0 4 3
cap enc_attn enc_attention 4 1 3
timeout 300 ncu --set full --clock-control none -k regex:'median_filter|dtw_' -c 4 -o gpurun_out/r2_timing -f \
  python -m pytest tests/test_timing_gpu.py -q -x > gpurun_out/ncu_timing.log 2>&1
[ -f gpurun_out/r2_timing.ncu-rep ] && ncu -i gpurun_out/r2_timing.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_select.py > gpurun_out/r2_timing_ncu_full_selected.csv
ls -la gpurun_out/*.ncu-rep | head -20
