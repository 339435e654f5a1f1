#!/bin/bash
# Round-2 profile captures (run on the GPU box through gpurun; outputs land in gpurun_out/).
set -u
mkdir -p gpurun_out
B="python bench.py --images 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
# 1. launch list of one bench step (cold-cache, serialised): the kernels' SHARES of a step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_raw.csv $B > gpurun_out/r02_launches_bench.log 2>&1
# 2. one steady-state launch of every kernel of a decode step + tail, full sections
timeout 1200 ncu --set full --import-source on --clock-control none \
  -k regex:"wino_gemm_tc|lstm_cell_wino|wino_input|conv_gemm_tc|head_gather|head_drt|head_finish|semantic_feat|sgemm_nt|attention_update|sample_actions|score_pairs_g8|reduce_pairs|prep_paths_kernel" \
  --launch-skip 60 --launch-count 22 -o gpurun_out/r02_step_kernels $B > gpurun_out/r02_step_kernels.log 2>&1
ncu -i gpurun_out/r02_step_kernels.ncu-rep --page raw --csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_active.avg,smsp__cycles_active.avg,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,sm__cycles_elapsed.avg.per_second \
  > gpurun_out/r02_step_kernels_raw.csv 2>/dev/null
tail -2 gpurun_out/r02_step_kernels.log
