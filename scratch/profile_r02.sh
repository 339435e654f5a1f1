#!/bin/bash
# Round-2 profile captures (run on the GPU box through gpurun; outputs land in gpurun_out/).
set -u
mkdir -p gpurun_out
B="python bench.py --images 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_active.avg,smsp__cycles_active.avg,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,sm__cycles_elapsed.avg.per_second
# 1. launch list of one bench command (cold-cache, serialised): the kernels' SHARES of a step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_raw.csv $B > gpurun_out/r02_launches_bench.log 2>&1
# 2. one steady-state launch of every kernel of a decode step + tail, full sections
timeout 1200 ncu --set full --import-source on --clock-control none \
  -k regex:"wino_gemm_tc|lstm_cell_wino|wino_input_kernel|conv_gemm_tc|head_gather|head_drt|head_finish|semantic_feat|sgemm_nt|attention_update|sample_actions|score_pairs_g8|reduce_pairs|prep_paths_kernel" \
  --launch-skip 60 --launch-count 22 -f -o gpurun_out/r02_step_kernels $B > gpurun_out/r02_step_kernels.log 2>&1
ncu -i gpurun_out/r02_step_kernels.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_step_kernels_raw.csv 2>/dev/null
# 3. the once-per-wave x-gate convolution (Winograd F(2x2)): input transform, GEMM (16 positions), output transform --
#    the first three matching launches of the second wave (a wave has 2 + 2 + 2 + 15 of them)
timeout 900 ncu --set full --import-source on --clock-control none \
  -k regex:"wino_input22|wino_output22|wino_gemm_tc" --launch-skip 21 --launch-count 3 -f -o gpurun_out/r02_xgate_kernels $B > gpurun_out/r02_xgate_kernels.log 2>&1
ncu -i gpurun_out/r02_xgate_kernels.ncu-rep --page raw --csv --metrics $M > gpurun_out/r02_xgate_kernels_raw.csv 2>/dev/null
for k in wino_gemm_tc lstm_cell_wino wino_input_kernel conv_gemm_tc semantic_feat; do
  ncu -i gpurun_out/r02_step_kernels.ncu-rep --page details --kernel-name regex:$k 2>/dev/null | head -220 > gpurun_out/r02_${k}_ncu_details.txt
done
tail -2 gpurun_out/r02_step_kernels.log; tail -2 gpurun_out/r02_xgate_kernels.log
