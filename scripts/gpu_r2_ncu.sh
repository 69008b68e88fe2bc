#!/bin/bash
# round 2: ncu metrics of one generation kernel on a short horizon.  `--set full` (and any multi-section capture) dies with
# LaunchFailed on its third replay pass of the layer-pipelined kernels, so the metrics are collected in small groups of one
# or two passes each; every group is its own run of the same command.   usage: gpu_r2_ncu.sh <tag> <kernel regex> <bench args...>
mkdir -p gpurun_out
TAG=$1; KRE=$2; shift 2
CMD="python bench.py $* --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
run() { # name, metrics
  timeout 200 ncu --clock-control none --metrics "$2" -k regex:$KRE -s 1 -c 1 --csv --log-file gpurun_out/${TAG}_$1.csv $CMD > gpurun_out/ncu_$1_${TAG}.log 2>&1
  echo "== $TAG $1 rc=$? rows=$(grep -c $KRE gpurun_out/${TAG}_$1.csv)"
}
run dram "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_elapsed.max,lts__t_bytes.sum"
run issue "smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__cycles_active.avg"
run pipes "sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_xu.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum"
run smem "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum"
run stall "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"
cat gpurun_out/${TAG}_*.csv | grep $KRE | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
