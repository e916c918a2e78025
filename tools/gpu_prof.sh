#!/bin/bash
# developer tool: ncu metrics for both kernel variants (fast mode, Cornell A, 1200x1200 @ 64 spp)
cat > /tmp/run_variant.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import tpt_b200 as T, common
kern = int(sys.argv[1]); spp = int(sys.argv[2]); fov=float(sys.argv[3]); depth=int(sys.argv[4])
sc = T.Scene(common.host_scene(T, "cornell_box"))
cam = T.cornell_camera(1200, 1200, fov=fov)
for i in range(2):
    st = sc.render_device(cam, T.make_params(1200, 1200, spp, depth, mode=T.MODE_FAST, seed=1, kernel=kern))
print(kern, st["render_ms"], st["paths"]/st["render_ms"]/1e3)
PY
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_cbu.sum,sm__inst_executed_pipe_adu.sum
for k in 0 1; do
  ncu --metrics $M --clock-control none -k regex:render_ -s 1 -c 1 --csv --log-file gpurun_out/prof_k${k}.csv python /tmp/run_variant.py $k ${1:-64} 90 15 > gpurun_out/prof_k${k}.log 2>&1
done
tail -2 gpurun_out/prof_k*.log
