#!/bin/bash
# one pass of BASELINE config $1 under ncu with the handful of metrics of tools/ncu_to_json.py (a few replays per kernel, so the
# whole pass of the large configurations fits in minutes) -> gpurun_out/$2.ncu-rep
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:"k_vv_|k_vert_volume|k_resolve_|k_seg_prep|k_seg_cells|k_cell3_flags|k_cell_planes|k_block_list|k_fill_tiles|k_build_list" -o gpurun_out/$2 -f python tools/profile_pass.py $1 > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log
