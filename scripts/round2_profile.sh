#!/bin/bash
# Source-level ncu captures of the kernels round 1 did not get to (about 1 min of GPU time each); read the reports
# with scripts/ncu_stall_by_role.py.  Usage on a GPU box: bash scripts/round2_profile.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEC="--section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section InstructionStats"
timeout 200 ncu $SEC --import-source on --clock-control none -k regex:"stage1|side_conv|tail_fwd" -c 6 -f \
    -o gpurun_out/r02_first_side_tail python scripts/one_forward.py > gpurun_out/r02_first_side_tail.log 2>&1
timeout 300 ncu $SEC --import-source on --clock-control none -k regex:"wgrad_tc|unpool|conv_first_wgrad|tail_bwd|side_folded" -c 14 -f \
    -o gpurun_out/r02_backward python scripts/one_train_step.py > gpurun_out/r02_backward.log 2>&1
timeout 200 ncu $SEC --import-source on --clock-control none -k regex:"conv3x3_halo|stage1" -c 3 -f \
    -o gpurun_out/r02_halo_lean python scripts/one_forward.py > gpurun_out/r02_halo_lean.log 2>&1
ls -la gpurun_out/*.ncu-rep
