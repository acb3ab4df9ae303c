#!/bin/bash
# ncu captures of the descriptor kernel: the staged launch on a C3 step and the dense launch on a C5 step.
# Usage (on the GPU box): bash tools/ncu_shot.sh <tag>
set -u
tag=${1:-r02}
SECS="--section SpeedOfLight --section WarpStateStats --section SourceCounters --section Occupancy --section LaunchStats --section SchedulerStats --section MemoryWorkloadAnalysis"
ncu $SECS --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:k_shot<\(bool\)0, \(bool\)0>' --launch-skip 16 -c 1 -o gpurun_out/${tag}_shot_c3 -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --shard-legs off --label-check 2 > gpurun_out/${tag}_ncu_shot_c3.log 2>&1
ncu $SECS --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:k_shot<\(bool\)0, \(bool\)1>' --launch-skip 1 -c 1 -o gpurun_out/${tag}_shot_c5 -f \
  python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --shard-legs off --label-check 2 > gpurun_out/${tag}_ncu_shot_c5.log 2>&1
ls -la gpurun_out/${tag}_shot_c*.ncu-rep
