#!/bin/bash
# Round-2 evidence run (on the GPU box): ncu captures of the dominant kernels and the launch list of a C3 step.
# Every ncu command runs a program that has already exited 0 without ncu in this round.  Output: gpurun_out/r02_*
set -u
O=gpurun_out
SECS="--section SpeedOfLight --section WarpStateStats --section SourceCounters --section Occupancy --section LaunchStats --section SchedulerStats --section MemoryWorkloadAnalysis"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --shard-legs off --label-check 2"
# 1. launch list of one C3 run (training set-up + warm-up step + timed step + single-cloud calls)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_c3.csv $B > $O/r02_ncu_launches.log 2>&1
# 2. the pooled sweep of the pre-filter and the bound sweep at C3 size (300 k queries x 1.07 M words), full set
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:k_knn_gemm<\(bool\)1, \(int\)1, \(bool\)1>' --launch-skip 1 -c 1 -o $O/r02_gemm_pool_c3 -f python tools/pca_profile.py 1024 > $O/r02_ncu_gemm_pool.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:k_knn_gemm<\(bool\)1, \(int\)1, \(bool\)0>' --launch-skip 2 -c 1 -o $O/r02_gemm_bound_c3 -f python tools/pca_profile.py 1024 > $O/r02_ncu_gemm_bound.log 2>&1
# 3. descriptor kernel: staged launches (frames, descriptors) of a C3 step, dense launch of a C5 step
ncu $SECS --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:k_shot<\(bool\)0, \(bool\)0>' --launch-skip 32 -c 2 -o $O/r02_shot_c3 -f $B > $O/r02_ncu_shot_c3.log 2>&1
ncu $SECS --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:k_shot<\(bool\)0, \(bool\)1>' --launch-skip 1 -c 1 -o $O/r02_shot_c5 -f $B --workload c5 > $O/r02_ncu_shot_c5.log 2>&1
# 4. C4 (CSHOT-1344, streaming-query variant): last GEMM launch of the run
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:k_knn_gemm<\(bool\)0' --launch-skip 3 -c 1 \
  -o $O/r02_gemm_c4 -f $B --workload c4 --batch 256 > $O/r02_ncu_gemm_c4.log 2>&1
ls -la $O/r02_*.ncu-rep
