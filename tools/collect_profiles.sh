#!/bin/bash
# Round-2 evidence run (on the GPU box, one B200): GPU tests, smoke, one bench line per workload, the launch list of a C3
# step and ncu captures of the dominant kernels.  Every ncu command runs a program that has already exited 0 without
# ncu in this round; a number printed under ncu is never a bench value.  Output: gpurun_out/r02_*  (copied by hand into
# profiles/ after reading).  usage: bash tools/collect_profiles.sh [quick]
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --shard-legs off"
run() {  # name, timeout, command...
  local name=$1 t=$2; shift 2
  timeout $t "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(cut -c1-140 $O/$name.json | head -1)"
}
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $O/r02_pytest_gpu.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/r02_smoke.log | cut -c1-160)"
run r02_bench_c3_1gpu 300 python bench.py
run r02_bench_c3_chi2 240 $B --dist chisquared --label-check 8
run r02_bench_c2_1gpu 240 $B --workload c2 --label-check 8
run r02_bench_c4_1gpu 400 $B --workload c4 --label-check 4
run r02_bench_c5_1gpu 300 $B --workload c5 --label-check 2
[ "${1:-}" = quick ] && exit 0
[ "${1:-}" = plain ] && { PCDB_GEMM_PCA=0 run r02_bench_c3_plain_sweep 240 $B --label-check 2; PCDB_GEMM_PCA_WIDE=0 run r02_bench_c4_plain_sweep 400 $B --workload c4 --label-check 2; }
# launch list of one C3 run (training set-up + warm-up step + timed step + single-cloud calls)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_c3.csv \
  $B --steps 1 --warmup 1 --label-check 2 > $O/r02_ncu_launches.log 2>&1; echo "launch list rc=$?"
python tools/launch_list.py $O/r02_launches_c3.csv > $O/r02_launches_c3_step.csv; head -8 $O/r02_launches_c3_step.csv
# the pooled sweep of the pre-filter at C3 size (300 k queries x 1.07 M words), full set
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:k_knn_gemm<\(bool\)1, \(int\)1, \(bool\)1>' --launch-skip 1 -c 1 -o $O/r02_gemm_pool_c3 -f \
  python tools/pca_profile.py 1024 > $O/r02_ncu_gemm_pool.log 2>&1; echo "ncu pool rc=$?"
ls -la $O/*.ncu-rep
