#!/bin/bash
# Run under gpurun: writes the ncu evidence of this round into gpurun_out/ (copied to profiles/ afterwards).
#   1. launch list (device time of every launch) of the bench command
#   2. ncu --set full of the dominant contraction kernels (eager launches, one HVP)
#   3. ncu --set full of the eigen-iteration vector kernels at P = 2^26
set -u
R=${1:-r1}
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --no-vec --no-reg"
$BENCH > gpurun_out/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${R}_launches_bench.csv $BENCH > gpurun_out/${R}_bench_under_ncu.log 2>&1
echo "launch list rc $?"
T="python tools/ncu_target.py cifar_densenet 32 2"
$T > gpurun_out/${R}_target_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tma_kernel -s 158 -c 4 -o gpurun_out/${R}_conv_tma -f $T > gpurun_out/${R}_ncu_conv_tma.log 2>&1
echo "conv_tma rc $?"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_wgrad_kernel -s 80 -c 2 -o gpurun_out/${R}_conv_wgrad -f $T > gpurun_out/${R}_ncu_wgrad.log 2>&1
echo "wgrad rc $?"
ncu --set full --clock-control none -k regex:bn_ -s 160 -c 2 -o gpurun_out/${R}_bn -f $T > gpurun_out/${R}_ncu_bn.log 2>&1
echo "bn rc $?"
V="python tools/bench_vec.py"
$V > gpurun_out/${R}_vec_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:pi_ -s 130 -c 2 -o gpurun_out/${R}_vec -f $V > gpurun_out/${R}_ncu_vec.log 2>&1
echo "vec rc $?"
ls -la gpurun_out/${R}_*
