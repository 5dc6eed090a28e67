#!/bin/bash
# Run under gpurun: writes the ncu evidence of this round into gpurun_out/ (copied to profiles/ afterwards).
#   1. launch list (device time of every launch) of the bench command
#   2. ncu --set full of the dominant contraction kernels (eager launches, one HVP)
#   3. ncu --set full of the eigen-iteration vector kernels at P = 2^26
set -u
R=${1:-r1}
# the .ncu-rep files stay on the GPU box (tens of MB each; gpurun brings back at most 64 MiB): they are summarised there
# by tools/summarize_profiles.py and only the summaries travel (gpurun_out/profiles_$R/ -> profiles/)
REP=/tmp/ncu_$R
mkdir -p $REP gpurun_out/profiles_$R
# the launch list is taken with the host-driven loop (B2S_DEVICE_LOOP=0): one HVP graph per iteration, every kernel a node ncu can name
export B2S_DEVICE_LOOP=0
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --no-vec --no-reg --no-table --no-yardstick"
$BENCH > $REP/${R}_bench_plain.json 2> gpurun_out/${R}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $REP/${R}_launches_bench.csv $BENCH > gpurun_out/${R}_bench_under_ncu.log 2>&1
echo "launch list rc $?"
T="python tools/ncu_target.py cifar_densenet 32 2"
$T > gpurun_out/${R}_target_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tma_kernel -s 158 -c 4 -o $REP/${R}_conv_tma -f $T > gpurun_out/${R}_ncu_conv_tma.log 2>&1
echo "conv_tma rc $?"
# DenseNet3's weight gradients run on conv_wgrad_tma_kernel (conv_wgrad_tma.cu): launches 62.. are block-1 layers of the Hv pass
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tma_kernel -s 62 -c 4 -o $REP/${R}_conv_wgrad -f $T > gpurun_out/${R}_ncu_wgrad.log 2>&1
echo "wgrad rc $?"
ncu --set full --clock-control none -k regex:bn_ -s 160 -c 2 -o $REP/${R}_bn -f $T > gpurun_out/${R}_ncu_bn.log 2>&1
echo "bn rc $?"
# the chest VGG16-bn HVP: the wide layers where the tensor pipe can be busy (conv3_x .. conv5_x of the order-1 forward sweep)
G="python tools/ncu_target.py chest_vgg 4 1"
$G > gpurun_out/${R}_target_vgg_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 62 -c 8 -o $REP/${R}_vgg_conv_tc -f $G > gpurun_out/${R}_ncu_vgg_conv_tc.log 2>&1
echo "vgg conv_tc rc $?"
ncu --set full --clock-control none -k regex:conv_tc_wgrad_kernel -s 36 -c 6 -o $REP/${R}_vgg_wgrad -f $G > gpurun_out/${R}_ncu_vgg_wgrad.log 2>&1
echo "vgg wgrad rc $?"
ncu --set full --clock-control none -k regex:bn_bwd -s 30 -c 4 -o $REP/${R}_vgg_bn -f $G > gpurun_out/${R}_ncu_vgg_bn.log 2>&1
echo "vgg bn rc $?"
V="python tools/bench_vec.py"
$V > gpurun_out/${R}_vec_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:pi_ -s 130 -c 2 -o $REP/${R}_vec -f $V > gpurun_out/${R}_ncu_vec.log 2>&1
echo "vec rc $?"
B2S_PROFILE_SRC=$REP B2S_PROFILE_DST=gpurun_out/profiles_$R python tools/summarize_profiles.py $R
ls -la $REP gpurun_out/profiles_$R
