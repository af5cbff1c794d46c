#!/bin/bash
# ncu --set full capture of kernels matching a regex inside a short bench run.  usage: tools/ncu_kernel.sh <tag> <regex> [skip] [count]
T=$1; RX=$2; SKIP=${3:-6}; CNT=${4:-2}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -f -o gpurun_out/${T} \
   python bench.py --steps 2 --warmup 6 --no-cpu-baseline --no-latency --no-gpu-eager > gpurun_out/${T}_ncu.log 2>&1
echo "rc=$?"; ls -la gpurun_out/${T}.ncu-rep
