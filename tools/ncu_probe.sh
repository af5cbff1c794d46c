#!/bin/bash
# Development aid: does the pipeline survive ncu's serialised, cache-flushed launches?  usage: tools/ncu_probe.sh <lib dir or -> <runs>
for i in $(seq 1 ${2:-2}); do
  if [ "$1" != "-" ]; then export VT_LIB_DIR=$1; fi
  timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 70 --csv --log-file gpurun_out/ncu_probe.csv python bench.py --steps 2 --warmup 8 --no-cpu-baseline --no-latency > gpurun_out/ncu_probe.log 2>&1
  echo "lib=$1 run=$i rc=$? profiled=$(grep -c '^"' gpurun_out/ncu_probe.csv)"
done
