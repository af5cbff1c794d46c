#!/bin/bash
# Round-2 baseline on one B200: GPU tests, bench line, and a compute-sanitizer feasibility probe on the smoke path.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > $O/r02a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02a_pytest_gpu.txt
timeout 600 python bench.py --latency-frames 300 > $O/r02a_bench.json 2> $O/r02a_bench.err
which compute-sanitizer > $O/r02a_sanitizer_probe.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke >> $O/r02a_sanitizer_probe.txt 2>&1
echo "memcheck exit $?" >> $O/r02a_sanitizer_probe.txt
cat $O/r02a_pytest_gpu.txt; head -c 1500 $O/r02a_bench.json; echo; tail -15 $O/r02a_sanitizer_probe.txt
