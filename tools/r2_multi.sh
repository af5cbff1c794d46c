#!/bin/bash
# Multi-GPU evidence (run under gpurun --gpus N): tools/r2_multi.sh <tag> <N> [what...]   what: weak strong widest nccltest topo
# BENCH_ARGS (environment) is appended to the bench command lines, e.g. "--no-cpu-baseline --no-latency --no-gpu-eager".
T=$1; N=$2; shift; shift
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for what in "$@"; do
  case $what in
    weak)    timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 $BENCH_ARGS > $O/${T}_bench_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err
             python tools/bench_brief.py $O/${T}_bench_${N}gpu.json; python - <<PY
import json
d=json.loads(open("$O/${T}_bench_${N}gpu.json").read().strip().splitlines()[-1])
print(json.dumps({"strong_8192": d.get("strong_8192"), "e2e": {k: d["e2e"][k] for k in ("value","h2d_cap_gbs","h2d_cap_gbs_per_rank","frac_of_h2d_cap","binding")}}))
PY
             ;;
    strong)  timeout 900 $TR --master-port 29522 bench.py --gpus $N --total-tracks 8192 --steps 20 --warmup 3 --no-strong-leg > $O/${T}_bench_strong8192_${N}gpu.json 2> $O/${T}_bench_strong8192_${N}gpu.err
             python tools/bench_brief.py $O/${T}_bench_strong8192_${N}gpu.json ;;
    widest)  timeout 1200 $TR --master-port 29523 bench.py --gpus $N --config vit_768_h256_d12 --tracks 512 --steps 3 --warmup 3 $BENCH_ARGS > $O/${T}_bench_widest_${N}gpu.json 2> $O/${T}_bench_widest_${N}gpu.err
             python tools/bench_brief.py $O/${T}_bench_widest_${N}gpu.json; tail -2 $O/${T}_bench_widest_${N}gpu.err ;;
    nccltest) timeout 300 python -m pytest tests/test_gpu_parity.py -q -k nccl 2>&1 | tail -2 ;;
    topo)    nvidia-smi topo -m > $O/${T}_topo_${N}gpu.txt 2>&1; lscpu | head -25 >> $O/${T}_topo_${N}gpu.txt; numactl -H >> $O/${T}_topo_${N}gpu.txt 2>&1; head -30 $O/${T}_topo_${N}gpu.txt ;;
  esac
done
