#!/bin/bash
# Round-2 GPU session: usage tools/r2_run.sh <tag> <what...>   (what: tests bench sanitize soak ncu argmax)
T=$1; shift
O=gpurun_out
mkdir -p $O
for what in "$@"; do
  case $what in
    tests)   VT_SOAK_STEPS=${VT_SOAK_STEPS:-20000} timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/${T}_pytest_gpu.txt; cat $O/${T}_pytest_gpu.txt ;;
    bench)   timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; head -c 3000 $O/${T}_bench.json; echo; tail -5 $O/${T}_bench.err ;;
    benchq)  timeout 600 python bench.py --no-cpu-baseline --no-latency --no-gpu-eager > $O/${T}_benchq.json 2> $O/${T}_benchq.err; python tools/bench_brief.py $O/${T}_benchq.json; tail -3 $O/${T}_benchq.err ;;
    ref)     timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err; head -c 1500 $O/${T}_bench_reference_arm.json; echo ;;
    sanitize) bash tools/sanitize.sh $T memcheck synccheck ;;
    sanitize2) bash tools/sanitize.sh $T racecheck initcheck ;;
    soak)    timeout 900 python tools/soak.py --steps 100000 > $O/${T}_soak.txt 2>&1; tail -3 $O/${T}_soak.txt ;;
    argmax)  timeout 900 python tools/argmax_parity.py --blocks tcgen05 --out $O/${T}_argmax_parity_tcgen05.json | head -c 1500; echo
             timeout 900 python tools/argmax_parity.py --blocks simt --out $O/${T}_argmax_parity_simt.json | head -c 600; echo ;;
    widest)  timeout 900 python bench.py --config vit_768_h256_d12 --tracks 512 --steps 3 --warmup 3 > $O/${T}_bench_widest.json 2> $O/${T}_bench_widest.err; head -c 2500 $O/${T}_bench_widest.json; echo; tail -3 $O/${T}_bench_widest.err ;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 70 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 8 --no-cpu-baseline --no-latency --no-gpu-eager > /dev/null 2> $O/${T}_ncu.err; python tools/summarize_launches.py $O/${T}_launches.csv 2>/dev/null | head -20 ;;
    *) echo "unknown $what" ;;
  esac
done
