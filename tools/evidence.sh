#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU tests (incl. the 10^5-step soak), bench lines, reference arm, parity gates for the three
# block implementations.  Usage: tools/evidence.sh <tag>
T=${1:-r02x}
O=gpurun_out
mkdir -p $O
VT_SOAK_STEPS=100000 timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/${T}_pytest_gpu.txt
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err
timeout 600 python bench.py --blocks tcgen05_3term --no-cpu-baseline --no-latency --no-gpu-eager > $O/${T}_bench_3term.json 2>> $O/${T}_bench.err
timeout 900 python tools/argmax_parity.py --blocks tcgen05 --out $O/${T}_argmax_parity_30k_tcgen05.json > /dev/null 2>&1
timeout 900 python tools/argmax_parity.py --blocks tcgen05_3term --out $O/${T}_argmax_parity_30k_tcgen05_3term.json > /dev/null 2>&1
timeout 900 python tools/argmax_parity.py --blocks simt --out $O/${T}_argmax_parity_30k_simt.json > /dev/null 2>&1
cat $O/${T}_pytest_gpu.txt; python tools/bench_brief.py $O/${T}_bench.json; python tools/bench_brief.py $O/${T}_bench_3term.json | cut -c1-200
head -c 500 $O/${T}_bench_reference_arm.json; echo
for b in tcgen05 tcgen05_3term simt; do python -c "import json; d=json.load(open('$O/${T}_argmax_parity_30k_$b.json')); print('$b', {k: d[k] for k in ('frames','argmax_flips','ties_excluded','boxes_outside_tolerance','status_nonzero','max_box_abs_err')})"; done
