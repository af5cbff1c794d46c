#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU tests, bench lines, parity gate, widest configuration.  Usage: tools/evidence.sh <tag>
T=${1:-r01x}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${T}_pytest_gpu.txt
timeout 600 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err
timeout 600 python tools/argmax_parity.py --blocks tcgen05 --out $O/${T}_argmax_parity_10k_tcgen05.json > /dev/null 2>&1
timeout 600 python tools/argmax_parity.py --blocks simt --out $O/${T}_argmax_parity_10k_simt.json > /dev/null 2>&1
timeout 300 python tools/bench_generic.py > $O/${T}_bench_widest_config.json 2>> $O/${T}_bench.err
cat $O/${T}_pytest_gpu.txt; head -c 600 $O/${T}_bench.json; echo; cat $O/${T}_bench_reference_arm.json | head -c 400; echo; cat $O/${T}_bench_widest_config.json
