#!/bin/bash
# Development aid: bench.py at several stem chunk sizes (does the conv1 -> conv2 -> conv3 -> conv4 hand-over stay in L2?)
for c in "$@"; do
  timeout 100 python bench.py --no-cpu-baseline --no-latency --chunk $c --steps 10 2>/dev/null | C=$c python -c '
import sys, json, os
d = json.loads(sys.stdin.read())
print(os.environ["C"], round(d["value"]), round(d["ms_per_step"], 4), {k: round(v["ms_per_step"], 4) for k, v in d["stages"].items()}, "e2e", round(d["e2e"]["value"]))'
done
