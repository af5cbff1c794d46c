#!/bin/bash
# compute-sanitizer over every hand-written kernel (run under gpurun).  Usage: tools/sanitize.sh <tag> [tools...]
# memcheck: out-of-bounds / misaligned global + shared accesses;  racecheck: shared-memory hazards between warps (the mbarrier / bar.sync
# protocols of the tcgen05 kernels);  synccheck: divergent or mismatched barrier use;  initcheck: reads of uninitialised device memory.
T=${1:-r02}
shift
TOOLS=${@:-memcheck racecheck synccheck initcheck}
O=gpurun_out
mkdir -p $O
for tool in $TOOLS; do
  for part in fast generic; do
    log=$O/${T}_sanitizer_${tool}_${part}.txt
    extra=""
    [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
    t0=$SECONDS
    timeout 900 compute-sanitizer --tool $tool $extra --print-limit 30 python tools/sanitize_target.py $part > $log 2>&1
    echo "exit $? wall $((SECONDS - t0)) s" >> $log
    echo "== $tool $part: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|exit |wall ' $log | tr '\n' ' ')"
  done
done
