"""Instruction mix per kernel from an ncu source-page CSV holding one or more kernels."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:80], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for k in kernels:
    hdr = k["hdr"]
    si, ei = hdr.index("Source"), hdr.index("Instructions Executed")
    ops, tot = collections.Counter(), 0
    for r in k["rows"]:
        try:
            n = int(r[ei])
        except Exception:
            continue
        toks = r[si].strip().split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0]] += n
        tot += n
    print("==", k["name"], "| warp-instrs", tot)
    print("   " + "  ".join(f"{o}:{100 * v / tot:.1f}%" for o, v in ops.most_common(14)))
