#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): per kernel, the SASS instruction count and how many of the
mnemonics that matter are in it - UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st), UBLKCP
(cp.async.bulk), SYNCS (mbarrier), FFMA2 / FMUL2 / FADD2 (packed fp32), FHFMA (fp16 x fp16 + fp32), MUFU, IDP (dp2a).

    python tools/sass_summary.py [path/to/libvittrack_b200.so] > profiles/<tag>_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FHFMA", "FFMA", "MUFU", "IDP", "LDGSTS"]


def main() -> None:
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vittracker_b200", "lib", "libvittrack_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    print(f"`cuobjdump -sass {os.path.relpath(lib, ROOT)}` (sm_100a), instruction and mnemonic counts per kernel\n")
    print("| kernel | SASS instructions | " + " | ".join(KEYS) + " |")
    print("|---|---:|" + "---:|" * len(KEYS))
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name = part.split("\n", 1)[0].strip()
        ops = re.findall(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", part, re.M)
        c = collections.Counter(ops)
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem).replace("void ", "").replace("vt::", "")
        m = re.search(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{8}(\d+)", name)     # nvcc's internal-linkage mangling
        if m:
            n, rest = int(m.group(1)), name[m.end():]
            targs = re.search(r"ILi(\d+)ELi(\d+)", rest[n:])
            dem = rest[:n] + (f"<{targs.group(1)}, {targs.group(2)}>" if targs else "")
        print(f"| `{dem[:90]}` | {len(ops)} | " + " | ".join(str(c.get(k, 0) or "") for k in KEYS) + " |")


if __name__ == "__main__":
    main()
