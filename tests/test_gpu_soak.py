"""Soak of the persistent tcgen05 kernels under a host watchdog (tools/soak.py): 10^5 steps of the benchmarked workload by default
(VT_SOAK_STEPS overrides).  Runs in a child process so that a deadlocked kernel ends the child (exit code 3), not the test session."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_soak_no_hang_and_bit_stable():
    steps = int(os.environ.get("VT_SOAK_STEPS", "100000"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "soak.py"), "--steps", str(steps)], capture_output=True, text=True,
                       timeout=120 + steps * 0.004, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    res = json.loads(lines[-1])
    print(res)
    assert res["soak"] == "OK" and res["steps"] == steps and res["repeat_mismatches"] == 0
