"""Per-step cycle trace of the tcgen05 block kernel (CTA 0). Build with VT_TC_TRACE defined:
   NVCC_EXTRA=-DVT_TC_TRACE python -m vittracker_b200.build --force ; python tools/tc_trace.py"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vt_oracle as O
from vittracker_b200 import load_cfg, _lib
from vittracker_b200.engine import Engine
cfg = load_cfg()
sd = O.make_state_dict(seed=1, stress=True)
e = Engine(cfg, max_tracks=148, chunk_tracks=148)
e.load_state_dict(sd)
z = torch.randn(148, 3, 128, 128); x = torch.randn(148, 3, 256, 256)
lib = _lib.load()
buf = (C.c_longlong * (2 * 4096))(); n = (C.c_int * 2)()
for it in range(3):
    e.forward(z, x)
    lib.vt_tc_trace_read(buf, n)
a = np.frombuffer(buf, dtype=np.int64).reshape(2, 4096)
ctl, epi = a[0, :n[0]], a[1, :n[1]]
print("control events", n[0], "epilogue events", n[1])
# control: pairs (before wait_go, after wait_go); epilogue: pairs (before wait_done, after wait_done)
cw = ctl.reshape(-1, 2); ew = epi.reshape(-1, 2)
t0 = min(cw[0, 0], ew[0, 0])
names = ["qkv0", "qkv1", "qkv2", "S0", "PV0+S1", "PV1+S2", "PV2", "h_A(fc1_0)", "h_B(fc1_1)", "y_A(fc2_0)", "h_A(fc1_2)", "y_B(fc2_1)", "y_A(fc2_2)"]
print("epilogue-side view (control-side columns are not aligned with these rows in the MLP phase)")
print(f"{'step':10s} {'ctl_wait_go':>11s} {'issue':>7s} {'epi_wait_done':>13s} {'epilogue':>9s}")
tot = dict(wg=0, iss=0, wd=0, ep=0)
for i in range(min(len(ew), 3 * len(names))):
    wait_go = cw[i, 1] - cw[i, 0] if i < len(cw) else 0
    issue = (cw[i + 1, 0] - cw[i, 1]) if i + 1 < len(cw) else 0          # after wait_go -> next wait_go start = issue + commit
    wait_done = ew[i, 1] - ew[i, 0]
    epil = (ew[i + 1, 0] - ew[i, 1]) if i + 1 < len(ew) else 0            # after wait_done -> next wait_done start = epilogue + signal
    tot["wg"] += wait_go; tot["iss"] += issue; tot["wd"] += wait_done; tot["ep"] += epil
    print(f"{names[i % len(names)]:10s} {wait_go:11d} {issue:7d} {wait_done:13d} {epil:9d}")
print("totals per track (3 blocks):", tot, "span", max(cw[-1, 1], ew[-1, 1]) - t0)
