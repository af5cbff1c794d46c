"""Phase timing of head_kernel (CTA 0): NVCC_EXTRA=-DVT_HEAD_TRACE python -m vittracker_b200.build --force"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vt_oracle as O
from vittracker_b200 import load_cfg, _lib
from vittracker_b200.engine import Engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
e = Engine(load_cfg(), max_tracks=N, chunk_tracks=N)
e.load_state_dict(O.make_state_dict(seed=1, stress=True))
z = torch.randn(N, 3, 128, 128); x = torch.randn(N, 3, 256, 256)
lib = _lib.load(); buf = (C.c_longlong * 32)()
for _ in range(3):
    e.forward(z, x); lib.vt_head_trace_read(buf)
t = np.frombuffer(buf, dtype=np.int64)
names = ["prologue+LN -> A image", "conv1 (MMA + 2 epilogues)", "conv2 (MMA + epilogue)", "conv3 (MMA + epilogue)", "conv4 (CUDA cores)", "conv5 + arg-max + decode"]
for i, n in enumerate(names): print(f"{n:28s} {t[i+1]-t[i]:8d} cycles")
print("total", t[6] - t[0])
if t[28] > 0: print(f"CTA 0, all its tracks: {t[31] - t[30]} cycles in {t[29] - t[28]} ns = {(t[31] - t[30]) / max(1, t[29] - t[28]):.3f} GHz")
if t[7] > 0: print("box decode by thread 0 (float64)", t[7] - t[6], " end-of-track barrier", t[8] - t[7])
if t[10] > 0:          # head_tc_kernel's finer stamps
    print("  prologue (zero rows, biases, TMEM alloc, barriers, sync)", t[10] - t[0])
    if t[20] > 0:
        print("    of which: up to the TMEM allocation", t[20] - t[0], " allocation", t[21] - t[20], " barrier init + CTA barrier", t[10] - t[21])
    print("  LayerNorm rows -> operand image                         ", t[11] - t[10])
    print("  barrier                                                 ", t[1] - t[11])
    print("  conv1 half 0: MMAs (wait)", t[12] - t[1], " epilogue", t[13] - t[12])
    print("  conv1 half 1: MMAs (wait)", t[15] - t[13], " epilogue", t[16] - t[15], " (barriers included in the waits)")
    print("  conv2: MMAs (wait)", t[18] - t[2], " zeroing + epilogue", t[3] - t[18])
