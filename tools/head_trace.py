"""Phase timing of head_kernel (CTA 0): NVCC_EXTRA=-DVT_HEAD_TRACE python -m vittracker_b200.build --force"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vt_oracle as O
from vittracker_b200 import load_cfg, _lib
from vittracker_b200.engine import Engine
e = Engine(load_cfg(), max_tracks=148, chunk_tracks=148)
e.load_state_dict(O.make_state_dict(seed=1, stress=True))
z = torch.randn(148, 3, 128, 128); x = torch.randn(148, 3, 256, 256)
lib = _lib.load(); buf = (C.c_longlong * 32)()
for _ in range(3):
    e.forward(z, x); lib.vt_head_trace_read(buf)
t = np.frombuffer(buf, dtype=np.int64)[:7]
names = ["zero+LN+A", "conv1", "conv2", "conv3", "conv4", "conv5+argmax"]
for i, n in enumerate(names): print(f"{n:14s} {t[i+1]-t[i]:8d} cycles")
print("total", t[6] - t[0])
