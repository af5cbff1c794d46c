#!/usr/bin/env python
"""Throughput of the generic-configuration path on the widest vit_dist configuration (BASELINE configs[4]: C = 768, 12 heads,
depth 12, head 256; 33.32 GMAC = 66.65 GFLOP per tracked frame).  Not the headline bench - a reported number."""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vt_oracle as O
from vittracker_b200 import BatchedTracker, FramePool, parameters

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=64)
ap.add_argument("--chunk", type=int, default=16)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
cfg = parameters("vit_768_h256_d12").cfg
sd = O.make_state_dict(seed=1, stress=True, C=768, depth=12, head_ch=256)
frames = O.synth_frames(8, 720, 1280, seed=3)
bt = BatchedTracker(cfg, sd, max_tracks=a.tracks, chunk_tracks=a.chunk)
pool = FramePool(frames, bt.device)
n = a.tracks
boxes = O.synth_boxes(n, 720, 1280, seed=4)
fidx = torch.arange(n) % 8
assert int(bt.initialize(pool, fidx, boxes).abs().sum()) == 0
bt.track(pool, (fidx + 1) % 8, update_state=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(a.steps):
    bt.track(pool, (fidx + s) % 8, update_state=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"config": "vit_768_h256_d12 (widest)", "tracks": n, "chunk": a.chunk, "ms_per_step": ms, "frames_per_s": n / ms * 1e3,
                  "algorithmic_tflops": 66.65e9 * n / (ms * 1e-3) / 1e12, "path": "generic: im2col + tcgen05 GEMM (fp16 hi/lo split, fp32 accumulate), fp32 CUDA-core GEMM for small shapes"}))
