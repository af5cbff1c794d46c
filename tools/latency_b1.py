#!/usr/bin/env python
"""BASELINE configs[1] / SURVEY 8d C2 on its own: batch-1 Vit_dist.initialize() + track() latency over N open-loop frames
(default 10000) plus closed-loop bursts of 8 frames after each re-initialisation; prints one JSON object.

    python tools/latency_b1.py [--frames 10000] > gpurun_out/latency_b1_10k.json"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    a = ap.parse_args()
    import bench
    from oracle import vt_oracle as O          # seeded synthetic weights / frames / boxes only
    from vittracker_b200 import load_cfg
    out = bench.latency_b1(load_cfg(), O.make_state_dict(seed=1, stress=True), iters=a.frames)
    print(json.dumps(out))
