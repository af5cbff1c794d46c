#!/usr/bin/env python
"""Per-layer precision budget of the tensor-core stem, search branch (CPU emulation - test infrastructure; companion of
tools/precision_budget.py).

conv2 / conv3 / conv4 (3x3 stride 2, BatchNorm folded) take fp16 hi + lo operands as three products hi*hi + lo*hi + hi*lo (fp32
accumulate); conv1 takes the uint8 crop exactly and hi | lo weights.  This tool emulates that arithmetic in PyTorch on the CPU, one layer
and one dropped term at a time, on N synthetic frames (template branch, blocks and head by the fp32 oracle) and counts Hann-weighted
arg-max flips against the fp32 oracle (ties = top-1 - top-2 < 1e-5 excluded) and the score-map error.  Dropping `lo*hi` leaves the
layer's INPUT at fp16 precision - its producer would no longer have to write, nor the layer read, the low half of the operand image.

    python tools/precision_budget_stem.py [--n 10240] [--out profiles/r02_precision_budget_stem.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import vt_oracle as O  # noqa: E402
from precision_budget import mm_split  # noqa: E402

LAYERS = ("conv2", "conv3", "conv4")


class SplitStem(O.OracleModel):
    def __init__(self, sd, drops):
        super().__init__(sd)
        self.drops = drops
        self.folded = []
        for i in range(4):
            p = f"patch_embed.net.{2 * i}"
            w = self.sd[f"{p}.c.weight"].double()
            g, beta = self.sd[f"{p}.bn.weight"].double(), self.sd[f"{p}.bn.bias"].double()
            mu, var = self.sd[f"{p}.bn.running_mean"].double(), self.sd[f"{p}.bn.running_var"].double()
            s = g / torch.sqrt(var + O.BN_EPS)
            self.folded.append(((w * s[:, None, None, None]).float(), (beta - mu * s).float()))

    def patch_embed(self, img, taps=None, tag=""):
        if tag != "_x":                                   # the template branch runs on the fp32 kernels at initialize()
            return super().patch_embed(img, taps, tag)
        x = img
        for i in range(4):
            w, b = self.folded[i]
            if i == 0:
                x = F.conv2d(x, w, b, stride=2, padding=1)
            else:
                B, C, H, W = x.shape
                cols = F.unfold(x, 3, padding=1, stride=2).transpose(1, 2)
                y = mm_split(cols, w.reshape(w.shape[0], -1).t(), self.drops.get(f"conv{i + 1}")) + b
                x = y.transpose(1, 2).reshape(B, -1, H // 2, W // 2)
            if i < 3:
                x = F.hardswish(x)
        return x.flatten(2).transpose(1, 2)


def run(n, weights, H=360, W=640, Fn=16, group=64):
    sd = O.make_state_dict(**weights)
    frames = np.concatenate([O.synth_frames(Fn // 2, H, W, seed=81, smooth=True), O.synth_frames(Fn // 2, H, W, seed=82)])
    init_boxes, step_boxes = O.synth_boxes(n, H, W, seed=83), O.synth_boxes(n, H, W, seed=84)
    win = O.hann2d(16, 16)
    configs = [("three_terms", {})]
    for c in LAYERS:
        for drop in ("lo_hi", "hi_lo"):
            configs.append((f"{c}:drop_{drop}", {c: drop}))
    configs.append(("all:drop_lo_hi", {c: "lo_hi" for c in LAYERS}))
    ref = O.OracleModel(sd)
    models = [(name, SplitStem(sd, d)) for name, d in configs]
    stats = {name: dict(flips=0, max_err=0.0) for name, _ in configs}
    ties = 0
    t0 = time.time()
    for g0 in range(0, n, group):
        idx = range(g0, min(n, g0 + group))
        z = torch.cat([O.preprocess(O.sample_target_cv(frames[i % Fn], list(init_boxes[i]), 2.0, 128)[0]) for i in idx])
        x = torch.cat([O.preprocess(O.sample_target_cv(frames[(i * 7 + 3) % Fn], list(step_boxes[i]), 4.0, 256)[0]) for i in idx])
        with torch.no_grad():
            o = ref.forward(z, x)
            resp = (win * o["score_map"]).flatten(1)
            top = torch.topk(resp, 2, dim=1).values
            tie = (top[:, 0] - top[:, 1]) < 1e-5
            ties += int(tie.sum())
            am = resp.argmax(dim=1)
            for name, m in models:
                om = m.forward(z, x)
                r2 = (win * om["score_map"]).flatten(1)
                st = stats[name]
                st["flips"] += int(((r2.argmax(dim=1) != am) & ~tie).sum())
                st["max_err"] = max(st["max_err"], float((om["score_map"] - o["score_map"]).abs().max()))
        if (g0 // group) % 20 == 0:
            print(f"{g0 + len(idx)} / {n} frames, {time.time() - t0:.0f} s", file=sys.stderr, flush=True)
    return {"frames": n, "ties_excluded": ties, "weights": weights, "frame_hw": [H, W], "configs": stats}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10240)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_precision_budget_stem.json"))
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    res = {"what": "arg-max flips / max |score_map error| vs the fp32 oracle when ONE split term of ONE tensor-core stem layer (search branch) is "
                   "dropped (CPU emulation of the fp16 hi/lo arithmetic of conv_s2_tc_kernel / stem12_fused_kernel; everything else in fp32)",
           "runs": [run(a.n, dict(seed=11, stress=True)), run(a.n // 2, dict(seed=0, stress=False))]}
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
