#!/usr/bin/env python
"""Per-layer precision budget of the tensor-core head (CPU emulation - test infrastructure; companion of tools/precision_budget.py).

head_tc_kernel evaluates conv1 (48 -> 96 merged), conv2 (32 -> 16 per tower) and conv3 (16 -> 8 per tower) with fp16 hi + lo operands as
three products hi*hi + lo*hi + hi*lo (fp32 accumulate, BatchNorm folded into the weights); conv4 / conv5 run in fp32.  This tool emulates
that arithmetic in PyTorch on the CPU for the three layers, one layer and one dropped term at a time, on N synthetic frames whose stem and
blocks are evaluated by the fp32 oracle, and counts Hann-weighted arg-max flips against the fp32 oracle (ties = top-1 - top-2 < 1e-5
excluded) and the score-map error.  Dropping `lo*hi` leaves the activations at fp16 precision, dropping `hi*lo` the weights.

    python tools/precision_budget_head.py [--n 10240] [--out profiles/r02_precision_budget_head.json]
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

LAYERS = ("conv1", "conv2", "conv3")


class SplitHead:
    """The CENTER head of the oracle with layers 1-3 evaluated like head_tc_kernel (BN folded, im2col x weights with split operands)."""

    def __init__(self, model, drops):
        self.m, self.drops = model, drops
        sd = model.sd
        self.w = {}
        for t in O.TOWERS:
            for i in range(4):
                p = f"box_head.conv{i + 1}_{t}"
                w, b = sd[f"{p}.0.weight"].double(), sd[f"{p}.0.bias"].double()
                g, beta = sd[f"{p}.1.weight"].double(), sd[f"{p}.1.bias"].double()
                mu, var = sd[f"{p}.1.running_mean"].double(), sd[f"{p}.1.running_var"].double()
                s = g / torch.sqrt(var + O.BN_EPS)
                self.w[(t, i)] = ((w * s[:, None, None, None]).float(), ((b - mu) * s + beta).float())

    def conv(self, x, w, b, drop, split):
        if not split:
            return F.relu(F.conv2d(x, w, b, stride=1, padding=1))
        B, C, H, W = x.shape
        cols = F.unfold(x, 3, padding=1).transpose(1, 2)                  # [B, HW, C * 9]
        y = mm_split(cols, w.reshape(w.shape[0], -1).t(), drop) + b       # [B, HW, Cout]
        return F.relu(y.transpose(1, 2).reshape(B, -1, H, W))

    def __call__(self, feat):
        outs = {}
        sd = self.m.sd
        for t in O.TOWERS:
            x = feat
            for i in range(4):
                w, b = self.w[(t, i)]
                x = self.conv(x, w, b, self.drops.get(f"conv{i + 1}"), i < 3)
            outs[t] = F.conv2d(x, sd[f"box_head.conv5_{t}.weight"], sd[f"box_head.conv5_{t}.bias"])
        sig = lambda v: torch.clamp(torch.sigmoid(v), min=1e-4, max=1 - 1e-4)
        return sig(outs["ctr"]), sig(outs["size"]), outs["offset"]


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
    configs.append(("all:single_pass_fp16", {c: "both" for c in LAYERS}))
    ref = O.OracleModel(sd)
    heads = [(name, SplitHead(ref, d)) for name, d in configs]
    stats = {name: dict(flips=0, max_err=0.0) for name, _ in configs}
    ties = 0
    t0 = time.time()
    for g0 in range(0, n, group):
        idx = range(g0, min(n, g0 + group))
        z = torch.cat([O.preprocess(O.sample_target_cv(frames[i % Fn], list(init_boxes[i]), 2.0, 128)[0]) for i in idx])
        x = torch.cat([O.preprocess(O.sample_target_cv(frames[(i * 7 + 3) % Fn], list(step_boxes[i]), 4.0, 256)[0]) for i in idx])
        with torch.no_grad():
            taps = {}
            o = ref.forward(z, x, taps)
            t = taps["tokens_norm"]
            feat = t[:, -ref.feat_sz ** 2:].unsqueeze(-1).permute(0, 3, 2, 1).contiguous().view(t.shape[0], ref.C, ref.feat_sz, ref.feat_sz)
            resp = (win * o["score_map"]).flatten(1)
            top = torch.topk(resp, 2, dim=1).values
            tie = (top[:, 0] - top[:, 1]) < 1e-5
            ties += int(tie.sum())
            am = resp.argmax(dim=1)
            for name, h in heads:
                score = h(feat)[0]
                r2 = (win * score).flatten(1)
                st = stats[name]
                st["flips"] += int(((r2.argmax(dim=1) != am) & ~tie).sum())
                st["max_err"] = max(st["max_err"], float((score - o["score_map"]).abs().max()))
        if (g0 // group) % 20 == 0:
            print(f"{g0 + len(idx)} / {n} frames, {time.time() - t0:.0f} s", file=sys.stderr, flush=True)
    return {"frames": n, "ties_excluded": ties, "weights": weights, "frame_hw": [H, W], "configs": stats}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10240)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_precision_budget_head.json"))
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    res = {"what": "arg-max flips / max |score_map error| vs the fp32 oracle when ONE split term of ONE tensor-core head layer is dropped "
                   "(CPU emulation of the fp16 hi/lo arithmetic of head_tc_kernel; stem and blocks in fp32)",
           "runs": [run(a.n, dict(seed=11, stress=True)), run(a.n // 2, dict(seed=0, stress=False))]}
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
