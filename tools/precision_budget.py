#!/usr/bin/env python
"""Per-contraction precision budget of the tensor-core ViT blocks (CPU emulation - test infrastructure).

The block kernel evaluates every contraction with fp16 hi + lo operands as three products, hi*hi + lo*hi + hi*lo (fp32 accumulate).
Dropping `lo*hi` leaves the A operand (activations) at fp16 precision, dropping `hi*lo` the B operand (weights / K / V').  This tool
emulates the split arithmetic in PyTorch on the CPU (operands rounded exactly as the kernel rounds them, fp32 accumulation), one
contraction and one dropped term at a time, on N synthetic frames, and counts Hann-weighted arg-max flips against the fp32 oracle
(ties = oracle top-1 - top-2 < 1e-5 excluded) and the score-map error.  A term may be dropped where flips stay 0.

    python tools/precision_budget.py [--n 10240] [--out profiles/r02_precision_budget.json]
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
from oracle import vt_oracle as O  # noqa: E402

CONTRACTIONS = ("qkv", "scores", "pv", "fc1", "fc2")
ONLY = []


def split(x):
    hi = x.half().float()
    lo = (x - hi).half().float()
    return hi, lo


def mm_split(a, bt, drop):
    """a [.., M, K] @ bt [.., K, N] with fp16 hi/lo operands; drop in {None, 'lo_hi', 'hi_lo', 'both'}."""
    ah, al = split(a)
    bh, bl = split(bt)
    r = ah @ bh
    if drop not in ("lo_hi", "both"):
        r = r + al @ bh
    if drop not in ("hi_lo", "both"):
        r = r + ah @ bl
    return r


class SplitModel(O.OracleModel):
    """Oracle graph with the blocks' contractions evaluated like vt_block_tc.cu (proj folded into V, q pre-scaled)."""

    def __init__(self, sd, drops):
        super().__init__(sd)
        self.drops = drops
        self.folded = {}
        for b in range(self.depth):
            p = f"blocks.{b}"
            wqkv, bqkv = self.sd[f"{p}.attn.qkv.weight"].double(), self.sd[f"{p}.attn.qkv.bias"].double()
            wp = self.sd[f"{p}.attn.proj.weight"].double()
            C = self.C
            w = wqkv.clone()
            w[2 * C:] = wp @ wqkv[2 * C:]
            bias = bqkv.clone()
            bias[2 * C:] = wp @ bqkv[2 * C:]
            self.folded[b] = (w.float(), bias.float())

    def block(self, x, b):
        p = f"blocks.{b}"
        C = self.C
        d = self.drops
        h = F.layer_norm(x, (C,), self.sd[f"{p}.norm1.weight"], self.sd[f"{p}.norm1.bias"], O.LN_EPS)
        w, bias = self.folded[b]
        qkv = mm_split(h, w.t(), d.get("qkv")) + bias
        q, k, v = qkv[..., :C] * (C ** -0.5), qkv[..., C:2 * C], qkv[..., 2 * C:]
        s = mm_split(q, k.transpose(-2, -1), d.get("scores"))
        m = s.max(dim=-1, keepdim=True).values
        pr = torch.exp(s - m)
        l = pr.sum(dim=-1, keepdim=True)
        o = mm_split(pr, v, d.get("pv")) / l
        x = x + o + self.sd[f"{p}.attn.proj.bias"]
        h = F.layer_norm(x, (C,), self.sd[f"{p}.norm2.weight"], self.sd[f"{p}.norm2.bias"], O.LN_EPS)
        h = F.gelu(mm_split(h, self.sd[f"{p}.mlp.fc1.weight"].t(), d.get("fc1")) + self.sd[f"{p}.mlp.fc1.bias"])
        return x + mm_split(h, self.sd[f"{p}.mlp.fc2.weight"].t(), d.get("fc2")) + self.sd[f"{p}.mlp.fc2.bias"]


def run(n, weights, H=360, W=640, Fn=16, group=64):
    sd = O.make_state_dict(**weights)
    frames = np.concatenate([O.synth_frames(Fn // 2, H, W, seed=81, smooth=True), O.synth_frames(Fn // 2, H, W, seed=82)])
    init_boxes, step_boxes = O.synth_boxes(n, H, W, seed=83), O.synth_boxes(n, H, W, seed=84)
    win = O.hann2d(16, 16)
    configs = [("three_terms", {})]
    for c in CONTRACTIONS:
        for drop in ("lo_hi", "hi_lo"):
            configs.append((f"{c}:drop_{drop}", {c: drop}))
    configs.append(("pv+fc2:drop_lo_hi", {"pv": "lo_hi", "fc2": "lo_hi"}))
    configs.append(("scores:single_pass_fp16", {"scores": "both"}))
    configs.append(("scores:single_pass+qkv:drop_lo_hi", {"scores": "both", "qkv": "lo_hi"}))
    configs.append(("all:single_pass_fp16", {c: "both" for c in CONTRACTIONS}))
    if ONLY:
        configs = [c for c in configs if any(c[0].startswith(p) for p in ONLY)]
    ref = O.OracleModel(sd)
    models = [(name, SplitModel(sd, d)) for name, d in configs]
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
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_precision_budget.json"))
    ap.add_argument("--only", default=None, help="comma-separated config name prefixes to run (default: all)")
    a = ap.parse_args()
    ONLY[:] = a.only.split(",") if a.only else []
    torch.set_num_threads(os.cpu_count() or 1)
    res = {"what": "arg-max flips / max |score_map error| vs the fp32 oracle when ONE split term of ONE block contraction is dropped "
                   "(CPU emulation of the fp16 hi/lo arithmetic of vt_block_tc.cu; stem and head in fp32)",
           "runs": [run(a.n, dict(seed=11, stress=True)), run(a.n // 2, dict(seed=0, stress=False))]}
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
