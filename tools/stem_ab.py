#!/usr/bin/env python
"""A / B of the fused stem front (vt_stem_fused.cu) against the previous three-kernel front (VT_STEM_UNFUSED=1) and the CPU oracle:
one open-loop step of n tracks in each mode (child processes: the switch is read once per process), maps and boxes compared.
    python tools/stem_ab.py [--n 96] [--hw 360x640]"""
import argparse, json, os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=96)
ap.add_argument("--hw", default="360x640")
ap.add_argument("--child", default=None)
ap.add_argument("--oracle", type=int, default=8, help="tracks also compared with the CPU oracle")
a = ap.parse_args()
H, W = (int(v) for v in a.hw.split("x"))

if a.child:
    import torch
    from oracle import vt_oracle as O
    from vittracker_b200 import BatchedTracker, FramePool, load_cfg
    n, F = a.n, 4
    sd = O.make_state_dict(seed=9, stress=True)
    frames = np.concatenate([O.synth_frames(F // 2, H, W, seed=51, smooth=True), O.synth_frames(F // 2, H, W, seed=52)])
    init_boxes, step_boxes = O.synth_boxes(n, H, W, seed=52), O.synth_boxes(n, H, W, seed=53)
    fi, fs = np.arange(n) % F, (np.arange(n) + 1) % F
    bt = BatchedTracker(load_cfg(), sd, max_tracks=n, chunk_tracks=min(n, 64))
    pool = FramePool(frames, bt.device)
    assert int(bt.initialize(pool, torch.from_numpy(fi), init_boxes).abs().sum()) == 0
    bt.set_state(step_boxes)
    out, det = bt.track(pool, torch.from_numpy(fs), update_state=False, detail=True)
    torch.cuda.synchronize()
    maps = bt.engine.tracks_last_maps(0, n)
    np.savez(a.child, out=out.cpu().numpy(), det=det.cpu().numpy(), score=maps["score_map"].cpu().numpy(), size=maps["size_map"].cpu().numpy(),
             off=maps["offset_map"].cpu().numpy())
    sys.exit(0)

res = {}
with tempfile.TemporaryDirectory() as td:
    for mode in ("fused", "unfused"):
        env = dict(os.environ, VT_STEM_UNFUSED="1" if mode == "unfused" else "0")
        p = os.path.join(td, mode + ".npz")
        r = subprocess.run([sys.executable, __file__, "--n", str(a.n), "--hw", a.hw, "--child", p], env=env, capture_output=True, text=True, timeout=300)
        if r.returncode != 0:
            print(json.dumps({"mode": mode, "rc": r.returncode, "stderr": r.stderr[-1500:]}))
            sys.exit(2)
        res[mode] = dict(np.load(p))
f, u = res["fused"], res["unfused"]
rep = {"n": a.n, "hw": [H, W]}
for k in ("score", "size", "off"):
    rep[f"max_abs_diff_{k}"] = float(np.abs(f[k] - u[k]).max())
rep["argmax_equal"] = int((f["det"][:, 5] == u["det"][:, 5]).sum())
rep["max_box_diff"] = float(np.abs(f["out"][:, :4] - u["out"][:, :4]).max())
rep["status_fused"] = f["det"][:, 6].astype(int).tolist()[:8]
rep["nan_fused"] = int(np.isnan(f["score"]).sum())
if a.oracle > 0:
    import torch
    from oracle import vt_oracle as O
    n, F = a.n, 4
    sd = O.make_state_dict(seed=9, stress=True)
    model = O.OracleModel(sd)
    frames = np.concatenate([O.synth_frames(F // 2, H, W, seed=51, smooth=True), O.synth_frames(F // 2, H, W, seed=52)])
    init_boxes, step_boxes = O.synth_boxes(n, H, W, seed=52), O.synth_boxes(n, H, W, seed=53)
    errs = {"fused": 0.0, "unfused": 0.0}
    for i in range(min(a.oracle, n)):
        z = O.preprocess(O.sample_target_cv(frames[i % F], list(init_boxes[i]), 2.0, 128)[0])
        x = O.preprocess(O.sample_target_cv(frames[(i + 1) % F], list(step_boxes[i]), 4.0, 256)[0])
        o = model.forward(z, x)
        for m in errs:
            errs[m] = max(errs[m], float(np.abs(res[m]["score"][i].ravel() - o["score_map"].numpy().ravel()).max()),
                          float(np.abs(res[m]["off"][i].ravel() - o["offset_map"].numpy().ravel()).max()))
    rep["max_abs_err_vs_oracle"] = errs
print(json.dumps(rep))
