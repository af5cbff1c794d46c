#!/usr/bin/env python
"""Soak of the persistent tcgen05 kernels: N open-loop steps of the benchmarked workload (1024 tracks, 64 x 720p frames resident in HBM,
states re-seeded every step) under a HOST watchdog.  The kernels hand work between warps through mbarrier parity protocols; a protocol bug shows
as a hang, and a hung persistent kernel never returns - so the loop synchronises every `--beat` steps and a watchdog thread ends the process
(exit code 3, after printing which profiled stage launch is stuck: vt_debug_pending) if a synchronisation does not come back in time.
Process exit tears the context down, which is what frees the GPU.  Also checks that the boxes of repeated identical steps stay bit-identical.

    python tools/soak.py [--steps 100000] [--tracks 1024] [--blocks tcgen05] [--stall-s 30]
"""
import argparse, ctypes as C, json, os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vt_oracle as O                      # synthetic workload generators only
from vittracker_b200 import BatchedTracker, FramePool, load_cfg, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=100000)
ap.add_argument("--tracks", type=int, default=1024)
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--blocks", default="tcgen05")
ap.add_argument("--beat", type=int, default=500, help="synchronise (and feed the watchdog) every this many steps")
ap.add_argument("--stall-s", type=float, default=30.0)
a = ap.parse_args()
n, F = a.tracks, a.frames
bt = BatchedTracker(load_cfg(), O.make_state_dict(seed=1, stress=True), max_tracks=n, chunk_tracks=min(1024, n), blocks_impl=a.blocks)
dev = bt.device
pool = FramePool(O.synth_frames(F, 720, 1280, seed=1000), dev)
lib = _lib.load()
beat = [time.time(), "start"]


def watchdog():
    while True:
        time.sleep(1.0)
        if time.time() - beat[0] > a.stall_s:
            buf = (C.c_int32 * 4096)()
            k = lib.vt_debug_pending(bt.engine.handle, buf, 4096)
            pend = [(i, buf[i] // 4, buf[i] % 4) for i in range(k) if buf[i] % 4 != 3]
            print(json.dumps({"soak": "HANG", "at": beat[1], "unfinished_stage_launches(index, stage, started|finished<<1)": pend[:16]}), flush=True)
            os._exit(3)


threading.Thread(target=watchdog, daemon=True).start()
assert int(bt.initialize(pool, torch.arange(n, device=dev) % F, O.synth_boxes(n, 720, 1280, seed=2000)).abs().sum()) == 0
step_boxes = torch.stack([torch.tensor(O.synth_boxes(n, 720, 1280, seed=3000 + s)) for s in range(8)]).to(dev)
fh = np.arange(n, dtype=np.int64) % F
offs = [torch.from_numpy(((fh + t) % F) * pool.frame_bytes).to(dev) for t in range(F)]
ref = {}
mismatch = 0
bt.engine.profile(True)
t0 = time.time()
for t in range(a.steps):
    bt.engine.tracks_set_state(step_boxes[t % 8], first=0)
    out = bt.track_offsets(pool.data, offs[t % F], update_state=True)
    key = (t % 8, t % F)                               # same state set and frame assignment => same boxes, bit for bit
    if t < 64:
        ref[key] = out.clone()
    elif t % a.beat == a.beat - 1 and key in ref:
        mismatch += int(not torch.equal(out, ref[key]))
    if t % a.beat == a.beat - 1:
        torch.cuda.synchronize(dev)
        bt.engine.profile_read()
        beat[:] = [time.time(), f"step {t}"]
torch.cuda.synchronize(dev)
el = time.time() - t0
print(json.dumps({"soak": "OK", "steps": a.steps, "tracks": n, "blocks": a.blocks, "ms_per_step": el / a.steps * 1e3,
                  "repeat_mismatches": mismatch, "wall_s": round(el, 1)}), flush=True)
sys.exit(0 if mismatch == 0 else 4)
