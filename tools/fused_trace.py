"""Per-phase cycle trace of the fused stem front (one worker thread of CTA 0).  Builds a traced copy of the library in a scratch directory:
   python tools/fused_trace.py [--tracks 1024]"""
import argparse, ctypes as C, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=1024)
ap.add_argument("--child", action="store_true")
a = ap.parse_args()
if not a.child:
    td = tempfile.mkdtemp(prefix="vt_trace_")
    env = dict(os.environ, VT_LIB_DIR=td, NVCC_EXTRA="-DVT_FUSED_TRACE")
    sys.exit(subprocess.run([sys.executable, __file__, "--child", "--tracks", str(a.tracks)], env=env).returncode)
import numpy as np, torch
from oracle import vt_oracle as O
from vittracker_b200 import BatchedTracker, FramePool, load_cfg, _lib
n, F, H, W = a.tracks, 64, 720, 1280
sd = O.make_state_dict(seed=1, stress=True)
frames = np.concatenate([O.synth_frames(F // 2, H, W, seed=51, smooth=True), O.synth_frames(F // 2, H, W, seed=52)])
init_boxes, step_boxes = O.synth_boxes(n, H, W, seed=52), O.synth_boxes(n, H, W, seed=53)
fi, fs = np.arange(n) % F, (np.arange(n) + 1) % F
bt = BatchedTracker(load_cfg(), sd, max_tracks=n, chunk_tracks=n)
pool = FramePool(frames, bt.device)
bt.initialize(pool, torch.from_numpy(fi), init_boxes)
lib = _lib.load()
buf = (C.c_longlong * (8192 + 4096))(); cnt = C.c_int(0)
lib.vt_fused_trace_read(buf, C.byref(cnt))
for it in range(3):
    bt.set_state(step_boxes)
    bt.track(pool, torch.from_numpy(fs), update_state=False)
    torch.cuda.synchronize()
    lib.vt_fused_trace_read(buf, C.byref(cnt))
t = np.frombuffer(buf, dtype=np.int64)[:cnt.value]
cta = np.frombuffer(buf, dtype=np.int64)[8192:].reshape(1024, 4)
k = 8                                    # events per item
m = len(t) // k
ev = t[:m * k].reshape(m, k)
d = np.diff(ev, axis=1)
names = ["barrier 1", "wait conv2 (i-1)", "epilogue 2 (i-1)", "wait conv1", "epilogue 1", "barrier 2", "gather (i+1)"]
print(f"{m} items traced by CTA 0; cycles per item (mean / median / max), share of the item")
tot = (ev[1:, 0] - ev[:-1, 0]).mean()
for j, nm in enumerate(names):
    print(f"  {nm:32s} {d[1:, j].mean():8.0f} {np.median(d[1:, j]):8.0f} {d[1:, j].max():8.0f}   {d[1:, j].mean() / tot * 100:5.1f} %")
print(f"  item period {tot:.0f} cycles = {tot / 1.965e3:.2f} us; kernel ~ {tot * m / 1.965e6:.3f} ms")
g = cta[cta[:, 0] > 0]
t0 = g[:, 0].min()
print(f"{len(g)} CTAs on {len(set(g[:, 3]))} SMs: entry {np.percentile(g[:, 0] - t0, [0, 50, 100])} ns, loop start {np.percentile(g[:, 1] - t0, [0, 50, 100])} ns, "
      f"loop end {np.percentile(g[:, 2] - t0, [0, 10, 50, 90, 100])} ns; loop duration {np.percentile(g[:, 2] - g[:, 1], [0, 10, 50, 90, 100])} ns")
