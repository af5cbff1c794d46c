"""Per-item cycle trace of conv4's control warp and of one epilogue warp (CTA 0): builds a traced copy of the library in a scratch directory.
   python tools/conv_trace.py"""
import ctypes as C, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--child" not in sys.argv:
    td = tempfile.mkdtemp(prefix="vt_trace_")
    sys.exit(subprocess.run([sys.executable, __file__, "--child"], env=dict(os.environ, VT_LIB_DIR=td, NVCC_EXTRA="-DVT_CONV_TRACE")).returncode)
import numpy as np, torch
from oracle import vt_oracle as O
from vittracker_b200 import BatchedTracker, FramePool, load_cfg, _lib
n, F, H, W = 1024, 64, 720, 1280
sd = O.make_state_dict(seed=1, stress=True)
frames = np.concatenate([O.synth_frames(F // 2, H, W, seed=51, smooth=True), O.synth_frames(F // 2, H, W, seed=52)])
init_boxes, step_boxes = O.synth_boxes(n, H, W, seed=52), O.synth_boxes(n, H, W, seed=53)
fi, fs = np.arange(n) % F, (np.arange(n) + 1) % F
bt = BatchedTracker(load_cfg(), sd, max_tracks=n, chunk_tracks=n)
pool = FramePool(frames, bt.device)
bt.initialize(pool, torch.from_numpy(fi), init_boxes)
lib = _lib.load()
buf = (C.c_longlong * 6144)(); cnt = (C.c_int * 3)()
lib.vt_conv_trace_read(buf, cnt)
for it in range(3):
    bt.set_state(step_boxes)
    bt.track(pool, torch.from_numpy(fs), update_state=False)
    torch.cuda.synchronize()
    lib.vt_conv_trace_read(buf, cnt)
a = np.frombuffer(buf, dtype=np.int64).reshape(3, 2048)
mma = a[0, :cnt[0] // 4 * 4].reshape(-1, 4)
d = np.diff(mma, axis=1)
print(f"MMA warp, {len(mma)} items; cycles per item (mean / max)")
for j, nm in enumerate(["wait: operands of item i landed", "wait: accumulators free (epilogue of i-2)", "issue MMAs + commits"]):
    print(f"  {nm:46s} {d[1:, j].mean():8.0f} {d[1:, j].max():8.0f}")
print(f"  item period {np.diff(mma[:, 0]).mean():.0f} cycles")
pr = a[2, :cnt[2] // 3 * 3].reshape(-1, 3)
dp = np.diff(pr, axis=1)
print(f"copy warp, {len(pr)} items: wait for the stage {dp[2:, 0].mean():.0f}, arm + start the copies {dp[:, 1].mean():.0f} cycles")
ep = a[1, :cnt[1] // 3 * 3].reshape(-1, 3)
de = np.diff(ep, axis=1)
print(f"epilogue warp 0 (accumulator set 0), {len(ep)} items: wait for accumulators {de[:, 0].mean():.0f}, epilogue {de[:, 1].mean():.0f} cycles; period {np.diff(ep[:, 0]).mean():.0f}")
