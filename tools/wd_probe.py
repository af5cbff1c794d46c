"""Development aid: steps under the in-kernel watchdog build (-DVT_TC_WATCHDOG); prints the stuck waits if a kernel deadlocked."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vt_oracle as O
from vittracker_b200 import BatchedTracker, FramePool, load_cfg, _lib
n, F, steps = 1024, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 20
bt = BatchedTracker(load_cfg(), O.make_state_dict(seed=1, stress=True), max_tracks=n, chunk_tracks=1024)
pool = FramePool(O.synth_frames(F, 720, 1280, seed=1000), bt.device)
boxes = O.synth_boxes(n, 720, 1280, seed=2000)
fidx = torch.arange(n, device=bt.device) % F
assert int(bt.initialize(pool, fidx, boxes).abs().sum()) == 0
lib = _lib.load()
rec = (C.c_int * (256 * 6))(); cnt = C.c_int(0)
for s in range(steps):
    bt.track(pool, (fidx + s) % F, update_state=False)
    torch.cuda.synchronize()
    lib.vt_tc_watchdog_read(rec, C.byref(cnt))
    if cnt.value:
        a = np.frombuffer(rec, dtype=np.int32).reshape(256, 6)[:min(cnt.value, 256)]
        print("WATCHDOG step", s, "records", cnt.value)
        for r in a[np.lexsort((a[:, 1], a[:, 0]))][:120]:
            print("  block", r[0], "warp", r[1], "bar", hex(r[2]), "parity", r[3])
        sys.exit(3)
print("no deadlock in", steps, "steps")
