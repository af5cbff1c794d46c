"""Development aid: bench-like stress loop (open-loop steps with resident frames, then end-to-end steps with uploads overlapping) under a
watchdog that reports which pipeline stage is stuck if a synchronisation does not return.  Use under `timeout`."""
import argparse, ctypes as C, os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vt_oracle as O
from vittracker_b200 import BatchedTracker, FramePool, PipelinedFrameFeeder, load_cfg, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=1024)
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--blocks", default="tcgen05")
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--e2e", type=int, default=60)
a = ap.parse_args()
n, F = a.tracks, a.frames
cfg = load_cfg()
sd = O.make_state_dict(seed=1, stress=True)
frames = O.synth_frames(F, 720, 1280, seed=1000)
bt = BatchedTracker(cfg, sd, max_tracks=n, chunk_tracks=1024, blocks_impl=a.blocks)
dev = bt.device
pool = FramePool(frames, dev)
lib = _lib.load()
lib.vt_debug_pending.restype = C.c_int
lib.vt_debug_pending.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32]
beat = [time.time(), "start"]
def watchdog():
    while True:
        time.sleep(1.0)
        if time.time() - beat[0] > 8.0:
            buf = (C.c_int32 * 4096)()
            k = lib.vt_debug_pending(bt.engine.handle, buf, 4096)
            v = [buf[i] for i in range(k)]
            pend = [(i, x // 4, x % 4) for i, x in enumerate(v) if x % 4 != 3]
            print("HANG at", beat[1], "records", k, "unfinished (index, stage, started|finished<<1):", pend[:12], flush=True)
            os._exit(3)
threading.Thread(target=watchdog, daemon=True).start()
boxes = O.synth_boxes(n, 720, 1280, seed=2000)
fidx0 = torch.arange(n, device=dev) % F
assert int(bt.initialize(pool, fidx0, boxes).abs().sum()) == 0
step_boxes = torch.stack([torch.tensor(O.synth_boxes(n, 720, 1280, seed=3000 + s)) for s in range(8)]).to(dev)
fh = np.arange(n, dtype=np.int64) % F
offs = [torch.from_numpy(((fh + t) % F) * pool.frame_bytes).to(dev) for t in range(F)]
bt.engine.profile(True)
t0 = time.time()
for t in range(a.steps):
    bt.engine.tracks_set_state(step_boxes[t % 8], first=0)
    bt.track_offsets(pool.data, offs[t % F], update_state=True)
    if t % 10 == 9:
        beat[:] = [time.time(), f"resident step {t}"]
        torch.cuda.synchronize(); bt.engine.profile_read()
torch.cuda.synchronize(); bt.engine.profile_read()
print("resident steps OK", a.steps, round((time.time() - t0) / a.steps * 1e3, 3), "ms/step", flush=True)
host_pools = [torch.from_numpy(O.synth_frames(F, 720, 1280, seed=5000 + k)).pin_memory() for k in range(2)]
host_boxes = step_boxes.cpu().pin_memory()
host_out = torch.empty((n, 5), dtype=torch.float64).pin_memory()
feeder = PipelinedFrameFeeder(F, 720, 1280, dev, max_tracks=n)
beat[:] = [time.time(), "e2e start"]
feeder.upload(host_pools[0], host_boxes[0])
for t in range(a.e2e):
    fp = feeder.acquire()
    if t + 1 < a.e2e:
        feeder.upload(host_pools[(t + 1) % 2], host_boxes[(t + 1) % 8])
    bt.engine.tracks_set_state(fp.boxes, first=0)
    out = bt.track_offsets(fp.data, offs[t % F], update_state=True)
    feeder.release(fp)
    host_out.copy_(out, non_blocking=True)
    if t % 5 == 4:
        beat[:] = [time.time(), f"e2e step {t}"]
        torch.cuda.synchronize(); bt.engine.profile_read()
torch.cuda.synchronize()
print("e2e steps OK", a.e2e, flush=True)
