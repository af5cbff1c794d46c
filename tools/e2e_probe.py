"""Where does the end-to-end step time go?  copies only / compute only / pipelined."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vt_oracle as O
from vittracker_b200 import BatchedTracker, FramePool, PipelinedFrameFeeder, load_cfg
dev = torch.device("cuda", 0)
cfg = load_cfg(); sd = O.make_state_dict(seed=1, stress=True)
n, F, H, W = 1024, 64, 720, 1280
frames = O.synth_frames(F, H, W, seed=1000)
bt = BatchedTracker(cfg, sd, max_tracks=n, chunk_tracks=1024)
pool = FramePool(frames, dev)
boxes = O.synth_boxes(n, H, W, seed=2000)
fidx = torch.arange(n, device=dev) % F
bt.initialize(pool, fidx, boxes)
step_boxes = torch.tensor(O.synth_boxes(n, H, W, seed=3000)).to(dev)
offs = pool.offsets(fidx)
host = [torch.from_numpy(O.synth_frames(F, H, W, seed=5000 + k)).pin_memory() for k in range(2)]
feeder = PipelinedFrameFeeder(F, H, W, dev)
host_out = torch.empty((n, 5), dtype=torch.float64).pin_memory()
def timeit(fn, reps=10):
    fn(3); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(reps); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def copies(k):
    for t in range(k):
        feeder.upload(host[t % 2]); fp = feeder.acquire(); feeder.release(fp)
def compute(k):
    for t in range(k):
        bt.engine.tracks_set_state(step_boxes, first=0)
        out = bt.track_offsets(pool.data, offs, update_state=True)
        host_out.copy_(out, non_blocking=True)
def piped(k):
    feeder.upload(host[0])
    for t in range(k):
        fp = feeder.acquire()
        if t + 1 < k: feeder.upload(host[(t + 1) % 2])
        bt.engine.tracks_set_state(step_boxes, first=0)
        out = bt.track_offsets(fp.data, offs, update_state=True)
        host_out.copy_(out, non_blocking=True)
        feeder.release(fp)
print("copies only   ms/step", timeit(copies))
print("compute only  ms/step", timeit(compute))
print("pipelined(10) ms/step", timeit(piped, 10))
print("pipelined(40) ms/step", timeit(piped, 40))
