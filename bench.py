#!/usr/bin/env python
"""Benchmark of the VitTracker per-frame hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this implementation (N>1: under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

A "step" advances every track by one frame: crop + normalise + stem + 3 ViT blocks + head + Hann /
arg-max / box decode (+ for N>1 one NCCL all-gather of the boxes).  Workload per GPU (BASELINE.json
configs[2]; configs[3] at N=8): 1024 concurrent synthetic tracks over 64 distinct 720x1280 uint8
frames resident in HBM (177 MB > L2), open-loop seeded boxes re-seeded every step, stress-init
weights.  Prints ONE JSON line (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_JSON_OUT = sys.stdout
METRIC = "tracked frames/sec over B concurrent tracks"
UNIT = "frames/s"
FLOP_BLOCKS = 112.07e6        # algorithmic FLOP per tracked frame in the 3 ViT blocks
FLOP_STEM = 21.23e6
FLOP_HEAD = 30.53e6
FRAME_H, FRAME_W = 720, 1280


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(blocks, items_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the block kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, bytes per track), scaled to this launch size."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            per_track = json.load(f)[blocks]["dram_bytes_per_track"]
        return per_track * items_per_launch
    except Exception:
        return None


def hbm_rooflines(boxes_xywh, stages, peaks, S=256, factor=4.0):
    """The HBM-bound stages against the measured copy bandwidth (north star: achieved HBM GB/s for the crop and the head), on
    ALGORITHMIC bytes.  Stem (crop + resize + normalise fused into conv1, conv2-4): read = the unique source bytes the
    crops touch, 3 * min(crop_sz^2, 4 S^2) per track clipped to the image, write = the 256 x 48 fp32 search tokens.  Head + decode:
    read = the tokens, write = (x, y, w, h, conf)."""
    b = np.asarray(boxes_xywh, dtype=np.float64)
    crop_sz = np.ceil(np.sqrt(b[:, 2] * b[:, 3]) * factor)
    x1 = np.rint(b[:, 0] + 0.5 * b[:, 2] - 0.5 * crop_sz)
    y1 = np.rint(b[:, 1] + 0.5 * b[:, 3] - 0.5 * crop_sz)
    wx = np.clip(np.minimum(x1 + crop_sz, FRAME_W - 1) - np.maximum(x1, 0), 0, None)       # columns / rows 0 .. W-2 / H-2 are readable
    wy = np.clip(np.minimum(y1 + crop_sz, FRAME_H - 1) - np.maximum(y1, 0), 0, None)
    frac_in = np.where(crop_sz > 0, (wx * wy) / np.maximum(crop_sz * crop_sz, 1), 0.0)
    src = 3.0 * np.minimum(crop_sz * crop_sz, 4.0 * S * S) * frac_in
    tokens = 256 * 48 * 4
    per_track = {"stem": float(src.mean()) + tokens, "head": tokens + 5 * 4}
    out = {}
    for k, bytes_per_track in per_track.items():
        st = stages[k]
        launches = max(1, st["launches"])
        ms, items = st["ms"] / launches, st["items"] / launches
        achieved = bytes_per_track * items / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f)[k]["dram_bytes_per_track"] * items
        except Exception:
            pass
        out[k] = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                  "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_track * items, "avg_launch_ms": ms}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------------
# CPU path (the reference's algorithm; oracle port) - used for cpu_baseline and --impl reference
# --------------------------------------------------------------------------------------------------
class CpuPath:
    """Reference per-frame path on the host: cv2 crop + normalise + PyTorch fp32 forward (template
    stem re-run every frame, as the reference does) + Hann / arg-max / decode / clip.  Tracks are
    processed in groups of `group` so that the forward is batched (best CPU throughput)."""

    def __init__(self, sd, frames, group=16):
        from oracle import vt_oracle as O
        self.O = O
        self.model = O.OracleModel(sd)
        self.win = O.hann2d(16, 16)
        self.frames = frames
        self.group = group

    def prepare(self, init_boxes):
        O = self.O
        self.z = torch.cat([O.preprocess(O.sample_target_cv(self.frames[i % len(self.frames)], list(b), 2.0, 128)[0])
                            for i, b in enumerate(init_boxes)])

    def step(self, boxes, t):
        O = self.O
        n = len(boxes)
        out_states = []
        for g0 in range(0, n, self.group):
            idx = range(g0, min(n, g0 + self.group))
            crops, rfs = [], []
            for i in idx:
                p, rf, _ = O.sample_target_cv(self.frames[(i + t) % len(self.frames)], list(boxes[i]), 4.0, 256)
                crops.append(O.preprocess(p)); rfs.append(rf)
            out = self.model.forward(self.z[g0:g0 + len(crops)], torch.cat(crops))
            resp = self.win * out["score_map"]
            pb = self.model.cal_bbox(resp, out["size_map"], out["offset_map"])
            for k, i in enumerate(idx):
                pred = (pb[k] * 256 / rfs[k]).tolist()
                out_states.append(O.clip_box(O.map_box_back(list(boxes[i]), pred, rfs[k]), FRAME_H, FRAME_W, margin=10))
        return out_states


def cpu_sample(sd, budget_s=12.0, tracks=64, frames_n=4):
    """Bounded sample of the same workload on the host cores: `tracks` tracks over 720p frames,
    repeated until ~budget_s seconds have passed.  Returns (frames/s, cores, description)."""
    from oracle import vt_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frames = O.synth_frames(frames_n, FRAME_H, FRAME_W, seed=100)
    cp = CpuPath(sd, frames)
    cp.prepare(O.synth_boxes(tracks, FRAME_H, FRAME_W, seed=101))
    cp.step(O.synth_boxes(tracks, FRAME_H, FRAME_W, seed=102), 0)          # warm-up
    done, t0, step = 0, time.perf_counter(), 0
    while True:
        cp.step(O.synth_boxes(tracks, FRAME_H, FRAME_W, seed=103 + step), step)
        done += tracks; step += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return done / el, torch.get_num_threads(), f"{done} tracked frames ({step} steps x {tracks} tracks, forward batched by 16) in {el:.1f} s"


def cpu_model_name() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_b1_sample(sd, threads, budget_s=2.0):
    """The reference's own shape of the path - one sequence, batch 1 (B = 1, all threads and one thread) - on the host:
    frames/s of crop + forward + decode for a single track."""
    from oracle import vt_oracle as O
    torch.set_num_threads(threads)
    frames = O.synth_frames(2, FRAME_H, FRAME_W, seed=110)
    cp = CpuPath(sd, frames, group=1)
    cp.prepare(O.synth_boxes(1, FRAME_H, FRAME_W, seed=111))
    boxes = O.synth_boxes(64, FRAME_H, FRAME_W, seed=112)
    cp.step(boxes[:1], 0)
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        cp.step(boxes[done % 64:done % 64 + 1], done)
        done += 1
    return done / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    from oracle import vt_oracle as O
    sd = O.make_state_dict(seed=1, stress=True)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tracks = 32                                        # bounded sample per step
    frames = O.synth_frames(4, FRAME_H, FRAME_W, seed=100)
    cp = CpuPath(sd, frames)
    cp.prepare(O.synth_boxes(tracks, FRAME_H, FRAME_W, seed=101))
    for w in range(args.warmup):
        cp.step(O.synth_boxes(tracks, FRAME_H, FRAME_W, seed=200 + w), w)
    t0 = time.perf_counter()
    for s in range(args.steps):
        cp.step(O.synth_boxes(tracks, FRAME_H, FRAME_W, seed=300 + s), s)
    el = time.perf_counter() - t0
    v = tracks * args.steps / el
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world, note=f"CPU sample: {tracks} tracks per step"),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{tracks} tracks x {args.steps} steps, cv2 crop + torch fp32 forward batched by 16 + decode"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def workload_config(args, world, note=None):
    c = {"workload": f"vit_48_h32_noKD, {args.tracks} concurrent synthetic tracks per GPU (BASELINE configs[2]; "
                     f"{args.tracks * world} total), {args.frames} distinct {FRAME_H}x{FRAME_W} uint8 frames per GPU in HBM, "
                     f"open-loop seeded boxes, stress-init weights",
         "tracks_per_gpu": args.tracks, "total_tracks": args.tracks * world, "frames_per_gpu": args.frames,
         "frame_hw": [FRAME_H, FRAME_W], "chunk_tracks": args.chunk, "blocks_impl": args.blocks,
         "parallelism": f"tracks sharded x{world}, all-gather of boxes" if world > 1 else "single GPU",
         "l2": "inputs larger than L2 (frame pool %.0f MB + per-chunk intermediates)" % (args.frames * FRAME_H * FRAME_W * 3 / 1e6)}
    if note:
        c["note"] = note
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tracks", type=int, default=1024, help="concurrent tracks per GPU")
    ap.add_argument("--frames", type=int, default=64, help="distinct frames resident per GPU")
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--blocks", default="tcgen05", choices=["simt", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--latency-frames", type=int, default=300, help="open-loop frames of the batch-1 latency leg (SURVEY C2 asks for 10000)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries the one JSON line and nothing else: library chatter (e.g. NCCL's version banner) goes to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from oracle import vt_oracle as O          # synthetic workload generators only (frames / boxes / weights)
    from vittracker_b200 import BatchedTracker, FramePool, ShardedTracker, load_cfg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = load_cfg()
    sd = O.make_state_dict(seed=1, stress=True)
    n, F, K, W = args.tracks, args.frames, args.steps, args.warmup
    frames = O.synth_frames(F, FRAME_H, FRAME_W, seed=1000 + rank)
    bt = BatchedTracker(cfg, sd, max_tracks=n, chunk_tracks=args.chunk, blocks_impl=args.blocks)
    pool = FramePool(frames, dev)
    sharded = ShardedTracker(n * world, bt) if world > 1 else None
    init_boxes = O.synth_boxes(n, FRAME_H, FRAME_W, seed=2000 + rank)
    fidx0 = torch.arange(n, device=dev) % F
    status = bt.initialize(pool, fidx0, init_boxes)
    assert int(status.abs().sum()) == 0
    # open-loop: a fresh seeded state per step, generated up front on the device
    nsets = min(K + W, 8)
    step_boxes = torch.stack([torch.tensor(O.synth_boxes(n, FRAME_H, FRAME_W, seed=3000 + 97 * rank + s)) for s in range(nsets)]).to(dev)

    # per-step frame offsets prepared up front: the timed loop launches only this library's kernels
    fidx_host = np.arange(n, dtype=np.int64) % F
    step_offsets = [torch.from_numpy(((fidx_host + t) % F) * pool.frame_bytes).to(dev) for t in range(F)]

    pending = [None]

    def finish_gather():
        if pending[0] is not None:
            pending[0].wait()                 # current stream waits for the gather of the previous step
            pending[0] = None

    def step(t):
        bt.engine.tracks_set_state(step_boxes[t % nsets], first=0)
        out = bt.track_offsets(pool.data, step_offsets[t % F], update_state=True)
        if sharded is not None:
            # the only exchange step: all-gather of (x, y, w, h, conf); it overlaps the next step's kernels
            finish_gather()
            pending[0], out = sharded.gather_async(out)
        return out

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for t in range(W):
        step(t)
    finish_gather()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = bt.engine.launch_count
    bt.engine.profile(True)
    bt.engine.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):
        step(W + t)
    finish_gather()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    stages = bt.engine.profile_read()
    bt.engine.profile(False)
    launches = bt.engine.launch_count - launches0 + (K if world > 1 else 0)
    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = n * world * K / (ms / 1e3)

    # ---- end to end through the public API with HOST buffers: every step uploads its 64 frames (pinned host
    # memory -> HBM on a copy stream, double-buffered so that the upload of step t+1 overlaps the compute of
    # step t), uploads the boxes, runs the step and reads the boxes back
    from vittracker_b200 import PipelinedFrameFeeder
    host_pools = [torch.from_numpy(O.synth_frames(F, FRAME_H, FRAME_W, seed=5000 + rank + k)).pin_memory() for k in range(2)]
    host_boxes = step_boxes.cpu().pin_memory()
    host_out = torch.empty((n * world if world > 1 else n, 5), dtype=torch.float64).pin_memory()
    feeder = PipelinedFrameFeeder(F, FRAME_H, FRAME_W, dev, max_tracks=n)

    prev = [None]

    def e2e_run(steps):
        feeder.upload(host_pools[0], host_boxes[0])
        for t in range(steps):
            fp = feeder.acquire()
            if t + 1 < steps:
                feeder.upload(host_pools[(t + 1) % 2], host_boxes[(t + 1) % nsets])
            bt.engine.tracks_set_state(fp.boxes, first=0)
            out = bt.track_offsets(fp.data, step_offsets[t % F], update_state=True)
            feeder.release(fp)
            if sharded is not None:
                # read back the PREVIOUS step's gathered boxes (its gather overlapped this step), then start this step's
                if pending[0] is not None:
                    finish_gather()
                    host_out.copy_(prev[0], non_blocking=True)
                pending[0], prev[0] = sharded.gather_async(out)
            else:
                host_out.copy_(out, non_blocking=True)
        if sharded is not None:
            finish_gather()
            host_out.copy_(prev[0], non_blocking=True)

    e2e_run(3)
    barrier()
    Ke = max(3, min(K, 10))
    e0.record()
    e2e_run(Ke)
    e1.record()
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_value = n * world * Ke / (float(tms.item()) / 1e3)
    h2d = F * FRAME_H * FRAME_W * 3 + n * 32
    d2h = host_out.numel() * 8

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    blk = stages["blocks"]
    blk_ms = blk["ms"] / max(1, blk["launches"])
    blk_items = blk["items"] / max(1, blk["launches"])
    achieved = FLOP_BLOCKS * blk_items / (blk_ms * 1e-3) / 1e12 if blk_ms > 0 else 0.0
    total_stage_ms = sum(s["ms"] for s in stages.values()) or 1.0
    crop = stages["crop"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "roofline": {"kernel": "blocks_simt_kernel" if args.blocks == "simt" else "blocks_tc_kernel", "bound": "tensor",
                     "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": ncu_traffic(args.blocks, blk_items),
                     "peak_source": peaks["source"] + ", bf16 dense sustained",
                     "algorithmic_flop_per_launch": FLOP_BLOCKS * blk_items, "avg_launch_ms": blk_ms,
                     "share_of_step": blk["ms"] / total_stage_ms},
        "stages": {k: {"ms_per_step": v["ms"] / K, "share": v["ms"] / total_stage_ms, "launches_per_step": v["launches"] / K}
                   for k, v in stages.items()},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "per step: H2D of all %d frames + boxes from pinned host memory (upload of step t+1 overlaps compute "
                        "of step t on a copy stream), step, D2H of the boxes; H2D link measured at 55 GB/s (tools/h2d_probe.py)" % F},
        "gpu_launches": int(launches),
    }
    if crop["ms"] > 0:
        line["stages"]["crop"]["note"] = "HBM-bound gather; see profiles/ for achieved GB/s"
    try:
        line["roofline_hbm"] = hbm_rooflines(O.synth_boxes(n, FRAME_H, FRAME_W, seed=3000 + 97 * rank), stages, peaks)
    except Exception as e:                             # secondary figures: never at the expense of the line
        line["roofline_hbm"] = {"error": f"{type(e).__name__}: {e}"}

    if world == 1 and not args.no_latency:
        line["latency_b1"] = latency_b1(cfg, sd, iters=args.latency_frames)
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_sample(sd)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "cpu_model": cpu_model_name(), "os_cpu_count": os.cpu_count()}
        try:                                           # the reference's own batch-1 shape, all threads and one thread
            line["cpu_baseline"]["batch1_frames_per_s"] = {"all_threads": cpu_b1_sample(sd, cores), "one_thread": cpu_b1_sample(sd, 1)}
        except Exception as e:
            line["cpu_baseline"]["batch1_frames_per_s"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line, default=float), file=_JSON_OUT, flush=True)     # default: NumPy scalars, should one slip in
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def latency_b1(cfg, sd, frames_n=8, iters=300, burst=8):
    """BASELINE configs[1]: batch-1 initialize() + track() loop through the drop-in tracker (host numpy
    frame in, Python list out), wall clock per call.  `iters` open-loop frames (state re-seeded per frame), then
    closed-loop bursts of `burst` frames after a re-initialisation (iters // 40 bursts), initialize() timed as well."""
    from oracle import vt_oracle as O
    from vittracker_b200 import get_tracker_class, parameters
    params = parameters("vit_48_h32_noKD")
    params.state_dict = sd
    trk = get_tracker_class()(params, "synthetic")
    frames = O.synth_frames(frames_n, FRAME_H, FRAME_W, seed=77)
    boxes = O.synth_boxes(iters + 20, FRAME_H, FRAME_W, seed=78)
    trk.initialize(frames[0], {"init_bbox": list(boxes[0])})
    lat = []
    for i in range(iters + 20):
        trk.state = list(boxes[i])                      # open loop: re-seeded state per frame
        t0 = time.perf_counter()
        trk.track(frames[i % frames_n], {})
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat[20:]) * 1e3
    out = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "frames": iters,
           "path": "Vit_dist.track(): host frame rows H2D + crop + forward + decode + D2H, blocking"}
    n_bursts = iters // 40
    if n_bursts > 0 and burst > 0:
        try:
            out.update(_latency_bursts(trk, frames, boxes, frames_n, n_bursts, burst))
        except Exception as e:                         # the headline line must not depend on this leg
            out["closed_loop"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def _latency_bursts(trk, frames, boxes, frames_n, n_bursts, burst):
    lat_c, lat_i = [], []
    for b in range(n_bursts):
        t0 = time.perf_counter()
        trk.initialize(frames[b % frames_n], {"init_bbox": list(boxes[b])})
        lat_i.append(time.perf_counter() - t0)
        for k in range(burst):                          # closed loop: the tracker follows its own state
            t0 = time.perf_counter()
            trk.track(frames[(b + 1 + k) % frames_n], {})
            lat_c.append(time.perf_counter() - t0)
    lat_c, lat_i = np.array(lat_c) * 1e3, np.array(lat_i) * 1e3
    return {"closed_loop": {"p50_ms": float(np.percentile(lat_c, 50)), "p99_ms": float(np.percentile(lat_c, 99)),
                            "frames": int(lat_c.size), "bursts": n_bursts, "burst": burst},
            "initialize": {"p50_ms": float(np.percentile(lat_i, 50)), "p99_ms": float(np.percentile(lat_i, 99)), "calls": n_bursts}}


if __name__ == "__main__":
    main()
