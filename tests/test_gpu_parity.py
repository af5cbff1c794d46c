"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors recorded
from the reference.  Tolerances are the north star's: crop pixels and arg-max index bit-exact, maps
and boxes within 1e-2 abs / 1e-3 rel of the fp32 reference (the asserts below are tighter where the
fp32 kernels allow it, so that regressions show early)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import state_dict_from_npz
from oracle import vt_oracle as O

pytestmark = pytest.mark.gpu

ABS_TOL, REL_TOL = 1e-2, 1e-3          # north-star tolerance for maps and boxes
TIE_GAP = 1e-5                         # oracle top-1 - top-2 below this is a tie


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def close(a, b, atol=ABS_TOL, rtol=REL_TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= atol + rtol * np.abs(b)))


@pytest.fixture(scope="module")
def cfg():
    from vittracker_b200 import load_cfg
    return load_cfg()


@pytest.fixture(scope="module")
def stress_sd(golden_model):
    return state_dict_from_npz(golden_model)


# fp32 CUDA-core kernels, tcgen05 kernels (scores q k^T as one fp16 pass), tcgen05 kernels with three-term scores: same parity bar
BLOCKS = ["simt", "tcgen05", "tcgen05_3term"]


@pytest.fixture(scope="module", params=BLOCKS)
def engine(request, cfg, stress_sd):
    from vittracker_b200.engine import Engine
    e = Engine(cfg, max_tracks=1024, chunk_tracks=128, blocks_impl=request.param)
    e.load_state_dict(stress_sd)
    return e


def _dev_inputs(engine, frame, boxes):
    dev = engine.device
    H, W = frame.shape[:2]
    n = len(boxes)
    f = torch.from_numpy(np.ascontiguousarray(frame)).to(dev).reshape(-1)
    return (f, torch.zeros(n, dtype=torch.int64, device=dev),
            torch.tensor([[H, W]] * n, dtype=torch.int32, device=dev),
            torch.tensor(np.asarray(boxes, dtype=np.float64), device=dev).contiguous())


# ------------------------------------------------------------------------------------------------
# K1: crop / resize / normalise
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("factor,S,key", [(4.0, 256, "search"), (2.0, 128, "template")])
def test_crop_bit_exact_against_reference_golden(engine, golden_crops, factor, S, key):
    g = golden_crops
    frame, boxes = g["frame"], g["boxes"]
    f, off, hw, bx = _dev_inputs(engine, frame, boxes)
    out = engine.crop_normalize(f, off, hw, bx, factor, S, want_u8=True, want_mask=True)
    u8 = out["u8"].cpu().numpy()
    assert out["status"].cpu().numpy().tolist() == [0] * len(boxes)
    bad = [i for i in range(len(boxes)) if sha(u8[i]) != str(g[f"sha_{key}"][i])]
    assert not bad, f"{len(bad)} / {len(boxes)} crops differ from the reference, first {bad[:5]}"
    assert np.array_equal(out["resize_factor"].cpu().numpy(), g[f"rf_{key}"])
    if key == "search":
        mask = out["mask"].cpu().numpy().astype(bool)
        badm = [i for i in range(len(boxes)) if sha(mask[i]) != str(g["sha_mask_search"][i])]
        assert not badm, f"masks differ for {badm[:5]}"
    # normalised fp32 is bit-identical to Preprocessor.process of the u8 crop
    lut = torch.from_numpy(O.preprocess_lut())
    t = out["tensors"].cpu()
    for i in range(0, len(boxes), 17):
        want = torch.stack([lut[c][torch.from_numpy(u8[i][:, :, c].astype(np.int64))] for c in range(3)])
        assert torch.equal(t[i], want)


def test_crop_720p_random_boxes_and_status(engine):
    frame = O.synth_frames(1, 720, 1280, seed=31)[0]
    boxes = O.synth_boxes(48, 720, 1280, seed=32)
    boxes = np.concatenate([boxes, [[100, 100, 0, 0], [5000, 5000, 20, 20], [-400, 300, 30, 30], [1279, 719, 1, 1]]])
    f, off, hw, bx = _dev_inputs(engine, frame, boxes)
    out = engine.crop_normalize(f, off, hw, bx, 4.0, 256, want_u8=True, want_mask=True)
    u8, st = out["u8"].cpu().numpy(), out["status"].cpu().numpy()
    mk = out["mask"].cpu().numpy().astype(bool)
    for i, b in enumerate(boxes):
        w, h = b[2], b[3]
        import math
        if math.ceil(math.sqrt(w * h) * 4.0) < 1:
            assert st[i] == 1
        elif not O.crop_in_domain(b, 4.0, 720, 1280):
            assert st[i] == 2
        else:
            p, r, m = O.sample_target_spec(frame, list(b), 4.0, 256)
            assert st[i] == 0 and np.array_equal(u8[i], p), f"box {i} {b}: {np.abs(u8[i].astype(int) - p).max()}"
            assert np.array_equal(mk[i], m)


# ------------------------------------------------------------------------------------------------
# K2-K4: forward(z, x)
# ------------------------------------------------------------------------------------------------
def test_forward_against_reference_golden(engine, golden_model):
    g = golden_model
    z = torch.cat([O.preprocess(p) for p in g["z_patch"]])
    x = torch.cat([O.preprocess(p) for p in g["x_patch"]])
    out = engine.forward(z, x, taps=True)
    taps = out["taps"].cpu().numpy()
    names = ["tokens0", "tokens1", "tokens2", "tokens3", "tokens_norm"]
    report = {n: float(np.abs(taps[i] - g[f"tap::{n}"]).max()) for i, n in enumerate(names)}
    for k in ("score_map", "size_map", "offset_map", "pred_boxes"):
        report[k] = float(np.abs(out[k].cpu().numpy() - g[k]).max())
    print("max abs err vs reference:", report)
    for n in names:
        assert report[n] < 5e-4, report
    for k in ("score_map", "size_map", "offset_map", "pred_boxes"):
        assert close(out[k].cpu().numpy(), g[k]), report
        assert report[k] < 1e-4, report
    resp = torch.from_numpy(g["hann"]).to(out["score_map"].device) * out["score_map"]
    assert np.array_equal(resp.flatten(1).argmax(1).cpu().numpy(), g["argmax_windowed"])
    boxes = engine.cal_bbox(resp, out["size_map"], out["offset_map"]).cpu().numpy()
    assert close(boxes, g["pred_boxes_windowed"], atol=1e-4)


@pytest.mark.parametrize("blocks", BLOCKS)
@pytest.mark.parametrize("stress", [False, True])
def test_forward_random_batch_against_oracle(cfg, stress, blocks):
    from vittracker_b200.engine import Engine
    sd = O.make_state_dict(seed=5 if stress else 0, stress=stress)
    e = Engine(cfg, max_tracks=4, chunk_tracks=3, blocks_impl=blocks)   # chunk < batch: exercises chunking
    e.load_state_dict(sd)
    frame = O.synth_frames(1, 360, 480, seed=41, smooth=True)[0]
    boxes = O.synth_boxes(7, 360, 480, seed=42)
    zs = torch.cat([O.preprocess(O.sample_target_spec(frame, list(b), 2.0, 128)[0]) for b in boxes])
    xs = torch.cat([O.preprocess(O.sample_target_spec(frame, list(b), 4.0, 256)[0]) for b in boxes])
    torch.set_num_threads(max(1, torch.get_num_threads()))
    want = O.OracleModel(sd).forward(zs, xs)
    got = e.forward(zs, xs)
    for k in want:
        d = float((got[k].cpu() - want[k]).abs().max())
        assert close(got[k].cpu().numpy(), want[k].numpy()) and d < 1e-4, (k, d)
    assert torch.equal(got["score_map"].flatten(1).argmax(1).cpu(), want["score_map"].flatten(1).argmax(1))


@pytest.mark.parametrize("blocks", BLOCKS)
def test_model_dropin_surface(cfg, stress_sd, blocks):
    from vittracker_b200 import build_ostrack_dist
    net = build_ostrack_dist(cfg, blocks_impl=blocks)
    net.load_state_dict(stress_sd, strict=True)
    net = net.cuda().eval()
    z, x = torch.randn(1, 3, 128, 128), torch.randn(1, 3, 256, 256)
    out = net.forward(z=z, x=x)
    assert out["pred_boxes"].shape == (1, 1, 4) and out["score_map"].shape == (1, 1, 16, 16)
    assert out["size_map"].shape == (1, 2, 16, 16) and out["offset_map"].shape == (1, 2, 16, 16)
    want = O.OracleModel(stress_sd).forward(z, x)
    assert close(out["score_map"].cpu().numpy(), want["score_map"].numpy())
    b = net.box_head.cal_bbox(out["score_map"], out["size_map"], out["offset_map"])
    assert close(b.cpu().numpy(), want["pred_boxes"].view(-1, 4).numpy())
    with pytest.raises(RuntimeError):
        net.load_state_dict({"bogus": torch.zeros(1)}, strict=True)


# ------------------------------------------------------------------------------------------------
# tracker state machine
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("blocks", BLOCKS)
@pytest.mark.parametrize("tag", ["stress", "stable"])
def test_tracker_closed_loop_against_reference_golden(golden_track, golden_model, tag, blocks):
    from vittracker_b200 import get_tracker_class, parameters
    g = golden_track
    sd = state_dict_from_npz(golden_model) if tag == "stress" else state_dict_from_npz(g, "w_stable::")
    params = parameters("vit_48_h32_noKD")
    params.state_dict = sd
    params.blocks_impl = blocks
    trk = get_tracker_class()(params, "synthetic")
    frames = g["frames"]
    init = [float(v) for v in g[f"{tag}_init"]]
    assert trk.initialize(frames[0], {"init_bbox": init}) is None
    assert trk.state is init
    for t in range(1, 9):
        out = trk.track(frames[t % 4], {})
        assert out["target_bbox"] is trk.state and len(out["target_bbox"]) == 4
        want = g[f"{tag}_states"][t - 1]
        assert close(out["target_bbox"], want), (tag, t, out["target_bbox"], want.tolist())
        assert abs(float(out["confidence"]) - g[f"{tag}_conf"][t - 1]) < 1e-4
        # the device-side decode (used by the batched path) equals the host decode of the same frame
        assert close(trk.last_detail["device_box"], out["target_bbox"], atol=1e-9, rtol=0)


def test_tracker_errors(golden_track, stress_sd):
    from vittracker_b200 import get_tracker_class, parameters
    params = parameters("vit_48_h32_noKD")
    params.state_dict = stress_sd
    trk = get_tracker_class()(params, "synthetic")
    frame = golden_track["frames"][0]
    with pytest.raises(Exception, match="Too small bounding box"):
        trk.initialize(frame, {"init_bbox": [10, 10, 0, 0]})
    params2 = parameters("vit_48_h32_noKD")
    params2.checkpoint = "/nonexistent/ckpt.pth.tar"
    with pytest.raises(FileNotFoundError):
        get_tracker_class()(params2, "synthetic")


def test_checkpoint_file_roundtrip(tmp_path, stress_sd, golden_track):
    from vittracker_b200 import get_tracker_class, parameters
    ck = tmp_path / "OstrackDist_ep0300.pth.tar"
    torch.save({"epoch": 300, "net": stress_sd, "net_type": "OstrackDist"}, ck)
    params = parameters("vit_48_h32_noKD")
    params.checkpoint = str(ck)
    trk = get_tracker_class()(params, "synthetic")
    trk.initialize(golden_track["frames"][0], {"init_bbox": [140.0, 100.0, 36.0, 28.0]})
    out = trk.track(golden_track["frames"][1], {})
    assert close(out["target_bbox"], golden_track["stress_states"][0])


# ------------------------------------------------------------------------------------------------
# batched path
# ------------------------------------------------------------------------------------------------
def _oracle_open_loop(sd, frames, fidx_init, fidx_step, init_boxes, step_boxes):
    model = O.OracleModel(sd)
    win = O.hann2d(16, 16)
    res = []
    for i in range(len(init_boxes)):
        z = O.preprocess(O.sample_target_spec(frames[fidx_init[i]], list(init_boxes[i]), 2.0, 128)[0])
        xp, rf, _ = O.sample_target_spec(frames[fidx_step[i]], list(step_boxes[i]), 4.0, 256)
        out = model.forward(z, O.preprocess(xp))
        resp = (win * out["score_map"]).flatten()
        top = torch.topk(resp, 2).values
        pb = model.cal_bbox(resp.view(1, 1, 16, 16), out["size_map"], out["offset_map"]).view(-1, 4)
        pred = (pb.mean(0) * 256 / rf).tolist()
        H, W = frames.shape[1:3]
        state = O.clip_box(O.map_box_back(list(step_boxes[i]), pred, rf), H, W, margin=10)
        res.append(dict(argmax=int(resp.argmax()), gap=float(top[0] - top[1]), state=state,
                        conf=float(out["score_map"].max()), score=out["score_map"].flatten().numpy()))
    return res


@pytest.mark.parametrize("blocks,n", [("simt", 96), ("tcgen05", 96), ("tcgen05", 1200)])
def test_batched_open_loop_argmax_and_boxes_against_oracle(cfg, blocks, n):
    from vittracker_b200 import BatchedTracker, FramePool
    sd = O.make_state_dict(seed=9, stress=True)
    F = 3
    frames = O.synth_frames(F, 360, 640, seed=51, smooth=True)
    init_boxes = O.synth_boxes(n, 360, 640, seed=52)
    step_boxes = O.synth_boxes(n, 360, 640, seed=53)
    fi = np.arange(n) % F
    fs = (np.arange(n) + 1) % F
    bt = BatchedTracker(cfg, sd, max_tracks=n, chunk_tracks=40 if n < 200 else 256, blocks_impl=blocks)
    pool = FramePool(frames, bt.device)
    st = bt.initialize(pool, torch.from_numpy(fi), init_boxes)
    assert int(st.abs().sum()) == 0
    bt.set_state(step_boxes)
    out, det = bt.track(pool, torch.from_numpy(fs), update_state=True, detail=True)
    out, det = out.cpu().numpy(), det.cpu().numpy()
    new_state = bt.get_state().cpu().numpy()
    want = _oracle_open_loop(sd, frames, fi, fs, init_boxes, step_boxes)
    ties = flips = 0
    for i, w in enumerate(want):
        if w["gap"] < TIE_GAP:
            ties += 1
            continue
        if int(det[i, 5]) != w["argmax"]:
            flips += 1
            continue
        assert close(out[i, :4], w["state"]), (i, out[i], w["state"])
        assert abs(out[i, 4] - w["conf"]) < 1e-4
        assert np.array_equal(new_state[i], out[i, :4])
    print(f"open loop [{blocks}]: {n} tracks, {ties} ties excluded, {flips} arg-max flips")
    assert flips == 0
    maps = bt.engine.tracks_last_maps(0, n)
    assert close(maps["score_map"][5].flatten().cpu().numpy(), want[5]["score"], atol=1e-4)


@pytest.mark.parametrize("weights,n", [("default", 3414), ("stress1", 3413), ("stress2", 3413)])
def test_argmax_bit_exact_on_10k_frames(weights, n):
    """North-star gate on the benchmarked workload shape: Hann-weighted arg-max index identical to the reference algorithm on
    10 240 synthetic 720 x 1280 frames (64 distinct, half smooth / half white noise) split over three weight sets - default-init
    and two stress-init seeds - ties (oracle top-1 - top-2 < 1e-5) excepted and counted, decoded boxes within 1e-2 abs / 1e-3 rel.
    (tools/argmax_parity.py runs 10 240 frames per weight set; its records are under profiles/.)"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import argmax_parity
    torch.set_num_threads(os.cpu_count() or 1)
    res = argmax_parity.run(n=n, blocks="tcgen05", weights=weights)
    print(res)
    assert res["argmax_flips"] == 0, res
    assert res["boxes_outside_tolerance"] == 0, res
    assert res["status_nonzero"] == 0, res
    assert res["ties_excluded"] < 0.01 * res["frames"], res


@pytest.mark.parametrize("blocks", BLOCKS)
def test_batched_full_size_properties(cfg, stress_sd, blocks):
    """BASELINE config 3 size (1024 concurrent tracks): size-independent properties."""
    from vittracker_b200 import BatchedTracker, FramePool
    n, F = 1024, 4
    frames = O.synth_frames(F, 720, 1280, seed=61)
    boxes = O.synth_boxes(n, 720, 1280, seed=62)
    fidx = torch.arange(n) % F
    bt = BatchedTracker(cfg, stress_sd, max_tracks=n, chunk_tracks=256, blocks_impl=blocks)
    pool = FramePool(frames, bt.device)
    assert int(bt.initialize(pool, fidx, boxes).abs().sum()) == 0
    a = bt.track(pool, (fidx + 1) % F, update_state=False).clone()
    b = bt.track(pool, (fidx + 1) % F, update_state=False).clone()
    assert torch.equal(a, b), "step is not deterministic / idempotent without a state update"
    assert torch.isfinite(a).all()
    H, W = 720, 1280
    x, y, w, h = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    assert (x >= 0).all() and (y >= 0).all() and (w >= 10).all() and (h >= 10).all()
    assert (x <= W - 10).all() and (y <= H - 10).all() and (x + w <= W + 1e-9).all() and (y + h <= H + 1e-9).all()
    assert ((a[:, 4] >= 1e-4) & (a[:, 4] <= 0.9999 + 1e-7)).all()
    # permutation equivariance: tracks are independent, so reversing the batch reverses the result
    perm = torch.arange(n - 1, -1, -1)
    bt2 = BatchedTracker(cfg, stress_sd, max_tracks=n, chunk_tracks=96, blocks_impl=blocks)
    bt2.initialize(pool, fidx[perm], boxes[perm.numpy()])
    c = bt2.track(pool, ((fidx + 1) % F)[perm], update_state=False)
    assert torch.equal(c, a[perm.to(a.device)]), "result depends on batch position / chunking"
    # a sample of tracks against the oracle at full frame size
    want = _oracle_open_loop(stress_sd, frames, fidx.numpy()[:6], ((fidx + 1) % F).numpy()[:6], boxes[:6], boxes[:6])
    for i, wnt in enumerate(want):
        if wnt["gap"] >= TIE_GAP:
            assert close(a[i, :4].cpu().numpy(), wnt["state"]), (i, a[i], wnt["state"])


@pytest.mark.parametrize("blocks", BLOCKS)
def test_batched_closed_loop_matches_single_tracker(cfg, golden_track, blocks):
    from vittracker_b200 import BatchedTracker, FramePool, get_tracker_class, parameters
    g = golden_track
    sd = state_dict_from_npz(g, "w_stable::")
    frames = g["frames"]
    bt = BatchedTracker(cfg, sd, max_tracks=2, blocks_impl=blocks)
    pool = FramePool(frames, bt.device)
    init = np.array([g["stable_init"], g["stable_init"] + [3, -2, 4, 1]])
    bt.initialize(pool, torch.zeros(2, dtype=torch.int64), init)
    params = parameters("vit_48_h32_noKD")
    params.state_dict = sd
    params.blocks_impl = blocks
    singles = []
    for k in range(2):
        t = get_tracker_class()(params, "synthetic")
        t.initialize(frames[0], {"init_bbox": [float(v) for v in init[k]]})
        singles.append(t)
    for t in range(1, 9):
        out = bt.track(pool, torch.full((2,), t % 4, dtype=torch.int64)).cpu().numpy()
        for k in range(2):
            s = singles[k].track(frames[t % 4], {})
            assert close(out[k, :4], s["target_bbox"], atol=1e-9, rtol=0), (t, k, out[k], s["target_bbox"])
    assert close(out[0, :4], g["stable_states"][7])


# ------------------------------------------------------------------------------------------------
# multi-GPU: tracks sharded over ranks, one NCCL all-gather of the boxes (needs >= 2 GPUs)
# ------------------------------------------------------------------------------------------------
_NCCL_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VT_ROOT"])
from oracle import vt_oracle as O
from vittracker_b200 import BatchedTracker, FramePool, ShardedTracker, load_cfg, shard_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%s" % os.environ["VT_PORT"], rank=rank, world_size=world,
                        device_id=torch.device("cuda", rank))
total, F = 37, 3                                   # ragged split on purpose
cfg = load_cfg()
sd = O.make_state_dict(seed=4, stress=True)
frames = O.synth_frames(F, 240, 320, seed=5, smooth=True)
boxes = O.synth_boxes(total, 240, 320, seed=6)
lo, hi = shard_range(total, rank, world)
bt = BatchedTracker(cfg, sd, max_tracks=hi - lo)
pool = FramePool(frames, bt.device)
fidx = torch.arange(total) % F
bt.initialize(pool, fidx[lo:hi], boxes[lo:hi])
sh = ShardedTracker(total, bt)
local = bt.track(pool, ((fidx + 1) % F)[lo:hi])
full = sh.gather(local).cpu()
assert full.shape == (total, 5)
if rank == 0:                                      # single-GPU run of all tracks must give the same boxes
    ref = BatchedTracker(cfg, sd, max_tracks=total)
    ref.initialize(pool, fidx, boxes)
    want = ref.track(pool, (fidx + 1) % F).cpu()
    assert torch.equal(full, want), (full - want).abs().max()
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_tracks_nccl_gather_matches_single_gpu(tmp_path):
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER)
    port = str(29600 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", VT_ROOT=root, VT_PORT=port, MASTER_ADDR="127.0.0.1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0 and "OK" in out, out[-3000:]


# ------------------------------------------------------------------------------------------------
# edge cases of the batched path: empty / ragged batches, extreme frame sizes, flagged tracks, odd row pitch
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("blocks", BLOCKS)
def test_batched_edge_cases_against_oracle(cfg, blocks):
    from vittracker_b200 import BatchedTracker
    sd = O.make_state_dict(seed=12, stress=True)
    model = O.OracleModel(sd)
    dev = torch.device("cuda", torch.cuda.current_device())
    # frames of very different sizes in one buffer: tiny, odd width (row pitch not a multiple of 4), 4K
    shapes = [(48, 64), (241, 321), (2160, 3840)]
    frames = [O.synth_frames(2, H, W, seed=70 + i, smooth=(i != 1)) for i, (H, W) in enumerate(shapes)]
    flat = np.concatenate([f.reshape(-1) for f in frames])
    base = np.cumsum([0] + [f.size for f in frames])[:-1]
    buf = torch.from_numpy(flat).to(dev)
    #          frame set, init box                       (tracks 0-1 tiny, 2-4 odd pitch incl. a border-touching one, 5-6 4K)
    tracks = [(0, [12.0, 10.0, 20.0, 16.0]), (0, [30.0, 20.0, 9.0, 11.0]),
              (1, [100.5, 80.25, 40.0, 30.0]), (1, [0.0, 200.0, 25.0, 37.0]), (1, [290.0, 10.0, 30.0, 50.0]),
              (2, [1800.0, 900.0, 400.0, 300.0]), (2, [3700.0, 2000.0, 90.0, 120.0])]
    n = len(tracks)
    bt = BatchedTracker(cfg, sd, max_tracks=n + 2, chunk_tracks=3, blocks_impl=blocks)          # ragged chunks: 3 + 3 + 1
    hw = torch.tensor([list(shapes[s]) for s, _ in tracks], dtype=torch.int32, device=dev)
    boxes = torch.tensor([b for _, b in tracks], dtype=torch.float64, device=dev)

    def offsets(t):
        return torch.tensor([int(base[s]) + t * int(np.prod(shapes[s])) * 3 for s, _ in tracks], dtype=torch.int64, device=dev)

    # an empty batch is a no-op
    empty = bt.engine.tracks_init(buf, offsets(0)[:0], hw[:0], boxes[:0], first=0)
    assert empty.numel() == 0
    st = bt.engine.tracks_init(buf, offsets(0), hw, boxes, first=0)
    assert st.cpu().tolist() == [0] * n
    out, det = bt.engine.tracks_step(buf, offsets(1), hw, first=0, n=n, update_state=True, detail=True)
    out, det = out.cpu().numpy(), det.cpu().numpy()
    for i, (s, b) in enumerate(tracks):
        trk = O.OracleTracker(model)
        trk.initialize(frames[s][0], {"init_bbox": list(b)})
        want = trk.track(frames[s][1], {})
        resp = trk.last["response"].flatten()
        top = torch.topk(resp, 2).values
        if float(top[0] - top[1]) < TIE_GAP:
            continue
        assert int(det[i, 5]) == int(resp.argmax()), (i, shapes[s])
        assert close(out[i, :4], want["target_bbox"]), (i, out[i], want["target_bbox"])
    # a track whose state was forced out of the image is flagged (confidence -1, state kept), its neighbours are unaffected
    state = bt.engine.tracks_get_state(0, n).clone()
    bad = state.clone()
    bad[2] = torch.tensor([5000.0, 5000.0, 20.0, 20.0], dtype=torch.float64)
    bad[5] = torch.tensor([10.0, 10.0, 0.0, 0.0], dtype=torch.float64)
    bt.engine.tracks_set_state(bad.contiguous(), first=0)
    out2 = bt.engine.tracks_step(buf, offsets(0), hw, first=0, n=n, update_state=True).cpu().numpy()
    assert out2[2, 4] == -1.0 and out2[5, 4] == -1.0
    assert np.array_equal(out2[2, :4], [5000.0, 5000.0, 20.0, 20.0])
    bt.engine.tracks_set_state(state.contiguous(), first=0)
    ref = bt.engine.tracks_step(buf, offsets(0), hw, first=0, n=n, update_state=False).cpu().numpy()
    for i in (0, 1, 3, 4, 6):
        assert np.array_equal(out2[i], ref[i]), i
    # sub-range step: tracks [3, 6) only, results land at the front of the output
    part = bt.engine.tracks_step(buf, offsets(0)[3:6].contiguous(), hw[3:6].contiguous(), first=3, n=3, update_state=False).cpu().numpy()
    assert np.array_equal(part, ref[3:6])
