"""GPU tests of the SURVEY 8(f) "next" rows against the checker (CPU oracle / reference golden), not against the product itself:
frame ingest (PipelinedFrameFeeder), the sample_target + Preprocessor drop-in (CropPreprocessor), the batched multi-sequence driver
(run_sequences) and the row-staged uploads of its backend."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import state_dict_from_npz
from oracle import vt_oracle as O

pytestmark = pytest.mark.gpu
TIE_GAP = 1e-5


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def close(a, b, atol=1e-2, rtol=1e-3):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= atol + rtol * np.abs(b)))


def test_crop_preprocessor_matches_reference_golden(golden_crops):
    """CropPreprocessor.crop = sample_target + Preprocessor.process of the reference (processing_utils.py:12-79, data_utils.py:11-17):
    uint8 patch, resize factor and attention mask against the fixtures recorded from the reference itself."""
    from vittracker_b200 import CropPreprocessor, Engine, load_cfg
    g = golden_crops
    eng = Engine(load_cfg(), max_tracks=1)
    eng.load_state_dict(O.make_state_dict(seed=0))
    cp = CropPreprocessor(eng)
    frame, boxes = g["frame"], g["boxes"]
    lut = torch.from_numpy(O.preprocess_lut())
    for i in list(range(0, len(boxes), 9)) + [int(k) for k in g["full_idx"]]:
        for factor, S, key in ((4.0, 256, "search"), (2.0, 128, "template")):
            patch, rf, mask, nested = cp.crop(frame, list(boxes[i]), factor, S)
            assert patch.dtype == np.uint8 and patch.shape == (S, S, 3) and sha(patch) == str(g[f"sha_{key}"][i]), (i, key)
            assert rf == float(g[f"rf_{key}"][i])
            if key == "search":
                assert mask.dtype == np.bool_ and sha(mask) == str(g["sha_mask_search"][i])
            want = torch.stack([lut[c][torch.from_numpy(patch[:, :, c].astype(np.int64))] for c in range(3)])[None]
            assert nested.tensors.shape == (1, 3, S, S) and torch.equal(nested.tensors.cpu(), want)
            assert nested.mask.shape == (1, S, S) and nested.mask.dtype == torch.bool
    for k, i in enumerate(g["full_idx"]):
        assert np.array_equal(cp.crop(frame, boxes[int(i)], 4.0, 256)[0], g["full_search"][k])        # numpy box accepted (tolist)
    with pytest.raises(Exception, match="Too small bounding box"):
        cp.crop(frame, [10.0, 10.0, 0.0, 0.0], 4.0, 256)


def _oracle_step(model, frames, fi, fs, init_boxes, step_boxes):
    win = O.hann2d(16, 16)
    H, W = frames.shape[1:3]
    res = []
    for i in range(len(init_boxes)):
        z = O.preprocess(O.sample_target_cv(frames[fi[i]], list(init_boxes[i]), 2.0, 128)[0])
        xp, rf, _ = O.sample_target_cv(frames[fs[i]], list(step_boxes[i]), 4.0, 256)
        out = model.forward(z, O.preprocess(xp))
        resp = (win * out["score_map"]).flatten()
        top = torch.topk(resp, 2).values
        pb = model.cal_bbox(resp.view(1, 1, 16, 16), out["size_map"], out["offset_map"])[0]
        pred = (pb * 256 / rf).tolist()
        res.append(dict(argmax=int(resp.argmax()), gap=float(top[0] - top[1]), conf=float(out["score_map"].max()),
                        state=O.clip_box(O.map_box_back(list(step_boxes[i]), pred, rf), H, W, margin=10)))
    return res


def test_pipelined_frame_feeder_against_oracle():
    """Frame ingest: frames + boxes uploaded from pinned host memory through the feeder's copy stream, double
    buffered over several steps with different frame sets; every step's result is checked against the CPU oracle on the frames that
    SHOULD have been resident for it (a stale or half-written pool would show as arg-max / box errors)."""
    from vittracker_b200 import BatchedTracker, PipelinedFrameFeeder, load_cfg
    n, F, H, W, steps = 24, 3, 360, 640, 5
    sd = O.make_state_dict(seed=13, stress=True)
    model = O.OracleModel(sd)
    sets = [O.synth_frames(F, H, W, seed=300 + s, smooth=(s % 2 == 0)) for s in range(steps)]
    host = [torch.from_numpy(f).pin_memory() for f in sets]
    init_boxes = O.synth_boxes(n, H, W, seed=41)
    step_boxes = [O.synth_boxes(n, H, W, seed=50 + s) for s in range(steps)]
    host_boxes = [torch.tensor(b).pin_memory() for b in step_boxes]
    bt = BatchedTracker(load_cfg(), sd, max_tracks=n)
    dev = bt.device
    feeder = PipelinedFrameFeeder(F, H, W, dev, max_tracks=n)
    fidx = np.arange(n) % F
    # initialise on the first set through the feeder as well
    feeder.upload(host[0])
    p0 = feeder.acquire()
    assert int(bt.initialize(p0, torch.from_numpy(fidx), init_boxes).abs().sum()) == 0
    feeder.release(p0)
    outs = []
    feeder.upload(host[0], host_boxes[0])
    for s in range(steps):
        fp = feeder.acquire()
        if s + 1 < steps:
            feeder.upload(host[s + 1], host_boxes[s + 1])           # overlaps this step
        bt.engine.tracks_set_state(fp.boxes[:n], first=0)
        out, det = bt.track(fp, torch.from_numpy((fidx + s) % F), update_state=True, detail=True)
        outs.append((out.clone(), det.clone()))
        feeder.release(fp)
    torch.cuda.synchronize()
    flips = 0
    for s in range(steps):
        out, det = outs[s][0].cpu().numpy(), outs[s][1].cpu().numpy()
        want = _oracle_step(model, np.concatenate([sets[0], sets[s]]), fidx, F + (fidx + s) % F, init_boxes, step_boxes[s])
        for i, w in enumerate(want):
            assert det[i, 6] == 0
            if w["gap"] < TIE_GAP:
                continue
            if int(det[i, 5]) != w["argmax"]:
                flips += 1
                continue
            assert close(out[i, :4], w["state"]), (s, i, out[i], w["state"])
            assert abs(out[i, 4] - w["conf"]) < 1e-4
    assert flips == 0


def test_run_sequences_against_oracle_tracker(tmp_path):
    """Batched multi-sequence driver against the CPU oracle's tracker run sequence by sequence (closed loop, ragged
    lengths, mixed frame sizes), and the written files against what the reference's writer makes of the oracle's boxes."""
    from vittracker_b200 import load_cfg
    from vittracker_b200.sequences import Sequence, run_sequences
    sd = O.make_state_dict(seed=31, stress=True, stable_size=True)
    model = O.OracleModel(sd)
    rng = np.random.default_rng(5)
    seqs = []
    for i, (n, (H, W)) in enumerate(zip([4, 1, 6, 3, 5, 7], [(360, 640), (240, 320), (360, 640), (300, 500), (240, 320), (720, 1280)])):
        frames = O.synth_frames(n, H, W, seed=40 + i, smooth=True)
        box = [float(rng.uniform(40, W - 140)), float(rng.uniform(40, H - 120)), float(rng.uniform(30, 90)), float(rng.uniform(30, 70))]
        seqs.append(Sequence(f"seq{i}", list(frames), box))
    res = run_sequences(seqs, load_cfg(), sd, slots=3, results_dir=str(tmp_path))
    # row-staged uploads (only the rows a crop can read) change nothing: bit-identical to uploading whole frames
    res_full = run_sequences(seqs, load_cfg(), sd, slots=3, row_staging=False)
    for s in seqs:
        assert res[s.name].get("target_bbox") == res_full[s.name].get("target_bbox"), s.name
    for s in seqs:
        if len(s.frames) == 1:
            assert "target_bbox" not in res[s.name]
            continue
        trk = O.OracleTracker(model, use_cv=True)
        trk.initialize(s.frames[0], s.init_info())
        want = [list(s.init_bbox)]
        tie = False
        for f in s.frames[1:]:
            want.append(list(trk.track(f, {})["target_bbox"]))
            top = torch.topk(trk.last["response"].flatten(), 2).values
            tie = tie or float(top[0] - top[1]) < TIE_GAP
        got = res[s.name]["target_bbox"]
        assert len(got) == len(want) and len(res[s.name]["time"]) == len(want)
        if tie:
            continue                                                 # a tie anywhere makes the rest of a closed loop incomparable
        assert close(np.array(got), np.array(want, dtype=np.float64)), (s.name, got, want)
        on_disk = np.loadtxt(os.path.join(str(tmp_path), s.name + ".txt"), delimiter="\t")
        want_int = np.array(want).astype(int)
        assert on_disk.shape == want_int.shape and np.abs(on_disk - want_int).max() <= 1       # truncation of boxes within 1e-2 px


def test_upload_frame_rect_writes_the_rectangle_and_nothing_else():
    """vt_upload_frame_rect (the batch-1 tracker's frame staging): rows x columns of a pageable host frame land at their place in the device
    frame buffer - through the device-side spread and through the strided fallback - and every other byte keeps its value; bad
    rectangles are refused."""
    from vittracker_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(12)
    st = torch.cuda.current_stream().cuda_stream
    for (H, W) in ((720, 1280), (333, 501)):
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        pin = torch.empty(H * W * 3, dtype=torch.uint8).pin_memory()
        stage = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda")
        rects = [(0, H, 0, W), (5, H - 7, 3, W - 9), (H // 3, H // 3 + 1, 10, 11), (0, H, W - 1, W), (17, 300, 40, 460), (1, 2, 0, W)]
        for use_stage in (True, False):
            for (ya, yb, xa, xb) in rects:
                dev = torch.full((H, W, 3), 7, dtype=torch.uint8, device="cuda")
                rc = lib.vt_upload_frame_rect(img.ctypes.data, H, W, ya, yb, xa, xb, pin.data_ptr(), stage.data_ptr() if use_stage else None,
                                              dev.data_ptr(), st)
                assert rc == 0
                torch.cuda.synchronize()
                want = np.full((H, W, 3), 7, dtype=np.uint8)
                want[ya:yb, xa:xb] = img[ya:yb, xa:xb]
                assert np.array_equal(dev.cpu().numpy(), want), (H, W, ya, yb, xa, xb, use_stage)
        dev = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
        for bad in ((5, 5, 0, W), (0, H + 1, 0, W), (-1, 4, 0, W), (0, H, 9, 9), (0, H, 0, W + 1)):
            assert lib.vt_upload_frame_rect(img.ctypes.data, H, W, *bad, pin.data_ptr(), stage.data_ptr(), dev.data_ptr(), st) != 0
