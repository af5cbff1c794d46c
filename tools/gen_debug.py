"""Development aid: widest-config forward vs oracle, per-stage taps, to localise a discrepancy."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vt_oracle as O
from vittracker_b200 import load_cfg
from vittracker_b200.model import build_ostrack_dist
c = dict(C=768, heads=12, depth=12, hc=256)
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 3
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = load_cfg()
cfg.MODEL.BACKBONE.CHANNELS, cfg.MODEL.BACKBONE.HEADS, cfg.MODEL.BACKBONE.DEPTH, cfg.MODEL.HEAD.NUM_CHANNELS = 768, 12, 12, 256
torch.set_num_threads(16)
sd = O.make_state_dict(seed=21, stress=True, C=768, depth=12, head_ch=256)
oracle = O.OracleModel(sd, depth=12, num_heads=12)
frames = O.synth_frames(2, 360, 640, seed=5, smooth=True)
boxes = O.synth_boxes(nt, 360, 640, seed=6)
z = torch.cat([O.preprocess(O.sample_target_cv(frames[0], list(b), 2.0, 128)[0]) for b in boxes])
x = torch.cat([O.preprocess(O.sample_target_cv(frames[1], list(b), 4.0, 256)[0]) for b in boxes])
taps = {}
want = oracle.forward(z, x, taps)
net = build_ostrack_dist(cfg, depth=12, max_tracks=nt, chunk_tracks=chunk)
net.load_state_dict(sd, strict=True)
got = net.cuda().forward(z=z, x=x, return_taps=True)
tp = got["taps"].cpu()
for i in range(nt):
    e0z = float((tp[0, i, :64] - taps["tokens0"][i, :64]).abs().max())
    e0x = float((tp[0, i, 64:] - taps["tokens0"][i, 64:]).abs().max())
    eb = [float((tp[b + 1, i] - taps[f"tokens{b + 1}"][i]).abs().max()) for b in (0, 5, 11)]
    en = float((tp[13, i] - taps["tokens_norm"][i]).abs().max())
    es = float((got["score_map"][i].cpu() - want["score_map"][i]).abs().max())
    print(f"track {i}: tokens0 z {e0z:.2e} x {e0x:.2e} | blocks {eb} | norm {en:.2e} | score {es:.2e}")
