"""CPU oracle for VitTracker's per-frame inference hot path (vit_dist / vit_48_h32_noKD).

TEST INFRASTRUCTURE ONLY.  Nothing under ``vittracker_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do, and only as the
checker / the reported CPU baseline - never as the thing measured or shipped.

It restates, function by function, what the reference computes on the path
``lib/test/tracker/vit_dist.py initialize()/track()`` + ``lib/models/vit_dist forward(z, x)``
(citations are relative to /root/reference):

* integer crop/resize     -> ``lib/train/data/processing_utils.py:12-79`` (+ OpenCV ``cv.resize`` u8 INTER_LINEAR)
* normalisation           -> ``lib/test/tracker/data_utils.py:6-17``
* Hann window             -> ``lib/test/utils/hann.py:6-16``
* conv stem               -> ``lib/models/vit_dist/vit_dist.py:10-54``
* ViT block               -> ``timm.models.vision_transformer.Block`` (EXTERNAL, un-vendored, unpinned:
                             ``install.sh:95``); in-repo restatements ``tracking/onnxexport.py:126-225`` and
                             ``lib/models/ostrack/vit.py:39-91`` were followed
* forward / forward_head  -> ``lib/models/vit_dist/vit_dist.py:77-100,122-153``
* CENTER head, cal_bbox   -> ``lib/models/layers/head.py:8-21,98-201``
* box decode / clip       -> ``lib/test/tracker/vit_dist.py:103-111,150-156``, ``lib/utils/box_ops.py:97-106``

Pinning status: the reference ships NO golden vectors, tests or fixtures, so
the oracle is pinned the other way the task allows: ``oracle/make_golden.py`` imports the reference's
own modules from /root/reference (through ``oracle/ref_shim.py``, which stubs the absent third-party
packages) and asserts this restatement reproduces the reference's outputs (crop bit-exact, tracker
boxes equal, model maps to float32 round-off); the fixtures it wrote are committed under
``tests/golden/``.  The timm ``Block`` itself is absent from /root/reference and from this image; at
that one boundary parity is "unpinned" against timm proper and rests on the authors' restatement.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# Static configuration of vit_48_h32_noKD  (experiments/vit_dist/vit_48_h32_noKD.yaml:56-64,89-92)
# ----------------------------------------------------------------------------------------------
EMBED_DIM = 48
NUM_HEADS = 1
DEPTH = 3            # build_ostrack_dist default, lib/models/vit_dist/vit_dist.py:159
MLP_RATIO = 4
HEAD_CHANNELS = 32
STRIDE = 16
TEMPLATE_SIZE, TEMPLATE_FACTOR = 128, 2.0
SEARCH_SIZE, SEARCH_FACTOR = 256, 4.0
FEAT_SZ = SEARCH_SIZE // STRIDE      # 16
LN_EPS = 1e-5
BN_EPS = 1e-5
MEAN = (0.485, 0.456, 0.406)         # lib/test/tracker/data_utils.py:8
STD = (0.229, 0.224, 0.225)          # lib/test/tracker/data_utils.py:9
TOWERS = ("ctr", "offset", "size")


# ----------------------------------------------------------------------------------------------
# R3  sample_target: crop geometry + u8 bilinear (integer spec)
# ----------------------------------------------------------------------------------------------
def crop_geometry(box, factor: float, out_sz: int) -> Tuple[int, int, int, float]:
    """(crop_sz, x1, y1, resize_factor) exactly as processing_utils.py:30-38,67 computes them."""
    x, y, w, h = [float(v) for v in box]
    crop_sz = math.ceil(math.sqrt(w * h) * factor)
    if crop_sz < 1:
        raise Exception('Too small bounding box.')          # processing_utils.py:32-33
    x1 = round(x + 0.5 * w - crop_sz * 0.5)                   # Python round: half-to-even
    y1 = round(y + 0.5 * h - crop_sz * 0.5)
    return crop_sz, x1, y1, out_sz / crop_sz


def _resize_taps_x(src: int, dst: int):
    """OpenCV resizeGeneric INTER_LINEAR u8 horizontal taps: (s0, s1, a0, a1) per output column."""
    scale = 1.0 / (dst / src)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s = np.where(lo, 0, s)
    f = np.where(lo, np.float32(0), f)
    hi = s >= src - 1
    s = np.where(hi, src - 1, s)
    f = np.where(hi, np.float32(0), f).astype(np.float32)
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int64)
    s1 = np.minimum(s + 1, src - 1)
    return s, s1, a0, a1


def _resize_taps_y(src: int, dst: int):
    """Vertical taps: weights from the UNCLAMPED fractional part, rows clamped afterwards."""
    scale = 1.0 / (dst / src)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    b0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int64)
    b1 = np.rint(f * np.float32(2048.0)).astype(np.int64)
    r0 = np.clip(s, 0, src - 1)
    r1 = np.clip(s + 1, 0, src - 1)
    return r0, r1, b0, b1


def sample_target_spec(im: np.ndarray, box, factor: float, out_sz: int):
    """NumPy integer restatement of sample_target(im, box, factor, output_sz) with mask=None
    (processing_utils.py:12-71): returns (patch u8 SxSx3, resize_factor, att_mask bool SxS).

    The padded crop is never materialised: P[j][i] = im[y1+j][x1+i] iff 0<=x1+i<=W-2 and
    0<=y1+j<=H-2, else 0 (constant border; the '+1' of x2_pad/y2_pad at :42,45 drops the last
    image row and column).  cv.resize u8 INTER_LINEAR is 11-bit fixed point.
    """
    H, W = im.shape[:2]
    crop_sz, x1, y1, resize_factor = crop_geometry(box, factor, out_sz)
    sx0, sx1, a0, a1 = _resize_taps_x(crop_sz, out_sz)
    ry0, ry1, b0, b1 = _resize_taps_y(crop_sz, out_sz)

    def gather_cols(rows_crop):
        # rows_crop: crop-space row index per output row; returns Hrow (S, S, 3) int64
        yy = y1 + rows_crop
        vy = (yy >= 0) & (yy <= H - 2)
        xa, xb = x1 + sx0, x1 + sx1
        va = (xa >= 0) & (xa <= W - 2)
        vb = (xb >= 0) & (xb <= W - 2)
        yyc = np.clip(yy, 0, H - 1)
        pa = im[yyc[:, None], np.clip(xa, 0, W - 1)[None, :], :].astype(np.int64)
        pb = im[yyc[:, None], np.clip(xb, 0, W - 1)[None, :], :].astype(np.int64)
        pa = pa * (vy[:, None] & va[None, :])[:, :, None]
        pb = pb * (vy[:, None] & vb[None, :])[:, :, None]
        return pa * a0[None, :, None] + pb * a1[None, :, None]

    h0 = gather_cols(ry0)
    h1 = gather_cols(ry1)
    out = (((b0[:, None, None] * (h0 >> 4)) >> 16) + ((b1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2
    patch = np.clip(out, 0, 255).astype(np.uint8)
    mask = att_mask_spec(H, W, crop_sz, x1, y1, out_sz)
    return patch, resize_factor, mask


def att_mask_spec(H: int, W: int, crop_sz: int, x1: int, y1: int, out_sz: int) -> np.ndarray:
    """att_mask of processing_utils.py:55-62,69: float64 ones with zeros on the valid region,
    cv.resize (float bilinear) then astype(bool).  A resized pixel is non-zero iff one of its
    bilinear taps with a non-zero weight lies on a padded pixel."""
    scale = 1.0 / (out_sz / crop_sz)
    d = np.arange(out_sz, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s = np.where(lo, 0, s)
    f = np.where(lo, np.float32(0), f)
    hi = s >= crop_sz - 1
    s = np.where(hi, crop_sz - 1, s)
    f = np.where(hi, np.float32(0), f)
    s1 = np.minimum(s + 1, crop_sz - 1)
    w0 = (np.float32(1.0) - f) != 0
    w1 = f != 0

    def pad1d(idx, origin, size):
        p = origin + idx
        return ~((p >= 0) & (p <= size - 2))

    px = (pad1d(s, x1, W) & w0) | (pad1d(s1, x1, W) & w1)       # column touches padding
    py = (pad1d(s, y1, H) & w0) | (pad1d(s1, y1, H) & w1)
    return py[:, None] | px[None, :]


def sample_target_cv(im: np.ndarray, box, factor: float, out_sz: int):
    """The reference's own recipe, line for line, calling OpenCV (processing_utils.py:24-71).
    Used to pin ``sample_target_spec``; cv2 is part of this image (4.13.0)."""
    import cv2 as cv
    x, y, w, h = [float(v) for v in box]
    crop_sz = math.ceil(math.sqrt(w * h) * factor)
    if crop_sz < 1:
        raise Exception('Too small bounding box.')
    x1 = round(x + 0.5 * w - crop_sz * 0.5)
    x2 = x1 + crop_sz
    y1 = round(y + 0.5 * h - crop_sz * 0.5)
    y2 = y1 + crop_sz
    x1_pad = max(0, -x1)
    x2_pad = max(x2 - im.shape[1] + 1, 0)
    y1_pad = max(0, -y1)
    y2_pad = max(y2 - im.shape[0] + 1, 0)
    im_crop = im[y1 + y1_pad:y2 - y2_pad, x1 + x1_pad:x2 - x2_pad, :]
    padded = cv.copyMakeBorder(im_crop, y1_pad, y2_pad, x1_pad, x2_pad, cv.BORDER_CONSTANT)
    Hh, Ww = padded.shape[0], padded.shape[1]
    att = np.ones((Hh, Ww))
    end_x, end_y = -x2_pad, -y2_pad
    if y2_pad == 0:
        end_y = None
    if x2_pad == 0:
        end_x = None
    att[y1_pad:end_y, x1_pad:end_x] = 0
    resize_factor = out_sz / crop_sz
    padded = cv.resize(padded, (out_sz, out_sz))
    att = cv.resize(att, (out_sz, out_sz)).astype(np.bool_)
    return padded, resize_factor, att


def crop_in_domain(box, factor: float, H: int, W: int) -> bool:
    """True when the reference's slicing (processing_utils.py:48) is well defined: the crop must
    overlap the image's reachable region [0, W-2] x [0, H-2] in at least one pixel.  Outside it the
    slice is empty or gets a negative stop that NumPy wraps: undefined behaviour
    that the new implementation rejects instead of reproducing."""
    x, y, w, h = [float(v) for v in box]
    if not (w * h >= 0):
        return False
    crop_sz = math.ceil(math.sqrt(w * h) * factor)
    if crop_sz < 1:
        return False
    x1 = round(x + 0.5 * w - crop_sz * 0.5)
    y1 = round(y + 0.5 * h - crop_sz * 0.5)
    return max(x1, 0) < min(x1 + crop_sz, W - 1) and max(y1, 0) < min(y1 + crop_sz, H - 1)


# ----------------------------------------------------------------------------------------------
# R4  Preprocessor.process
# ----------------------------------------------------------------------------------------------
def preprocess(patch_u8: np.ndarray) -> torch.Tensor:
    """uint8 HWC -> fp32 (1,3,S,S): ((x / 255.0) - mean) / std, data_utils.py:13-14 (div, sub, div)."""
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    t = torch.tensor(patch_u8).float().permute(2, 0, 1).unsqueeze(0)
    return ((t / 255.0) - mean) / std


def preprocess_lut() -> np.ndarray:
    """The 3x256 fp32 table the map above takes its values from (one entry per channel, byte)."""
    v = torch.arange(256, dtype=torch.float32).view(1, 1, 256, 1).expand(1, 3, 256, 1)
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    return (((v / 255.0) - mean) / std).reshape(3, 256).numpy().copy()


# ----------------------------------------------------------------------------------------------
# R5  Hann window
# ----------------------------------------------------------------------------------------------
def hann1d(sz: int) -> torch.Tensor:
    """hann.py:6-9 (centered=True)."""
    return 0.5 * (1 - torch.cos((2 * math.pi / (sz + 1)) * torch.arange(1, sz + 1).float()))


def hann2d(sz_y: int, sz_x: int) -> torch.Tensor:
    """hann.py:14-16 -> (1,1,sz_y,sz_x)."""
    return hann1d(sz_y).reshape(1, 1, -1, 1) * hann1d(sz_x).reshape(1, 1, 1, -1)


# ----------------------------------------------------------------------------------------------
# R6  parameters: names, shapes, synthetic initialisation
# ----------------------------------------------------------------------------------------------
def param_shapes(C: int = EMBED_DIM, depth: int = DEPTH, head_ch: int = HEAD_CHANNELS,
                 mlp_ratio: int = MLP_RATIO) -> Dict[str, Tuple[int, ...]]:
    """Every state_dict entry of build_ostrack_dist(cfg) in eval mode."""
    s: Dict[str, Tuple[int, ...]] = {"pos_embed_z": (1, 64, C), "pos_embed_x": (1, 256, C)}
    chans = [3, C // 8, C // 4, C // 2, C]
    for i in range(4):
        p = f"patch_embed.net.{2 * i}"
        s[f"{p}.c.weight"] = (chans[i + 1], chans[i], 3, 3)
        for n in ("weight", "bias", "running_mean", "running_var"):
            s[f"{p}.bn.{n}"] = (chans[i + 1],)
        s[f"{p}.bn.num_batches_tracked"] = ()
    for b in range(depth):
        p = f"blocks.{b}"
        s[f"{p}.norm1.weight"] = (C,); s[f"{p}.norm1.bias"] = (C,)
        s[f"{p}.attn.qkv.weight"] = (3 * C, C); s[f"{p}.attn.qkv.bias"] = (3 * C,)
        s[f"{p}.attn.proj.weight"] = (C, C); s[f"{p}.attn.proj.bias"] = (C,)
        s[f"{p}.norm2.weight"] = (C,); s[f"{p}.norm2.bias"] = (C,)
        s[f"{p}.mlp.fc1.weight"] = (mlp_ratio * C, C); s[f"{p}.mlp.fc1.bias"] = (mlp_ratio * C,)
        s[f"{p}.mlp.fc2.weight"] = (C, mlp_ratio * C); s[f"{p}.mlp.fc2.bias"] = (C,)
    s["norm.weight"] = (C,); s["norm.bias"] = (C,)
    hc = [C, head_ch, head_ch // 2, head_ch // 4, head_ch // 8]
    outs = {"ctr": 1, "offset": 2, "size": 2}
    for t in TOWERS:
        for i in range(4):
            p = f"box_head.conv{i + 1}_{t}"
            s[f"{p}.0.weight"] = (hc[i + 1], hc[i], 3, 3); s[f"{p}.0.bias"] = (hc[i + 1],)
            for n in ("weight", "bias", "running_mean", "running_var"):
                s[f"{p}.1.{n}"] = (hc[i + 1],)
            s[f"{p}.1.num_batches_tracked"] = ()
        s[f"box_head.conv5_{t}.weight"] = (outs[t], hc[4], 1, 1)
        s[f"box_head.conv5_{t}.bias"] = (outs[t],)
    return s


def make_state_dict(seed: int = 0, stress: bool = False, stable_size: bool = False,
                    **shape_kw) -> Dict[str, torch.Tensor]:
    """Synthetic 'random-init' weights with the reference's distributions:
    PyTorch defaults for conv/linear (kaiming-uniform a=sqrt(5) + fan-in bias), xavier-uniform for
    every >1-D head parameter (head.py:126-128), LayerNorm/BN identity, pos-embeds zero
    (vit_dist.py:61-62).  ``stress=True`` randomises pos-embeds, BN statistics/affine and LN affine so
    that BN folding and the pos-embed add are not no-ops.  ``stable_size`` biases conv5_size to
    sigmoid^-1(0.25) so closed-loop boxes stay scale-stable."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    for name, shape in param_shapes(**shape_kw).items():
        leaf = name.rsplit(".", 1)[-1]
        is_bn = ".bn." in name or (name.startswith("box_head.conv") and ".1." in name)
        is_ln = "norm" in name
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif name.startswith("pos_embed"):
            t = torch.randn(shape, generator=g) * 0.02 if stress else torch.zeros(shape)
        elif is_bn:
            if not stress:
                t = torch.ones(shape) if leaf in ("weight", "running_var") else torch.zeros(shape)
            elif leaf == "running_mean":
                t = torch.randn(shape, generator=g) * 0.1
            elif leaf == "running_var":
                t = torch.rand(shape, generator=g) + 0.5
            elif leaf == "weight":
                t = torch.rand(shape, generator=g) * 0.4 + 0.8
            else:
                t = torch.randn(shape, generator=g) * 0.1
        elif is_ln:
            if not stress:
                t = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
            elif leaf == "weight":
                t = torch.rand(shape, generator=g) * 0.4 + 0.8
            else:
                t = torch.randn(shape, generator=g) * 0.1
        elif len(shape) > 1:
            fan_in = int(np.prod(shape[1:]))
            fan_out = shape[0] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[0]
            if name.startswith("box_head"):
                t = uni(shape, math.sqrt(6.0 / (fan_in + fan_out)))      # xavier_uniform_
            else:
                t = uni(shape, 1.0 / math.sqrt(fan_in))                  # kaiming_uniform_(a=sqrt(5))
        else:  # bias of conv / linear
            wshape = param_shapes(**shape_kw)[name[:-4] + "weight"]
            t = uni(shape, 1.0 / math.sqrt(int(np.prod(wshape[1:]))))
        sd[name] = t.contiguous()
    if stable_size:
        sd["box_head.conv5_size.bias"] = torch.full((2,), math.log(1.0 / 3.0))
    return sd


# ----------------------------------------------------------------------------------------------
# R7-R12  the model
# ----------------------------------------------------------------------------------------------
class OracleModel:
    """fp32 eval-mode restatement of OstrackDist (vit_dist.py:57-155) on the stock torch CPU ops the
    reference itself dispatches to (conv2d, batch_norm, layer_norm, linear, softmax, gelu)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], depth: int = DEPTH, num_heads: int = NUM_HEADS):
        self.sd = {k: v.detach().clone().float() if v.is_floating_point() else v.clone()
                   for k, v in state_dict.items()}
        self.depth = depth
        self.num_heads = num_heads
        self.C = self.sd["pos_embed_x"].shape[-1]
        self.feat_sz = FEAT_SZ

    def to(self, device) -> "OracleModel":
        """Move the parameters (bench.py's `gpu_eager` baseline leg: the same stock torch ops, dispatched to the GPU)."""
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        return self

    # -- R7: LevitPatchEmbedding (vit_dist.py:36-54) ------------------------------------------
    def patch_embed(self, img: torch.Tensor, taps: Optional[dict] = None, tag: str = "") -> torch.Tensor:
        x = img
        for i in range(4):
            p = f"patch_embed.net.{2 * i}"
            x = F.conv2d(x, self.sd[f"{p}.c.weight"], None, stride=2, padding=1)
            x = F.batch_norm(x, self.sd[f"{p}.bn.running_mean"], self.sd[f"{p}.bn.running_var"],
                             self.sd[f"{p}.bn.weight"], self.sd[f"{p}.bn.bias"], False, 0.0, BN_EPS)
            if i < 3:
                x = F.hardswish(x)
            if taps is not None:
                taps[f"stem{i + 1}{tag}"] = x
        return x.flatten(2).transpose(1, 2)

    # -- R9: timm Block (onnxexport.py:153-225; vit.py:39-91) ---------------------------------
    def block(self, x: torch.Tensor, b: int) -> torch.Tensor:
        p = f"blocks.{b}"
        B, N, C = x.shape
        hd = C // self.num_heads
        h = F.layer_norm(x, (C,), self.sd[f"{p}.norm1.weight"], self.sd[f"{p}.norm1.bias"], LN_EPS)
        qkv = F.linear(h, self.sd[f"{p}.attn.qkv.weight"], self.sd[f"{p}.attn.qkv.bias"])
        qkv = qkv.reshape(B, N, 3, self.num_heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q * (hd ** -0.5)) @ k.transpose(-2, -1)
        attn = attn.softmax(dim=-1)
        a = (attn @ v).transpose(1, 2).reshape(B, N, C)
        x = x + F.linear(a, self.sd[f"{p}.attn.proj.weight"], self.sd[f"{p}.attn.proj.bias"])
        h = F.layer_norm(x, (C,), self.sd[f"{p}.norm2.weight"], self.sd[f"{p}.norm2.bias"], LN_EPS)
        h = F.gelu(F.linear(h, self.sd[f"{p}.mlp.fc1.weight"], self.sd[f"{p}.mlp.fc1.bias"]))
        return x + F.linear(h, self.sd[f"{p}.mlp.fc2.weight"], self.sd[f"{p}.mlp.fc2.bias"])

    # -- R11: CenterPredictor.get_score_map (head.py:175-201) ---------------------------------
    def head(self, feat: torch.Tensor):
        outs = {}
        for t in TOWERS:
            x = feat
            for i in range(4):
                p = f"box_head.conv{i + 1}_{t}"
                x = F.conv2d(x, self.sd[f"{p}.0.weight"], self.sd[f"{p}.0.bias"], stride=1, padding=1)
                x = F.batch_norm(x, self.sd[f"{p}.1.running_mean"], self.sd[f"{p}.1.running_var"],
                                 self.sd[f"{p}.1.weight"], self.sd[f"{p}.1.bias"], False, 0.0, BN_EPS)
                x = F.relu(x)
            outs[t] = F.conv2d(x, self.sd[f"box_head.conv5_{t}.weight"], self.sd[f"box_head.conv5_{t}.bias"])
        sig = lambda v: torch.clamp(torch.sigmoid(v), min=1e-4, max=1 - 1e-4)
        return sig(outs["ctr"]), sig(outs["size"]), outs["offset"]

    # -- R12: cal_bbox (head.py:142-160) ------------------------------------------------------
    def cal_bbox(self, score: torch.Tensor, size_map: torch.Tensor, offset_map: torch.Tensor) -> torch.Tensor:
        _, idx = torch.max(score.flatten(1), dim=1, keepdim=True)
        idx_y = idx // self.feat_sz
        idx_x = idx % self.feat_sz
        idx2 = idx.unsqueeze(1).expand(idx.shape[0], 2, 1)
        size = size_map.flatten(2).gather(dim=2, index=idx2)
        offset = offset_map.flatten(2).gather(dim=2, index=idx2).squeeze(-1)
        return torch.cat([(idx_x.to(torch.float) + offset[:, :1]) / self.feat_sz,
                          (idx_y.to(torch.float) + offset[:, 1:]) / self.feat_sz,
                          size.squeeze(-1)], dim=1)

    # -- R8 + R10: forward (vit_dist.py:77-100,122-153) ---------------------------------------
    @torch.no_grad()
    def forward(self, z: torch.Tensor, x: torch.Tensor, taps: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        zt = self.patch_embed(z, taps, "_z") + self.sd["pos_embed_z"]
        xt = self.patch_embed(x, taps, "_x") + self.sd["pos_embed_x"]
        t = torch.cat((zt, xt), dim=1)
        if taps is not None:
            taps["tokens0"] = t
        for b in range(self.depth):
            t = self.block(t, b)
            if taps is not None:
                taps[f"tokens{b + 1}"] = t
        t = F.layer_norm(t, (self.C,), self.sd["norm.weight"], self.sd["norm.bias"], LN_EPS)
        if taps is not None:
            taps["tokens_norm"] = t
        B = t.shape[0]
        feat = t[:, -self.feat_sz ** 2:].unsqueeze(-1).permute(0, 3, 2, 1).contiguous()
        feat = feat.view(-1, self.C, self.feat_sz, self.feat_sz)
        score, size_map, offset_map = self.head(feat)
        bbox = self.cal_bbox(score, size_map, offset_map)
        return {"pred_boxes": bbox.view(B, 1, 4), "score_map": score, "size_map": size_map,
                "offset_map": offset_map}

    __call__ = forward


# ----------------------------------------------------------------------------------------------
# R13  post-processing and the tracker state machine
# ----------------------------------------------------------------------------------------------
def clip_box(box: list, H, W, margin=0) -> list:
    """lib/utils/box_ops.py:97-106 (Python scalars; ints survive where a clamp constant wins)."""
    x1, y1, w, h = box
    x2, y2 = x1 + w, y1 + h
    x1 = min(max(0, x1), W - margin)
    x2 = min(max(margin, x2), W)
    y1 = min(max(0, y1), H - margin)
    y2 = min(max(margin, y2), H)
    w = max(margin, x2 - x1)
    h = max(margin, y2 - y1)
    return [x1, y1, w, h]


def map_box_back(state: list, pred_box: list, resize_factor: float, search_size: int = SEARCH_SIZE) -> list:
    """lib/test/tracker/vit_dist.py:150-156."""
    cx_prev, cy_prev = state[0] + 0.5 * state[2], state[1] + 0.5 * state[3]
    cx, cy, w, h = pred_box
    half_side = 0.5 * search_size / resize_factor
    cx_real = cx + (cx_prev - half_side)
    cy_real = cy + (cy_prev - half_side)
    return [cx_real - 0.5 * w, cy_real - 0.5 * h, w, h]


class OracleTracker:
    """State machine of lib/test/tracker/vit_dist.py:53-148 (debug/visdom/save_all_boxes branches
    omitted: params.debug = 0 and save_all_boxes = False on the evaluated path)."""

    def __init__(self, model: OracleModel, use_cv: bool = False):
        self.model = model
        self.window = hann2d(FEAT_SZ, FEAT_SZ)
        self.crop = sample_target_cv if use_cv else sample_target_spec
        self.state = None
        self.frame_id = 0
        self.z = None
        self.last = {}

    def initialize(self, image: np.ndarray, info: dict):
        z_patch, _, _ = self.crop(image, info['init_bbox'], TEMPLATE_FACTOR, TEMPLATE_SIZE)
        self.z_patch_arr = z_patch
        self.z = preprocess(z_patch)
        self.state = info['init_bbox']
        self.frame_id = 0

    @torch.no_grad()
    def track(self, image: np.ndarray, info: dict = None):
        H, W, _ = image.shape
        self.frame_id += 1
        x_patch, resize_factor, _ = self.crop(image, self.state, SEARCH_FACTOR, SEARCH_SIZE)
        x = preprocess(x_patch)
        out = self.model.forward(self.z, x)
        response = self.window * out['score_map']
        pred_boxes = self.model.cal_bbox(response, out['size_map'], out['offset_map']).view(-1, 4)
        pred_box = (pred_boxes.mean(dim=0) * SEARCH_SIZE / resize_factor).tolist()
        self.state = clip_box(map_box_back(self.state, pred_box, resize_factor), H, W, margin=10)
        self.last = {"x_patch": x_patch, "resize_factor": resize_factor, "out": out, "response": response,
                     "pred_boxes": pred_boxes}
        return {"target_bbox": self.state, "confidence": out['score_map'].max()}


# ----------------------------------------------------------------------------------------------
# Synthetic workload generators shared by tests and bench
# ----------------------------------------------------------------------------------------------
def synth_frames(n: int, H: int = 720, W: int = 1280, seed: int = 0, smooth: bool = False) -> np.ndarray:
    """uint8 frames (n,H,W,3).  ``smooth`` low-pass filters the noise so bilinear taps differ less
    trivially (both kinds are used by the tests)."""
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)
    small = rng.integers(0, 256, size=(n, H // 8 + 2, W // 8 + 2, 3)).astype(np.float32)
    up = np.repeat(np.repeat(small, 8, axis=1), 8, axis=2)[:, :H, :W]
    up += rng.normal(0, 12, size=up.shape).astype(np.float32)
    return np.clip(up, 0, 255).astype(np.uint8)


def synth_boxes(n: int, H: int = 720, W: int = 1280, seed: int = 0) -> np.ndarray:
    """Open-loop boxes (n,4) float64 [x,y,w,h]: w,h ~ U(16,400); 80% strictly inside with >=10 px
    margin, 10% touching a border, 10% small (crop_sz < S, the up-scaling path)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 4), dtype=np.float64)
    for i in range(n):
        kind = rng.random()
        if kind < 0.1:
            w, h = rng.uniform(4, 40), rng.uniform(4, 40)
        else:
            w, h = rng.uniform(16, 400), rng.uniform(16, 400)
        w, h = min(w, W - 21), min(h, H - 21)
        x = rng.uniform(10, W - 10 - w)
        y = rng.uniform(10, H - 10 - h)
        if 0.1 <= kind < 0.2:
            side = rng.integers(0, 4)
            if side == 0: x = 0.0
            elif side == 1: y = 0.0
            elif side == 2: x = W - w
            else: y = H - h
        if rng.random() < 0.15:       # integer boxes hit the .5 rounding case of round()
            x, y, w, h = float(int(x)), float(int(y)), float(max(4, int(w))), float(max(4, int(h)))
        out[i] = (x, y, w, h)
    return out
