"""State-dict plumbing: key names/shapes of ``build_ostrack_dist(cfg)`` in eval mode
(lib/models/vit_dist/vit_dist.py:57-75, lib/models/layers/head.py:98-128), a random initialiser with
the reference's distributions, and checkpoint loading in the reference's format
(``{'net': state_dict, ...}``, lib/train/trainers/base_trainer.py:116-148)."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

TOWERS = ("ctr", "offset", "size")


def param_shapes(cfg, depth: int = 3, mlp_ratio: int = 4) -> Dict[str, Tuple[int, ...]]:
    C = int(cfg.MODEL.BACKBONE.CHANNELS)
    hc = int(cfg.MODEL.HEAD.NUM_CHANNELS)
    s: Dict[str, Tuple[int, ...]] = {"pos_embed_z": (1, 64, C), "pos_embed_x": (1, 256, C)}
    ch = [3, C // 8, C // 4, C // 2, C]
    for i in range(4):
        p = f"patch_embed.net.{2 * i}"
        s[f"{p}.c.weight"] = (ch[i + 1], ch[i], 3, 3)
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            s[f"{p}.bn.{leaf}"] = (ch[i + 1],)
        s[f"{p}.bn.num_batches_tracked"] = ()
    for b in range(depth):
        p = f"blocks.{b}"
        for ln in ("norm1", "norm2"):
            s[f"{p}.{ln}.weight"] = (C,)
            s[f"{p}.{ln}.bias"] = (C,)
        s[f"{p}.attn.qkv.weight"], s[f"{p}.attn.qkv.bias"] = (3 * C, C), (3 * C,)
        s[f"{p}.attn.proj.weight"], s[f"{p}.attn.proj.bias"] = (C, C), (C,)
        s[f"{p}.mlp.fc1.weight"], s[f"{p}.mlp.fc1.bias"] = (mlp_ratio * C, C), (mlp_ratio * C,)
        s[f"{p}.mlp.fc2.weight"], s[f"{p}.mlp.fc2.bias"] = (C, mlp_ratio * C), (C,)
    s["norm.weight"], s["norm.bias"] = (C,), (C,)
    hch = [C, hc, hc // 2, hc // 4, hc // 8]
    for t, nout in zip(TOWERS, (1, 2, 2)):
        for i in range(4):
            p = f"box_head.conv{i + 1}_{t}"
            s[f"{p}.0.weight"], s[f"{p}.0.bias"] = (hch[i + 1], hch[i], 3, 3), (hch[i + 1],)
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                s[f"{p}.1.{leaf}"] = (hch[i + 1],)
            s[f"{p}.1.num_batches_tracked"] = ()
        s[f"box_head.conv5_{t}.weight"], s[f"box_head.conv5_{t}.bias"] = (nout, hch[4], 1, 1), (nout,)
    return s


def random_init_state_dict(cfg, depth: int = 3, generator: torch.Generator = None) -> Dict[str, torch.Tensor]:
    """What a freshly constructed reference model holds: PyTorch default init for conv / linear
    (uniform +-1/sqrt(fan_in)), identity BatchNorm / LayerNorm, zero positional embeddings
    (vit_dist.py:61-62) and xavier-uniform on every >1-D head parameter (head.py:126-128)."""
    shapes = param_shapes(cfg, depth)
    sd: Dict[str, torch.Tensor] = {}

    def uniform(shape, bound):
        return (torch.rand(shape, generator=generator) * 2 - 1) * bound

    for name, shape in shapes.items():
        leaf = name.rsplit(".", 1)[-1]
        is_bn = ".bn." in name or (name.startswith("box_head.conv") and ".1." in name)
        if leaf == "num_batches_tracked":
            sd[name] = torch.zeros((), dtype=torch.long)
        elif name.startswith("pos_embed"):
            sd[name] = torch.zeros(shape)
        elif is_bn or "norm" in name:
            sd[name] = torch.ones(shape) if leaf in ("weight", "running_var") else torch.zeros(shape)
        elif len(shape) > 1:
            fan_in = math.prod(shape[1:])
            fan_out = shape[0] * math.prod(shape[2:])
            bound = math.sqrt(6.0 / (fan_in + fan_out)) if name.startswith("box_head") else 1.0 / math.sqrt(fan_in)
            sd[name] = uniform(shape, bound)
        else:
            sd[name] = uniform(shape, 1.0 / math.sqrt(math.prod(shapes[name[:-4] + "weight"][1:])))
    return sd


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """``torch.load(path, map_location='cpu')['net']`` (lib/test/tracker/vit_dist.py:25)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    return ckpt["net"] if isinstance(ckpt, dict) and "net" in ckpt else ckpt
