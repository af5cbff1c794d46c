"""Configuration tree for the vit_dist tracker family.

Mirrors the interface of the reference's ``lib/config/vit_dist/config.py`` (``cfg`` attribute tree,
``update_config_from_file(yaml)`` that raises ``ValueError`` on a key the defaults do not know,
:128-149) for the keys the inference path reads; training-only keys are accepted and carried but
never interpreted.
"""
from __future__ import annotations

import copy
import os

import yaml

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
DEFAULT_YAML = os.path.join(PKG_DIR, "experiments", "vit_dist", "vit_48_h32_noKD.yaml")


class Node(dict):
    """dict with attribute access (what easydict gives the reference)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, Node(v) if isinstance(v, dict) and not isinstance(v, Node) else v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__

    def __deepcopy__(self, memo):
        return Node({k: copy.deepcopy(v, memo) for k, v in self.items()})


def default_cfg() -> Node:
    """Family defaults (lib/config/vit_dist/config.py:7-106)."""
    return Node({
        "MODEL": {
            "PRETRAIN_FILE": "mae_pretrain_vit_base.pth", "EXTRA_MERGER": False, "RETURN_INTER": False,
            "RETURN_STAGES": [],
            "BACKBONE": {"TYPE": "vit_base_patch16_224", "STRIDE": 16, "MID_PE": False, "SEP_SEG": False,
                         "CAT_MODE": "direct", "MERGE_LAYER": 0, "ADD_CLS_TOKEN": False,
                         "CLS_TOKEN_USE_MODE": "ignore", "CHANNELS": 768, "HEADS": 12, "CE_LOC": [],
                         "CE_KEEP_RATIO": [], "CE_TEMPLATE_RANGE": "ALL",
                         "DEPTH": 3},      # DEPTH: this package only (build_ostrack_dist(cfg, depth=3) in the reference)
            "HEAD": {"TYPE": "CENTER", "NUM_CHANNELS": 256},
        },
        "TRAIN": {"LR": 0.0001, "WEIGHT_DECAY": 0.0001, "EPOCH": 500, "LR_DROP_EPOCH": 400, "BATCH_SIZE": 16,
                  "NUM_WORKER": 8, "OPTIMIZER": "ADAMW", "BACKBONE_MULTIPLIER": 0.1, "GIOU_WEIGHT": 2.0,
                  "L1_WEIGHT": 5.0, "AUX_WEIGHT": 1.0, "AUX_TYPE": "3 output", "FREEZE_LAYERS": [0],
                  "PRINT_INTERVAL": 50, "VAL_EPOCH_INTERVAL": 20, "GRAD_CLIP_NORM": 0.1, "AMP": False,
                  "TEACHER": "ostrack", "CE_START_EPOCH": 20, "CE_WARM_EPOCH": 80, "DROP_PATH_RATE": 0.1,
                  "SCHEDULER": {"TYPE": "step", "DECAY_RATE": 0.1}},
        "DATA": {"SAMPLER_MODE": "causal", "MEAN": [0.485, 0.456, 0.406], "STD": [0.229, 0.224, 0.225],
                 "MAX_SAMPLE_INTERVAL": 200,
                 "TRAIN": {"DATASETS_NAME": ["LASOT", "GOT10K_vottrain"], "DATASETS_RATIO": [1, 1],
                           "SAMPLE_PER_EPOCH": 60000},
                 "VAL": {"DATASETS_NAME": ["GOT10K_votval"], "DATASETS_RATIO": [1], "SAMPLE_PER_EPOCH": 10000},
                 "SEARCH": {"SIZE": 320, "FACTOR": 5.0, "CENTER_JITTER": 4.5, "SCALE_JITTER": 0.5, "NUMBER": 1},
                 "TEMPLATE": {"NUMBER": 1, "SIZE": 128, "FACTOR": 2.0, "CENTER_JITTER": 0, "SCALE_JITTER": 0}},
        "TEST": {"TEMPLATE_FACTOR": 2.0, "TEMPLATE_SIZE": 128, "SEARCH_FACTOR": 5.0, "SEARCH_SIZE": 320,
                 "EPOCH": 500},
    })


cfg = default_cfg()


def _update_config(base: Node, exp: dict) -> None:
    for k, v in exp.items():
        if k not in base:
            raise ValueError("{} not exist in config.py".format(k))      # config.py:137
        if isinstance(v, dict):
            _update_config(base[k], v)
        else:
            base[k] = v


def update_config_from_file(filename: str, base_cfg: Node = None) -> None:
    with open(filename) as f:
        exp = yaml.safe_load(f) or {}
    _update_config(cfg if base_cfg is None else base_cfg, exp)


def load_cfg(yaml_file: str = DEFAULT_YAML) -> Node:
    """Fresh defaults overlaid with ``yaml_file`` (does not touch the module-level ``cfg``)."""
    c = default_cfg()
    update_config_from_file(yaml_file, c)
    return c
