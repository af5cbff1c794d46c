"""Tracker parameters: mirrors ``lib/test/utils/params.py`` (TrackerParams attribute bag) and
``lib/test/parameter/vit_dist.py:7-30`` (``parameters(yaml_name)``)."""
from __future__ import annotations

import os

from . import config as _config


class TrackerParams:
    """Attribute bag with the helper methods the reference's harness relies on (params.py:5-25)."""

    def set_default_values(self, default_vals: dict):
        for name, val in default_vals.items():
            if not hasattr(self, name):
                setattr(self, name, val)

    def get(self, name: str, *default):
        if len(default) > 1:
            raise ValueError('Can only give one default value.')
        if not default:
            return getattr(self, name)
        return getattr(self, name, default[0])

    def has(self, name: str):
        return hasattr(self, name)


def parameters(yaml_name: str, prj_dir: str = None, save_dir: str = None) -> TrackerParams:
    """Build the parameter bag for experiment ``yaml_name``.

    ``prj_dir`` / ``save_dir`` stand in for the reference's ``env_settings()`` paths
    (lib/test/evaluation/environment.py); by default the packaged experiments directory and
    ``$VT_SAVE_DIR`` (or ./output) are used.  The checkpoint path keeps the reference's naming
    (``checkpoints/train/vit_dist/<yaml>/OstrackDist_ep%04d.pth.tar``)."""
    params = TrackerParams()
    prj_dir = prj_dir or _config.PKG_DIR
    save_dir = save_dir or os.environ.get("VT_SAVE_DIR", os.path.join(os.getcwd(), "output"))
    yaml_file = os.path.join(prj_dir, 'experiments/vit_dist/%s.yaml' % yaml_name)
    cfg = _config.load_cfg(yaml_file)
    params.cfg = cfg
    params.template_factor = cfg.TEST.TEMPLATE_FACTOR
    params.template_size = cfg.TEST.TEMPLATE_SIZE
    params.search_factor = cfg.TEST.SEARCH_FACTOR
    params.search_size = cfg.TEST.SEARCH_SIZE
    params.checkpoint = os.path.join(save_dir, "checkpoints/train/vit_dist/%s/OstrackDist_ep%04d.pth.tar" %
                                     (yaml_name, cfg.TEST.EPOCH))
    params.save_all_boxes = False
    params.debug = 0
    return params
