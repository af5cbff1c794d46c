"""Pre-processing drop-ins: ``sample_target`` (lib/train/data/processing_utils.py:12-79, mask=None,
output_sz given) and ``Preprocessor`` (lib/test/tracker/data_utils.py:6-17), both served by the
single fused crop/resize/normalise kernel (vt_crop_normalize)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .engine import Engine


class NestedTensor(object):
    """lib/utils/misc.py:284-305."""

    def __init__(self, tensors, mask: Optional[torch.Tensor]):
        self.tensors = tensors
        self.mask = mask

    def to(self, device):
        return NestedTensor(self.tensors.to(device), self.mask.to(device) if self.mask is not None else None)

    def decompose(self):
        return self.tensors, self.mask

    def __repr__(self):
        return str(self.tensors)


class CropPreprocessor:
    """``sample_target`` and ``Preprocessor.process`` in one call on the GPU."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def crop(self, im: np.ndarray, target_bb, search_area_factor: float, output_sz: int):
        """Returns (patch uint8 [S,S,3] numpy, resize_factor float, att_mask bool [S,S] numpy, NestedTensor)
        - the first three are what ``sample_target`` returns, the last what ``Preprocessor.process``
        would make of them (normalised fp32 (1,3,S,S) + bool mask (1,S,S), on the device)."""
        if not isinstance(target_bb, list):
            target_bb = target_bb.tolist()
        dev = self.engine.device
        im = np.ascontiguousarray(im)
        H, W, _ = im.shape
        frame = torch.from_numpy(im).to(dev).reshape(-1)
        out = self.engine.crop_normalize(frame, torch.zeros(1, dtype=torch.int64, device=dev),
                                         torch.tensor([[H, W]], dtype=torch.int32, device=dev),
                                         torch.tensor([[float(v) for v in target_bb]], dtype=torch.float64, device=dev),
                                         float(search_area_factor), int(output_sz), want_u8=True, want_mask=True)
        st = int(out["status"].item())
        if st == 1:
            raise Exception('Too small bounding box.')
        if st != 0:
            raise ValueError("crop lies outside the image (undefined behaviour in the reference)")
        mask = out["mask"][0].to(torch.bool)
        return (out["u8"][0].cpu().numpy(), float(out["resize_factor"].item()), mask.cpu().numpy(),
                NestedTensor(out["tensors"], mask.unsqueeze(0)))
