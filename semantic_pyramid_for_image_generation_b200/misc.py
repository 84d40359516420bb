"""Semantic-level masks and the metric logger, API-compatible with the reference's misc.py (misc.py:13-159).

Design: a training mask set is fully described by a tiny *descriptor* -- the selected pyramid stage plus, for the
spatially varying case, one low-resolution keep/drop bitmap -- and expanded to the seven per-level tensors by pure
integer indexing.  `draw_mask_descriptor` consumes the Python / NumPy random streams in the reference's order
(misc.py:28,32,37), `expand_mask_descriptor` reproduces its outputs bit for bit (ones / zeros / nearest-neighbour
resized bitmap, deepest-first list reversed to VGG order).  The descriptor form is what a data loader ships to the
GPU (a few bytes per sample instead of ~90 KB of float masks).

`skimage.draw.random_shapes` (misc.py:8) is used when scikit-image is installed; otherwise a seeded rectangle/disc
rasteriser with the same output contract (uint8 image, 255 = background) stands in.
"""
import json
import os
import random
from collections import namedtuple
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

PYRAMID_SHAPES = ((1, 128, 128), (1, 64, 64), (1, 32, 32), (1, 16, 16), (1, 8, 8), (4096,), (365,))

# stage counts from the deepest level (0 = logits, 1 = fc7, 2 = pool5, ... 6 = pool1); bitmap is uint8 {0,1} at the
# resolution of level stage+1 or None
MaskDescriptor = namedtuple("MaskDescriptor", ["stage", "bitmap"])


def _builtin_random_shapes(image_shape, min_shapes=1, max_shapes=4, min_size=2, allow_overlap=True):
    h, w = image_shape
    canvas = np.full((h, w, 3), 255, dtype=np.uint8)
    count = np.random.randint(min_shapes, max_shapes + 1)
    found = []
    for _ in range(count):
        extent = np.random.randint(min_size, max(min_size + 1, min(h, w)))
        top = np.random.randint(0, max(1, h - extent + 1))
        left = np.random.randint(0, max(1, w - extent + 1))
        colour = np.random.randint(0, 255, size=3)
        if np.random.rand() < 0.5:
            canvas[top:top + extent, left:left + extent] = colour
            found.append(("rectangle", ((top, top + extent), (left, left + extent))))
        else:
            rows, cols = np.ogrid[:h, :w]
            radius = extent / 2.0
            canvas[(rows - top - radius) ** 2 + (cols - left - radius) ** 2 <= radius ** 2] = colour
            found.append(("circle", ((top, top + extent), (left, left + extent))))
    return canvas, found


try:  # pragma: no cover - environment dependent
    from skimage.draw import random_shapes as _shape_rasteriser  # type: ignore
except Exception:
    _shape_rasteriser = _builtin_random_shapes


def draw_mask_descriptor(mask_shapes: Sequence[Tuple] = PYRAMID_SHAPES, p_random_mask: float = 0.3,
                         rasteriser=None) -> MaskDescriptor:
    """Samples what misc.py:28-45 samples, in the same order: stage (two extra chances for the two vector levels),
    then the spatial-mask coin, then -- only for 0 < stage < 6 -- the shape image at the next shallower level."""
    deepest_first = tuple(reversed(tuple(mask_shapes)))
    levels = len(deepest_first)
    stage = random.choice(list(range(levels)) + [0, 1])
    wants_spatial = np.random.rand() < p_random_mask
    if not (wants_spatial and 0 < stage < levels - 1):
        return MaskDescriptor(stage, None)
    canvas_hw = deepest_first[stage + 1][1:]
    draw = rasteriser or _shape_rasteriser
    image = draw(tuple(canvas_hw), min_shapes=1, max_shapes=4, min_size=min(8, deepest_first[stage + 1][1] // 2),
                 allow_overlap=True)[0]
    keep = (np.asarray(image)[:, :, 0] == 255).astype(np.uint8)  # red channel; background (255) is kept
    return MaskDescriptor(stage, keep)


def _nearest_rows(src: int, dst: int) -> torch.Tensor:
    return (torch.arange(dst) * src // dst).clamp_(max=src - 1)


def expand_mask_descriptor(desc: MaskDescriptor, mask_shapes: Sequence[Tuple] = PYRAMID_SHAPES, device='cpu',
                           add_batch_size: bool = False) -> List[torch.Tensor]:
    """Seven float masks in VGG order (shallowest first), values exactly 0.0 / 1.0."""
    deepest_first = tuple(reversed(tuple(mask_shapes)))
    bitmap = None if desc.bitmap is None else torch.as_tensor(np.asarray(desc.bitmap), dtype=torch.float32)
    out = []
    for depth, shape in enumerate(deepest_first):
        if depth == desc.stage:
            level = torch.ones(shape, dtype=torch.float32)
        elif bitmap is not None and depth > desc.stage:
            rows = _nearest_rows(bitmap.shape[0], shape[1])
            cols = _nearest_rows(bitmap.shape[1], shape[2])
            level = bitmap[rows[:, None], cols[None, :]].unsqueeze(0).contiguous()
        else:
            level = torch.zeros(shape, dtype=torch.float32)
        level = level.to(device)
        out.append(level.unsqueeze(0) if add_batch_size else level)
    out.reverse()
    return out


def get_masks_for_training(mask_shapes: Sequence[Tuple] = PYRAMID_SHAPES, device: str = 'cpu',
                           add_batch_size: bool = False, p_random_mask: float = 0.3) -> List[torch.Tensor]:
    '''Random masks of section 3.2 of the paper; drop-in for reference misc.py:13-68.'''
    return expand_mask_descriptor(draw_mask_descriptor(mask_shapes, p_random_mask), mask_shapes, device, add_batch_size)


def get_masks_for_inference(stage_index_to_choose: int, mask_shapes: Sequence[Tuple] = PYRAMID_SHAPES,
                            device: str = 'cpu', add_batch_size: bool = False) -> List[torch.Tensor]:
    '''Only level `stage_index_to_choose` (counted from the deepest) is kept; reference misc.py:78-97.'''
    return expand_mask_descriptor(MaskDescriptor(stage_index_to_choose, None), mask_shapes, device, add_batch_size)


def get_masks_for_validation(mask_shapes: Sequence[Tuple] = PYRAMID_SHAPES, device: str = 'cpu',
                             add_batch_size: bool = False) -> List[torch.Tensor]:
    '''Uniformly random single level; reference misc.py:71-75.'''
    return get_masks_for_inference(random.choice(range(len(mask_shapes))), mask_shapes, device, add_batch_size)


def _minmax(batch: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    flat = batch.reshape(batch.shape[0], -1)
    view = (-1,) + (1,) * (batch.dim() - 1)
    return flat.min(dim=1)[0].view(view), flat.max(dim=1)[0].view(view)


def normalize_0_1_batch(input: torch.Tensor) -> torch.Tensor:
    '''Per-sample min-max to [0, 1] (reference misc.py:100-109); plotting helper, host side.'''
    lo, hi = _minmax(input)
    return (input - lo) / (hi - lo)


def normalize_m1_1_batch(input: torch.Tensor) -> torch.Tensor:
    '''Per-sample min-max to [-1, 1] (reference misc.py:112-121).'''
    return 2 * normalize_0_1_batch(input) - 1


class Logger(object):
    """Metric store of reference misc.py:124-159: `log(name, value)` appends, `save_metrics(path)` writes
    hyperparameter.txt (JSON) and one `<metric>.pt` tensor per metric."""

    def __init__(self) -> None:
        self.metrics = dict()
        self.hyperparameter = dict()

    def log(self, metric_name: str, value: float) -> None:
        self.metrics.setdefault(metric_name, []).append(value)

    def save_metrics(self, path: str) -> None:
        with open(os.path.join(path, 'hyperparameter.txt'), 'w') as handle:
            json.dump(self.hyperparameter, handle)
        for name, series in self.metrics.items():
            torch.save(torch.tensor(series), os.path.join(path, name + '.pt'))
