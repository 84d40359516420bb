"""Device-side input pipeline: the work of the reference's DataLoader workers after JPEG decoding.

The reference builds, per sample and on the host, a float image normalised to [-1, 1] (data.py:49-53) and seven float
mask tensors (misc.py:13-68, ~90 KB), collates them (data.py:76-90) and copies the batch to the GPU synchronously from
pageable memory (model_wrapper.py:139-142).  Here a batch crosses PCIe as

    uint8 images (B, 3, H, W)  +  int64 class indices (B,)  +  mask descriptors (stage, optional low-res bitmap)

from pinned staging buffers on a copy stream, and two kernels of the C-ABI library rebuild exactly what the reference
would have produced: `spyr_image_u8_minmax_normalize` and `spyr_expand_mask_level` (bit-exact masks, values 0.0 / 1.0).
The random draws stay on the host in the reference's order (`misc.draw_mask_descriptor`).

`DeviceBatchLoader` wraps any iterable of host batches and yields device-resident `(images, labels, masks)` triples
in the collate function's format, one batch ahead of the consumer.
"""
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import misc
from ._native import call

MAX_BITMAP = 128  # the shallowest level a spatial mask is drawn at (misc.py:37: level stage + 1 <= pool1 = 128x128)


class PackedDescriptors(object):
    """Mask descriptors of a batch as three flat host tensors (pinned when CUDA is available)."""

    def __init__(self, batch: int, pin: bool = True) -> None:
        pin = pin and torch.cuda.is_available()
        self.stage = torch.zeros(batch, dtype=torch.int32, pin_memory=pin)
        self.bitmap_hw = torch.zeros(batch, dtype=torch.int32, pin_memory=pin)
        self.bitmaps = torch.zeros((batch, MAX_BITMAP * MAX_BITMAP), dtype=torch.uint8, pin_memory=pin)

    def fill(self, descriptors: Sequence[misc.MaskDescriptor]) -> "PackedDescriptors":
        if len(descriptors) != self.stage.shape[0]:
            raise ValueError("expected %d descriptors, got %d" % (self.stage.shape[0], len(descriptors)))
        for b, d in enumerate(descriptors):
            self.stage[b] = int(d.stage)
            if d.bitmap is None:
                self.bitmap_hw[b] = 0
                continue
            bm = np.asarray(d.bitmap)
            if bm.ndim != 2 or bm.shape[0] != bm.shape[1] or bm.shape[0] > MAX_BITMAP:
                raise ValueError("mask bitmap must be square and at most %dx%d, got %s" % (MAX_BITMAP, MAX_BITMAP, bm.shape))
            n = bm.shape[0]
            self.bitmap_hw[b] = n
            self.bitmaps[b, :n * n] = torch.from_numpy((bm != 0).astype(np.uint8).reshape(-1))
        return self


def expand_masks(stage: torch.Tensor, bitmap_hw: torch.Tensor, bitmaps: torch.Tensor,
                 mask_shapes: Sequence[Tuple] = misc.PYRAMID_SHAPES) -> List[torch.Tensor]:
    """Seven float masks (VGG order, shallowest first, each with a leading batch dimension) from device-resident
    descriptors -- `torch.stack` of `misc.expand_mask_descriptor` over the batch, bit for bit."""
    if not stage.is_cuda:
        raise RuntimeError("expand_masks: descriptors must be CUDA tensors (host expansion is misc.expand_mask_descriptor)")
    B = stage.shape[0]
    deepest_first = tuple(reversed(tuple(mask_shapes)))
    out = []
    for depth, shape in enumerate(deepest_first):
        level = torch.empty((B,) + tuple(shape), dtype=torch.float32, device=stage.device)
        H, W = (shape[1], shape[2]) if len(shape) == 3 else (1, shape[0])
        call("spyr_expand_mask_level", stage.data_ptr(), bitmap_hw.data_ptr(), bitmaps.data_ptr(), bitmaps.stride(0), B, depth,
             H, W, level.data_ptr())
        out.append(level)
    out.reverse()
    return out


def normalize_images(images_u8: torch.Tensor) -> torch.Tensor:
    """uint8 (B, C, H, W) -> float32 in [-1, 1]: `to_tensor` then per-(sample, channel) min-max (data.py:49-53)."""
    if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or not images_u8.is_contiguous():
        raise RuntimeError("normalize_images: needs a contiguous CUDA uint8 tensor")
    B, C, H, W = images_u8.shape
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=images_u8.device)
    call("spyr_image_u8_minmax_normalize", images_u8.data_ptr(), B * C, H * W, out.data_ptr())
    return out


def one_hot_labels(class_index: torch.Tensor, number_of_classes: int) -> torch.Tensor:
    """int64 one-hot rows as the dataset emits them (data.py:58-59)."""
    out = torch.zeros((class_index.shape[0], number_of_classes), dtype=torch.long, device=class_index.device)
    out.scatter_(1, class_index.view(-1, 1).long(), 1)
    return out


class _Slot(object):
    def __init__(self, batch, image_shape, device):
        pin = True
        self.images = torch.empty((batch,) + tuple(image_shape), dtype=torch.uint8, pin_memory=pin)
        self.classes = torch.empty(batch, dtype=torch.int64, pin_memory=pin)
        self.desc = PackedDescriptors(batch)
        self.d_images = torch.empty(self.images.shape, dtype=torch.uint8, device=device)
        self.d_classes = torch.empty(batch, dtype=torch.int64, device=device)
        self.d_stage = torch.empty(batch, dtype=torch.int32, device=device)
        self.d_hw = torch.empty(batch, dtype=torch.int32, device=device)
        self.d_bitmaps = torch.empty(self.desc.bitmaps.shape, dtype=torch.uint8, device=device)
        self.ready = torch.cuda.Event()
        self.consumed = torch.cuda.Event()
        self.result = None


class DeviceBatchLoader(object):
    """Iterates `(images, labels, masks)` on the device, in the format of the reference's collate function.

    `source` yields host batches `(images_u8, class_index, descriptors)`: a uint8 tensor / array (B, 3, H, W), B class
    indices and B `misc.MaskDescriptor`s (e.g. from `misc.draw_mask_descriptor()` in the dataset, where the reference
    calls `get_masks_for_training()`).  Two pinned staging slots: while the consumer trains on batch i, batch i+1 is
    copied on a separate stream and normalised / expanded there; the consumer's stream only waits for an event."""

    def __init__(self, source: Iterable, batch_size: int, device="cuda", image_shape=(3, 256, 256),
                 number_of_classes: int = 365, mask_shapes: Sequence[Tuple] = misc.PYRAMID_SHAPES) -> None:
        self.source = source
        self.batch_size = batch_size
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceBatchLoader feeds the CUDA path; there is no CPU fallback")
        self.number_of_classes = number_of_classes
        self.mask_shapes = mask_shapes
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [_Slot(batch_size, image_shape, self.device) for _ in range(2)]
        self.dataset = getattr(source, "dataset", None)  # ModelWrapper.train reads len(loader.dataset)

    def __len__(self) -> int:
        return len(self.source)

    def _stage(self, slot: _Slot, host_batch) -> None:
        images_u8, class_index, descriptors = host_batch
        images_u8 = torch.as_tensor(images_u8)
        if images_u8.dtype != torch.uint8 or tuple(images_u8.shape) != tuple(slot.images.shape):
            raise ValueError("expected uint8 images of shape %s, got %s %s" % (tuple(slot.images.shape), images_u8.dtype,
                                                                               tuple(images_u8.shape)))
        # the staging buffers may still be read by the previous copy out of this slot
        slot.consumed.synchronize()
        slot.images.copy_(images_u8)
        slot.classes.copy_(torch.as_tensor(class_index, dtype=torch.int64))
        slot.desc.fill(descriptors)
        with torch.cuda.stream(self.copy_stream):
            slot.d_images.copy_(slot.images, non_blocking=True)
            slot.d_classes.copy_(slot.classes, non_blocking=True)
            slot.d_stage.copy_(slot.desc.stage, non_blocking=True)
            slot.d_hw.copy_(slot.desc.bitmap_hw, non_blocking=True)
            slot.d_bitmaps.copy_(slot.desc.bitmaps, non_blocking=True)
            slot.consumed.record(self.copy_stream)
            images = normalize_images(slot.d_images)
            labels = one_hot_labels(slot.d_classes, self.number_of_classes)
            masks = expand_masks(slot.d_stage, slot.d_hw, slot.d_bitmaps, self.mask_shapes)
            slot.ready.record(self.copy_stream)
        slot.result = (images, labels, masks)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, List[torch.Tensor]]]:
        it = iter(self.source)
        pending: Optional[_Slot] = None
        index = 0
        for host_batch in it:
            slot = self.slots[index & 1]
            index += 1
            self._stage(slot, host_batch)
            if pending is not None:
                yield self._hand_over(pending)
            pending = slot
        if pending is not None:
            yield self._hand_over(pending)

    def _hand_over(self, slot: _Slot):
        main = torch.cuda.current_stream(self.device)
        main.wait_event(slot.ready)
        images, labels, masks = slot.result
        slot.result = None
        # produced on the copy stream, consumed on the caller's: keep the allocator from recycling them early
        for t in [images, labels] + list(masks):
            t.record_stream(main)
        return images, labels, masks
