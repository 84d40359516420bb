"""GPU parity of the device-side input pipeline against the host code it replaces: masks bit-exact
(misc.py:13-68), image normalisation bit-exact with the loader's FP32 expression (data.py:49-53)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _descriptors(n, seed):
    from semantic_pyramid_for_image_generation_b200 import misc
    random.seed(seed)
    np.random.seed(seed)
    # p_random_mask raised so that spatial descriptors of every level show up among n draws
    return [misc.draw_mask_descriptor(p_random_mask=0.7) for _ in range(n)]


def _host_masks(descs):
    from semantic_pyramid_for_image_generation_b200 import misc
    per_sample = [misc.expand_mask_descriptor(d) for d in descs]
    return [torch.stack([m[level] for m in per_sample], dim=0) for level in range(7)]


def _reference_normalize(images_u8):
    # TVF.to_tensor followed by kornia.normalize_min_max(image[None], -1., 1.) (data.py:49-53), FP32 on the host
    x = images_u8.float() / 255.0
    B, C = x.shape[:2]
    flat = x.view(B, C, -1)
    lo = flat.min(-1)[0].view(B, C, 1)
    hi = flat.max(-1)[0].view(B, C, 1)
    return ((1.0 - -1.0) * (flat - lo) / (hi - lo + 1e-6) + -1.0).view(x.shape)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_mask_expansion_is_bit_exact(seed):
    from semantic_pyramid_for_image_generation_b200 import input_pipeline as ip
    descs = _descriptors(48, seed)
    assert any(d.bitmap is not None for d in descs) and any(d.bitmap is None for d in descs)
    packed = ip.PackedDescriptors(len(descs)).fill(descs)
    got = ip.expand_masks(packed.stage.cuda(), packed.bitmap_hw.cuda(), packed.bitmaps.cuda())
    want = _host_masks(descs)
    assert [tuple(t.shape) for t in got] == [tuple(t.shape) for t in want]
    for level, (g, w) in enumerate(zip(got, want)):
        assert g.dtype == torch.float32
        assert torch.equal(g.cpu(), w), "level %d differs" % level


def test_inference_style_descriptors_select_one_level():
    from semantic_pyramid_for_image_generation_b200 import input_pipeline as ip, misc
    descs = [misc.MaskDescriptor(k, None) for k in range(7)]
    packed = ip.PackedDescriptors(7).fill(descs)
    got = ip.expand_masks(packed.stage.cuda(), packed.bitmap_hw.cuda(), packed.bitmaps.cuda())
    for b in range(7):
        want = misc.get_masks_for_inference(b)
        for level in range(7):
            assert torch.equal(got[level][b].cpu(), want[level])


def test_image_normalisation_is_bit_exact():
    from semantic_pyramid_for_image_generation_b200 import input_pipeline as ip
    g = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (5, 3, 256, 256), dtype=torch.uint8, generator=g)
    img[1, 0] = 17                      # constant plane: (x - min) / (0 + eps)
    img[2, :, :, :] = img[2].clamp(40, 200)
    img[3] = img[3][:, :, :].clamp(0, 1)  # two-valued plane
    got = ip.normalize_images(img.cuda()).cpu()
    want = _reference_normalize(img)
    assert torch.equal(got, want), float((got - want).abs().max())
    assert float(got.min()) >= -1.0 and float(got.max()) <= 1.0


def test_device_batch_loader_matches_host_collate():
    from semantic_pyramid_for_image_generation_b200 import input_pipeline as ip
    B, steps = 4, 3
    g = torch.Generator().manual_seed(11)
    host = []
    for s in range(steps):
        images = torch.randint(0, 256, (B, 3, 64, 64), dtype=torch.uint8, generator=g)
        classes = torch.randint(0, 365, (B,), generator=g)
        host.append((images, classes, _descriptors(B, 100 + s)))
    loader = ip.DeviceBatchLoader(host, batch_size=B, image_shape=(3, 64, 64))
    seen = 0
    for (images, labels, masks), (h_img, h_cls, h_desc) in zip(loader, host):
        torch.cuda.synchronize()
        assert torch.equal(images.cpu(), _reference_normalize(h_img))
        assert labels.dtype == torch.long and tuple(labels.shape) == (B, 365)
        assert torch.equal(labels.argmax(dim=1).cpu(), h_cls) and int(labels.sum()) == B
        for got, want in zip(masks, _host_masks(h_desc)):
            assert torch.equal(got.cpu(), want)
        seen += 1
    assert seen == steps
