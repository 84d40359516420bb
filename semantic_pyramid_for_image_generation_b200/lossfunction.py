"""Drop-in replacements of the reference's loss modules (lossfunction.py:8-164) over warp-shuffle reduction kernels.

Same class names, constructor signatures, forward arguments, return shapes and `__repr__` strings.  The adversarial
loss is least-squares (LSGAN), exactly as the reference.
"""
from typing import List, Tuple

import torch
import torch.nn as nn

from ._native import call

F32 = torch.float32
BF16 = torch.bfloat16


def _batch_mask(m, batch):
    """Masks built by get_masks_for_inference(add_batch_size=True) have batch 1: broadcast like the reference's `*`."""
    if m.shape[0] == batch:
        return m
    if m.shape[0] == 1:
        return m.expand(batch, *m.shape[1:])
    raise RuntimeError("mask batch %d does not match the feature batch %d" % (m.shape[0], batch))


def _nhwc(t):
    """NCHW-shaped feature -> NHWC BF16 contiguous (zero-copy for the views VGG16 returns)."""
    from . import ops
    return ops.as_nhwc_bf16(t)


class _RecLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n, *tensors):
        reals, fakes, masks = tensors[:n], tensors[n:2 * n], tensors[2 * n:3 * n]
        dev = fakes[0].device
        from . import ops
        loss = torch.zeros(1, dtype=F32, device=dev)
        sc = ops.scratch(1, dev)  # the level kernels run back to back on this stream: one scratch serves all of them
        saved = []
        for fr, ff, m in zip(reals, fakes, masks):
            m = _batch_mask(m.float(), ff.shape[0]).contiguous()
            if ff.dim() == 4:
                a, b = _nhwc(fr), _nhwc(ff)
                B, H, W, C = b.shape
                if a.shape != b.shape or m.numel() != B * H * W:
                    raise RuntimeError("SemanticReconstructionLoss: level shapes differ (real %s, fake %s, mask %s)" %
                                       (tuple(a.shape), tuple(b.shape), tuple(m.shape)))
                call("spyr_rec_level_fwd", a.data_ptr(), b.data_ptr(), m.data_ptr(), B, H, W, C, loss.data_ptr(),
                     sc.data_ptr())
                saved.append((a, b, m))
            else:
                a, b = fr.float().contiguous(), ff.float().contiguous()
                if a.shape != b.shape or m.shape != b.shape:
                    raise RuntimeError("SemanticReconstructionLoss: vector level shapes differ (real %s, fake %s, mask %s)" %
                                       (tuple(a.shape), tuple(b.shape), tuple(m.shape)))
                call("spyr_rec_vec_fwd", a.data_ptr(), b.data_ptr(), m.data_ptr(), a.shape[0], a.shape[1], loss.data_ptr(),
                     sc.data_ptr())
                saved.append((a, b, m))
        ctx.saved, ctx.n = saved, n
        return loss

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        n = ctx.n
        grads = []
        for i, (a, b, m) in enumerate(ctx.saved):
            if not ctx.needs_input_grad[1 + n + i]:
                grads.append(None)
                continue
            if b.dim() == 4:
                from . import ops
                B, H, W, C = b.shape
                gff = ops.act_like(b)
                call("spyr_rec_level_bwd", a.data_ptr(), b.data_ptr(), m.data_ptr(), B, H, W, C, g.data_ptr(),
                     gff.data_ptr())
                grads.append(gff.permute(0, 3, 1, 2))
            else:
                gff = torch.empty_like(b)
                call("spyr_rec_vec_bwd", a.data_ptr(), b.data_ptr(), m.data_ptr(), b.shape[0], b.shape[1], g.data_ptr(),
                     gff.data_ptr())
                grads.append(gff)
        ctx.saved = None
        return (None,) + (None,) * n + tuple(grads) + (None,) * n


class SemanticReconstructionLoss(nn.Module):
    '''
    Semantic reconstruction loss (reference lossfunction.py:8-68): sum over the pyramid levels of
    mean(|maxpool2(real) - maxpool2(fake)| * maxpool2(mask)); returns a tensor of shape [1]
    '''

    def __init__(self) -> None:
        super(SemanticReconstructionLoss, self).__init__()
        self.max_pooling_2d = nn.MaxPool2d(2)
        self.max_pooling_1d = nn.MaxPool1d(2)

    def __repr__(self):
        return '{}, maxpool kernel size{}'.format(self.__class__.__name__, self.max_pooling_1d.kernel_size)

    def forward(self, features_real: List[torch.Tensor], features_fake: List[torch.Tensor],
                masks: List[torch.Tensor]) -> torch.Tensor:
        assert len(features_real) == len(features_fake) == len(masks)
        n = len(features_fake)
        return _RecLossFn.apply(n, *features_real, *features_fake, *masks)


class _DiversityFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images_fake, latent_inputs):
        img = images_fake.float().contiguous()
        z = latent_inputs.float().contiguous()
        B = img.shape[0]
        img_half = (B // 2) * (img.numel() // B)
        z_half = (z.shape[0] // 2) * (z.numel() // z.shape[0])
        work = torch.empty(2, dtype=F32, device=img.device)
        loss = torch.empty((), dtype=F32, device=img.device)
        from . import ops
        sc = ops.scratch(2, img.device)
        call("spyr_diversity_fwd", img.data_ptr(), img_half, z.data_ptr(), z_half, work.data_ptr(), loss.data_ptr(),
             sc.data_ptr())
        ctx.img, ctx.work, ctx.img_half = img, work, img_half
        return loss

    @staticmethod
    def backward(ctx, g):
        if not ctx.needs_input_grad[0]:
            return None, None
        img = ctx.img
        gimg = torch.zeros_like(img) if 2 * ctx.img_half != img.numel() else torch.empty_like(img)
        call("spyr_diversity_bwd", img.data_ptr(), ctx.img_half, ctx.work.data_ptr(), g.contiguous().data_ptr(),
             gimg.data_ptr())
        return gimg, None


class DiversityLoss(nn.Module):
    '''
    Mini-batch diversity loss (reference lossfunction.py:71-110): L1(z first half, z second half) /
    (L1(images first half, images second half) + 1e-8).  The latent gradient is not produced (the reference
    discards it, model_wrapper.py:168-190).
    '''

    def __init__(self) -> None:
        super(DiversityLoss, self).__init__()
        self.l1_loss = nn.L1Loss(reduction='mean')

    def __repr__(self):
        return self.__class__.__name__

    def forward(self, images_fake: torch.Tensor, latent_inputs: torch.Tensor) -> torch.Tensor:
        assert images_fake.shape[0] > 1
        return _DiversityFn.apply(images_fake, latent_inputs)


class _LsganFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, target):
        p = p.float().contiguous()
        out = torch.empty((), dtype=F32, device=p.device)
        call("spyr_lsgan_fwd", p.data_ptr(), p.numel(), target, out.data_ptr())
        ctx.p, ctx.target = p, target
        return out

    @staticmethod
    def backward(ctx, g):
        p = ctx.p
        gp = torch.empty_like(p)
        call("spyr_lsgan_bwd", p.data_ptr(), p.numel(), ctx.target, g.contiguous().data_ptr(), gp.data_ptr())
        return gp, None


class LSGANGeneratorLoss(nn.Module):
    '''
    Least squares generator loss (reference lossfunction.py:115-137): 0.5 * mean((D(G(z)) - 1)^2)
    '''

    def __init__(self) -> None:
        super(LSGANGeneratorLoss, self).__init__()

    def __repr__(self):
        return '{}'.format(self.__class__.__name__)

    def forward(self, images_fake: torch.Tensor) -> torch.Tensor:
        return _LsganFn.apply(images_fake, 1.0)


def lsgan_term(prediction: torch.Tensor, target: float) -> torch.Tensor:
    """One least-squares term 0.5 * mean((prediction - target)^2) (reference lossfunction.py:131-137,156-164)."""
    return _LsganFn.apply(prediction, float(target))


class LSGANDiscriminatorLoss(nn.Module):
    '''
    Least squares discriminator loss (reference lossfunction.py:140-164): (0.5 * mean((D(x) - 1)^2), 0.5 * mean(D(G(z))^2))
    '''

    def __init__(self) -> None:
        super(LSGANDiscriminatorLoss, self).__init__()

    def __repr__(self):
        return '{}'.format(self.__class__.__name__)

    def forward(self, images_real: torch.Tensor, images_fake: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return _LsganFn.apply(images_real, 1.0), _LsganFn.apply(images_fake, 0.0)
