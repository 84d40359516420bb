"""TEST INFRASTRUCTURE ONLY -- the oracle's networks with BF16 rounding inserted exactly where the B200 path stores BF16.

Why this exists.  The B200 path keeps activations in BF16 between kernels (north_star: BF16 operands, FP32 accumulate).
Against the plain FP32 oracle every kernel is within 5e-3, but whole-network *gradients* differ by ~1e-1 because a
1e-2 forward perturbation flips ~1 % of the non-smooth gates (ReLU / LeakyReLU signs, max-pool arg-max) and each flip
changes its gradient entry by O(1).  That number says nothing about the correctness of the backward kernels.  This
module therefore restates the *same* forward functions (oracle/spyramid_oracle.py, i.e. reference models.py:65-99,
140-155, 183-216, 249-275) with a straight-through `q()` (round to BF16, identity gradient) at the B200 path's storage
points.  Its forward then reproduces the B200 activations to ~1e-3 -- so both sides take the same gates -- and its FP32
autograd backward is the yardstick for the hand-written backward schedules (tests/test_gpu_emulated_parity.py).

Rounding points mirror semantic_pyramid_for_image_generation_b200/engine.py and vgg_engine.py:
conv inputs/outputs, packed weights W/sigma, fused 3-source sums (rounded once), pooled maps, attention P and O; FP32
stays FP32 where the B200 path keeps FP32 (statistics, linear heads, the C->3 tail, losses).
"""
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import spyramid_oracle as O

State = Dict[str, torch.Tensor]


ROUNDING = "bf16"  # "bf16": one BF16 plane (8 mantissa bits); "split": hi + lo BF16 planes (~16 bits), the strict mode


class rounding(object):
    """Context manager selecting what the storage points round to: with rounding("split"): ...  The gradient difference
    between this emulation and the plain FP32 oracle is the sensitivity floor of the network's gradients to forward
    rounding of that size (non-smooth gates flip), the yardstick printed next to the GPU figures in tests/."""

    def __init__(self, mode):
        self.mode, self.prev = mode, None

    def __enter__(self):
        global ROUNDING
        self.prev, ROUNDING = ROUNDING, self.mode
        return self

    def __exit__(self, *exc):
        global ROUNDING
        ROUNDING = self.prev
        return False


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        hi = x.bfloat16().float()
        if ROUNDING == "split":
            return hi + (x - hi).bfloat16().float()
        return hi

    @staticmethod
    def backward(ctx, g):
        return g


def q(x: torch.Tensor) -> torch.Tensor:
    return _RoundSTE.apply(x)


def _w(sd: State, name: str, training: bool) -> torch.Tensor:
    return q(O.sn_weight(sd, name, training))


def _lrelu(x):
    return F.leaky_relu(x, 0.2)


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def self_attention(sd: State, name: str, x: torch.Tensor, training: bool) -> torch.Tensor:
    b, c, h, w = x.shape
    qq = q(F.conv2d(x, _w(sd, name + ".query_convolution", training), sd[name + ".query_convolution.bias"]))
    xp = F.max_pool2d(x, 2)
    k = q(F.conv2d(xp, _w(sd, name + ".key_convolution", training), sd[name + ".key_convolution.bias"]))
    v = q(F.conv2d(xp, _w(sd, name + ".value_convolution", training), sd[name + ".value_convolution.bias"]))
    qq = qq.reshape(b, c // 8, h * w)
    k = k.reshape(b, c // 8, h * w // 4)
    v = v.reshape(b, c // 2, h * w // 4)
    att = q(torch.softmax(torch.einsum("bcq,bck->bqk", qq, k), dim=-1))
    o = q(torch.einsum("bck,bqk->bcq", v, att)).reshape(b, c // 2, h, w)
    t = q(F.conv2d(o, _w(sd, name + ".attention_convolution", training), sd[name + ".attention_convolution.bias"]))
    return q(sd[name + ".gamma"] * t + x)


def vgg16_features(sd: State, images: torch.Tensor) -> List[torch.Tensor]:
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    x = q((images - mean) * (1.0 / std))  # im2col rows are BF16; the kernel multiplies by 1/std
    feats = []
    for idx in O.VGG_CONV_IDX:
        x = q(F.relu(F.conv2d(x, q(sd["vgg16.features.%d.weight" % idx]), sd["vgg16.features.%d.bias" % idx], padding=1)))
        if idx in O.VGG_POOL_AFTER:
            x = F.max_pool2d(x, 2)
            feats.append(x)
    x = q(F.adaptive_avg_pool2d(x, (7, 7))).flatten(1)
    y6 = F.relu(F.linear(x, q(sd["vgg16.classifier.0.weight"]), sd["vgg16.classifier.0.bias"]))
    y7 = F.relu(F.linear(q(y6), q(sd["vgg16.classifier.3.weight"]), sd["vgg16.classifier.3.bias"]))
    feats.append(y7)
    feats.append(F.linear(q(y7), q(sd["vgg16.classifier.6.weight"]), sd["vgg16.classifier.6.bias"]))
    return feats


def _generator_block(sd, p, x, feat, mask, cls, training):
    h = O.conditional_batch_norm(sd, p + ".main_block.0", x, cls, training)
    a = q(_up2(_lrelu(h)))
    xu = q(_up2(x))
    h1 = q(F.conv2d(a, _w(sd, p + ".main_block.3", training), sd[p + ".main_block.3.bias"], padding=1))
    a2 = q(_lrelu(O.conditional_batch_norm(sd, p + ".main_block.4", h1, cls, training)))
    main = F.conv2d(a2, _w(sd, p + ".main_block.6", training), sd[p + ".main_block.6.bias"], padding=1)
    skip = F.conv2d(xu, _w(sd, p + ".residual_mapping.1", training), sd[p + ".residual_mapping.1.bias"])
    # masked_feature_mapping: the feature channels go through the tensor cores in BF16, the mask channel stays FP32
    wf = O.sn_weight(sd, p + ".masked_feature_mapping", training)
    cf = wf.shape[1] - 1
    fm = q(q(feat) * mask)
    mapped = F.conv2d(fm, q(wf[:, :cf]), sd[p + ".masked_feature_mapping.bias"], padding=1) + \
        F.conv2d(mask, wf[:, cf:], None, padding=1)
    return q(main + skip + mapped)


def generator_forward(sd: State, z, features, masks, class_onehot, training: bool = True) -> torch.Tensor:
    out = F.linear(z, O.sn_weight(sd, "linear_layer", training), sd["linear_layer.bias"])
    out = O._linear_block(sd, "linear_block_1", out, features[6] * masks[6], training)
    out = O._linear_block(sd, "linear_block_2", out, features[5] * masks[5], training)
    out = q(_lrelu(out.reshape(out.shape[0], -1, 4, 4)))
    out = q(F.conv2d(out, _w(sd, "convolution_layer.1", training), sd["convolution_layer.1.bias"]))
    level = 4
    for idx in (0, 1, 2, 3, 4, 5):
        if idx == 3:
            out = self_attention(sd, "main_path.3", out, training)
        else:
            out = _generator_block(sd, "main_path.%d" % idx, out, features[level], masks[level], class_onehot, training)
            level -= 1
    xh = O._bn_train(_up2(out), sd, "final_block.1.", 0.1, training)
    a = q(_lrelu(xh * sd["final_block.1.weight"].view(1, -1, 1, 1) + sd["final_block.1.bias"].view(1, -1, 1, 1)))
    a3 = q(_lrelu(F.conv2d(a, _w(sd, "final_block.3", training), sd["final_block.3.bias"], padding=1)))
    return torch.tanh(F.conv2d(a3, O.sn_weight(sd, "final_block.5", training), sd["final_block.5.bias"]))


def discriminator_forward(sd: State, x, labels_onehot, training: bool = True) -> torch.Tensor:
    p = "layers.0"
    h = q(_lrelu(F.conv2d(q(x), _w(sd, p + ".main_block.0", training), sd[p + ".main_block.0.bias"], padding=1)))
    s = q(F.conv2d(h, _w(sd, p + ".main_block.2", training), sd[p + ".main_block.2.bias"], padding=1))
    r = q(F.conv2d(q(F.avg_pool2d(x, 2)), _w(sd, p + ".residual_mapping", training), sd[p + ".residual_mapping.bias"]))
    pooled = F.avg_pool2d(s, 2) + r
    out, act = q(pooled), q(_lrelu(pooled))
    for idx in (1, 2, 3, 4, 5, 6, 7):
        p = "layers.%d" % idx
        if idx == 3:
            out = self_attention(sd, p, out, training)
            act = q(_lrelu(out))
        else:
            h = q(_lrelu(F.conv2d(act, _w(sd, p + ".main_block.1", training), sd[p + ".main_block.1.bias"], padding=1)))
            s = q(F.conv2d(h, _w(sd, p + ".main_block.3", training), sd[p + ".main_block.3.bias"], padding=1) +
                  F.conv2d(out, _w(sd, p + ".residual_mapping", training), sd[p + ".residual_mapping.bias"]))
            pooled = F.avg_pool2d(s, 2)
            out, act = q(pooled), q(_lrelu(pooled))
    feat = _lrelu(out).mean(dim=(2, 3))
    feat = _lrelu(F.linear(feat, O.sn_weight(sd, "layers.11", training), sd["layers.11.bias"]))
    emb = O.sn_weight(sd, "embedding", training)[labels_onehot.argmax(dim=-1, keepdim=True)]
    cls = F.linear(feat, O.sn_weight(sd, "classification", training), sd["classification.bias"])
    return cls + feat * emb
