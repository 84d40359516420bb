"""VGG-16 feature-pyramid encoder on the tensor-core conv kernels: frozen for the GAN step, trainable for fine-tuning.

Mirrors VGG16.forward (reference models.py:183-216): ImageNet normalisation of the [-1,1] image as is
(SURVEY Q8), 13 conv3x3+ReLU, a tap after each of the five max-pools, adaptive 7x7 average pool, fc6 / fc7 / fc8
with taps at classifier index 3 (which the following in-place ReLU turns into ReLU(fc7)) and 6.
In the GAN step the weights are frozen (model_wrapper.py:67-68): packed to BF16 once, only input gradients computed.
Fine-tuning (vgg_16_train.py:134-165) adds dropout in the classifier and the weight / bias gradients of all 16 layers.
"""
import torch

from . import ops
from ._native import call
from .ops import Src

F32 = torch.float32
BF16 = torch.bfloat16

CONV_IDX = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)
POOL_AFTER = (1, 3, 6, 9, 12)  # positions in CONV_IDX followed by MaxPool2d(2)


class VGGPack(object):
    """BF16 operands of the frozen network (built once per weight version and precision mode; split-BF16 mode packs a
    (hi, lo) plane pair per operand)."""

    def __init__(self, vgg16):
        dev = vgg16.features[0].weight.device
        self.split = ops.SPLIT
        self.w, self.b, self.ch = [], [], []
        for j, idx in enumerate(CONV_IDX):
            conv = vgg16.features[idx]
            w = conv.weight.detach().float()
            cout, cin = w.shape[0], w.shape[1]
            if j == 0:
                # im2col operand [cout][32]: k = tap*3 + c
                pk = torch.zeros(cout, 32, dtype=F32, device=dev)
                pk[:, :27] = w.permute(0, 2, 3, 1).reshape(cout, 27)
            else:
                pk = w.permute(2, 3, 0, 1).reshape(9, cout, cin)
            self.w.append(ops.to_act(pk))
            self.b.append(conv.bias.detach().float().contiguous())
            self.ch.append((cin, cout))
        fc6, fc7, fc8 = vgg16.classifier[0], vgg16.classifier[3], vgg16.classifier[6]
        c5 = self.ch[-1][1]
        w6 = fc6.weight.detach().float()
        # torch flattens (C,7,7); the NHWC pipeline flattens (7,7,C)
        w6 = w6.view(w6.shape[0], c5, 7, 7).permute(0, 2, 3, 1).reshape(w6.shape[0], -1)
        n8 = fc8.weight.shape[0]
        self.n8, self.n8_pad = n8, (n8 + 7) // 8 * 8
        w8 = torch.zeros(self.n8_pad, fc8.weight.shape[1], dtype=F32, device=dev)
        w8[:n8] = fc8.weight.detach().float()
        self.fc_w = [ops.to_act(w6), ops.to_act(fc7.weight.detach()), ops.to_act(w8)]
        self.fc_b = [m.bias.detach().float().contiguous() for m in (fc6, fc7, fc8)]
        self.mean = torch.tensor([0.485, 0.456, 0.406], dtype=F32, device=dev)
        self.invstd = 1.0 / torch.tensor([0.229, 0.224, 0.225], dtype=F32, device=dev)


def _splits(ktotal, ntiles, mtiles=1):
    s = max(1, 148 // max(1, ntiles * mtiles))
    return min(s, ktotal)


def _fc_forward(x_bf16, w, bias, B, K, O, relu, dev, want_bf16=True):
    # split-K over the reduction: one FP32 slice per split, summed in split order by the epilogue (no atomics)
    ns = _splits(K // 64, (O + 255) // 256)
    acc = torch.empty((ns, B, O), dtype=F32, device=dev)
    # w may hold more rows than O (fc8: 368 for 365 classes): tell the library where its lo plane really starts
    ops.conv(B, 1, 1, O, [Src(x_bf16, w, K, 1, w_lo_off=w.numel())], f32_out=acc, splits=ns, f32_store=(ns == 1))
    y = torch.empty((B, O), dtype=F32, device=dev)
    yb = ops.act_empty((B, O), dev) if want_bf16 else None
    call("spyr_vec_epilogue", acc.data_ptr(), ns, bias.data_ptr(), None, None, 1 if relu else 0, y.data_ptr(),
         yb.data_ptr() if want_bf16 else None, O, B, O)
    return y, yb


def _dropout(y, p, rng):
    """Inverted dropout of an FP32 activation; returns (y_dropped FP32, BF16 copy, u8 keep mask)."""
    out, outb = torch.empty_like(y), ops.act_empty(y.shape, y.device)
    mask = torch.empty(y.shape, dtype=torch.uint8, device=y.device)
    seed, offset = rng
    call("spyr_dropout_fwd", y.data_ptr(), y.numel(), p, seed, offset, out.data_ptr(), outb.data_ptr(), mask.data_ptr())
    return out, outb, mask


def vgg_forward(pk, img, save, dropout=None):
    """img: (B,3,H,W) FP32 NCHW.  Returns ([p1..p5] NHWC BF16, fc7 post-ReLU (B,4096) FP32, logits (B,365) FP32), ctx.

    dropout = (p, seed, offset) applies nn.Dropout(p) after ReLU(fc6) and ReLU(fc7) (torchvision classifier[2], [5])."""
    B, Ci, H, W = img.shape
    dev = img.device
    col = ops.act_empty((B, H, W, 32), dev)
    call("spyr_im2col3x3", img.data_ptr(), B, H, W, pk.mean.data_ptr(), pk.invstd.data_ptr(), col.data_ptr())
    acts, pools = [], []
    h, w = H, W
    x = col
    for j, (cin, cout) in enumerate(pk.ch):
        src = Src(x, pk.w[j], 32, 1) if j == 0 else Src(x, pk.w[j], cin, 3)
        _, x = ops.conv(B, h, w, cout, [src], bias=pk.b[j], want_raw=False, want_act=True, act=1)
        acts.append(x)
        if j in POOL_AFTER:
            x = ops.maxpool2(x)
            pools.append(x)
            h, w = h // 2, w // 2
    c5 = pk.ch[-1][1]
    pooled = ops.act_empty((B, 7, 7, c5), dev)
    call("spyr_adaptive_avgpool_fwd", x.data_ptr(), pooled.data_ptr(), B, h, w, 7, 7, c5)
    K6 = 49 * c5
    y6, y6b = _fc_forward(pooled, pk.fc_w[0], pk.fc_b[0], B, K6, pk.fc_w[0].shape[0], True, dev)
    drop = None
    x7, x8 = y6, None
    if dropout is not None:
        p, seed, offset = dropout
        x7, y6b, m6 = _dropout(y6, p, (seed, offset))
    y7, y7b = _fc_forward(y6b, pk.fc_w[1], pk.fc_b[1], B, pk.fc_w[1].shape[1], pk.fc_w[1].shape[0], True, dev)
    x8 = y7
    if dropout is not None:
        x8, y7b, m7 = _dropout(y7, p, (seed, offset + y6.numel()))
        drop = (p, m6, m7, x7, x8)
    y8, _ = _fc_forward(y7b, pk.fc_w[2], pk.fc_b[2], B, pk.fc_w[2].shape[1], pk.n8, False, dev, want_bf16=False)
    ctx = (col, acts, pools, y6, y7, (B, H, W), pooled, drop) if save else None
    return pools, y7, y8, ctx


def _fc_dgrad(g_bf16, w, B, K, O, dev):
    """g (B,1,1,K) @ W[K][O] -> FP32 split-K slices (ns,B,O) via the MN-major view of the forward pack; returns
    (slices, ns) for spyr_vec_epilogue / _sum_slices."""
    ns = _splits(K // 64, (O + 255) // 256)
    acc = torch.empty((ns, B, O), dtype=F32, device=dev)
    ops.conv(B, 1, 1, O, [Src(g_bf16, w, K, 1, mn=True)], f32_out=acc, splits=ns, f32_store=(ns == 1))
    return acc, ns


def _sum_slices(acc, ns, B, O, add=None, want_bf16=False):
    """FP32 sum of the split-K slices in split order (+ add); optionally also as a BF16 map."""
    out = torch.empty((B, O), dtype=F32, device=acc.device)
    outb = ops.act_empty((B, O), acc.device) if want_bf16 else None
    call("spyr_vec_epilogue", acc.data_ptr(), ns, None, add.data_ptr() if add is not None else None, None, 0,
         out.data_ptr(), outb.data_ptr() if want_bf16 else None, O, B, O)
    return out, outb


def _conv_param_grads(wg, j, x_in, g, B, h, w, cin, cout):
    """Weight and bias gradient of conv j from its input and the gradient of its pre-activation output."""
    weight, bias = wg["conv"][j]
    taps, k_in, stride = (1, 32, 27) if j == 0 else (9, cin, cin)
    gw = torch.zeros(taps * stride * cout, dtype=F32, device=g.device)
    ops.wgrad(x_in, g, gw.data_ptr(), B, h, w, k_in, cout, 3 if j else 1, cin_stride=stride if j == 0 else 0)
    if j == 0:
        # im2col rows k = tap*3 + c  ->  (Cout, 3, 3, 3): nine taps of three channels
        call("spyr_wgrad_to_oihw", gw.data_ptr(), weight.data_ptr(), 9, 3, cout, 3, 0)
    else:
        call("spyr_wgrad_to_oihw", gw.data_ptr(), weight.data_ptr(), 9, cin, cout, cin, 0)
    bias.zero_()
    ops.colsum(g, cout, bias.data_ptr())
    wg["keep"].append(gw)


def _fc_param_grads(wg, i, gy, x):
    """gw[o][k] = sum_b gy[b][o] x[b][k], gb[o] = sum_b gy[b][o] for classifier layer i (batch in chunks of <= 32 rows)."""
    weight, bias = wg["fc"][i]
    weight.zero_()
    bias.zero_()
    for b0 in range(0, gy.shape[0], 32):
        ops.linear_bwd_w(gy[b0:b0 + 32].contiguous(), x[b0:b0 + 32].contiguous(), weight.data_ptr(), bias.data_ptr())


def vgg_backward(pk, ctx, g_pools, g7, g8, wg=None):
    """d/dimage (NCHW FP32) from the gradients of the seven taps (any may be None).

    wg (fine-tuning): {"conv": [(weight_grad, bias_grad)] * 13, "fc": [...] * 3, "keep": []} of FP32 tensors in the
    parameters' own layouts, filled here; the returned image gradient is None when `wg["image"]` is False."""
    col, acts, pools, y6, y7, (B, H, W), pooled, drop = ctx
    dev = col.device
    c5 = pk.ch[-1][1]
    n7, n6 = y7.shape[1], y6.shape[1]
    g_pool_in = None  # gradient w.r.t. the (B,h5,w5,c5) input of the adaptive pool
    hp, wp = pools[-1].shape[1], pools[-1].shape[2]
    if g7 is not None or g8 is not None:
        acc7 = None  # FP32 (B, n7): d/d(ReLU(fc7)) = W8^T g8 (+ g7)
        g7c = g7.contiguous().float() if g7 is not None else None
        if g8 is not None:
            g8b = ops.act_zeros((B, pk.n8_pad), dev)
            call("spyr_vec_epilogue", g8.contiguous().data_ptr(), 1, None, None, None, 0, None, g8b.data_ptr(), pk.n8_pad, B,
                 pk.n8)
            sl7, ns7 = _fc_dgrad(g8b, pk.fc_w[2], B, pk.n8_pad, n7, dev)
            acc7, _ = _sum_slices(sl7, ns7, B, n7, add=g7c)
        else:
            acc7 = g7c
        if wg is not None and g8 is not None:
            _fc_param_grads(wg, 2, g8.contiguous(), drop[4] if drop is not None else y7)
        if drop is not None:
            # through the dropout that follows ReLU(fc7): d/dy7 = d/dx8 * mask / (1 - p)
            if g8 is None:
                acc7 = acc7.clone()
            call("spyr_dropout_bwd", acc7.data_ptr(), drop[2].data_ptr(), acc7.numel(), drop[0], acc7.data_ptr())
        g7b = ops.act_empty((B, n7), dev)
        g7f = torch.empty((B, n7), dtype=F32, device=dev) if wg is not None else None
        call("spyr_vec_epilogue", acc7.data_ptr(), 1, None, None, y7.data_ptr(), 2,
             g7f.data_ptr() if g7f is not None else None, g7b.data_ptr(), n7, B, n7)
        if wg is not None:
            _fc_param_grads(wg, 1, g7f, drop[3] if drop is not None else y6)
        sl6, ns6 = _fc_dgrad(g7b, pk.fc_w[1], B, n7, n6, dev)
        if drop is not None:
            acc6, _ = _sum_slices(sl6, ns6, B, n6)
            call("spyr_dropout_bwd", acc6.data_ptr(), drop[1].data_ptr(), acc6.numel(), drop[0], acc6.data_ptr())
            sl6, ns6 = acc6, 1
        g6b = ops.act_empty((B, n6), dev)
        g6f = torch.empty((B, n6), dtype=F32, device=dev) if wg is not None else None
        call("spyr_vec_epilogue", sl6.data_ptr(), ns6, None, None, y6.data_ptr(), 2,
             g6f.data_ptr() if g6f is not None else None, g6b.data_ptr(), n6, B, n6)
        if wg is not None:
            # fc6 reads torch's (C,7,7) flattening: present the pooled map in that order so the gradient lands in the
            # parameter's own layout
            x6 = ops.nhwc_to_nchw(pooled).view(B, -1)
            _fc_param_grads(wg, 0, g6f, x6)
        slp, nsp = _fc_dgrad(g6b, pk.fc_w[0], B, n6, 49 * c5, dev)
        _, g_pooled = _sum_slices(slp, nsp, B, 49 * c5, want_bf16=True)
        g_pooled = g_pooled.view(B, 7, 7, c5)
        g_pool_in = ops.act_empty((B, hp, wp, c5), dev)
        call("spyr_adaptive_avgpool_bwd", g_pooled.data_ptr(), g_pools[4].data_ptr() if g_pools[4] is not None else None,
             g_pool_in.data_ptr(), B, hp, wp, 7, 7, c5)
    else:
        g_pool_in = g_pools[4]
    if g_pool_in is None:
        g_pool_in = ops.act_zeros((B, hp, wp, c5), dev)
    # walk the conv stack backwards; g = gradient w.r.t. the output of the current stage
    g = g_pool_in
    level = 4
    for j in range(len(pk.ch) - 1, -1, -1):
        cin, cout = pk.ch[j]
        a = acts[j]
        h, w = a.shape[1], a.shape[2]
        if j in POOL_AFTER:
            # g is d/d(pool output): route to the arg-max and gate by the ReLU
            g = ops.maxpool2_bwd(a, g, True)
            level -= 1
        if wg is not None:
            # g is now the gradient of conv j's pre-activation output
            x_in = col if j == 0 else (pools[level] if (j - 1) in POOL_AFTER else acts[j - 1])
            _conv_param_grads(wg, j, x_in, g, B, h, w, cin, cout)
        if j == 0:
            if wg is not None and not wg.get("image", True):
                return None
            g_col, _ = ops.conv(B, h, w, 32, [Src(g, pk.w[0], cout, 1, mn=True)])
            g_img = torch.empty((B, 3, h, w), dtype=F32, device=dev)
            call("spyr_col2im3x3", g_col.data_ptr(), B, h, w, pk.invstd.data_ptr(), g_img.data_ptr(), 0)
            return g_img
        if (j - 1) in POOL_AFTER:
            # the input of conv j is a pooled tap: add the tap's own gradient, no gate here
            res = g_pools[level]
            g, _ = ops.conv(B, h, w, cin, [Src(g, pk.w[j], cout, 3, mn=True)], residual=res)
        else:
            g, _ = ops.conv(B, h, w, cin, [Src(g, pk.w[j], cout, 3, mn=True)], dmask=acts[j - 1], dmask_slope=0.0)
    raise AssertionError("unreachable")
