"""Forward / backward schedules of the Generator, Discriminator, SelfAttention and VGG-16 over the C-ABI kernels.

Each `*_forward` returns its outputs plus a context object holding exactly the tensors the matching `*_backward`
needs; `*_backward` writes parameter gradients into a flat FP32 arena (one slice per parameter, the order of
`module.parameters()`) and returns input gradients.  The kernel sequence follows the reference's forward order
(models.py:65-99, 140-155, 183-216, 249-275, 317-339, 362-375, 408-419, 453-466) so that spectral-norm power
iterations and batch-norm running statistics advance exactly as in the reference.
"""
import torch

from . import ops
from ._native import call, ptr
from .ops import Src, LRELU

F32 = torch.float32
BF16 = torch.bfloat16


class GradArena(object):
    """Flat FP32 gradient storage for all parameters of a module (offsets in floats, 64-float aligned).

    Offsets are kept per position in `module.parameters()` and resolved through the module's CURRENT parameter objects,
    so a `copy.deepcopy` of the module (an EMA copy, say) or a `module.to(...)` that re-creates parameters keeps working:
    the id -> position map is rebuilt whenever the parameter objects are not the ones it was built from."""

    def __init__(self, module):
        self.module = module
        self.sizes = [p.numel() for p in module.parameters()]
        self.offset_list = []
        off = 0
        for n in self.sizes:
            self.offset_list.append(off)
            off += (n + 63) // 64 * 64
        self.total = max(off, 64)
        self._index = None
        self._params = None

    def __deepcopy__(self, memo):
        import copy
        new = GradArena.__new__(GradArena)
        new.module = copy.deepcopy(self.module, memo)  # memo holds the copy being built: no second copy is made
        new.sizes, new.offset_list, new.total = list(self.sizes), list(self.offset_list), self.total
        new._index = new._params = None
        return new

    @property
    def params(self):
        cur = list(self.module.parameters())
        if self._params is None or len(cur) != len(self._params) or any(a is not b for a, b in zip(cur, self._params)):
            if [p.numel() for p in cur] != self.sizes:
                raise RuntimeError("the parameter list of the module changed shape since construction")
            self._params = cur
            self._index = {id(p): i for i, p in enumerate(cur)}
        return self._params

    def offset(self, param):
        self.params  # refresh the id map if the parameter objects changed
        return self.offset_list[self._index[id(param)]]

    def new(self, device):
        return torch.zeros(self.total, dtype=F32, device=device)

    def ptr(self, arena, param):
        return arena.data_ptr() + 4 * self.offset(param)

    def views(self, arena, needs):
        out = []
        for p, o, need in zip(self.params, self.offset_list, needs):
            out.append(arena[o:o + p.numel()].view(p.shape) if need else None)
        return out


def _f32c(t):
    if t.dtype != F32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------------------
# SelfAttention (models.py:249-275); x is the raw NHWC residual stream
# ------------------------------------------------------------------------------------------------
FUSED_ATTENTION = True  # tests flip this to compare the fused kernel with the per-image GEMM path


def fused_attention_ok(hw, d, nk, dv):
    # split-BF16 mode: the fused kernel keeps P in single-plane BF16 -> per-image GEMMs on (hi, lo) operands instead
    return FUSED_ATTENTION and not ops.SPLIT and hw % 128 == 0 and 8 <= d <= 64 and d % 8 == 0 and nk % 64 == 0 and 64 <= nk <= 256 \
        and dv in (64, 128)


def attention_forward(att, key, st, x, want_act, save):
    B, H, W, Cc = x.shape
    d, dv, nk = Cc // 8, Cc // 2, (H * W) // 4
    q, _ = ops.conv(B, H, W, d, [Src(x, st.w(key + ".query_convolution"), Cc, 1)], bias=att.query_convolution.bias)
    xp = ops.maxpool2(x)
    k, _ = ops.conv(B, H // 2, W // 2, d, [Src(xp, st.w(key + ".key_convolution"), Cc, 1)], bias=att.key_convolution.bias)
    v, _ = ops.conv(B, H // 2, W // 2, dv, [Src(xp, st.w(key + ".value_convolution"), Cc, 1)],
                    bias=att.value_convolution.bias)
    if fused_attention_ok(H * W, d, nk, dv):
        # one kernel: S = Q K^T, softmax over keys, O = P V; the attention map is written once (BF16) only for backward
        Pm = ops.act_empty((B, H, W, nk), x.device) if save else None
        O = ops.act_empty((B, H, W, dv), x.device)
        call("spyr_sagan_attention_fwd", q.data_ptr(), k.data_ptr(), v.data_ptr(), O.data_ptr(), ptr(Pm), B, H * W, d, nk, dv)
    else:
        S = torch.empty((B, H * W, nk), dtype=F32, device=x.device)
        ops.conv(B, H, W, nk, [Src(q, k, d, 1, per_image=True)], f32_out=S, f32_store=True)
        Pm = ops.act_empty((B, H, W, nk), x.device)
        call("spyr_softmax_rows_fwd", S.data_ptr(), Pm.data_ptr(), B * H * W, nk)
        O, _ = ops.conv(B, H, W, dv, [Src(Pm, v, nk, 1, mn=True, per_image=True)])
    t, _ = ops.conv(B, H, W, Cc, [Src(O, st.w(key + ".attention_convolution"), dv, 1)],
                    bias=att.attention_convolution.bias)
    out = ops.act_like(x)
    out_act = ops.act_like(x) if want_act else None
    call("spyr_gamma_residual_fwd", t.data_ptr(), x.data_ptr(), att.gamma.data_ptr(), out.data_ptr(), ptr(out_act), LRELU,
         x.numel())
    ctx = (x, xp, q, k, v, Pm, O, t) if save else None
    return out, out_act, ctx


def attention_backward(att, key, st, sn, ga, gw, grad, ctx, g_out, want_wgrad=True):
    """Returns d/dx.  Parameter gradients go to gw (spectral-normed weights) and grad (biases, gamma)."""
    x, xp, q, k, v, Pm, O, t = ctx
    B, H, W, Cc = x.shape
    d, dv, nk = Cc // 8, Cc // 2, (H * W) // 4
    dev = x.device
    gt = ops.act_like(g_out)
    sc = ops.scratch(1, dev)
    call("spyr_gamma_residual_bwd", g_out.data_ptr(), t.data_ptr(), att.gamma.data_ptr(), gt.data_ptr(),
         ga.ptr(grad, att.gamma), g_out.numel(), sc.data_ptr())
    ko = key + ".attention_convolution"
    if want_wgrad:
        ops.wgrad(O, gt, sn.gw_ptr(gw, ko), B, H, W, dv, Cc, 1)
        ops.colsum(gt, Cc, ga.ptr(grad, att.attention_convolution.bias))
    gO, _ = ops.conv(B, H, W, dv, [Src(gt, st.w(ko), Cc, 1, mn=True)])
    # dV = P^T dO, dP = dO V^T, dS = softmax', dQ = dS K, dK = dS^T Q
    dV = torch.zeros((B, nk, dv), dtype=F32, device=dev)
    ops.wgrad(Pm, gO, dV.data_ptr(), B, H, W, nk, dv, 1, per_image=True)
    dP = torch.empty((B, H * W, nk), dtype=F32, device=dev)
    ops.conv(B, H, W, nk, [Src(gO, v, dv, 1, per_image=True)], f32_out=dP, f32_store=True)
    dS = ops.act_empty((B, H, W, nk), dev)
    call("spyr_softmax_rows_bwd", Pm.data_ptr(), dP.data_ptr(), dS.data_ptr(), B * H * W, nk)
    gq, _ = ops.conv(B, H, W, d, [Src(dS, k, nk, 1, mn=True, per_image=True)])
    dK = torch.zeros((B, nk, d), dtype=F32, device=dev)
    ops.wgrad(dS, q, dK.data_ptr(), B, H, W, nk, d, 1, per_image=True)
    gk = ops.cast_bf16(dK).view(B, H // 2, W // 2, d)
    gv = ops.cast_bf16(dV).view(B, H // 2, W // 2, dv)
    kq, kk, kv = key + ".query_convolution", key + ".key_convolution", key + ".value_convolution"
    if want_wgrad:
        ops.wgrad(x, gq, sn.gw_ptr(gw, kq), B, H, W, Cc, d, 1)
        ops.colsum(gq, d, ga.ptr(grad, att.query_convolution.bias))
        ops.wgrad(xp, gk, sn.gw_ptr(gw, kk), B, H // 2, W // 2, Cc, d, 1)
        ops.colsum(gk, d, ga.ptr(grad, att.key_convolution.bias))
        ops.wgrad(xp, gv, sn.gw_ptr(gw, kv), B, H // 2, W // 2, Cc, dv, 1)
        ops.colsum(gv, dv, ga.ptr(grad, att.value_convolution.bias))
    gxp, _ = ops.conv(B, H // 2, W // 2, Cc, [Src(gk, st.w(kk), d, 1, mn=True), Src(gv, st.w(kv), dv, 1, mn=True)])
    gx, _ = ops.conv(B, H, W, Cc, [Src(gq, st.w(kq), d, 1, mn=True)], residual=g_out)
    ops.maxpool2_bwd(x, gxp, False, out=gx)
    return gx


# ------------------------------------------------------------------------------------------------
# Generator (models.py:65-99)
# ------------------------------------------------------------------------------------------------
def _cbn_forward(cbn, x, cls, mode, training, want_xu=False):
    B, H, W, Cc = x.shape
    bn = cbn.batch_norm
    sums = ops.bn_stats(x) if training else None
    mr = ops.bn_finalize(sums, B * H * W, Cc, bn.eps, bn.momentum, bn.running_mean, bn.running_var,
                         bn.num_batches_tracked, training)
    emb = cbn.embedding.weight
    a, xu = ops.bn_act(x, mr, emb.data_ptr(), emb.data_ptr() + 4 * Cc, 2 * Cc, cls, mode, want_xu=want_xu)
    return a, xu, mr


def _cbn_backward(cbn, ga, grad, x, mr, cls, g, mode, residual=None):
    """mode 0: g = d/dy at x's resolution (gated).  mode 1: g = d/da at 2x resolution, a = up2(lrelu(y))."""
    B, H, W, Cc = x.shape
    emb = cbn.embedding.weight
    sp, hp = emb.data_ptr(), emb.data_ptr() + 4 * Cc
    S = ops.bn_bwd_partials(B, H * W, Cc, x.device)
    gy = ops.act_like(x) if mode else g
    call("spyr_bn_bwd_reduce", g.data_ptr(), x.data_ptr(), mr.data_ptr(), sp, hp, 2 * Cc, cls.data_ptr(), LRELU, mode,
         gy.data_ptr() if mode else None, S.data_ptr(), B, H, W, Cc)
    M = torch.empty(2 * Cc, dtype=F32, device=x.device)
    ge = ga.ptr(grad, emb)
    call("spyr_bn_bwd_finalize", S.data_ptr(), B, Cc, float(B * H * W), sp, 2 * Cc, cls.data_ptr(), M.data_ptr(), None,
         None)
    # the gain / bias gradients of the class rows are leaves: off the reduce -> finalize -> apply chain
    ops.LEAF.run([S, cls], "spyr_bn_bwd_params", S.data_ptr(), B, Cc, 2 * Cc, cls.data_ptr(), ge, ge + 4 * Cc)
    gx = ops.act_like(x)
    call("spyr_bn_bwd_apply", gy.data_ptr(), x.data_ptr(), mr.data_ptr(), sp, 2 * Cc, cls.data_ptr(), M.data_ptr(),
         ptr(residual), gx.data_ptr(), B, H, W, Cc, 0)
    return gx


def _gblock_forward(blk, key, st, x, feat, mask, cls, training, save):
    B, H, W, Cin = x.shape
    k3, k6, kr, kf = key + ".main_block.3", key + ".main_block.6", key + ".residual_mapping.1", key + ".masked_feature_mapping"
    c3, c6, cr, cf_ = blk.main_block[3], blk.main_block[6], blk.residual_mapping[1], blk.masked_feature_mapping
    Cout = c3.shape[0]
    Cf = cf_.shape[1] - 1
    a, xu, mr1 = _cbn_forward(blk.main_block[0], x, cls, 1, training, want_xu=True)
    h1, _ = ops.conv(B, 2 * H, 2 * W, Cout, [Src(a, st.w(k3), Cin, 3)], bias=c3.bias)
    a2, _, mr2 = _cbn_forward(blk.main_block[4], h1, cls, 0, training)
    fm = ops.as_nhwc_bf16(feat, mask=mask)
    if fm.shape != (B, 2 * H, 2 * W, Cf):
        raise RuntimeError("generator block %s expects features of shape (B,%d,%d,%d), got NHWC %s" %
                           (key, Cf, 2 * H, 2 * W, tuple(fm.shape)))
    out, _ = ops.conv(B, 2 * H, 2 * W, Cout,
                      [Src(a2, st.w(k6), Cout, 3), Src(xu, st.w(kr), Cin, 1), Src(fm, st.w(kf), Cf, 3)],
                      bias=c6.bias, bias2=cr.bias, bias3=cf_.bias, stencil_mask=mask, stencil_w=st.stencil_w(kf))
    ctx = (x, a, xu, mr1, h1, a2, mr2, fm, mask) if save else None
    return out, ctx


def _gblock_backward(blk, key, st, sn, ga, gw, grad, ctx, cls, g_out):
    x, a, xu, mr1, h1, a2, mr2, fm, mask = ctx
    B, H, W, Cin = x.shape
    k3, k6, kr, kf = key + ".main_block.3", key + ".main_block.6", key + ".residual_mapping.1", key + ".masked_feature_mapping"
    c3, c6, cr, cf_ = blk.main_block[3], blk.main_block[6], blk.residual_mapping[1], blk.masked_feature_mapping
    Cout = c3.shape[0]
    Cf = cf_.shape[1] - 1
    H2, W2 = 2 * H, 2 * W
    ops.colsum(g_out, Cout, ga.ptr(grad, c6.bias), ga.ptr(grad, cr.bias), ga.ptr(grad, cf_.bias))
    ops.wgrad(a2, g_out, sn.gw_ptr(gw, k6), B, H2, W2, Cout, Cout, 3)
    ops.wgrad(xu, g_out, sn.gw_ptr(gw, kr), B, H2, W2, Cin, Cout, 1)
    ops.wgrad(fm, g_out, sn.gw_ptr(gw, kf), B, H2, W2, Cf, Cout, 3, cin_stride=Cf + 1)
    ops.stencil_wgrad(mask, g_out, B, H2, W2, Cout, sn.gw_ptr(gw, kf), Cf + 1, Cf)
    gy2, _ = ops.conv(B, H2, W2, Cout, [Src(g_out, st.w(k6), Cout, 3, mn=True)], dmask=a2, dmask_slope=LRELU)
    g_h1 = _cbn_backward(blk.main_block[4], ga, grad, h1, mr2, cls, gy2, 0)
    ops.wgrad(a, g_h1, sn.gw_ptr(gw, k3), B, H2, W2, Cin, Cout, 3)
    ops.colsum(g_h1, Cout, ga.ptr(grad, c3.bias))
    g_a, _ = ops.conv(B, H2, W2, Cin, [Src(g_h1, st.w(k3), Cout, 3, mn=True)])
    # skip path: up2^T commutes with the 1x1 conv, so transpose-upsample the Cout-channel gradient first
    g_lo = ops.act_empty((B, H, W, Cout), x.device)
    call("spyr_up2_bwd", g_out.data_ptr(), g_lo.data_ptr(), B, H, W, Cout)
    g_skip, _ = ops.conv(B, H, W, Cin, [Src(g_lo, st.w(kr), Cout, 1, mn=True)])
    return _cbn_backward(blk.main_block[0], ga, grad, x, mr1, cls, g_a, 1, residual=g_skip)


def _mask_for(mask, batch, shape_tail, what):
    """FP32 contiguous mask of `batch` rows; batch-1 masks (get_masks_for_inference(add_batch_size=True), misc.py:86-96)
    broadcast like the reference's `features * masks`; anything else that does not match raises instead of letting a
    kernel read out of bounds."""
    m = mask if mask.dtype == F32 else mask.float()
    if m.shape[0] == 1 and batch > 1:
        m = m.expand(batch, *m.shape[1:])
    if m.shape[0] != batch or m.numel() != batch * shape_tail:
        raise RuntimeError("%s: mask of shape %s does not match batch %d x %d elements" %
                           (what, tuple(mask.shape), batch, shape_tail))
    return m.contiguous()


def prepare_generator(G):
    """Runs the spectral-norm pass of the NEXT generator forward (power iteration, sigma, BF16 operand pack: four
    bandwidth-bound kernels over all 30 M weights, ~0.2 ms) on the current stream, ahead of the forward itself.  The
    weights do not depend on the VGG features the forward waits for, so ModelWrapper issues this on a side stream next
    to VGG(real) instead of behind it.  generator_forward consumes the prepared state once (same kernels, same order on
    the same u, v: results are unchanged)."""
    st = G._sn.forward(G.training)
    done = torch.cuda.Event()
    done.record(torch.cuda.current_stream())
    G._sn_prepared = (st, done, G.training)


def generator_forward(G, z, features, masks, class_id, save):
    training = G.training
    if save and not training:
        raise RuntimeError("Generator: backward through an eval()-mode forward is not implemented (the batch-norm backward "
                           "kernels assume batch statistics); call .train() or wrap the call in torch.no_grad()")
    sn = G._sn
    prepared = G.__dict__.pop("_sn_prepared", None)
    if prepared is not None and prepared[2] == training:
        st = prepared[0]
        cur = torch.cuda.current_stream()
        cur.wait_event(prepared[1])
        for t in (st.packed, st.stencil, st.saved):  # allocated on the preparing stream, used (and freed) on this one
            if t is not None:
                t.record_stream(cur)
    else:
        st = sn.forward(training)
    B = z.shape[0]
    if class_id.shape[0] != B:
        raise RuntimeError("Generator: class_id batch %d does not match the latent batch %d" % (class_id.shape[0], B))
    cls = ops.argmax_rows(class_id)
    z = _f32c(z)
    f6, f5 = _f32c(features[6]), _f32c(features[5])
    if f6.shape[0] != B or f5.shape[0] != B:
        raise RuntimeError("Generator: feature batch does not match the latent batch %d" % B)
    m6 = _mask_for(masks[6], B, f6.shape[1], "Generator (logits level)")
    m5 = _mask_for(masks[5], B, f5.shape[1], "Generator (fc7 level)")
    lb1, lb2 = G.linear_block_1, G.linear_block_2
    h0 = ops.linear_fwd(z, G.linear_layer.weight_orig, st.sigma("linear_layer"), G.linear_layer.bias)
    t1 = ops.linear_fwd(f6, lb1.masked_feature_mapping.weight_orig, st.sigma("linear_block_1.masked_feature_mapping"),
                        lb1.masked_feature_mapping.bias, xmask=m6)
    h1 = ops.linear_fwd(h0, lb1.main_block[1].weight_orig, st.sigma("linear_block_1.main_block.1"), lb1.main_block[1].bias,
                        in_slope=LRELU, y_add=t1)
    t2 = ops.linear_fwd(f5, lb2.masked_feature_mapping.weight_orig, st.sigma("linear_block_2.masked_feature_mapping"),
                        lb2.masked_feature_mapping.bias, xmask=m5)
    h2 = ops.linear_fwd(h1, lb2.main_block[1].weight_orig, st.sigma("linear_block_2.main_block.1"), lb2.main_block[1].bias,
                        in_slope=LRELU, y_add=t2)
    c_in = h2.shape[1] // 16
    a0 = ops.nchw_to_nhwc(h2.view(B, c_in, 4, 4), slope=LRELU)
    cl = G.convolution_layer[1]
    x, _ = ops.conv(B, 4, 4, cl.shape[0], [Src(a0, st.w("convolution_layer.1"), c_in, 1)], bias=cl.bias)
    level = 4
    block_ctx = []
    mask_list = []
    for idx, layer in enumerate(G.main_path):
        key = "main_path.%d" % idx
        if idx == 3:
            x, _, c = attention_forward(layer, key, st, x, False, save)
        else:
            hw = 4 * x.shape[1] * x.shape[2]  # the block's output resolution = the feature level's
            mask = _mask_for(masks[level], B, hw, "Generator (pyramid level %d)" % level)
            x, c = _gblock_forward(layer, key, st, x, features[level], mask, cls, training, save)
            level -= 1
        block_ctx.append(c)
    # final block: up2 -> BN -> LeakyReLU -> conv3x3 -> LeakyReLU -> conv1x1 -> tanh
    Bx, H, W, c5 = x.shape
    bn = G.final_block[1]
    xu = None
    if training:
        # the normalised tensor is up2(x): materialise it once (168 MB at cf=1) -- statistics, activation and both
        # backward passes then run as plain same-resolution kernels instead of re-interpolating x four times
        xu, sums = ops.up2_stats(x)
        mr = ops.bn_finalize(sums, B * 4 * H * W, c5, bn.eps, bn.momentum, bn.running_mean, bn.running_var,
                             bn.num_batches_tracked, True)
        a, _ = ops.bn_act(xu, mr, bn.weight.data_ptr(), bn.bias.data_ptr(), 0, None, 0)
    else:
        mr = ops.bn_finalize(None, B * 4 * H * W, c5, bn.eps, bn.momentum, bn.running_mean, bn.running_var,
                             bn.num_batches_tracked, False)
        a, _ = ops.bn_act(x, mr, bn.weight.data_ptr(), bn.bias.data_ptr(), 0, None, 2)
    f3, f5_ = G.final_block[3], G.final_block[5]
    _, a3 = ops.conv(B, 2 * H, 2 * W, c5, [Src(a, st.w("final_block.3"), c5, 3)], bias=f3.bias, want_raw=False,
                     want_act=True)
    oc = f5_.shape[0]
    img = torch.empty((B, oc, 2 * H, 2 * W), dtype=F32, device=x.device)
    call("spyr_conv1x1_tanh_fwd", a3.data_ptr(), f5_.weight_orig.data_ptr(), st.sigma("final_block.5"),
         f5_.bias.data_ptr(), img.data_ptr(), B, 4 * H * W, c5, oc)
    ctx = None
    if save:
        ctx = dict(st=st, cls=cls, z=z, f6=f6, f5=f5, m6=m6, m5=m5, h0=h0, h1=h1, h2=h2, a0=a0, blocks=block_ctx,
                   xf=x, xu=xu, mr=mr, a=a, a3=a3, img=img)
    return img, ctx


def generator_backward(G, ctx, g_img):
    """Returns the flat parameter-gradient arena."""
    sn, ga = G._sn, G._ga
    st, cls = ctx["st"], ctx["cls"]
    dev = g_img.device
    grad = ga.new(dev)
    gw = torch.zeros(sn.gw_floats, dtype=F32, device=dev)
    g_img = _f32c(g_img)
    ops.LEAF.begin(dev)  # weight/bias-gradient kernels from here on overlap the input-gradient chain
    ops.DEFER.begin()    # ... and their fixed-order second stages are batched into one launch at the end
    x, mr, a, a3, img = ctx["xf"], ctx["mr"], ctx["a"], ctx["a3"], ctx["img"]
    B, H, W, c5 = x.shape
    f3, f5_ = G.final_block[3], G.final_block[5]
    bn = G.final_block[1]
    oc = f5_.shape[0]
    g_h3 = ops.act_like(a3)
    sc = ops.scratch(oc * c5 + oc, dev)
    call("spyr_conv1x1_tanh_bwd", g_img.data_ptr(), img.data_ptr(), a3.data_ptr(), f5_.weight_orig.data_ptr(),
         st.sigma("final_block.5"), LRELU, g_h3.data_ptr(), sn.gw_ptr(gw, "final_block.5"), ga.ptr(grad, f5_.bias), B,
         4 * H * W, c5, oc, sc.data_ptr())
    ops.wgrad(a, g_h3, sn.gw_ptr(gw, "final_block.3"), B, 2 * H, 2 * W, c5, c5, 3)
    ops.colsum(g_h3, c5, ga.ptr(grad, f3.bias))
    g_pre, _ = ops.conv(B, 2 * H, 2 * W, c5, [Src(g_h3, st.w("final_block.3"), c5, 3, mn=True)], dmask=a,
                        dmask_slope=LRELU)
    wp, bp = bn.weight.data_ptr(), bn.bias.data_ptr()
    xu = ctx.get("xu")
    # training forward: BN ran on the materialised up2(x) -> plain reductions at 2H x 2W; eval forward: x interpolated
    xs, hs, ws, mode = (xu, 2 * H, 2 * W, 0) if xu is not None else (x, H, W, 3)
    S = ops.bn_bwd_partials(B, 4 * H * W, c5, dev)
    call("spyr_bn_bwd_reduce", g_pre.data_ptr(), xs.data_ptr(), mr.data_ptr(), wp, bp, 0, None, LRELU, mode, None,
         S.data_ptr(), B, hs, ws, c5)
    M = torch.empty(2 * c5, dtype=F32, device=dev)
    call("spyr_bn_bwd_finalize", S.data_ptr(), B, c5, float(B * 4 * H * W), wp, 0, None, M.data_ptr(), None, None)
    ops.LEAF.run([S], "spyr_bn_bwd_params", S.data_ptr(), B, c5, 0, None, ga.ptr(grad, bn.weight), ga.ptr(grad, bn.bias))
    g_hi = ops.act_like(g_pre)
    call("spyr_bn_bwd_apply", g_pre.data_ptr(), xs.data_ptr(), mr.data_ptr(), wp, 0, None, M.data_ptr(), None,
         g_hi.data_ptr(), B, hs, ws, c5, 0 if xu is not None else 1)
    g = ops.act_like(x)
    call("spyr_up2_bwd", g_hi.data_ptr(), g.data_ptr(), B, H, W, c5)
    del g_hi, g_pre, g_h3
    for idx in range(len(G.main_path) - 1, -1, -1):
        key = "main_path.%d" % idx
        layer = G.main_path[idx]
        c = ctx["blocks"][idx]
        if idx == 3:
            g = attention_backward(layer, key, st, sn, ga, gw, grad, c, g)
        else:
            g = _gblock_backward(layer, key, st, sn, ga, gw, grad, c, cls, g)
    # head
    cl = G.convolution_layer[1]
    a0, h0, h1, h2 = ctx["a0"], ctx["h0"], ctx["h1"], ctx["h2"]
    c0, c_in = cl.shape[0], cl.shape[1]
    ops.wgrad(a0, g, sn.gw_ptr(gw, "convolution_layer.1"), B, 4, 4, c_in, c0, 1)
    ops.colsum(g, c0, ga.ptr(grad, cl.bias))
    g_a0, _ = ops.conv(B, 4, 4, c_in, [Src(g, st.w("convolution_layer.1"), c0, 1, mn=True)])
    g_h2 = ops.nhwc_to_nchw(g_a0, gate_x=h2.view(B, c_in, 4, 4), slope=LRELU).view(B, -1)
    lb1, lb2 = G.linear_block_1, G.linear_block_2
    m2, f2 = lb2.main_block[1], lb2.masked_feature_mapping
    ops.linear_bwd_w(g_h2, h1, sn.gw_ptr(gw, "linear_block_2.main_block.1"), ga.ptr(grad, m2.bias), in_slope=LRELU)
    ops.linear_bwd_w(g_h2, ctx["f5"], sn.gw_ptr(gw, "linear_block_2.masked_feature_mapping"), ga.ptr(grad, f2.bias),
                     xmask=ctx["m5"])
    g_h1 = ops.linear_bwd_x(g_h2, m2.weight_orig, st.sigma("linear_block_2.main_block.1"), x=h1, in_slope=LRELU)
    m1, f1 = lb1.main_block[1], lb1.masked_feature_mapping
    ops.linear_bwd_w(g_h1, h0, sn.gw_ptr(gw, "linear_block_1.main_block.1"), ga.ptr(grad, m1.bias), in_slope=LRELU)
    ops.linear_bwd_w(g_h1, ctx["f6"], sn.gw_ptr(gw, "linear_block_1.masked_feature_mapping"), ga.ptr(grad, f1.bias),
                     xmask=ctx["m6"])
    g_h0 = ops.linear_bwd_x(g_h1, m1.weight_orig, st.sigma("linear_block_1.main_block.1"), x=h0, in_slope=LRELU)
    ops.linear_bwd_w(g_h0, ctx["z"], sn.gw_ptr(gw, "linear_layer"), ga.ptr(grad, G.linear_layer.bias))
    ops.LEAF.join()
    ops.DEFER.flush()
    sn.backward(st, gw, grad)
    return grad


# ------------------------------------------------------------------------------------------------
# Discriminator (models.py:140-155)
# ------------------------------------------------------------------------------------------------
def _dblock_forward(blk, key, st, x, a, want_act, save):
    B, H, W, Cin = x.shape
    k1, k3, kr = key + ".main_block.1", key + ".main_block.3", key + ".residual_mapping"
    c1, c3, cr = blk.main_block[1], blk.main_block[3], blk.residual_mapping
    Cout = c1.shape[0]
    _, h = ops.conv(B, H, W, Cout, [Src(a, st.w(k1), Cin, 3)], bias=c1.bias, want_raw=False, want_act=True)
    srcs = [Src(h, st.w(k3), Cout, 3), Src(x, st.w(kr), Cin, 1)]
    if ops.can_pool(H, W, Cout):
        # AvgPool2d of models.py:451 in the convolution's epilogue: the full-resolution sum is never written
        out, out_act = ops.conv(B, H, W, Cout, srcs, bias=c3.bias, bias2=cr.bias, want_act=want_act, pool=True)
    else:
        s, _ = ops.conv(B, H, W, Cout, srcs, bias=c3.bias, bias2=cr.bias)
        out, out_act = ops.avgpool2(s, want_act=want_act)
    return out, out_act, ((x, a, h) if save else None)


def _dblock_backward(blk, key, st, sn, ga, gw, grad, ctx, g_out, want_wgrad):
    x, a, h = ctx
    B, H, W, Cin = x.shape
    k1, k3, kr = key + ".main_block.1", key + ".main_block.3", key + ".residual_mapping"
    c1, c3, cr = blk.main_block[1], blk.main_block[3], blk.residual_mapping
    Cout = c1.shape[0]
    g_s = ops.avgpool2_bwd(g_out)
    if want_wgrad:
        ops.wgrad(h, g_s, sn.gw_ptr(gw, k3), B, H, W, Cout, Cout, 3)
        ops.wgrad(x, g_s, sn.gw_ptr(gw, kr), B, H, W, Cin, Cout, 1)
        # sum over the full-resolution gradient = sum over the pooled one (each value is replicated 4 x 0.25)
        ops.colsum(g_out, Cout, ga.ptr(grad, c3.bias), ga.ptr(grad, cr.bias))
    g_h, _ = ops.conv(B, H, W, Cout, [Src(g_s, st.w(k3), Cout, 3, mn=True)], dmask=h, dmask_slope=LRELU)
    if want_wgrad:
        ops.wgrad(a, g_h, sn.gw_ptr(gw, k1), B, H, W, Cin, Cout, 3)
        ops.colsum(g_h, Cout, ga.ptr(grad, c1.bias))
    if ops.can_pool(H, W, Cin):
        # the 1x1 skip path is pointwise, so its input gradient commutes with the pooling backward: run it on the pooled
        # gradient (a quarter of the pixels) and let the main path's epilogue read it as 0.25 * g_skip[h/2][w/2]
        g_skip, _ = ops.conv(B, H // 2, W // 2, Cin, [Src(g_out, st.w(kr), Cout, 1, mn=True)])
        g_x, _ = ops.conv(B, H, W, Cin, [Src(g_h, st.w(k1), Cout, 3, mn=True)], dmask=a, dmask_slope=LRELU,
                          residual=g_skip, residual_pooled=True)
    else:
        g_skip, _ = ops.conv(B, H, W, Cin, [Src(g_s, st.w(kr), Cout, 1, mn=True)])
        g_x, _ = ops.conv(B, H, W, Cin, [Src(g_h, st.w(k1), Cout, 3, mn=True)], dmask=a, dmask_slope=LRELU,
                          residual=g_skip)
    return g_x


def discriminator_forward(D, img, class_id, save):
    training = D.training
    sn = D._sn
    st = sn.forward(training)
    img = _f32c(img)
    B, Ci, H, W = img.shape
    if Ci != 3:
        raise RuntimeError("the B200 discriminator path supports 3-channel images (got %d channels)" % Ci)
    cls = ops.argmax_rows(class_id)
    dev = img.device
    blk0 = D.layers[0]
    c0, c2, cr = blk0.main_block[0], blk0.main_block[2], blk0.residual_mapping
    C0 = c0.shape[0]
    col = ops.act_empty((B, H, W, 32), dev)
    call("spyr_im2col3x3", img.data_ptr(), B, H, W, None, None, col.data_ptr())
    _, h0 = ops.conv(B, H, W, C0, [Src(col, st.w("layers.0.main_block.0"), 32, 1)], bias=c0.bias, want_raw=False,
                     want_act=True)
    xp8 = ops.act_empty((B, H // 2, W // 2, 8), dev)
    call("spyr_img_avgpool_pad8", img.data_ptr(), B, H, W, xp8.data_ptr())
    r, _ = ops.conv(B, H // 2, W // 2, C0, [Src(xp8, st.w("layers.0.residual_mapping"), 8, 1)], bias=cr.bias)
    if ops.can_pool(H, W, C0):
        # avgpool(conv2(...)) + conv1x1(avgpool(x)) (models.py:415-418): pooling and the skip add in conv2's epilogue
        x, a = ops.conv(B, H, W, C0, [Src(h0, st.w("layers.0.main_block.2"), C0, 3)], bias=c2.bias, residual=r,
                        want_act=True, pool=True)
    else:
        s, _ = ops.conv(B, H, W, C0, [Src(h0, st.w("layers.0.main_block.2"), C0, 3)], bias=c2.bias)
        x, a = ops.avgpool2(s, residual=r, want_act=True)
        del s
    del r
    ctxs = [(col, h0, xp8) if save else None]
    for idx in range(1, 8):
        key = "layers.%d" % idx
        layer = D.layers[idx]
        if idx == 3:
            x, a, c = attention_forward(layer, key, st, x, True, save)
        else:
            x, a, c = _dblock_forward(layer, key, st, x, a, idx not in (2, 7), save)
        ctxs.append(c)
    Bx, h, w, C7 = x.shape
    feat0 = torch.empty((B, C7), dtype=F32, device=dev)
    call("spyr_global_avgpool_lrelu_fwd", x.data_ptr(), LRELU, feat0.data_ptr(), B, h * w, C7)
    l11 = D.layers[11]
    feat = ops.linear_fwd(feat0, l11.weight_orig, st.sigma("layers.11"), l11.bias, out_slope=LRELU)
    cls_out = ops.linear_fwd(feat, D.classification.weight_orig, st.sigma("classification"), D.classification.bias)
    E = feat.shape[1]
    out = torch.empty((B, B, E), dtype=F32, device=dev)
    call("spyr_dhead_out_fwd", cls_out.data_ptr(), feat.data_ptr(), D.embedding.weight_orig.data_ptr(),
         st.sigma("embedding"), cls.data_ptr(), out.data_ptr(), B, E)
    ctx = None
    if save:
        ctx = dict(st=st, cls=cls, blocks=ctxs, x7=x, feat0=feat0, feat=feat, shape=(B, H, W))
    return out, ctx


def discriminator_backward(D, ctx, g_out, want_wgrad, want_input_grad):
    """Returns (grad_arena or None, d/dimage NCHW FP32 or None)."""
    sn, ga = D._sn, D._ga
    st, cls = ctx["st"], ctx["cls"]
    B, H, W = ctx["shape"]
    dev = g_out.device
    g_out = _f32c(g_out)
    grad = ga.new(dev)
    gw = torch.zeros(sn.gw_floats, dtype=F32, device=dev)
    if want_wgrad:
        ops.LEAF.begin(dev)
        ops.DEFER.begin()
    feat0, feat, x7 = ctx["feat0"], ctx["feat"], ctx["x7"]
    E = feat.shape[1]
    l11 = D.layers[11]
    g_cls = torch.empty((B, 1), dtype=F32, device=dev)
    g_feat = torch.empty((B, E), dtype=F32, device=dev)
    call("spyr_dhead_out_bwd", g_out.data_ptr(), feat.data_ptr(), D.embedding.weight_orig.data_ptr(), st.sigma("embedding"),
         cls.data_ptr(), g_cls.data_ptr(), g_feat.data_ptr(), sn.gw_ptr(gw, "embedding") if want_wgrad else None, B, E)
    if want_wgrad:
        ops.linear_bwd_w(g_cls, feat, sn.gw_ptr(gw, "classification"), ga.ptr(grad, D.classification.bias))
    ops.linear_bwd_x(g_cls, D.classification.weight_orig, st.sigma("classification"), out=g_feat)
    if want_wgrad:
        ops.linear_bwd_w(g_feat, feat0, sn.gw_ptr(gw, "layers.11"), ga.ptr(grad, l11.bias), y=feat, out_slope=LRELU)
    g_feat0 = ops.linear_bwd_x(g_feat, l11.weight_orig, st.sigma("layers.11"), y=feat, out_slope=LRELU)
    Bx, h, w, C7 = x7.shape
    g = ops.act_like(x7)
    call("spyr_global_avgpool_lrelu_bwd", x7.data_ptr(), g_feat0.data_ptr(), LRELU, g.data_ptr(), B, h * w, C7)
    for idx in range(7, 0, -1):
        key = "layers.%d" % idx
        layer = D.layers[idx]
        c = ctx["blocks"][idx]
        if idx == 3:
            g = attention_backward(layer, key, st, sn, ga, gw, grad, c, g, want_wgrad)
        else:
            g = _dblock_backward(layer, key, st, sn, ga, gw, grad, c, g, want_wgrad)
    # input block
    col, h0, xp8 = ctx["blocks"][0]
    blk0 = D.layers[0]
    c0, c2, cr = blk0.main_block[0], blk0.main_block[2], blk0.residual_mapping
    C0 = c0.shape[0]
    g_s = ops.avgpool2_bwd(g)
    if want_wgrad:
        ops.wgrad(h0, g_s, sn.gw_ptr(gw, "layers.0.main_block.2"), B, H, W, C0, C0, 3)
        # column sums of g_s = those of g (each pooled value replicated 4 x 0.25): one pass serves both biases
        ops.colsum(g, C0, ga.ptr(grad, c2.bias), ga.ptr(grad, cr.bias))
        ops.wgrad(xp8, g, sn.gw_ptr(gw, "layers.0.residual_mapping"), B, H // 2, W // 2, 8, C0, 1, cin_stride=3)
    g_h0, _ = ops.conv(B, H, W, C0, [Src(g_s, st.w("layers.0.main_block.2"), C0, 3, mn=True)], dmask=h0,
                       dmask_slope=LRELU)
    del g_s
    if want_wgrad:
        ops.wgrad(col, g_h0, sn.gw_ptr(gw, "layers.0.main_block.0"), B, H, W, 32, C0, 1, cin_stride=27)
        ops.colsum(g_h0, C0, ga.ptr(grad, c0.bias))
    g_img = None
    if want_input_grad:
        g_col, _ = ops.conv(B, H, W, 32, [Src(g_h0, st.w("layers.0.main_block.0"), C0, 1, mn=True)])
        g_img = torch.empty((B, 3, H, W), dtype=F32, device=dev)
        call("spyr_col2im3x3", g_col.data_ptr(), B, H, W, None, g_img.data_ptr(), 0)
        g8, _ = ops.conv(B, H // 2, W // 2, 8, [Src(g, st.w("layers.0.residual_mapping"), C0, 1, mn=True)])
        call("spyr_img_avgpool_pad8_bwd", g8.data_ptr(), B, H, W, g_img.data_ptr(), 1)
    if want_wgrad:
        ops.LEAF.join()
        ops.DEFER.flush()
        sn.backward(st, gw, grad)
        return grad, g_img
    return None, g_img
