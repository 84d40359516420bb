"""Kernel-level parity (through the C-ABI) against plain PyTorch FP32 ops evaluated on the CPU, on inputs that are exactly
representable in the mode under test.  Every test runs in both precision modes:
  bf16   BF16 operands / one-plane maps: rel-L2 <= 5e-3 (north_star) for BF16-output kernels (one output rounding = 2^-9
         relative, rel-L2 ~1.7e-3), <= 1e-4 for FP32-output kernels;
  split  (hi, lo) BF16 plane pairs, three tensor-core products per convolution: rel-L2 <= 1e-4 (measured ~1e-5).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-4


@pytest.fixture(autouse=True, params=["bf16", "split"])
def precision_mode(request):
    from semantic_pyramid_for_image_generation_b200 import ops as o
    o.set_precision(request.param)
    yield request.param
    o.set_precision("bf16")


def _o():
    from semantic_pyramid_for_image_generation_b200 import ops as o
    return o


def tol16():
    """Tolerance of a kernel whose output is a BF16 map."""
    return 1e-4 if _o().SPLIT else 5e-3


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def q(t):
    """Rounds to what a map of the current mode represents exactly (BF16, or hi + lo)."""
    hi = t.bfloat16().float()
    if _o().SPLIT:
        return hi + (t - hi).bfloat16().float()
    return hi


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def nhwc(t):  # NCHW f32 cpu -> NHWC map on the GPU (one BF16 plane, or hi + lo)
    return _o().to_act(t.permute(0, 2, 3, 1).contiguous().cuda())


def nchw(t):  # NHWC map on the GPU -> NCHW f32 cpu
    return _o().act_value(t).cpu().permute(0, 3, 1, 2).contiguous()


def pack(w):  # (Cout,Cin,k,k) -> [taps][Cout][Cin] operand on the GPU
    co, ci, k, _ = w.shape
    return _o().to_act(w.permute(2, 3, 0, 1).reshape(k * k, co, ci).contiguous().cuda())


def amap(*shape):  # uninitialised map of the current mode
    return _o().act_empty(shape, "cuda")


@pytest.fixture(scope="module")
def ops():
    from semantic_pyramid_for_image_generation_b200 import ops as o
    return o


@pytest.mark.parametrize("shape", [(2, 64, 64, 16, 3), (3, 128, 256, 8, 3), (2, 256, 64, 32, 1), (5, 512, 512, 4, 3),
                                   (2, 32, 16, 16, 1)])
def test_conv_fprop_dgrad_wgrad(ops, shape):
    B, Cin, Cout, H, k = shape
    g = gen(1)
    x = q(torch.randn(B, Cin, H, H, generator=g))
    w = q(torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k))
    b = torch.randn(Cout, generator=g)
    dy = q(torch.randn(B, Cout, H, H, generator=g))
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    y_ref = F.conv2d(xr, wr, b, padding=k // 2)
    y_ref.backward(dy)
    xc, wc, dyc = nhwc(x), pack(w), nhwc(dy)
    y, ya = ops.conv(B, H, H, Cout, [ops.Src(xc, wc, Cin, k)], bias=b.cuda(), want_act=True)
    assert rel_l2(nchw(y), y_ref) < tol16()
    assert rel_l2(nchw(ya), F.leaky_relu(y_ref, 0.2)) < tol16()
    gx, _ = ops.conv(B, H, H, Cin, [ops.Src(dyc, wc, Cout, k, mn=True)])
    assert rel_l2(nchw(gx), xr.grad) < tol16()
    dw = torch.zeros(k * k, Cin, Cout, device="cuda")
    ops.wgrad(xc, dyc, dw.data_ptr(), B, H, H, Cin, Cout, k)
    dw_ref = wr.grad.permute(2, 3, 1, 0).reshape(k * k, Cin, Cout)
    assert rel_l2(dw, dw_ref) < TOL_F32
    db = torch.zeros(Cout, device="cuda")
    ops.colsum(dyc, Cout, db.data_ptr())
    assert rel_l2(db, dy.sum(dim=(0, 2, 3))) < TOL_F32


def test_conv_fused_sources_stencil_gate_residual(ops):
    """out = conv3(a) + conv1(xu) + conv3(cat(f*m, m)) as ONE launch (models.py:333-338) and the gated dgrad."""
    B, C, Cf, H = 2, 64, 128, 16
    g = gen(2)
    a = q(torch.randn(B, C, H, H, generator=g))
    xu = q(torch.randn(B, C, H, H, generator=g))
    f = q(torch.randn(B, Cf, H, H, generator=g))
    m = (torch.rand(B, 1, H, H, generator=g) > 0.5).float()
    m[0] = 1.0
    w6 = q(torch.randn(C, C, 3, 3, generator=g) * 0.05)
    wr = q(torch.randn(C, C, 1, 1, generator=g) * 0.1)
    wf = torch.randn(C, Cf + 1, 3, 3, generator=g) * 0.05
    wf[:, :Cf] = q(wf[:, :Cf])
    b1, b2, b3 = (torch.randn(C, generator=g) for _ in range(3))
    ref = F.conv2d(a, w6, b1, padding=1) + F.conv2d(xu, wr, b2) + F.conv2d(torch.cat((f * m, m), 1), wf, b3, padding=1)
    st = torch.zeros(10, C)
    st[:9] = wf[:, Cf].reshape(C, 9).t()
    st[9] = st[:9].sum(0)
    from semantic_pyramid_for_image_generation_b200 import ops as o
    fm = o.maskgate(nhwc(f), m.cuda())
    y, _ = ops.conv(B, H, H, C, [ops.Src(nhwc(a), pack(w6), C, 3), ops.Src(nhwc(xu), pack(wr), C, 1),
                                 ops.Src(fm, pack(wf[:, :Cf].contiguous()), Cf, 3)],
                    bias=b1.cuda(), bias2=b2.cuda(), bias3=b3.cuda(), stencil_mask=m.cuda(), stencil_w=st.cuda())
    assert rel_l2(nchw(y), ref) < tol16()
    # mask-channel weight gradient
    dy = q(torch.randn(B, C, H, H, generator=g))
    dw = torch.zeros(9, Cf + 1, C, device="cuda")
    o.stencil_wgrad(m.cuda(), nhwc(dy), B, H, H, C, dw.data_ptr(), Cf + 1, Cf)
    mr = m.clone().requires_grad_(False)
    wm = torch.zeros(C, 1, 3, 3, requires_grad=True)
    (F.conv2d(mr, wm, padding=1) * dy).sum().backward()
    assert rel_l2(dw[:, Cf, :], wm.grad.reshape(C, 9).t()) < TOL_F32
    # gated input gradient + residual (LeakyReLU backward fused in the epilogue)
    act = q(torch.randn(B, C, H, H, generator=g))
    res = q(torch.randn(B, C, H, H, generator=g))
    gx, _ = ops.conv(B, H, H, C, [ops.Src(nhwc(dy), pack(w6), C, 3, mn=True)], dmask=nhwc(act), dmask_slope=0.2,
                     residual=nhwc(res))
    gref = F.conv_transpose2d(dy, w6, padding=1) * torch.where(act > 0, 1.0, 0.2) + res
    assert rel_l2(nchw(gx), gref) < tol16()


def test_first_layer_im2col_path(ops):
    """3-channel 3x3 conv as im2col + 1x1 tensor-core conv; its weight and image gradients (models.py:195-202,393)."""
    from semantic_pyramid_for_image_generation_b200 import ops as o
    B, H, C = 2, 32, 64
    g = gen(3)
    img = torch.rand(B, 3, H, H, generator=g) * 2 - 1
    mean = torch.tensor([0.485, 0.456, 0.406])
    std = torch.tensor([0.229, 0.224, 0.225])
    w = q(torch.randn(C, 3, 3, 3, generator=g) * 0.2)
    b = torch.randn(C, generator=g)
    xn = ((img - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1)).requires_grad_(True)
    col = amap(B, H, H, 32)
    o.call("spyr_im2col3x3", img.cuda(), B, H, H, mean.cuda(), (1 / std).cuda(),
           col.data_ptr())
    ref_col = F.unfold(q(xn.detach()), 3, padding=1).view(B, 3, 9, H, H).permute(0, 3, 4, 2, 1).reshape(B, H, H, 27)
    colv = o.act_value(col)
    assert rel_l2(colv[..., :27], ref_col) < 1e-6 and float(colv[..., 27:].abs().max()) == 0.0
    wp = torch.zeros(C, 32)
    wp[:, :27] = w.permute(0, 2, 3, 1).reshape(C, 27)
    wp = o.to_act(wp.cuda())
    y, _ = ops.conv(B, H, H, C, [ops.Src(col, wp, 32, 1)], bias=b.cuda())
    y_ref = F.conv2d(q(xn), w, b, padding=1)
    assert rel_l2(nchw(y), y_ref) < tol16()
    dy = q(torch.randn(B, C, H, H, generator=g))
    xr = q(xn.detach()).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    F.conv2d(xr, wr, b, padding=1).backward(dy)
    dw = torch.zeros(32, C, device="cuda")
    ops.wgrad(col, nhwc(dy), dw.data_ptr(), B, H, H, 32, C, 1)
    assert rel_l2(dw[:27], wr.grad.permute(2, 3, 1, 0).reshape(27, C)) < TOL_F32
    gcol, _ = ops.conv(B, H, H, 32, [ops.Src(nhwc(dy), wp, C, 1, mn=True)])
    gimg = torch.empty(B, 3, H, H, device="cuda")
    o.call("spyr_col2im3x3", gcol.data_ptr(), B, H, H, (1 / std).cuda(), gimg.data_ptr(), 0)
    assert rel_l2(gimg, xr.grad / std.view(1, 3, 1, 1)) < tol16()
    # skip path: avgpool(img) padded to 8 channels and its transpose
    p8 = amap(B, H // 2, H // 2, 8)
    o.call("spyr_img_avgpool_pad8", img.cuda(), B, H, H, p8.data_ptr())
    assert rel_l2(nchw(p8)[:, :3], F.avg_pool2d(img, 2)) < tol16() and float(o.act_value(p8)[..., 3:].abs().max()) == 0.0
    g8 = q(torch.randn(B, H // 2, H // 2, 8, generator=g))
    acc = torch.ones(B, 3, H, H, device="cuda")
    o.call("spyr_img_avgpool_pad8_bwd", o.to_act(g8.cuda()), B, H, H, acc.data_ptr(), 1)
    ref = 1.0 + 0.25 * F.interpolate(g8.permute(0, 3, 1, 2)[:, :3], scale_factor=2, mode="nearest")
    assert rel_l2(acc, ref) < 1e-6


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_batchnorm_act_upsample_forward_backward(ops, mode):
    """CBN/BN + LeakyReLU (+ bilinear x2, align_corners=True) against torch autograd (models.py:295-298,491-506,51-54)."""
    from semantic_pyramid_for_image_generation_b200 import ops as o
    B, C, H, ncls = 3, 32, 8, 5
    g = gen(4 + mode)
    x = q(torch.randn(B, C, H, H, generator=g) * 1.5 + 0.3)
    emb = torch.randn(ncls, 2 * C, generator=g)
    cls = torch.tensor([1, 4, 1])
    rm, rv = torch.zeros(C), torch.ones(C)
    xr = x.clone().requires_grad_(True)
    er = emb.clone().requires_grad_(True)
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    rows = er[cls]
    sc, sh = rows[:, :C, None, None], rows[:, C:, None, None]
    if mode == 2:
        y = F.batch_norm(up(xr), rm, rv, None, None, True, 0.1, 1e-5)
        a_ref = F.leaky_relu(sc * y + sh, 0.2)
    else:
        y = F.batch_norm(xr, rm, rv, None, None, True, 0.1, 1e-5)
        a_ref = F.leaky_relu(sc * y + sh, 0.2)
        if mode == 1:
            a_ref = up(a_ref)
    xc = nhwc(x)
    embc = emb.cuda()
    clsc = cls.int().cuda()
    rmc, rvc = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    nbt = torch.zeros((), dtype=torch.long, device="cuda")
    sums = o.bn_stats(xc, up2=(mode == 2))
    cnt = B * H * H * (4 if mode == 2 else 1)
    mr = o.bn_finalize(sums, cnt, C, 1e-5, 0.1, rmc, rvc, nbt, True)
    a, xu = o.bn_act(xc, mr, embc.data_ptr(), embc.data_ptr() + 4 * C, 2 * C, clsc, mode, want_xu=(mode == 1))
    assert rel_l2(nchw(a), a_ref) < tol16()
    assert rel_l2(rmc, rm) < 1e-4 and rel_l2(rvc, rv) < 1e-4 and int(nbt) == 1
    if mode == 1:
        assert rel_l2(nchw(xu), up(x)) < tol16()
    # backward
    ga = q(torch.randn(a_ref.shape, generator=g))
    a_ref.backward(ga)
    S = o.bn_bwd_partials(B, H * H * (4 if mode == 2 else 1), C, "cuda")
    M = torch.empty(2 * C, device="cuda")
    demb = torch.zeros(ncls, 2 * C, device="cuda")
    sp, hp = embc.data_ptr(), embc.data_ptr() + 4 * C
    if mode == 0:
        gy = nhwc(ga * torch.where(a_ref.detach() > 0, 1.0, 0.2))  # the conv epilogue applies this gate
        o.call("spyr_bn_bwd_reduce", gy.data_ptr(), xc.data_ptr(), mr.data_ptr(), sp, hp, 2 * C, clsc.data_ptr(), 0.2, 0,
               None, S.data_ptr(), B, H, H, C)
        gsrc, up_flag = gy, 0
    elif mode == 1:
        gy = o.act_like(xc)
        o.call("spyr_bn_bwd_reduce", nhwc(ga), xc.data_ptr(), mr.data_ptr(), sp, hp, 2 * C, clsc.data_ptr(), 0.2,
               1, gy.data_ptr(), S.data_ptr(), B, H, H, C)
        gsrc, up_flag = gy, 0
    else:
        gy = nhwc(ga * torch.where(a_ref.detach() > 0, 1.0, 0.2))
        o.call("spyr_bn_bwd_reduce", gy.data_ptr(), xc.data_ptr(), mr.data_ptr(), sp, hp, 2 * C, clsc.data_ptr(), 0.2, 3,
               None, S.data_ptr(), B, H, H, C)
        gsrc, up_flag = gy, 1
    o.call("spyr_bn_bwd_finalize", S.data_ptr(), B, C, float(cnt), sp, 2 * C, clsc.data_ptr(), M.data_ptr(),
           demb.data_ptr(), demb.data_ptr() + 4 * C)
    gx = o.act_like(gsrc)
    o.call("spyr_bn_bwd_apply", gsrc.data_ptr(), xc.data_ptr(), mr.data_ptr(), sp, 2 * C, clsc.data_ptr(), M.data_ptr(), None,
           gx.data_ptr(), B, H, H, C, up_flag)
    if mode == 2:
        lo = o.act_like(xc)
        o.call("spyr_up2_bwd", gx.data_ptr(), lo.data_ptr(), B, H, H, C)
        gx = lo
    # bf16 mode: gy is re-read as BF16 by the second pass, two roundings on the path
    assert rel_l2(nchw(gx), xr.grad) < (1e-4 if o.SPLIT else 8e-3)
    assert rel_l2(demb, er.grad) < (1e-4 if o.SPLIT else 5e-3)


def test_pooling_kernels(ops):
    from semantic_pyramid_for_image_generation_b200 import ops as o
    B, C, H = 2, 32, 16
    g = gen(8)
    x = q(torch.randn(B, C, H, H, generator=g))
    r = q(torch.randn(B, C, H // 2, H // 2, generator=g))
    y, ya = o.avgpool2(nhwc(x), residual=nhwc(r), want_act=True)
    ref = F.avg_pool2d(x, 2) + r
    assert rel_l2(nchw(y), ref) < tol16() and rel_l2(nchw(ya), F.leaky_relu(q(ref), 0.2)) < tol16()
    gl = q(torch.randn(B, C, H // 2, H // 2, generator=g))
    assert rel_l2(nchw(o.avgpool2_bwd(nhwc(gl))), 0.25 * F.interpolate(gl, scale_factor=2, mode="nearest")) < tol16()
    xr = F.relu(x).requires_grad_(True)
    p = F.max_pool2d(xr, 2)
    p.backward(gl)
    assert rel_l2(nchw(o.maxpool2(nhwc(F.relu(x)))), p) == 0.0
    gx = o.maxpool2_bwd(nhwc(F.relu(x)), nhwc(gl), True)
    assert rel_l2(nchw(gx), xr.grad * (xr.detach() > 0)) < 1e-6
    # adaptive 8x8 -> 7x7 and its transpose
    x8 = q(torch.randn(B, C, 8, 8, generator=g)).requires_grad_(True)
    a7 = F.adaptive_avg_pool2d(x8, (7, 7))
    g7 = q(torch.randn(B, C, 7, 7, generator=g))
    a7.backward(g7)
    y7 = amap(B, 7, 7, C)
    o.call("spyr_adaptive_avgpool_fwd", nhwc(x8.detach()), y7.data_ptr(), B, 8, 8, 7, 7, C)
    assert rel_l2(nchw(y7), a7) < tol16()
    g8 = amap(B, 8, 8, C)
    o.call("spyr_adaptive_avgpool_bwd", nhwc(g7), None, g8.data_ptr(), B, 8, 8, 7, 7, C)
    assert rel_l2(nchw(g8), x8.grad) < tol16()
    # bilinear transpose
    gh = q(torch.randn(B, C, 2 * H, 2 * H, generator=g))
    xl = x.clone().requires_grad_(True)
    F.interpolate(xl, scale_factor=2, mode="bilinear", align_corners=True).backward(gh)
    lo = amap(B, H, H, C)
    o.call("spyr_up2_bwd", nhwc(gh), lo.data_ptr(), B, H, H, C)
    assert rel_l2(nchw(lo), xl.grad) < tol16()


def test_spectral_norm_forward_backward(ops):
    """Power iteration, sigma, BF16 pack and the backward through sigma against torch.nn.utils.spectral_norm."""
    from semantic_pyramid_for_image_generation_b200.spectral import LayerSpec, SNSet, SpectralNormHolder
    from semantic_pyramid_for_image_generation_b200.engine import GradArena
    import torch.nn as nn
    torch.manual_seed(0)

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.a = SpectralNormHolder(96, 40, 3, 3)
            self.b = SpectralNormHolder(70, 300)
            self.c = SpectralNormHolder(64, 33, 3, 3)
            self.d = SpectralNormHolder(365, 128, bias=False, init="normal")

    m = M().cuda()
    ga = GradArena(m)
    sn = SNSet([LayerSpec("a", m.a, pack_cin=40), LayerSpec("b", m.b), LayerSpec("c", m.c, pack_cin=32, stencil=True),
                LayerSpec("d", m.d)], ga.offset)
    refs = {}
    for key in "abcd":
        h = getattr(m, key)
        w = h.weight_orig.detach().cpu().clone().requires_grad_(True)
        u, v = h.weight_u.cpu().clone(), h.weight_v.cpu().clone()
        mat = w.reshape(w.shape[0], -1)
        with torch.no_grad():
            v = F.normalize(torch.mv(mat.t(), u), dim=0, eps=1e-12)
            u = F.normalize(torch.mv(mat, v), dim=0, eps=1e-12)
        sigma = torch.dot(u, torch.mv(mat, v))
        refs[key] = (w, u, v, sigma)
    st = sn.forward(True)
    for key in "abcd":
        w, u, v, sigma = refs[key]
        h = getattr(m, key)
        assert rel_l2(h.weight_u, u) < 1e-5 and rel_l2(h.weight_v, v) < 1e-5
        off = sn.by_key[key].saved_off
        assert abs(float(st.saved[off]) - float(sigma)) < 1e-5 * float(sigma)
    wa = refs["a"][0].detach() / refs["a"][3]
    def packed(key, n):  # value of a layer's packed operand: hi plane (+ lo plane right behind it in split mode)
        off = sn.by_key[key].pack_off
        v = st.packed[off:off + n].float()
        if _o().SPLIT:
            v = v + st.packed[off + n:off + 2 * n].float()
        return v

    pa = packed("a", 9 * 96 * 40).view(9, 96, 40)
    assert rel_l2(pa, wa.permute(2, 3, 0, 1).reshape(9, 96, 40)) < tol16()
    wc = refs["c"][0].detach() / refs["c"][3]
    pc = packed("c", 9 * 64 * 32).view(9, 64, 32)
    assert rel_l2(pc, wc[:, :32].permute(2, 3, 0, 1).reshape(9, 64, 32)) < tol16()
    sc = st.stencil[sn.by_key["c"].stencil_off:sn.by_key["c"].stencil_off + 640].view(10, 64)
    assert rel_l2(sc[:9], wc[:, 32].reshape(64, 9).t()) < 1e-5 and rel_l2(sc[9], wc[:, 32].sum(dim=(1, 2))) < 1e-5
    # backward: G random, layouts as produced by the wgrad kernel (conv) / linear kernels
    g = gen(9)
    gw = torch.zeros(sn.gw_floats, device="cuda")
    grad = ga.new("cuda")
    for key in "abcd":
        w, u, v, sigma = refs[key]
        G = torch.randn(w.shape, generator=g)
        ((w / sigma) * G).sum().backward()
        spec = sn.by_key[key]
        if spec.gw_layout == 1:
            flat = G.permute(2, 3, 1, 0).reshape(-1)
        else:
            flat = G.reshape(-1)
        gw[spec.gw_off:spec.gw_off + flat.numel()] = flat.cuda()
    sn.backward(st, gw, grad)
    for key in "abcd":
        h = getattr(m, key)
        o_ = ga.offset(h.weight_orig)
        mine = grad[o_:o_ + h.weight_orig.numel()].view(h.weight_orig.shape)
        assert rel_l2(mine, refs[key][0].grad) < 1e-4, key
    # eval mode: no iteration, sigma from the stored vectors
    u_before = m.a.weight_u.clone()
    st2 = sn.forward(False)
    assert torch.equal(u_before, m.a.weight_u)
    w, u, v, _ = refs["a"]
    sig_eval = torch.dot(u, torch.mv(w.detach().reshape(96, -1), v))
    assert abs(float(st2.saved[sn.by_key["a"].saved_off]) - float(sig_eval)) < 1e-5 * float(sig_eval)


def test_linear_and_head_kernels(ops):
    from semantic_pyramid_for_image_generation_b200 import ops as o
    g = gen(10)
    B, K, O_ = 5, 365, 300
    x = torch.randn(B, K, generator=g)
    m = (torch.rand(B, K, generator=g) > 0.3).float()
    w = torch.randn(O_, K, generator=g) * 0.05
    b = torch.randn(O_, generator=g)
    add = torch.randn(B, O_, generator=g)
    sigma = torch.tensor([1.7])
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y_ref = F.leaky_relu(F.linear(F.leaky_relu(xr * m, 0.2), wr / sigma, b) + add, 0.2)
    gy = torch.randn(B, O_, generator=g)
    y_ref.backward(gy)
    sg = sigma.cuda()
    y = o.linear_fwd(x.cuda(), w.cuda(), sg.data_ptr(), b.cuda(), xmask=m.cuda(), in_slope=0.2, y_add=add.cuda(),
                     out_slope=0.2)
    assert rel_l2(y, y_ref) < 1e-5
    gx = o.linear_bwd_x(gy.cuda(), w.cuda(), sg.data_ptr(), y=y, out_slope=0.2, x=(x * m).cuda(), in_slope=0.2)
    assert rel_l2(gx * m.cuda(), xr.grad) < 1e-5
    gw = torch.zeros(O_, K, device="cuda")
    gb = torch.zeros(O_, device="cuda")
    o.linear_bwd_w(gy.cuda(), x.cuda(), gw.data_ptr(), gb.data_ptr(), y=y, out_slope=0.2, xmask=m.cuda(), in_slope=0.2)
    assert rel_l2(gw / sigma.cuda(), wr.grad) < 1e-5
    # (B,B,E) discriminator output
    E = 128
    cls = torch.randn(B, 1, generator=g)
    feat = torch.randn(B, E, generator=g)
    emb = torch.randn(365, E, generator=g)
    idx = torch.tensor([3, 7, 3, 100, 364])
    fr, cr, er = feat.clone().requires_grad_(True), cls.clone().requires_grad_(True), emb.clone().requires_grad_(True)
    out_ref = cr + fr * (er / sigma)[idx].unsqueeze(1)
    go = torch.randn(B, B, E, generator=g)
    out_ref.backward(go)
    out = torch.empty(B, B, E, device="cuda")
    o.call("spyr_dhead_out_fwd", cls.cuda(), feat.cuda(), emb.cuda(), sg.data_ptr(),
           idx.int().cuda(), out.data_ptr(), B, E)
    assert rel_l2(out, out_ref) < 1e-6
    gc, gf, ge = torch.empty(B, 1, device="cuda"), torch.empty(B, E, device="cuda"), torch.zeros(365, E, device="cuda")
    o.call("spyr_dhead_out_bwd", go.cuda(), feat.cuda(), emb.cuda(), sg.data_ptr(),
           idx.int().cuda(), gc.data_ptr(), gf.data_ptr(), ge.data_ptr(), B, E)
    assert rel_l2(gc, cr.grad) < 1e-5 and rel_l2(gf, fr.grad) < 1e-5 and rel_l2(ge / sigma.cuda(), er.grad) < 1e-5


def test_generator_tail_and_adam(ops):
    from semantic_pyramid_for_image_generation_b200 import ops as o
    from semantic_pyramid_for_image_generation_b200.optim import FusedAdam
    g = gen(12)
    B, C, H = 2, 64, 16
    a = q(F.leaky_relu(torch.randn(B, C, H, H, generator=g), 0.2))
    w = torch.randn(3, C, generator=g) * 0.2
    b = torch.randn(3, generator=g) * 0.1
    sigma = torch.tensor([1.3])
    hr = a.clone().requires_grad_(True)  # stands for the post-LeakyReLU activation
    wr = w.clone().requires_grad_(True)
    img_ref = torch.tanh(F.conv2d(hr, (wr / sigma).view(3, C, 1, 1), b))
    gi = torch.randn(B, 3, H, H, generator=g)
    img_ref.backward(gi)
    img = torch.empty(B, 3, H, H, device="cuda")
    ac = nhwc(a)
    o.call("spyr_conv1x1_tanh_fwd", ac.data_ptr(), w.cuda(), sigma.cuda(), b.cuda(),
           img.data_ptr(), B, H * H, C, 3)
    assert rel_l2(img, img_ref) < 1e-5
    gh = o.act_like(ac)
    dw, db = torch.zeros(3, C, device="cuda"), torch.zeros(3, device="cuda")
    o.call("spyr_conv1x1_tanh_bwd", gi.cuda(), img.data_ptr(), ac.data_ptr(), w.cuda(),
           sigma.cuda(), 0.2, gh.data_ptr(), dw.data_ptr(), db.data_ptr(), B, H * H, C, 3,
           o.scratch(3 * C + 3, "cuda"))
    assert rel_l2(nchw(gh), hr.grad * torch.where(a > 0, 1.0, 0.2)) < tol16()
    assert rel_l2(dw / sigma.cuda(), wr.grad) < 1e-4
    # Adam against torch.optim.Adam for three steps
    ps = [torch.randn(n, generator=g) for n in (5, 1000, 70001)]
    mine = [p.clone().cuda().requires_grad_(True) for p in ps]
    ref = [p.clone().requires_grad_(True) for p in ps]
    om, orf = FusedAdam(mine, lr=1e-3), torch.optim.Adam(ref, lr=1e-3)
    for step in range(3):
        for pm, pr in zip(mine, ref):
            gr = torch.randn(pr.shape, generator=g)
            pm.grad, pr.grad = gr.cuda(), gr.clone()
        om.step()
        orf.step()
    for pm, pr in zip(mine, ref):
        assert rel_l2(pm, pr) < 1e-6
    assert float(om.state_dict()["state"][0]["step"]) == 3.0


def test_attention_forward_backward(ops):
    """SelfAttention (models.py:249-275) through the per-image tensor-core GEMMs against torch autograd."""
    from semantic_pyramid_for_image_generation_b200 import models, engine
    from oracle import spyramid_oracle as O
    import torch.nn as nn

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.att = models.SelfAttention(64)

    torch.manual_seed(3)
    net = Net().cuda()
    net.att.gamma.data.fill_(0.7)
    from semantic_pyramid_for_image_generation_b200.spectral import LayerSpec, SNSet
    ga = engine.GradArena(net)
    specs = [LayerSpec("att." + n, getattr(net.att, n), pack_cin=getattr(net.att, n).shape[1])
             for n in ("query_convolution", "key_convolution", "value_convolution", "attention_convolution")]
    sn = SNSet(specs, ga.offset)
    sd = {k[len("att."):] if k.startswith("att.") else k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    sd = {"a." + k: v for k, v in sd.items()}
    g = gen(13)
    B, H = 2, 16
    x = q(torch.randn(B, 64, H, H, generator=g))
    for k in list(sd):
        if k.endswith(("weight_orig", "bias", "gamma")):
            sd[k].requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    out_ref = O.self_attention(sd, "a", xr, training=True)
    go = q(torch.randn(out_ref.shape, generator=g))
    out_ref.backward(go)
    st = sn.forward(True)
    xc = nhwc(x)
    out, _, ctx = engine.attention_forward(net.att, "att", st, xc, False, True)
    assert rel_l2(nchw(out), out_ref) < tol16()
    grad = ga.new("cuda")
    gw = torch.zeros(sn.gw_floats, device="cuda")
    gx = engine.attention_backward(net.att, "att", st, sn, ga, gw, grad, ctx, nhwc(go))
    sn.backward(st, gw, grad)
    assert rel_l2(nchw(gx), xr.grad) < (2e-4 if _o().SPLIT else 1e-2)
    for name, p in net.named_parameters():
        ref = sd["a." + name[len("att."):]].grad
        o_ = ga.offset(p)
        mine = grad[o_:o_ + p.numel()].view(p.shape)
        if float(ref.norm()) < 1e-6 * float(sd["a.value_convolution.bias"].grad.norm()):
            continue  # key bias: softmax shift invariance makes this gradient analytically zero
        assert rel_l2(mine, ref) < (5e-4 if _o().SPLIT else 2e-2), name


@pytest.mark.parametrize("shape", [(3, 32, 32, 32, 128), (2, 16, 32, 16, 64), (2, 16, 16, 8, 64)])
def test_fused_sagan_attention_forward(ops, shape):
    """One-kernel attention (S = QK^T, softmax over keys, O = PV) against torch on BF16-representable q, k, v."""
    from semantic_pyramid_for_image_generation_b200 import ops as o
    if o.SPLIT:
        with pytest.raises(RuntimeError, match="split-BF16"):  # single-plane kernel: refuses instead of losing the lo planes
            o.call("spyr_sagan_attention_fwd", 0, 0, 0, 0, 0, 1, 128, 32, 64, 64)
        return
    B, H, W, d, dv = shape
    hw, nk = H * W, (H * W) // 4
    g = gen(17)
    qq = q(torch.randn(B, hw, d, generator=g))
    kk = q(torch.randn(B, nk, d, generator=g) * 0.7)
    vv = q(torch.randn(B, nk, dv, generator=g))
    p_ref = torch.softmax(torch.einsum("bqd,bkd->bqk", qq, kk), dim=-1)
    o_ref = torch.einsum("bqk,bkc->bqc", p_ref, vv)
    out = torch.empty(B, hw, dv, dtype=torch.bfloat16, device="cuda")
    pm = torch.empty(B, hw, nk, dtype=torch.bfloat16, device="cuda")
    o.call("spyr_sagan_attention_fwd", qq.bfloat16().cuda(), kk.bfloat16().cuda(), vv.bfloat16().cuda(), out, pm, B, hw, d,
           nk, dv)
    assert rel_l2(pm, p_ref) < tol16()
    assert rel_l2(out, o_ref) < tol16()
    assert float((pm.float().sum(-1) - 1).abs().max()) < 2e-2


def test_arena_add_and_weight_relayout(ops):
    """spyr_add_inplace (second D backward pass into the first arena) and spyr_weight_transpose_flip (K-major operand of a
    64-wide input gradient) are exact data movement / FP32 adds."""
    from semantic_pyramid_for_image_generation_b200._native import call
    a = torch.randn(1 << 18, generator=gen(1)).cuda()
    b = torch.randn(1 << 18, generator=gen(2)).cuda()
    want = a + b
    call("spyr_add_inplace", a.data_ptr(), b.data_ptr(), a.numel())
    assert torch.equal(a, want)
    wf = q(torch.randn(9, 128, 64, generator=gen(3)))
    w = _o().to_act(wf.cuda())  # forward pack [tap][Cout=128][Cin=64]
    out = amap(9, 64, 128)
    call("spyr_weight_transpose_flip", w.data_ptr(), out.data_ptr(), 9, 128, 64)
    assert torch.equal(_o().act_value(out).cpu(), wf.flip(0).transpose(1, 2).contiguous())


def test_input_gradient_of_64_wide_layer_uses_relaid_weights(ops):
    """ops.conv routes a 64-wide 3x3 input gradient through the K-major re-lay (CTA-pair kernel); same numbers as the
    MN-major route of the other widths, checked against torch's conv_transpose."""
    B, H, W, cin_f, cout_f = 2, 32, 32, 64, 128  # forward conv 64 -> 128; its input gradient has 64 channels
    g = q(torch.randn(B, cout_f, H, W, generator=gen(4)))
    w = q(torch.randn(cout_f, cin_f, 3, 3, generator=gen(5)) * 0.05)
    want = F.conv_transpose2d(g, w, padding=1)
    got, _ = ops.conv(B, H, W, cin_f, [ops.Src(nhwc(g), pack(w), cout_f, 3, mn=True)])
    assert rel_l2(nchw(got), want) < tol16()


def test_up2_stats_materialises_the_upsampled_map(ops):
    B, H, W, C = 2, 16, 16, 64
    x = q(torch.randn(B, C, H, W, generator=gen(6)))
    xu, sums = ops.up2_stats(nhwc(x))
    want = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    assert tuple(xu.shape) == (B, 2 * H, 2 * W, C)
    assert rel_l2(nchw(xu), want) < tol16()
    xs = nchw(xu).double()  # the statistics are those of the stored values
    sums = ops.bn_sums(sums, B * 4 * H * W, C)
    assert torch.allclose(sums[:C].cpu(), xs.sum(dim=(0, 2, 3)), rtol=1e-6, atol=1e-6)
    assert torch.allclose(sums[C:].cpu(), (xs * xs).sum(dim=(0, 2, 3)), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 64, False), (2, 32, 64, 64, 128, True), (1, 16, 16, 128, 256, True)])
def test_conv_with_fused_average_pool(ops, shape):
    """pool=True: avgpool2(conv(x) + bias) (+ residual at the pooled resolution), raw and LeakyReLU outputs, on both the
    single-CTA (Cout 64, MN... K-major) and the CTA-pair kernels (models.py:406,415-418,451,465)."""
    B, H, W, cin, cout, with_res = shape
    if ops.SPLIT:
        assert not ops.can_pool(H, W, cout)  # the split-BF16 mode keeps the separate pooling kernel
        return
    assert ops.can_pool(H, W, cout)
    x = q(torch.randn(B, cin, H, W, generator=gen(7)))
    w = q(torch.randn(cout, cin, 3, 3, generator=gen(8)) * 0.05)
    b = torch.randn(cout, generator=gen(9)) * 0.1
    res = q(torch.randn(B, cout, H // 2, W // 2, generator=gen(10))) if with_res else None
    want = F.avg_pool2d(F.conv2d(x, w, b, padding=1), 2)
    if with_res:
        want = want + res
    raw, act = ops.conv(B, H, W, cout, [ops.Src(nhwc(x), pack(w), cin, 3)], bias=b.cuda(),
                        residual=nhwc(res) if with_res else None, want_raw=True, want_act=True, pool=True)
    assert tuple(raw.shape) == (B, H // 2, W // 2, cout)
    assert rel_l2(nchw(raw), want) < tol16()
    assert rel_l2(nchw(act), F.leaky_relu(want, 0.2)) < tol16()


def test_conv_with_pooled_residual(ops):
    """residual_pooled=True: out = gate(conv(x)) + 0.25 * upsample_nearest(residual) -- the average-pooled skip branch's
    gradient joining the main path of a discriminator block without its full-resolution copy."""
    B, H, W, cin, cout = 2, 32, 32, 128, 64
    if ops.SPLIT:
        return  # fused only in the single-plane mode (engine falls back to the full-resolution skip gradient)
    x = q(torch.randn(B, cin, H, W, generator=gen(11)))
    w = q(torch.randn(cout, cin, 3, 3, generator=gen(12)) * 0.05)
    gate = q(torch.randn(B, cout, H, W, generator=gen(13)))
    res = q(torch.randn(B, cout, H // 2, W // 2, generator=gen(14)))
    want = F.conv2d(x, w, None, padding=1)
    want = torch.where(gate > 0, want, 0.2 * want) + 0.25 * F.interpolate(res, scale_factor=2, mode="nearest")
    got, _ = ops.conv(B, H, W, cout, [ops.Src(nhwc(x), pack(w), cin, 3)], dmask=nhwc(gate), dmask_slope=0.2,
                      residual=nhwc(res), residual_pooled=True)
    assert rel_l2(nchw(got), want) < tol16()


@pytest.mark.parametrize("shape", [(2, 64, 128, 128), (1, 128, 256, 256), (1, 32, 8, 128)])
def test_conv_64_channel_rows_stacked(ops, shape):
    """Cout = 64, 3x3, maps >= 128 wide run on conv_stack3.cu (the three taps of a kernel row stacked along N, outputs
    combined across neighbouring lanes): plain, bias + activation, gated input gradient + residual, two sources, the
    ragged last column tile (128 = 4 * 30 + 8, 256 = 8 * 30 + 16) and a partial 64-channel chunk (Cin = 32)."""
    B, cin, H, W = shape
    cout = 64
    g = gen(31)
    x = q(torch.randn(B, cin, H, W, generator=g))
    x2 = q(torch.randn(B, 64, H, W, generator=g))
    w = q(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
    w2 = q(torch.randn(cout, 64, 3, 3, generator=g) / 24)
    b = torch.randn(cout, generator=g)
    gate = q(torch.randn(B, cout, H, W, generator=g))
    res = q(torch.randn(B, cout, H, W, generator=g))
    ref = F.conv2d(x, w, b, padding=1)
    raw, act = ops.conv(B, H, W, cout, [ops.Src(nhwc(x), pack(w), cin, 3)], bias=b.cuda(), want_raw=True, want_act=True)
    assert ops.last_conv_kernel() == "conv_stack3_kernel"
    assert rel_l2(nchw(raw), ref) < tol16()
    assert rel_l2(nchw(act), F.leaky_relu(ref, 0.2)) < tol16()
    ref2 = F.conv2d(x, w, None, padding=1) + F.conv2d(x2, w2, None, padding=1)
    ref2 = torch.where(gate > 0, ref2, 0.2 * ref2) + res
    got, _ = ops.conv(B, H, W, cout, [ops.Src(nhwc(x), pack(w), cin, 3), ops.Src(nhwc(x2), pack(w2), 64, 3)],
                      dmask=nhwc(gate), dmask_slope=0.2, residual=nhwc(res))
    if W >= 256 or ops.SPLIT:  # two 64-channel chunks on a 128-wide map: the pair kernel is ahead (conv_stack3.cu)
        assert ops.last_conv_kernel() == "conv_stack3_kernel"
    assert rel_l2(nchw(got), ref2) < tol16()
    again, _ = ops.conv(B, H, W, cout, [ops.Src(nhwc(x), pack(w), cin, 3), ops.Src(nhwc(x2), pack(w2), 64, 3)],
                        dmask=nhwc(gate), dmask_slope=0.2, residual=nhwc(res))
    assert torch.equal(ops.act_value(got), ops.act_value(again))


def test_argmax_rows_first_maximum(ops):
    """class_id.argmax(dim=-1) (models.py:151,501) for int64 one-hot labels and float rows; ties resolve to the first
    maximum like torch."""
    g = gen(41)
    B, n = 20, 365
    idx = torch.randint(0, n, (B,), generator=g)
    onehot = F.one_hot(idx, n).long()
    assert torch.equal(ops.argmax_rows(onehot.cuda()).cpu().long(), idx)
    x = torch.randint(0, 5, (B, n), generator=g).float()  # many ties
    x[3] = 0.0
    x[4, n - 1] = 9.0
    assert torch.equal(ops.argmax_rows(x.cuda()).cpu().long(), x.argmax(dim=-1))


def test_reductions_are_bit_reproducible(ops):
    """No floating-point atomics: split-K convolutions, weight gradients, bias column sums, BN statistics and the stencil
    gradient give bit-identical results on repeated launches (the small-map split-K and every weight gradient used
    red.global.add before, which made the discriminator's prediction move by 1e-2 from run to run)."""
    o = ops
    g = gen(21)
    B, Cin, Cout, H = 4, 256, 512, 8  # small map: split-K over the reduction
    x, w, dy = nhwc(q(torch.randn(B, Cin, H, H, generator=g))), pack(q(torch.randn(Cout, Cin, 3, 3, generator=g) * 0.03)), \
        nhwc(q(torch.randn(B, Cout, H, H, generator=g)))
    runs = []
    for _ in range(3):
        y, _ = o.conv(B, H, H, Cout, [o.Src(x, w, Cin, 3)])
        dw = torch.zeros(9, Cin, Cout, device="cuda")
        o.wgrad(x, dy, dw.data_ptr(), B, H, H, Cin, Cout, 3)
        db = torch.zeros(Cout, device="cuda")
        o.colsum(dy, Cout, db.data_ptr())
        runs.append((o.act_value(y).clone(), dw, db))
    for r in runs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(runs[0], r))
    B, C, H = 3, 64, 64  # large map: halo-tiled weight gradient with many pixel splits, BN statistics
    x, dy = nhwc(q(torch.randn(B, C, H, H, generator=g))), nhwc(q(torch.randn(B, C, H, H, generator=g)))
    runs = []
    for _ in range(3):
        dw = torch.zeros(9, C, C, device="cuda")
        o.wgrad(x, dy, dw.data_ptr(), B, H, H, C, C, 3)
        runs.append((dw, o.bn_sums(o.bn_stats(x), B * H * H, C)))
    for r in runs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(runs[0], r))
    torch.cuda.synchronize()
