"""Parity of the B200 forward AND backward schedules against the oracle with BF16 storage points emulated
(oracle/bf16_emulation.py), at two granularities.

* Block level (one residual block = 3-4 convolutions + norms/pools, fed identical BF16-representable inputs): both
  sides take the same ReLU / LeakyReLU / arg-max gates, so the comparison measures the kernels and their scheduling:
  forward rel-L2 <= 5e-3 (north_star), input/weight gradients <= 2e-2 (the B200 backward rounds the gradient to BF16
  once per stage; the oracle's autograd is FP32).
* Whole networks: measured on B200, the emulation tracks the B200 activations to 5e-5 after two convolutions and then
  drifts (VGG taps 4e-4, 2.4e-3, 4.2e-3, 5.2e-3): random-init deep nets amplify the 1-ulp BF16 flips caused by a
  different FP32 summation order ~2x per layer, so ANY two correct BF16 implementations differ by ~5e-3 at pool5 depth
  and by a few 1e-2 in deep gradients.  Whole-network bounds are therefore the measured floors with head-room."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import bf16_emulation as E  # noqa: E402
from oracle import spyramid_oracle as O  # noqa: E402


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _clone(sd):
    return {k: v.clone() for k, v in sd.items()}


def _global_rel(named_params, ref_sd, min_numel=1024):
    num = den = 0.0
    worst = ("", 0.0)
    for name, p in named_params:
        gr = ref_sd[name].grad
        if gr is None or p.numel() < min_numel or float(gr.norm()) < 1e-12:
            continue
        d = p.grad.detach().float().cpu() - gr
        num += float(d.pow(2).sum())
        den += float(gr.pow(2).sum())
        e = rel_l2(p.grad, gr)
        if e > worst[1]:
            worst = (name, e)
    return (num / den) ** 0.5, worst


@pytest.fixture(scope="module")
def batch():
    images, labels, masks, z_d, _ = O.synthetic_batch(2, seed=3, mask_mode="blob")
    vsd = O.init_vgg_state(seed=5)
    with torch.no_grad():
        feats = E.vgg16_features(vsd, images)
    return dict(images=images, labels=labels, masks=masks, z=z_d, vsd=vsd, feats=feats)


def test_vgg_forward_and_input_gradient(batch):
    from semantic_pyramid_for_image_generation_b200 import models
    v = models.VGG16()
    v.load_state_dict(batch["vsd"])
    v.cuda().eval()
    for p in v.parameters():
        p.requires_grad = False
    gen = torch.Generator().manual_seed(11)
    ws = [torch.randn(f.shape, generator=gen) / f.numel() ** 0.5 for f in batch["feats"]]
    x = batch["images"].clone().requires_grad_(True)
    fr = E.vgg16_features(batch["vsd"], x)
    sum((f * w).sum() for f, w in zip(fr, ws)).backward()
    xc = batch["images"].cuda().requires_grad_(True)
    fm = v(xc)
    for lvl, (a, b) in enumerate(zip(fm, fr)):
        e = rel_l2(a, b)
        print("vgg level %d rel-L2 vs emulated oracle %.3e" % (lvl, e))
        assert e < (1e-3 if lvl < 2 else 1.2e-2), (lvl, e)
    sum((f.float() * w.cuda()).sum() for f, w in zip(fm, ws)).backward()
    e = rel_l2(xc.grad, x.grad)
    print("vgg d/dimage rel-L2 vs emulated oracle %.3e" % e)
    assert e < 0.25, e


@pytest.mark.parametrize("cf", [1, 2])
def test_generator_forward_and_gradients(batch, cf):
    from semantic_pyramid_for_image_generation_b200 import models
    g_sd = O.init_generator_state(cf, seed=3)
    G = models.Generator(channels_factor=cf)
    G.load_state_dict(_clone(g_sd))
    G.cuda().train()
    cls = batch["labels"].float()
    ref_sd = _clone(g_sd)
    O._with_grad(ref_sd)
    img_ref = E.generator_forward(ref_sd, batch["z"], batch["feats"], batch["masks"], cls, training=True)
    r = torch.randn(img_ref.shape, generator=torch.Generator().manual_seed(7))
    (img_ref * r).sum().backward()
    img = G(input=batch["z"].cuda(), features=[f.cuda() for f in batch["feats"]],
            masks=[m.cuda() for m in batch["masks"]], class_id=cls.cuda())
    e = rel_l2(img, img_ref)
    print("generator cf=%s image rel-L2 vs emulated oracle %.3e" % (cf, e))
    assert e < 1.5e-2, e
    (img * r.cuda()).sum().backward()
    g_all, worst = _global_rel(G.named_parameters(), ref_sd)
    print("generator cf=%s weight gradients: global rel-L2 %.3e, worst %s %.3e" % (cf, g_all, worst[0], worst[1]))
    assert g_all < 0.15, g_all


@pytest.mark.parametrize("cf", [1, 2])
def test_discriminator_forward_and_gradients(batch, cf):
    from semantic_pyramid_for_image_generation_b200 import models
    d_sd = O.init_discriminator_state(cf, seed=4)
    D = models.Discriminator(channel_factor=cf)
    D.load_state_dict(_clone(d_sd))
    D.cuda().train()
    ref_sd = _clone(d_sd)
    O._with_grad(ref_sd)
    x = batch["images"].clone().requires_grad_(True)
    p_ref = E.discriminator_forward(ref_sd, x, batch["labels"], training=True)
    r = torch.randn(p_ref.shape, generator=torch.Generator().manual_seed(9))
    (p_ref * r).sum().backward()
    xc = batch["images"].cuda().requires_grad_(True)
    p = D(xc, batch["labels"].cuda())
    e = rel_l2(p, p_ref)
    print("discriminator cf=%s prediction rel-L2 vs emulated oracle %.3e" % (cf, e))
    # whole-network figure: BF16 gate flips + the run-to-run order of the split-K FP32 atomics on the 4x4/8x8 maps
    # move it between 1.2e-2 and 2.3e-2 (measured over repeated runs); the tight bounds are the block-level tests
    assert e < 4e-2, e
    (p * r.cuda()).sum().backward()
    e = rel_l2(xc.grad, x.grad)
    print("discriminator cf=%s d/dimage rel-L2 vs emulated oracle %.3e" % (cf, e))
    assert e < 0.2, e
    g_all, worst = _global_rel(D.named_parameters(), ref_sd)
    print("discriminator cf=%s weight gradients: global rel-L2 %.3e, worst %s %.3e" % (cf, g_all, worst[0], worst[1]))
    assert g_all < 6e-2, g_all


# ------------------------------------------------------------------------------------------------
# block level: identical inputs, same gates -> tight bounds
# ------------------------------------------------------------------------------------------------
def _q(t):
    return t.bfloat16().float()


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().bfloat16().cuda()


def _nchw(t):
    return t.float().cpu().permute(0, 3, 1, 2).contiguous()


def _block_param_check(module, prefix, ref_sd, grad_arena, ga, tol):
    worst = ("", 0.0)
    for name, p in module.named_parameters():
        if not name.startswith(prefix + "."):
            continue
        gr = ref_sd[name].grad
        if gr is None or p.numel() < 1024 or float(gr.norm()) < 1e-12:
            continue
        o = ga.offset(p)
        mine = grad_arena[o:o + p.numel()].view(p.shape)
        e = rel_l2(mine, gr)
        if e > worst[1]:
            worst = (name, e)
    assert worst[1] < tol, worst
    return worst


@pytest.mark.parametrize("cf,idx", [(1, 2), (1, 5), (2, 0), (2, 4)])
def test_generator_block_forward_backward_tight(cf, idx):
    from semantic_pyramid_for_image_generation_b200 import engine, models, ops
    g_sd = O.init_generator_state(cf, seed=3)
    G = models.Generator(channels_factor=cf)
    G.load_state_dict(_clone(g_sd))
    G.cuda().train()
    blk = G.main_path[idx]
    key = "main_path.%d" % idx
    cin, cout = blk.main_block[3].shape[1], blk.main_block[3].shape[0]
    cfeat = blk.masked_feature_mapping.shape[1] - 1
    res = {0: 4, 1: 8, 2: 16, 4: 32, 5: 64}[idx]
    gen = torch.Generator().manual_seed(31 + idx)
    B = 2
    x = _q(torch.randn(B, cin, res, res, generator=gen))
    feat = _q(torch.randn(B, cfeat, 2 * res, 2 * res, generator=gen).relu())
    mask = (torch.rand(B, 1, 2 * res, 2 * res, generator=gen) > 0.4).float()
    mask[0] = 1.0
    labels = torch.nn.functional.one_hot(torch.tensor([5, 200]), 365).float()
    gout = _q(torch.randn(B, cout, 2 * res, 2 * res, generator=gen))
    ref_sd = _clone(g_sd)
    O._with_grad(ref_sd)
    xr = x.clone().requires_grad_(True)
    out_ref = E._generator_block(ref_sd, key, xr, feat, mask, labels, True)
    out_ref.backward(gout)
    st = G._sn.forward(True)
    cls = ops.argmax_rows(labels.cuda())
    xc = _nhwc(x)
    out, ctx = engine._gblock_forward(blk, key, st, xc, feat.cuda(), mask.cuda(), cls, True, True)
    e = rel_l2(_nchw(out), out_ref)
    print("G block %s cf=%s out rel-L2 %.3e" % (key, cf, e))
    assert e < 5e-3, e
    grad = G._ga.new("cuda")
    gw = torch.zeros(G._sn.gw_floats, device="cuda")
    gx = engine._gblock_backward(blk, key, st, G._sn, G._ga, gw, grad, ctx, cls, _nhwc(gout))
    G._sn.backward(st, gw, grad)
    e = rel_l2(_nchw(gx), xr.grad)
    print("G block %s cf=%s d/dx rel-L2 %.3e" % (key, cf, e))
    assert e < 2e-2, e
    worst = _block_param_check(G, key, ref_sd, grad, G._ga, 2e-2)
    print("G block %s cf=%s worst weight gradient %s %.3e" % (key, cf, worst[0], worst[1]))


@pytest.mark.parametrize("cf,idx", [(1, 1), (1, 4), (2, 7)])
def test_discriminator_block_forward_backward_tight(cf, idx):
    from semantic_pyramid_for_image_generation_b200 import engine, models
    import torch.nn.functional as F
    d_sd = O.init_discriminator_state(cf, seed=4)
    D = models.Discriminator(channel_factor=cf)
    D.load_state_dict(_clone(d_sd))
    D.cuda().train()
    blk = D.layers[idx]
    key = "layers.%d" % idx
    cin, cout = blk.main_block[1].shape[1], blk.main_block[1].shape[0]
    res = {1: 32, 4: 16, 7: 16}[idx]
    gen = torch.Generator().manual_seed(41 + idx)
    B = 2
    x = _q(torch.randn(B, cin, res, res, generator=gen))
    gout = _q(torch.randn(B, cout, res // 2, res // 2, generator=gen))
    ref_sd = _clone(d_sd)
    O._with_grad(ref_sd)
    xr = x.clone().requires_grad_(True)
    act = E.q(F.leaky_relu(xr, 0.2))
    h = E.q(F.leaky_relu(F.conv2d(act, E._w(ref_sd, key + ".main_block.1", True), ref_sd[key + ".main_block.1.bias"],
                                  padding=1), 0.2))
    s = E.q(F.conv2d(h, E._w(ref_sd, key + ".main_block.3", True), ref_sd[key + ".main_block.3.bias"], padding=1) +
            F.conv2d(xr, E._w(ref_sd, key + ".residual_mapping", True), ref_sd[key + ".residual_mapping.bias"]))
    out_ref = F.avg_pool2d(s, 2)
    out_ref.backward(gout)
    st = D._sn.forward(True)
    xc = _nhwc(x)
    ac = _nhwc(F.leaky_relu(x, 0.2))
    out, _, ctx = engine._dblock_forward(blk, key, st, xc, ac, False, True)
    e = rel_l2(_nchw(out), out_ref)
    print("D block %s cf=%s out rel-L2 %.3e" % (key, cf, e))
    assert e < 5e-3, e
    grad = D._ga.new("cuda")
    gw = torch.zeros(D._sn.gw_floats, device="cuda")
    gx = engine._dblock_backward(blk, key, st, D._sn, D._ga, gw, grad, ctx, _nhwc(gout), True)
    D._sn.backward(st, gw, grad)
    e = rel_l2(_nchw(gx), xr.grad)
    print("D block %s cf=%s d/dx rel-L2 %.3e" % (key, cf, e))
    assert e < 2e-2, e
    worst = _block_param_check(D, key, ref_sd, grad, D._ga, 2e-2)
    print("D block %s cf=%s worst weight gradient %s %.3e" % (key, cf, worst[0], worst[1]))
