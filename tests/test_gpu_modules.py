"""GPU parity of the drop-in modules against the CPU oracle (oracle/spyramid_oracle.py) on identical inputs/weights.

Tolerances: north_star asks rel-L2 <= 5e-3 for BF16-operand / FP32-accumulate kernels against the FP32 reference.
Single modules fed identical inputs meet it; deep chains accumulate BF16 rounding (SURVEY 7.2-1 measured the floor by
emulating BF16 operands inside the reference), so whole-network gradients are checked against the looser, stated
bounds below and reported with their measured value.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import spyramid_oracle as O  # noqa: E402


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _clone(sd):
    return {k: v.clone() for k, v in sd.items()}


def _to_cuda(ts):
    return [t.cuda() for t in ts]


@pytest.fixture(scope="module")
def batch():
    images, labels, masks, z_d, z_g = O.synthetic_batch(2, seed=3, mask_mode="blob")
    vsd = O.init_vgg_state(seed=5)
    with torch.no_grad():
        feats = O.vgg16_features(vsd, images)
    return dict(images=images, labels=labels, masks=masks, z=z_d, z2=z_g, vsd=vsd, feats=feats)


def test_vgg_features_match_oracle(batch):
    from semantic_pyramid_for_image_generation_b200 import models
    v = models.VGG16()
    v.load_state_dict(batch["vsd"])
    v.cuda().eval()
    with torch.no_grad():
        mine = v(batch["images"].cuda())
    assert len(mine) == 7
    for lvl, (m, r) in enumerate(zip(mine, batch["feats"])):
        assert tuple(m.shape) == tuple(r.shape)
        e = rel_l2(m, r)
        print("vgg level %d rel-L2 %.3e" % (lvl, e))
        assert e < (6e-3 if lvl < 3 else 1.2e-2), (lvl, e)
    assert float(mine[5].min()) >= 0.0  # fc7 tap is post-ReLU


def test_vgg_input_gradient_matches_oracle(batch):
    from semantic_pyramid_for_image_generation_b200 import models
    v = models.VGG16()
    v.load_state_dict(batch["vsd"])
    v.cuda().eval()
    for p in v.parameters():
        p.requires_grad = False
    gen = torch.Generator().manual_seed(11)
    ws = [torch.randn(f.shape, generator=gen) / f.numel() ** 0.5 for f in batch["feats"]]
    x = batch["images"].clone().requires_grad_(True)
    fr = O.vgg16_features(batch["vsd"], x)
    sum((f * w).sum() for f, w in zip(fr, ws)).backward()
    xc = batch["images"].cuda().requires_grad_(True)
    fm = v(xc)
    sum((f.float() * w.cuda()).sum() for f, w in zip(fm, ws)).backward()
    e = rel_l2(xc.grad, x.grad)
    print("vgg d/dimage rel-L2 %.3e" % e)
    assert e < 3e-2, e


@pytest.mark.parametrize("cf", [1, 2])
def test_generator_forward_backward_match_oracle(batch, cf):
    from semantic_pyramid_for_image_generation_b200 import models
    g_sd = O.init_generator_state(cf, seed=3)
    G = models.Generator(channels_factor=cf)
    G.load_state_dict(_clone(g_sd))
    G.cuda().train()
    cls = batch["labels"].float()
    # oracle
    ref_sd = _clone(g_sd)
    O._with_grad(ref_sd)
    img_ref = O.generator_forward(ref_sd, batch["z"], batch["feats"], batch["masks"], cls, training=True)
    r = torch.randn(img_ref.shape, generator=torch.Generator().manual_seed(7))
    (img_ref * r).sum().backward()
    # B200
    img = G(input=batch["z"].cuda(), features=_to_cuda(batch["feats"]), masks=_to_cuda(batch["masks"]),
            class_id=cls.cuda())
    assert tuple(img.shape) == tuple(img_ref.shape)
    e = rel_l2(img, img_ref)
    print("generator cf=%s image rel-L2 %.3e" % (cf, e))
    assert e < 2e-2, e
    (img * r.cuda()).sum().backward()
    sd = G.state_dict()
    for k in ("linear_layer.weight_u", "main_path.0.main_block.3.weight_v", "main_path.5.masked_feature_mapping.weight_u",
              "main_path.0.main_block.0.batch_norm.running_mean", "main_path.4.main_block.4.batch_norm.running_var",
              "final_block.1.running_var", "final_block.1.num_batches_tracked"):
        e = rel_l2(sd[k], ref_sd[k])
        assert e < 2e-2, (k, e)
    worst = 0.0
    num = den = 0.0
    for name, p in G.named_parameters():
        gr = ref_sd[name].grad
        assert p.grad is not None, name
        if gr is None or float(gr.norm()) < 1e-12:
            continue
        d = (p.grad.detach().float().cpu() - gr)
        num += float(d.pow(2).sum())
        den += float(gr.pow(2).sum())
        worst = max(worst, rel_l2(p.grad, gr))
    g_all = (num / den) ** 0.5
    print("generator cf=%s grads: global rel-L2 %.3e, worst tensor %.3e" % (cf, g_all, worst))
    assert g_all < 8e-2, g_all


@pytest.mark.parametrize("cf", [1, 2])
def test_discriminator_forward_backward_match_oracle(batch, cf):
    from semantic_pyramid_for_image_generation_b200 import models
    d_sd = O.init_discriminator_state(cf, seed=4)
    D = models.Discriminator(channel_factor=cf)
    D.load_state_dict(_clone(d_sd))
    D.cuda().train()
    ref_sd = _clone(d_sd)
    O._with_grad(ref_sd)
    x = batch["images"].clone().requires_grad_(True)
    p_ref = O.discriminator_forward(ref_sd, x, batch["labels"], training=True)
    r = torch.randn(p_ref.shape, generator=torch.Generator().manual_seed(9))
    (p_ref * r).sum().backward()
    xc = batch["images"].cuda().requires_grad_(True)
    p = D(xc, batch["labels"].cuda())
    assert tuple(p.shape) == (2, 2, 128)
    e = rel_l2(p, p_ref)
    print("discriminator cf=%s prediction rel-L2 %.3e" % (cf, e))
    assert e < 2e-2, e
    (p * r.cuda()).sum().backward()
    e = rel_l2(xc.grad, x.grad)
    print("discriminator cf=%s d/dimage rel-L2 %.3e" % (cf, e))
    assert e < 5e-2, e
    num = den = 0.0
    worst = ("", 0.0)
    for name, prm in D.named_parameters():
        gr = ref_sd[name].grad
        assert prm.grad is not None, name
        if gr is None or float(gr.norm()) < 1e-12:
            continue
        d = (prm.grad.detach().float().cpu() - gr)
        num += float(d.pow(2).sum())
        den += float(gr.pow(2).sum())
        el = rel_l2(prm.grad, gr)
        if el > worst[1]:
            worst = (name, el)
    g_all = (num / den) ** 0.5
    print("discriminator cf=%s grads: global rel-L2 %.3e, worst %s %.3e" % (cf, g_all, worst[0], worst[1]))
    assert g_all < 5e-2, g_all
    sd = D.state_dict()
    for k in ("layers.0.main_block.0.weight_u", "layers.7.main_block.3.weight_v", "embedding.weight_u"):
        assert rel_l2(sd[k], ref_sd[k]) < 1e-3, k


def test_losses_match_oracle(batch):
    from semantic_pyramid_for_image_generation_b200 import lossfunction as L
    gen = torch.Generator().manual_seed(21)
    p1 = torch.randn(2, 2, 128, generator=gen)
    p2 = torch.randn(2, 2, 128, generator=gen)
    a, b = L.LSGANDiscriminatorLoss()(p1.cuda(), p2.cuda())
    ra, rb = O.lsgan_discriminator_loss(p1, p2)
    assert abs(float(a) - float(ra)) < 1e-5 and abs(float(b) - float(rb)) < 1e-5
    assert abs(float(L.LSGANGeneratorLoss()(p2.cuda())) - float(O.lsgan_generator_loss(p2))) < 1e-5
    # gradient of the LSGAN loss
    pc = p2.cuda().requires_grad_(True)
    L.LSGANGeneratorLoss()(pc).backward()
    pr = p2.clone().requires_grad_(True)
    O.lsgan_generator_loss(pr).backward()
    assert rel_l2(pc.grad, pr.grad) < 1e-5
    # diversity
    img = torch.rand(4, 3, 32, 32, generator=gen) * 2 - 1
    z = torch.randn(4, 128, generator=gen)
    ic = img.cuda().requires_grad_(True)
    lv = L.DiversityLoss()(ic, z.cuda())
    ir = img.clone().requires_grad_(True)
    lr_ = O.diversity_loss(ir, z)
    assert abs(float(lv) - float(lr_)) < 1e-4 * abs(float(lr_))
    lv.backward()
    lr_.backward()
    assert rel_l2(ic.grad, ir.grad) < 1e-4
    # semantic reconstruction on BF16-representable features (so the only difference is the reduction order)
    feats_r = [f.bfloat16().float() if f.dim() == 4 else f for f in batch["feats"]]
    feats_f = [(f + 0.3 * torch.randn(f.shape, generator=gen)) for f in batch["feats"]]
    feats_f = [f.bfloat16().float() if f.dim() == 4 else f for f in feats_f]
    masks = batch["masks"]
    fc = [f.cuda().requires_grad_(True) for f in feats_f]
    lm = L.SemanticReconstructionLoss()(_to_cuda(feats_r), fc, _to_cuda(masks))
    fr = [f.clone().requires_grad_(True) for f in feats_f]
    lo = O.semantic_reconstruction_loss(feats_r, fr, masks)
    assert tuple(lm.shape) == (1,)
    assert abs(float(lm) - float(lo)) < 1e-4 * abs(float(lo)) + 1e-7
    lm.backward()
    lo.backward()
    for lvl, (gm, go) in enumerate(zip(fc, fr)):
        if float(go.grad.norm()) == 0.0:
            assert float(gm.grad.float().norm()) == 0.0
        else:
            assert rel_l2(gm.grad, go.grad) < 5e-3, lvl
