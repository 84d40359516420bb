"""Pins oracle/spyramid_oracle.py against the UNMODIFIED reference (imported through shims) on CPU.

Runs only where /root/reference exists (the build container); on the GPU box the committed goldens in
tests/golden/ (generated from the reference by tests/golden/make_golden.py) take over (test_golden.py).
"""
import copy
import random

import numpy as np
import pytest
import torch

from oracle import reference_shims
from oracle import spyramid_oracle as O

pytestmark = pytest.mark.skipif(not reference_shims.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return reference_shims.import_reference()


def _clone(sd):
    return {k: v.clone() for k, v in sd.items()}


def test_state_dict_keys_and_shapes_match_reference(ref):
    models, _, _ = ref
    for cf in (1, 2):
        g = models.Generator(channels_factor=cf)
        d = models.Discriminator(channel_factor=cf)
        for module, mine in ((g, O.init_generator_state(cf)), (d, O.init_discriminator_state(cf))):
            theirs = module.state_dict()
            assert list(theirs.keys()) == list(mine.keys())
            for k in theirs:
                assert tuple(theirs[k].shape) == tuple(mine[k].shape), k
    v = models.VGG16()
    mine = O.init_vgg_state()
    theirs = v.state_dict()
    assert set(theirs.keys()) == set(mine.keys())
    for k in theirs:
        assert tuple(theirs[k].shape) == tuple(mine[k].shape), k


def test_vgg_features_match_reference(ref):
    models, _, _ = ref
    sd = O.init_vgg_state(seed=5)
    v = models.VGG16()
    v.load_state_dict(sd)
    v.eval()
    x = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(0)) * 2 - 1
    with torch.no_grad():
        theirs = v(x)
        mine = O.vgg16_features(sd, x)
    assert [tuple(t.shape) for t in theirs] == [tuple(t.shape) for t in mine]
    for a, b in zip(theirs, mine):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    assert float(theirs[5].min()) == 0.0  # fc7 tap is post-ReLU


@pytest.mark.parametrize("cf", [1, 2])
def test_generator_discriminator_forward_and_state_match_reference(ref, cf):
    models, _, _ = ref
    g_sd, d_sd = O.init_generator_state(cf, seed=3), O.init_discriminator_state(cf, seed=4)
    g, d = models.Generator(channels_factor=cf), models.Discriminator(channel_factor=cf)
    g.load_state_dict(_clone(g_sd))
    d.load_state_dict(_clone(d_sd))
    g.train()
    d.train()
    images, labels, masks, z, _ = O.synthetic_batch(3, seed=1, mask_mode="blob")
    feats = O.vgg16_features(O.init_vgg_state(seed=5), images)
    with torch.no_grad():
        for _ in range(2):  # two forwards: power iteration / running stats advance identically
            theirs = g(input=z, features=feats, masks=masks, class_id=labels.float())
            mine = O.generator_forward(g_sd, z, feats, masks, labels.float(), training=True)
            assert torch.allclose(theirs, mine, rtol=1e-4, atol=1e-5)
            p_theirs = d(theirs, labels)
            p_mine = O.discriminator_forward(d_sd, mine, labels, training=True)
            assert p_theirs.shape == (3, 3, 128) and p_mine.shape == (3, 3, 128)
            assert torch.allclose(p_theirs, p_mine, rtol=1e-4, atol=1e-5)
    for module, sd in ((g, g_sd), (d, d_sd)):
        for k, t in module.state_dict().items():
            assert torch.allclose(t.float(), sd[k].float(), rtol=1e-4, atol=1e-6), k
    # eval mode: no power iteration, running statistics
    g.eval()
    with torch.no_grad():
        theirs = g(input=z, features=feats, masks=masks, class_id=labels.float())
        mine = O.generator_forward(g_sd, z, feats, masks, labels.float(), training=False)
    assert torch.allclose(theirs, mine, rtol=1e-4, atol=1e-5)


def test_losses_match_reference(ref):
    _, lossfunction, _ = ref
    gen = torch.Generator().manual_seed(0)
    fr = [torch.randn(2, 8, 16, 16, generator=gen), torch.randn(2, 4096, generator=gen)]
    ff = [torch.randn(2, 8, 16, 16, generator=gen), torch.randn(2, 4096, generator=gen)]
    ms = [(torch.rand(2, 1, 16, 16, generator=gen) > 0.5).float(), torch.ones(2, 4096)]
    assert torch.allclose(lossfunction.SemanticReconstructionLoss()(fr, ff, ms),
                          O.semantic_reconstruction_loss(fr, ff, ms))
    img, z = torch.randn(4, 3, 8, 8, generator=gen), torch.randn(4, 128, generator=gen)
    assert torch.allclose(lossfunction.DiversityLoss()(img, z), O.diversity_loss(img, z))
    p, q = torch.randn(4, 4, 128, generator=gen), torch.randn(4, 4, 128, generator=gen)
    assert torch.allclose(lossfunction.LSGANGeneratorLoss()(p), O.lsgan_generator_loss(p))
    a, b = lossfunction.LSGANDiscriminatorLoss()(p, q)
    c, d = O.lsgan_discriminator_loss(p, q)
    assert torch.allclose(a, c) and torch.allclose(b, d)


def test_full_train_step_matches_reference(ref):
    """model_wrapper.py:136-190 executed with the reference modules vs oracle.train_step (B=2, config 1)."""
    models, lossfunction, _ = ref
    g_sd, d_sd, v_sd = O.init_generator_state(2, seed=3), O.init_discriminator_state(2, seed=4), O.init_vgg_state(5)
    g, d, v = models.Generator(channels_factor=2), models.Discriminator(channel_factor=2), models.VGG16()
    g.load_state_dict(_clone(g_sd))
    d.load_state_dict(_clone(d_sd))
    v.load_state_dict(_clone(v_sd))
    g.train(); d.train(); v.eval()
    for p in v.parameters():
        p.requires_grad = False
    g_opt = torch.optim.Adam(g.parameters(), lr=1e-5)
    d_opt = torch.optim.Adam(d.parameters(), lr=1e-5)
    images, labels, masks, z_d, z_g = O.synthetic_batch(2, seed=0, mask_mode="inference")
    # --- the reference step body, verbatim order ---
    g.zero_grad(); d.zero_grad()
    with torch.no_grad():
        features_real = v(images)
        images_fake = g(input=z_d, features=features_real, masks=masks, class_id=labels.float())
    prediction_real = d(images, labels)
    prediction_fake = d(images_fake, labels)
    l_real, l_fake = lossfunction.LSGANDiscriminatorLoss()(prediction_real, prediction_fake)
    (l_real + l_fake).backward()
    d_grads_ref = {k: p.grad.clone() for k, p in d.named_parameters()}
    d_opt.step()
    g.zero_grad(); d.zero_grad()
    images_fake = g(input=z_g, features=features_real, masks=masks, class_id=labels.float())
    prediction_fake = d(images_fake, labels)
    l_g = lossfunction.LSGANGeneratorLoss()(prediction_fake)
    l_div = 0.1 * lossfunction.DiversityLoss()(images_fake, z_g)
    features_fake = v(images_fake)
    l_rec = 0.1 * lossfunction.SemanticReconstructionLoss()(features_real, features_fake, masks)
    (l_g + l_rec + l_div).backward()
    g_grads_ref = {k: p.grad.clone() for k, p in g.named_parameters()}
    g_opt.step()
    # --- oracle ---
    out = O.train_step(g_sd, d_sd, v_sd, images, labels, masks, z_d, z_g, {}, {}, lr=1e-5)
    assert out["loss_discriminator_real"] == pytest.approx(float(l_real), rel=1e-4)
    assert out["loss_discriminator_fake"] == pytest.approx(float(l_fake), rel=1e-3, abs=1e-7)
    assert out["loss_generator"] == pytest.approx(float(l_g), rel=1e-4)
    assert out["loss_generator_semantic_reconstruction"] == pytest.approx(float(l_rec), rel=1e-4)
    assert out["loss_generator_diversity"] == pytest.approx(float(l_div), rel=1e-4)

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-20))

    for k, gref in d_grads_ref.items():
        if float(gref.abs().max()) > 1e-7:  # skips analytically-zero grads (bias feeding a BatchNorm)
            tol = 3e-2 if gref.numel() == 1 else 2e-3
            assert rel(out["d_grads"][k], gref) < tol, (k, rel(out["d_grads"][k], gref))
    for k, gref in g_grads_ref.items():
        if float(gref.abs().max()) > 1e-7:  # skips analytically-zero grads (bias feeding a BatchNorm)
            tol = 3e-2 if gref.numel() == 1 else 2e-3  # scalar gamma: one fp32 sum with heavy cancellation
            assert rel(out["g_grads"][k], gref) < tol, (k, rel(out["g_grads"][k], gref))
    # Adam's first step moves every weight by ~lr*sign(g): entries whose gradient is at rounding-noise level may
    # flip sign between two fp32 evaluations, so post-step parameters agree to 2*lr, buffers (u, v, BN) tightly.
    for module, sd in ((g, g_sd), (d, d_sd)):
        for k, t in module.state_dict().items():
            atol = 2.5e-5 if k in O.trainable_keys(sd) else 2e-6
            assert torch.allclose(t.float(), sd[k].float(), rtol=1e-3, atol=atol), k


def test_masks_bit_exact(ref):
    _, _, misc = ref
    for stage in range(7):
        for a, b in zip(misc.get_masks_for_inference(stage), O.masks_for_inference(stage)):
            assert torch.equal(a, b)
    # training masks: same RNG consumption and selection logic; the rasteriser is injected on both sides
    rng = np.random.RandomState(7)
    canned = {}

    def shape_image(hw, min_size):
        key = (hw, min_size, len(canned))
        img = np.full(hw, 255, dtype=np.uint8)
        y, x = rng.randint(0, hw[0] // 2), rng.randint(0, hw[1] // 2)
        img[y:y + max(min_size, 2), x:x + max(min_size, 2)] = rng.randint(0, 200)
        canned[key] = img
        return img

    def fake_random_shapes(shape, min_shapes, max_shapes, min_size, allow_overlap):
        img = shape_image(tuple(shape), min_size)
        return np.stack([img, img, img], axis=-1), None

    misc.random_shapes = fake_random_shapes
    n_spatial = 0
    for seed in range(60):
        random.seed(seed); np.random.seed(seed); rng.seed(seed)
        theirs = misc.get_masks_for_training()
        random.seed(seed); np.random.seed(seed); rng.seed(seed)
        mine = O.masks_for_training(shape_image)
        assert len(theirs) == len(mine) == 7
        for a, b in zip(theirs, mine):
            assert a.shape == b.shape and torch.equal(a, b)
        n_spatial += int(any(0 < float(m.mean()) < 1 for m in mine))
    assert n_spatial > 0


def test_checkpoint_interchange_with_reference_modules(ref, tmp_path):
    """a reference checkpoint (model_wrapper.py:215-223) loads into the B200 modules and back, including the
    optimizer state, with strict key/shape matching."""
    from semantic_pyramid_for_image_generation_b200 import models as new_models
    from semantic_pyramid_for_image_generation_b200.optim import FusedAdam
    ref_models, _, _ = ref
    for cf in (1, 2):
        g_ref, d_ref = ref_models.Generator(channels_factor=cf), ref_models.Discriminator(channel_factor=cf)
        g_new, d_new = new_models.Generator(channels_factor=cf), new_models.Discriminator(channel_factor=cf)
        g_opt_ref = torch.optim.Adam(g_ref.parameters(), lr=1e-5)
        for p in g_ref.parameters():
            p.grad = torch.ones_like(p)
        g_opt_ref.step()
        path = tmp_path / ("checkpoint_%d.pt" % cf)
        torch.save({"generator": g_ref.state_dict(), "discriminator": d_ref.state_dict(),
                    "generator_optimizer": g_opt_ref.state_dict()}, path)
        ckpt = torch.load(path, weights_only=False)
        g_new.load_state_dict(ckpt["generator"])  # strict
        d_new.load_state_dict(ckpt["discriminator"])
        g_opt_new = FusedAdam(g_new.parameters(), lr=1e-5)
        g_opt_new.load_state_dict(ckpt["generator_optimizer"])
        assert len(g_opt_new.state_dict()["state"]) == len(ckpt["generator_optimizer"]["state"])
        for k, v in g_ref.state_dict().items():
            assert torch.equal(v, g_new.state_dict()[k]), k
        # and back: a checkpoint written by the B200 modules loads into the reference
        g_ref2, d_ref2 = ref_models.Generator(channels_factor=cf), ref_models.Discriminator(channel_factor=cf)
        g_ref2.load_state_dict(g_new.state_dict())
        d_ref2.load_state_dict(d_new.state_dict())
        torch.optim.Adam(g_ref2.parameters(), lr=1e-5).load_state_dict(g_opt_new.state_dict())
        assert [n for n, _ in g_ref2.named_parameters()] == [n for n, _ in g_new.named_parameters()]
        assert [n for n, _ in d_ref2.named_parameters()] == [n for n, _ in d_new.named_parameters()]
    v_ref, v_new = ref_models.VGG16(), new_models.VGG16()
    v_new.load_state_dict(v_ref.state_dict())
    v_ref.load_state_dict(v_new.state_dict())


def _realistic_running_stats(g_sd, z, feats, masks, cls, seed=17):
    """Fills the batch-norm running statistics with what a long training run would hold: the batch statistics of a
    training-mode forward (momentum forced to 1), jittered by 10 %% so that eval mode is distinguishable from train mode.
    (With the constructor's mean 0 / var 1 a random-init generator saturates tanh everywhere and eval parity is vacuous.)"""
    from unittest import mock
    real = O._bn_train
    with mock.patch.object(O, "_bn_train", lambda x, sd, prefix, momentum, training: real(x, sd, prefix, 1.0, training)):
        with torch.no_grad():
            O.generator_forward(g_sd, z, feats, masks, cls, training=True)
    gen = torch.Generator().manual_seed(seed)
    for k, v in g_sd.items():
        if k.endswith("running_mean"):
            v.add_(0.1 * v.abs().mean() * torch.randn(v.shape, generator=gen))
        elif k.endswith("running_var"):
            v.mul_(1.0 + 0.1 * (2 * torch.rand(v.shape, generator=gen) - 1))


def test_eval_mode_generator_matches_reference(ref):
    """Pins the oracle's eval-mode generator (running batch-norm statistics, no power iteration: what `inference()` and
    `validate()` run, model_wrapper.py:231-296) against the reference's `generator.eval()`."""
    models, _, _ = ref
    cf = 2
    g_sd = O.init_generator_state(cf, seed=3)
    images, labels, masks, z, _ = O.synthetic_batch(2, seed=2, mask_mode="inference")
    feats = O.vgg16_features(O.init_vgg_state(seed=5), images)
    _realistic_running_stats(g_sd, z, feats, masks, labels.float())
    g = models.Generator(channels_factor=cf)
    g.load_state_dict(_clone(g_sd))
    g.eval()
    before = _clone(g_sd)
    with torch.no_grad():
        theirs = g(input=z, features=feats, masks=masks, class_id=labels.float())
        mine = O.generator_forward(g_sd, z, feats, masks, labels.float(), training=False)
    assert torch.allclose(theirs, mine, rtol=1e-4, atol=1e-5)
    assert float(mine.abs().mean()) < 0.9  # not saturated: the comparison means something
    for k in g_sd:  # eval mode leaves u, v and the running statistics untouched
        assert torch.equal(g_sd[k], before[k]), k


def test_data_parallel_oracle_reduces_to_the_single_process_step():
    """train_step_data_parallel (reference step per shard, gradients averaged) with ONE replica is train_step,
    and with two replicas each phase's gradient is the mean of the per-shard gradients of the phase functions."""
    cf = 2
    images, labels, masks, z_d, z_g = O.synthetic_batch(2, seed=4, mask_mode="inference")
    v_sd = O.init_vgg_state(seed=5)
    g1, d1 = O.init_generator_state(cf, seed=3), O.init_discriminator_state(cf, seed=4)
    g2, d2 = _clone(g1), _clone(d1)
    a = O.train_step(g1, d1, v_sd, images, labels, masks, z_d, z_g, {}, {}, lr=1e-4)
    b = O.train_step_data_parallel([(g2, d2)], v_sd, [(images, labels, masks, z_d, z_g)], {}, {}, lr=1e-4)[0]
    for k in ("loss_discriminator_real", "loss_generator", "loss_generator_semantic_reconstruction",
              "loss_generator_diversity"):
        assert a[k] == b[k], k
    for k in g1:
        assert torch.equal(g1[k], g2[k]), k
    for k in d1:
        assert torch.equal(d1[k], d2[k]), k
    # two replicas, one sample each: the D-phase gradient is the mean of the two per-shard gradients
    shards = []
    for r in range(2):
        im, lb, mk, zd, zg = O.synthetic_batch(2, seed=10 + r, mask_mode="inference")
        shards.append((im, lb, mk, zd, zg))
    g0, d0 = O.init_generator_state(cf, seed=3), O.init_discriminator_state(cf, seed=4)
    per_shard = []
    for sh in shards:
        _, _, _, gr = O._discriminator_phase(_clone(g0), _clone(d0), v_sd, sh[0], sh[1], sh[2], sh[3])
        per_shard.append(gr)
    reps = [(_clone(g0), _clone(d0)), (_clone(g0), _clone(d0))]
    out = O.train_step_data_parallel(reps, v_sd, shards, {}, {}, lr=1e-4)
    for k in per_shard[0]:
        assert torch.allclose(out[0]["d_grads"][k], (per_shard[0][k] + per_shard[1][k]) / 2, rtol=0, atol=0), k
    for k in O.trainable_keys(reps[0][1]):
        assert torch.equal(reps[0][1][k], reps[1][1][k]), k  # replicas stay identical after the update
