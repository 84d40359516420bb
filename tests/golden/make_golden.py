"""Generates tests/golden/*.pt by executing the UNMODIFIED reference (/root/reference, imported through
oracle/reference_shims.py) on seeded synthetic inputs.  Run in the build container:

    python tests/golden/make_golden.py

The fixtures pin oracle/spyramid_oracle.py and the product's mask builders on machines where the reference tree is
absent (the GPU box).  Weights and inputs are NOT stored: they are regenerated from the seeds by the oracle's
deterministic initialisers; the fixtures hold what the reference computed from them.
"""
import hashlib
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shims  # noqa: E402
from oracle import spyramid_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CF, BATCH, LR = 2, 2, 1e-5
SEEDS = dict(g=3, d=4, v=5, batch=0)


def digest(tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().float().numpy().tobytes())
    return h.hexdigest()


def _head(named_parameters, n=256):
    """First n entries of every gradient tensor: element-level pins (norms alone cannot see a wrong direction)."""
    return {k: p.grad.flatten()[:n].clone() for k, p in named_parameters if p.grad is not None}


def golden_step(CF=CF):
    models, lossfunction, _ = reference_shims.import_reference()
    g_sd, d_sd, v_sd = (O.init_generator_state(CF, seed=SEEDS["g"]), O.init_discriminator_state(CF, seed=SEEDS["d"]),
                        O.init_vgg_state(SEEDS["v"]))
    g, d, v = models.Generator(channels_factor=CF), models.Discriminator(channel_factor=CF), models.VGG16()
    g.load_state_dict(g_sd)
    d.load_state_dict(d_sd)
    v.load_state_dict(v_sd)
    g.train(); d.train(); v.eval()
    for p in v.parameters():
        p.requires_grad = False
    g_opt, d_opt = torch.optim.Adam(g.parameters(), lr=LR), torch.optim.Adam(d.parameters(), lr=LR)
    images, labels, masks, z_d, z_g = O.synthetic_batch(BATCH, seed=SEEDS["batch"], mask_mode="blob")
    # reference step body (model_wrapper.py:136-190), its own order
    g.zero_grad(); d.zero_grad()
    with torch.no_grad():
        features_real = v(images)
        images_fake_d = g(input=z_d, features=features_real, masks=masks, class_id=labels.float())
    prediction_real = d(images, labels)
    prediction_fake = d(images_fake_d, labels)
    l_real, l_fake = lossfunction.LSGANDiscriminatorLoss()(prediction_real, prediction_fake)
    (l_real + l_fake).backward()
    d_grad_norms = {k: float(p.grad.norm()) for k, p in d.named_parameters()}
    d_grad_head = {k: p.grad.flatten()[:64].clone() for k, p in d.named_parameters() if k.endswith("main_block.3.weight_orig")}
    d_grad_sub = _head(d.named_parameters())
    prediction_fake_d = prediction_fake.detach().clone()
    d_opt.step()
    g.zero_grad(); d.zero_grad()
    images_fake = g(input=z_g, features=features_real, masks=masks, class_id=labels.float())
    prediction_fake_g = d(images_fake, labels)
    l_g = lossfunction.LSGANGeneratorLoss()(prediction_fake_g)
    l_div = 0.1 * lossfunction.DiversityLoss()(images_fake, z_g)
    features_fake = v(images_fake)
    l_rec = 0.1 * lossfunction.SemanticReconstructionLoss()(features_real, features_fake, masks)
    (l_g + l_rec + l_div).backward()
    g_grad_norms = {k: float(p.grad.norm()) for k, p in g.named_parameters()}
    g_grad_sub = _head(g.named_parameters())
    g_opt.step()
    gsd, dsd = g.state_dict(), d.state_dict()
    return {
        "config": dict(channel_factor=CF, batch=BATCH, lr=LR, seeds=SEEDS, mask_mode="blob"),
        "losses": {"loss_discriminator_real": float(l_real), "loss_discriminator_fake": float(l_fake),
                   "loss_generator": float(l_g), "loss_generator_semantic_reconstruction": float(l_rec),
                   "loss_generator_diversity": float(l_div)},
        "features_real_sub": [f[:, ::8, ::8, ::8].clone() if f.dim() == 4 else f[:, ::16].clone() for f in features_real],
        "features_real_norm": [float(f.norm()) for f in features_real],
        "images_fake_sub": images_fake.detach()[:, :, ::8, ::8].clone(),
        "images_fake_norm": float(images_fake.norm()),
        "prediction_real": prediction_real.detach().clone(),
        "prediction_fake_g": prediction_fake_g.detach().clone(),
        "prediction_fake_d": prediction_fake_d,
        "d_grad_norms": d_grad_norms, "g_grad_norms": g_grad_norms, "d_grad_head": d_grad_head,
        "d_grad_sub": d_grad_sub, "g_grad_sub": g_grad_sub,
        "post_step": {k: gsd[k].clone() for k in ("linear_layer.weight_u", "main_path.2.main_block.3.weight_v",
                                                  "main_path.0.main_block.0.batch_norm.running_mean",
                                                  "final_block.1.running_var", "final_block.1.num_batches_tracked")},
        "post_step_d": {k: dsd[k].clone() for k in ("layers.0.main_block.0.weight_u", "embedding.weight_v",
                                                    "classification.weight_orig")},
    }


def golden_masks():
    """get_masks_for_training / _for_inference of the reference with an injected, seeded rasteriser."""
    _, _, misc = reference_shims.import_reference()
    from semantic_pyramid_for_image_generation_b200 import misc as product_misc
    misc.random_shapes = product_misc._builtin_random_shapes
    out = {"training": [], "inference": []}
    for seed in range(48):
        random.seed(seed)
        np.random.seed(seed)
        masks = misc.get_masks_for_training()
        out["training"].append(dict(seed=seed, sha256=digest(masks), means=[float(m.mean()) for m in masks]))
    for stage in range(7):
        out["inference"].append(dict(stage=stage, sha256=digest(misc.get_masks_for_inference(stage))))
    return out


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.save(golden_step(2), os.path.join(HERE, "step_cf2_b2.pt"))
    torch.manual_seed(0)
    torch.save(golden_step(1), os.path.join(HERE, "step_cf1_b2.pt"))  # BASELINE.json configs[0]: cf=1, batch 2
    torch.save(golden_masks(), os.path.join(HERE, "masks.pt"))
    for name in ("step_cf2_b2.pt", "step_cf1_b2.pt", "masks.pt"):
        print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")
