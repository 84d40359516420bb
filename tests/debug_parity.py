"""Developer script (not a pytest file): prints per-tensor parity of the B200 modules against the CPU oracle."""
import sys

import torch

sys.path.insert(0, ".")
from oracle import spyramid_oracle as O  # noqa: E402
from semantic_pyramid_for_image_generation_b200 import models  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def clone(sd):
    return {k: v.clone() for k, v in sd.items()}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    cf = 1
    images, labels, masks, z, _ = O.synthetic_batch(2, seed=3, mask_mode="blob")
    vsd = O.init_vgg_state(seed=5)
    with torch.no_grad():
        feats = O.vgg16_features(vsd, images)
    cu = lambda ts: [t.cuda() for t in ts]
    if which in ("all", "vgg"):
        v = models.VGG16()
        v.load_state_dict(vsd)
        v.cuda().eval()
        for p in v.parameters():
            p.requires_grad = False
        gen = torch.Generator().manual_seed(11)
        ws = [torch.randn(f.shape, generator=gen) / f.numel() ** 0.5 for f in feats]
        for only in list(range(7)) + [None]:
            x = images.clone().requires_grad_(True)
            fr = O.vgg16_features(vsd, x)
            sum((f * w).sum() for i, (f, w) in enumerate(zip(fr, ws)) if only is None or i == only).backward()
            xc = images.cuda().requires_grad_(True)
            fm = v(xc)
            sum((f.float() * w.cuda()).sum() for i, (f, w) in enumerate(zip(fm, ws)) if only is None or i == only).backward()
            print("vgg d/dimage through tap %s: rel %.3e  (|ref| %.3e)" % (only, rel(xc.grad, x.grad), float(x.grad.norm())))
    if which in ("all", "d"):
        d_sd = O.init_discriminator_state(cf, seed=4)
        D = models.Discriminator(channel_factor=cf)
        D.load_state_dict(clone(d_sd))
        D.cuda().train()
        ref_sd = clone(d_sd)
        O._with_grad(ref_sd)
        x = images.clone().requires_grad_(True)
        taps = {}
        p_ref = O.discriminator_forward(ref_sd, x, labels, training=True, taps=taps)
        r = torch.randn(p_ref.shape, generator=torch.Generator().manual_seed(9))
        (p_ref * r).sum().backward()
        xc = images.cuda().requires_grad_(True)
        p = D(xc, labels.cuda())
        print("D out rel %.3e" % rel(p, p_ref))
        (p * r.cuda()).sum().backward()
        print("D d/dimage rel %.3e" % rel(xc.grad, x.grad))
        for name, prm in D.named_parameters():
            gr = ref_sd[name].grad
            print("  %-45s rel %.3e |ref| %.3e" % (name, rel(prm.grad, gr), float(gr.norm())))
    if which in ("all", "g"):
        g_sd = O.init_generator_state(cf, seed=3)
        G = models.Generator(channels_factor=cf)
        G.load_state_dict(clone(g_sd))
        G.cuda().train()
        cls = labels.float()
        ref_sd = clone(g_sd)
        O._with_grad(ref_sd)
        taps = {}
        img_ref = O.generator_forward(ref_sd, z, feats, masks, cls, training=True, taps=taps)
        r = torch.randn(img_ref.shape, generator=torch.Generator().manual_seed(7))
        (img_ref * r).sum().backward()
        img = G(input=z.cuda(), features=cu(feats), masks=cu(masks), class_id=cls.cuda())
        print("G image rel %.3e" % rel(img, img_ref))
        (img * r.cuda()).sum().backward()
        sd = G.state_dict()
        for k in sd:
            if k.endswith(("weight_u", "weight_v", "running_mean", "running_var")):
                e = rel(sd[k], ref_sd[k])
                if e > 1e-3:
                    print("  state %-55s rel %.3e" % (k, e))
        for name, prm in G.named_parameters():
            gr = ref_sd[name].grad
            if gr is None:
                print("  %-55s ref grad None" % name)
                continue
            print("  %-55s rel %.3e |ref| %.3e" % (name, rel(prm.grad, gr), float(gr.norm())))


if __name__ == "__main__":
    main()
