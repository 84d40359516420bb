import sys, os, math, torch
import torch.nn.functional as F
sys.path.insert(0, "/root/repo")
from semantic_pyramid_for_image_generation_b200 import ops as o
o.set_precision("split")
g = torch.Generator().manual_seed(0)
for Cin, H in ((64, 128), (128, 256), (64, 256)):
    B, Cout = 1, 64
    x = torch.randn(B, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    q = lambda t: (lambda hi: hi + (t - hi).bfloat16().float())(t.bfloat16().float())
    x, w = q(x), q(w)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    xc = o.to_act(x.permute(0, 2, 3, 1).contiguous().cuda())
    wc = o.to_act(w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous().cuda())
    y, _ = o.conv(B, H, H, Cout, [o.Src(xc, wc, Cin, 3)])
    got = o.act_value(y).cpu().permute(0, 3, 1, 2).double()
    err = got - ref
    print(o.last_conv_kernel(), "Cin", Cin, "H", H, "rel-L2 %.3e" % float(err.norm() / ref.norm()),
          "projection %.3e" % float((err * ref).sum() / (ref * ref).sum()))
