"""Developer probe: error of a split-BF16 3x3 convolution against a float64 reference as a function of the reduction
length K = 9 * Cin, and the SIGN of the error relative to the result (a truncating FP32 accumulator in the tensor pipe
shows up as a systematic shrink that grows with K / 16, the number of accumulations per output)."""
import sys, os, math
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_pyramid_for_image_generation_b200 import ops as o

o.set_precision("split")
g = torch.Generator().manual_seed(0)
for Cin, Cout, H in ((64, 64, 32), (128, 128, 32), (256, 256, 32), (512, 512, 16), (512, 512, 32)):
    B = 2
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
    rel = float(err.norm() / ref.norm())
    # projection of the error on the result: negative = shrink toward zero
    shrink = float((err * ref).sum() / (ref * ref).sum())
    # positive-only operands: every partial sum has the same sign, truncation bias fully visible
    xp, wp = x.abs(), w.abs()
    refp = F.conv2d(xp.double(), wp.double(), padding=1)
    yp, _ = o.conv(B, H, H, Cout, [o.Src(o.to_act(xp.permute(0, 2, 3, 1).contiguous().cuda()),
                                         o.to_act(wp.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous().cuda()), Cin, 3)])
    gp = o.act_value(yp).cpu().permute(0, 3, 1, 2).double()
    relp = float(((gp - refp) / refp).mean())
    print("K=%5d (%d MMAs/output x3 products): rel-L2 %.3e  projection on result %.3e | positive operands: mean rel err %.3e"
          % (9 * Cin, 9 * Cin // 16, rel, shrink, relp))
