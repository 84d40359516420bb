"""Developer micro-benchmark: small-map 3x3 convolutions (wave quantisation on the CTA-pair kernel)."""
import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_pyramid_for_image_generation_b200 import ops as o

B = 20
for (H, cin, cout) in ((16, 256, 256), (32, 256, 256), (32, 512, 512), (16, 512, 512), (64, 256, 256)):
    x = o.to_act(torch.randn(B, H, H, cin, device="cuda"))
    w = o.to_act((torch.randn(9, cout, cin, device="cuda") / math.sqrt(9 * cin)))
    for _ in range(3):
        o.conv(B, H, H, cout, [o.Src(x, w, cin, 3)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for i in range(n):
        o.conv(B, H, H, cout, [o.Src(x, w, cin, 3)])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * B * H * H * cout * cin * 9
    print("%dx%d %d->%d %s: %.1f us  %.0f TFLOP/s" % (H, H, cin, cout, o.last_conv_kernel(), us, fl / us / 1e6))
