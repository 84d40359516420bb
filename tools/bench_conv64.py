"""Developer micro-benchmark: the 64 -> 64 3x3 convolution at 256x256, batch 20 (the largest row of the step), alone on
the GPU, L2 flushed by rotating over more activations than the L2 holds."""
import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semantic_pyramid_for_image_generation_b200 import ops as o

B, C, H = 20, 64, 256
cin = int(os.environ.get("CIN", "64"))
xs = [o.to_act(torch.randn(B, H, H, cin, device="cuda")) for _ in range(3)]
w = o.to_act((torch.randn(9, C, cin, device="cuda") / math.sqrt(9 * cin)))
gate = o.to_act(torch.randn(B, H, H, C, device="cuda"))
for mode in ("plain", "act", "gate"):
    kw = {}
    if mode == "act":
        kw = dict(bias=torch.randn(C, device="cuda"), want_raw=False, want_act=True)
    if mode == "gate":
        kw = dict(dmask=gate, dmask_slope=0.2)
    for _ in range(3):
        o.conv(B, H, H, C, [o.Src(xs[0], w, cin, 3)], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for i in range(n):
        o.conv(B, H, H, C, [o.Src(xs[i % 3], w, cin, 3)], **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * B * H * H * C * cin * 9
    print("%-6s %s: %.1f us  %.0f TFLOP/s" % (mode, o.last_conv_kernel(), us, fl / us / 1e6))
