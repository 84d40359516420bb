"""Thin tensor-level wrappers over the C-ABI kernels (allocation + argument marshalling only).

Feature maps are torch tensors of shape (B, H, W, C), dtype bfloat16, contiguous (NHWC).  Vectors, statistics,
losses and master weights are float32.  Nothing here computes on the host or through torch kernels.
"""
import ctypes as C
import os

import torch

from . import _native as N
from ._native import call, ptr

BF16 = torch.bfloat16
F32 = torch.float32
LRELU = 0.2

# ------------------------------------------------------------------------------------------------
# Precision mode (include/spyramid_b200.h, spyr_set_precision).  "bf16": BF16 operands / FP32 accumulate, one plane per
# feature map (the throughput mode).  "split": every BF16 map is a (hi, lo) plane pair and the tensor cores compute
# three products per convolution -- ~16 mantissa bits end to end, the mode in which network-level outputs meet the
# reference's FP32 results within rel-L2 5e-3 (the "strict" mode of the parity tests).  Maps are torch tensors viewing
# the hi plane; the lo plane follows it in the same allocation, which is why every map must come from `act_empty`.
# ------------------------------------------------------------------------------------------------
SPLIT = False


def set_precision(mode):
    """mode: "bf16" (default) or "split" / "strict".  Process-wide; switch only between forward/backward passes."""
    global SPLIT
    mode = {"fast": "bf16", "strict": "split"}.get(mode, mode)
    if mode not in ("bf16", "split"):
        raise ValueError("precision mode must be 'bf16' or 'split' (got %r)" % (mode,))
    rc = N.lib().spyr_set_precision(1 if mode == "split" else 0)
    if rc != 0:
        raise RuntimeError("spyr_set_precision failed: %s" % N.lib().spyr_last_error().decode())
    SPLIT = mode == "split"


def precision():
    return "split" if SPLIT else "bf16"


if os.environ.get("SPYR_PRECISION"):
    set_precision(os.environ["SPYR_PRECISION"])


def act_empty(shape, device):
    """Uninitialised BF16 feature map (or packed operand) of `shape`; in split mode the lo plane follows in the same storage."""
    shape = tuple(int(v) for v in shape)
    if SPLIT:
        return torch.empty((2,) + shape, dtype=BF16, device=device)[0]
    return torch.empty(shape, dtype=BF16, device=device)


def act_zeros(shape, device):
    shape = tuple(int(v) for v in shape)
    if SPLIT:
        return torch.zeros((2,) + shape, dtype=BF16, device=device)[0]
    return torch.zeros(shape, dtype=BF16, device=device)


def act_like(t):
    return act_empty(t.shape, t.device)


def has_planes(t):
    """Whether the storage behind the contiguous BF16 map `t` holds what the current mode needs (split: a lo plane)."""
    if not SPLIT:
        return True
    return t.untyped_storage().nbytes() >= 2 * (t.storage_offset() + 2 * t.numel())


def to_act(values):
    """FP32 tensor -> BF16 map in the current mode (split: hi = bf16(v), lo = bf16(v - hi)).  Host-side packing of frozen
    operands (VGG weights); activations are converted by the kernels."""
    values = values.float().contiguous()
    out = act_empty(values.shape, values.device)
    hi = values.to(BF16)
    out.copy_(hi)
    if SPLIT:
        lo = torch.as_strided(out, out.shape, out.stride(), out.storage_offset() + out.numel())
        lo.copy_((values - hi.float()).to(BF16))
    return out


def act_value(t):
    """FP32 value of a map (hi + lo in split mode): for tests and debugging."""
    v = t.float()
    if SPLIT and has_planes(t):
        v = v + torch.as_strided(t, t.shape, t.stride(), t.storage_offset() + t.numel()).float()
    return v


def scratch(n_outputs, device):
    """Scratch of a deterministic grid reduction over `n_outputs` values (SPYR_REDUCE_SCRATCH_BYTES)."""
    return torch.empty(N.REDUCE_BLOCKS * int(n_outputs) * 8, dtype=torch.uint8, device=device)

# bench.py sets this to a list to time every tensor-core launch with CUDA events on the launching stream:
# entries are (kernel, algorithmic_flops, algorithmic_bytes, start_event, end_event)
PROFILE = None


def _profiled(kernel, flops, nbytes, name, desc, label=""):
    if PROFILE is None:
        call(name, desc)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, desc)
    e1.record()
    real = N.lib().spyr_last_conv_kernel().decode() or kernel  # the kernel the dispatcher actually chose
    PROFILE.append((real, flops, nbytes, e0, e1, label))


def last_conv_kernel():
    """Name of the tensor-core kernel the last conv / wgrad call of this thread launched (spyr_last_conv_kernel)."""
    return N.lib().spyr_last_conv_kernel().decode()


class LeafStream(object):
    """Second CUDA stream for the leaves of a backward pass.

    Weight/bias-gradient kernels (conv wgrad, colsum, stencil wgrad, linear wgrad) consume a gradient map and a saved
    activation and produce a slice of the gradient arenas that nothing reads before the spectral-norm backward at the
    end of the pass.  Run after `begin()`, they are enqueued on a side stream that waits for the main stream's current
    point, so the small-map launches (a few dozen CTAs) overlap the input-gradient chain instead of each taking a turn
    on a mostly idle GPU.  `join()` makes the main stream wait for them.  Inputs are kept alive until `join()` (the
    caching allocator would otherwise hand a freed gradient map to the next main-stream allocation).  Works the same
    under CUDA-graph capture, where the waits become graph edges.  SPYR_LEAF_STREAM=0 disables it."""

    def __init__(self):
        self.enabled = os.environ.get("SPYR_LEAF_STREAM", "1") != "0"
        self.streams = {}
        self.active = None
        self.keep = []

    def begin(self, device):
        if self.active is not None:  # a pass that raised before its join(): close it
            self.join()
        if not self.enabled or PROFILE is not None:
            return False
        # one side stream per parent stream: backward passes that run concurrently on different streams (D(real) next to
        # D(fake), model_wrapper._phase_discriminator) must not queue their leaves behind each other
        key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
        if key not in self.streams:
            self.streams[key] = torch.cuda.Stream(device=device)
        self.active = self.streams[key]
        return True

    def run(self, tensors, name, *args):
        """`tensors`: everything the call reads or writes that must outlive the enqueue (inputs, scratch)."""
        if self.active is None:
            call(name, *args)
            return
        side = self.active
        side.wait_stream(torch.cuda.current_stream())
        self.keep.extend(tensors)
        with torch.cuda.stream(side):
            call(name, *args)

    def join(self):
        if self.active is None:
            return
        torch.cuda.current_stream().wait_stream(self.active)
        self.active = None
        del self.keep[:]


LEAF = LeafStream()


class DeferredReductions(object):
    """Second stages (fixed-order sums of partial results) of the weight / bias gradients of ONE backward pass, batched.

    Between `begin()` and `flush()`, `wgrad`, `colsum` and `stencil_wgrad` launch only their first stage and leave a
    description of the pending sum here; `flush()` (after the leaf stream has joined) performs all of them in one launch per
    32 entries instead of ~90 small kernels.  Same arithmetic in the same order as the immediate form: bit-identical.
    SPYR_DEFER_REDUCE=0 disables it."""

    def __init__(self):
        self.enabled = os.environ.get("SPYR_DEFER_REDUCE", "1") != "0"
        self.entries = None
        self.keep = []

    def begin(self):
        if self.entries is not None:
            self.flush()
        if self.enabled and PROFILE is None:
            self.entries = []

    def active(self):
        return self.entries is not None

    def add(self, entry, *tensors):
        if entry.kind != 0:
            self.entries.append(entry)
        self.keep.extend(t for t in tensors if t is not None)

    def flush(self):
        entries, self.entries = self.entries, None
        if entries:
            for i in range(0, len(entries), N.REDUCE_BATCH):
                chunk = entries[i:i + N.REDUCE_BATCH]
                arr = (N.ReduceEntry * len(chunk))(*chunk)
                call("spyr_reduce_batched", arr, len(chunk))
        del self.keep[:]


DEFER = DeferredReductions()


def empty_bf16(*shape, device=None):
    return act_empty(shape, device or "cuda")


def empty_f32(*shape, device=None):
    return torch.empty(shape, dtype=F32, device=device or "cuda")


def zeros_f32(*shape, device=None):
    return torch.zeros(shape, dtype=F32, device=device or "cuda")


class Src(object):
    """One accumulation source of a tensor-core convolution (see spyr_conv_src)."""
    __slots__ = ("x", "w", "cin", "ksize", "mn", "per_image", "w_lo_off")

    def __init__(self, x, w, cin, ksize, mn=False, per_image=False, w_lo_off=0):
        self.x, self.w, self.cin, self.ksize, self.mn, self.per_image = x, w, cin, ksize, mn, per_image
        self.w_lo_off = w_lo_off  # split mode: elements from w to its lo plane when not slices * Cout * cin


def _run_fprop(d, srcs, B, H, W, Cout, out_bytes):
    if PROFILE is None:
        call("spyr_conv2d_fprop", C.byref(d))
        return
    kred = sum(s.cin * s.ksize * s.ksize for s in srcs)
    npx = B * H * W
    nbytes = 2 * npx * sum(s.cin for s in srcs) + 2 * Cout * kred + out_bytes
    label = "B%d %dx%d Cout=%d %s" % (B, H, W, Cout, "+".join(
        "%dk%d%s%s" % (s.cin, s.ksize, "T" if s.mn else "", "b" if s.per_image else "") for s in srcs))
    _profiled("conv_fprop_kernel", 2.0 * npx * Cout * kred, float(nbytes), "spyr_conv2d_fprop", C.byref(d), label)


def conv(B, H, W, Cout, srcs, bias=None, bias2=None, bias3=None, stencil_mask=None, stencil_w=None, dmask=None, dmask_slope=1.0, residual=None,
         want_raw=True, want_act=False, act=2, act_slope=LRELU, f32_out=None, f32_store=False, splits=1, device=None, pool=False,
         residual_pooled=False):
    """out = sum_src conv(src) (+bias, mask stencil, gate, residual).  Returns (y_raw, y_act) (None when not asked).

    pool=True: 2x2 average pool fused into the epilogue -- outputs (and `residual`) are (B, H/2, W/2, Cout); callers use
    `can_pool(H, W, Cout)` first (small maps and ragged widths keep the separate pooling kernel).
    residual_pooled=True: `residual` is (B, H/2, W/2, Cout) and enters as 0.25 * residual[h/2][w/2] (same condition).

    `w` of a source may be a tensor or an int device address (a slice of a packed-weight arena)."""
    dev = device or srcs[0].x.device
    if Cout == 64 and len(srcs) == 1 and srcs[0].mn and srcs[0].ksize == 3 and not srcs[0].per_image and H >= 16 \
            and f32_out is None and os.environ.get("SPYR_DGRAD_KMAJOR", "1") != "0":
        # 64-wide input gradient: re-lay the (tiny) forward weights as a K-major operand with flipped taps so the layer
        # runs on the CTA-pair kernel (its MN-major form is limited to the single-CTA kernel)
        s0 = srcs[0]
        wt = act_empty((9, Cout, s0.cin), dev)
        call("spyr_weight_transpose_flip", s0.w if isinstance(s0.w, int) else s0.w.data_ptr(), wt.data_ptr(), 9, s0.cin, Cout)
        srcs = [Src(s0.x, wt, s0.cin, 3)]
    d = N.ConvDesc()
    d.B, d.H, d.W, d.Cout, d.nsrc = B, H, W, Cout, len(srcs)
    for i, s in enumerate(srcs):
        d.src[i].x = s.x.data_ptr()
        d.src[i].w = s.w if isinstance(s.w, int) else s.w.data_ptr()
        d.src[i].cin, d.src[i].ksize = s.cin, s.ksize
        d.src[i].w_mn_major, d.src[i].w_per_image = int(s.mn), int(s.per_image)
        d.src[i].w_lo_off = int(s.w_lo_off)
    d.bias = bias if isinstance(bias, int) else ptr(bias)
    d.bias2, d.bias3 = ptr(bias2), ptr(bias3)
    d.stencil_mask = ptr(stencil_mask)
    d.stencil_w = stencil_w if isinstance(stencil_w, int) else ptr(stencil_w)
    d.dmask, d.dmask_slope = ptr(dmask), dmask_slope
    d.residual = ptr(residual)
    d.residual_pooled = int(residual_pooled)
    y_raw = y_act = None
    npix = B * H * W
    if f32_out is None and H * W < 128 and Cout % 8 == 0 and not any(s.per_image for s in srcs):
        # 4x4 / 8x8 maps: a handful of output tiles cannot fill 148 SMs -> split the reduction over CTAs into an FP32
        # accumulator and apply the fused epilogue in a second, tiny kernel
        ksteps = sum(((s.cin + 63) // 64) * s.ksize * s.ksize for s in srcs)
        tiles = ((npix + 127) // 128) * ((Cout + 255) // 256)
        if SPLIT:
            ksteps *= 3  # three operand products per source
        nsplit = max(1, min(ksteps // 4, 148 // tiles))
        if nsplit > 1:
            # one FP32 slice per split, summed in split order by the epilogue kernel (no atomics: reproducible)
            acc = torch.empty((nsplit, npix, Cout), dtype=F32, device=dev)
            d.y_f32, d.splits = acc.data_ptr(), nsplit
            _run_fprop(d, srcs, B, H, W, Cout, 4 * npix * Cout)
            d.y_f32 = None
            if want_raw:
                y_raw = act_empty((B, H, W, Cout), dev)
                d.y_raw = y_raw.data_ptr()
            if want_act:
                y_act = act_empty((B, H, W, Cout), dev)
                d.y_act = y_act.data_ptr()
            d.act, d.act_slope = act, act_slope
            call("spyr_conv2d_epilogue", C.byref(d), acc.data_ptr())
            return y_raw, y_act
    if f32_out is not None:
        d.y_f32, d.f32_store, d.splits = f32_out.data_ptr(), int(f32_store), splits
    else:
        oh, ow = (H // 2, W // 2) if pool else (H, W)
        d.pool = int(pool)
        if want_raw:
            y_raw = act_empty((B, oh, ow, Cout), dev)
            d.y_raw = y_raw.data_ptr()
        if want_act:
            y_act = act_empty((B, oh, ow, Cout), dev)
            d.y_act = y_act.data_ptr()
        d.act, d.act_slope = act, act_slope
    out_bytes = (4 if f32_out is not None else 2) * npix * Cout * (int(want_raw) + int(want_act) if f32_out is None else 1)
    _run_fprop(d, srcs, B, H, W, Cout, out_bytes)
    return y_raw, y_act


def can_pool(H, W, Cout):
    """Whether conv(..., pool=True) applies: a halo-tiled map (>= 16x8, not the split-K small maps) and full 32-channel
    epilogue chunks."""
    return H >= 16 and W >= 8 and H % 16 == 0 and W % 8 == 0 and H * W >= 128 and Cout % 32 == 0 \
        and os.environ.get("SPYR_POOL_FUSION", "1") != "0" and not SPLIT


def wgrad(x, dy, dw_ptr, B, H, W, Cin, Cout, ksize, cin_stride=0, per_image=False):
    """dw[tap][ci][co] += sum_pixels x[p+tap][ci] * dy[p][co]  (FP32 accumulate into a zero-filled arena slice)."""
    d = N.WgradDesc()
    d.B, d.H, d.W, d.Cin, d.Cout, d.ksize = B, H, W, Cin, Cout, ksize
    d.x, d.dy, d.dw = x.data_ptr(), dy.data_ptr(), dw_ptr
    d.cin_stride, d.per_image = cin_stride, int(per_image)
    # partial-sum slices of the pixel splits (summed in a fixed order by a second kernel: no atomics)
    need = int(N.lib().spyr_conv2d_wgrad_scratch_floats(C.byref(d)))
    if need < 0:
        raise RuntimeError("spyr_conv2d_wgrad_scratch_floats: %s" % N.lib().spyr_last_error().decode())
    sc = torch.empty(need, dtype=F32, device=x.device) if need > 0 else None
    d.scratch, d.scratch_floats = ptr(sc), need
    if PROFILE is None:
        if per_image:
            call("spyr_conv2d_wgrad", C.byref(d))  # attention dK / dV feed the rest of the backward: not a leaf
        elif DEFER.active():
            entry = N.ReduceEntry()
            LEAF.run((x, dy, sc), "spyr_conv2d_wgrad_deferred", C.byref(d), C.byref(entry))
            DEFER.add(entry, sc)
        else:
            LEAF.run((x, dy, sc), "spyr_conv2d_wgrad", C.byref(d))
    else:
        npx = B * H * W
        _profiled("conv_wgrad_kernel", 2.0 * npx * Cin * Cout * ksize * ksize,
                  float(2 * npx * (Cin + Cout) + 4 * Cin * Cout * ksize * ksize), "spyr_conv2d_wgrad", C.byref(d),
                  "B%d %dx%d Cin=%d Cout=%d k%d%s" % (B, H, W, Cin, Cout, ksize, "b" if per_image else ""))


def colsum(g, C_, out0, out1=None, out2=None):
    rows = g.numel() // C_
    sc = scratch(C_, g.device)
    if DEFER.active():
        entry = N.ReduceEntry()
        LEAF.run((g, sc), "spyr_colsum_deferred", g.data_ptr(), rows, C_, out0, out1, out2, sc.data_ptr(), C.byref(entry))
        DEFER.add(entry, sc)
        return
    LEAF.run((g, sc), "spyr_colsum", g.data_ptr(), rows, C_, out0, out1, out2, sc.data_ptr())


def stencil_wgrad(mask, dy, B, H, W, Cout, gw_ptr, cin_stride, cin_index):
    """Weight gradient of the mask channel of cat(feature*mask, mask) (models.py:94): nine masked sums of dy."""
    sc = scratch(9 * Cout, dy.device)
    if DEFER.active():
        entry = N.ReduceEntry()
        LEAF.run((mask, dy, sc), "spyr_stencil_wgrad_deferred", mask.data_ptr(), dy.data_ptr(), B, H, W, Cout, gw_ptr,
                 cin_stride, cin_index, sc.data_ptr(), C.byref(entry))
        DEFER.add(entry, sc)
        return
    LEAF.run((mask, dy, sc), "spyr_stencil_wgrad", mask.data_ptr(), dy.data_ptr(), B, H, W, Cout, gw_ptr, cin_stride,
             cin_index, sc.data_ptr())


def maskgate(f, mask):
    out = act_like(f)
    call("spyr_maskgate", f.data_ptr(), mask.data_ptr(), out.data_ptr(), f.numel() // f.shape[-1], f.shape[-1])
    return out


def nchw_to_nhwc(src, mask=None, slope=1.0):
    B, Cc, H, W = src.shape
    out = act_empty((B, H, W, Cc), src.device)
    call("spyr_nchw_to_nhwc", src.data_ptr(), ptr(mask), slope, out.data_ptr(), B, Cc, H * W)
    return out


def nhwc_to_nchw(src, gate_x=None, slope=1.0):
    B, H, W, Cc = src.shape
    out = torch.empty((B, Cc, H, W), dtype=F32, device=src.device)
    call("spyr_nhwc_to_nchw", src.data_ptr(), ptr(gate_x), slope, out.data_ptr(), B, Cc, H * W)
    return out


def as_nhwc_bf16(t, mask=None):
    """Accepts an NCHW-shaped tensor: a permuted view of NHWC BF16 storage is used in place, anything else is
    converted (FP32 NCHW -> BF16 NHWC).  The optional (B,1,H,W) mask gates the feature (models.py:94)."""
    if t.dtype == BF16 and t.dim() == 4:
        v = t.permute(0, 2, 3, 1)
        if v.is_contiguous() and has_planes(v):  # split mode: only maps of this package carry their lo plane
            return maskgate(v, mask) if mask is not None else v
    if t.dtype != F32 or not t.is_contiguous():
        t = t.float().contiguous()
    return nchw_to_nhwc(t, mask)


def avgpool2(x, residual=None, want_raw=True, want_act=False, slope=LRELU):
    B, H, W, Cc = x.shape
    y_raw = act_empty((B, H // 2, W // 2, Cc), x.device) if want_raw else None
    y_act = act_empty((B, H // 2, W // 2, Cc), x.device) if want_act else None
    call("spyr_avgpool2_fwd", x.data_ptr(), ptr(residual), ptr(y_raw), ptr(y_act), slope, B, H, W, Cc)
    return y_raw, y_act


def avgpool2_bwd(g_lo):
    B, h, w, Cc = g_lo.shape
    g_hi = act_empty((B, 2 * h, 2 * w, Cc), g_lo.device)
    call("spyr_avgpool2_bwd", g_lo.data_ptr(), g_hi.data_ptr(), B, 2 * h, 2 * w, Cc)
    return g_hi


def maxpool2(x):
    B, H, W, Cc = x.shape
    y = act_empty((B, H // 2, W // 2, Cc), x.device)
    call("spyr_maxpool2_fwd", x.data_ptr(), y.data_ptr(), B, H, W, Cc)
    return y


def maxpool2_bwd(x, gy, relu_gate, out=None):
    B, H, W, Cc = x.shape
    acc = out is not None
    gx = out if acc else act_like(x)
    call("spyr_maxpool2_bwd", x.data_ptr(), gy.data_ptr(), gx.data_ptr(), B, H, W, Cc, int(relu_gate), int(acc))
    return gx


def bn_stats(x, up2=False):
    B, H, W, Cc = x.shape
    sums = scratch(2 * Cc, x.device)  # per-block partial sums (FP64); bn_finalize adds them in block order
    call("spyr_bn_stats", x.data_ptr(), B, H, W, Cc, int(up2), sums.data_ptr())
    return sums


def up2_stats(x):
    """xu = up2(x) (bilinear, align_corners=True) materialised in BF16 together with its per-channel sum / sum of squares."""
    B, H, W, Cc = x.shape
    xu = act_empty((B, 2 * H, 2 * W, Cc), x.device)
    sums = scratch(2 * Cc, x.device)
    call("spyr_up2_stats", x.data_ptr(), B, H, W, Cc, xu.data_ptr(), sums.data_ptr())
    return xu, sums


def bn_stats_blocks(npix, Cc):
    """Number of partial vectors spyr_bn_stats writes for `npix` pixels (mirrors bn_stats_blocks in csrc/norm.cu)."""
    prows = max(1, 256 // (Cc // 8))
    return max(1, min((npix + prows * 16 - 1) // (prows * 16), N.REDUCE_BLOCKS))


def bn_sums(partials, npix, Cc):
    """(sum, sum of squares) per channel from the partials of bn_stats / up2_stats, as FP64 [2C] (tests, debugging)."""
    nb = bn_stats_blocks(npix, Cc)
    return partials.view(torch.float64)[:nb * 2 * Cc].view(nb, 2 * Cc).sum(dim=0)


def bn_bwd_partials(B, npix_per_image, Cc, device):
    """Buffer for the block partials of spyr_bn_bwd_reduce ([B][gx][2][C] FP32; gx mirrors bn_bwd_blocks in csrc/norm.cu)."""
    prows = max(1, 256 // (Cc // 8))
    gx = max(1, min((npix_per_image + prows * 8 - 1) // (prows * 8), (4 * N.REDUCE_BLOCKS) // B))
    return torch.empty(B * (gx + 1) * 2 * Cc, dtype=F32, device=device)  # [B][2][C] sums + [B][gx][2][C] partials


def bn_finalize(sums, count, Cc, eps, momentum, running_mean, running_var, nbt, training):
    mean_rstd = torch.empty(2 * Cc, dtype=F32, device=running_mean.device if running_mean is not None else sums.device)
    call("spyr_bn_finalize", ptr(sums), float(count), Cc, eps, momentum, ptr(running_mean), ptr(running_var), ptr(nbt),
         mean_rstd.data_ptr(), int(training))
    return mean_rstd


def bn_act(x, mean_rstd, scale_ptr, shift_ptr, row_stride, cls, mode, want_xu=False, slope=LRELU):
    B, H, W, Cc = x.shape
    f = 2 if mode else 1
    a = act_empty((B, H * f, W * f, Cc), x.device)
    xu = act_empty((B, H * f, W * f, Cc), x.device) if want_xu else None
    if PROFILE is None:
        call("spyr_bn_act", x.data_ptr(), mean_rstd.data_ptr(), scale_ptr, shift_ptr, row_stride, ptr(cls), slope, mode,
             a.data_ptr(), ptr(xu), B, H, W, Cc)
    else:
        # bandwidth-bound pass: algorithmic bytes = x read once + every output written once (per plane)
        planes = 2 if SPLIT else 1
        nbytes = 2 * planes * (x.numel() + a.numel() + (xu.numel() if xu is not None else 0))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call("spyr_bn_act", x.data_ptr(), mean_rstd.data_ptr(), scale_ptr, shift_ptr, row_stride, ptr(cls), slope, mode,
             a.data_ptr(), ptr(xu), B, H, W, Cc)
        e1.record()
        PROFILE.append(("bn_act_kernel<%d>" % mode, 0.0, float(nbytes), e0, e1, "B%d %dx%d C=%d" % (B, H, W, Cc)))
    return a, xu


def linear_fwd(x, w, sigma_ptr, bias, xmask=None, in_slope=1.0, y_add=None, out_slope=1.0):
    B, K = x.shape
    O = w.shape[0]
    y = torch.empty((B, O), dtype=F32, device=x.device)
    call("spyr_linear_fwd", x.data_ptr(), ptr(xmask), in_slope, w.data_ptr(), sigma_ptr, ptr(bias), ptr(y_add), out_slope,
         y.data_ptr(), B, K, O)
    return y


def linear_bwd_x(gy, w, sigma_ptr, y=None, out_slope=1.0, x=None, in_slope=1.0, out=None):
    B, O = gy.shape
    K = w.shape[1]
    acc = out is not None
    gx = out if acc else torch.empty((B, K), dtype=F32, device=gy.device)
    call("spyr_linear_bwd_x", gy.data_ptr(), ptr(y), out_slope, w.data_ptr(), sigma_ptr, ptr(x), in_slope, gx.data_ptr(),
         int(acc), B, K, O)
    return gx


def linear_bwd_w(gy, x, gw_ptr, gb_ptr, y=None, out_slope=1.0, xmask=None, in_slope=1.0):
    B, O = gy.shape
    K = x.shape[1]
    LEAF.run((gy, x, y, xmask), "spyr_linear_bwd_w", gy.data_ptr(), ptr(y), out_slope, x.data_ptr(), ptr(xmask), in_slope,
             gw_ptr, gb_ptr, B, K, O)


def argmax_rows(onehot):
    B, n = onehot.shape
    if onehot.dtype not in (F32, torch.int64):
        onehot = onehot.float()
    onehot = onehot.contiguous()
    out = torch.empty(B, dtype=torch.int32, device=onehot.device)
    call("spyr_argmax_rows", onehot.data_ptr(), int(onehot.dtype == torch.int64), B, n, out.data_ptr())
    return out


def cast_bf16(src):
    out = act_empty(src.shape, src.device)
    call("spyr_cast_f32_bf16", src.data_ptr(), out.data_ptr(), src.numel())
    return out
