"""Batched spectral normalisation state for one model (Generator or Discriminator).

Host-side mirror of torch.nn.utils.spectral_norm (torch/nn/utils/spectral_norm.py:92-114) as used at every
Conv2d / Linear / Embedding of the reference (models.py:28,34,55,58,128,132,135,232-243,299-313,356-360,393-403,
438-448): the parameters stay `weight_orig` / `weight_u` / `weight_v` with identical shapes and state_dict keys,
while the arithmetic (power iteration, sigma, W/sigma packed to BF16, backward through sigma) runs in
csrc/sn.cu for all layers of the model at once.
"""
import ctypes as C
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N
from ._native import call, call_nostream


class SpectralNormHolder(nn.Module):
    """Parameter/buffer container with the state_dict layout of a spectral-normed layer
    (`bias`, `weight_orig`, `weight_u`, `weight_v`).  init: Xavier-uniform weights, zero bias (models.py:509-519,
    which does reach `weight_orig`) or N(0,1) for nn.Embedding."""

    def __init__(self, *shape, bias=True, init="xavier"):
        super().__init__()
        if bias:
            self.bias = nn.Parameter(torch.zeros(shape[0]))
        else:
            self.register_parameter("bias", None)
        w = torch.empty(*shape)
        if init == "xavier":
            nn.init.xavier_uniform_(w)
        else:
            nn.init.normal_(w)
        self.weight_orig = nn.Parameter(w)
        rows = shape[0]
        cols = w.numel() // rows
        self.register_buffer("weight_u", F.normalize(torch.randn(rows), dim=0, eps=1e-12))
        self.register_buffer("weight_v", F.normalize(torch.randn(cols), dim=0, eps=1e-12))
        self.shape = tuple(shape)

    def extra_repr(self):
        return "spectral_norm, weight shape %s, bias=%s" % (self.shape, self.bias is not None)


class LayerSpec(object):
    __slots__ = ("key", "holder", "pack_cin", "pack_mode", "stencil", "index", "pack_off", "stencil_off", "gw_off",
                 "gw_layout", "rows", "cols", "taps", "cin", "saved_off")

    def __init__(self, key, holder, pack_cin=0, pack_mode=0, stencil=False, gw_layout=None):
        self.key, self.holder = key, holder
        self.pack_cin, self.pack_mode, self.stencil = pack_cin, pack_mode, stencil
        shape = holder.shape
        self.rows = shape[0]
        self.cin = shape[1]
        self.taps = 1
        for s in shape[2:]:
            self.taps *= s
        self.cols = self.cin * self.taps
        # gradients of tensor-core layers arrive in the wgrad kernel's [tap][cin][cout] order
        self.gw_layout = (1 if pack_cin > 0 else 0) if gw_layout is None else gw_layout


class SNState(object):
    """Per-forward outputs: packed BF16 weights, mask-channel stencils, (sigma, u, v) clones."""

    def __init__(self, owner, packed, stencil, saved):
        self.owner, self.packed, self.stencil, self.saved = owner, packed, stencil, saved
        self._pbase = packed.data_ptr() if packed is not None else 0
        self._sbase = stencil.data_ptr() if stencil is not None else 0
        self._vbase = saved.data_ptr()

    def w(self, key):
        return self._pbase + 2 * self.owner.by_key[key].pack_off

    def stencil_w(self, key):
        return self._sbase + 4 * self.owner.by_key[key].stencil_off

    def sigma(self, key):
        return self._vbase + 4 * self.owner.by_key[key].saved_off


class SNSet(object):
    def __init__(self, specs, param_offsets):
        """specs: list of LayerSpec in forward order; param_offsets(param) -> float offset in the grad arena."""
        self.specs = specs
        self.by_key = {s.key: s for s in specs}
        self.param_offsets = param_offsets
        self.n = len(specs)
        sten = gw = 0
        for s in specs:
            s.pack_off = 0  # assigned per precision mode in _ensure_table
            if s.stencil:
                s.stencil_off = sten
                sten += 10 * s.rows
            else:
                s.stencil_off = -1
            s.gw_off = gw  # 16-byte aligned slices: the wgrad kernels store float4 vectors
            gw += (s.rows * s.cols + 3) // 4 * 4
        self.pack_elems, self.stencil_floats, self.gw_floats = 0, sten, gw
        self._ptr_sig = None
        self.host_tab = None
        self.dev_tab = None
        self.plan = N.SnPlan()
        self.scratch = None

    def __deepcopy__(self, memo):
        # the device/host tables hold raw pointers of the original module: a copy plans its own
        import copy
        return SNSet(copy.deepcopy(self.specs, memo), copy.deepcopy(self.param_offsets, memo))

    def _signature(self):
        from . import ops
        return tuple(s.holder.weight_orig.data_ptr() for s in self.specs) + tuple(
            s.holder.weight_u.data_ptr() for s in self.specs) + (ops.SPLIT,)

    def _ensure_table(self):
        sig = self._signature()
        if sig == self._ptr_sig:
            return
        from . import ops
        planes = 2 if ops.SPLIT else 1  # split-BF16 mode: the lo plane of a layer's pack follows its hi plane
        pack = 0
        for s in self.specs:
            if s.pack_cin > 0:
                s.pack_off = pack
                per = s.rows * s.pack_cin * (1 if s.pack_mode == 1 else s.taps)
                pack += (planes * per + 63) // 64 * 64
        self.pack_elems = pack
        tab = (N.SnLayer * self.n)()
        for i, s in enumerate(self.specs):
            h = s.holder
            if not (h.weight_orig.is_cuda and h.weight_orig.dtype == torch.float32 and h.weight_orig.is_contiguous()):
                raise RuntimeError("spectral-normed layer %s must hold contiguous float32 CUDA parameters" % s.key)
            L = tab[i]
            L.w, L.u, L.v = h.weight_orig.data_ptr(), h.weight_u.data_ptr(), h.weight_v.data_ptr()
            L.rows, L.cols, L.taps, L.cin = s.rows, s.cols, s.taps, s.cin
            L.pack_cin, L.pack_mode, L.pack_off, L.stencil_off = s.pack_cin, s.pack_mode, s.pack_off, s.stencil_off
            L.gw_off, L.gw_layout = s.gw_off, s.gw_layout
            L.grad_off = self.param_offsets(h.weight_orig)
        call_nostream("spyr_sn_plan", tab, self.n, C.byref(self.plan))
        for i, s in enumerate(self.specs):
            s.saved_off = tab[i].saved_off
        dev = self.specs[0].holder.weight_orig.device
        raw = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8)
        self.dev_tab = raw.to(dev)
        self.host_tab = tab
        self.scratch = torch.empty(max(int(self.plan.scratch_floats), 1), dtype=torch.float32, device=dev)
        self.device = dev
        self._ptr_sig = sig

    def forward(self, training):
        """One power iteration (training) / sigma only (eval) for every layer; returns the SNState."""
        self._ensure_table()
        dev = self.device
        packed = torch.empty(max(self.pack_elems, 8), dtype=torch.bfloat16, device=dev)
        stencil = torch.empty(max(self.stencil_floats, 1), dtype=torch.float32, device=dev)
        saved = torch.empty(int(self.plan.saved_floats), dtype=torch.float32, device=dev)
        call("spyr_sn_forward", self.dev_tab.data_ptr(), self.n, C.byref(self.plan), int(training), 1e-12,
             self.scratch.data_ptr(), packed.data_ptr(), stencil.data_ptr(), saved.data_ptr())
        return SNState(self, packed, stencil, saved)

    def backward(self, state, gw_arena, grad_arena):
        """dL/dweight_orig for every layer from dL/d(W/sigma) (gw_arena) into grad_arena."""
        dots = torch.empty(max(int(self.plan.tiles_bwd), 1), dtype=torch.float32, device=self.device)
        call("spyr_sn_backward", self.dev_tab.data_ptr(), self.n, C.byref(self.plan), gw_arena.data_ptr(),
             state.saved.data_ptr(), dots.data_ptr(), grad_arena.data_ptr())

    def gw_ptr(self, gw_arena, key):
        return gw_arena.data_ptr() + 4 * self.by_key[key].gw_off


def xavier_bound(shape):
    rf = 1
    for s in shape[2:]:
        rf *= s
    return math.sqrt(6.0 / ((shape[0] + shape[1]) * rf))
