"""Fused multi-tensor Adam with torch.optim.Adam's default semantics (reference main.py:64-65:
`torch.optim.Adam(params, lr)`, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).

The step counter lives on the device and every launch goes to the current stream, so `step()` can be captured in a
CUDA graph.  `state_dict()` uses torch.optim.Adam's keys (`step`, `exp_avg`, `exp_avg_sq`) so checkpoints written by
either optimizer load into the other (model_wrapper.py:215-223).
"""
import ctypes as C

import torch

from . import _native as N
from ._native import call


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        defaults = dict(lr=lr, betas=betas, eps=eps)
        super().__init__(params, defaults)
        self._steps = {}  # id(group) -> device int32 counter

    def _init_state(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["step"] = torch.zeros((), dtype=torch.float32)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            # the kernel writes through raw pointers: tell autograd / version-keyed caches (VGG16._pack) that the
            # parameters changed.  Under CUDA-graph replay this Python code does not run; consumers that cache derived
            # operands across replays must not rely on versions (the GAN path re-reads weight_orig every forward).
            torch._C._increment_version(params)
            dev = params[0].device
            counter = self._steps.get(gi)
            if counter is None or counter.device != dev:
                first = self.state.get(params[0], {})
                start = int(first["step"]) if "step" in first else 0
                counter = torch.full((1,), start, dtype=torch.int32, device=dev)
                self._steps[gi] = counter
            call("spyr_adam_tick", counter.data_ptr())
            b1, b2 = group["betas"]
            chunk = N.AdamChunk()
            k = 0
            for p in params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam needs contiguous float32 CUDA parameters")
                g = p.grad
                if not g.is_contiguous() or g.dtype != torch.float32:
                    raise RuntimeError("FusedAdam needs contiguous float32 gradients")
                st = self._init_state(p)
                chunk.p[k], chunk.g[k] = p.data_ptr(), g.data_ptr()
                chunk.m[k], chunk.v[k] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                chunk.n[k] = p.numel()
                k += 1
                if k == N.ADAM_MAX_TENSORS:
                    chunk.count = k
                    call("spyr_adam_step", C.byref(chunk), counter.data_ptr(), group["lr"], b1, b2, group["eps"])
                    k = 0
            if k:
                chunk.count = k
                call("spyr_adam_step", C.byref(chunk), counter.data_ptr(), group["lr"], b1, b2, group["eps"])
        return loss

    def state_dict(self):
        # refresh the per-parameter `step` entries (torch.optim.Adam layout) from the device counters
        for gi, group in enumerate(self.param_groups):
            counter = self._steps.get(gi)
            if counter is None:
                continue
            steps = float(counter.item())
            for p in group["params"]:
                if p in self.state:
                    self.state[p]["step"] = torch.tensor(steps, dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._steps = {}
