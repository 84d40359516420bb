"""One process per GPU data parallelism: replaces the reference's nn.DataParallel wrap (main.py:91-94).

Every rank holds a full replica and its own 20-image shard; batch-norm statistics and the (B,B,128) discriminator
cross terms stay per replica exactly as under nn.DataParallel.  The only exchange is one gradient
all-reduce per optimizer step.  Because each model's parameter gradients live in ONE flat FP32 arena
(engine.GradArena), the exchange is a handful of large NCCL calls over NVLink/NVSwitch instead of ~170 small ones;
it runs on a side stream so the caller can overlap it with independent work (`average(..., wait=False)`).
"""
import os
from typing import List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> "GradientReducer":
    """Initialises torch.distributed from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT / LOCAL_RANK (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    return GradientReducer()


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `total` samples for `rank` (sizes differ by at most one)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GradientReducer(object):
    def __init__(self, group=None, bucket_bytes: int = 64 << 20):
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.world = dist.get_world_size(group) if self.active else 1
        self.bucket_bytes = bucket_bytes
        self._stream = None
        self._pending: List = []

    def _buckets(self, flat: torch.Tensor):
        step = max(1, self.bucket_bytes // flat.element_size())
        return [flat[i:i + step] for i in range(0, flat.numel(), step)]

    def average_flat(self, flat: torch.Tensor, wait: bool = True) -> None:
        """In-place mean over ranks of a flat tensor, bucketed; NCCL averages in the collective, gloo sums then scales."""
        if not self.active:
            return
        nccl = flat.is_cuda
        if nccl:
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                for b in self._buckets(flat):
                    dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self.group)
            flat.record_stream(self._stream)
            if wait:
                self.wait()
        else:
            for b in self._buckets(flat):
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)

    def wait(self) -> None:
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)

    @staticmethod
    def grads_alias_arena(module, arena) -> bool:
        """True when every existing `.grad` of `module` is a view into `arena` at its GradArena offset.  That is what
        autograd leaves behind after a backward from zeroed (set_to_none) gradients; it does NOT hold after
        `zero_grad(set_to_none=False)`, gradient accumulation over several backwards, or a gradient hook that replaces
        the tensor -- then the arena is a dead buffer and must not be what gets averaged."""
        ga = getattr(module, "_ga", None)
        if ga is None or arena is None:
            return False
        base = arena.data_ptr()
        seen = False
        for p, off in zip(ga.params, ga.offset_list):
            if p.grad is None:
                continue
            seen = True
            if p.grad.data_ptr() != base + 4 * off or p.grad.dtype != arena.dtype:
                return False
        return seen

    def average(self, module, wait: bool = True) -> None:
        """Averages the gradients of `module` over the ranks: ONE collective sequence over the flat arena of the last
        backward when the `.grad` tensors are views of it (checked), else one all-reduce per gradient tensor."""
        if not self.active:
            return
        arena = getattr(module, "_last_grad_arena", None)
        if arena is not None and self.grads_alias_arena(module, arena):
            self.average_flat(arena, wait=wait)
            return
        for p in module.parameters():
            if p.grad is not None:
                self.average_flat(p.grad.view(-1), wait=wait)

    def broadcast_module(self, module, src: int = 0) -> None:
        """Parameters and buffers of rank `src` to every rank (what nn.DataParallel's per-forward `replicate` guaranteed,
        main.py:91-94): replicas then start identical whatever seed each process used."""
        if not self.active:
            return
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=self.group)

    def max_over_ranks(self, value: float, device=None) -> float:
        if not self.active:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if torch.cuda.is_available() else "cpu"))
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def barrier(self) -> None:
        if self.active:
            dist.barrier(group=self.group)
