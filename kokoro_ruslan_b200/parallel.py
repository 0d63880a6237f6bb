"""Data-parallel plumbing (new functionality: the reference is single-process).

One process per GPU.  Utterance batches shard naturally, so the only collective of a step is ONE
all-reduce(SUM) of the flat fp32 gradient buffer after the backward pass (197.7 MB over NVLink 5 /
NVSwitch via NCCL); the per-rank loss gradients are pre-scaled by 1 / world_size inside the fused
loss kernel, which makes the reduced buffer the mean of the per-rank (per-batch-mean) gradients —
the same semantics as the reference's gradient accumulation (trainer.py:2284-2294).  Grad-norm,
clipping, explosion detection, AdamW, EMA and the weight-norm projection then run replicated on the
reduced buffer, so every rank takes identical decisions without further communication.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; initialises the default process
    group when WORLD_SIZE > 1 (NCCL on GPUs, gloo on CPU-only hosts)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kwargs)
    return rank, local, world


def all_reduce_gradients(flat_grads: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the flat gradient buffer (the single collective of a step)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return flat_grads


def symmetric_memory_usable(group, device) -> bool:
    """True on every rank iff every rank can allocate and rendezvous a symmetric-memory buffer (collective call)."""
    ok = 1
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(1024, dtype=torch.float32, device=device)
        symm.rendezvous(t, group)
    except Exception:        # noqa: BLE001 — any failure means "not usable here"
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(int(flag.item()))


class SymmetricGradReducer:
    """The step's collective as ONE hand-written kernel over NVLink / NVSwitch peer memory
    (csrc/kr_comm.cu): all-reduce of the flat gradient buffer fused with the optimizer's per-chunk
    squared-norm pass.  torch.distributed's symmetric-memory allocator is used for the plumbing only
    (same-layout allocations on every rank, peer / multicast mappings, exchange of the handles);
    the data path has no NCCL call.

    The gradient buffer of `store` and the optimizer's per-chunk sum buffer are re-homed into
    symmetric memory; `reduce()` enqueues the kernel on the current stream (graph-capturable
    plumbing is not needed: it is launched between the two step graphs)."""

    # measured on 2 / 4 / 8 x B200 (tools/ar_bench.py, 197.7 MB): the multicast path is fastest with about one block
    # per SM (8 GPUs: 0.44 ms vs NCCL 0.55 ms and 0.57 ms for peer loads); with 2 GPUs the switch cannot save
    # traffic and plain peer loads with 4 blocks per SM win (0.34 ms vs NCCL 0.40 ms, multicast 0.51 ms)
    GRID_MULTICAST = 148
    GRID_PEER = 592

    def __init__(self, store, opt, group, use_multicast: Optional[bool] = None):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        self.store, self.opt, self.group = store, opt, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("SymmetricGradReducer: one NVSwitch domain (<= 8 GPUs)")
        dev = store.device
        self.grads = symm.empty(store.total, dtype=torch.float32, device=dev)
        # + one slot per rank behind the chunk sums: the ranks' clip norms of the step (global MIN, csrc/kr_comm.cu)
        self.sq_chunk = symm.empty(opt.n_chunks + 8, dtype=torch.float32, device=dev)
        self.flags = symm.empty(2048 * self.world, dtype=torch.int32, device=dev)      # room for grids up to 2048 blocks
        self.grads.zero_()
        self.sq_chunk.zero_()
        self.flags.zero_()
        torch.cuda.synchronize(dev)
        self._h = [symm.rendezvous(t, group) for t in (self.grads, self.sq_chunk, self.flags)]
        if use_multicast is None:
            use_multicast = os.environ.get("KR_MULTICAST", "1" if self.world >= 4 else "0") != "0"
        mc_ok = bool(use_multicast) and all(bool(getattr(h, "has_multicast_support", False)) and int(h.multicast_ptr) != 0
                                            for h in self._h[:2])
        self.multicast = mc_ok
        self.GRID = self.GRID_MULTICAST if mc_ok else self.GRID_PEER
        self._mc = [int(self._h[0].multicast_ptr) if mc_ok else 0, int(self._h[1].multicast_ptr) if mc_ok else 0]
        arr = ctypes.c_void_p * self.world
        self._ptrs = [arr(*[int(p) for p in h.buffer_ptrs]) for h in self._h]
        store.grads = self.grads                  # every gradient view is taken from store.grads at call time
        store._views.clear()                      # (cached views of the old buffer would keep its 198 MB alive)
        opt.sq_chunk = self.sq_chunk[:opt.n_chunks]
        self.clip_global = torch.zeros(1, dtype=torch.float32, device=dev)
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.watchdog_seconds = float(os.environ.get("KR_COMM_WATCHDOG_S", "300"))
        dist.barrier(group=group, device_ids=[dev.index])

    def reduce(self, clip_local: Optional[torch.Tensor] = None, chunk_ranges: Optional[Tuple[int, int, int, int]] = None,
               grid: Optional[int] = None) -> Optional[torch.Tensor]:
        """All ranks' gradients -> their sum on every rank, plus the per-chunk squared sums of the reduced buffer.
        ``chunk_ranges`` = (begin0, end0, begin1, end1): only these two chunk-index ranges (the step reduces what is final
        half-way through the backward pass underneath the rest of it); None = everything.
        ``clip_local`` (device scalar): this rank's clip norm for the step; returns the device scalar holding the MINIMUM
        over all ranks (the clip every replica must apply — a per-rank clip would let the replicas diverge when one rank
        draws a long utterance and the stabiliser of trainer.py:2218-2255 tightens only its clip)."""
        import ctypes
        from ._lib import check, lib
        opt = self.opt
        rc = lib().kr_allreduce_sqnorm(ctypes.c_void_p(self._mc[0]), ctypes.c_void_p(self._mc[1]), self._ptrs[0],
                                       self._ptrs[1], self._ptrs[2], ctypes.c_int(self.rank), ctypes.c_int(self.world),
                                       ctypes.c_void_p(opt.chunk_start.data_ptr()), ctypes.c_void_p(opt.chunk_len.data_ptr()),
                                       ctypes.c_int(opt.n_chunks),
                                       (ctypes.c_int * 4)(*chunk_ranges) if chunk_ranges is not None else None,
                                       ctypes.c_int(grid or self.GRID),
                                       ctypes.c_void_p(clip_local.data_ptr() if clip_local is not None else None),
                                       ctypes.c_void_p(self.clip_global.data_ptr() if clip_local is not None else None),
                                       ctypes.c_double(self.watchdog_seconds), ctypes.c_void_p(self.error_flag.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        check(rc, "kr_allreduce_sqnorm")
        return self.clip_global if clip_local is not None else None


def broadcast_parameters(flat_params: torch.Tensor, src: int = 0, group=None) -> None:
    """Rank `src`'s weights to everyone (start of training / after loading a checkpoint on rank 0)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat_params, src=src, group=group)


def max_over_ranks(value: float, device=None, group=None) -> float:
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
