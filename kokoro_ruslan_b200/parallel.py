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
