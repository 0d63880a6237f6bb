"""Micro-benchmark of the data-parallel collective (run under torchrun): the fused symmetric-memory all-reduce +
grad-norm kernel (multicast / peer path, several grid sizes) vs torch.distributed NCCL all_reduce of the same
197.7 MB flat gradient buffer.  Prints ms per call (max over ranks) and the algorithmic bus bandwidth."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokoro_ruslan_b200.optim import FusedAdamW  # noqa: E402
from kokoro_ruslan_b200.parallel import SymmetricGradReducer  # noqa: E402
from kokoro_ruslan_b200.params import ModelConfig, ParamStore  # noqa: E402


def timed(fn, dev, iters=20):
    for _ in range(3):
        fn()
    dist.barrier(device_ids=[dev.index])
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    store = ParamStore(ModelConfig(), dev, with_ema=False)
    opt = FusedAdamW(store)
    nbytes = store.total * 4
    plain = torch.randn(store.total, device=dev)
    t_nccl = timed(lambda: dist.all_reduce(plain), dev)
    rows = [("nccl all_reduce", t_nccl)]
    for mc in (True, False):
        red = SymmetricGradReducer(store, opt, dist.group.WORLD, use_multicast=mc)
        if mc and not red.multicast:
            rows.append(("fused multicast: not supported on this fabric", float("nan")))
            continue
        store.grads.normal_()
        valid = torch.zeros(store.total, dtype=torch.bool, device=dev)     # alignment padding is never reduced (zero in real use)
        for e in store.entries.values():
            valid[e.offset:e.offset + e.numel] = True
        store.grads.mul_(valid)
        mine = [torch.empty_like(store.grads) for _ in range(world)]
        dist.all_gather(mine, store.grads.clone())
        want = torch.stack(mine).double().sum(0)
        red.GRID = 592
        red.reduce()
        torch.cuda.synchronize(dev)
        err = float((store.grads.double() - want).abs().max() / want.abs().max())
        sq_want = torch.stack([store.grads[int(a):int(a) + int(n)].double().pow(2).sum() for a, n in
                               zip(opt.chunk_start[:64].tolist(), opt.chunk_len[:64].tolist())])
        sq_err = float((opt.sq_chunk[:64].double() - sq_want).abs().max() / sq_want.abs().max())
        rows.append((f"  sum error {err:.1e}, chunk-norm error {sq_err:.1e}", 0.0))
        store.grads.mul_(1e-30)      # the timing loop doubles the buffer every call
        for grid in ((64, 96, 148, 222, 296) if mc else (148, 296, 592)):
            red.GRID = grid
            if grid * world > red.flags.numel():
                continue
            rows.append((f"fused {'multicast' if mc else 'peer'} grid={grid}", timed(red.reduce, dev)))
        # correctness: every rank holds the same reduced buffer and chunk sums match a direct computation
        g = store.grads.clone()
        ref = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(ref, g)
        same = all(torch.equal(ref[0], r) for r in ref)
        rows.append((f"  replicas identical: {same}", 0.0))
    if rank == 0:
        for name, ms in rows:
            bw = nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9 if ms and ms == ms else 0.0
            print(f"N={world} {name:48s} {ms:8.3f} ms  busbw {bw:7.1f} GB/s")
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
