"""world_size-2 gloo test of the data-parallel step protocol on CPU: identical epoch partition on both
ranks, one SUM all-reduce of the flat gradient buffer with 1/world pre-scaling == gradient of the mean
of the per-rank losses, and replicated optimizer decisions."""
import os
import random
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class DummyDataset:
    def __init__(self, lengths):
        self.samples = [{"audio_length": int(v)} for v in lengths]

    def __len__(self):
        return len(self.samples)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from kokoro_ruslan_b200.data import DistributedBatchSampler, DynamicFrameBatchSampler
    from kokoro_ruslan_b200.parallel import all_reduce_gradients, broadcast_parameters, init_distributed, max_over_ranks
    r, _, w = init_distributed("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(50, 900, (101,), generator=g).tolist()
    random.seed(rank * 17 + 1)                               # different RNG state per rank on purpose
    sampler = DistributedBatchSampler(DynamicFrameBatchSampler(DummyDataset(lens), max_frames=4000, min_batch_size=1),
                                      rank, world, seed=7)
    sampler.set_epoch(2)
    mine = list(iter(sampler))
    # toy "model": flat parameter vector, per-rank loss = mean over its batches of a quadratic
    params = torch.full((64,), float(rank))                  # deliberately different before the broadcast
    broadcast_parameters(params, src=0)
    target = torch.tensor([float(sum(b) % 13) for b in mine])
    local_loss_grad = (params.unsqueeze(0) - target.unsqueeze(1)).mean(dim=0)          # d/dp of 0.5*mean((p-t)^2)
    flat = local_loss_grad / world                           # 1/world folded into the loss scale
    all_reduce_gradients(flat)
    t_max = max_over_ranks(float(rank + 1))
    q.put((rank, mine, flat.tolist(), params.tolist(), target.tolist(), t_max))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_protocol():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, g0, p0, t0, m0), (r1, b1, g1, p1, t1, m1) = res
    assert len(b0) == len(b1) and not set(map(tuple, b0)) & set(map(tuple, b1))
    assert p0 == p1 == [0.0] * 64                             # broadcast from rank 0
    assert g0 == g1                                           # identical reduced gradients on both ranks
    want = 0.5 * ((torch.tensor(p0).unsqueeze(0) - torch.tensor(t0).unsqueeze(1)).mean(0)
                  + (torch.tensor(p1).unsqueeze(0) - torch.tensor(t1).unsqueeze(1)).mean(0))
    assert torch.allclose(torch.tensor(g0), want, atol=1e-6)
    assert m0 == m1 == 2.0
