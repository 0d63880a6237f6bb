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


def _shard_worker(rank, world, port, out_dir, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from kokoro_ruslan_b200 import checkpoint as ck
    from kokoro_ruslan_b200.parallel import broadcast_parameters, init_distributed
    init_distributed("gloo")
    flat = torch.arange(4096, dtype=torch.float32) * (rank + 1)      # different before the broadcast, identical after
    broadcast_parameters(flat, src=0)
    payload = {"epoch": 1, "model_state_dict": {f"p{i}": flat[i * 512:(i + 1) * 512].clone() for i in range(8)},
               "optimizer_state_dict": {"state": {0: {"exp_avg": flat[:100] * 0.5}}, "param_groups": [{"lr": 1e-4}]}}
    writer = ck.AsyncCheckpointWriter()                              # the async writer path, one writer per rank
    path = os.path.join(out_dir, "checkpoint_epoch_2.pth")
    files = ck.save_sharded(path, payload, rank, world, writer)
    writer.wait()
    dist.barrier()                                                    # every shard is on disk
    merged = ck.load_sharded(path)
    ok = all(torch.equal(merged["model_state_dict"][f"p{i}"], torch.arange(4096, dtype=torch.float32)[i * 512:(i + 1) * 512])
             for i in range(8)) and torch.equal(merged["optimizer_state_dict"]["state"][0]["exp_avg"],
                                                torch.arange(100, dtype=torch.float32) * 0.5)
    q.put((rank, [os.path.basename(f) for f in files], ok, merged["epoch"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_checkpoint(tmp_path):
    """Sharded checkpoints with two gloo ranks: each rank writes its share of the replicated state
    through its own asynchronous writer, rank 0 adds the head file, and both ranks read the same merged payload back."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == ["checkpoint_epoch_2.pth.shard0of2", "checkpoint_epoch_2.pth"]
    assert res[1][1] == ["checkpoint_epoch_2.pth.shard1of2"]
    assert all(r[2] and r[3] == 1 for r in res)
    sizes = [os.path.getsize(os.path.join(str(tmp_path), f"checkpoint_epoch_2.pth.shard{r}of2")) for r in range(2)]
    assert abs(sizes[0] - sizes[1]) < 4096                             # balanced shares
