"""Data-parallel parity check, run under torchrun with N ranks (one per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py

Every rank r trains on its own batch b_r for a few optimizer steps through TrainStep (the fused all-reduce +
grad-norm + global-clip kernel over symmetric memory).  Averaging per-rank batch-mean gradients is exactly the
reference's gradient-accumulation semantics (trainer.py:2284-2294), so rank 0 repeats the run single-process
as ONE accumulation window [b_0 .. b_{N-1}] per step and compares losses and weights.  Dropout off.

The LAST rank's batch carries one 200-frame token (> the 150-frame stabiliser threshold of trainer.py:2218-2255): that
rank alone computes a tightened clip (0.433) and loss scale (0.75).  The clip every replica applies must be the minimum
over the ranks — replicas stay bit-identical — and equals what the single-process window applies (its last
micro-batch's clip).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokoro_ruslan_b200.optim import OptimConfig  # noqa: E402
from kokoro_ruslan_b200.params import ModelConfig  # noqa: E402
from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep  # noqa: E402
from bench import synthetic_batch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = ModelConfig(vocab_size=59, mel_dim=80, hidden_dim=128, n_encoder_layers=2, n_heads=2, encoder_ff_dim=256,
                      n_decoder_layers=4, decoder_ff_dim=256, max_decoder_seq_len=1200, variance_filter_size=64)
    batches = [synthetic_batch(3, 24, 300, 80, 59, seed=100 + r) for r in range(world)]
    long_d = batches[world - 1]["phoneme_durations"]          # one long token on the last rank only, sum(d) unchanged
    take = 200 - int(long_d[0, 0])
    long_d[0, 0] += take
    k = 1
    while take > 0:
        t = min(take, int(long_d[0, k]) - 1)
        long_d[0, k] -= t
        take -= t
        k += 1
    assert int(long_d[0].sum()) == 300 and int(long_d.max()) == 200
    sched = ScheduleConfig(total_steps=1000, use_warmup=False, pct_start=0.5)
    ok = True
    for graphs in (False, True):
        ts = TrainStep(cfg, OptimConfig(learning_rate=1e-3), sched, device=dev, use_graphs=graphs,
                       process_group=dist.group.WORLD)
        ts.store.init_default(seed=0)          # identical weights on every rank
        if rank == 0:
            print(f"graphs={graphs}: multicast={ts.reducer.multicast}, overlapped early launches after decoder layers "
                  f"{ts.split_layer}")
        mine = [ts.train_step(batches[rank]).cpu() for _ in range(4)]
        torch.cuda.synchronize()
        w_dp = ts.store.params.clone()
        all_w = [torch.empty_like(w_dp) for _ in range(world)]
        dist.all_gather(all_w, w_dp)
        all_l = [torch.empty(4, 6, device=dev) for _ in range(world)]
        dist.all_gather(all_l, torch.stack(mine).to(dev))
        clip_used = ts.opt.read_ctrl()["clip_used"]
        if abs(clip_used - 0.5 / (200 / 150) ** 0.5) > 1e-6:
            ok = False
            print(f"rank {rank}: clip used {clip_used}, expected the last rank's stabiliser clip")
        if rank == 0:
            for r in range(1, world):      # replicas stay bit-identical
                same = torch.equal(all_w[0], all_w[r])
                ok &= same
                print(f"graphs={graphs}: rank {r} weights identical to rank 0: {same}")
            ref = TrainStep(cfg, OptimConfig(learning_rate=1e-3), sched, device=dev, use_graphs=False)
            ref.store.init_default(seed=0)
            for k in range(4):
                want = ref.train_window(batches)
                for r in range(world):
                    got = all_l[r][k].cpu()
                    close = torch.allclose(got, want[r].cpu(), rtol=2e-3, atol=1e-5)
                    ok &= close
                    if not close:
                        print("loss mismatch", k, r, got.tolist(), want[r].cpu().tolist())
            base = TrainStep(cfg, device=dev, use_graphs=False)
            base.store.init_default(seed=0)
            p0 = base.store.params
            err = float(((w_dp - p0) - (ref.store.params - p0)).norm() / (ref.store.params - p0).norm())
            print(f"graphs={graphs}: relative L2 error of the accumulated update vs the single-process window: {err:.3e}")
            ok &= err < 2e-2
    if rank == 0:
        print("DP_CHECK", "OK" if ok else "FAILED")
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
