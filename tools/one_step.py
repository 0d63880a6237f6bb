"""Runs N eager (non-graph) training steps at the bench workload — the command profiled under ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import B_PER_GPU, N_MELS, P_LEN, T_LEN, synthetic_batch  # noqa: E402
from kokoro_ruslan_b200.engine import DropoutConfig  # noqa: E402
from kokoro_ruslan_b200.params import ModelConfig  # noqa: E402
from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = ModelConfig()
dropout = None if os.environ.get("KR_DROPOUT", "reference") == "off" else DropoutConfig.reference_training()
ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=1000), device="cuda:0", use_graphs=False, dropout=dropout)
ts.store.init_default(seed=0)
host = {k: v.pin_memory() for k, v in synthetic_batch(B_PER_GPU, P_LEN, T_LEN, N_MELS, cfg.vocab_size, 1).items()}
for i in range(n):
    torch.cuda.synchronize()
    print("step", i, ts.train_step(host).cpu().tolist(), "launches", ts.launches_last_step)
