"""Race / reproducibility stress: two independent TrainStep instances, same weights and batches, 12 steps each
(eager step, capture, replays) — per-step losses must agree to 1e-3 relative although fp32 atomics reorder."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import B_PER_GPU, N_MELS, P_LEN, T_LEN, synthetic_batch
from kokoro_ruslan_b200.params import ModelConfig
from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep

def run(graphs):
    cfg = ModelConfig()
    ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=1000), device="cuda:0", use_graphs=graphs)
    ts.store.init_default(seed=0)
    out = []
    for i in range(12):
        host = {k: v.pin_memory() for k, v in synthetic_batch(B_PER_GPU, P_LEN, T_LEN, N_MELS, cfg.vocab_size, 1 + i % 3).items()}
        out.append(ts.train_step(host).cpu().clone())
    del ts
    torch.cuda.empty_cache()
    return torch.stack(out)

a, b, c = run(True), run(True), run(False)
for name, x in (("graph vs graph", b), ("graph vs eager", c)):
    rel = ((a - x).abs() / (a.abs() + 1e-6)).max().item()
    print(name, "max rel diff of the 12x6 losses:", rel)
    assert rel < 2e-3, rel
print("total loss trajectory:", [round(float(v), 4) for v in a[:, 0]])
print("stress ok")
