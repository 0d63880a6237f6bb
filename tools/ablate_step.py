"""In-step marginal cost of each kernel family: the bench-shape training step (CUDA graphs, dropout on) is timed with
one family's C-ABI entry points replaced by no-ops, one subprocess per family.  step(all) - step(without F) is what F
costs ON THE CRITICAL PATH of the multi-stream step — unlike the serialised ncu launch list or per-op event pairs, which
count work that overlaps other streams.  A profiling tool only: the ablated runs compute garbage.
usage: python tools/ablate_step.py            (all families)      python tools/ablate_step.py --one kr_attn_bwd"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = ["", "kr_attn_bwd", "kr_attn_fwd", "kr_attn_bwd_prep", "kr_qkv_prep_bwd", "kr_qkv_prep_fwd", "kr_layernorm_bwd",
            "kr_layernorm_fwd", "kr_glu_bwd", "kr_glu_fwd", "kr_rmsnorm_resid_bwd", "kr_rmsnorm_resid_ln_fwd", "kr_resid_drop_ln_fwd", "kr_colsum",
            "kr_gn_", "kr_adamw_step", "kr_gemm"]


def one(skip: str) -> None:
    sys.path.insert(0, ROOT)
    import torch
    import kokoro_ruslan_b200._lib as L
    from bench import B_PER_GPU, N_MELS, P_LEN, T_LEN, synthetic_batch
    from kokoro_ruslan_b200.engine import DropoutConfig
    from kokoro_ruslan_b200.params import ModelConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep
    real = L.lib()
    orig = L.lib

    class Proxy:
        def __getattr__(self, name):
            f = getattr(real, name)
            if skip and name.startswith(skip) and not (skip == "kr_attn_bwd" and name.startswith("kr_attn_bwd_prep")):
                return lambda *a, **k: 0
            return f
    proxy = Proxy()
    for m in list(sys.modules.values()):
        if getattr(m, "__name__", "").startswith("kokoro_ruslan_b200") and getattr(m, "lib", None) is orig:
            m.lib = lambda: proxy
    cfg = ModelConfig()
    ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=1000), device="cuda:0", use_graphs=True,
                   dropout=DropoutConfig.reference_training())
    ts.store.init_default(seed=0)
    host = {k: v.pin_memory() for k, v in synthetic_batch(B_PER_GPU, P_LEN, T_LEN, N_MELS, cfg.vocab_size, 1).items()}
    for _ in range(4):
        loss = ts.train_step(host)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        loss = ts.train_step(host)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"skip": skip, "ms_per_step": e0.elapsed_time(e1) / n, "loss0": float(loss.cpu()[0])}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        one(sys.argv[2] if sys.argv[2] != "none" else "")
    else:
        base = None
        for fam in FAMILIES:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", fam or "none"], capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if not line:
                print(fam, "FAILED", r.stderr[-400:])
                continue
            d = json.loads(line[-1])
            if not fam:
                base = d["ms_per_step"]
            print(f"{fam or 'full step':24s} {d['ms_per_step']:7.3f} ms   marginal {0.0 if base is None else base - d['ms_per_step']:+7.3f} ms   loss {d['loss0']}")
