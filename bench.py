#!/usr/bin/env python
"""bench.py — acoustic-model training step throughput (mel-frames/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json north_star / SURVEY.md §8(d)): synthetic batch B=8 utterances per GPU,
phoneme_len 128, mel_frames 800, n_mels 80, default 49.4 M-parameter model, bf16 tensor-core GEMMs
with fp32 residual stream / accumulation / optimizer.  One "step" = H2D-resident batch ->
forward -> fused losses -> backward -> (all-reduce) -> clip + AdamW + EMA (one optimizer step).

Prints ONE JSON line (rank 0).  `value` = device-resident whole-job throughput; `e2e` = the same
through TrainStep.train_step() with pinned HOST batches (H2D inside the timed region) and a D2H
read of the losses every step.  `roofline` = the tcgen05 GEMM family (dominant kernel) timed per
launch with CUDA events right after the timed region — and, inside the same object so that the driver keeps
them, `top_ops` (per-op table of the step), `hifigan` (second half of BASELINE.json's metric: HiFi-GAN samples/s
with its own e2e and roofline fractions), `features` and `decode` (micro-benchmarks of the feature-extraction and
autoregressive-decode kernels, isolated subprocesses with hard timeouts after every headline measurement).
`cpu_baseline` = the reference's own CPU path (below) on a bounded sample, plus `gpu_eager_ms`: the unmodified
reference model trained with plain PyTorch (bf16 autocast, fused AdamW) on the SAME GPU.

`--impl reference` runs the UNMODIFIED reference from baseline/_ref on the host cores through its own code: the
reference `KokoroTrainer.train_epoch` (forward, reference losses, backward, pre-clip, explosion detector, clip, AdamW,
EMA, projection) around the reference `KokoroModel`, fp32, reference dropout defaults, at the SAME fixed batch
(B = 8, P = 128, T = 800).  A CPU step takes many seconds, so when K + W steps do not fit the arm's time budget the
number of timed steps is cut (never the batch) and the line says so (`steps` = timed steps, `requested_steps` = K).
Falls back to the oracle port (`kind: "port"`) only when baseline/_ref is absent.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU, P_LEN, T_LEN, N_MELS = 8, 128, 800, 80
METRIC = "mel_frames_per_sec_train_step"
UNIT = "mel-frames/s"
WORKLOAD = "acoustic train step: B=8/GPU, phoneme_len=128, mel_frames=800, n_mels=80 (configs[1]/[2] synthetic point)"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return {"hbm_gbs": float(pk["hbm_gbs"]), "tf_burst": float(pk["bf16_tflops"]),
                "tf_sustained": float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def synthetic_batch(B, P, T, n_mels, vocab, seed):
    """Seeded synthetic batch of SURVEY.md §8(d) (same recipe as oracle.acoustic.synthetic_batch,
    restated here so the product arm never imports oracle/)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ph = torch.randint(1, vocab, (B, P), generator=g)
    stress = torch.randint(0, 3, (B, P), generator=g)
    base, extra = T // P, T - (T // P) * P
    dur = torch.full((B, P), base, dtype=torch.long)
    dur[:, :extra] += 1
    mel = torch.randn(B, T, n_mels, generator=g) * 2.0 - 5.0
    pitch = torch.rand(B, T, generator=g)
    energy = torch.rand(B, T, generator=g)
    stop = torch.zeros(B, T)
    for k in range(7):
        stop[:, T - 1 - k] = 0.5 ** k
    return {"phoneme_indices": ph, "stress_indices": stress, "phoneme_durations": dur, "mel_specs": mel,
            "pitches": pitch, "energies": energy, "stop_token_targets": stop,
            "mel_lengths": torch.full((B,), T, dtype=torch.long), "phoneme_lengths": torch.full((B,), P, dtype=torch.long)}


# ----------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline and --impl reference)
# ----------------------------------------------------------------------------------------------
DROPOUT_DESC = {"reference": "reference training defaults: encoder 0.15, decoder 0.20, decoder input 0.15, "
                             "variance 0.1, stochastic depth 0.1 (training/config.py:108-121,195)",
                "off": "0.0 (deterministic parity configuration)"}


def _host_threads():
    """All the host cores the box has: torchrun exports OMP_NUM_THREADS=1, which would make the CPU arm
    single-threaded under N > 1 launches."""
    import torch
    try:
        import psutil
        want = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    except Exception:
        want = os.cpu_count() or 1
    try:
        want = min(want, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    if torch.get_num_threads() < want:
        torch.set_num_threads(want)
    return torch.get_num_threads()


REF_TRAINER_OVERRIDES = dict(  # the bench point: default model (training/config.py), one optimizer step per batch
    gradient_accumulation_steps=1, num_epochs=1, use_fused_adamw=False, enable_profiling=False, profile_epoch_start=999,
    use_torch_compile=False, use_spec_augment=False)


def cpu_reference_run(steps: int, warmup: int, budget_s: float, dropout: str = "reference"):
    """The reference's own CPU implementation of the step: the unmodified KokoroTrainer.train_epoch from baseline/_ref
    around the reference KokoroModel (oracle/ref_trainer.py builds the trainer object without a dataset on disk), fp32,
    fixed B = 8 batch.  Steps are cut to the time budget, never the batch.
    Returns (frames/s, ms/step, cores, sample text, timed steps, kind)."""
    import torch
    from oracle import ref_trainer as rt
    if not rt.reference_available():
        v, ms, cores, sample, n = cpu_oracle_run(steps, warmup, budget_s, dropout)
        return v, ms, cores, sample, n, "port"
    cores = _host_threads()
    over = dict(REF_TRAINER_OVERRIDES)
    if dropout == "off":
        over.update(encoder_dropout=0.0, decoder_dropout=0.0, decoder_input_dropout=0.0, variance_dropout=0.0,
                    use_stochastic_depth=False)
    torch.manual_seed(0)
    tr = rt.build_trainer(None, torch.device("cpu"), 59, over)
    batch = synthetic_batch(B_PER_GPU, P_LEN, T_LEN, N_MELS, 59, seed=1)
    t0 = time.perf_counter()
    rt.run_epoch(tr, [batch])                                   # first step: allocator / thread-pool warm-up
    t_first = time.perf_counter() - t0
    n_warm = 1
    left = budget_s - t_first
    n_timed = max(1, min(steps, int(left / t_first) - max(0, min(warmup, 2) - 1)))
    extra_warm = max(0, min(warmup - 1, int(left / t_first) - n_timed))
    if extra_warm:
        rt.run_epoch(tr, [batch] * extra_warm)
        n_warm += extra_warm
    done0 = tr.optimizer_steps_completed
    t0 = time.perf_counter()
    rt.run_epoch(tr, [batch] * n_timed)
    dt = time.perf_counter() - t0
    assert tr.optimizer_steps_completed - done0 == n_timed, "the reference trainer skipped a batch"
    sample = (f"{n_timed} optimizer steps of the UNMODIFIED reference KokoroTrainer.train_epoch (baseline/_ref: forward, "
              f"losses, backward, pre-clip, explosion detector, clip, AdamW, EMA, projection; fp32, gradient checkpointing "
              f"as the reference configures it, dropout {'0' if dropout == 'off' else 'at the reference training defaults'}) on "
              f"B={B_PER_GPU} utterances x P={P_LEN} x T={T_LEN} after {n_warm} warm-up step(s), torch {torch.__version__} CPU, "
              f"{cores} threads")
    return n_timed * B_PER_GPU * T_LEN / dt, dt / n_timed * 1e3, cores, sample, n_timed, "reference"


def cpu_oracle_run(steps: int, warmup: int, budget_s: float, dropout: str = "reference"):
    """Fallback when baseline/_ref is absent: the oracle port of the training step (oracle/train_step.py, pinned to the live
    trainer by tests/golden/trainer_step.npz), fixed B = 8, steps cut to the budget."""
    import torch
    from oracle import acoustic as oa
    from oracle.train_step import CpuTrainStep
    cores = _host_threads()
    cfg = oa.AcousticConfig()
    drop = None
    if dropout != "off":      # training/config.py:108-121,195
        drop = oa.TorchDropout(p_enc=0.15, p_dec=0.20, p_in=0.15, p_var=0.1, sd_rate=0.1,
                               n_enc=cfg.n_encoder_layers, n_dec=cfg.n_decoder_layers)
    step = CpuTrainStep(cfg, oa.seeded_state_dict(cfg, seed=0), drop=drop)
    batch = oa.synthetic_batch(B=B_PER_GPU, P=P_LEN, T=T_LEN, seed=1)
    t0 = time.perf_counter()
    step.train_step(batch)
    t_first = time.perf_counter() - t0
    n_timed = max(1, min(steps, int((budget_s - t_first) / t_first)))
    t0 = time.perf_counter()
    for _ in range(n_timed):
        step.train_step(batch)
    dt = time.perf_counter() - t0
    sample = (f"{n_timed} full optimizer steps of the oracle PORT (fwd+losses+bwd+pre-clip+clip+AdamW+EMA, fp32, dropout "
              f"{'0' if drop is None else 'at the reference training defaults'}) of "
              f"B={B_PER_GPU} utterances x P={P_LEN} x T={T_LEN}, torch {torch.__version__} CPU, {cores} threads")
    return n_timed * B_PER_GPU * T_LEN / dt, dt / n_timed * 1e3, cores, sample, n_timed


def bench_config(world: int, dropout: str, comm=None):
    """The `config` object shared by both arms (same workload, same keys)."""
    return {"workload": WORKLOAD, "global_batch": B_PER_GPU * world, "params": 49432276,
            "precision": "bf16 tcgen05 GEMM/attention operands, fp32 accumulate/residual/optimizer",
            "dropout": DROPOUT_DESC[dropout], "grad_accum": 1, "parallelism": f"dp{world}", "comm": comm,
            "l2": "no flush: a step streams >1 GB of weights/optimizer state/activations (> 126 MB L2)",
            "cuda_graphs": True}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, cores, sample, n_timed, kind = cpu_reference_run(args.steps, args.warmup, budget_s=240.0, dropout=args.dropout)
    cfg = bench_config(1, args.dropout)
    cfg.update(precision="fp32 (the reference's CPU path)", cuda_graphs=False, parallelism="cpu",
               l2="n/a (host cores)")
    if n_timed != args.steps:
        cfg["reference_arm_note"] = (f"same fixed batch (B=8, P=128, T=800); {n_timed} of the requested {args.steps} steps "
                                     f"timed to fit the arm's 240 s budget (a CPU step takes ~{ms / 1e3:.0f} s)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": n_timed, "requested_steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# HiFi-GAN vocoder leg (second half of BASELINE.json's metric): 16 mels x 800 frames -> 22.05 kHz audio
# ----------------------------------------------------------------------------------------------
def hifigan_leg(peaks, steps: int = 10, B: int = 16, T: int = 800):
    import torch
    from kokoro_ruslan_b200 import _lib
    from kokoro_ruslan_b200.hifigan import HiFiGANConfig, HiFiGANGenerator
    gen = HiFiGANGenerator(HiFiGANConfig.get_default_config(), device="cuda")
    g = torch.Generator().manual_seed(0)
    mel_host = (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).pin_memory()
    mel_dev = mel_host.cuda()
    out_host = torch.empty(B, 1, T * 256).pin_memory()
    for _ in range(3):
        gen(mel_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        gen(mel_dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    e0.record()
    for _ in range(steps):
        out_host.copy_(gen(mel_host), non_blocking=True)      # H2D of the mels + D2H of the audio, every call
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / steps
    samples = B * T * 256
    flops = 614.1e6 * B * T                                   # 614.1 MFLOP per mel frame
    alg_bytes = 2.03e6 * B * T                                # ... and 2.03 MB (bf16) of activation traffic
    traffic = None
    try:
        import glob
        with open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_hifigan_traffic*.json")))[-1]) as f:
            tr = json.load(f)
        convs = [tr[k] for k in ("kr_gemm_kernel", "hifi_resblock_kernel") if k in tr]
        traffic = sum(c["dram_read_bytes"] + c["dram_write_bytes"] for c in convs) / float(tr.get("_forwards_in_capture", 2))
    except Exception:
        pass
    return {"metric": "hifigan_audio_samples_per_sec", "value": samples / (ms * 1e-3), "unit": "samples/s",
            "ms_per_batch": ms, "config": {"workload": "HiFi-GAN v1 generator, batch=16 mels x 800 frames -> 204800 samples each",
                                           "dtype": "bf16 tcgen05 implicit-GEMM convs, fp32 residual stream"},
            "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_batch": ms_e2e,
                    "h2d_bytes_per_step": mel_host.numel() * 4, "d2h_bytes_per_step": samples * 4},
            "roofline": {"bound": "hbm", "kernel": "kr_gemm_kernel conv mode + hifi_resblock_kernel (%d launches per forward)" % gen.launches_last_forward,
                         "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": alg_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": traffic,
                         "tensor_frac": flops / (ms * 1e-3) / 1e12 / peaks["tf_sustained"],
                         "note": "algorithmic bytes = 2.03 MB/frame (bf16, each conv reads its input and writes its output "
                                 "once; the fp32 residual stream this implementation keeps for accuracy is NOT in it); traffic = ncu dram bytes "
                                 "of the conv kernels per forward (profiles/r02_hifigan_traffic.json)"},
            "gpu_launches": gen.launches_last_forward * steps * 2}


# ----------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------
def _json_subprocess(cmd, env=None, timeout=120.0):
    """Runs `python cmd...` in its own process with a hard timeout; returns the JSON lines it printed (or an error dict):
    an auxiliary measurement can never take the bench line down."""
    import subprocess
    root = os.path.dirname(os.path.abspath(__file__))
    try:
        r = subprocess.run([sys.executable] + cmd, cwd=root, env={**os.environ, **(env or {})}, capture_output=True,
                           text=True, timeout=timeout)
        rows = []
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                try:
                    rows.append(json.loads(ln))
                except ValueError:
                    pass
        return rows if (r.returncode == 0 and rows) else {"rc": r.returncode, "error": (r.stderr or r.stdout)[-400:]}
    except subprocess.TimeoutExpired:
        return {"error": "timeout"}
    except Exception as exc:
        return {"error": repr(exc)}


def kernel_legs():
    """Micro-benchmarks of the kernels beside the training step: feature extraction at the bench
    shape with the HBM fraction of each kernel, autoregressive decode frames/s — isolated subprocesses after every
    headline measurement.  They land inside `roofline` so that the driver's record keeps them."""
    return {"features": _json_subprocess(["tools/features_bench.py"]),
            "decode": _json_subprocess(["tools/decode_bench.py", "1", "64", "400"], timeout=200.0)}


def reference_gpu_eager_ms():
    """The unmodified reference model (baseline/_ref) trained with plain PyTorch — bf16 autocast, fused AdamW, its own loss
    function — on the SAME GPU at the bench shape: the like-for-like bar of SURVEY.md 8(d).  ms per step, or an error."""
    rows = _json_subprocess(["tools/ref_gpu_step.py"], timeout=240.0)
    if isinstance(rows, list):
        for r in rows:
            if r.get("autocast") == "bf16":
                return r
    return rows


def run_ours(args):
    import torch
    import torch.distributed as dist
    from kokoro_ruslan_b200.build import build
    build()
    from kokoro_ruslan_b200 import _lib
    from kokoro_ruslan_b200.engine import DropoutConfig
    from kokoro_ruslan_b200.kprof import OpTimer
    from kokoro_ruslan_b200.params import ModelConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)

    cfg = ModelConfig()
    dropout = DropoutConfig.reference_training() if args.dropout != "off" else None
    if dropout is not None:
        dropout.seed = 1234 + rank      # data-parallel ranks draw independent masks
    ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=100000), device=dev, use_graphs=True, process_group=pg,
                   dropout=dropout)
    ts.store.init_default(seed=0)          # same weights on every rank
    host = synthetic_batch(B_PER_GPU, P_LEN, T_LEN, N_MELS, cfg.vocab_size, seed=1 + rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    frames_per_step = B_PER_GPU * T_LEN * world

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up through the public API (eager pass, graph capture, replays) --------------------
    W = max(3, args.warmup)
    for _ in range(W):
        losses = ts.train_step(host)
    torch.cuda.synchronize(dev)
    first_losses = losses.cpu().tolist()
    launches_per_step = ts.launches_last_step
    # device-resident copy of the batch for the `value` leg (same static buffers: no copies at all)
    st, key = ts.stage(host)
    resident = st.dev

    # ---- e2e leg: host batches in, losses out, every step ---------------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        l_host = ts.train_step_host(host)     # host batch in, the step's six losses out as Python floats (D2H inside the step)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps

    # ---- value leg: inputs resident in HBM --------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ts.train_step(resident)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop()
    final_losses = ts._staged[key].losses.cpu().tolist() if ts._staged[key].losses is not None else None

    if rank != 0:
        if world > 1:
            dist.barrier(device_ids=[local])
            dist.destroy_process_group()
        return

    # ---- roofline: instrumented eager step (per-launch CUDA events) ---------------------------------
    peaks = _peaks()
    roof, top_ops = None, []
    try:
        ts_use_graphs = ts.use_graphs
        ts.use_graphs = False
        world_saved = ts.world
        ts.world = 1                              # rank-0-only pass: no collective
        with OpTimer() as tm:
            ts.stage(resident)
            ts.opt.set_lrs(ts.sched.lrs())
            ts._run_fwd_bwd(st, key)
            with tm.region("optimizer(sqnorm+control+adamw_ema+wn_project)", byts=9 * 4.0 * ts.store.total):
                ts._run_optimizer()
        ts.use_graphs, ts.world = ts_use_graphs, world_saved
        rows = tm.summary()
        replay_ms, replay_n = tm.replay_gemms()
        step_ms_eager = sum(r["ms"] for r in rows)
        gemm = [r for r in rows if r["op"].startswith("gemm")]
        g_ms, g_fl = sum(r["ms"] for r in gemm), sum(r["flops"] for r in gemm)
        g_n = sum(r["launches"] for r in gemm)
        all_fl = sum(r["flops"] for r in rows)
        eager_achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
        achieved = g_fl / (replay_ms * 1e-3) / 1e12 if replay_ms > 0 else eager_achieved
        roof = {"bound": "tensor", "kernel": "kr_gemm_kernel (tcgen05 bf16 GEMM, all %d launches of a step)" % g_n,
                "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tf_sustained"], "traffic": None,
                "avg_launch_us": replay_ms * 1e3 / max(1, replay_n),
                "frac_eager_event_pairs": eager_achieved / peaks["tf_sustained"],
                "peak_source": peaks["src"] + " (bf16_tflops_sustained: kernel timed inside a long step)",
                "share_of_step": g_ms / step_ms_eager if step_ms_eager else None,
                "algorithmic_flops_per_step": all_fl,
                "step_tensor_frac": all_fl / (ms * 1e-3) / 1e12 / peaks["tf_sustained"],
                "how": "the step's %d GEMM launches (same operands, same order) replayed back to back in a CUDA graph, "
                       "CUDA events around 5 replays, after the timed region; frac_eager_event_pairs = the same launches "
                       "timed one by one with an event pair each in an eager pass (adds ~3 us to every ~15 us launch); "
                       "share_of_step from the eager pass" % replay_n}
        top_ops = [{"op": r["op"], "n": r["launches"], "ms": round(r["ms"], 4), "share": round(r["share"], 4),
                    "tflops": round(r["tflops"], 1), "gbs": round(r["gbs"], 1)} for r in rows[:48]]
    except Exception as exc:  # the bench line must still be printed
        roof = {"bound": "tensor", "error": repr(exc)}

    # ---- measured DRAM traffic of the dominant kernel (ncu capture of this command, profiles/) -------------
    try:
        import glob
        tfile = sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_step_traffic*.json")))[-1]
        with open(tfile) as f:
            tr = json.load(f)["kr_gemm_kernel"]
        if roof is not None and "error" not in roof:
            roof["traffic"] = tr["traffic_per_launch"]
            roof["traffic_note"] = ("mean dram__bytes_read+write per kr_gemm_kernel launch, ncu --clock-control none on "
                                    "tools/one_step.py (profiles/" + os.path.basename(tfile) + "); algorithmic bytes per launch = "
                                    "%.1f MB" % (sum(r["bytes"] for r in rows if r["op"].startswith("gemm")) / max(1, g_n) / 1e6))
    except Exception:
        pass

    # ---- second headline metric: HiFi-GAN vocoder inference (BASELINE config 5) ----------------------------
    hifi = None
    if world == 1 and not args.no_hifigan:
        try:
            hifi = hifigan_leg(peaks, steps=max(5, args.steps // 2))
        except Exception as exc:
            hifi = {"error": repr(exc)}

    # ---- CPU baseline: the reference's own trainer step on the host cores (bounded sample) + the reference on this GPU
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, ms_cpu, cores, sample, _, kind = cpu_reference_run(steps=2, warmup=1, budget_s=45.0, dropout=args.dropout)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "ms_per_step": ms_cpu}
            eager = reference_gpu_eager_ms()
            if isinstance(eager, dict) and "ms_per_step" in eager:
                cpu["gpu_eager_ms"] = eager["ms_per_step"]
                cpu["gpu_eager_note"] = eager.get("what")
            else:
                cpu["gpu_eager_ms"] = None
                cpu["gpu_eager_error"] = eager
        except Exception as exc:
            cpu = {"error": repr(exc)}

    # ---- the kernels beside the step (features, decode): isolated subprocesses, after everything above ------
    if roof is not None:
        roof["top_ops"] = top_ops
        roof["top_ops_note"] = ("event pairs around each C-ABI call in an EAGER multi-stream pass: a row's ms is the op's duration in "
                                "situ, i.e. it includes the slow-down from kernels running concurrently on the side streams "
                                "(weight gradients, encoder, predictors) and ~3 us of event overhead per launch; the rows therefore "
                                "sum to far more than the graph-replayed step.  An op family's marginal cost ON the step's critical "
                                "path is in marginal_in_step_ms")
        try:   # tools/ablate_step.py: step time with one family's entry points turned into no-ops (committed capture)
            marg = {}
            with open(os.path.join(ROOT, "profiles", "r02_step_ablation.txt")) as f:
                for line in f:
                    parts = line.split()
                    if len(parts) >= 5 and parts[0].startswith("kr_") and parts[2] == "ms":
                        marg[parts[0]] = float(parts[4])
            roof["marginal_in_step_ms"] = {"source": "profiles/r02_step_ablation.txt (tools/ablate_step.py on a B200, bench shape, "
                                                     "graph replay; step(all) - step(without the family))", **marg}
        except Exception:
            pass
        roof["hifigan"] = hifi
        if world == 1 and not args.no_extras:
            try:
                roof.update(kernel_legs())
            except Exception as exc:
                roof["kernel_legs_error"] = repr(exc)

    comm = None
    if world > 1:
        comm = "fused all-reduce + grad-norm + global-clip kernel over symmetric memory (%s; %s)" % (
            "NVSwitch multicast" if ts.reducer.multicast else "peer loads/stores",
            ("early gradient ranges reduced underneath the backward after decoder layers %s" % ts.split_layer)
            if ts.split_layer is not None else "one launch after the backward")
    line = {"metric": METRIC, "value": frames_per_step / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": bench_config(world, args.dropout, comm),
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": frames_per_step / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d_bytes + 12 + 40, "d2h_bytes_per_step": 28,
                    "api": "TrainStep.train_step_host(host batch) -> six losses as Python floats, every step: the batch goes "
                           "host -> device in two copies before the step's graph, the losses (+ a sequence id) come back by a "
                           "D2H copy inside the step right after the loss kernel, so the host stages the next batch while the "
                           "backward pass and optimizer run; the next step starts after this one has finished (stream order)"},
            "gpu_launches": launches_per_step * args.steps * 2, "launches_per_step": launches_per_step,
            "clocks": clocks, "losses_first": first_losses, "losses_last": final_losses, "lib": str(_lib.LIB_PATH)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dropout", default="reference", choices=["reference", "off"],
                    help="reference = the trainer's dropout / stochastic-depth defaults (the real training step); "
                         "off = the deterministic parity configuration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hifigan", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the isolated micro-benchmarks of the feature / decode kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
