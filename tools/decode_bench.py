"""Autoregressive decode throughput: frames / s of InferenceEngine.generate at the full model width
(random-init weights, B utterances x P tokens), CUDA-graph step, and the per-step floor it should be read against:
the decoder's bf16 weights (27 M parameters = 54 MB) are re-read from L2 every frame, plus the KV caches
(6 layers x 2 x t x 512 x 2 B per utterance) and the cross-attention memory K/V.

    python tools/decode_bench.py [B] [P] [frames]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_ruslan_b200.model import KokoroModel  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    frames = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    m = KokoroModel(vocab_size=59, encoder_ff_dim=1536, decoder_ff_dim=1536, qk_norm=True)
    m.eval()
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(1, 59, (B, P), generator=g).cuda()
    # random weights neither predict sensible durations nor raise the stop probability in a controlled way: give every
    # token 6 frames (a 6 * P frame cross-attention memory, as at the bench shape) and pin the length with the loop bounds
    from kokoro_ruslan_b200.inference import InferenceEngine
    inf = InferenceEngine(m.engine)
    dur = torch.full((B, P), 6, dtype=torch.int64, device="cuda")
    kw = dict(durations=dur, min_len_floor=frames, max_len_cap=frames + 1, max_len_ratio=1000.0, min_len_ratio=0.0)
    inf.generate(idx, None, **kw)                        # warm-up (builds kernels' lazy state)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mel = inf.generate(idx, None, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = mel.shape[1]
    dec_params = sum(p.numel() for name, p in m.named_parameters() if name.startswith("decoder."))
    print(json.dumps({"workload": f"AR decode B={B} P={P} memory={6 * P} frames", "frames": n,
                      "projections": "kr_dec_gemv" if inf.be.use_gemv and B <= 8 else "tcgen05 GEMM, 128 padded rows",
                      "s": round(dt, 4),
                      "frames_per_s": round(B * n / dt), "us_per_step": round(dt / n * 1e6, 1),
                      "x_realtime": round(n * 256 / 22050 / dt, 1),
                      "weights_mb_per_step": round(dec_params * 2 / 1e6, 1)}))


if __name__ == "__main__":
    main()
