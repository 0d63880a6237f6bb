"""Regenerates include/kokoro_b200.h from the extern "C" definitions in kokoro_ruslan_b200/csrc.

The prototypes are copied verbatim from the sources (single source of truth); the per-function
comment (what it replaces in the reference, file:line) comes from DOCS below.  A function without a
DOCS entry fails the generation so the header can never silently lose its citations.
"""
import glob
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DOCS = {
    "kr_last_error": "Thread-local message of the last failing call (every entry point returns 0 or a negative KR_ERR_* code; the Python side maps non-zero to RuntimeError, which the reference trainer's per-batch handler expects: training/trainer.py:2679-2686).",
    "kr_abi_version": "ABI version of this header (3: kr_allreduce_sqnorm chunk ranges / global clip / watchdog).",
    "kr_launch_count": "Number of CUDA kernels this library has launched since it was loaded (bench.py's gpu_launches evidence).",
    "kr_memset_zero": "Asynchronous zero fill (memset node under graph capture) — replaces the torch fill kernels for gradient / scratch buffers.",
    "kr_device_cc": "Compute capability (major*10+minor) of the current device; 100 on B200.",
    "kr_gemm_bf16": "tcgen05 GEMM C[M,N] (+)= alpha*A[M,K]*B[N,K]^T (+bias[n]) (+resid[m % resid_mod, n]); bf16 operands staged by TMA, fp32 accumulation in TMEM. a_mn_major / b_mn_major select the [K,rows] storage that weight- and data-gradient GEMMs read in place. epi_mode 0 = bf16 store, 1 = fp32 store, 2 = fp32 atomic accumulate (split-K). Replaces every nn.Linear on the path: model/transformers.py:228,258-259,434 (Q/K/V/out projections), transformers.py:105-111 (GLU FFN), model/model.py:519-531,561 (mel in/out projections), and — through an overlapping-row view — the k=3 nn.Conv1d of model/variance_predictor.py:46.",
    "kr_gemm_ex": "Persistent tcgen05 GEMM / implicit-GEMM conv1d with the full fused epilogue (see kr_gemm_args). kr_gemm_bf16 is the plain-argument subset. The conv mode replaces nn.Conv1d / nn.ConvTranspose1d (polyphase) of inference/hifigan_vocoder.py:31-133.",
    "kr_hifi_pack_mel": "Mel (B,80,T) [time_major=0] or (B,T,80) [1] fp32 -> channels-last bf16 [B, T+2*halo, c_phys] (interior rows; halos and padded channels stay zero): the input-layout handling of inference/hifigan_vocoder.py:112-117 fused with the bf16 cast.",
    "kr_hifi_resblock": "One HiFi-GAN ResBlock step in one kernel (inference/hifigan_vocoder.py:31-83: xt = c1(lrelu(x)); xt = c2(lrelu(xt)); x = xt + x): t = lrelu(conv1(x_act) + b1, 0.1) stays in shared memory (bf16); v = conv2(t) + b2 + resid; v = v * beta + resid2 (MRF accumulation, optional); out = v (fp32, optional); out_act = bf16(lrelu(v, slope)) (optional). x_act = padded channels-last bf16 activation [B, L + 2*halo, 64] with zero halos (the 32-channel stage arrives time-folded). Both convs are lists of K-half blocks (host arrays): block i = bf16 weights w[:, 32 i .. 32 i + 31] ([64 output channels x 32 input channels], w is [64, n*32]) applied to the activation row at offset off[i] relative to the output row and to input-channel half kh[i] (0 / 1); a plain tap is two blocks, the zero halves of the block-sparse time-folded convs are not listed. resid / resid2 / out / out_act start at time 0 of item 0 with row (ld) and batch strides in elements.",
    "kr_hifi_resblock_resident": "Non-zero if kr_hifi_resblock can keep the weights of both convs (n1 / n2 K-half blocks of 4 KB) resident in shared memory next to its activation slabs — the only regime it supports: the number of tile slots it will run with (2 = two tiles ping-pong, 1); 0 = use two kr_gemm_ex launches.",
    "kr_hifi_post_tanh": "conv_post (C -> 1, k=7, pad 3) + tanh on the channels-last activation, inference/hifigan_vocoder.py:131-132.",
    "kr_wave_peak": "peak[b] = max |wav[b, :len[b]]| — the peak normalisation x / (max|x| + 1e-9) of data/dataset.py:672 is applied inside kr_mel_stft.",
    "kr_mel_stft": "Log-mel features out[B, n_mels, frames_max] = log(melfb(|STFT|^2) + log_eps): reflect pad 512, periodic Hann 1024, hop 256, 513 bins, dense filterbank fb_t[n_mels, 513] (HTK, norm=None) with optional fb_ranges[n_mels][2] = [first non-zero bin, one past the last) so that the exact zeros of the triangular filters are skipped (bit-identical to the dense product); frames beyond 1 + len//256 are zero. Replaces torchaudio.transforms.MelSpectrogram + log of data/dataset.py:162-178,694-697.",
    "kr_pitch_frames": "Per-frame YIN / CMND analysis of PitchExtractor.extract_pitch, model/variance_predictor.py:492-563: zero pad to 2048, pre-emphasis 0.97, reflect pad 1024, periodic Hann 2048, 4096-point FFT autocorrelation, cumulative-mean-normalised difference over the lags sample_rate/fmax .. sample_rate/fmin, first dip below 0.15 else the minimum, parabolic interpolation. One CTA per (frame, utterance); writes the frequency candidate, the autocorrelation peak and the mean frame energy to [B, frames_max] fp32 rows.",
    "kr_pitch_track": "Per-utterance part of extract_pitch, variance_predictor.py:566-615: voicing threshold clip(0.8 * quantile_25(ac peak), .15, .35), energy gate 0.05 * median, fmin / fmax gate, linear interpolation of unvoiced gaps <= 5 frames, median-5 with reflect padding, normalisation to [0, 1] (0 = unvoiced). Quantiles are exact order statistics by rank counting (no sort). Frames beyond 1 + max(len, 2048) // 256 are zero.",
    "kr_energy_frames": "Per-frame energy of EnergyExtractor.extract_energy_from_mel, variance_predictor.py:659-668: mean over the mel bins (log-domain input) or log1p(max(mean, 0)) (linear power, what data/dataset.py:813 passes); mel is (B, T, n_mels) [time_major = 1] or (B, n_mels, T) [0, the layout kr_mel_stft writes]; exp_input = 1 exponentiates the input first so that kr_mel_stft's log-mel can feed the linear-power branch.",
    "kr_energy_norm": "5 / 95-percentile normalisation of the per-frame energies to [0, 1] per utterance (min / max below 3 frames), variance_predictor.py:675-687; zeros beyond frames[b].",
    "kr_dec_state_size": "sizeof the device-resident generation state (krd::DecState in csrc/kr_decode_core.cuh: t, done, n_frames, lo, hi, expected, stop_thr, post_thr, ring[30]; 256 bytes; mirrored by kokoro_ruslan_b200/inference.py STATE_FIELDS).",
    "kr_dec_feed": "Decoder input of frame t of the autoregressive loop: mel_projection_in on the previous output frame + bias + PE row t (model/model.py:519-531 in eval mode, model/generator.py:52-57); t is read from the device state.",
    "kr_dec_attn": "One decode-step attention per (utterance, head) with the KV cache of model/transformers.py:237-277: per-head RMSNorm of the new query (rotated as position 0 — the reference's q_offset = 0), and for self-attention RMSNorm + RoPE(position t) of the new key and RMSNorm of the new value appended to the bf16 caches, then softmax(q K^T * scale) V over the cached rows; cross-attention reads the pre-normalised memory keys / values with the key-padding mask. rotate_q = 1 rotates the query to position t instead (opt-in fix of that train / inference mismatch).",
    "kr_dec_finish": "End of a decode step: decoder.norm + mel_projection_out + stop head (model/model.py:547-563) and the generator's stop rules on the device (model/generator.py:58-103: min / max length, stop probability mean over the batch against the pre / post-expected-length thresholds, 30-frame silence rule); stores the clamped frame, the un-clamped feedback frame and the stop probability, advances t.",
    "kr_val_metrics_acc_floats": "Length (floats) of the validation-metrics accumulator of kr_val_metrics.",
    "kr_val_metrics": "Validation metrics of KokoroTrainer.validate_epoch as device reductions, training/trainer.py:1868-1916: per utterance ||mel - pred||_F / ||mel||_F over its valid frames (spectral convergence) and sqrt(mean((pitch - pred)^2)) (frame-level F0 RMSE); the batch means of the valid utterances are added to epoch accumulators by the last block to finish. Replaces a Python loop with four .item() syncs per utterance.",
    "kr_resample_length": "ceil(new_freq * n / orig_freq): the output length of torchaudio.functional.resample for n input samples.",
    "kr_resample": "Hann-windowed sinc resampling with torchaudio.functional.resample's defaults (lowpass_filter_width 6, rolloff 0.99) — the speed perturbation of data/dataset.py:674-684. Every output sample evaluates its own ~14 taps (float64 like torchaudio's filter-bank construction, including its float32 phase offsets, rounded to float32) instead of a [new, 2 * width + orig] filter bank and a strided conv1d; rows are zero beyond ceil(new * len[b] / orig).",
    "kr_average_by_duration": "Frame-level values averaged to token level through the durations, bit-identical to the scatter formulation of utils/lengths.py:156-208 (clamped starts / ends, frames no token covers fall to token 0, masked or zero-duration tokens give 0); sums run in ascending frame order per token.",
    "kr_dec_gemv": "Skinny projection of a decode step (B <= 8 rows): out = x . W^T (+ bias) (+ residual), the weight matrix streamed once by the whole grid, one warp per output feature, fp32 accumulation; optional LayerNorm prologue (the pre-norm of model/transformers.py:564,572,581) and GLU epilogue (gelu_erf(gate) * lin, :105-108) — the weight-bandwidth-shaped replacement of the nn.Linear calls of model/transformers.py:228,258-259,434,105-111 inside the autoregressive loop (model/generator.py:44-103).",
    "kr_trim_end": "Trailing-silence trim point of a generated mel before vocoding, inference/inference.py:590-621: threshold clamp(0.5 * (q10 + q20) of the frame means, -9.8, -9.2), last frame above it + 24 frames of margin, at least 60 frames, at most T; T when nothing is above the threshold. Input = kr_energy_frames' per-frame means.",
    "kr_spec_augment": "SpecAugment on the cross-attention memory (bf16) or its gradient (fp32): zeroes the per-sample frame / hidden-dim spans in spans[B, n_time+n_feat, 2] = (start, length). training/trainer.py:1578-1604, applied at model/model.py:636-639.",
    "kr_attn_fwd": "tcgen05 flash attention forward, head_dim 64, on token-major [B,S,H,64] bf16 tensors (q_ss/q_bs = seq/batch strides in elements). Causal and per-key padding (key_mask[B,Sk], 1 = masked) are predicates; lse[B,H,Sq] is the log2-domain log-sum-exp kept for the backward. Replaces F.scaled_dot_product_attention with the dense additive mask, model/transformers.py:299-316,393-398.",
    "kr_attn_bwd": "Flash attention backward: dq (fp32 [B,Sq,H,64], zeroed by the call's own prep kernel, then atomically accumulated), dk/dv (bf16). delta[B,H,Sq] is scratch. Autograd of model/transformers.py:393-398.",
    "kr_stop_head_fwd": "Stop-token logits z[n] = x[n,:].w + b on the (detached) decoder output, model/model.py:562.",
    "kr_stop_head_bwd": "Weight/bias gradient of the stop head (no data gradient: the input is detached, model/model.py:562).",
    "kr_losses_fwd_bwd": "Fused masked losses + gradients w.r.t. the five model outputs: L1 mel, Huber(1) log1p-duration, BCE-with-logits(pos_weight) stop, Huber(delta_var) pitch/energy, clamps 100/100/100/10/10, weighted total; losses[6] = total, mel, dur, stop, pitch, energy. training/losses.py:9-216, criteria training/trainer.py:410-444.",
    "kr_layernorm_fwd": "LayerNorm over the last dim (D in {128,256,512}), fp32 in, bf16 and/or fp32 out, saves mean/rstd. model/transformers.py:478,485,564,572,581,660; model/model.py:122.",
    "kr_layernorm_bwd": "LayerNorm backward fused with the residual-gradient add (dx = dres + ...), optional bf16 copy of dx (with the dropout / stochastic-depth factors of the residual branch it feeds, and its column sums = that branch's output-bias gradient accumulated into dcol_bf16), dgamma/dbeta accumulated with atomics.",
    "kr_rmsnorm_resid_fwd": "FFN output RMSNorm(eps = fp32 finfo.eps) + residual add, model/transformers.py:94,109-111.",
    "kr_resid_drop_ln_fwd": "Tail of an attention sub-layer fused with the LayerNorm that follows it: out = resid + dropout(y) (y = the out-projection's fp32 output incl. bias; the dropout / stochastic-depth factors and their element index row * D + col are those of kr_gemm_ex's dropout epilogue, model/transformers.py:482-483,569,578) and h = LayerNorm(out) (bf16 and / or fp32 copy, mean / rstd saved for the backward), model/transformers.py:474,563 pre-norm blocks. Bit-identical to the GEMM epilogue + kr_layernorm_fwd it replaces.",
    "kr_rmsnorm_resid_ln_fwd": "kr_rmsnorm_resid_fwd fused with the LayerNorm of the next sub-layer (or the final norm): out as there, h = LayerNorm(out), mean / rstd saved.",
    "kr_rmsnorm_resid_bwd": "Backward of kr_rmsnorm_resid_fwd w.r.t. y (bf16) and the gain; dcol += column sums of dy (bias gradient of linear2).",
    "kr_qkv_prep_fwd": "Per-head RMSNorm(64) with learned gain on up to three column blocks (q|k|v) + rotate-half RoPE on the blocks selected by rope_mask; position = row % S. model/transformers.py:145-148,260-272; model/positional_encoding.py:196-209.",
    "kr_qkv_prep_bwd": "Backward of kr_qkv_prep_fwd: incoming gradients may be fp32 (grad_f32_mask) or bf16; writes d(raw projection) bf16 and accumulates the gain gradients.",
    "kr_optim_ctrl_size": "sizeof the device-resident optimizer control block (64 bytes; layout in kokoro_ruslan_b200/optim.py CTRL_FIELDS).",
    "kr_grad_sqnorm": "Per-tensor squared L2 norms of the flat gradient buffer + non-finite flag. Replaces the 308 `.norm().item()` + 2x308 `isfinite().all()` host syncs of training/trainer.py:2355-2362,1308-1313.",
    "kr_step_control": "Device-side step control: per-tensor spike pre-clip scales (training/trainer.py:1332-1407), total norm, explosion detector (trainer.py:1315-1330,2367-2405), clip coefficient (clip_grad_norm_, training/runtime_policies.py:33-79), skip-on-non-finite (trainer.py:2407-2463), bias corrections.",
    "kr_adamw_step": "Fused clip + AdamW (per-group lr / weight decay, trainer.py:446-689) + EMA (trainer.py:1491-1517) + bf16 shadow write over the flat buffers; accumulates squared norms of the tensors subject to the weight-norm projection.",
    "kr_wn_project": "Post-step projection ||W||_2 <= limit of the decoder FFN matrices, training/trainer.py:883-912.",
    "kr_glu_fwd": "u = gelu_erf(gate) * lin over h = [gate | lin], model/transformers.py:105-108.",
    "kr_glu_bwd": "Backward of kr_glu_fwd.",
    "kr_colsum_bf16": "out[c] += sum_n x[n,c] (bias gradients).",
    "kr_embed_fwd": "Encoder input emb[idx]*sqrt(D) + stress_emb[s] + PE[n % P], model/model.py:375-378, model/positional_encoding.py:66-74.",
    "kr_embed_bwd": "Scatter-add into the token / stress embedding gradients (stress row 0 = padding_idx gets none, model/model.py:92).",
    "kr_shift_cast": "Teacher-forcing shift-right of the mel target + bf16 cast, model/model.py:519.",
    "kr_cast_bf16": "fp32 -> bf16 copy (initial shadow of the master weights).",
    "kr_scatter_rows": "dst[map[r]] = src[r] row copy (fp32 -> bf16/fp32): token order -> predictor padded layout.",
    "kr_gather_rows": "dst[r] = src[map[r]] row copy (fp32), zero rows where map[r] < 0.",
    "kr_eq_mask_i64": "Byte mask idx == value: the reference text padding mask `phoneme_indices == 0`, model/model.py:586-587.",
    "kr_nonfinite_flag": "Sets `bit` in *flag if x holds a NaN/Inf — the finite-output guard of training/trainer.py:3233-3256 without host syncs.",
    "kr_lr_index": "LengthRegulator index tensor: idx[b,f] = min{j : cumsum(max(0,d))[b,j] > f} for f < L[b], else -1; lengths[b] = L[b]. Bit-exact restatement of utils/lengths.py:16-96 (repeat_interleave + scatter on the CPU in the reference).",
    "kr_lr_index_masked": "Index tensor of the `length_regulate` fallback (use_variance_predictor=False path, utils/lengths.py:108-153): padded tokens (pad_mask = 1) are skipped, every other duration is clamped to >= 1; lengths[b] = expanded length.",
    "kr_expand_rows_fwd": "Duration-expand gather out[b,f,:] = x[b, idx[b,f], :] (zeros and frame_mask = 1 where idx < 0), exact fp32 copy: torch.repeat_interleave + left-packed scatter of utils/lengths.py:139-147.",
    "kr_expand_rows_bwd": "Backward of the gather: deterministic segment sums dx[b,j,:] = sum of dout over the frames of token j (the fallback path keeps autograd, unlike the detached LengthRegulator).",
    "kr_range_flag": "flag |= any(x > 1 or x < 0): the data-dependent normalisation test of model/variance_predictor.py:244,268.",
    "kr_expand_adapt": "Duration-expand gather + pitch/energy bucketize(255 bins) + embedding add + frame masks; writes the predictor input (padded layout) and the decoder memory aligned to the mel length. model/variance_predictor.py:345-437, model/model.py:607-628.",
    "kr_adapt_bwd": "Memory gradient -> pitch/energy embedding rows (the only path through the detached expansion, utils/lengths.py:30).",
    "kr_gn_fwd": "GroupNorm(1 group, eps 1e-5) + ReLU per independent (sample, 512-frame chunk) on the padded layout, model/variance_predictor.py:55,70-115.",
    "kr_gn_bwd": "Backward of kr_gn_fwd (ReLU mask recomputed), dgamma/dbeta accumulated.",
    "kr_vp_head_fwd": "Predictor head Linear(F->1) + masked_fill(mask, 0); a chunk of length 1 yields zeros, model/variance_predictor.py:95-115.",
    "kr_vp_head_bwd": "Backward of kr_vp_head_fwd.",
    "kr_drop_begin": "Once per training forward: state[1] += 1 (new dropout step) and table[s, b] = stochastic-depth factor of residual branch s for sample b (0 with probability path_p[s], else 1/(1-p)): drop_path of model/transformers.py:16-40 with the per-layer rates of model/model.py:99-107.",
    "kr_dec_in_drop": "Decoder input y = drop_b(drop_a(t) + PE[row % T]) for t = mel_projection_in(shifted mel): F.dropout(p = decoder_input_dropout) then PositionalEncoding's own dropout, model/model.py:525-531, model/positional_encoding.py:72-74. scale_a = 1/keep_a; drop->scale = 1/(keep_a*keep_b).",
    "kr_drop_export_mask": "Test aid: out[r*cols + c] = keep(element r*ld + c) of one dropout site as bytes, so a CPU oracle can apply exactly the masks the fused kernels regenerate.",
    "kr_allreduce_sqnorm": "Data-parallel gradient all-reduce FUSED with the optimizer's squared-norm pass, over symmetric memory (NVSwitch multicast multimem.ld_reduce / multimem.st when mc_grads != NULL, else peer loads / stores): chunk c is reduced by rank c % world, broadcast, and its squared sum written to sq[.][c] on every rank. grads / sq / flags are HOST arrays of `world` device pointers (each rank's buffers as mapped into this process); flags = zero-initialised unsigned [grid * world] per rank. chunk_ranges = HOST {begin0, end0, begin1, end1} chunk-index ranges to reduce (NULL = all): the step reduces the ranges that are final half-way through the backward from a side stream underneath the rest of it. clip_local / clip_out (device scalars, both or neither): this rank's clip norm for the step in, the minimum over all ranks out (slot n_chunks + rank of the sq buffers carries it, so they hold n_chunks + world floats). New functionality (the reference is single-process); replaces the per-parameter norm loop of training/trainer.py:2355-2362 on the reduced buffer.",
    "kr_chunk_sqnorm": "Per-chunk squared sums of the flat gradient buffer (single-GPU first optimizer phase; deterministic, no atomics).",
    "kr_chunk_to_tensor_sq": "sq[t] = sum over the chunks of tensor t in chunk order (first_chunk[n_tensors + 1]); raises the control block's non-finite flag (training/trainer.py:1308-1313).",
    "kr_conv_dgrad_shadow": "bf16 tap-reversed transpose of a tap-major conv weight: the B operand of the conv data-gradient GEMM.",
    "kr_conv_dgrad_shadow_multi": "kr_conv_dgrad_shadow for n <= 8 convs of one (Co, Ci) shape in one launch: w2 / wd are HOST arrays of n device pointers.",
}

STRUCTS = """/* Dropout / stochastic-depth descriptor of ONE site (HOST struct, passed by pointer; NULL or
 * state == NULL = disabled).  The keep decision of element e is a pure function of
 * (state[0] = seed, state[1] = step, site, e): keep iff lane16(hash(e >> 1, key(seed, step, site)), e & 1) >= thr
 * with thr = round(p * 65536), so backward kernels regenerate the mask instead of storing it
 * (kr_common.cuh).  Kept elements are multiplied by `scale` (1 / keep probability, product over
 * both masks) and, when row_scale != NULL, by row_scale[row / rows_per_sample] — the per-sample
 * stochastic-depth factor (0 or 1 / (1 - p_path)) written by kr_drop_begin. */
typedef struct kr_drop_spec {
  const unsigned long long* state;   /* device {seed, step} */
  unsigned int site_a, thr_a;        /* first mask (thr_a == 0: none) */
  unsigned int site_b, thr_b;        /* optional second, independent mask (thr_b == 0: none) */
  float scale;
  const float* row_scale;            /* device [n_samples] or NULL */
  int rows_per_sample;
} kr_drop_spec;

/* Argument block of kr_gemm_ex (plain C, zero-initialise then fill what you need). */
typedef struct kr_gemm_args {
  const void* A;            /* bf16; logical [M,K]: stored [M,K] (K-major) or [K,M] if a_mn_major */
  const void* B;            /* bf16; logical [N,K]: stored [N,K] (K-major) or [K,N] if b_mn_major */
  int M, N, K, batch;       /* per batch item; stride_b == 0 with batch > 1 = weights shared */
  long long lda, ldb, stride_a, stride_b;
  int a_mn_major, b_mn_major;
  /* implicit-GEMM conv1d on A: channels-last activation [batch, a_rows, conv_cin] that already
   * contains its zero halos; K = conv_taps * conv_cin, K-block (tap, chunk) is read at row
   * conv_row0 + m + tap * conv_dil.  conv_taps == 0: plain GEMM. */
  int conv_taps, conv_dil, conv_row0, conv_cin, a_rows;
  /* epilogue: v = alpha*acc + bias[n] + resid[m % resid_mod, n];  v = v*beta + resid2[m, n] */
  float alpha, beta;
  const float* bias;
  const void* resid;  int resid_dtype;  long long ldr, stride_r;  int resid_mod;   /* dtype 0 f32, 1 bf16 */
  const void* resid2; int resid2_dtype; long long ldr2, stride_r2;
  void* C;  int c_mode; long long ldc, stride_c;    /* 0 bf16, 1 f32, 2 f32 atomic add, 3 none */
  void* C2; long long ldc2, stride_c2; float act_slope;   /* optional bf16 leaky_relu(v, act_slope) */
  int splits;               /* split-K (c_mode 2 only) */
  int force_block_n;        /* 0 = heuristic, else 64 / 128 / 192 / 256 */
  int no_slab;              /* 1 = conv mode re-fetches every tap (debug / A-B comparison) */
  /* dropout on the GEMM result BEFORE the residual adds: v = drop(alpha*acc + bias) + resid ...; element index
   * m * N + n (single batch only).  transformers.py:482-483,569,578 (dropout(drop_path(attn_out)) + residual) */
  const kr_drop_spec* drop;
} kr_gemm_args;
"""


def main():
    protos = []
    for f in sorted(glob.glob(os.path.join(ROOT, "kokoro_ruslan_b200", "csrc", "*.cu"))):
        s = open(f).read()
        group = []
        for m in re.finditer(r'extern "C"\s+((?:const\s+)?(?:long\s+long|\w+)\*?)\s+(kr_\w+)\s*\(([^)]*)\)', s):
            ret, name, args = m.groups()
            args = re.sub(r"\s+", " ", args.strip())
            if name not in DOCS:
                sys.exit(f"gen_header: no DOCS entry for {name}")
            group.append((ret, name, args))
        if group:
            protos.append((os.path.basename(f), group))
    out = ["/* kokoro_b200.h — C ABI of libkokoro_b200.so (GENERATED by tools/gen_header.py; do not edit).",
           " *",
           " * Drop-in boundary of the Blackwell-native hot path of igorshmukler/kokoro-ruslan.  The reference",
           " * has no FFI of its own (it is pure PyTorch); each entry point below replaces the ATen/torch call",
           " * sites cited in its comment (paths relative to the reference's src/kokoro/).",
           " *",
           " * Conventions: plain pointers and sizes only (no torch types); all pointers are DEVICE pointers",
           " * unless noted; the caller allocates every output / workspace; every launch goes to the",
           " * cudaStream_t passed as `stream` (graph-capturable, no allocation, no synchronisation);",
           " * return 0 on success or a negative KR_ERR_* code with a message in kr_last_error().",
           " */",
           "#ifndef KOKORO_B200_H", "#define KOKORO_B200_H", "", "#ifdef __cplusplus", 'extern "C" {', "#endif", "",
           "#define KR_OK 0", "#define KR_ERR_ARG (-1)", "#define KR_ERR_CUDA (-2)", "#define KR_ERR_TMAP (-3)",
           "#define KR_ERR_UNSUPPORTED (-4)", "", STRUCTS]
    for fname, group in protos:
        out.append(f"/* ---- {fname} " + "-" * max(4, 88 - len(fname)) + " */")
        for ret, name, args in group:
            doc = DOCS[name]
            words, line, lines = doc.split(), "", []
            for w in words:
                if len(line) + len(w) + 1 > 96:
                    lines.append(line)
                    line = w
                else:
                    line = (line + " " + w).strip()
            lines.append(line)
            out.append("/* " + ("\n * ".join(lines)) + " */")
            proto = f"{ret} {name}({args});"
            # wrap long prototypes
            if len(proto) > 100:
                parts = args.split(", ")
                cur = f"{ret} {name}("
                wrapped = []
                for i, p in enumerate(parts):
                    piece = p + (", " if i < len(parts) - 1 else ");")
                    if len(cur) + len(piece) > 100:
                        wrapped.append(cur.rstrip())
                        cur = "    " + piece
                    else:
                        cur += piece
                wrapped.append(cur)
                proto = "\n".join(wrapped)
            out.append(proto)
            out.append("")
    out += ["#ifdef __cplusplus", "}", "#endif", "#endif /* KOKORO_B200_H */", ""]
    path = os.path.join(ROOT, "include", "kokoro_b200.h")
    open(path, "w").write("\n".join(out))
    print(f"wrote {path}: {sum(len(g) for _, g in protos)} entry points")


if __name__ == "__main__":
    main()
