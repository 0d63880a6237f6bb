// kr_attn.cu — tcgen05 flash attention, forward + backward, head_dim 64
// (reference: MultiHeadAttentionImproved.forward, model/transformers.py:299-316,393-398 — SDPA with an
//  additive causal(-inf) + key-padding(-inf) mask and scale 1/sqrt(d_k)).
//
// Q/K/V/O/dO live token-major as [batch, seq, heads, 64] bf16 (arbitrary seq/batch strides) and are
// reached through 4-D TMA maps (d, head, seq, batch), so the projections' [tokens, H*64] GEMM
// outputs are consumed in place.  The T x T mask is never materialised: causality and the per-key
// padding byte-mask are predicates on the S tile.
//
// forward : one CTA per (pair of 128-query tiles, head, batch), warp-specialised (TMA warp, MMA warp, two softmax
//           warpgroups ping-ponging): S = Q K^T lands in TMEM, each softmax thread owns one query row, P goes back
//           into TMEM as the bf16 A operand of O += P V (V is the MN-major B operand straight from its natural
//           [kv, d] tile), O accumulates in TMEM with lazy rescaling.
// backward: one CTA per (128-key tile, head, batch) looping over query tiles: S and dP = dO V^T in
//           TMEM, P / dS rebuilt per row, then dV += P^T dO, dK += dS^T Q (P, dS as MN-major A
//           operands, dO / Q as MN-major B operands — no transposed copies) and dQ_tile = dS K,
//           which is reduced into the fp32 dQ buffer with vector atomics.
#include "kr_common.cuh"
#include "kokoro_b200.h"
#include <stdlib.h>

namespace {
using namespace kr;

constexpr int TQ = 128, TK = 128, HD = 64;
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB: 128 rows x 128 B

struct AttnParams {
  int B, H, Sq, Sk;
  float scale, scale_log2;
  const uint8_t* key_mask;  // [B, Sk] 1 = masked key, or null
  bf16* O; long long o_ss, o_bs;
  float* lse;               // [B, H, Sq], log2 domain
  const float* delta;       // [B, H, Sq]
  float* dQ; long long dq_ss, dq_bs;
  bf16* dK; long long dk_ss, dk_bs;
  bf16* dV; long long dv_ss, dv_bs;
  // attention-probability dropout (F.scaled_dot_product_attention(dropout_p), transformers.py:393-398): element
  // index of P[b, h, q, k] = ((b*H + h)*Sq + q) * sk_pad + k, sk_pad = Sk rounded up to the 128-key tile.
  // This site uses the generator's BYTE-lane mode (one 32-bit hash serves 4 keys, thr quantised to 1/256): the
  // softmax threads of these kernels are issue-bound and the 16-bit mode doubled the forward's time.
  DropSpec drop;
  uint32_t sk_pad_quarter, thr4;     // thr4 = the 8-bit threshold replicated into 4 bytes
};

// keep bytes (0xff / 0x00) of the 4 keys served by hash `quad`
__device__ __forceinline__ uint32_t drop_quad_bytes(uint32_t quad, uint2 key, uint32_t thr4) {
  return __vcmpgeu4(drop_hash(quad, key), thr4);
}
// packed-bf16 AND masks of key pairs (0,1) and (2,3) of a quad from its keep bytes
__device__ __forceinline__ uint32_t keep_lo_pair(uint32_t m) { return __byte_perm(m, 0u, 0x1100); }
__device__ __forceinline__ uint32_t keep_hi_pair(uint32_t m) { return __byte_perm(m, 0u, 0x3322); }

// [128 rows x 64 K] bf16 tile, K-major, SW128: k-th UMMA_K slice.
__device__ __forceinline__ uint64_t desc_k64(uint32_t base, int k) {
  return make_smem_desc_sw128(base + k * 32, 16, 1024);
}
// [128 rows x 128 K] bf16 tile stored as two 64-wide K chunks (16 KB each), K-major.
__device__ __forceinline__ uint64_t desc_k128(uint32_t base, int k) {
  return make_smem_desc_sw128(base + (k >> 2) * TILE_BYTES + (k & 3) * 32, 16, 1024);
}
// Same two-chunk tile read as an MN-major operand: MN = the 128 contiguous columns (two 64-chunks,
// LBO apart), K = the rows (8-row groups SBO apart); k-th slice = rows [16k, 16k+16).
__device__ __forceinline__ uint64_t desc_mn128(uint32_t base, int k) {
  return make_smem_desc_sw128(base + k * 2048, TILE_BYTES, 1024);
}
// [rows x 64] tile read as MN-major operand with MN = the 64 columns, K = rows.
__device__ __forceinline__ uint64_t desc_mn64(uint32_t base, int k) {
  return make_smem_desc_sw128(base + k * 2048, 2048, 1024);
}

// store 32 packed-bf16 columns [32c, 32c+32) of `row` into a two-chunk [128 x 128] SW128 tile
__device__ __forceinline__ void store_chunk_sw128(uint8_t* tile, int row, int c, const uint32_t* pk) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int u = c * 4 + q;
    uint8_t* dst = tile + (u >> 3) * TILE_BYTES + row * 128 + (((u & 7) ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
}

// Bit e of the result is set iff key (key0 + e) may be attended by query row qi: not padded (pad_bits bit e
// clear) and, if causal, key <= qi.  All-ones is the common case and takes a predicate-free fast path.
__device__ __forceinline__ uint32_t allowed_bits(uint32_t pad_bits, bool causal, int key0, int qi) {
  uint32_t a = ~pad_bits;
  if (causal) {
    const int n = qi - key0 + 1;                       // number of leading keys of this chunk that are <= qi
    a &= n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
  }
  return a;
}

// ---------------------------------------------------------------------------------------------
// forward — warp-specialised, two query tiles ping-ponging on one SM
//
// One CTA per (PAIR of adjacent 128-query tiles, head, batch), 320 threads:
//   warps 0-3 = softmax warpgroup 0 (query tile 2i),  warps 4-7 = softmax warpgroup 1 (query tile 2i+1):
//               one thread per query row (= TMEM lane), the whole 128-key S row in registers
//   warp 8    = MMA issuer (one elected lane issues every tcgen05.mma / commit)
//   warp 9    = TMA producer (Q tiles once, K / V tiles through a 3-stage ring shared by both warpgroups)
// Tensor memory (all 512 columns): S_w (128 fp32 columns), P_w (64 columns of packed bf16 pairs) and O_w (64 columns)
// per warpgroup.  O stays in TMEM for the whole key loop — the P V MMA accumulates into it — and is rescaled in
// place only when a row's running maximum moves by more than 2^8 (lazy rescale: the softmax is computed against a
// possibly stale maximum, P <= 256 is harmless in bf16 / fp32; S is not overwritten by P, so the rare rescale simply
// redoes the pass), so the per-tile O read-back + 64 FMAs per row of the previous kernel are gone, and the S row is
// streamed through registers 32 keys at a time (no 128-register row, no spills).  While warpgroup 0 runs its softmax
// on the CUDA cores (MUFU-bound: 128 x 128 exp2 per tile), the tensor pipe runs warpgroup 1's P V and next
// Q K^T, and vice versa; the issue order per key tile j is  [P_0 ready] PV_0(j), S_0(j+1), [P_1 ready] PV_1(j), S_1(j+1).
// In-order MMA execution makes "S_w(j+1) complete" imply "PV_w(j) complete", so one barrier per warpgroup covers
// both the next softmax and the lazy rescale of O.  The 1/sqrt(d) scale is folded into the exp2 argument
// (packed FFMA2), row sums add the fp32 probabilities (FADD2).  Masks (causal, key padding) are predicates on the
// S registers; the key-padding bits of ALL key tiles are ballotted into shared memory once, in the prologue.
// ---------------------------------------------------------------------------------------------
constexpr int FWD_STAGES = 3;
constexpr int FWD_THREADS = 320;
constexpr int FWD_MAX_KTILES = 64;                                   // key-padding bit table: Sk <= 8192
constexpr int FWD_SMEM = (2 + 2 * FWD_STAGES) * TILE_BYTES + 256 + FWD_MAX_KTILES * 16 + 1024;
constexpr float FWD_RESCALE_LOG2 = 8.f;                              // rescale O when the row max grows by > 2^8

template <bool CAUSAL, bool DROP>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;                                   // [2] query tiles
  uint8_t* sK = smem + 2 * TILE_BYTES;                  // [FWD_STAGES]
  uint8_t* sV = sK + FWD_STAGES * TILE_BYTES;           // [FWD_STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FWD_STAGES * TILE_BYTES);
  uint64_t* q_bar = bars;                               // Q tiles landed (tx)
  uint64_t* k_full = bars + 1;                          // [FWD_STAGES] (tx)
  uint64_t* v_full = k_full + FWD_STAGES;               // [FWD_STAGES] (tx)
  uint64_t* kv_empty = v_full + FWD_STAGES;             // [FWD_STAGES] commit: every MMA reading the stage is done
  uint64_t* s_full = kv_empty + FWD_STAGES;             // [2] commit: S_w(j) (and with it PV_w(j-1)) complete
  uint64_t* p_ready = s_full + 2;                       // [2] one arrival per warp of the warpgroup: P_w(j) (and a rescaled O_w) are in TMEM
  uint64_t* o_final = p_ready + 2;                      // [2] commit: the last PV_w is complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 2);
  uint32_t* mask_words = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [n_ktiles][4]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_qtiles = (p.Sq + TQ - 1) / TQ;
  // 1-D grid, pair of query tiles OUTERMOST: CTAs are handed out in blockIdx order, so the whole machine takes the heavy
  // pairs first (longest-processing-time order).  Causal: the last pair attends to the most keys; otherwise the last pair
  // (often a single, partly filled tile) is the light one and goes last.  With the pair innermost (round 1-2: a 3-D grid)
  // every (head, batch) group put one heavy CTA at the END of the schedule: 13 vs 10 key-tile times for T = 800.
  const int per_pair = p.H * p.B;
  const int pi = (int)blockIdx.x / per_pair, hb = (int)blockIdx.x - pi * per_pair;
  const int pair = CAUSAL ? (n_qtiles + 1) / 2 - 1 - pi : pi;
  const int h = hb % p.H, b = hb / p.H;
  const int n_ktiles = (p.Sk + TK - 1) / TK;
  int n_w[2];
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const int qt = 2 * pair + w;
    n_w[w] = qt < n_qtiles ? (CAUSAL ? min(n_ktiles, qt + 1) : n_ktiles) : 0;
  }
  const int n_max = max(n_w[0], n_w[1]);

  pdl_launch_dependents();
  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_bar, 1);
    for (int i = 0; i < FWD_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int w = 0; w < 2; ++w) { mbar_init(&s_full[w], 1); mbar_init(&p_ready[w], 4); mbar_init(&o_final[w], 1); }
    fence_barrier_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  pdl_wait();
  // key-padding bits of every key tile (bit e of word [t][c] set = key 128 t + 32 c + e is padding / out of range)
  for (int wi = warp; wi < n_max * 4; wi += FWD_THREADS / 32) {
    const int key = wi * 32 + lane;
    bool masked = key >= p.Sk;
    if (!masked && p.key_mask != nullptr) masked = p.key_mask[(long long)b * p.Sk + key] != 0;
    const unsigned bits = __ballot_sync(0xffffffffu, masked);
    if (lane == 0) mask_words[wi] = bits;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
  // A partial last key tile is computed only on its 32-key chunks that hold real keys: S MMA with N = 32 * nck, the
  // softmax over nck chunks, P V over 32 * nck keys (at Sk = 800 the seventh tile has 32 keys: 3/4 of its work is skipped).
  auto nck_of = [&](int j) { return min(4, (p.Sk - j * TK + 31) >> 5); };

  if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_bar, TILE_BYTES * (n_w[1] > 0 ? 2 : 1));
      tma_load_4d(sQ, &tmQ, q_bar, 0, h, (2 * pair) * TQ, b);
      if (n_w[1] > 0) tma_load_4d(sQ + TILE_BYTES, &tmQ, q_bar, 0, h, (2 * pair + 1) * TQ, b);
      for (int j = 0; j < n_max; ++j) {
        const int st = j % FWD_STAGES;
        if (j >= FWD_STAGES) mbar_wait(&kv_empty[st], ((j / FWD_STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_4d(sK + st * TILE_BYTES, &tmK, &k_full[st], 0, h, j * TK, b);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_4d(sV + st * TILE_BYTES, &tmV, &v_full[st], 0, h, j * TK, b);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(q_bar, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        if (n_w[w] > 0) {
          const uint32_t aQ = smem_u32(sQ + w * TILE_BYTES), aK = smem_u32(sK);
          const uint32_t idesc_s = make_idesc_bf16(128, 32 * nck_of(0), 0, 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16_ss(tmem_base + w * 128, desc_k64(aQ, k), desc_k64(aK, k), idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&s_full[w]);
        }
      }
      for (int j = 0; j < n_max; ++j) {
        const int st = j % FWD_STAGES;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          if (j >= n_w[w]) continue;
          mbar_wait(&p_ready[w], j & 1);
          mbar_wait(&v_full[st], (j / FWD_STAGES) & 1);
          tc_fence_after();
          const uint32_t tS = tmem_base + w * 128, tP = tmem_base + 256 + w * 64, tO = tmem_base + 384 + w * 64;
          const uint32_t aV = smem_u32(sV + st * TILE_BYTES);
          const int ksteps = 2 * nck_of(j);       // 16 keys = 8 packed TMEM columns per MMA
          if (ksteps == TK / 16) {                // full tile: unrolled (one thread issues every MMA of the CTA)
#pragma unroll
            for (int k = 0; k < TK / 16; ++k)
              umma_bf16_ts(tO, tP + k * 8, desc_mn64(aV, k), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          } else {
            for (int k = 0; k < ksteps; ++k)
              umma_bf16_ts(tO, tP + k * 8, desc_mn64(aV, k), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          }
          if (j + 1 < n_w[w]) {
            const int ns = (j + 1) % FWD_STAGES;
            mbar_wait(&k_full[ns], ((j + 1) / FWD_STAGES) & 1);
            tc_fence_after();
            const uint32_t aQ = smem_u32(sQ + w * TILE_BYTES), aK = smem_u32(sK + ns * TILE_BYTES);
            const uint32_t idesc_s = make_idesc_bf16(128, 32 * nck_of(j + 1), 0, 0);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              umma_bf16_ss(tS, desc_k64(aQ, k), desc_k64(aK, k), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&s_full[w]);
          } else {
            umma_commit(&o_final[w]);
          }
        }
        umma_commit(&kv_empty[st]);               // S_*(j) and PV_*(j) were all issued before this point
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const int w = warp >> 2;                       // warpgroup = query tile of the pair
    const int row = tid & 127;                     // query row inside the tile = TMEM lane
    const int q0 = (2 * pair + w) * TQ, qi = q0 + row;
    const uint32_t t_lane = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + w * 128 + t_lane, tP = tmem_base + 256 + w * 64 + t_lane,
                   tO = tmem_base + 384 + w * 64 + t_lane;
    // a warp whose 32 rows all lie beyond Sq only keeps the barriers moving
    const bool warp_live = q0 + (warp & 3) * 32 < p.Sq;
    float m_run = -INFINITY, l_run = 0.f;
    uint2 dkey = make_uint2(0u, 0u);
    uint32_t drow = 0;
    if (DROP) {
      dkey = drop_key(p.drop.state, p.drop.site_a);
      drow = (uint32_t)(((long long)b * p.H + h) * p.Sq + min(qi, p.Sq - 1)) * p.sk_pad_quarter;
    }
    const int nt = n_w[w];
    // One pass over the S row in 32-key chunks (the next chunk's TMEM load is in flight while the current one is
    // processed): masks, running tile maximum of the RAW logits, P = exp2(s * scale - m_use) -> packed bf16 (+ dropout)
    // -> TMEM, fp32 row sum.  WRITE = false only finds the maximum (first key tile).
    auto row_pass = [&](int j, float m_use, bool write, float& tile_max, float& tile_sum) {
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_use, -m_use);
      float2 rs2 = make_float2(0.f, 0.f);
      float mx0 = -INFINITY, mx1 = -INFINITY;
      uint32_t buf[2][32];
      const int nck = nck_of(j);
      tmem_ld_32x32(tS, buf[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_wait();
        if (c + 1 < 4) tmem_ld_32x32(tS + (c + 1) * 32, buf[(c + 1) & 1]);   // (stale columns beyond nck are loaded, not used)
        if (c < nck) {                              // CTA-uniform: chunks beyond the last real key are skipped
        uint32_t (&r)[32] = buf[c & 1];
        const uint32_t ok = allowed_bits(mask_words[j * 4 + c], CAUSAL, j * TK + c * 32, qi);
        if (!__all_sync(0xffffffffu, ok == 0xffffffffu)) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = ((ok >> i) & 1u) ? r[i] : 0xff800000u;      // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(r[i]));
          mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        }
        if (write) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float2 a0 = __ffma2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, nm2);
            const float2 a1 = __ffma2_rn(make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), sc2, nm2);
            const float2 e0 = make_float2(fast_exp2(a0.x), fast_exp2(a0.y));   // exp2(-inf) = 0 for masked keys
            const float2 e1 = make_float2(fast_exp2(a1.x), fast_exp2(a1.y));
            rs2 = __fadd2_rn(rs2, e0);               // the softmax denominator is that of the un-dropped probabilities
            rs2 = __fadd2_rn(rs2, e1);
            uint32_t u0 = pack_bf16(e0.x, e0.y), u1 = pack_bf16(e1.x, e1.y);
            if (DROP) {
              const uint32_t m = drop_quad_bytes(drow + j * (TK / 4) + c * 8 + (i >> 2), dkey, p.thr4);
              u0 &= keep_lo_pair(m);
              u1 &= keep_hi_pair(m);
            }
            pk[i >> 1] = u0;
            pk[(i >> 1) + 1] = u1;
          }
          tmem_st_32x16(tP + c * 16, pk);            // packed: 32-bit column k holds keys (2k, 2k+1)
        }
        }
      }
      tile_max = fmaxf(mx0, mx1) * p.scale_log2;     // scale > 0: max commutes with it
      tile_sum = rs2.x + rs2.y;
    };
    for (int j = 0; j < nt; ++j) {
      mbar_wait(&s_full[w], j & 1);
      tc_fence_after();
      if (warp_live) {
        float mx, rs;
        if (j == 0) {                                // first key tile: find the maximum first
          row_pass(0, 0.f, false, mx, rs);
          m_run = mx;
        }
        row_pass(j, (m_run == -INFINITY) ? 0.f : m_run, true, mx, rs);
        // Lazy rescale: the pass above used the running (possibly stale) maximum; only when some row of the warp
        // outgrew it by more than 2^8 are O and l rescaled and the pass redone (warp-uniform branch: the TMEM accesses
        // are warp-collective).  S is still intact: P lives in its own TMEM columns.
        if (__any_sync(0xffffffffu, mx > m_run + FWD_RESCALE_LOG2)) {
          const float m_new = fmaxf(m_run, mx);
          const float alpha = fast_exp2(m_run - ((m_new == -INFINITY) ? 0.f : m_new));   // m_run = -inf -> 0
          if (j > 0) {                               // PV(j-1) is complete (s_full ordering): O can be touched
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t r[32];
              tmem_ld_32x32(tO + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_32x32(tO + c * 32, r);
            }
          }
          l_run *= alpha;
          m_run = m_new;
          row_pass(j, (m_run == -INFINITY) ? 0.f : m_run, true, mx, rs);
        }
        l_run += rs;
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();                                  // one elected arrival per warp (128 arrivals on one mbarrier word serialise)
      if (lane == 0) mbar_arrive(&p_ready[w]);
    }
    if (nt > 0) {
      mbar_wait(&o_final[w], 0);
      tc_fence_after();
      if (warp_live) {
        float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        if (DROP) inv *= p.drop.scale;
        bf16* orow = p.O + (long long)b * p.o_bs + (long long)qi * p.o_ss + h * HD;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tO + c * 32, r);
          tmem_ld_wait();
          if (qi < p.Sq) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 q;
              q.x = pack_bf16(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
              q.y = pack_bf16(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
              q.z = pack_bf16(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
              q.w = pack_bf16(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + c * 32 + 8 * i) = q;
            }
          }
        }
        if (qi < p.Sq) p.lse[((long long)b * p.H + h) * p.Sq + qi] = l_run > 0.f ? m_run + log2f(l_run) : 0.f;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------
// backward prep: delta[b,h,q] = sum_d dO * O   (one warp per token row of H*64 elements)
// ---------------------------------------------------------------------------------------------
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ O, long long o_ss, long long o_bs,
                                     const bf16* __restrict__ dO, long long do_ss, long long do_bs,
                                     float* __restrict__ delta, int B, int H, int Sq,
                                     float* __restrict__ dq, long long dq_ss, long long dq_bs) {
  pdl_entry();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * Sq) return;
  const int b = row / Sq, q = row % Sq;
  if (dq != nullptr) {      // the main kernel accumulates dQ with atomics: zero it here instead of a separate memset
    float4* z = reinterpret_cast<float4*>(dq + (long long)b * dq_bs + (long long)q * dq_ss);
    for (int v = lane; v < H * 16; v += 32) z[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const bf16* o = O + (long long)b * o_bs + (long long)q * o_ss;
  const bf16* d = dO + (long long)b * do_bs + (long long)q * do_ss;
  // H*64 elements, 8 per 16-byte vector -> H*8 vectors; vector v belongs to head v/8
  for (int v0 = 0; v0 < H * 8; v0 += 32) {
    const int v = v0 + lane;
    float acc = 0.f;
    if (v < H * 8) {
      const uint4 a = *reinterpret_cast<const uint4*>(o + v * 8);
      const uint4 g = *reinterpret_cast<const uint4*>(d + v * 8);
      const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = unpack_bf16(aa[i]), y = unpack_bf16(gg[i]);
        acc += x.x * y.x + x.y * y.y;
      }
    }
    // reduce groups of 8 lanes (one head)
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if ((lane & 7) == 0 && v < H * 8) delta[((long long)b * H + (v >> 3)) * Sq + q] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// backward
//
// One CTA per (128-key tile, head, batch), looping over query tiles i.  Per iteration:
//   MMA1(i): S = Q K^T, dP = dO V^T                     -> TMEM (128 cols each)
//   soft(i): P = exp2(S*scale - lse), dS = P*(dP - delta)*scale  -> bf16 smem tiles (8 warps: lane group =
//            warp % 4 = 32 query rows, column half = warp / 4 = 64 keys)
//   MMA2(i): dV += P^T dO, dK += dS^T Q (TMEM, accumulated over i), dQ_i = dS K (TMEM, fresh)
//   dQ_i is read back by the 8 warps and reduced into the fp32 dQ buffer with vector atomics.
// A 9th CONTROL warp issues all TMA loads and tcgen05.mma; the issue order "MMA1(i+1), then MMA2(i)"
// lets soft(i+1) run on the CUDA cores while the tensor pipe executes MMA2(i).  Q/dO and P/dS tiles are
// double-buffered in shared memory (224 KB), S/dP/dQ are single-buffered in TMEM and handed over with
// mbarriers (sdp: MMA1 done, out: MMA2 done, soft: the 256 compute threads are done with iteration i).
// ---------------------------------------------------------------------------------------------
constexpr int BWD_SMEM = 14 * TILE_BYTES + 256 + 1024;
constexpr int BWD_THREADS = 288;

template <bool CAUSAL, bool DROP>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE_BYTES;
  uint8_t* sQ = smem + 2 * TILE_BYTES;    // [2] buffers
  uint8_t* sdO = smem + 4 * TILE_BYTES;   // [2]
  uint8_t* sP = smem + 6 * TILE_BYTES;    // [2] x 32 KB
  uint8_t* sdS = smem + 10 * TILE_BYTES;  // [2] x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 14 * TILE_BYTES);
  uint64_t* kv_bar = bars;
  uint64_t* qdo_bar = bars + 1;   // [2]
  uint64_t* s_bar = bars + 3;     // S(it) complete: the exp2 half of the softmax starts while dP is still in the tensor pipe
  uint64_t* dp_bar = bars + 4;    // dP(it) complete
  uint64_t* out_bar = bars + 5;
  uint64_t* soft_bar = bars + 6;  // [2] one per half: one arrival per warp (4 each)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint32_t* mask_words = tmem_slot + 2;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 1-D grid, key tile outermost: the CTAs of key tile 0 (the longest under a causal mask) are scheduled first and those
  // of the last, usually partial, key tile last — they fill the tail of the final wave (448 CTAs on 148 SMs at 8 x 8 x 800)
  const int per_tile = p.H * p.B;
  const int ktile = (int)blockIdx.x / per_tile, hb = (int)blockIdx.x - ktile * per_tile;
  const int kv0 = ktile * TK, h = hb % p.H, b = hb / p.H;
  const int n_q_tiles = (p.Sq + TQ - 1) / TQ;
  const int i_begin = CAUSAL ? ktile : 0;
  const int n_it = n_q_tiles - i_begin;
  // partial key tile: only the 32-key chunks that hold real keys are computed — S / dP MMAs with N = 32 * nck, the
  // softmax on nck chunks, the dQ contraction over 32 * nck keys.  The other columns of the P / dS tiles are never
  // written; they only reach rows of dV / dK beyond Sk, which are not stored.
  const int nck = (min(TK, p.Sk - kv0) + 31) >> 5;
  pdl_launch_dependents();

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(kv_bar, 1); mbar_init(&qdo_bar[0], 1); mbar_init(&qdo_bar[1], 1);
    mbar_init(s_bar, 1); mbar_init(dp_bar, 1); mbar_init(out_bar, 1);
    mbar_init(&soft_bar[0], 4); mbar_init(&soft_bar[1], 4);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  pdl_wait();
  if (tid < 128) {
    const int key = kv0 + tid;
    bool masked = key >= p.Sk;
    if (!masked && p.key_mask != nullptr) masked = p.key_mask[(long long)b * p.Sk + key] != 0;
    const unsigned w = __ballot_sync(0xffffffffu, masked);
    if (lane == 0) mask_words[warp] = w;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dV = tmem_base + 256,
                 tmem_dK = tmem_base + 320, tmem_dQ = tmem_base + 384;
  const uint32_t idesc_s = make_idesc_bf16(128, 32 * nck, 0, 0);   // S, dP
  constexpr uint32_t idesc_tt = make_idesc_bf16(128, 64, 1, 1);    // dV, dK
  constexpr uint32_t idesc_dq = make_idesc_bf16(128, 64, 0, 1);    // dQ

  if (warp == 8) {
    // ===================== control warp: TMA + MMA issue =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_bar, 2 * TILE_BYTES);
      tma_load_4d(sK, &tmK, kv_bar, 0, h, kv0, b);
      tma_load_4d(sV, &tmV, kv_bar, 0, h, kv0, b);
      for (int j = 0; j < 2 && j < n_it; ++j) {
        mbar_arrive_expect_tx(&qdo_bar[j], 2 * TILE_BYTES);
        tma_load_4d(sQ + j * TILE_BYTES, &tmQ, &qdo_bar[j], 0, h, (i_begin + j) * TQ, b);
        tma_load_4d(sdO + j * TILE_BYTES, &tmdO, &qdo_bar[j], 0, h, (i_begin + j) * TQ, b);
      }
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      // MMA1: S = Q K^T, dP = dO V^T as full 128-key MMAs (every tcgen05.mma re-reads its 128-row A operand from
      // shared memory, and smem operand bandwidth is what bounds these d = 64 shapes: N = 128 halves the A traffic
      // of the two-half variant); both half barriers are signalled by the same completion.
      auto issue_mma1 = [&](int it) {
        const uint32_t aQ = smem_u32(sQ + (it & 1) * TILE_BYTES), adO = smem_u32(sdO + (it & 1) * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_S, desc_k64(aQ, k), desc_k64(aK, k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_bar);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_bf16_ss(tmem_dP, desc_k64(adO, k), desc_k64(aV, k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(dp_bar);
      };
      mbar_wait(kv_bar, 0);
      mbar_wait(&qdo_bar[0], 0);
      tc_fence_after();
      issue_mma1(0);
      for (int it = 0; it < n_it; ++it) {
        const int bsel = it & 1;
        mbar_wait(&soft_bar[0], it & 1);      // P/dS(it) in smem; S, dP and dQ TMEM regions are free again
        mbar_wait(&soft_bar[1], it & 1);
        tc_fence_after();
        if (it + 1 < n_it) {
          mbar_wait(&qdo_bar[bsel ^ 1], ((it + 1) >> 1) & 1);
          tc_fence_after();
          issue_mma1(it + 1);                 // first in the tensor pipe: soft(it+1) can start early ...
        }
        {                                     // ... while MMA2(it) executes underneath it
          const uint32_t aP = smem_u32(sP + bsel * 2 * TILE_BYTES), adS = smem_u32(sdS + bsel * 2 * TILE_BYTES);
          const uint32_t aQ = smem_u32(sQ + bsel * TILE_BYTES), adO = smem_u32(sdO + bsel * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < TQ / 16; ++k)   // dV[kv,d] += P^T dO, contraction over the 128 query rows
            umma_bf16_ss(tmem_dV, desc_mn128(aP, k), desc_mn64(adO, k), idesc_tt, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < TQ / 16; ++k)   // dK[kv,d] += dS^T Q
            umma_bf16_ss(tmem_dK, desc_mn128(adS, k), desc_mn64(aQ, k), idesc_tt, (it > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          if (nck == 4) {                     // dQ[q,d] = dS K, contraction over the (computed) keys
#pragma unroll
            for (int k = 0; k < TK / 16; ++k)
              umma_bf16_ss(tmem_dQ, desc_k128(adS, k), desc_mn64(aK, k), idesc_dq, k > 0 ? 1u : 0u);
          } else {
            for (int k = 0; k < 2 * nck; ++k)
              umma_bf16_ss(tmem_dQ, desc_k128(adS, k), desc_mn64(aK, k), idesc_dq, k > 0 ? 1u : 0u);
          }
          umma_commit(out_bar);
        }
        if (it + 2 < n_it) {                  // refill this Q/dO buffer once MMA2(it) has drained it
          mbar_wait(out_bar, it & 1);
          mbar_arrive_expect_tx(&qdo_bar[bsel], 2 * TILE_BYTES);
          tma_load_4d(sQ + bsel * TILE_BYTES, &tmQ, &qdo_bar[bsel], 0, h, (i_begin + it + 2) * TQ, b);
          tma_load_4d(sdO + bsel * TILE_BYTES, &tmdO, &qdo_bar[bsel], 0, h, (i_begin + it + 2) * TQ, b);
        }
      }
    }
  } else {
    // ===================== compute warps =====================
    const int lg = warp & 3, ch = warp >> 2;
    const int row = lg * 32 + lane;
    const uint32_t t_lane = static_cast<uint32_t>(lg * 32) << 16;
    const uint32_t mw0 = mask_words[2 * ch], mw1 = mask_words[2 * ch + 1];
    const long long bh = (long long)b * p.H + h;
    uint2 dkey = make_uint2(0u, 0u);
    if (DROP) dkey = drop_key(p.drop.state, p.drop.site_a);
    const float dp_scale = DROP ? p.scale * p.drop.scale : p.scale;   // dP_eff = mask * dP / keep

    // dQ tile of iteration it_done -> fp32 dQ buffer (vector atomics).  The thread's 32 columns of its TMEM row are
    // transposed through a swizzled 4 KB staging tile so that every warp instruction adds 4 rows x 128 CONTIGUOUS bytes
    // (a row-per-thread RED touched 32 separate lines per instruction: 2048 LSU wavefronts per iteration and SM, measured
    // 12 of 97 us at 8 x 8 x 800 x 800).  The staging tile is the warp's own 32-row x 64-key block of the P tile of that
    // iteration, which MMA2(it_done) has finished reading and which this warp rewrites next in soft(it_done + 2).
    auto dq_flush = [&](int it_done) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_dQ + t_lane + ch * 32, r);
      tmem_ld_wait();
      if (p.dQ != nullptr) {
        float* stg = reinterpret_cast<float*>(sP + (it_done & 1) * 2 * TILE_BYTES + ch * TILE_BYTES + lg * 4096);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
              make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        __syncwarp();
        const int crow = lane >> 3, xs = lane & 7;
        const int q_first = (i_begin + it_done) * TQ + lg * 32 + crow;       // this lane's rows: q_first + 4 * i
        float* dq = p.dQ + (long long)b * p.dq_bs + (long long)q_first * p.dq_ss + h * HD + ch * 32 + 4 * xs;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(stg + (4 * i + crow) * 32 + ((xs ^ (((i & 1) << 2) | crow)) << 2));
          if (q_first + 4 * i < p.Sq) atomicAdd(reinterpret_cast<float4*>(dq + (long long)(4 * i) * p.dq_ss), v);
        }
      }
      __syncwarp();
    };

    // lse / delta of the NEXT query tile are requested one iteration ahead (their L2 latency was 10 % of the samples)
    float lse_nx = 0.f, delta_nx = 0.f;
    if (i_begin * TQ + row < p.Sq) {
      lse_nx = p.lse[bh * p.Sq + i_begin * TQ + row];
      delta_nx = p.delta[bh * p.Sq + i_begin * TQ + row];
    }
    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_begin + it) * TQ;
      const int qi = q0 + row;
      const bool row_ok = qi < p.Sq;
      const float lse_i = lse_nx, delta_i = delta_nx;
      if (it + 1 < n_it && qi + TQ < p.Sq) {
        lse_nx = p.lse[bh * p.Sq + qi + TQ];
        delta_nx = p.delta[bh * p.Sq + qi + TQ];
      } else {
        lse_nx = 0.f; delta_nx = 0.f;
      }
      uint8_t* tP = sP + (it & 1) * 2 * TILE_BYTES;
      uint8_t* tdS = sdS + (it & 1) * 2 * TILE_BYTES;
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nlse2 = make_float2(-lse_i, -lse_i);
      const float nds = -delta_i * p.scale;
      const float2 dps2 = make_float2(dp_scale, dp_scale), nds2 = make_float2(nds, nds);
      // The softmax half of the backward is ISSUE- and latency-bound (2 warps per sub-partition, ~950 instructions per
      // thread and iteration before this version; 16 compute warps were measured slower — 96 vs 86 us — because the
      // per-warp overhead doubles): packed fp32x2 FMAs, masks applied as bit operations (-inf on the S bits of disallowed
      // keys; dropout keep bytes expanded by PRMT and ANDed), a warp-uniform branch for the common nothing-masked tile.
      // Both 32-key chunks of the thread are in flight together: the dropout hashes run BEFORE the wait for S (while
      // MMA1 executes), the two S loads share one tcgen05.wait, and the second dP load flies under the first dS chunk.
      uint32_t rs0[32], rs1[32], dm0[8], dm1[8];
      uint32_t ok0 = allowed_bits(mw0, CAUSAL, kv0 + (2 * ch) * 32, qi);
      uint32_t ok1 = allowed_bits(mw1, CAUSAL, kv0 + (2 * ch + 1) * 32, qi);
      if (!row_ok) { ok0 = 0u; ok1 = 0u; }
      if (DROP && 2 * ch < nck) {
        const uint32_t prow = (uint32_t)(bh * p.Sq + min(qi, p.Sq - 1)) * p.sk_pad_quarter + ((kv0 + 2 * ch * 32) >> 2);
#pragma unroll
        for (int e = 0; e < 8; ++e) { dm0[e] = drop_quad_bytes(prow + e, dkey, p.thr4); dm1[e] = drop_quad_bytes(prow + 8 + e, dkey, p.thr4); }
      }
      const bool all_ok = __all_sync(0xffffffffu, (ok0 & ok1) == 0xffffffffu);
      // P = exp2(S * scale - lse): fp32 kept in rs for dS, bf16 (dropout-masked: dV += (mask P)^T dO, 1/keep at the store) -> smem
      auto p_chunk = [&](uint32_t (&rs)[32], const uint32_t (&dm)[8], uint32_t ok, int c) {
        uint32_t pk[16];
        if (!all_ok) {
#pragma unroll
          for (int e = 0; e < 32; ++e) rs[e] = ((ok >> e) & 1u) ? rs[e] : 0xff800000u;   // -inf: exp2 -> 0
        }
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float2 a = __ffma2_rn(make_float2(__uint_as_float(rs[e]), __uint_as_float(rs[e + 1])), sc2, nlse2);
          const float p0 = fast_exp2(a.x), p1 = fast_exp2(a.y);
          rs[e] = __float_as_uint(p0); rs[e + 1] = __float_as_uint(p1);
          pk[e >> 1] = pack_bf16(p0, p1);
        }
        if (DROP) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { pk[2 * e] &= keep_lo_pair(dm[e]); pk[2 * e + 1] &= keep_hi_pair(dm[e]); }
        }
        store_chunk_sw128(tP, row, c, pk);
      };
      // dS = P * (mask * dP / keep - delta) * scale -> smem
      auto ds_chunk = [&](uint32_t (&rp)[32], const uint32_t (&rs)[32], const uint32_t (&dm)[8], int c) {
        uint32_t pk[16];
        if (DROP) {
#pragma unroll
          for (int e = 0; e < 32; ++e) rp[e] &= __byte_perm(dm[e >> 2], 0u, 0x1111u * (e & 3));
        }
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(rp[e]), __uint_as_float(rp[e + 1])), dps2, nds2);
          const float2 d = __fmul2_rn(make_float2(__uint_as_float(rs[e]), __uint_as_float(rs[e + 1])), t);
          pk[e >> 1] = pack_bf16(d.x, d.y);
        }
        store_chunk_sw128(tdS, row, c, pk);
      };
      const bool do0 = 2 * ch < nck, do1 = 2 * ch + 1 < nck;      // CTA-uniform: chunks beyond the last real key are skipped
      mbar_wait(s_bar, it & 1);
      tc_fence_after();
      if (do0) {
        tmem_ld_32x32(tmem_S + t_lane + (2 * ch) * 32, rs0);
        if (do1) tmem_ld_32x32(tmem_S + t_lane + (2 * ch + 1) * 32, rs1);
        tmem_ld_wait();
        p_chunk(rs0, dm0, ok0, 2 * ch);
        if (do1) p_chunk(rs1, dm1, ok1, 2 * ch + 1);
      }
      mbar_wait(dp_bar, it & 1);
      tc_fence_after();
      if (do0) {
        uint32_t rp0[32], rp1[32];
        tmem_ld_32x32(tmem_dP + t_lane + (2 * ch) * 32, rp0);
        tmem_ld_wait();
        if (do1) tmem_ld_32x32(tmem_dP + t_lane + (2 * ch + 1) * 32, rp1);
        ds_chunk(rp0, rs0, dm0, 2 * ch);
        tmem_ld_wait();
        if (do1) ds_chunk(rp1, rs1, dm1, 2 * ch + 1);
      }
      if (it > 0) {
        mbar_wait(out_bar, (it - 1) & 1);     // MMA2(it-1) done: its dQ tile is complete
        tc_fence_after();
        dq_flush(it - 1);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&soft_bar[ch]);
    }
    mbar_wait(out_bar, (n_it - 1) & 1);
    tc_fence_after();
    dq_flush(n_it - 1);

    // dK / dV: this thread owns key row kv0 + row, 32 of the 64 columns
    const int kv = kv0 + row;
    const bool ok = kv < p.Sk;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      bf16* base = which == 0 ? p.dV : p.dK;
      const long long ss = which == 0 ? p.dv_ss : p.dk_ss, bs = which == 0 ? p.dv_bs : p.dk_bs;
      const uint32_t tm = which == 0 ? tmem_dV : tmem_dK;
      uint32_t r[32];
      tmem_ld_32x32(tm + t_lane + ch * 32, r);
      tmem_ld_wait();
      if (DROP && which == 0) {
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * p.drop.scale);
      }
      if (ok) {
        bf16* orow = base + (long long)b * bs + (long long)kv * ss + h * HD + ch * 32;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint4 q;
          q.x = pack_bf16(__uint_as_float(r[8 * e]), __uint_as_float(r[8 * e + 1]));
          q.y = pack_bf16(__uint_as_float(r[8 * e + 2]), __uint_as_float(r[8 * e + 3]));
          q.z = pack_bf16(__uint_as_float(r[8 * e + 4]), __uint_as_float(r[8 * e + 5]));
          q.w = pack_bf16(__uint_as_float(r[8 * e + 6]), __uint_as_float(r[8 * e + 7]));
          *reinterpret_cast<uint4*>(orow + 8 * e) = q;
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int make_head_map(CUtensorMap* m, const void* ptr, int H, int S, int B, long long ss, long long bs) {
  return kr_make_tmap_bf16_heads(m, ptr, H, S, B, HD, ss, bs, 128);
}

int set_drop(AttnParams& p, const kr_drop_spec* drop, const char* who) {
  p.drop = kr_drop_to_device(drop);
  const long long sk_pad = (long long)((p.Sk + TK - 1) / TK) * TK;
  p.sk_pad_quarter = (uint32_t)(sk_pad / 4);
  if (p.drop.state != nullptr) {
    if (p.drop.thr_b != 0 || p.drop.row_scale != nullptr) { kr_set_error("attention dropout takes a single-mask spec"); return KR_ERR_ARG; }
    if (p.drop.thr_a & 0xffu) { kr_set_error("attention dropout uses the byte-lane mode: thr must be a multiple of 256"); return KR_ERR_ARG; }
    if ((long long)p.B * p.H * p.Sq * (sk_pad / 4) >= (1LL << 32)) { kr_set_error(who); return KR_ERR_UNSUPPORTED; }
    const uint32_t t8 = p.drop.thr_a >> 8;
    p.thr4 = t8 * 0x01010101u;
    if (p.drop.thr_a == 0) p.drop.state = nullptr;     // p = 0: plain kernels
  }
  return KR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int kr_attn_fwd(const void* q, long long q_ss, long long q_bs, const void* k, long long k_ss,
                           long long k_bs, const void* v, long long v_ss, long long v_bs, void* o,
                           long long o_ss, long long o_bs, float* lse, const unsigned char* key_mask,
                           int B, int H, int Sq, int Sk, int causal, float scale, const kr_drop_spec* drop,
                           void* stream) {
  if (B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0) { kr_set_error("kr_attn_fwd: empty problem"); return KR_ERR_ARG; }
  if (causal && Sq != Sk) { kr_set_error("kr_attn_fwd: causal needs Sq == Sk"); return KR_ERR_ARG; }
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_head_map(&tq, q, H, Sq, B, q_ss, q_bs)) != KR_OK) return rc;
  if ((rc = make_head_map(&tk, k, H, Sk, B, k_ss, k_bs)) != KR_OK) return rc;
  if ((rc = make_head_map(&tv, v, H, Sk, B, v_ss, v_bs)) != KR_OK) return rc;
  AttnParams p{};
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.key_mask = key_mask; p.O = reinterpret_cast<bf16*>(o); p.o_ss = o_ss; p.o_bs = o_bs; p.lse = lse;
  if ((rc = set_drop(p, drop, "kr_attn_fwd")) != KR_OK) return rc;
  if ((Sk + TK - 1) / TK > FWD_MAX_KTILES) { kr_set_error("kr_attn_fwd: Sk > 8192 keys is not supported"); return KR_ERR_UNSUPPORTED; }
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    cudaFuncSetAttribute(attn_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    cudaFuncSetAttribute(attn_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    cudaFuncSetAttribute(attn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
    attr = true;
  }
  const int n_qtiles = (Sq + TQ - 1) / TQ;
  dim3 grid(((n_qtiles + 1) / 2) * H * B, 1, 1);   // one CTA per PAIR of query tiles, pair outermost (LPT order)
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool dr = p.drop.state != nullptr;
  if (causal && dr)  kr::launch(attn_fwd_kernel<true, true>, grid, FWD_THREADS, FWD_SMEM, st, tq, tk, tv, p);
  else if (causal)   kr::launch(attn_fwd_kernel<true, false>, grid, FWD_THREADS, FWD_SMEM, st, tq, tk, tv, p);
  else if (dr)       kr::launch(attn_fwd_kernel<false, true>, grid, FWD_THREADS, FWD_SMEM, st, tq, tk, tv, p);
  else               kr::launch(attn_fwd_kernel<false, false>, grid, FWD_THREADS, FWD_SMEM, st, tq, tk, tv, p);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_attn_bwd(const void* q, long long q_ss, long long q_bs, const void* k, long long k_ss,
                           long long k_bs, const void* v, long long v_ss, long long v_bs, const void* o,
                           long long o_ss, long long o_bs, const void* d_o, long long do_ss,
                           long long do_bs, const float* lse, float* delta, float* dq, long long dq_ss,
                           long long dq_bs, void* dk, long long dk_ss, long long dk_bs, void* dv,
                           long long dv_ss, long long dv_bs, const unsigned char* key_mask, int B, int H,
                           int Sq, int Sk, int causal, float scale, const kr_drop_spec* drop, void* stream) {
  if (B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0) { kr_set_error("kr_attn_bwd: empty problem"); return KR_ERR_ARG; }
  if (causal && Sq != Sk) { kr_set_error("kr_attn_bwd: causal needs Sq == Sk"); return KR_ERR_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    const int rows = B * Sq, wpb = 8;
    kr::launch(attn_bwd_prep_kernel, (rows + wpb - 1) / wpb, wpb * 32, 0, st, 
        reinterpret_cast<const bf16*>(o), o_ss, o_bs, reinterpret_cast<const bf16*>(d_o), do_ss, do_bs,
        delta, B, H, Sq, dq, dq_ss, dq_bs);
    KR_CHECK_LAUNCH();
  }
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_head_map(&tq, q, H, Sq, B, q_ss, q_bs)) != KR_OK) return rc;
  if ((rc = make_head_map(&tk, k, H, Sk, B, k_ss, k_bs)) != KR_OK) return rc;
  if ((rc = make_head_map(&tv, v, H, Sk, B, v_ss, v_bs)) != KR_OK) return rc;
  if ((rc = make_head_map(&tdo, d_o, H, Sq, B, do_ss, do_bs)) != KR_OK) return rc;
  AttnParams p{};
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.key_mask = key_mask; p.lse = const_cast<float*>(lse); p.delta = delta;
  p.dQ = dq; p.dq_ss = dq_ss; p.dq_bs = dq_bs;
  p.dK = reinterpret_cast<bf16*>(dk); p.dk_ss = dk_ss; p.dk_bs = dk_bs;
  p.dV = reinterpret_cast<bf16*>(dv); p.dv_ss = dv_ss; p.dv_bs = dv_bs;
  if ((rc = set_drop(p, drop, "kr_attn_bwd")) != KR_OK) return rc;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
    cudaFuncSetAttribute(attn_bwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
    cudaFuncSetAttribute(attn_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
    cudaFuncSetAttribute(attn_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
    attr = true;
  }
  dim3 grid(((Sk + TK - 1) / TK) * H * B, 1, 1);
  const bool dr = p.drop.state != nullptr;
  if (causal && dr)  kr::launch(attn_bwd_kernel<true, true>, grid, BWD_THREADS, BWD_SMEM, st, tq, tk, tv, tdo, p);
  else if (causal)   kr::launch(attn_bwd_kernel<true, false>, grid, BWD_THREADS, BWD_SMEM, st, tq, tk, tv, tdo, p);
  else if (dr)       kr::launch(attn_bwd_kernel<false, true>, grid, BWD_THREADS, BWD_SMEM, st, tq, tk, tv, tdo, p);
  else               kr::launch(attn_bwd_kernel<false, false>, grid, BWD_THREADS, BWD_SMEM, st, tq, tk, tv, tdo, p);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
