// kr_comm.cu — the data-parallel collective of a training step, fused with the first phase of the optimizer:
// ONE kernel all-reduces the flat fp32 gradient buffer over NVLink 5 / NVSwitch peer memory AND produces the
// per-chunk squared gradient norms the step control needs (the reference has no data-parallel path; this
// replaces "NCCL all-reduce, then kr_grad_sqnorm re-reads 198 MB": SURVEY.md 8(e), trainer.py:2355-2362).
//
// The gradient buffer of every rank lives in symmetric memory (same virtual layout on all ranks, peers mapped,
// plus an NVSwitch MULTICAST mapping when the fabric supports it).  4096-element chunk c — the optimizer's own
// chunk table, chunks never straddle tensors — is owned by rank c % world.  The owner
//   * multicast path: `multimem.ld_reduce.add.v4.f32` — the SWITCH sums the chunk over all ranks and returns it
//     once (one pass of link traffic instead of a ring's two), `multimem.st.v4.f32` broadcasts the sum back;
//   * peer path (no multicast): 16-byte loads from every rank in rank order, stores to every rank;
//   * squares and sums what it just reduced (the data is in registers) and broadcasts the chunk's sum.
// Every element is reduced exactly once and then replicated, and the per-tensor norms are later summed from the
// chunk sums in a fixed order, so all replicas see bit-identical gradients and norms (NCCL + atomics gave
// neither guarantee for the norms) and take identical clip / skip decisions without further communication.
// Cross-GPU ordering: a per-block flag barrier (system-scope release / acquire CAS on flags in symmetric
// memory) before the first load — the peer's backward has finished — and after the last store.
#include "kr_common.cuh"

namespace {
using namespace kr;

constexpr int MAX_RANKS = 8;
constexpr int AR_THREADS = 256;

struct ArParams {
  float* mc;                      // multicast address of the gradient buffer, or nullptr
  float* sq_mc;                   // multicast address of the per-chunk squared-sum buffer, or nullptr
  float* grads[MAX_RANKS];        // every rank's gradient buffer (peer mappings)
  float* sq[MAX_RANKS];           // every rank's per-chunk squared-sum buffer
  unsigned* flags[MAX_RANKS];     // every rank's barrier flags [grid][world]
  const long long* chunk_start;
  const int* chunk_len;
  int n_chunks, rank, world;
  // the chunks this launch reduces: two half-open chunk-index ranges (the step reduces the gradient ranges that are final
  // half-way through the backward pass from a side stream underneath the rest of it, and the remainder afterwards);
  // virtual index v = position in the concatenation of the two ranges, owned by rank v % world
  int r0_begin, r0_len, r1_begin, r1_len;
  // the step's clip norm is a GLOBAL decision: slot [n_chunks + r] of rank r's sq buffer carries rank r's adaptive clip
  // (long-utterance stabiliser, trainer.py:2218-2255, computed from ITS batch); every rank takes the minimum over all
  // slots, so the replicas apply the same clip coefficient and stay bit-identical
  const float* clip_local;        // this rank's clip for the step (device scalar), or nullptr
  float* clip_out;                // min over ranks, read by kr_step_control as clip_override
  long long watchdog_cycles;      // spin limit of the cross-GPU barrier
  int* error_flag;                // raised (1 = barrier timeout) before the trap, for the host-side diagnosis
};

__device__ __forceinline__ float4 multimem_ld_reduce_f32x4(const float* addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f32x4(float* addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_f32(float* addr, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 ld_sys_f32x4(const float* addr) {      // peer memory: never through L1
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned cas_release_sys(unsigned* addr, unsigned cmp, unsigned val) {
  unsigned old;
  asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ unsigned cas_acquire_sys(unsigned* addr, unsigned cmp, unsigned val) {
  unsigned old;
  asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// Block b of every rank meets block b of every other rank.  Slot [b][src] of rank dst's flags is raised by
// rank src (0 -> 1) and lowered by rank dst (1 -> 0), so the flags are back at zero afterwards.  A watchdog
// traps instead of hanging the box if a peer never arrives.
__device__ __forceinline__ void cross_gpu_barrier(const ArParams& p) {
  __syncthreads();
  if ((int)threadIdx.x < p.world) {
    const int q = threadIdx.x;
    unsigned* put = p.flags[q] + (size_t)blockIdx.x * p.world + p.rank;
    unsigned* wait = p.flags[p.rank] + (size_t)blockIdx.x * p.world + q;
    const long long t0 = clock64();
    while (cas_release_sys(put, 0u, 1u) != 0u) {
      if (clock64() - t0 > p.watchdog_cycles) {
        if (p.error_flag != nullptr) { *p.error_flag = 1; __threadfence_system(); }
        printf("kr_comm: barrier put watchdog (rank %d block %d): a peer never lowered its flag\n", p.rank, blockIdx.x); __trap();
      }
    }
    while (cas_acquire_sys(wait, 1u, 0u) != 1u) {
      if (clock64() - t0 > p.watchdog_cycles) {
        if (p.error_flag != nullptr) { *p.error_flag = 1; __threadfence_system(); }
        printf("kr_comm: barrier wait watchdog (rank %d block %d): peer %d never arrived\n", p.rank, blockIdx.x, q); __trap();
      }
    }
  }
  __syncthreads();
}

// loads of one chunk: all issued before the first use (NVLink / switch latency: keep the link full)
__device__ __forceinline__ void ar_load_chunk(const ArParams& p, int c, float4 (&v)[4]) {
  const long long off = p.chunk_start[c];
  const int n4 = (p.chunk_len[c] + 3) >> 2;      // tensors are 64-element aligned and zero padded: whole float4s are safe
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * AR_THREADS;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      if (p.mc != nullptr) {
        v[k] = multimem_ld_reduce_f32x4(p.mc + off + 4 * i);
      } else {
        for (int q = 0; q < p.world; ++q) {
          const float4 t = ld_sys_f32x4(p.grads[q] + off + 4 * i);
          v[k].x += t.x; v[k].y += t.y; v[k].z += t.z; v[k].w += t.w;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(AR_THREADS) allreduce_sqnorm_kernel(const ArParams p) {
  __shared__ float red[32];
  if (p.clip_local != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    p.sq[p.rank][p.n_chunks + p.rank] = *p.clip_local;      // published by the release of the barrier below
    __threadfence_system();
  }
  cross_gpu_barrier(p);                          // every rank's gradients are final
  if (p.clip_local != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    float c = *p.clip_local;
    for (int q = 0; q < p.world; ++q) {
      float v;
      asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p.sq[q] + p.n_chunks + q) : "memory");
      c = fminf(c, v);
    }
    *p.clip_out = c;
  }
  const int stride = p.world * (int)gridDim.x;
  const int n_virtual = p.r0_len + p.r1_len;
  auto chunk_of = [&](int vi) { return vi < p.r0_len ? p.r0_begin + vi : p.r1_begin + (vi - p.r0_len); };
  int vi = p.rank + p.world * (int)blockIdx.x;
  float4 v[4], vn[4];
  if (vi < n_virtual) ar_load_chunk(p, chunk_of(vi), vn);
  for (; vi < n_virtual; vi += stride) {
    const int c = chunk_of(vi);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = vn[k];
    if (vi + stride < n_virtual) ar_load_chunk(p, chunk_of(vi + stride), vn);   // next chunk's loads fly under this chunk's stores
    const long long off = p.chunk_start[c];
    const int n4 = (p.chunk_len[c] + 3) >> 2;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = threadIdx.x + k * AR_THREADS;
      if (i < n4) {
        s += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
        if (p.mc != nullptr) {
          multimem_st_f32x4(p.mc + off + 4 * i, v[k]);
        } else {
          for (int q = 0; q < p.world; ++q) *reinterpret_cast<float4*>(p.grads[q] + off + 4 * i) = v[k];
        }
      }
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
      if (p.sq_mc != nullptr) {
        multimem_st_f32(p.sq_mc + c, s);
      } else {
        for (int q = 0; q < p.world; ++q) p.sq[q][c] = s;
      }
    }
  }
  __threadfence_system();
  cross_gpu_barrier(p);                          // every rank's stores have landed everywhere
}

// single-GPU / after-NCCL variant of the same first optimizer phase: per-chunk squared sums, no atomics
__global__ void __launch_bounds__(AR_THREADS) chunk_sqnorm_kernel(const float* __restrict__ g,
                                                                 const long long* __restrict__ chunk_start,
                                                                 const int* __restrict__ chunk_len,
                                                                 float* __restrict__ sq_chunk) {
  kr::pdl_entry();
  __shared__ float red[32];
  const int c = blockIdx.x;
  const float* p = g + chunk_start[c];
  const int n4 = (chunk_len[c] + 3) >> 2;
  float s = 0.f;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) sq_chunk[c] = s;
}

// sq[t] = sum of the chunk sums of tensor t in a fixed order (deterministic, identical on every replica);
// a non-finite sum raises the control block's non-finite flag (word `nonfinite_word`).
__global__ void chunk_to_tensor_kernel(const float* __restrict__ sq_chunk, const int* __restrict__ first_chunk,
                                       float* __restrict__ sq, int n_tensors, int* ctrl_nonfinite) {
  kr::pdl_entry();
  // one warp per tensor: lane l sums chunks l, l+32, ... in order, then a fixed shuffle tree -> the same bits on
  // every replica and every run
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= n_tensors) return;
  float s = 0.f;
  for (int c = first_chunk[t] + lane; c < first_chunk[t + 1]; c += 32) s += sq_chunk[c];
  s = warp_sum(s);
  if (lane == 0) {
    sq[t] = s;
    if (!isfinite(s)) atomicExch(ctrl_nonfinite, 1);
  }
}

}  // namespace

extern "C" int kr_allreduce_sqnorm(void* mc_grads, void* mc_sq, void* const* grads, void* const* sq,
                                   void* const* flags, int rank, int world, const long long* chunk_start,
                                   const int* chunk_len, int n_chunks, const int* chunk_ranges, int grid,
                                   const float* clip_local, float* clip_out, double watchdog_seconds, int* error_flag,
                                   void* stream) {
  if (world < 1 || world > MAX_RANKS || rank < 0 || rank >= world) { kr_set_error("kr_allreduce_sqnorm: world must be 1..8"); return KR_ERR_ARG; }
  if (grid < 1 || n_chunks <= 0) { kr_set_error("kr_allreduce_sqnorm: empty problem"); return KR_ERR_ARG; }
  ArParams p{};
  p.mc = reinterpret_cast<float*>(mc_grads);
  p.sq_mc = reinterpret_cast<float*>(mc_sq);
  for (int q = 0; q < world; ++q) {
    p.grads[q] = reinterpret_cast<float*>(grads[q]);
    p.sq[q] = reinterpret_cast<float*>(sq[q]);
    p.flags[q] = reinterpret_cast<unsigned*>(flags[q]);
  }
  p.chunk_start = chunk_start; p.chunk_len = chunk_len; p.n_chunks = n_chunks; p.rank = rank; p.world = world;
  // chunk_ranges (HOST array {begin0, end0, begin1, end1}, may be NULL = all chunks)
  p.r0_begin = 0; p.r0_len = n_chunks; p.r1_begin = 0; p.r1_len = 0;
  if (chunk_ranges != nullptr) {
    p.r0_begin = chunk_ranges[0]; p.r0_len = chunk_ranges[1] - chunk_ranges[0];
    p.r1_begin = chunk_ranges[2]; p.r1_len = chunk_ranges[3] - chunk_ranges[2];
    if (p.r0_begin < 0 || p.r0_len < 0 || p.r1_begin < 0 || p.r1_len < 0 || p.r0_begin + p.r0_len > n_chunks ||
        p.r1_begin + p.r1_len > n_chunks || p.r0_len + p.r1_len == 0) {
      kr_set_error("kr_allreduce_sqnorm: bad chunk ranges"); return KR_ERR_ARG;
    }
  }
  if ((clip_local == nullptr) != (clip_out == nullptr)) { kr_set_error("kr_allreduce_sqnorm: clip_local and clip_out go together"); return KR_ERR_ARG; }
  p.clip_local = clip_local; p.clip_out = clip_out; p.error_flag = error_flag;
  // seconds -> SM cycles at a nominal 2 GHz (the limit only has to be generous: rank skew from one-sided host work)
  p.watchdog_cycles = (long long)((watchdog_seconds > 0.0 ? watchdog_seconds : 300.0) * 2.0e9);
  // plain launch: every block spins on its peers, so the whole grid must not depend on a later grid
  allreduce_sqnorm_kernel<<<grid, AR_THREADS, 0, (cudaStream_t)stream>>>(p);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_chunk_sqnorm(const float* grads, const long long* chunk_start, const int* chunk_len, int n_chunks,
                               float* sq_chunk, void* stream) {
  if (n_chunks <= 0) return KR_OK;
  kr::launch(chunk_sqnorm_kernel, n_chunks, AR_THREADS, 0, (cudaStream_t)stream, grads, chunk_start, chunk_len, sq_chunk);
  KR_CHECK_LAUNCH();
  return KR_OK;
}

extern "C" int kr_chunk_to_tensor_sq(const float* sq_chunk, const int* first_chunk, float* sq, int n_tensors,
                                     void* ctrl, void* stream) {
  if (n_tensors <= 0) return KR_OK;
  int* nonfinite = reinterpret_cast<int*>(ctrl) + 10;      // Ctrl::nonfinite (kr_optim.cu)
  kr::launch(chunk_to_tensor_kernel, (n_tensors + 7) / 8, 256, 0, (cudaStream_t)stream, sq_chunk, first_chunk, sq,
             n_tensors, nonfinite);
  KR_CHECK_LAUNCH();
  return KR_OK;
}
