// TEST INFRASTRUCTURE — a minimal SIMT emulator for the dual-compiled kernel bodies (csrc/kr_*_core.cuh with
// -DKR_HOST_EMU -DKR_HOST_EMU_SIMT): one host thread per CUDA thread of a block, __syncthreads() as a pthread barrier over
// the block, warp shuffles / reductions as exchanges through a per-warp buffer guarded by a per-warp barrier.
//
// What it adds over the one-"thread" emulation: the kernels run with their REAL block and warp geometry, so the lane /
// warp index arithmetic, the warp-level reductions and the shared-memory hand-offs between phases execute as written.
// A shuffle reached by only part of a warp blocks forever (the device's undefined behaviour becomes a test timeout), and
// because pthread barriers are understood by ThreadSanitizer, building the harness with -fsanitize=thread turns a missing
// __syncthreads() into a reported data race — the host-side counterpart of `compute-sanitizer --tool racecheck`.
// Blocks of a grid run one after the other on a persistent pool of threads.
#pragma once
#include <pthread.h>
#include <stdint.h>

#include <functional>
#include <thread>
#include <vector>

namespace emu {

struct Block {
  int nthreads = 0;
  pthread_barrier_t bar;
  std::vector<pthread_barrier_t> warp_bar;
  std::vector<float> wf;       // [warps][32]
  std::vector<int> wi;
  std::vector<float> wv;       // [warps][32][kVec]: vector reductions
};

constexpr int kVec = 64;
inline thread_local int t_tid = 0;
inline thread_local Block* t_blk = nullptr;

inline int tid() { return t_tid; }
inline int nthreads() { return t_blk->nthreads; }
inline void syncthreads() { pthread_barrier_wait(&t_blk->bar); }

template <class T, class Op>
inline T warp_reduce(T v, std::vector<T>& buf, Op op) {
  Block* b = t_blk;
  const int w = t_tid >> 5, l = t_tid & 31;
  buf[w * 32 + l] = v;
  pthread_barrier_wait(&b->warp_bar[w]);          // every lane of the warp must arrive: a divergent shuffle hangs here
  T r = buf[w * 32];
  for (int i = 1; i < 32; ++i) r = op(r, buf[w * 32 + i]);
  pthread_barrier_wait(&b->warp_bar[w]);          // the buffer may be reused
  return r;
}
inline float warp_sum(float v) { return warp_reduce(v, t_blk->wf, [](float a, float b) { return a + b; }); }
inline float warp_max(float v) { return warp_reduce(v, t_blk->wf, [](float a, float b) { return a > b ? a : b; }); }
inline float warp_min(float v) { return warp_reduce(v, t_blk->wf, [](float a, float b) { return a < b ? a : b; }); }
inline int warp_min_int(int v) { return warp_reduce(v, t_blk->wi, [](int a, int b) { return a < b ? a : b; }); }
inline int warp_sum_int(int v) { return warp_reduce(v, t_blk->wi, [](int a, int b) { return a + b; }); }

// element-wise sum of n <= kVec values per lane across the warp (one exchange instead of n)
inline void warp_sum_vec(float* v, int n) {
  Block* b = t_blk;
  const int w = t_tid >> 5, l = t_tid & 31;
  float* buf = b->wv.data() + (size_t)w * 32 * kVec;
  for (int i = 0; i < n; ++i) buf[l * kVec + i] = v[i];
  pthread_barrier_wait(&b->warp_bar[w]);
  for (int i = 0; i < n; ++i) {
    float s = 0.f;
    for (int k = 0; k < 32; ++k) s += buf[k * kVec + i];
    v[i] = s;
  }
  pthread_barrier_wait(&b->warp_bar[w]);
}

// the device's kr::block_sum pattern (kr_common.cuh): warp reduce, one value per warp through shared memory, reduce again
template <class T, class WR>
inline T block_reduce(T v, T* red, T fill, WR wr) {
  v = wr(v);
  const int w = t_tid >> 5, l = t_tid & 31, nw = (t_blk->nthreads + 31) >> 5;
  syncthreads();
  if (l == 0) red[w] = v;
  syncthreads();
  T r = l < nw ? red[l] : fill;
  return wr(r);
}
inline float block_sum(float v, float* red) { return block_reduce<float>(v, red, 0.f, warp_sum); }
inline float block_max(float v, float* red) { return block_reduce<float>(v, red, -3.402823466e38f, warp_max); }
inline int block_sum_int(int v, int* red) { return block_reduce<int>(v, red, 0, warp_sum_int); }

// Persistent pool: run(fn) executes fn() on `nthreads` threads with tid 0..nthreads-1 and returns when all are done.
class Pool {
 public:
  explicit Pool(int nthreads) : n_(nthreads) {
    blk_.nthreads = nthreads;
    pthread_barrier_init(&blk_.bar, nullptr, nthreads);
    const int nw = (nthreads + 31) / 32;
    blk_.warp_bar.resize(nw);
    for (int w = 0; w < nw; ++w) {
      const int lanes = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
      pthread_barrier_init(&blk_.warp_bar[w], nullptr, lanes);
    }
    blk_.wf.assign(nw * 32, 0.f);
    blk_.wi.assign(nw * 32, 0);
    blk_.wv.assign((size_t)nw * 32 * kVec, 0.f);
    pthread_barrier_init(&gate_, nullptr, nthreads + 1);
    for (int t = 0; t < nthreads; ++t) threads_.emplace_back([this, t] { loop(t); });
  }
  ~Pool() {
    stop_ = true;
    pthread_barrier_wait(&gate_);
    for (auto& th : threads_) th.join();
  }
  void run(const std::function<void()>& fn) {
    fn_ = &fn;
    pthread_barrier_wait(&gate_);   // release the workers
    pthread_barrier_wait(&gate_);   // wait for them
  }

 private:
  void loop(int t) {
    t_tid = t;
    t_blk = &blk_;
    for (;;) {
      pthread_barrier_wait(&gate_);
      if (stop_) return;
      (*fn_)();
      pthread_barrier_wait(&gate_);
    }
  }
  int n_;
  Block blk_;
  pthread_barrier_t gate_;
  std::vector<std::thread> threads_;
  const std::function<void()>* fn_ = nullptr;
  bool stop_ = false;
};

}  // namespace emu
