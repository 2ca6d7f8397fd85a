// Test infrastructure: a tiny CUDA-on-CPU execution model, so that block-level kernels
// (shared memory, __syncthreads, shared/global atomics, warp shuffles) can be EXECUTED
// in the CPU test suite of a container that has no GPU.  Not part of the product.
//
// Model: one OS thread per CUDA thread of a block; the blocks of a grid run one after
// the other on the same team of threads.  `__shared__` becomes a function-local static
// (one copy, i.e. per block since blocks are sequential), `__syncthreads()` a team
// barrier, atomics are real atomics, warp shuffles exchange through a per-warp buffer.
// What this checks: indexing, tails, barriers placed consistently, accumulation logic.
// What it cannot check: memory-model races between warps, alignment traps, performance.
//
// Usage (see tests/test_cuda_emu.py): compile a translation unit with
//   g++ -std=c++17 -O1 -pthread -x c++ -include tests/cuda_emu/cuda_emu.h kernel_source ...
// and launch with emu::launch(grid, block, [&]{ kernel(args...); }).
#pragma once
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <unordered_map>
#include <memory>
#include <mutex>
#include <pthread.h>
#include <thread>
#include <vector>

#define PSB_CUDA_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __grid_constant__
#define __noinline__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_emu { unsigned x, y, z; };
struct double2 { double x, y; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline double2 make_double2(double a, double b) { return {a, b}; }
static inline float2 make_float2(float a, float b) { return {a, b}; }

namespace emu {

class Barrier {      // reusable barrier (futex-based pthread barrier: far cheaper than a
 public:            // mutex + condition-variable broadcast for teams of 256-512 threads)
  Barrier() = default;
  Barrier(const Barrier&) = delete;
  ~Barrier() { if (live_) pthread_barrier_destroy(&b_); }
  void reset(int n) {
    if (live_) pthread_barrier_destroy(&b_);
    pthread_barrier_init(&b_, nullptr, (unsigned)n);
    live_ = true;
  }
  void wait() { pthread_barrier_wait(&b_); }
 private:
  pthread_barrier_t b_;
  bool live_ = false;
};

struct Team {
  Barrier block;
  std::mutex named_mutex;
  std::map<int, std::unique_ptr<Barrier>> named;
  std::unique_ptr<Barrier[]> warps;
  std::vector<std::vector<unsigned long long>> warp_buf;   // 32 slots of 8 bytes per warp
};

inline Team*& team() { static Team* t = nullptr; return t; }
inline thread_local int t_linear = 0;

}  // namespace emu

// the CUDA built-in index variables: real objects (not macros), so that members of the same
// name elsewhere (cudaLaunchConfig_t::gridDim) are left alone
inline thread_local uint3_emu threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace emu {

// Worker threads are kept between launches (creating 256-512 threads per launch costs more
// than most of the small kernels of the tests); the pool is leaked on purpose.
struct Pool {
  std::mutex m;
  std::condition_variable cv;
  std::vector<std::thread> ths;
  unsigned long gen = 0;
  int nt = 0;
  std::function<void(int)> job;
  Barrier done;

  void worker(int t) {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return gen != seen; });
        seen = gen;
        if (t >= nt) continue;
      }
      job(t);
      done.wait();
    }
  }
  void run(int n, std::function<void(int)> f) {
    {
      std::unique_lock<std::mutex> lk(m);
      while ((int)ths.size() < n) {
        const int t = (int)ths.size();
        ths.emplace_back([this, t] { worker(t); });
        ths.back().detach();
      }
      job = std::move(f);
      nt = n;
      done.reset(n + 1);
      ++gen;
    }
    cv.notify_all();
    done.wait();
  }
};
inline Pool& pool() { static Pool* p = new Pool; return *p; }

// Block order: CUDA promises none.  Forward by default; `reverse` runs the highest linear
// block index first, which a kernel whose block 0 WAITS for the others needs on a model that
// runs blocks one after the other (the watcher CTA of csrc/stage_tail.cuh).
inline bool& reverse_blocks() { static bool r = false; return r; }

template <class F>
void launch(dim3 grid, dim3 block, F&& body) {
  const int nt = int(block.x * block.y * block.z);
  Team tm;
  tm.block.reset(nt);
  const int nw = (nt + 31) / 32;
  tm.warps.reset(new Barrier[nw]);
  tm.warp_buf.assign(nw, std::vector<unsigned long long>(32, 0));
  for (int w = 0; w < nw; ++w) tm.warps[w].reset(std::min(32, nt - 32 * w));
  team() = &tm;
  blockDim = block;
  gridDim = grid;
  pool().run(nt, [&](int t) {
    t_linear = t;
    threadIdx = {unsigned(t % block.x), unsigned((t / block.x) % block.y),
                 unsigned(t / (block.x * block.y))};
    const unsigned long nb = (unsigned long)grid.x * grid.y * grid.z;
    for (unsigned long n = 0; n < nb; ++n) {
      const unsigned long b = reverse_blocks() ? nb - 1 - n : n;
      blockIdx = {unsigned(b % grid.x), unsigned((b / grid.x) % grid.y),
                  unsigned(b / ((unsigned long)grid.x * grid.y))};
      body();
      tm.block.wait();          // the next block reuses the statics ("shared memory")
    }
  });
  team() = nullptr;
}

// dynamic shared memory of the running block (size fixed by launch_smem)
inline std::vector<double>& dyn_smem_store() { static std::vector<double> v; return v; }
inline double* dynamic_smem() { return dyn_smem_store().data(); }

template <class F>
void launch_smem(dim3 grid, dim3 block, size_t smem_bytes, F&& body) {
  // fill with a non-zero pattern: kernels must initialise what they read
  dyn_smem_store().assign((smem_bytes + 7) / 8, -12345.678);
  launch(grid, block, body);
}

template <class T>
T shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  Team& tm = *team();
  const int w = t_linear / 32, lane = t_linear % 32;
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  tm.warp_buf[w][lane] = bits;
  tm.warps[w].wait();
  unsigned long long got = tm.warp_buf[w][src_lane & 31];
  tm.warps[w].wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}

inline unsigned ballot(bool pred) {
  Team& tm = *team();
  const int w = t_linear / 32, lane = t_linear % 32;
  tm.warp_buf[w][lane] = pred ? 1ull : 0ull;
  tm.warps[w].wait();
  unsigned bits = 0;
  for (int l = 0; l < 32; ++l) bits |= unsigned(tm.warp_buf[w][l] & 1ull) << l;
  tm.warps[w].wait();
  return bits;
}

// whether tensor maps can be encoded (off: the library behaves as without the driver entry
// point and every kernel takes its non-TMA instance)
inline bool& tma_enabled() { static bool e = true; return e; }

// ---- mbarrier + TMA (cp.async.bulk.tensor) model -------------------------------------------
// State is kept beside the 64-bit words (keyed by address); one lock and one condition
// variable for all barriers: the lattices of the tests are tiny.
struct Mbar { int expected = 0, pending = 0; long tx = 0; unsigned phase = 0; };
inline std::mutex& mbar_mutex() { static std::mutex m; return m; }
inline std::condition_variable& mbar_cv() { static std::condition_variable c; return c; }
inline std::unordered_map<const void*, Mbar>& mbars() {
  static std::unordered_map<const void*, Mbar> t; return t;
}
inline void mbar_settle(Mbar& b) {           // a phase completes: all arrivals and all bytes
  if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.expected; }
}
inline void mbar_init(const void* p, unsigned count) {
  std::lock_guard<std::mutex> lk(mbar_mutex());
  Mbar b; b.expected = b.pending = int(count); mbars()[p] = b;
}
inline void mbar_arrive(const void* p) {
  { std::lock_guard<std::mutex> lk(mbar_mutex()); Mbar& b = mbars()[p]; --b.pending; mbar_settle(b); }
  mbar_cv().notify_all();
}
inline void mbar_expect_tx(const void* p, unsigned bytes) {
  { std::lock_guard<std::mutex> lk(mbar_mutex()); Mbar& b = mbars()[p]; b.tx += bytes; --b.pending;
    mbar_settle(b); }
  mbar_cv().notify_all();
}
inline void mbar_complete_tx(const void* p, unsigned bytes) {
  { std::lock_guard<std::mutex> lk(mbar_mutex()); Mbar& b = mbars()[p]; b.tx -= bytes; mbar_settle(b); }
  mbar_cv().notify_all();
}
inline void mbar_wait(const void* p, unsigned parity) {   // until the phase of that parity is over
  std::unique_lock<std::mutex> lk(mbar_mutex());
  mbar_cv().wait(lk, [&] { return mbars()[p].phase != (parity & 1u); });
}

// software tensor map: what cuTensorMapEncodeTiled is given, stored in the 128 opaque bytes
struct TensorMap {
  char* base; int es; unsigned rank; unsigned long long dim[4]; unsigned long long stride[3];
  unsigned box[4];
};
inline int encode_tiled(void* map, int dtype, unsigned rank, void* base,
                        const unsigned long long* gdim, const unsigned long long* gstride,
                        const unsigned* box, const unsigned*, int, int, int, int) {
  if (!tma_enabled()) return 1;           // as if the driver refused: callers fall back
  TensorMap t;
  std::memset(&t, 0, sizeof(t));
  t.base = static_cast<char*>(base);
  t.es = dtype == 8 ? 8 : 4;              // CU_TENSOR_MAP_DATA_TYPE_FLOAT64 = 8 in the stub
  t.rank = rank;
  for (unsigned k = 0; k < 4; ++k) { t.dim[k] = k < rank ? gdim[k] : 1; t.box[k] = k < rank ? box[k] : 1; }
  for (unsigned k = 0; k + 1 < rank; ++k) t.stride[k] = gstride[k];
  static_assert(sizeof(TensorMap) <= 128, "tensor map model");
  std::memcpy(map, &t, sizeof(t));
  return 0;
}
// tile load: box elements innermost coordinate first, zeros outside the tensor, then the
// byte count of the WHOLE box is credited to the mbarrier (as the hardware does)
inline void tma_load_4d(void* dst, unsigned long long map, const void* bar, int c0, int c1, int c2,
                        int c3) {
  TensorMap t;
  std::memcpy(&t, reinterpret_cast<const void*>(map), sizeof(t));
  char* out = static_cast<char*>(dst);
  const long long c[4] = {c0, c1, c2, c3};
  for (unsigned i3 = 0; i3 < t.box[3]; ++i3)
    for (unsigned i2 = 0; i2 < t.box[2]; ++i2)
      for (unsigned i1 = 0; i1 < t.box[1]; ++i1)
        for (unsigned i0 = 0; i0 < t.box[0]; ++i0) {
          const long long g[4] = {c[0] + i0, c[1] + i1, c[2] + i2, c[3] + i3};
          bool in = true;
          for (int k = 0; k < 4; ++k) in = in && g[k] >= 0 && (unsigned long long)g[k] < t.dim[k];
          if (in)
            std::memcpy(out, t.base + g[0] * t.es + g[1] * (long long)t.stride[0] +
                                 g[2] * (long long)t.stride[1] + g[3] * (long long)t.stride[2], t.es);
          else
            std::memset(out, 0, t.es);
          out += t.es;
        }
  mbar_complete_tx(bar, t.box[0] * t.box[1] * t.box[2] * t.box[3] * t.es);
}

// PTX of hardware-only paths without a model here: reaching it is a bug
inline void unsupported_ptx() {
  fprintf(stderr, "cuda_emu: PTX of a hardware-only path was reached\n");
  abort();
}
// bar.sync id, count: the whole block, or a named barrier among `count` threads (the consumer
// warps of the warp-specialised kernels)
inline void named_barrier(int id, int count) {
  Team& tm = *team();
  if (id == 0 && count == int(blockDim.x * blockDim.y * blockDim.z)) { tm.block.wait(); return; }
  Barrier* b;
  {
    std::lock_guard<std::mutex> lk(tm.named_mutex);
    auto& slot = tm.named[id];
    if (!slot) { slot.reset(new Barrier); slot->reset(count); }
    b = slot.get();
  }
  b->wait();
}

template <class T>
T atomic_add(T* p, T v) {
  // CAS loop on the bit pattern: works for float / double / integers
  using U = typename std::conditional<sizeof(T) == 8, unsigned long long, unsigned>::type;
  U* up = reinterpret_cast<U*>(p);
  U old = __atomic_load_n(up, __ATOMIC_RELAXED);
  for (;;) {
    T cur;
    std::memcpy(&cur, &old, sizeof(T));
    T nxt = cur + v;
    U nbits;
    std::memcpy(&nbits, &nxt, sizeof(T));
    if (__atomic_compare_exchange_n(up, &old, nbits, false, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED))
      return cur;
  }
}

}  // namespace emu

static inline void __syncthreads() { emu::team()->block.wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::team()->warps[emu::t_linear / 32].wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }

static inline double atomicAdd(double* p, double v) { return emu::atomic_add(p, v); }
static inline float atomicAdd(float* p, float v) { return emu::atomic_add(p, v); }
static inline int atomicAdd(int* p, int v) { return emu::atomic_add(p, v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return emu::atomic_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return emu::atomic_add(p, v);
}

template <class T> static inline T atomicExch(T* p, T v) {
  return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST);
}
template <class T> static inline T atomicCAS(T* p, T cmp, T v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}

template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu::shfl(v, lane); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
  const int lane = emu::t_linear % 32;
  return emu::shfl(v, lane + int(d) < 32 ? lane + int(d) : lane);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  const int lane = emu::t_linear % 32;
  return emu::shfl(v, lane - int(d) >= 0 ? lane - int(d) : lane);
}
static inline unsigned __ballot_sync(unsigned, bool pred) { return emu::ballot(pred); }
static inline int __ffs(unsigned x) { return __builtin_ffs(int(x)); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
  return emu::shfl(v, (emu::t_linear % 32) ^ m);
}

// streaming (evict-first) loads / stores: plain accesses here
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline T __ldcg(const T* p) { return *p; }

// CUDA's global integer min / max
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }

static inline double __ldg(const double* p) { return *p; }
static inline float __ldg(const float* p) { return *p; }
static inline int __ldg(const int* p) { return *p; }
static inline double2 __ldg(const double2* p) { return *p; }
static inline float2 __ldg(const float2* p) { return *p; }
