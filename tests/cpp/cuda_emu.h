// TEST INFRASTRUCTURE (CPU tier): the few CUDA built-ins the kernels of this repository use, for a host build in which one
// OS thread plays one CUDA thread and the blocks of a grid run one after the other:
//   threadIdx / blockIdx   thread-local;  blockDim / gridDim  globals set by run_grid
//   __shared__             function-local static (only one block is alive at a time)
//   __syncthreads()        pthread barrier over the block's threads
//   __shfl_xor_sync / __shfl_up_sync   exchange through a per-warp buffer between two warp barriers (full masks only)
//   __shfl_sync / __shfl_down_sync / __ballot_sync / __syncwarp   the same way;  atomic*  GCC atomics on the same address
//   __fmul_rn ...          plain float operations (build with -ffp-contract=off), __float2int_rn = nearbyintf (half to even)
//   __popc, __ffs, __vsadu4, __ldcg, __threadfence, clock64 (0)
// It reproduces data flow and arithmetic, not timing or scheduling.  Include BEFORE the kernel header.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>

struct EmuIdx { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuIdx threadIdx, blockIdx;
static EmuIdx blockDim, gridDim;
static pthread_barrier_t g_block_barrier;
static pthread_barrier_t g_warp_barrier[32];
static unsigned long long g_xchg[32][32];
static unsigned char *g_dyn_smem = nullptr;                      // dynamic shared memory of the running block

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __align__
#define __align__(n) __attribute__((aligned(n)))

static inline void __syncthreads() { pthread_barrier_wait(&g_block_barrier); }
template <class T> static inline T emu_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  g_xchg[w][l] = bits;
  pthread_barrier_wait(&g_warp_barrier[w]);
  const unsigned long long got = g_xchg[w][src_lane];
  pthread_barrier_wait(&g_warp_barrier[w]);
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int d) { return emu_exchange(v, (int)((threadIdx.x & 31) ^ d)); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int o) {
  const int l = threadIdx.x & 31;
  return emu_exchange(v, l >= o ? l - o : l);                    // lanes below the offset get their own value back
}
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __float2int_rn(float x) { return (int)nearbyintf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdividef(float a, float b) { return a / b; }        // approximate on the GPU: do not expect bit parity through it
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __vsadu4(unsigned a, unsigned b) {                   // sum of absolute differences of the four unsigned bytes
  unsigned s = 0;
  for (int i = 0; i < 4; ++i) { const int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF; s += (unsigned)(x > y ? x - y : y - x); }
  return s;
}
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  const int w = threadIdx.x >> 5;
  pthread_barrier_wait(&g_warp_barrier[w]);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int o) {
  const int l = threadIdx.x & 31;
  return emu_exchange(v, l + o < 32 ? l + o : l);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  g_xchg[w][l] = pred ? 1ull : 0ull;
  pthread_barrier_wait(&g_warp_barrier[w]);
  unsigned m = 0;
  const unsigned lanes = std::min(32u, blockDim.x - 32u * (unsigned)w);
  for (unsigned i = 0; i < lanes; ++i) m |= (unsigned)g_xchg[w][i] << i;
  pthread_barrier_wait(&g_warp_barrier[w]);
  return m;
}
static inline unsigned emu_lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }     // %lanemask_lt (kernels read it through inline PTX)
// atomics on shared or global memory: the GCC builtins on the same address
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double *p, double v) {
  double old = *p, next;
  do { next = old + v; } while (!__atomic_compare_exchange(p, &old, &next, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return old;
}
static inline float atomicAdd(float *p, float v) {
  float old = *p, next;
  do { next = old + v; } while (!__atomic_compare_exchange(p, &old, &next, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return old;
}
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicCAS(T *p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
template <class T> static inline T atomicMin(T *p, T v) { T old = *p; while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return old; }
template <class T> static inline T atomicMax(T *p, T v) { T old = *p; while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return old; }
static inline long long clock64() { return 0; }
using std::max;
using std::min;

template <class F> static void run_grid(unsigned gx, unsigned gy, unsigned gz, unsigned threads, size_t dyn_smem_bytes, F &&kernel) {
  gridDim = EmuIdx{gx, gy, gz};
  blockDim = EmuIdx{threads, 1, 1};
  const unsigned warps = (threads + 31) / 32;
  std::vector<unsigned char> dyn(dyn_smem_bytes + 64);
  g_dyn_smem = dyn.data();
  // one set of OS threads for the whole grid: every thread plays the same threadIdx in block after block; a barrier at the end of
  // each block keeps the blocks strictly one after the other (the __shared__ statics and the block barrier are reused)
  pthread_barrier_t end_of_block;
  pthread_barrier_init(&end_of_block, nullptr, threads);
  pthread_barrier_init(&g_block_barrier, nullptr, threads);
  for (unsigned w = 0; w < warps; ++w) pthread_barrier_init(&g_warp_barrier[w], nullptr, std::min(32u, threads - 32 * w));
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (unsigned t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      threadIdx = EmuIdx{t, 0, 0};
      for (unsigned bz = 0; bz < gz; ++bz)
        for (unsigned by = 0; by < gy; ++by)
          for (unsigned bx = 0; bx < gx; ++bx) {
            blockIdx = EmuIdx{bx, by, bz};
            kernel();
            pthread_barrier_wait(&end_of_block);
          }
    });
  for (auto &th : pool) th.join();
  pthread_barrier_destroy(&end_of_block);
  pthread_barrier_destroy(&g_block_barrier);
  for (unsigned w = 0; w < warps; ++w) pthread_barrier_destroy(&g_warp_barrier[w]);
  g_dyn_smem = nullptr;
}

// kernel launch as written in the product (`k<<<grid, block, smem, stream>>>(args)`, rewritten by tests/emu_build.py):
// grid and block may be ints or dim3; blocks are one-dimensional in this code base.
static inline EmuIdx emu_dim(const dim3 &d) { return EmuIdx{d.x, d.y, d.z}; }
static inline EmuIdx emu_dim(int v) { return EmuIdx{(unsigned)v, 1, 1}; }
static inline EmuIdx emu_dim(unsigned v) { return EmuIdx{v, 1, 1}; }
template <class G, class B, class F> static void emu_launch(G grid, B block, size_t dyn_smem_bytes, F &&kernel) {
  const EmuIdx g = emu_dim(grid), b = emu_dim(block);
  run_grid(g.x, g.y, g.z, b.x, dyn_smem_bytes, kernel);
}

// the C++ convenience overloads of the runtime that cuda_runtime.h only declares for nvcc
template <class R, class... A> static inline cudaError_t cudaFuncSetAttribute(R (*)(A...), cudaFuncAttribute, int) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbolAsync(const T &symbol, const void *src, size_t n, size_t off, cudaMemcpyKind, cudaStream_t) {
  memcpy((char *)&symbol + off, src, n);
  return cudaSuccess;
}
