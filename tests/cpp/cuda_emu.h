// TEST INFRASTRUCTURE (CPU tier): the few CUDA built-ins the kernels of this repository use, for a host build in which one
// OS thread plays one CUDA thread and the blocks of a grid run one after the other:
//   threadIdx / blockIdx   thread-local;  blockDim / gridDim  globals set by run_grid
//   __shared__             function-local static (only one block is alive at a time)
//   __syncthreads()        pthread barrier over the block's threads
//   __shfl_xor_sync / __shfl_up_sync   exchange through a per-warp buffer between two warp barriers (full masks only)
//   __fmul_rn ...          plain float operations (build with -ffp-contract=off), __float2int_rn = nearbyintf (half to even)
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
using std::max;
using std::min;

template <class F> static void run_grid(unsigned gx, unsigned gy, unsigned gz, unsigned threads, size_t dyn_smem_bytes, F &&kernel) {
  gridDim = EmuIdx{gx, gy, gz};
  blockDim = EmuIdx{threads, 1, 1};
  const unsigned warps = (threads + 31) / 32;
  std::vector<unsigned char> dyn(dyn_smem_bytes + 64);
  g_dyn_smem = dyn.data();
  for (unsigned bz = 0; bz < gz; ++bz)
    for (unsigned by = 0; by < gy; ++by)
      for (unsigned bx = 0; bx < gx; ++bx) {
        pthread_barrier_init(&g_block_barrier, nullptr, threads);
        for (unsigned w = 0; w < warps; ++w) pthread_barrier_init(&g_warp_barrier[w], nullptr, std::min(32u, threads - 32 * w));
        std::vector<std::thread> pool;
        pool.reserve(threads);
        for (unsigned t = 0; t < threads; ++t)
          pool.emplace_back([&, t] {
            threadIdx = EmuIdx{t, 0, 0};
            blockIdx = EmuIdx{bx, by, bz};
            kernel();
          });
        for (auto &th : pool) th.join();
        pthread_barrier_destroy(&g_block_barrier);
        for (unsigned w = 0; w < warps; ++w) pthread_barrier_destroy(&g_warp_barrier[w]);
      }
  g_dyn_smem = nullptr;
}
