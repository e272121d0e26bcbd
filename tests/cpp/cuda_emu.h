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
#include <stdio.h>
#include <stdlib.h>
#include <thread>
#include <vector>

#include <mutex>

struct EmuIdx { unsigned x = 0, y = 0, z = 0; };
struct EmuCluster;
struct EmuBlock {                                                 // what the threads of one block share
  pthread_barrier_t bar;
  pthread_barrier_t wbar[32];
  unsigned long long xchg[32][32];
  unsigned char *dyn = nullptr;                                   // dynamic shared memory
  size_t dyn_bytes = 0;
  unsigned char *stat = nullptr;                                  // per-block storage of rewritten __shared__ declarations (EMU_SHARED)
  unsigned rank = 0;                                              // rank inside its cluster
  EmuCluster *cluster = nullptr;
};
struct EmuCluster {
  pthread_barrier_t bar;
  std::vector<EmuBlock *> blocks;
};
constexpr size_t EMU_STATIC_ARENA = 256 * 1024;
static thread_local EmuIdx threadIdx, blockIdx;
static thread_local EmuBlock *tl_blk = nullptr;
static thread_local EmuIdx blockDim, gridDim;                  // thread-local: launches may come from several host threads at once
#define g_dyn_smem (tl_blk->dyn)
#define g_block_barrier (tl_blk->bar)
#define g_warp_barrier (tl_blk->wbar)
#define g_xchg (tl_blk->xchg)

// Per-block storage for `__shared__` declarations that tests/emu_build.py rewrote (needed where the blocks of a cluster run
// at the same time; a function-local static would be one array for all of them).  Every declaration has an id; its offset in
// the block's arena is fixed the first time any thread asks for it.
static inline void *emu_block_static(int id, size_t bytes) {
  static std::mutex mu;
  static std::vector<size_t> offset;
  static size_t used = 0;
  std::lock_guard<std::mutex> lk(mu);
  if ((int)offset.size() <= id) offset.resize((size_t)id + 1, (size_t)-1);
  if (offset[(size_t)id] == (size_t)-1) {
    used = (used + 15) & ~(size_t)15;
    offset[(size_t)id] = used;
    used += bytes;
    if (used > EMU_STATIC_ARENA) { fprintf(stderr, "cuda_emu: static shared arena exhausted\n"); abort(); }
  }
  return tl_blk->stat + offset[(size_t)id];
}

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
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }       // approximate on the GPU (2 ulp): no bit parity through it
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
static inline unsigned __vabsdiffu4(unsigned a, unsigned b) {               // per-byte absolute difference
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) { const int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF; r |= (unsigned)(x > y ? x - y : y - x) << (8 * i); }
  return r;
}
static inline int __any_sync(unsigned, int pred);                           // defined with the other warp votes below
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
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline unsigned __match_any_sync(unsigned, unsigned v) {            // lanes of the warp that hold the same value
  const int w = threadIdx.x >> 5;
  g_xchg[w][threadIdx.x & 31] = v;
  pthread_barrier_wait(&g_warp_barrier[w]);
  unsigned m = 0;
  const unsigned lanes = std::min(32u, blockDim.x - 32u * (unsigned)w);
  for (unsigned i = 0; i < lanes; ++i) m |= (unsigned)(g_xchg[w][i] == (unsigned long long)v) << i;
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

static inline void emu_block_init(EmuBlock *b, unsigned threads, size_t dyn_bytes) {
  const unsigned warps = (threads + 31) / 32;
  pthread_barrier_init(&b->bar, nullptr, threads);
  for (unsigned w = 0; w < warps; ++w) pthread_barrier_init(&b->wbar[w], nullptr, std::min(32u, threads - 32 * w));
  b->dyn_bytes = dyn_bytes;
  b->dyn = (unsigned char *)calloc(dyn_bytes + 64, 1);
  b->stat = (unsigned char *)calloc(EMU_STATIC_ARENA, 1);
}
static inline void emu_block_destroy(EmuBlock *b, unsigned threads) {
  const unsigned warps = (threads + 31) / 32;
  pthread_barrier_destroy(&b->bar);
  for (unsigned w = 0; w < warps; ++w) pthread_barrier_destroy(&b->wbar[w]);
  free(b->dyn);
  free(b->stat);
}

// blocks strictly one after the other (function-local `static` __shared__ arrays are reused block after block)
template <class F> static void run_grid(unsigned gx, unsigned gy, unsigned gz, unsigned threads, size_t dyn_smem_bytes, F &&kernel) {
  EmuBlock blk;
  emu_block_init(&blk, threads, dyn_smem_bytes);
  pthread_barrier_t end_of_block;
  pthread_barrier_init(&end_of_block, nullptr, threads);
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (unsigned t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      threadIdx = EmuIdx{t, 0, 0};
      gridDim = EmuIdx{gx, gy, gz};
      blockDim = EmuIdx{threads, 1, 1};
      tl_blk = &blk;
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
  emu_block_destroy(&blk, threads);
}

// one thread-block cluster after the other; the blocks of a cluster run at the same time (distributed shared memory, cluster
// barriers).  Kernels launched this way must not keep block state in function-local statics: tests/emu_build.py rewrites their
// `__shared__` declarations into per-block storage (EMU_SHARED).
template <class F> static void run_clusters(unsigned grid_x, unsigned cluster_size, unsigned threads, size_t dyn_smem_bytes, F &&kernel) {
  for (unsigned c0 = 0; c0 < grid_x; c0 += cluster_size) {
    EmuCluster cl;
    std::vector<EmuBlock> blocks(cluster_size);
    pthread_barrier_init(&cl.bar, nullptr, cluster_size * threads);
    for (unsigned r = 0; r < cluster_size; ++r) {
      emu_block_init(&blocks[r], threads, dyn_smem_bytes);
      blocks[r].rank = r;
      blocks[r].cluster = &cl;
      cl.blocks.push_back(&blocks[r]);
    }
    std::vector<std::thread> pool;
    pool.reserve((size_t)cluster_size * threads);
    for (unsigned r = 0; r < cluster_size; ++r)
      for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([&, r, t] {
          threadIdx = EmuIdx{t, 0, 0};
          blockIdx = EmuIdx{c0 + r, 0, 0};
          gridDim = EmuIdx{grid_x, 1, 1};
          blockDim = EmuIdx{threads, 1, 1};
          tl_blk = &blocks[r];
          kernel();
        });
    for (auto &th : pool) th.join();
    for (unsigned r = 0; r < cluster_size; ++r) emu_block_destroy(&blocks[r], threads);
    pthread_barrier_destroy(&cl.bar);
  }
}

// cooperative_groups, as far as csrc/ba.cu uses it
namespace cooperative_groups {
struct cluster_group {
  unsigned block_rank() const { return tl_blk->rank; }
  unsigned num_blocks() const { return (unsigned)tl_blk->cluster->blocks.size(); }
  void sync() const { pthread_barrier_wait(&tl_blk->cluster->bar); }
  template <class T> T *map_shared_rank(T *p, unsigned r) const {
    const unsigned char *q = (const unsigned char *)p;
    EmuBlock *me = tl_blk, *other = me->cluster->blocks[r];
    if (q >= me->dyn && q < me->dyn + me->dyn_bytes + 64) return (T *)(other->dyn + (q - me->dyn));
    if (q >= me->stat && q < me->stat + EMU_STATIC_ARENA) return (T *)(other->stat + (q - me->stat));
    fprintf(stderr, "cuda_emu: map_shared_rank of an address outside the block's shared memory\n");
    abort();
  }
};
static inline cluster_group this_cluster() { return cluster_group(); }
}  // namespace cooperative_groups

// cudaLaunchKernelEx with a cluster-dimension attribute (the only extended launch in this code base; tests/emu_build.py
// renames the call: cuda_runtime.h has its own host template of that name)
template <class... P, class... A> static inline cudaError_t emu_cudaLaunchKernelEx(const cudaLaunchConfig_t *cfg, void (*kernel)(P...), A &&...args) {
  unsigned csize = 1;
  for (unsigned i = 0; i < cfg->numAttrs; ++i)
    if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension) csize = cfg->attrs[i].val.clusterDim.x;
  run_clusters(cfg->gridDim.x, csize, cfg->blockDim.x, cfg->dynamicSmemBytes, [&] { kernel(args...); });
  return cudaSuccess;
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
