// Self-test of tests/cpp/cuda_emu.h: a kernel that uses every emulated collective and atomic, with a known answer.
#include "cuda_emu.h"

namespace {
__global__ void __launch_bounds__(96)
k_selftest(const int *__restrict__ in, int n, int *__restrict__ out) {
  __shared__ int s_hist[4];
  __shared__ unsigned s_bits;
  __shared__ int s_scan[96];
  const int tid = threadIdx.x, lane = tid & 31, gid = blockIdx.x * blockDim.x + tid;
  if (tid < 4) s_hist[tid] = 0;
  if (tid == 0) s_bits = 0;
  __syncthreads();
  const int v = gid < n ? in[gid] : 0;
  atomicAdd(&s_hist[v & 3], 1);
  const unsigned odd = __ballot_sync(0xffffffffu, v & 1);
  if (lane == 0) atomicOr(&s_bits, odd ? 1u << (tid >> 5) : 0u);
  int sum = v;                                                    // warp sum by butterfly
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  int incl = v;                                                   // warp inclusive scan
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
  s_scan[tid] = incl;
  __syncwarp();
  const int first = __shfl_sync(0xffffffffu, v, 0), next = __shfl_down_sync(0xffffffffu, v, 1);
  __syncthreads();
  int *o = out + (size_t)gid * 6;
  o[0] = sum; o[1] = incl; o[2] = first; o[3] = next; o[4] = __popc(odd & emu_lanemask_lt()); o[5] = s_scan[(tid & ~31) + 31];
  if (tid < 4) atomicAdd(&out[(size_t)gridDim.x * blockDim.x * 6 + tid], s_hist[tid]);
  if (tid == 0) atomicMax(&out[(size_t)gridDim.x * blockDim.x * 6 + 4], (int)s_bits);
}
}  // namespace

extern "C" int emu_selftest(const int *in, int n, int blocks, int *out) {
  run_grid((unsigned)blocks, 1, 1, 96, 0, [&] { k_selftest(in, n, out); });
  return (int)__vsadu4(0x10FF0580u, 0x20F00A7Fu) + 1000 * __ffs(0x50) + 100000 * __popc(0xF0F0u);
}
