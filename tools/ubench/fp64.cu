// micro-benchmark: dependent-chain latency and per-SM throughput of DFMA / FFMA / DADD / SHFL on this GPU
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void chain(T *out, T a, T b, int iters, long long *cyc) {
  T x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = (T)threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = x[i] * a + b;
  }
  long long t1 = clock64();
  T s = 0;
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_chain(double *out, int iters, long long *cyc) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) x += __shfl_xor_sync(0xffffffffu, x, 1 + (it & 15));
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void div_chain(double *out, double a, int iters, long long *cyc) {
  double x = 1.0 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) x = a / x + 1.0;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *o; long long *c, h;
  cudaMalloc(&o, 1 << 24); cudaMalloc(&c, 8);
  const int iters = 4096;
#define RUN(name, call, ops) call; cudaDeviceSynchronize(); call; cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-40s %8.2f cycles/op\n", name, (double)h / (ops));
  RUN("DFMA latency (1 warp, ILP1)", (chain<double,1><<<1,32>>>(o, 1.0000001, 1e-9, iters, c)), iters);
  RUN("DFMA 1 warp ILP8 (per op)", (chain<double,8><<<1,32>>>(o, 1.0000001, 1e-9, iters, c)), iters*8);
  RUN("DFMA 16 warps ILP8 (per warp-op/SM)", (chain<double,8><<<1,512>>>(o, 1.0000001, 1e-9, iters, c)), iters*8*16);
  RUN("DFMA 32 warps ILP8 (per warp-op/SM)", (chain<double,8><<<1,1024>>>(o, 1.0000001, 1e-9, iters, c)), iters*8*32);
  RUN("FFMA latency (1 warp, ILP1)", (chain<float,1><<<1,32>>>((float*)o, 1.0000001f, 1e-9f, iters, c)), iters);
  RUN("FFMA 32 warps ILP8 (per warp-op/SM)", (chain<float,8><<<1,1024>>>((float*)o, 1.0000001f, 1e-9f, iters, c)), iters*8*32);
  RUN("SHFL.64+DADD dependent chain", (shfl_chain<<<1,32>>>(o, iters, c)), iters);
  RUN("fp64 divide dependent chain", (div_chain<<<1,32>>>(o, 3.0, iters, c)), iters);
  return 0;
}
