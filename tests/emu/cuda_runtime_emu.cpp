// TEST INFRASTRUCTURE (CPU tier): the handful of CUDA runtime entry points the library's host code calls, over host memory —
// "device" memory is malloc'ed host memory, streams and events are inert (work is synchronous), copies are memcpy.  Linked
// with host builds of the product's translation units (tests/emu_build.py) in place of libcudart.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>

namespace {
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct EmuEvent { double t = 0; };
}  // namespace

extern "C" {
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
#ifdef cudaGetDeviceProperties
#undef cudaGetDeviceProperties
#endif
static void fill_prop(cudaDeviceProp *p) {
  memset(p, 0, sizeof *p);
  strcpy(p->name, "emulated sm_100 (tests/emu)");
  p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
  p->sharedMemPerBlockOptin = 227 * 1024; p->sharedMemPerBlock = 48 * 1024;
}
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp *p, int) { fill_prop(p); return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { fill_prop(p); return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
  for (size_t r = 0; r < h; ++r) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemcpyToSymbolAsync(const void *sym, const void *s, size_t n, size_t off, cudaMemcpyKind, cudaStream_t) {
  memcpy((char *)sym + off, s, n);
  return cudaSuccess;
}
// no driver behind the emulation: the product then stages the FAST bands with plain loads instead of TMA
cudaError_t cudaGetDriverEntryPoint(const char *, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q) {
  if (fn) *fn = nullptr;
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorNotSupported;
}
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "emulated runtime"; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t) new EmuEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { ((EmuEvent *)e)->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(((EmuEvent *)b)->t - ((EmuEvent *)a)->t); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete (EmuEvent *)e; return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { memset(a, 0, sizeof *a); a->type = cudaMemoryTypeHost; return cudaSuccess; }
}
