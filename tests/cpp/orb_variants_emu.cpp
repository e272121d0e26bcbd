// TEST INFRASTRUCTURE (CPU tier): csrc/orb_variants.cuh — the very text nvcc compiles for k_blur2, the shipped descriptor blur —
// compiled for the host and executed one OS thread per CUDA thread, one block at a time:
// threadIdx / blockIdx are thread-local, __shared__ arrays are function-local statics (one block runs at a time),
// __syncthreads() is a barrier over the block's threads, __shfl_xor_sync an exchange through a per-warp buffer with two
// warp barriers, and the round-to-nearest intrinsics are plain float operations (built with -ffp-contract=off).
// This checks the index arithmetic and the data flow of the kernels against the oracle without a GPU; what it cannot
// show is timing and anything specific to the hardware's scheduling.
#include "cuda_emu.h"

#include "orb.cuh"
#include "pdl_device.cuh"   // pdl_wait / pdl_launch_dependents: no-ops on the CPU tier
namespace {
#include "orb_pattern.inc"
float c_gauss7[7];
constexpr int BLUR_TW = 128, BLUR_TH = 16;                        // as in orb.cu
struct BlurTiles { int first[MVO_MAX_LEVELS + 1]; int nx[MVO_MAX_LEVELS]; };
constexpr int DESC_WARPS = 8;
#include "orb_variants.cuh"

// one pyramid level per "plan level", planes laid out like orb_host.cpp does (pitch multiple of 128, 256-byte aligned offsets)
OrbPlanDev make_plan(int nlevels, const int *w, const int *h, const float *scale, std::vector<uint8_t> *planes) {
  OrbPlanDev pl;
  memset(&pl, 0, sizeof pl);
  pl.nlevels = nlevels;
  size_t off = 0;
  for (int l = 0; l < nlevels; ++l) {
    OrbLevelDev &L = pl.lv[l];
    L.w = w[l]; L.h = h[l]; L.scale = scale[l];
    L.pitch = (L.w + 127) & ~127;
    L.img_off = (uint32_t)off; off = (off + (size_t)L.pitch * L.h + 255) & ~(size_t)255;
    L.blur_off = (uint32_t)off; off = (off + (size_t)L.pitch * L.h + 255) & ~(size_t)255;
  }
  pl.slot_bytes = (uint32_t)off;
  planes->assign(off + 64, 0xA5);                                  // padding bytes hold junk, like recycled device memory
  return pl;
}
}  // namespace

extern "C" {

// imgs: the level images back to back (w[l] x h[l] bytes each); out: the blurred levels in the same packing
int emu_blur2(int nlevels, const int *w, const int *h, const uint8_t *imgs, uint8_t *out) {
  double k[7], s = 0;                                              // cv::getGaussianKernel(7, 2, CV_32F), as upload_gauss (orb.cu)
  for (int i = 0; i < 7; ++i) { const double x = i - 3; k[i] = exp(-x * x / 8.0); s += k[i]; }
  for (int i = 0; i < 7; ++i) c_gauss7[i] = (float)(k[i] / s);
  std::vector<uint8_t> planes;
  std::vector<float> scale((size_t)nlevels, 1.f);
  const OrbPlanDev pl = make_plan(nlevels, w, h, scale.data(), &planes);
  const uint8_t *src = imgs;
  for (int l = 0; l < nlevels; ++l) {
    for (int y = 0; y < h[l]; ++y) memcpy(&planes[pl.lv[l].img_off + (size_t)y * pl.lv[l].pitch], src + (size_t)y * w[l], (size_t)w[l]);
    src += (size_t)w[l] * h[l];
  }
  BlurTiles tiles;
  int total = 0;
  for (int l = 0; l < nlevels; ++l) {                              // as orb_launch_blur
    tiles.first[l] = total;
    tiles.nx[l] = (w[l] + BLUR_TW - 1) / BLUR_TW;
    total += tiles.nx[l] * ((h[l] + BLUR_TH - 1) / BLUR_TH);
  }
  tiles.first[nlevels] = total;
  uint8_t *p = planes.data();
  run_grid((unsigned)total, 1, 1, 256, 0, [&] { k_blur2(pl, tiles, p); });
  uint8_t *dst = out;
  for (int l = 0; l < nlevels; ++l) {
    for (int y = 0; y < h[l]; ++y) memcpy(dst + (size_t)y * w[l], &planes[pl.lv[l].blur_off + (size_t)y * pl.lv[l].pitch], (size_t)w[l]);
    dst += (size_t)w[l] * h[l];
  }
  return 0;
}


}  // extern "C"
