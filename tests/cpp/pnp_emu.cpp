// TEST INFRASTRUCTURE (CPU tier): the RANSAC PnP kernels (csrc/pnp_kernels.cuh) built for the host with tests/cpp/cuda_emu.h
// and launched the way pnp_enqueue (csrc/pnp.cu) launches them: P3P hypotheses, scoring, arg-max + consensus set.  The
// least-squares refit that follows on the GPU is the cluster LM kernel of ba.cu and is not part of this build.
#include "cuda_emu.h"

#include <string.h>
#include "mvo_internal.h"
namespace {
#define PNP_DYN_SMEM(type, name) type *name = (type *)g_dyn_smem
#include "pnp_kernels.cuh"
}  // namespace

extern "C" int emu_pnp_ransac(const float *p3, const float *p2, int n, const double *K, double reproj_error, int H, uint64_t seed,
                              double *poses /* H x 12 */, int32_t *counts /* H */, double *pose_best /* 12 */, int32_t *out_i /* 3 */,
                              int32_t *inl /* n */) {
  PnpCam cam;
  cam.fx = K[0]; cam.fy = K[4]; cam.cx = K[2]; cam.cy = K[5];
  const double thr2 = reproj_error * reproj_error;
  std::vector<int32_t> valid((size_t)H), ef((size_t)n + 8);
  std::vector<double> ex((size_t)n * 3 + 8), eo((size_t)n * 2 + 8);
  run_grid((unsigned)((H + 127) / 128), 1, 1, 128, 0, [&] { k_pnp_hypotheses(p3, p2, n, nullptr, cam, seed, H, poses, valid.data()); });
  const int grid = std::min((H + 7) / 8, 48);                    // grid-stride loop: any grid covers every hypothesis
  run_grid((unsigned)grid, 1, 1, 256, (size_t)n * 5 * sizeof(float), [&] { k_pnp_score(p3, p2, n, nullptr, cam, thr2, H, poses, valid.data(), counts); });
  run_grid(1, 1, 1, FIN_T, 0, [&] { k_pnp_finish(p3, p2, n, nullptr, cam, thr2, H, poses, counts, 0, 0, pose_best, out_i, inl, ex.data(), eo.data(), ef.data()); });
  return out_i[0];
}
