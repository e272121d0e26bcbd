// TEST INFRASTRUCTURE (CPU tier): k_match_filter (csrc/track_filter.cuh) built for the host with tests/cpp/cuda_emu.h and
// launched the way mvo_track_match_filter (csrc/track.cu) launches it: one block of 1024 threads, the same capacity and
// dynamic shared-memory size.  The reference it is compared with is the library's own host filter running the real
// libstdc++ std::sort (mvo_test_match_filter_host), as in tests/test_track_filter_gpu.py.
#include "cuda_emu.h"

#include <string.h>
#include "mvo_internal.h"
namespace {
#define MVO_DYN_SMEM(type, name) type *name = (type *)g_dyn_smem
#include "track_filter.cuh"
}  // namespace

extern "C" {

// a context object for the host-side reference (it reads the match ratios from ctx->prm, nothing else)
mvo_ctx *emu_params_ctx(double xiang_gao_ratio, double lowe_ratio) {
  mvo_ctx *c = new mvo_ctx();
  mvo_default_params(&c->prm);
  c->prm.xiang_gao_ratio = xiang_gao_ratio;
  c->prm.lowe_ratio = lowe_ratio;
  return c;
}
void emu_params_ctx_free(mvo_ctx *c) { delete c; }

int emu_match_filter(const uint32_t *keys, const uint8_t *vis, int nmap, int nk, int method, double xiang_gao_ratio, double lowe_ratio,
                     int32_t *pairs, int32_t *info /* 64 */) {
  int cap = 2048;
  while (cap < std::max(std::min(nmap, 65535), nk) && cap < MF_MAXN) cap *= 2;
  FilterArgs a;
  memset(&a, 0, sizeof a);
  std::vector<uint8_t> v(vis, vis + nmap);
  a.keys = keys; a.vis = v.data(); a.nmap = nmap; a.nk = nk; a.method = method; a.n_cap = cap;
  a.xg_ratio = xiang_gao_ratio; a.lowe_ratio = lowe_ratio;
  a.pairs = (int2 *)pairs; a.info = info;
  const size_t smem = (size_t)cap * (4 + 4 + 2 + 2 + 2) + 3 * ((size_t)cap / 16 + 2) * 4 + ((size_t)cap / 16 + 2) + 64;
  run_grid(1, 1, 1, MF_T, smem, [&] { k_match_filter(a); });
  return 0;
}

}  // extern "C"
