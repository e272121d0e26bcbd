// Host side of the matching entry points: staging, the reference's threshold rules and the
// libstdc++-dependent duplicate removal (kept on the host on purpose, SURVEY.md §7 hard part 3).
#include <algorithm>
#include <string.h>
#include "mvo_internal.h"

namespace {

// Upload both descriptor sets (+ optional coordinates), run the all-pairs kernel, bring the
// packed keys back.  Returns a pointer to the keys in pinned memory (valid until next call).
int run_match(mvo_ctx *ctx, int mode, const uint8_t *d1, const float *xy1, int n1,
              const uint8_t *d2, const float *xy2, int n2, float radius, const uint32_t **keys_out,
              bool d2_on_device = false) {
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if (n1 < 0 || n2 < 0 || (n1 > 0 && !d1) || (n2 > 0 && !d2))
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "match: null descriptors or negative count");
  if (n1 > 65535 || n2 > 65535)
    return mvo_fail(ctx, MVO_ERR_UNSUPPORTED, "match: more than 65535 descriptors in a set");
  if (mode == 2 && n1 > 0 && n2 > 0 && (!xy1 || !xy2))
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "match: method 3 needs keypoint coordinates");
  MVO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int W = mode == 1 ? 2 : 1;
  const size_t b1 = (size_t)n1 * 32, b2 = d2_on_device ? 0 : (size_t)n2 * 32;
  const size_t x1 = mode == 2 ? (size_t)n1 * 8 : 0, x2 = mode == 2 ? (size_t)n2 * 8 : 0;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_d1 = 0, o_d2 = al(o_d1 + b1), o_x1 = al(o_d2 + b2), o_x2 = al(o_x1 + x1);
  const size_t in_bytes = al(o_x2 + x2);
  const size_t key_bytes = (size_t)n1 * W * 4;
  MVO_TRY(mvo_reserve(ctx, ctx->match_in, in_bytes + 256));
  MVO_TRY(mvo_reserve(ctx, ctx->match_keys, key_bytes + 256));
  MVO_TRY(mvo_reserve_pinned(ctx, ctx->match_h, in_bytes + key_bytes + 512));
  uint8_t *h = (uint8_t *)ctx->match_h.p;
  uint8_t *d = (uint8_t *)ctx->match_in.p;
  if (b1) memcpy(h + o_d1, d1, b1);
  if (b2) memcpy(h + o_d2, d2, b2);
  if (x1) memcpy(h + o_x1, xy1, x1);
  if (x2) memcpy(h + o_x2, xy2, x2);
  if (in_bytes) MVO_CUDA(ctx, cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  MVO_TRY(mvo_match_launch(ctx, mode, d + o_d1, (const float *)(d + o_x1), n1, d2_on_device ? d2 : d + o_d2,
                           (const float *)(d + o_x2), n2, radius, (uint32_t *)ctx->match_keys.p));
  uint32_t *hk = (uint32_t *)(h + in_bytes + 256);
  if (key_bytes)
    MVO_CUDA(ctx, cudaMemcpyAsync(hk, ctx->match_keys.p, key_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *keys_out = hk;
  return MVO_OK;
}

inline mvo_dmatch unpack(int q, uint32_t key, int img_idx, bool sad) {
  mvo_dmatch m;
  m.query_idx = q;
  m.train_idx = (int)(key & 0xFFFFu);
  m.img_idx = img_idx;
  const uint32_t d = key >> 16;
  // method 3: cv::sum(diff)[0] / descriptors_1.cols as double, then static_cast<float>
  // (feature_match.cpp:111,122)
  m.distance = sad ? (float)((double)d / 32.0) : (float)d;
  return m;
}

}  // namespace

extern "C" {

int mvo_match_hamming_nn(mvo_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2,
                         mvo_dmatch *out) {
  const uint32_t *k = nullptr;
  MVO_TRY(run_match(ctx, 0, d1, nullptr, n1, d2, nullptr, n2, 0.f, &k));
  if (n1 > 0 && !out) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "null output");
  if (n2 == 0) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "empty train set");
  for (int i = 0; i < n1; ++i) out[i] = unpack(i, k[i], 0, false);
  return MVO_OK;
}

int mvo_match_hamming_knn2(mvo_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2,
                           mvo_dmatch *out) {
  if (ctx && n2 < 2) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "knn2 needs >= 2 train descriptors");
  const uint32_t *k = nullptr;
  MVO_TRY(run_match(ctx, 1, d1, nullptr, n1, d2, nullptr, n2, 0.f, &k));
  if (n1 > 0 && !out) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "null output");
  for (int i = 0; i < n1; ++i) {
    out[2 * i] = unpack(i, k[2 * i], 0, false);
    out[2 * i + 1] = unpack(i, k[2 * i + 1], 0, false);
  }
  return MVO_OK;
}

int mvo_match_radius_sad(mvo_ctx *ctx, const uint8_t *d1, const float *xy1, int n1,
                         const uint8_t *d2, const float *xy2, int n2, float radius,
                         mvo_dmatch *out, int *n_out) {
  if (!n_out) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "null n_out");
  *n_out = 0;
  const uint32_t *k = nullptr;
  MVO_TRY(run_match(ctx, 2, d1, xy1, n1, d2, xy2, n2, radius, &k));
  if (n1 > 0 && !out) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "null output");
  int n = 0;
  for (int i = 0; i < n1; ++i)
    if (k[i] != 0xFFFFFFFFu) out[n++] = unpack(i, k[i], -1, true);
  *n_out = n;
  return MVO_OK;
}

int mvo_remove_duplicated_matches(mvo_dmatch *matches, int *n) {
  if (!n || (*n > 0 && !matches)) return MVO_ERR_INVALID_ARG;
  // feature_match.cpp:244-247: std::sort (unstable) on trainIdx only, then unique-by-run.
  std::sort(matches, matches + *n,
            [](const mvo_dmatch &a, const mvo_dmatch &b) { return a.train_idx < b.train_idx; });
  int w = 0;
  for (int i = 0; i < *n; ++i)
    if (i == 0 || matches[i].train_idx != matches[i - 1].train_idx) matches[w++] = matches[i];
  *n = w;
  return MVO_OK;
}

int mvo_match_features(mvo_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2,
                       int method_index, const float *xy1, const float *xy2, float radius,
                       mvo_dmatch *out, int *n_out) {
  return mvo_match_features_ex(ctx, d1, n1, d2, n2, 0, method_index, xy1, xy2, radius, out, n_out);
}

}  // extern "C"

// The tail of matchFeatures over the packed keys of the all-pairs kernel (one key per query for methods 1 / 3, two for method 2):
// the distance thresholds (feature_match.cpp:179-196 / :210-217) and removeDuplicatedMatches (:229, :241-260)
int mvo_match_filter_keys(mvo_ctx *ctx, int method_index, const uint32_t *k, int n1, mvo_dmatch *out, int *n_out) {
  int n = 0;
  if (method_index == 1 || method_index == 3) {
    const bool sad = method_index == 3;
    // feature_match.cpp:179-187
    double min_dis = 9999999, max_dis = 0;
    for (int i = 0; i < n1; ++i) {
      if (k[i] == 0xFFFFFFFFu) continue;
      const double dist = unpack(i, k[i], 0, sad).distance;
      if (dist < min_dis) min_dis = dist;
      if (dist > max_dis) max_dis = dist;
    }
    const double thr = std::max<float>(min_dis * ctx->prm.xiang_gao_ratio, 30.0);
    for (int i = 0; i < n1; ++i) {   // :194-196
      if (k[i] == 0xFFFFFFFFu) continue;
      const mvo_dmatch m = unpack(i, k[i], sad ? -1 : 0, sad);
      if (m.distance < thr) out[n++] = m;
    }
  } else {
    for (int i = 0; i < n1; ++i) {   // :210-217
      const mvo_dmatch m0 = unpack(i, k[2 * i], 0, false), m1 = unpack(i, k[2 * i + 1], 0, false);
      const double dist = m0.distance;
      if (dist < ctx->prm.lowe_ratio * m1.distance) out[n++] = m0;
    }
  }
  mvo_remove_duplicated_matches(out, &n);   // :229
  *n_out = n;
  return MVO_OK;
}

int mvo_match_features_ex(mvo_ctx *ctx, const uint8_t *d1, int n1, const uint8_t *d2, int n2, int d2_on_device,
                          int method_index, const float *xy1, const float *xy2, float radius, mvo_dmatch *out, int *n_out) {
  const bool dev2 = d2_on_device != 0;
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if (!n_out) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "null n_out");
  *n_out = 0;
  if (method_index < 1 || method_index > 3)   // feature_match.cpp:225 throws
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "matchFeatures: wrong method index %d", method_index);
  if (n1 > 0 && !out) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "null output");
  const uint32_t *k = nullptr;
  if (method_index == 1 || method_index == 3) {
    const bool sad = method_index == 3;
    if (!sad && n2 == 0) return MVO_OK;   // BFMatcher on an empty train set returns no matches
    MVO_TRY(run_match(ctx, sad ? 2 : 0, d1, xy1, n1, d2, xy2, n2, radius, &k, dev2));
  } else {
    if (n2 < 2) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "method 2 needs >= 2 train descriptors");
    MVO_TRY(run_match(ctx, 1, d1, nullptr, n1, d2, nullptr, n2, 0.f, &k, dev2));
  }
  return mvo_match_filter_keys(ctx, method_index, k, n1, out, n_out);
}

extern "C" {

int mvo_match_dev(mvo_ctx *ctx, int mode, const uint8_t *d_d1, const float *d_xy1, int n1,
                  const uint8_t *d_d2, const float *d_xy2, int n2, float radius, uint32_t *d_keys) {
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if ((n1 > 0 && (!d_d1 || !d_keys)) || (n2 > 0 && !d_d2))
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "match_dev: null pointer");
  return mvo_match_launch(ctx, mode, d_d1, d_xy1, n1, d_d2, d_xy2, n2, radius, d_keys);
}

}  // extern "C"
