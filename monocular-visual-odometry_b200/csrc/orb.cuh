// Device-visible plan of one ORB extraction (pyramid geometry, FAST bands, selection grid).
// Shared by orb.cu (kernels) and orb_host.cpp (orchestration + the libstdc++-dependent slow path).
#pragma once
#include <stdint.h>
#include "mvo_internal.h"

#define ORB_BAND_H 8          // NMS rows per FAST band (one CTA per band)
#define ORB_EDGE 31           // cv::ORB edgeThreshold
#define ORB_HALF_PATCH 15
#define ORB_MAX_BANDS 1024
#define ORB_MAX_W 4096

struct OrbLevelDev {
  int w, h, pitch;
  int band_first, nbands;
  int cap;                 // featuresPerLevel
  float scale;
  uint32_t img_off, blur_off;          // byte offsets inside a frame slot
  uint32_t tab_off;                    // int32 offsets into the resize tables: xofs,xw1,yofs,yw1
};

struct OrbPlanDev {
  int nlevels, rows, cols;
  int total_bands, band_cap;           // staging entries per band
  int grid_rows, grid_cols, grid_size, max_per_cell, max_kpts;
  int cand_cap;                        // compact candidate capacity per frame (upper bound of NMS output)
  int sel_cap;                         // fast-path bound on candidates (= nfeatures = sum of level caps)
  int fast_threshold;
  uint32_t slot_bytes;
  // fused gray + pyramid kernel (k_pyramid): the image is cut into pyr_nb horizontal bands; per band and level the rows it has to
  // hold in shared memory and the rows it owns (writes), 4 ints each, at tables + pyr_tab_off; pyr_rows[l] = most rows of level l a band holds
  int pyr_nb;
  uint32_t pyr_tab_off;
  int pyr_rows[MVO_MAX_LEVELS];
  OrbLevelDev lv[MVO_MAX_LEVELS];
};

// Per-frame metadata written by the selection kernel.
struct OrbFrameMeta {
  int32_t n_sel;                       // keypoints selected on the fast path
  int32_t overflow;                    // 0: no level exceeds featuresPerLevel; 1: some do (k_select) -> 3: retainBest + selection done on
                                       // the device (k_retain, k_select_kept); 2: beyond the device path -> host retainBest
  int32_t n_cand;                      // total FAST candidates (all levels)
  int32_t lvl_count[MVO_MAX_LEVELS];
  int32_t pad[5];
};
static_assert(sizeof(OrbFrameMeta) == 64, "meta layout");

// candidate packing: x | y << 12 | score << 24   (level coordinates, x,y < 4096, score < 256)
__host__ __device__ inline uint32_t orb_pack(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
__host__ __device__ inline int orb_px(uint32_t p) { return p & 0xFFF; }
__host__ __device__ inline int orb_py(uint32_t p) { return (p >> 12) & 0xFFF; }
__host__ __device__ inline int orb_ps(uint32_t p) { return p >> 24; }

// launchers (orb.cu), all asynchronous on ctx->stream
int orb_launch_gray(mvo_ctx *ctx, const OrbPlanDev &plan, const uint8_t *d_in, int channels, size_t stride,
                    size_t frame_stride, uint8_t *planes, int batch);
int orb_launch_pyramid(mvo_ctx *ctx, const OrbPlanDev &plan, const int32_t *tables, uint8_t *planes, int batch);
// both in one launch (k_pyramid) when the bands fit in shared memory, else the two launchers above
int orb_launch_gray_pyramid(mvo_ctx *ctx, const OrbPlanDev &plan, const uint8_t *d_in, int channels, size_t stride, size_t frame_stride,
                            const int32_t *tables, uint8_t *planes, int batch);
int orb_launch_fast(mvo_ctx *ctx, const OrbPlanDev &plan, const uint8_t *planes, uint32_t *staging,
                    int32_t *bandcnt, int batch);
int orb_launch_select(mvo_ctx *ctx, const OrbPlanDev &plan, const uint32_t *staging, const int32_t *bandcnt,
                      uint32_t *cand, uint2 *sel, OrbFrameMeta *meta, int batch);
int orb_launch_blur(mvo_ctx *ctx, const OrbPlanDev &plan, uint8_t *planes, int batch);
int orb_launch_harris_all(mvo_ctx *ctx, const OrbPlanDev &plan, const uint8_t *planes, const uint32_t *cand,
                          const OrbFrameMeta *meta, float *harris, int batch);
int orb_retain_max(void);
int orb_launch_retain(mvo_ctx *ctx, const OrbPlanDev &plan, const uint32_t *cand, const float *harris, OrbFrameMeta *meta,
                      uint16_t *kept, int32_t *kept_cnt, uint2 *sel, int batch);
// mode 0: from selection list (sel) -> keypoints (+ descriptors if with_desc)
int orb_launch_describe_sel(mvo_ctx *ctx, const OrbPlanDev &plan, const uint8_t *planes, const uint2 *sel,
                            const OrbFrameMeta *meta, const int32_t *n_override, mvo_keypoint *kpts, uint8_t *desc,
                            int32_t *counts, int out_cap, int with_desc, int batch);
// mode 1: descriptors for caller keypoints (single frame)
int orb_launch_describe_kpts(mvo_ctx *ctx, const OrbPlanDev &plan, const uint8_t *planes, const mvo_keypoint *kpts,
                             int n, uint8_t *desc, int32_t *bad_flag);
