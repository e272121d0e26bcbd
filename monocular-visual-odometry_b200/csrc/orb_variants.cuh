// Experimental lower-instruction-count variants of two ORB kernels (selected with MVO_BLUR2=1 / MVO_DESCRIBE2=1, off by
// default, not yet run on hardware).  Kept in a header of their own so that the CPU test tier can compile the very same
// kernel text for the host and run it thread by thread (tests/cpp/orb_variants_emu.cpp) against the oracle.
//
// Included by orb.cu INSIDE its anonymous namespace, after these have been defined there: c_gauss7[7] (the Gaussian taps),
// BLUR_TW / BLUR_TH, struct BlurTiles, DESC_WARPS, kOrbPattern[512][2] and the declarations of orb.cuh.  Helpers that the
// shipped kernels also have are repeated here under a _v suffix so that the shipped kernels' code is not touched.
#pragma once

__device__ __forceinline__ int warp_sum_v(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cv::fastAtan2 scalar path (degrees); every product/sum rounded separately, like the C++ build.
__device__ __forceinline__ float fast_atan2_deg_v(float y, float x) {
  constexpr float r2d = (float)(180.0 / 3.1415926535897932384626433832795);
  constexpr float p1 = 0.9997878412794807f * r2d, p3 = -0.3258083974640975f * r2d;
  constexpr float p5 = 0.1555786518463281f * r2d, p7 = -0.04432655554792128f * r2d;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  const float eps = 2.2204460492503131e-16f;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = __fsub_rn(180.f, a);
  if (y < 0.f) a = __fsub_rn(360.f, a);
  return a;
}

// OpenCV HarrisResponses: 7x7 block, Sobel-like 3x3 gradients, integer sums.
__device__ __forceinline__ float harris_warp_v(const uint8_t *__restrict__ img, int pitch, int x0, int y0, int lane) {
  int a = 0, b = 0, c = 0;
#pragma unroll
  for (int rep = 0; rep < 2; ++rep) {
    const int k = lane + 32 * rep;
    if (k < 49) {
      const int i = k / 7, j = k - 7 * i;
      const uint8_t *p = img + (y0 - 3 + i) * pitch + (x0 - 3 + j);
      const int Ix = (p[1] - p[-1]) * 2 + (p[-pitch + 1] - p[-pitch - 1]) + (p[pitch + 1] - p[pitch - 1]);
      const int Iy = (p[pitch] - p[-pitch]) * 2 + (p[pitch - 1] - p[-pitch - 1]) + (p[pitch + 1] - p[-pitch + 1]);
      a += Ix * Ix;
      b += Iy * Iy;
      c += Ix * Iy;
    }
  }
  a = warp_sum_v(a);
  b = warp_sum_v(b);
  c = warp_sum_v(c);
  const float scale = __fdiv_rn(1.f, __fmul_rn(28.f, 255.f));            // 1.f/((1<<2)*7*255.f)
  const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
  const float fa = (float)a, fb = (float)b, fc = (float)c;
  const float ab = __fadd_rn(fa, fb);
  return __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, ab), ab)), s4);
}

// Same filter, same arithmetic (identical operand order, no FMA), fewer instructions: the tile is moved as 32-bit words,
// the row pass produces four outputs per thread from three shared-memory words, the column pass two rows of four outputs
// per thread from eight float4 rows.  ~45 thread-instructions per pixel against ~185 for k_blur (profiles/ncu_r1_k_blur.json:
// 4.4 M warp-instructions per 771 K pixels).  NOT yet run on hardware: selected with MVO_BLUR2=1 only.
constexpr int BLUR2_INW = BLUR_TW / 4 + 2;      // input words per tile row: 4-px halo word left and right
constexpr int BLUR2_ROWF = BLUR_TW + 4;         // floats per row of the row-pass result (16-byte aligned rows, skewed banks)

__global__ void __launch_bounds__(256)
k_blur2(OrbPlanDev plan, BlurTiles tiles, uint8_t *__restrict__ planes) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ uint32_t s_in[BLUR_TH + 6][BLUR2_INW];
  __shared__ __align__(16) float s_row[BLUR_TH + 6][BLUR2_ROWF];
  int level = 0;
  while (level + 1 < plan.nlevels && (int)blockIdx.x >= tiles.first[level + 1]) ++level;
  const int tile = (int)blockIdx.x - tiles.first[level];
  const OrbLevelDev &L = plan.lv[level];
  const int f = blockIdx.z, tx0 = (tile % tiles.nx[level]) * BLUR_TW, ty0 = (tile / tiles.nx[level]) * BLUR_TH;
  const int w = L.w, h = L.h;
  const uint8_t *img = planes + (size_t)f * plan.slot_bytes + L.img_off;
  uint8_t *out = planes + (size_t)f * plan.slot_bytes + L.blur_off;
  const int tid = threadIdx.x;
  // ---- tile rows ty0-3 .. ty0+TH+2, columns tx0-4 .. tx0+TW+3, BORDER_REFLECT_101 on both axes ----
  for (int i = tid; i < (BLUR_TH + 6) * BLUR2_INW; i += 256) {
    const int r = i / BLUR2_INW, wc = i - r * BLUR2_INW;
    int y = ty0 + r - 3;
    y = y < 0 ? -y : (y >= h ? 2 * h - 2 - y : y);
    y = max(0, min(y, h - 1));
    const uint8_t *row = img + (size_t)y * L.pitch;
    const int x = tx0 + 4 * wc - 4;
    uint32_t v;
    if (x >= 0 && x + 3 < w) {
      v = *reinterpret_cast<const uint32_t *>(row + x);          // planes and pitches are 128-byte aligned, x is a multiple of 4
    } else {
      v = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int xx = x + j;
        xx = xx < 0 ? -xx : (xx >= w ? 2 * w - 2 - xx : xx);
        xx = max(0, min(xx, w - 1));
        v |= (uint32_t)row[xx] << (8 * j);
      }
    }
    s_in[r][wc] = v;
  }
  __syncthreads();
  // ---- row pass: outputs x = tx0 + 4q + j (j = 0..3) need inputs x-3 .. x+3 = bytes 4q+1+j .. 4q+7+j of the tile row ----
  for (int i = tid; i < (BLUR_TH + 6) * (BLUR_TW / 4); i += 256) {
    const int r = i / (BLUR_TW / 4), q = i - r * (BLUR_TW / 4);
    const uint32_t w0 = s_in[r][q], w1 = s_in[r][q + 1], w2 = s_in[r][q + 2];
    float v[10];
    v[0] = (float)((w0 >> 8) & 0xFF); v[1] = (float)((w0 >> 16) & 0xFF); v[2] = (float)(w0 >> 24);
    v[3] = (float)(w1 & 0xFF); v[4] = (float)((w1 >> 8) & 0xFF); v[5] = (float)((w1 >> 16) & 0xFF); v[6] = (float)(w1 >> 24);
    v[7] = (float)(w2 & 0xFF); v[8] = (float)((w2 >> 8) & 0xFF); v[9] = (float)((w2 >> 16) & 0xFF);
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = __fmul_rn(c_gauss7[0], v[j]);
#pragma unroll
      for (int k = 1; k < 7; ++k) s = __fadd_rn(s, __fmul_rn(c_gauss7[k], v[j + k]));
      o[j] = s;
    }
    *reinterpret_cast<float4 *>(&s_row[r][4 * q]) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  // ---- column pass: thread = 4 columns x 2 output rows (r0, r0 + 1) from row-pass rows r0 .. r0 + 7 ----
  {
    const int q = tid & 31, r0 = (tid >> 5) * 2;
    float4 win[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) win[k] = *reinterpret_cast<const float4 *>(&s_row[r0 + k][4 * q]);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int y = ty0 + r0 + rr;
      if (y >= h || tx0 + 4 * q >= L.pitch) continue;
      uint32_t pk = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float c3 = j == 0 ? win[rr + 3].x : j == 1 ? win[rr + 3].y : j == 2 ? win[rr + 3].z : win[rr + 3].w;
        float s = __fmul_rn(c_gauss7[3], c3);
#pragma unroll
        for (int k = 1; k <= 3; ++k) {
          const float4 &hi = win[rr + 3 + k], &lo = win[rr + 3 - k];
          const float a = j == 0 ? hi.x : j == 1 ? hi.y : j == 2 ? hi.z : hi.w;
          const float b = j == 0 ? lo.x : j == 1 ? lo.y : j == 2 ? lo.z : lo.w;
          s = __fadd_rn(s, __fmul_rn(c_gauss7[3 + k], __fadd_rn(a, b)));
        }
        int vq = __float2int_rn(s);
        vq = max(0, min(255, vq));
        pk |= (uint32_t)vq << (8 * j);
      }
      *reinterpret_cast<uint32_t *>(out + (size_t)y * L.pitch + tx0 + 4 * q) = pk;
    }
  }
}

// ---- experimental variant of k_describe<0> (MVO_DESCRIBE2=1; NOT yet run on hardware) ----
// Same arithmetic, fewer instructions (k_describe executes ~1556 warp-instructions per keypoint, profiles/ncu_r1_k_describe.json):
// the centroid loop is fully unrolled so that the disc half-widths are compile-time constants and the row address is a
// running 32-bit offset; the BRIEF pattern sits in shared memory as floats (no int->float conversion per sample).
__device__ __forceinline__ float ic_angle_warp2(const uint8_t *__restrict__ img, int pitch, int x0, int y0, int lane) {
  constexpr int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};      // OpenCV's ORB umax table
  const int u = lane - ORB_HALF_PATCH;                            // lanes 0..30 -> u = -15..15
  const int au = u < 0 ? -u : u;
  const uint8_t *p = img + (y0 - ORB_HALF_PATCH) * pitch + x0 + u;
  int m10 = 0, m01 = 0;
#pragma unroll
  for (int v = -ORB_HALF_PATCH; v <= ORB_HALF_PATCH; ++v) {
    const int um = umax[v < 0 ? -v : v];
    if (lane < 31 && au <= um) {
      const int val = p[(v + ORB_HALF_PATCH) * pitch];
      m10 += u * val;
      m01 += v * val;
    }
  }
  m10 = warp_sum_v(m10);
  m01 = warp_sum_v(m01);
  return fast_atan2_deg_v((float)m01, (float)m10);
}

__device__ __forceinline__ uint32_t brief_byte2(const uint8_t *__restrict__ blur, int pitch, int cx, int cy, float angle_deg,
                                                int lane, const float2 *__restrict__ s_patf) {
  const float th = __fmul_rn(angle_deg, 0.017453292519943295f);
  const float a = (float)cos((double)th), b = (float)sin((double)th);
  const uint8_t *center = blur + cy * pitch + cx;
  uint32_t byte = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 p0 = s_patf[(lane * 8 + j) * 2], p1 = s_patf[(lane * 8 + j) * 2 + 1];
    const float x0 = __fsub_rn(__fmul_rn(p0.x, a), __fmul_rn(p0.y, b));
    const float y0 = __fadd_rn(__fmul_rn(p0.x, b), __fmul_rn(p0.y, a));
    const float x1 = __fsub_rn(__fmul_rn(p1.x, a), __fmul_rn(p1.y, b));
    const float y1 = __fadd_rn(__fmul_rn(p1.x, b), __fmul_rn(p1.y, a));
    const int t0 = center[__float2int_rn(y0) * pitch + __float2int_rn(x0)];
    const int t1 = center[__float2int_rn(y1) * pitch + __float2int_rn(x1)];
    byte |= (uint32_t)(t0 < t1) << j;
  }
  return byte;
}

__global__ void __launch_bounds__(DESC_WARPS * 32)
k_describe_sel2(OrbPlanDev plan, const uint8_t *__restrict__ planes, const uint2 *__restrict__ sel, const OrbFrameMeta *__restrict__ meta,
                const int32_t *__restrict__ n_override, mvo_keypoint *__restrict__ kout, uint8_t *__restrict__ desc,
                int32_t *__restrict__ counts, int out_cap, int with_desc) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float2 s_patf[512];
  const int f = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 512; i += DESC_WARPS * 32) s_patf[i] = make_float2((float)kOrbPattern[i][0], (float)kOrbPattern[i][1]);
  __syncthreads();
  int n = n_override ? n_override[f] : meta[f].n_sel;
  n = min(n, out_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0 && counts) counts[f] = n;
  const uint8_t *slot = planes + (size_t)f * plan.slot_bytes;
  for (int k = blockIdx.x * DESC_WARPS + warp; k < n; k += gridDim.x * DESC_WARPS) {
    const uint2 s = sel[(size_t)f * (plan.max_kpts + 1) + k];
    const int l = (int)s.y, x = orb_px(s.x), y = orb_py(s.x);
    const OrbLevelDev &L = plan.lv[l];
    const uint8_t *img = slot + L.img_off;
    const float resp = harris_warp_v(img, L.pitch, x, y, lane);
    const float angle = ic_angle_warp2(img, L.pitch, x, y, lane);
    if (lane == 0) {
      mvo_keypoint kp;
      kp.x = l ? __fmul_rn((float)x, L.scale) : (float)x;
      kp.y = l ? __fmul_rn((float)y, L.scale) : (float)y;
      kp.size = __fmul_rn(31.f, L.scale);
      kp.angle = angle;
      kp.response = resp;
      kp.octave = l;
      kp.class_id = -1;
      kout[(size_t)f * out_cap + k] = kp;
    }
    if (with_desc) {
      const uint32_t byte = brief_byte2(slot + L.blur_off, L.pitch, x, y, angle, lane, s_patf);
      desc[((size_t)f * out_cap + k) * 32 + lane] = (uint8_t)byte;
    }
  }
}

