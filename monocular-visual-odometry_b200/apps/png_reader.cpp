#include "png_reader.h"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace mvo_app {
namespace {

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool fail(std::string *err, const std::string &msg) {
  if (err) *err = msg;
  return false;
}

}  // namespace

bool read_png_bgr(const std::string &path, std::vector<uint8_t> *bgr, int *rows, int *cols, std::string *err) {
  FILE *fp = fopen(path.c_str(), "rb");
  if (!fp) return fail(err, "cannot open " + path);
  std::vector<uint8_t> file;
  uint8_t buf[65536];
  size_t got;
  while ((got = fread(buf, 1, sizeof buf, fp)) > 0) file.insert(file.end(), buf, buf + got);
  fclose(fp);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8) != 0) return fail(err, path + ": not a PNG file");
  uint32_t w = 0, h = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat, palette;
  size_t pos = 8;
  bool seen_iend = false;
  while (pos + 12 <= file.size()) {
    const uint32_t len = be32(&file[pos]);
    const uint8_t *type = &file[pos + 4], *data = &file[pos + 8];
    if (pos + 12 + (size_t)len > file.size()) return fail(err, path + ": truncated chunk");
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) return fail(err, path + ": bad IHDR");
      w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
    } else if (!memcmp(type, "PLTE", 4)) {
      palette.assign(data, data + len);
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!memcmp(type, "IEND", 4)) {
      seen_iend = true;
      break;
    }
    pos += 12 + (size_t)len;
  }
  if (!seen_iend || ctype < 0) return fail(err, path + ": missing IHDR / IEND");
  if (w == 0 || h == 0 || w > 32768 || h > 32768 || (uint64_t)w * h > (64ull << 20)) return fail(err, path + ": unsupported image size");   // <= 64 Mpixel
  if (interlace) return fail(err, path + ": interlaced PNG is not supported");
  if (!((depth == 8 || depth == 16) && (ctype == 0 || ctype == 2 || ctype == 4 || ctype == 6)) && !(ctype == 3 && depth == 8))
    return fail(err, path + ": unsupported colour type / bit depth");
  const int samples = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
  const int bps = depth / 8, bpp = samples * bps;                  // bytes per pixel (filter distance)
  const size_t stride = (size_t)w * bpp;
  std::vector<uint8_t> raw((stride + 1) * h);
  uLongf out_len = (uLongf)raw.size();
  const int zrc = uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size());
  if (zrc != Z_OK || out_len != raw.size()) return fail(err, path + ": corrupt image data");
  // ---- undo the scanline filters (PNG specification, section 9) ----
  std::vector<uint8_t> img(stride * h);
  for (uint32_t y = 0; y < h; ++y) {
    const uint8_t ft = raw[(stride + 1) * y];
    const uint8_t *in = &raw[(stride + 1) * y + 1];
    uint8_t *cur = &img[stride * y];
    const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
      int v;
      switch (ft) {
        case 0: v = in[i]; break;
        case 1: v = in[i] + a; break;
        case 2: v = in[i] + b; break;
        case 3: v = in[i] + ((a + b) >> 1); break;
        case 4: v = in[i] + paeth(a, b, c); break;
        default: return fail(err, path + ": bad filter type");
      }
      cur[i] = (uint8_t)v;
    }
  }
  // ---- to 8-bit BGR ----
  bgr->assign((size_t)w * h * 3, 0);
  for (size_t p = 0; p < (size_t)w * h; ++p) {
    const uint8_t *s = &img[p * bpp];
    uint8_t r, g, b;
    if (ctype == 0 || ctype == 4) { r = g = b = s[0]; }           // 16-bit: s[0] is the high byte (big endian)
    else if (ctype == 3) {
      const size_t k = (size_t)s[0] * 3;
      if (k + 3 > palette.size()) return fail(err, path + ": palette index out of range");
      r = palette[k]; g = palette[k + 1]; b = palette[k + 2];
    } else { r = s[0]; g = s[bps]; b = s[2 * bps]; }
    (*bgr)[3 * p] = b; (*bgr)[3 * p + 1] = g; (*bgr)[3 * p + 2] = r;
  }
  *rows = (int)h;
  *cols = (int)w;
  return true;
}

}  // namespace mvo_app

extern "C" int mvo_app_read_png_bgr(const char *path, uint8_t *out, int cap, int *rows, int *cols) {     // test hook (ctypes)
  std::vector<uint8_t> bgr;
  std::string err;
  if (!mvo_app::read_png_bgr(path, &bgr, rows, cols, &err)) return -1;
  if ((int)bgr.size() > cap) return -4;
  memcpy(out, bgr.data(), bgr.size());
  return 0;
}
