// Minimal PNG decoder for the run_vo application: what cv::imread(path) (IMREAD_COLOR, reference run_vo.cpp:114) yields
// for the PNG files of the reference's datasets — an 8-bit, 3-channel BGR image.  Supports non-interlaced PNG of colour
// type 0 (gray), 2 (RGB), 3 (palette), 4 (gray + alpha), 6 (RGBA) at 8 or 16 bits per sample (16-bit samples keep their
// high byte, alpha is dropped, as IMREAD_COLOR does).  zlib does the inflate.  Not part of libmvo.so.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mvo_app {
// Returns false (with a message in *err) if the file is missing, not a PNG this reader supports, or corrupt.
bool read_png_bgr(const std::string &path, std::vector<uint8_t> *bgr, int *rows, int *cols, std::string *err);
}
