// pp2d/map_io.hpp -- map loading for the C++ host mirror, without OpenCV.
//
// Reference: MdpPathPlanning2d::loadMapFromFile / PomdpPathPlanning2d::
// loadMapFromFile (src/mdp/path_planning_2d.cu:191-205,
// src/pomdp/path_planning_2d.cu:243-257):
//     img = cv::imread(path, IMREAD_GRAYSCALE);
//     cv::threshold(img, grid, 250.0, 1.0, THRESH_BINARY_INV);   // >250 -> 0
// This header decodes 8-bit non-interlaced PNG (gray, gray+alpha, RGB, RGBA,
// palette) with zlib and converts colour to gray exactly as
// cv::imread(IMREAD_GRAYSCALE) does for PNG files: OpenCV lets libpng do it
// (png_set_rgb_to_gray, 0.299 / 0.587) which evaluates
//     gray = (9798*R + 19235*G + 3735*B) >> 15        (no rounding)
// -- checked against cv2 4.13 on RGB fixtures in tests/test_host_mirror_cpu.py;
// two of the five maps the reference bundles are RGB.
// Binary PGM (P5) is accepted too.  Link with -lz.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace pp2d {

inline bool read_file(const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  size_t got = n > 0 ? fread(out.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == out.size();
}

inline uint32_t be32(const uint8_t* p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}

inline uint8_t cv_gray(uint8_t r, uint8_t g, uint8_t b) {
  return (uint8_t)((9798u * r + 19235u * g + 3735u * b) >> 15);
}

// Decodes into 8-bit gray, row major.  Returns false on anything unsupported.
inline bool load_gray_png(const std::vector<uint8_t>& d, uint32_t& w, uint32_t& h,
                          std::vector<uint8_t>& gray) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (d.size() < 33 || memcmp(d.data(), sig, 8) != 0) return false;
  size_t pos = 8;
  int color = -1, depth = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  while (pos + 12 <= d.size()) {
    const uint32_t len = be32(&d[pos]);
    const char* type = reinterpret_cast<const char*>(&d[pos + 4]);
    const uint8_t* body = &d[pos + 8];
    if (pos + 12 + len > d.size()) return false;
    if (!memcmp(type, "IHDR", 4)) {
      w = be32(body); h = be32(body + 4);
      depth = body[8]; color = body[9]; interlace = body[12];
    } else if (!memcmp(type, "PLTE", 4)) {
      plte.assign(body, body + len);
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + len;
  }
  if (depth != 8 || interlace != 0 || w == 0 || h == 0) return false;
  int ch;
  switch (color) {
    case 0: ch = 1; break;   // gray
    case 2: ch = 3; break;   // RGB
    case 3: ch = 1; break;   // palette
    case 4: ch = 2; break;   // gray + alpha
    case 6: ch = 4; break;   // RGBA
    default: return false;
  }
  const size_t stride = (size_t)w * ch;
  std::vector<uint8_t> raw((stride + 1) * h);
  uLongf raw_len = raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), idat.size()) != Z_OK ||
      raw_len != raw.size())
    return false;
  std::vector<uint8_t> img(stride * h);
  for (uint32_t y = 0; y < h; ++y) {            // undo the PNG row filters
    const uint8_t ft = raw[(stride + 1) * y];
    const uint8_t* in = &raw[(stride + 1) * y + 1];
    uint8_t* out = &img[stride * y];
    const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= (size_t)ch ? out[i - ch] : 0;
      const int b = up ? up[i] : 0;
      const int c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
      int pred = 0;
      switch (ft) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: {
          const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
          pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
          break;
        }
        default: return false;
      }
      out[i] = (uint8_t)(in[i] + pred);
    }
  }
  gray.resize((size_t)w * h);
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    const uint8_t* px = &img[i * ch];
    switch (color) {
      case 0: case 4: gray[i] = px[0]; break;
      case 2: case 6: gray[i] = cv_gray(px[0], px[1], px[2]); break;
      case 3: {
        if ((size_t)px[0] * 3 + 2 >= plte.size()) return false;
        const uint8_t* q = &plte[px[0] * 3];
        gray[i] = cv_gray(q[0], q[1], q[2]);
        break;
      }
    }
  }
  return true;
}

inline bool load_gray_pgm(const std::vector<uint8_t>& d, uint32_t& w, uint32_t& h,
                          std::vector<uint8_t>& gray) {
  if (d.size() < 7 || d[0] != 'P' || d[1] != '5') return false;
  size_t pos = 2;
  unsigned vals[3];
  for (int k = 0; k < 3; ++k) {
    while (pos < d.size() && (isspace(d[pos]) || d[pos] == '#')) {
      if (d[pos] == '#') while (pos < d.size() && d[pos] != '\n') ++pos;
      else ++pos;
    }
    unsigned v = 0;
    while (pos < d.size() && isdigit(d[pos])) v = v * 10 + (d[pos++] - '0');
    vals[k] = v;
  }
  ++pos;
  w = vals[0]; h = vals[1];
  if (vals[2] != 255 || pos + (size_t)w * h > d.size()) return false;
  gray.assign(d.begin() + pos, d.begin() + pos + (size_t)w * h);
  return true;
}

// loadMapFromFile: gray > 250 -> 0 (free), else 1 (occupied).
inline bool load_occupancy(const std::string& path, uint32_t& w, uint32_t& h,
                           std::vector<uint8_t>& grid) {
  std::vector<uint8_t> d, gray;
  if (!read_file(path, d)) return false;
  if (!load_gray_png(d, w, h, gray) && !load_gray_pgm(d, w, h, gray)) return false;
  grid.resize(gray.size());
  for (size_t i = 0; i < gray.size(); ++i) grid[i] = gray[i] > 250 ? 0 : 1;
  return true;
}

}  // namespace pp2d
