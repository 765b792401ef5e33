// envmap_read.h — lat-long environment maps for `-e <file>` (reference src/application/main.cpp:40-78,115-118, which
// reads OpenEXR through the vendored tinyexr).  Own readers, no third-party code:
//   .exr  single-part scan-line OpenEXR, HALF or FLOAT channels named R, G, B, compression NONE / RLE / ZIPS / ZIP
//         (the PIZ / PXR24 / B44 / DWA coders and tiled files are reported as unsupported)
//   .pfm  portable float map ("PF": RGB, bottom row first)
//   r,g,b three numbers: a constant environment (1 x 2 map)
// Rows come out top first (row 0 = straight up), RGB floats.
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace bdpt_host {

struct EnvMap { int w = 0, h = 0; std::vector<float> rgb; };

inline float half_to_float(uint16_t hbits) {
  const uint32_t s = (hbits >> 15) & 1u, e = (hbits >> 10) & 0x1fu, m = hbits & 0x3ffu;
  uint32_t f;
  if (e == 0) {
    if (m == 0) f = s << 31;
    else { int sh = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; sh++; } f = (s << 31) | ((uint32_t)(113 - sh) << 23) | ((mm & 0x3ffu) << 13); }
  } else if (e == 31) f = (s << 31) | 0x7f800000u | (m << 13);
  else f = (s << 31) | ((e + 112u) << 23) | (m << 13);
  float out; memcpy(&out, &f, 4);
  return out;
}

inline bool read_file(const std::string& path, std::vector<unsigned char>* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  out->resize(n > 0 ? (size_t)n : 0);
  bool ok = n >= 0 && fread(out->data(), 1, out->size(), f) == out->size();
  fclose(f);
  return ok;
}

inline bool read_exr(const std::string& path, EnvMap* out, std::string* err) {
  std::vector<unsigned char> b;
  if (!read_file(path, &b)) { *err = "cannot read " + path; return false; }
  size_t p = 0;
  auto need = [&](size_t n) { return p + n <= b.size(); };
  auto u32 = [&](size_t at) { uint32_t v; memcpy(&v, &b[at], 4); return v; };
  if (!need(8) || u32(0) != 20000630u) { *err = path + " is not an OpenEXR file"; return false; }
  const uint32_t version = u32(4);
  if ((version & 0xff) != 2 || (version & 0x1a00u)) { *err = "only single-part scan-line OpenEXR files are supported"; return false; }
  p = 8;
  struct Chan { std::string name; int type; };
  std::vector<Chan> chans;
  int compression = -1, xmin = 0, ymin = 0, xmax = -1, ymax = -1, line_order = 0;
  for (;;) {
    if (!need(1)) { *err = "truncated OpenEXR header"; return false; }
    if (b[p] == 0) { p++; break; }
    std::string name((const char*)&b[p]); p += name.size() + 1;
    if (!need(1)) { *err = "truncated OpenEXR header"; return false; }
    std::string type((const char*)&b[p]); p += type.size() + 1;
    if (!need(4)) { *err = "truncated OpenEXR header"; return false; }
    const uint32_t size = u32(p); p += 4;
    if (!need(size)) { *err = "truncated OpenEXR header"; return false; }
    if (name == "channels") {
      size_t q = p;
      while (q < p + size && b[q] != 0) {
        Chan c; c.name = (const char*)&b[q]; q += c.name.size() + 1;
        c.type = (int)u32(q); q += 16;    // type, pLinear + 3 reserved, xSampling, ySampling
        chans.push_back(c);
      }
    } else if (name == "compression") compression = b[p];
    else if (name == "dataWindow") { xmin = (int)u32(p); ymin = (int)u32(p + 4); xmax = (int)u32(p + 8); ymax = (int)u32(p + 12); }
    else if (name == "lineOrder") line_order = b[p];
    p += size;
  }
  const int w = xmax - xmin + 1, h = ymax - ymin + 1;
  if (w <= 0 || h <= 0 || chans.empty()) { *err = "OpenEXR header lacks channels / dataWindow"; return false; }
  if (compression < 0 || compression > 3) { *err = "OpenEXR compression " + std::to_string(compression) + " is not supported (NONE, RLE, ZIPS, ZIP are)"; return false; }
  int ci[3] = {-1, -1, -1};
  size_t line_bytes = 0;
  std::vector<size_t> chan_off(chans.size());
  for (size_t c = 0; c < chans.size(); c++) {
    chan_off[c] = line_bytes;
    if (chans[c].type != 1 && chans[c].type != 2) { *err = "OpenEXR UINT channels are not supported"; return false; }
    line_bytes += (size_t)w * (chans[c].type == 1 ? 2 : 4);
    if (chans[c].name == "R") ci[0] = (int)c; else if (chans[c].name == "G") ci[1] = (int)c; else if (chans[c].name == "B") ci[2] = (int)c;
  }
  if (ci[0] < 0 || ci[1] < 0 || ci[2] < 0) { *err = "OpenEXR file has no R, G, B channels"; return false; }
  const int lines_per_block = compression == 3 ? 16 : 1;
  const int nblocks = (h + lines_per_block - 1) / lines_per_block;
  if (!need((size_t)nblocks * 8)) { *err = "truncated OpenEXR offset table"; return false; }
  out->w = w; out->h = h; out->rgb.assign((size_t)w * h * 3, 0.f);
  std::vector<unsigned char> raw, tmp;
  for (int k = 0; k < nblocks; k++) {
    uint64_t at; memcpy(&at, &b[p + 8 * (size_t)k], 8);
    if (at + 8 > b.size()) { *err = "bad OpenEXR block offset"; return false; }
    const int y0 = (int)u32(at) - ymin; const uint32_t dsize = u32(at + 4);
    if (y0 < 0 || y0 >= h || at + 8 + dsize > b.size()) { *err = "bad OpenEXR block"; return false; }
    const int nl = std::min(lines_per_block, h - y0);
    const size_t want = line_bytes * nl;
    const unsigned char* src = &b[at + 8];
    raw.resize(want);
    if (compression == 0 || dsize == want) memcpy(raw.data(), src, std::min<size_t>(want, dsize));
    else {
      tmp.resize(want);
      if (compression == 1) {   // RLE
        size_t o = 0, i = 0;
        while (i < dsize && o < want) {
          int c = (signed char)src[i++];
          if (c < 0) { size_t n = (size_t)(-c); if (i + n > dsize || o + n > want) break; memcpy(&tmp[o], &src[i], n); o += n; i += n; }
          else { size_t n = (size_t)c + 1; if (i >= dsize || o + n > want) break; memset(&tmp[o], src[i++], n); o += n; }
        }
        if (o != want) { *err = "corrupt RLE block in OpenEXR file"; return false; }
      } else {
        uLongf n = (uLongf)want;
        if (uncompress(tmp.data(), &n, src, dsize) != Z_OK || n != want) { *err = "corrupt ZIP block in OpenEXR file"; return false; }
      }
      for (size_t i = 1; i < want; i++) tmp[i] = (unsigned char)(tmp[i - 1] + tmp[i] - 128);   // undo the predictor
      const size_t half = (want + 1) / 2;                                                       // undo the interleave
      for (size_t i = 0; i < want; i++) raw[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
    }
    for (int l = 0; l < nl; l++) {
      const int y = line_order == 1 ? y0 + l : y0 + l;
      const unsigned char* line = &raw[line_bytes * l];
      for (int c = 0; c < 3; c++) {
        const Chan& ch = chans[ci[c]];
        const unsigned char* d = line + chan_off[ci[c]];
        for (int x = 0; x < w; x++) {
          float v;
          if (ch.type == 1) { uint16_t hb; memcpy(&hb, d + 2 * x, 2); v = half_to_float(hb); }
          else memcpy(&v, d + 4 * x, 4);
          out->rgb[3 * ((size_t)y * w + x) + c] = v;
        }
      }
    }
  }
  return true;
}

inline bool read_pfm(const std::string& path, EnvMap* out, std::string* err) {
  std::vector<unsigned char> b;
  if (!read_file(path, &b)) { *err = "cannot read " + path; return false; }
  int w = 0, h = 0, used = 0; float scale = 0; char magic[3] = {0};
  if (sscanf((const char*)b.data(), "%2s %d %d %f%n", magic, &w, &h, &scale, &used) != 4 || strcmp(magic, "PF") || w <= 0 || h <= 0) { *err = path + " is not an RGB .pfm file"; return false; }
  size_t p = (size_t)used + 1;
  if (p + (size_t)w * h * 12 > b.size()) { *err = "truncated .pfm file"; return false; }
  out->w = w; out->h = h; out->rgb.resize((size_t)w * h * 3);
  const bool swap = scale > 0;   // positive scale: big-endian data
  for (int y = 0; y < h; y++)
    for (int i = 0; i < 3 * w; i++) {
      unsigned char q[4]; memcpy(q, &b[p + 4 * ((size_t)y * 3 * w + i)], 4);
      if (swap) { std::swap(q[0], q[3]); std::swap(q[1], q[2]); }
      float v; memcpy(&v, q, 4);
      out->rgb[(size_t)(h - 1 - y) * 3 * w + i] = v * std::fabs(scale);   // .pfm stores the bottom row first
    }
  return true;
}

inline bool read_envmap(const std::string& spec, EnvMap* out, std::string* err) {
  float r, g, bl; char tail;
  if (sscanf(spec.c_str(), "%f,%f,%f%c", &r, &g, &bl, &tail) == 3) { out->w = 1; out->h = 2; out->rgb = {r, g, bl, r, g, bl}; return true; }
  if (spec.size() > 4 && spec.substr(spec.size() - 4) == ".pfm") return read_pfm(spec, out, err);
  return read_exr(spec, out, err);
}

}  // namespace bdpt_host
