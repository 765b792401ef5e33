// png_write.h — RGBA8 PNG writer on zlib (the reference vendors lodepng for this).
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace bdpt_host {

inline bool write_png_rgba8(const std::string& path, const uint8_t* rgba, int w, int h) {
  std::vector<uint8_t> raw((size_t)h * (1 + (size_t)w * 4));
  for (int y = 0; y < h; y++) {
    raw[(size_t)y * (1 + (size_t)w * 4)] = 0;  // filter: none
    std::copy(rgba + (size_t)y * w * 4, rgba + (size_t)(y + 1) * w * 4, raw.begin() + (size_t)y * (1 + (size_t)w * 4) + 1);
  }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  auto be32 = [](uint32_t v, uint8_t* o) { o[0] = v >> 24; o[1] = v >> 16; o[2] = v >> 8; o[3] = v; };
  auto chunk = [&](const char* type, const uint8_t* data, uint32_t len) {
    uint8_t hdr[8]; be32(len, hdr); hdr[4] = type[0]; hdr[5] = type[1]; hdr[6] = type[2]; hdr[7] = type[3];
    fwrite(hdr, 1, 8, f);
    if (len) fwrite(data, 1, len, f);
    uLong c = crc32(0L, hdr + 4, 4);
    if (len) c = crc32(c, data, len);
    uint8_t crc[4]; be32((uint32_t)c, crc); fwrite(crc, 1, 4, f);
  };
  const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  fwrite(sig, 1, 8, f);
  uint8_t ihdr[13]; be32((uint32_t)w, ihdr); be32((uint32_t)h, ihdr + 4);
  ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", comp.data(), (uint32_t)clen);
  chunk("IEND", nullptr, 0);
  return fclose(f) == 0;
}

}  // namespace bdpt_host
