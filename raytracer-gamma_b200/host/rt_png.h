/* rt_png.h — minimal PNG writer for the host program (8-bit RGB, zlib "stored" blocks: no
 * compression, no dependency).  SURVEY.md 8f row 1 (PPM/PNG encode); the reference only writes
 * PPM (main.cpp:43-91).  Header-only so the tests can compile it on its own. */
#ifndef RT_PNG_H
#define RT_PNG_H

#include <stdint.h>
#include <stdio.h>
#include <vector>

namespace rtpng {

inline uint32_t crc32(const unsigned char* p, size_t n, uint32_t crc = 0) {
  static uint32_t table[256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    ready = true;
  }
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
  return ~crc;
}

inline void be32(std::vector<unsigned char>& v, uint32_t x) {
  v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16));
  v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}

inline bool chunk(FILE* f, const char type[4], const std::vector<unsigned char>& data) {
  std::vector<unsigned char> buf;
  buf.reserve(data.size() + 12);
  be32(buf, (uint32_t)data.size());
  buf.insert(buf.end(), type, type + 4);
  buf.insert(buf.end(), data.begin(), data.end());
  be32(buf, crc32(buf.data() + 4, buf.size() - 4));
  return fwrite(buf.data(), 1, buf.size(), f) == buf.size();
}

/* rgb: height rows of width*3 bytes, top row first (the layout rt_cuda_readback_rgb8 returns) */
inline bool write_rgb8(const char* filename, const unsigned char* rgb, unsigned width, unsigned height) {
  if (!filename || !rgb || width == 0 || height == 0) return false;
  FILE* f = fopen(filename, "wb");
  if (!f) return false;
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  bool ok = fwrite(sig, 1, 8, f) == 8;
  std::vector<unsigned char> ihdr;
  be32(ihdr, width); be32(ihdr, height);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   /* 8-bit, RGB, deflate, no filter set, no interlace */
  ok = ok && chunk(f, "IHDR", ihdr);
  /* the raw scanlines: filter byte 0 + the row */
  const size_t row = (size_t)width * 3 + 1, total = row * height;
  std::vector<unsigned char> raw(total);
  for (unsigned y = 0; y < height; ++y) {
    raw[y * row] = 0;
    for (size_t i = 0; i < (size_t)width * 3; ++i) raw[y * row + 1 + i] = rgb[(size_t)y * width * 3 + i];
  }
  /* zlib stream of stored blocks (at most 65535 bytes each) + Adler-32 */
  std::vector<unsigned char> z;
  z.reserve(total + total / 65535 * 5 + 16);
  z.push_back(0x78); z.push_back(0x01);
  uint32_t a = 1, b = 0;
  for (size_t off = 0; off < total;) {
    const size_t n = (total - off < 65535) ? total - off : 65535;
    z.push_back(off + n == total ? 1 : 0);
    z.push_back((unsigned char)(n & 0xFF)); z.push_back((unsigned char)(n >> 8));
    z.push_back((unsigned char)(~n & 0xFF)); z.push_back((unsigned char)((~n >> 8) & 0xFF));
    for (size_t i = 0; i < n; ++i) {
      const unsigned char c = raw[off + i];
      z.push_back(c);
      a += c; if (a >= 65521u) a -= 65521u;
      b += a; if (b >= 65521u) b -= 65521u;
    }
    off += n;
  }
  be32(z, (b << 16) | a);
  ok = ok && chunk(f, "IDAT", z);
  ok = ok && chunk(f, "IEND", std::vector<unsigned char>());
  return (fclose(f) == 0) && ok;
}

}  // namespace rtpng
#endif
