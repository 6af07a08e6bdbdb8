// nvpyr_io.cpp -- image files on either side of the path (host only, no CUDA).
//
//   * writing: the reference dumps every level of a MipmapStorage as `name.tga`, `name.1.tga`, ...
//     through stbi_write_tga (include/mipmap_storage.hpp:441-479).  stb_image_write is a third-party
//     dependency that is NOT part of /root/reference (nvpro_core bundles it); its published TGA writer
//     (stb_image_write v1.16, stbi_write_tga_core with stbi_write_tga_with_rle = 1, the default) is
//     restated here: 18-byte header "111 221 2222 11" = {0, 0, 10, 0, 0, 0, 0, 0, w, h, 32, 8}, rows
//     bottom-up, texels B,G,R,A, and its greedy run-length packets (runs and literal packets of at
//     most 128 texels, a literal packet ends one texel before a repeat starts).  Byte-for-byte parity
//     with stb is UNPINNED (no stb here to run); the decoded pixels and header are pinned by reading the
//     files back with our reader and with PIL (tests/test_io.py).
//   * reading: the reference loads its inputs with stbi_load(..., 4) (include/scoped_image.hpp:217-218).
//     No JPEG/PNG decoder is built here; the reader takes TGA (true colour / grey, raw or RLE, 8/24/32
//     bits) and binary PPM/PGM, and returns top-down R,G,B,A with A = 255 when the file has none -- the
//     same convention as stbi_load with req_comp = 4.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/nvpyr.h"
#include "nvpyr_plan.hpp"

namespace {

struct File
{
  FILE* f;
  explicit File(const char* name, const char* mode) : f(fopen(name, mode)) {}
  ~File()
  {
    if(f)
      fclose(f);
  }
};

// ------------------------------------------------------------------ TGA out
void put16(std::vector<unsigned char>& o, uint32_t v)
{
  o.push_back(uint8_t(v & 0xFFu));
  o.push_back(uint8_t((v >> 8) & 0xFFu));
}
void putTexel(std::vector<unsigned char>& o, const unsigned char* t)
{
  o.push_back(t[2]), o.push_back(t[1]), o.push_back(t[0]), o.push_back(t[3]);  // B, G, R, A
}

// One level as an RLE TGA, the way stbi_write_tga(filename, w, h, 4, data) packs it.
void encodeTga(const unsigned char* rgba, uint32_t w, uint32_t h, std::vector<unsigned char>& o)
{
  o.clear();
  o.reserve(size_t(w) * h * 4u + 18u);
  o.push_back(0), o.push_back(0), o.push_back(10);  // no id, no colour map, RLE true colour
  put16(o, 0), put16(o, 0), o.push_back(0);         // colour-map spec
  put16(o, 0), put16(o, 0), put16(o, w), put16(o, h);
  o.push_back(32), o.push_back(8);  // 32 bits per texel, 8 alpha bits, origin bottom-left
  for(uint32_t j = h; j-- > 0;)
  {
    const unsigned char* row = rgba + size_t(j) * w * 4u;
    uint32_t             len = 0;
    for(uint32_t i = 0; i < w; i += len)
    {
      const unsigned char* begin = row + size_t(i) * 4u;
      bool                 diff  = true;
      len                        = 1;
      if(i + 1 < w)
      {
        ++len;
        diff = memcmp(begin, row + size_t(i + 1) * 4u, 4) != 0;
        if(diff)
        {
          const unsigned char* prev = begin;
          for(uint32_t k = i + 2; k < w && len < 128; ++k)
          {
            if(memcmp(prev, row + size_t(k) * 4u, 4))
            {
              prev += 4;
              ++len;
            }
            else
            {
              --len;
              break;
            }
          }
        }
        else
        {
          for(uint32_t k = i + 2; k < w && len < 128; ++k)
          {
            if(memcmp(begin, row + size_t(k) * 4u, 4))
              break;
            ++len;
          }
        }
      }
      if(diff)
      {
        o.push_back(uint8_t(len - 1));
        for(uint32_t k = 0; k < len; ++k)
          putTexel(o, begin + size_t(k) * 4u);
      }
      else
      {
        o.push_back(uint8_t(len - 129));  // 0x80 | (len - 1)
        putTexel(o, begin);
      }
    }
  }
}

nvpyrStatus writeTga(const char* filename, const unsigned char* rgba, uint32_t w, uint32_t h)
{
  if(w > 0xFFFFu || h > 0xFFFFu)
    return NVPYR_ERROR_INVALID_VALUE;  // 16-bit header fields
  std::vector<unsigned char> bytes;
  encodeTga(rgba, w, h, bytes);
  File out(filename, "wb");
  if(!out.f || fwrite(bytes.data(), 1, bytes.size(), out.f) != bytes.size())
    return NVPYR_ERROR_IO;
  return NVPYR_SUCCESS;
}

// image.name.tga -> image.name.<level>.tga; level 0 keeps the base name (mipmap_storage.hpp:447-462).
std::string levelFilename(const char* base, uint32_t level)
{
  if(level == 0)
    return base;
  const char* dot = strrchr(base, '.');
  std::string s   = dot ? std::string(base, size_t(dot - base) + 1) : std::string(base);
  s += std::to_string(level);
  if(dot)
    s += dot;
  return s;
}

// ------------------------------------------------------------------- readers
bool readAll(const char* filename, std::vector<unsigned char>& bytes)
{
  File in(filename, "rb");
  if(!in.f || fseek(in.f, 0, SEEK_END) != 0)
    return false;
  const long n = ftell(in.f);
  if(n < 0 || fseek(in.f, 0, SEEK_SET) != 0)
    return false;
  bytes.resize(size_t(n));
  return fread(bytes.data(), 1, bytes.size(), in.f) == bytes.size();
}

nvpyrStatus decodeTga(const std::vector<unsigned char>& b, unsigned char** out, nvpyrExtent2D* extent)
{
  if(b.size() < 18)
    return NVPYR_ERROR_IO;
  const uint32_t idLen = b[0], cmapType = b[1], type = b[2];
  const uint32_t w = b[12] | (b[13] << 8), h = b[14] | (b[15] << 8), bpp = b[16], desc = b[17];
  const bool     rle = type == 10 || type == 11, grey = type == 3 || type == 11;
  if(cmapType != 0 || !(type == 2 || type == 3 || rle) || w == 0 || h == 0)
    return NVPYR_ERROR_UNSUPPORTED;
  if(!((grey && bpp == 8) || (!grey && (bpp == 24 || bpp == 32))))
    return NVPYR_ERROR_UNSUPPORTED;
  const uint32_t bytesPer = bpp / 8;
  size_t         pos      = 18 + size_t(idLen);
  unsigned char* img      = static_cast<unsigned char*>(malloc(size_t(w) * h * 4u));
  if(!img)
    return NVPYR_ERROR_OUT_OF_MEMORY;
  const bool topDown = desc & 0x20u, rightToLeft = desc & 0x10u;
  auto       store   = [&](uint64_t index, const unsigned char* t) {
    uint32_t x = uint32_t(index % w), y = uint32_t(index / w);
    if(!topDown)
      y = h - 1 - y;
    if(rightToLeft)
      x = w - 1 - x;
    unsigned char* d = img + (size_t(y) * w + x) * 4u;
    if(grey)
      d[0] = d[1] = d[2] = t[0], d[3] = 255;
    else
      d[0] = t[2], d[1] = t[1], d[2] = t[0], d[3] = bytesPer == 4 ? t[3] : 255;
  };
  const uint64_t total = uint64_t(w) * h;
  uint64_t       n     = 0;
  bool           ok    = true;
  if(!rle)
  {
    ok = b.size() >= pos + total * bytesPer;
    for(; ok && n < total; ++n, pos += bytesPer)
      store(n, &b[pos]);
  }
  else
  {
    while(ok && n < total)
    {
      if(pos >= b.size())
      {
        ok = false;
        break;
      }
      const uint32_t head = b[pos++], count = (head & 0x7Fu) + 1u;
      if(n + count > total)
      {
        ok = false;
        break;
      }
      if(head & 0x80u)
      {
        if(pos + bytesPer > b.size())
        {
          ok = false;
          break;
        }
        for(uint32_t k = 0; k < count; ++k)
          store(n++, &b[pos]);
        pos += bytesPer;
      }
      else
      {
        if(pos + size_t(count) * bytesPer > b.size())
        {
          ok = false;
          break;
        }
        for(uint32_t k = 0; k < count; ++k, pos += bytesPer)
          store(n++, &b[pos]);
      }
    }
  }
  if(!ok)
  {
    free(img);
    return NVPYR_ERROR_IO;
  }
  *out    = img;
  *extent = nvpyrExtent2D{w, h};
  return NVPYR_SUCCESS;
}

// Binary PGM (P5) / PPM (P6), maxval <= 255.
nvpyrStatus decodePnm(const std::vector<unsigned char>& b, unsigned char** out, nvpyrExtent2D* extent)
{
  size_t pos   = 2;
  auto   token = [&](uint32_t& v) {
    for(;;)
    {
      while(pos < b.size() && (b[pos] == ' ' || b[pos] == '\t' || b[pos] == '\n' || b[pos] == '\r'))
        ++pos;
      if(pos < b.size() && b[pos] == '#')
        while(pos < b.size() && b[pos] != '\n')
          ++pos;
      else
        break;
    }
    if(pos >= b.size() || b[pos] < '0' || b[pos] > '9')
      return false;
    uint64_t x = 0;
    while(pos < b.size() && b[pos] >= '0' && b[pos] <= '9' && x < (1ull << 32))
      x = x * 10 + uint64_t(b[pos++] - '0');
    v = uint32_t(x);
    return x < (1ull << 32);
  };
  const uint32_t comps = b[1] == '6' ? 3u : 1u;
  uint32_t       w = 0, h = 0, maxval = 0;
  if(!token(w) || !token(h) || !token(maxval) || w == 0 || h == 0 || maxval == 0 || maxval > 255 || pos >= b.size())
    return NVPYR_ERROR_UNSUPPORTED;
  ++pos;  // the single whitespace byte after maxval
  const uint64_t total = uint64_t(w) * h;
  if(b.size() < pos + total * comps)
    return NVPYR_ERROR_IO;
  unsigned char* img = static_cast<unsigned char*>(malloc(size_t(total) * 4u));
  if(!img)
    return NVPYR_ERROR_OUT_OF_MEMORY;
  for(uint64_t i = 0; i < total; ++i)
  {
    const unsigned char* s = &b[pos + i * comps];
    unsigned char*       d = img + i * 4u;
    d[0] = s[0], d[1] = s[comps == 3 ? 1 : 0], d[2] = s[comps == 3 ? 2 : 0], d[3] = 255;
  }
  *out    = img;
  *extent = nvpyrExtent2D{w, h};
  return NVPYR_SUCCESS;
}

}  // namespace

extern "C" {

nvpyrStatus nvpyrWriteTga(const char* filename, const void* rgba8, nvpyrExtent2D extent)
{
  if(filename == nullptr || rgba8 == nullptr || extent.width == 0 || extent.height == 0)
    return NVPYR_ERROR_INVALID_VALUE;
  return writeTga(filename, static_cast<const unsigned char*>(rgba8), extent.width, extent.height);
}

nvpyrStatus nvpyrGetLevelFilename(const char* baseFilename, uint32_t level, char* out, size_t outSize)
{
  if(baseFilename == nullptr || out == nullptr)
    return NVPYR_ERROR_INVALID_VALUE;
  const std::string s = levelFilename(baseFilename, level);
  if(s.size() + 1 > outSize)
    return NVPYR_ERROR_INVALID_VALUE;
  memcpy(out, s.c_str(), s.size() + 1);
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrWriteChainTga(const void* hostChain, nvpyrExtent2D extent, uint32_t levelCount, const char* baseFilename)
{
  if(hostChain == nullptr || baseFilename == nullptr || extent.width == 0 || extent.height == 0)
    return NVPYR_ERROR_INVALID_VALUE;
  const uint32_t maxLevels = nvpyr::levelCountFor(extent.width, extent.height);
  if(levelCount == 0)
    levelCount = maxLevels;
  if(levelCount > maxLevels)
    return NVPYR_ERROR_INVALID_VALUE;
  std::vector<nvpyrStatus> status(levelCount, NVPYR_SUCCESS);
  auto                     one = [&](uint32_t level) {
    const unsigned char* data = static_cast<const unsigned char*>(hostChain)
                                + nvpyr::levelOffsetTexels(extent.width, extent.height, level) * 4u;
    status[level] = writeTga(levelFilename(baseFilename, level).c_str(), data, nvpyr::levelDim(extent.width, level),
                             nvpyr::levelDim(extent.height, level));
  };
  // Like the reference: the big levels on one thread each, the last eight in the caller (mipmap_storage.hpp:464-478).
  std::vector<std::thread> threads;
  const uint32_t           parallel = levelCount > 8u ? levelCount - 8u : 0u;
  for(uint32_t level = 0; level < parallel; ++level)
    threads.emplace_back(one, level);
  for(uint32_t level = parallel; level < levelCount; ++level)
    one(level);
  for(std::thread& t : threads)
    t.join();
  for(nvpyrStatus st : status)
    if(st != NVPYR_SUCCESS)
      return st;
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrReadImage(const char* filename, void** rgba8, nvpyrExtent2D* extent)
{
  if(filename == nullptr || rgba8 == nullptr || extent == nullptr)
    return NVPYR_ERROR_INVALID_VALUE;
  std::vector<unsigned char> bytes;
  if(!readAll(filename, bytes))
    return NVPYR_ERROR_IO;
  unsigned char* img = nullptr;
  nvpyrStatus    st;
  if(bytes.size() >= 2 && bytes[0] == 'P' && (bytes[1] == '5' || bytes[1] == '6'))
    st = decodePnm(bytes, &img, extent);
  else
    st = decodeTga(bytes, &img, extent);  // TGA has no magic number: everything else is tried as TGA
  if(st == NVPYR_SUCCESS)
    *rgba8 = img;
  return st;
}

void nvpyrFree(void* p)
{
  free(p);
}

}  // extern "C"
