// ref_harness.cpp -- compiles the REFERENCE's own headers, unmodified and in
// place under $(REF) (= /root/reference), into oracle/_ref/libnvpyr_ref.so:
//   * nvpro_pyramid/nvpro_pyramid_dispatch.hpp  (host scheduler) against the
//     mock Vulkan recorder of ref_shims/vulkan/vulkan_core.h
//   * include/mipmap_storage.hpp + shaders/srgb.h (CPU generator, comparator,
//     transfer functions) against ref_shims/glm and ref_shims/stb
// TEST INFRASTRUCTURE ONLY: used to pin oracle/nvpyr_oracle.c and as the
// "reference" CPU baseline of bench.py.  No reference source is copied.
#include <stdio.h>
#include <string.h>
#include <vector>

#include "nvpro_pyramid/nvpro_pyramid_dispatch.hpp"
#include "include/mipmap_storage.hpp"

// ---------------------------------------------------------------- mock Vulkan
namespace {
struct Event
{
  int      type;  // 0 bind, 1 push, 2 dispatch, 3 barrier
  uint64_t a;
  uint32_t b;
};
thread_local std::vector<Event> g_log;
MockPipeline_T*                 kGeneral = reinterpret_cast<MockPipeline_T*>(0x1000);
MockPipeline_T*                 kFast    = reinterpret_cast<MockPipeline_T*>(0x2000);
}  // namespace

void vkCmdBindPipeline(VkCommandBuffer, VkPipelineBindPoint, VkPipeline p)
{
  g_log.push_back({0, reinterpret_cast<uint64_t>(p), 0});
}
void vkCmdPushConstants(VkCommandBuffer, VkPipelineLayout, VkShaderStageFlags, uint32_t offset, uint32_t size,
                        const void* pValues)
{
  uint32_t v = 0;
  memcpy(&v, pValues, size < 4 ? size : 4);
  g_log.push_back({1, v, offset});
}
void vkCmdDispatch(VkCommandBuffer, uint32_t x, uint32_t y, uint32_t z)
{
  g_log.push_back({2, x, y << 16 | z});
}
void vkCmdPipelineBarrier(VkCommandBuffer, VkPipelineStageFlags s, VkPipelineStageFlags d, VkDependencyFlags,
                          uint32_t n, const VkMemoryBarrier* b, uint32_t, const VkBufferMemoryBarrier*, uint32_t,
                          const VkImageMemoryBarrier*)
{
  g_log.push_back({3, (uint64_t)s << 32 | d, n ? (b->srcAccessMask << 16 | b->dstAccessMask) : 0});
}

extern "C" {

// Same field order as nvo_step in nvpyr_oracle.c.
struct ref_step
{
  uint32_t pipeline, input_level, level_count, src_w, src_h, workgroups, push_constant, bind, barrier_after;
};

// Runs the reference's nvproCmdPyramidDispatch and returns the recorded steps.
int ref_record_dispatch(uint32_t w, uint32_t h, uint32_t mipLevels, uint32_t haveFast, ref_step* out,
                        uint32_t maxSteps)
{
  g_log.clear();
  NvproPyramidPipelines p;
  p.generalPipeline = kGeneral;
  p.fastPipeline    = haveFast ? kFast : VK_NULL_HANDLE;
  nvproCmdPyramidDispatch(nullptr, p, w, h, mipLevels);
  uint32_t n = 0, pc = 0, bound = 0xFFFFFFFFu, pendingBind = 0;
  for(const Event& e : g_log)
  {
    switch(e.type)
    {
      case 0:
        bound       = e.a == reinterpret_cast<uint64_t>(kFast) ? 1u : 0u;
        pendingBind = 1;
        break;
      case 1: pc = (uint32_t)e.a; break;
      case 2:
        if(n >= maxSteps)
          return -1;
        out[n]               = ref_step{};
        out[n].pipeline      = bound;
        out[n].input_level   = pc >> 5;
        out[n].level_count   = pc & 31u;
        out[n].workgroups    = (uint32_t)e.a;
        out[n].push_constant = pc;
        out[n].bind          = pendingBind;
        if(e.b != (1u << 16 | 1u))
          return -2;  // y,z must be 1
        pendingBind = 0;
        ++n;
        break;
      case 3:
        if(n == 0)
          return -3;
        // COMPUTE->COMPUTE, SHADER_WRITE->SHADER_READ
        if(e.a != ((uint64_t)VK_PIPELINE_STAGE_COMPUTE_SHADER_BIT << 32 | VK_PIPELINE_STAGE_COMPUTE_SHADER_BIT)
           || e.b != (VK_ACCESS_SHADER_WRITE_BIT << 16 | VK_ACCESS_SHADER_READ_BIT))
          return -4;
        out[n - 1].barrier_after = 1;
        break;
    }
  }
  return (int)n;
}

// The template knobs the reference exposes: <DivisibilityRequirement, MaxLevels>.
int ref_record_dispatch_variant(uint32_t w, uint32_t h, uint32_t mipLevels, uint32_t div, uint32_t maxLevels,
                                ref_step* out, uint32_t maxSteps)
{
  nvpro_pyramid_dispatcher_t fast = nullptr;
  if(div == 4 && maxLevels == 6) fast = nvproPyramidDefaultFastDispatcher<4, 6>;
  if(div == 2 && maxLevels == 6) fast = nvproPyramidDefaultFastDispatcher<2, 6>;
  if(div == 2 && maxLevels == 5) fast = nvproPyramidDefaultFastDispatcher<2, 5>;
  if(div == 2 && maxLevels == 3) fast = nvproPyramidDefaultFastDispatcher<2, 3>;
  if(div == 8 && maxLevels == 3) fast = nvproPyramidDefaultFastDispatcher<8, 3>;
  if(!fast)
    return -10;
  g_log.clear();
  NvproPyramidPipelines p;
  p.generalPipeline = kGeneral;
  p.fastPipeline    = kFast;
  nvproCmdPyramidDispatch(nullptr, p, w, h, mipLevels, nvproPyramidDefaultGeneralDispatcher, fast);
  // Re-run the decoder of the plain entry point on the log.
  std::vector<Event> log = g_log;
  uint32_t           n = 0, pc = 0, bound = 0, pendingBind = 0;
  for(const Event& e : log)
  {
    if(e.type == 0) { bound = e.a == reinterpret_cast<uint64_t>(kFast); pendingBind = 1; }
    else if(e.type == 1) pc = (uint32_t)e.a;
    else if(e.type == 2)
    {
      if(n >= maxSteps) return -1;
      out[n] = ref_step{};
      out[n].pipeline = bound; out[n].input_level = pc >> 5; out[n].level_count = pc & 31u;
      out[n].workgroups = (uint32_t)e.a; out[n].push_constant = pc; out[n].bind = pendingBind;
      pendingBind = 0; ++n;
    }
    else if(n) out[n - 1].barrier_after = 1;
  }
  return (int)n;
}

// ------------------------------------------------------- transfer functions
float    ref_linear_from_srgb(uint32_t c) { return linearFromSrgb(c); }
uint32_t ref_srgb_from_linear(float x) { return srgbFromLinear(x); }

// ------------------------------------------------------- CPU generator
// chain: packed MipmapStorage layout, level 0 filled on entry.
int ref_cpu_generate_srgba8(uint8_t* chain, uint32_t w, uint32_t h)
{
  MipmapStorage<uint8_t, 4> m(w, h);
  memcpy(m.levelData(0), chain, m.getLevelByteSize(0));
  cpuGenerateMipmaps_sRGBA(&m);
  memcpy(chain, m.levelData(0), m.getByteSize());
  return (int)m.getLevelOffsets().size();
}

// In-place variant without the copies, for timing (the reference generates in
// its own MipmapStorage; the copy-in of level 0 is outside the timed region).
struct ref_storage
{
  MipmapStorage<uint8_t, 4> m;
  ref_storage(uint32_t w, uint32_t h) : m(w, h) {}
};
void* ref_storage_create(uint32_t w, uint32_t h, const uint8_t* level0)
{
  auto* s = new ref_storage(w, h);
  memcpy(s->m.levelData(0), level0, s->m.getLevelByteSize(0));
  return s;
}
void     ref_storage_generate(void* s) { cpuGenerateMipmaps_sRGBA(&static_cast<ref_storage*>(s)->m); }
uint64_t ref_storage_bytes(void* s) { return static_cast<ref_storage*>(s)->m.getByteSize(); }
void     ref_storage_read(void* s, uint8_t* out)
{
  auto* p = static_cast<ref_storage*>(s);
  memcpy(out, p->m.levelData(0), p->m.getByteSize());
}
void ref_storage_destroy(void* s) { delete static_cast<ref_storage*>(s); }

// Layout as the reference computes it.
uint32_t ref_layout(uint32_t w, uint32_t h, uint64_t* offsets, uint32_t* widths, uint32_t* heights, uint32_t max)
{
  MipmapStorage<uint8_t, 4> m(1, 1);  // placeholder to keep a default path simple
  MipmapStorage<uint8_t, 4> real(w, h);
  uint32_t                  n = (uint32_t)real.getLevelOffsets().size();
  for(uint32_t i = 0; i < n && i < max; ++i)
  {
    offsets[i] = real.getLevelOffsets()[i];
    widths[i]  = real.getWidthHeight()[i].x;
    heights[i] = real.getWidthHeight()[i].y;
  }
  return n;
}

// MipmapStorage::compare
uint32_t ref_compare(const uint8_t* a, const uint8_t* b, uint32_t w, uint32_t h, uint32_t* xyzc)
{
  MipmapStorage<uint8_t, 4> m(w, h);
  memcpy(m.levelData(0), a, m.getByteSize());
  glm::uvec3 coord;
  uint32_t   ch    = 0;
  uint8_t    delta = m.compare(static_cast<const void*>(b), &coord, &ch);
  if(xyzc)
  {
    xyzc[0] = coord.x, xyzc[1] = coord.y, xyzc[2] = coord.z, xyzc[3] = ch;
  }
  return delta;
}

// Needed to link mipmap_storage.hpp's writeMipmapsTga; uncompressed 32-bit TGA.
int stbi_write_tga(char const* filename, int w, int h, int comp, const void* data)
{
  FILE* f = fopen(filename, "wb");
  if(!f || comp != 4)
    return 0;
  unsigned char hdr[18] = {0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, (unsigned char)(w & 255), (unsigned char)(w >> 8),
                           (unsigned char)(h & 255), (unsigned char)(h >> 8), 32, 0x28};
  fwrite(hdr, 1, 18, f);
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for(long i = 0; i < (long)w * h; ++i)
  {
    unsigned char bgra[4] = {p[4 * i + 2], p[4 * i + 1], p[4 * i + 0], p[4 * i + 3]};
    fwrite(bgra, 1, 4, f);
  }
  fclose(f);
  return 1;
}
}  // extern "C"
