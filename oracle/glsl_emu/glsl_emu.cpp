// glsl_emu.cpp -- driver of the GLSL-as-C++ execution of the reference path (TEST INFRASTRUCTURE ONLY).
//
// nvproCmdPyramidDispatch from the reference's own nvpro_pyramid_dispatch.hpp runs against a mock
// Vulkan (oracle/ref_shims/vulkan/vulkan_core.h) whose vkCmdDispatch EXECUTES the bound shader:
// emu_fast.cpp / emu_general.cpp are the reference's shader sources compiled as C++.  Each work group
// runs as 256 / 128 fibers (ucontext); subgroupShuffleXor and barrier() are the scheduling points.
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include <vector>

#include "glsl_shim.hpp"
#include "nvpro_pyramid/nvpro_pyramid_dispatch.hpp"
// the reference's CPU twin of the transfer functions
#include "shaders/srgb.h"

void emuFastMain();
void emuGeneralMain();
void emuFastSetImage(const uimage2D* levels);
void emuGeneralSetImage(const uimage2D* levels);
// the same shaders built with F16_SHARED = 1 (emu_fast_f16.cpp, emu_general_f16.cpp)
void emuFastMainF16();
void emuGeneralMainF16();
void emuFastSetImageF16(const uimage2D* levels);
void emuGeneralSetImageF16(const uimage2D* levels);
// ... and with SRGB_SHARED = 1 (emu_fast_srgb.cpp, emu_general_srgb.cpp)
void emuFastMainSrgb();
void emuGeneralMainSrgb();
void emuFastSetImageSrgb(const uimage2D* levels);
void emuGeneralSetImageSrgb(const uimage2D* levels);
uint emuGlslSrgbFromLinear(float x);

EmuInvocation* g_emuCur          = nullptr;
uint           g_emuPushConstant = 0;
uint64_t       g_emuStores       = 0;
float          emuLinearFromSrgb(uint c) { return linearFromSrgb(c); }

// ------------------------------------------------------------------------------- fibers
namespace {
constexpr size_t kStack = 128 * 1024;
struct Fiber
{
  ucontext_t        ctx;
  EmuInvocation     inv;
  int               state;  // 0 runnable, 1 waiting at barrier, 2 done
  std::vector<vec4> published;  // values this invocation offered to subgroupShuffleXor, in order
  uint              shuffles;
  char*             stack = nullptr;
};
std::vector<Fiber> g_fibers;
ucontext_t         g_sched;
Fiber*             g_fiber = nullptr;
void (*g_entry)()          = nullptr;

void trampoline()
{
  g_entry();
  g_fiber->state = 2;
  swapcontext(&g_fiber->ctx, &g_sched);
}
void yieldToScheduler() { swapcontext(&g_fiber->ctx, &g_sched); }

void runWorkgroup(uint wg, uint localSize, void (*entry)())
{
  g_entry = entry;
  if(g_fibers.size() < localSize)
    g_fibers.resize(localSize);
  for(uint l = 0; l < localSize; ++l)
  {
    Fiber& f = g_fibers[l];
    if(!f.stack)
      f.stack = static_cast<char*>(malloc(kStack));
    f.inv      = EmuInvocation{l, wg * localSize + l, wg};
    f.state    = 0;
    f.shuffles = 0;
    f.published.clear();
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp   = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link          = nullptr;
    makecontext(&f.ctx, trampoline, 0);
  }
  for(;;)
  {
    uint done = 0, atBarrier = 0;
    for(uint l = 0; l < localSize; ++l)
    {
      Fiber& f = g_fibers[l];
      if(f.state == 0)
      {
        g_fiber  = &f;
        g_emuCur = &f.inv;
        swapcontext(&g_sched, &f.ctx);  // runs until its next scheduling point
      }
      done += f.state == 2;
      atBarrier += f.state == 1;
    }
    if(done == localSize)
      break;
    if(done + atBarrier == localSize)  // everybody still alive has arrived: release
      for(uint l = 0; l < localSize; ++l)
        if(g_fibers[l].state == 1)
          g_fibers[l].state = 0;
  }
}
}  // namespace

void barrier()
{
  g_fiber->state = 1;
  yieldToScheduler();
}

// All invocations that execute the same shuffle publish in the same scheduler pass, so after the yield
// the partner's value for this shuffle number is available.  A partner that is inactive (returned, or
// on another path) has not published: the result is undefined in GLSL and unused by the shader.
vec4 subgroupShuffleXor(const vec4& v, uint mask)
{
  Fiber*     self = g_fiber;
  const uint n    = self->shuffles++;
  self->published.push_back(v);
  yieldToScheduler();
  const uint partner = (self->inv.local & ~31u) | ((self->inv.local & 31u) ^ mask);
  if(partner < g_fibers.size() && g_fibers[partner].published.size() > n && g_fibers[partner].state != 2)
    return g_fibers[partner].published[n];
  return v;
}

// ------------------------------------------------------------------------------- mock Vulkan that executes
namespace {
int            g_bound = -1;  // 0 general, 1 fast
MockPipeline_T* kGeneral = reinterpret_cast<MockPipeline_T*>(0x1000);
MockPipeline_T* kFast    = reinterpret_cast<MockPipeline_T*>(0x2000);
uint           g_dispatches = 0;
int            g_sharedMode = 0;  // which build of the shaders the mock pipelines stand for: 0 default, 1 F16_SHARED, 2 SRGB_SHARED
}  // namespace
void vkCmdBindPipeline(VkCommandBuffer, VkPipelineBindPoint, VkPipeline p) { g_bound = p == kFast ? 1 : 0; }
void vkCmdPushConstants(VkCommandBuffer, VkPipelineLayout, VkShaderStageFlags, uint32_t, uint32_t size, const void* v)
{
  memcpy(&g_emuPushConstant, v, size < 4 ? size : 4);
}
void vkCmdDispatch(VkCommandBuffer, uint32_t x, uint32_t, uint32_t)
{
  ++g_dispatches;
  for(uint32_t wg = 0; wg < x; ++wg)
    runWorkgroup(wg, g_bound == 1 ? 256u : 128u,
                 g_bound == 1 ? (g_sharedMode == 1 ? emuFastMainF16 : g_sharedMode == 2 ? emuFastMainSrgb : emuFastMain)
                              : (g_sharedMode == 1 ? emuGeneralMainF16 : g_sharedMode == 2 ? emuGeneralMainSrgb : emuGeneralMain));
}
void vkCmdPipelineBarrier(VkCommandBuffer, VkPipelineStageFlags, VkPipelineStageFlags, VkDependencyFlags, uint32_t,
                          const VkMemoryBarrier*, uint32_t, const VkBufferMemoryBarrier*, uint32_t,
                          const VkImageMemoryBarrier*)
{
}

extern "C" {
// Runs the reference's nvproCmdPyramidDispatch + shaders on a packed RGBA8 chain (level 0 filled).
// Returns the number of dispatches executed; *stores receives the number of imageStore calls.
int emu_run_chain_ex(uint8_t* chain, uint32_t w, uint32_t h, uint32_t mipLevels, uint32_t haveFast, uint32_t f16Shared,
                     uint64_t* stores);
int emu_run_chain(uint8_t* chain, uint32_t w, uint32_t h, uint32_t mipLevels, uint32_t haveFast, uint64_t* stores)
{
  return emu_run_chain_ex(chain, w, h, mipLevels, haveFast, 0, stores);
}
// f16Shared = 1: the F16_SHARED build of both shaders; 2: the SRGB_SHARED build.
int emu_run_chain_ex(uint8_t* chain, uint32_t w, uint32_t h, uint32_t mipLevels, uint32_t haveFast, uint32_t f16Shared,
                     uint64_t* stores)
{
  g_sharedMode = int(f16Shared);
  uint32_t levels = mipLevels;
  if(levels == 0)
    for(uint32_t a = w, b = h; a != 0 || b != 0; a >>= 1, b >>= 1)
      ++levels;
  if(levels > 16)
    return -1;  // the sRGBA8 instance has 16 storage views (srgba8_mipmap_preamble.glsl:13)
  uimage2D imageMipLevels_shared[16];
  size_t   off = 0;
  for(uint32_t i = 0; i < 16; ++i)
  {
    // views beyond the last level alias the last one, like ScopedImage pads its descriptor array
    const uint32_t l  = i < levels ? i : levels - 1;
    uint32_t       lw = w >> l, lh = h >> l;
    lw = lw ? lw : 1, lh = lh ? lh : 1;
    if(i < levels)
    {
      imageMipLevels_shared[i] = uimage2D{chain + 4 * off, int(lw), int(lh)};
      off += size_t(lw) * lh;
    }
    else
      imageMipLevels_shared[i] = imageMipLevels_shared[levels - 1];
  }
  emuFastSetImage(imageMipLevels_shared);
  emuGeneralSetImage(imageMipLevels_shared);
  emuFastSetImageF16(imageMipLevels_shared);
  emuGeneralSetImageF16(imageMipLevels_shared);
  emuFastSetImageSrgb(imageMipLevels_shared);
  emuGeneralSetImageSrgb(imageMipLevels_shared);
  g_emuStores           = 0;
  g_dispatches          = 0;
  g_bound               = -1;
  NvproPyramidPipelines p;
  p.generalPipeline = kGeneral;
  p.fastPipeline    = haveFast ? kFast : VK_NULL_HANDLE;
  nvproCmdPyramidDispatch(nullptr, p, w, h, mipLevels);
  if(stores)
    *stores = g_emuStores;
  return int(g_dispatches);
}
// The GLSL twin of the encode (srgba8_mipmap_preamble.glsl:110-121) as compiled here, for comparison with
// the pinned thresholds.
uint32_t emu_glsl_srgb_from_linear(float x) { return emuGlslSrgbFromLinear(x); }
}
