// nvpyr_plan.hpp -- host-side scheduler (no CUDA, no Vulkan).
//
// Re-creates the decisions of the reference host scheduler
// (nvpro_pyramid/nvpro_pyramid_dispatch.hpp:109-292): which pipeline handles
// which run of levels.  On B200 those runs are not needed for performance
// (kernels are fused differently) but they are OBSERVABLE: inside one reference
// dispatch deeper levels are computed from un-quantised float32 values, between
// dispatches the 8-bit image is re-read.  The plan therefore defines the
// "carry groups" our kernels must honour to reproduce the reference bits.
//
// Like the reference (dispatch.hpp:99-104) the policy is pluggable: a dispatcher
// is a callback that inspects the state and claims 0..n levels.
#pragma once
#include <stdint.h>

#include "../../include/nvpyr.h"

namespace nvpyr {

// Mirrors NvproPyramidState (dispatch.hpp:63-75).
struct PyramidState
{
  uint32_t currentLevel    = 0;  // input level of the next step
  uint32_t remainingLevels = 0;  // never 0 when handed to a dispatcher
  uint32_t currentX = 0, currentY = 0;
};

// A dispatcher fills `step` (workgroups/pushConstant) and returns the number of
// levels it will fill; a fast dispatcher may decline with 0, a general one may not.
using dispatcher_t = uint32_t (*)(const PyramidState& state, nvpyrPlanStep& step);

inline uint32_t levelCountFor(uint32_t w, uint32_t h)
{
  if(w == 0 || h == 0)
    return 0;
  uint32_t m = w > h ? w : h, n = 0;
  for(; m != 0; m >>= 1)
    ++n;
  return n;
}
inline uint32_t levelDim(uint32_t d, uint32_t level)
{
  uint32_t v = level < 32 ? d >> level : 0;
  return v ? v : 1u;
}
inline uint64_t levelOffsetTexels(uint32_t w, uint32_t h, uint32_t level)
{
  uint64_t off = 0;
  for(uint32_t i = 0; i < level; ++i)
    off += uint64_t(levelDim(w, i)) * levelDim(h, i);
  return off;
}

// Equivalent of nvproPyramidDefaultFastDispatcher<Div, Max> (dispatch.hpp:195-242).
// Eligible when both edges are multiples of Div; claims as many levels as both
// edges stay even, bounded by what remains and by Max.
template <uint32_t Div = 4, uint32_t Max = 6>
inline uint32_t defaultFastDispatcher(const PyramidState& s, nvpyrPlanStep& step)
{
  static_assert(Div > 0 && Div % 2 == 0 && Max >= 1 && Max <= 6, "fast pipeline limits");
  if((s.currentX % Div) | (s.currentY % Div))
    return 0;
  const uint32_t cap = s.remainingLevels < Max ? s.remainingLevels : Max;
  uint32_t       n   = 0;
  while(n < cap && !((s.currentX >> n) & 1u) && !((s.currentY >> n) & 1u))
    ++n;
  // Work-group count the reference would launch: 4096 input texels per group for
  // a 6-level step, 1024 otherwise; 32-bit product as in dispatch.hpp:237.
  const uint32_t texels = s.currentX * s.currentY;
  step.workgroups       = n == 6 ? (texels + 4095u) >> 12 : (texels + 1023u) >> 10;
  step.pushConstant     = s.currentLevel << 5 | n;
  return n;
}

// Equivalent of nvproPyramidDefaultGeneralDispatcher (dispatch.hpp:247-292):
// two levels at a time (py2_4_8_8: 128 threads, 8x8 tile of the second level).
inline uint32_t defaultGeneralDispatcher(const PyramidState& s, nvpyrPlanStep& step)
{
  const uint32_t n  = s.remainingLevels < 2 ? s.remainingLevels : 2;
  const uint32_t dw = levelDim(s.currentX, n), dh = levelDim(s.currentY, n);
  step.workgroups   = n == 1 ? (dw * dh + 127u) / 128u : ((dw + 7u) / 8u) * ((dh + 7u) / 8u);
  step.pushConstant = s.currentLevel << 5 | n;
  return n;
}

// The blit fallback of demo_app/mipmap_pipelines.cpp:404-441 as a "general dispatcher": one level per step, no compute
// dispatch (workgroups 0); the step is executed by a blit of level currentLevel into currentLevel + 1.
inline uint32_t blitDispatcher(const PyramidState& s, nvpyrPlanStep& step)
{
  step.workgroups   = 0;
  step.pushConstant = s.currentLevel << 5 | 1u;
  return 1;
}

// Equivalent of the 7-argument nvproCmdPyramidDispatch (dispatch.hpp:109-188)
// with "record" replaced by "append to steps[]".  fast == nullptr means the fast
// pipeline is unavailable.  Returns the number of steps, or -1 on overflow / bad
// arguments / a dispatcher that violates its contract.
// A dispatcher callback in either form: a C++ function (the defaults above, nvpyr.cuh users) or the C ABI's
// nvpyrDispatcher + user data (nvpyrDispatchWithDispatchers).
struct DispatcherRef
{
  dispatcher_t    fn   = nullptr;
  nvpyrDispatcher cfn  = nullptr;
  void*           user = nullptr;
  DispatcherRef() = default;
  DispatcherRef(dispatcher_t f) : fn(f) {}
  DispatcherRef(decltype(nullptr)) {}
  DispatcherRef(nvpyrDispatcher f, void* u) : cfn(f), user(u) {}
  explicit operator bool() const { return fn != nullptr || cfn != nullptr; }
  bool     operator==(const DispatcherRef& o) const { return fn == o.fn && cfn == o.cfn && user == o.user; }
  bool     operator!=(const DispatcherRef& o) const { return !(*this == o); }
  uint32_t operator()(const PyramidState& s, nvpyrPlanStep& step) const
  {
    if(fn != nullptr)
      return fn(s, step);
    const nvpyrPyramidState cs{s.currentLevel, s.remainingLevels, s.currentX, s.currentY};
    return cfn(&cs, &step, user);
  }
};

inline int buildPlan(uint32_t baseWidth, uint32_t baseHeight, uint32_t mipLevels, const DispatcherRef& general,
                     const DispatcherRef& fast, nvpyrPlanStep* steps, uint32_t maxSteps)
{
  if(baseWidth == 0 || baseHeight == 0 || !general)
    return -1;
  if(mipLevels == 0)
    mipLevels = levelCountFor(baseWidth, baseHeight);
  PyramidState st;
  st.remainingLevels = mipLevels - 1;
  st.currentX        = baseWidth;
  st.currentY        = baseHeight;
  int  count         = 0;
  int  bound         = -1;  // -1 nothing, 0 general, 1 fast
  while(st.remainingLevels != 0)
  {
    if(uint32_t(count) >= maxSteps)
      return -1;
    nvpyrPlanStep step{};
    uint32_t      filled = fast ? fast(st, step) : 0;
    int           which  = 1;
    if(filled == 0)
    {
      which  = 0;
      filled = general(st, step);
    }
    if(filled == 0 || filled > st.remainingLevels)
      return -1;  // the reference asserts here (dispatch.hpp:169,172)
    step.pipeline     = uint32_t(which);
    step.inputLevel   = st.currentLevel;
    step.levelCount   = filled;
    step.srcWidth     = st.currentX;
    step.srcHeight    = st.currentY;
    step.bindPipeline = bound != which;
    bound             = which;
    st.currentLevel += filled;
    st.remainingLevels -= filled;
    st.currentX       = levelDim(st.currentX, filled);
    st.currentY       = levelDim(st.currentY, filled);
    step.barrierAfter = st.remainingLevels != 0;
    steps[count++]    = step;
  }
  return count;
}

// Run-time selection of the <Div, Max> instantiations the tests exercise.
inline dispatcher_t selectFastDispatcher(uint32_t div, uint32_t maxLevels)
{
  if(div == 0)
    div = 4;
  if(maxLevels == 0)
    maxLevels = 6;
#define NVPYR_FD(D, M)                                                                                            \
  if(div == D && maxLevels == M)                                                                                  \
    return defaultFastDispatcher<D, M>;
  NVPYR_FD(4, 6) NVPYR_FD(4, 5) NVPYR_FD(4, 4) NVPYR_FD(4, 3) NVPYR_FD(4, 2)
  NVPYR_FD(2, 6) NVPYR_FD(2, 5) NVPYR_FD(2, 4) NVPYR_FD(2, 3) NVPYR_FD(2, 2) NVPYR_FD(2, 1)
  NVPYR_FD(8, 6) NVPYR_FD(8, 5) NVPYR_FD(8, 4) NVPYR_FD(8, 3)
#undef NVPYR_FD
  return nullptr;
}

}  // namespace nvpyr
