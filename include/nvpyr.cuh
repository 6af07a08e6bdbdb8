/* nvpyr.cuh -- C++/CUDA template API of the B200 mip-pyramid generator: the USER-DEFINED instance.
 *
 * The reference library is a GLSL template (nvpro_pyramid/nvpro_pyramid.glsl) that the user configures with
 * macros before including it -- NVPRO_PYRAMID_TYPE / _LOAD / _REDUCE / _STORE / _LEVEL_SIZE and the optional
 * _REDUCE2 / _REDUCE4 / _LOAD_REDUCE4 / _SHUFFLE_XOR / _SHARED_* (glsl:27-120, defaults :164-207) -- and a host
 * scheduler, nvproCmdPyramidDispatch, that takes two optional dispatcher callbacks
 * (nvpro_pyramid_dispatch.hpp:99-116).  The shipped sRGBA8 configuration is
 * srgba8_mipmap_preamble.glsl; that instance (and an rgba32f one) is what the C ABI in nvpyr.h runs.
 *
 * This header is the same extension point for CUDA: the macros become a C++ functor set, the dispatcher
 * callbacks become nvpyr::dispatcher_t function pointers, and nvpyr::dispatch<Set>() plays the role of
 * nvproCmdPyramidDispatch with the user's pipelines bound.  The kernels it instantiates are the library's
 * functor-template kernels (vk_compute_mipmaps_b200/csrc/nvpyr_kernels.cuh): same schedule, same work
 * decomposition per dispatch, same float32 expression trees as the reference shaders.
 *
 *   struct DepthMax : nvpyr::PyramidFunctors<DepthMax>            // a hi-z pyramid: max of the footprint
 *   {
 *     using Value = float;                                          // NVPRO_PYRAMID_TYPE
 *     static constexpr int kTexelBytes = 4;                         // size of one texel in memory
 *     struct Params {};                                             // optional device-resident user data
 *     __device__ static Value load(const Params*, const void* t) { return *static_cast<const float*>(t); }     // _LOAD
 *     __device__ static void  store(const Params*, void* t, Value v) { *static_cast<float*>(t) = v; }          // _STORE
 *     __device__ static Value reduce(float a0, Value v0, float a1, Value v1, float a2, Value v2)               // _REDUCE
 *     { return fmaxf(v0, fmaxf(v1, v2)); }
 *   };
 *   nvpyrDispatchDesc d = {...};  nvpyr::dispatch<DepthMax>(d);
 *
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -I include your.cu   (header-only; libnvpyr.so is NOT
 * needed by this path).  Compile with -fmad=false if the reduce functions must not be contracted.
 *
 * Contract of a functor set S (static members; everything is __device__):
 *   S::Value, S::kTexelBytes, S::load, S::store, S::reduce                                   -- required
 *   S::Params                                   user data in DEVICE memory, pointer handed to load/store;
 *                                               stands for the descriptor sets / push constants a GLSL user
 *                                               binds around the dispatch (dispatch.hpp:48-53)
 *   S::reduce2(v0, v1)                          default reduce(0.5, v0, 0.5, v1, 0, v1)          (glsl:164-167)
 *   S::reduce4(v00, v01, v10, v11)              default reduce2(reduce2(v00, v01), reduce2(v10, v11)) (:169-177)
 *   S::sharedRound(v)                           SHARED_STORE followed by SHARED_LOAD, default identity (:196-207)
 *   S::loadReduce4(params, texel00, rowPitchBytes, x, y, level)
 *                                               _LOAD_REDUCE4: load the 2x2 square whose upper-left texel is (x, y) of
 *                                               mip level `level` (its address: texel00) and reduce it.  Fast pipeline
 *                                               only, first level of every dispatch, exactly where the GLSL template
 *                                               expands the macro (glsl:78-88); default = the four loads + reduce4 of
 *                                               glsl:179-189.  A set may fetch through a texture object kept in its
 *                                               Params here to use the hardware's bilinear filter.
 * _LEVEL_SIZE has no counterpart: in GLSL it is how the shader learns a size the host scheduler has already assumed
 * (dispatch.hpp derives every level size from the base size); here the kernels get the same sizes from the descriptor.
 * Value must be trivially copyable and made of 1..16 32-bit words (it travels through warp shuffles and
 * shared memory).  reduce4's argument order tells which neighbours share a bracket; the kernels call it with
 * the reference's per-site pairing (SURVEY.md section 8a note 1).
 */
#pragma once
#ifndef __CUDACC__
#error "nvpyr.cuh is a CUDA C++ header: compile with nvcc"
#endif
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <utility>

#include "nvpyr.h"
#include "../vk_compute_mipmaps_b200/csrc/nvpyr_kernels.cuh"
#include "../vk_compute_mipmaps_b200/csrc/nvpyr_plan.hpp"

namespace nvpyr {

// Defaults of the optional members, as in nvpro_pyramid.glsl:164-207 (CRTP: S derives from PyramidFunctors<S>).
template <class S>
struct PyramidFunctors
{
  struct Params
  {
  };
  template <class V>
  __device__ __forceinline__ static V reduce2(V v0, V v1)
  {
    return S::reduce(0.5f, v0, 0.5f, v1, 0.0f, v1);
  }
  template <class V>
  __device__ __forceinline__ static V reduce4(V v00, V v01, V v10, V v11)
  {
    return S::reduce2(S::reduce2(v00, v01), S::reduce2(v10, v11));
  }
  template <class V>
  __device__ __forceinline__ static V sharedRound(V v)
  {
    return v;
  }
  // NVPRO_PYRAMID_LOAD_REDUCE4's default (glsl:179-189): loads (0,0), (0,1), (1,0), (1,1), then reduce4 in that order.
  template <class P>
  __device__ __forceinline__ static auto loadReduce4(const P* params, const void* texel00, size_t rowPitchBytes, uint32_t,
                                                     uint32_t, uint32_t)
  {
    const unsigned char* t   = static_cast<const unsigned char*>(texel00);
    const auto           v00 = S::load(params, t), v01 = S::load(params, t + rowPitchBytes);
    const auto           v10 = S::load(params, t + S::kTexelBytes), v11 = S::load(params, t + rowPitchBytes + S::kTexelBytes);
    return S::reduce4(v00, v01, v10, v11);
  }
};

namespace detail {

// Presents a user set to the kernel templates (the contract of nvpyr_functors.cuh).
template <class S>
struct UserSet
{
  using Value                      = typename S::Value;
  using Params                     = typename S::Params;
  static constexpr int kTexelBytes = S::kTexelBytes;
  struct Shared
  {
    const Params* params;
  };
  // FastParams::tables / GeneralParams::tables carry the user's Params pointer for user sets.
  __device__ static void sharedInit(Shared& s, const DeviceTables* t)
  {
    if(threadIdx.x == 0)
      s.params = reinterpret_cast<const Params*>(t);
  }
  __device__ __forceinline__ static Value load(const Shared& s, const void* p) { return S::load(s.params, p); }
  __device__ __forceinline__ static void  load4(const Shared& s, const void* p, Value out[4])
  {
#pragma unroll
    for(int i = 0; i < 4; ++i)
      out[i] = S::load(s.params, static_cast<const unsigned char*>(p) + i * kTexelBytes);
  }
  template <bool kClampHigh = true>
  __device__ __forceinline__ static void store(const Shared& s, void* p, Value v)
  {
    S::store(s.params, p, v);
  }
  template <bool kClampHigh = true>
  __device__ __forceinline__ static void store2(const Shared& s, void* p, Value v0, Value v1)
  {
    S::store(s.params, p, v0);
    S::store(s.params, static_cast<unsigned char*>(p) + kTexelBytes, v1);
  }
  __device__ __forceinline__ static Value reduce(float a0, Value v0, float a1, Value v1, float a2, Value v2)
  {
    return S::reduce(a0, v0, a1, v1, a2, v2);
  }
  __device__ __forceinline__ static Value reduce2(Value v0, Value v1) { return S::reduce2(v0, v1); }
  __device__ __forceinline__ static Value reduce4(Value a, Value b, Value c, Value d) { return S::reduce4(a, b, c, d); }
  __device__ __forceinline__ static Value sharedRound(Value v) { return S::sharedRound(v); }
  __device__ __forceinline__ static Value loadReduce4(const Shared& s, const void* texel00, size_t rowPitchBytes, uint32_t x,
                                                      uint32_t y, uint32_t level)
  {
    return S::loadReduce4(s.params, texel00, rowPitchBytes, x, y, level);
  }
};

// Blocks per SM of a kernel on the current device (and the opt-in to its dynamic shared memory), looked up once per
// (kernel, device): the two runtime calls cost as much as a small launch.
inline int blocksPerSm(const void* kernel, int threads, size_t smem, int device)
{
  static std::mutex                                 m;
  static std::map<std::pair<const void*, int>, int> cache;
  std::lock_guard<std::mutex>                       lock(m);
  const auto                                        it = cache.find({kernel, device});
  if(it != cache.end())
    return it->second;
  int perSm = 0;
  if(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess
     || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, threads, smem) != cudaSuccess)
    return -1;
  cache[{kernel, device}] = perSm;
  return perSm;
}

template <class K>
inline nvpyrStatus launchOn(K kernel, uint64_t work, int threads, size_t smem, cudaStream_t stream, int device,
                            int smCount, const void* params)
{
  const int perSm = blocksPerSm(reinterpret_cast<const void*>(kernel), threads, smem, device);
  if(perSm < 0)
    return NVPYR_ERROR_CUDA;
  if(perSm < 1)
    return NVPYR_ERROR_UNSUPPORTED;
  uint64_t grid = uint64_t(perSm) * uint64_t(smCount);
  grid          = grid > work ? work : grid;
  grid          = grid < 1 ? 1 : grid;
  void* args[]  = {const_cast<void*>(params)};
  if(cudaLaunchKernel(reinterpret_cast<const void*>(kernel), dim3(unsigned(grid)), dim3(unsigned(threads)), args, smem,
                      stream)
     != cudaSuccess)
    return NVPYR_ERROR_CUDA;
  return NVPYR_SUCCESS;
}

template <class F, int M>
inline nvpyrStatus launchFastStep(const FastParams& p, cudaStream_t stream, int device, int smCount)
{
  return launchOn(fastKernel<F, M, false>, uint64_t(p.tilesX) * p.tilesY, 256, sizeof(FastSmem<F>), stream, device, smCount,
                  &p);
}

}  // namespace detail

// nvproCmdPyramidDispatch (nvpro_pyramid_dispatch.hpp:109-188) for a user-defined functor set: fills levels
// 1..levelCount-1 of the image described by `desc` from its level 0, asynchronously on desc.stream.
//   desc.format is ignored (the texel format is S's business); desc.flags may hold NVPYR_FLAG_FORCE_GENERAL
//   (= NvproPyramidPipelines::fastPipeline == VK_NULL_HANDLE); packed MipmapStorage layout unless desc.levels
//   / rowPitchBytes say otherwise.
//   deviceParams: pointer to an S::Params in device memory (or nullptr), handed to S::load / S::store.
//   general / fast: the dispatcher callbacks of dispatch.hpp:99-116.  A dispatcher returns how many levels
//   its pipeline fills from the state it is given (fast may return 0 = "not eligible", at most 6; general
//   1 or 2) -- the same contract as nvpro_pyramid_dispatcher_t; fast == nullptr selects the default
//   <DivisibilityRequirement 4, MaxLevels 6>.  One kernel launch per dispatch, stream order in place of the
//   reference's pipeline barriers.
template <class S>
inline nvpyrStatus dispatch(const nvpyrDispatchDesc& desc, const typename S::Params* deviceParams = nullptr,
                            dispatcher_t general = nullptr, dispatcher_t fast = nullptr)
{
  using F = detail::UserSet<S>;
  static_assert(sizeof(typename F::Value) % 4 == 0 && sizeof(typename F::Value) <= 64, "Value: 1..16 32-bit words");
  if(desc.structSize != sizeof(nvpyrDispatchDesc) || desc.extent.width == 0 || desc.extent.height == 0)
    return NVPYR_ERROR_INVALID_VALUE;
  if(desc.flags & ~uint32_t(NVPYR_FLAG_FORCE_GENERAL))
    return NVPYR_ERROR_UNSUPPORTED;
  const uint32_t w = desc.extent.width, h = desc.extent.height, maxLevels = levelCountFor(w, h);
  const uint32_t levels = desc.levelCount == 0 ? maxLevels : desc.levelCount;
  if(levels > maxLevels || levels > NVPYR_MAX_LEVELS)
    return NVPYR_ERROR_INVALID_VALUE;
  if(general == nullptr)
    general = defaultGeneralDispatcher;
  if(desc.flags & NVPYR_FLAG_FORCE_GENERAL)
    fast = nullptr;
  else if(fast == nullptr)
  {
    fast = selectFastDispatcher(desc.fastDivisibility, desc.fastMaxLevels);
    if(fast == nullptr)
      return NVPYR_ERROR_UNSUPPORTED;  // like the C ABI: an unknown <Div, Max> pair is an error, never a silent general-only plan
  }

  LevelView lv[NVPYR_MAX_LEVELS];
  uint64_t  off = 0;
  for(uint32_t i = 0; i < levels; ++i)
  {
    lv[i].w     = levelDim(w, i);
    lv[i].h     = levelDim(h, i);
    lv[i].level = i;
    if(desc.levels[i] != nullptr)
    {
      lv[i].ptr   = static_cast<unsigned char*>(desc.levels[i]);
      lv[i].pitch = desc.rowPitchBytes[i] ? desc.rowPitchBytes[i] : lv[i].w * uint32_t(F::kTexelBytes);
    }
    else
    {
      if(desc.base == nullptr)
        return NVPYR_ERROR_INVALID_VALUE;
      lv[i].ptr   = static_cast<unsigned char*>(desc.base) + off * uint64_t(F::kTexelBytes);
      lv[i].pitch = lv[i].w * uint32_t(F::kTexelBytes);
    }
    if(lv[i].pitch < lv[i].w * uint32_t(F::kTexelBytes))
      return NVPYR_ERROR_INVALID_VALUE;
    off += uint64_t(lv[i].w) * lv[i].h;
  }

  nvpyrPlanStep steps[NVPYR_MAX_STEPS];
  const int     n = buildPlan(w, h, levels, general, fast, steps, NVPYR_MAX_STEPS);
  if(n < 0)
    return NVPYR_ERROR_INVALID_VALUE;  // a dispatcher filled 0 levels or too many (the reference asserts)
  // Limits of the kernels, checked for the whole plan before anything is enqueued.
  for(int i = 0; i < n; ++i)
  {
    const nvpyrPlanStep& s = steps[i];
    if(s.pipeline == 1 ? (s.levelCount < 1 || s.levelCount > 6 || ((s.srcWidth | s.srcHeight) & ((1u << s.levelCount) - 1u)))
                       : (s.levelCount < 1 || s.levelCount > 2))
      return NVPYR_ERROR_INVALID_VALUE;  // a custom dispatcher promised levels its pipeline cannot fill
  }
  int device = 0, smCount = 0;
  if(cudaGetDevice(&device) != cudaSuccess
     || cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
    return NVPYR_ERROR_CUDA;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(desc.stream);
  const auto*  tables = reinterpret_cast<const DeviceTables*>(deviceParams);

  for(int i = 0; i < n; ++i)
  {
    const nvpyrPlanStep& s  = steps[i];
    nvpyrStatus          st = NVPYR_SUCCESS;
    if(s.pipeline == 1)
    {
      FastParams p{};
      for(uint32_t k = 0; k <= s.levelCount; ++k)
        p.lv[k] = lv[s.inputLevel + k];
      p.tilesX = (p.lv[0].w + 63u) / 64u;
      p.tilesY = (p.lv[0].h + 63u) / 64u;
      p.tables = tables;
      switch(s.levelCount)
      {
        case 1:
          st = detail::launchOn(fastKernel1<F>, (uint64_t(p.lv[1].w) * p.lv[1].h + 255u) / 256u, 256, sizeof(FastSmem<F>),
                                stream, device, smCount, &p);
          break;
        case 2: st = detail::launchFastStep<F, 2>(p, stream, device, smCount); break;
        case 3: st = detail::launchFastStep<F, 3>(p, stream, device, smCount); break;
        case 4: st = detail::launchFastStep<F, 4>(p, stream, device, smCount); break;
        case 5: st = detail::launchFastStep<F, 5>(p, stream, device, smCount); break;
        default: st = detail::launchFastStep<F, 6>(p, stream, device, smCount); break;
      }
    }
    else
    {
      GeneralParams p{};
      for(uint32_t k = 0; k <= s.levelCount; ++k)
        p.lv[k] = lv[s.inputLevel + k];
      p.levels            = s.levelCount;
      const LevelView& o  = s.levelCount == 1 ? p.lv[1] : p.lv[2];
      const uint32_t   t  = s.levelCount == 1 ? 2u * kGenTile2 : uint32_t(kGenTile2);
      p.tilesX            = (o.w + t - 1u) / t;
      p.tilesY            = (o.h + t - 1u) / t;
      p.tables            = tables;
      st = detail::launchOn(generalKernel<F>, uint64_t(p.tilesX) * p.tilesY, 256, sizeof(GeneralSmem<F>), stream, device,
                            smCount, &p);
    }
    if(st != NVPYR_SUCCESS)
      return st;
  }
  return NVPYR_SUCCESS;
}

}  // namespace nvpyr
