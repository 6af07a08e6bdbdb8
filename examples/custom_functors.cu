// custom_functors.cu -- the user-defined instance of the pyramid template (include/nvpyr.cuh), i.e. what a
// user of the reference does by defining the NVPRO_PYRAMID_* macros before including nvpro_pyramid.glsl
// (nvpro_pyramid/nvpro_pyramid.glsl:27-120) and by passing their own dispatcher callbacks to
// nvproCmdPyramidDispatch (nvpro_pyramid_dispatch.hpp:99-116).
//
//   1. DepthMax   -- a hi-z pyramid over an R32F image: Value = float, reduce = max of the footprint, a
//                    device-resident Params block (an upper clamp applied by the store).  max is exact and
//                    associative, so EVERY schedule must give exactly the CPU result: checked bit for bit for
//                    power-of-two, odd and mixed sizes, with the default dispatchers, with the fast pipeline
//                    absent, with <4, 3> fast limits and with a custom general dispatcher that fills one level
//                    per dispatch.
//   2. MyRgba32f  -- the sRGBA8 preamble's reduce functions (srgba8_mipmap_preamble.glsl:24-25,:35,:37-38) on
//                    float4 texels, written as a user set: must reproduce the library's own RGBA32F instance
//                    (nvpyrDispatchEx) bit for bit.
//   3. MinFirst   -- DepthMax with its own NVPRO_PYRAMID_LOAD_REDUCE4 (nvpro_pyramid.glsl:78-88): the hook returns the
//                    MINIMUM of the 2x2 square, so its effect is visible -- the first level of every fast dispatch
//                    must hold minima, every other level maxima, and a chain without the fast pipeline none at all
//                    (the general pipeline never expands the macro).  Checked against a CPU loop that follows the
//                    plan, for one-dispatch images, multi-dispatch images and M = 1 steps.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -std=c++17 -I include examples/custom_functors.cu
//             -L vk_compute_mipmaps_b200 -lnvpyr -o examples/custom_functors
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nvpyr.cuh"

struct DepthMax : nvpyr::PyramidFunctors<DepthMax>
{
  using Value                      = float;
  static constexpr int kTexelBytes = 4;
  struct Params
  {
    float clampMax;
  };
  __device__ static Value load(const Params*, const void* t) { return *static_cast<const float*>(t); }
  __device__ static void  store(const Params* p, void* t, Value v) { *static_cast<float*>(t) = fminf(v, p->clampMax); }
  __device__ static Value reduce(float, Value v0, float, Value v1, float, Value v2) { return fmaxf(v0, fmaxf(v1, v2)); }
};

struct MinFirst : nvpyr::PyramidFunctors<MinFirst>
{
  using Value                      = float;
  static constexpr int kTexelBytes = 4;
  __device__ static Value load(const Params*, const void* t) { return *static_cast<const float*>(t); }
  __device__ static void  store(const Params*, void* t, Value v) { *static_cast<float*>(t) = v; }
  __device__ static Value reduce(float, Value v0, float, Value v1, float, Value v2) { return fmaxf(v0, fmaxf(v1, v2)); }
  // NVPRO_PYRAMID_LOAD_REDUCE4(srcCoord, srcLevel, out_); checks the coordinates it is given against the address
  __device__ static Value loadReduce4(const Params*, const void* texel00, size_t pitch, uint32_t x, uint32_t y, uint32_t level)
  {
    const unsigned char* t = static_cast<const unsigned char*>(texel00);
    const float a = *reinterpret_cast<const float*>(t), b = *reinterpret_cast<const float*>(t + 4);
    const float c = *reinterpret_cast<const float*>(t + pitch), d = *reinterpret_cast<const float*>(t + pitch + 4);
    const bool  coordsOk = (x & 1u) == 0u && (y & 1u) == 0u && level < 32u;
    return coordsOk ? fminf(fminf(a, b), fminf(c, d)) : -1.0f;
  }
};

struct MyRgba32f : nvpyr::PyramidFunctors<MyRgba32f>
{
  using Value                      = float4;
  static constexpr int kTexelBytes = 16;
  __device__ static Value load(const Params*, const void* t) { return *static_cast<const float4*>(t); }
  __device__ static void  store(const Params*, void* t, Value v) { *static_cast<float4*>(t) = v; }
  __device__ static float r1(float a0, float v0, float a1, float v1, float a2, float v2)
  {
    return __fmaf_rn(a2, v2, __fmaf_rn(a1, v1, __fmul_rn(a0, v0)));  // the library's numerics contract
  }
  __device__ static Value reduce(float a0, Value v0, float a1, Value v1, float a2, Value v2)
  {
    return make_float4(r1(a0, v0.x, a1, v1.x, a2, v2.x), r1(a0, v0.y, a1, v1.y, a2, v2.y), r1(a0, v0.z, a1, v1.z, a2, v2.z),
                       r1(a0, v0.w, a1, v1.w, a2, v2.w));
  }
  __device__ static float h(float a, float b) { return __fmul_rn(0.5f, __fadd_rn(a, b)); }
  __device__ static Value reduce2(Value a, Value b) { return make_float4(h(a.x, b.x), h(a.y, b.y), h(a.z, b.z), h(a.w, b.w)); }
  __device__ static float q(float a, float b, float c, float d)
  {
    return __fmul_rn(0.25f, __fadd_rn(__fadd_rn(a, b), __fadd_rn(c, d)));
  }
  __device__ static Value reduce4(Value a, Value b, Value c, Value d)
  {
    return make_float4(q(a.x, b.x, c.x, d.x), q(a.y, b.y, c.y, d.y), q(a.z, b.z, c.z, d.z), q(a.w, b.w, c.w, d.w));
  }
};

// Dispatchers the library must reject (the reference asserts, dispatch.hpp:169,172): nothing may be enqueued.
static uint32_t fillsNothing(const nvpyr::PyramidState&, nvpyrPlanStep&) { return 0; }
static uint32_t fillsTooMuch(const nvpyr::PyramidState& s, nvpyrPlanStep&) { return s.remainingLevels + 1; }
static uint32_t fastOnOddSizes(const nvpyr::PyramidState& s, nvpyrPlanStep&) { return s.remainingLevels < 3 ? s.remainingLevels : 3; }

// A user dispatcher (nvpro_pyramid_dispatcher_t): general pipeline, one level per dispatch.
static uint32_t oneLevelGeneral(const nvpyr::PyramidState& s, nvpyrPlanStep& step)
{
  const uint32_t dw = nvpyr::levelDim(s.currentX, 1), dh = nvpyr::levelDim(s.currentY, 1);
  step.workgroups   = (dw * dh + 127u) / 128u;
  step.pushConstant = s.currentLevel << 5 | 1u;
  return 1;
}

#define CK(x)                                                                                                     \
  do                                                                                                              \
  {                                                                                                               \
    cudaError_t e_ = (x);                                                                                         \
    if(e_ != cudaSuccess)                                                                                         \
    {                                                                                                             \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));                                   \
      exit(2);                                                                                                    \
    }                                                                                                             \
  } while(0)

static uint32_t dim(uint32_t d, uint32_t l)
{
  const uint32_t v = d >> l;
  return v ? v : 1u;
}
static int taps(uint32_t size)
{
  return size == 1 ? 1 : (size & 1u) ? 3 : 2;  // kernelSizeFromInputSize_, glsl:557-561
}

// CPU hi-z chain: level i+1 = max over the reference's footprint of level i, clamp applied by every store.
static std::vector<float> cpuDepthChain(const std::vector<float>& l0, uint32_t w, uint32_t h, uint32_t levels, float clampMax)
{
  std::vector<float> chain(l0);
  size_t             src = 0;
  for(uint32_t l = 0; l + 1 < levels; ++l)
  {
    const uint32_t sw = dim(w, l), sh = dim(h, l), dw = dim(w, l + 1), dh = dim(h, l + 1);
    const int      kx = taps(sw), ky = taps(sh);
    const size_t   dst = chain.size();
    chain.resize(dst + size_t(dw) * dh);
    for(uint32_t y = 0; y < dh; ++y)
      for(uint32_t x = 0; x < dw; ++x)
      {
        float m = -INFINITY;
        for(int j = 0; j < ky; ++j)
          for(int i = 0; i < kx; ++i)
            m = fmaxf(m, chain[src + size_t(2 * y + j) * sw + (2 * x + i)]);
        chain[dst + size_t(y) * dw + x] = fminf(m, clampMax);
      }
    src = dst;
  }
  return chain;
}

static int failures = 0;

static void checkDepth(uint32_t w, uint32_t h, const char* what, uint32_t flags, uint32_t div, uint32_t maxLevels,
                       nvpyr::dispatcher_t general, uint32_t levelCount = 0)
{
  const uint32_t     levels = levelCount ? levelCount : nvpyr::levelCountFor(w, h);
  std::vector<float> l0(size_t(w) * h);
  uint32_t           s = 12345u + w * 31u + h;
  for(float& v : l0)
  {
    s = s * 1664525u + 1013904223u;
    v = float(s >> 8) * (1.0f / 16777216.0f);
  }
  const float              clampMax = 0.999f;
  const std::vector<float> want     = cpuDepthChain(l0, w, h, levels, clampMax);
  float*                   dev      = nullptr;
  DepthMax::Params*        params   = nullptr;
  CK(cudaMalloc(&dev, want.size() * 4));
  CK(cudaMemset(dev, 0xFF, want.size() * 4));
  CK(cudaMemcpy(dev, l0.data(), l0.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&params, sizeof(DepthMax::Params)));
  const DepthMax::Params hp{clampMax};
  CK(cudaMemcpy(params, &hp, sizeof(hp), cudaMemcpyHostToDevice));
  nvpyrDispatchDesc d;
  memset(&d, 0, sizeof(d));
  d.structSize       = sizeof(d);
  d.flags            = flags;
  d.extent           = {w, h};
  d.levelCount       = levelCount;
  d.base             = dev;
  d.fastDivisibility = div;
  d.fastMaxLevels    = maxLevels;
  const nvpyrStatus st = nvpyr::dispatch<DepthMax>(d, params, general);
  CK(cudaDeviceSynchronize());
  std::vector<float> got(want.size());
  CK(cudaMemcpy(got.data(), dev, got.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for(size_t i = l0.size(); i < want.size(); ++i)
    bad += memcmp(&got[i], &want[i], 4) != 0;
  printf("DepthMax  %5ux%-5u %-34s status %d, %zu of %zu texels differ\n", w, h, what, int(st), bad, want.size() - l0.size());
  failures += st != NVPYR_SUCCESS || bad != 0;
  cudaFree(dev);
  cudaFree(params);
}

// MinFirst against a CPU loop that follows the plan: level l + 1 holds 2x2 minima when it is the first level of a
// fast dispatch, footprint maxima otherwise.
static void checkMinFirst(uint32_t w, uint32_t h, const char* what, uint32_t flags, uint32_t div, uint32_t maxLevels)
{
  const uint32_t     levels = nvpyr::levelCountFor(w, h);
  std::vector<float> chain(size_t(w) * h);
  uint32_t           s = 4242u + w * 17u + h;
  for(float& v : chain)
  {
    s = s * 1664525u + 1013904223u;
    v = float(s >> 8) * (1.0f / 16777216.0f);
  }
  const size_t       l0Texels = chain.size();
  nvpyrPlanStep      steps[NVPYR_MAX_STEPS];
  nvpyr::dispatcher_t fast = (flags & NVPYR_FLAG_FORCE_GENERAL) ? nullptr : nvpyr::selectFastDispatcher(div, maxLevels);
  const int           n    = nvpyr::buildPlan(w, h, levels, nvpyr::defaultGeneralDispatcher, fast, steps, NVPYR_MAX_STEPS);
  std::vector<bool>   minLevel(levels + 1, false);  // minLevel[l]: level l is the first output of a fast dispatch
  size_t              hooks = 0;
  for(int i = 0; i < n; ++i)
    if(steps[i].pipeline == 1)
      minLevel[steps[i].inputLevel + 1] = true, ++hooks;
  size_t src = 0;
  for(uint32_t l = 0; l + 1 < levels; ++l)
  {
    const uint32_t sw = dim(w, l), sh = dim(h, l), dw = dim(w, l + 1), dh = dim(h, l + 1);
    const int      kx = taps(sw), ky = taps(sh);
    const size_t   dst = chain.size();
    chain.resize(dst + size_t(dw) * dh);
    for(uint32_t y = 0; y < dh; ++y)
      for(uint32_t x = 0; x < dw; ++x)
      {
        float m = minLevel[l + 1] ? INFINITY : -INFINITY;
        for(int j = 0; j < ky; ++j)
          for(int i = 0; i < kx; ++i)
          {
            const float v = chain[src + size_t(2 * y + j) * sw + (2 * x + i)];
            m             = minLevel[l + 1] ? fminf(m, v) : fmaxf(m, v);
          }
        chain[dst + size_t(y) * dw + x] = m;
      }
    src = dst;
  }
  float* dev = nullptr;
  CK(cudaMalloc(&dev, chain.size() * 4));
  CK(cudaMemset(dev, 0xFF, chain.size() * 4));
  CK(cudaMemcpy(dev, chain.data(), l0Texels * 4, cudaMemcpyHostToDevice));
  nvpyrDispatchDesc d;
  memset(&d, 0, sizeof(d));
  d.structSize       = sizeof(d);
  d.flags            = flags;
  d.extent           = {w, h};
  d.base             = dev;
  d.fastDivisibility = div;
  d.fastMaxLevels    = maxLevels;
  const nvpyrStatus st = nvpyr::dispatch<MinFirst>(d);
  CK(cudaDeviceSynchronize());
  std::vector<float> got(chain.size());
  CK(cudaMemcpy(got.data(), dev, got.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for(size_t i = l0Texels; i < chain.size(); ++i)
    bad += memcmp(&got[i], &chain[i], 4) != 0;
  printf("MinFirst  %5ux%-5u %-34s status %d, %zu fast dispatches use the hook, %zu of %zu texels differ\n", w, h, what, int(st),
         hooks, bad, chain.size() - l0Texels);
  failures += st != NVPYR_SUCCESS || bad != 0;
  cudaFree(dev);
}

static void checkRgba32f(uint32_t w, uint32_t h)
{
  const uint32_t levels = nvpyr::levelCountFor(w, h);
  size_t         texels = 0;
  for(uint32_t l = 0; l < levels; ++l)
    texels += size_t(dim(w, l)) * dim(h, l);
  std::vector<float> l0(size_t(w) * h * 4);
  uint32_t           s = 777u + w + 7u * h;
  for(float& v : l0)
  {
    s = s * 1664525u + 1013904223u;
    v = float(s >> 8) * (1.0f / 16777216.0f);
  }
  float *a = nullptr, *b = nullptr;
  CK(cudaMalloc(&a, texels * 16));
  CK(cudaMalloc(&b, texels * 16));
  CK(cudaMemset(a, 0, texels * 16));
  CK(cudaMemset(b, 0, texels * 16));
  CK(cudaMemcpy(a, l0.data(), l0.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b, l0.data(), l0.size() * 4, cudaMemcpyHostToDevice));
  nvpyrDispatchDesc d;
  memset(&d, 0, sizeof(d));
  d.structSize = sizeof(d);
  d.format     = NVPYR_FORMAT_RGBA32F;
  d.extent     = {w, h};
  d.base       = a;
  const nvpyrStatus s1 = nvpyr::dispatch<MyRgba32f>(d);
  d.base               = b;
  const nvpyrStatus s2 = nvpyrDispatchEx(&d);  // the library's own RGBA32F instance
  CK(cudaDeviceSynchronize());
  std::vector<float> ga(texels * 4), gb(texels * 4);
  CK(cudaMemcpy(ga.data(), a, texels * 16, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gb.data(), b, texels * 16, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for(size_t i = 0; i < ga.size(); ++i)
    bad += memcmp(&ga[i], &gb[i], 4) != 0;
  printf("MyRgba32f %5ux%-5u user set vs library RGBA32F          status %d/%d, %zu of %zu floats differ\n", w, h, int(s1),
         int(s2), bad, ga.size());
  failures += s1 != NVPYR_SUCCESS || s2 != NVPYR_SUCCESS || bad != 0;
  cudaFree(a);
  cudaFree(b);
}

int main()
{
  const uint32_t sizes[][2] = {{1024, 512}, {256, 256}, {1000, 700}, {333, 97}, {2052, 1028}, {1080, 4096}, {1, 37}, {64, 1}, {5, 5}};
  for(const auto& sz : sizes)
  {
    checkDepth(sz[0], sz[1], "default dispatchers", 0, 0, 0, nullptr);
    checkDepth(sz[0], sz[1], "no fast pipeline", NVPYR_FLAG_FORCE_GENERAL, 0, 0, nullptr);
    checkDepth(sz[0], sz[1], "fast <4, 3>", 0, 4, 3, nullptr);
    checkDepth(sz[0], sz[1], "fast <2, 5> + one-level general", 0, 2, 5, oneLevelGeneral);
  }
  checkDepth(1024, 1024, "levelCount 4 (partial chain)", 0, 0, 0, nullptr, 4);
  {
    float* dev = nullptr;
    CK(cudaMalloc(&dev, 4 * 333 * 201 * 2));
    CK(cudaMemset(dev, 0, 4 * 333 * 201 * 2));
    nvpyrDispatchDesc d;
    memset(&d, 0, sizeof(d));
    d.structSize = sizeof(d);
    d.extent     = {333, 201};
    d.base       = dev;
    const nvpyrStatus a = nvpyr::dispatch<DepthMax>(d, nullptr, fillsNothing);
    const nvpyrStatus b = nvpyr::dispatch<DepthMax>(d, nullptr, fillsTooMuch);
    const nvpyrStatus c = nvpyr::dispatch<DepthMax>(d, nullptr, nullptr, fastOnOddSizes);
    CK(cudaDeviceSynchronize());
    std::vector<float> got(333 * 201 * 2);
    CK(cudaMemcpy(got.data(), dev, got.size() * 4, cudaMemcpyDeviceToHost));
    size_t written = 0;
    for(float v : got)
      written += v != 0.0f;
    const bool ok = a == NVPYR_ERROR_INVALID_VALUE && b == NVPYR_ERROR_INVALID_VALUE && c == NVPYR_ERROR_INVALID_VALUE && written == 0;
    printf("bad dispatchers: status %d %d %d, %zu texels written -> %s\n", int(a), int(b), int(c), written, ok ? "rejected" : "NOT rejected");
    failures += !ok;
    cudaFree(dev);
  }
  const uint32_t msizes[][2] = {{64, 64}, {256, 256}, {1024, 512}, {260, 260}, {2052, 1028}, {2, 2}, {333, 97}};
  for(const auto& sz : msizes)
  {
    checkMinFirst(sz[0], sz[1], "LOAD_REDUCE4 hook, default", 0, 0, 0);
    checkMinFirst(sz[0], sz[1], "LOAD_REDUCE4 hook, fast <2, 5>", 0, 2, 5);
    checkMinFirst(sz[0], sz[1], "no fast pipeline: hook unused", NVPYR_FLAG_FORCE_GENERAL, 0, 0);
  }
  const uint32_t fsizes[][2] = {{512, 256}, {255, 129}, {260, 260}, {96, 1000}};
  for(const auto& sz : fsizes)
    checkRgba32f(sz[0], sz[1]);
  printf(failures ? "FAILED (%d)\n" : "custom functor sets ok\n", failures);
  return failures ? 1 : 0;
}
