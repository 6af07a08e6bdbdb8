// The reference's FAST pipeline shader, compiled as C++ (see glsl_shim.hpp).  Mirrors
// nvpro_pyramid/srgba8_mipmap_fast_pipeline.comp: same define, same two includes, same main().
// The preamble's resource declarations (sampler2D srgbTex; uimage2D imageMipLevels[16];) become
// namespace-scope objects of this translation unit; the driver fills them through emuFastSetImage.
#include "glsl_shim.hpp"
namespace emu_fast {
#define float Float
#define NVPRO_PYRAMID_IS_FAST_PIPELINE 1
#if defined(SRGB_SHARED) && SRGB_SHARED
#undef in
#define in  // parameter qualifiers of srgbUnpack / srgbPack (preamble.glsl:81,91)
#define out
#endif
#include "nvpro_pyramid/srgba8_mipmap_preamble.glsl"
#if defined(SRGB_SHARED) && SRGB_SHARED
#undef in
#undef out
#define in int emuLocalSizeDecl_
#include "emu_srgb_shared_glue.inc"
#endif
#include "nvpro_pyramid/nvpro_pyramid.glsl"
#undef float
void mainEntry() { nvproPyramidMain(); }
uint encode(Float x) { return srgbFromLinear(x); }
#if defined(SRGB_SHARED) && SRGB_SHARED
// SHARED_STORE then SHARED_LOAD of (x, x, x, x): returns the red component, *code receives the packed byte
float sharedRoundTrip(float x, uint* code)
{
  uint packed;
  const Float fx(x);
  vec4        v(fx, fx, fx, fx), back;
  NVPRO_PYRAMID_SHARED_STORE(packed, v);
  NVPRO_PYRAMID_SHARED_LOAD(packed, back);
  *code = packed & 255u;
  return back.x.v;
}
#endif
void setImage(const uimage2D* levels)
{
  for(int i = 0; i < 16; ++i)
    imageMipLevels[i] = levels[i];
  srgbTex.levels = imageMipLevels;
}
}  // namespace emu_fast
void emuFastMain() { emu_fast::mainEntry(); }
void emuFastSetImage(const uimage2D* levels) { emu_fast::setImage(levels); }
uint emuGlslSrgbFromLinear(float x) { return emu_fast::encode(Float(x)); }
#if defined(SRGB_SHARED) && SRGB_SHARED
extern "C" float emu_srgb_shared_round_trip(float x, uint* code) { return emu_fast::sharedRoundTrip(x, code); }
#endif
