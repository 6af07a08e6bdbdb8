// The reference's GENERAL pipeline shader, compiled as C++ (see glsl_shim.hpp).  Mirrors
// nvpro_pyramid/srgba8_mipmap_general_pipeline.comp: same define, same two includes, same main().
// The preamble's resource declarations (sampler2D srgbTex; uimage2D imageMipLevels[16];) become
// namespace-scope objects of this translation unit; the driver fills them through emuGeneralSetImage.
#include "glsl_shim.hpp"
namespace emu_general {
#define float Float
#define NVPRO_PYRAMID_IS_FAST_PIPELINE 0
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
void setImage(const uimage2D* levels)
{
  for(int i = 0; i < 16; ++i)
    imageMipLevels[i] = levels[i];
  srgbTex.levels = imageMipLevels;
}
}  // namespace emu_general
void emuGeneralMain() { emu_general::mainEntry(); }
void emuGeneralSetImage(const uimage2D* levels) { emu_general::setImage(levels); }
