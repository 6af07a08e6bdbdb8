// The reference's FAST pipeline shader built with SRGB_SHARED = 1 (srgba8_mipmap_preamble.glsl:60-101; the demo's
// "srgbShared" alternative, demo_app/mipmap_pipelines.cpp:109-112): the same translation unit as emu_fast.cpp with
// the macro set and its entry points renamed.
#define SRGB_SHARED 1
#define emu_fast emu_fast_srgb
#define emuFastMain emuFastMainSrgb
#define emuFastSetImage emuFastSetImageSrgb
#define emuGlslSrgbFromLinear emuGlslSrgbFromLinearSrgb
#include "emu_fast.cpp"
