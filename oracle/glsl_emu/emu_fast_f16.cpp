// The reference's FAST pipeline shader built with F16_SHARED = 1 (srgba8_mipmap_preamble.glsl:103-108; the
// demo's "f16Shared" alternative, demo_app/mipmap_pipelines.cpp:109-116): the same translation unit as
// emu_fast.cpp with the macro set and its entry points renamed.
#define F16_SHARED 1
#define emu_fast emu_fast_f16
#define emuFastMain emuFastMainF16
#define emuFastSetImage emuFastSetImageF16
#define emuGlslSrgbFromLinear emuGlslSrgbFromLinearF16
#include "emu_fast.cpp"
