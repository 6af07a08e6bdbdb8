// The reference's GENERAL pipeline shader built with SRGB_SHARED = 1 ("srgbSharedGeneral"): emu_general.cpp with
// the macro set and its entry points renamed.
#define SRGB_SHARED 1
#define emu_general emu_general_srgb
#define emuGeneralMain emuGeneralMainSrgb
#define emuGeneralSetImage emuGeneralSetImageSrgb
#include "emu_general.cpp"
