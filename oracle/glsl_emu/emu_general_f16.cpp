// The reference's GENERAL pipeline shader built with F16_SHARED = 1 ("f16SharedGeneral"): emu_general.cpp with
// the macro set and its entry points renamed.
#define F16_SHARED 1
#define emu_general emu_general_f16
#define emuGeneralMain emuGeneralMainF16
#define emuGeneralSetImage emuGeneralSetImageF16
#include "emu_general.cpp"
