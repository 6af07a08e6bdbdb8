"""ctypes binding of libnvpyr.so (C ABI in include/nvpyr.h).

The product path has NO fallback: if the CUDA library has not been built
(``python -c "import __graft_entry__ as g; g.build()"``) importing this module
raises.  Nothing here touches ``oracle/``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NVPYR_LIB_PATH: load another build of the same ABI (A/B timing of kernel variants).
LIB_PATH = os.environ.get("NVPYR_LIB_PATH") or os.path.join(_HERE, "libnvpyr.so")

NVPYR_MAX_LEVELS = 32
NVPYR_MAX_STEPS = 40

SUCCESS, ERROR_INVALID_VALUE, ERROR_UNSUPPORTED, ERROR_CUDA, ERROR_OUT_OF_MEMORY, ERROR_IO = range(6)
FORMAT_SRGBA8, FORMAT_RGBA32F = 0, 1
FLAG_NONE, FLAG_FORCE_GENERAL, FLAG_PREMULTIPLY_ALPHA, FLAG_F16_SHARED, FLAG_SRGB_SHARED = 0, 1, 2, 4, 8
FLAG_GENERAL_BLIT = 16


class Extent2D(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32)]


class PlanStep(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "pipeline", "inputLevel", "levelCount", "srcWidth", "srcHeight",
        "workgroups", "pushConstant", "bindPipeline", "barrierAfter")]


class PyramidState(C.Structure):
    """NvproPyramidState (nvpro_pyramid_dispatch.hpp:63-75)."""
    _fields_ = [(n, C.c_uint32) for n in ("currentLevel", "remainingLevels", "currentX", "currentY")]


class PlanOptions(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("fastDivisibility", C.c_uint32), ("fastMaxLevels", C.c_uint32)]


class DispatchDesc(C.Structure):
    _fields_ = [
        ("structSize", C.c_uint32),
        ("format", C.c_int),
        ("flags", C.c_uint32),
        ("extent", Extent2D),
        ("levelCount", C.c_uint32),
        ("base", C.c_void_p),
        ("levels", C.c_void_p * NVPYR_MAX_LEVELS),
        ("rowPitchBytes", C.c_uint32 * NVPYR_MAX_LEVELS),
        ("fastDivisibility", C.c_uint32),
        ("fastMaxLevels", C.c_uint32),
        ("stream", C.c_void_p),
    ]


# nvpro_pyramid_dispatcher_t without the Vulkan arguments: levels filled = f(state*, step*, userData)
Dispatcher = C.CFUNCTYPE(C.c_uint32, C.POINTER(PyramidState), C.POINTER(PlanStep), C.c_void_p)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    u32, u64, vp, st = C.c_uint32, C.c_uint64, C.c_void_p, C.c_int
    sigs = {
        "nvpyrGetLevelCount": (u32, [Extent2D]),
        "nvpyrGetLevelExtent": (st, [Extent2D, u32, C.POINTER(Extent2D)]),
        "nvpyrGetLevelOffsetTexels": (st, [Extent2D, u32, C.POINTER(u64)]),
        "nvpyrGetChainBytes": (st, [Extent2D, u32, C.c_int, C.POINTER(u64)]),
        "nvpyrGetPlan": (st, [Extent2D, u32, C.POINTER(PlanOptions), C.POINTER(PlanStep), u32, C.POINTER(u32)]),
        "nvpyrDispatch": (st, [vp, u32, Extent2D, vp]),
        "nvpyrDispatchEx": (st, [C.POINTER(DispatchDesc)]),
        "nvpyrDispatchWithDispatchers": (st, [C.POINTER(DispatchDesc), Dispatcher, Dispatcher, vp]),
        "nvpyrDispatchBatch": (st, [C.POINTER(DispatchDesc), u32]),
        "nvpyrPremultiplyAlpha": (st, [vp, vp, u64, vp]),
        "nvpyrGenerateHost": (st, [vp, vp, Extent2D, u32, C.c_int, u32]),
        "nvpyrImportExternalMemoryFd": (st, [C.c_int, u64, u64, u64, C.POINTER(vp), C.POINTER(vp)]),
        "nvpyrReleaseExternalMemory": (st, [vp]),
        "nvpyrWriteTga": (st, [C.c_char_p, vp, Extent2D]),
        "nvpyrGetLevelFilename": (st, [C.c_char_p, u32, C.c_char_p, C.c_size_t]),
        "nvpyrWriteChainTga": (st, [vp, Extent2D, u32, C.c_char_p]),
        "nvpyrReadImage": (st, [C.c_char_p, C.POINTER(vp), C.POINTER(Extent2D)]),
        "nvpyrFree": (None, [vp]),
        "nvpyrGetErrorString": (C.c_char_p, [st]),
        "nvpyrGetLastCudaError": (C.c_int, []),
        "nvpyrGetLaunchCount": (u64, []),
        "nvpyrSelfTestEncodeTable": (u64, []),
        "nvpyrInit": (st, []),
        "nvpyrShutdown": (st, []),
        "nvpyrGetVersion": (u32, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


class NvpyrError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib.nvpyrGetErrorString(status).decode()
        if status == ERROR_CUDA:
            msg += f" (cudaError {lib.nvpyrGetLastCudaError()})"
        super().__init__(f"{where}: {msg}")


def check(status, where):
    if status != SUCCESS:
        raise NvpyrError(status, where)
