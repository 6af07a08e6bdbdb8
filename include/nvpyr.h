/*
 * nvpyr.h -- C ABI of the B200-native mip-pyramid generator.
 *
 * Drop-in boundary for ONE path of nvpro-samples/vk_compute_mipmaps: "fill mip
 * levels 1..N-1 of an image from level 0" -- what the reference does with
 *     nvproCmdPyramidDispatch(cmdBuf, pipelines, baseWidth, baseHeight, mipLevels)
 *     (reference nvpro_pyramid/nvpro_pyramid_dispatch.hpp:54-59, :109-188, :294-304)
 * plus the two compute shaders it dispatches (nvpro_pyramid/nvpro_pyramid.glsl with
 * nvpro_pyramid/srgba8_mipmap_preamble.glsl).  Plain C types only; the stream
 * argument is a CUstream/cudaStream_t passed as an opaque pointer.
 *
 * Differences from the Vulkan original, by construction:
 *   - commands are ENQUEUED on a CUDA stream instead of recorded into a
 *     VkCommandBuffer; stream order replaces the inter-dispatch pipeline barriers
 *     (dispatch.hpp:180-186);
 *   - the image is a LINEAR buffer in device memory, packed exactly like the
 *     reference's MipmapStorage / staging buffer (include/mipmap_storage.hpp:35-39,
 *     :53-76; include/scoped_image.hpp:436-453): level i starts at texel offset
 *     sum_{j<i} W_j*H_j, rows are tight, W_i = max(1, W >> i), texel = R,G,B,A;
 *   - errors are returned, never asserted (dispatch.hpp:169,172 assert).
 *
 * Every entry point is reentrant and uses the CURRENT CUDA device.
 */
#ifndef NVPYR_H_
#define NVPYR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define NVPYR_API __declspec(dllexport)
#else
#define NVPYR_API __attribute__((visibility("default")))
#endif

#define NVPYR_VERSION 100 /* 0.1.0 */
#define NVPYR_MAX_LEVELS 32u
#define NVPYR_MAX_STEPS 40u

typedef enum nvpyrStatus
{
  NVPYR_SUCCESS             = 0,
  NVPYR_ERROR_INVALID_VALUE = 1, /* null pointer, zero extent, levelCount > max, misaligned base */
  NVPYR_ERROR_UNSUPPORTED   = 2, /* unknown format / flag, device is not sm_100 */
  NVPYR_ERROR_CUDA          = 3, /* a CUDA call failed; see nvpyrGetLastCudaError */
  NVPYR_ERROR_OUT_OF_MEMORY = 4,
  NVPYR_ERROR_IO            = 5 /* a file could not be opened, read or written, or is truncated */
} nvpyrStatus;

typedef struct nvpyrExtent2D
{
  uint32_t width;
  uint32_t height;
} nvpyrExtent2D;

/* Texel formats = shipped instances of the load/reduce/store functor set
 * (the reference's NVPRO_PYRAMID_* macro set, nvpro_pyramid.glsl:27-120). */
typedef enum nvpyrFormat
{
  NVPYR_FORMAT_SRGBA8  = 0, /* srgba8_mipmap_preamble.glsl: decode sRGB, average in float32, encode */
  NVPYR_FORMAT_RGBA32F = 1  /* identity load/store, same reductions; 16 bytes per texel */
} nvpyrFormat;

typedef enum nvpyrFlags
{
  NVPYR_FLAG_NONE = 0,
  /* Never use the fast pipeline (== NvproPyramidPipelines::fastPipeline = VK_NULL_HANDLE,
   * minimal_app's -force-no-fast-pipeline, minimal_mipmaps.cpp:42-43,357-360). */
  NVPYR_FLAG_FORCE_GENERAL = 1u << 0,
  /* Run the premultiply-alpha pre-pass of include/scoped_image.hpp:233-255 on level 0
   * (in place) before generating. sRGBA8 only. */
  NVPYR_FLAG_PREMULTIPLY_ALPHA = 1u << 1,
  /* The reference's optional F16_SHARED build of the sRGBA8 shaders (srgba8_mipmap_preamble.glsl:103-108,
   * demo_app alternative "f16Shared"): values that pass through shared memory inside a dispatch are rounded
   * to IEEE binary16.  Non-default and lossy; sRGBA8 only; runs on the functor-template kernels. */
  NVPYR_FLAG_F16_SHARED = 1u << 2,
  /* The reference's optional SRGB_SHARED build (srgba8_mipmap_preamble.glsl:60-101, demo_app alternative
   * "srgbShared"): values that pass through shared memory inside a dispatch are packed to 8-bit sRGB and unpacked
   * again.  Non-default and lossy; sRGBA8 only; mutually exclusive with F16_SHARED; functor-template kernels. */
  NVPYR_FLAG_SRGB_SHARED = 1u << 3,
  /* demo_app's "generalblit" alternative (demo_app/pipeline_alternative.cpp:16, mipmap_pipelines.cpp:350-453): levels
   * the fast pipeline does not take are filled ONE AT A TIME by a linear-filter blit of the previous level
   * (vkCmdBlitImage with VK_FILTER_LINEAR, mipmap_pipelines.cpp:418-426) instead of by the general pipeline.  Together
   * with NVPYR_FLAG_FORCE_GENERAL it is the "blit" alternative (every level blitted).  Cheaper and, on odd sizes, WRONG
   * in the reference's own words ("may trade correctness for performance"): a blit samples two source texels per axis
   * whatever the scale, so a 5 -> 2 reduction ignores a fifth of the image.  Vulkan leaves a blit's filtering arithmetic
   * to the implementation; ours is pinned in DESIGN.md section 4.10, restated by the oracle, and reproduces the worst
   * deltas the reference recorded for these alternatives on its 13 test images (demo_app/rtx3090.json) exactly on 9 and
   * within 3 code values on all.  Not combinable with the shared-type flags. */
  NVPYR_FLAG_GENERAL_BLIT = 1u << 4
} nvpyrFlags;

typedef struct CUstream_st* nvpyrStream; /* == cudaStream_t == CUstream */

/* ------------------------------------------------------------------ layout */

/* floor(log2(max(W,H))) + 1 -- the default of dispatch.hpp:122-132. 0 if an extent is 0. */
NVPYR_API uint32_t nvpyrGetLevelCount(nvpyrExtent2D extent);
/* max(1, W >> level) x max(1, H >> level); mipmap_storage.hpp:67-69. */
NVPYR_API nvpyrStatus nvpyrGetLevelExtent(nvpyrExtent2D extent, uint32_t level, nvpyrExtent2D* out);
/* Texel offset of `level` in the packed chain; mipmap_storage.hpp:56-73. */
NVPYR_API nvpyrStatus nvpyrGetLevelOffsetTexels(nvpyrExtent2D extent, uint32_t level, uint64_t* out);
/* Bytes of a packed chain of levelCount levels (0 = all). */
NVPYR_API nvpyrStatus nvpyrGetChainBytes(nvpyrExtent2D extent, uint32_t levelCount, nvpyrFormat format,
                                         uint64_t* out);

/* -------------------------------------------------------------------- plan */

/* One dispatch of the reference-equivalent schedule (what nvproCmdPyramidDispatch would
 * record): defines the float "carry groups" that are observable in the output bits. */
typedef struct nvpyrPlanStep
{
  uint32_t pipeline;     /* 1 = fast (nvproPyramidDefaultFastDispatcher, dispatch.hpp:195-242),
                            0 = general (nvproPyramidDefaultGeneralDispatcher, :247-292) */
  uint32_t inputLevel;   /* NvproPyramidState::currentLevel */
  uint32_t levelCount;   /* levels filled by the dispatch */
  uint32_t srcWidth;     /* NvproPyramidState::currentX */
  uint32_t srcHeight;    /* NvproPyramidState::currentY */
  uint32_t workgroups;   /* groupCountX the reference passes to vkCmdDispatch */
  uint32_t pushConstant; /* inputLevel << 5 | levelCount (dispatch.hpp:77, glsl:157-162) */
  uint32_t bindPipeline; /* reference records vkCmdBindPipeline before this dispatch */
  uint32_t barrierAfter; /* reference records a pipeline barrier after this dispatch */
} nvpyrPlanStep;

typedef struct nvpyrPlanOptions
{
  uint32_t flags;                /* NVPYR_FLAG_FORCE_GENERAL and NVPYR_FLAG_GENERAL_BLIT (pipeline 0 steps = one blit each,
                                    workgroups 0) honoured */
  uint32_t fastDivisibility;     /* template arg DivisibilityRequirement; 0 = default 4 */
  uint32_t fastMaxLevels;        /* template arg MaxLevels (<= 6);        0 = default 6 */
} nvpyrPlanOptions;

/* Host only, no CUDA.  options may be NULL.  *count receives the number of steps. */
NVPYR_API nvpyrStatus nvpyrGetPlan(nvpyrExtent2D extent, uint32_t levelCount, const nvpyrPlanOptions* options,
                                   nvpyrPlanStep* steps, uint32_t maxSteps, uint32_t* count);

/* ---------------------------------------------------------------- dispatch */

/* Replaces nvproCmdPyramidDispatch(cmdBuf, pipelines, baseWidth, baseHeight, mipLevels)
 * (dispatch.hpp:54-59) for an sRGBA8 image with both pipelines available.
 *   srcLevel0  device pointer to the packed chain (level 0 filled, 16-byte aligned);
 *              levels 1..levelCount-1 are written behind it
 *   levelCount 0 = all levels (dispatch.hpp:122-132)
 * Asynchronous on `stream`. */
NVPYR_API nvpyrStatus nvpyrDispatch(void* srcLevel0, uint32_t levelCount, nvpyrExtent2D extent, nvpyrStream stream);

typedef struct nvpyrDispatchDesc
{
  uint32_t      structSize; /* sizeof(nvpyrDispatchDesc) */
  nvpyrFormat   format;
  uint32_t      flags; /* nvpyrFlags */
  nvpyrExtent2D extent;
  uint32_t      levelCount; /* 0 = all */
  void*         base;       /* packed chain; may be NULL if levels[] is given */
  /* Optional override of the packed layout (the analogue of the reference's per-level
   * image views, scoped_image.hpp:331-344): levels[i] != NULL gives the address of
   * level i, rowPitchBytes[i] its pitch (0 = tight). */
  void*       levels[NVPYR_MAX_LEVELS];
  uint32_t    rowPitchBytes[NVPYR_MAX_LEVELS];
  uint32_t    fastDivisibility; /* 0 = 4 */
  uint32_t    fastMaxLevels;    /* 0 = 6 */
  nvpyrStream stream;
} nvpyrDispatchDesc;

NVPYR_API nvpyrStatus nvpyrDispatchEx(const nvpyrDispatchDesc* desc);

/* The 7-argument overload of nvproCmdPyramidDispatch (dispatch.hpp:109-116): the caller supplies the dispatcher
 * callbacks that decide, dispatch by dispatch, which pipeline runs and how many levels it fills.
 *   nvpyrPyramidState = NvproPyramidState (dispatch.hpp:63-75); remainingLevels is never 0 when a callback is called.
 *   nvpyrDispatcher   = nvpro_pyramid_dispatcher_t (dispatch.hpp:99-104) without the Vulkan arguments: returns the
 *                       number of levels filled from `state` (the fast dispatcher may return 0 = "not eligible",
 *                       the general one never); it may fill step->workgroups / step->pushConstant (informational).
 * general == NULL: nvproPyramidDefaultGeneralDispatcher.  fast == NULL: the default fast dispatcher selected by
 * desc->fastDivisibility / fastMaxLevels, or none with NVPYR_FLAG_FORCE_GENERAL.  Limits of the kernels, checked
 * before anything is enqueued (NVPYR_ERROR_INVALID_VALUE): a fast dispatch fills 1..6 levels and both edges of its
 * input level are multiples of 2^levels; a general dispatch fills 1 or 2 levels; a dispatcher that fills 0 levels
 * (general) or more than remain is the reference's assert (dispatch.hpp:169,172).  Callbacks run on the calling
 * thread, inside this call, and may be called more than once per state (validation, then execution): like the
 * reference's dispatchers they must be pure functions of the state. */
typedef struct nvpyrPyramidState
{
  uint32_t currentLevel, remainingLevels, currentX, currentY;
} nvpyrPyramidState;
typedef uint32_t (*nvpyrDispatcher)(const nvpyrPyramidState* state, nvpyrPlanStep* step, void* userData);
NVPYR_API nvpyrStatus nvpyrDispatchWithDispatchers(const nvpyrDispatchDesc* desc, nvpyrDispatcher general,
                                                   nvpyrDispatcher fast, void* userData);

/* Independent images (no data crosses images or GPUs); descs[i].stream is honoured. */
NVPYR_API nvpyrStatus nvpyrDispatchBatch(const nvpyrDispatchDesc* descs, uint32_t count);

/* Premultiply-alpha pre-pass alone (scoped_image.hpp:233-255), in place when in == out. */
NVPYR_API nvpyrStatus nvpyrPremultiplyAlpha(const void* in, void* out, uint64_t texels, nvpyrStream stream);

/* Whole round trip with HOST buffers, the shape of minimal_app (minimal_mipmaps.cpp:59-241):
 * upload level 0, generate, download the packed chain.  Synchronous.  hostChain receives all
 * levelCount levels (level 0 included, as the reference's download does). */
NVPYR_API nvpyrStatus nvpyrGenerateHost(const void* hostLevel0, void* hostChain, nvpyrExtent2D extent,
                                        uint32_t levelCount, nvpyrFormat format, uint32_t flags);

/* ------------------------------------------------- Vulkan external memory  */

typedef struct nvpyrExternalMemory_t* nvpyrExternalMemory;
/* Imports a VK_KHR_external_memory_fd (OPAQUE_FD) allocation backing a linear VkBuffer and maps
 * `size` bytes at `offset`; the fd is owned by the library on success (cudaImportExternalMemory). */
NVPYR_API nvpyrStatus nvpyrImportExternalMemoryFd(int fd, uint64_t allocationSize, uint64_t offset, uint64_t size,
                                                  nvpyrExternalMemory* outHandle, void** outDevicePtr);
NVPYR_API nvpyrStatus nvpyrReleaseExternalMemory(nvpyrExternalMemory handle);

/* ------------------------------------------------------------- image files */

/* Replaces stbi_write_tga(filename, w, h, 4, data) as the reference calls it (mipmap_storage.hpp:462):
 * 32-bit run-length-encoded TGA, rows bottom-up, texels B,G,R,A.  Host memory, synchronous. */
NVPYR_API nvpyrStatus nvpyrWriteTga(const char* filename, const void* rgba8, nvpyrExtent2D extent);
/* Name of the file that holds `level`: "image.name.tga" -> "image.name.<level>.tga", level 0 keeps the
 * base name (mipmap_storage.hpp:447-460). */
NVPYR_API nvpyrStatus nvpyrGetLevelFilename(const char* baseFilename, uint32_t level, char* out, size_t outSize);
/* Replaces writeMipmapsTga(mips, pBaseFilename) (mipmap_storage.hpp:441-479): one TGA per level of a packed
 * sRGBA8 host chain (levelCount 0 = all). */
NVPYR_API nvpyrStatus nvpyrWriteChainTga(const void* hostChain, nvpyrExtent2D extent, uint32_t levelCount,
                                         const char* baseFilename);
/* Stands where the reference calls stbi_load(filename, &w, &h, &n, 4) (scoped_image.hpp:217-218): returns
 * top-down R,G,B,A texels (A = 255 if the file has no alpha) in a buffer to be released with nvpyrFree.
 * Formats: TGA (true colour or grey, raw or RLE, 8/24/32 bits) and binary PGM/PPM; no JPEG/PNG decoder. */
NVPYR_API nvpyrStatus nvpyrReadImage(const char* filename, void** rgba8, nvpyrExtent2D* extent);
NVPYR_API void        nvpyrFree(void* p);

/* -------------------------------------------------------------------- misc */

NVPYR_API const char* nvpyrGetErrorString(nvpyrStatus status);
/* Last CUDA error code seen by this thread inside the library (cudaError_t value), 0 if none. */
NVPYR_API int nvpyrGetLastCudaError(void);
/* Number of kernels the library has launched in this process (for launch accounting). */
NVPYR_API uint64_t nvpyrGetLaunchCount(void);
/* Host-only self-test of the tuned fast kernel's sRGB encode table: every float32 pattern the kernel can present
 * (linearFromSrgb(1)/4 .. 1.0 and zero, ~110 M values) goes through the kernel's own look-up arithmetic and is compared
 * with srgbFromLinear restated as "number of pinned thresholds <= x" (reference shaders/srgb.h:30-41).  Returns the
 * number of mismatches: 0.  Needs no GPU; takes about a second. */
NVPYR_API uint64_t nvpyrSelfTestEncodeTable(void);
/* Creates the library's per-device state (tables, counters) for the CURRENT device now instead of inside the first
 * dispatch.  Call it before capturing nvpyrDispatch* into a CUDA graph: the first call on a device allocates and
 * copies, which a stream capture forbids.  NVPYR_ERROR_UNSUPPORTED on anything but an sm_100 device. */
NVPYR_API nvpyrStatus nvpyrInit(void);
/* Frees per-device cached tables. */
NVPYR_API nvpyrStatus nvpyrShutdown(void);
NVPYR_API uint32_t    nvpyrGetVersion(void);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* NVPYR_H_ */
