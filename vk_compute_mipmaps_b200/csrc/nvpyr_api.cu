// nvpyr_api.cu -- C ABI (include/nvpyr.h) over the planner and the sm_100a kernels.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <atomic>
#include <deque>
#include <map>
#include <type_traits>
#include <utility>
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/nvpyr.h"
#include "nvpyr_fast_srgba8.cuh"
#include "nvpyr_general_srgba8.cuh"
#include "nvpyr_kernels.cuh"
#include "nvpyr_plan.hpp"
#include "srgb_tables.inc"

namespace nvpyr {
namespace {

thread_local int      g_lastCudaError = 0;
// NVPYR_GENERIC_FAST=1 routes sRGBA8 through the generic functor kernel (fastKernel<Srgba8, M>)
// instead of the tuned one; both must produce the same bits (tests/test_gpu_parity.py).
const bool g_forceGenericFast = [] {
  const char* e = getenv("NVPYR_GENERIC_FAST");
  return e != nullptr && e[0] == '1';
}();
// NVPYR_NO_TAIL_FUSION=1: one launch per reference dispatch (no cooperative tail kernel).
const bool g_noTailFusion = [] {
  const char* e = getenv("NVPYR_NO_TAIL_FUSION");
  return e != nullptr && e[0] == '1';
}();
// NVPYR_NO_PDL=1: plain stream-ordered launches (no programmatic dependent launch), for A/B timing.
const bool g_noPdl = [] {
  const char* e = getenv("NVPYR_NO_PDL");
  return e != nullptr && e[0] == '1';
}();
std::atomic<uint64_t> g_launchCount{0};

// Every kernel goes out with the programmatic-stream-serialization attribute: its CTAs may be scheduled
// (and set up their shared-memory tables) while the previous kernel of the stream is still finishing;
// the kernels call griddepcontrol.wait before touching any level (nvpyr_functors.cuh).
template <class... KArgs, class... Args>
cudaError_t launchKernel(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args)
{
  cudaLaunchConfig_t cfg{};
  cfg.gridDim          = dim3(unsigned(grid));
  cfg.blockDim         = dim3(unsigned(block));
  cfg.dynamicSmemBytes = smem;
  cfg.stream           = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs                                          = attr;
  cfg.numAttrs                                       = g_noPdl ? 0u : 1u;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

#define NVPYR_CUDA(call)                                                                                          \
  do                                                                                                              \
  {                                                                                                               \
    cudaError_t e_ = (call);                                                                                      \
    if(e_ != cudaSuccess)                                                                                         \
    {                                                                                                             \
      g_lastCudaError = int(e_);                                                                                  \
      return e_ == cudaErrorMemoryAllocation ? NVPYR_ERROR_OUT_OF_MEMORY : NVPYR_ERROR_CUDA;                      \
    }                                                                                                             \
  } while(0)

// ---------------------------------------------------------------- host tables
float bitsToFloat(uint32_t b)
{
  float f;
  memcpy(&f, &b, 4);
  return f;
}

// Builds one encode bucket table of nvpyr_functors.cuh (keys = float bits >> shift) from the 255 pinned
// thresholds.  Returns false if a bucket would hold two thresholds (cannot happen with the pinned data;
// checked anyway so a regenerated table cannot silently break the encode).
bool buildEncodeTable(uint32_t* out, uint32_t shift, uint32_t entriesPadded)
{
  const uint32_t* thr = NVPYR_SRGB_ENCODE_THRESHOLD_BITS;  // thr[c-1] = first bits with code >= c
  if(thr[0] <= kEncMinBits || thr[254] > kEncMaxBits)
    return false;
  const uint32_t minKey = kEncMinBits >> shift, maxKey = kEncMaxBits >> shift;
  uint32_t       code   = 0;  // number of thresholds <= bucket start
  for(uint32_t key = minKey; key <= maxKey; ++key)
  {
    const uint32_t lo = key << shift, hi = lo + (1u << shift);  // [lo, hi)
    while(code < 255 && thr[code] <= lo)
      ++code;
    uint32_t entry = code << 16;
    if(code < 255 && thr[code] < hi)
    {
      if(code + 1 < 255 && thr[code + 1] < hi)
        return false;
      entry += 0x10000u - (thr[code] - lo);
    }
    out[key - minKey] = entry - lo;  // pre-biased: kernel adds the full bit pattern
  }
  for(uint32_t i = maxKey - minKey + 1; i < entriesPadded; ++i)
    out[i] = 0;
  return true;
}

// Row of the tuned fast kernel's encode table (nvpyr_functors.cuh): key of RN(x + kRowEncC).  IEEE float32 addition,
// the same value the kernel's FFMA produces (x = S' * 2^k is exact).
uint32_t rowOfBits(uint32_t xbits)
{
  volatile float z = bitsToFloat(xbits) + bitsToFloat(kRowEncCBits);
  const float    f = z;
  uint32_t       b;
  memcpy(&b, &f, 4);
  return (b >> 16) - kRowEncFirstKey;
}
// The smallest non-zero value a level +1 encode (the unclamped one) can see: linearFromSrgb(1) / 4.
inline uint32_t rowEncFloorBits() { return NVPYR_SRGB_DECODE_BITS[1] - (2u << 23); }

// entry[row] + bits(x) carries srgbFromLinear(x) in bits 24..31 for every x of the row that the kernel can present.
// Fails (returns false) if a row held two thresholds or spanned 2^24 patterns or more -- impossible with the pinned
// thresholds, checked so that a regenerated table cannot silently break the encode.
bool buildRowEncodeTable(uint32_t* out)
{
  const uint32_t* thr       = NVPYR_SRGB_ENCODE_THRESHOLD_BITS;
  const uint32_t  floorBits = rowEncFloorBits();
  if(rowOfBits(0u) != 0u || rowOfBits(floorBits) < 1u || rowOfBits(kEncMinBits) < 1u || floorBits > kEncMinBits
     || rowOfBits(kEncMaxBits) != kRowEncRows - 2u || rowOfBits(kRowEncTopBits) != kRowEncRows - 1u)
    return false;
  auto firstOfRow = [](uint32_t row) {  // rowOfBits is monotone over the non-negative floats
    uint32_t lo = 0u, hi = kRowEncTopBits + 1u;
    while(lo < hi)
    {
      const uint32_t mid = lo + (hi - lo) / 2u;
      if(rowOfBits(mid) >= row)
        hi = mid;
      else
        lo = mid + 1u;
    }
    return lo;
  };
  out[0] = 0u - kRowEncZeroBits;  // exact zero of level +1, the only pattern that reaches row 0
  uint32_t code = 0;              // thresholds <= first pattern of the row
  for(uint32_t row = 1; row < kRowEncRows; ++row)
  {
    const uint32_t lo = std::max(firstOfRow(row), floorBits), hi = firstOfRow(row + 1u);  // [lo, hi)
    if(hi <= lo)
    {
      out[row] = 0u;  // below the floor: never read
      continue;
    }
    while(code < 255u && thr[code] <= lo)
      ++code;
    if(code < 255u && thr[code] < hi)
    {
      const uint32_t t = thr[code];
      if((code + 1u < 255u && thr[code + 1u] < hi) || hi - t > (1u << 24) || t - lo > (1u << 24))
        return false;
      out[row] = ((code + 1u) << 24) - t;  // x >= t: code + 1;  x < t: the borrow leaves code
    }
    else
    {
      if(hi - lo > (1u << 24))
        return false;
      out[row] = (code << 24) - lo;
    }
  }
  for(uint32_t i = kRowEncRows; i < ((kRowEncRows + 3u) & ~3u); ++i)
    out[i] = 0u;
  return true;
}

bool buildHostTables(DeviceTables& t)
{
  static_assert(kEncShift <= 16 && kFastEncShift <= 16, "entry + bits must not carry past the code byte");
  for(int c = 0; c < 256; ++c)
    t.decode[c] = bitsToFloat(NVPYR_SRGB_DECODE_BITS[c]);
  return buildEncodeTable(t.encode, kEncShift, kEncEntriesPadded)
         && buildEncodeTable(t.encodeFast, kFastEncShift, kFastEncEntriesPadded) && buildRowEncodeTable(t.encodeRows);
}

// Host-only proof of the row table: EVERY float pattern the kernel can present (the floor .. 1.0, and the zero of
// level +1) is pushed through the kernel's own arithmetic -- row from RN(x + c), entry + bits, byte 3 -- and compared
// with "number of pinned thresholds <= x".  Returns the number of mismatches (0 expected), ~0 if the build failed.
uint64_t checkRowEncodeTable()
{
  static uint32_t rows[(kRowEncRows + 3u) & ~3u];
  if(!buildRowEncodeTable(rows))
    return ~0ull;
  const uint32_t* thr  = NVPYR_SRGB_ENCODE_THRESHOLD_BITS;
  uint64_t        bad  = ((rows[0] + kRowEncZeroBits) >> 24) != 0u;
  uint32_t        code = 0;
  for(uint32_t x = rowEncFloorBits(); x <= kRowEncTopBits; ++x)
  {
    while(code < 255u && thr[code] <= x)
      ++code;
    const uint32_t row = rowOfBits(x);
    bad += row >= kRowEncRows || ((rows[row] + x) >> 24) != code;
  }
  return bad;
}

// ------------------------------------------------------------ device context
// Counters of tailKernel ("which CTA finished the grid step last") and of the tuned fast kernel (dynamic tile
// hand-out); both return to zero when their launch ends.  A counter may only be shared by
// launches that are ordered one after the other, so: every STREAM owns one slot (launches of a stream are ordered;
// each kernel executes griddepcontrol.wait before it touches anything, so with programmatic dependent launch the
// counter is still used by one launch at a time), and every launch recorded into a CUDA GRAPH gets a dedicated slot
// that is never recycled (a graph may be replayed on any stream while live launches go on; launches of one
// executable graph are ordered among themselves by CUDA).  Exhaustion is an error (NVPYR_ERROR_OUT_OF_MEMORY),
// never a silently shared counter: at most kTicketPool streams + captured tail launches per device and process.
constexpr uint32_t kTicketPool = 4096;

constexpr uint32_t kMaxHostBands = 64;
constexpr uint32_t kBatchRing    = 1u << 17;  // base pointers (1 MB); a batch takes a contiguous slice

struct HostPipeline;

struct DeviceContext
{
  int           device   = -1;
  int           smCount  = 0;
  DeviceTables* tables   = nullptr;
  uint32_t*     tickets  = nullptr;  // kTicketPool zero-initialised slots of two counters: [0] tailKernel's ticket, [1] the fast kernel's tile hand-out
  std::mutex                       ticketMutex;
  std::map<cudaStream_t, uint32_t> ticketOfStream;
  uint32_t                         ticketsUsed = 0;
  // nvpyrDispatchBatch: ring of device words holding the chain base pointers of batched launches.  A slice stays
  // reserved until the event recorded behind the batch's last kernel has completed.
  const unsigned char** batchBases = nullptr;
  CUtensorMap*          batchMaps  = nullptr;  // same ring positions: one tensor map (level 0) per image; allocated by the first fused batch
  struct BatchSlice
  {
    uint64_t    begin, end;  // positions in the unwrapped ring
    cudaEvent_t done;
  };
  std::mutex                    batchMutex;
  std::deque<BatchSlice>        batchInFlight;
  std::vector<cudaEvent_t>      batchEventPool;
  uint64_t                      batchHead = 0;     // next free position (unwrapped)
  bool          genWindowOk = false;  // the strip kernels' absolute shared-memory addresses are valid on this device
  // nvpyrGenerateHost: a small pool of independent pipelines (device scratch chain + three streams + events), so
  // that concurrent round trips on one device do not serialise on one scratch buffer.
  std::mutex                 hostMutex;
  std::vector<HostPipeline*> hostIdle;
};

std::mutex                  g_ctxMutex;
std::vector<DeviceContext*> g_ctx;

nvpyrStatus getContext(DeviceContext** out)
{
  int dev = 0;
  NVPYR_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_ctxMutex);
  for(DeviceContext* c : g_ctx)
    if(c->device == dev)
    {
      *out = c;
      return NVPYR_SUCCESS;
    }
  cudaDeviceProp prop;
  NVPYR_CUDA(cudaGetDeviceProperties(&prop, dev));
  if(prop.major != 10 || prop.minor != 0)
    return NVPYR_ERROR_UNSUPPORTED;  // the kernels are built for sm_100a only (arch-specific: no other 10.x part runs them)
  static DeviceTables host;
  static const bool   hostOk = buildHostTables(host);
  if(!hostOk)
    return NVPYR_ERROR_UNSUPPORTED;
  DeviceTables*         d       = nullptr;
  uint32_t*             tickets = nullptr;
  const unsigned char** ring    = nullptr;
  uint32_t*             probe   = nullptr;
  cudaError_t           e       = cudaMalloc(&d, sizeof(DeviceTables));
  if(e == cudaSuccess)
    e = cudaMemcpy(d, &host, sizeof(DeviceTables), cudaMemcpyHostToDevice);
  if(e == cudaSuccess)
    e = cudaMalloc(&tickets, 2 * kTicketPool * sizeof(uint32_t));
  if(e == cudaSuccess)
    e = cudaMemset(tickets, 0, 2 * kTicketPool * sizeof(uint32_t));
  if(e == cudaSuccess)
    e = cudaMalloc(&ring, size_t(kBatchRing) * sizeof(void*));
  // Where does the dynamic shared-memory window of a kernel without static shared memory start on this
  // device / driver?  The strip kernels address their tables absolutely (nvpyr_general_srgba8.cuh); if the answer
  // is not what they were built for, sRGBA8 general steps take the functor-template kernel instead.
  uint32_t windowBase = 0;
  if(e == cudaSuccess)
    e = cudaMalloc(&probe, sizeof(uint32_t));
  if(e == cudaSuccess)
  {
    genWindowProbeKernel<<<1, 32, 1024>>>(probe);
    e = cudaMemcpy(&windowBase, probe, sizeof(uint32_t), cudaMemcpyDeviceToHost);
  }
  cudaFree(probe);
  if(e != cudaSuccess)
  {
    cudaFree(d);
    cudaFree(tickets);
    cudaFree(ring);
    g_lastCudaError = int(e);
    return e == cudaErrorMemoryAllocation ? NVPYR_ERROR_OUT_OF_MEMORY : NVPYR_ERROR_CUDA;
  }
  DeviceContext* c = new DeviceContext;
  c->batchBases    = ring;
  c->device        = dev;
  c->smCount       = prop.multiProcessorCount;
  c->tables        = d;
  c->tickets       = tickets;
  c->genWindowOk   = windowBase == kGenWindowBase;
  g_ctx.push_back(c);
  *out = c;
  return NVPYR_SUCCESS;
}

// Is `stream` being captured into a CUDA graph?  (Legacy-stream queries fail while another stream captures in
// global mode; that counts as "capturing": the caller must not do anything a capture forbids.)
bool streamIsCapturing(cudaStream_t stream)
{
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if(cudaStreamIsCapturing(stream, &st) != cudaSuccess)
  {
    cudaGetLastError();
    return true;
  }
  return st != cudaStreamCaptureStatusNone;
}

// The ticket counter a tail launch on `stream` may use (see kTicketPool).
nvpyrStatus acquireTicket(DeviceContext& ctx, cudaStream_t stream, uint32_t** ticket)
{
  const bool                  capturing = streamIsCapturing(stream);
  std::lock_guard<std::mutex> lock(ctx.ticketMutex);
  if(!capturing)
  {
    auto it = ctx.ticketOfStream.find(stream);
    if(it != ctx.ticketOfStream.end())
    {
      *ticket = ctx.tickets + 2u * it->second;
      return NVPYR_SUCCESS;
    }
  }
  if(ctx.ticketsUsed >= kTicketPool)
    return NVPYR_ERROR_OUT_OF_MEMORY;  // documented limit; never share a counter between unordered launches
  const uint32_t slot = ctx.ticketsUsed++;
  if(!capturing)
    ctx.ticketOfStream[stream] = slot;
  *ticket = ctx.tickets + 2u * slot;
  return NVPYR_SUCCESS;
}

// ------------------------------------------------------------------ launches
// Blocks per SM of a kernel (and the > 48 KB shared-memory opt-in), cached per (kernel, device):
// the two runtime calls cost several microseconds, as much as a small launch.
nvpyrStatus blocksPerSm(const void* kernel, size_t smem, int threads, int device, int* perSm)
{
  static std::mutex                                  m;
  static std::map<std::pair<const void*, int>, int>  cache;
  std::lock_guard<std::mutex>                        lock(m);
  auto                                               it = cache.find({kernel, device});
  if(it != cache.end())
  {
    *perSm = it->second;
    return NVPYR_SUCCESS;
  }
  NVPYR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int n = 0;
  NVPYR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem));
  if(n < 1)
    return NVPYR_ERROR_UNSUPPORTED;
  cache[{kernel, device}] = n;
  *perSm                  = n;
  return NVPYR_SUCCESS;
}

template <class K>
nvpyrStatus persistentGrid(K kernel, size_t smem, const DeviceContext& ctx, uint64_t workItems, int* grid,
                           int threads = 256)
{
  int         perSm = 0;
  nvpyrStatus st    = blocksPerSm(reinterpret_cast<const void*>(kernel), smem, threads, ctx.device, &perSm);
  if(st != NVPYR_SUCCESS)
    return st;
  uint64_t g = uint64_t(perSm) * uint64_t(ctx.smCount);
  if(g > workItems)
    g = workItems;
  *grid = int(g < 1 ? 1 : g);
  return NVPYR_SUCCESS;
}

template <class F, int M, bool kVec>
nvpyrStatus launchFastT(const DeviceContext& ctx, const FastParams& p, cudaStream_t stream)
{
  const size_t smem = sizeof(FastSmem<F>);
  int          grid = 1;
  nvpyrStatus  st   = persistentGrid(fastKernel<F, M, kVec>, smem, ctx, uint64_t(p.tilesX) * p.tilesY, &grid);
  if(st != NVPYR_SUCCESS)
    return st;
  NVPYR_CUDA(launchKernel(fastKernel<F, M, kVec>, grid, 256, smem, stream, p));
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

// Vector path of the fast kernels: 16-byte aligned input rows, output rows of level +1 aligned
// for a two-texel store.
template <class F>
bool fastVectorOk(const LevelView* lv)
{
  const uint32_t a1 = 2u * F::kTexelBytes > 16u ? 16u : 2u * F::kTexelBytes;
  return (reinterpret_cast<uintptr_t>(lv[0].ptr) % 16u == 0) && (lv[0].pitch % 16u == 0)
         && (reinterpret_cast<uintptr_t>(lv[1].ptr) % a1 == 0) && (lv[1].pitch % a1 == 0);
}

// The tuned sRGBA8 kernel (nvpyr_fast_srgba8.cuh): warp-autonomous 64 x 2^M tiles, 32 warps per CTA.
// batch != nullptr: one launch for batch->count chains of this size (p.lv[].ptr = offsets inside a chain).
// The tuned sRGBA8 fast kernel takes this step (else the functor-template kernel does).
template <class F>
bool tunedFastOk(const LevelView* lv)
{
  return std::is_same<F, Srgba8>::value && !g_forceGenericFast && fastVectorOk<F>(lv);
}

// cuTensorMapEncodeTiled, looked up at run time (libcuda is not linked).
using TensorMapEncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensorMapEncoder()
{
  static const TensorMapEncodeFn encode = [] {
    void*                           f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess)
      f = nullptr;
    return reinterpret_cast<TensorMapEncodeFn>(f);
  }();
  return encode;
}

// cuTensorMapReplaceAddress (CUDA 12.0+): the maps of a batch differ in their base address only.
using TensorMapReplaceFn = CUresult (*)(CUtensorMap*, void*);
TensorMapReplaceFn tensorMapReplacer()
{
  static const TensorMapReplaceFn replace = [] {
    void*                           f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if(cudaGetDriverEntryPoint("cuTensorMapReplaceAddress", &f, cudaEnableDefault, &q) != cudaSuccess)
      f = nullptr;
    return reinterpret_cast<TensorMapReplaceFn>(f);
  }();
  return replace;
}

// NVPYR_NO_SLAB_TASKS=1: always one warp per tile in the tuned fast kernel (A/B timing).
const bool g_noSlabTasks = [] {
  const char* e = getenv("NVPYR_NO_SLAB_TASKS");
  return e != nullptr && e[0] == '1';
}();

// NVPYR_SLAB_MAX_TILES_PER_WARP_X100: up to how many tiles per resident warp (x 100) a single-image fast step runs in
// slab-task mode (A/B timing; see launchFastSrgba8T).
const uint32_t g_slabMaxTilesPerWarpX100 = [] {
  const char* e = getenv("NVPYR_SLAB_MAX_TILES_PER_WARP_X100");
  return e != nullptr ? uint32_t(atoi(e)) : 25u;
}();

template <int M, bool kBatch, bool kPremul, bool kSlabTasks, int kWarps>
nvpyrStatus launchFastSrgba8W(const DeviceContext& ctx, const FastParams& p, const FastBatch& b, uint64_t work,
                              cudaStream_t stream)
{
  const size_t smem = fastSmemBytes(kFastTma && !kPremul && !kSlabTasks);
  int          grid = 1;
  nvpyrStatus  st   = persistentGrid(fastSrgba8Kernel<M, kBatch, kPremul, kSlabTasks, kWarps>, smem, ctx, work, &grid, kWarps * 32);
  if(st != NVPYR_SUCCESS)
    return st;
  FastTensorMap tmap{};
#if NVPYR_FAST_TMA == 2
  if(!kBatch && !kPremul && !kSlabTasks)
  {
    // 2-D tensor map of the step's input level: W x H texels of 4 bytes, row pitch in bytes, box 64 x 8
    const TensorMapEncodeFn encode = tensorMapEncoder();
    if(encode == nullptr)
      return NVPYR_ERROR_UNSUPPORTED;
    const cuuint64_t dims[2]    = {p.lv[0].w, p.lv[0].h};
    const cuuint64_t strides[1] = {p.lv[0].pitch};
    const cuuint32_t box[2]     = {64u, 8u};
    const cuuint32_t estr[2]    = {1u, 1u};
    if(encode(&tmap.map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, p.lv[0].ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
       != CUDA_SUCCESS)
      return NVPYR_ERROR_CUDA;
  }
#endif
  FastParams pp = p;
  if(!kSlabTasks)
  {
    uint32_t*   slot = nullptr;
    nvpyrStatus tst  = acquireTicket(const_cast<DeviceContext&>(ctx), stream, &slot);
    if(tst != NVPYR_SUCCESS)
      return tst;
    pp.tileCounter = slot + 1;
  }
  NVPYR_CUDA(launchKernel(fastSrgba8Kernel<M, kBatch, kPremul, kSlabTasks, kWarps>, grid, kWarps * 32, smem, stream, pp, b, tmap));
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

// Warps per CTA of a tile-mode launch (see fastSrgba8Kernel): the launch takes ceil(tiles / resident warps) rounds of
// one tile per warp; a round of the 24-warp build is shorter (measured 12.8 vs 18.1 us for 64 x 64 tiles at 16384^2) but
// holds fewer tiles.  NVPYR_FAST_WARPS_LARGE=32 pins the 32-warp build (A/B timing).
#ifndef NVPYR_FAST_WARPS_LARGE_N
#define NVPYR_FAST_WARPS_LARGE_N 24
#endif
constexpr int kFastWarpsLarge = NVPYR_FAST_WARPS_LARGE_N;
const bool g_noLargeWarps = [] {
  const char* e = getenv("NVPYR_FAST_WARPS_LARGE");
  return e != nullptr && atoi(e) == 32;
}();
const bool g_alwaysLargeWarps = [] {  // NVPYR_FAST_WARPS_LARGE=24: the 24-warp build for every tile-mode launch (tests, sanitizer)
  const char* e = getenv("NVPYR_FAST_WARPS_LARGE");
  return e != nullptr && atoi(e) == kFastWarpsLarge;
}();
inline bool preferFewerWarps(const DeviceContext& ctx, uint64_t tiles)
{
  if(g_noLargeWarps)
    return false;
  if(g_alwaysLargeWarps)
    return true;
  const uint64_t r32 = (tiles + uint64_t(ctx.smCount) * kFastWarps - 1) / (uint64_t(ctx.smCount) * kFastWarps);
  const uint64_t r24 = (tiles + uint64_t(ctx.smCount) * kFastWarpsLarge - 1) / (uint64_t(ctx.smCount) * kFastWarpsLarge);
  return r24 * 128u < r32 * 181u;
}

template <int M, bool kBatch, bool kPremul, bool kSlabTasks>
nvpyrStatus launchFastSrgba8K(const DeviceContext& ctx, const FastParams& p, const FastBatch& b, uint64_t work,
                              cudaStream_t stream)
{
  if(!kSlabTasks && !kPremul && preferFewerWarps(ctx, work))
    return launchFastSrgba8W<M, kBatch, false, false, kFastWarpsLarge>(ctx, p, b, work, stream);
  return launchFastSrgba8W<M, kBatch, kPremul, kSlabTasks, kFastWarps>(ctx, p, b, work, stream);
}

// premul: level 0 has straight alpha; premultiply it on the fly (scoped_image.hpp:233-255 fused into the read).
template <int M>
nvpyrStatus launchFastSrgba8T(const DeviceContext& ctx, FastParams p, cudaStream_t stream, const FastBatch* batch = nullptr,
                              bool premul = false)
{
  constexpr uint32_t tileH = M >= 3 ? (1u << M) : 8u;  // one warp per 64 x tileH tile
  p.tilesX                 = (p.lv[0].w + 63u) / 64u;
  p.tilesY                 = (p.lv[0].h + tileH - 1u) / tileH;
  FastBatch b              = batch ? *batch : FastBatch{nullptr, 0u, 0u, nullptr};
  b.tilesPerImage          = p.tilesX * p.tilesY;
  const uint64_t work      = uint64_t(b.tilesPerImage) * (batch ? b.count : 1u);
  if(work > 0xFFFFFFFFull)
    return NVPYR_ERROR_INVALID_VALUE;
  if(batch)
    return premul ? launchFastSrgba8K<M, true, true, false>(ctx, p, b, work, stream)
                  : launchFastSrgba8K<M, true, false, false>(ctx, p, b, work, stream);
  // Few tiles for the resident warps (images up to ~2048^2): warps take single slabs, not whole tiles.  With a
  // tile or more per warp the tile mode wins (the slab-to-slab prefetch stays inside one warp): measured
  // 1024^2 9.7 -> 5.7 us, 2048^2 14.2 -> 12.8 us, but 2560x1440 16.1 -> 16.5 us and 4096^2 26.6 -> 28.4 us.
  // Round 2: the slab-task kernel recycles its stash slots, so it is no longer limited to 32 tiles per CTA, and was
  // measured on larger images (NVPYR_SLAB_MAX_TILES_PER_WARP_X100): 4096^2 22.4 -> 27.1 us, 8192^2 69 -> 90 us.  The
  // kernel is issue-bound and a slab task costs ~25 % more instructions than a slab of a tile walk (tile coordinates,
  // output pointers, hand-off), which is more than the partly filled last round of whole tiles wastes (~4 % at 8192^2).
  constexpr bool kCanSlab = M >= 4;
  const uint64_t resident = uint64_t(ctx.smCount) * kFastWarps;
  if(kCanSlab && !g_noSlabTasks && 100u * work <= uint64_t(premul ? 25u : g_slabMaxTilesPerWarpX100) * resident)
    return premul ? launchFastSrgba8K<M, false, true, kCanSlab>(ctx, p, b, work, stream)
                  : launchFastSrgba8K<M, false, false, kCanSlab>(ctx, p, b, work, stream);
  return premul ? launchFastSrgba8K<M, false, true, false>(ctx, p, b, work, stream)
                : launchFastSrgba8K<M, false, false, false>(ctx, p, b, work, stream);
}

template <class F>
nvpyrStatus launchFast(const DeviceContext& ctx, FastParams p, uint32_t M, cudaStream_t stream, bool premul = false)
{
  p.tables = ctx.tables;
  if(M == 1)
  {
    const size_t   smem  = sizeof(FastSmem<F>);
    const uint64_t items = (uint64_t(p.lv[1].w) * p.lv[1].h + 255u) / 256u;
    int            grid  = 1;
    nvpyrStatus    st    = persistentGrid(fastKernel1<F>, smem, ctx, items, &grid);
    if(st != NVPYR_SUCCESS)
      return st;
    NVPYR_CUDA(launchKernel(fastKernel1<F>, grid, 256, smem, stream, p));
    ++g_launchCount;
    return NVPYR_SUCCESS;
  }
  p.tilesX = (p.lv[0].w + 63u) / 64u;
  p.tilesY = (p.lv[0].h + 63u) / 64u;
  const bool vec = fastVectorOk<F>(p.lv);
#define NVPYR_FAST_CASE(m)                                                                                        \
  case m:                                                                                                         \
    if(tunedFastOk<F>(p.lv))                                                                                      \
      return launchFastSrgba8T<m>(ctx, p, stream, nullptr, premul);                                               \
    return vec ? launchFastT<F, m, true>(ctx, p, stream) : launchFastT<F, m, false>(ctx, p, stream);
  switch(M)
  {
    NVPYR_FAST_CASE(2)
    NVPYR_FAST_CASE(3)
    NVPYR_FAST_CASE(4)
    NVPYR_FAST_CASE(5)
    NVPYR_FAST_CASE(6)
    default: return NVPYR_ERROR_INVALID_VALUE;
  }
#undef NVPYR_FAST_CASE
}

inline void generalTiles(const LevelView* lv, uint32_t levels, uint32_t tile2, uint32_t* tx, uint32_t* ty)
{
  const LevelView& out = levels == 1 ? lv[1] : lv[2];
  const uint32_t   t   = levels == 1 ? 2u * tile2 : tile2;
  *tx                  = (out.w + t - 1u) / t;
  *ty                  = (out.h + t - 1u) / t;
}

// Task granularity of the strip kernels: a warp task is a strip x a segment of rows.  Measured on B200
// (tools/bench_configs.py): ONE task per resident warp with segments as short as one row of level +2 wins --
// large levels get long segments (few prologues), small levels get enough tasks to occupy every warp
// (4095^2 55.7 -> 50.8 us, 2052^2 22.7 -> 18.2 us against "4 tasks per warp, at least 4 rows").
// NVPYR_GEN_TPW / NVPYR_GEN_MINSEG override for experiments.
const uint32_t g_genTasksPerWarp = getenv("NVPYR_GEN_TPW") ? uint32_t(std::max(1, atoi(getenv("NVPYR_GEN_TPW")))) : 1u;
const uint32_t g_genMinSegRows   = getenv("NVPYR_GEN_MINSEG") ? uint32_t(std::max(1, atoi(getenv("NVPYR_GEN_MINSEG")))) : 1u;

// The tuned general kernel (nvpyr_general_srgba8.cuh): warp strips, streaming rows; C = texel codec.
template <class C, int kLevels, bool kX3, bool kY3>
nvpyrStatus launchGeneralStripT(const DeviceContext& ctx, const GeneralParams& gp, cudaStream_t stream)
{
  GenStripParams p{};
  p.lv[0] = gp.lv[0], p.lv[1] = gp.lv[1], p.lv[2] = gp.lv[2];
  p.tables            = ctx.tables;
  p.stripsX           = std::max(1u, (gp.lv[1].w - 1u + 29u) / 30u);
  const size_t smem   = C::kSmemBytes;
  int          perSm  = 0;
  nvpyrStatus  st     = blocksPerSm(reinterpret_cast<const void*>(generalStripKernel<C, kLevels, kX3, kY3>), smem,
                                    C::kWarps * 32, ctx.device, &perSm);
  if(st != NVPYR_SUCCESS)
    return st;
  // Rows per warp task: see g_genTasksPerWarp.
  const uint32_t rows    = kLevels == 2 ? gp.lv[2].h : gp.lv[1].h;
  const uint64_t warps   = uint64_t(perSm) * ctx.smCount * C::kWarps;
  const uint32_t wantSeg = uint32_t(std::max<uint64_t>(1, (g_genTasksPerWarp * warps + p.stripsX - 1) / p.stripsX));
  p.segRows = std::min(64u, std::max(kLevels == 2 ? g_genMinSegRows : 2u * g_genMinSegRows, (rows + wantSeg - 1) / wantSeg));
  p.segsY                = (rows + p.segRows - 1) / p.segRows;
  const uint64_t tasks   = uint64_t(p.stripsX) * p.segsY;
  const uint64_t ctas    = std::min<uint64_t>(uint64_t(perSm) * ctx.smCount, (tasks + C::kWarps - 1) / C::kWarps);
  NVPYR_CUDA(launchKernel(generalStripKernel<C, kLevels, kX3, kY3>, int(std::max<uint64_t>(1, ctas)), C::kWarps * 32, smem,
                          stream, p));
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

// NVPYR_GEN_STRIP2=1: sRGBA8 through the two-column strip kernel instead of the four-column one (A/B).
const bool g_genStrip2 = [] {
  const char* e = getenv("NVPYR_GEN_STRIP2");
  return e != nullptr && e[0] == '1';
}();

// NVPYR_GEN_STRIP4_MIN_TEXELS: smallest input level (texels) that takes the four-column kernel (tests use 0 to
// run it on every size).  Measured: 2047^2 (4.19 M texels) 25.7 -> 24.8 us with the four-column kernel, 3095x990
// (3.06 M) 22.6 -> 23.2 us: the crossover lies between them.
const uint64_t g_genStrip4MinTexels = [] {
  const char* e = getenv("NVPYR_GEN_STRIP4_MIN_TEXELS");
  return e != nullptr ? uint64_t(strtoull(e, nullptr, 10)) : 4000000ull;
}();

// NVPYR_GEN_STAGED=0: the four-column strip kernel loads its rows with per-lane 4-byte loads into registers instead
// of staging them in shared memory with cp.async (A/B timing).
const bool g_genStaged = [] {
  const char* e = getenv("NVPYR_GEN_STAGED");
  return e == nullptr || e[0] != '0';
}();

// The four-columns-per-lane sRGBA8 strip kernel (generalStrip4Kernel): strips of 62 (+1 halo) level +1 columns.
template <int kLevels, bool kX3, bool kY3, bool kStaged>
nvpyrStatus launchGeneralStrip4K(const DeviceContext& ctx, const GeneralParams& gp, cudaStream_t stream)
{
  GenStripParams p{};
  p.lv[0] = gp.lv[0], p.lv[1] = gp.lv[1], p.lv[2] = gp.lv[2];
  p.tables            = ctx.tables;
  p.stripsX           = std::max(1u, (gp.lv[1].w - 1u + 61u) / 62u);
  const size_t smem   = kStaged ? kGenStagedSmemBytes : kGenSmemBytes;
  int          perSm  = 0;
  nvpyrStatus  st     = blocksPerSm(reinterpret_cast<const void*>(generalStrip4Kernel<kLevels, kX3, kY3, kStaged>), smem,
                                    kGen4Warps * 32, ctx.device, &perSm);
  if(st != NVPYR_SUCCESS)
    return st;
  // Rows per warp task: see g_genTasksPerWarp.
  const uint32_t rows    = kLevels == 2 ? gp.lv[2].h : gp.lv[1].h;
  const uint64_t warps   = uint64_t(perSm) * ctx.smCount * kGen4Warps;
  const uint32_t wantSeg = uint32_t(std::max<uint64_t>(1, (g_genTasksPerWarp * warps + p.stripsX - 1) / p.stripsX));
  p.segRows = std::min(64u, std::max(kLevels == 2 ? g_genMinSegRows : 2u * g_genMinSegRows, (rows + wantSeg - 1) / wantSeg));
  p.segsY                = (rows + p.segRows - 1) / p.segRows;
  const uint64_t tasks   = uint64_t(p.stripsX) * p.segsY;
  const uint64_t ctas    = std::min<uint64_t>(uint64_t(perSm) * ctx.smCount, (tasks + kGen4Warps - 1) / kGen4Warps);
  NVPYR_CUDA(launchKernel(generalStrip4Kernel<kLevels, kX3, kY3, kStaged>, int(std::max<uint64_t>(1, ctas)), kGen4Warps * 32, smem,
                          stream, p));
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

template <int kLevels, bool kX3, bool kY3>
nvpyrStatus launchGeneralStrip4T(const DeviceContext& ctx, const GeneralParams& gp, cudaStream_t stream)
{
  // (4-byte copies: the level base and pitch are multiples of the texel size by contract)
  if(g_genStaged)
    return launchGeneralStrip4K<kLevels, kX3, kY3, true>(ctx, gp, stream);
  return launchGeneralStrip4K<kLevels, kX3, kY3, false>(ctx, gp, stream);
}

template <int kLevels>
nvpyrStatus launchGeneralStrip4(const DeviceContext& ctx, const GeneralParams& gp, cudaStream_t stream)
{
  const bool x3 = gp.lv[0].w & 1u, y3 = gp.lv[0].h & 1u;
  if(x3)
    return y3 ? launchGeneralStrip4T<kLevels, true, true>(ctx, gp, stream)
              : launchGeneralStrip4T<kLevels, true, false>(ctx, gp, stream);
  return y3 ? launchGeneralStrip4T<kLevels, false, true>(ctx, gp, stream)
            : launchGeneralStrip4T<kLevels, false, false>(ctx, gp, stream);
}

template <class C, int kLevels>
nvpyrStatus launchGeneralStrip(const DeviceContext& ctx, const GeneralParams& gp, cudaStream_t stream)
{
  const bool x3 = gp.lv[0].w & 1u, y3 = gp.lv[0].h & 1u;
  if(x3)
    return y3 ? launchGeneralStripT<C, kLevels, true, true>(ctx, gp, stream)
              : launchGeneralStripT<C, kLevels, true, false>(ctx, gp, stream);
  return y3 ? launchGeneralStripT<C, kLevels, false, true>(ctx, gp, stream)
            : launchGeneralStripT<C, kLevels, false, false>(ctx, gp, stream);
}

// The strip kernel's codec for a functor set (void: only the functor-template kernel applies).
template <class F>
struct StripCodec
{
  using type = void;
};
template <>
struct StripCodec<Srgba8>
{
  using type = GenCodecSrgba8;
};
template <>
struct StripCodec<Rgba32f>
{
  using type = GenCodecRgba32f;
};
template <class C>
nvpyrStatus launchGeneralTuned(const DeviceContext& ctx, const GeneralParams& p, cudaStream_t stream)
{
  // Four columns per lane (fewer instructions per texel, fewer warps) pays on large levels; small levels need
  // the warp-level parallelism of the two-column kernel (measured cross-over around 2048^2).
  if(std::is_same<C, GenCodecSrgba8>::value && !g_genStrip2 && uint64_t(p.lv[0].w) * p.lv[0].h >= g_genStrip4MinTexels)
    return p.levels == 1 ? launchGeneralStrip4<1>(ctx, p, stream) : launchGeneralStrip4<2>(ctx, p, stream);
  return p.levels == 1 ? launchGeneralStrip<C, 1>(ctx, p, stream) : launchGeneralStrip<C, 2>(ctx, p, stream);
}
template <>
nvpyrStatus launchGeneralTuned<void>(const DeviceContext&, const GeneralParams&, cudaStream_t)
{
  return NVPYR_ERROR_UNSUPPORTED;
}

template <class F>
nvpyrStatus launchGeneral(const DeviceContext& ctx, GeneralParams p, cudaStream_t stream)
{
  p.tables = ctx.tables;
  // Tuned path (sRGBA8, rgba32f): no 1-texel-wide/high level involved (those use kernel size 1).
  using Codec = typename StripCodec<F>::type;
  constexpr bool kAbsoluteTables = std::is_same<Codec, GenCodecSrgba8>::value;  // needs the expected window base
  if(!std::is_same<Codec, void>::value && !g_forceGenericFast && (ctx.genWindowOk || !kAbsoluteTables) && p.lv[0].w >= 2
     && p.lv[0].h >= 2
     && (p.levels == 1 || (p.lv[1].w >= 2 && p.lv[1].h >= 2)))
    return launchGeneralTuned<Codec>(ctx, p, stream);
  generalTiles(p.lv, p.levels, kGenTile2, &p.tilesX, &p.tilesY);
  const size_t smem = sizeof(GeneralSmem<F>);
  int          grid = 1;
  nvpyrStatus  st   = persistentGrid(generalKernel<F>, smem, ctx, uint64_t(p.tilesX) * p.tilesY, &grid);
  if(st != NVPYR_SUCCESS)
    return st;
  NVPYR_CUDA(launchKernel(generalKernel<F>, grid, 256, smem, stream, p));
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

nvpyrStatus launchPremultiply(const DeviceContext& ctx, const void* in, void* out, uint64_t texels,
                              cudaStream_t stream)
{
  // 16-byte aligned buffers: groups of four texels through the conflict-free tables of the fast kernel;
  // the (at most three) texels left over, and unaligned buffers, through the functor-table kernel.
  uint64_t done = 0;
  if(reinterpret_cast<uintptr_t>(in) % 16u == 0 && reinterpret_cast<uintptr_t>(out) % 16u == 0 && texels >= 4u
     && !g_forceGenericFast)
  {
    const uint64_t n4   = texels / 4u;
    const size_t   smem = fastSmemBytes(false);
    int            grid = 1;
    nvpyrStatus    st   = persistentGrid(premultiplySrgba8Kernel, smem, ctx, (n4 + kFastWarps * 32 - 1) / (kFastWarps * 32), &grid,
                                         kFastWarps * 32);
    if(st != NVPYR_SUCCESS)
      return st;
    NVPYR_CUDA(launchKernel(premultiplySrgba8Kernel, grid, kFastWarps * 32, smem, stream, static_cast<const uint4*>(in),
                            static_cast<uint4*>(out), n4, static_cast<const DeviceTables*>(ctx.tables)));
    ++g_launchCount;
    done = n4 * 4u;
    if(done == texels)
      return NVPYR_SUCCESS;
  }
  const size_t smem = sizeof(Srgba8::Shared);
  int          grid = 1;
  nvpyrStatus  st   = persistentGrid(premultiplyKernel, smem, ctx, (texels - done + 255u) / 256u, &grid);
  if(st != NVPYR_SUCCESS)
    return st;
  NVPYR_CUDA(launchKernel(premultiplyKernel, grid, 256, smem, stream, static_cast<const uint32_t*>(in) + done,
                          static_cast<uint32_t*>(out) + done, texels - done, static_cast<const DeviceTables*>(ctx.tables)));
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

// ------------------------------------------------------------------ dispatch
struct ResolvedDesc
{
  nvpyrFormat  format;
  uint32_t     flags, w, h, levels, texelBytes;
  LevelView    lv[NVPYR_MAX_LEVELS];
  DispatcherRef general = defaultGeneralDispatcher, fast;
  bool          customDispatchers = false;  // nvpyrDispatchWithDispatchers: plan limits are checked, no fused paths
  cudaStream_t  stream;
};

nvpyrStatus resolve(const nvpyrDispatchDesc* d, ResolvedDesc& r)
{
  if(d == nullptr || d->structSize != sizeof(nvpyrDispatchDesc))
    return NVPYR_ERROR_INVALID_VALUE;
  if(d->extent.width == 0 || d->extent.height == 0)
    return NVPYR_ERROR_INVALID_VALUE;
  if(d->format != NVPYR_FORMAT_SRGBA8 && d->format != NVPYR_FORMAT_RGBA32F)
    return NVPYR_ERROR_UNSUPPORTED;
  constexpr uint32_t kSharedFlags = NVPYR_FLAG_F16_SHARED | NVPYR_FLAG_SRGB_SHARED;
  if(d->flags & ~uint32_t(NVPYR_FLAG_FORCE_GENERAL | NVPYR_FLAG_PREMULTIPLY_ALPHA | kSharedFlags | NVPYR_FLAG_GENERAL_BLIT))
    return NVPYR_ERROR_UNSUPPORTED;
  if((d->flags & NVPYR_FLAG_GENERAL_BLIT) && (d->flags & kSharedFlags))
    return NVPYR_ERROR_INVALID_VALUE;  // a blit has no shared-memory carry the shared types could apply to
  if((d->flags & (NVPYR_FLAG_PREMULTIPLY_ALPHA | kSharedFlags)) && d->format != NVPYR_FORMAT_SRGBA8)
    return NVPYR_ERROR_UNSUPPORTED;
  if((d->flags & kSharedFlags) == kSharedFlags)
    return NVPYR_ERROR_INVALID_VALUE;  // the two shared types are alternative builds of the shaders
  r.format         = d->format;
  r.flags          = d->flags;
  r.w              = d->extent.width;
  r.h              = d->extent.height;
  r.texelBytes     = d->format == NVPYR_FORMAT_SRGBA8 ? 4u : 16u;
  r.stream         = reinterpret_cast<cudaStream_t>(d->stream);
  const uint32_t maxLevels = levelCountFor(r.w, r.h);
  r.levels                 = d->levelCount == 0 ? maxLevels : d->levelCount;
  if(r.levels > maxLevels || r.levels > NVPYR_MAX_LEVELS)
    return NVPYR_ERROR_INVALID_VALUE;
  r.general = (d->flags & NVPYR_FLAG_GENERAL_BLIT) ? DispatcherRef(blitDispatcher) : DispatcherRef(defaultGeneralDispatcher);
  r.fast    = nullptr;
  if(!(d->flags & NVPYR_FLAG_FORCE_GENERAL))
  {
    r.fast = selectFastDispatcher(d->fastDivisibility, d->fastMaxLevels);
    if(!r.fast)
      return NVPYR_ERROR_UNSUPPORTED;
  }
  uint64_t off = 0;
  for(uint32_t i = 0; i < r.levels; ++i)
  {
    LevelView& v = r.lv[i];
    v.w          = levelDim(r.w, i);
    v.h          = levelDim(r.h, i);
    v.level      = i;
    if(d->levels[i] != nullptr)
    {
      v.ptr   = static_cast<unsigned char*>(d->levels[i]);
      v.pitch = d->rowPitchBytes[i] ? d->rowPitchBytes[i] : v.w * r.texelBytes;
      if(v.pitch < v.w * r.texelBytes)
        return NVPYR_ERROR_INVALID_VALUE;
    }
    else
    {
      if(d->base == nullptr)
        return NVPYR_ERROR_INVALID_VALUE;
      v.ptr   = static_cast<unsigned char*>(d->base) + off * r.texelBytes;
      v.pitch = v.w * r.texelBytes;
    }
    if(reinterpret_cast<uintptr_t>(v.ptr) % r.texelBytes != 0 || v.pitch % r.texelBytes != 0)
      return NVPYR_ERROR_INVALID_VALUE;
    off += uint64_t(v.w) * v.h;
  }
  if(d->base != nullptr && reinterpret_cast<uintptr_t>(d->base) % 16u != 0)
    return NVPYR_ERROR_INVALID_VALUE;
  return NVPYR_SUCCESS;
}

// Steps whose input level has at most kTailMaxTexels texels are candidates for tailKernel: a
// "grid step" spread over all CTAs followed by "solo" steps (input <= kSoloMaxTexels) that the
// last CTA runs alone.
const uint64_t kTailMaxTexels = [] {
  // NVPYR_TAIL_MAX_TEXELS overrides the threshold (tests use 0 to force the stand-alone kernels
  // onto tiny levels).
  const char* e = getenv("NVPYR_TAIL_MAX_TEXELS");
  return e != nullptr ? uint64_t(strtoull(e, nullptr, 10)) : 512ull * 512ull;
}();
constexpr uint64_t kSoloMaxTexelsFast    = 64ull * 64ull;  // one 64x64 tile
// general steps run solo when the input is at most kSoloMaxEdgeGeneral wide and high (one tile, nvpyr_kernels.cuh)

// NVPYR_NO_SOLO_SMEM=1: solo general steps read their input from global memory tile by tile as in round 1 (A/B timing).
const bool g_noSoloSmem = [] {
  const char* e = getenv("NVPYR_NO_SOLO_SMEM");
  return e != nullptr && e[0] == '1';
}();
// Largest input level (texels) of a solo general step in shared memory.  One CTA with un-replicated tables is
// LSU-bound: measured (round 2) with 127^2 inputs allowed, 2047^2 27.2 -> 40.2 us and 3095x990 24.1 -> 31.7 us (the
// 127 -> 63 -> 31 step alone takes ~20 us on one SM); with 63^2 the saved launch and the slower step cancel (4095^2
// 51.0 us either way); up to 32^2 -- the steps that ran solo anyway -- shared memory wins: 4095^2 51.3 -> 48.4 us,
// 4094^2 48.4 -> 45.8 us.
const uint64_t g_soloSmemMaxTexels = [] {
  const char* e = getenv("NVPYR_SOLO_SMEM_MAX_TEXELS");
  return e != nullptr ? uint64_t(strtoull(e, nullptr, 10)) : 32ull * 32ull;
}();
template <class TF>
bool soloSmemOk(uint32_t w, uint32_t h, uint32_t levels)
{
  return !g_noSoloSmem && uint64_t(w) * h <= g_soloSmemMaxTexels && soloSmemFits<TF>(w, h, levels);
}

template <class F>
struct TailFunctors
{
  using type = F;
};
template <>
struct TailFunctors<Srgba8>
{
  using type = Srgba8Lite;  // small levels: 1 KB decode table, cheap to set up
};

// NVPYR_TAIL_DEBUG_CLOCKS=1 (development): every tail launch is followed by a synchronisation and prints the clock64()
// stamps of its first CTA and of the CTA that ran the solo steps (cycles since the CTA started).
const bool g_tailDebugClocks = [] {
  const char* e = getenv("NVPYR_TAIL_DEBUG_CLOCKS");
  return e != nullptr && e[0] == '1';
}();
// NVPYR_TAIL_DEFER_WAIT: 0 = tailKernel waits for the previous kernel at its entry, 1 = right before the grid step's
// first load, 2 = the same and the next kernel is let go (griddepcontrol.launch_dependents) at the kernel's entry.
const uint32_t g_tailDeferWait = [] {
  const char* e = getenv("NVPYR_TAIL_DEFER_WAIT");
  return e != nullptr ? uint32_t(atoi(e)) : 1u;
}();
long long* tailDebugBuffer()
{
  static long long* buf = nullptr;
  if(g_tailDebugClocks && buf == nullptr)
    cudaMalloc(&buf, 1024 * 32 * sizeof(long long));
  if(buf != nullptr)
    cudaMemset(buf, 0, 1024 * 32 * sizeof(long long));
  return buf;
}
void tailDebugPrint(const TailParams& tp, int grid, cudaStream_t stream)
{
  if(tp.debugClocks == nullptr)
    return;
  cudaStreamSynchronize(stream);
  std::vector<long long> h(size_t(grid) * 32u);
  cudaMemcpy(h.data(), tp.debugClocks, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  int last = 0;
  for(int b = 0; b < grid; ++b)
    if(h[size_t(b) * 32u + 4u] != 0)
      last = b;
  for(int b : {0, last})
  {
    fprintf(stderr, "nvpyr tail clocks: grid %d steps %u cta %d:", grid, tp.numSteps, b);
    for(int k = 1; k < 32; ++k)
      if(h[size_t(b) * 32u + k] != 0)
        fprintf(stderr, " [%d] %lld", k, h[size_t(b) * 32u + k] - h[size_t(b) * 32u]);
    fprintf(stderr, "\n");
    if(last == 0)
      break;
  }
}

// NVPYR_NO_FAST_TINY=1: the last tiny fast steps of a chain run through fastTileLoop as before (A/B timing).
const bool g_noFastTiny = [] {
  const char* e = getenv("NVPYR_NO_FAST_TINY");
  return e != nullptr && e[0] == '1';
}();
// TailStep::vec bit 1: a (solo) fast step small enough for fastTinyStep.
inline uint32_t fastTinyBit(const TailStep& ts)
{
  return !g_noFastTiny && ts.levels <= 3u && ts.lv[0].w <= kFastTinyMaxEdge && ts.lv[0].h <= kFastTinyMaxEdge ? 2u : 0u;
}

// One tail launch: steps[0] on the whole grid, steps[1..count) solo.
template <class F>
nvpyrStatus launchTail(DeviceContext& ctx, const ResolvedDesc& r, const nvpyrPlanStep* steps, int count)
{
  using TF = typename TailFunctors<F>::type;
  TailParams tp{};
  bool       anySoloSmem = false;
  tp.numSteps = uint32_t(count);
  tp.tables   = ctx.tables;
  nvpyrStatus tst = acquireTicket(ctx, r.stream, &tp.ticket);
  if(tst != NVPYR_SUCCESS)
    return tst;
  for(int i = 0; i < count; ++i)
  {
    const nvpyrPlanStep& s  = steps[i];
    TailStep&            ts = tp.steps[i];
    const bool           isBlit = s.pipeline == 0 && (r.flags & NVPYR_FLAG_GENERAL_BLIT);
    ts.pipeline             = isBlit ? 2u : s.pipeline;
    ts.levels               = s.levelCount;
    for(uint32_t k = 0; k <= s.levelCount; ++k)
      ts.lv[k] = r.lv[s.inputLevel + k];
    if(isBlit)
    {
      // thread-strided over the destination texels: "tiles" = chunks of kTailThreads texels
      ts.tilesX = uint32_t((uint64_t(ts.lv[1].w) * ts.lv[1].h + uint64_t(kTailThreads) - 1u) / uint64_t(kTailThreads));
      ts.tilesY = 1u;
    }
    else if(s.pipeline == 1)
    {
      ts.vec    = (fastVectorOk<F>(ts.lv) ? 1u : 0u) | fastTinyBit(ts);
      ts.tilesX = (ts.lv[0].w + 63u) / 64u;
      ts.tilesY = (ts.lv[0].h + 63u) / 64u;
    }
    else
    {
      generalTiles(ts.lv, s.levelCount, i == 0 ? kGenTile2Small : SoloTile2<typename TF::Value>::value, &ts.tilesX, &ts.tilesY);
      ts.soloSmem = i > 0 && soloSmemOk<TF>(ts.lv[0].w, ts.lv[0].h, s.levelCount) ? 1u : 0u;
      anySoloSmem |= ts.soloSmem != 0u;
      fillTailStepReciprocals(ts);
    }
  }
  uint64_t work = uint64_t(tp.steps[0].tilesX) * tp.steps[0].tilesY;
  if(tp.steps[0].pipeline == 1 && tp.steps[0].levels == 1)  // fastLoop1 is thread-strided, not tiled
    work = (uint64_t(tp.steps[0].lv[1].w) * tp.steps[0].lv[1].h + uint64_t(kTailThreads) - 1u) / uint64_t(kTailThreads);
  // The whole-level buffers of the solo steps are only allocated by launches that use them (a small footprint lets
  // the next kernel's CTAs move in early); the opt-in and the grid size are those of the larger footprint.
  const size_t smemBase = (sizeof(TailSmem<TF>) + 15u) & ~size_t(15);
  const size_t smemSolo = smemBase + sizeof(SoloSmem);
  const size_t smemMax  = smemBase + std::max<size_t>(sizeof(SoloSmem), kCascadeAreaMax);  // the opt-in is made once per kernel
  const size_t smem     = anySoloSmem ? smemSolo : sizeof(TailSmem<TF>);
  int          grid = 1;
  nvpyrStatus  st   = persistentGrid(tailKernel<TF>, smemMax, ctx, work, &grid, kTailThreads);
  if(st != NVPYR_SUCCESS)
    return st;
  tp.debugClocks = tailDebugBuffer();
  tp.deferWait   = g_tailDeferWait;
  NVPYR_CUDA(launchKernel(tailKernel<TF>, grid, kTailThreads, smem, r.stream, tp));
  tailDebugPrint(tp, grid, r.stream);
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

// ------------------------------------------------------------------ cascades (cascadeRun, nvpyr_kernels.cuh)
// NVPYR_CASCADE=1 switches the cascades on.  They are OFF by default: bit-exact on the whole GPU suite, but measured
// no faster than one dispatch per grid step (a level costs ~0.5 us of dependent latency on one CTA
// whichever kernel runs it, and the launch boundary a cascade saves costs no more than the staging and the halo it adds).
const bool g_noCascade = [] {
  const char* e = getenv("NVPYR_CASCADE");
  return !(e != nullptr && e[0] == '1');
}();
// Largest input level (texels) of a general step that STARTS a tail launch as a cascade over the whole grid
// (NVPYR_CASCADE_MAX_TEXELS); larger levels take the strip kernels.
const uint64_t g_cascadeMaxTexels = [] {
  const char* e = getenv("NVPYR_CASCADE_MAX_TEXELS");
  return e != nullptr ? uint64_t(strtoull(e, nullptr, 10)) : 1100ull * 1100ull;
}();
// Largest input level (texels) of a cascade that ONE CTA runs alone after the grid step (NVPYR_CASCADE_SOLO_MAX_TEXELS).
const uint64_t g_cascadeSoloMaxTexels = [] {
  const char* e = getenv("NVPYR_CASCADE_SOLO_MAX_TEXELS");
  return e != nullptr ? uint64_t(strtoull(e, nullptr, 10)) : 64ull * 64ull;
}();
// A deeper grid cascade (more dispatches per launch, more halo recomputed) is taken while the texels one SM stages
// stay below max(g_cascadeCostFactor x its fair share of the input level, g_cascadeCostFloor).
const double g_cascadeCostFactor = [] {
  const char* e = getenv("NVPYR_CASCADE_COST_FACTOR");
  return e != nullptr ? atof(e) : 2.5;
}();
const uint64_t g_cascadeCostFloor = [] {
  const char* e = getenv("NVPYR_CASCADE_COST_FLOOR");
  return e != nullptr ? uint64_t(strtoull(e, nullptr, 10)) : 4096ull;
}();

const bool g_cascadeDebug = [] {
  const char* e = getenv("NVPYR_CASCADE_DEBUG");
  return e != nullptr && e[0] == '1';
}();
// Largest cascade area a launch may use (NVPYR_CASCADE_AREA_MAX, at most kCascadeAreaMax): a small footprint lets the
// CTAs start -- and set up their tables -- while the previous kernel still occupies the SMs.
const uint64_t g_cascadeAreaMax = [] {
  const char* e = getenv("NVPYR_CASCADE_AREA_MAX");
  return std::min<uint64_t>(kCascadeAreaMax, e != nullptr ? uint64_t(strtoull(e, nullptr, 10)) : uint64_t(kCascadeAreaMax));
}();
struct CascadeGeom
{
  uint32_t tileW = 0, tileH = 0, tilesX = 0, tilesY = 0, atStage = 0, off0 = 0, offA = 0, offB = 0;
  size_t   areaBytes = 0;
  uint64_t cost      = 0;  // texels of the input level the busiest SM stages
  uint64_t rounds    = 0;  // tiles the busiest CTA walks
};
// What a round of a cascade costs besides its texels (staging round trip, a CTA barrier per level), in texels.
const uint64_t g_cascadeRoundCost = [] {
  const char* e = getenv("NVPYR_CASCADE_ROUND_COST");
  return e != nullptr ? uint64_t(strtoull(e, nullptr, 10)) : 8192ull;
}();

// Shared-memory layout and cost of a cascade over lv[0..n] whose CTAs own tw x th tiles of level n.
template <class TF>
bool cascadeGeometry(const LevelView* lv, uint32_t n, uint32_t tw, uint32_t th, uint32_t smCount, CascadeGeom& g)
{
  auto     taps = [](uint32_t size) { return size == 1u ? 1u : (2u | (size & 1u)); };
  uint32_t fx[7], fy[7];
  fx[n] = std::min(tw, lv[n].w), fy[n] = std::min(th, lv[n].h);
  for(int l = int(n) - 1; l >= 0; --l)
  {
    fx[l] = std::min(2u * (fx[l + 1] - 1u) + taps(lv[l].w), lv[l].w);
    fy[l] = std::min(2u * (fy[l + 1] - 1u) + taps(lv[l].h), lv[l].h);
  }
  auto         align16 = [](uint64_t b) { return (b + 15u) & ~uint64_t(15); };
  const size_t vb = sizeof(typename TF::Value), tb = TF::kTexelBytes;
  const uint64_t a  = n >= 2u ? align16(uint64_t(fx[1]) * fy[1] * vb) : 0u;  // levels 1, 3, 5 (the last level is not kept)
  const uint64_t b  = n >= 3u ? align16(uint64_t(fx[2]) * fy[2] * vb) : 0u;  // levels 2, 4
  const uint64_t v0 = align16(uint64_t(fx[0]) * fy[0] * vb), r0 = align16(uint64_t(fx[0]) * fy[0] * tb);
  uint64_t       s0;
  if(kCascadeHeaderBytes + v0 + a + b <= g_cascadeAreaMax)
    g.atStage = 1u, s0 = v0;
  else if(kCascadeHeaderBytes + r0 + a + b <= g_cascadeAreaMax)
    g.atStage = 0u, s0 = r0;
  else
    return false;
  g.tileW = fx[n], g.tileH = fy[n];
  g.tilesX = (lv[n].w + fx[n] - 1u) / fx[n], g.tilesY = (lv[n].h + fy[n] - 1u) / fy[n];
  g.off0      = kCascadeHeaderBytes;
  g.offA      = uint32_t(g.off0 + s0);
  g.offB      = uint32_t(g.offA + a);
  g.areaBytes = size_t(g.offB + b);
  const uint64_t tiles = uint64_t(g.tilesX) * g.tilesY;
  g.rounds             = (tiles + smCount - 1u) / smCount;
  g.cost               = g.rounds * uint64_t(fx[0]) * fy[0];
  return true;
}

// The cheapest tile size for a grid cascade over lv[0..n].
template <class TF>
bool cascadeBestGrid(const LevelView* lv, uint32_t n, uint32_t smCount, CascadeGeom& best)
{
  bool found = false;
  for(uint32_t t = 1; t <= 256u; ++t)
  {
    CascadeGeom g;
    if(!cascadeGeometry<TF>(lv, n, t, t, smCount, g))
      break;  // footprints only grow with t
    if(!found || g.cost + g.rounds * g_cascadeRoundCost < best.cost + best.rounds * g_cascadeRoundCost)
      best = g, found = true;
    if(t >= lv[n].w && t >= lv[n].h)
      break;
  }
  return found;
}

void fillCascadeStep(TailStep& ts, const ResolvedDesc& r, const nvpyrPlanStep* steps, int count, const CascadeGeom& g)
{
  ts          = TailStep{};
  ts.pipeline = 3u;
  uint32_t n  = 0;
  ts.lv[0]    = r.lv[steps[0].inputLevel];
  for(int k = 0; k < count; ++k)
  {
    for(uint32_t j = 1; j <= steps[k].levelCount; ++j)
      ts.lv[n + j] = r.lv[steps[k].inputLevel + j];
    n += steps[k].levelCount;
    ts.boundaryMask |= 1u << n;
  }
  ts.levels = n;
  ts.tileW = g.tileW, ts.tileH = g.tileH, ts.tilesX = g.tilesX, ts.tilesY = g.tilesY;
  ts.atStage = g.atStage, ts.off0 = g.off0, ts.offA = g.offA, ts.offB = g.offB;
  fillTailStepReciprocals(ts);
}

// One tail launch that starts at plan step steps[0] (at most `avail` steps follow): general steps run as cascades --
// the first as a grid cascade of as many dispatches as the cost limit allows, the following ones solo.  *consumed =
// plan steps taken.
template <class F>
nvpyrStatus launchTailCascade(DeviceContext& ctx, const ResolvedDesc& r, const nvpyrPlanStep* steps, int avail, int* consumed)
{
  using TF = typename TailFunctors<F>::type;
  TailParams tp{};
  tp.tables        = ctx.tables;
  nvpyrStatus tst  = acquireTicket(ctx, r.stream, &tp.ticket);
  if(tst != NVPYR_SUCCESS)
    return tst;
  size_t   areaBytes = 0;
  int      i         = 0;
  uint32_t items     = 0;
  uint64_t work      = 1;
  auto     texels    = [&](int k) { return uint64_t(steps[k].srcWidth) * steps[k].srcHeight; };
  // consecutive general steps from k on, at most three dispatches and six levels: lv[] and level count
  auto gather = [&](int k, int maxSteps, LevelView* lv, int* stepLevels) {
    int      cnt = 0;
    uint32_t n   = 0;
    lv[0]        = r.lv[steps[k].inputLevel];
    while(k + cnt < avail && cnt < maxSteps && steps[k + cnt].pipeline == 0 && n + steps[k + cnt].levelCount <= 6u)
    {
      for(uint32_t j = 1; j <= steps[k + cnt].levelCount; ++j)
        lv[n + j] = r.lv[steps[k + cnt].inputLevel + j];
      n += steps[k + cnt].levelCount;
      stepLevels[cnt++] = int(n);
    }
    return cnt;
  };
  while(i < avail && items < kMaxTailSteps)
  {
    const nvpyrPlanStep& s  = steps[i];
    TailStep&            ts = tp.steps[items];
    if(s.pipeline == 1)
    {
      if(items > 0 && texels(i) > kSoloMaxTexelsFast)
        break;
      ts          = TailStep{};
      ts.pipeline = 1u;
      ts.levels   = s.levelCount;
      for(uint32_t k = 0; k <= s.levelCount; ++k)
        ts.lv[k] = r.lv[s.inputLevel + k];
      ts.vec    = (fastVectorOk<F>(ts.lv) ? 1u : 0u) | fastTinyBit(ts);
      ts.tilesX = (ts.lv[0].w + 63u) / 64u;
      ts.tilesY = (ts.lv[0].h + 63u) / 64u;
      if(items == 0)
      {
        work = uint64_t(ts.tilesX) * ts.tilesY;
        if(ts.levels == 1)  // fastLoop1 is thread-strided, not tiled
          work = (uint64_t(ts.lv[1].w) * ts.lv[1].h + uint64_t(kTailThreads) - 1u) / uint64_t(kTailThreads);
      }
      ++i;
    }
    else
    {
      LevelView   lv[7];
      int         stepLevels[3];
      CascadeGeom g;
      int         take = 0;
      if(items == 0)
      {
        const int cnt = gather(i, 3, lv, stepLevels);
        const double fair = double(texels(i)) / double(ctx.smCount);
        const uint64_t limit = std::max<uint64_t>(uint64_t(g_cascadeCostFactor * fair), g_cascadeCostFloor);
        for(int d = cnt; d >= 1 && take == 0; --d)
        {
          CascadeGeom c;
          if(cascadeBestGrid<TF>(lv, uint32_t(stepLevels[d - 1]), uint32_t(ctx.smCount), c) && (d == 1 || c.cost <= limit))
            g = c, take = d;
        }
        if(take == 0)
          return NVPYR_ERROR_UNSUPPORTED;  // (a one-texel tile of a one-dispatch cascade always fits)
        work = uint64_t(g.tilesX) * g.tilesY;
      }
      else
      {
        if(texels(i) > g_cascadeSoloMaxTexels)
          break;
        const int cnt = gather(i, 3, lv, stepLevels);
        for(int d = cnt; d >= 1 && take == 0; --d)
        {
          const uint32_t n = uint32_t(stepLevels[d - 1]);
          if(cascadeGeometry<TF>(lv, n, lv[n].w, lv[n].h, uint32_t(ctx.smCount), g))
            take = d;
        }
        if(take == 0)
          break;
      }
      fillCascadeStep(ts, r, steps + i, take, g);
      if(g_cascadeDebug)
        fprintf(stderr, "nvpyr cascade: %s %ux%u -> %ux%u (%d dispatches, %u levels) tile %ux%u, %ux%u tiles, %s, %zu bytes, cost %llu\n",
                items == 0 ? "grid" : "solo", ts.lv[0].w, ts.lv[0].h, ts.lv[ts.levels].w, ts.lv[ts.levels].h, take, ts.levels, g.tileW,
                g.tileH, g.tilesX, g.tilesY, g.atStage ? "values" : "raw", g.areaBytes, (unsigned long long)g.cost);
      areaBytes = std::max(areaBytes, g.areaBytes);
      i += take;
    }
    ++items;
  }
  tp.numSteps = items;
  *consumed   = i;
  const size_t smemBase = (sizeof(TailSmem<TF>) + 15u) & ~size_t(15);
  const size_t smemMax  = smemBase + std::max<size_t>(sizeof(SoloSmem), kCascadeAreaMax);
  int          grid     = 1;
  nvpyrStatus  st       = persistentGrid(tailKernel<TF>, smemMax, ctx, work, &grid, kTailThreads);
  if(st != NVPYR_SUCCESS)
    return st;
  tp.debugClocks = tailDebugBuffer();
  tp.deferWait   = g_tailDeferWait;
  NVPYR_CUDA(launchKernel(tailKernel<TF>, grid, kTailThreads, areaBytes ? smemBase + areaBytes : sizeof(TailSmem<TF>), r.stream, tp));
  tailDebugPrint(tp, grid, r.stream);
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

// One level by a linear-filter blit (NVPYR_FLAG_GENERAL_BLIT).
template <class F>
nvpyrStatus launchBlit(DeviceContext& ctx, const LevelView& src, const LevelView& dst, cudaStream_t stream)
{
  using TF = typename TailFunctors<F>::type;
  BlitParams p{src, dst, ctx.tables};
  const size_t smem = sizeof(typename TF::Shared);
  int          grid = 1;
  nvpyrStatus  st   = persistentGrid(blitKernel<TF>, smem, ctx, (uint64_t(dst.w) * dst.h + 255u) / 256u, &grid);
  if(st != NVPYR_SUCCESS)
    return st;
  NVPYR_CUDA(launchKernel(blitKernel<TF>, grid, 256, smem, stream, p));
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

// firstStep > 0: the steps before it have been enqueued by the caller (nvpyrGenerateHost runs step 0
// band by band).
// premulFirst: step 0 also premultiplies level 0 on the fly (the caller checked canFusePremultiply).
template <class F>
nvpyrStatus runPlan(DeviceContext& ctx, const ResolvedDesc& r, int firstStep = 0, bool premulFirst = false)
{
  nvpyrPlanStep steps[NVPYR_MAX_STEPS];
  const int     n = buildPlan(r.w, r.h, r.levels, r.general, r.fast, steps, NVPYR_MAX_STEPS);
  if(n < 0)
    return NVPYR_ERROR_INVALID_VALUE;
  auto texels = [&](int i) { return uint64_t(steps[i].srcWidth) * steps[i].srcHeight; };
  const bool blit = (r.flags & NVPYR_FLAG_GENERAL_BLIT) != 0;  // general steps are one-level blits (tailKernel pipeline 2 when small)
  constexpr uint64_t kSoloMaxTexelsBlit = 128ull * 128ull;      // a blit of a level this small runs solo (<= 8 texels per thread)
  for(int i = firstStep; i < n; ++i)
    if(blit && steps[i].pipeline == 0 && steps[i].levelCount != 1)
      return NVPYR_ERROR_INVALID_VALUE;  // (a user dispatcher together with the blit flag must fill one level per step)
  for(int i = firstStep; i < n;)
  {
    const nvpyrPlanStep& s = steps[i];
    nvpyrStatus          st;
    if(blit && s.pipeline == 0 && (g_noTailFusion || texels(i) > kTailMaxTexels))
    {
      st = launchBlit<F>(ctx, r.lv[s.inputLevel], r.lv[s.inputLevel + 1], r.stream);
      ++i;
    }
    else if(!g_noTailFusion && !g_noCascade && !blit
            && (s.pipeline == 0 ? texels(i) <= std::max(g_cascadeMaxTexels, kTailMaxTexels) : texels(i) <= kTailMaxTexels))
    {
      // small steps with the general ones as cascades: several dispatches per launch, no boundary between them
      int consumed = 0;
      st           = launchTailCascade<F>(ctx, r, steps + i, n - i, &consumed);
      i += consumed;
    }
    else if(!g_noTailFusion && texels(i) <= kTailMaxTexels)
    {
      // grid step i, then as many solo steps as follow (level sizes only shrink)
      int count = 1;
      while(i + count < n && count < int(kMaxTailSteps)
            && (steps[i + count].pipeline == 1
                    ? texels(i + count) <= kSoloMaxTexelsFast
                : blit
                    ? texels(i + count) <= kSoloMaxTexelsBlit
                    : (std::max(steps[i + count].srcWidth, steps[i + count].srcHeight) <= kSoloMaxEdgeGeneral
                       || soloSmemOk<typename TailFunctors<F>::type>(steps[i + count].srcWidth, steps[i + count].srcHeight,
                                                                     steps[i + count].levelCount))))
        ++count;
      st = launchTail<F>(ctx, r, steps + i, count);
      i += count;
    }
    else
    {
      if(s.pipeline == 1)
      {
        FastParams p{};
        for(uint32_t k = 0; k <= s.levelCount; ++k)
          p.lv[k] = r.lv[s.inputLevel + k];
        st = launchFast<F>(ctx, p, s.levelCount, r.stream, premulFirst && i == 0);
      }
      else
      {
        GeneralParams p{};
        for(uint32_t k = 0; k <= s.levelCount; ++k)
          p.lv[k] = r.lv[s.inputLevel + k];
        p.levels = s.levelCount;
        st       = launchGeneral<F>(ctx, p, r.stream);
      }
      ++i;
    }
    if(st != NVPYR_SUCCESS)
      return st;
  }
  return NVPYR_SUCCESS;
}

// The premultiply pre-pass can ride in the first launch when that launch is the tuned sRGBA8 fast kernel
// (tiles never overlap, so rewriting level 0 in place is safe; the general pipeline re-reads halo columns).
bool canFusePremultiply(const ResolvedDesc& r)
{
  if(r.format != NVPYR_FORMAT_SRGBA8 || r.levels < 2 || !r.fast || (r.flags & (NVPYR_FLAG_F16_SHARED | NVPYR_FLAG_SRGB_SHARED)))
    return false;
  nvpyrPlanStep steps[NVPYR_MAX_STEPS];
  const int     n = buildPlan(r.w, r.h, r.levels, r.general, r.fast, steps, NVPYR_MAX_STEPS);
  if(n < 1 || steps[0].pipeline != 1 || steps[0].levelCount < 2)
    return false;
  if(!g_noTailFusion && uint64_t(steps[0].srcWidth) * steps[0].srcHeight <= kTailMaxTexels)
    return false;  // small images start in tailKernel
  return tunedFastOk<Srgba8>(r.lv);
}

nvpyrStatus dispatchResolved(const ResolvedDesc& r)
{
  DeviceContext* ctx = nullptr;
  nvpyrStatus    st  = getContext(&ctx);
  if(st != NVPYR_SUCCESS)
    return st;
  bool fusePremul = false;
  if(r.flags & NVPYR_FLAG_PREMULTIPLY_ALPHA)
  {
    fusePremul = canFusePremultiply(r);
    // Level 0 is tight in the packed layout; with a pitched level 0 go row by row.
    const LevelView& v = r.lv[0];
    if(fusePremul)
      ;
    else if(v.pitch == v.w * 4u)
      st = launchPremultiply(*ctx, v.ptr, v.ptr, uint64_t(v.w) * v.h, r.stream);
    else
      for(uint32_t y = 0; y < v.h && st == NVPYR_SUCCESS; ++y)
        st = launchPremultiply(*ctx, v.ptr + size_t(y) * v.pitch, v.ptr + size_t(y) * v.pitch, v.w, r.stream);
    if(st != NVPYR_SUCCESS)
      return st;
  }
  if(r.levels <= 1)
    return NVPYR_SUCCESS;
  if(r.flags & NVPYR_FLAG_F16_SHARED)
    return runPlan<Srgba8F16Shared>(*ctx, r);
  if(r.flags & NVPYR_FLAG_SRGB_SHARED)
    return runPlan<Srgba8SrgbShared>(*ctx, r);
  return r.format == NVPYR_FORMAT_SRGBA8 ? runPlan<Srgba8>(*ctx, r, 0, fusePremul) : runPlan<Rgba32f>(*ctx, r);
}

// ------------------------------------------------------------- fused batches
// nvpyrDispatchBatch on a HOMOGENEOUS batch (sRGBA8, one size, one stream, packed chains whose plan is
// "one big fast step, then small steps"): two launches for the whole batch instead of two per image.  The
// big step streams the tiles of all images through one persistent grid (no per-image launch latency, table
// set-up or tail imbalance: a batch of 4096^2 images runs at the speed of one 16384^2-class image); the
// small steps run one CTA per image.  Anything else falls back to one dispatch per image.
constexpr uint64_t kBatchSoloMaxTexels = 256ull * 256ull;

template <int M>
nvpyrStatus launchFastBatchM(const DeviceContext& ctx, const FastParams& p, cudaStream_t stream, const FastBatch& b,
                             bool premul)
{
  return launchFastSrgba8T<M>(ctx, p, stream, &b, premul);
}

nvpyrStatus dispatchBatchFused(DeviceContext& ctx, const std::vector<ResolvedDesc>& r, bool* handled)
{
  *handled = false;
  const ResolvedDesc& a     = r[0];
  const uint32_t      count = uint32_t(r.size());
  if(count < 2 || count > kBatchRing / 4 || a.format != NVPYR_FORMAT_SRGBA8 || a.levels < 2 || g_forceGenericFast
     || g_noTailFusion || !a.fast || a.customDispatchers || (a.flags & (NVPYR_FLAG_F16_SHARED | NVPYR_FLAG_SRGB_SHARED)))
    return NVPYR_SUCCESS;
  for(const ResolvedDesc& d : r)
  {
    if(d.format != a.format || d.flags != a.flags || d.w != a.w || d.h != a.h || d.levels != a.levels || d.fast != a.fast
       || d.stream != a.stream || reinterpret_cast<uintptr_t>(d.lv[0].ptr) % 16u != 0)
      return NVPYR_SUCCESS;
    for(uint32_t k = 0; k < d.levels; ++k)  // packed layout only
      if(d.lv[k].pitch != d.lv[k].w * 4u
         || size_t(d.lv[k].ptr - d.lv[0].ptr) != size_t(levelOffsetTexels(d.w, d.h, k)) * 4u)
        return NVPYR_SUCCESS;
  }
  if(a.flags & NVPYR_FLAG_GENERAL_BLIT)
    return NVPYR_SUCCESS;  // blitted levels are separate launches: one dispatch per image
  nvpyrPlanStep steps[NVPYR_MAX_STEPS];
  const int     n = buildPlan(a.w, a.h, a.levels, a.general, a.fast, steps, NVPYR_MAX_STEPS);
  if(n < 1)
    return NVPYR_SUCCESS;
  auto texels = [&](int i) { return uint64_t(steps[i].srcWidth) * steps[i].srcHeight; };
  if(steps[0].pipeline != 1 || steps[0].levelCount < 2 || texels(0) <= kTailMaxTexels || n - 1 > int(kMaxTailSteps))
    return NVPYR_SUCCESS;
  for(int i = 1; i < n; ++i)
    if(texels(i) > kBatchSoloMaxTexels)
      return NVPYR_SUCCESS;
  if(!fastVectorOk<Srgba8>(a.lv))
    return NVPYR_SUCCESS;

  // Under stream capture the fused path is not taken: its base-pointer upload would be recorded as a copy from
  // host memory that is gone at replay time, and the ring slice would be baked into the graph.
  if(streamIsCapturing(a.stream))
    return NVPYR_SUCCESS;

  // chain base pointers -> a slice of the device ring, reserved until this batch's last kernel has completed
  const unsigned char** devBases = nullptr;
  cudaEvent_t           doneEvent = nullptr;
  {
    std::lock_guard<std::mutex> lock(ctx.batchMutex);
    auto reclaim = [&](bool wait) {
      while(!ctx.batchInFlight.empty())
      {
        DeviceContext::BatchSlice& f = ctx.batchInFlight.front();
        const cudaError_t q = wait ? cudaEventSynchronize(f.done) : cudaEventQuery(f.done);
        if(q != cudaSuccess)
        {
          cudaGetLastError();  // cudaErrorNotReady is not an error
          return;
        }
        ctx.batchEventPool.push_back(f.done);
        ctx.batchInFlight.pop_front();
        if(wait)
          return;
      }
    };
    reclaim(false);
    for(;;)
    {
      uint64_t begin = ctx.batchHead;
      if(begin % kBatchRing + count > kBatchRing)
        begin += kBatchRing - begin % kBatchRing;  // a slice never wraps: skip to the start of the ring
      const uint64_t tail = ctx.batchInFlight.empty() ? begin : ctx.batchInFlight.front().begin;
      if(begin + count - tail <= kBatchRing)
      {
        if(ctx.batchMaps == nullptr)  // first fused batch on this device
          NVPYR_CUDA(cudaMalloc(&ctx.batchMaps, size_t(kBatchRing) * sizeof(CUtensorMap)));
        ctx.batchHead = begin + count;
        devBases      = ctx.batchBases + begin % kBatchRing;
        if(ctx.batchEventPool.empty())
          NVPYR_CUDA(cudaEventCreateWithFlags(&doneEvent, cudaEventDisableTiming));
        else
        {
          doneEvent = ctx.batchEventPool.back();
          ctx.batchEventPool.pop_back();
        }
        ctx.batchInFlight.push_back({begin, begin + count, doneEvent});
        break;
      }
      reclaim(true);  // ring full of batches still in flight: wait for the oldest one instead of overwriting it
    }
  }
  // From here on every exit records doneEvent on the stream (a slice whose event was never recorded would be
  // reclaimed at once, which is right if nothing was enqueued, and the stream order covers what was).
  struct RecordDone
  {
    cudaEvent_t  e;
    cudaStream_t s;
    ~RecordDone() { cudaEventRecord(e, s); }
  } recordDone{doneEvent, a.stream};
  // The pointers travel as kernel-launch-sized chunks of a memcpy node fed from a std::vector: cudaMemcpyAsync from
  // pageable memory returns once the source has been staged, so the vector may die when this function returns.
  std::vector<const unsigned char*> hostBases(count);
  for(uint32_t i = 0; i < count; ++i)
    hostBases[i] = r[i].lv[0].ptr;
  NVPYR_CUDA(cudaMemcpyAsync(devBases, hostBases.data(), size_t(count) * sizeof(void*), cudaMemcpyHostToDevice, a.stream));
  *handled = true;  // from here on errors are real errors

  const bool premul = a.flags & NVPYR_FLAG_PREMULTIPLY_ALPHA;  // fused into the level-0 read of the big step
  // TMA staging of the batch kernel: one 2-D tensor map per image (level 0: W x H texels, box 64 x 8), encoded on the
  // host and uploaded next to the base pointers (the premultiplying variant rewrites level 0 and keeps the register path).
  const CUtensorMap* devMaps = nullptr;
#if NVPYR_FAST_TMA == 2
  if(!premul)
  {
    const TensorMapEncodeFn encode = tensorMapEncoder();
    if(encode == nullptr)
      return NVPYR_ERROR_UNSUPPORTED;
    std::vector<CUtensorMap> hostMaps(count);
    const cuuint64_t         dims[2]    = {a.lv[0].w, a.lv[0].h};
    const cuuint64_t         strides[1] = {a.lv[0].pitch};
    const cuuint32_t         box[2] = {64u, 8u}, estr[2] = {1u, 1u};
    const TensorMapReplaceFn replace = tensorMapReplacer();
    for(uint32_t i = 0; i < count; ++i)
    {
      // the maps differ in their base address only: encode the first one, then copy + replace the address
      if(i > 0 && replace != nullptr)
      {
        hostMaps[i] = hostMaps[0];
        if(replace(&hostMaps[i], r[i].lv[0].ptr) == CUDA_SUCCESS)
          continue;
      }
      if(encode(&hostMaps[i], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, r[i].lv[0].ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
         != CUDA_SUCCESS)
        return NVPYR_ERROR_CUDA;
    }
    CUtensorMap* dst = ctx.batchMaps + (devBases - ctx.batchBases);
    NVPYR_CUDA(cudaMemcpyAsync(dst, hostMaps.data(), size_t(count) * sizeof(CUtensorMap), cudaMemcpyHostToDevice, a.stream));
    devMaps = dst;
  }
#endif

  auto offsetView = [&](uint32_t level) {
    LevelView v = a.lv[level];
    v.ptr       = reinterpret_cast<unsigned char*>(size_t(a.lv[level].ptr - a.lv[0].ptr));
    return v;
  };
  FastParams p{};
  for(uint32_t k = 0; k <= steps[0].levelCount; ++k)
    p.lv[k] = offsetView(k);
  p.tables = ctx.tables;
  FastBatch   b{devBases, 0u, count, devMaps};
  nvpyrStatus st;
  switch(steps[0].levelCount)
  {
    case 2: st = launchFastBatchM<2>(ctx, p, a.stream, b, premul); break;
    case 3: st = launchFastBatchM<3>(ctx, p, a.stream, b, premul); break;
    case 4: st = launchFastBatchM<4>(ctx, p, a.stream, b, premul); break;
    case 5: st = launchFastBatchM<5>(ctx, p, a.stream, b, premul); break;
    default: st = launchFastBatchM<6>(ctx, p, a.stream, b, premul); break;
  }
  if(st != NVPYR_SUCCESS || n == 1)
    return st;

  using TF = Srgba8Lite;
  TailParams tp{};
  tp.numSteps = uint32_t(n - 1);
  tp.tables   = ctx.tables;
  for(int i = 1; i < n; ++i)
  {
    const nvpyrPlanStep& s  = steps[i];
    TailStep&            ts = tp.steps[i - 1];
    ts.pipeline             = s.pipeline;
    ts.levels               = s.levelCount;
    LevelView real[7];
    for(uint32_t k = 0; k <= s.levelCount; ++k)
    {
      real[k]  = a.lv[s.inputLevel + k];
      ts.lv[k] = offsetView(s.inputLevel + k);
    }
    if(s.pipeline == 1)
    {
      ts.vec    = (fastVectorOk<Srgba8>(real) ? 1u : 0u) | fastTinyBit(ts);  // every base is 16-byte aligned: same answer for all images
      ts.tilesX = (ts.lv[0].w + 63u) / 64u;
      ts.tilesY = (ts.lv[0].h + 63u) / 64u;
    }
    else
    {
      generalTiles(ts.lv, s.levelCount, SoloTile2<TF::Value>::value, &ts.tilesX, &ts.tilesY);  // every batch-tail step is solo
      fillTailStepReciprocals(ts);
    }
  }
  const size_t smem = sizeof(TailSmem<TF>);
  int          grid = 1;
  st                = persistentGrid(tailBatchKernel<TF>, smem, ctx, count, &grid, kTailThreads);
  if(st != NVPYR_SUCCESS)
    return st;
  NVPYR_CUDA(launchKernel(tailBatchKernel<TF>, grid, kTailThreads, smem, a.stream, tp,
                          static_cast<const unsigned char* const*>(devBases), count));
  ++g_launchCount;
  return NVPYR_SUCCESS;
}

// ------------------------------------------------- host round trip, pipelined
// One independent round-trip pipeline: device scratch chain, three streams (upload, compute, download), the events
// that chain them band by band, and -- for callers with pageable memory -- a pinned staging chain.  A call takes an
// idle pipeline from the device's pool (or makes a new one) and gives it back when it returns, so concurrent
// nvpyrGenerateHost calls on one device run side by side.
struct HostPipeline
{
  void*          scratch = nullptr;
  size_t         scratchBytes = 0;
  cudaStream_t   up = nullptr, run = nullptr, down = nullptr;
  cudaEvent_t    evUp[kMaxHostBands] = {}, evRun[kMaxHostBands + 1] = {}, evDown[kMaxHostBands + 1] = {};
  unsigned char* stage = nullptr;  // pinned, stageBytes
  size_t         stageBytes = 0;
};

void destroyHostPipeline(HostPipeline* hp)
{
  if(hp == nullptr)
    return;
  if(hp->scratch)
    cudaFree(hp->scratch);
  if(hp->stage)
    cudaFreeHost(hp->stage);
  for(cudaStream_t st : {hp->up, hp->run, hp->down})
    if(st)
      cudaStreamDestroy(st);
  for(cudaEvent_t e : hp->evUp)
    if(e)
      cudaEventDestroy(e);
  for(cudaEvent_t e : hp->evRun)
    if(e)
      cudaEventDestroy(e);
  for(cudaEvent_t e : hp->evDown)
    if(e)
      cudaEventDestroy(e);
  delete hp;
}

nvpyrStatus createHostPipeline(HostPipeline** out)
{
  HostPipeline* hp = new HostPipeline;
  auto          fail = [&](cudaError_t e) {
    g_lastCudaError = int(e);
    destroyHostPipeline(hp);
    return NVPYR_ERROR_CUDA;
  };
  cudaError_t e;
  for(cudaStream_t* st : {&hp->up, &hp->run, &hp->down})
    if((e = cudaStreamCreateWithFlags(st, cudaStreamNonBlocking)) != cudaSuccess)
      return fail(e);
  for(uint32_t i = 0; i < kMaxHostBands; ++i)
    if((e = cudaEventCreateWithFlags(&hp->evUp[i], cudaEventDisableTiming)) != cudaSuccess)
      return fail(e);
  for(uint32_t i = 0; i <= kMaxHostBands; ++i)
  {
    if((e = cudaEventCreateWithFlags(&hp->evRun[i], cudaEventDisableTiming)) != cudaSuccess)
      return fail(e);
    if((e = cudaEventCreateWithFlags(&hp->evDown[i], cudaEventDisableTiming)) != cudaSuccess)
      return fail(e);
  }
  *out = hp;
  return NVPYR_SUCCESS;
}

// Takes an idle pipeline whose scratch chain holds `bytes` (growing it if need be).
nvpyrStatus acquireHostPipeline(DeviceContext& ctx, uint64_t bytes, HostPipeline** out)
{
  HostPipeline* hp = nullptr;
  {
    std::lock_guard<std::mutex> lock(ctx.hostMutex);
    for(size_t i = 0; i < ctx.hostIdle.size(); ++i)  // prefer one that is already large enough
      if(ctx.hostIdle[i]->scratchBytes >= bytes)
      {
        hp = ctx.hostIdle[i];
        ctx.hostIdle.erase(ctx.hostIdle.begin() + i);
        break;
      }
    if(hp == nullptr && !ctx.hostIdle.empty())
    {
      hp = ctx.hostIdle.back();
      ctx.hostIdle.pop_back();
    }
  }
  if(hp == nullptr)
  {
    nvpyrStatus st = createHostPipeline(&hp);
    if(st != NVPYR_SUCCESS)
      return st;
  }
  if(hp->scratchBytes < bytes)
  {
    if(hp->scratch)
      cudaFree(hp->scratch);
    hp->scratch      = nullptr;
    hp->scratchBytes = 0;
    const cudaError_t e = cudaMalloc(&hp->scratch, bytes);
    if(e != cudaSuccess)
    {
      g_lastCudaError = int(e);
      destroyHostPipeline(hp);
      return e == cudaErrorMemoryAllocation ? NVPYR_ERROR_OUT_OF_MEMORY : NVPYR_ERROR_CUDA;
    }
    hp->scratchBytes = bytes;
  }
  *out = hp;
  return NVPYR_SUCCESS;
}

void releaseHostPipeline(DeviceContext& ctx, HostPipeline* hp)
{
  std::lock_guard<std::mutex> lock(ctx.hostMutex);
  ctx.hostIdle.push_back(hp);
}

// Band height (rows of level 0) of the pipelined round trip; 0 = do not band.  NVPYR_HOST_BAND_BYTES
// overrides the target band size (0 disables banding).
const uint64_t kHostBandBytes = [] {
  const char* e = getenv("NVPYR_HOST_BAND_BYTES");
  return e != nullptr ? uint64_t(strtoull(e, nullptr, 10)) : 32ull << 20;
}();
// Levels that the pipelined round trip downloads band by band (NVPYR_HOST_BAND_LEVELS, A/B timing).
const uint32_t g_hostBandLevels = [] {
  const char* e = getenv("NVPYR_HOST_BAND_LEVELS");
  return e != nullptr ? std::max(1u, uint32_t(atoi(e))) : 6u;
}();
// NVPYR_HOST_TAPER=0: the last band of the pipelined round trip is not cut into smaller ones (A/B timing).
const bool g_hostTaper = [] {
  const char* e = getenv("NVPYR_HOST_TAPER");
  return !(e != nullptr && e[0] == '0');
}();
// Host threads that move a pageable caller's bands into / out of the pinned staging chain (NVPYR_HOST_COPY_THREADS).
const unsigned kHostCopyThreads = [] {
  const char* e = getenv("NVPYR_HOST_COPY_THREADS");
  if(e != nullptr)
    return unsigned(std::max(1, atoi(e)));
  const unsigned hw = std::thread::hardware_concurrency();
  return std::max(1u, std::min(8u, hw ? hw : 4u));
}();

// memcpy split over host threads (one thread cannot saturate the host memory system: ~10 GB/s against a PCIe 5 link
// that takes 50+).
void parallelCopy(void* dst, const void* src, size_t bytes)
{
  const size_t   kMinChunk = 4u << 20;
  const unsigned n = unsigned(std::min<size_t>(kHostCopyThreads, std::max<size_t>(1, bytes / kMinChunk)));
  if(n <= 1)
  {
    memcpy(dst, src, bytes);
    return;
  }
  const size_t             chunk = ((bytes + n - 1) / n + 4095) & ~size_t(4095);
  std::vector<std::thread> ts;
  for(unsigned i = 1; i < n; ++i)
  {
    const size_t off = size_t(i) * chunk;
    if(off < bytes)
      ts.emplace_back([=] { memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, std::min(chunk, bytes - off)); });
  }
  memcpy(dst, src, std::min(chunk, bytes));
  for(std::thread& t : ts)
    t.join();
}

// Is this host pointer pageable (neither cudaHostAlloc'ed nor cudaHostRegister'ed)?  Copies from / to such memory are
// staged by the driver and block the calling thread, which would serialise the three streams of the pipeline.
bool isPageable(const void* p)
{
  cudaPointerAttributes a;
  if(cudaPointerGetAttributes(&a, p) != cudaSuccess)
  {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// Upload, generate, download -- the shape of minimal_app (minimal_mipmaps.cpp:134-217) -- with the three
// stages overlapped.  When the chain starts with a fast-pipeline step of M >= 2 levels, level 0 is cut
// into bands of whole tile rows (a multiple of 2^M rows: no texel of levels 1..M depends on two bands,
// and the float32 expression trees are untouched).  Band b is uploaded on stream `up`; the step runs on it on
// `run` as soon as it has arrived; its rows of levels 1 and 2 (and 0, unless the caller's chain
// already holds it) go back on `down` while band b+1 is still arriving: both PCIe directions and the
// SMs are busy at once.  The rest of the plan and the small levels follow in one piece.
// In place (hostChain == hostLevel0, the reference's single staging buffer, scoped_image.hpp:436-453)
// level 0 is not downloaded again unless the premultiply pre-pass changed it.
// Pageable callers (a std::vector, malloc): the bands pass through the pipeline's pinned staging chain, moved by a
// few host threads while the previous band is on the wire, so the device-side overlap is the same.
template <class F>
nvpyrStatus generateHostPipelined(DeviceContext& ctx, HostPipeline& hp, const ResolvedDesc& r, const void* hostLevel0,
                                  void* hostChain, uint64_t chainBytes)
{
  const unsigned char* hin      = static_cast<const unsigned char*>(hostLevel0);
  unsigned char*       hout     = static_cast<unsigned char*>(hostChain);
  unsigned char*       dev      = r.lv[0].ptr;
  const uint64_t       rowBytes = uint64_t(r.w) * r.texelBytes, level0Bytes = rowBytes * r.h;
  const bool           premul   = r.flags & NVPYR_FLAG_PREMULTIPLY_ALPHA;
  const bool           level0Back = premul || hostChain != hostLevel0;

  nvpyrPlanStep steps[NVPYR_MAX_STEPS];
  const int     n = r.levels > 1 ? buildPlan(r.w, r.h, r.levels, r.general, r.fast, steps, NVPYR_MAX_STEPS) : 0;
  if(n < 0)
    return NVPYR_ERROR_INVALID_VALUE;

  uint32_t bandRows = 0, M = 0, bandUnit = 0;
  if(n > 0 && steps[0].pipeline == 1 && steps[0].levelCount >= 2 && kHostBandBytes != 0 && level0Bytes >= 2 * kHostBandBytes)
  {
    M                   = steps[0].levelCount;
    const uint32_t unit = std::max(8u, 1u << M);  // tile height of the fast kernels
    bandUnit            = unit;
    uint64_t       rows = std::max<uint64_t>(1, kHostBandBytes / rowBytes);
    rows                = (rows + unit - 1) / unit * unit;
    const uint64_t minRows = (uint64_t(r.h) + kMaxHostBands - 1) / kMaxHostBands;
    if(rows < minRows)
      rows = (minRows + unit - 1) / unit * unit;
    bandRows = uint32_t(std::min<uint64_t>(rows, r.h));
  }

  if(bandRows == 0 || bandRows >= r.h)
  {
    // One piece: upload, whole plan, download (pageable memory: the driver stages these few megabytes).
    NVPYR_CUDA(cudaMemcpyAsync(dev, hin, level0Bytes, cudaMemcpyHostToDevice, hp.run));
    nvpyrStatus st = dispatchResolved(r);
    if(st != NVPYR_SUCCESS)
      return st;
    const uint64_t from = level0Back ? 0 : level0Bytes;
    if(chainBytes > from)
      NVPYR_CUDA(cudaMemcpyAsync(hout + from, dev + from, chainBytes - from, cudaMemcpyDeviceToHost, hp.run));
    return NVPYR_SUCCESS;
  }

  // Pageable source / destination: stage through pinned memory (same offsets as the chain).
  const bool stageIn = isPageable(hin), stageOut = isPageable(hout);
  if((stageIn || stageOut) && hp.stageBytes < chainBytes)
  {
    if(hp.stage)
      cudaFreeHost(hp.stage);
    hp.stage      = nullptr;
    hp.stageBytes = 0;
    NVPYR_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&hp.stage), chainBytes, cudaHostAllocDefault));
    hp.stageBytes = chainBytes;
  }
  const unsigned char* upSrc   = stageIn ? hp.stage : hin;
  unsigned char*       downDst = stageOut ? hp.stage : hout;

  // Levels 1 .. bandLevels travel back band by band, the small rest at the end.  In place (the upload is the
  // bottleneck) every level of the step does: what is left for the end of a 16384^2 chain is then 87 KB instead of 22 MB,
  // 0.35 ms less after the last upload (20.55 -> 20.2 ms).  When level 0 goes back too the download is the bottleneck and
  // small pieces between its 32 MB bands only slow it down (29.8 -> 31 ms with four levels): two levels as before.
  const uint32_t bandLevels = std::min(M, level0Back ? std::min(2u, g_hostBandLevels) : g_hostBandLevels);
  struct OutPiece
  {
    size_t off, bytes;
  };
  std::vector<OutPiece> outPieces[kMaxHostBands + 1];  // what band b's download event covers
  // Pageable destination: a helper thread follows the download events and moves each band out of the staging chain
  // while the main thread is still feeding later bands in (both directions of the host copy run at once).
  std::atomic<uint32_t>    recorded{0};      // download events recorded so far
  std::atomic<bool>        stop{false};      // main thread gave up (error): the helper must not wait for more events
  std::atomic<cudaError_t> drainError{cudaSuccess};
  uint32_t                 totalEvents = 0;  // set before the last event is published
  std::thread              drainer;
  struct JoinDrainer
  {
    std::thread&       t;
    std::atomic<bool>& stop;
    ~JoinDrainer()
    {
      stop = true;
      if(t.joinable())
        t.join();
    }
  } joinDrainer{drainer, stop};
  if(stageOut)
    drainer = std::thread([&] {
      cudaSetDevice(ctx.device);
      for(uint32_t b = 0;; ++b)
      {
        while(recorded.load(std::memory_order_acquire) <= b)
        {
          if(stop.load())
            return;
          std::this_thread::yield();
        }
        const cudaError_t q = cudaEventSynchronize(hp.evDown[b]);
        if(q != cudaSuccess)
        {
          drainError = q;
          return;
        }
        for(const OutPiece& o : outPieces[b])
          parallelCopy(hout + o.off, hp.stage + o.off, o.bytes);
        if(totalEvents != 0 && b + 1 == totalEvents)
          return;
      }
    });

  // In place the LAST band is cut into halves of halves (down to one tile row): what remains to be done after the last
  // upload -- that band's kernel and the download of its levels -- shrinks with it.
  const bool taper = !level0Back && g_hostTaper && (r.h + bandRows - 1) / bandRows + 10u <= kMaxHostBands;
  uint32_t   band = 0, rows = 0;
  for(uint32_t row0 = 0; row0 < r.h; row0 += rows, ++band)
  {
    const uint32_t remaining = r.h - row0;
    rows                     = std::min(bandRows, remaining);
    if(taper && remaining <= bandRows && remaining > bandUnit)
      rows = std::max(bandUnit, remaining / 2u / bandUnit * bandUnit);
    if(stageIn)
      parallelCopy(hp.stage + row0 * rowBytes, hin + row0 * rowBytes, rows * rowBytes);
    NVPYR_CUDA(cudaMemcpyAsync(dev + row0 * rowBytes, upSrc + row0 * rowBytes, rows * rowBytes, cudaMemcpyHostToDevice, hp.up));
    NVPYR_CUDA(cudaEventRecord(hp.evUp[band], hp.up));
    NVPYR_CUDA(cudaStreamWaitEvent(hp.run, hp.evUp[band], 0));
    FastParams p{};
    for(uint32_t k = 0; k <= M; ++k)
    {
      p.lv[k]     = r.lv[k];
      p.lv[k].ptr = r.lv[k].ptr + size_t(row0 >> k) * r.lv[k].pitch;
      p.lv[k].h   = rows >> k;  // rows and row0 are multiples of 2^M (H is, too)
    }
    const bool fusePremul = premul && tunedFastOk<F>(p.lv);
    if(premul && !fusePremul)
    {
      nvpyrStatus st = launchPremultiply(ctx, dev + row0 * rowBytes, dev + row0 * rowBytes, uint64_t(rows) * r.w, hp.run);
      if(st != NVPYR_SUCCESS)
        return st;
    }
    nvpyrStatus st = launchFast<F>(ctx, p, M, hp.run, fusePremul);
    if(st != NVPYR_SUCCESS)
      return st;
    NVPYR_CUDA(cudaEventRecord(hp.evRun[band], hp.run));
    NVPYR_CUDA(cudaStreamWaitEvent(hp.down, hp.evRun[band], 0));
    for(uint32_t k = level0Back ? 0u : 1u; k <= bandLevels; ++k)
    {
      const size_t off = size_t(p.lv[k].ptr - dev), sz = size_t(p.lv[k].h) * p.lv[k].pitch;
      NVPYR_CUDA(cudaMemcpyAsync(downDst + off, dev + off, sz, cudaMemcpyDeviceToHost, hp.down));
      outPieces[band].push_back({off, sz});
    }
    NVPYR_CUDA(cudaEventRecord(hp.evDown[band], hp.down));
    recorded.store(band + 1, std::memory_order_release);
  }
  // the rest of the plan, then levels 3.. in one piece
  nvpyrStatus st = runPlan<F>(ctx, r, 1);
  if(st != NVPYR_SUCCESS)
    return st;
  NVPYR_CUDA(cudaEventRecord(hp.evRun[kMaxHostBands], hp.run));
  NVPYR_CUDA(cudaStreamWaitEvent(hp.down, hp.evRun[kMaxHostBands], 0));
  if(r.levels > bandLevels + 1)
  {
    const size_t off = size_t(r.lv[bandLevels + 1].ptr - dev);
    NVPYR_CUDA(cudaMemcpyAsync(downDst + off, dev + off, chainBytes - off, cudaMemcpyDeviceToHost, hp.down));
    outPieces[band].push_back({off, size_t(chainBytes - off)});
  }
  NVPYR_CUDA(cudaEventRecord(hp.evDown[band], hp.down));
  totalEvents = band + 1;
  recorded.store(band + 1, std::memory_order_release);
  if(stageOut)
  {
    drainer.join();
    NVPYR_CUDA(drainError.load());
  }
  return NVPYR_SUCCESS;
}

}  // namespace
}  // namespace nvpyr

using namespace nvpyr;

// ============================================================================
extern "C" {

uint32_t nvpyrGetLevelCount(nvpyrExtent2D e)
{
  return levelCountFor(e.width, e.height);
}

nvpyrStatus nvpyrGetLevelExtent(nvpyrExtent2D e, uint32_t level, nvpyrExtent2D* out)
{
  if(out == nullptr || e.width == 0 || e.height == 0 || level >= levelCountFor(e.width, e.height))
    return NVPYR_ERROR_INVALID_VALUE;
  out->width  = levelDim(e.width, level);
  out->height = levelDim(e.height, level);
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrGetLevelOffsetTexels(nvpyrExtent2D e, uint32_t level, uint64_t* out)
{
  if(out == nullptr || e.width == 0 || e.height == 0 || level > levelCountFor(e.width, e.height))
    return NVPYR_ERROR_INVALID_VALUE;
  *out = levelOffsetTexels(e.width, e.height, level);
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrGetChainBytes(nvpyrExtent2D e, uint32_t levelCount, nvpyrFormat format, uint64_t* out)
{
  if(out == nullptr || e.width == 0 || e.height == 0)
    return NVPYR_ERROR_INVALID_VALUE;
  if(format != NVPYR_FORMAT_SRGBA8 && format != NVPYR_FORMAT_RGBA32F)
    return NVPYR_ERROR_UNSUPPORTED;
  const uint32_t maxLevels = levelCountFor(e.width, e.height);
  if(levelCount == 0)
    levelCount = maxLevels;
  if(levelCount > maxLevels)
    return NVPYR_ERROR_INVALID_VALUE;
  *out = levelOffsetTexels(e.width, e.height, levelCount) * (format == NVPYR_FORMAT_SRGBA8 ? 4u : 16u);
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrGetPlan(nvpyrExtent2D e, uint32_t levelCount, const nvpyrPlanOptions* options, nvpyrPlanStep* steps,
                         uint32_t maxSteps, uint32_t* count)
{
  if(steps == nullptr || count == nullptr || e.width == 0 || e.height == 0)
    return NVPYR_ERROR_INVALID_VALUE;
  const uint32_t maxLevels = levelCountFor(e.width, e.height);
  if(levelCount > maxLevels)
    return NVPYR_ERROR_INVALID_VALUE;
  nvpyrPlanOptions o{};
  if(options)
    o = *options;
  dispatcher_t fast = nullptr;
  if(!(o.flags & NVPYR_FLAG_FORCE_GENERAL))
  {
    fast = selectFastDispatcher(o.fastDivisibility, o.fastMaxLevels);
    if(fast == nullptr)
      return NVPYR_ERROR_UNSUPPORTED;
  }
  const dispatcher_t general = (o.flags & NVPYR_FLAG_GENERAL_BLIT) ? blitDispatcher : defaultGeneralDispatcher;
  const int          n       = buildPlan(e.width, e.height, levelCount, general, fast, steps, maxSteps);
  if(n < 0)
    return NVPYR_ERROR_INVALID_VALUE;
  *count = uint32_t(n);
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrDispatchEx(const nvpyrDispatchDesc* desc)
{
  ResolvedDesc r;
  nvpyrStatus  st = resolve(desc, r);
  if(st != NVPYR_SUCCESS)
    return st;
  return dispatchResolved(r);
}

// Limits of the kernels for a plan made by user dispatchers (the defaults satisfy them by construction).
nvpyrStatus checkCustomPlan(const ResolvedDesc& r)
{
  nvpyrPlanStep steps[NVPYR_MAX_STEPS];
  const int     n = r.levels > 1 ? buildPlan(r.w, r.h, r.levels, r.general, r.fast, steps, NVPYR_MAX_STEPS) : 0;
  if(n < 0)
    return NVPYR_ERROR_INVALID_VALUE;  // 0 levels, more levels than remain, or too many dispatches
  for(int i = 0; i < n; ++i)
  {
    const nvpyrPlanStep& s = steps[i];
    if(s.pipeline == 1)
    {
      if(s.levelCount < 1 || s.levelCount > 6 || ((s.srcWidth | s.srcHeight) & ((1u << s.levelCount) - 1u)))
        return NVPYR_ERROR_INVALID_VALUE;
    }
    else if(s.levelCount < 1 || s.levelCount > 2)
      return NVPYR_ERROR_INVALID_VALUE;
  }
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrDispatchWithDispatchers(const nvpyrDispatchDesc* desc, nvpyrDispatcher general, nvpyrDispatcher fast,
                                         void* userData)
{
  ResolvedDesc r;
  nvpyrStatus  st = resolve(desc, r);
  if(st != NVPYR_SUCCESS)
    return st;
  if(general != nullptr)
    r.general = DispatcherRef(general, userData);
  if(fast != nullptr && !(r.flags & NVPYR_FLAG_FORCE_GENERAL))
    r.fast = DispatcherRef(fast, userData);
  r.customDispatchers = general != nullptr || fast != nullptr;
  // The callbacks are called once here (validation) and once more when the plan is executed: like the reference's,
  // they must be pure functions of the state they are given.
  st = checkCustomPlan(r);
  if(st != NVPYR_SUCCESS)
    return st;
  return dispatchResolved(r);
}

nvpyrStatus nvpyrDispatch(void* srcLevel0, uint32_t levelCount, nvpyrExtent2D extent, nvpyrStream stream)
{
  nvpyrDispatchDesc d;
  memset(&d, 0, sizeof d);
  d.structSize = sizeof d;
  d.format     = NVPYR_FORMAT_SRGBA8;
  d.extent     = extent;
  d.levelCount = levelCount;
  d.base       = srcLevel0;
  d.stream     = stream;
  return nvpyrDispatchEx(&d);
}

nvpyrStatus nvpyrDispatchBatch(const nvpyrDispatchDesc* descs, uint32_t count)
{
  if(descs == nullptr && count != 0)
    return NVPYR_ERROR_INVALID_VALUE;
  // Validate everything before enqueueing anything.
  std::vector<ResolvedDesc> r(count);
  for(uint32_t i = 0; i < count; ++i)
  {
    nvpyrStatus st = resolve(&descs[i], r[i]);
    if(st != NVPYR_SUCCESS)
      return st;
  }
  if(count == 0)
    return NVPYR_SUCCESS;
  DeviceContext* ctx = nullptr;
  nvpyrStatus    st  = getContext(&ctx);
  if(st != NVPYR_SUCCESS)
    return st;
  bool fused = false;
  st         = dispatchBatchFused(*ctx, r, &fused);
  if(fused || st != NVPYR_SUCCESS)
    return st;
  for(uint32_t i = 0; i < count; ++i)
  {
    st = dispatchResolved(r[i]);
    if(st != NVPYR_SUCCESS)
      return st;
  }
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrPremultiplyAlpha(const void* in, void* out, uint64_t texels, nvpyrStream stream)
{
  if(in == nullptr || out == nullptr)
    return NVPYR_ERROR_INVALID_VALUE;
  if(reinterpret_cast<uintptr_t>(in) % 4u || reinterpret_cast<uintptr_t>(out) % 4u)
    return NVPYR_ERROR_INVALID_VALUE;
  if(texels == 0)
    return NVPYR_SUCCESS;
  DeviceContext* ctx = nullptr;
  nvpyrStatus    st  = getContext(&ctx);
  if(st != NVPYR_SUCCESS)
    return st;
  return launchPremultiply(*ctx, in, out, texels, reinterpret_cast<cudaStream_t>(stream));
}

nvpyrStatus nvpyrGenerateHost(const void* hostLevel0, void* hostChain, nvpyrExtent2D extent, uint32_t levelCount,
                              nvpyrFormat format, uint32_t flags)
{
  if(hostLevel0 == nullptr || hostChain == nullptr)
    return NVPYR_ERROR_INVALID_VALUE;
  uint64_t    bytes = 0;
  nvpyrStatus st    = nvpyrGetChainBytes(extent, levelCount, format, &bytes);
  if(st != NVPYR_SUCCESS)
    return st;
  DeviceContext* ctx = nullptr;
  st                 = getContext(&ctx);
  if(st != NVPYR_SUCCESS)
    return st;
  HostPipeline* hp = nullptr;
  st               = acquireHostPipeline(*ctx, bytes, &hp);
  if(st != NVPYR_SUCCESS)
    return st;
  nvpyrDispatchDesc d;
  memset(&d, 0, sizeof d);
  d.structSize = sizeof d;
  d.format     = format;
  d.flags      = flags;
  d.extent     = extent;
  d.levelCount = levelCount;
  d.base       = hp->scratch;
  d.stream     = reinterpret_cast<nvpyrStream>(hp->run);
  ResolvedDesc r;
  st = resolve(&d, r);
  if(st == NVPYR_SUCCESS)
  {
    if(r.flags & NVPYR_FLAG_F16_SHARED)
      st = generateHostPipelined<Srgba8F16Shared>(*ctx, *hp, r, hostLevel0, hostChain, bytes);
    else if(r.flags & NVPYR_FLAG_SRGB_SHARED)
      st = generateHostPipelined<Srgba8SrgbShared>(*ctx, *hp, r, hostLevel0, hostChain, bytes);
    else
      st = r.format == NVPYR_FORMAT_SRGBA8 ? generateHostPipelined<Srgba8>(*ctx, *hp, r, hostLevel0, hostChain, bytes)
                                           : generateHostPipelined<Rgba32f>(*ctx, *hp, r, hostLevel0, hostChain, bytes);
  }
  // Never leave work in flight on the pipeline's scratch chain, success or not.
  const cudaError_t e0 = cudaStreamSynchronize(hp->up), e1 = cudaStreamSynchronize(hp->run),
                    e2 = cudaStreamSynchronize(hp->down);
  releaseHostPipeline(*ctx, hp);
  if(st != NVPYR_SUCCESS)
    return st;
  NVPYR_CUDA(e0);
  NVPYR_CUDA(e1);
  NVPYR_CUDA(e2);
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrInit(void)
{
  DeviceContext* ctx = nullptr;
  return getContext(&ctx);
}

struct nvpyrExternalMemory_t
{
  cudaExternalMemory_t mem;
  void*                ptr;
};

nvpyrStatus nvpyrImportExternalMemoryFd(int fd, uint64_t allocationSize, uint64_t offset, uint64_t size,
                                        nvpyrExternalMemory* outHandle, void** outDevicePtr)
{
  if(fd < 0 || outHandle == nullptr || outDevicePtr == nullptr || size == 0 || offset + size > allocationSize)
    return NVPYR_ERROR_INVALID_VALUE;
  cudaExternalMemoryHandleDesc hd;
  memset(&hd, 0, sizeof hd);
  hd.type      = cudaExternalMemoryHandleTypeOpaqueFd;
  hd.handle.fd = fd;
  hd.size      = allocationSize;
  cudaExternalMemory_t mem;
  NVPYR_CUDA(cudaImportExternalMemory(&mem, &hd));
  cudaExternalMemoryBufferDesc bd;
  memset(&bd, 0, sizeof bd);
  bd.offset = offset;
  bd.size   = size;
  void*       ptr = nullptr;
  cudaError_t e   = cudaExternalMemoryGetMappedBuffer(&ptr, mem, &bd);
  if(e != cudaSuccess)
  {
    cudaDestroyExternalMemory(mem);
    g_lastCudaError = int(e);
    return NVPYR_ERROR_CUDA;
  }
  nvpyrExternalMemory h = new nvpyrExternalMemory_t{mem, ptr};
  *outHandle            = h;
  *outDevicePtr         = ptr;
  return NVPYR_SUCCESS;
}

nvpyrStatus nvpyrReleaseExternalMemory(nvpyrExternalMemory handle)
{
  if(handle == nullptr)
    return NVPYR_ERROR_INVALID_VALUE;
  cudaFree(handle->ptr);
  cudaError_t e = cudaDestroyExternalMemory(handle->mem);
  delete handle;
  if(e != cudaSuccess)
  {
    g_lastCudaError = int(e);
    return NVPYR_ERROR_CUDA;
  }
  return NVPYR_SUCCESS;
}

const char* nvpyrGetErrorString(nvpyrStatus status)
{
  switch(status)
  {
    case NVPYR_SUCCESS: return "NVPYR_SUCCESS";
    case NVPYR_ERROR_INVALID_VALUE: return "NVPYR_ERROR_INVALID_VALUE";
    case NVPYR_ERROR_UNSUPPORTED: return "NVPYR_ERROR_UNSUPPORTED";
    case NVPYR_ERROR_CUDA: return "NVPYR_ERROR_CUDA";
    case NVPYR_ERROR_OUT_OF_MEMORY: return "NVPYR_ERROR_OUT_OF_MEMORY";
    case NVPYR_ERROR_IO: return "NVPYR_ERROR_IO";
  }
  return "NVPYR_ERROR_UNKNOWN";
}

int nvpyrGetLastCudaError(void)
{
  return g_lastCudaError;
}

uint64_t nvpyrGetLaunchCount(void)
{
  return g_launchCount.load();
}

uint64_t nvpyrSelfTestEncodeTable(void)
{
  return checkRowEncodeTable();
}

nvpyrStatus nvpyrShutdown(void)
{
  std::lock_guard<std::mutex> lock(g_ctxMutex);
  int                         prev = 0;
  cudaGetDevice(&prev);
  for(DeviceContext* c : g_ctx)
  {
    cudaSetDevice(c->device);
    cudaFree(c->tables);
    cudaFree(c->tickets);
    cudaFree(c->batchBases);
    cudaFree(c->batchMaps);
    for(HostPipeline* hp : c->hostIdle)
      destroyHostPipeline(hp);
    for(DeviceContext::BatchSlice& f : c->batchInFlight)
      cudaEventDestroy(f.done);
    for(cudaEvent_t e : c->batchEventPool)
      cudaEventDestroy(e);
    delete c;
  }
  g_ctx.clear();
  cudaSetDevice(prev);
  return NVPYR_SUCCESS;
}

uint32_t nvpyrGetVersion(void)
{
  return NVPYR_VERSION;
}

}  // extern "C"
