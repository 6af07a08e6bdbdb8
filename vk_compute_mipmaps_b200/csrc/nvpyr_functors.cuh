// nvpyr_functors.cuh -- the user-customisable part of the pyramid generator.
//
// The reference configures its GLSL template with preprocessor macros
// (nvpro_pyramid/nvpro_pyramid.glsl:27-120): NVPRO_PYRAMID_TYPE, _LOAD, _REDUCE,
// _REDUCE2, _REDUCE4, _LOAD_REDUCE4, _STORE, _SHARED_*.  Here the same contract
// is a C++ "functor set": a struct with static __device__ members that the
// kernel templates of nvpyr_kernels.cuh are instantiated with.
//
//   struct Functors {
//     using Value = float4;                       // NVPRO_PYRAMID_TYPE
//     static constexpr int kTexelBytes;           // storage bytes per texel
//     struct Shared;                              // per-CTA tables in shared memory
//     static void sharedInit(Shared&, const DeviceTables*);  // cooperative, once per CTA
//     static Value load(const Shared&, const void* texel);   // NVPRO_PYRAMID_LOAD
//     static void  load4(const Shared&, const void* p, Value out[4]);   // 4 texels in a row, 16B-aligned
//     static void  store(const Shared&, void* texel, Value);            // NVPRO_PYRAMID_STORE
//     static void  store2(const Shared&, void* p, Value, Value);        // 2 texels in a row, aligned
//     static Value reduce(float a0, Value v0, float a1, Value v1, float a2, Value v2);  // _REDUCE
//     static Value reduce2(Value, Value);                               // _REDUCE2
//     static Value reduce4(Value v00, Value v01, Value v10, Value v11); // _REDUCE4
//     static Value sharedRound(Value);   // _SHARED_STORE followed by _SHARED_LOAD (identity by default)
//   };
//
// Shipped instances: Srgba8 (nvpro_pyramid/srgba8_mipmap_preamble.glsl) and
// Rgba32f (identity load/store).  All float arithmetic uses the *_rn intrinsics
// so that nvcc never decides about contraction: the numerics contract
// is float32, round-to-nearest, the pairing order of the reference shader, and
// exactly one explicit contraction (the 3-tap REDUCE = mul, fma, fma).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nvpyr {

// ---------------------------------------------------------------------------
// sRGB encode bucket table.
//
// srgbFromLinear(x) (shaders/srgb.h:30-41) is monotone, so it is fully described
// by 255 float thresholds (srgb_tables.inc).  To evaluate it without a search:
// the float's upper bits (sign/exponent + 8 mantissa bits) select a bucket that
// contains at most ONE threshold; the entry holds the code at the bucket's lower
// edge and where inside the bucket the next code starts, pre-biased so that
//     t = entry[bits >> 15] + bits;   code = (t >> 16) & 0xFF
// (one shift, one table read, one add).  bits must be clamped to
// [kEncMinBits, kEncMaxBits] first.
constexpr uint32_t kEncShift   = 15;
constexpr uint32_t kEncMinBits = 0x39000000u;  // 2^-13 < threshold of code 1
constexpr uint32_t kEncMaxBits = 0x3F800000u;  // 1.0f -> code 255
constexpr uint32_t kEncMinKey  = kEncMinBits >> kEncShift;
constexpr uint32_t kEncMaxKey  = kEncMaxBits >> kEncShift;
constexpr uint32_t kEncEntries = kEncMaxKey - kEncMinKey + 1;  // 3329

constexpr uint32_t kEncEntriesPadded = (kEncEntries + 3u) & ~3u;  // whole uint4s for the copy into shared memory

// The tuned fast kernel (nvpyr_fast_srgba8.cuh) keys the same kind of table on 7 mantissa bits: with the pinned
// thresholds a bucket of 2^16 float patterns still holds at most one (checked when the table is built), and the
// table is half as long -- which pays for twice as many bank-partitioned copies of every entry.
#ifndef NVPYR_FAST_ENC_SHIFT
#define NVPYR_FAST_ENC_SHIFT 16
#endif
constexpr uint32_t kFastEncShift   = NVPYR_FAST_ENC_SHIFT;
constexpr uint32_t kFastEncMinKey  = kEncMinBits >> kFastEncShift;
constexpr uint32_t kFastEncMaxKey  = kEncMaxBits >> kFastEncShift;
constexpr uint32_t kFastEncEntries = kFastEncMaxKey - kFastEncMinKey + 1;
constexpr uint32_t kFastEncEntriesPadded = (kFastEncEntries + 3u) & ~3u;

// Row table of the tuned fast kernel (round 2).  The bucket of a value x in [0, 1] is taken from the float
//     z = RN(x + kRowEncC),   row = (bits(z) >> 16) - kRowEncFirstKey        (sign/exponent + 7 mantissa bits of z)
// instead of from x itself.  Adding the constant compresses the twelve low octaves of x, where the sRGB thresholds
// are far apart, into the linear part of z's first octave, so 645 rows cover [0, 1] where keys on x need 1665 --
// (646 with the row that takes the general pipeline's sums just above 1) --
// few enough rows to give EVERY LANE ITS OWN COPY of every entry (a row = 32 lanes x 4 bytes): a warp-wide
// look-up is one conflict-free wavefront whatever the data.  RN(x + c) is monotone in x, so a row is an interval
// of x; it holds at most one threshold and spans less than 2^24 float patterns (checked when the table is built),
// and the entry is biased so that
//     t = entry[row] + bits(x);   code = t >> 24
// (x clamped from below to 2^-13; exact zero has row 0 to itself -- kRowEncC lies half a row below 1/32 -- whose
// entry is made for the bit pattern the kernel presents for a zero of level +1, the only level encoded unclamped).
constexpr uint32_t kRowEncCBits    = 0x3CFF8000u;  // 1/32 - 2^-14
constexpr uint32_t kRowEncFirstKey = kRowEncCBits >> 16;
constexpr uint32_t kRowEncTopBits  = 0x3F810000u;  // 1 + 2^-7: the general pipeline's weighted sums may exceed 1 by a few ulp
constexpr uint32_t kRowEncRows     = 646;          // keys 0x3CFF (x = 0) .. 0x3F83 (x = 1), 0x3F84 (x up to kRowEncTopBits)
constexpr int      kFastDecScaleExp = 100;         // the tuned fast kernel carries 2^-100 * 4^K * x
constexpr uint32_t kRowEncZeroBits = uint32_t(kFastDecScaleExp - 2) << 23;  // what bits(0) + kAdd<1> presents

struct alignas(16) DeviceTables
{
  float    decode[256];                // linearFromSrgb(c), shaders/srgb.h:18-28 (pinned bits)
  uint32_t encode[kEncEntriesPadded];  // bucket table described above
  uint32_t encodeFast[kFastEncEntriesPadded];  // the same, keyed on bits >> kFastEncShift
  uint32_t encodeRows[(kRowEncRows + 3u) & ~3u];  // row table described above
};

// Programmatic dependent launch (sm_90+): every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, sets up its shared-memory tables (which depend on
// nothing a previous kernel wrote) and only then waits for the previous kernel of the stream to complete
// and flush.  Launch latency and table set-up of kernel n+1 thereby overlap the tail of kernel n -- the
// reference pays a full pipeline barrier + dispatch at that point (dispatch.hpp:180-186).
// No global memory written by an earlier kernel may be touched before gridDependencyWait().
__device__ __forceinline__ void gridDependencyWait()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void gridLaunchDependents()
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Copies n4 uint4s global -> shared with every load of a thread in flight before its first
// store: the per-CTA table set-up is pure latency, and it is paid by every launch.
template <int kThreads>
__device__ __forceinline__ void copyTableWide(uint4* dst, const uint4* src, uint32_t n4)
{
  constexpr int kMaxPerThread = 4;
  uint4         v[kMaxPerThread];
#pragma unroll
  for(int k = 0; k < kMaxPerThread; ++k)
  {
    const uint32_t i = threadIdx.x + uint32_t(k) * kThreads;
    if(i < n4)
      v[k] = __ldg(src + i);
  }
#pragma unroll
  for(int k = 0; k < kMaxPerThread; ++k)
  {
    const uint32_t i = threadIdx.x + uint32_t(k) * kThreads;
    if(i < n4)
      dst[i] = v[k];
  }
}

__device__ __forceinline__ float4 f4add(float4 a, float4 b)
{
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 f4scale(float s, float4 a)
{
  return make_float4(__fmul_rn(s, a.x), __fmul_rn(s, a.y), __fmul_rn(s, a.z), __fmul_rn(s, a.w));
}

// Reductions shared by both shipped instances (srgba8_mipmap_preamble.glsl:24-25,:35,:37-38).
struct LinearReduce
{
  using Value = float4;
  // a0*v0 + a1*v1 + a2*v2 contracted the way a GPU compiler contracts the GLSL expression
  // (srgba8_mipmap_preamble.glsl:24-25 has no `precise`): mul, fma, fma, left to right.
  // This is the only contraction of the numerics contract, and it is explicit.
  __device__ __forceinline__ static float reduce1(float a0, float v0, float a1, float v1, float a2, float v2)
  {
    return __fmaf_rn(a2, v2, __fmaf_rn(a1, v1, __fmul_rn(a0, v0)));
  }
  __device__ __forceinline__ static Value reduce(float a0, Value v0, float a1, Value v1, float a2, Value v2)
  {
    return make_float4(reduce1(a0, v0.x, a1, v1.x, a2, v2.x), reduce1(a0, v0.y, a1, v1.y, a2, v2.y),
                       reduce1(a0, v0.z, a1, v1.z, a2, v2.z), reduce1(a0, v0.w, a1, v1.w, a2, v2.w));
  }
  __device__ __forceinline__ static Value reduce2(Value v0, Value v1) { return f4scale(0.5f, f4add(v0, v1)); }
  // NVPRO_PYRAMID_SHARED_STORE + NVPRO_PYRAMID_SHARED_LOAD: what a value becomes on its way through the
  // reference's shared memory (sharedTile_, glsl:253,395,487-512; sharedLevel_, glsl:555,664,717).  The
  // default shared type is the value type itself (glsl:204-206): identity.  Our kernels keep such values in
  // registers or float4 shared memory and apply this function where the reference stores them.
  __device__ __forceinline__ static Value sharedRound(Value v) { return v; }
  // 0.25 * ((v00 + v01) + (v10 + v11)): the caller chooses which neighbours share a
  // bracket (pairing order differs per site in the reference, SURVEY.md section 8a note 1).
  __device__ __forceinline__ static Value reduce4(Value v00, Value v01, Value v10, Value v11)
  {
    return f4scale(0.25f, f4add(f4add(v00, v01), f4add(v10, v11)));
  }
};

// ---------------------------------------------------------------------------
// sRGBA8: srgba8_mipmap_preamble.glsl.  kRep = 32: the decode table is replicated per lane
// (entry (code, lane) at [code * 32 + lane]) so a warp-wide lookup with arbitrary codes never
// has a bank conflict -- for kernels that stream large levels.  kRep = 1: a single 1 KB copy,
// cheap to set up -- for the small tail levels.
template <int kRep>
struct Srgba8T : LinearReduce
{
  static_assert(kRep == 1 || kRep == 32, "decode table replication");
  static constexpr int kTexelBytes = 4;

  struct alignas(16) Shared
  {
    float                decode[256 * kRep];
    alignas(16) uint32_t encode[kEncEntriesPadded];
  };

  // Called by all 256 threads of a CTA.
  __device__ static void sharedInit(Shared& s, const DeviceTables* t)
  {
    static_assert(kEncEntriesPadded / 4 <= 4 * 256, "copyTableWide capacity");
    if(kRep == 1)
      s.decode[threadIdx.x & 255u] = __ldg(&t->decode[threadIdx.x & 255u]);
    else
    {
      // thread -> one code, its 32 lane copies as 8 x float4
      const float  v  = __ldg(&t->decode[threadIdx.x & 255u]);
      const float4 v4 = make_float4(v, v, v, v);
      float4*      d  = reinterpret_cast<float4*>(&s.decode[(threadIdx.x & 255u) * kRep]);
#pragma unroll
      for(int k = 0; k < kRep / 4; ++k)
        d[k] = v4;
    }
    copyTableWide<256>(reinterpret_cast<uint4*>(s.encode), reinterpret_cast<const uint4*>(t->encode),
                       kEncEntriesPadded / 4);
  }

  // texelFetch on the sRGB view: RGB through the decode table, alpha = a * (1/255)
  // (shaders/srgb.h:60, mipmap_storage.hpp:400).
  __device__ __forceinline__ static Value decodeWord(const Shared& s, uint32_t w)
  {
    const uint32_t lane = kRep == 32 ? (threadIdx.x & 31u) : 0u;
    Value          v;
    v.x = s.decode[((w & 0xFFu) * kRep) | lane];
    v.y = s.decode[(((w >> 8) & 0xFFu) * kRep) | lane];
    v.z = s.decode[(((w >> 16) & 0xFFu) * kRep) | lane];
    v.w = __fmul_rn(float(w >> 24), 1.0f / 255.0f);
    return v;
  }

  template <bool kClampHigh>
  __device__ __forceinline__ static uint32_t encodeChannel(const Shared& s, float x)
  {
    uint32_t b = __float_as_uint(x);
    b          = max(b, kEncMinBits);
    if(kClampHigh)
      b = min(b, kEncMaxBits);
    return s.encode[(b >> kEncShift) - kEncMinKey] + b;  // code in bits 16..23
  }

  // srgbFromLinearVec, srgba8_mipmap_preamble.glsl:122-128.  kClampHigh = false is
  // valid when every channel is known to be <= 1 (2x2 box reductions of values <= 1).
  template <bool kClampHigh>
  __device__ __forceinline__ static uint32_t encodeWord(const Shared& s, Value v)
  {
    const uint32_t r = encodeChannel<kClampHigh>(s, v.x);
    const uint32_t g = encodeChannel<kClampHigh>(s, v.y);
    const uint32_t b = encodeChannel<kClampHigh>(s, v.z);
    // uint(a * 255 + 0.5), clamped to 255.
    const float    af = __fadd_rn(__fmul_rn(v.w, 255.0f), 0.5f);
    uint32_t       a  = __float2uint_rz(af);
    a                 = min(a, 255u);
    const uint32_t rg = __byte_perm(r, g, 0x0062);  // byte0 = r.byte2, byte1 = g.byte2
    const uint32_t ba = __byte_perm(b, a, 0x0042);  // byte0 = b.byte2, byte1 = a.byte0
    return __byte_perm(rg, ba, 0x5410);
  }

  __device__ __forceinline__ static Value load(const Shared& s, const void* p)
  {
    return decodeWord(s, *reinterpret_cast<const uint32_t*>(p));
  }
  __device__ __forceinline__ static void load4(const Shared& s, const void* p, Value out[4])
  {
    const uint4 w = *reinterpret_cast<const uint4*>(p);
    out[0]        = decodeWord(s, w.x);
    out[1]        = decodeWord(s, w.y);
    out[2]        = decodeWord(s, w.z);
    out[3]        = decodeWord(s, w.w);
  }
  template <bool kClampHigh = true>
  __device__ __forceinline__ static void store(const Shared& s, void* p, Value v)
  {
    *reinterpret_cast<uint32_t*>(p) = encodeWord<kClampHigh>(s, v);
  }
  template <bool kClampHigh = true>
  __device__ __forceinline__ static void store2(const Shared& s, void* p, Value v0, Value v1)
  {
    *reinterpret_cast<uint2*>(p) = make_uint2(encodeWord<kClampHigh>(s, v0), encodeWord<kClampHigh>(s, v1));
  }
};

using Srgba8     = Srgba8T<32>;
using Srgba8Lite = Srgba8T<1>;

// F16_SHARED variant of the sRGBA8 instance (srgba8_mipmap_preamble.glsl:103-108): the shared type is
// f16vec4, i.e. values that pass through shared memory are rounded to IEEE binary16 (round to nearest even,
// the conversion of f16vec4(vec4)) and widened again.  Non-default in the reference; runs on the
// functor-template kernels only.
struct Srgba8F16Shared : Srgba8T<1>
{
  __device__ __forceinline__ static Value sharedRound(Value v)
  {
    return make_float4(__half2float(__float2half_rn(v.x)), __half2float(__float2half_rn(v.y)),
                       __half2float(__float2half_rn(v.z)), __half2float(__float2half_rn(v.w)));
  }
};

// SRGB_SHARED variant of the sRGBA8 instance (srgba8_mipmap_preamble.glsl:60-101, the demo's "srgbShared"
// alternative): the shared type is a packed 8-bit sRGB texel -- SHARED_STORE = packUnorm4x8 of
// srgbComponentFromLinear, SHARED_LOAD = linearFromSrgbComponent of unpackUnorm4x8.  Numerics as pinned by
// tools/gen_srgb_tables.c (and checked by the test suite against the executed shaders): the pack is monotone, so it equals the number of pinned
// SHARED thresholds <= x (binary search in constant memory); the unpack is a 256-entry table; alpha is
// round-half-even(clamp(a, 0, 1) * 255) / 255 (IEEE division).  Non-default, lossy and slow in the reference too;
// runs on the functor-template kernels only.
namespace devtables {
#define NVPYR_TABLES_SHARED_ONLY
#define NVPYR_SHARED_TABLE_DECL static __constant__ unsigned int
#include "srgb_tables.inc"
#undef NVPYR_SHARED_TABLE_DECL
#undef NVPYR_TABLES_SHARED_ONLY
}  // namespace devtables

struct Srgba8SrgbShared : Srgba8T<1>
{
  __device__ __forceinline__ static float roundTrip(float x)
  {
    uint32_t lo = 0u, hi = 255u;  // the code is in [lo, hi]
#pragma unroll
    for(int i = 0; i < 8; ++i)
    {
      const uint32_t mid = (lo + hi + 1u) >> 1;
      if(mid > lo && x >= __uint_as_float(devtables::NVPYR_SRGB_SHARED_ENCODE_THRESHOLD_BITS[mid - 1u]))
        lo = mid;
      else
        hi = mid > lo ? mid - 1u : hi;
    }
    return __uint_as_float(devtables::NVPYR_SRGB_SHARED_DECODE_BITS[lo]);
  }
  __device__ __forceinline__ static Value sharedRound(Value v)
  {
    const float a = fminf(fmaxf(v.w, 0.0f), 1.0f);
    return make_float4(roundTrip(v.x), roundTrip(v.y), roundTrip(v.z), __fdiv_rn(rintf(__fmul_rn(a, 255.0f)), 255.0f));
  }
};

// ---------------------------------------------------------------------------
// RGBA32F: the template of nvpro_pyramid.glsl:27-49 instantiated with identity
// load/store (the reference ships no such shader).
struct Rgba32f : LinearReduce
{
  static constexpr int kTexelBytes = 16;
  struct Shared
  {
    int unused;
  };
  __device__ static void sharedInit(Shared&, const DeviceTables*) {}
  __device__ __forceinline__ static Value load(const Shared&, const void* p)
  {
    return *reinterpret_cast<const float4*>(p);
  }
  __device__ __forceinline__ static void load4(const Shared&, const void* p, Value out[4])
  {
    const float4* q = reinterpret_cast<const float4*>(p);
    out[0] = q[0], out[1] = q[1], out[2] = q[2], out[3] = q[3];
  }
  template <bool kClampHigh = true>
  __device__ __forceinline__ static void store(const Shared&, void* p, Value v)
  {
    *reinterpret_cast<float4*>(p) = v;
  }
  template <bool kClampHigh = true>
  __device__ __forceinline__ static void store2(const Shared&, void* p, Value v0, Value v1)
  {
    float4* q = reinterpret_cast<float4*>(p);
    q[0] = v0, q[1] = v1;
  }
};

}  // namespace nvpyr
