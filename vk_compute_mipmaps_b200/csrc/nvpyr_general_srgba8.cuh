// nvpyr_general_srgba8.cuh -- tuned general (NPOT) pipeline for sRGBA8: 1 or 2 levels from an
// input level of any size >= 2x2, the reference's energy-conserving separable 1/2/3-tap kernel
// (nvpro_pyramid/nvpro_pyramid.glsl:575-656) with its float32 carry from level +1 to level +2.
//
// Same float32 expression tree per output texel as generalKernel<Srgba8> (nvpyr_kernels.cuh) --
// vertical reduction of each source column, then horizontal reduction, weights
// (n - i, n, 1 - w0 - w1) / (2n + 1) -- hence the same bits, but organised as a stream:
//
//   * one WARP owns a strip of 31 columns of level +1 (lane l <-> column c0 + l, strips advance by
//     30: the last column is the halo the next strip recomputes, like the reference's overlapping
//     work groups) and walks DOWN a segment of rows.  Each lane loads its own two source columns
//     (a warp reads 256 contiguous bytes per source row), decodes them once (PRMT + conflict-free
//     LDS, as in the fast kernel), keeps source row 2y+2 in registers as row 2(y+1) of the next
//     output row, reduces its two columns vertically and fetches the third column's vertical sum
//     from lane l+1 with shuffles.  The reference fetches and decodes up to 9 texels per output
//     texel; here it is 2 (+ 2 shuffled floats4).
//   * level +2 is accumulated on the fly: each lane keeps the last two level +1 values of its
//     column (float32, never re-quantised), reduces vertically every second row, and even lanes
//     finish horizontally with two shuffles.  No shared-memory tile, no barrier.
//
// 1/(2n+1) is one IEEE division per kernel; per-row weights cost a subtract and two multiplies.
#pragma once
#include <type_traits>

#include "nvpyr_fast_srgba8.cuh"

namespace nvpyr {

// Shared-memory layout chosen so that table look-ups need NO address arithmetic beyond the one instruction
// that extracts the index (the kernel is instruction-issue bound; a base-address add per look-up was 8 % of
// all instructions).  Addresses below are in the CTA's shared window, whose dynamic part starts at
// kGenWindowBase (1 KB is reserved by the system on sm_100; checked at run time, the kernel traps otherwise):
//   * decode table at window address 0x10000, [code][64 floats] (floats 0..31 = one copy per lane):
//     address = 0x10000 | code << 8 | lane << 2 is produced by ONE PRMT from the packed texel and
//     (0x10000 | lane << 2);
//   * encode bucket table placed so that the entry of key k sits at window address (stride * k) mod 2^16: the
//     masked, shifted float bits ARE the address (no wrap inside the key range; see NVPYR_GEN_ENC_WAYS below).
// 127 KB per CTA, one CTA per SM; ~100 KB stay free for the next kernel's CTAs (programmatic dependent launch).
constexpr uint32_t kGenWindowBase = 0x400u;
constexpr uint32_t kGenDecodeAddr = 0x10000u;
// NVPYR_GEN_ENC_WAYS = 1 (default): 8 mantissa bits, one copy, the entry of key k at window address (4 k) mod 2^16.
// NVPYR_GEN_ENC_WAYS = 8: the bucket table keyed on 7 mantissa bits (DeviceTables::encodeFast), every entry stored
// 8 times (32 bytes), lane l reads copy l & 7 -- lanes with different l & 7 never collide on a bank (un-replicated,
// 35 % of the shared-load wavefronts of the four-column kernel are bank conflicts on noisy input).  The entry of
// key k sits at window address (32 k) mod 2^16 (keys 0x3900..0x3F80 -> 0x2000..0xF000, 52 KB, below the decode
// table).  Bit-exact, but no faster (4095^2 52.4 -> 52.0 us, 4094^2 50.9 -> 51.7 us, the rest within noise): these
// kernels are bound by load latency, not by the LSU pipe -- so the cheaper set-up stays the default.
#ifndef NVPYR_GEN_ENC_WAYS
#define NVPYR_GEN_ENC_WAYS 1
#endif
constexpr uint32_t kGenEncWays    = NVPYR_GEN_ENC_WAYS;
static_assert(kGenEncWays == 1 || kGenEncWays == 8, "one copy (8-bit buckets) or eight (7-bit buckets)");
constexpr uint32_t kGenEncShift   = kGenEncWays == 1 ? kEncShift : kFastEncShift;
constexpr uint32_t kGenEncMinKey  = kEncMinBits >> kGenEncShift, kGenEncMaxKey = kEncMaxBits >> kGenEncShift;
constexpr uint32_t kGenEncEntries = kGenEncWays == 1 ? kEncEntriesPadded : kFastEncEntriesPadded;
constexpr uint32_t kGenEncStride  = 4u * kGenEncWays;  // bytes per entry
constexpr uint32_t kGenEncodeMask = 0x10000u - kGenEncStride;
constexpr uint32_t kGenEncodeAddr = (kGenEncMinKey * kGenEncStride) & kGenEncodeMask;  // window address of the first entry
static_assert(((kGenEncMaxKey * kGenEncStride) & kGenEncodeMask) == kGenEncodeAddr + (kGenEncMaxKey - kGenEncMinKey) * kGenEncStride,
              "the key range must not wrap inside the 16-bit window");
// NVPYR_GEN_ENC_ROWS = 1 (default, round 2): the strip kernels use the fast kernel's row table (nvpyr_functors.cuh):
// ONE table of kRowEncRows rows x 256 bytes at window address 0x10000 -- bytes 0..127 of row r = the 32 lane copies of
// linearFromSrgb(r) (r < 256), bytes 128..255 = the 32 lane copies of encode entry r -- so that the encode look-up is
// as conflict-free as the decode (ncu, round 2: 31 % of the four-column kernel's shared-load wavefronts were bank
// conflicts of the single-copy encode table).  Look-up: FADD (z = x + c), PRMT (key(z) << 8 | lane << 2), LDS with an
// immediate that folds the table's window address, IADD.  The staging ring of the four-column kernel moves below the
// table (3 stages per warp: 48 KB between the window base and 0x10000).
#ifndef NVPYR_GEN_ENC_ROWS
#define NVPYR_GEN_ENC_ROWS 1
#endif
constexpr bool     kGenRows       = NVPYR_GEN_ENC_ROWS != 0;
constexpr uint32_t kGenRowsEnd    = kGenDecodeAddr + kRowEncRows * 256u;  // window address behind the row table
constexpr int32_t  kGenRowEncImm  = int32_t(kGenDecodeAddr + 128u) - int32_t(kRowEncFirstKey << 8);  // LDS immediate of the encode look-up
constexpr uint32_t kGenSmemBytes  = (kGenRows ? kGenRowsEnd : kGenDecodeAddr + 256u * 256u) - kGenWindowBase;
static_assert(kGenRowsEnd <= kGenWindowBase + 227u * 1024u, "row table must end inside the largest dynamic window");
static_assert(kGenEncodeAddr >= kGenWindowBase && kGenEncodeAddr + kGenEncEntries * kGenEncStride <= kGenDecodeAddr,
              "encode table must fit below the decode table");
static_assert(kGenEncodeAddr % 16u == 0, "encode table is copied as uint4");

struct GenStripParams
{
  LevelView           lv[3];
  uint32_t            stripsX;   // strips of 30 (+1 halo) level +1 columns
  uint32_t            segsY;     // row segments
  uint32_t            segRows;   // rows per segment: of level +2 (two levels) or of level +1 (one level)
  const DeviceTables* tables;
};

#ifndef NVPYR_GEN_WARPS
#define NVPYR_GEN_WARPS 24
#endif
constexpr int kGenWarps = NVPYR_GEN_WARPS;  // one CTA of 24 warps per SM (80 registers per thread)

template <int kThreads = kGenWarps * 32>
__device__ __forceinline__ void genSrgba8Init(unsigned char* smemRaw, const DeviceTables* t)
{
  if(kGenRows)
  {
    // thread -> (row, 16-byte column): eight consecutive threads write the 128 contiguous bytes of one half row
    // (every load of a thread in flight before its first store, as in srgba8FastInit: the set-up is pure latency)
    uint4*             tab    = reinterpret_cast<uint4*>(smemRaw + (kGenDecodeAddr - kGenWindowBase));  // 16 uint4 per row
    constexpr uint32_t kItems = kRowEncRows * 8u, kPerThread = (kItems + kThreads - 1) / kThreads;
    uint32_t           e[kPerThread], d[kPerThread];
#pragma unroll
    for(uint32_t k = 0; k < kPerThread; ++k)
    {
      const uint32_t i = threadIdx.x + k * kThreads, row = i >> 3;
      if(i < kItems)
      {
        e[k] = __ldg(&t->encodeRows[row]);
        if(row < 256u)
          d[k] = __float_as_uint(__ldg(&t->decode[row]));
      }
    }
#pragma unroll
    for(uint32_t k = 0; k < kPerThread; ++k)
    {
      const uint32_t i = threadIdx.x + k * kThreads, row = i >> 3, col = i & 7u;
      if(i < kItems)
      {
        tab[row * 16u + 8u + col] = make_uint4(e[k], e[k], e[k], e[k]);
        if(row < 256u)
          tab[row * 16u + col] = make_uint4(d[k], d[k], d[k], d[k]);
      }
    }
    return;
  }
  float*    decode = reinterpret_cast<float*>(smemRaw + (kGenDecodeAddr - kGenWindowBase));
  uint32_t* encode = reinterpret_cast<uint32_t*>(smemRaw + (kGenEncodeAddr - kGenWindowBase));
  for(uint32_t i = threadIdx.x; i < 512u; i += kThreads)
  {
    const uint32_t code = i >> 1, half = i & 1u;
    const float    v    = __ldg(&t->decode[code]);
    float4*        d    = reinterpret_cast<float4*>(&decode[code * 64u + half * 16u]);
    const float4   v4   = make_float4(v, v, v, v);
    d[0] = v4, d[1] = v4, d[2] = v4, d[3] = v4;
  }
  if(kGenEncWays == 1)
    copyTableWide<kThreads>(reinterpret_cast<uint4*>(encode), reinterpret_cast<const uint4*>(t->encode),
                            kEncEntriesPadded / 4);
  else
  {
    // thread -> (entry, half): 4 of its 8 copies as one 16-byte store
    uint4* e4 = reinterpret_cast<uint4*>(encode);
    for(uint32_t i = threadIdx.x; i < kGenEncEntries * 2u; i += kThreads)
    {
      const uint32_t v = __ldg(&t->encodeFast[i >> 1]);
      e4[i]            = make_uint4(v, v, v, v);
    }
  }
}

// linearFromSrgb of byte kByte (0..2) of a packed texel: PRMT builds the absolute shared address
// (laneAddr = kGenDecodeAddr | lane << 2), LDS reads it.
template <int kByte>
__device__ __forceinline__ float genDec8(uint32_t w, uint32_t laneAddr)
{
  const uint32_t addr = __byte_perm(w, laneAddr, 0x7604u | (uint32_t(kByte) << 4));
  float          v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ V4 genDecodeTexel(uint32_t laneAddr, uint32_t w)
{
  V4 r;
  r.rg = pack2(genDec8<0>(w, laneAddr), genDec8<1>(w, laneAddr));
  r.ba = pack2(genDec8<2>(w, laneAddr), decAlpha(w));
  return r;
}

// Packed float32 pairs (FMUL2 / FADD2, two IEEE roundings per issue slot): see nvpyr_fast_srgba8.cuh.
__device__ __forceinline__ F2 mul2(F2 a, F2 b)
{
  F2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c)
{
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
// a0*v0 + a1*v1 + a2*v2 as mul, fma, fma (the contract's one explicit contraction, see
// LinearReduce::reduce); weights come as (w, w) pairs.  FMUL2/FFMA2: two lanes per issue slot.
__device__ __forceinline__ V4 genReduce3(F2 a0, V4 v0, F2 a1, V4 v1, F2 a2, V4 v2)
{
  V4 r;
  r.rg = fma2(a2, v2.rg, fma2(a1, v1.rg, mul2(a0, v0.rg)));
  r.ba = fma2(a2, v2.ba, fma2(a1, v1.ba, mul2(a0, v0.ba)));
  return r;
}
// 0.5 * (v0 + v1) (srgba8_mipmap_preamble.glsl:35)
__device__ __forceinline__ V4 genReduce2(V4 v0, V4 v1)
{
  const F2 half = pack2(0.5f, 0.5f);
  V4       r;
  r.rg = mul2(add2(v0.rg, v1.rg), half);
  r.ba = mul2(add2(v0.ba, v1.ba), half);
  return r;
}

// srgbFromLinear with both clamps (weighted sums may exceed 1 by an ulp); code in bits 16..23.
// Row-table variant: code in bits 24..31.  laneAddr = kGenDecodeAddr | lane << 2 (byte 0 = lane << 2, byte 3 = 0).
__device__ __forceinline__ uint32_t genEncChannelRows(float x, uint32_t laneAddr)
{
  // Weighted sums are >= 0 and stay below 1 + 2^-7 (the table's last row): only the lower clamp is needed.
  const uint32_t b    = max(__float_as_uint(x), kEncMinBits);
  const float    z    = __fadd_rn(__uint_as_float(b), __uint_as_float(kRowEncCBits));  // RN(x + c)
  const uint32_t addr = __byte_perm(__float_as_uint(z), laneAddr, 0x7324);           // key(z) << 8 | lane << 2
  uint32_t       e;
  asm("ld.shared.u32 %0, [%1+%2];" : "=r"(e) : "r"(addr), "n"(kGenRowEncImm));
  return e + b;
}
__device__ __forceinline__ uint32_t genEncChannel(float x)
{
  // Weighted sums stay below 1 + 2^-8, i.e. inside the table's last bucket (key of 1.0f): only the lower
  // clamp is needed.
  const uint32_t b    = max(__float_as_uint(x), kEncMinBits);
  uint32_t addr = (b >> (kGenEncShift - (kGenEncWays == 1 ? 2 : 5))) & kGenEncodeMask;  // (stride * key) mod 2^16 = the entry's shared address
  if(kGenEncWays != 1)
    addr |= (threadIdx.x & 7u) << 2;  // this lane's copy
  uint32_t       e;
  asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(addr));
  return e + b;
}
__device__ __forceinline__ uint32_t genEncWord(float4 v)
{
  // uint(a * 255 + 0.5): a <= 1 + 2 ulp, so the truncation never exceeds 255
  const uint32_t a = __float_as_uint(__fadd_rz(__fadd_rn(__fmul_rn(v.w, 255.0f), 0.5f), 8388608.0f));
  if(kGenRows)
  {
    const uint32_t laneAddr = kGenDecodeAddr | ((threadIdx.x & 31u) << 2);
    const uint32_t r = genEncChannelRows(v.x, laneAddr), g = genEncChannelRows(v.y, laneAddr), b = genEncChannelRows(v.z, laneAddr);
    return __byte_perm(__byte_perm(r, g, 0x0073), __byte_perm(b, a, 0x0043), 0x5410);
  }
  const uint32_t r = genEncChannel(v.x), g = genEncChannel(v.y), b = genEncChannel(v.z);
  return __byte_perm(__byte_perm(r, g, 0x0062), __byte_perm(b, a, 0x0042), 0x5410);
}

// Weights of destination index i of n (glsl:582-586): w0 = rcp*(n-i), w1 = rcp*n, w2 = 1-w0-w1.
struct Taps
{
  F2 w0, w1, w2;  // each weight duplicated into both halves of a pair
};
__device__ __forceinline__ Taps genTaps(float rcp, float fn, uint32_t i)
{
  const float w0 = __fmul_rn(rcp, __fsub_rn(fn, float(i)));
  const float w1 = __fmul_rn(rcp, fn);
  const float w2 = __fsub_rn(__fsub_rn(1.0f, w0), w1);
  Taps        t;
  t.w0 = pack2(w0, w0), t.w1 = pack2(w1, w1), t.w2 = pack2(w2, w2);
  return t;
}
__device__ __forceinline__ float genRcp(uint32_t n)
{
  const float fn = float(n);
  return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(2.0f, fn), 1.0f));
}
__device__ __forceinline__ V4 shflDown(V4 v, int d)
{
  V4 r;
  r.rg.v = __shfl_down_sync(0xffffffffu, v.rg.v, d);
  r.ba.v = __shfl_down_sync(0xffffffffu, v.ba.v, d);
  return r;
}

// Reports where the dynamic shared-memory window of a kernel without static shared memory starts (run once per
// device by the host; the strip kernels below are only used where the answer is kGenWindowBase).
__global__ void genWindowProbeKernel(uint32_t* out)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  if(threadIdx.x == 0)
    *out = uint32_t(__cvta_generic_to_shared(smemRaw));
}

// The strip kernel is written once and instantiated with a texel codec: how a raw texel is fetched, turned
// into a float vector and stored again (the NVPRO_PYRAMID_LOAD / _STORE macros of this kernel).
struct GenCodecSrgba8
{
  using Raw                              = uint32_t;
  static constexpr uint32_t kTexelBytes  = 4;
  static constexpr bool     kTables      = true;  // sRGB tables in shared memory (absolute addresses above)
  static constexpr int      kWarps       = kGenWarps;
  static constexpr size_t   kSmemBytes   = kGenSmemBytes;
  __device__ __forceinline__ static Raw  zero() { return 0u; }
  __device__ __forceinline__ static Raw  load(const unsigned char* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
  __device__ __forceinline__ static V4   decode(uint32_t laneAddr, Raw w) { return genDecodeTexel(laneAddr, w); }
  __device__ __forceinline__ static void store(unsigned char* p, V4 v) { *reinterpret_cast<uint32_t*>(p) = genEncWord(toFloat4(v)); }
};
// rgba32f: identity load / store (nvpro_pyramid.glsl:27-49 instantiated with trivial macros), 16-byte texels.
struct GenCodecRgba32f
{
  using Raw                              = float4;
  static constexpr uint32_t kTexelBytes  = 16;
  static constexpr bool     kTables      = false;
  static constexpr int      kWarps       = 16;  // 512 threads, up to 128 registers: the raw prefetch slots are 4x larger
  static constexpr size_t   kSmemBytes   = 0;
  __device__ __forceinline__ static Raw  zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ static Raw  load(const unsigned char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ static V4   decode(uint32_t, Raw w) { return toV4(w); }
  __device__ __forceinline__ static void store(unsigned char* p, V4 v) { *reinterpret_cast<float4*>(p) = toFloat4(v); }
};

// kLevels: 1 or 2.  kX3 / kY3: the first level uses 3 taps (odd source size) along x / y; otherwise 2.
template <class C, int kLevels, bool kX3, bool kY3>
__global__ void __launch_bounds__(C::kWarps * 32, 1) generalStripKernel(const GenStripParams p)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  using Raw                 = typename C::Raw;
  constexpr uint32_t TB     = C::kTexelBytes;
  constexpr uint32_t kWarps = uint32_t(C::kWarps);
  if(C::kTables)
  {
    if(uint32_t(__cvta_generic_to_shared(smemRaw)) != kGenWindowBase)
      __trap();  // the absolute table addresses above assume this window layout: fail loudly, never silently
    genSrgba8Init(smemRaw, p.tables);
    __syncthreads();
  }
  const uint32_t  lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, laneAddr = kGenDecodeAddr | (lane * 4u);
  const LevelView L0 = p.lv[0], L1 = p.lv[1], L2 = p.lv[2];
  const float     fH1 = float(L1.h), fW1 = float(L1.w);
  const float     rcpY1 = kY3 ? genRcp(L1.h) : 0.f, rcpX1 = kX3 ? genRcp(L1.w) : 0.f;
  // second level (uniform run-time branches)
  const bool  x3b = kLevels == 2 && (L1.w & 1u), y3b = kLevels == 2 && (L1.h & 1u);
  const float fH2 = kLevels == 2 ? float(L2.h) : 0.f, fW2 = kLevels == 2 ? float(L2.w) : 0.f;
  const float rcpY2 = y3b ? genRcp(L2.h) : 0.f, rcpX2 = x3b ? genRcp(L2.w) : 0.f;

  const uint32_t numTasks = p.stripsX * p.segsY;
  gridDependencyWait();    // the previous kernel's levels are complete and visible (the constants above need none of them)
  gridLaunchDependents();  // the next kernel may start its own set-up as SMs become free
  for(uint32_t task = blockIdx.x + gridDim.x * warp; task < numTasks; task += gridDim.x * kWarps)
  {
    const uint32_t sx = task % p.stripsX, sy = task / p.stripsX;
    const uint32_t x1 = sx * 30u + lane;  // this lane's column of level +1
    const uint32_t c0 = 2u * x1;          // its first source column
    const bool     srcA = c0 < L0.w, srcB = c0 + 1u < L0.w;
    const bool     out1 = x1 < L1.w && lane < 31u;
    const V4       zero = toV4(make_float4(0.f, 0.f, 0.f, 0.f));
    Taps           tx1{zero.rg, zero.rg, zero.rg};
    if(kX3)
      tx1 = genTaps(rcpX1, fW1, x1);
    // level +2: even lanes own column x2 = 15 sx + lane / 2
    const uint32_t x2   = sx * 15u + (lane >> 1);
    const bool     out2 = kLevels == 2 && !(lane & 1u) && lane < 30u && x2 < L2.w;
    Taps           tx2{zero.rg, zero.rg, zero.rg};
    if(kLevels == 2 && x3b)
      tx2 = genTaps(rcpX2, fW2, x2);

    // rows of level +1 handled by this task: [ya, yb]
    uint32_t ya, yb, r2a = 0;
    if(kLevels == 2)
    {
      r2a                = sy * p.segRows;
      const uint32_t r2b = min(r2a + p.segRows, L2.h);
      ya                 = 2u * r2a;
      yb                 = min(y3b ? 2u * r2b : 2u * r2b - 1u, L1.h - 1u);
    }
    else
    {
      ya = sy * p.segRows;
      yb = min(ya + p.segRows, L1.h) - 1u;
    }

    const unsigned char* src = L0.ptr + size_t(2u * ya) * L0.pitch + size_t(c0) * TB;  // source row 2*ya
    auto                 load2 = [&](const unsigned char* row, Raw& a, Raw& b) {
      a = srcA ? C::load(row) : C::zero();
      b = srcB ? C::load(row + TB) : C::zero();
    };

    V4 carryA = zero, carryB = zero;  // decoded source row 2y (3-tap only)
    if(kY3)
    {
      Raw a, b;
      load2(src, a, b);
      carryA = C::decode(laneAddr, a);
      carryB = C::decode(laneAddr, b);
    }
    // Raw words of the two new source rows of an output row.  Two output rows are in flight ahead of the
    // one being computed; the row loop is unrolled by two so that both slots, the vertical carry and the
    // level +2 history (q0, q1) are compile-time registers that never need to be moved.
    struct Slot
    {
      Raw a0, b0, a1, b1;
    };
    const unsigned char* nextRows = kY3 ? src + L0.pitch : src;  // new source rows of the next row to prefetch
    const size_t         rowStep  = 2u * size_t(L0.pitch);
    auto                 loadRow  = [&](bool valid, Slot& r) {
      if(valid)
      {
        load2(nextRows, r.a0, r.b0);
        load2(nextRows + L0.pitch, r.a1, r.b1);
      }
      nextRows += rowStep;
    };
    Slot r0{C::zero(), C::zero(), C::zero(), C::zero()}, r1{C::zero(), C::zero(), C::zero(), C::zero()};
    loadRow(true, r0);
    loadRow(ya + 1u <= yb, r1);
    V4 q0 = zero, q1 = zero;  // last level +1 values of this column

    unsigned char* d1 = L1.ptr + size_t(ya) * L1.pitch + size_t(x1) * TB;
    unsigned char* d2 = kLevels == 2 ? L2.ptr + size_t(r2a) * L2.pitch + size_t(x2) * TB : nullptr;
    uint32_t       y2 = r2a;  // next row of level +2 to emit

    // One output row of level +1 (and what it completes of level +2).  kOdd: parity of the row inside the
    // segment (segments of two-level steps start on even rows).
    auto row = [&](uint32_t y, Slot& slot, auto parity) {
      constexpr bool kOdd = decltype(parity)::value;
      const Raw      m0a = slot.a0, m0b = slot.b0, m1a = slot.a1, m1b = slot.b1;
      loadRow(y + 2u <= yb, slot);
      // ---- vertical reduction of this lane's two source columns ----
      const V4 vA0 = C::decode(laneAddr, m0a), vB0 = C::decode(laneAddr, m0b);
      const V4 vA1 = C::decode(laneAddr, m1a), vB1 = C::decode(laneAddr, m1b);
      V4       hA, hB;
      if(kY3)
      {
        const Taps ty = genTaps(rcpY1, fH1, y);
        hA            = genReduce3(ty.w0, carryA, ty.w1, vA0, ty.w2, vA1);
        hB            = genReduce3(ty.w0, carryB, ty.w1, vB0, ty.w2, vB1);
        carryA        = vA1;
        carryB        = vB1;
      }
      else
      {
        hA = genReduce2(vA0, vA1);
        hB = genReduce2(vB0, vB1);
      }
      // ---- horizontal reduction ----
      V4 o;
      if(kX3)
      {
        const V4 hC = shflDown(hA, 1);  // column 2 x1 + 2 = first column of lane + 1
        o           = genReduce3(tx1.w0, hA, tx1.w1, hB, tx1.w2, hC);
      }
      else
        o = genReduce2(hA, hB);
      if(out1)
        C::store(d1, o);
      d1 += L1.pitch;

      // ---- level +2, float32 carry ----
      if(kLevels == 2)
      {
        V4   g    = zero;
        bool emit = false;
        if(y3b)
        {
          // rows 2 y2, 2 y2 + 1, 2 y2 + 2: emit when the third arrives (even row, not the segment's first)
          if(!kOdd)
          {
            if(y != ya)
            {
              const Taps ty = genTaps(rcpY2, fH2, y2);
              g             = genReduce3(ty.w0, q0, ty.w1, q1, ty.w2, o);
              emit          = true;
            }
            q0 = o;
          }
          else
            q1 = o;
        }
        else
        {
          if(kOdd)
          {
            g    = genReduce2(q0, o);
            emit = true;
          }
          else
            q0 = o;
        }
        if(emit)  // warp-uniform
        {
          const V4 g1 = shflDown(g, 1);
          V4       o2;
          if(x3b)
          {
            const V4 g2 = shflDown(g, 2);
            o2          = genReduce3(tx2.w0, g, tx2.w1, g1, tx2.w2, g2);
          }
          else
            o2 = genReduce2(g, g1);
          if(out2)
            C::store(d2, o2);
          d2 += L2.pitch;
          ++y2;
        }
      }
    };
    uint32_t y = ya;
    for(; y < yb; y += 2u)
    {
      row(y, r0, std::false_type{});
      row(y + 1u, r1, std::true_type{});
    }
    if(y == yb)
      row(y, r0, std::false_type{});
  }
}

// ---------------------------------------------------------------------------------------------------
// generalStrip4Kernel: the sRGBA8 strip kernel with FOUR source columns per lane.
//
// The 2-column strip kernel above is instruction-issue bound: per source texel it spends ~47 warp
// instructions, more than half of them per-row overhead (row weights, pointer updates, control flow, the
// horizontal reduction and both encodes) that does not depend on how many columns a lane owns, and its
// level +2 encode runs on half the lanes.  Here a lane owns source columns c0..c0+3 = level +1 columns
// xa, xa+1 = ONE level +2 column; a strip is 63 columns of level +1 (the last one is the halo the next
// strip recomputes, strips advance by 62) and 31 columns of level +2.  Same float32 expression tree per
// output texel (vertical reduction of each source column, then horizontal; weights (n - i, n, 1 - w0 - w1)
// / (2n + 1); float32 carry to level +2), hence the same bits, with ~half the instructions per texel.
#ifndef NVPYR_GEN4_WARPS
#define NVPYR_GEN4_WARPS 16
#endif
constexpr int kGen4Warps = NVPYR_GEN4_WARPS;  // 16: 512 threads per CTA, one CTA per SM, up to 128 registers per thread

// kStaged: the source rows of a strip are staged in shared memory by asynchronous copies instead of per-lane loads
// into registers.  NPOT levels have row pitches that are not multiples of 16 bytes (4095 texels = 16380 bytes), so
// rows start at 0 / 12 / 8 / 4 bytes modulo 16: a lane cannot fetch its four columns with one 16-byte load.  The
// register path issues four 4-byte loads per row and lane, 16 bytes apart from lane to lane -- every instruction
// touches 16 sectors and every sector is fetched four times from L1 (ncu, round 1: 302 MB of L1 sector traffic for
// 67 MB of input) -- and keeps two output rows of raw texels in prefetch registers (121 registers per thread).
// The TMA unit cannot help: a tensor-map box must start at a 16-byte aligned global address (measured with
// tools/tma1d_probe.cu: a 1-D map accepts element 0, element 1 raises an illegal-instruction fault; 2-D maps need
// 16-byte strides).  cp.async (LDGSTS) can: it copies 4-byte words, so lane l moves WORD l + 32 j of the strip's
// 512-byte row segment -- consecutive lanes, consecutive words: 4-5 sectors per instruction, each fetched once --
// into a 16-byte aligned stage of the warp's ring, and every lane then reads its four columns of both rows back with
// two LDS.128.  No prefetch registers; kGenStages output rows are in flight ahead of the one being computed instead
// of two.  Copies beyond the end of a row are suppressed (src-size 0: zero fill, no global access).
#ifndef NVPYR_GEN_STAGES
#define NVPYR_GEN_STAGES (NVPYR_GEN_ENC_ROWS ? 3 : 4)
#endif
constexpr uint32_t kGenStages     = NVPYR_GEN_STAGES;
constexpr uint32_t kGenStageBytes = 1024;                     // two source rows x 128 texels
// window address of the first warp's ring: behind the decode table, or (row table) below it
constexpr uint32_t kGenRingAddr   = kGenRows ? kGenWindowBase : kGenDecodeAddr + 0x10000u;
constexpr uint32_t kGenStagedSmemBytes =
    kGenRows ? kGenSmemBytes : kGenRingAddr + kGen4Warps * kGenStages * kGenStageBytes - kGenWindowBase;
static_assert(kGenRingAddr % 16u == 0 && kGenStagedSmemBytes + 1024u <= 228u * 1024u, "ring placement");
static_assert(!kGenRows || kGenRingAddr + kGen4Warps * kGenStages * kGenStageBytes <= kGenDecodeAddr, "ring must end below the table");

__device__ __forceinline__ void cpAsync4(uint32_t dst, const void* src, uint32_t srcBytes)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(srcBytes) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cpAsyncWait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

template <int kLevels, bool kX3, bool kY3, bool kStaged>
__global__ void __launch_bounds__(kGen4Warps * 32, 1) generalStrip4Kernel(const GenStripParams p)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  if(uint32_t(__cvta_generic_to_shared(smemRaw)) != kGenWindowBase)
    __trap();  // the absolute table addresses assume this window layout: fail loudly, never silently
  genSrgba8Init<kGen4Warps * 32>(smemRaw, p.tables);
  __syncthreads();
  const uint32_t  lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, laneAddr = kGenDecodeAddr | (lane * 4u);
  const LevelView L0 = p.lv[0], L1 = p.lv[1], L2 = p.lv[2];
  const float     fH1 = float(L1.h), fW1 = float(L1.w);
  const float     rcpY1 = kY3 ? genRcp(L1.h) : 0.f, rcpX1 = kX3 ? genRcp(L1.w) : 0.f;
  // second level (uniform run-time branches)
  const bool  x3b = kLevels == 2 && (L1.w & 1u), y3b = kLevels == 2 && (L1.h & 1u);
  const float fH2 = kLevels == 2 ? float(L2.h) : 0.f, fW2 = kLevels == 2 ? float(L2.w) : 0.f;
  const float rcpY2 = y3b ? genRcp(L2.h) : 0.f, rcpX2 = x3b ? genRcp(L2.w) : 0.f;
  const V4    zero = toV4(make_float4(0.f, 0.f, 0.f, 0.f));

  // staging ring of this warp: stage = (rows consumed so far) mod kGenStages
  const uint32_t ringBase = kGenRingAddr + warp * kGenStages * kGenStageBytes;
  uint32_t       consumed = 0, issued = 0;  // stage indices (wrap at kGenStages)

  const uint32_t numTasks = p.stripsX * p.segsY;
  gridDependencyWait();    // the previous kernel's levels are complete and visible (the constants above need none of them)
  gridLaunchDependents();  // the next kernel may start its own set-up as SMs become free
  for(uint32_t task = blockIdx.x + gridDim.x * warp; task < numTasks; task += gridDim.x * kGen4Warps)
  {
    const uint32_t sx = task % p.stripsX, sy = task / p.stripsX;
    const uint32_t xa = sx * 62u + 2u * lane;  // this lane's level +1 columns: xa, xa + 1
    const uint32_t c0 = 2u * xa;               // its source columns: c0 .. c0 + 3
    const bool     src0 = c0 < L0.w, src1 = c0 + 1u < L0.w, src2 = c0 + 2u < L0.w, src3 = c0 + 3u < L0.w;
    const bool     out1a = xa < L1.w, out1b = xa + 1u < L1.w && lane < 31u;  // lane 31 has no right-hand halo
    Taps           txa{zero.rg, zero.rg, zero.rg}, txb{zero.rg, zero.rg, zero.rg};
    if(kX3)
    {
      txa = genTaps(rcpX1, fW1, xa);
      txb = genTaps(rcpX1, fW1, xa + 1u);
    }
    const uint32_t x2   = sx * 31u + lane;  // its level +2 column
    const bool     out2 = kLevels == 2 && lane < 31u && x2 < L2.w;
    Taps           tx2{zero.rg, zero.rg, zero.rg};
    if(kLevels == 2 && x3b)
      tx2 = genTaps(rcpX2, fW2, x2);

    // rows of level +1 handled by this task: [ya, yb]
    uint32_t ya, yb, r2a = 0;
    if(kLevels == 2)
    {
      r2a                = sy * p.segRows;
      const uint32_t r2b = min(r2a + p.segRows, L2.h);
      ya                 = 2u * r2a;
      yb                 = min(y3b ? 2u * r2b : 2u * r2b - 1u, L1.h - 1u);
    }
    else
    {
      ya = sy * p.segRows;
      yb = min(ya + p.segRows, L1.h) - 1u;
    }

    struct Row4
    {
      uint32_t w0, w1, w2, w3;
    };
    const unsigned char* src   = L0.ptr + size_t(2u * ya) * L0.pitch + size_t(c0) * 4u;  // source row 2*ya
    auto                 load4 = [&](const unsigned char* row, Row4& r) {
      r.w0 = src0 ? __ldg(reinterpret_cast<const uint32_t*>(row)) : 0u;
      r.w1 = src1 ? __ldg(reinterpret_cast<const uint32_t*>(row + 4)) : 0u;
      r.w2 = src2 ? __ldg(reinterpret_cast<const uint32_t*>(row + 8)) : 0u;
      r.w3 = src3 ? __ldg(reinterpret_cast<const uint32_t*>(row + 12)) : 0u;
    };

    V4 carry0 = zero, carry1 = zero, carry2 = zero, carry3 = zero;  // decoded source row 2y (3-tap only)
    if(kY3)
    {
      Row4 r;
      load4(src, r);
      carry0 = genDecodeTexel(laneAddr, r.w0), carry1 = genDecodeTexel(laneAddr, r.w1);
      carry2 = genDecodeTexel(laneAddr, r.w2), carry3 = genDecodeTexel(laneAddr, r.w3);
    }
    struct Slot
    {
      Row4 a, b;  // the two new source rows of an output row
    };
    const unsigned char* nextRows = kY3 ? src + L0.pitch : src;
    const size_t         rowStep  = 2u * size_t(L0.pitch);
    auto                 loadRow  = [&](bool valid, Slot& s) {
      if(valid)
      {
        load4(nextRows, s.a);
        load4(nextRows + L0.pitch, s.b);
      }
      nextRows += rowStep;
    };
    Slot r0{{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}}, r1{{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}};
    // staged variant: this lane's four words of a row segment (word lane + 32 j) and whether they exist
    const unsigned char* segSrc = L0.ptr + size_t(2u * 62u * sx + lane) * 4u;  // word `lane` of the strip's segment in row 0
    uint32_t             segBytes[4];
#pragma unroll
    for(uint32_t j = 0; j < 4u; ++j)
      segBytes[j] = 2u * 62u * sx + lane + 32u * j < L0.w ? 4u : 0u;
    // The copies of output row yy: its two new source rows.  One commit group per output row, also when the row does
    // not exist (an empty group keeps the wait count uniform).
    auto stageIssue = [&](uint32_t yy) {
      if(yy <= yb)
      {
        const uint32_t       dst = ringBase + issued * kGenStageBytes + lane * 4u;
        const unsigned char* r0  = segSrc + size_t(kY3 ? 2u * yy + 1u : 2u * yy) * L0.pitch;
#pragma unroll
        for(uint32_t j = 0; j < 4u; ++j)
        {
          cpAsync4(dst + 128u * j, r0 + 128u * j, segBytes[j]);
          cpAsync4(dst + 512u + 128u * j, r0 + L0.pitch + 128u * j, segBytes[j]);
        }
      }
      cpAsyncCommit();
      issued = issued + 1u == kGenStages ? 0u : issued + 1u;
    };
    if(kStaged)
    {
#pragma unroll
      for(uint32_t i = 0; i < kGenStages; ++i)
        stageIssue(ya + i);
    }
    else
    {
      loadRow(true, r0);
      loadRow(ya + 1u <= yb, r1);
    }
    V4 q0a = zero, q1a = zero, q0b = zero, q1b = zero;  // last level +1 values of the lane's two columns

    unsigned char* d1 = L1.ptr + size_t(ya) * L1.pitch + size_t(xa) * 4u;
    unsigned char* d2 = kLevels == 2 ? L2.ptr + size_t(r2a) * L2.pitch + size_t(x2) * 4u : nullptr;
    uint32_t       y2 = r2a;  // next row of level +2 to emit

    auto row = [&](uint32_t y, Slot& slot, auto parity) {
      constexpr bool kOdd = decltype(parity)::value;
      Slot           m    = slot;
      if(kStaged)
      {
        // wait for this row's stage (the oldest of kGenStages groups), read the lane's four columns of both source
        // rows, hand the stage back
        cpAsyncWait<kGenStages - 1>();
        __syncwarp();  // every lane's words of the stage have landed
        const uint32_t mine = ringBase + consumed * kGenStageBytes + lane * 16u;
        consumed            = consumed + 1u == kGenStages ? 0u : consumed + 1u;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(m.a.w0), "=r"(m.a.w1), "=r"(m.a.w2), "=r"(m.a.w3) : "r"(mine));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+512];" : "=r"(m.b.w0), "=r"(m.b.w1), "=r"(m.b.w2), "=r"(m.b.w3) : "r"(mine));
        __syncwarp();  // all lanes have read the stage before anyone overwrites it
        stageIssue(y + kGenStages);
      }
      else
        loadRow(y + 2u <= yb, slot);
      // ---- vertical reduction of the lane's four source columns ----
      const V4 a0 = genDecodeTexel(laneAddr, m.a.w0), a1 = genDecodeTexel(laneAddr, m.a.w1);
      const V4 a2 = genDecodeTexel(laneAddr, m.a.w2), a3 = genDecodeTexel(laneAddr, m.a.w3);
      const V4 b0 = genDecodeTexel(laneAddr, m.b.w0), b1 = genDecodeTexel(laneAddr, m.b.w1);
      const V4 b2 = genDecodeTexel(laneAddr, m.b.w2), b3 = genDecodeTexel(laneAddr, m.b.w3);
      V4       h0, h1, h2, h3;
      if(kY3)
      {
        const Taps ty = genTaps(rcpY1, fH1, y);
        h0            = genReduce3(ty.w0, carry0, ty.w1, a0, ty.w2, b0);
        h1            = genReduce3(ty.w0, carry1, ty.w1, a1, ty.w2, b1);
        h2            = genReduce3(ty.w0, carry2, ty.w1, a2, ty.w2, b2);
        h3            = genReduce3(ty.w0, carry3, ty.w1, a3, ty.w2, b3);
        carry0 = b0, carry1 = b1, carry2 = b2, carry3 = b3;
      }
      else
      {
        h0 = genReduce2(a0, b0), h1 = genReduce2(a1, b1);
        h2 = genReduce2(a2, b2), h3 = genReduce2(a3, b3);
      }
      // ---- horizontal reduction: level +1 columns xa (source c0, c0+1[, c0+2]) and xa+1 (c0+2, c0+3[, c0+4]) ----
      V4 oa, ob;
      if(kX3)
      {
        const V4 h4 = shflDown(h0, 1);  // source column c0 + 4 = first column of lane + 1
        oa          = genReduce3(txa.w0, h0, txa.w1, h1, txa.w2, h2);
        ob          = genReduce3(txb.w0, h2, txb.w1, h3, txb.w2, h4);
      }
      else
      {
        oa = genReduce2(h0, h1);
        ob = genReduce2(h2, h3);
      }
      const uint32_t wa = genEncWord(toFloat4(oa)), wb = genEncWord(toFloat4(ob));
      if(out1a)
        *reinterpret_cast<uint32_t*>(d1) = wa;
      if(out1b)
        *reinterpret_cast<uint32_t*>(d1 + 4) = wb;
      d1 += L1.pitch;

      // ---- level +2, float32 carry ----
      if(kLevels == 2)
      {
        V4   ga = zero, gb = zero;
        bool emit = false;
        if(y3b)
        {
          // rows 2 y2, 2 y2 + 1, 2 y2 + 2: emit when the third arrives (even row, not the segment's first)
          if(!kOdd)
          {
            if(y != ya)
            {
              const Taps ty = genTaps(rcpY2, fH2, y2);
              ga            = genReduce3(ty.w0, q0a, ty.w1, q1a, ty.w2, oa);
              gb            = genReduce3(ty.w0, q0b, ty.w1, q1b, ty.w2, ob);
              emit          = true;
            }
            q0a = oa, q0b = ob;
          }
          else
            q1a = oa, q1b = ob;
        }
        else
        {
          if(kOdd)
          {
            ga   = genReduce2(q0a, oa);
            gb   = genReduce2(q0b, ob);
            emit = true;
          }
          else
            q0a = oa, q0b = ob;
        }
        if(emit)  // warp-uniform
        {
          V4 o2;
          if(x3b)
          {
            const V4 gc = shflDown(ga, 1);  // level +1 column xa + 2 = first column of lane + 1
            o2          = genReduce3(tx2.w0, ga, tx2.w1, gb, tx2.w2, gc);
          }
          else
            o2 = genReduce2(ga, gb);
          const uint32_t w2 = genEncWord(toFloat4(o2));
          if(out2)
            *reinterpret_cast<uint32_t*>(d2) = w2;
          d2 += L2.pitch;
          ++y2;
        }
      }
    };
    uint32_t y = ya;
    for(; y < yb; y += 2u)
    {
      row(y, r0, std::false_type{});
      row(y + 1u, r1, std::true_type{});
    }
    if(y == yb)
      row(y, r0, std::false_type{});
  }
}

}  // namespace nvpyr
