// nvpyr_fast_srgba8.cuh -- the hot kernel: sRGBA8 fast pipeline, 2..6 levels per launch.
//
// Same float32 expression trees as fastKernel<Srgba8, M> (nvpyr_kernels.cuh) and therefore
// the same bits, but arranged for the B200's instruction and shared-memory budgets.  At HBM speed an SM may
// spend 121 cycles on a 64x8 slab (2 KB in, 0.67 KB out): 484 warp-instruction issue slots and 121 LSU
// wavefronts.  The exact path needs 48 table decodes and 16 channel encodes per lane and slab, so this kernel
// is issue / LSU bound, not DRAM bound; as measured (round 2): 405 instructions and ~106
// wavefronts per slab, 0.94 of the measured HBM peak for the launch, on any content.
//
//  * decode: one PRMT builds the shared-memory address (code << 8 | lane << 2) straight from
//    the packed texel word, one LDS reads the lane-private copy of the 256-entry table
//    (stride 256 B per code: no bank conflict for any data).  Alpha goes through I2F (XU
//    pipe) + FMUL to keep the LSU free.
//  * encode (round 2, NVPYR_FAST_ENC_ROWS): the bucket is the key of z = RN(x + 1/32 - 2^-14) instead of the key of
//    x: 645 rows cover [0, 1], few enough for a lane-private copy of every entry, so the look-up -- FFMA, PRMT, LDS,
//    IADD3 -- is one conflict-free wavefront like the decode.  Decode and encode share one table of 646 rows x 256
//    bytes (low half rows: decode copies / level +3 stashes, high half rows: encode copies).  Level +1 (12 of every
//    16 encodes) needs no clamp: exact zero has a row to itself.
//  * deferred normalisation: 0.25*((a+b)+(c+d)) is carried as the un-normalised sum
//    S = (a+b)+(c+d) = 4^k * value, and the decode table is pre-scaled by 2^-100.  Scaling
//    by a power of two commutes with rounding (nothing here is subnormal or overflows), so
//    every deeper sum and every stored code is bit-identical; the FMULs disappear.  Alpha uses the exact
//    constant 255 / 4^k.
//  * level +3 is a "transpose-reduce": in two shuffle steps the four threads of a 2x2 block
//    end up holding ONE channel each of the common result (3 SHFL + 3 FADD per thread
//    instead of 12 + 12), encode it in parallel and gather the four bytes with two PRMTs.
//  * warp-autonomous tiles: one WARP owns a 64 x max(8, 2^M) input tile and walks it as
//    64x8 slabs (lane = 4x4 texels), keeping the level +3 sums in a 1 KB warp-private stash;
//    levels +4..+M are finished by the same warp.  No CTA-wide barrier exists after
//    the table set-up.  Tiles are only as tall as the step needs, so steps with M < 6 expose 2-8x more
//    independent warp tasks.  The warp index is taken through a shuffle so that the compiler keeps the tile
//    bookkeeping, the TMA coordinates and the mbarrier address in uniform registers (no spill, no ELECT/R2UR loops).
//  * packed adds: FADD2 (add.rn.f32x2) performs two IEEE float32 additions per issue slot;
//    the sums run on (R, G) and (B, A) pairs.
//  * the level-0 slab of a warp is staged in shared memory by the TMA unit (one 2-D tensor-map
//    copy per slab on the warp's own mbarrier, NVPYR_FAST_TMA below) while the previous slab is
//    processed; the batch / fused-premultiply / slab-task kernels prefetch the next slab's four
//    16-byte rows into registers instead.
//
// CTA = 32 or 24 warps (template parameter) sharing the tables; one CTA per SM; 161.5 KB of tables and stashes
// + 64 KB of TMA ring for the kernels that use it.  With LDG loads, shared memory beyond ~200 KB
// leaves the SM too little L1 for the loads in flight (a 213 KB variant of the register path ran
// 19 % slower); loads staged by the TMA unit do not pass through L1.
//
// The compile-time switches below select measured-and-rejected variants (each with its numbers) for A/B runs; the
// round-1 bucket tables (NVPYR_FAST_ENC_ROWS = 0: NVPYR_ENC_WAYS, NVPYR_FAST_ENC_CLAMP, ..._LOW_OCTAVES,
// NVPYR_FAST_L3_IN_DECODE) are among them.
#pragma once
#include <stddef.h>

#include "nvpyr_kernels.cuh"
#include <cuda.h>  // CUtensorMap (types only: the encoder is looked up at run time, libcuda is not linked)

namespace nvpyr {

#ifndef NVPYR_FAST_WARPS
#define NVPYR_FAST_WARPS 32
#endif
#ifndef NVPYR_FAST_PREFETCH
#define NVPYR_FAST_PREFETCH 1
#endif
// NVPYR_ENC_WAYS = W: every bucket entry is stored W times (4 W bytes) and lane l reads copy l & (W - 1),
// so lanes with different l & (W - 1) can never collide on a bank (the un-replicated table costs ~3.8
// wavefronts per warp-wide lookup on random data).  Needs one 1024-thread CTA per SM.
// NVPYR_FAST_ENC_LOW_OCTAVES = n: the table is extended n octaves below 2^-13 (everything there encodes to 0).
// A level whose smallest non-zero value, linearFromSrgb(1) / 4^K, still lies inside the table is encoded
// without a clamp; the other levels clamp the value from below with one FMNMX first.  n = 11 (default) covers
// all six levels; n = 3 covers levels +1 and +2 (15 of every 16 encodes) with a table 33 % shorter (164 KB per
// CTA instead of 197 KB).  NVPYR_FAST_ENC_CLAMP = 1 clamps everywhere (n = 0).  Measured at 16384^2 (us, Julia /
// uniform-random input): n = 11: 242.4 / 273.4;  n = 3: 247.4 / 268.1;  clamp everywhere: 249.3 / 268.8.
// Shared memory beyond ~200 KB leaves the SM too little L1 for the loads in flight: the same code padded to
// 213 KB runs at 311 / 335 us.
#ifndef NVPYR_ENC_WAYS
#define NVPYR_ENC_WAYS 8
#endif
#ifndef NVPYR_FAST_ENC_CLAMP
#define NVPYR_FAST_ENC_CLAMP 1  // round 2: with the stashes inside the decode table (below) every encode clamps
#endif
#ifndef NVPYR_FAST_TMA
#define NVPYR_FAST_TMA 2  // see below
#endif
#ifndef NVPYR_FAST_ENC_LOW_OCTAVES
#if NVPYR_FAST_TMA
#define NVPYR_FAST_ENC_LOW_OCTAVES 3  // the TMA ring needs the short table
#else
#define NVPYR_FAST_ENC_LOW_OCTAVES 11
#endif
#endif
// NVPYR_FAST_SLAB_UNROLL = 2: the prefetch registers of one slab are the working registers of the next, so the
// register moves of the rotating double buffer disappear (422 -> 350 instructions per slab) -- and the kernel
// gets SLOWER (258 us): ptxas tracks every LDG of the loop with one scoreboard, so the first use of slab s+1
// also waits for the just-issued loads of slab s+2 (long-scoreboard stalls 0.29 -> 3.6 warps per issue, ncu).
// NVPYR_FAST_PIN_PREFETCH = 1 orders the new loads after that first use through a fake address dependency
// (260 us: ptxas then spills inside the loop).  NVPYR_FAST_UNCOND_LOADS = 1 loads unconditionally (lanes outside
// the image read its first rows).  All three are kept for A/B runs only; the measured best is the default.
#ifndef NVPYR_FAST_SLAB_UNROLL
#define NVPYR_FAST_SLAB_UNROLL 1
#endif
#ifndef NVPYR_FAST_PAD_BYTES
#define NVPYR_FAST_PAD_BYTES 16
#endif
// NVPYR_FAST_TMA = 2 (default): the level-0 slab of a warp (8 rows x 256 bytes) is staged in a 2 KB shared-memory
// buffer by the TMA unit -- ONE 2-D tensor-map copy per slab (cp.async.bulk.tensor.2d, box 64 x 8 texels, parts
// outside the image zero-filled) issued by lane 0 and completing on the warp's own mbarrier -- and read back with four
// LDS.128, instead of four LDG.128 into prefetch registers.  Single-buffered: the copy of slab s+1 is issued as soon
// as slab s has been read into registers (a proxy fence orders the reads before the asynchronous writes).  The ring
// of the 32 warps is 64 KB, which only fits next to the 3-low-octave encode table (224.6 KB of the 227 KB a CTA may
// have); with the loads no longer passing through L1 the large carve-out is harmless.  The host encodes the tensor
// map of the step's input level per launch (a kernel parameter); a batch launch gets one map per image in device memory
// (FastBatch::maps).  The fused premultiply and slab tasks keep the register path (and do not allocate the ring).
// Measured at 16384^2 against the register path with the same table (us, 6-level kernel): Julia 242.3 -> 237.2
// (92.3 % of HBM peak), smooth gradient 247.7 -> 245.7, uniform-random bytes 273.5 -> 278.0; the other sizes of the
// config table within +-1 %.
// NVPYR_FAST_TMA = 1: eight cp.async.bulk ROW copies per slab instead (no tensor map): the per-lane issue loops cost
// more than they save (Julia 286 us).  NVPYR_FAST_TMA = 0: the register path for every kernel.
#ifndef NVPYR_FAST_PIN_PREFETCH
#define NVPYR_FAST_PIN_PREFETCH 0
#endif
#ifndef NVPYR_FAST_UNCOND_LOADS
#define NVPYR_FAST_UNCOND_LOADS 0
#endif
// NVPYR_FAST_L3_IN_DECODE = 1: the per-warp stashes of level +3 sums (32 KB) live in the UNUSED HALVES of the decode
// table's rows (row = 256 bytes per code, of which the 32 lane copies take 128): stash row (warp, slab) -- eight
// texels of 16 bytes -- is the spare half of decode row 8 * warp + slab.  Frees 32 KB of shared memory, which is what
// lets a 16-way encode table (131 KB) fit under the ~200 KB above which per-lane global loads lose their L1.  The
// zero words of the clamp-free encode would collide with the stashes, so this layout clamps every encode.
// Measured at 16384^2 (round 2, one box, us whole chain / 6-level kernel; Julia | uniform random | gradient):
//   round-1 layout (TMA ring, 8-way, 3 low octaves, stashes apart, 224.6 KB)  252.1 / 243.4 | 291.2 / 283.2 | 260.1 / 250.9
//   this layout    (TMA ring, 8-way, clamp everywhere, 192 KB): the default   253.1 / 242.8 | 283.3 / 274.8 | 255.1 / 245.2
//   register path, 16-way, clamp, this layout (193 KB)                        265.3 / 256.5 | 271.5 / 263.2 | 265.7 / 257.1
//   register path, 8-way, clamp, this layout (132 KB)                         264.5 / 255.4 | 282.9 / 273.6 | 267.0 / 256.7
//   register path, 8-way, 3 low octaves, stashes apart (164 KB)               259.9 / 249.9 | 282.5 / 273.9 | 263.6 / 253.1
// The 32 KB given back to L1 are worth 3 % on high-entropy input and cost nothing on the reference's Julia texture.
// A 16-way table takes another 10 us off the random input but only fits without the TMA ring (64 + 131 + 64 KB >
// 227 KB), and the register path costs the Julia input 13 us.
#ifndef NVPYR_FAST_L3_IN_DECODE
#define NVPYR_FAST_L3_IN_DECODE 1
#endif
// NVPYR_FAST_DYNAMIC_TILES = 1: after its first (statically dealt) tile a warp takes further tiles from one global
// counter, so that CTAs that start late (their SM still runs the previous chain's tail kernel) take fewer tiles.
// Measured (round 2, same box, 16384^2 whole chain / kernel alone): static 252.7 / 244.6 us on Julia, 284.3 / 275.4 on
// random bytes; dynamic 261.4 / 254.1 and 292.8 / 283.4 -- the hand-out costs 9 us even with the atomic issued a whole
// tile ahead of its use (8192^2 80.5 -> 86.9 us, 1080p 14.7 -> 16.5 us).  Kept for A/B runs; static dealing is the default.
#ifndef NVPYR_FAST_DYNAMIC_TILES
#define NVPYR_FAST_DYNAMIC_TILES 0
#endif
// NVPYR_FAST_ENC_ROWS = 1 (default, round 2): the encode bucket is keyed on RN(x + 1/32 - 2^-14) instead of on x
// (nvpyr_functors.cuh "Row table"): 645 rows cover [0, 1], so every lane owns a private copy of every entry and a
// warp-wide encode look-up is ONE conflict-free wavefront for any data, like the decode.  Decode and encode share one
// table of 645 rows x 256 bytes: bytes 0..127 of row r = the 32 lane copies of linearFromSrgb(r) (r < 256) or a
// level +3 stash row (256 <= r < 512), bytes 128..255 = the 32 lane copies of encode entry r.  The look-up is
// FFMA (z = S' * 2^k + c, on the idle FMA pipe), PRMT (address = key(z) << 8 | lane << 2), LDS, IADD3 -- one
// 16-lane-ALU instruction fewer than shift + mask -- and level +1 (12 of every 16 encodes) needs no clamp: exact
// zero has row 0 to itself.  161 KB of tables + 64 KB TMA ring = 225.5 KB.
// NVPYR_FAST_ENC_ROWS = 0: the bank-partitioned bucket tables of round 1 (NVPYR_ENC_WAYS etc. below apply).
#ifndef NVPYR_FAST_ENC_ROWS
#define NVPYR_FAST_ENC_ROWS 1
#endif
// NVPYR_FAST_OPAQUE_PATH = 1: slabs whose texels are all opaque skip the alpha arithmetic (see the slab loop): ~10 % fewer
// instructions on images without alpha -- and SLOWER everywhere (measured, round 2, 16384^2 chain: uniform random bytes
// 238.5 -> 250.7 us, the same colours with alpha 255 237.7 -> 244.0 us; 4096^2 21.6 -> 22.7 / 22.6 us): the second copy
// of the level +1 / +2 code costs registers (spills return in the 64- and 80-register builds) and the test itself sits
// on the critical path of every slab.  Off by default; kept for A/B runs.
#ifndef NVPYR_FAST_OPAQUE_PATH
#define NVPYR_FAST_OPAQUE_PATH 0
#endif
constexpr bool     kEncRows         = NVPYR_FAST_ENC_ROWS != 0;
constexpr bool     kDynTiles        = NVPYR_FAST_DYNAMIC_TILES != 0;
constexpr bool     kL3InDecode      = kEncRows || NVPYR_FAST_L3_IN_DECODE != 0;
constexpr uint32_t kL3RowFloats     = kL3InDecode ? 64u : 32u;  // floats from one stash row (8 texels) to the next
constexpr bool     kFastUncondLoads = NVPYR_FAST_UNCOND_LOADS != 0;
constexpr int      kFastSlabUnroll = NVPYR_FAST_SLAB_UNROLL;
constexpr uint32_t kEncWays       = NVPYR_ENC_WAYS;
constexpr bool     kEncClamp      = NVPYR_FAST_ENC_CLAMP != 0;
constexpr int      kFastWarps     = NVPYR_FAST_WARPS;
constexpr int      kFastCtasPerSm = NVPYR_ENC_WAYS == 1 ? 2 : 1;
constexpr bool     kFastPrefetch  = NVPYR_FAST_PREFETCH != 0;
constexpr int      kDecScaleExp   = kFastDecScaleExp;  // decode table holds 2^-100 * linearFromSrgb(code)
constexpr uint32_t kEncKeysPerOctave = 1u << (23 - kFastEncShift);
constexpr uint32_t kEncLowOctaves = kEncClamp ? 0 : NVPYR_FAST_ENC_LOW_OCTAVES;  // bucket table extended below 2^-13
// Is every non-zero value of level K inside the extended table?  linearFromSrgb(1) = 2^-11.7: sums of level K
// are >= 2^-(11.7 + 2K); the premultiply pre-pass (K = 0) makes products down to 2^-19.7.
constexpr bool encCovered(int K)
{
  return K == 0 ? kEncLowOctaves >= 7 : 2u * uint32_t(K) <= kEncLowOctaves + 1u;
}
constexpr uint32_t kEncMinKeyExt  = kFastEncMinKey - kEncLowOctaves * kEncKeysPerOctave;
constexpr uint32_t kEncEntriesExt = kFastEncEntries + kEncLowOctaves * kEncKeysPerOctave;
constexpr uint32_t kEncStride     = 4u * kEncWays;  // bytes per bucket entry (all copies)
constexpr uint32_t kEncStrideLog2 = kEncWays == 1 ? 2 : kEncWays == 2 ? 3 : kEncWays == 4 ? 4 : kEncWays == 8 ? 5 : 6;
static_assert((1u << kEncStrideLog2) == kEncStride, "1, 2, 4, 8 or 16 copies");

#if NVPYR_FAST_ENC_ROWS
struct Srgba8FastSmem
{
  // row r: floats 0..31 = lane copies of 2^-100 * linearFromSrgb(r) (r < 256) / stash row r - 256 (256 <= r < 512),
  //        floats 32..63 = lane copies of encode entry r
  float decode[kRowEncRows * 64];
  alignas(128) unsigned char ring[kFastWarps][2048];  // per warp: the level-0 slab in flight (8 rows x 256 bytes)
  unsigned long long tmaBar[kFastWarps];              // per warp: mbarrier the slab's copy completes on
};
static_assert(NVPYR_FAST_TMA != 1, "row-table layout: tensor-map staging or the register path");
#else
struct Srgba8FastSmem
{
  float    decode[256 * 64];        // [code][64]: floats 0..31 = per-lane copies, 32..63 spare (zero words live there)
  float    pad[32];                 // keeps the zero words in the spare halves (see encScaled)
  alignas(16) uint32_t encode[(kEncEntriesExt + 3) * kEncWays];  // bucket table, extended downwards, kEncWays copies per entry
  alignas(16) float l3[kL3InDecode ? 1 : kFastWarps][8][8][4];  // per warp: level +3 sums of its 64x64 tile, [slab][x][channel]
  unsigned char unused[NVPYR_FAST_PAD_BYTES];  // A/B experiments on the shared-memory carve-out
#if NVPYR_FAST_TMA
  alignas(128) unsigned char ring[kFastWarps][2048];  // per warp: the level-0 slab in flight (8 rows x 256 bytes)
  unsigned long long tmaBar[kFastWarps];              // per warp: mbarrier the slab's row copies complete on
#endif
};
#endif
#if NVPYR_FAST_TMA || NVPYR_FAST_ENC_ROWS
static_assert(sizeof(Srgba8FastSmem) + 1024 <= 227 * 1024, "TMA ring: build with NVPYR_FAST_ENC_LOW_OCTAVES=3");
static_assert(offsetof(Srgba8FastSmem, ring) % 128 == 0 && alignof(Srgba8FastSmem) >= 128,
              "cp.async.bulk.tensor destinations must be 128-byte aligned inside a 128-byte aligned block");
#endif
constexpr bool kFastTma = NVPYR_FAST_TMA != 0;
static_assert(kEncRows || !kL3InDecode || (NVPYR_FAST_ENC_CLAMP != 0 && kFastWarps * 8 <= 256),
              "stashes in the decode table's spare halves: every encode clamps (no zero words), at most 256 stash rows");
// First float of the stash of warp (or local tile) j.
__device__ __forceinline__ float* stashOf(Srgba8FastSmem& sm, uint32_t j)
{
#if NVPYR_FAST_ENC_ROWS
  static_assert(256 + kFastWarps * 8 <= kRowEncRows, "stash rows live below the encode entries of rows 256..");
  return &sm.decode[(256u + j * 8u) * 64u];
#else
  return kL3InDecode ? &sm.decode[j * 8u * 64u + 32u] : &sm.l3[kL3InDecode ? 0u : j][0][0][0];
#endif
}
// Dynamic shared memory of a launch: only the kernels that stage through the TMA ring pay for it (the others keep
// the SM's L1 for their loads in flight).
constexpr size_t fastSmemBytes(bool tma)
{
#if NVPYR_FAST_TMA || NVPYR_FAST_ENC_ROWS
  return tma ? sizeof(Srgba8FastSmem) : offsetof(Srgba8FastSmem, ring);
#else
  return sizeof(Srgba8FastSmem) + 0 * size_t(tma);
#endif
}

// mbarrier / bulk-copy primitives of the TMA variant (PTX ISA 8.x, sm_90+)
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulkCopyG2S(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#if NVPYR_FAST_TMA == 2
__device__ __forceinline__ void tensorCopy2D(uint32_t dst, const CUtensorMap* map, uint32_t x, uint32_t y, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(x), "r"(y), "r"(bar)
               : "memory");
}
struct FastTensorMap
{
  alignas(64) CUtensorMap map;
};
#else
struct FastTensorMap
{
  int unused;
};
#endif
#if !NVPYR_FAST_ENC_ROWS
static_assert(offsetof(Srgba8FastSmem, decode) == 0 && offsetof(Srgba8FastSmem, encode) == 65536 + 128,
              "encScaled's zero words assume this layout");
#endif

#if NVPYR_FAST_ENC_ROWS
// Per-level constants of the scaled encode.  The carried value is S' = 2^-E * 4^K * x, so
// bits(x) = bits(S') + ((E - 2K) << 23) and RN(x + c) = fma(S', 2^(E - 2K), c).
template <int K>
struct EncConst
{
  static constexpr uint32_t kAdd = uint32_t(kDecScaleExp - 2 * K) << 23;
  // Level +1 sees exact zero or values >= linearFromSrgb(1) / 4: zero has row 0 to itself and its entry is made for
  // the pattern bits(0) + kAdd, everything else lies inside the table.  Deeper levels (and the premultiply, K = 0)
  // can see smaller non-zero values: they clamp S' from below to 2^-13 (below the first threshold) first.
  static constexpr bool     kClamp   = K != 1;
  static constexpr uint32_t kMinBits = kEncMinBits - kAdd;
  static constexpr uint32_t kScaleBits = uint32_t(127 + kDecScaleExp - 2 * K) << 23;  // 2^(E - 2K)
};
static_assert(EncConst<1>::kAdd == kRowEncZeroBits, "row 0 of the encode table is built for the zero of level +1");

// Address bias of the encode half-rows: row key k lives at byte (k - kRowEncFirstKey) * 256 + 128 of the table.
constexpr int32_t kRowEncBias = int32_t(kRowEncFirstKey << 8) - 128;

// kThreads = threads of the CTA.  Every load of a thread is in flight before its first store: the set-up is pure
// latency (five to seven table entries per thread, each an L2 round trip if taken one after the other), and it is on the
// critical path of every chain whose main launch is short -- the next chain's CTAs build their tables while the tail of
// this one runs, and the CTA on the SM that was freed last finishes building after that tail has ended.
template <int kThreads>
__device__ __forceinline__ void srgba8FastInit(Srgba8FastSmem& sm, const DeviceTables* t)
{
  // thread -> (row, 16-byte column): eight consecutive threads write the 128 contiguous bytes of one half row
  uint4*             tab    = reinterpret_cast<uint4*>(sm.decode);  // 16 uint4 per row: 0..7 decode / stash, 8..15 encode
  constexpr uint32_t kItems = kRowEncRows * 8u, kPerThread = (kItems + kThreads - 1) / kThreads;
  uint32_t           e[kPerThread];
  float              d[kPerThread];
#pragma unroll
  for(uint32_t k = 0; k < kPerThread; ++k)
  {
    const uint32_t i = threadIdx.x + k * kThreads, row = i >> 3;
    if(i < kItems)
    {
      e[k] = __ldg(&t->encodeRows[row]);
      if(row < 256u)
        d[k] = __ldg(&t->decode[row]);
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
      {
        const uint32_t v = __float_as_uint(__fmul_rn(d[k], 7.888609052210118e-31f));  // * 2^-100, exact
        tab[row * 16u + col] = make_uint4(v, v, v, v);
      }
    }
  }
}
#else
// Per-level constants of the scaled encode.  The carried value is S' = 2^-E * 4^K * x, so
// bits(x) = bits(S') + ((E - 2K) << 23) and key(x) = key(S') + (E - 2K) * kEncKeysPerOctave.
template <int K>
struct EncConst
{
  static constexpr uint32_t kAdd  = uint32_t(kDecScaleExp - 2 * K) << 23;
  static constexpr int32_t  kBias = (int32_t(kEncMinKeyExt) - (kDecScaleExp - 2 * K) * int32_t(kEncKeysPerOctave)) * int32_t(kEncStride);  // bytes, > 0
  // float index (into Srgba8FastSmem::decode) of the word that an exact zero reads
  static constexpr int32_t kZeroIndex = (65536 + 128 - kBias) / 4;  // meaningful where !kClamp
  static constexpr bool kClamp = !encCovered(K);
  static_assert(kClamp || (kBias > 0 && kBias <= 65536 && (65536 + 128 - kBias) % 256 == 128),
                "zero word must fall into a spare half row");
  // clamped levels: S' of 2^-13 (below the first threshold; inside the table)
  static constexpr uint32_t kMinBits = kEncMinBits - kAdd;
};

template <int K>
__device__ __forceinline__ void putZeroWord(Srgba8FastSmem& sm)
{
  if constexpr(!EncConst<K>::kClamp)
    for(uint32_t w = 0; w < kEncWays; ++w)
      sm.decode[EncConst<K>::kZeroIndex + w] = __uint_as_float(0u - EncConst<K>::kAdd);
}

template <int kThreads>
__device__ __forceinline__ void srgba8FastInit(Srgba8FastSmem& sm, const DeviceTables* t)
{
  // 512 threads: thread -> (code, half): 16 lane slots = 4 x float4
  for(uint32_t i = threadIdx.x; i < 512u; i += blockDim.x)
  {
    const uint32_t code = i >> 1, half = i & 1u;
    const float    v    = __fmul_rn(__ldg(&t->decode[code]), 7.888609052210118e-31f);  // * 2^-100, exact
    float4*        d    = reinterpret_cast<float4*>(&sm.decode[code * 64u + half * 16u]);
    const float4   v4   = make_float4(v, v, v, v);
    d[0] = v4, d[1] = v4, d[2] = v4, d[3] = v4;
  }
  constexpr uint32_t kLow = kEncEntriesExt - kFastEncEntries;
  static_assert(kLow % 4 == 0, "upper part of the table must stay 16-byte aligned");
  if(kEncWays == 1)
  {
    for(uint32_t i = threadIdx.x; i < kLow; i += blockDim.x)
      sm.encode[i] = 0u - ((kEncMinKeyExt + i) << kFastEncShift);  // code 0, no threshold, pre-biased
    copyTableWide<kFastWarps * 32>(reinterpret_cast<uint4*>(&sm.encode[kLow]),
                                   reinterpret_cast<const uint4*>(t->encodeFast), kFastEncEntriesPadded / 4);
  }
  else
  {
    // thread -> (entry, group of four copies): one 16-byte store
    constexpr uint32_t kQuads = kEncWays >= 4 ? kEncWays / 4 : 1;
    for(uint32_t i = threadIdx.x; i < kEncEntriesExt * kQuads; i += blockDim.x)
    {
      const uint32_t k = i / kQuads;
      const uint32_t v = k < kLow ? 0u - ((kEncMinKeyExt + k) << kFastEncShift) : __ldg(&t->encodeFast[k - kLow]);
      if(kEncWays >= 4)
        reinterpret_cast<uint4*>(sm.encode)[i] = make_uint4(v, v, v, v);
      else
        reinterpret_cast<uint2*>(sm.encode)[i] = make_uint2(v, v);
    }
  }
  if(threadIdx.x == 0)
  {
    putZeroWord<0>(sm);  // premultiply pre-pass (K = 0: values are not sums)
    putZeroWord<1>(sm), putZeroWord<2>(sm), putZeroWord<3>(sm);
    putZeroWord<4>(sm), putZeroWord<5>(sm), putZeroWord<6>(sm);
  }
}

#endif

// 2^-100 * linearFromSrgb of byte k (0..2) of a packed texel: PRMT + LDS.
template <int kByte>
__device__ __forceinline__ float dec8(const unsigned char* decodeBytes, uint32_t w, uint32_t laneOff)
{
  const uint32_t off = __byte_perm(w, laneOff, 0x5504u | (uint32_t(kByte) << 4));  // code << 8 | lane << 2
  return *reinterpret_cast<const float*>(decodeBytes + off);
}
// alpha * (1/255): shaders/srgb.h:60
__device__ __forceinline__ float decAlpha(uint32_t w)
{
  return __fmul_rn(float(w >> 24), 1.0f / 255.0f);
}

// Packed float32 pairs: Blackwell's FADD2 performs two IEEE round-to-nearest float32 additions
// in ONE issue slot (add.rn.f32x2).  The kernel is issue-bound, so the pairwise sums of
// (R, G) and (B, A) are done two channels at a time; each lane of the pair is rounded exactly
// like a scalar __fadd_rn.
struct F2
{
  unsigned long long v;
};
__device__ __forceinline__ F2 pack2(float lo, float hi)
{
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(F2 a, float& lo, float& hi)
{
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ F2 add2(F2 a, F2 b)
{
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
// A texel value as two packed pairs: rg = (R, G), ba = (B, A).
struct V4
{
  F2 rg, ba;
};
__device__ __forceinline__ V4 add4(V4 a, V4 b)
{
  V4 r;
  r.rg = add2(a.rg, b.rg);
  r.ba = add2(a.ba, b.ba);
  return r;
}
__device__ __forceinline__ float4 toFloat4(V4 a)
{
  float4 f;
  unpack2(a.rg, f.x, f.y);
  unpack2(a.ba, f.z, f.w);
  return f;
}
__device__ __forceinline__ V4 toV4(float4 f)
{
  V4 r;
  r.rg = pack2(f.x, f.y);
  r.ba = pack2(f.z, f.w);
  return r;
}

// Decoded texel (rgb pre-scaled by 2^-100, alpha = a * (1/255)) as packed pairs.  kOpaque: the caller knows that
// alpha is 255, i.e. 255 * (1/255) = 1.0f exactly (the product rounds to 1: checked by the opaque-path parity tests).
template <bool kOpaque = false>
__device__ __forceinline__ V4 decodeTexel(const unsigned char* dec, uint32_t laneOff, uint32_t w)
{
  V4 r;
  r.rg = pack2(dec8<0>(dec, w, laneOff), dec8<1>(dec, w, laneOff));
  r.ba = pack2(dec8<2>(dec, w, laneOff), kOpaque ? 1.0f : decAlpha(w));
  return r;
}

// Un-normalised sum of one 2x2 quad, vertical pairing: (UL + LL) + (UR + LR)  (glsl:180-188).
template <bool kOpaque = false>
__device__ __forceinline__ V4 quadSumV(const unsigned char* dec, uint32_t laneOff, uint32_t ul, uint32_t ur,
                                       uint32_t ll, uint32_t lr)
{
  return add4(add4(decodeTexel<kOpaque>(dec, laneOff, ul), decodeTexel<kOpaque>(dec, laneOff, ll)),
              add4(decodeTexel<kOpaque>(dec, laneOff, ur), decodeTexel<kOpaque>(dec, laneOff, lr)));
}

#if NVPYR_FAST_ENC_ROWS
// Encode of one RGB channel carried as S' = 2^-100 * 4^K * x; the code lands in bits 24..31 (kEncCodeByte).
// encSel = lane * 4: the lane's private column of the row table.  encBytes = table base - kRowEncBias.
constexpr uint32_t kEncCodeByte = 3;
template <int K>
__device__ __forceinline__ uint32_t encScaled(const unsigned char* encBytes, float s, uint32_t encSel = 0)
{
  if(EncConst<K>::kClamp)
    s = fmaxf(s, __uint_as_float(EncConst<K>::kMinBits));  // below the first threshold everything encodes to 0
  const float    z   = __fmaf_rn(s, __uint_as_float(EncConst<K>::kScaleBits), __uint_as_float(kRowEncCBits));  // RN(x + c)
  const uint32_t off = __byte_perm(__float_as_uint(z), encSel, 0x5324);  // key(z) << 8 | lane << 2
  const uint32_t e   = *reinterpret_cast<const uint32_t*>(encBytes + off);
  return e + __float_as_uint(s) + EncConst<K>::kAdd;
}
__device__ __forceinline__ uint32_t encSelOfLane(uint32_t lane) { return lane * 4u; }
__device__ __forceinline__ const unsigned char* encBaseOf(const Srgba8FastSmem& sm)
{
  return reinterpret_cast<const unsigned char*>(sm.decode) - kRowEncBias;
}
#else
constexpr uint32_t kEncCodeByte = 2;
// Encode of one RGB channel carried as S' = 2^-100 * 4^K * x; the code lands in bits 16..23.
// Covered levels (encCovered): every non-zero S' lies inside the extended table, zero reads its dedicated word.
// encWay = (lane & (kEncWays - 1)) * 4: which copy of the entry this lane reads.
template <int K>
__device__ __forceinline__ uint32_t encScaled(const unsigned char* encBytes, float s, uint32_t encWay = 0)
{
  if(EncConst<K>::kClamp)
    s = fmaxf(s, __uint_as_float(EncConst<K>::kMinBits));  // below the first threshold everything encodes to 0
  const uint32_t b = __float_as_uint(s);
  uint32_t       off;
  if(kEncWays == 1)
    off = (b >> (kFastEncShift - 2)) & 0xFFFFFFFCu;
  else
    off = ((b >> (kFastEncShift - kEncStrideLog2)) & ~(kEncStride - 1u)) | encWay;  // key * stride + copy * 4
  const uint32_t e = *reinterpret_cast<const uint32_t*>(encBytes + off - EncConst<K>::kBias);
  return e + b + EncConst<K>::kAdd;
}
__device__ __forceinline__ uint32_t encSelOfLane(uint32_t lane) { return kEncWays == 1 ? 0u : (lane & (kEncWays - 1u)) * 4u; }
__device__ __forceinline__ const unsigned char* encBaseOf(const Srgba8FastSmem& sm)
{
  return reinterpret_cast<const unsigned char*>(sm.encode);
}
#endif
// PRMT selectors that gather code bytes: (own code byte, other's code byte) etc.
constexpr uint32_t kSelCC = kEncCodeByte | ((kEncCodeByte + 4u) << 4);  // byte0 = a.code, byte1 = b.code
constexpr uint32_t kSelCA = kEncCodeByte | (4u << 4);                   // byte0 = a.code, byte1 = b.byte0 (alpha)
// uint(a * 255 + 0.5) for a = S / 4^K (alpha is not pre-scaled); code in bits 0..7 (a <= 1: no clamp).
template <int K>
__device__ __forceinline__ uint32_t encAlphaScaled(float s)
{
  constexpr float kMul = 255.0f / float(1 << (2 * K));  // exact
  const float     v    = __fadd_rn(__fmul_rn(s, kMul), 0.5f);
  return __float_as_uint(__fadd_rz(v, 8388608.0f));  // 2^23 + trunc(v)
}
// kOpaque: the alpha sum is exactly 4^K (every contributing texel has alpha 255), which encodes to 255.
template <int K, bool kOpaque = false>
__device__ __forceinline__ uint32_t encWordScaled(const unsigned char* encBytes, float4 s)
{
  const uint32_t w = encSelOfLane(threadIdx.x & 31u);
  const uint32_t r = encScaled<K>(encBytes, s.x, w), g = encScaled<K>(encBytes, s.y, w), b = encScaled<K>(encBytes, s.z, w);
  const uint32_t a = kOpaque ? 255u : encAlphaScaled<K>(s.w);
  return __byte_perm(__byte_perm(r, g, kSelCC), __byte_perm(b, a, kSelCA), 0x5410);
}

template <int K, bool kOpaque = false>
__device__ __forceinline__ uint32_t encWordScaled(const unsigned char* encBytes, V4 s)
{
  return encWordScaled<K, kOpaque>(encBytes, toFloat4(s));
}

__device__ __forceinline__ V4 sum4Paired(bool horizontal, V4 ul, V4 ur, V4 ll, V4 lr)
{
  const V4 p = horizontal ? add4(ul, ur) : add4(ul, ll);
  const V4 q = horizontal ? add4(ll, lr) : add4(ur, lr);
  return add4(p, q);
}
__device__ __forceinline__ V4 shflXor(V4 a, int mask)
{
  V4 r;
  r.rg.v = __shfl_xor_sync(0xffffffffu, a.rg.v, mask);
  r.ba.v = __shfl_xor_sync(0xffffffffu, a.ba.v, mask);
  return r;
}

// Premultiply-alpha pre-pass of the reference's image loader (include/scoped_image.hpp:233-255), fused into
// the level-0 read: c' = srgbFromLinear(linearFromSrgb(c) * (A * (1/255))) per colour channel, alpha kept.
// Opaque texels are returned unchanged without any work: srgbFromLinear(linearFromSrgb(c)) == c for all 256
// codes (tests/test_oracle_pins.py).  Uses the scaled tables: 2^-100 commutes with the rounding of the product.
__device__ __forceinline__ uint32_t premultiplyWord(const unsigned char* dec, const unsigned char* enc, uint32_t laneOff,
                                                    uint32_t encWay, uint32_t w)
{
  if((w >> 24) == 255u)
    return w;
  const float    a = decAlpha(w);
  const uint32_t r = encScaled<0>(enc, __fmul_rn(dec8<0>(dec, w, laneOff), a), encWay);
  const uint32_t g = encScaled<0>(enc, __fmul_rn(dec8<1>(dec, w, laneOff), a), encWay);
  const uint32_t b = encScaled<0>(enc, __fmul_rn(dec8<2>(dec, w, laneOff), a), encWay);
  // codes sit in byte kEncCodeByte of r, g, b; alpha stays in byte 3 of w
  return __byte_perm(__byte_perm(r, g, kSelCC), __byte_perm(b, w, kEncCodeByte | (7u << 4)), 0x5410);
}
__device__ __forceinline__ bool premultiplyRow(const unsigned char* dec, const unsigned char* enc, uint32_t laneOff,
                                               uint32_t encWay, uint4& c)
{
  const uint4 o = c;
  c.x = premultiplyWord(dec, enc, laneOff, encWay, c.x), c.y = premultiplyWord(dec, enc, laneOff, encWay, c.y);
  c.z = premultiplyWord(dec, enc, laneOff, encWay, c.z), c.w = premultiplyWord(dec, enc, laneOff, encWay, c.w);
  return (o.x != c.x) | (o.y != c.y) | (o.z != c.z) | (o.w != c.w);
}

// Stand-alone premultiply pre-pass over n4 groups of four texels (16-byte aligned in and out; in == out
// allowed), with the conflict-free tables of the fast kernel.  Used where the pre-pass cannot ride in a
// fast step (chains that start with the general pipeline, nvpyrPremultiplyAlpha).
__global__ void __launch_bounds__(kFastWarps * 32, kFastCtasPerSm)
    premultiplySrgba8Kernel(const uint4* in, uint4* out, uint64_t n4, const DeviceTables* tables)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  Srgba8FastSmem& sm = *reinterpret_cast<Srgba8FastSmem*>(smemRaw);
  srgba8FastInit<kFastWarps * 32>(sm, tables);
  __syncthreads();
  gridDependencyWait();
  gridLaunchDependents();
  const unsigned char* dec = reinterpret_cast<const unsigned char*>(sm.decode);
  const unsigned char* enc = encBaseOf(sm);
  const uint32_t       lane = threadIdx.x & 31u, laneOff = lane * 4u;
  const uint32_t       way  = encSelOfLane(lane);
  const uint64_t       step = uint64_t(gridDim.x) * blockDim.x;
  for(uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += 2u * step)
  {
    // two independent groups per trip: more loads in flight
    const bool second = i + step < n4;
    uint4      a = __ldg(in + i), b = second ? __ldg(in + i + step) : make_uint4(0u, 0u, 0u, 0u);
    const bool ca = premultiplyRow(dec, enc, laneOff, way, a);
    if(ca || in != out)
      out[i] = a;
    if(second)
    {
      const bool cb = premultiplyRow(dec, enc, laneOff, way, b);
      if(cb || in != out)
        out[i + step] = b;
    }
  }
}

// Batch mode (nvpyrDispatchBatch on images of one size): ONE launch streams the same step of `count`
// independent packed chains.  p.lv[k].ptr then holds the byte offset of level k inside a chain, bases[i]
// the chain of image i; tile t belongs to image t / tilesPerImage.
struct FastBatch
{
  const unsigned char* const* bases;
  uint32_t                    tilesPerImage, count;
  const CUtensorMap*          maps;  // one tensor map per image (its level 0), in device memory: the batch kernel's TMA staging
};

// kPremul: level 0 holds straight (un-premultiplied) alpha; it is premultiplied on the fly, written back in
// place (only 4x4 blocks that changed) and the chain is generated from the premultiplied codes -- exactly what
// premultiplyKernel followed by this kernel produce, minus one full read + write pass over level 0.
// kSlabTasks (M >= 4; launches of few tiles per resident warp, up to ~8192^2): the unit of work of a warp is ONE
// 64x8 slab instead of a whole 64 x 2^M tile.  A CTA owns tiles c, c + G, ... (local index j); its warps take the
// slabs of those tiles round-robin, drop their row of level +3 sums into the tile's stash (slot j mod 32) and the
// warp that arrives last at the tile (shared-memory counter, nobody waits) finishes levels +4..+M.  A slot is handed
// to local tile j + 32 by a generation counter that the finishing warp bumps (eight warps work on a tile at a time, so
// a slot is free again long before it is needed: the check never spins in practice, it is there for correctness).
// A 1024^2 image has 256 tiles for 4736 resident warps: in tile mode 5 % of the warps walk 8 slabs each, in
// slab mode 43 % walk one; 8192^2 is 3.46 tiles per warp -- the fourth round of whole tiles runs 46 % full -- but 27.7
// slabs.  Same expression trees, same bits.
// kWarps: warps per CTA (one CTA per SM).  32 warps leave 64 registers per thread -- ptxas then spills a tile
// coordinate and re-derives lane constants inside the slab loop -- 24 warps get 80 and need neither: with enough tiles
// to keep every warp busy for several rounds the 24-warp build is 4 % faster (16384^2: 262 -> 251 us), while images
// of about one tile per warp want the 32-warp build's extra tasks in flight (4096^2: 24.1 vs 27.1 us).
template <int M, bool kBatch, bool kPremul, bool kSlabTasks, int kWarps = kFastWarps>
__global__ void __launch_bounds__(kWarps * 32, kFastCtasPerSm)
    fastSrgba8Kernel(const FastParams p, const FastBatch batch, const __grid_constant__ FastTensorMap tmap)
{
  static_assert(kWarps <= kFastWarps, "the shared-memory layout is sized for kFastWarps");
  static_assert(M >= 2 && M <= 6, "2..6 levels");
  static_assert(!kSlabTasks || (M >= 4 && !kBatch), "slab tasks: single image, levels beyond +3");
  constexpr uint32_t kTileH = M >= 3 ? (1u << M) : 8u, kSlabs = kTileH / 8u;
  constexpr int      kSlabUnroll = kPremul ? 1 : kFastSlabUnroll;
  constexpr bool     kOpaquePath = NVPYR_FAST_OPAQUE_PATH != 0 && !kPremul;  // (a premultiplying launch has translucent texels by definition)
  constexpr bool     kPinPrefetch = NVPYR_FAST_PIN_PREFETCH != 0 && kSlabUnroll > 1 && kSlabs > 1;
  extern __shared__ __align__(128) unsigned char smemRaw[];  // the TMA ring inside needs 128-byte alignment
  constexpr uint32_t kSlots = kFastWarps;          // stash slots of the slab-task mode
  __shared__ uint32_t tileArrivals[kSlots];        // slab tasks: slabs of the slot's current tile that have arrived
  __shared__ uint32_t slotGeneration[kSlots];      // slab tasks: tiles the slot has seen completed
  Srgba8FastSmem& sm = *reinterpret_cast<Srgba8FastSmem*>(smemRaw);
  if(kSlabTasks && threadIdx.x < kSlots)
    tileArrivals[threadIdx.x] = 0u, slotGeneration[threadIdx.x] = 0u;
  srgba8FastInit<kWarps * 32>(sm, p.tables);
  __syncthreads();  // the only CTA-wide barrier
  // (the wait for the previous kernel comes right before the first access to a level, below: the lane constants and
  // the mbarrier set-up need nothing the previous kernel wrote)
  const unsigned char* dec = reinterpret_cast<const unsigned char*>(sm.decode);
  const unsigned char* enc = encBaseOf(sm);

  // (the shuffle tells the compiler that the warp index -- and with it the tile bookkeeping -- is warp-uniform)
  const uint32_t lane = threadIdx.x & 31u, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t tx = lane & 15u, ty = lane >> 4;  // lane = 4x4 texels at (4 tx, 4 ty) of a 64x8 slab
  const uint32_t laneOff = lane * 4u;
  const uint32_t W = p.lv[0].w, H = p.lv[0].h;
  const uint32_t numTiles = p.tilesX * p.tilesY * (kBatch ? batch.count : 1u);
  // base address of level k of the image that tile t belongs to
  auto levelPtr = [&](int k, uint32_t t) -> unsigned char* {
    if(!kBatch)
      return p.lv[k].ptr;
    return reinterpret_cast<unsigned char*>(__ldg(reinterpret_cast<const unsigned long long*>(batch.bases) + t / batch.tilesPerImage)) + reinterpret_cast<size_t>(p.lv[k].ptr);
  };
  const size_t   pitch0 = p.lv[0].pitch, pitch1 = p.lv[1].pitch, pitch2 = p.lv[2].pitch;
  float*         myL3 = stashOf(sm, warp);

  // level +3 transpose-reduce roles
  const bool     xOdd = lane & 1u, yOdd = lane & 16u;
  const uint32_t ch   = (xOdd ? 2u : 0u) + (yOdd ? 1u : 0u);  // channel this lane ends up owning
  // gather selectors: step 1 (xor 16) builds the 2-byte pair of this column, step 2 (xor 1) the word
  //   x even: pair = (R from y-even lane, G from y-odd lane), codes in byte 2 of both
  //   x odd : pair = (B from y-even lane (byte 2), A from y-odd lane (byte 0))
  constexpr uint32_t kOwn = kEncCodeByte, kOther = kEncCodeByte + 4u;  // PRMT indices of the two code bytes
  const uint32_t sel1 = xOdd ? (yOdd ? (kOther | 0u << 4) : (kOwn | 4u << 4)) : (yOdd ? (kOther | kOwn << 4) : (kOwn | kOther << 4));
  const uint32_t sel2 = xOdd ? 0x1054u : 0x5410u;

  // Tiles are dealt CTA-major so that mid-size images still spread over every SM.
  const uint32_t tileStep = gridDim.x * kWarps;
  uint32_t       task     = warp;  // slab tasks: local task index = local tile * kSlabs + slab
  uint32_t       tile     = kSlabTasks ? blockIdx.x + gridDim.x * (task / kSlabs) : blockIdx.x + gridDim.x * warp;
  uint32_t       slab0    = kSlabTasks ? task % kSlabs : 0u;  // first slab of this warp's current unit of work

  // Per-lane cursor of the slab being prefetched: source pointer + "inside the image".  Edges are multiples of
  // 2^M and a tile is 2^M rows tall, so a lane is inside or outside for a whole tile (outside: a narrow level
  // whose last tile is cut in x, or no tile left).  Outside lanes read the first rows of the image instead
  // -- the loads are unconditional, which spares the compiler a second copy of the 16 prefetch registers.
  struct Cursor
  {
    const unsigned char* src;
    bool                 active;
  };
  auto tileCursor = [&](uint32_t t, uint32_t firstSlab) {
    const uint32_t tt = kBatch ? t % batch.tilesPerImage : t;  // tile inside its image
    const uint32_t x0 = (tt % p.tilesX) * 64u + tx * 4u, y0 = (tt / p.tilesX) * kTileH + firstSlab * 8u + ty * 4u;
    Cursor         c;
    c.active = t < numTiles && x0 < W && y0 < H;  // (y0 >= H: the 8-row tiles of a 2-level step on H % 8 == 4)
    const unsigned char* base = levelPtr(0, kBatch && t >= numTiles ? 0u : t);
    c.src    = c.active || !kFastUncondLoads ? base + size_t(y0) * pitch0 + size_t(x0) * 4u : base;
    return c;
  };
  uint4 row[4];
  auto  loadRows = [&](const Cursor& c) {
    if(kFastUncondLoads || c.active)
    {
#pragma unroll
      for(int i = 0; i < 4; ++i)
        row[i] = __ldg(reinterpret_cast<const uint4*>(c.src + size_t(i) * pitch0));
    }
  };
  // TMA variant: this warp's staging buffer, its mbarrier and the phase the next wait expects
  // (slab tasks keep the register path: a warp runs one or two tasks there, so there is little to overlap, and the
  // first copy of a CTA would pay the tensor-map fetch on top of the DRAM latency; TMA staging brought no gain outside
  // the box-to-box noise of the 2048^2 config when tried in round 2)
  constexpr bool kTma = kFastTma && !kPremul && !kSlabTasks;  // (batches: one tensor map per image, FastBatch::maps)
  uint32_t       tmaBar = 0u, tmaRing = 0u, tmaPhase = 0u;
  // Issues the row copies of slab s of tile t (rows cut at the image edges; nothing for a tile that does not exist).
  auto tmaIssue = [&](uint32_t t, uint32_t s) {
    if(t >= numTiles)
      return;
    const uint32_t tt = kBatch ? t % batch.tilesPerImage : t;  // tile inside its image
    const uint32_t xt = (tt % p.tilesX) * 64u, ys = (tt / p.tilesX) * kTileH + s * 8u;
#if NVPYR_FAST_TMA == 2
    const CUtensorMap* map = kBatch ? batch.maps + t / batch.tilesPerImage : &tmap.map;
    if(lane == 0u)
    {
      mbarExpectTx(tmaBar, 2048u);  // the whole box, zero-filled outside the image
      tensorCopy2D(tmaRing, map, xt, ys, tmaBar);
    }
#else
    const uint32_t rowBytes = min(64u, W - xt) * 4u, rows = min(8u, H - ys);
    if(lane == 0u)
      mbarExpectTx(tmaBar, rowBytes * rows);
    __syncwarp();
    if(lane < rows)
      bulkCopyG2S(tmaRing + lane * 256u, p.lv[0].ptr + size_t(ys + lane) * pitch0 + size_t(xt) * 4u, rowBytes, tmaBar);
#endif
  };
  if(kTma)
  {
#if NVPYR_FAST_TMA
    tmaBar  = smemAddr(&sm.tmaBar[warp]);
    tmaRing = smemAddr(&sm.ring[warp][0]);
#endif
    if(lane == 0u)
      mbarInit(tmaBar, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fenceProxyAsync();
    __syncwarp();
  }
  gridDependencyWait();    // the previous kernel's levels are complete and visible
  gridLaunchDependents();  // the next kernel may start its own set-up as SMs become free
  Cursor nxt = tileCursor(tile, slab0);
  if(kTma)
    tmaIssue(tile, slab0);
  else if(kFastPrefetch)
    loadRows(nxt);

  while(tile < numTiles)
  {
    // this unit of work, and the one after it (for the prefetch)
    const uint32_t nextTask  = task + kWarps;
    const uint32_t nextSlab0 = kSlabTasks ? nextTask % kSlabs : 0u;
    uint32_t       nextTileI = blockIdx.x + gridDim.x * (nextTask / kSlabs), fetched = 0;
    if(!kSlabTasks && !kDynTiles)
      nextTileI = tile + tileStep;
    if(!kSlabTasks && kDynTiles)
    {
      // Tile mode: the first tile of every warp is dealt statically (CTA-major), all further ones are handed out
      // through one global counter, so that a CTA that starts late -- its SM was still busy with the previous
      // chain's tail kernel -- simply takes fewer tiles instead of finishing late.  One atomic per 64 x 2^M tile,
      // fetched a whole tile ahead of its use.  Every processed tile fetches exactly once, so the fetch that
      // returns numTiles - 1 is the last of the launch: it puts the counter back to zero for the next one.
      // (lane 0 issues the atomic here; the warp only looks at the answer when it reaches the tile's last slab)
      if(lane == 0u)
      {
        fetched = atomicAdd(p.tileCounter, 1u);
        if(fetched == numTiles - 1u)
          *p.tileCounter = 0u;
      }
    }
    // The next unit of work becomes known (and its cursor is computed) at the last slab of this one.
    auto resolveNext = [&]() {
      if(!kSlabTasks && kDynTiles)
        nextTileI = __shfl_sync(0xffffffffu, fetched, 0) + tileStep;
      return tileCursor(nextTileI, nextSlab0);
    };
    const uint32_t slot = kSlabTasks ? (task / kSlabs) % kSlots : 0u, generation = kSlabTasks ? (task / kSlabs) / kSlots : 0u;
    if(kSlabTasks)
    {
      myL3 = stashOf(sm, slot);
      // the slot's previous tile (local tile - kSlots) must have been finished before its stash is overwritten
      if(lane == 0u)
        while(atomicAdd(&slotGeneration[slot], 0u) != generation)  // (an atomic read: the flag is written atomically too)
          ;
      __syncwarp();
    }
    const uint32_t tileInImage = kBatch ? tile % batch.tilesPerImage : tile;
    const uint32_t tileX = tileInImage % p.tilesX, tileY = tileInImage / p.tilesX;
    const uint32_t x0 = tileX * 64u + tx * 4u;
    uint32_t       y0 = tileY * kTileH + slab0 * 8u + ty * 4u;
    // Output cursors of this lane (advance by one slab = 8 input rows per iteration).
    unsigned char* d1 = levelPtr(1, tile) + size_t(y0 >> 1) * pitch1 + size_t(x0 >> 1) * 4u;
    unsigned char* d2 = levelPtr(2, tile) + size_t(y0 >> 2) * pitch2 + size_t(x0 >> 2) * 4u;
    unsigned char* d3 = M >= 3 ? levelPtr(3, tile) + size_t(y0 >> 3) * p.lv[3].pitch + size_t(x0 >> 3) * 4u : nullptr;
    const uint32_t slabEnd  = kSlabTasks ? slab0 + 1u : kSlabs;
#pragma unroll kSlabUnroll
    for(uint32_t slab = slab0; slab < slabEnd; ++slab, y0 += 8u, d1 += 4u * pitch1, d2 += 2u * pitch2)
    {
      // Edges are multiples of 2^M >= 4: a 4x4 block is entirely inside or outside.
      const bool active = kTma ? (x0 < W && y0 < H) : nxt.active;
      if(!kTma && !kFastPrefetch)
        loadRows(nxt);  // no software prefetch: more resident warps hide the latency instead
      uint4 c0 = row[0], c1 = row[1], c2 = row[2], c3 = row[3];
      if(kTma)
      {
        // wait for this slab, read the lane's four 16-byte rows, hand the buffer back to the TMA unit for the next slab
        mbarWait(tmaBar, tmaPhase);
        tmaPhase ^= 1u;
        const uint32_t mine = tmaRing + ty * 1024u + tx * 16u;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c0.x), "=r"(c0.y), "=r"(c0.z), "=r"(c0.w) : "r"(mine));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+256];" : "=r"(c1.x), "=r"(c1.y), "=r"(c1.z), "=r"(c1.w) : "r"(mine));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+512];" : "=r"(c2.x), "=r"(c2.y), "=r"(c2.z), "=r"(c2.w) : "r"(mine));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+768];" : "=r"(c3.x), "=r"(c3.y), "=r"(c3.z), "=r"(c3.w) : "r"(mine));
        __syncwarp();
        fenceProxyAsync();  // the generic-proxy reads above are ordered before the async-proxy writes below
        if(kSlabTasks)
          tmaIssue(nextTileI, nextSlab0);
        else if(slab + 1u < kSlabs)
          tmaIssue(tile, slab + 1u);
        else
        {
          resolveNext();
          tmaIssue(nextTileI, 0u);
        }
      }
      unsigned char* const curSrc = const_cast<unsigned char*>(nxt.src);
      // the next slab (of this tile, or the first one of this warp's next tile)
      if(kTma)
      {
      }
      else if(!kSlabTasks && slab + 1u < kSlabs)
      {
        nxt.src += 8u * pitch0;  // (with unconditional loads an outside lane walks down the first 2^M rows of the image)
        if(!kFastUncondLoads)
          nxt.active = x0 < W && y0 + 8u < H;
      }
      else
        nxt = resolveNext();
      if(!kTma && kFastPrefetch)
      {
        if(kPinPrefetch)
        {
          // ptxas tracks every LDG of this loop with ONE scoreboard.  In the unrolled loop the prefetch of slab s+2
          // would be issued just before the first use of slab s+1's registers, and that use would then wait for
          // both.  The address below depends (by a bit that is always 0: decoded values are never negative) on the
          // first decode of this slab, so the new loads are issued after the wait for the previous ones.
          Cursor pinned = nxt;
          pinned.src += __float_as_uint(dec8<0>(dec, c0.x, laneOff)) >> 31;
          loadRows(pinned);
        }
        else
          loadRows(nxt);
      }

      if(kPremul && active)
      {
        const uint32_t way = encSelOfLane(lane);
        bool           changed = premultiplyRow(dec, enc, laneOff, way, c0);
        changed |= premultiplyRow(dec, enc, laneOff, way, c1);
        changed |= premultiplyRow(dec, enc, laneOff, way, c2);
        changed |= premultiplyRow(dec, enc, laneOff, way, c3);
        if(changed)
        {
          *reinterpret_cast<uint4*>(curSrc)              = c0;
          *reinterpret_cast<uint4*>(curSrc + pitch0)      = c1;
          *reinterpret_cast<uint4*>(curSrc + 2u * pitch0) = c2;
          *reinterpret_cast<uint4*>(curSrc + 3u * pitch0) = c3;
        }
      }
      V4 s2 = toV4(make_float4(0.f, 0.f, 0.f, 0.f));
      // level +1 (K = 1): four quads, vertical pairing inside each; level +2 (K = 2) from the
      // thread's own 2x2.  The quads are visited in the order of the level +2 pairing so that
      // only one partial sum stays live (register pressure).
      auto levels12 = [&](auto opaqueTag) {
        constexpr bool kOpaque = decltype(opaqueTag)::value;
        if(fastPairingIsHorizontal(2, M))
        {
          const V4 s00 = quadSumV<kOpaque>(dec, laneOff, c0.x, c0.y, c1.x, c1.y);
          const V4 s01 = quadSumV<kOpaque>(dec, laneOff, c0.z, c0.w, c1.z, c1.w);
          *reinterpret_cast<uint2*>(d1) = make_uint2(encWordScaled<1, kOpaque>(enc, s00), encWordScaled<1, kOpaque>(enc, s01));
          const V4 top = add4(s00, s01);
          const V4 s10 = quadSumV<kOpaque>(dec, laneOff, c2.x, c2.y, c3.x, c3.y);
          const V4 s11 = quadSumV<kOpaque>(dec, laneOff, c2.z, c2.w, c3.z, c3.w);
          *reinterpret_cast<uint2*>(d1 + pitch1) = make_uint2(encWordScaled<1, kOpaque>(enc, s10), encWordScaled<1, kOpaque>(enc, s11));
          s2 = add4(top, add4(s10, s11));  // (UL + UR) + (LL + LR)
        }
        else
        {
          const V4       s00 = quadSumV<kOpaque>(dec, laneOff, c0.x, c0.y, c1.x, c1.y);
          const uint32_t e00 = encWordScaled<1, kOpaque>(enc, s00);
          const V4       s10 = quadSumV<kOpaque>(dec, laneOff, c2.x, c2.y, c3.x, c3.y);
          const uint32_t e10 = encWordScaled<1, kOpaque>(enc, s10);
          const V4       left = add4(s00, s10);
          const V4       s01 = quadSumV<kOpaque>(dec, laneOff, c0.z, c0.w, c1.z, c1.w);
          *reinterpret_cast<uint2*>(d1) = make_uint2(e00, encWordScaled<1, kOpaque>(enc, s01));
          const V4 s11 = quadSumV<kOpaque>(dec, laneOff, c2.z, c2.w, c3.z, c3.w);
          *reinterpret_cast<uint2*>(d1 + pitch1) = make_uint2(e10, encWordScaled<1, kOpaque>(enc, s11));
          s2 = add4(left, add4(s01, s11));  // (UL + LL) + (UR + LR)
        }
        *reinterpret_cast<uint32_t*>(d2) = encWordScaled<2, kOpaque>(enc, s2);
      };
      // Opaque fast path (NVPYR_FAST_OPAQUE_PATH, off by default: measured slower): when every texel of the warp's slab
      // has alpha 255 the 16 alpha conversions (I2F + FMUL) and the 5 alpha encodes of levels +1 and +2 are constants
      // (255 * (1/255) rounds to exactly 1.0f, sums of ones are exact).  Same bits by construction.
      bool opaqueSlab = false;
      if(kOpaquePath)
      {
        const uint32_t am = (c0.x & c0.y & c0.z) & (c0.w & c1.x & c1.y) & (c1.z & c1.w & c2.x) & (c2.y & c2.z & c2.w) & (c3.x & c3.y & c3.z) & c3.w;
        opaqueSlab        = __all_sync(0xffffffffu, !active || am >= 0xFF000000u);
      }
      if(active)
      {
        if(opaqueSlab)
          levels12(std::true_type{});
        else
          levels12(std::false_type{});
      }

      if(M >= 3)
      {
        // level +3 (K = 3), horizontal pairing for every M: (self + x) + (y + xy).
        // Step x: even lanes keep (R, G) and send (B, A); odd lanes the other way round.
        const F2 keep = xOdd ? s2.ba : s2.rg, send = xOdd ? s2.rg : s2.ba;
        F2       got;
        got.v      = __shfl_xor_sync(0xffffffffu, send.v, 1);
        const F2 t = add2(keep, got);
        float    t0, t1;
        unpack2(t, t0, t1);
        // Step y: y-even lanes keep the first of their two channels, y-odd lanes the second.
        const float u = __fadd_rn(yOdd ? t1 : t0, __shfl_xor_sync(0xffffffffu, yOdd ? t0 : t1, 16));
        // encode this lane's channel, gather the four bytes
        const uint32_t code =
            ch == 3u ? encAlphaScaled<3>(u) : encScaled<3>(enc, u, encSelOfLane(lane));
        const uint32_t pair = __byte_perm(code, __shfl_xor_sync(0xffffffffu, code, 16), sel1);
        const uint32_t word = __byte_perm(pair, __shfl_xor_sync(0xffffffffu, pair, 1), sel2);
        if(active)
        {
          if(ch == 0u)
            *reinterpret_cast<uint32_t*>(d3) = word;
          if(M >= 4)
            myL3[slab * kL3RowFloats + (tx >> 1) * 4u + ch] = u;
        }
        d3 += p.lv[3].pitch;
      }
    }

    bool finishTile = M >= 4;
    if(M >= 4 && kSlabTasks)
    {
      // Publish this slab's row of level +3 sums; the warp that completes the tile finishes it.
      __syncwarp();
      __threadfence_block();
      uint32_t arrived = 0;
      if(lane == 0u)
        arrived = atomicAdd(&tileArrivals[slot], 1u);
      arrived    = __shfl_sync(0xffffffffu, arrived, 0);
      finishTile = arrived == kSlabs - 1u;
      if(finishTile)
        __threadfence_block();
    }
    if(finishTile)
    {
      __syncwarp();
      // 16 lanes <-> 4 x 4 texels of level +4; +5 and +6 with butterflies.
      const uint32_t i = lane & 3u, j = (lane >> 2) & 3u;
      const uint32_t ox = tileX * 64u + i * 16u, oy = tileY * kTileH + j * 16u;  // origin in the input level
      const bool     valid = lane < 16u && j * 16u < kTileH && ox < W && oy < H;
      V4             s4    = toV4(make_float4(0.f, 0.f, 0.f, 0.f));
      if(valid)
      {
        const float4* l3 = reinterpret_cast<const float4*>(myL3);
        constexpr uint32_t kRow4 = kL3RowFloats / 4u;  // float4s per stash row
        const V4      ul = toV4(l3[(2 * j) * kRow4 + 2 * i]), ur = toV4(l3[(2 * j) * kRow4 + 2 * i + 1]);
        const V4      ll = toV4(l3[(2 * j + 1) * kRow4 + 2 * i]), lr = toV4(l3[(2 * j + 1) * kRow4 + 2 * i + 1]);
        s4               = sum4Paired(fastPairingIsHorizontal(4, M), ul, ur, ll, lr);
        *reinterpret_cast<uint32_t*>(levelPtr(4, tile) + size_t(oy >> 4) * p.lv[4].pitch + size_t(ox >> 4) * 4u) =
            encWordScaled<4>(enc, s4);
      }
      if(M >= 5)
      {
        // vertical pairing (shared-memory tail of the reference): (self + y) + (x + xy)
        const V4 a  = add4(s4, shflXor(s4, 4));
        const V4 s5 = add4(a, shflXor(a, 1));
        if(valid && !(i & 1u) && !(j & 1u))
          *reinterpret_cast<uint32_t*>(levelPtr(5, tile) + size_t(oy >> 5) * p.lv[5].pitch + size_t(ox >> 5) * 4u) =
              encWordScaled<5>(enc, s5);
        if(M >= 6)
        {
          const V4 c  = add4(s5, shflXor(s5, 8));
          const V4 s6 = add4(c, shflXor(c, 2));
          if(valid && lane == 0u)
            *reinterpret_cast<uint32_t*>(levelPtr(6, tile) + size_t(oy >> 6) * p.lv[6].pitch + size_t(ox >> 6) * 4u) =
                encWordScaled<6>(enc, s6);
        }
      }
      __syncwarp();  // the warp's level +3 tile is free again
      if(kSlabTasks && lane == 0u)
      {
        // hand the slot to local tile + kSlots: counter back to zero, then the generation (in this order)
        atomicExch(&tileArrivals[slot], 0u);
        __threadfence_block();
        atomicExch(&slotGeneration[slot], generation + 1u);
      }
    }
    // next unit of work
    if(kSlabTasks)
    {
      task  = nextTask;
      tile  = nextTileI;
      slab0 = nextSlab0;
    }
    else
      tile = nextTileI;
  }
}

}  // namespace nvpyr
