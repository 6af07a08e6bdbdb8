// nvpyr_kernels.cuh -- sm_100a kernels of the mip-pyramid generator.
//
// Two pipelines, like the reference (nvpro_pyramid/nvpro_pyramid.glsl):
//
//  fastKernel<F, M>     M = 2..6 levels from one input level whose edges are
//                       multiples of 2^M (glsl:216-532).  One CTA owns a 64x64
//                       input tile (a partial one at the right/bottom edge); a
//                       thread owns a 4x4 block: levels +1 and +2 in registers,
//                       +3 with warp shuffles, +4..+6 by one warp after a single
//                       barrier through a 1 KB shared tile.  Persistent grid.
//  fastKernel1<F>       M = 1 (only at the end of a chain).
//  generalKernel<F>     1 or 2 levels from an input level of ANY size with the
//                       energy-conserving 1/2/3-tap separable kernel
//                       (glsl:543-882); level +1 is carried to level +2 as
//                       float32 through shared memory.
//
// The thread<->texel mapping is ours (row-major 16x16 threads, coalesced 16-byte
// loads) -- NOT the reference's Morton order -- but every output texel is produced
// by the same float32 expression tree as in the reference shader, including which
// neighbours share a bracket in 0.25*((a+b)+(c+d)) at each level ("pairing"), so
// the stored bits are identical to the shader-order oracle.
#pragma once
#include <type_traits>
#include <utility>

#include "nvpyr_functors.cuh"

namespace nvpyr {

struct LevelView
{
  unsigned char* ptr;
  uint32_t       pitch;  // bytes
  uint32_t       w, h;
  uint32_t       level;  // index of the level in its chain (handed to user hooks; the kernels do not use it)
};

// NVPRO_PYRAMID_LOAD_REDUCE4 (nvpro_pyramid.glsl:78-88, default :179-189): a functor set MAY define
//   static Value loadReduce4(const Shared&, const void* texel00, size_t rowPitchBytes, uint32_t x, uint32_t y, uint32_t level)
// = "load the 2x2 square whose upper-left texel is (x, y) of mip level `level` (at texel00) and reduce it".  Like the
// macro it is used by the fast pipeline only, for the first level of every dispatch, and replaces load + reduce4
// there (a user set may fetch through a texture object in its Params to use the hardware's bilinear filter).
template <class F, class = void>
struct HasLoadReduce4 : std::false_type
{
};
template <class F>
struct HasLoadReduce4<F, std::void_t<decltype(F::loadReduce4(std::declval<const typename F::Shared&>(), static_cast<const void*>(nullptr),
                                                             size_t(0), 0u, 0u, 0u))>> : std::true_type
{
};

// Pairing of the 2x2 reduction that produces level (input + k) in an M-level fast
// step.  false = "vertical": (UL+LL)+(UR+LR); true = "horizontal": (UL+UR)+(LL+LR).
//   k = 1            software LOAD_REDUCE4, glsl:180-188            vertical
//   k = 2, M = 6     the thread's own 2x2 of level +1, glsl:321-342 vertical
//   next two levels  subgroup shuffles xor 1,2,3 / 4,8,12, glsl:366-391  horizontal
//   last one or two  shared-memory tail, glsl:487-491,:509-526      vertical
__host__ __device__ constexpr bool fastPairingIsHorizontal(int k, int M)
{
  return M == 6 ? (k == 3 || k == 4) : (k == 2 || k == 3);
}

template <class F>
__device__ __forceinline__ typename F::Value reduce4Paired(bool horizontal, typename F::Value ul,
                                                           typename F::Value ur, typename F::Value ll,
                                                           typename F::Value lr)
{
  return horizontal ? F::reduce4(ul, ur, ll, lr) : F::reduce4(ul, ll, ur, lr);
}

__device__ __forceinline__ float4 shflXor(float4 v, int mask)
{
  return make_float4(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask),
                     __shfl_xor_sync(0xffffffffu, v.z, mask), __shfl_xor_sync(0xffffffffu, v.w, mask));
}
// NVPRO_PYRAMID_SHUFFLE_XOR (glsl:191-193) for any value type made of 32-bit words (user functor sets).
template <class V>
__device__ __forceinline__ V shflXor(V v, int mask)
{
  static_assert(sizeof(V) % 4 == 0 && sizeof(V) <= 64, "value types are made of 1..16 32-bit words");
  union
  {
    V        v;
    uint32_t w[sizeof(V) / 4];
  } a, b;
  a.v = v;
#pragma unroll
  for(int i = 0; i < int(sizeof(V) / 4); ++i)
    b.w[i] = __shfl_xor_sync(0xffffffffu, a.w[i], mask);
  return b.v;
}

struct FastParams
{
  LevelView           lv[7];  // lv[0] = input level, lv[k] = input + k
  uint32_t            tilesX, tilesY;
  const DeviceTables* tables;
  uint32_t*           tileCounter;  // tuned sRGBA8 kernel: dynamic tile hand-out (zero on entry, zero again on exit)
};

template <class F>
struct FastSmem
{
  typename F::Shared tables;
  typename F::Value  l3[2][8][8];  // level +3 of the current tile, double buffered
};

// Tile loop of an M-level fast step, executed by a whole 256-thread CTA: tiles firstTile,
// firstTile + tileStride, ...  `tables` is the functor set's shared-memory state, `l3` a
// [2][8][8] float4 scratch.  Contains CTA barriers: every thread of the CTA must call it with
// the same arguments.
// kVec: all row starts of lv[0] are 16-byte aligned and those of lv[1] are aligned for a
// 2-texel store.
template <class F, int M, bool kVec>
__device__ __forceinline__ void fastTileLoop(const FastParams& p, const typename F::Shared& tables,
                                             typename F::Value (*l3buf)[8][8], uint32_t firstTile,
                                             uint32_t tileStride, uint32_t deferredWait = 0u)
{
  static_assert(M >= 2 && M <= 6, "fastTileLoop handles 2..6 levels");
  using V = typename F::Value;
  const uint32_t tid = threadIdx.x, tx = tid & 15u, ty = tid >> 4;
  const uint32_t W = p.lv[0].w, H = p.lv[0].h;
  const uint32_t numTiles = p.tilesX * p.tilesY;
  uint32_t       parity   = 0;

  for(uint32_t tile = firstTile; tile < numTiles; tile += tileStride, parity ^= 1u)
  {
    const uint32_t tileX = tile % p.tilesX, tileY = tile / p.tilesX;
    const uint32_t x0 = tileX * 64u + tx * 4u, y0 = tileY * 64u + ty * 4u;
    // Edges are multiples of 2^M >= 4: a 4x4 block is entirely inside or outside.  A CTA may have more than
    // 256 threads (tailKernel): the extra ones only take part in the barriers.
    const bool active = tid < 256u && x0 < W && y0 < H;

    // deferredWait (tailKernel's grid step): the wait for the previous kernel comes HERE, right before the first load
    // of a level it wrote -- the parameter reads, the tile arithmetic and the instruction fetches up to this point
    // (~800 cycles, cold) then overlap the previous kernel's last CTAs instead of following them.
    if(deferredWait)
    {
      gridDependencyWait();
      if(deferredWait == 1u)
        gridLaunchDependents();
      deferredWait = 0u;
    }
    V l1[2][2];
    V l2{};
    if(active)
    {
#pragma unroll
      for(int qy = 0; qy < 2; ++qy)
      {
        V                    a[4], b[4];
        const unsigned char* r0 = p.lv[0].ptr + size_t(y0 + 2 * qy) * p.lv[0].pitch + size_t(x0) * F::kTexelBytes;
        const unsigned char* r1 = r0 + p.lv[0].pitch;
        if constexpr(HasLoadReduce4<F>::value && !kVec)
        {
          // the set's own NVPRO_PYRAMID_LOAD_REDUCE4
          l1[qy][0] = F::loadReduce4(tables, r0, p.lv[0].pitch, x0, y0 + 2 * qy, p.lv[0].level);
          l1[qy][1] = F::loadReduce4(tables, r0 + 2 * F::kTexelBytes, p.lv[0].pitch, x0 + 2u, y0 + 2 * qy, p.lv[0].level);
        }
        else
        {
          if(kVec)
          {
            F::load4(tables, r0, a);
            F::load4(tables, r1, b);
          }
          else
          {
#pragma unroll
            for(int i = 0; i < 4; ++i)
            {
              a[i] = F::load(tables, r0 + i * F::kTexelBytes);
              b[i] = F::load(tables, r1 + i * F::kTexelBytes);
            }
          }
          // level +1: vertical pairing (k = 1)
          l1[qy][0] = F::reduce4(a[0], b[0], a[1], b[1]);
          l1[qy][1] = F::reduce4(a[2], b[2], a[3], b[3]);
        }
        unsigned char* d = p.lv[1].ptr + size_t((y0 >> 1) + qy) * p.lv[1].pitch + size_t(x0 >> 1) * F::kTexelBytes;
        if(kVec)
          F::template store2<false>(tables, d, l1[qy][0], l1[qy][1]);
        else
        {
          F::template store<false>(tables, d, l1[qy][0]);
          F::template store<false>(tables, d + F::kTexelBytes, l1[qy][1]);
        }
      }
      // level +2 from the thread's own 2x2 of level +1
      l2 = reduce4Paired<F>(fastPairingIsHorizontal(2, M), l1[0][0], l1[0][1], l1[1][0], l1[1][1]);
      F::template store<false>(tables,
                               p.lv[2].ptr + size_t(y0 >> 2) * p.lv[2].pitch + size_t(x0 >> 2) * F::kTexelBytes, l2);
    }

    if(M >= 3)
    {
      // level +3: 2x2 threads = lanes l, l^1 (x), l^16 (y), l^17.  All four compute the
      // same bits (float add is commutative), the even/even one stores.
      const V sx = shflXor(l2, 1), sy = shflXor(l2, 16), sxy = shflXor(l2, 17);
      const V l3 = reduce4Paired<F>(fastPairingIsHorizontal(3, M), l2, sx, sy, sxy);
      if(active && !(tx & 1u) && !(ty & 1u))
      {
        F::template store<false>(
            tables, p.lv[3].ptr + size_t(y0 >> 3) * p.lv[3].pitch + size_t(x0 >> 3) * F::kTexelBytes, l3);
        // The reference hands the last shuffle-made level to its tail through sharedTile_ (glsl:395): that is
        // level +3 for M = 4, 5 and level +4 for M = 6 (where +3 travels by shuffle, full precision).
        if(M >= 4)
          l3buf[parity][ty >> 1][tx >> 1] = M == 6 ? l3 : F::sharedRound(l3);
      }
    }

    if(M >= 4)
    {
      __syncthreads();
      if(tid < 32)
      {
        // 16 lanes <-> 4x4 texels of level +4; then +5 (2x2) and +6 (1) by shuffles.
        const uint32_t i = tid & 3u, j = (tid >> 2) & 3u;
        const uint32_t ox = tileX * 64u + i * 16u, oy = tileY * 64u + j * 16u;  // origin in level 0
        const bool     valid = tid < 16 && ox < W && oy < H;
        V              l4{};
        if(valid)
        {
          const V ul = l3buf[parity][2 * j][2 * i], ur = l3buf[parity][2 * j][2 * i + 1];
          const V ll = l3buf[parity][2 * j + 1][2 * i], lr = l3buf[parity][2 * j + 1][2 * i + 1];
          l4         = reduce4Paired<F>(fastPairingIsHorizontal(4, M), ul, ur, ll, lr);
          F::template store<false>(
              tables, p.lv[4].ptr + size_t(oy >> 4) * p.lv[4].pitch + size_t(ox >> 4) * F::kTexelBytes, l4);
        }
        if(M >= 5)
        {
          if(M == 6)
            l4 = F::sharedRound(l4);  // sharedTile_ of the 6-level step holds level +4
          const V sx = shflXor(l4, 1), sy = shflXor(l4, 4), sxy = shflXor(l4, 5);
          const V l5 = reduce4Paired<F>(fastPairingIsHorizontal(5, M), l4, sx, sy, sxy);
          if(valid && !(i & 1u) && !(j & 1u))
            F::template store<false>(
                tables, p.lv[5].ptr + size_t(oy >> 5) * p.lv[5].pitch + size_t(ox >> 5) * F::kTexelBytes, l5);
          if(M >= 6)
          {
            const V tx2 = shflXor(l5, 2), ty2 = shflXor(l5, 8), txy2 = shflXor(l5, 10);
            const V l6  = reduce4Paired<F>(fastPairingIsHorizontal(6, M), l5, tx2, ty2, txy2);
            if(valid && tid == 0)
              F::template store<false>(
                  tables, p.lv[6].ptr + size_t(oy >> 6) * p.lv[6].pitch + size_t(ox >> 6) * F::kTexelBytes, l6);
          }
        }
      }
      // l3buf is double buffered: the next iteration writes the other half, and the
      // barrier of that iteration orders this read before the overwrite after it.
    }
  }
  if(deferredWait)  // (a CTA without a tile)
  {
    gridDependencyWait();
    if(deferredWait == 1u)
      gridLaunchDependents();
  }
}

template <class F, int M, bool kVec>
__global__ void __launch_bounds__(256) fastKernel(const FastParams p)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  FastSmem<F>& sm = *reinterpret_cast<FastSmem<F>*>(smemRaw);
  F::sharedInit(sm.tables, p.tables);
  __syncthreads();
  // (the tile loop waits for the previous kernel right before its first load and lets the next one go: deferredWait 1)
  fastTileLoop<F, M, kVec>(p, sm.tables, sm.l3, blockIdx.x, gridDim.x, 1u);
}

// M = 1 (glsl:345-357 with levelCount_ == 1): one thread per output texel, grid-strided from
// `first` with stride `stride` (in threads).
template <class F>
__device__ __forceinline__ void fastLoop1(const FastParams& p, const typename F::Shared& tables, uint64_t first,
                                          uint64_t stride)
{
  using V = typename F::Value;
  const uint32_t W1 = p.lv[1].w, H1 = p.lv[1].h;
  const uint64_t n = uint64_t(W1) * H1;
  for(uint64_t g = first; g < n; g += stride)
  {
    const uint32_t       x = uint32_t(g % W1), y = uint32_t(g / W1);
    const unsigned char* s = p.lv[0].ptr + size_t(2 * y) * p.lv[0].pitch + size_t(2 * x) * F::kTexelBytes;
    V out;
    if constexpr(HasLoadReduce4<F>::value)
      out = F::loadReduce4(tables, s, p.lv[0].pitch, 2u * x, 2u * y, p.lv[0].level);
    else
    {
      const V ul = F::load(tables, s), ur = F::load(tables, s + F::kTexelBytes);
      const V ll = F::load(tables, s + p.lv[0].pitch), lr = F::load(tables, s + p.lv[0].pitch + F::kTexelBytes);
      out        = F::reduce4(ul, ll, ur, lr);
    }
    F::template store<false>(tables, p.lv[1].ptr + size_t(y) * p.lv[1].pitch + size_t(x) * F::kTexelBytes, out);
  }
}

// A fast step of at most three levels on an input level of at most 8 x 8 texels (the last dispatch of a power-of-two
// chain: 4 x 4 -> 1 after the 256^2 tail of 16384^2, 2 x 2 -> 1 for 8192^2), by ONE warp with one lane per texel of
// level +1.  fastTileLoop would give the whole level to a single thread -- some 700 dependent instructions, 4 us on
// B200 -- here a lane does 1/16 of that.  Same expression trees: level +1 pairs vertically (k = 1), the later levels
// as fastPairingIsHorizontal says, over lanes arranged 4 x 4 (x neighbour = lane ^ 1, y neighbour = lane ^ 4) with
// the even/even lane holding (ul, ur, ll, lr) in this order.  No barrier inside; lanes >= 32 return at once.
template <class F>
__device__ __forceinline__ void fastTinyStep(const LevelView* lv, uint32_t M, const typename F::Shared& tables)
{
  using V               = typename F::Value;
  constexpr uint32_t TB = F::kTexelBytes;
  if(threadIdx.x >= 32u)
    return;
  const uint32_t lane = threadIdx.x, x = lane & 3u, y = (lane >> 2) & 3u;
  const bool     valid = lane < 16u && x < lv[1].w && y < lv[1].h;
  V              l1{};
  if(valid)
  {
    const unsigned char* s = lv[0].ptr + size_t(2u * y) * lv[0].pitch + size_t(2u * x) * TB;
    if constexpr(HasLoadReduce4<F>::value)
      l1 = F::loadReduce4(tables, s, lv[0].pitch, 2u * x, 2u * y, lv[0].level);
    else
    {
      const V ul = F::load(tables, s), ur = F::load(tables, s + TB);
      const V ll = F::load(tables, s + lv[0].pitch), lr = F::load(tables, s + lv[0].pitch + TB);
      l1         = F::reduce4(ul, ll, ur, lr);
    }
    F::template store<false>(tables, lv[1].ptr + size_t(y) * lv[1].pitch + size_t(x) * TB, l1);
  }
  if(M < 2u)
    return;
  const V        sx = shflXor(l1, 1), sy = shflXor(l1, 4), sxy = shflXor(l1, 5);
  const V        l2 = reduce4Paired<F>(fastPairingIsHorizontal(2, int(M)), l1, sx, sy, sxy);
  const bool     even = !(x & 1u) && !(y & 1u);
  if(valid && even)
    F::template store<false>(tables, lv[2].ptr + size_t(y >> 1) * lv[2].pitch + size_t(x >> 1) * TB, l2);
  if(M < 3u)
    return;
  const V tx = shflXor(l2, 2), ty = shflXor(l2, 8), txy = shflXor(l2, 10);
  const V l3 = reduce4Paired<F>(fastPairingIsHorizontal(3, int(M)), l2, tx, ty, txy);
  if(valid && lane == 0u)
    F::template store<false>(tables, lv[3].ptr, l3);
}
constexpr uint32_t kFastTinyMaxEdge = 8;  // fastTinyStep: input levels up to 8 x 8, at most three levels

template <class F>
__global__ void __launch_bounds__(256) fastKernel1(const FastParams p)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  FastSmem<F>& sm = *reinterpret_cast<FastSmem<F>*>(smemRaw);
  F::sharedInit(sm.tables, p.tables);
  __syncthreads();
  gridDependencyWait();    // the previous kernel's levels are complete and visible
  gridLaunchDependents();  // the next kernel may start its own set-up as SMs become free
  fastLoop1<F>(p, sm.tables, uint64_t(blockIdx.x) * blockDim.x + threadIdx.x, uint64_t(gridDim.x) * blockDim.x);
}

// ---------------------------------------------------------------------------
// General pipeline.

struct GeneralParams
{
  LevelView           lv[3];  // input, +1, +2
  uint32_t            levels; // 1 or 2
  uint32_t            tilesX, tilesY;
  const DeviceTables* tables;
  // 1 / (2 n + 1) for n = width / height of lv[1], lv[2], divided on the host (the same IEEE float32 division); 0 =
  // not supplied, the kernel divides.  The small steps at the end of a chain run cold code once, and there an IEEE
  // division -- a call into a slow-path subroutine somewhere else in the instruction stream -- costs several hundred
  // cycles per output texel (tools/tail_clocks.sh: level +2 of a 15 x 15 -> 7 x 7 -> 3 x 3 step, two divisions per
  // texel, 2200-4400 cycles; of 15 x 8 -> 7 x 4 -> 3 x 2, one division, 700).
  float               rcpW[3] = {0.f, 0.f, 0.f}, rcpH[3] = {0.f, 0.f, 0.f};
};

constexpr int kGenTile2 = 16;  // tile edge in level +2 of the stand-alone general kernel
constexpr int kGenTile2Small = 8;  // ... of the tail kernel's grid step (more, smaller tiles: the levels are tiny)
// ... and of its SOLO steps, which one CTA runs alone.  Measured (round 2): making the solo tile as large as shared
// memory allows (32 x 32 of level +2, so that a 127^2 or 63^2 input is ONE tile and the chain needs one tail launch
// instead of two) is SLOWER -- 4095^2 47.3 -> 52.4 us, 2047^2 24.7 -> 31.5 us: one CTA walks the 3969 (961) texels
// of level +1 in 8 (2) rounds of nine L2-latency loads each, whereas a second launch spreads them over 16 (4) CTAs in
// one round and hides its launch latency behind the first through programmatic dependent launch.
template <class V>
struct SoloTile2
{
  static constexpr int value = kGenTile2Small;
};
constexpr uint32_t kSoloMaxEdgeGeneral = 32;  // a general step whose input is at most this wide and high runs solo

// Shared scratch of a T2 x T2 tile of level +2: (2 T2 + 1)^2 texels of level +1 (float32 carry).
template <int T2, class V = float4>
struct GenTile
{
  static constexpr int kTile1 = 2 * T2 + 1;
  static constexpr int kPitch = kTile1 + 1;
  V                    l1[kTile1][kPitch];  // [y][x]
};

template <class F>
struct GeneralSmem
{
  typename F::Shared                    tables;
  GenTile<kGenTile2, typename F::Value> tile;
};

// kernelSizeFromInputSize_, glsl:557-561
__device__ __forceinline__ int kernelTaps(uint32_t size)
{
  return size == 1u ? 1 : int(2u | (size & 1u));
}

// Weights of the 3-tap kernel for destination index i of n (glsl:582-586, :639-643):
// (n - i, n, 1 + i) / (2n + 1) with w2 evaluated as 1 - w0 - w1.
__device__ __forceinline__ void taps3(uint32_t n, uint32_t i, float& w0, float& w1, float& w2, float hostRcp = 0.f)
{
  const float fn  = float(n);
  const float rcp = hostRcp != 0.f ? hostRcp : __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(2.0f, fn), 1.0f));
  w0              = __fmul_rn(rcp, __fsub_rn(fn, float(i)));
  w1              = __fmul_rn(rcp, fn);
  w2              = __fsub_rn(__fsub_rn(1.0f, w0), w1);
}

// reduceStoreSample_ without the store (glsl:575-651): vertical reduction of each
// source column, then horizontal reduction.  `fetch(dx, dy)` returns the source
// sample at (srcX + dx, srcY + dy).
template <class F, class Fetch>
__device__ __forceinline__ typename F::Value reduceSample(int kx, int ky, uint32_t dstW, uint32_t dstH,
                                                          uint32_t dx, uint32_t dy, Fetch fetch, float rcpW = 0.f,
                                                          float rcpH = 0.f)
{
  using V = typename F::Value;
  V     hcol[3];
  float w0 = 0.f, w1 = 0.f, w2 = 0.f;
  if(ky == 3)
    taps3(dstH, dy, w0, w1, w2, rcpH);
#pragma unroll
  for(int c = 0; c < 3; ++c)
  {
    if(c < kx)
    {
      const V v0 = fetch(c, 0);
      if(ky == 3)
        hcol[c] = F::reduce(w0, v0, w1, fetch(c, 1), w2, fetch(c, 2));
      else if(ky == 2)
        hcol[c] = F::reduce2(v0, fetch(c, 1));
      else
        hcol[c] = v0;
    }
  }
  if(kx == 3)
  {
    taps3(dstW, dx, w0, w1, w2, rcpW);
    return F::reduce(w0, hcol[0], w1, hcol[1], w2, hcol[2]);
  }
  if(kx == 2)
    return F::reduce2(hcol[0], hcol[1]);
  return hcol[0];
}

// Tile loop of a 1- or 2-level general step, executed by a whole CTA (contains CTA barriers).
// p.tilesX/Y must have been computed for T2 (generalTiles()).
template <class F, int T2>
__device__ __forceinline__ void generalTileLoop(const GeneralParams& p, const typename F::Shared& tables,
                                                GenTile<T2, typename F::Value>& scratch, uint32_t firstTile,
                                                uint32_t tileStride, uint32_t deferredWait = 0u)
{
  using V                             = typename F::Value;
  V(*l1buf)[GenTile<T2, V>::kPitch] = scratch.l1;
  const LevelView L0 = p.lv[0], L1 = p.lv[1], L2 = p.lv[2];
  const int       k1x = kernelTaps(L0.w), k1y = kernelTaps(L0.h);
  const uint32_t  numTiles = p.tilesX * p.tilesY;

  for(uint32_t tile = firstTile; tile < numTiles; tile += tileStride)
  {
    const uint32_t tileX = tile % p.tilesX, tileY = tile / p.tilesX;
    if(deferredWait)  // (see fastTileLoop)
    {
      gridDependencyWait();
      if(deferredWait == 1u)
        gridLaunchDependents();
      deferredWait = 0u;
    }
    if(p.levels == 1)
    {
      // (2 T2) x (2 T2) tile of level +1, no carry needed.
      for(uint32_t t = threadIdx.x; t < 4u * T2 * T2; t += blockDim.x)
      {
        const uint32_t x = tileX * (2u * T2) + (t % (2u * T2)), y = tileY * (2u * T2) + (t / (2u * T2));
        if(x >= L1.w || y >= L1.h)
          continue;
        const unsigned char* s   = L0.ptr + size_t(2 * y) * L0.pitch + size_t(2 * x) * F::kTexelBytes;
        const V              out = reduceSample<F>(k1x, k1y, L1.w, L1.h, x, y, [&](int dx, int dy) {
          return F::load(tables, s + size_t(dy) * L0.pitch + size_t(dx) * F::kTexelBytes);
        }, p.rcpW[1], p.rcpH[1]);
        F::template store<true>(tables, L1.ptr + size_t(y) * L1.pitch + size_t(x) * F::kTexelBytes, out);
      }
      continue;
    }

    // Two levels.  Tile of level +2: [x2a, x2b) x [y2a, y2b); the level +1 footprint it
    // needs starts at (2*x2a, 2*y2a) and spans 2*n + taps - 2 texels per axis.
    const int      k2x = kernelTaps(L1.w), k2y = kernelTaps(L1.h);
    const uint32_t x2a = tileX * T2, y2a = tileY * T2;
    const uint32_t x2b = min(x2a + T2, L2.w), y2b = min(y2a + T2, L2.h);
    const uint32_t fw = min(2u * (x2b - x2a) + uint32_t(k2x) - 2u, L1.w - 2u * x2a);
    const uint32_t fh = min(2u * (y2b - y2a) + uint32_t(k2y) - 2u, L1.h - 2u * y2a);

    __syncthreads();  // previous tile's level +2 pass is done with sm.l1
    for(uint32_t t = threadIdx.x; t < fw * fh; t += blockDim.x)
    {
      const uint32_t lx = t % fw, ly = t / fw;
      const uint32_t x = 2u * x2a + lx, y = 2u * y2a + ly;
      const unsigned char* s   = L0.ptr + size_t(2 * y) * L0.pitch + size_t(2 * x) * F::kTexelBytes;
      const V              out = reduceSample<F>(k1x, k1y, L1.w, L1.h, x, y, [&](int dx, int dy) {
        return F::load(tables, s + size_t(dy) * L0.pitch + size_t(dx) * F::kTexelBytes);
      }, p.rcpW[1], p.rcpH[1]);
      // The halo column/row is also produced (with identical bits) by the neighbouring
      // tile, exactly like the reference's overlapping work groups.
      F::template store<true>(tables, L1.ptr + size_t(y) * L1.pitch + size_t(x) * F::kTexelBytes, out);
      l1buf[ly][lx] = F::sharedRound(out);  // sharedLevel_ (glsl:717)
    }
    __syncthreads();
    const uint32_t tw = x2b - x2a, th = y2b - y2a;
    for(uint32_t t = threadIdx.x; t < tw * th; t += blockDim.x)
    {
      const uint32_t lx = t % tw, ly = t / tw;
      const V        out = reduceSample<F>(k2x, k2y, L2.w, L2.h, x2a + lx, y2a + ly,
                                           [&](int dx, int dy) { return l1buf[2 * ly + dy][2 * lx + dx]; }, p.rcpW[2], p.rcpH[2]);
      F::template store<true>(tables, L2.ptr + size_t(y2a + ly) * L2.pitch + size_t(x2a + lx) * F::kTexelBytes,
                              out);
    }
  }
  if(deferredWait)  // (a CTA without a tile)
  {
    gridDependencyWait();
    if(deferredWait == 1u)
      gridLaunchDependents();
  }
}

template <class F>
__global__ void __launch_bounds__(256) generalKernel(const GeneralParams p)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  GeneralSmem<F>& sm = *reinterpret_cast<GeneralSmem<F>*>(smemRaw);
  F::sharedInit(sm.tables, p.tables);
  __syncthreads();
  // (the tile loop waits for the previous kernel right before its first load and lets the next one go: deferredWait 1)
  generalTileLoop<F, kGenTile2>(p, sm.tables, sm.tile, blockIdx.x, gridDim.x, 1u);
}

// ---------------------------------------------------------------------------
// Linear-filter blit of one level into the next (NVPYR_FLAG_GENERAL_BLIT; demo_app/mipmap_pipelines.cpp:418-426:
// vkCmdBlitImage(level -> level + 1, whole extents, VK_FILTER_LINEAR)).  Vulkan's blit rule: the centre of destination
// texel (i, j) maps to source coordinates ((i + 0.5) * srcW / dstW, (j + 0.5) * srcH / dstH), which are sampled with
// an unnormalised, clamp-to-edge linear filter: texels floor(u - 0.5) and floor(u - 0.5) + 1 with weights (1 - a, a),
// a = frac(u - 0.5).  Arithmetic pinned here (Vulkan leaves it implementation-defined): float32; scale = srcW / dstW
// (IEEE division); u = (i + 0.5) * scale - 0.5 (two roundings); the three lerps go through the functor set's own
// REDUCE as reduce(1 - a, p, a, q, 0, q): horizontally in both rows, then vertically -- so the blit works for any
// functor set and an sRGB image is filtered in linear space, like a texture unit does it.
__device__ __forceinline__ void blitTap(uint32_t i, float scale, uint32_t srcSize, uint32_t& i0, uint32_t& i1, float& a)
{
  const float u = __fsub_rn(__fmul_rn(__fadd_rn(float(i), 0.5f), scale), 0.5f);
  const float f = floorf(u);
  a             = __fsub_rn(u, f);
  const int   k = int(f), last = int(srcSize) - 1;
  i0            = uint32_t(min(max(k, 0), last));
  i1            = uint32_t(min(max(k + 1, 0), last));
}
// Destination texels first, first + stride, ... (thread indices), by the calling thread.
template <class F>
__device__ __forceinline__ void blitLoop(const LevelView& src, const LevelView& dst, const typename F::Shared& tables, uint64_t first,
                                         uint64_t stride)
{
  using V                = typename F::Value;
  constexpr uint32_t TB  = F::kTexelBytes;
  const float        sx  = __fdiv_rn(float(src.w), float(dst.w)), sy = __fdiv_rn(float(src.h), float(dst.h));
  const uint64_t     n   = uint64_t(dst.w) * dst.h;
  for(uint64_t t = first; t < n; t += stride)
  {
    const uint32_t y = uint32_t(t / dst.w), x = uint32_t(t - uint64_t(y) * dst.w);
    uint32_t       x0, x1, y0, y1;
    float          a, b;
    blitTap(x, sx, src.w, x0, x1, a);
    blitTap(y, sy, src.h, y0, y1, b);
    const unsigned char* r0  = src.ptr + size_t(y0) * src.pitch;
    const unsigned char* r1  = src.ptr + size_t(y1) * src.pitch;
    const V              t00 = F::load(tables, r0 + size_t(x0) * TB), t10 = F::load(tables, r0 + size_t(x1) * TB);
    const V              t01 = F::load(tables, r1 + size_t(x0) * TB), t11 = F::load(tables, r1 + size_t(x1) * TB);
    const float          ia = __fsub_rn(1.0f, a), ib = __fsub_rn(1.0f, b);
    const V              top = F::reduce(ia, t00, a, t10, 0.0f, t10), bot = F::reduce(ia, t01, a, t11, 0.0f, t11);
    F::template store<true>(tables, dst.ptr + size_t(y) * dst.pitch + size_t(x) * TB, F::reduce(ib, top, b, bot, 0.0f, bot));
  }
}

// ---------------------------------------------------------------------------
// Tail kernel: several consecutive small steps of a plan in ONE launch.
//
// The reference records one dispatch + pipeline barrier per step
// (nvpro_pyramid_dispatch.hpp:142-187); on B200 a launch costs more than the work of a small
// level.  Step 0 ("grid step") is spread over all CTAs; the CTA that finishes it LAST (atomic
// ticket) then runs the remaining ("solo") steps alone, with CTA barriers where the reference
// has pipeline barriers.  No CTA ever waits for another one, so no cooperative launch is
// needed.  Carry groups (and therefore bits) are unchanged: each step still re-reads the
// 8-bit level written by the previous step.
constexpr uint32_t kMaxTailSteps = 12;
// 512 threads: the general steps stride over all of them (fewer serial rounds per tile: 2047^2 29.1 -> 27.5 us,
// 1080p 15.3 -> 14.7 us; 1024 brings nothing more), fast tiles use the first 256.
#ifndef NVPYR_TAIL_THREADS
#define NVPYR_TAIL_THREADS 512
#endif
constexpr int kTailThreads = NVPYR_TAIL_THREADS;

struct TailStep
{
  uint32_t  pipeline;  // 1 fast, 0 general, 2 one level by a linear-filter blit (NVPYR_FLAG_GENERAL_BLIT),
                       // 3 cascade: SEVERAL consecutive general dispatches, `levels` levels in all (cascadeRun)
  uint32_t  levels;
  uint32_t  vec;       // fast: bit 0 vector loads/stores allowed, bit 1 solo step small enough for fastTinyStep
  uint32_t  soloSmem;  // solo general step that runs on whole levels held in shared memory (soloGeneralSmem)
  uint32_t  tilesX, tilesY;
  LevelView lv[7];
  // cascade only
  uint32_t  tileW, tileH;    // tile of the LAST level that one CTA owns (solo: the whole level)
  uint32_t  boundaryMask;    // bit l: level l ends a reference dispatch (the next level re-reads it as stored texels)
  uint32_t  atStage;         // level 0 footprint is decoded while it is staged (values in shared memory), else raw texels
  uint32_t  off0, offA, offB;  // byte offsets of the three buffers inside the cascade area
  uint32_t  pad_;
  // general steps: 1 / (2 n + 1) for n = width / height of lv[k], k >= 1, divided on the host (GeneralParams::rcpW)
  float     rcpW[7], rcpH[7];
};
// (host) fills the reciprocals of a step whose lv[1 .. levels] are set
inline void fillTailStepReciprocals(TailStep& ts)
{
  for(uint32_t k = 0; k < 7u; ++k)
  {
    const bool used = k >= 1u && k <= ts.levels;
    ts.rcpW[k]      = used ? 1.0f / (2.0f * float(ts.lv[k].w) + 1.0f) : 0.f;
    ts.rcpH[k]      = used ? 1.0f / (2.0f * float(ts.lv[k].h) + 1.0f) : 0.f;
  }
}

struct TailParams
{
  uint32_t            numSteps;
  uint32_t*           ticket;  // zero on entry, reset to zero by the last CTA
  const DeviceTables* tables;
  uint32_t            deferWait;    // the grid step's tile loop executes griddepcontrol.wait itself (NVPYR_TAIL_DEFER_WAIT=0: at kernel entry)
  long long*          debugClocks;  // NVPYR_TAIL_DEBUG_CLOCKS=1: 32 clock64() stamps per CTA (phase timeline); else null
  TailStep            steps[kMaxTailSteps];
};
#define NVPYR_TAIL_STAMP(k)                                                                                          \
  if(tp.debugClocks != nullptr && threadIdx.x == 0)                                                                  \
  tp.debugClocks[blockIdx.x * 32u + (k)] = clock64()

template <class F>
struct TailSmem
{
  typename F::Shared tables;
  union
  {
    typename F::Value                                        l3[2][8][8];
    GenTile<kGenTile2Small, typename F::Value>               tile;
    GenTile<SoloTile2<typename F::Value>::value, typename F::Value> soloTile;
  };
  uint32_t isLast;
};

// ---------------------------------------------------------------------------
// Solo general steps on levels held in shared memory.
//
// The last steps of every NPOT chain are general dispatches on tiny levels (127^2, 63^2, 31^2, ...): three to
// five dependent dispatches that the reference separates with pipeline barriers.  Run by one CTA straight from
// global memory, each costs several dependent L2 round trips (nine texel fetches per output texel, 2 us or more
// per step); spread over a second launch they cost its latency.  Here the CTA that runs the solo steps copies the
// step's whole input level into shared memory once (coalesced), computes level +1 from there, keeps it as float32
// values in shared memory for level +2 -- the reference's sharedLevel_ carry, glsl:555,664,717, for the whole level
// instead of per work group: the same values -- and leaves the step's last level, as stored texels, in a small
// shared buffer that is the next solo step's input.  Every level is also written to global memory as before; what
// changes is only where the next step reads it from (the same bytes).  The functor set is used unchanged: its
// load / store work on generic pointers.
constexpr uint32_t kSoloInBytes  = 65536;  // a step's input level, stored texels, tight rows (127^2 sRGBA8, 63^2 rgba32f)
constexpr uint32_t kSoloMidBytes = 65536;  // its level +1 as values (63^2 float4)
constexpr uint32_t kSoloOutBytes = 16384;  // its last level, stored texels = the next step's input (ping-pong)
struct SoloSmem
{
  alignas(16) unsigned char in[kSoloInBytes];
  alignas(16) unsigned char mid[kSoloMidBytes];
  alignas(16) unsigned char out[2][kSoloOutBytes];
};
// Can a general step (input w0 x h0, `levels` levels) run on SoloSmem with functor set F?
template <class F>
__host__ __device__ inline bool soloSmemFits(uint32_t w0, uint32_t h0, uint32_t levels)
{
  const uint64_t w1 = w0 > 1u ? w0 / 2u : 1u, h1 = h0 > 1u ? h0 / 2u : 1u;
  const uint64_t w2 = w1 > 1u ? w1 / 2u : 1u, h2 = h1 > 1u ? h1 / 2u : 1u;
  const uint64_t last = levels == 2u ? w2 * h2 : w1 * h1;
  return uint64_t(w0) * h0 * F::kTexelBytes <= kSoloInBytes && w1 * h1 * sizeof(typename F::Value) <= kSoloMidBytes
         && last * F::kTexelBytes <= kSoloOutBytes;
}

// One general step.  inSmem: where the input level lies in shared memory as tight rows of stored texels (nullptr:
// copy it from global memory into solo.in first).  Leaves the step's last level in `outSmem` the same way.
template <class F>
__device__ __forceinline__ void soloGeneralSmem(const TailStep& st, const typename F::Shared& tables, SoloSmem& solo,
                                                const unsigned char* inSmem, unsigned char* outSmem)
{
  using V                  = typename F::Value;
  constexpr uint32_t TB    = F::kTexelBytes;
  static_assert(TB % 4u == 0, "levels are staged in 4- or 16-byte words");
  const LevelView    L0 = st.lv[0], L1 = st.lv[1], L2 = st.lv[2];
  const uint32_t     pitchIn = L0.w * TB;
  if(inSmem == nullptr)
  {
    // rows are tight in shared memory; in global memory they start on texel boundaries only (NPOT pitches)
    constexpr uint32_t kWord = TB % 16u == 0 ? 16u : 4u;
    const uint32_t     wordsPerRow = pitchIn / kWord, words = wordsPerRow * L0.h;
    for(uint32_t i = threadIdx.x; i < words; i += blockDim.x)
    {
      const uint32_t       y = i / wordsPerRow, x = i - y * wordsPerRow;
      const unsigned char* g = L0.ptr + size_t(y) * L0.pitch + size_t(x) * kWord;
      if(kWord == 16u)
        *reinterpret_cast<uint4*>(solo.in + size_t(i) * 16u) = __ldcg(reinterpret_cast<const uint4*>(g));
      else
        *reinterpret_cast<uint32_t*>(solo.in + size_t(i) * 4u) = __ldcg(reinterpret_cast<const uint32_t*>(g));
    }
    inSmem = solo.in;
    __syncthreads();
  }
  const int k1x = kernelTaps(L0.w), k1y = kernelTaps(L0.h);
  V*        mid = reinterpret_cast<V*>(solo.mid);
  for(uint32_t t = threadIdx.x; t < L1.w * L1.h; t += blockDim.x)
  {
    const uint32_t       y = t / L1.w, x = t - y * L1.w;
    const unsigned char* s = inSmem + size_t(2u * y) * pitchIn + size_t(2u * x) * TB;
    const V out = reduceSample<F>(k1x, k1y, L1.w, L1.h, x, y,
                                  [&](int dx, int dy) { return F::load(tables, s + size_t(dy) * pitchIn + size_t(dx) * TB); },
                                  st.rcpW[1], st.rcpH[1]);
    F::template store<true>(tables, L1.ptr + size_t(y) * L1.pitch + size_t(x) * TB, out);
    if(st.levels == 2u)
      mid[t] = F::sharedRound(out);  // sharedLevel_ (glsl:717)
    else
      F::template store<true>(tables, outSmem + size_t(t) * TB, out);
  }
  if(st.levels == 2u)
  {
    __syncthreads();
    const int k2x = kernelTaps(L1.w), k2y = kernelTaps(L1.h);
    for(uint32_t t = threadIdx.x; t < L2.w * L2.h; t += blockDim.x)
    {
      const uint32_t y = t / L2.w, x = t - y * L2.w;
      const V*       m = mid + size_t(2u * y) * L1.w + 2u * x;
      const V        out = reduceSample<F>(k2x, k2y, L2.w, L2.h, x, y, [&](int dx, int dy) { return m[size_t(dy) * L1.w + dx]; },
                                           st.rcpW[2], st.rcpH[2]);
      F::template store<true>(tables, L2.ptr + size_t(y) * L2.pitch + size_t(x) * TB, out);
      F::template store<true>(tables, outSmem + size_t(t) * TB, out);
    }
  }
}

// ---------------------------------------------------------------------------
// Cascade: several consecutive GENERAL dispatches on shared-memory tiles, no launch boundary between them.
//
// An NPOT chain ends in five to seven general dispatches on levels of 1023^2 texels and less.  One launch per
// dispatch costs more in launch boundaries (2-3 us each, even with programmatic dependent launch) than in work.
// Here ONE CTA owns a tile of the LAST level of a group of up to three dispatches (up to six levels) and computes
// everything that tile depends on by itself: its footprint on the group's input level is staged in shared memory
// once, every level between is produced into shared memory for exactly the footprint the tile needs (halo
// texels are recomputed by the neighbouring CTA -- the same expression on the same inputs, hence the same bits, just
// as the reference's overlapping work groups recompute them), and each CTA writes to global memory
// the part of every level that lies above its own tile.  No CTA waits for another.  What the reference's dispatch
// boundaries mean is kept: INSIDE a dispatch the first level is handed to the second as float32 values
// (sharedLevel_, glsl:555,664,717: F::sharedRound); ACROSS a dispatch boundary the next level sees what was stored,
// so the value goes through F::store and F::load (for sRGBA8: encode to 8 bits, decode again) -- in registers.
// With a tile that covers the whole last level the same function is the "solo" continuation of the chain on one CTA.
struct CascadeRects
{
  uint32_t x0[7], y0[7], w[7], h[7];  // footprint of the CTA's tile on level l of the group (texels)
  uint32_t ownX1[7], ownY1[7];        // the CTA writes texels [x0, ownX1) x [y0, ownY1) of level l to global memory
};
constexpr uint32_t kCascadeHeaderBytes = 256;         // CascadeRects at the start of the cascade area
constexpr uint32_t kCascadeAreaMax     = 200u * 1024u;  // the area's largest size (host: cascadeGeometry)
static_assert(sizeof(CascadeRects) <= kCascadeHeaderBytes, "cascade header");

// t / d for t, d < 2^16 with magic = ceil(2^32 / d) (exact: t * (magic * d - 2^32) < 2^32); d = 1 has no 32-bit magic.
__device__ __forceinline__ uint32_t cascadeMagic(uint32_t d)
{
  return d > 1u ? 0xFFFFFFFFu / d + 1u : 0u;  // = ceil(2^32 / d), also when d divides 2^32
}
__device__ __forceinline__ uint32_t cascadeDivide(uint32_t t, uint32_t d, uint32_t magic)
{
  return d > 1u ? __umulhi(t, magic) : t;
}

// The weights of one axis of a level: everything of taps3() that does not depend on the destination index.
struct CascadeAxis
{
  int   taps;
  float fn, rcp, w1;
};
__device__ __forceinline__ CascadeAxis cascadeAxis(uint32_t srcSize, uint32_t dstSize, float hostRcp = 0.f)
{
  CascadeAxis a;
  a.taps = kernelTaps(srcSize);
  a.fn   = float(dstSize);
  a.rcp  = hostRcp != 0.f ? hostRcp : __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(2.0f, a.fn), 1.0f));
  a.w1   = __fmul_rn(a.rcp, a.fn);
  return a;
}
// reduceSample() with the per-level part of the weights precomputed (the same operations on the same values).
template <class F, class Fetch>
__device__ __forceinline__ typename F::Value cascadeSample(const CascadeAxis& ax, const CascadeAxis& ay, uint32_t dx, uint32_t dy,
                                                           Fetch fetch)
{
  using V = typename F::Value;
  V     hcol[3];
  float w0 = 0.f, w2 = 0.f;
  if(ay.taps == 3)
  {
    w0 = __fmul_rn(ay.rcp, __fsub_rn(ay.fn, float(dy)));
    w2 = __fsub_rn(__fsub_rn(1.0f, w0), ay.w1);
  }
#pragma unroll
  for(int c = 0; c < 3; ++c)
  {
    if(c < ax.taps)
    {
      const V v0 = fetch(c, 0);
      if(ay.taps == 3)
        hcol[c] = F::reduce(w0, v0, ay.w1, fetch(c, 1), w2, fetch(c, 2));
      else if(ay.taps == 2)
        hcol[c] = F::reduce2(v0, fetch(c, 1));
      else
        hcol[c] = v0;
    }
  }
  if(ax.taps == 3)
  {
    w0 = __fmul_rn(ax.rcp, __fsub_rn(ax.fn, float(dx)));
    w2 = __fsub_rn(__fsub_rn(1.0f, w0), ax.w1);
    return F::reduce(w0, hcol[0], ax.w1, hcol[1], w2, hcol[2]);
  }
  if(ax.taps == 2)
    return F::reduce2(hcol[0], hcol[1]);
  return hcol[0];
}

template <class F>
__device__ __forceinline__ void cascadeRun(const TailStep& st, const typename F::Shared& tables, unsigned char* area,
                                           uint32_t firstTile, uint32_t tileStride, long long* dbg = nullptr)
{
  uint32_t dbgSlot = 16u;
#define NVPYR_CASCADE_STAMP()                                                                                        \
  if(dbg != nullptr && threadIdx.x == 0 && dbgSlot < 32u)                                                            \
  dbg[dbgSlot++] = clock64()
  using V                  = typename F::Value;
  constexpr uint32_t TB    = F::kTexelBytes;
  static_assert(TB % 4u == 0 && TB <= 16u, "texels are staged as 4- or 16-byte words");
  constexpr uint32_t kWord = TB % 16u == 0 ? 16u : 4u;
  CascadeRects&      R     = *reinterpret_cast<CascadeRects*>(area);
  unsigned char*     raw0  = area + st.off0;
  V*                 buf0  = reinterpret_cast<V*>(area + st.off0);
  V*                 bufA  = reinterpret_cast<V*>(area + st.offA);
  V*                 bufB  = reinterpret_cast<V*>(area + st.offB);
  const uint32_t     n = st.levels, numTiles = st.tilesX * st.tilesY;

  for(uint32_t tile = firstTile; tile < numTiles; tile += tileStride)
  {
    __syncthreads();  // the previous tile (or step) is done with the area
    if(threadIdx.x < 2u)
    {
      // thread 0: x ranges, thread 1: y ranges.  Level l + 1 texel i reads texels 2i .. 2i + taps - 1 of level l.
      const bool     isY  = threadIdx.x == 1u;
      const uint32_t t    = isY ? tile / st.tilesX : tile % st.tilesX;
      const uint32_t tsz  = isY ? st.tileH : st.tileW;
      const uint32_t endN = isY ? st.lv[n].h : st.lv[n].w;
      uint32_t*      x0   = isY ? R.y0 : R.x0;
      uint32_t*      w    = isY ? R.h : R.w;
      uint32_t*      own  = isY ? R.ownY1 : R.ownX1;
      const uint32_t a = t * tsz, b = min(a + tsz, endN);
      x0[n] = a, w[n] = b - a, own[n] = b;
      for(int l = int(n) - 1; l >= 0; --l)
      {
        const uint32_t size = isY ? st.lv[l].h : st.lv[l].w;
        const uint32_t lo = 2u * x0[l + 1], hi = min(2u * (x0[l + 1] + w[l + 1] - 1u) + uint32_t(kernelTaps(size)), size);
        x0[l]  = lo;
        w[l]   = hi - lo;
        own[l] = b == endN ? size : min(size, b << (n - uint32_t(l)));
      }
    }
    __syncthreads();
    NVPYR_CASCADE_STAMP();

    // Stage the footprint on the group's input level (written by an earlier launch, or by other CTAs of this one
    // before the ticket: read through L2).
    {
      const LevelView      L0 = st.lv[0];
      const uint32_t       perRow = st.atStage ? R.w[0] : R.w[0] * TB / kWord, magic = cascadeMagic(perRow), total = perRow * R.h[0];
      const unsigned char* g0 = L0.ptr + size_t(R.y0[0]) * L0.pitch + size_t(R.x0[0]) * TB;
      if(st.atStage)
      {
        // kStageBatch loads of a thread in flight before the first is decoded (the footprint is a few texels per
        // thread: one L2 round trip instead of one per texel)
        constexpr uint32_t kStageBatch = 8;
        for(uint32_t base = threadIdx.x; base < total; base += blockDim.x * kStageBatch)
        {
          alignas(16) uint32_t texel[kStageBatch][TB / 4u];
#pragma unroll
          for(uint32_t k = 0; k < kStageBatch; ++k)
          {
            const uint32_t i = base + k * blockDim.x;
            if(i < total)
            {
              const uint32_t       y = cascadeDivide(i, perRow, magic), x = i - y * perRow;
              const unsigned char* g = g0 + size_t(y) * L0.pitch + size_t(x) * TB;
              if(kWord == 16u)
                *reinterpret_cast<uint4*>(texel[k]) = __ldcg(reinterpret_cast<const uint4*>(g));
              else
              {
#pragma unroll
                for(uint32_t j = 0; j < TB / 4u; ++j)
                  texel[k][j] = __ldcg(reinterpret_cast<const uint32_t*>(g) + j);
              }
            }
          }
#pragma unroll
          for(uint32_t k = 0; k < kStageBatch; ++k)
          {
            const uint32_t i = base + k * blockDim.x;
            if(i < total)
              buf0[i] = F::load(tables, texel[k]);
          }
        }
      }
      else
      {
        // raw texels: asynchronous copies straight into shared memory (LDGSTS), all in flight at once
        const uint32_t dst0 = uint32_t(__cvta_generic_to_shared(raw0));
        for(uint32_t i = threadIdx.x; i < total; i += blockDim.x)
        {
          const uint32_t       y = cascadeDivide(i, perRow, magic), x = i - y * perRow;
          const unsigned char* g = g0 + size_t(y) * L0.pitch + size_t(x) * kWord;
          if(kWord == 16u)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + i * 16u), "l"(g) : "memory");
          else
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + i * 4u), "l"(g) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
      }
    }
    __syncthreads();
    NVPYR_CASCADE_STAMP();

    const V* in = buf0;
    for(uint32_t l = 1; l <= n; ++l)
    {
      V*                out = (l & 1u) ? bufA : bufB;
      const LevelView   Lo  = st.lv[l];
      const CascadeAxis ax = cascadeAxis(st.lv[l - 1u].w, Lo.w, st.rcpW[l]), ay = cascadeAxis(st.lv[l - 1u].h, Lo.h, st.rcpH[l]);
      const uint32_t    inW = R.w[l - 1u], outW = R.w[l], total = outW * R.h[l], magic = cascadeMagic(outW);
      const uint32_t    ox = R.x0[l], oy = R.y0[l], ownX1 = R.ownX1[l], ownY1 = R.ownY1[l];
      const bool        rawInput = l == 1u && !st.atStage, isLast = l == n, boundary = (st.boundaryMask >> l) & 1u;
      for(uint32_t t = threadIdx.x; t < total; t += blockDim.x)
      {
        const uint32_t ly = cascadeDivide(t, outW, magic), lx = t - ly * outW;
        const uint32_t x = ox + lx, y = oy + ly;
        V              v;
        if(rawInput)
        {
          const unsigned char* s = raw0 + (2u * ly * inW + 2u * lx) * TB;
          v = cascadeSample<F>(ax, ay, x, y, [&](int dx, int dy) { return F::load(tables, s + (uint32_t(dy) * inW + uint32_t(dx)) * TB); });
        }
        else
        {
          const V* m = in + (2u * ly * inW + 2u * lx);
          v          = cascadeSample<F>(ax, ay, x, y, [&](int dx, int dy) { return m[uint32_t(dy) * inW + uint32_t(dx)]; });
        }
        if(x < ownX1 && y < ownY1)
          F::template store<true>(tables, Lo.ptr + size_t(y) * Lo.pitch + size_t(x) * TB, v);
        if(!isLast)
        {
          if(boundary)
          {
            alignas(16) uint32_t texel[TB / 4u];  // what the next dispatch reads back: the stored texel
            F::template store<true>(tables, texel, v);
            out[t] = F::load(tables, texel);
          }
          else
            out[t] = F::sharedRound(v);  // sharedLevel_ (glsl:717)
        }
      }
      __syncthreads();
      NVPYR_CASCADE_STAMP();
      in = out;
    }
  }
#undef NVPYR_CASCADE_STAMP
}

// kSolo: the step is run by one CTA alone (its general tiles were counted for soloTile2).
template <class F, bool kSolo>
__device__ __forceinline__ void tailRunStep(const TailStep& st, TailSmem<F>& sm, const DeviceTables* tables,
                                            uint32_t first, uint32_t stride, unsigned char* area = nullptr, long long* dbg = nullptr,
                                            uint32_t deferredWait = 0u)
{
  if(st.pipeline == 3u)
    cascadeRun<F>(st, sm.tables, area, first, stride, dbg);
  else if(st.pipeline == 2u)
    blitLoop<F>(st.lv[0], st.lv[1], sm.tables, uint64_t(first) * blockDim.x + threadIdx.x, uint64_t(stride) * blockDim.x);
  else if(st.pipeline == 1u && kSolo && (st.vec & 2u))
    fastTinyStep<F>(st.lv, st.levels, sm.tables);
  else if(st.pipeline == 1u)
  {
    FastParams p;
#pragma unroll
    for(int k = 0; k < 7; ++k)
      p.lv[k] = st.lv[k];
    p.tilesX = st.tilesX;
    p.tilesY = st.tilesY;
    p.tables = tables;
#define NVPYR_TAIL_FAST(m)                                                                                        \
  case m:                                                                                                         \
    if(st.vec & 1u)                                                                                               \
      fastTileLoop<F, m, true>(p, sm.tables, sm.l3, first, stride, deferredWait);                                 \
    else                                                                                                          \
      fastTileLoop<F, m, false>(p, sm.tables, sm.l3, first, stride, deferredWait);                                \
    break;
    switch(st.levels)
    {
      case 1:
        fastLoop1<F>(p, sm.tables, uint64_t(first) * blockDim.x + threadIdx.x, uint64_t(stride) * blockDim.x);
        break;
        NVPYR_TAIL_FAST(2)
        NVPYR_TAIL_FAST(3)
        NVPYR_TAIL_FAST(4)
        NVPYR_TAIL_FAST(5)
        NVPYR_TAIL_FAST(6)
    }
#undef NVPYR_TAIL_FAST
  }
  else
  {
    GeneralParams p;
    p.lv[0] = st.lv[0], p.lv[1] = st.lv[1], p.lv[2] = st.lv[2];
    p.levels = st.levels;
    p.tilesX = st.tilesX;
    p.tilesY = st.tilesY;
    p.tables = tables;
    p.rcpW[1] = st.rcpW[1], p.rcpW[2] = st.rcpW[2], p.rcpH[1] = st.rcpH[1], p.rcpH[2] = st.rcpH[2];
    if(kSolo)
      generalTileLoop<F, SoloTile2<typename F::Value>::value>(p, sm.tables, sm.soloTile, first, stride);
    else
      generalTileLoop<F, kGenTile2Small>(p, sm.tables, sm.tile, first, stride, deferredWait);
  }
}

template <class F>
__global__ void __launch_bounds__(kTailThreads) tailKernel(const __grid_constant__ TailParams tp)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  TailSmem<F>& sm = *reinterpret_cast<TailSmem<F>*>(smemRaw);
  NVPYR_TAIL_STAMP(0);
  F::sharedInit(sm.tables, tp.tables);
  __syncthreads();
  NVPYR_TAIL_STAMP(1);
  // Tile steps wait for the previous kernel themselves, right before their first load (see fastTileLoop).
  // (tp.deferWait 2: the next kernel is let go at once, 1: after the wait as in the other kernels)
  const uint32_t deferWait = tp.steps[0].pipeline == 0u || (tp.steps[0].pipeline == 1u && tp.steps[0].levels >= 2u) ? tp.deferWait : 0u;
  if(deferWait == 0u)
    gridDependencyWait();  // the previous kernel's levels are complete and visible
  if(deferWait != 1u)
    gridLaunchDependents();  // the next kernel may start its own set-up as SMs become free
  NVPYR_TAIL_STAMP(2);

  // (launched with sizeof(TailSmem<F>) + the solo buffers or the cascade area when a step needs them)
  unsigned char* area = smemRaw + ((sizeof(TailSmem<F>) + 15u) & ~size_t(15));
  long long* dbg = tp.debugClocks != nullptr ? tp.debugClocks + blockIdx.x * 32u : nullptr;
  tailRunStep<F, false>(tp.steps[0], sm, tp.tables, blockIdx.x, gridDim.x, area, dbg, deferWait);
  NVPYR_TAIL_STAMP(3);
  if(tp.numSteps == 1u)
    return;

  // Publish this CTA's part of step 0, then find out whether it was the last one.  (A grid of one CTA -- images of a
  // few thousand texels -- needs neither the fences nor the atomic, ~1 us: a barrier orders the CTA's own accesses.)
  if(gridDim.x != 1u)
  {
    __threadfence();
    __syncthreads();
    if(threadIdx.x == 0)
    {
      const uint32_t t = atomicAdd(tp.ticket, 1u);
      sm.isLast        = t == gridDim.x - 1u;
      if(sm.isLast)
        *tp.ticket = 0u;  // every CTA has taken its ticket: safe to recycle the counter
    }
    __syncthreads();
    if(!sm.isLast)
      return;
    __threadfence();  // acquire side of the ticket
  }
  else
  {
    __threadfence_block();
    __syncthreads();
  }
  NVPYR_TAIL_STAMP(4);

  SoloSmem&            solo = *reinterpret_cast<SoloSmem*>(area);
  const unsigned char* levelInSmem = nullptr;  // the previous step's last level, if that step left it in shared memory
  uint32_t             pingPong = 0;
  for(uint32_t s = 1; s < tp.numSteps; ++s)
  {
    if(tp.steps[s].soloSmem)
    {
      unsigned char* out = solo.out[pingPong];
      soloGeneralSmem<F>(tp.steps[s], sm.tables, solo, levelInSmem, out);
      levelInSmem = out;
      pingPong ^= 1u;
    }
    else
    {
      tailRunStep<F, true>(tp.steps[s], sm, tp.tables, 0u, 1u, area, dbg);
      levelInSmem = nullptr;
    }
    __threadfence_block();
    __syncthreads();  // the reference's inter-dispatch pipeline barrier (one CTA: a CTA barrier suffices)
    NVPYR_TAIL_STAMP(4u + min(s, 9u));
  }
}

// Batch tail: the remaining (small) steps of `count` independent chains of one size in ONE launch.  CTA c
// runs every step of images c, c + gridDim.x, ... alone; tp.steps[].lv[].ptr hold byte offsets inside a
// chain, bases[i] the chain of image i.  (tp.ticket is unused: nothing crosses CTAs.)
template <class F>
__global__ void __launch_bounds__(kTailThreads) tailBatchKernel(const __grid_constant__ TailParams tp,
                                                        const unsigned char* const* bases, uint32_t count)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  TailSmem<F>& sm = *reinterpret_cast<TailSmem<F>*>(smemRaw);
  F::sharedInit(sm.tables, tp.tables);
  __syncthreads();
  gridDependencyWait();
  gridLaunchDependents();
  for(uint32_t image = blockIdx.x; image < count; image += gridDim.x)
  {
    unsigned char* base = reinterpret_cast<unsigned char*>(__ldg(reinterpret_cast<const unsigned long long*>(bases) + image));
    for(uint32_t s = 0; s < tp.numSteps; ++s)
    {
      TailStep st = tp.steps[s];
#pragma unroll
      for(int k = 0; k < 7; ++k)
        st.lv[k].ptr = base + reinterpret_cast<size_t>(st.lv[k].ptr);
      tailRunStep<F, true>(st, sm, tp.tables, 0u, 1u);
      __threadfence_block();
      __syncthreads();  // the reference's inter-dispatch pipeline barrier
    }
  }
}

// ---------------------------------------------------------------------------
// Stand-alone blit of one (large) level; small levels are blitted by tailKernel (pipeline 2).
struct BlitParams
{
  LevelView           src, dst;
  const DeviceTables* tables;
};
template <class F>
__global__ void __launch_bounds__(256) blitKernel(const BlitParams p)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  typename F::Shared& tables = *reinterpret_cast<typename F::Shared*>(smemRaw);
  F::sharedInit(tables, p.tables);
  __syncthreads();
  gridDependencyWait();
  gridLaunchDependents();
  blitLoop<F>(p.src, p.dst, tables, uint64_t(blockIdx.x) * blockDim.x + threadIdx.x, uint64_t(gridDim.x) * blockDim.x);
}

// ---------------------------------------------------------------------------
// Premultiply-alpha pre-pass, include/scoped_image.hpp:233-255 (sRGBA8 only).
__global__ void __launch_bounds__(256) premultiplyKernel(const uint32_t* in, uint32_t* out, uint64_t texels,
                                                         const DeviceTables* tables)
{
  extern __shared__ __align__(128) unsigned char smemRaw[];
  Srgba8::Shared& sm = *reinterpret_cast<Srgba8::Shared*>(smemRaw);
  Srgba8::sharedInit(sm, tables);
  __syncthreads();
  gridDependencyWait();    // the previous kernel's levels are complete and visible
  gridLaunchDependents();  // the next kernel may start its own set-up as SMs become free
  for(uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < texels; i += uint64_t(gridDim.x) * blockDim.x)
  {
    const uint32_t w = in[i];
    float4         v = Srgba8::decodeWord(sm, w);
    const uint32_t r = Srgba8::encodeChannel<true>(sm, __fmul_rn(v.x, v.w));
    const uint32_t g = Srgba8::encodeChannel<true>(sm, __fmul_rn(v.y, v.w));
    const uint32_t b = Srgba8::encodeChannel<true>(sm, __fmul_rn(v.z, v.w));
    out[i]           = ((r >> 16) & 0xFFu) | (((g >> 16) & 0xFFu) << 8) | (((b >> 16) & 0xFFu) << 16) | (w & 0xFF000000u);
  }
}

}  // namespace nvpyr
