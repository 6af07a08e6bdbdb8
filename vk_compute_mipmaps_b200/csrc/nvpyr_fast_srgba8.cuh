// nvpyr_fast_srgba8.cuh -- the hot kernel: sRGBA8 fast pipeline, 2..6 levels per launch.
//
// Same float32 expression trees as fastKernel<Srgba8, M> (nvpyr_kernels.cuh) and therefore
// the same bits, but arranged for the B200's instruction budget.  At HBM speed the SMs can
// issue only ~118 warp-lane instructions per 2x2 input quad (16 bytes), so the kernel is
// instruction/LSU bound, not DRAM bound (DESIGN.md "Why the fast kernel is issue-bound"):
//
//  * decode: one PRMT builds the shared-memory address (code << 8 | lane << 2) straight from
//    the packed texel word, one LDS reads the lane-private copy of the 256-entry table
//    (stride 256 B per code: no bank conflict for any data).  Alpha goes through I2F (XU
//    pipe) + FMUL to keep the LSU free.
//  * deferred normalisation: 0.25*((a+b)+(c+d)) is carried as the un-normalised sum
//    S = (a+b)+(c+d) = 4^k * value.  Scaling by a power of two commutes with rounding (no
//    value here is subnormal or overflows), so every deeper sum and every stored code is
//    bit-identical, and the FMULs disappear: the encode looks S up directly with a biased
//    clamp/offset, alpha uses the exact constant 255 / 4^k.
//  * level +3 is a "transpose-reduce": in two shuffle steps the four threads of a 2x2 block
//    end up holding ONE channel each of the common result (3 SHFL + 3 FADD per thread
//    instead of 12 + 12), encode it in parallel and gather the four bytes with two PRMTs.
//  * the next tile's four 16-byte rows are prefetched into registers before the current
//    tile is processed.
//
// CTA = 512 threads = 16 warps (2 across x 8 down), warp = 16 x 2 threads, thread = 4x4
// texels: a CTA tile is 128 x 64 texels of the input level; two CTAs per SM.
#pragma once
#include "nvpyr_kernels.cuh"

namespace nvpyr {

struct Srgba8FastSmem
{
  float    decode[256 * 64];      // [code][64]: floats 0..31 = linearFromSrgb(code) per lane, 32..63 unused
  uint32_t encode[kEncEntries];   // bucket table (nvpyr_functors.cuh)
  alignas(16) float l3[2][8][16][4];  // level +3 sums of the current tile (per channel), double buffered
};

__device__ __forceinline__ void srgba8FastInit(Srgba8FastSmem& sm, const DeviceTables* t)
{
  // 512 threads: thread -> (code, half): 16 lane slots = 4 x float4
  for(uint32_t i = threadIdx.x; i < 512u; i += blockDim.x)
  {
    const uint32_t code = i >> 1, half = i & 1u;
    const float    v    = __ldg(&t->decode[code]);
    float4*        d    = reinterpret_cast<float4*>(&sm.decode[code * 64u + half * 16u]);
    const float4   v4   = make_float4(v, v, v, v);
    d[0] = v4, d[1] = v4, d[2] = v4, d[3] = v4;
  }
  for(uint32_t i = threadIdx.x; i < kEncEntries; i += blockDim.x)
    sm.encode[i] = __ldg(&t->encode[i]);
}

// linearFromSrgb of byte k (0..2) of a packed texel: PRMT + LDS.
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

// Un-normalised sum of one 2x2 quad, vertical pairing: (UL + LL) + (UR + LR)  (glsl:180-188).
__device__ __forceinline__ float4 quadSumV(const unsigned char* dec, uint32_t laneOff, uint32_t ul, uint32_t ur,
                                           uint32_t ll, uint32_t lr)
{
  float4 s;
  s.x = __fadd_rn(__fadd_rn(dec8<0>(dec, ul, laneOff), dec8<0>(dec, ll, laneOff)),
                  __fadd_rn(dec8<0>(dec, ur, laneOff), dec8<0>(dec, lr, laneOff)));
  s.y = __fadd_rn(__fadd_rn(dec8<1>(dec, ul, laneOff), dec8<1>(dec, ll, laneOff)),
                  __fadd_rn(dec8<1>(dec, ur, laneOff), dec8<1>(dec, lr, laneOff)));
  s.z = __fadd_rn(__fadd_rn(dec8<2>(dec, ul, laneOff), dec8<2>(dec, ll, laneOff)),
                  __fadd_rn(dec8<2>(dec, ur, laneOff), dec8<2>(dec, lr, laneOff)));
  s.w = __fadd_rn(__fadd_rn(decAlpha(ul), decAlpha(ll)), __fadd_rn(decAlpha(ur), decAlpha(lr)));
  return s;
}

// Encode of S = 4^K * x for one RGB channel; result has the code in bits 16..23.
// bits(S) = bits(x) + (2K << 23) for every non-zero x here, so the clamp and the table
// offset are shifted by 2K exponent steps and the surplus is removed in the final 3-input add.
template <int K>
__device__ __forceinline__ uint32_t encScaled(const unsigned char* encBytes, float s)
{
  constexpr uint32_t kExp  = uint32_t(2 * K) << 23;
  constexpr uint32_t kBias = (kEncMinKey + (uint32_t(2 * K) << (23 - kEncShift))) * 4u;
  uint32_t           b     = max(__float_as_uint(s), kEncMinBits + kExp);
  const uint32_t     off   = (b >> (kEncShift - 2)) & 0x3FFFCu;
  const uint32_t     e     = *reinterpret_cast<const uint32_t*>(encBytes + off - kBias);
  return e + b - kExp;
}
// uint(a * 255 + 0.5) for a = S / 4^K, result has the code in bits 0..7 (a <= 1, no clamp needed).
template <int K>
__device__ __forceinline__ uint32_t encAlphaScaled(float s)
{
  constexpr float kMul = 255.0f / float(1 << (2 * K));  // exact
  const float     v    = __fadd_rn(__fmul_rn(s, kMul), 0.5f);
  return __float_as_uint(__fadd_rz(v, 8388608.0f));  // 2^23 + trunc(v)
}
template <int K>
__device__ __forceinline__ uint32_t encWordScaled(const unsigned char* encBytes, float4 s)
{
  const uint32_t r = encScaled<K>(encBytes, s.x), g = encScaled<K>(encBytes, s.y), b = encScaled<K>(encBytes, s.z);
  const uint32_t a = encAlphaScaled<K>(s.w);
  return __byte_perm(__byte_perm(r, g, 0x0062), __byte_perm(b, a, 0x0042), 0x5410);
}

__device__ __forceinline__ float4 sum4Paired(bool horizontal, float4 ul, float4 ur, float4 ll, float4 lr)
{
  const float4 p = horizontal ? f4add(ul, ur) : f4add(ul, ll);
  const float4 q = horizontal ? f4add(ll, lr) : f4add(ur, lr);
  return f4add(p, q);
}

template <int M>
__global__ void __launch_bounds__(512, 2) fastSrgba8Kernel(const FastParams p)
{
  static_assert(M >= 2 && M <= 6, "2..6 levels");
  extern __shared__ __align__(16) unsigned char smemRaw[];
  Srgba8FastSmem& sm = *reinterpret_cast<Srgba8FastSmem*>(smemRaw);
  srgba8FastInit(sm, p.tables);
  __syncthreads();
  const unsigned char* dec = reinterpret_cast<const unsigned char*>(sm.decode);
  const unsigned char* enc = reinterpret_cast<const unsigned char*>(sm.encode);

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t tx = (warp & 1u) * 16u + (lane & 15u);  // 0..31  (x 4 texels)
  const uint32_t ty = (warp >> 1) * 2u + (lane >> 4);    // 0..15  (x 4 texels)
  const uint32_t laneOff = lane * 4u;
  const uint32_t W = p.lv[0].w, H = p.lv[0].h;
  const uint32_t numTiles = p.tilesX * p.tilesY;
  const size_t   pitch0 = p.lv[0].pitch;

  // level +3 transpose-reduce roles
  const bool     xOdd = lane & 1u, yOdd = lane & 16u;
  const uint32_t ch   = (xOdd ? 2u : 0u) + (yOdd ? 1u : 0u);  // channel this lane ends up owning
  // gather selectors: step 1 (xor 16) builds the 2-byte pair of this column, step 2 (xor 1) the word
  //   x even: pair = (R from y-even lane, G from y-odd lane), codes in byte 2 of both
  //   x odd : pair = (B from y-even lane (byte 2), A from y-odd lane (byte 0))
  const uint32_t sel1 = xOdd ? (yOdd ? 0x0006u : 0x0042u) : (yOdd ? 0x0026u : 0x0062u);
  const uint32_t sel2 = xOdd ? 0x1054u : 0x5410u;

  uint4 row[4];
  auto  loadTile = [&](uint32_t tile, uint4 (&r)[4]) {
    const uint32_t tileX = tile % p.tilesX, tileY = tile / p.tilesX;
    const uint32_t x0 = tileX * 128u + tx * 4u, y0 = tileY * 64u + ty * 4u;
    if(x0 < W && y0 < H)
    {
      const unsigned char* s = p.lv[0].ptr + size_t(y0) * pitch0 + size_t(x0) * 4u;
#pragma unroll
      for(int i = 0; i < 4; ++i)
        r[i] = __ldg(reinterpret_cast<const uint4*>(s + size_t(i) * pitch0));
    }
  };

  uint32_t tile = blockIdx.x;
  if(tile < numTiles)
    loadTile(tile, row);
  uint32_t parity = 0;
  for(; tile < numTiles; tile += gridDim.x, parity ^= 1u)
  {
    const uint32_t tileX = tile % p.tilesX, tileY = tile / p.tilesX;
    const uint32_t x0 = tileX * 128u + tx * 4u, y0 = tileY * 64u + ty * 4u;
    const bool     active = x0 < W && y0 < H;
    uint4          cur[4] = {row[0], row[1], row[2], row[3]};
    if(tile + gridDim.x < numTiles)
      loadTile(tile + gridDim.x, row);  // prefetch

    float4 s2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if(active)
    {
      // level +1 (K = 1): four quads, vertical pairing
      const float4 s00 = quadSumV(dec, laneOff, cur[0].x, cur[0].y, cur[1].x, cur[1].y);
      const float4 s01 = quadSumV(dec, laneOff, cur[0].z, cur[0].w, cur[1].z, cur[1].w);
      unsigned char* d1 = p.lv[1].ptr + size_t(y0 >> 1) * p.lv[1].pitch + size_t(x0 >> 1) * 4u;
      *reinterpret_cast<uint2*>(d1) = make_uint2(encWordScaled<1>(enc, s00), encWordScaled<1>(enc, s01));
      const float4 s10 = quadSumV(dec, laneOff, cur[2].x, cur[2].y, cur[3].x, cur[3].y);
      const float4 s11 = quadSumV(dec, laneOff, cur[2].z, cur[2].w, cur[3].z, cur[3].w);
      *reinterpret_cast<uint2*>(d1 + p.lv[1].pitch) =
          make_uint2(encWordScaled<1>(enc, s10), encWordScaled<1>(enc, s11));
      // level +2 (K = 2) from the thread's own 2x2
      s2 = sum4Paired(fastPairingIsHorizontal(2, M), s00, s01, s10, s11);
      *reinterpret_cast<uint32_t*>(p.lv[2].ptr + size_t(y0 >> 2) * p.lv[2].pitch + size_t(x0 >> 2) * 4u) =
          encWordScaled<2>(enc, s2);
    }

    if(M >= 3)
    {
      // level +3 (K = 3), horizontal pairing for every M: (self + x) + (y + xy).
      // Step x: even lanes keep (R, G), odd lanes keep (B, A).
      const float keep0 = xOdd ? s2.z : s2.x, keep1 = xOdd ? s2.w : s2.y;
      const float send0 = xOdd ? s2.x : s2.z, send1 = xOdd ? s2.y : s2.w;
      const float t0    = __fadd_rn(keep0, __shfl_xor_sync(0xffffffffu, send0, 1));
      const float t1    = __fadd_rn(keep1, __shfl_xor_sync(0xffffffffu, send1, 1));
      // Step y: y-even lanes keep the first of their two channels, y-odd lanes the second.
      const float u = __fadd_rn(yOdd ? t1 : t0, __shfl_xor_sync(0xffffffffu, yOdd ? t0 : t1, 16));
      // encode this lane's channel, gather the four bytes
      const uint32_t code  = ch == 3u ? encAlphaScaled<3>(u) : encScaled<3>(enc, u);
      const uint32_t pair  = __byte_perm(code, __shfl_xor_sync(0xffffffffu, code, 16), sel1);
      const uint32_t word  = __byte_perm(pair, __shfl_xor_sync(0xffffffffu, pair, 1), sel2);
      if(active)
      {
        if(ch == 0u)
          *reinterpret_cast<uint32_t*>(p.lv[3].ptr + size_t(y0 >> 3) * p.lv[3].pitch + size_t(x0 >> 3) * 4u) = word;
        if(M >= 4)
          sm.l3[parity][ty >> 1][tx >> 1][ch] = u;
      }
    }

    if(M >= 4)
    {
      __syncthreads();
      if(tid < 32)
      {
        // 32 lanes <-> 8 x 4 texels of level +4; +5 and +6 with butterflies.
        const uint32_t i = lane & 7u, j = lane >> 3;
        const uint32_t ox = tileX * 128u + i * 16u, oy = tileY * 64u + j * 16u;  // origin in the input level
        const bool     valid = ox < W && oy < H;
        float4         s4    = make_float4(0.f, 0.f, 0.f, 0.f);
        if(valid)
        {
          const float4* l3 = reinterpret_cast<const float4*>(&sm.l3[parity][0][0][0]);
          const float4  ul = l3[(2 * j) * 16 + 2 * i], ur = l3[(2 * j) * 16 + 2 * i + 1];
          const float4  ll = l3[(2 * j + 1) * 16 + 2 * i], lr = l3[(2 * j + 1) * 16 + 2 * i + 1];
          s4               = sum4Paired(fastPairingIsHorizontal(4, M), ul, ur, ll, lr);
          *reinterpret_cast<uint32_t*>(p.lv[4].ptr + size_t(oy >> 4) * p.lv[4].pitch + size_t(ox >> 4) * 4u) =
              encWordScaled<4>(enc, s4);
        }
        if(M >= 5)
        {
          // vertical pairing (shared-memory tail of the reference): (self + y) + (x + xy)
          const float4 a  = f4add(s4, shflXor(s4, 8));
          const float4 s5 = f4add(a, shflXor(a, 1));
          if(valid && !(i & 1u) && !(j & 1u))
            *reinterpret_cast<uint32_t*>(p.lv[5].ptr + size_t(oy >> 5) * p.lv[5].pitch + size_t(ox >> 5) * 4u) =
                encWordScaled<5>(enc, s5);
          if(M >= 6)
          {
            const float4 c  = f4add(s5, shflXor(s5, 16));
            const float4 s6 = f4add(c, shflXor(c, 2));
            if(valid && !(i & 3u) && j == 0u)
              *reinterpret_cast<uint32_t*>(p.lv[6].ptr + size_t(oy >> 6) * p.lv[6].pitch + size_t(ox >> 6) * 4u) =
                  encWordScaled<6>(enc, s6);
          }
        }
      }
    }
  }
}

}  // namespace nvpyr
