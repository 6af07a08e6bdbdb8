/*
 * nvpyr_oracle.c -- CPU restatement of the nvpro_pyramid mip-chain algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the CUDA library
 * libnvpyr.so, the package vk_compute_mipmaps_b200, bench.py's GPU arm) may
 * import, link or execute this file.  It is the checker: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 *
 * Parity pin (see DESIGN.md "Oracle"): every function here is checked in
 * tests/test_oracle_pins.py against the reference's OWN code compiled in place
 * from /root/reference (oracle/_ref/libnvpyr_ref.so, built by oracle/Makefile):
 *   - nvo_plan                 vs nvproCmdPyramidDispatch recorded through a
 *                                 mock Vulkan command buffer
 *   - nvo_linear_from_srgb /
 *     nvo_srgb_from_linear     vs shaders/srgb.h compiled as C++
 *   - nvo_cpu_chain_srgba8     vs cpuGenerateMipmaps_sRGBA (bit-exact)
 *   - nvo_shader_chain_*       vs the reference GLSL (nvpro_pyramid.glsl +
 *                                 srgba8_mipmap_preamble.glsl) executed on the
 *                                 CPU by oracle/glsl_emu (bit-exact), and vs
 *                                 the recorded worst-delta known answers of
 *                                 demo_app/rtx3090.json (<=2 opaque, <=5 alpha)
 *
 * Two oracles:
 *   Oracle A  "shader order": follows the GLSL schedule literally -- work
 *             groups, invocations, Morton sub-tiles, subgroup shuffles, shared
 *             memory, float32 carry inside one dispatch, re-read of the 8-bit
 *             image between dispatches.  Software LOAD_REDUCE4
 *             (USE_BILINEAR_SAMPLING 0 semantics, nvpro_pyramid.glsl:179-189).
 *   Oracle B  "cpu": restatement of the reference's own CPU generator
 *             include/mipmap_storage.hpp:207-414 (re-quantises every level,
 *             truncates alpha).
 *
 * Numerics pinned for Oracle A (implementation-defined in GLSL):
 *   - float32 everywhere, round-to-nearest-even; the compiler may contract nothing
 *     (compile with -ffp-contract=off); the ONE contraction of the contract is explicit:
 *     the 3-tap REDUCE a0*v0 + a1*v1 + a2*v2 is mul, fma, fma (see a_reduce)
 *   - sRGB decode = 256-entry table, sRGB encode = 255 thresholds, both
 *     committed as bit patterns in vk_compute_mipmaps_b200/csrc/srgb_tables.inc
 *     and verified against the srgb.h formulas (nvo_*_formula below)
 *   - alpha decode a*(1/255) as shaders/srgb.h:60, alpha encode
 *     uint(a*255+0.5) clamped as srgba8_mipmap_preamble.glsl:125
 *   - 1/(2n+1) is an IEEE division
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../vk_compute_mipmaps_b200/csrc/srgb_tables.inc"

/* ------------------------------------------------------------------------ */
/* Transfer functions                                                        */

static inline float bits_to_float(uint32_t b)
{
  float f;
  memcpy(&f, &b, 4);
  return f;
}
static inline uint32_t float_to_bits(float f)
{
  uint32_t b;
  memcpy(&b, &f, 4);
  return b;
}

/* shaders/srgb.h:18-28, literal formula (uses libm powf). */
float nvo_linear_from_srgb_formula(uint32_t arg)
{
  arg     = arg > 255u ? 255u : arg;
  float u = (float)arg * (1.0f / 255.0f);
  return u <= 0.04045f ? u * (25.0f / 323.0f) :
                         powf((200.0f * u + 11.0f) * (1.0f / 211.0f), 2.4f);
}

/* shaders/srgb.h:30-41, literal formula. */
uint32_t nvo_srgb_from_linear_bias_formula(float arg, float bias)
{
  float srgb = arg <= 0.0031308f ? (323.0f / 25.0f) * arg :
                                   1.055f * powf(arg, 1.0f / 2.4f) - 0.055f;
  float v = srgb * 255.0f + bias;
  v       = v < 0.f ? 0.f : (v > 255.f ? 255.f : v); /* glm::clamp = min(max(x,lo),hi) */
  return (uint32_t)v;
}
uint32_t nvo_srgb_from_linear_formula(float arg)
{
  return nvo_srgb_from_linear_bias_formula(arg, 0.5f);
}

/* Pinned-table versions (what Oracle A, Oracle B and the CUDA kernels use). */
float nvo_linear_from_srgb(uint32_t arg)
{
  arg = arg > 255u ? 255u : arg;
  return bits_to_float(NVPYR_SRGB_DECODE_BITS[arg]);
}

/* Number of thresholds <= x.  NVPYR_SRGB_ENCODE_THRESHOLD_BITS[c-1] is the
 * smallest float whose code is >= c (c = 1..255).  NaN encodes to 0 like
 * uint(clamp(NaN)) does with glm::clamp's comparisons on x86. */
uint32_t nvo_srgb_from_linear(float x)
{
  uint32_t lo = 0, hi = 255; /* answer in [lo, hi] */
  if(!(x == x))
    return 0;
  while(lo < hi)
  {
    uint32_t mid = (lo + hi + 1) >> 1; /* candidate code */
    if(x >= bits_to_float(NVPYR_SRGB_ENCODE_THRESHOLD_BITS[mid - 1]))
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

/* srgba8_mipmap_preamble.glsl:125 */
static inline uint32_t alpha_round(float a)
{
  float v = a * 255.0f + 0.5f;
  if(!(v > 0.f))
    return 0u; /* uint() of a negative float is undefined in GLSL; clamp */
  uint32_t u = v >= 4294967040.f ? 0xFFFFFFFFu : (uint32_t)v;
  return u > 255u ? 255u : u;
}

/* include/mipmap_storage.hpp:403-404 */
static inline uint32_t alpha_trunc(float a)
{
  float v = a * 255.0f;
  v       = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
  return (uint32_t)v;
}

/* ------------------------------------------------------------------------ */
/* Layout: include/mipmap_storage.hpp:53-76                                  */

uint32_t nvo_level_count(uint32_t w, uint32_t h)
{
  /* nvpro_pyramid_dispatch.hpp:122-132 (== MipmapStorage's level count) */
  uint32_t n = 0;
  while(w != 0 || h != 0)
  {
    w >>= 1;
    h >>= 1;
    ++n;
  }
  return n;
}

void nvo_level_dims(uint32_t w, uint32_t h, uint32_t level, uint32_t* lw, uint32_t* lh)
{
  uint32_t a = w >> level, b = h >> level;
  if(level >= 32)
    a = b = 0;
  *lw = a ? a : 1u;
  *lh = b ? b : 1u;
}

uint64_t nvo_level_offset(uint32_t w, uint32_t h, uint32_t level)
{
  uint64_t off = 0;
  for(uint32_t i = 0; i < level; ++i)
  {
    uint32_t lw, lh;
    nvo_level_dims(w, h, i, &lw, &lh);
    off += (uint64_t)lw * lh;
  }
  return off;
}

uint64_t nvo_chain_texels(uint32_t w, uint32_t h, uint32_t levels)
{
  return nvo_level_offset(w, h, levels);
}

/* ------------------------------------------------------------------------ */
/* Planner: nvpro_pyramid_dispatch.hpp:109-292                               */

typedef struct
{
  uint32_t pipeline;      /* 1 = fast, 0 = general */
  uint32_t input_level;   /* state.currentLevel */
  uint32_t level_count;   /* levels filled */
  uint32_t src_w, src_h;  /* state.currentX/Y */
  uint32_t workgroups;    /* vkCmdDispatch groupCountX */
  uint32_t push_constant; /* input_level << 5 | level_count */
  uint32_t bind;          /* vkCmdBindPipeline recorded before the dispatch */
  uint32_t barrier_after; /* vkCmdPipelineBarrier recorded after it */
} nvo_step;

/* nvproPyramidDefaultFastDispatcher<Div, Max>, dispatch.hpp:195-242 */
static uint32_t fast_dispatcher(uint32_t x, uint32_t y, uint32_t remaining, uint32_t div,
                                uint32_t max_levels, uint32_t* workgroups)
{
  if(!(x % div == 0u && y % div == 0u))
    return 0u;
  uint32_t cx = x, cy = y, levels = 0u;
  while(cx % 2u == 0u && cy % 2u == 0u && levels < remaining && levels < max_levels)
  {
    cx /= 2u;
    cy /= 2u;
    levels++;
  }
  const uint32_t shift   = levels > 5 ? 12u : 10u;
  const uint32_t mask    = levels > 5 ? 4095u : 1023u;
  const uint32_t samples = x * y; /* uint32 wrap is the reference's behaviour */
  *workgroups            = (samples + mask) >> shift;
  return levels;
}

/* nvproPyramidDefaultGeneralDispatcher, dispatch.hpp:247-292 */
static uint32_t general_dispatcher(uint32_t x, uint32_t y, uint32_t remaining, uint32_t* workgroups)
{
  const uint32_t levels = remaining >= 2u ? 2u : remaining;
  uint32_t       dw     = x >> levels;
  dw                    = dw ? dw : 1u;
  uint32_t dh           = y >> levels;
  dh                    = dh ? dh : 1u;
  if(levels == 1u)
    *workgroups = (dw * dh + 127u) / 128u;
  else
    *workgroups = ((dw + 7u) / 8u) * ((dh + 7u) / 8u);
  return levels;
}

/* Returns the number of steps, or -1 if max_steps is too small / bad input. */
int nvo_plan(uint32_t w, uint32_t h, uint32_t mip_levels, uint32_t have_fast, uint32_t fast_div,
             uint32_t fast_max_levels, nvo_step* steps, uint32_t max_steps)
{
  if(w == 0 || h == 0)
    return -1;
  if(mip_levels == 0)
    mip_levels = nvo_level_count(w, h);
  if(fast_div == 0)
    fast_div = 4;
  if(fast_max_levels == 0)
    fast_max_levels = 6;
  uint32_t level = 0, remaining = mip_levels - 1u, x = w, y = h;
  uint32_t fast_bound = 0, general_bound = 0; /* which pipeline is currently bound */
  int      n = 0;
  if(remaining == 0)
    return 0; /* reference would loop with remainingLevels==0: caller error */
  for(;;)
  {
    uint32_t done = 0, wg = 0;
    nvo_step s;
    memset(&s, 0, sizeof s);
    if(have_fast)
      done = fast_dispatcher(x, y, remaining, fast_div, fast_max_levels, &wg);
    if(done != 0)
    {
      s.pipeline    = 1;
      s.bind        = !fast_bound;
      fast_bound    = 1;
      general_bound = 0;
    }
    else
    {
      done          = general_dispatcher(x, y, remaining, &wg);
      s.pipeline    = 0;
      s.bind        = !general_bound;
      general_bound = 1;
      fast_bound    = 0;
    }
    s.input_level   = level;
    s.level_count   = done;
    s.src_w         = x;
    s.src_h         = y;
    s.workgroups    = wg;
    s.push_constant = level << 5 | done;
    level += done;
    remaining -= done;
    x >>= done;
    x = x ? x : 1u;
    y >>= done;
    y               = y ? y : 1u;
    s.barrier_after = remaining != 0u;
    if((uint32_t)n >= max_steps)
      return -1;
    steps[n++] = s;
    if(remaining == 0u)
      break;
  }
  return n;
}

/* ------------------------------------------------------------------------ */
/* Oracle A: shader-order emulation                                          */

typedef struct
{
  float x, y, z, w;
} vec4;
typedef struct
{
  int x, y;
} ivec2;

typedef struct
{
  int      fmt; /* 0 = sRGBA8, 1 = RGBA32F */
  uint32_t w, h, levels;
  uint8_t* u8;  /* packed chain, fmt 0 */
  float*   f32; /* packed chain, fmt 1 */
  uint64_t off[33];
  uint32_t lw[33], lh[33];
  uint64_t stores; /* number of NVPRO_PYRAMID_STORE executed */
  int      shared_f16; /* F16_SHARED build of the shaders (srgba8_mipmap_preamble.glsl:103-108) */
  int      shared_srgb; /* SRGB_SHARED build (srgba8_mipmap_preamble.glsl:60-101) */
} actx;

/* NVPRO_PYRAMID_SHARED_STORE followed by NVPRO_PYRAMID_SHARED_LOAD.  Default: the shared type is the value
 * type (nvpro_pyramid.glsl:204-206).  F16_SHARED: f16vec4(in_) then vec4(smem_), i.e. every component is
 * rounded to IEEE binary16 (round to nearest even) and widened again. */
/* SRGB_SHARED: srgbPack then srgbUnpack (srgba8_mipmap_preamble.glsl:81-99).  packUnorm4x8 of
 * srgbComponentFromLinear(x) == number of pinned SHARED thresholds <= x (monotone, checked exhaustively by
 * tools/gen_srgb_tables.c); unpackUnorm4x8 + linearFromSrgbComponent == the pinned SHARED decode table.  Alpha:
 * round-half-even(clamp(a, 0, 1) * 255) and code / 255.0 (IEEE division). */
uint32_t nvo_srgb_shared_pack(float x)
{
  uint32_t lo = 0, hi = 255;
  if(!(x == x))
    return 0;
  while(lo < hi)
  {
    uint32_t mid = (lo + hi + 1) >> 1;
    if(x >= bits_to_float(NVPYR_SRGB_SHARED_ENCODE_THRESHOLD_BITS[mid - 1]))
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}
float nvo_srgb_shared_unpack(uint32_t code)
{
  return bits_to_float(NVPYR_SRGB_SHARED_DECODE_BITS[code > 255u ? 255u : code]);
}
static inline float a_unorm8_round_trip(float a)
{
  float s = a < 0.f ? 0.f : (a > 1.f ? 1.f : a);
  if(!(a == a))
    s = 0.f;
  return rintf(s * 255.0f) / 255.0f;
}

static inline vec4 a_shared_round(const actx* c, vec4 v)
{
  if(c->shared_srgb)
  {
    v.x = nvo_srgb_shared_unpack(nvo_srgb_shared_pack(v.x));
    v.y = nvo_srgb_shared_unpack(nvo_srgb_shared_pack(v.y));
    v.z = nvo_srgb_shared_unpack(nvo_srgb_shared_pack(v.z));
    v.w = a_unorm8_round_trip(v.w);
  }
  if(c->shared_f16)
  {
    v.x = (float)(_Float16)v.x, v.y = (float)(_Float16)v.y;
    v.z = (float)(_Float16)v.z, v.w = (float)(_Float16)v.w;
  }
  return v;
}

static void actx_init(actx* c, int fmt, void* chain, uint32_t w, uint32_t h, uint32_t levels)
{
  memset(c, 0, sizeof *c);
  c->fmt    = fmt;
  c->w      = w;
  c->h      = h;
  c->levels = levels;
  c->u8     = (uint8_t*)chain;
  c->f32    = (float*)chain;
  uint64_t o = 0;
  for(uint32_t i = 0; i < levels && i < 33; ++i)
  {
    nvo_level_dims(w, h, i, &c->lw[i], &c->lh[i]);
    c->off[i] = o;
    o += (uint64_t)c->lw[i] * c->lh[i];
  }
}

static inline vec4 v_add(vec4 a, vec4 b)
{
  vec4 r = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
  return r;
}
static inline vec4 v_scale(float s, vec4 a)
{
  vec4 r = {s * a.x, s * a.y, s * a.z, s * a.w};
  return r;
}

/* NVPRO_PYRAMID_LOAD: srgba8_mipmap_preamble.glsl:21-22 (texelFetch on the
 * sRGB view == decode table; alpha as shaders/srgb.h:60). */
static inline vec4 a_load(const actx* c, ivec2 p, int level)
{
  uint64_t i = c->off[level] + (uint64_t)p.y * c->lw[level] + (uint64_t)p.x;
  vec4     r;
  if(c->fmt == 0)
  {
    const uint8_t* t = c->u8 + 4 * i;
    r.x              = nvo_linear_from_srgb(t[0]);
    r.y              = nvo_linear_from_srgb(t[1]);
    r.z              = nvo_linear_from_srgb(t[2]);
    r.w              = (float)t[3] * (1.0f / 255.0f);
  }
  else
  {
    const float* t = c->f32 + 4 * i;
    r.x = t[0], r.y = t[1], r.z = t[2], r.w = t[3];
  }
  return r;
}

/* NVPRO_PYRAMID_STORE: srgba8_mipmap_preamble.glsl:27-28, :122-128 */
static inline void a_store(actx* c, ivec2 p, int level, vec4 v)
{
  uint64_t i = c->off[level] + (uint64_t)p.y * c->lw[level] + (uint64_t)p.x;
  c->stores++;
  if(c->fmt == 0)
  {
    uint8_t* t = c->u8 + 4 * i;
    t[0]       = (uint8_t)nvo_srgb_from_linear(v.x);
    t[1]       = (uint8_t)nvo_srgb_from_linear(v.y);
    t[2]       = (uint8_t)nvo_srgb_from_linear(v.z);
    t[3]       = (uint8_t)alpha_round(v.w);
  }
  else
  {
    float* t = c->f32 + 4 * i;
    t[0] = v.x, t[1] = v.y, t[2] = v.z, t[3] = v.w;
  }
}

/* srgba8_mipmap_preamble.glsl:24-25: out_ = a0 * v0 + a1 * v1 + a2 * v2.  GLSL lets the
 * implementation contract this (no `precise`), and a GPU compiler does: one multiply and two
 * fused multiply-adds, left to right.  Pinned to exactly that form (fmaf is correctly rounded):
 *   fma(a2, v2, fma(a1, v1, a0 * v0)) */
static inline float a_reduce1(float a0, float v0, float a1, float v1, float a2, float v2)
{
  return fmaf(a2, v2, fmaf(a1, v1, a0 * v0));
}
static inline vec4 a_reduce(float a0, vec4 v0, float a1, vec4 v1, float a2, vec4 v2)
{
  vec4 r = {a_reduce1(a0, v0.x, a1, v1.x, a2, v2.x), a_reduce1(a0, v0.y, a1, v1.y, a2, v2.y),
            a_reduce1(a0, v0.z, a1, v1.z, a2, v2.z), a_reduce1(a0, v0.w, a1, v1.w, a2, v2.w)};
  return r;
}
/* :35 */
static inline vec4 a_reduce2(vec4 v0, vec4 v1)
{
  return v_scale(0.5f, v_add(v0, v1));
}
/* :37-38 */
static inline vec4 a_reduce4(vec4 v00, vec4 v01, vec4 v10, vec4 v11)
{
  return v_scale(0.25f, v_add(v_add(v00, v01), v_add(v10, v11)));
}
/* nvpro_pyramid.glsl:179-189 (software LOAD_REDUCE4) */
static inline vec4 a_load_reduce4(const actx* c, ivec2 p, int level)
{
  ivec2 p01 = {p.x, p.y + 1}, p10 = {p.x + 1, p.y}, p11 = {p.x + 1, p.y + 1};
  vec4  v00 = a_load(c, p, level);
  vec4  v01 = a_load(c, p01, level);
  vec4  v10 = a_load(c, p10, level);
  vec4  v11 = a_load(c, p11, level);
  return a_reduce4(v00, v01, v10, v11);
}

/* ---- fast pipeline ------------------------------------------------------ */

#define FAST_WG 256

typedef struct
{
  int      active[FAST_WG]; /* invocation reached handleTile_ */
  vec4     out[FAST_WG];
  ivec2    dstSubTile[FAST_WG];
  int      dstLevel[FAST_WG];
  int      returned[FAST_WG];
  vec4     sharedTile[16];
} fast_wg_state;

/* handleTile_, nvpro_pyramid.glsl:272-398, executed in lock step by all
 * active invocations of one work group.  gl_SubgroupInvocationID is taken as
 * gl_LocalInvocationIndex & 31 (32-wide subgroups laid out linearly). */
static void a_handle_tile(actx* c, fast_wg_state* s, const ivec2* srcTileOffset, int inputLevel,
                          uint32_t levelCount, int sharedMemoryWrite, const uint32_t* sharedMemoryIdx)
{
  const uint32_t teamMask = levelCount >= 3 ? 15u : levelCount == 2 ? 3u : 0u;
  vec4           nxt[FAST_WG];

  for(uint32_t l = 0; l < FAST_WG; ++l)
  {
    s->returned[l] = !s->active[l];
    if(!s->active[l])
      continue;
    const uint32_t idxInTeam = l & teamMask;
    int            dstLevel  = inputLevel + 1;
    ivec2          dstSubTile;
    vec4           out;
    if(sharedMemoryWrite && levelCount == 4)
    { /* :305-344 */
      uint32_t xo         = (idxInTeam & 1) << 2 | (idxInTeam & 4) << 1;
      uint32_t yo         = (idxInTeam & 2) << 1 | (idxInTeam & 8);
      ivec2    srcSubTile = {srcTileOffset[l].x + (int)xo, srcTileOffset[l].y + (int)yo};
      dstSubTile.x        = srcSubTile.x >> 1;
      dstSubTile.y        = srcSubTile.y >> 1;
      ivec2 sc, dc;
      sc          = srcSubTile;
      dc          = dstSubTile;
      vec4 s00    = a_load_reduce4(c, sc, inputLevel);
      a_store(c, dc, dstLevel, s00);
      sc.x = srcSubTile.x, sc.y = srcSubTile.y + 2;
      dc.x = dstSubTile.x, dc.y = dstSubTile.y + 1;
      vec4 s01 = a_load_reduce4(c, sc, inputLevel);
      a_store(c, dc, dstLevel, s01);
      sc.x = srcSubTile.x + 2, sc.y = srcSubTile.y;
      dc.x = dstSubTile.x + 1, dc.y = dstSubTile.y;
      vec4 s10 = a_load_reduce4(c, sc, inputLevel);
      a_store(c, dc, dstLevel, s10);
      sc.x = srcSubTile.x + 2, sc.y = srcSubTile.y + 2;
      dc.x = dstSubTile.x + 1, dc.y = dstSubTile.y + 1;
      vec4 s11 = a_load_reduce4(c, sc, inputLevel);
      a_store(c, dc, dstLevel, s11);
      dstLevel++;
      dstSubTile.x >>= 1;
      dstSubTile.y >>= 1;
      out = a_reduce4(s00, s01, s10, s11);
      a_store(c, dstSubTile, dstLevel, out);
    }
    else
    { /* :345-357 */
      uint32_t xo         = (idxInTeam & 1) << 1 | (idxInTeam & 4);
      uint32_t yo         = (idxInTeam & 2) | (idxInTeam & 8) >> 1;
      ivec2    srcSubTile = {srcTileOffset[l].x + (int)xo, srcTileOffset[l].y + (int)yo};
      dstSubTile.x        = srcSubTile.x >> 1;
      dstSubTile.y        = srcSubTile.y >> 1;
      out                 = a_load_reduce4(c, srcSubTile, inputLevel);
      a_store(c, dstSubTile, dstLevel, out);
    }
    s->out[l]        = out;
    s->dstSubTile[l] = dstSubTile;
    s->dstLevel[l]   = dstLevel;
    if(!sharedMemoryWrite && levelCount == 1)
      s->returned[l] = 1; /* :359 */
  }

  /* :361-375  shuffle xor 1,2,3 */
  memcpy(nxt, s->out, sizeof nxt);
  for(uint32_t l = 0; l < FAST_WG; ++l)
  {
    if(s->returned[l])
      continue;
    s->dstLevel[l]++;
    s->dstSubTile[l].x >>= 1;
    s->dstSubTile[l].y >>= 1;
    vec4 s00 = s->out[l];
    vec4 s01 = s->out[l ^ 1];
    vec4 s10 = s->out[l ^ 2];
    vec4 s11 = s->out[l ^ 3];
    if(0 == ((l & 31u) & 3u))
    {
      nxt[l] = a_reduce4(s00, s01, s10, s11);
      a_store(c, s->dstSubTile[l], s->dstLevel[l], nxt[l]);
    }
  }
  memcpy(s->out, nxt, sizeof nxt);
  for(uint32_t l = 0; l < FAST_WG; ++l)
    if(!s->returned[l] && !sharedMemoryWrite && levelCount == 2)
      s->returned[l] = 1; /* :377 */

  /* :379-397  shuffle xor 4,8,12 */
  for(uint32_t l = 0; l < FAST_WG; ++l)
  {
    if(s->returned[l])
      continue;
    s->dstLevel[l]++;
    s->dstSubTile[l].x >>= 1;
    s->dstSubTile[l].y >>= 1;
    vec4 s00 = s->out[l];
    vec4 s01 = s->out[l ^ 4];
    vec4 s10 = s->out[l ^ 8];
    vec4 s11 = s->out[l ^ 12];
    if(0 == ((l & 31u) & 15u))
    {
      nxt[l] = a_reduce4(s00, s01, s10, s11);
      a_store(c, s->dstSubTile[l], s->dstLevel[l], nxt[l]);
      if(sharedMemoryWrite)
        s->sharedTile[sharedMemoryIdx[l]] = a_shared_round(c, nxt[l]);
    }
  }
  memcpy(s->out, nxt, sizeof nxt);
}

/* nvproPyramidMain (fast), nvpro_pyramid.glsl:401-532, one work group. */
static void a_fast_workgroup(actx* c, uint32_t wg, uint32_t pc)
{
  static _Thread_local fast_wg_state s;
  const int      levelCount0 = (int)(pc & 31u);
  const int      inputLevel0 = (int)(pc >> 5u);
  const uint32_t srcW = c->lw[inputLevel0], srcH = c->lh[inputLevel0];
  const uint32_t horizontalTiles = srcW >> levelCount0;
  const uint32_t verticalTiles   = srcH >> levelCount0;
  uint32_t       teamSizeLog2    = (uint32_t)levelCount0 * 2u - 2u;
  teamSizeLog2                   = teamSizeLog2 < 8u ? teamSizeLog2 : 8u;

  ivec2    tileOffset[FAST_WG];
  uint32_t smemIdx[FAST_WG];
  int      levelCount[FAST_WG], inputLevel[FAST_WG];

  memset(&s, 0, sizeof s);
  for(uint32_t l = 0; l < FAST_WG; ++l)
  {
    uint32_t g     = wg * FAST_WG + l;
    uint32_t tile  = g >> teamSizeLog2;
    uint32_t hi    = tile % horizontalTiles;
    uint32_t vi    = tile / horizontalTiles;
    tileOffset[l].x = (int)(hi << levelCount0);
    tileOffset[l].y = (int)(vi << levelCount0);
    s.active[l]     = vi < verticalTiles;
    levelCount[l]   = levelCount0;
    inputLevel[l]   = inputLevel0;
    smemIdx[l]      = 0;
  }

  if(levelCount0 <= 3)
  { /* :422-431 */
    a_handle_tile(c, &s, tileOffset, inputLevel0, (uint32_t)levelCount0, 0, smemIdx);
    return;
  }

  /* :437-465 */
  const int subLevelCount = levelCount0 == 6 ? 4 : 3;
  const int subTeamMask   = levelCount0 == 4 ? 3 : 15;
  for(uint32_t l = 0; l < FAST_WG; ++l)
  {
    uint32_t g          = wg * FAST_WG + l;
    int      subTeamIdx = (int)(g >> 4) & subTeamMask;
    ivec2    o;
    o.x = (subTeamIdx & 1) << 3 | (subTeamIdx & 4) << 2;
    o.y = (subTeamIdx & 2) << 2 | (subTeamIdx & 8) << 1;
    if(subLevelCount == 4)
    {
      o.x <<= 1;
      o.y <<= 1;
    }
    tileOffset[l].x += o.x;
    tileOffset[l].y += o.y;
    smemIdx[l] = (g >> 4u) & 15u;
    if(s.active[l])
    {
      inputLevel[l] += subLevelCount;
      levelCount[l] -= subLevelCount;
    }
  }
  a_handle_tile(c, &s, tileOffset, inputLevel0, (uint32_t)subLevelCount, 1, smemIdx);

  /* barrier(); :468 -- then :472-531 */
  vec4 out[4];
  int  did[4] = {0, 0, 0, 0};
  for(uint32_t l = 0; l < 4; ++l)
  {
    if(levelCount[l] == 1)
    { /* :475-494 */
      uint32_t tile = wg * 4 + l;
      uint32_t hi = tile % horizontalTiles, vi = tile / horizontalTiles;
      ivec2    to = {(int)hi, (int)vi};
      uint32_t so = l * 4u;
      if(vi < verticalTiles)
      {
        vec4 in00 = s.sharedTile[so + 0u];
        vec4 in10 = s.sharedTile[so + 1u];
        vec4 in01 = s.sharedTile[so + 2u];
        vec4 in11 = s.sharedTile[so + 3u];
        vec4 o    = a_reduce4(in00, in01, in10, in11);
        a_store(c, to, inputLevel[l] + 1, o);
      }
    }
    else
    { /* :495-530 (first half) */
      uint32_t tile = wg;
      uint32_t hi = tile % horizontalTiles, vi = tile / horizontalTiles;
      uint32_t so = l * 4u;
      if(vi < verticalTiles)
      {
        vec4 in00 = s.sharedTile[so + 0u];
        vec4 in10 = s.sharedTile[so + 1u];
        vec4 in01 = s.sharedTile[so + 2u];
        vec4 in11 = s.sharedTile[so + 3u];
        out[l]    = a_reduce4(in00, in01, in10, in11);
        ivec2 p   = {(int)hi * 2 + (int)(l & 1), (int)vi * 2 + (int)((l & 2) >> 1)};
        a_store(c, p, inputLevel[l] + 1, out[l]);
        did[l] = 1;
      }
    }
  }
  /* :518-528 shuffle among invocations 0..3, invocation 0 stores */
  if(did[0])
  {
    uint32_t tile = wg;
    uint32_t hi = tile % horizontalTiles, vi = tile / horizontalTiles;
    ivec2    to   = {(int)hi, (int)vi};
    vec4     in00 = out[0], in10 = out[1], in01 = out[2], in11 = out[3];
    vec4     o = a_reduce4(in00, in01, in10, in11);
    a_store(c, to, inputLevel[0] + 2, o);
  }
}

/* ---- general pipeline ---------------------------------------------------- */

typedef struct
{
  vec4 sharedLevel[17][17]; /* [y][x] */
} gen_wg_state;

/* kernelSizeFromInputSize_, nvpro_pyramid.glsl:557-561 */
static inline ivec2 a_kernel_size(uint32_t w, uint32_t h)
{
  ivec2 k = {w == 1 ? 1 : (int)(2 | (w & 1)), h == 1 ? 1 : (int)(2 | (h & 1))};
  return k;
}

/* loadSample_, :658-671 */
static inline vec4 a_load_sample(const actx* c, const gen_wg_state* s, ivec2 p, int level, int fromShared)
{
  if(fromShared)
    return s->sharedLevel[p.y][p.x];
  return a_load(c, p, level);
}

/* reduceStoreSample_, :575-656 */
static vec4 a_reduce_store_sample(actx* c, const gen_wg_state* s, ivec2 src, int srcLevel, int lfs,
                                  ivec2 kernelSize, ivec2 dstImageSize, ivec2 dst, int dstLevel)
{
  float n   = (float)dstImageSize.y;
  float rcp = 1.0f / (2 * n + 1);
  float w0  = rcp * (n - (float)dst.y);
  float w1  = rcp * n;
  float w2  = 1.0f - w0 - w1;
  vec4  v0 = {0, 0, 0, 0}, v1 = v0, v2 = v0, h[3] = {v0, v0, v0}, out = v0;

  /* columns are visited 2,1,0 (switch fall-through); order is irrelevant to
   * the values */
  for(int col = kernelSize.x - 1; col >= 0; --col)
  {
    ivec2 p;
    p.x = src.x + col;
    if(kernelSize.y >= 3)
    {
      p.y = src.y + 2;
      v2  = a_load_sample(c, s, p, srcLevel, lfs);
    }
    if(kernelSize.y >= 2)
    {
      p.y = src.y + 1;
      v1  = a_load_sample(c, s, p, srcLevel, lfs);
    }
    p.y = src.y;
    v0  = a_load_sample(c, s, p, srcLevel, lfs);
    switch(kernelSize.y)
    {
      case 3: h[col] = a_reduce(w0, v0, w1, v1, w2, v2); break;
      case 2: h[col] = a_reduce2(v0, v1); break;
      default: h[col] = v0; break;
    }
  }
  switch(kernelSize.x)
  {
    case 3:
      n   = (float)dstImageSize.x;
      rcp = 1.0f / (2 * n + 1);
      w0  = rcp * (n - (float)dst.x);
      w1  = rcp * n;
      w2  = 1.0f - w0 - w1;
      out = a_reduce(w0, h[0], w1, h[1], w2, h[2]);
      break;
    case 2: out = a_reduce2(h[0], h[1]); break;
    default: out = h[0]; break;
  }
  a_store(c, dst, dstLevel, out);
  return out;
}

/* nvproPyramidMain (general), :824-882, one work group of 128 invocations. */
static void a_general_workgroup(actx* c, uint32_t wg, uint32_t pc)
{
  static _Thread_local gen_wg_state s;
  const int levelCount = (int)(pc & 31u);
  const int inputLevel = (int)(pc >> 5u);

  if(levelCount == 1)
  { /* :828-842 */
    ivec2 kernelSize = a_kernel_size(c->lw[inputLevel], c->lh[inputLevel]);
    ivec2 dstSize    = {(int)c->lw[inputLevel + 1], (int)c->lh[inputLevel + 1]};
    for(uint32_t l = 0; l < 128; ++l)
    {
      int   g   = (int)(wg * 128u + l);
      ivec2 dst = {g % dstSize.x, g / dstSize.x};
      ivec2 src = {dst.x * 2, dst.y * 2};
      if(dst.y < dstSize.y)
        a_reduce_store_sample(c, &s, src, inputLevel, 0, kernelSize, dstSize, dst, inputLevel + 1);
    }
    return;
  }

  /* :843-881 */
  const int level1 = inputLevel + 1, level2 = inputLevel + 2;
  ivec2     level2Size = {(int)c->lw[level2], (int)c->lh[level2]};
  ivec2     tileCount  = {(int)((uint32_t)(level2Size.x + 7) / 8u), (int)((uint32_t)(level2Size.y + 7) / 8u)};
  ivec2     tileIdx    = {(int)(wg % (uint32_t)tileCount.x), (int)(wg / (uint32_t)tileCount.x)};
  const int boundsCheck = tileIdx.x >= tileCount.x - 1 || tileIdx.y >= tileCount.y - 1;

  /* fillIntermediateTile_, :732-781 */
  {
    ivec2 dstTile    = {tileIdx.x * 16, tileIdx.y * 16};
    ivec2 dstSize    = {(int)c->lw[level1], (int)c->lh[level1]};
    ivec2 futureK    = a_kernel_size(c->lw[level1], c->lh[level1]);
    ivec2 kernelSize = a_kernel_size(c->lw[inputLevel], c->lh[inputLevel]);
    for(uint32_t l = 0; l < 128; ++l)
    {
      ivec2 init, step;
      int   iterations;
      if(futureK.x == 3)
      {
        if(futureK.y == 3)
        {
          init.x = (int)(l % 17u), init.y = (int)(l / 17u);
          step.x = 0, step.y = 7;
          iterations = l >= 7 * 17 ? 0 : l < 3 * 17 ? 3 : 2;
        }
        else
        {
          init.x = (int)(l / 16u), init.y = (int)(l % 16u);
          step.x = 8, step.y = 0;
          iterations = l < 16 ? 3 : 2;
        }
      }
      else
      {
        init.x = (int)(l % 16u), init.y = (int)(l / 16u);
        step.x = 0, step.y = 8;
        iterations = (futureK.y == 3) ? (l < 16 ? 3 : 2) : 2;
      }
      /* intermediateLevelLoop_, :685-722 */
      ivec2 dst = {dstTile.x + init.x, dstTile.y + init.y};
      ivec2 sh  = init;
      for(int i = 0; i < iterations; ++i, dst.x += step.x, dst.y += step.y, sh.x += step.x, sh.y += step.y)
      {
        ivec2 src = {dst.x * 2, dst.y * 2};
        if(boundsCheck)
        {
          /* NOTE: the GLSL `continue` skips the increments (:706-707 jump to
           * ++i_ without dstCoord_ += step_), so once a coordinate is out of
           * bounds every later iteration of this invocation is too. */
          if((uint32_t)dst.x >= (uint32_t)dstSize.x || (uint32_t)dst.y >= (uint32_t)dstSize.y)
            break;
        }
        vec4 v = a_reduce_store_sample(c, &s, src, inputLevel, 0, kernelSize, dstSize, dst, level1);
        s.sharedLevel[sh.y][sh.x] = a_shared_round(c, v);
      }
    }
  }
  /* barrier(); fillLastTile_, :791-820 */
  {
    ivec2 dstTile    = {tileIdx.x * 8, tileIdx.y * 8};
    ivec2 srcSize    = {(int)c->lw[level1], (int)c->lh[level1]};
    ivec2 kernelSize = a_kernel_size((uint32_t)srcSize.x, (uint32_t)srcSize.y);
    for(uint32_t l = 0; l < 64; ++l)
    {
      ivec2 to  = {(int)(l % 8u), (int)(l / 8u)};
      ivec2 ssc = {to.x * 2, to.y * 2};
      ivec2 dst = {to.x + dstTile.x, to.y + dstTile.y};
      int   inb = 1;
      if(boundsCheck)
        inb = ((uint32_t)dst.x < (uint32_t)level2Size.x) && ((uint32_t)dst.y < (uint32_t)level2Size.y);
      if(inb)
        a_reduce_store_sample(c, &s, ssc, 0, 1, kernelSize, level2Size, dst, level2);
    }
  }
}

/* ---- blit fallback -------------------------------------------------------- */
/* demo_app/mipmap_pipelines.cpp:404-441: when the fast dispatcher declines, ONE level is filled by
 * vkCmdBlitImage(level -> level + 1, whole extents, VK_FILTER_LINEAR).  Vulkan's rule for a scaled blit: the centre of
 * destination texel (i, j) maps to source coordinates u = (i + 0.5) * srcW / dstW, v likewise, sampled with an
 * unnormalised clamp-to-edge linear filter (texels floor(u - 0.5), floor(u - 0.5) + 1; weight of the second =
 * frac(u - 0.5)).  The precision of that arithmetic is implementation-defined in Vulkan and no Vulkan device exists
 * here to pin it bit for bit: PARITY UNPINNED at that level.  Pinned instead: the worst deltas against the CPU generator
 * that the reference recorded for these alternatives on its 13 test images (demo_app/rtx3090.json) -- this restatement
 * reproduces them exactly on 9 images and within 3 code values on all (tests/test_blit.py) -- and OUR arithmetic
 * contract, restated here: float32,
 * scale by IEEE division, u = (i + 0.5) * scale - 0.5 in two roundings, lerps through NVPRO_PYRAMID_REDUCE as
 * reduce(1 - a, p, a, q, 0, q), rows first. */
static void a_blit_tap(uint32_t i, float scale, uint32_t src_size, int* i0, int* i1, float* a)
{
  float u    = ((float)i + 0.5f) * scale;
  u          = u - 0.5f;
  float f    = floorf(u);
  *a         = u - f;
  int k      = (int)f, last = (int)src_size - 1;
  *i0        = k < 0 ? 0 : (k > last ? last : k);
  *i1        = k + 1 < 0 ? 0 : (k + 1 > last ? last : k + 1);
}
static void a_blit_level(actx* c, int src_level)
{
  const uint32_t sw = c->lw[src_level], sh = c->lh[src_level], dw = c->lw[src_level + 1], dh = c->lh[src_level + 1];
  const float    sx = (float)sw / (float)dw, sy = (float)sh / (float)dh;
  for(uint32_t y = 0; y < dh; ++y)
    for(uint32_t x = 0; x < dw; ++x)
    {
      int   x0, x1, y0, y1;
      float a, b;
      a_blit_tap(x, sx, sw, &x0, &x1, &a);
      a_blit_tap(y, sy, sh, &y0, &y1, &b);
      ivec2 p00 = {x0, y0}, p10 = {x1, y0}, p01 = {x0, y1}, p11 = {x1, y1}, d = {(int)x, (int)y};
      vec4  t00 = a_load(c, p00, src_level), t10 = a_load(c, p10, src_level);
      vec4  t01 = a_load(c, p01, src_level), t11 = a_load(c, p11, src_level);
      float ia = 1.0f - a, ib = 1.0f - b;
      vec4  top = a_reduce(ia, t00, a, t10, 0.0f, t10), bot = a_reduce(ia, t01, a, t11, 0.0f, t11);
      a_store(c, d, src_level + 1, a_reduce(ib, top, b, bot, 0.0f, bot));
    }
}

/* Whole chain in shader order.  fmt 0: chain = uint8 RGBA; fmt 1: float RGBA.
 * flags bit0: force general pipeline (no fast pipeline available); bit1: F16_SHARED build; bit2: SRGB_SHARED build;
 * bit3: the blit fallback replaces the general pipeline (demo_app alternatives "generalblit", with bit0 "blit").
 * Returns number of dispatches, <0 on error.  stores_out (optional) receives
 * the number of texel stores executed (coverage accounting). */
int nvo_shader_chain(int fmt, void* chain, uint32_t w, uint32_t h, uint32_t mip_levels, uint32_t flags,
                     uint32_t fast_div, uint32_t fast_max_levels, uint64_t* stores_out)
{
  nvo_step steps[40];
  if(w == 0 || h == 0 || !chain)
    return -1;
  if(mip_levels == 0)
    mip_levels = nvo_level_count(w, h);
  if(mip_levels > 32)
    return -1;
  actx c;
  actx_init(&c, fmt, chain, w, h, mip_levels);
  c.shared_f16  = (flags & 2u) != 0;
  c.shared_srgb = (flags & 4u) != 0;
  if(flags & 8u)
  {
    /* the loop of demo_app/mipmap_pipelines.cpp:376-453 */
    uint32_t level = 0, remaining = mip_levels - 1u, x = w, y = h;
    int      count = 0;
    while(remaining != 0u)
    {
      uint32_t wg = 0, done = 0;
      if(!(flags & 1u))
        done = fast_dispatcher(x, y, remaining, fast_div ? fast_div : 4u, fast_max_levels ? fast_max_levels : 6u, &wg);
      if(done != 0u)
      {
        for(uint32_t g = 0; g < wg; ++g)
          a_fast_workgroup(&c, g, level << 5 | done);
      }
      else
      {
        a_blit_level(&c, (int)level);
        done = 1u;
      }
      level += done;
      remaining -= done;
      x = (x >> done) ? (x >> done) : 1u;
      y = (y >> done) ? (y >> done) : 1u;
      ++count;
    }
    if(stores_out)
      *stores_out = c.stores;
    return count;
  }
  int n = nvo_plan(w, h, mip_levels, !(flags & 1u), fast_div, fast_max_levels, steps, 40);
  if(n < 0)
    return n;
  for(int i = 0; i < n; ++i)
  {
    for(uint32_t wg = 0; wg < steps[i].workgroups; ++wg)
    {
      if(steps[i].pipeline)
        a_fast_workgroup(&c, wg, steps[i].push_constant);
      else
        a_general_workgroup(&c, wg, steps[i].push_constant);
    }
  }
  if(stores_out)
    *stores_out = c.stores;
  return n;
}

/* One dispatch only (used by tests that exercise a single push constant). */
int nvo_shader_dispatch(int fmt, void* chain, uint32_t w, uint32_t h, uint32_t mip_levels, uint32_t pipeline,
                        uint32_t push_constant, uint32_t workgroups)
{
  actx c;
  if(mip_levels == 0)
    mip_levels = nvo_level_count(w, h);
  actx_init(&c, fmt, chain, w, h, mip_levels);
  for(uint32_t wg = 0; wg < workgroups; ++wg)
  {
    if(pipeline)
      a_fast_workgroup(&c, wg, push_constant);
    else
      a_general_workgroup(&c, wg, push_constant);
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Oracle B: include/mipmap_storage.hpp:207-414                              */

typedef struct
{
  float c[4];
} sample4;

static inline void b_mac(sample4* lhs, const sample4* rhs, float wt)
{
  for(int c = 0; c < 4; ++c)
    lhs->c[c] += rhs->c[c] * wt;
}

/* generateLevel<WE,HE>, :246-392, for one level of an 8-bit sRGBA chain
 * (fmt 0) or a float chain (fmt 1; toLinear/fromLinear = identity).
 * Rows [y0, y1) only, so callers can split rows across threads. */
static void b_generate_level_rows(int fmt, void* chain, uint32_t w, uint32_t h, uint32_t level, uint32_t y0,
                                  uint32_t y1)
{
  uint32_t sw, sh, dw, dh;
  nvo_level_dims(w, h, level - 1, &sw, &sh);
  nvo_level_dims(w, h, level, &dw, &dh);
  const uint64_t so = nvo_level_offset(w, h, level - 1), dof = nvo_level_offset(w, h, level);
  const int      WE = !(sw & 1), HE = !(sh & 1);
  const uint8_t* src8 = (const uint8_t*)chain + 4 * so;
  uint8_t*       dst8 = (uint8_t*)chain + 4 * dof;
  const float*   srcf = (const float*)chain + 4 * so;
  float*         dstf = (float*)chain + 4 * dof;

  for(uint32_t y = y0; y < y1; ++y)
  {
    for(uint32_t x = 0; x < dw; ++x)
    {
      sample4 s[3][3]; /* s[xOffset][yOffset] */
      memset(s, 0, sizeof s);
#define B_LOAD(xo, yo)                                                                                       \
  do                                                                                                         \
  {                                                                                                          \
    uint64_t i_ = (uint64_t)(2 * x + (xo)) + (uint64_t)sw * (2 * y + (yo));                                  \
    if(fmt == 0)                                                                                             \
    {                                                                                                        \
      const uint8_t* t_ = src8 + 4 * i_;                                                                     \
      s[xo][yo].c[0]    = nvo_linear_from_srgb(t_[0]);                                                       \
      s[xo][yo].c[1]    = nvo_linear_from_srgb(t_[1]);                                                       \
      s[xo][yo].c[2]    = nvo_linear_from_srgb(t_[2]);                                                       \
      s[xo][yo].c[3]    = (float)t_[3] * (1.f / 255.f);                                                      \
    }                                                                                                        \
    else                                                                                                     \
      memcpy(s[xo][yo].c, srcf + 4 * i_, 16);                                                                \
  } while(0)

      B_LOAD(0, 0);
      if(WE || sw != 1)
        B_LOAD(1, 0);
      if(HE || sh != 1)
      {
        B_LOAD(0, 1);
        if(WE || sw != 1)
          B_LOAD(1, 1);
      }
      if(!WE && sw != 1)
      {
        B_LOAD(2, 0);
        /* The reference reads (2,1) unconditionally here (:299); for sh == 1
         * that texel does not exist and its value is never used. */
        if(HE || sh != 1)
          B_LOAD(2, 1);
        if(!HE && sh != 1)
          B_LOAD(2, 2);
      }
      if(!HE && sh != 1)
      {
        B_LOAD(0, 2);
        if(WE || sw != 1) /* same remark for (1,2) when sw == 1 (:306) */
          B_LOAD(1, 2);
      }
#undef B_LOAD

      sample4 s0, s1, s2, r;
      memset(&s0, 0, sizeof s0);
      s1 = s2 = r = s0;
      if(HE)
      {
        b_mac(&s0, &s[0][0], 0.5f);
        b_mac(&s0, &s[0][1], 0.5f);
        b_mac(&s1, &s[1][0], 0.5f);
        b_mac(&s1, &s[1][1], 0.5f);
        if(!WE)
        {
          b_mac(&s2, &s[2][0], 0.5f);
          b_mac(&s2, &s[2][1], 0.5f);
        }
      }
      else if(sh == 1)
      {
        s0 = s[0][0];
        if(WE || sw != 1)
          s1 = s[1][0];
        if(!WE)
          s2 = s[2][0];
      }
      else
      {
        const float n   = (float)dh;
        const float rcp = 1.0f / (2 * n + 1);
        const float w0  = rcp * (n - (float)y);
        const float w1  = rcp * n;
        const float w2  = rcp * (float)(1 + y);
        b_mac(&s0, &s[0][0], w0);
        b_mac(&s0, &s[0][1], w1);
        b_mac(&s0, &s[0][2], w2);
        if(WE || sw != 1)
        {
          b_mac(&s1, &s[1][0], w0);
          b_mac(&s1, &s[1][1], w1);
          b_mac(&s1, &s[1][2], w2);
        }
        if(!WE)
        {
          b_mac(&s2, &s[2][0], w0);
          b_mac(&s2, &s[2][1], w1);
          b_mac(&s2, &s[2][2], w2);
        }
      }

      if(WE)
      {
        b_mac(&r, &s0, 0.5f);
        b_mac(&r, &s1, 0.5f);
      }
      else if(sw == 1)
      {
        r = s0;
      }
      else
      {
        const float n   = (float)dw;
        const float rcp = 1.0f / (2 * n + 1);
        const float w0  = rcp * (n - (float)x);
        const float w1  = rcp * n;
        const float w2  = rcp * (float)(1 + x);
        b_mac(&r, &s0, w0);
        b_mac(&r, &s1, w1);
        b_mac(&r, &s2, w2);
      }

      uint64_t o = (uint64_t)dw * y + x;
      if(fmt == 0)
      {
        uint8_t* t = dst8 + 4 * o;
        t[0]       = (uint8_t)nvo_srgb_from_linear(r.c[0]);
        t[1]       = (uint8_t)nvo_srgb_from_linear(r.c[1]);
        t[2]       = (uint8_t)nvo_srgb_from_linear(r.c[2]);
        t[3]       = (uint8_t)alpha_trunc(r.c[3]);
      }
      else
        memcpy(dstf + 4 * o, r.c, 16);
    }
  }
}

/* cpuGenerateMipmaps_sRGBA, :395-414 (fmt 0), single-threaded as the
 * reference runs it. */
int nvo_cpu_chain(int fmt, void* chain, uint32_t w, uint32_t h)
{
  if(!chain || w == 0 || h == 0)
    return -1;
  uint32_t levels = nvo_level_count(w, h);
  for(uint32_t level = 1; level < levels; ++level)
  {
    uint32_t dw, dh;
    nvo_level_dims(w, h, level, &dw, &dh);
    b_generate_level_rows(fmt, chain, w, h, level, 0, dh);
  }
  return (int)levels;
}

/* Row-range entry point so a harness can spread one level over host threads
 * (levels remain serially dependent). */
int nvo_cpu_level_rows(int fmt, void* chain, uint32_t w, uint32_t h, uint32_t level, uint32_t y0, uint32_t y1)
{
  b_generate_level_rows(fmt, chain, w, h, level, y0, y1);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Comparator: MipmapStorage::compare, mipmap_storage.hpp:159-202            */

typedef struct
{
  uint32_t worst_delta;
  uint32_t x, y, level, channel;
  uint64_t mismatched_texels; /* texels (levels >= 1) with any channel differing */
  uint64_t compared_texels;
} nvo_compare_result;

void nvo_compare_srgba8(const uint8_t* a, const uint8_t* b, uint32_t w, uint32_t h, uint32_t levels,
                        nvo_compare_result* out)
{
  nvo_compare_result r;
  memset(&r, 0, sizeof r);
  if(levels == 0)
    levels = nvo_level_count(w, h);
  for(uint32_t level = 1; level < levels; ++level)
  {
    uint32_t lw, lh;
    nvo_level_dims(w, h, level, &lw, &lh);
    uint64_t o = nvo_level_offset(w, h, level);
    for(uint32_t y = 0; y < lh; ++y)
      for(uint32_t x = 0; x < lw; ++x)
      {
        const uint8_t* ta  = a + 4 * (o + (uint64_t)lw * y + x);
        const uint8_t* tb  = b + 4 * (o + (uint64_t)lw * y + x);
        int            any = 0;
        for(uint32_t c = 0; c < 4; ++c)
        {
          uint32_t d = ta[c] > tb[c] ? ta[c] - tb[c] : tb[c] - ta[c];
          any |= d != 0;
          if(d > r.worst_delta)
          {
            r.worst_delta = d;
            r.x = x, r.y = y, r.level = level, r.channel = c;
          }
        }
        r.mismatched_texels += any;
        r.compared_texels++;
      }
  }
  *out = r;
}

/* ------------------------------------------------------------------------ */
/* Premultiply-alpha pre-pass: include/scoped_image.hpp:233-255              */

void nvo_premultiply_srgba8(const uint8_t* in, uint8_t* out, uint64_t texels)
{
  for(uint64_t i = 0; i < texels; ++i)
  {
    const uint8_t* p     = in + 4 * i;
    const float    alpha = (float)p[3] * (1.f / 255.f);
    const float    red   = nvo_linear_from_srgb(p[0]) * alpha;
    const float    green = nvo_linear_from_srgb(p[1]) * alpha;
    const float    blue  = nvo_linear_from_srgb(p[2]) * alpha;
    uint8_t        a     = p[3];
    out[4 * i + 0]       = (uint8_t)nvo_srgb_from_linear(red);
    out[4 * i + 1]       = (uint8_t)nvo_srgb_from_linear(green);
    out[4 * i + 2]       = (uint8_t)nvo_srgb_from_linear(blue);
    out[4 * i + 3]       = a;
  }
}

/* ------------------------------------------------------------------------ */
/* Synthetic level-0 generators (bench / tests)                              */

/* Restatement of shaders/julia.comp:27-63 with the push constants of
 * demo_app/julia.cpp:65-81 (alphaNormalized as given, maxIterations 64). */
void nvo_julia_srgba8_rows(uint8_t* out, uint32_t w, uint32_t h, uint32_t y_begin, uint32_t y_end,
                           uint32_t alpha_normalized, int max_iterations);
void nvo_julia_srgba8(uint8_t* out, uint32_t w, uint32_t h, uint32_t alpha_normalized, int max_iterations)
{
  nvo_julia_srgba8_rows(out, w, h, 0, h, alpha_normalized, max_iterations);
}
/* Rows [y_begin, y_end) of the same image (out = base of the whole image): lets callers fill a large level 0 from
 * several host threads. */
void nvo_julia_srgba8_rows(uint8_t* out, uint32_t w, uint32_t h, uint32_t y_begin, uint32_t y_end,
                           uint32_t alpha_normalized, int max_iterations)
{
  const double alphaRadians = alpha_normalized * 1.4629180792671596e-09;
  const float  c_real       = (float)(0.7885 * sin(alphaRadians));
  const float  c_imag       = (float)(0.7885 * cos(alphaRadians));
  const float  offset_real  = -2.0f;
  const float  scale        = 4.0f / (float)w;
  const float  offset_imag  = 2.0f * (float)h / (float)w;
  for(uint32_t y = y_begin; y < y_end && y < h; ++y)
    for(uint32_t x = 0; x < w; ++x)
    {
      float zr = (float)x * scale + offset_real;
      float zi = (float)y * -scale + offset_imag;
      int   it = 0;
      while(it < max_iterations && zr * zr + zi * zi <= 4)
      {
        float tr = zr * zr - zi * zi;
        float ti = zr * zi + zr * zi;
        zr       = tr + c_real;
        zi       = ti + c_imag;
        ++it;
      }
      uint8_t* t = out + 4 * ((uint64_t)y * w + x);
      if(it < 16)
      {
        float s = (float)(4 + it) * (float)(1 / 20.);
        t[0]    = (uint8_t)(0 * s);
        t[1]    = (uint8_t)(128 * s);
        t[2]    = (uint8_t)(255 * s);
        t[3]    = (uint8_t)(255 * s);
      }
      else
      {
        /* it == max_iterations divides by zero in the shader (undefined
         * uint(inf)); pinned here to the saturated value. */
        uint32_t n = max_iterations == it ? 255u : (uint32_t)(127.0f * (float)(it - 16) / (float)(max_iterations - it));
        n          = n > 255u ? 255u : n;
        t[0]       = (uint8_t)n;
        t[1]       = (uint8_t)(128 + n / 4);
        t[2]       = (uint8_t)(255 - n);
        t[3]       = 255;
      }
    }
}
