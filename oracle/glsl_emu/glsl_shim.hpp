// glsl_shim.hpp -- just enough of GLSL 4.60, as C++, to compile the REFERENCE's shader sources
// (nvpro_pyramid/srgba8_mipmap_preamble.glsl + nvpro_pyramid/nvpro_pyramid.glsl) unmodified and in
// place, and execute them on the CPU: every invocation of a work group is a fiber, subgroup
// shuffles and barrier() are scheduling points.  TEST INFRASTRUCTURE ONLY: this is how the
// shader-order oracle (oracle/nvpyr_oracle.c, "Oracle A") is pinned against the reference's own
// schedule, thread<->texel mapping, pairing order and carry groups.
//
// Implementation-defined float behaviour is pinned here exactly as in DESIGN.md:
//   * `float` is IEEE binary32; literals/ints are narrowed to binary32 BEFORE they take part in an
//     operation (GLSL has no double arithmetic in these shaders) -- class Float;
//   * a0*v0 + a1*v1 + a2*v2 on vectors is contracted like a GPU compiler does: mul, fma, fma
//     (lazy product type Prod); nothing else is contracted;
//   * texelFetch on the sRGB view = the reference's linearFromSrgb (shaders/srgb.h), alpha a*(1/255);
//   * subgroups are 32 wide and laid out linearly (gl_SubgroupInvocationID = local index & 31).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef uint32_t uint;

// ------------------------------------------------------------------ scalars
struct Float
{
  float v;
  Float() = default;
  Float(float f) : v(f) {}
  Float(double d) : v(static_cast<float>(d)) {}
  Float(int i) : v(static_cast<float>(i)) {}
  Float(uint u) : v(static_cast<float>(u)) {}
  explicit operator uint() const { return v > 0.f ? (v >= 4294967296.f ? 0xFFFFFFFFu : static_cast<uint>(v)) : 0u; }
  explicit operator int() const { return static_cast<int>(v); }
};
inline Float operator+(Float a, Float b) { return Float(a.v + b.v); }
inline Float operator-(Float a, Float b) { return Float(a.v - b.v); }
inline Float operator*(Float a, Float b) { return Float(a.v * b.v); }
inline Float operator/(Float a, Float b) { return Float(a.v / b.v); }
inline bool  operator<=(Float a, Float b) { return a.v <= b.v; }
inline bool  operator<(Float a, Float b) { return a.v < b.v; }
inline Float pow(Float a, Float b) { return Float(powf(a.v, b.v)); }
inline Float clamp(Float x, Float lo, Float hi) { return x.v < lo.v ? lo : (x.v > hi.v ? hi : x); }
template <class T>
inline T clamp(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
template <class T>
inline T min(T a, T b) { return b < a ? b : a; }

// ------------------------------------------------------------------ vectors
struct ivec2
{
  int x, y;
  ivec2() = default;
  template <class A, class B>
  ivec2(A a, B b) : x(static_cast<int>(a)), y(static_cast<int>(b)) {}
};
inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator*(ivec2 a, ivec2 b) { return ivec2(a.x * b.x, a.y * b.y); }
inline ivec2 operator*(ivec2 a, int s) { return ivec2(a.x * s, a.y * s); }
inline ivec2 operator>>(ivec2 a, int s) { return ivec2(a.x >> s, a.y >> s); }
inline ivec2 operator<<(ivec2 a, int s) { return ivec2(a.x << s, a.y << s); }
inline ivec2& operator>>=(ivec2& a, int s) { a = a >> s; return a; }
inline ivec2& operator<<=(ivec2& a, int s) { a = a << s; return a; }
inline ivec2& operator+=(ivec2& a, ivec2 b) { a = a + b; return a; }

struct uvec4
{
  uint x, y, z, w;
  uvec4() = default;
  uvec4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
};

struct vec4;
struct f16vec4;  // GL_EXT_shader_explicit_arithmetic_types, used by the F16_SHARED build of the preamble
struct Prod  // a * v, not yet rounded: lets "p + q" and "v + p" contract into an fma like a GPU compiler
{
  float a;
  float v[4];
};
struct vec4
{
  Float x, y, z, w;
  Float &r = x, &g = y, &b = z, &a = w;  // colour aliases (srgbFromLinearVec uses .r .g .b .a)
  vec4() {}
  vec4(Float x_, Float y_, Float z_, Float w_) : x(x_), y(y_), z(z_), w(w_) {}
  vec4(const Prod& p) : x(p.a * p.v[0]), y(p.a * p.v[1]), z(p.a * p.v[2]), w(p.a * p.v[3]) {}
  vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
  explicit vec4(const f16vec4& h);
  vec4& operator=(const vec4& o)
  {
    x = o.x, y = o.y, z = o.z, w = o.w;
    return *this;
  }
};
// f16vec4(vec4) rounds every component to IEEE binary16, round to nearest even (the conversion mode GPUs
// use for OpFConvert unless decorated otherwise); vec4(f16vec4) widens exactly.
struct f16vec4
{
  _Float16 x, y, z, w;
  f16vec4() = default;
  explicit f16vec4(const vec4& v) : x(_Float16(v.x.v)), y(_Float16(v.y.v)), z(_Float16(v.z.v)), w(_Float16(v.w.v)) {}
};
inline vec4::vec4(const f16vec4& h) : x(float(h.x)), y(float(h.y)), z(float(h.z)), w(float(h.w)) {}
inline vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline Prod operator*(Float s, const vec4& v) { return Prod{s.v, {v.x.v, v.y.v, v.z.v, v.w.v}}; }
inline vec4 operator+(const vec4& c, const Prod& p)
{
  return vec4(fmaf(p.a, p.v[0], c.x.v), fmaf(p.a, p.v[1], c.y.v), fmaf(p.a, p.v[2], c.z.v), fmaf(p.a, p.v[3], c.w.v));
}
inline vec4 operator+(const Prod& p, const Prod& q) { return vec4(p) + q; }

// packUnorm4x8 / unpackUnorm4x8 (used by the SRGB_SHARED build of the preamble): GLSL 4.60 section 8.4 --
// pack: round(clamp(c, 0, +1) * 255.0), component x in the least significant byte; unpack: byte / 255.0.
// The direction of round()'s ties is implementation-defined in GLSL; pinned here as round-half-even.
inline uint packUnorm4x8(const vec4& v)
{
  auto q = [](Float c) {
    float s = c.v < 0.f ? 0.f : (c.v > 1.f ? 1.f : c.v);
    if(!(c.v == c.v))
      s = 0.f;
    return uint(rintf(s * 255.0f));
  };
  return q(v.x) | (q(v.y) << 8) | (q(v.z) << 16) | (q(v.w) << 24);
}
inline vec4 unpackUnorm4x8(uint p)
{
  return vec4(Float(p & 255u) / Float(255.0), Float((p >> 8) & 255u) / Float(255.0), Float((p >> 16) & 255u) / Float(255.0),
              Float(p >> 24) / Float(255.0));
}

// ------------------------------------------------------------------ resources
struct uimage2D
{
  uint8_t* data;  // RGBA8, tight rows
  int      w, h;
};
struct sampler2D
{
  const uimage2D* levels;  // the same image's levels, read through the sRGB view
};
float emuLinearFromSrgb(uint c);  // the reference's shaders/srgb.h (glsl_emu.cpp)
inline vec4 texelFetch(const sampler2D& s, ivec2 p, int level)
{
  const uimage2D& im = s.levels[level];
  const uint8_t*  t  = im.data + 4 * (size_t(p.y) * im.w + p.x);
  return vec4(emuLinearFromSrgb(t[0]), emuLinearFromSrgb(t[1]), emuLinearFromSrgb(t[2]), float(t[3]) * (1.0f / 255.0f));
}
extern uint64_t g_emuStores;
inline void imageStore(const uimage2D& im, ivec2 p, uvec4 c)
{
  uint8_t* t = im.data + 4 * (size_t(p.y) * im.w + p.x);
  t[0] = uint8_t(c.x), t[1] = uint8_t(c.y), t[2] = uint8_t(c.z), t[3] = uint8_t(c.w);
  ++g_emuStores;
}
inline ivec2 imageSize(const uimage2D& im) { return ivec2(im.w, im.h); }

// ------------------------------------------------------------------ invocation state + scheduling points
struct EmuInvocation
{
  uint local, global, workgroup;
};
extern EmuInvocation* g_emuCur;
extern uint           g_emuPushConstant;
struct uvec3x { uint x; };
#define gl_LocalInvocationIndex (g_emuCur->local)
#define gl_SubgroupInvocationID (g_emuCur->local & 31u)
#define gl_GlobalInvocationID (uvec3x{g_emuCur->global})
#define gl_WorkGroupID (uvec3x{g_emuCur->workgroup})
void barrier();
vec4 subgroupShuffleXor(const vec4& v, uint mask);

// GLSL declaration syntax that has no meaning here
#define layout(...)
#define uniform
#define writeonly
#define in int emuLocalSizeDecl_  // "layout(local_size_x = N) in;" becomes a harmless extern declaration
#define shared static
#define NVPRO_PYRAMID_PUSH_CONSTANT g_emuPushConstant  // documented hook, nvpro_pyramid.glsl:65-69
