// Minimal stand-in for the parts of GLM that the reference's
// include/mipmap_storage.hpp and shaders/srgb.h use (glm is an un-vendored
// third-party dependency of the reference).  TEST INFRASTRUCTURE ONLY: lets
// oracle/Makefile compile those reference headers unmodified, in place.
#pragma once
#include <stdint.h>
namespace glm {
struct uvec2
{
  uint32_t x = 0, y = 0;
  uvec2() = default;
  uvec2(uint32_t x_, uint32_t y_) : x(x_), y(y_) {}
  bool operator==(const uvec2& o) const { return x == o.x && y == o.y; }
  bool operator!=(const uvec2& o) const { return !(*this == o); }
};
struct uvec3
{
  uint32_t x = 0, y = 0, z = 0;
  uvec3() = default;
  uvec3(uint32_t x_, uint32_t y_, uint32_t z_) : x(x_), y(y_), z(z_) {}
};
struct vec4f
{
  float x, y, z, w;
};
// glm::clamp(x, lo, hi) == min(max(x, lo), hi) with glm's comparison forms.
template <typename T>
inline T clamp(T x, T lo, T hi)
{
  T m = (x < lo) ? lo : x;
  return (hi < m) ? hi : m;
}
}  // namespace glm
