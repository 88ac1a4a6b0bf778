// Minimal glm-compatible shim (TEST INFRASTRUCTURE, not product code).
//
// The reference fetches glm 0.9.9.8 through CMake FetchContent (CMakeLists.txt:26-30); glm is not
// vendored and there is no network here.  This header supplies just the subset of the glm surface
// that src/omp/ompsph.hpp, src/sph.hpp and src/utils.hpp touch, so that the UNMODIFIED reference
// OpenMP backend can be compiled where it lies under /root/reference (see oracle/Makefile).
//
// Arithmetic follows glm 0.9.9.8's published scalar definitions:
//   dot(vec3)   = (x*x' + y*y') + z*z'          length = sqrt(dot)       distance(a,b) = length(b-a)
//   mix(x,y,a)  = x*(1-a) + y*a                 clamp  = min(max(x,lo),hi)
//   min(a,b)    = (b<a) ? b : a                 max(a,b) = (a<b) ? b : a
//   vec / s     = component / s  (true division, no reciprocal)
//   fastSqrt(x) = 1 / inversesqrt(x), inversesqrt(x) = 1 / sqrt(x)  (highp path, no bit tricks)
#pragma once

#include <cassert>
#include <cmath>
#include <cstddef>
#include <functional>
#include <iomanip>
#include <string>
#include <type_traits>
#include <vector>

namespace glm {

template <size_t L, typename T> struct vec;

template <typename T> struct vec<3, T> {
  T x, y, z;
  constexpr vec() : x(0), y(0), z(0) {}
  template <typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>
  constexpr explicit vec(S s) : x(T(s)), y(T(s)), z(T(s)) {}
  template <typename A, typename B, typename C,
            typename = std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B> && std::is_arithmetic_v<C>>>
  constexpr vec(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {}
  template <typename U> constexpr vec(const vec<3, U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
  template <typename U> constexpr explicit vec(const vec<4, U> &o);
  vec &operator+=(const vec &o) { x += o.x; y += o.y; z += o.z; return *this; }
  vec &operator-=(const vec &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  bool operator==(const vec &o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const vec &o) const { return !(*this == o); }
};

template <typename T> struct vec<4, T> {
  T x, y, z, w;
  constexpr vec() : x(0), y(0), z(0), w(0) {}
  template <typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>
  constexpr explicit vec(S s) : x(T(s)), y(T(s)), z(T(s)), w(T(s)) {}
  template <typename A, typename B, typename C, typename D,
            typename = std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B> &&
                                        std::is_arithmetic_v<C> && std::is_arithmetic_v<D>>>
  constexpr vec(A a, B b, C c, D d) : x(T(a)), y(T(b)), z(T(c)), w(T(d)) {}
  template <typename U, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>
  constexpr vec(const vec<3, U> &o, S s) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(s)) {}
  template <typename U> constexpr vec(const vec<4, U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(o.w)) {}
  vec &operator+=(const vec &o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
  bool operator==(const vec &o) const { return x == o.x && y == o.y && z == o.z && w == o.w; }
  bool operator!=(const vec &o) const { return !(*this == o); }
};

template <typename T> template <typename U> constexpr vec<3, T>::vec(const vec<4, U> &o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}

template <typename T> using tvec3 = vec<3, T>;
template <typename T> using tvec4 = vec<4, T>;
using vec3 = vec<3, float>;
using vec4 = vec<4, float>;

#define GLM_SHIM_BINOP(OP)                                                                                             \
  template <typename T> constexpr vec<3, T> operator OP(const vec<3, T> &a, const vec<3, T> &b) {                      \
    return vec<3, T>(a.x OP b.x, a.y OP b.y, a.z OP b.z);                                                              \
  }                                                                                                                    \
  template <typename T> constexpr vec<3, T> operator OP(const vec<3, T> &a, T s) {                                     \
    return vec<3, T>(a.x OP s, a.y OP s, a.z OP s);                                                                    \
  }                                                                                                                    \
  template <typename T> constexpr vec<3, T> operator OP(T s, const vec<3, T> &a) {                                     \
    return vec<3, T>(s OP a.x, s OP a.y, s OP a.z);                                                                    \
  }                                                                                                                    \
  template <typename T> constexpr vec<4, T> operator OP(const vec<4, T> &a, const vec<4, T> &b) {                      \
    return vec<4, T>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w);                                                  \
  }                                                                                                                    \
  template <typename T> constexpr vec<4, T> operator OP(const vec<4, T> &a, T s) {                                     \
    return vec<4, T>(a.x OP s, a.y OP s, a.z OP s, a.w OP s);                                                          \
  }                                                                                                                    \
  template <typename T> constexpr vec<4, T> operator OP(T s, const vec<4, T> &a) {                                     \
    return vec<4, T>(s OP a.x, s OP a.y, s OP a.z, s OP a.w);                                                          \
  }
GLM_SHIM_BINOP(+)
GLM_SHIM_BINOP(-)
GLM_SHIM_BINOP(*)
GLM_SHIM_BINOP(/)
#undef GLM_SHIM_BINOP

template <typename T> constexpr vec<3, T> operator-(const vec<3, T> &a) { return vec<3, T>(-a.x, -a.y, -a.z); }
template <typename T> constexpr vec<4, T> operator-(const vec<4, T> &a) { return vec<4, T>(-a.x, -a.y, -a.z, -a.w); }

// scalar helpers
template <typename T> constexpr T min(T a, T b) { return (b < a) ? b : a; }
template <typename T> constexpr T max(T a, T b) { return (a < b) ? b : a; }
template <typename T> constexpr T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
template <typename T> inline T pow(T b, T e) { return std::pow(b, e); }
template <typename T> inline T sqrt(T x) { return std::sqrt(x); }
template <typename T> inline T inversesqrt(T x) { return T(1) / std::sqrt(x); }

template <typename T> constexpr vec<3, T> min(const vec<3, T> &a, const vec<3, T> &b) {
  return vec<3, T>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z));
}
template <typename T> constexpr vec<3, T> max(const vec<3, T> &a, const vec<3, T> &b) {
  return vec<3, T>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z));
}
template <typename T> constexpr vec<3, T> clamp(const vec<3, T> &v, T lo, T hi) {
  return vec<3, T>(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi));
}
template <typename T> constexpr vec<4, T> clamp(const vec<4, T> &v, T lo, T hi) {
  return vec<4, T>(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi), clamp(v.w, lo, hi));
}
template <typename T> constexpr vec<3, T> mix(const vec<3, T> &x, const vec<3, T> &y, T a) {
  return x * (T(1) - a) + y * a;
}
template <typename T> constexpr vec<4, T> mix(const vec<4, T> &x, const vec<4, T> &y, T a) {
  return x * (T(1) - a) + y * a;
}
template <typename T> inline vec<3, T> floor(const vec<3, T> &v) {
  return vec<3, T>(std::floor(v.x), std::floor(v.y), std::floor(v.z));
}

template <typename T> constexpr T dot(const vec<3, T> &a, const vec<3, T> &b) {
  return (a.x * b.x + a.y * b.y) + a.z * b.z;
}
template <typename T> inline T length(const vec<3, T> &v) { return std::sqrt(dot(v, v)); }
template <typename T> inline T distance(const vec<3, T> &a, const vec<3, T> &b) { return length(b - a); }
template <typename T> inline vec<3, T> normalize(const vec<3, T> &v) { return v * inversesqrt(dot(v, v)); }

} // namespace glm
