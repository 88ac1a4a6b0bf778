// glm shim (test infrastructure): gtx/fast_square_root, highp definitions of glm 0.9.9.8:
//   fastInverseSqrt(x) = 1/sqrt(x); fastSqrt(x) = 1/fastInverseSqrt(x);
//   fastLength(v) = fastSqrt(dot(v,v)); fastDistance(a,b) = fastLength(b-a);
//   fastNormalize(v) = v * fastInverseSqrt(dot(v,v)).
#pragma once
#include "../glm.hpp"
namespace glm {
template <typename T> inline T fastInverseSqrt(T x) { return T(1) / std::sqrt(x); }
template <typename T> inline T fastSqrt(T x) { return T(1) / fastInverseSqrt(x); }
template <typename T> inline T fastLength(const vec<3, T> &v) { return fastSqrt(dot(v, v)); }
template <typename T> inline T fastDistance(const vec<3, T> &a, const vec<3, T> &b) { return fastLength(b - a); }
template <typename T> inline vec<3, T> fastNormalize(const vec<3, T> &v) { return v * fastInverseSqrt(dot(v, v)); }
} // namespace glm
