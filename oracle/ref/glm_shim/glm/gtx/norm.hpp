// glm shim (test infrastructure): gtx/norm — length2(v) = dot(v,v).
#pragma once
#include "../glm.hpp"
namespace glm {
template <typename T> constexpr T length2(const vec<3, T> &v) { return dot(v, v); }
} // namespace glm
