// glm shim (test infrastructure): gtx/optimum_pow — pow2(x) = x*x, pow3(x) = x*x*x.
#pragma once
#include "../glm.hpp"
namespace glm {
template <typename T> constexpr T pow2(T x) { return x * x; }
template <typename T> constexpr T pow3(T x) { return x * x * x; }
} // namespace glm
