// glm shim (test infrastructure): gtc/constants — only pi<T>() is provided.
#pragma once
#include "../glm.hpp"
namespace glm {
template <typename T> constexpr T pi() { return T(3.14159265358979323846264338327950288); }
} // namespace glm
