// vec.hpp — the few lines of vector type the host-side mirror needs.  The reference uses glm::vec (glm 0.9.9.8,
// fetched by CMake; not available offline); any type with public .x .y .z (.w) members, component constructors and
// the same memory layout works with the adaptor in cudasph.hpp, glm::vec included.
#pragma once

#include <cstddef>

namespace pbf {

template <std::size_t L, typename T> struct vec;

template <typename T> struct vec<3, T> {
  T x{}, y{}, z{};
  constexpr vec() = default;
  constexpr vec(T a, T b, T c) : x(a), y(b), z(c) {}
  template <typename A, typename B, typename C> constexpr vec(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {}
  constexpr vec &operator+=(const vec &o) { x += o.x; y += o.y; z += o.z; return *this; }
  constexpr bool operator==(const vec &o) const { return x == o.x && y == o.y && z == o.z; }
};

template <typename T> struct vec<4, T> {
  T x{}, y{}, z{}, w{};
  constexpr vec() = default;
  constexpr vec(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
  template <typename A, typename B, typename C, typename D>
  constexpr vec(A a, B b, C c, D d) : x(T(a)), y(T(b)), z(T(c)), w(T(d)) {}
  constexpr bool operator==(const vec &o) const { return x == o.x && y == o.y && z == o.z && w == o.w; }
};

template <typename T> constexpr vec<3, T> operator+(const vec<3, T> &a, const vec<3, T> &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> constexpr vec<3, T> operator*(const vec<3, T> &a, T s) { return {a.x * s, a.y * s, a.z * s}; }

}  // namespace pbf
