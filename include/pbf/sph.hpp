// sph.hpp — host-side mirror of the reference's solver interface and data model (reference src/sph.hpp), so that
// code written against `sph::Solver<T,N,V>::advance` compiles unchanged against this repository.  Same namespace,
// type names, field names and field order as the reference (that is the interface); the vector type is a template
// template parameter exactly as there, so the reference's glm::vec and pbf::vec both fit.
//
//   data model        reference src/sph.hpp:15-125
//   scene factory     reference src/sph.hpp:127-186  (makeCube, applyMotionSinXCosZ, simpleConfigWith2Cubes)
//   + damBreak        the scaled-up dam-break family used by BASELINE.json's configs
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <optional>
#include <tuple>
#include <vector>

namespace sph {

enum class Type : uint8_t { Fluid = 0, Obstacle = 1 };

template <typename T, typename N, template <size_t, typename C = N> typename V> struct Query {
  T id;
  V<3> point;
};

template <typename T, typename N, template <size_t, typename C = N> typename V> struct QueryResult {
  T id{};
  V<3> point{};
  std::vector<T> neighbours{};
};

// 56 bytes for <size_t, float>: id@0 type@8 mass@12 position@16 velocity@28 colour@40 — this is pbf_particle.
template <typename T, typename N, template <size_t, typename C = N> typename V> struct Particle {
  T id{};
  Type type{Type::Fluid};
  N mass{};
  V<3> position{}, velocity{};
  V<4> colour{};

  constexpr Particle() = default;
  constexpr explicit Particle(T id, Type type, N mass, const V<4> &colour, const V<3> &position, const V<3> &velocity)
      : id(id), type(type), mass(mass), position(position), velocity(velocity), colour(colour) {}
};

template <typename T, typename N, template <size_t, typename C = N> typename V> struct Well {
  T tag;
  V<3> centre;
  N force;
};

template <typename T, typename N, template <size_t, typename C = N> typename V> struct Source {
  T tag;
  V<3> centre, velocity;
  V<4> colour;
  N rate;
};

template <typename T, typename N, template <size_t, typename C = N> typename V> struct Drain {
  T tag;
  V<3> centre;
  N width, depth;
};

template <typename T, typename N, template <size_t S, typename C = N> typename V> struct Scene {
  std::vector<Well<T, N, V>> wells;
  std::vector<Source<T, N, V>> sources;
  std::vector<Drain<T, N, V>> drains;
  std::vector<Query<T, N, V>> queries;
};

template <typename N> struct McParams {
  N resolution, isolevel, particleSize, particleInfluence;
};

template <typename T, typename N, template <size_t S, typename C = N> typename V> struct SphParams {
  N h, dt, scale;  // h is carried for interface parity only: every solver takes h as a constructor argument
  size_t iteration;
  V<3> constantForce, minBound, maxBound;
  bool wait;
  std::optional<McParams<N>> surface;
};

template <typename N, template <size_t S, typename C = N> typename V> struct ColouredMesh {
  std::vector<V<3>> vs{}, ns{};
  std::vector<V<4>> cs{};
};

template <typename T, typename N, template <size_t, typename C = N> typename V> struct Result {
  ColouredMesh<N, V> mesh{};
  std::vector<QueryResult<T, N, V>> queries{};
};

template <typename T, typename N, template <size_t, typename _ = N> typename V> class Solver {
public:
  virtual ~Solver() = default;
  // One PBF step.  xs is caller-owned, advanced in place and returned in Z-sorted order with ids carried.
  virtual Result<T, N, V> advance(const SphParams<T, N, V> &config, const Scene<T, N, V> &scene,
                                  std::vector<Particle<T, N, V>> &xs) = 0;
};

// ---- scene factory ----------------------------------------------------------------------------------------------------
// side^3 lattice, x slowest / z fastest, ids consecutive from `offset`; returns the next free id.
template <typename T, typename N, template <size_t, typename C = N> typename V>
T makeCube(T offset, N spacing, size_t count, V<3> origin, V<4> colour, std::vector<Particle<T, N, V>> &xs) {
  const auto side = static_cast<size_t>(std::cbrt(count));
  xs.reserve(xs.size() + side * side * side);
  for (size_t x = 0; x < side; ++x)
    for (size_t y = 0; y < side; ++y)
      for (size_t z = 0; z < side; ++z)
        xs.emplace_back(offset++, Type::Fluid, N(1), colour,
                        V<3>(N(x) * spacing + origin.x, N(y) * spacing + origin.y, N(z) * spacing + origin.z), V<3>(0, 0, 0));
  return offset;
}

// The benchmark's moving wall: the box slides by (300 sin(f/20), 0, 90 cos(f/20)).
template <typename T, typename N, template <size_t, typename C = N> typename V>
SphParams<T, N, V> applyMotionSinXCosZ(const SphParams<T, N, V> &config, size_t frame) {
  const float rate = 20.f, amplitude = 300.f;
  const N dx = N(std::sin(float(frame) / rate) * amplitude);
  const N dz = N(std::cos(float(frame) / rate) * amplitude * 0.3);
  auto moved = config;
  moved.minBound.x += dx; moved.maxBound.x += dx;
  moved.minBound.z += dz; moved.maxBound.z += dz;
  return moved;
}

template <typename T, typename N, template <size_t, typename C = N> typename V>
SphParams<T, N, V> defaultParams(size_t solverIter, N scaling) {
  SphParams<T, N, V> c{};
  c.h = N(0.1);
  c.dt = N(0.0083 * 1.5f);
  c.scale = scaling;
  c.iteration = solverIter;
  c.constantForce = V<3>(0, 9.8, 0);
  c.minBound = V<3>(0, 0, 0);
  c.maxBound = V<3>(1000, 1000, 1000);
  c.wait = true;
  return c;
}

// Two cubes of (count/2) particles each in a 1000^3 box; returns {McParams, SphParams, particles} like the reference.
template <typename T, typename N, template <size_t, typename C = N> typename V>
std::tuple<McParams<N>, SphParams<T, N, V>, std::vector<Particle<T, N, V>>> simpleConfigWith2Cubes(size_t count,
                                                                                                   size_t solverIter,
                                                                                                   N scaling) {
  std::vector<Particle<T, N, V>> xs;
  T id{};
  id = makeCube<T, N, V>(id, N(22), count / 2, V<3>(100, 0, 100), V<4>(0, 0.1, 0.8, 1), xs);
  id = makeCube<T, N, V>(id, N(22), count / 2, V<3>(600, 0, 600), V<4>(0.1, 0.8, 0.1, 1), xs);
  return {McParams<N>{N(2), N(100), N(25), N(0.5)}, defaultParams<T, N, V>(solverIter, scaling), xs};
}

// dam(side): one side^3 block, spacing 22, against the +y (gravity) wall of a (44 side + 200) x (22 side + 200)^2 box.
template <typename T, typename N, template <size_t, typename C = N> typename V>
std::tuple<McParams<N>, SphParams<T, N, V>, std::vector<Particle<T, N, V>>> damBreak(size_t side, size_t solverIter,
                                                                                     N scaling) {
  const N lx = N(44) * N(side) + N(200), ly = N(22) * N(side) + N(200);
  auto c = defaultParams<T, N, V>(solverIter, scaling);
  c.maxBound = V<3>(lx, ly, ly);
  // (not makeCube: its side = (size_t)cbrt(count) rounds 30^3 down to 29, like the reference's would)
  std::vector<Particle<T, N, V>> xs;
  xs.reserve(side * side * side);
  const V<3> origin(100, ly - N(22) * N(side) - N(50), 100);
  T id{};
  for (size_t x = 0; x < side; ++x)
    for (size_t y = 0; y < side; ++y)
      for (size_t z = 0; z < side; ++z)
        xs.emplace_back(id++, Type::Fluid, N(1), V<4>(0, 0.1, 0.8, 1),
                        V<3>(N(x) * N(22) + origin.x, N(y) * N(22) + origin.y, N(z) * N(22) + origin.z), V<3>(0, 0, 0));
  return {McParams<N>{N(2), N(100), N(25), N(0.5)}, c, xs};
}

}  // namespace sph
