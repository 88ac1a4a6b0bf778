// cudasph.hpp — sph::cuda_impl::Solver: the B200 backend behind the reference's solver interface.
//
// Include AFTER an `sph.hpp` (the reference's src/sph.hpp inside the reference tree, or include/pbf/sph.hpp here).
// Construction mirrors the other backends (omp_impl::Solver(h) benchmark.cpp:160-163, ocl_impl::Solver(h, ..., device)
// oclsph.hpp:145-148); advance() has sph::Solver::advance's contract (sph.hpp:119-125):
//   * xs is caller-owned, advanced in place, possibly resized by sources/drains, returned Z-sorted (ompsph.hpp:479-481);
//   * empty xs: prints "Particles depleted", sleeps 5 ms, returns an empty Result (ompsph.hpp:122-126);
//   * failures surface as std::runtime_error, which the drivers catch and rethrow (benchmark.cpp:34-37);
//   * may be called from any thread, serially (visualise.cpp:85-109): the C ABI sets the device on every entry.
// Sources and drains are the reference's host-side list edits (ompsph.hpp:91-120) and are reproduced here; wells and
// queries are not part of the accelerated path (both drivers pass an empty Scene) and are rejected.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "pbf_cuda.h"

namespace sph::cuda_impl {

template <typename T, typename N, template <size_t, typename> typename V> class Solver final : public sph::Solver<T, N, V> {
  static_assert(std::is_same_v<T, size_t> && std::is_same_v<N, float>,
                "the CUDA backend is fp32 only (like the OpenCL backend, benchmark.cpp:140-141)");
  using P = sph::Particle<T, N, V>;
  static_assert(sizeof(P) == sizeof(pbf_particle) && alignof(P) == alignof(pbf_particle), "Particle layout");
  static_assert(sizeof(V<3, N>) == 12 && sizeof(V<4, N>) == 16, "vector layout");

  pbf_ctx *ctx = nullptr;
  const N h;

  [[noreturn]] void raise(const char *where) const {
    throw std::runtime_error(std::string("sph::cuda_impl::Solver: ") + where + ": " + pbf_last_error(ctx));
  }

public:
  explicit Solver(N h, int device = 0) : h(h) {
    if (pbf_create(&ctx, h, device) != PBF_OK) raise("pbf_create");
  }
  ~Solver() override { pbf_destroy(ctx); }
  Solver(const Solver &) = delete;
  Solver &operator=(const Solver &) = delete;

  pbf_ctx *handle() const { return ctx; }  // for the resident path (pbf_upload / pbf_step / pbf_download)

  static pbf_params toParams(const sph::SphParams<T, N, V> &c) {
    pbf_params p{};
    p.dt = c.dt;
    p.scale = c.scale;
    p.iteration = c.iteration;
    p.constant_force[0] = c.constantForce.x; p.constant_force[1] = c.constantForce.y; p.constant_force[2] = c.constantForce.z;
    p.min_bound[0] = c.minBound.x; p.min_bound[1] = c.minBound.y; p.min_bound[2] = c.minBound.z;
    p.max_bound[0] = c.maxBound.x; p.max_bound[1] = c.maxBound.y; p.max_bound[2] = c.maxBound.z;
    p.wait = c.wait;
    p.surface_enabled = c.surface.has_value();
    if (c.surface) p.surface = {c.surface->resolution, c.surface->isolevel, c.surface->particleSize, c.surface->particleInfluence};
    return p;
  }

  sph::Result<T, N, V> advance(const sph::SphParams<T, N, V> &config, const sph::Scene<T, N, V> &scene,
                               std::vector<P> &xs) override {
    if (!scene.wells.empty() || !scene.queries.empty())
      throw std::runtime_error("sph::cuda_impl::Solver: wells and queries are not supported by the CUDA backend");
    // sources: a floor(sqrt(rate)) x ceil(sqrt(rate)) sheet of new particles, spacing h*scale/2, centred on the source
    const N spacing = h * config.scale / 2;
    for (const auto &src : scene.sources) {
      const N side = std::sqrt(static_cast<N>(src.rate));
      const size_t width = size_t(std::floor(side)), depth = size_t(std::ceil(side));
      for (size_t x = 0; x < width; ++x)
        for (size_t z = 0; z < depth; ++z)
          xs.emplace_back(src.tag, sph::Type::Fluid, N(1), src.colour,
                          V<3, N>(src.centre.x - N(width) * N(0.5) * spacing + N(x) * spacing, src.centre.y,
                                  src.centre.z - N(depth) * N(0.5) * spacing + N(z) * spacing),
                          src.velocity);
    }
    // drains: fluid particles within `width` of a drain centre disappear
    if (!scene.drains.empty())
      xs.erase(std::remove_if(xs.begin(), xs.end(),
                              [&](const P &p) {
                                if (p.type == sph::Type::Obstacle) return false;
                                for (const auto &d : scene.drains) {
                                  const N dx = d.centre.x - p.position.x, dy = d.centre.y - p.position.y, dz = d.centre.z - p.position.z;
                                  if (std::sqrt(dx * dx + dy * dy + dz * dz) < d.width) return true;
                                }
                                return false;
                              }),
               xs.end());
    if (xs.empty()) {
      std::cout << "Particles depleted" << std::endl;
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
      return {};
    }
    const pbf_params p = toParams(config);
    uint64_t nv = 0;
    if (pbf_advance_host(ctx, &p, reinterpret_cast<pbf_particle *>(xs.data()), xs.size(), &nv) != PBF_OK) raise("advance");
    sph::Result<T, N, V> r;
    if (nv) {
      r.mesh.vs.resize(nv); r.mesh.ns.resize(nv); r.mesh.cs.resize(nv);
      if (pbf_mesh_download(ctx, reinterpret_cast<float *>(r.mesh.vs.data()), reinterpret_cast<float *>(r.mesh.ns.data()),
                            reinterpret_cast<float *>(r.mesh.cs.data()), nv) != PBF_OK)
        raise("mesh download");
    }
    return r;
  }
};

}  // namespace sph::cuda_impl
