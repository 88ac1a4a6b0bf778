// cudasph.hpp — sph::cuda_impl::Solver: the B200 backend behind the reference's solver interface.
//
// Include AFTER an `sph.hpp` (the reference's src/sph.hpp inside the reference tree, or include/pbf/sph.hpp here).
// Construction mirrors the other backends (omp_impl::Solver(h) benchmark.cpp:160-163, ocl_impl::Solver(h, ..., device)
// oclsph.hpp:145-148); advance() has sph::Solver::advance's contract (sph.hpp:119-125):
//   * xs is caller-owned, advanced in place, possibly resized by sources/drains, returned Z-sorted (ompsph.hpp:479-481);
//   * empty xs: prints "Particles depleted", sleeps 5 ms, returns an empty Result (ompsph.hpp:122-126);
//   * failures surface as std::runtime_error, which the drivers catch and rethrow (benchmark.cpp:34-37);
//   * may be called from any thread, serially (visualise.cpp:85-109): the C ABI sets the device on every entry.
// The whole sph::Scene is honoured on the device (scene.cu): sources emit and drains remove on the resident arrays
// (ompsph.hpp:91-120), wells pull inside the prediction (:141-148), queries are answered from the step's cell table
// (:167-186).
//
// Several devices: Solver(h, {0, 1, 2, 3}) — the reference's `-d/--devices` is a list (args.cpp:20-23, args.hpp:46) —
// runs the step slab-decomposed along the Z-curve over those GPUs (csrc/dist.cu: one rank per device inside this
// process, ghost layers exchanged once per solver iteration) and returns the very particles, in the very order, one
// device returns.  Scene dynamics and the surface are single-device features for now: with several devices a
// non-empty Scene or config.surface raises std::runtime_error.
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

  pbf_ctx *ctx = nullptr;            // rank 0 (the only context on one device)
  std::vector<pbf_ctx *> ranks;      // every context, rank order
  const N h;

  [[noreturn]] void raise(const char *where) const {
    throw std::runtime_error(std::string("sph::cuda_impl::Solver: ") + where + ": " + pbf_last_error(ctx));
  }

public:
  // pinCallerMemory: page-lock the caller's std::vector storage while it keeps coming back (PBF_FLAG_PIN_HOST), so that
  // the copies of advance() run at the PCIe rate instead of through the runtime's pageable staging.  Safe whenever the
  // caller keeps its vector alive between calls, as both reference drivers do (benchmark.cpp:22-58, visualise.cpp:85-109);
  // this adaptor releases the lock itself before it grows the vector.  Call unpin() before freeing the vector early.
  explicit Solver(N h, int device = 0, bool pinCallerMemory = false) : Solver(h, std::vector<int>{device}, pinCallerMemory) {}
  // One slab rank per entry of `devices` (ordinals may repeat: several ranks then share a GPU).
  Solver(N h, const std::vector<int> &devices, bool pinCallerMemory = false) : h(h) {
    if (devices.empty()) throw std::runtime_error("sph::cuda_impl::Solver: empty device list");
    try {
      for (int dev : devices) {
        pbf_ctx *c = nullptr;
        if (pbf_create(&c, h, dev) != PBF_OK) {
          const std::string why = pbf_last_error(nullptr);
          throw std::runtime_error("sph::cuda_impl::Solver: pbf_create: " + why);
        }
        ranks.push_back(c);
      }
      ctx = ranks[0];
      if (pinCallerMemory && pbf_set_flags(ctx, PBF_FLAG_PIN_HOST) != PBF_OK) raise("pbf_set_flags");
      if (ranks.size() > 1 && pbf_dist_init_local(ranks.data(), int(ranks.size())) != PBF_OK) raise("pbf_dist_init_local");
    } catch (...) {
      for (pbf_ctx *c : ranks) pbf_destroy(c);
      throw;
    }
  }
  void unpin() { pbf_unpin_host(ctx); }
  size_t deviceCount() const { return ranks.size(); }
  ~Solver() override {
    for (pbf_ctx *c : ranks) pbf_destroy(c);
  }
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
    if (ranks.size() > 1) return advanceSlabs(config, scene, xs);
    // sph::Scene -> pbf_scene (plain arrays; the library copies them)
    std::vector<pbf_well> wells;
    std::vector<pbf_source> sources;
    std::vector<pbf_drain> drains;
    std::vector<pbf_query> queries;
    for (const auto &w : scene.wells) wells.push_back({uint64_t(w.tag), {w.centre.x, w.centre.y, w.centre.z}, w.force});
    for (const auto &s : scene.sources)
      sources.push_back({uint64_t(s.tag), {s.centre.x, s.centre.y, s.centre.z}, {s.velocity.x, s.velocity.y, s.velocity.z},
                         {s.colour.x, s.colour.y, s.colour.z, s.colour.w}, s.rate});
    for (const auto &d : scene.drains) drains.push_back({uint64_t(d.tag), {d.centre.x, d.centre.y, d.centre.z}, d.width, d.depth});
    for (const auto &q : scene.queries) queries.push_back({uint64_t(q.id), {q.point.x, q.point.y, q.point.z}});
    const pbf_scene sc{wells.data(),  uint32_t(wells.size()),  sources.data(), uint32_t(sources.size()),
                       drains.data(), uint32_t(drains.size()), queries.data(), uint32_t(queries.size())};
    // sources append a floor(sqrt(rate)) x ceil(sqrt(rate)) sheet each (ompsph.hpp:93-104): make room for them
    size_t emitted = 0;
    for (const auto &s : scene.sources) {
      const N side = std::sqrt(static_cast<N>(s.rate));
      emitted += size_t(std::floor(side)) * size_t(std::ceil(side));
    }
    const size_t n = xs.size();
    if (n + emitted > xs.capacity()) pbf_unpin_host(ctx);  // the resize below moves the storage
    xs.resize(n + emitted);
    const pbf_params p = toParams(config);
    uint64_t nv = 0, n_out = 0;
    const int rc = pbf_advance_scene_host(ctx, &p, &sc, reinterpret_cast<pbf_particle *>(xs.data()), n, xs.size(), &n_out, &nv);
    if (rc != PBF_OK) {
      xs.resize(n);  // the particles themselves are untouched on failure
      raise("advance");
    }
    xs.resize(n_out);
    if (xs.empty()) {
      std::cout << "Particles depleted" << std::endl;  // ompsph.hpp:122-126
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
      return {};
    }
    sph::Result<T, N, V> r;
    if (nv) {
      r.mesh.vs.resize(nv); r.mesh.ns.resize(nv); r.mesh.cs.resize(nv);
      if (pbf_mesh_download(ctx, reinterpret_cast<float *>(r.mesh.vs.data()), reinterpret_cast<float *>(r.mesh.ns.data()),
                            reinterpret_cast<float *>(r.mesh.cs.data()), nv) != PBF_OK)
        raise("mesh download");
    }
    for (size_t i = 0; i < scene.queries.size(); ++i) {  // sph::QueryResult, sph.hpp:27-31
      uint64_t count = 0;
      pbf_query_result(ctx, uint32_t(i), nullptr, 0, &count);
      std::vector<uint64_t> ids(count);
      if (count && pbf_query_result(ctx, uint32_t(i), ids.data(), count, &count) != PBF_OK) raise("query");
      r.queries.push_back({scene.queries[i].id, scene.queries[i].point, std::vector<T>(ids.begin(), ids.end())});
    }
    return r;
  }

private:
  sph::Result<T, N, V> advanceSlabs(const sph::SphParams<T, N, V> &config, const sph::Scene<T, N, V> &scene, std::vector<P> &xs) {
    if (!scene.wells.empty() || !scene.sources.empty() || !scene.drains.empty() || !scene.queries.empty())
      throw std::runtime_error("sph::cuda_impl::Solver: a non-empty Scene needs a single device");
    if (xs.empty()) {
      std::cout << "Particles depleted" << std::endl;  // ompsph.hpp:122-126
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
      return {};
    }
    const pbf_params p = toParams(config);
    uint64_t nv = 0;
    if (pbf_dist_advance_host(ctx, &p, reinterpret_cast<pbf_particle *>(xs.data()), xs.size(), &nv) != PBF_OK) raise("advance");
    sph::Result<T, N, V> r;
    if (nv) {  // the surface of the whole fluid, extracted on rank 0 from the lattice all ranks filled (same mesh as one device)
      r.mesh.vs.resize(nv); r.mesh.ns.resize(nv); r.mesh.cs.resize(nv);
      if (pbf_mesh_download(ctx, reinterpret_cast<float *>(r.mesh.vs.data()), reinterpret_cast<float *>(r.mesh.ns.data()),
                            reinterpret_cast<float *>(r.mesh.cs.data()), nv) != PBF_OK)
        raise("mesh download");
    }
    return r;
  }
};

}  // namespace sph::cuda_impl
