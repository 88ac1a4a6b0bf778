// ref_driver.cpp — C entry point around the UNMODIFIED reference OpenMP backend.  TEST INFRASTRUCTURE.
//
// Compiled by oracle/Makefile directly against the sources where they lie:
//   /root/reference/src/omp/ompsph.hpp (+ src/sph.hpp, curves.h, sph_constants.h, mc_constants.h, utils.hpp,
//   src/ocl/{cl_types.h,oclsph_type.h}, include/CL/*) with oracle/ref/glm_shim standing in for glm 0.9.9.8.
// Outputs go to oracle/_ref/ only (git-ignored, shipped to the GPU box by gpurun).  No reference source
// is copied into this repository.
//
// -DPBF_REF_STABLE_SORT swaps the reference's std::sort (ompsph.hpp:158, unstable on ties) for
// std::stable_sort by a token macro that is active only while ompsph.hpp is parsed; everything else
// in the reference is untouched.  That variant is the "reference with a stable sort" used to pin the
// oracle's tie order; the plain variant is the reference exactly as shipped.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <numeric>
#include <optional>
#include <sstream>
#include <thread>
#include <tuple>
#include <type_traits>
#include <vector>

#include <omp.h>

#include "glm/glm.hpp"
#include "glm/gtc/constants.hpp"
#include "glm/gtx/fast_exponential.hpp"
#include "glm/gtx/fast_square_root.hpp"
#include "glm/gtx/norm.hpp"
#include "glm/gtx/optimum_pow.hpp"

// reference headers that do not sort: parse them before the macro below can touch them
#include "curves.h"
#include "sph_constants.h"
#include "utils.hpp"
#include "mc_constants.h"
#include "oclsph_type.h"
#include "sph.hpp"

#ifdef PBF_REF_STABLE_SORT
#define sort stable_sort
#endif
#include "ompsph.hpp"
#ifdef PBF_REF_STABLE_SORT
#undef sort
#endif

#include "pbf_cuda.h"

namespace {
template <size_t L, typename T> using V = glm::vec<L, T>;
using Particle = sph::Particle<size_t, float, V>;
using Params = sph::SphParams<size_t, float, V>;

Params to_ref(const pbf_params &p) {
  Params c{};
  c.h = 0.f; // never read by any solver
  c.dt = p.dt;
  c.scale = p.scale;
  c.iteration = size_t(p.iteration);
  c.constantForce = V<3, float>(p.constant_force[0], p.constant_force[1], p.constant_force[2]);
  c.minBound = V<3, float>(p.min_bound[0], p.min_bound[1], p.min_bound[2]);
  c.maxBound = V<3, float>(p.max_bound[0], p.max_bound[1], p.max_bound[2]);
  c.wait = p.wait != 0;
  if (p.surface_enabled)
    c.surface = sph::McParams<float>{p.surface.resolution, p.surface.isolevel, p.surface.particle_size,
                                     p.surface.particle_influence};
  return c;
}
} // namespace

extern "C" {

const char *pbf_ref_variant() {
#ifdef PBF_REF_STABLE_SORT
  return "reference ompsph.hpp, std::sort->std::stable_sort, " PBF_REF_FLAGS;
#else
  return "reference ompsph.hpp unmodified, " PBF_REF_FLAGS;
#endif
}

int pbf_ref_max_threads() { return omp_get_max_threads(); }
void pbf_ref_set_threads(int n) { omp_set_num_threads(n); }

// One sph::Solver::advance on the reference OMP backend.  xs is rewritten in the reference's output order.
int pbf_ref_advance(float h, const pbf_params *p, pbf_particle *xs, uint64_t n, float *vs, float *ns, float *cs,
                    uint64_t cap_vertices, uint64_t *n_vertices) {
  std::vector<Particle> ps;
  ps.reserve(n);
  for (uint64_t i = 0; i < n; ++i) {
    const pbf_particle &q = xs[i];
    ps.emplace_back(size_t(q.id), static_cast<sph::Type>(q.type), q.mass,
                    V<4, float>(q.colour[0], q.colour[1], q.colour[2], q.colour[3]),
                    V<3, float>(q.position[0], q.position[1], q.position[2]),
                    V<3, float>(q.velocity[0], q.velocity[1], q.velocity[2]));
  }
  std::ostringstream sink; // the solver prints a stopwatch table every step (ompsph.hpp:394,482)
  std::streambuf *old = std::cout.rdbuf(sink.rdbuf());
  int rc = 0;
  try {
    sph::omp_impl::Solver<size_t, float> solver(h);
    auto result = solver.advance(to_ref(*p), sph::Scene<size_t, float, V>{}, ps);
    std::cout.rdbuf(old);
    if (ps.size() != n) return -4;
    for (uint64_t i = 0; i < n; ++i) {
      pbf_particle &q = xs[i];
      std::memset(&q, 0, sizeof(q));
      q.id = ps[i].id;
      q.type = uint8_t(ps[i].type);
      q.mass = ps[i].mass;
      q.position[0] = ps[i].position.x; q.position[1] = ps[i].position.y; q.position[2] = ps[i].position.z;
      q.velocity[0] = ps[i].velocity.x; q.velocity[1] = ps[i].velocity.y; q.velocity[2] = ps[i].velocity.z;
      q.colour[0] = ps[i].colour.x; q.colour[1] = ps[i].colour.y; q.colour[2] = ps[i].colour.z; q.colour[3] = ps[i].colour.w;
    }
    const uint64_t nv = result.mesh.vs.size();
    if (n_vertices) *n_vertices = nv;
    for (uint64_t i = 0; i < nv && i < cap_vertices; ++i) {
      if (vs) { vs[3 * i] = result.mesh.vs[i].x; vs[3 * i + 1] = result.mesh.vs[i].y; vs[3 * i + 2] = result.mesh.vs[i].z; }
      if (ns) { ns[3 * i] = result.mesh.ns[i].x; ns[3 * i + 1] = result.mesh.ns[i].y; ns[3 * i + 2] = result.mesh.ns[i].z; }
      if (cs) { cs[4 * i] = result.mesh.cs[i].x; cs[4 * i + 1] = result.mesh.cs[i].y; cs[4 * i + 2] = result.mesh.cs[i].z; cs[4 * i + 3] = result.mesh.cs[i].w; }
    }
  } catch (...) {
    std::cout.rdbuf(old);
    rc = -2;
  }
  return rc;
}

// The same with a Scene (sph.hpp:75-80): xs has room for `cap` particles, *n_out = particles after the call; the
// query results (ids) are written back to back into q_ids (cap q_ids_cap), q_counts[i] = ids of query i.
int pbf_ref_advance_scene(float h, const pbf_params *p, const pbf_scene *sc, pbf_particle *xs, uint64_t n, uint64_t cap,
                          uint64_t *n_out, uint64_t *q_ids, uint64_t q_ids_cap, uint64_t *q_counts) {
  std::vector<Particle> ps;
  ps.reserve(n);
  for (uint64_t i = 0; i < n; ++i) {
    const pbf_particle &q = xs[i];
    ps.emplace_back(size_t(q.id), static_cast<sph::Type>(q.type), q.mass,
                    V<4, float>(q.colour[0], q.colour[1], q.colour[2], q.colour[3]),
                    V<3, float>(q.position[0], q.position[1], q.position[2]),
                    V<3, float>(q.velocity[0], q.velocity[1], q.velocity[2]));
  }
  sph::Scene<size_t, float, V> scene{};
  if (sc) {
    for (uint32_t i = 0; i < sc->n_wells; ++i)
      scene.wells.push_back({size_t(sc->wells[i].tag),
                             V<3, float>(sc->wells[i].centre[0], sc->wells[i].centre[1], sc->wells[i].centre[2]),
                             sc->wells[i].force});
    for (uint32_t i = 0; i < sc->n_sources; ++i) {
      const pbf_source &s = sc->sources[i];
      scene.sources.push_back({size_t(s.tag), V<3, float>(s.centre[0], s.centre[1], s.centre[2]),
                               V<3, float>(s.velocity[0], s.velocity[1], s.velocity[2]),
                               V<4, float>(s.colour[0], s.colour[1], s.colour[2], s.colour[3]), s.rate});
    }
    for (uint32_t i = 0; i < sc->n_drains; ++i)
      scene.drains.push_back({size_t(sc->drains[i].tag),
                              V<3, float>(sc->drains[i].centre[0], sc->drains[i].centre[1], sc->drains[i].centre[2]),
                              sc->drains[i].width, sc->drains[i].depth});
    for (uint32_t i = 0; i < sc->n_queries; ++i)
      scene.queries.push_back({size_t(sc->queries[i].id),
                               V<3, float>(sc->queries[i].point[0], sc->queries[i].point[1], sc->queries[i].point[2])});
  }
  std::ostringstream sink;
  std::streambuf *old = std::cout.rdbuf(sink.rdbuf());
  int rc = 0;
  try {
    sph::omp_impl::Solver<size_t, float> solver(h);
    auto result = solver.advance(to_ref(*p), scene, ps);
    std::cout.rdbuf(old);
    if (n_out) *n_out = ps.size();
    if (ps.size() > cap) return -5;
    for (uint64_t i = 0; i < ps.size(); ++i) {
      pbf_particle &q = xs[i];
      std::memset(&q, 0, sizeof(q));
      q.id = ps[i].id;
      q.type = uint8_t(ps[i].type);
      q.mass = ps[i].mass;
      q.position[0] = ps[i].position.x; q.position[1] = ps[i].position.y; q.position[2] = ps[i].position.z;
      q.velocity[0] = ps[i].velocity.x; q.velocity[1] = ps[i].velocity.y; q.velocity[2] = ps[i].velocity.z;
      q.colour[0] = ps[i].colour.x; q.colour[1] = ps[i].colour.y; q.colour[2] = ps[i].colour.z; q.colour[3] = ps[i].colour.w;
    }
    uint64_t w = 0;
    for (size_t i = 0; i < result.queries.size(); ++i) {
      if (q_counts) q_counts[i] = result.queries[i].neighbours.size();
      for (size_t id : result.queries[i].neighbours) {
        if (q_ids && w < q_ids_cap) q_ids[w] = id;
        ++w;
      }
    }
  } catch (...) {
    std::cout.rdbuf(old);
    rc = -2;
  }
  return rc;
}

// The reference's scene factory (sph.hpp:160-186) + wall motion (sph.hpp:147-158), for fixture generation.
uint64_t pbf_ref_scene_2cubes(uint64_t count, uint64_t solver_iter, float scaling, pbf_params *out_p,
                              pbf_particle *xs, uint64_t cap) {
  auto [mc, cfg, prepared] = sph::simpleConfigWith2Cubes<size_t, float, V>(count, solver_iter, scaling);
  if (out_p) {
    std::memset(out_p, 0, sizeof(*out_p));
    out_p->dt = cfg.dt; out_p->scale = cfg.scale; out_p->iteration = cfg.iteration;
    out_p->constant_force[0] = cfg.constantForce.x; out_p->constant_force[1] = cfg.constantForce.y; out_p->constant_force[2] = cfg.constantForce.z;
    out_p->min_bound[0] = cfg.minBound.x; out_p->min_bound[1] = cfg.minBound.y; out_p->min_bound[2] = cfg.minBound.z;
    out_p->max_bound[0] = cfg.maxBound.x; out_p->max_bound[1] = cfg.maxBound.y; out_p->max_bound[2] = cfg.maxBound.z;
    out_p->wait = cfg.wait; out_p->surface_enabled = 0;
    out_p->surface.resolution = mc.resolution; out_p->surface.isolevel = mc.isolevel;
    out_p->surface.particle_size = mc.particleSize; out_p->surface.particle_influence = mc.particleInfluence;
  }
  for (uint64_t i = 0; i < prepared.size() && i < cap; ++i) {
    pbf_particle &q = xs[i];
    std::memset(&q, 0, sizeof(q));
    q.id = prepared[i].id; q.type = uint8_t(prepared[i].type); q.mass = prepared[i].mass;
    q.position[0] = prepared[i].position.x; q.position[1] = prepared[i].position.y; q.position[2] = prepared[i].position.z;
    q.velocity[0] = prepared[i].velocity.x; q.velocity[1] = prepared[i].velocity.y; q.velocity[2] = prepared[i].velocity.z;
    q.colour[0] = prepared[i].colour.x; q.colour[1] = prepared[i].colour.y; q.colour[2] = prepared[i].colour.z; q.colour[3] = prepared[i].colour.w;
  }
  return prepared.size();
}

void pbf_ref_apply_motion(const pbf_params *in, uint64_t frame, pbf_params *out) {
  Params c = sph::applyMotionSinXCosZ(to_ref(*in), size_t(frame));
  *out = *in;
  out->min_bound[0] = c.minBound.x; out->min_bound[1] = c.minBound.y; out->min_bound[2] = c.minBound.z;
  out->max_bound[0] = c.maxBound.x; out->max_bound[1] = c.maxBound.y; out->max_bound[2] = c.maxBound.z;
}

uint32_t pbf_ref_morton_encode(uint32_t x, uint32_t y, uint32_t z) { return uint32_t(zCurveGridIndexAtCoord(x, y, z)); }
void pbf_ref_morton_decode(uint32_t key, uint32_t xyz[3]) {
  xyz[0] = uint32_t(coordAtZCurveGridIndex0(key)); xyz[1] = uint32_t(coordAtZCurveGridIndex1(key)); xyz[2] = uint32_t(coordAtZCurveGridIndex2(key));
}
void pbf_ref_constants(float h, float out[3]) {
  out[0] = sph::poly6Factor(h);
  out[1] = sph::spikyKernelFactor(h);
  out[2] = sph::omp_impl::poly6Kernel(CorrDeltaQ * h, out[0], h);
}
} // extern "C"
