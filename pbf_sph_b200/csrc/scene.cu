// scene.cu — the scene dynamics of advance(): sources, drains and queries on the resident state.
//
// The reference edits the caller's particle vector on the host before the step (ompsph.hpp:91-120: sources append a
// sheet of particles, drains erase fluid near their centre) and answers queries from the step's cell table
// (ompsph.hpp:167-186).  Here the state is resident, so
//   sources  the emitted records (a few hundred at most) are formed on the host with the reference's arithmetic and
//            appended to the SoA arrays;
//   drains   keep-flags -> exclusive scan -> stable compaction into the other buffer set (the order of the
//            survivors is the reference's std::remove_if order); the new count is the one host read-back;
//   queries  each query resolves to ONE cell, i.e. one contiguous range of the Z-sorted id array: a tiny kernel turns
//            the query points into (first, count) ranges, pbf_query_result copies the ids out of the sorted array.
// Wells act inside the prediction (common.cuh predict(), ompsph.hpp:141-148).
#include <cmath>
#include <vector>

#include "common.cuh"

namespace pbf {

namespace {

constexpr int kBlock = 256;

// keep[i] = 0 when particle i is within `width` of a drain centre: glm::distance(drain.centre, x.position) < drain.width
// (ompsph.hpp:111), distance = sqrt((dx*dx + dy*dy) + dz*dz) of x.position - drain.centre, strict arithmetic.
__global__ void __launch_bounds__(kBlock) drain_flags_kernel(uint32_t n, const float4 *__restrict__ pos,
                                                             const float4 *__restrict__ drains, uint32_t n_drains,
                                                             uint32_t *__restrict__ keep) {
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const float4 p = ldg4(pos + i);
  uint32_t k = 1;
  for (uint32_t d = 0; d < n_drains; ++d) {
    const float4 dr = __ldg(drains + d);  // centre.xyz, width
    const float dx = fsub(p.x, dr.x), dy = fsub(p.y, dr.y), dz = fsub(p.z, dr.z);
    if (fsqrt(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz))) < dr.w) { k = 0; break; }
  }
  keep[i] = k;
}

__global__ void __launch_bounds__(kBlock) compact_kernel(uint32_t n, const uint32_t *__restrict__ keep,
                                                         const uint32_t *__restrict__ dst, const float4 *__restrict__ pos,
                                                         const float4 *__restrict__ vel, const float4 *__restrict__ col,
                                                         const unsigned long long *__restrict__ ids,
                                                         float4 *__restrict__ pos_out, float4 *__restrict__ vel_out,
                                                         float4 *__restrict__ col_out,
                                                         unsigned long long *__restrict__ ids_out) {
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const uint32_t o = dst[i];
  pos_out[o] = pos[i];
  vel_out[o] = vel[i];
  col_out[o] = col[i];
  ids_out[o] = ids[i];
}

// ompsph.hpp:172-182: scaled = point / scale - minExtent; zIdx = morton(size_t(scaled / h)); the result is the
// particle range of cell zIdx when zIdx < G and zIdx + 1 < G, else empty.
__global__ void query_ranges_kernel(StepConst c, uint32_t n_queries, const float4 *__restrict__ points,
                                    const uint32_t *__restrict__ table, uint2 *__restrict__ ranges) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_queries) return;
  const float4 pt = points[q];
  const float v[3] = {pt.x, pt.y, pt.z};
  uint32_t cc[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) cc[a] = cell_coord(fdiv(fsub(fdiv(v[a], c.scale), c.min_extent[a]), c.h));
  const uint32_t z = morton3(cc[0], cc[1], cc[2]);
  uint2 r = make_uint2(0u, 0u);
  if (z < c.G && z + 1u < c.G) {
    r.x = table[z];
    r.y = table[z + 1u] - r.x;
  }
  ranges[q] = r;
}

}  // namespace

// The particles the scene's sources emit in one call, in the reference's order (ompsph.hpp:92-104).
void scene_emit(float h, float scale, const std::vector<pbf_source> &sources, std::vector<pbf_particle> &out) {
  const float spacing = h * scale / 2;
  for (const pbf_source &s : sources) {
    const float size = std::sqrt(s.rate);
    const size_t width = (size_t)std::floor(size), depth = (size_t)std::ceil(size);
    // offset = centre - (V3(width, 0, depth) * 0.5 * spacing); pos = offset + V3(x, 0, z) * spacing
    const float ox = s.centre[0] - ((float)width * 0.5f) * spacing;
    const float oy = s.centre[1] - (0.0f * 0.5f) * spacing;
    const float oz = s.centre[2] - ((float)depth * 0.5f) * spacing;
    for (size_t x = 0; x < width; ++x)
      for (size_t z = 0; z < depth; ++z) {
        pbf_particle p{};
        p.id = s.tag;
        p.type = PBF_TYPE_FLUID;
        p.mass = 1.0f;
        p.position[0] = ox + (float)x * spacing;
        p.position[1] = oy + 0.0f * spacing;
        p.position[2] = oz + (float)z * spacing;
        for (int a = 0; a < 3; ++a) p.velocity[a] = s.velocity[a];
        for (int a = 0; a < 4; ++a) p.colour[a] = s.colour[a];
        out.push_back(p);
      }
  }
}

// Sources then drains on the resident state (the reference's order).  Synchronises the stream when there are drains.
int scene_edit_particles(pbf_ctx *ctx, const pbf_params &p) {
  pbf_scene_state &sc = ctx->scene;
  if (!sc.sources.empty()) {
    std::vector<pbf_particle> fresh;
    scene_emit(ctx->h, p.scale, sc.sources, fresh);
    const uint64_t k = fresh.size(), n = ctx->n;
    if (k) {
      if (n + k >= 0xFFFFFFF0ull) return fail(ctx, PBF_ERR_INVALID, "n", "more than 2^32 particles on one device");
      PBF_CUDA(ctx, ctx->pos[ctx->cur].reserve(n + k, true, ctx->stream));
      PBF_CUDA(ctx, ctx->vel[ctx->cur].reserve(n + k, true, ctx->stream));
      PBF_CUDA(ctx, ctx->col[ctx->cur_col].reserve(n + k, true, ctx->stream));
      PBF_CUDA(ctx, ctx->ids[ctx->cur].reserve(n + k, true, ctx->stream));
      PBF_CUDA(ctx, ctx->aos.reserve(k + 1));
      // pageable source: the copy is staged by the runtime before the call returns
      PBF_CUDA(ctx, cudaMemcpyAsync(ctx->aos.p, fresh.data(), k * sizeof(pbf_particle), cudaMemcpyHostToDevice, ctx->stream));
      PBF_TRY(launch_unpack_aos(ctx, ctx->aos.p, k, ctx->pos[ctx->cur].p + n, ctx->vel[ctx->cur].p + n,
                                ctx->col[ctx->cur_col].p + n, ctx->ids[ctx->cur].p + n, ctx->flag_dev + 1));
      PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `fresh` goes out of scope
      ctx->n = n + k;
    }
  }
  if (!sc.drains.empty() && ctx->n) {
    const uint32_t n = (uint32_t)ctx->n;
    const int o = ctx->cur ^ 1, oc = ctx->cur_col ^ 1;
    PBF_CUDA(ctx, ctx->pos[o].reserve(n));
    PBF_CUDA(ctx, ctx->vel[o].reserve(n));
    PBF_CUDA(ctx, ctx->ids[o].reserve(n));
    PBF_CUDA(ctx, ctx->col[oc].reserve(n));
    PBF_CUDA(ctx, ctx->key_in.reserve(n));  // keep flags
    PBF_CUDA(ctx, ctx->key_a.reserve(n));   // destinations
    drain_flags_kernel<<<div_up(n, kBlock), kBlock, 0, ctx->stream>>>(n, ctx->pos[ctx->cur].p, sc.d_drains.p,
                                                                       (uint32_t)sc.drains.size(), ctx->key_in.p);
    PBF_LAUNCH_CHECK(ctx);
    PBF_TRY(exclusive_scan_u32(ctx, ctx->key_in.p, ctx->key_a.p, n, ctx->mc_total_dev + 3));
    compact_kernel<<<div_up(n, kBlock), kBlock, 0, ctx->stream>>>(n, ctx->key_in.p, ctx->key_a.p, ctx->pos[ctx->cur].p,
                                                                   ctx->vel[ctx->cur].p, ctx->col[ctx->cur_col].p,
                                                                   ctx->ids[ctx->cur].p, ctx->pos[o].p, ctx->vel[o].p,
                                                                   ctx->col[oc].p, ctx->ids[o].p);
    PBF_LAUNCH_CHECK(ctx);
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->mc_total_host + 3, ctx->mc_total_dev + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n = ctx->mc_total_host[3];
    ctx->cur = o;
    ctx->cur_col = oc;
  }
  return PBF_OK;
}

// After the cell table: (first, count) of every query's cell, copied to the pinned mirror (valid after a sync).
int scene_answer_queries(pbf_ctx *ctx) {
  pbf_scene_state &sc = ctx->scene;
  const uint32_t nq = (uint32_t)sc.queries.size();
  sc.answered = 0;
  if (nq == 0) return PBF_OK;
  query_ranges_kernel<<<div_up(nq, 64), 64, 0, ctx->stream>>>(ctx->sc, nq, sc.d_queries.p, ctx->table.p, sc.d_ranges.p);
  PBF_LAUNCH_CHECK(ctx);
  PBF_CUDA(ctx, cudaMemcpyAsync(sc.h_ranges, sc.d_ranges.p, nq * sizeof(uint2), cudaMemcpyDeviceToHost, ctx->stream));
  sc.answered = nq;
  return PBF_OK;
}

// Copies the caller's scene into the context (host vectors + the small device arrays the kernels read).
int scene_set(pbf_ctx *ctx, const pbf_scene *scene) {
  pbf_scene_state &sc = ctx->scene;
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  sc.wells.clear(); sc.sources.clear(); sc.drains.clear(); sc.queries.clear();
  sc.answered = 0;
  if (!scene) return PBF_OK;
  if ((scene->n_wells && !scene->wells) || (scene->n_sources && !scene->sources) || (scene->n_drains && !scene->drains) ||
      (scene->n_queries && !scene->queries))
    return fail(ctx, PBF_ERR_INVALID, "scene", "a non-empty list has a NULL pointer");
  if (scene->n_wells > PBF_MAX_WELLS) return fail(ctx, PBF_ERR_INVALID, "scene", "more than PBF_MAX_WELLS wells");
  sc.wells.assign(scene->wells, scene->wells + scene->n_wells);
  sc.sources.assign(scene->sources, scene->sources + scene->n_sources);
  sc.drains.assign(scene->drains, scene->drains + scene->n_drains);
  sc.queries.assign(scene->queries, scene->queries + scene->n_queries);
  for (const pbf_source &s : sc.sources)
    if (!(s.rate >= 0.f) || s.rate > 1.0e6f) return fail(ctx, PBF_ERR_INVALID, "scene", "source rate must be in [0, 1e6]");
  std::vector<float4> tmp;
  auto push = [&](DevBuf<float4> &buf) -> cudaError_t {
    if (tmp.empty()) return cudaSuccess;
    cudaError_t e = buf.reserve(tmp.size());
    if (e == cudaSuccess) e = cudaMemcpy(buf.p, tmp.data(), tmp.size() * sizeof(float4), cudaMemcpyHostToDevice);
    return e;
  };
  for (const pbf_well &w : sc.wells) tmp.push_back(make_float4(w.centre[0], w.centre[1], w.centre[2], w.force));
  PBF_CUDA(ctx, push(sc.d_wells));
  tmp.clear();
  for (const pbf_drain &d : sc.drains) tmp.push_back(make_float4(d.centre[0], d.centre[1], d.centre[2], d.width));
  PBF_CUDA(ctx, push(sc.d_drains));
  tmp.clear();
  for (const pbf_query &q : sc.queries) tmp.push_back(make_float4(q.point[0], q.point[1], q.point[2], 0.f));
  PBF_CUDA(ctx, push(sc.d_queries));
  if (!sc.queries.empty()) {
    PBF_CUDA(ctx, sc.d_ranges.reserve(sc.queries.size()));
    if (sc.h_ranges_cap < sc.queries.size()) {
      if (sc.h_ranges) cudaFreeHost(sc.h_ranges);
      sc.h_ranges = nullptr;
      sc.h_ranges_cap = 0;
      PBF_CUDA(ctx, cudaHostAlloc(&sc.h_ranges, sc.queries.size() * sizeof(uint2), cudaHostAllocDefault));
      sc.h_ranges_cap = sc.queries.size();
    }
  }
  return PBF_OK;
}

void scene_release(pbf_ctx *ctx) {
  pbf_scene_state &sc = ctx->scene;
  sc.d_wells.release(); sc.d_drains.release(); sc.d_queries.release(); sc.d_ranges.release();
  if (sc.h_ranges) cudaFreeHost(sc.h_ranges);
  sc.h_ranges = nullptr;
  sc.h_ranges_cap = 0;
}

}  // namespace pbf
