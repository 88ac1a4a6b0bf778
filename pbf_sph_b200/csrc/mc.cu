// mc.cu — optional marching-cubes surface extraction (SphParams::surface), ompsph.hpp:277-477:
//   mc_field   scalar field + normal + colour at every lattice point                (ompsph.hpp:288-356)
//   mc_count   triangles per cube                                                   (ompsph.hpp:372-393)
//   scan       exclusive prefix sum of the counts -> each cube's first triangle slot
//   mc_emit    edge interpolation and triangle output                               (ompsph.hpp:407-474)
// The reference hands out triangle slots with a global atomic counter (order depends on thread timing); the scan
// makes the order deterministic: ascending cube index, which is also what one reference thread produces.
// It uses the post-finalise positions with the cell table built before the solve (stale grid), as the reference does.
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "pbf/mc_tables.h"

namespace pbf {

namespace {

constexpr int kMcBlock = 128;

__constant__ uint64_t c_tri[256] = PBF_MC_TRI_WORDS_INIT;
__constant__ int c_corner[8][3] = PBF_MC_CORNER_OFFSETS_INIT;
__constant__ int c_edge[12][2] = PBF_MC_EDGE_CORNERS_INIT;

__device__ __forceinline__ float glm_fast_sqrt(float x) { return fdiv(1.0f, fdiv(1.0f, fsqrt(x))); }  // gtx/fast_square_root

// Lattice points are taken in 4 x 4 x 8 tiles (x, y, z), one tile per 128-thread block, so that a warp's 32 points
// (4 in y, 8 in z, one x) span only 2 x 4 cells at the stock resolution of 2: similar candidate counts in every lane
// and neighbouring table / particle reads.  Each point works in two phases like the solver's neighbour list:
//   1. the 27-cell walk tests every candidate on the squared distance only and appends the hits to the thread's list
//      in shared memory.  The reference tests  glm::fastSqrt(r2) < threshold  (fastSqrt = 1 / (1 / sqrt)): that map is
//      monotone in r2, so the host finds the largest r2 that passes (McConst::r2_hit) and the test is one compare;
//   2. the expensive part (two IEEE divides and a sqrt for the distance, powf, four divides, the colour gather) runs
//      over the hits only, every lane busy, in the order they were found — so every sum is the reference's sum.
// ncu on the one-phase form: 1.68 G warp instructions at 12 of 32 lanes, 90 % issue-bound, the hit code at 5 lanes.
constexpr int kFieldCap = 48;    // hits a point can hold; a full list is evaluated on the spot (rare: one lane works)
constexpr int kFieldFlush = 32;  // after each cell: if ANY lane holds this many, the whole warp evaluates its lists

// [key_lo, key_hi): the slab path's ownership filter — a rank evaluates the lattice points whose cell (clamped into the
// grid) it owns, and stores them into rank 0's lattice; one device passes [0, 2^32).
__global__ void __launch_bounds__(kMcBlock) mc_field_kernel(StepConst c, McConst m, const uint32_t *__restrict__ table,
                                                            const float4 *__restrict__ pos,
                                                            const float4 *__restrict__ col, float4 *__restrict__ PN,
                                                            float4 *__restrict__ LC, uint32_t key_lo, uint32_t key_hi) {
  __shared__ uint32_t s_list[kFieldCap * kMcBlock];  // [slot][thread]: conflict-free
  const uint32_t tid = threadIdx.x;
  const uint32_t ntz = (m.sample[2] + 7u) / 8u, nty = (m.sample[1] + 3u) / 4u;
  const uint32_t bz = blockIdx.x % ntz, by = (blockIdx.x / ntz) % nty, bx = blockIdx.x / (ntz * nty);
  const uint32_t x = bx * 4u + (tid >> 5), y = by * 4u + ((tid >> 3) & 3u), z = bz * 8u + (tid & 7u);
  // no early exit: every lane walks the 27 cells (lanes without a lattice point see empty cells), so the warp is
  // converged at the end of each cell and can decide together when to evaluate
  bool alive = x < m.sample[0] && y < m.sample[1] && z < m.sample[2];
  const uint64_t idx = ((uint64_t)x * m.sample[1] + y) * m.sample[2] + z;  // index3d, curves.h:17-19
  const float lp[3] = {(float)x, (float)y, (float)z};
  float a[3];
  int c0[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    a[k] = fmul(fadd(c.min_extent[k], fmul(lp[k], m.step)), c.scale);                       // ompsph.hpp:291
    c0[k] = (int)compact10(spread10(cell_coord(fdiv(lp[k], m.resolution))));                // ompsph.hpp:294-299
  }
  if (alive && key_hi - key_lo != 0xFFFFFFFFu) {  // slab path: is this lattice point mine?
    const uint32_t cc[3] = {min((uint32_t)c0[0], c.extent[0] - 1u), min((uint32_t)c0[1], c.extent[1] - 1u),
                            min((uint32_t)c0[2], c.extent[2] - 1u)};
    const uint32_t key = spread10(cc[0]) | (spread10(cc[1]) << 1) | (spread10(cc[2]) << 2);
    alive = key >= key_lo && key < key_hi;
  }
  if (alive && (uint32_t)c0[0] == c.extent[0] && (uint32_t)c0[1] == c.extent[1] && (uint32_t)c0[2] == c.extent[2]) {
    PN[idx] = make_float4(0.f, 0.f, 0.f, 0.f);  // ompsph.hpp:301-304: the entry keeps its zero initialisation
    LC[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    alive = false;
  }
  int nb[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // glm::clamp(int): ompsph.hpp:306-311
    const int top = (int)c.extent[k] - 1;
    nb[k][0] = min(max(c0[k] - 1, 0), top);
    nb[k][1] = c0[k];
    nb[k][2] = min(max(c0[k] + 1, 0), top);
  }
  float v = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  float cx = 0.f, cy = 0.f, cz = 0.f, cw = 0.f;
  uint32_t nn = 0, held = 0;
  const float w = fmul(-m.particle_influence, m.particle_size);
  const bool half_power = m.particle_influence == 0.5f;
  uint32_t *mine = s_list + tid;
  auto evaluate = [&]() {  // the sums of ompsph.hpp:329-345 over the held hits, in the order they were found
    for (uint32_t i = 0; i < held; ++i) {
      const uint32_t b = mine[i * kMcBlock];
      const float4 pb = ldg4(pos + b);
      const float ex = fsub(a[0], pb.x), ey = fsub(a[1], pb.y), ez = fsub(a[2], pb.z);
      const float d = glm_fast_sqrt(fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(ez, ez)));
      const float lx = fsub(pb.x, a[0]), ly = fsub(pb.y, a[1]), lz = fsub(pb.z, a[2]);
      // glm::pow(len, influence); |l| == d bit for bit (squares of negated terms).  The stock influence is 0.5, where
      // the correctly rounded square root is at least as close to glibc's powf as CUDA's powf is (both <= 1 ulp)
      const float den = half_power ? fsqrt(d) : powf(d, m.particle_influence);
      v = fadd(v, fdiv(m.particle_size, den));
      nx = fadd(nx, fmul(w, fdiv(lx, den)));
      ny = fadd(ny, fmul(w, fdiv(ly, den)));
      nz = fadd(nz, fmul(w, fdiv(lz, den)));
      const float4 cb = ldg4(col + b);
      cx = fadd(cx, cb.x); cy = fadd(cy, cb.y); cz = fadd(cz, cb.z); cw = fadd(cw, cb.w);
    }
    nn += held;
    held = 0;
  };
#pragma unroll 1
  for (int kz = 0; kz < 3; ++kz)
#pragma unroll 1
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll 1
      for (int kx = 0; kx < 3; ++kx) {  // order of ompsph.hpp:313-326
        const uint32_t o = morton3((uint32_t)nb[0][kx], (uint32_t)nb[1][ky], (uint32_t)nb[2][kz]);
        uint32_t s = 0, e = 0;
        if (alive && o < c.G) {
          s = __ldg(table + o);
          e = (o + 1 < c.G) ? __ldg(table + o + 1) : s;
        }
        for (uint32_t b = s; b < e; b += 4u) {  // four positions in flight per lane
          float4 pb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) pb[u] = ldg4(pos + min(b + u, e - 1u));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float ex = fsub(a[0], pb[u].x), ey = fsub(a[1], pb[u].y), ez = fsub(a[2], pb[u].z);
            if (b + u < e && fadd(fadd(fmul(ex, ex), fmul(ey, ey)), fmul(ez, ez)) <= m.r2_hit) {  // fastSqrt(r2) < threshold
              mine[held * kMcBlock] = b + u;
              if (++held == kFieldCap) evaluate();
            }
          }
        }
        if (__any_sync(0xFFFFFFFFu, held >= kFieldFlush)) evaluate();
      }
  evaluate();
  if (!alive) return;
  const float inv = fdiv(1.0f, fsqrt(fadd(fadd(fmul(nx, nx), fmul(ny, ny)), fmul(nz, nz))));  // fastNormalize
  PN[idx] = make_float4(v, fmul(nx, inv), fmul(ny, inv), fmul(nz, inv));
  const float fn = (float)nn;
  LC[idx] = make_float4(fdiv(cx, fn), fdiv(cy, fn), fdiv(cz, fn), fdiv(cw, fn));
}

struct Cube {
  uint32_t x, y, z;
};
__device__ __forceinline__ Cube cube_of(uint64_t i, const McConst &m) {  // utils.hpp:73-79
  const uint64_t yz = (uint64_t)m.march[1] * m.march[2];
  Cube q;
  q.x = (uint32_t)(i / yz);
  q.y = (uint32_t)((i - q.x * yz) / m.march[2]);
  q.z = (uint32_t)(i - q.x * yz - (uint64_t)q.y * m.march[2]);
  return q;
}
__device__ __forceinline__ uint64_t lattice_index(const McConst &m, uint32_t x, uint32_t y, uint32_t z) {  // curves.h:17-19
  return (uint64_t)x * m.sample[1] * m.sample[2] + (uint64_t)y * m.sample[2] + z;
}

__global__ void __launch_bounds__(kMcBlock) mc_count_kernel(McConst m, const float4 *__restrict__ PN,
                                                            uint32_t *__restrict__ counts) {
  const uint64_t i = (uint64_t)blockIdx.x * kMcBlock + threadIdx.x;
  if (i >= m.march_n) return;
  const Cube q = cube_of(i, m);
  uint32_t ci = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float v = __ldg(&PN[lattice_index(m, q.x + c_corner[k][0], q.y + c_corner[k][1], q.z + c_corner[k][2])].x);
    if (v < m.isolevel) ci |= 1u << k;
  }
  const uint64_t row = c_tri[ci];
  counts[i] = pbf_mc_edge_mask(row) == 0 ? 0u : pbf_mc_num_verts(row) / 3u;  // ompsph.hpp:388
}

__global__ void __launch_bounds__(kMcBlock) mc_emit_kernel(StepConst c, McConst m, const float4 *__restrict__ PN,
                                                           const float4 *__restrict__ LC,
                                                           const uint32_t *__restrict__ counts,
                                                           const uint32_t *__restrict__ offsets, float *__restrict__ vs,
                                                           float *__restrict__ ns, float *__restrict__ cs) {
  const uint64_t i = (uint64_t)blockIdx.x * kMcBlock + threadIdx.x;
  if (i >= m.march_n) return;
  const uint32_t ntri = __ldg(counts + i);
  if (ntri == 0) return;
  const Cube q = cube_of(i, m);
  float val[8];
  uint64_t li[8];
  uint32_t ci = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    li[k] = lattice_index(m, q.x + c_corner[k][0], q.y + c_corner[k][1], q.z + c_corner[k][2]);
    val[k] = __ldg(&PN[li[k]].x);
    if (val[k] < m.isolevel) ci |= 1u << k;
  }
  const uint64_t row = c_tri[ci];
  uint64_t slot = (uint64_t)__ldg(offsets + i) * 3;  // first vertex of this cube's first triangle
  for (uint32_t t = 0; t < ntri * 3; ++t, ++slot) {
    const uint32_t e = (uint32_t)((row >> (4 * t)) & 0xF);
    const int from = c_edge[e][0], to = c_edge[e][1];
    const float tt = fdiv(fsub(m.isolevel, val[from]), fsub(val[to], val[from]));  // utils.hpp:85
    const float omt = fsub(1.0f, tt);
    const float4 pf = ldg4(PN + li[from]), pt = ldg4(PN + li[to]);
    const float4 cf = ldg4(LC + li[from]), ct = ldg4(LC + li[to]);
    const uint32_t cq[3] = {q.x, q.y, q.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // ompsph.hpp:424, glm::mix = x*(1-t) + y*t
      const float of = fmul(fadd(c.min_extent[k], fmul((float)(cq[k] + c_corner[from][k]), m.step)), c.scale);
      const float ot = fmul(fadd(c.min_extent[k], fmul((float)(cq[k] + c_corner[to][k]), m.step)), c.scale);
      vs[slot * 3 + k] = fadd(fmul(of, omt), fmul(ot, tt));
    }
    ns[slot * 3 + 0] = fadd(fmul(pf.y, omt), fmul(pt.y, tt));
    ns[slot * 3 + 1] = fadd(fmul(pf.z, omt), fmul(pt.z, tt));
    ns[slot * 3 + 2] = fadd(fmul(pf.w, omt), fmul(pt.w, tt));
    cs[slot * 4 + 0] = fadd(fmul(cf.x, omt), fmul(ct.x, tt));
    cs[slot * 4 + 1] = fadd(fmul(cf.y, omt), fmul(ct.y, tt));
    cs[slot * 4 + 2] = fadd(fmul(cf.z, omt), fmul(ct.z, tt));
    cs[slot * 4 + 3] = fadd(fmul(cf.w, omt), fmul(ct.w, tt));
  }
}

}  // namespace

// constants of this step's surface (McParams -> McConst) and the lattice sizes; no device work
int mc_prepare(pbf_ctx *ctx, const pbf_params &p) {
  McConst &m = ctx->mc;
  m.resolution = p.surface.resolution;
  m.isolevel = p.surface.isolevel;
  m.particle_size = p.surface.particle_size;
  m.particle_influence = p.surface.particle_influence;
  m.step = ctx->h / p.surface.resolution;  // ompsph.hpp:290
  m.threshold = ctx->h * p.scale * 1;      // ompsph.hpp:292
  {  // largest r2 with glm::fastSqrt(r2) < threshold (fastSqrt = 1 / (1 / sqrt(x)) is monotone): bisection on the bits
    auto passes = [&](float x) { return 1.0f / (1.0f / std::sqrt(x)) < m.threshold; };
    uint32_t lo = 0u, hi = 0x7f7fffffu;  // passes(lo) holds for threshold > 0; find the last bit pattern that passes
    m.r2_hit = -1.0f;                     // nothing passes (threshold <= 0)
    float f;
    std::memcpy(&f, &lo, 4);
    if (passes(f)) {
      while (lo < hi) {
        const uint32_t mid = lo + (hi - lo + 1u) / 2u;
        std::memcpy(&f, &mid, 4);
        if (passes(f)) lo = mid; else hi = mid - 1u;
      }
      std::memcpy(&m.r2_hit, &lo, 4);
    }
  }
  for (int a = 0; a < 3; ++a) {
    m.sample[a] = ctx->grid.sample_size[a];
    m.march[a] = m.sample[a] - 1;
  }
  m.lattice_n = (uint64_t)m.sample[0] * m.sample[1] * m.sample[2];
  m.march_n = (uint64_t)m.march[0] * m.march[1] * m.march[2];
  if (m.lattice_n >= (1ull << 32)) return fail(ctx, PBF_ERR_INVALID, "surface", "lattice too large");
  const uint64_t tiles = (uint64_t)((m.sample[0] + 3u) / 4u) * ((m.sample[1] + 3u) / 4u) * ((m.sample[2] + 7u) / 8u);
  if (tiles >= (1ull << 31)) return fail(ctx, PBF_ERR_INVALID, "surface", "lattice too large");
  return PBF_OK;
}

// scalar field / normal / colour of the lattice points whose cell key lies in [key_lo, key_hi) -> PN, LC (lattice_n each)
int mc_field(pbf_ctx *ctx, const uint32_t *table, const float4 *pos, const float4 *col, float4 *PN, float4 *LC, uint32_t key_lo,
             uint32_t key_hi) {
  const McConst &m = ctx->mc;
  PhaseScope ps(ctx, PBF_PH_MC_FIELD);
  const uint64_t tiles = (uint64_t)((m.sample[0] + 3u) / 4u) * ((m.sample[1] + 3u) / 4u) * ((m.sample[2] + 7u) / 8u);
  mc_field_kernel<<<(unsigned)tiles, kMcBlock, 0, ctx->stream>>>(ctx->sc, m, table, pos, col, PN, LC, key_lo, key_hi);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

// triangles from a complete lattice: count, scan, emit (deterministic order: ascending cube index)
int mc_extract(pbf_ctx *ctx, const float4 *PN, const float4 *LC) {
  const McConst &m = ctx->mc;
  ctx->mc_valid = true;
  ctx->mc_total_host[0] = 0;
  if (m.march_n == 0) { ctx->n_triangles = 0; return PBF_OK; }
  PBF_CUDA(ctx, ctx->mc_count.reserve(m.march_n + 1));
  PBF_CUDA(ctx, ctx->mc_offset.reserve(m.march_n + 1));
  {
    PhaseScope ps(ctx, PBF_PH_MC_COUNT_SCAN);
    mc_count_kernel<<<div_up(m.march_n, kMcBlock), kMcBlock, 0, ctx->stream>>>(m, PN, ctx->mc_count.p);
    PBF_LAUNCH_CHECK(ctx);
    PBF_TRY(exclusive_scan_u32(ctx, ctx->mc_count.p, ctx->mc_offset.p, m.march_n, ctx->mc_total_dev));
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->mc_total_host, ctx->mc_total_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                  ctx->stream));
  }
  // the mesh size is data dependent: the host needs the total before it can size the output (the reference
  // sums the per-group counts on the host at the same point, ompsph.hpp:393-403)
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->n_triangles = ctx->mc_total_host[0];
  ctx->grid.n_triangles = (uint32_t)ctx->n_triangles;
  if (ctx->n_triangles == 0) return PBF_OK;
  PBF_CUDA(ctx, ctx->mesh_vs.reserve(ctx->n_triangles * 9));
  PBF_CUDA(ctx, ctx->mesh_ns.reserve(ctx->n_triangles * 9));
  PBF_CUDA(ctx, ctx->mesh_cs.reserve(ctx->n_triangles * 12));
  {
    PhaseScope ps(ctx, PBF_PH_MC_EMIT);
    mc_emit_kernel<<<div_up(m.march_n, kMcBlock), kMcBlock, 0, ctx->stream>>>(ctx->sc, m, PN, LC, ctx->mc_count.p, ctx->mc_offset.p,
                                                                             ctx->mesh_vs.p, ctx->mesh_ns.p, ctx->mesh_cs.p);
    PBF_LAUNCH_CHECK(ctx);
  }
  return PBF_OK;
}

int mc_run(pbf_ctx *ctx, const pbf_params &p, const uint32_t *table, const float4 *pos, const float4 *col) {
  PBF_TRY(mc_prepare(ctx, p));
  PBF_CUDA(ctx, ctx->mc_pn.reserve(ctx->mc.lattice_n));
  PBF_CUDA(ctx, ctx->mc_c.reserve(ctx->mc.lattice_n));
  ctx->mc_lattice_pn = ctx->mc_pn.p;
  ctx->mc_lattice_c = ctx->mc_c.p;
  PBF_TRY(mc_field(ctx, table, pos, col, ctx->mc_pn.p, ctx->mc_c.p, 0u, 0xFFFFFFFFu));
  return mc_extract(ctx, ctx->mc_pn.p, ctx->mc_c.p);
}

}  // namespace pbf
