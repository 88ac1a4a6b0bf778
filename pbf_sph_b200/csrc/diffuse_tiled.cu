// diffuse_tiled.cu — colour diffusion (ompsph.hpp:189-206, double-buffered like the reference's own OpenCL kernel
// oclsph_kernel.h:67-93) as a shared-memory tiled kernel over Morton cell blocks.
//
// The mixture of a particle is the plain sum of the colours of ALL candidates in its 27 cells (no distance test), so
// it depends only on the particle's cell.  Unit of work: one aligned 4x4x4 block of cells = 64 consecutive Morton
// keys.  A CTA looks up the block's 6x6x6 halo of cells in the global cell table (216 ranges), scans the counts and
// stages the halo's colours into shared memory re-laid out row-major with x fastest, so that the three x-neighbours
// of a cell are contiguous; one thread per occupied cell then forms the 27-cell sum sequentially over 9 contiguous
// shared-memory ranges — in exactly the reference's order (sph.hpp:215-236), so colours stay bit-identical to the
// oracle — and applies it to every particle of its cell.  CTAs are persistent and pull occupied blocks from a
// device-side queue.  Blocks whose halo exceeds the tile, and particles with keys >= G-1, take the global path.
//
// (The lambda/delta passes were also tried in this tiled form; under ncu the per-thread hit lists plus the tile left
// only 12 warps/SM and the kernel became latency-bound at 2x the time of the neighbour-list form, so they were
// dropped — see DESIGN.md §5.)
#include "cells.cuh"
#include "common.cuh"

namespace pbf {

namespace {

constexpr int kThreads = 128;
constexpr int kCap = 3072;       // halo particles staged per block (48 KB)
constexpr int kHaloCells = 216;  // 6*6*6

struct TiledArgs {
  const uint32_t *keys;
  const uint32_t *table;
  const float4 *col_in;
  float4 *col_out;
  const uint32_t *blk_list;   // occupied blocks
  const uint32_t *blk_count;
  uint32_t *queue;            // work counter of this launch (zeroed once per step)
  const uint32_t *n_end;      // slab path: {first, count} of the local array in device memory; else nullptr
};

struct __align__(16) Smem {
  float4 sp[kCap];
  uint32_t cell_start[kHaloCells + 1];  // exclusive scan of the halo cell populations (row-major halo order)
  uint32_t cell_src[kHaloCells];        // global index of each halo cell's first particle
  uint32_t ws0[kThreads / 32], ws1[kThreads / 32];
  uint32_t work;
};

// Fills cell_start / cell_src and, when the halo fits, sp[].  Returns the halo population.
__device__ __forceinline__ uint32_t stage_halo(Smem &sm, const StepConst &c, const TiledArgs &g, uint32_t x0, uint32_t y0,
                                               uint32_t z0) {
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 216 cells over 128 threads: thread t owns halo cells t and t+128 (if < 216)
  uint32_t cnt[2], src[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t l = tid + k * kThreads;
    cnt[k] = 0;
    src[k] = 0;
    if (l < kHaloCells) {
      const uint32_t lx = l % 6, ly = (l / 6) % 6, lz = l / 36;
      uint32_t s, e;
      cell_range(g.table, c.G, morton3(x0 + lx - 1u, y0 + ly - 1u, z0 + lz - 1u), s, e);
      cnt[k] = e - s;
      src[k] = s;
    }
  }
  // exclusive scan in halo-cell order l = 0..215: first all k=0 cells (0..127), then the k=1 cells (128..215)
  uint32_t incl0 = cnt[0], incl1 = cnt[1];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t0 = __shfl_up_sync(0xFFFFFFFFu, incl0, d), t1 = __shfl_up_sync(0xFFFFFFFFu, incl1, d);
    if (lane >= (unsigned)d) { incl0 += t0; incl1 += t1; }
  }
  if (lane == 31) { sm.ws0[warp] = incl0; sm.ws1[warp] = incl1; }
  __syncthreads();
  uint32_t base0 = 0, base1 = 0, total0 = 0, total1 = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    if (w < (int)warp) { base0 += sm.ws0[w]; base1 += sm.ws1[w]; }
    total0 += sm.ws0[w];
    total1 += sm.ws1[w];
  }
  sm.cell_start[tid] = base0 + incl0 - cnt[0];
  sm.cell_src[tid] = src[0];
  if (tid + kThreads < kHaloCells) {
    sm.cell_start[tid + kThreads] = total0 + base1 + incl1 - cnt[1];
    sm.cell_src[tid + kThreads] = src[1];
  }
  const uint32_t total = total0 + total1;
  if (tid == 0) sm.cell_start[kHaloCells] = total;
  __syncthreads();
  if (total <= kCap) {
    // flattened copy: slot j belongs to the halo cell found by binary search in cell_start
    for (uint32_t j = tid; j < total; j += kThreads) {
      uint32_t lo = 0, hi = kHaloCells;  // largest l with cell_start[l] <= j
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sm.cell_start[mid] <= j) lo = mid; else hi = mid;
      }
      sm.sp[j] = ldg4(g.col_in + sm.cell_src[lo] + (j - sm.cell_start[lo]));
    }
  }
  __syncthreads();
  return total;
}

__device__ __forceinline__ float4 diffuse_apply(const StepConst &c, float4 o, float4 y, uint32_t nn) {
  if (nn != 0) {  // ompsph.hpp:200
    const float t = c.diffuse_mix, omt = fsub(1.0f, t);  // glm::mix(x, y, t) = x*(1-t) + y*t, then clamp(0.03, 1)
    o.x = glm_min(glm_max(fadd(fmul(o.x, omt), fmul(y.x, t)), 0.03f), 1.0f);
    o.y = glm_min(glm_max(fadd(fmul(o.y, omt), fmul(y.y, t)), 0.03f), 1.0f);
    o.z = glm_min(glm_max(fadd(fmul(o.z, omt), fmul(y.z, t)), 0.03f), 1.0f);
    o.w = glm_min(glm_max(fadd(fmul(o.w, omt), fmul(y.w, t)), 0.03f), 1.0f);
  }
  return o;
}
__device__ __forceinline__ float4 diffuse_target(float4 m, uint32_t nn) {  // (mixture / n) * 1.33 — ompsph.hpp:202
  const float fn = (float)nn;
  return make_float4(fmul(fdiv(m.x, fn), 1.33f), fmul(fdiv(m.y, fn), 1.33f), fmul(fdiv(m.z, fn), 1.33f),
                     fmul(fdiv(m.w, fn), 1.33f));
}

__device__ __forceinline__ void diffuse_block_smem(Smem &sm, const StepConst &c, const TiledArgs &g) {
  for (uint32_t cell = threadIdx.x; cell < 64; cell += kThreads) {  // one thread per own cell (cx,cy,cz) in 0..3
    const uint32_t cx = cell & 3, cy = (cell >> 2) & 3, cz = cell >> 4;
    const uint32_t lc = ((cz + 1) * 6 + (cy + 1)) * 6 + (cx + 1);
    const uint32_t s = sm.cell_start[lc], e = sm.cell_start[lc + 1];
    if (s == e) continue;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t nn = 0;
#pragma unroll 1
    for (int row = 0; row < 9; ++row) {
      const int dz = row / 3 - 1, dy = row % 3 - 1;
      const uint32_t l0 = ((cz + 1 + dz) * 6 + (cy + 1 + dy)) * 6 + cx;
      const uint32_t j1 = sm.cell_start[l0 + 3];
      for (uint32_t j = sm.cell_start[l0]; j < j1; ++j) {
        const float4 cb = sm.sp[j];
        m.x = fadd(m.x, cb.x); m.y = fadd(m.y, cb.y); m.z = fadd(m.z, cb.z); m.w = fadd(m.w, cb.w);
        ++nn;
      }
    }
    const float4 y = diffuse_target(m, nn);
    const uint32_t dst = sm.cell_src[lc];
    for (uint32_t j = s; j < e; ++j) g.col_out[dst + (j - s)] = diffuse_apply(c, sm.sp[j], y, nn);
  }
}

__device__ __forceinline__ void diffuse_range_global(const StepConst &c, const TiledArgs &g, uint32_t first, uint32_t last,
                                                     uint32_t n) {
  for (uint32_t a = first + threadIdx.x; a < last; a += kThreads) {
    const uint32_t key = __ldg(g.keys + a);
    if (a > 0 && __ldg(g.keys + a - 1) == key) continue;  // only the first particle of a cell works
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t nn = 0;
    for_each_candidate(key, c.G, g.table, [&](uint32_t b) {
      const float4 cb = ldg4(g.col_in + b);
      m.x = fadd(m.x, cb.x); m.y = fadd(m.y, cb.y); m.z = fadd(m.z, cb.z); m.w = fadd(m.w, cb.w);
      ++nn;
    });
    const float4 y = diffuse_target(m, nn);
    for (uint32_t j = a; j < n && __ldg(g.keys + j) == key; ++j) g.col_out[j] = diffuse_apply(c, ldg4(g.col_in + j), y, nn);
  }
}

__global__ void __launch_bounds__(kThreads) diffuse_tiled_kernel(StepConst c, TiledArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
  // Particles with key >= G-1 take the global path: keys >= G are outside the grid (in no cell, but still processed
  // as `a`), and cell G-1 is never visible as a neighbour cell (sph.hpp:203-213), so its own particles cannot be
  // found through the staged halo.
  // (slab path: g.n_end = one past the last particle of the local array, in device memory; the table holds absolute indices)
  const uint32_t n_end = g.n_end ? __ldg(g.n_end) + __ldg(g.n_end + 1) : c.n;
  const uint32_t n_in = __ldg(g.table + (c.G - 1u));
  if (blockIdx.x == 0 && n_in < n_end) diffuse_range_global(c, g, n_in, n_end, n_end);
  const uint32_t n_work = __ldg(g.blk_count);
  while (true) {
    __syncthreads();  // the previous block's shared memory is no longer read
    if (threadIdx.x == 0) sm.work = atomicAdd(g.queue, 1u);
    __syncthreads();
    const uint32_t w = sm.work;
    if (w >= n_work) break;
    const uint32_t k0 = __ldg(g.blk_list + w) << 6;
    const uint32_t own_s = __ldg(g.table + k0);
    const uint32_t own_e = __ldg(g.table + min(k0 + 64u, c.G - 1u));
    const uint32_t total = stage_halo(sm, c, g, compact10(k0), compact10(k0 >> 1), compact10(k0 >> 2));
    if (total <= kCap) diffuse_block_smem(sm, c, g);
    else diffuse_range_global(c, g, own_s, own_e, n_in);
  }
}

// occupied-block queue
__global__ void __launch_bounds__(256) build_block_list_kernel(const uint32_t *__restrict__ table, uint32_t G,
                                                               uint32_t n_blocks, uint32_t *__restrict__ list,
                                                               uint32_t *__restrict__ ctl,
                                                               const uint32_t *__restrict__ own_range) {
  const uint32_t b = blockIdx.x * 256 + threadIdx.x;
  if (b >= n_blocks) return;
  const uint32_t k0 = b << 6;
  const uint32_t s = __ldg(table + k0), e = __ldg(table + min(k0 + 64u, G - 1u));
  bool take = e > s;
  if (take && own_range) {  // slab path: a block made of ghosts only is somebody else's work
    const uint32_t first = __ldg(own_range), last = first + __ldg(own_range + 1);
    take = s < last && e > first;
  }
  if (take) list[atomicAdd(ctl, 1u)] = b;
}

}  // namespace

int launch_diffuse_tiled(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *col_in,
                         float4 *col_out, const uint32_t *own_range_dev) {
  PhaseScope ps(ctx, PBF_PH_DIFFUSE);
  const uint32_t n_blocks = (ctx->sc.G - 1u + 63u) / 64u;  // blocks covering keys [0, G-1)
  PBF_CUDA(ctx, ctx->blk_list.reserve(n_blocks + 1));
  PBF_CUDA(ctx, ctx->blk_info.reserve(2));
  PBF_CUDA(ctx, cudaMemsetAsync(ctx->blk_info.p, 0, 2 * sizeof(uint32_t), ctx->stream));
  if (n_blocks) {
    build_block_list_kernel<<<div_up(n_blocks, 256), 256, 0, ctx->stream>>>(table, ctx->sc.G, n_blocks, ctx->blk_list.p,
                                                                            ctx->blk_info.p, own_range_dev);
    PBF_LAUNCH_CHECK(ctx);
  }
  // Function attributes are per device: the opt-in to > 48 KB of dynamic shared memory and the occupancy are set up once
  // per CONTEXT (a process may drive several devices: pbf_dist_init_local with distinct ordinals, or two Solvers).
  if (ctx->diffuse_blocks_per_sm == 0) {
    int per_sm = 0;
    PBF_CUDA(ctx, cudaFuncSetAttribute(diffuse_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    PBF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, diffuse_tiled_kernel, kThreads, sizeof(Smem)));
    ctx->diffuse_blocks_per_sm = per_sm < 1 ? 1 : per_sm;
  }
  const int per_sm = ctx->diffuse_blocks_per_sm;
  TiledArgs g{keys_sorted, table, col_in, col_out, ctx->blk_list.p, ctx->blk_info.p, ctx->blk_info.p + 1, ctx->diffuse_local_range};
  diffuse_tiled_kernel<<<(unsigned)(ctx->sm_count * per_sm), kThreads, sizeof(Smem), ctx->stream>>>(ctx->sc, g);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

}  // namespace pbf
