// neighbour.cu — the 27-cell neighbour passes in their simplest form (one thread per particle, every candidate
// evaluated in place, reads through L1/L2):
//   neighbour_counts   parity tap: candidates / in-radius count per particle
//   diffuse            colour averaging                         (ompsph.hpp:189-206, OCL oclsph_kernel.h:67-93)
//   lambda             density + constraint lambda               (ompsph.hpp:217-232)
//   delta              position correction + clamp to the box    (ompsph.hpp:235-248), Jacobi (second buffer)
//
// Cells and particles are visited in exactly the reference's order (sph.hpp:215-236: x fastest, then y, then z;
// ascending sorted index inside a cell).  These lambda/delta kernels are the cross-check selected by
// PBF_FLAG_GLOBAL_NEIGHBOURS; the production path is the neighbour-list form in neighbour_list.cu.
#include "cells.cuh"
#include "common.cuh"
#include "pair_math.cuh"

namespace pbf {

namespace {

constexpr int kNbBlock = 128;

__global__ void __launch_bounds__(kNbBlock) neighbour_counts_kernel(StepConst c, const uint32_t *__restrict__ keys,
                                                                    const uint32_t *__restrict__ table,
                                                                    const float4 *__restrict__ pstar,
                                                                    uint32_t *__restrict__ cand,
                                                                    uint32_t *__restrict__ nbr) {
  const uint32_t a = blockIdx.x * kNbBlock + threadIdx.x;
  if (a >= c.n) return;
  const float4 pa = ldg4(pstar + a);
  uint32_t nc = 0, nn = 0;
  for_each_candidate(__ldg(keys + a), c.G, table, [&](uint32_t b) {
    const float4 pb = ldg4(pstar + b);
    ++nc;
    if (strict_distance(pa, pb) <= c.h) ++nn;
  });
  cand[a] = nc;
  nbr[a] = nn;
}

// One thread per particle; only the first particle of each cell works: every particle of a cell sees the same
// 27 cells, hence the same candidate sum in the same order, so the (bit-exact, sequentially summed) mixture is
// computed once per occupied cell and applied to the whole cell.
__global__ void __launch_bounds__(kNbBlock) diffuse_kernel(StepConst c, const uint32_t *__restrict__ keys,
                                                           const uint32_t *__restrict__ table,
                                                           const float4 *__restrict__ col_in,
                                                           float4 *__restrict__ col_out) {
  const uint32_t a = blockIdx.x * kNbBlock + threadIdx.x;
  if (a >= c.n) return;
  const uint32_t key = __ldg(keys + a);
  if (a > 0 && __ldg(keys + a - 1) == key) return;
  float mx = 0.f, my = 0.f, mz = 0.f, mw = 0.f;
  uint32_t nn = 0;
  for_each_candidate(key, c.G, table, [&](uint32_t b) {
    const float4 cb = ldg4(col_in + b);
    mx = fadd(mx, cb.x); my = fadd(my, cb.y); mz = fadd(mz, cb.z); mw = fadd(mw, cb.w);
    ++nn;
  });
  const float fn = (float)nn;
  const float t = c.diffuse_mix, omt = fsub(1.0f, t);
  const float yx = fmul(fdiv(mx, fn), 1.33f), yy = fmul(fdiv(my, fn), 1.33f);
  const float yz = fmul(fdiv(mz, fn), 1.33f), yw = fmul(fdiv(mw, fn), 1.33f);
  for (uint32_t j = a; j < c.n && __ldg(keys + j) == key; ++j) {
    float4 o = ldg4(col_in + j);
    if (nn != 0) {  // ompsph.hpp:200 (a particle outside the grid may see no cell at all)
      o.x = glm_min(glm_max(fadd(fmul(o.x, omt), fmul(yx, t)), 0.03f), 1.0f);
      o.y = glm_min(glm_max(fadd(fmul(o.y, omt), fmul(yy, t)), 0.03f), 1.0f);
      o.z = glm_min(glm_max(fadd(fmul(o.z, omt), fmul(yz, t)), 0.03f), 1.0f);
      o.w = glm_min(glm_max(fadd(fmul(o.w, omt), fmul(yw, t)), 0.03f), 1.0f);
    }
    col_out[j] = o;
  }
}

template <bool kStrict>
__global__ void __launch_bounds__(kNbBlock) lambda_kernel(StepConst c, uint32_t first, uint32_t count,
                                                          const uint32_t *__restrict__ keys,
                                                          const uint32_t *__restrict__ table,
                                                          const float4 *__restrict__ pos_mass,
                                                          const float4 *__restrict__ pstar_in,
                                                          float4 *__restrict__ pstar_out, float *__restrict__ rho_out) {
  const uint32_t t = blockIdx.x * kNbBlock + threadIdx.x;
  if (t >= count) return;
  const uint32_t a = first + t;
  const float4 pa = ldg4(pstar_in + a);
  const float mass = __ldg(&pos_mass[a].w);
  LambdaAcc<kStrict> acc;
  acc.init();
  acc.set_mass(mass);
  for_each_candidate(__ldg(keys + a), c.G, table, [&](uint32_t b) { acc.add(c, pa, ldg4(pstar_in + b)); });
  float rho;
  const float lambda = acc.finish(c, mass, rho);
  pstar_out[a] = make_float4(pa.x, pa.y, pa.z, lambda);
  if (rho_out) rho_out[a] = rho;
}

template <bool kStrict>
__global__ void __launch_bounds__(kNbBlock) delta_kernel(StepConst c, uint32_t first, uint32_t count,
                                                         const uint32_t *__restrict__ keys,
                                                         const uint32_t *__restrict__ table,
                                                         const float4 *__restrict__ pstar_in,
                                                         float4 *__restrict__ pstar_out) {
  const uint32_t t = blockIdx.x * kNbBlock + threadIdx.x;
  if (t >= count) return;
  const uint32_t a = first + t;
  const float4 pa = ldg4(pstar_in + a);
  DeltaAcc<kStrict> acc;
  acc.init();
  for_each_candidate(__ldg(keys + a), c.G, table, [&](uint32_t b) { acc.add(c, pa, ldg4(pstar_in + b)); });
  pstar_out[a] = acc.finish(c, pa);
}

}  // namespace

int launch_neighbour_counts(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *pstar,
                            uint32_t *cand, uint32_t *nbr) {
  neighbour_counts_kernel<<<div_up(ctx->sc.n, kNbBlock), kNbBlock, 0, ctx->stream>>>(ctx->sc, keys_sorted, table, pstar,
                                                                                     cand, nbr);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int launch_diffuse(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *col_in,
                   float4 *col_out) {
  PhaseScope ps(ctx, PBF_PH_DIFFUSE);
  diffuse_kernel<<<div_up(ctx->sc.n, kNbBlock), kNbBlock, 0, ctx->stream>>>(ctx->sc, keys_sorted, table, col_in, col_out);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int launch_lambda_global(pbf_ctx *ctx, uint32_t first, uint32_t count, const uint32_t *keys_sorted,
                         const uint32_t *table, const float4 *pos_mass, const float4 *pstar_in, float4 *pstar_out,
                         float *rho_out) {
  if (count == 0) return PBF_OK;
  if (ctx->flags & PBF_FLAG_STRICT_FP)
    lambda_kernel<true><<<div_up(count, kNbBlock), kNbBlock, 0, ctx->stream>>>(ctx->sc, first, count, keys_sorted, table,
                                                                               pos_mass, pstar_in, pstar_out, rho_out);
  else
    lambda_kernel<false><<<div_up(count, kNbBlock), kNbBlock, 0, ctx->stream>>>(ctx->sc, first, count, keys_sorted, table,
                                                                                pos_mass, pstar_in, pstar_out, rho_out);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int launch_delta_global(pbf_ctx *ctx, uint32_t first, uint32_t count, const uint32_t *keys_sorted,
                        const uint32_t *table, const float4 *pstar_in, float4 *pstar_out) {
  if (count == 0) return PBF_OK;
  if (ctx->flags & PBF_FLAG_STRICT_FP)
    delta_kernel<true><<<div_up(count, kNbBlock), kNbBlock, 0, ctx->stream>>>(ctx->sc, first, count, keys_sorted, table,
                                                                              pstar_in, pstar_out);
  else
    delta_kernel<false><<<div_up(count, kNbBlock), kNbBlock, 0, ctx->stream>>>(ctx->sc, first, count, keys_sorted, table,
                                                                               pstar_in, pstar_out);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

}  // namespace pbf
