// neighbour_list.cu — the production form of the solver iteration (ompsph.hpp:215-249; 84 % of the reference's step).
//
// ncu on the plain one-pass kernels (neighbour.cu) shows an ISSUE-bound kernel (84 % issue-active) running at 16 of
// 32 lanes: a particle has ~130-270 candidates in its 27 cells but only ~25-50 lie within h, every lane hits at
// different candidates, so the expensive kernel-function code runs for almost every candidate at ~20 % utilisation.
// Positions do not change between the lambda pass and the delta pass of one iteration, so the in-radius set is
// found ONCE per iteration:
//   lambda pass  phase 1: walk the 27 cells (18 contiguous runs, cells.cuh) and test every candidate — a 7-flop test,
//                         nothing else — appending each hit's index to the particle's row of a neighbour list in
//                         global memory (column layout nl[k][particle]: phase-2 reads are fully coalesced);
//                phase 2: evaluate the density / gradient sums over the hits only, all lanes busy.
//   delta pass   phase 2 only: no cell walk, no tests — reads the list the lambda pass left.
// The list keeps the candidates in the reference's visiting order, so every sum is formed in the reference's
// order (sph.hpp:215-236).  A particle with more than kListMax hits (particles piled into one cell) is flagged and
// handled by the one-pass code in both passes.  The list costs ~3 x 4 B x hits of HBM/L2 traffic per particle and
// iteration — about 0.4 GB per iteration at 1 M particles — in exchange for ~3x fewer issued instructions.
#include "cells.cuh"
#include "common.cuh"
#include "pair_math.cuh"

namespace pbf {

namespace {

constexpr int kBlock = 128;

// Predicated 32-bit global store (one @p STG, no divergent branch).
__device__ __forceinline__ void store_if(uint32_t *p, uint32_t v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.b32 [%0], %1;\n\t}" ::"l"(p), "r"(v),
      "r"((uint32_t)pred)
      : "memory");
}

template <bool kStrict>
__global__ void __launch_bounds__(kBlock) lambda_list_kernel(StepConst c, uint32_t first, uint32_t count,
                                                             const uint32_t *__restrict__ keys,
                                                             const uint32_t *__restrict__ table,
                                                             const float4 *__restrict__ pos_mass,
                                                             const float4 *__restrict__ pstar_in,
                                                             float4 *__restrict__ pstar_out, float *__restrict__ rho_out,
                                                             uint32_t *nl, uint32_t stride, uint32_t *__restrict__ n_hits,
                                                             const uint32_t *__restrict__ role, uint32_t want) {
  const uint32_t t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= count) return;
  const uint32_t a = first + t;
  if (role && !(__ldg(role + a) & want)) return;  // multi-GPU: not this pass's particle (dist.cu roles)
  const float4 pa = ldg4(pstar_in + a);
  const uint32_t key = __ldg(keys + a);
  const float mass = __ldg(&pos_mass[a].w);
  // ---- phase 1: find the hits
  // The append is written so that it compiles to a predicated store plus a predicated increment (no divergent
  // branch): with ~1 hit in 5 candidates a branch would be taken by some lane for most candidates anyway.
  uint32_t k = 0;
  uint32_t slot = a;  // index of nl[k * stride + a]; fits 32 bits (checked by the launcher)
  for_each_run(key, c.G, table, [&](uint32_t s, uint32_t e) {
#pragma unroll 4
    for (uint32_t b = s; b < e; ++b) {
      const bool hit = LambdaAcc<kStrict>::test(c, pa, ldg4(pstar_in + b));
      store_if(nl + slot, b, hit && k < kListMax);
      if (hit) { ++k; slot += stride; }
    }
  });
  n_hits[a] = k;
  // ---- phase 2: the sums
  LambdaAcc<kStrict> acc;
  acc.init();
  acc.set_mass(mass);
  if (k <= kListMax) {
    const uint32_t *row = nl + a;
#pragma unroll 4
    for (uint32_t i = 0; i < k; ++i, row += stride) acc.add_in(c, pa, ldg4(pstar_in + __ldcg(row)));
  } else {
    for_each_candidate(key, c.G, table, [&](uint32_t b) { acc.add(c, pa, ldg4(pstar_in + b)); });
  }
  float rho;
  const float lambda = acc.finish(c, mass, rho);
  pstar_out[a] = make_float4(pa.x, pa.y, pa.z, lambda);
  if (rho_out) rho_out[a] = rho;
}

template <bool kStrict>
__global__ void __launch_bounds__(kBlock) delta_list_kernel(StepConst c, uint32_t first, uint32_t count,
                                                            const uint32_t *__restrict__ keys,
                                                            const uint32_t *__restrict__ table,
                                                            const float4 *__restrict__ pstar_in,
                                                            float4 *__restrict__ pstar_out,
                                                            const uint32_t *__restrict__ nl, uint32_t stride,
                                                            const uint32_t *__restrict__ n_hits,
                                                            const uint32_t *__restrict__ role, uint32_t want) {
  const uint32_t t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= count) return;
  const uint32_t a = first + t;
  if (role && !(__ldg(role + a) & want)) return;  // multi-GPU: not this pass's particle (dist.cu roles)
  const float4 pa = ldg4(pstar_in + a);
  const uint32_t k = __ldg(n_hits + a);
  DeltaAcc<kStrict> acc;
  acc.init();
  if (k <= kListMax) {
    const uint32_t *row = nl + a;
#pragma unroll 4
    for (uint32_t i = 0; i < k; ++i, row += stride) {
      acc.add_in(c, pa, ldg4(pstar_in + __ldg(row)));  // add_in skips the particle itself (r < EPSILON)
    }
  } else {
    for_each_candidate(__ldg(keys + a), c.G, table, [&](uint32_t b) { acc.add(c, pa, ldg4(pstar_in + b)); });
  }
  pstar_out[a] = acc.finish(c, pa);
}

}  // namespace

int launch_lambda_list(pbf_ctx *ctx, uint32_t first, uint32_t count, const uint32_t *keys_sorted, const uint32_t *table,
                       const float4 *pos_mass, const float4 *pstar_in, float4 *pstar_out, float *rho_out,
                       const uint32_t *role, uint32_t want) {
  if (count == 0) return PBF_OK;
  const uint32_t n = ctx->sc.n;
  const uint32_t stride = (n + 31u) & ~31u;  // rows start on 128-byte boundaries
  if ((uint64_t)stride * (kListMax + 1) >= (1ull << 32))
    return fail(ctx, PBF_ERR_INVALID, "n", "more than 2^32 / 65 particles on one device (neighbour-list indexing)");
  PBF_CUDA(ctx, ctx->nl.reserve((size_t)stride * kListMax));
  PBF_CUDA(ctx, ctx->nl_count.reserve(n));
  ctx->nl_stride = stride;
  if (ctx->flags & PBF_FLAG_STRICT_FP)
    lambda_list_kernel<true><<<div_up(count, kBlock), kBlock, 0, ctx->stream>>>(
        ctx->sc, first, count, keys_sorted, table, pos_mass, pstar_in, pstar_out, rho_out, ctx->nl.p, stride, ctx->nl_count.p,
        role, want);
  else
    lambda_list_kernel<false><<<div_up(count, kBlock), kBlock, 0, ctx->stream>>>(
        ctx->sc, first, count, keys_sorted, table, pos_mass, pstar_in, pstar_out, rho_out, ctx->nl.p, stride, ctx->nl_count.p,
        role, want);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int launch_delta_list(pbf_ctx *ctx, uint32_t first, uint32_t count, const uint32_t *keys_sorted, const uint32_t *table,
                      const float4 *pstar_in, float4 *pstar_out, const uint32_t *role, uint32_t want) {
  if (count == 0) return PBF_OK;
  if (ctx->flags & PBF_FLAG_STRICT_FP)
    delta_list_kernel<true><<<div_up(count, kBlock), kBlock, 0, ctx->stream>>>(
        ctx->sc, first, count, keys_sorted, table, pstar_in, pstar_out, ctx->nl.p, ctx->nl_stride, ctx->nl_count.p,
        role, want);
  else
    delta_list_kernel<false><<<div_up(count, kBlock), kBlock, 0, ctx->stream>>>(
        ctx->sc, first, count, keys_sorted, table, pstar_in, pstar_out, ctx->nl.p, ctx->nl_stride, ctx->nl_count.p,
        role, want);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

}  // namespace pbf
