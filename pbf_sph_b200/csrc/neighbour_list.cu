// neighbour_list.cu — the production form of the solver iteration (ompsph.hpp:215-249; 84 % of the reference's step).
//
// ncu on the plain one-pass kernels (neighbour.cu) shows an ISSUE-bound kernel running at 16 of 32 lanes: a particle
// has ~130-270 candidates in its 27 cells but only ~25-50 lie within h, every lane hits at different candidates, so
// the expensive kernel-function code runs for almost every candidate at ~20 % utilisation.  Positions do not change
// between the lambda pass and the delta pass of one iteration, so the in-radius set is found ONCE per iteration:
//
//   lambda pass  phase 1  walk the 27 cells as 9 (y,z) rows of two contiguous runs each (an even-x cell and its +x
//                         neighbour are consecutive Morton keys); the cell-table look-ups of row r+1 are issued
//                         before row r is scanned.  Every candidate gets the 7-flop distance test and nothing else;
//                         a hit appends the candidate's index to the particle's column of the list in global memory
//                         (nl[k][particle]: lanes with equal k write one 128-byte line) with ONE predicated store
//                         and a predicated 64-bit pointer bump — no divergent branch, no per-candidate bounds test.
//                phase 2  density / gradient sums over the hits only.
//   delta pass            no cell walk, no tests: reads the list the lambda pass left.
//
// (A shared-memory list — 2 instructions per append, indices for phase 2 from shared memory — was measured too: at
// 256 B per thread it leaves 24 warps and 64 KB of L1 per SM, and the candidate gathers, which live on the L1 hit
// rate, made the kernel 20 % slower than this form.)
//
// The list keeps the candidates in the reference's visiting order, so every sum is formed in the reference's
// order (sph.hpp:215-236).  A particle with more than kCap hits (particles piled into one cell) is flagged and
// handled by the one-pass code in both passes.
#include <algorithm>

#include "cells.cuh"
#include "common.cuh"
#include "pair_math.cuh"

namespace pbf {

namespace {

#ifndef PBF_NL_BLOCK
#define PBF_NL_BLOCK 256
#endif
constexpr int kBlock = PBF_NL_BLOCK;  // lambda pass
constexpr int kBlockD = 128;          // delta pass (larger blocks measured slower at every size)
// Layout of the hit list: particles in chunks of kChunk; inside a chunk hit k of particle a lives at
//   nl[(a / kChunk) * kChunk * (cap + 1) + k * kChunk + a % kChunk]
// so the 32 lanes of a warp still write / read one 128-byte line per row, but the rows of a particle are kChunk * 4 bytes
// apart WHATEVER the particle count.  (With rows n * 4 bytes apart the passes slowed down as n grew — the same 1 M particles
// with rows 32 MB apart, as in an 8 M-particle list: lambda 1.25 -> 1.59 ms per step; profiles/r02c_list_layout.txt.)
#ifndef PBF_NL_CHUNK_LOG2
#define PBF_NL_CHUNK_LOG2 15
#endif
constexpr uint32_t kChunkLog2 = PBF_NL_CHUNK_LOG2;
constexpr uint32_t kChunk = 1u << kChunkLog2;
__device__ __forceinline__ uint32_t list_base(uint32_t a, uint32_t chunk_words) {
  return (a >> kChunkLog2) * chunk_words + (a & (kChunk - 1u));
}

// Cache policy of the list's stores / loads (streaming: written once, read once per pass).  The -D switches exist for the
// A/B libraries of profiles/tools/build_variant.sh; .cg, L1::no_allocate and the default policy all measured within 1.5 %.
#ifndef PBF_NL_ST
#define PBF_NL_ST ".cs"
#endif
#ifndef PBF_NL_LDQ
#define PBF_NL_LDQ ".cs"
#endif

// Candidate-position gather of the search loop.  (L2 prefetch-size hints, L1 eviction priorities and coherent loads on this
// gather were all measured in round 2 and changed nothing: profiles/r02b_block_size.txt, section 2.)
__device__ __forceinline__ float4 ldg4s(const float4 *p) { return __ldg(p); }

// list entry load (coherent — never .nc: the lambda kernel reads what it wrote itself)
__device__ __forceinline__ uint32_t ld_list(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global" PBF_NL_LDQ ".b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Predicated 32-bit global store (one @p STG, no divergent branch: with ~1 hit in 5 candidates some lane would take
// a branch for most candidates anyway).  Streaming (.cs): the list is written once and read once per pass; it must not
// push the 16 B/particle pStar array — which every candidate test gathers from — out of L2.
__device__ __forceinline__ void store_if(uint32_t *p, uint32_t v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global" PBF_NL_ST ".b32 [%0], %1;\n\t}" ::"l"(p), "r"(v),
      "r"((uint32_t)pred)
      : "memory");
}

// The append of the search loop: store the candidate's index at the particle's next free slot and advance the slot,
// both under the hit predicate (@p STG + @p IADD: two instructions, where `slot += hit ? stride : 0` costs a select
// and an add on top of the store).
__device__ __forceinline__ void append_if(uint32_t *nl, uint32_t &slot, uint32_t v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.global" PBF_NL_ST ".b32 [%1], %2;\n\t@q add.u32 %0, %0, %4;\n\t}"
      : "+r"(slot)
      : "l"(nl + slot), "r"(v), "r"((uint32_t)pred), "n"(kChunk)
      : "memory");
}

// The sums over one particle's hit list, in list (= the reference's visiting) order: batches of kW (kW list entries,
// then kW position gathers in flight), then ONE masked batch for the 1..kW-1 hits left over — a scalar remainder loop
// waits out two dependent latencies (list entry, then position) per hit.  kW = 8 in the delta pass; 4 in the lambda
// pass, whose search loop needs the registers (56 -> 9 blocks per SM).
template <int kW, typename Acc>
__device__ __forceinline__ void sum_over_hits(Acc &acc, const StepConst &c, const float4 pa, const float4 *__restrict__ pstar,
                                              const uint32_t *row, uint32_t k, uint32_t self) {
  uint32_t i = 0;
  for (; i + kW <= k; i += kW, row += kW * kChunk) {
    uint32_t b[kW];
    float4 q[kW];
#pragma unroll
    for (int m = 0; m < kW; ++m) b[m] = ld_list(row + m * kChunk);
#pragma unroll
    for (int m = 0; m < kW; ++m) q[m] = ldg4(pstar + b[m]);
#pragma unroll
    for (int m = 0; m < kW; ++m) acc.add_in(c, pa, q[m]);
  }
  const uint32_t left = k - i;
  if (left) {
    uint32_t b[kW - 1];
    float4 q[kW - 1];
#pragma unroll
    for (int m = 0; m < kW - 1; ++m)
      b[m] = (uint32_t)m < left ? ld_list(row + m * kChunk) : self;
#pragma unroll
    for (int m = 0; m < kW - 1; ++m) q[m] = ldg4(pstar + b[m]);
#pragma unroll
    for (int m = 0; m < kW - 1; ++m)
      if ((uint32_t)m < left) acc.add_in(c, pa, q[m]);
  }
}

// cell-table bounds of one (y,z) row of the 27-cell neighbourhood: the merged pair run and the single cell
struct RowRuns {
  uint32_t ps, pe, ss, se;
};

__device__ __forceinline__ RowRuns load_row(const uint32_t *__restrict__ table, uint32_t G, uint32_t myz, uint32_t pair_lo,
                                            uint32_t single) {
  RowRuns r;
  const uint32_t op = myz | pair_lo, os = myz | single;
  if (op + 2u < G) {  // both cells of the pair, and the entry after them, exist
    r.ps = __ldg(table + op);
    r.pe = __ldg(table + op + 2);
  } else {            // at the end of the table fall back to the per-cell rule (cell G-1 is empty, sph.hpp:203-213)
    uint32_t s2, e2;
    cell_range(table, G, op, r.ps, r.pe);
    cell_range(table, G, op + 1u, s2, e2);
    if (r.pe == r.ps) { r.ps = s2; r.pe = e2; } else if (e2 != s2) r.pe = e2;  // adjacent when both non-empty
  }
  cell_range(table, G, os, r.ss, r.se);
  return r;
}

template <bool kStrict, int kCap>
__device__ __forceinline__ void lambda_particle(const StepConst &c, const uint32_t a, const uint32_t *__restrict__ keys,
                                                const uint32_t *__restrict__ table, const float4 *__restrict__ pos_mass,
                                                const float4 *__restrict__ pstar_in, float4 *__restrict__ pstar_out,
                                                float *__restrict__ rho_out, uint32_t *nl, uint32_t chunk_words,
                                                uint32_t *__restrict__ n_hits) {
  const float4 pa = ldg4(pstar_in + a);
  const uint32_t key = __ldg(keys + a);
  const float mass = __ldg(&pos_mass[a].w);

  // ---- phase 1: find the hits
  const uint32_t kx = key & kAxisMask, ky = (key >> 1) & kAxisMask, kz = (key >> 2) & kAxisMask;
  const uint32_t xm = dilated_dec(kx), xp = dilated_inc(kx);
  const bool x_even = (kx & 1u) == 0u;        // even: (x, x+1) are consecutive keys; odd: (x-1, x) are
  const uint32_t pair_lo = x_even ? kx : xm;  // first cell of the merged pair
  const uint32_t single = x_even ? xm : xp;
  const uint32_t ym = dilated_dec(ky) << 1, y0 = ky << 1, yp = dilated_inc(ky) << 1;
  const uint32_t zm = dilated_dec(kz) << 2, z0 = kz << 2, zp = dilated_inc(kz) << 2;
  auto row_key = [&](int r) {
    const int iz = r / 3, iy = r - 3 * iz;
    return (iz == 0 ? zm : (iz == 1 ? z0 : zp)) | (iy == 0 ? ym : (iy == 1 ? y0 : yp));
  };
  const uint32_t base = list_base(a, chunk_words);
  uint32_t slot = base;  // element index of the next free entry: hit k lives at nl[base + k * kChunk]
  uint32_t over = 0;  // hits that did not fit
  auto scan_run = [&](uint32_t s, uint32_t e) {
    uint32_t k = (slot - base) >> kChunkLog2;  // hits so far
    if (k + (e - s) <= (uint32_t)kCap) {  // cannot overflow: no per-candidate capacity test, no hit counter
      // Whole batches of four, then ONE masked batch for the 1-3 left over: its gathers go out together (indices
      // clamped into the run), where a scalar remainder loop waits out one gather latency per candidate.
      uint32_t b = s;
      for (; b + 4u <= e; b += 4u) {
        const float4 q0 = ldg4s(pstar_in + b), q1 = ldg4s(pstar_in + b + 1), q2 = ldg4s(pstar_in + b + 2), q3 = ldg4s(pstar_in + b + 3);
        append_if(nl, slot, b, LambdaAcc<kStrict>::test(c, pa, q0));
        append_if(nl, slot, b + 1u, LambdaAcc<kStrict>::test(c, pa, q1));
        append_if(nl, slot, b + 2u, LambdaAcc<kStrict>::test(c, pa, q2));
        append_if(nl, slot, b + 3u, LambdaAcc<kStrict>::test(c, pa, q3));
      }
      if (b < e) {
        const uint32_t last = e - 1u;
        const float4 q0 = ldg4s(pstar_in + b), q1 = ldg4s(pstar_in + min(b + 1u, last)), q2 = ldg4s(pstar_in + min(b + 2u, last));
        append_if(nl, slot, b, LambdaAcc<kStrict>::test(c, pa, q0));
        append_if(nl, slot, b + 1u, b + 1u < e && LambdaAcc<kStrict>::test(c, pa, q1));
        append_if(nl, slot, b + 2u, b + 2u < e && LambdaAcc<kStrict>::test(c, pa, q2));
      }
    } else {
#pragma unroll 1
      for (uint32_t b = s; b < e; ++b) {
        const bool hit = LambdaAcc<kStrict>::test(c, pa, ldg4s(pstar_in + b));
        const bool fits = hit && k < (uint32_t)kCap;
        store_if(nl + slot, b, fits);
        slot += fits ? kChunk : 0u;
        k += fits ? 1u : 0u;
        over += (hit && !fits) ? 1u : 0u;
      }
    }
  };
  RowRuns nxt = load_row(table, c.G, row_key(0), pair_lo, single);
#pragma unroll 1
  for (int r = 0; r < 9; ++r) {
    const RowRuns cur = nxt;
    if (r < 8) nxt = load_row(table, c.G, row_key(r + 1), pair_lo, single);
    // order along x is (x-1, x, x+1): the single cell comes first when x is even, last when x is odd
    scan_run(x_even ? cur.ss : cur.ps, x_even ? cur.se : cur.pe);
    scan_run(x_even ? cur.ps : cur.ss, x_even ? cur.pe : cur.se);
  }
  const uint32_t k = ((slot - base) >> kChunkLog2) + over;
  n_hits[a] = k;

  // ---- phase 2: the sums over the hits
  LambdaAcc<kStrict> acc;
  acc.init();
  acc.set_mass(mass);
  if (k <= (uint32_t)kCap) {
    sum_over_hits<4>(acc, c, pa, pstar_in, nl + base, k, a);
  } else {
    for_each_candidate(key, c.G, table, [&](uint32_t b) { acc.add(c, pa, ldg4(pstar_in + b)); });
  }
  float rho;
  const float lambda = acc.finish(c, mass, rho);
  pstar_out[a] = make_float4(pa.x, pa.y, pa.z, lambda);
  if (rho_out) rho_out[a] = rho;
}

#define PBF_LAMBDA_ARGS                                                                                                  \
  const uint32_t *__restrict__ keys, const uint32_t *__restrict__ table, const float4 *__restrict__ pos_mass,            \
      const float4 *__restrict__ pstar_in, float4 *__restrict__ pstar_out, float *__restrict__ rho_out, uint32_t *nl,    \
      uint32_t chunk_words, uint32_t *__restrict__ n_hits

// One thread per particle, 256 per block.  (With the list rows n * 4 bytes apart, 1 024-thread blocks were 17 % faster from
// ~3 M particles on and were shipped for a few hours; with the chunked list the 256-thread block wins at every size again:
// dam-8m lambda 13.99 ms per step originally, 11.5 with 1 024-thread blocks, 9.75 now.  profiles/r02b_block_size.txt)
template <bool kStrict, int kCap>
__global__ void __launch_bounds__(kBlock, 4) lambda_list_kernel(StepConst c, Sel sel, PBF_LAMBDA_ARGS) {
  uint32_t a;
  if (!sel_particle(sel, blockIdx.x * kBlock + threadIdx.x, a)) return;
  lambda_particle<kStrict, kCap>(c, a, keys, table, pos_mass, pstar_in, pstar_out, rho_out, nl, chunk_words, n_hits);
}

template <bool kStrict, int kCap>
__device__ __forceinline__ void delta_particle(const StepConst &c, const uint32_t a, const uint32_t *__restrict__ keys,
                                               const uint32_t *__restrict__ table, const float4 *__restrict__ pstar_in,
                                               float4 *__restrict__ pstar_out, const uint32_t *__restrict__ nl, uint32_t chunk_words,
                                               const uint32_t *__restrict__ n_hits) {
  const float4 pa = ldg4(pstar_in + a);
  const uint32_t k = __ldg(n_hits + a);
  DeltaAcc<kStrict> acc;
  acc.init();
  if (k <= (uint32_t)kCap) {
    sum_over_hits<8>(acc, c, pa, pstar_in, nl + list_base(a, chunk_words), k, a);  // add_in skips the particle itself (r < EPSILON)
  } else {
    for_each_candidate(__ldg(keys + a), c.G, table, [&](uint32_t b) { acc.add(c, pa, ldg4(pstar_in + b)); });
  }
  pstar_out[a] = acc.finish(c, pa);
}

#define PBF_DELTA_ARGS                                                                                                   \
  const uint32_t *__restrict__ keys, const uint32_t *__restrict__ table, const float4 *__restrict__ pstar_in,            \
      float4 *__restrict__ pstar_out, const uint32_t *__restrict__ nl, uint32_t chunk_words, const uint32_t *__restrict__ n_hits

template <bool kStrict, int kCap>
__global__ void __launch_bounds__(kBlockD, 8) delta_list_kernel(StepConst c, Sel sel, PBF_DELTA_ARGS) {
  uint32_t a;
  if (!sel_particle(sel, blockIdx.x * kBlockD + threadIdx.x, a)) return;
  delta_particle<kStrict, kCap>(c, a, keys, table, pstar_in, pstar_out, nl, chunk_words, n_hits);
}

template <bool kStrict, int kCap> int launch_lambda_mode(pbf_ctx *ctx, const Sel &sel, const uint32_t *keys_sorted,
                                                         const uint32_t *table, const float4 *pos_mass, const float4 *pstar_in,
                                                         float4 *pstar_out, float *rho_out, uint32_t chunk_words) {
  lambda_list_kernel<kStrict, kCap><<<div_up(sel.bound, (uint32_t)kBlock), kBlock, 0, ctx->stream>>>(
      ctx->sc, sel, keys_sorted, table, pos_mass, pstar_in, pstar_out, rho_out, ctx->nl.p, chunk_words, ctx->nl_count.p);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

template <int kCap> int launch_lambda_cap(pbf_ctx *ctx, const Sel &sel, const uint32_t *keys_sorted, const uint32_t *table,
                                          const float4 *pos_mass, const float4 *pstar_in, float4 *pstar_out, float *rho_out,
                                          uint32_t chunk_words) {
  if (ctx->flags & PBF_FLAG_STRICT_FP)
    return launch_lambda_mode<true, kCap>(ctx, sel, keys_sorted, table, pos_mass, pstar_in, pstar_out, rho_out, chunk_words);
  return launch_lambda_mode<false, kCap>(ctx, sel, keys_sorted, table, pos_mass, pstar_in, pstar_out, rho_out, chunk_words);
}

template <bool kStrict, int kCap> int launch_delta_mode(pbf_ctx *ctx, const Sel &sel, const uint32_t *keys_sorted,
                                                        const uint32_t *table, const float4 *pstar_in, float4 *pstar_out) {
  delta_list_kernel<kStrict, kCap><<<div_up(sel.bound, kBlockD), kBlockD, 0, ctx->stream>>>(
      ctx->sc, sel, keys_sorted, table, pstar_in, pstar_out, ctx->nl.p, ctx->nl_stride, ctx->nl_count.p);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

template <int kCap> int launch_delta_cap(pbf_ctx *ctx, const Sel &sel, const uint32_t *keys_sorted, const uint32_t *table,
                                         const float4 *pstar_in, float4 *pstar_out) {
  if (ctx->flags & PBF_FLAG_STRICT_FP) return launch_delta_mode<true, kCap>(ctx, sel, keys_sorted, table, pstar_in, pstar_out);
  return launch_delta_mode<false, kCap>(ctx, sel, keys_sorted, table, pstar_in, pstar_out);
}

}  // namespace

// The list spans the whole sorted array of the step: ctx->sc.n particles (slab path: the capacity of the local array,
// fixed between re-plans, so the list is sized once and never re-allocated in mid-step).
int launch_lambda_list(pbf_ctx *ctx, const Sel &sel, const uint32_t *keys_sorted, const uint32_t *table,
                       const float4 *pos_mass, const float4 *pstar_in, float4 *pstar_out, float *rho_out) {
  if (sel.bound == 0) return PBF_OK;
  const uint32_t n = ctx->sc.n;
  const uint32_t n_chunks = div_up(n, kChunk);
  // Rows cost address space, not bandwidth (a row is touched only by particles with that many hits), so the search keeps
  // up to kListWide hits: while a dam break splashes, particles clamped onto the walls pile up (dam-1m after 30 steps:
  // 431 particles with 97..229 neighbours) and every one that overflows drags its block through the one-pass fallback
  // in both passes.  The wide list needs n < 2^32 / 193 = 22 M particles per device; above that the list is kListMax deep.
  uint32_t cap = (uint32_t)ctx->list_cap;
  if (cap == kListWide && (uint64_t)n_chunks * kChunk * (kListWide + 1) >= (1ull << 32)) cap = kListMax;
  if ((uint64_t)n_chunks * kChunk * (cap + 1) >= (1ull << 32))
    return fail(ctx, PBF_ERR_INVALID, "n", "too many particles on one device (neighbour-list indexing)");
  const uint32_t chunk_words = kChunk * (cap + 1);
  PBF_CUDA(ctx, ctx->nl.reserve((size_t)n_chunks * chunk_words));
  PBF_CUDA(ctx, ctx->nl_count.reserve(n));
  ctx->nl_stride = chunk_words;  // words per chunk of the list as this launch writes it (the delta pass reads it back)
  ctx->nl_cap = cap;
  if (cap == kListWide)
    return launch_lambda_cap<kListWide>(ctx, sel, keys_sorted, table, pos_mass, pstar_in, pstar_out, rho_out, chunk_words);
  return launch_lambda_cap<kListMax>(ctx, sel, keys_sorted, table, pos_mass, pstar_in, pstar_out, rho_out, chunk_words);
}

int launch_delta_list(pbf_ctx *ctx, const Sel &sel, const uint32_t *keys_sorted, const uint32_t *table,
                      const float4 *pstar_in, float4 *pstar_out) {
  if (sel.bound == 0) return PBF_OK;
  if (ctx->nl_cap == kListWide)  // the capacity the lambda pass of this iteration wrote the list with
    return launch_delta_cap<kListWide>(ctx, sel, keys_sorted, table, pstar_in, pstar_out);
  return launch_delta_cap<kListMax>(ctx, sel, keys_sorted, table, pstar_in, pstar_out);
}

}  // namespace pbf
