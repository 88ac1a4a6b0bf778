// cell_search.cu — the neighbour search of one solver iteration with one WARP PER OCCUPIED CELL (PBF_SEARCH=cells).
//
// An alternative to the production search (neighbour_list.cu, lambda_list_kernel phase 1: one thread per particle walks
// the cell table; ncu: 17.8 of 32 lanes active, because the lanes of a warp sit in ~5 different cells whose 18 runs all
// differ in length).  All particles of one cell share the same 27 neighbour cells (sph.hpp:215-236), so here
//
//   once per step (keys and cell table are fixed for all iterations, ompsph.hpp:215-249) — cell_plan_kernel:
//   * a warp takes the cells whose first particle lies in its 32-particle window of the Z-sorted array;
//   * lanes 0..26 look up the 27 cell ranges (one table access each), a warp scan flattens them into ONE candidate
//     sequence in the reference's visiting order (x fastest, then y, then z; ascending index in a cell), and the
//     sequence is written out as the cell's candidate list;
//
//   once per iteration — search_cells_kernel:
//   * the cell's candidates are loaded ONCE, 32 at a time, coalesced (lane j holds candidate j), up to 8 chunks
//     (256 candidates) in registers, two chunks per 64-bit register pair; the next cell's list is already in flight;
//   * the cell's particles ("targets") are then taken one by one: every lane tests ITS candidates against the
//     target — two candidates per instruction with Blackwell's packed FADD2/FMUL2/FFMA2 — a ballot gives the hit
//     mask of the chunk, and the hit lanes append their candidate's index to the target's list at
//     count + popc(mask below me): the list is in visiting order, exactly the list the thread-per-particle search
//     writes, so the sums formed from it are bit-identical (tests/test_parity_gpu.py).
//
// No lane idles on another lane's run (31.5 of 32 active) and no candidate is loaded more than once per cell.  Hits are
// staged in shared memory, eight targets at a time, and written out as 32-byte row segments into the list layout
// lambda_sums / delta_list read (nl[k * stride + particle], n_hits[particle]); a particle with more than kCap hits is
// flagged by n_hits > kCap and its sums take the one-pass walk.
//
// MEASURED SLOWER than the production search (dam-1m: 320 us + 79 us for the sums against 340 us fused; lambda 1.50
// against 1.32 ms/step): the ballot/popc/append bookkeeping costs as many instructions per pair as the divergence it
// removes, and the per-cell load chain runs at 17-24 warps/SM.  Kept as an A/B option; profiles/r01c_search_experiments.txt
// has the ncu figures of every stage of this kernel.
#include "cells.cuh"
#include "common.cuh"
#include "pair_math.cuh"

namespace pbf {

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kChunks = 8;  // candidate chunks (of 32) held in registers at a time
constexpr unsigned kFull = 0xffffffffu;
constexpr float kFar = 1e18f;  // coordinate of a padding candidate: never within h of anything, squares stay finite

constexpr int kGroup = 8;      // targets whose hit lists are staged in shared memory at a time
constexpr int kPitch = 100;    // words per staged list: kCap + dump slot, and pitch = 4 (mod 32) makes the flush conflict-free

// Up to kGroup targets [g0, g0 + gn) of one cell against NCH resident candidate chunks.  X/Y/Z[m] hold chunks 2m (low
// half) and 2m+1 (high half).  Hits go to the group's staged lists s_list[g][*] (slot kCap is the dump slot); lane g
// keeps target g's running hit count in `mycnt`.  `more` = these are not the cell's first candidates.
template <bool kStrict, int kCap, int NCH>
__device__ __forceinline__ void scan_group(const StepConst &c, const float4 *__restrict__ pstar, uint32_t g0, uint32_t gn,
                                           bool more, const f2 (&X)[kChunks / 2], const f2 (&Y)[kChunks / 2],
                                           const f2 (&Z)[kChunks / 2], const uint32_t (&ci)[kChunks], uint32_t *s_list,
                                           uint32_t &mycnt, uint32_t lane, uint32_t lt) {
  constexpr int NP = (NCH + 1) / 2;
  float4 pa_next = ldg4(pstar + g0);
  for (uint32_t g = 0; g < gn; ++g) {
    const float4 pa = pa_next;
    if (g + 1 < gn) pa_next = ldg4(pstar + g0 + g + 1);
    uint32_t cnt = more ? __shfl_sync(kFull, mycnt, g) : 0u;
    const f2 AX = pack2(pa.x, pa.x), AY = pack2(pa.y, pa.y), AZ = pack2(pa.z, pa.z);
    float r2[2 * NP];
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      const f2 dx = sub2(X[m], AX), dy = sub2(Y[m], AY), dz = sub2(Z[m], AZ);
      // strict: (dx*dx + dy*dy) + dz*dz, no contraction — glm::distance's dot product (pair_math.cuh)
      const f2 r = kStrict ? add2(add2(mul2(dx, dx), mul2(dy, dy)), mul2(dz, dz)) : fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
      unpack2(r, r2[2 * m], r2[2 * m + 1]);
    }
    uint32_t *list = s_list + g * kPitch;
#pragma unroll
    for (int k = 0; k < 2 * NP; ++k) {
      const bool hit = r2[k] <= c.r2_max;
      const uint32_t m = __ballot_sync(kFull, hit);
      if (hit) list[min(cnt + __popc(m & lt), (uint32_t)kCap)] = ci[k];
      cnt += __popc(m);
    }
    if (lane == g) mycnt = cnt;
  }
}

// ---- once per step: the plan --------------------------------------------------------------------------------------
// Keys and the cell table are fixed for all iterations of a step (ompsph.hpp:215-249), so which cells exist, which
// particles they hold and which candidates they see is worked out ONCE per step:
//   head_mask[w]      bit l set: particle 32 w + l (+ first) is the first particle of a cell to be searched
//   cell_info[t0]     {offset of the cell's candidate list in cand[], number of candidates T, number of targets, 0}
//   cand[off + j]     sorted index of flat candidate j: the 27 neighbour cells in the reference's visiting order
//                     (sph.hpp:215-236), ascending index inside a cell.  Lists start on 128-byte boundaries.
__global__ void __launch_bounds__(kWarpsPerBlock * 32) cell_plan_kernel(
    uint32_t G, uint32_t first, uint32_t count, const uint32_t *__restrict__ keys, const uint32_t *__restrict__ table,
    const uint32_t *__restrict__ role, uint32_t want, uint32_t *__restrict__ head_mask, uint4 *__restrict__ cell_info,
    uint32_t *__restrict__ cand, uint32_t *cursor) {
  __shared__ uint32_t s_delta[kWarpsPerBlock][32];  // per non-empty neighbour cell, in visiting order: first particle - prefix
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u, le = lt | (1u << lane);
  const uint32_t end = first + count;
  const uint32_t w = blockIdx.x * kWarpsPerBlock + wib;
  const uint32_t base = first + w * 32u;
  if (base >= end) return;  // whole warp

  // ---- which cells start in this window
  const uint32_t a = base + lane;
  const bool in = a < end;
  const uint32_t key_a = in ? __ldg(keys + a) : 0xffffffffu;
  const bool change = in && (a == first || __ldg(keys + a - 1) != key_a);
  const bool wanted = !role || (in && (__ldg(role + a) & want) != 0u);
  const uint32_t changes = __ballot_sync(kFull, change);
  uint32_t heads = __ballot_sync(kFull, change && wanted);
  if (lane == 0) head_mask[w] = heads;

  // neighbour offset of lane l < 27 in the reference's visiting order
  const uint32_t ox = lane % 3u, oy = (lane / 3u) % 3u, oz = lane / 9u;

  while (heads) {
    const uint32_t hb = __ffs(heads) - 1u;
    heads &= heads - 1u;
    const uint32_t t0 = base + hb;  // first target
    const uint32_t ckey = __shfl_sync(kFull, key_a, hb);
    // last target + 1: the next key change — inside the window, or further down the sorted array
    uint32_t t1;
    const uint32_t later = changes & ~((2u << hb) - 1u);
    if (later) {
      t1 = base + __ffs(later) - 1u;
    } else {
      t1 = min(base + 32u, end);
      while (t1 < end) {
        const uint32_t i = t1 + lane;
        const uint32_t m = __ballot_sync(kFull, i >= end || __ldg(keys + i) != ckey);
        if (m) { t1 += __ffs(m) - 1u; break; }
        t1 += 32u;
      }
      t1 = min(t1, end);
    }

    // ---- the 27 cell ranges, flattened: cell r covers flat positions [pre, pre + len)
    uint32_t len = 0, s = 0;
    if (lane < 27u) {
      const uint32_t kx = ckey & kAxisMask, ky = (ckey >> 1) & kAxisMask, kz = (ckey >> 2) & kAxisMask;
      const uint32_t nx = ox == 0u ? dilated_dec(kx) : (ox == 1u ? kx : dilated_inc(kx));
      const uint32_t ny = oy == 0u ? dilated_dec(ky) : (oy == 1u ? ky : dilated_inc(ky));
      const uint32_t nz = oz == 0u ? dilated_dec(kz) : (oz == 1u ? kz : dilated_inc(kz));
      uint32_t e;
      cell_range(table, G, nx | (ny << 1) | (nz << 2), s, e);
      len = e - s;
    }
    uint32_t incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, incl, d);
      if ((int)lane >= d) incl += v;
    }
    const uint32_t T = __shfl_sync(kFull, incl, 31);
    const uint32_t pre = incl - len;
    const uint32_t nonempty = __ballot_sync(kFull, len != 0u);
    uint32_t off = 0;
    if (lane == 0) {
      off = atomicAdd(cursor, (T + 31u) & ~31u);
      cell_info[t0] = make_uint4(off, T, t1 - t0, 0u);
    }
    off = __shfl_sync(kFull, off, 0);
    __syncwarp();
    if (len) s_delta[wib][__popc(nonempty & lt)] = s - pre;
    __syncwarp();
    for (uint32_t Jk = 0; Jk < T; Jk += 32u) {
      // which non-empty cell holds flat position j: the cells starting inside this chunk mark their first position
      // in a 32-bit mask; r0 = non-empty cells that start before the chunk
      const uint32_t j = Jk + lane, rel = pre - Jk;
      const uint32_t starts = __reduce_or_sync(kFull, (len != 0u && rel < 32u) ? (1u << rel) : 0u);
      const uint32_t r0 = __popc(__ballot_sync(kFull, len != 0u && pre < Jk));
      if (j < T) cand[off + j] = j + s_delta[wib][r0 + __popc(starts & le) - 1u];
    }
  }
}

// ---- once per iteration: the search ---------------------------------------------------------------------------------
template <bool kStrict, int kCap>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) search_cells_kernel(
    StepConst c, uint32_t first, uint32_t count, const uint32_t *__restrict__ head_mask,
    const uint4 *__restrict__ cell_info, const uint32_t *__restrict__ cand, const float4 *__restrict__ pstar,
    uint32_t *__restrict__ nl, uint32_t stride, uint32_t *__restrict__ n_hits) {
  static_assert(kCap < kPitch, "staged list pitch");
  __shared__ uint32_t s_lists[kWarpsPerBlock][kGroup * kPitch];
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t w = blockIdx.x * kWarpsPerBlock + wib;
  const uint32_t base = first + w * 32u;
  if (base >= first + count) return;  // whole warp
  uint32_t heads = __ldg(head_mask + w);
  if (!heads) return;
  uint32_t *s_list = s_lists[wib];
  f2 zero2;
  asm volatile("mov.b64 %0, 0;" : "=l"(zero2));
  // every head lane fetches its cell's record now: one exposed latency per window instead of one per cell
  uint4 my_info = make_uint4(0u, 0u, 0u, 0u);
  if ((heads >> lane) & 1u) my_info = __ldg(cell_info + base + lane);

  auto load_indices = [&](uint32_t off, uint32_t T, uint32_t J0, uint32_t (&ci)[kChunks]) {
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      const uint32_t j = J0 + 32u * k + lane;
      ci[k] = (j < T) ? __ldg(cand + off + j) : 0xffffffffu;
    }
  };
  // the candidates' positions, two chunks per 64-bit register pair
  auto load_positions = [&](const uint32_t (&ci)[kChunks], f2 (&X)[kChunks / 2], f2 (&Y)[kChunks / 2], f2 (&Z)[kChunks / 2]) {
    float cx[kChunks], cy[kChunks], cz[kChunks];
#pragma unroll
    for (int k = 0; k < kChunks; ++k) {
      cx[k] = cy[k] = cz[k] = kFar;
      if (ci[k] != 0xffffffffu) {
        const float4 p = ldg4(pstar + ci[k]);
        cx[k] = p.x; cy[k] = p.y; cz[k] = p.z;
      }
    }
#pragma unroll
    for (int m = 0; m < kChunks / 2; ++m) {
      // "+ 0" (exact): one FADD2 per pair per cell lands the two chunks in an aligned 64-bit register pair; without
      // it ptxas keeps the halves where the 128-bit loads put them and re-assembles every pair for every target
      X[m] = add2(pack2(cx[2 * m], cx[2 * m + 1]), zero2);
      Y[m] = add2(pack2(cy[2 * m], cy[2 * m + 1]), zero2);
      Z[m] = add2(pack2(cz[2 * m], cz[2 * m + 1]), zero2);
    }
  };
  // straight-line code per (even) chunk count; an odd count tests one chunk of padding
  auto scan = [&](uint32_t g0, uint32_t gn, uint32_t T, uint32_t J0, const f2 (&X)[kChunks / 2], const f2 (&Y)[kChunks / 2],
                  const f2 (&Z)[kChunks / 2], const uint32_t (&ci)[kChunks], uint32_t &mycnt) {
#define PBF_SCAN(N) scan_group<kStrict, kCap, N>(c, pstar, g0, gn, J0 != 0u, X, Y, Z, ci, s_list, mycnt, lane, lt)
    switch ((min((T - J0 + 31u) >> 5, (uint32_t)kChunks) + 1u) >> 1) {
      case 1: PBF_SCAN(2); break;
      case 2: PBF_SCAN(4); break;
      case 3: PBF_SCAN(6); break;
      default: PBF_SCAN(8); break;
    }
#undef PBF_SCAN
  };
  // staged lists -> nl[k * stride + particle]: lane = (row within a block of four, target): each store covers four
  // rows of eight neighbouring particles
  auto flush = [&](uint32_t g0, uint32_t gn, uint32_t mycnt) {
    const uint32_t g = lane & (kGroup - 1), rr = lane / kGroup;
    const uint32_t cnt = __shfl_sync(kFull, mycnt, g);
    if (lane < gn) n_hits[g0 + lane] = mycnt;
    const uint32_t kept = g < gn ? min(cnt, (uint32_t)kCap) : 0u;
    uint32_t most = kept;
#pragma unroll
    for (int d = 1; d < kGroup; d <<= 1) most = max(most, __shfl_xor_sync(kFull, most, d));
    __syncwarp();
    uint32_t *out = nl + g0 + g;
    for (uint32_t r = rr; r < most; r += 32 / kGroup)
      if (r < kept) __stcs(out + (size_t)r * stride, s_list[g * kPitch + r]);
    __syncwarp();
  };

  // software pipeline over the window's cells: while cell i is searched, the candidate indices of cell i+1 are in
  // flight, so only the position gathers of a cell wait on memory
  uint32_t hb = __ffs(heads) - 1u;
  uint32_t off = __shfl_sync(kFull, my_info.x, hb), T = __shfl_sync(kFull, my_info.y, hb);
  uint32_t nt = __shfl_sync(kFull, my_info.z, hb);
  uint32_t ci_next[kChunks];
  load_indices(off, T, 0u, ci_next);
  for (;;) {
    const uint32_t t0 = base + hb, t1 = t0 + nt, cur_off = off, cur_T = T;
    uint32_t ci[kChunks];
#pragma unroll
    for (int k = 0; k < kChunks; ++k) ci[k] = ci_next[k];
    heads &= heads - 1u;
    if (heads) {
      hb = __ffs(heads) - 1u;
      off = __shfl_sync(kFull, my_info.x, hb);
      T = __shfl_sync(kFull, my_info.y, hb);
      nt = __shfl_sync(kFull, my_info.z, hb);
      load_indices(off, T, 0u, ci_next);
    }
    f2 X[kChunks / 2], Y[kChunks / 2], Z[kChunks / 2];
    if (cur_T == 0u) {
      for (uint32_t t = t0 + lane; t < t1; t += 32u) n_hits[t] = 0;
    } else if (cur_T <= 32u * kChunks) {  // the usual case: the cell's candidates stay resident for all its targets
      load_positions(ci, X, Y, Z);
      for (uint32_t g0 = t0; g0 < t1; g0 += kGroup) {
        const uint32_t gn = min((uint32_t)kGroup, t1 - g0);
        uint32_t mycnt = 0;
        scan(g0, gn, cur_T, 0u, X, Y, Z, ci, mycnt);
        flush(g0, gn, mycnt);
      }
    } else {
      for (uint32_t g0 = t0; g0 < t1; g0 += kGroup) {
        const uint32_t gn = min((uint32_t)kGroup, t1 - g0);
        uint32_t mycnt = 0;
        for (uint32_t J0 = 0; J0 < cur_T; J0 += 32u * kChunks) {
          load_indices(cur_off, cur_T, J0, ci);
          load_positions(ci, X, Y, Z);
          scan(g0, gn, cur_T, J0, X, Y, Z, ci, mycnt);
        }
        flush(g0, gn, mycnt);
      }
    }
    if (!heads) break;
  }
}

}  // namespace

// Builds the per-step plan when the cell table has changed since the last call (ctx->plan_valid is cleared by
// launch_cell_table) or the range / role selection differs.
int ensure_search_plan(pbf_ctx *ctx, uint32_t first, uint32_t count, const uint32_t *keys_sorted, const uint32_t *table,
                       const uint32_t *role, uint32_t want) {
  if (ctx->plan_valid && ctx->plan_first == first && ctx->plan_count == count && ctx->plan_role == role &&
      ctx->plan_want == want)
    return PBF_OK;
  const unsigned windows = div_up(count, 32);
  const unsigned blocks = div_up(windows, kWarpsPerBlock);
  const uint32_t n = ctx->sc.n;
  PBF_CUDA(ctx, ctx->plan_heads.reserve((size_t)windows + 1));
  PBF_CUDA(ctx, ctx->plan_info.reserve((size_t)first + count));
  // sum of T over cells <= 27 n (a particle is a candidate of at most 27 cells), + < 32 of padding per cell
  PBF_CUDA(ctx, ctx->plan_cand.reserve((size_t)27 * n + (size_t)32 * count + 32));
  uint32_t *cursor = ctx->plan_heads.p + windows;
  PBF_CUDA(ctx, cudaMemsetAsync(cursor, 0, sizeof(uint32_t), ctx->stream));
  cell_plan_kernel<<<blocks, kWarpsPerBlock * 32, 0, ctx->stream>>>(ctx->sc.G, first, count, keys_sorted, table, role, want,
                                                                    ctx->plan_heads.p, ctx->plan_info.p, ctx->plan_cand.p,
                                                                    cursor);
  PBF_LAUNCH_CHECK(ctx);
  ctx->plan_valid = true;
  ctx->plan_first = first; ctx->plan_count = count; ctx->plan_role = role; ctx->plan_want = want;
  return PBF_OK;
}

// Writes nl / n_hits for every particle of [first, first + count) whose role matches (see launch_lambda_list).
// nl must hold kCap + 1 rows of `stride` entries.
int launch_search_cells(pbf_ctx *ctx, uint32_t first, uint32_t count, const uint32_t *keys_sorted, const uint32_t *table,
                        const float4 *pstar_in, uint32_t stride, const uint32_t *role, uint32_t want) {
  if (count == 0) return PBF_OK;
  PBF_TRY(ensure_search_plan(ctx, first, count, keys_sorted, table, role, want));
  const unsigned blocks = div_up(div_up(count, 32), kWarpsPerBlock);
  const bool strict = (ctx->flags & PBF_FLAG_STRICT_FP) != 0;
#define PBF_SEARCH(S, CAP)                                                                                          \
  search_cells_kernel<S, CAP><<<blocks, kWarpsPerBlock * 32, 0, ctx->stream>>>(                                     \
      ctx->sc, first, count, ctx->plan_heads.p, ctx->plan_info.p, ctx->plan_cand.p, pstar_in, ctx->nl.p, stride, \
      ctx->nl_count.p)
  if (ctx->list_cap == 64) {
    if (strict) PBF_SEARCH(true, 64); else PBF_SEARCH(false, 64);
  } else {
    if (strict) PBF_SEARCH(true, (int)kListMax); else PBF_SEARCH(false, (int)kListMax);
  }
#undef PBF_SEARCH
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

}  // namespace pbf
