// cells.cuh — walking the 27-cell neighbourhood through the global cell table, in the reference's order.
#pragma once

#include "common.cuh"

namespace pbf {

// Particle range of cell `o` as the reference reads it (sph.hpp:203-213): cells >= G do not exist and the last
// cell G-1 is always empty (its end is read as its start).
__device__ __forceinline__ void cell_range(const uint32_t *__restrict__ table, uint32_t G, uint32_t o, uint32_t &s,
                                           uint32_t &e) {
  if (o >= G) { s = e = 0; return; }
  s = __ldg(table + o);
  e = (o + 1 < G) ? __ldg(table + o + 1) : s;
}

// Dilated-integer arithmetic on one Morton axis (bits 0,3,6,...,27): +1 / -1 without decoding.  The wrap-around
// is the reference's: 0 - 1 -> 1023 (size_t underflow masked by the spread, sph.hpp:221) and 1023 + 1 -> 0
// (bit 10 falls outside the spread masks, curves.h:73-76).
constexpr uint32_t kAxisMask = 0x09249249u;
__device__ __forceinline__ uint32_t dilated_dec(uint32_t v) { return (v - 1u) & kAxisMask; }
__device__ __forceinline__ uint32_t dilated_inc(uint32_t v) { return ((v | ~kAxisMask) + 1u) & kAxisMask; }

// Calls f(s, e) with the sorted-index range of each of the 27 neighbour cells of `key` in the reference's order
// (sph.hpp:215-236: x fastest, then y, then z, each -1,0,+1).
template <typename F> __device__ __forceinline__ void for_each_cell(uint32_t key, uint32_t G,
                                                                    const uint32_t *__restrict__ table, F &&f) {
  const uint32_t kx = key & kAxisMask, ky = (key >> 1) & kAxisMask, kz = (key >> 2) & kAxisMask;
  const uint32_t xm = dilated_dec(kx), xp = dilated_inc(kx);
  const uint32_t ym = dilated_dec(ky) << 1, y0 = ky << 1, yp = dilated_inc(ky) << 1;
  const uint32_t zm = dilated_dec(kz) << 2, z0 = kz << 2, zp = dilated_inc(kz) << 2;
#pragma unroll 1
  for (int iz = 0; iz < 3; ++iz) {
    const uint32_t mz = iz == 0 ? zm : (iz == 1 ? z0 : zp);
#pragma unroll 1
    for (int iy = 0; iy < 3; ++iy) {
      const uint32_t myz = mz | (iy == 0 ? ym : (iy == 1 ? y0 : yp));
#pragma unroll
      for (int ix = 0; ix < 3; ++ix) {
        uint32_t s, e;
        cell_range(table, G, myz | (ix == 0 ? xm : (ix == 1 ? kx : xp)), s, e);
        f(s, e);
      }
    }
  }
}

// Same walk with contiguous cells merged: in Morton order an even-x cell and its +x neighbour have consecutive keys,
// so two of the three x-cells of every (y,z) row form ONE contiguous particle range.  27 cell ranges become 18
// runs (a third fewer table look-ups and loop set-ups); the particles are still visited in the reference's order.
template <typename F> __device__ __forceinline__ void for_each_run(uint32_t key, uint32_t G,
                                                                   const uint32_t *__restrict__ table, F &&f) {
  const uint32_t kx = key & kAxisMask, ky = (key >> 1) & kAxisMask, kz = (key >> 2) & kAxisMask;
  const uint32_t xm = dilated_dec(kx), xp = dilated_inc(kx);
  const bool x_even = (kx & 1u) == 0u;          // even: (x, x+1) are consecutive keys; odd: (x-1, x) are
  const uint32_t pair_lo = x_even ? kx : xm;    // first cell of the merged pair
  const uint32_t single = x_even ? xm : xp;
  const uint32_t ym = dilated_dec(ky) << 1, y0 = ky << 1, yp = dilated_inc(ky) << 1;
  const uint32_t zm = dilated_dec(kz) << 2, z0 = kz << 2, zp = dilated_inc(kz) << 2;
#pragma unroll 1
  for (int iz = 0; iz < 3; ++iz) {
    const uint32_t mz = iz == 0 ? zm : (iz == 1 ? z0 : zp);
#pragma unroll 1
    for (int iy = 0; iy < 3; ++iy) {
      const uint32_t myz = mz | (iy == 0 ? ym : (iy == 1 ? y0 : yp));
      const uint32_t op = myz | pair_lo, os = myz | single;
      uint32_t ps, pe, ss, se;
      if (op + 2u < G) {  // both cells of the pair, and the entry after them, exist
        ps = __ldg(table + op);
        pe = __ldg(table + op + 2);
      } else {            // at the end of the table fall back to the per-cell rule (cell G-1 is empty)
        uint32_t s2, e2;
        cell_range(table, G, op, ps, pe);
        cell_range(table, G, op + 1u, s2, e2);
        if (pe == ps) { ps = s2; pe = e2; } else if (e2 != s2) pe = e2;  // the two ranges are adjacent when non-empty
      }
      cell_range(table, G, os, ss, se);
      // order along x is (x-1, x, x+1): the single cell comes first when x is even, last when x is odd
#pragma unroll 1
      for (int r = 0; r < 2; ++r) {
        const bool take_single = (r == 0) == x_even;
        f(take_single ? ss : ps, take_single ? se : pe);
      }
    }
  }
}

// Calls f(b) for every candidate b (ascending sorted index inside each cell).
template <typename F> __device__ __forceinline__ void for_each_candidate(uint32_t key, uint32_t G,
                                                                         const uint32_t *__restrict__ table, F &&f) {
  for_each_cell(key, G, table, [&](uint32_t s, uint32_t e) {
    for (uint32_t b = s; b < e; ++b) f(b);
  });
}

}  // namespace pbf
