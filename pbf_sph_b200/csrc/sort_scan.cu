// sort_scan.cu — device-wide exclusive scan and the stable LSD radix sort of (Morton key, index) pairs.
//
// Replaces the reference's host-side std::sort of 96-byte AoS records by zIndex (ompsph.hpp:158; unstable
// there, stable here, which is what the parity oracle defines — SURVEY.md F3).  Keys are 30-bit Morton
// codes (curves.h:72-88); a particle predicted outside the padded grid can carry a key >= G, and the
// reference sorts on the full key, so all 30 bits are sorted: three passes of 10-bit digits, regardless of G.
//
// Per pass:   digit histogram per 4096-key tile (+ the global digit totals, by atomics)  ->  exclusive scan of every
//             digit's row of the [digit][tile] matrix, one warp per row  ->  stable scatter, which scans the 1024
//             digit totals itself (a generic three-kernel scan of the whole 250 K-entry matrix took 18 us of the
//             41 us pass at 1 M keys; the row scan takes a third of that and a pass is 3 launches instead of 5).  Ranks inside a tile come from __match_any_sync warp multisplit plus per-warp
//             digit counters in shared memory, so equal digits keep their input order (stability) without
//             atomics on the ordering path.
#include "common.cuh"

namespace pbf {

namespace {

// ======================================= exclusive scan =======================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if (lane >= (unsigned)d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across a 256-thread block; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total) {
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  __shared__ uint32_t block_total;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t incl = warp_incl_scan(v);
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = lane < kScanThreads / 32 ? warp_sums[lane] : 0u;
    const uint32_t si = warp_incl_scan(s);
    if (lane < kScanThreads / 32) warp_sums[lane] = si - s;
    if (lane == kScanThreads / 32 - 1) block_total = si;
  }
  __syncthreads();
  *total = block_total;
  return warp_sums[warp] + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                      uint32_t *__restrict__ sums) {
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t i = base + (uint64_t)k * kScanThreads + threadIdx.x;
    if (i < n) s += __ldg(in + i);
  }
  uint32_t total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// one tile: thread t owns items [t*16, t*16+16) so the scan order is the memory order
// `out` may alias `in` (every thread reads its items before the block-wide barrier and writes them after it)
__global__ void __launch_bounds__(kScanThreads) scan_tile_apply_kernel(const uint32_t *in, uint64_t n,
                                                                       const uint32_t *tile_offsets, uint32_t *out,
                                                                       uint32_t *total_out) {
  const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    s += v[k];
  }
  uint32_t total;
  uint32_t run = block_excl_scan(s, &total) + (tile_offsets ? tile_offsets[blockIdx.x] : 0u);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *total_out = run;
}

// Up to 128 K entries in ONE launch by ONE block of 1 024 threads.  Warp w owns the contiguous segment w of 32; it walks
// the segment 128 entries at a time (four coalesced loads in flight, a shuffle scan each, a running carry) and writes the
// segment-local prefix; the block scans the 32 segment totals; every warp but the first adds its base in a second coalesced
// sweep.  The slab path scans a few thousand to a few ten thousand per-tile counters several times per step between
// dependent kernels: 4 x 60 us per rank-step at 8 ranks with the thread-per-slice form this replaces (strided loads), and
// 18 us for 16 K entries with the serial tile walk before that.  `out` may alias `in`.
constexpr int kScanWide = 1024;
constexpr uint64_t kScanSmall = 128ull * 1024;
// Launched with several blocks it scans ROWS: block r scans entries [r * n, (r + 1) * n) on their own and writes the row's
// total to total_out[r] (exclusive_scan_rows_u32: the slab path's per-destination tile counts).
__global__ void __launch_bounds__(kScanWide) scan_one_block_kernel(const uint32_t *in, uint32_t n, uint32_t *out,
                                                                   uint32_t *total_out) {
  __shared__ uint32_t wbase[kScanWide / 32];
  in += (size_t)blockIdx.x * n;
  out += (size_t)blockIdx.x * n;
  if (total_out) total_out += blockIdx.x;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t seg = (((n + 31u) / 32u) + 127u) & ~127u;  // entries per warp, a multiple of the 128-entry stride
  const uint32_t lo = min(n, warp * seg), hi = min(n, lo + seg);
  uint32_t carry = 0;
  uint32_t v[4], nx[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint32_t j = lo + (uint32_t)u * 32u + lane;
    v[u] = j < hi ? in[j] : 0u;
  }
  for (uint32_t i = lo; i < hi; i += 128u) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // the next 128 entries are in flight while these are scanned (in and out may alias: the
      const uint32_t j = i + 128u + (uint32_t)u * 32u + lane;  // compiler would not hoist these loads above the stores)
      nx[u] = j < hi ? in[j] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t j = i + (uint32_t)u * 32u + lane;
      const uint32_t incl = warp_incl_scan(v[u]);
      if (j < hi) out[j] = carry + incl - v[u];
      carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
      v[u] = nx[u];
    }
  }
  if (lane == 0) wbase[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = wbase[lane];
    const uint32_t wi = warp_incl_scan(w);
    wbase[lane] = wi - w;
    if (lane == 31 && total_out) *total_out = wi;
  }
  __syncthreads();
  const uint32_t base = wbase[warp];
  if (base)
    for (uint32_t j0 = lo + lane; j0 < hi; j0 += 256u) {  // eight independent read-modify-writes at a time
      uint32_t t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = j0 + (uint32_t)u * 32u < hi ? out[j0 + (uint32_t)u * 32u] : 0u;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (j0 + (uint32_t)u * 32u < hi) out[j0 + (uint32_t)u * 32u] = t[u] + base;
    }
}

// ======================================= radix sort ============================================================
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per block
constexpr int kDigitBits = 10;
constexpr int kBins = 1 << kDigitBits;  // 1024
constexpr int kPasses = 3;              // 30 bits

__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n,
                                                                 const uint32_t *__restrict__ n_dev, int shift,
                                                                 uint32_t n_tiles, uint32_t *__restrict__ tile_hist,
                                                                 uint32_t *__restrict__ digit_total) {
  __shared__ uint32_t hist[kBins];
  if (n_dev) n = __ldg(n_dev);  // slab path: the pair count lives in device memory; tiles beyond it hold nothing
  for (int b = threadIdx.x; b < kBins; b += kSortThreads) hist[b] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const uint32_t i = base + k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&hist[(__ldg(keys + i) >> shift) & (kBins - 1)], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kBins; b += kSortThreads) {
    const uint32_t h = hist[b];
    tile_hist[(uint32_t)b * n_tiles + blockIdx.x] = h;
    if (h) atomicAdd(digit_total + b, h);
  }
}

// In-place exclusive scan of every digit's row of the [digit][tile] histogram: one warp per row.
__global__ void __launch_bounds__(kSortThreads) sort_row_scan_kernel(uint32_t *__restrict__ tile_hist, uint32_t n_tiles) {
  const unsigned lane = threadIdx.x & 31;
  uint32_t *row = tile_hist + (size_t)(blockIdx.x * kSortWarps + (threadIdx.x >> 5)) * n_tiles;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n_tiles; base += 32u) {
    const uint32_t i = base + lane;
    const uint32_t v = i < n_tiles ? row[i] : 0u;
    const uint32_t incl = warp_incl_scan(v);
    if (i < n_tiles) row[i] = carry + incl - v;
    carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
  }
}

template <bool kIotaValues>
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                    const uint32_t *__restrict__ vals_in, uint32_t n,
                                                                    const uint32_t *__restrict__ n_dev, int shift,
                                                                    uint32_t n_tiles,
                                                                    const uint32_t *__restrict__ tile_offsets,
                                                                    const uint32_t *__restrict__ digit_total,
                                                                    uint32_t *__restrict__ keys_out,
                                                                    uint32_t *__restrict__ vals_out) {
  __shared__ uint32_t wc[kSortWarps][kBins];  // per-warp digit counters, then per-warp global bases (32 KB)
  if (n_dev) {
    n = __ldg(n_dev);
    if (blockIdx.x * kSortTile >= n) return;  // an empty tile of the capacity-sized grid (uniform for the block)
  }
  for (int b = threadIdx.x; b < kSortWarps * kBins; b += kSortThreads) (&wc[0][0])[b] = 0;
  __syncthreads();

  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint32_t warp_base = blockIdx.x * kSortTile + warp * (32 * kSortItems);
  uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];

  // warp w owns the contiguous range [warp_base, warp_base + 512); item k of lane l is element k*32 + l,
  // so (k, lane) lexicographic order == memory order == the order stability must preserve
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const uint32_t i = warp_base + k * 32 + lane;
    const bool valid = i < n;
    key[k] = valid ? __ldg(keys_in + i) : 0u;
    val[k] = kIotaValues ? i : (valid ? __ldg(vals_in + i) : 0u);
    const uint32_t d = valid ? ((key[k] >> shift) & (kBins - 1)) : 0xFFFFFFFFu;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
    const unsigned below = __popc(peers & lt_mask);
    uint32_t prev = 0;
    if (valid) prev = wc[warp][d];
    __syncwarp();
    if (valid && below == 0) wc[warp][d] = prev + __popc(peers);
    __syncwarp();
    rank[k] = prev + below;
  }
  // digit-major bases: keys with lower digits (exclusive scan of the 1024 digit totals: thread t scans digits
  // 4t..4t+3) + this digit's keys in lower tiles (the row scan) + the counts of the lower warps
  static_assert(kBins == 4 * kSortThreads, "four digits per thread");
  __shared__ uint32_t dbase[kBins];
  {
    uint32_t tot[4], mine = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { tot[j] = __ldg(digit_total + 4 * threadIdx.x + j); mine += tot[j]; }
    uint32_t all;
    uint32_t run = block_excl_scan(mine, &all);  // (its barriers also order the per-warp counters written above)
#pragma unroll
    for (int j = 0; j < 4; ++j) { dbase[4 * threadIdx.x + j] = run; run += tot[j]; }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kBins; d += kSortThreads) {
    uint32_t run = dbase[d] + __ldg(tile_offsets + (uint32_t)d * n_tiles + blockIdx.x);
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = wc[w][d];
      wc[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const uint32_t i = warp_base + k * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[k] >> shift) & (kBins - 1);
      const uint32_t dst = wc[warp][d] + rank[k];
      keys_out[dst] = key[k];
      vals_out[dst] = val[k];
    }
  }
}

}  // namespace

// Exclusive prefix sum of n u32 values (n up to 4096^3).  `out` may alias `in`.  If total_out_dev is non-null the
// grand total is written there (device memory).
int exclusive_scan_u32(pbf_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *total_out_dev) {
  if (n == 0) {
    if (total_out_dev) PBF_CUDA(ctx, cudaMemsetAsync(total_out_dev, 0, sizeof(uint32_t), ctx->stream));
    return PBF_OK;
  }
  const uint64_t t1 = (n + kScanTile - 1) / kScanTile;
  if (t1 == 1) {
    scan_tile_apply_kernel<<<1, kScanThreads, 0, ctx->stream>>>(in, n, nullptr, out, total_out_dev);
    PBF_LAUNCH_CHECK(ctx);
    return PBF_OK;
  }
  if (n <= kScanSmall) {
    scan_one_block_kernel<<<1, kScanWide, 0, ctx->stream>>>(in, (uint32_t)n, out, total_out_dev);
    PBF_LAUNCH_CHECK(ctx);
    return PBF_OK;
  }
  const uint64_t t2 = (t1 + kScanTile - 1) / kScanTile;
  const uint64_t t3 = (t2 + kScanTile - 1) / kScanTile;
  if (t3 > 1) return fail(ctx, PBF_ERR_INVALID, "exclusive_scan_u32", "input too large");
  PBF_CUDA(ctx, ctx->scan_tmp.reserve(t1 + t2 + 8));
  uint32_t *s1 = ctx->scan_tmp.p, *s2 = ctx->scan_tmp.p + t1;
  scan_tile_sums_kernel<<<(unsigned)t1, kScanThreads, 0, ctx->stream>>>(in, n, s1);
  PBF_LAUNCH_CHECK(ctx);
  if (t2 == 1) {
    scan_tile_apply_kernel<<<1, kScanThreads, 0, ctx->stream>>>(s1, t1, nullptr, s1, nullptr);
    PBF_LAUNCH_CHECK(ctx);
  } else {
    scan_tile_sums_kernel<<<(unsigned)t2, kScanThreads, 0, ctx->stream>>>(s1, t1, s2);
    PBF_LAUNCH_CHECK(ctx);
    scan_tile_apply_kernel<<<1, kScanThreads, 0, ctx->stream>>>(s2, t2, nullptr, s2, nullptr);
    PBF_LAUNCH_CHECK(ctx);
    scan_tile_apply_kernel<<<(unsigned)t2, kScanThreads, 0, ctx->stream>>>(s1, t1, s2, s1, nullptr);
    PBF_LAUNCH_CHECK(ctx);
  }
  scan_tile_apply_kernel<<<(unsigned)t1, kScanThreads, 0, ctx->stream>>>(in, n, s1, out, total_out_dev);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

// `rows` independent exclusive scans of row_len entries each, in place or not; totals[r] (device) = sum of row r.
int exclusive_scan_rows_u32(pbf_ctx *ctx, const uint32_t *in, uint32_t *out, uint32_t rows, uint32_t row_len, uint32_t *totals_dev) {
  if (rows == 0 || row_len == 0) return PBF_OK;
  if (row_len > kScanSmall) {  // very long rows (> 33 M particles of capacity per rank): one general scan per row
    for (uint32_t r = 0; r < rows; ++r)
      PBF_TRY(exclusive_scan_u32(ctx, in + (size_t)r * row_len, out + (size_t)r * row_len, row_len, totals_dev ? totals_dev + r : nullptr));
    return PBF_OK;
  }
  scan_one_block_kernel<<<rows, kScanWide, 0, ctx->stream>>>(in, row_len, out, totals_dev);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int radix_sort_pairs(pbf_ctx *ctx, const uint32_t *keys_in, uint32_t n, const uint32_t *vals_in, const uint32_t *n_dev) {
  PhaseScope ps(ctx, PBF_PH_SORT);
  if (n == 0) { ctx->keys_sorted = ctx->key_a.p; ctx->perm = ctx->idx_a.p; return PBF_OK; }
  const uint32_t n_tiles = div_up(n, kSortTile);
  PBF_CUDA(ctx, ctx->key_a.reserve(n));
  PBF_CUDA(ctx, ctx->key_b.reserve(n));
  PBF_CUDA(ctx, ctx->idx_a.reserve(n));
  PBF_CUDA(ctx, ctx->idx_b.reserve(n));
  PBF_CUDA(ctx, ctx->sort_hist.reserve((size_t)kBins * n_tiles + kPasses * kBins));
  uint32_t *digit_total = ctx->sort_hist.p + (size_t)kBins * n_tiles;  // one set of 1024 totals per pass
  PBF_CUDA(ctx, cudaMemsetAsync(digit_total, 0, kPasses * kBins * sizeof(uint32_t), ctx->stream));
  const uint32_t *src_k = keys_in, *src_v = vals_in;  // vals_in == nullptr: values are 0..n-1
  uint32_t *dst_k = ctx->key_a.p, *dst_v = ctx->idx_a.p;
  for (int pass = 0; pass < kPasses; ++pass) {
    const int shift = pass * kDigitBits;
    uint32_t *totals = digit_total + pass * kBins;
    sort_hist_kernel<<<n_tiles, kSortThreads, 0, ctx->stream>>>(src_k, n, n_dev, shift, n_tiles, ctx->sort_hist.p, totals);
    PBF_LAUNCH_CHECK(ctx);
    sort_row_scan_kernel<<<kBins / kSortWarps, kSortThreads, 0, ctx->stream>>>(ctx->sort_hist.p, n_tiles);
    PBF_LAUNCH_CHECK(ctx);
    if (pass == 0 && !vals_in)
      sort_scatter_kernel<true><<<n_tiles, kSortThreads, 0, ctx->stream>>>(src_k, nullptr, n, n_dev, shift, n_tiles,
                                                                           ctx->sort_hist.p, totals, dst_k, dst_v);
    else
      sort_scatter_kernel<false><<<n_tiles, kSortThreads, 0, ctx->stream>>>(src_k, src_v, n, n_dev, shift, n_tiles,
                                                                            ctx->sort_hist.p, totals, dst_k, dst_v);
    PBF_LAUNCH_CHECK(ctx);
    src_k = dst_k;
    src_v = dst_v;
    if (dst_k == ctx->key_a.p) { dst_k = ctx->key_b.p; dst_v = ctx->idx_b.p; } else { dst_k = ctx->key_a.p; dst_v = ctx->idx_a.p; }
  }
  ctx->keys_sorted = const_cast<uint32_t *>(src_k);
  ctx->perm = const_cast<uint32_t *>(src_v);
  return PBF_OK;
}

}  // namespace pbf
