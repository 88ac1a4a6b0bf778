// stream_kernels.cu — the HBM-bound streaming kernels of the PBF step:
//   unpack_aos / pack_aos   56-byte sph::Particle records <-> SoA float4 (drop-in boundary, sph.hpp:36-54)
//   predict_key             force -> velocity -> predicted position -> Morton key   (ompsph.hpp:137-154)
//   reorder                 gather by the sort permutation into Z-sorted SoA float4 (ompsph.hpp:158 moves AoS)
//   cell_table              first sorted index with key >= z for every z < G          (sph.hpp:238-250)
//   finalise                position / velocity update with damping                  (ompsph.hpp:256-264)
// All 128-bit coalesced; nothing here is a contraction, so no tensor cores.
#include "common.cuh"

namespace pbf {

namespace {

constexpr int kBlock = 256;

// ---- AoS <-> SoA --------------------------------------------------------------------------------------
// A block moves 256 records (14336 B) through shared memory with 16-byte coalesced global accesses; each
// thread then touches its own 56-byte record in shared memory (stride 14 words: 2-way bank conflict).
constexpr int kRecWords = 14;  // 56 B / 4

__global__ void __launch_bounds__(kBlock) unpack_aos_kernel(const pbf_particle *__restrict__ aos, uint64_t n,
                                                            float4 *__restrict__ pos, float4 *__restrict__ vel,
                                                            float4 *__restrict__ col,
                                                            unsigned long long *__restrict__ ids,
                                                            int *__restrict__ bad_type) {
  __shared__ uint4 stage[kBlock * kRecWords / 4];
  const uint64_t base = (uint64_t)blockIdx.x * kBlock;
  const uint64_t cnt = min((uint64_t)kBlock, n - base);
  const uint32_t n16 = (uint32_t)((cnt * 56 + 15) / 16);
  const uint4 *src = reinterpret_cast<const uint4 *>(aos + base);  // 256*56 is a multiple of 16
  const uint64_t total16 = (n * 56 + 15) / 16;
  const uint64_t off16 = base * 56 / 16;
  for (uint32_t i = threadIdx.x; i < n16; i += kBlock)
    if (off16 + i < total16) stage[i] = __ldg(src + i);
  __syncthreads();
  if (threadIdx.x < cnt) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(stage) + threadIdx.x * kRecWords;
    const uint64_t i = base + threadIdx.x;
    ids[i] = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
    if ((w[2] & 0xFFu) != PBF_TYPE_FLUID) *bad_type = 1;
    pos[i] = make_float4(__uint_as_float(w[4]), __uint_as_float(w[5]), __uint_as_float(w[6]), __uint_as_float(w[3]));
    vel[i] = make_float4(__uint_as_float(w[7]), __uint_as_float(w[8]), __uint_as_float(w[9]), 0.f);
    col[i] = make_float4(__uint_as_float(w[10]), __uint_as_float(w[11]), __uint_as_float(w[12]), __uint_as_float(w[13]));
  }
}

__global__ void __launch_bounds__(kBlock) pack_aos_kernel(pbf_particle *__restrict__ aos, uint64_t n,
                                                          const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                                          const float4 *__restrict__ col,
                                                          const unsigned long long *__restrict__ ids) {
  __shared__ uint4 stage[kBlock * kRecWords / 4];
  const uint64_t base = (uint64_t)blockIdx.x * kBlock;
  const uint64_t cnt = min((uint64_t)kBlock, n - base);
  if (threadIdx.x < cnt) {
    uint32_t *w = reinterpret_cast<uint32_t *>(stage) + threadIdx.x * kRecWords;
    const uint64_t i = base + threadIdx.x;
    const unsigned long long id = ids[i];
    const float4 p = pos[i], v = vel[i], c = col[i];
    w[0] = (uint32_t)id;
    w[1] = (uint32_t)(id >> 32);
    w[2] = PBF_TYPE_FLUID;
    w[3] = __float_as_uint(p.w);
    w[4] = __float_as_uint(p.x); w[5] = __float_as_uint(p.y); w[6] = __float_as_uint(p.z);
    w[7] = __float_as_uint(v.x); w[8] = __float_as_uint(v.y); w[9] = __float_as_uint(v.z);
    w[10] = __float_as_uint(c.x); w[11] = __float_as_uint(c.y); w[12] = __float_as_uint(c.z); w[13] = __float_as_uint(c.w);
  }
  __syncthreads();
  // the staging buffer of a full block is a whole number of 16-byte words; the ragged last block is
  // written word by word so nothing past record n-1 is touched
  if (cnt == kBlock) {
    uint4 *dst = reinterpret_cast<uint4 *>(aos + base);
    for (uint32_t i = threadIdx.x; i < kBlock * kRecWords / 4; i += kBlock) dst[i] = stage[i];
  } else {
    uint32_t *dst = reinterpret_cast<uint32_t *>(aos + base);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(stage);
    for (uint32_t i = threadIdx.x; i < cnt * kRecWords; i += kBlock) dst[i] = src[i];
  }
}

// ---- predict + key --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) predict_key_kernel(StepConst c, const float4 *__restrict__ pos,
                                                             const float4 *__restrict__ vel,
                                                             uint32_t *__restrict__ keys) {
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= count_of(c)) return;
  float v[3], ps[3];
  uint32_t key;
  predict(c, ldg4(pos + i), ldg4(vel + i), v, ps, key);
  keys[i] = key;
}

// ---- reorder (gather) ---------------------------------------------------------------------------------
// Re-derives v' and pStar from the gathered inputs (bit-identical to predict_key) instead of having
// predict_key write them and moving them again: 36 B/particle less traffic.
__global__ void __launch_bounds__(kBlock) reorder_kernel(StepConst c, const uint32_t *__restrict__ perm,
                                                         const float4 *__restrict__ pos_in,
                                                         const float4 *__restrict__ vel_in,
                                                         const float4 *__restrict__ col_in,
                                                         const unsigned long long *__restrict__ ids_in,
                                                         float4 *__restrict__ pos_out, float4 *__restrict__ vel_out,
                                                         float4 *__restrict__ col_out,
                                                         unsigned long long *__restrict__ ids_out,
                                                         float4 *__restrict__ pstar_out) {
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= count_of(c)) return;
  const uint32_t s = __ldg(perm + i);
  const float4 p = ldg4(pos_in + s);
  const float4 v = ldg4(vel_in + s);
  const float4 cc = ldg4(col_in + s);
  const unsigned long long id = __ldg(ids_in + s);
  float vn[3], ps[3];
  uint32_t key;
  predict(c, p, v, vn, ps, key);
  pos_out[i] = p;
  vel_out[i] = make_float4(vn[0], vn[1], vn[2], 0.f);
  col_out[i] = cc;
  ids_out[i] = id;
  pstar_out[i] = make_float4(ps[0], ps[1], ps[2], 0.f);
}

// ---- cell table -----------------------------------------------------------------------------------------
// table[z] = lower_bound(keys_sorted, z): identical to the reference's serial sweep (sph.hpp:243-248) for
// every z, including runs of empty cells, and balanced regardless of how the particles cluster.
__global__ void __launch_bounds__(kBlock) cell_table_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t G,
                                                            uint32_t *__restrict__ table,
                                                            const uint32_t *__restrict__ range_dev) {
  // (A block-level bracket first — two threads bound the block's 256 cells, the rest search inside — was measured in round
  // 2: 26 us against 20 us at 1 M particles, 44.5 against 46.8 on a slab rank with an 8 x larger table: the bracket's two
  // full searches and the barrier cost what the shorter per-thread searches save.)
  const uint32_t z = blockIdx.x * kBlock + threadIdx.x;
  if (z >= G) return;
  uint32_t lo = 0, hi = n;
  if (range_dev) { lo = __ldg(range_dev); hi = lo + __ldg(range_dev + 1); }
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < z) lo = mid + 1; else hi = mid;
  }
  table[z] = lo;
}

// ---- finalise ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) finalise_kernel(StepConst c, const float4 *__restrict__ pstar,
                                                          float4 *__restrict__ pos, float4 *__restrict__ vel) {
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= count_of(c)) return;
  const float4 ps = ldg4(pstar + i);
  float4 p = pos[i];
  float4 v = vel[i];
  const float s[3] = {ps.x, ps.y, ps.z};
  float pp[3] = {p.x, p.y, p.z};
  float vv[3] = {v.x, v.y, v.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float dx = fsub(s[a], fdiv(pp[a], c.scale));
    pp[a] = fmul(s[a], c.scale);
    vv[a] = fmul(fadd(fmul(dx, c.inv_dt), vv[a]), kVD);
  }
  pos[i] = make_float4(pp[0], pp[1], pp[2], p.w);
  vel[i] = make_float4(vv[0], vv[1], vv[2], 0.f);
}

}  // namespace

int launch_unpack_aos(pbf_ctx *ctx, const pbf_particle *aos, uint64_t n, float4 *pos, float4 *vel, float4 *col,
                      unsigned long long *ids, int *bad_type_flag) {
  if (n == 0) return PBF_OK;
  PhaseScope ps(ctx, PBF_PH_PACK);
  unpack_aos_kernel<<<div_up(n, kBlock), kBlock, 0, ctx->stream>>>(aos, n, pos, vel, col, ids, bad_type_flag);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int launch_pack_aos(pbf_ctx *ctx, pbf_particle *aos, uint64_t n, const float4 *pos, const float4 *vel, const float4 *col,
                    const unsigned long long *ids) {
  if (n == 0) return PBF_OK;
  PhaseScope ps(ctx, PBF_PH_PACK);
  pack_aos_kernel<<<div_up(n, kBlock), kBlock, 0, ctx->stream>>>(aos, n, pos, vel, col, ids);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int launch_predict_key(pbf_ctx *ctx, const float4 *pos, const float4 *vel, uint32_t *keys) {
  PhaseScope ps(ctx, PBF_PH_PREDICT_KEY);
  predict_key_kernel<<<div_up(ctx->sc.n, kBlock), kBlock, 0, ctx->stream>>>(ctx->sc, pos, vel, keys);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int launch_reorder(pbf_ctx *ctx, const uint32_t *perm, const float4 *pos_in, const float4 *vel_in, const float4 *col_in,
                   const unsigned long long *ids_in, float4 *pos_out, float4 *vel_out, float4 *col_out,
                   unsigned long long *ids_out, float4 *pstar_out) {
  PhaseScope ps(ctx, PBF_PH_REORDER);
  reorder_kernel<<<div_up(ctx->sc.n, kBlock), kBlock, 0, ctx->stream>>>(ctx->sc, perm, pos_in, vel_in, col_in, ids_in,
                                                                        pos_out, vel_out, col_out, ids_out, pstar_out);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int launch_cell_table(pbf_ctx *ctx, const uint32_t *keys_sorted, uint32_t *table, const uint32_t *range_dev) {
  if (ctx->sc.G == 0) return PBF_OK;
  PhaseScope ps(ctx, PBF_PH_CELL_TABLE);
  cell_table_kernel<<<div_up(ctx->sc.G, kBlock), kBlock, 0, ctx->stream>>>(keys_sorted, ctx->sc.n, ctx->sc.G, table, range_dev);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

int launch_finalise(pbf_ctx *ctx, const float4 *pstar, float4 *pos, float4 *vel) {
  PhaseScope ps(ctx, PBF_PH_FINALISE);
  finalise_kernel<<<div_up(ctx->sc.n, kBlock), kBlock, 0, ctx->stream>>>(ctx->sc, pstar, pos, vel);
  PBF_LAUNCH_CHECK(ctx);
  return PBF_OK;
}

}  // namespace pbf
