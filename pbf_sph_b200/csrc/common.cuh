// common.cuh — shared declarations of the sm_100a PBF backend (context, step constants, device helpers).
//
// Reference semantics cited as file:line are relative to UoB-HPC/pbf-sph (src/omp/ompsph.hpp is the
// backend whose behaviour is reproduced; see DESIGN.md for the kernel-by-kernel map).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "pbf_cuda.h"

namespace pbf {

// ---- physics constants: src/sph_constants.h:5-16 ---------------------------------------------------
constexpr float kVD = 0.49f;
constexpr float kRHO = 6378.0f;
constexpr float kRHO_RECIP = 1.f / kRHO;
constexpr float kEPSILON = 0.00000001f;
constexpr float kCFM_EPSILON = 600.0f;
constexpr float kCorrDeltaQ = 0.3f;
constexpr float kCorrK = 0.0001f;
constexpr float kCorrN = 4.f;

// ---- Morton curve, 10 bits per axis: src/curves.h:46-88 ----------------------------------------------
// Only bits 0..9 and 24..25 of the input survive the first mask, exactly as in the reference's size_t
// arithmetic, so 32-bit maths on the low word is equivalent (x-1 at x==0 wraps to 1023, x+1 at 1023 to 0).
__host__ __device__ __forceinline__ uint32_t spread10(uint32_t v) {
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__host__ __device__ __forceinline__ uint32_t compact10(uint32_t v) {
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030C30C3u;
  v = (v | (v >> 4)) & 0x0300F00Fu;
  v = (v | (v >> 8)) & 0x030000FFu;
  v = (v | (v >> 16)) & 0x000003FFu;
  return v;
}
__host__ __device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
  return spread10(x) | (spread10(y) << 1) | (spread10(z) << 2);
}

// ---- per-step constants handed to every kernel by value ------------------------------------------------
struct StepConst {
  float h, h2;              // smoothing length (solver ctor arg), h*h
  float scale, dt, inv_dt;  // SphParams::scale, dt, 1/dt
  float force[3];           // SphParams::constantForce
  float min_bound[3], max_bound[3];
  float min_extent[3];      // ompsph.hpp:133
  uint32_t extent[3];       // ompsph.hpp:135
  uint32_t G;               // sph.hpp:240
  uint32_t n;               // particles (slab path: an upper bound that sizes the grid when n_dyn is set)
  const uint32_t *n_dyn;    // slab path: the count lives in device memory (the host never reads it back); else nullptr
  float P6, SP, P6dq;       // sph.hpp:251-253, ompsph.hpp:211-213
  float r2_max;             // largest float r2 with sqrtf(r2) <= h  (same neighbour set as "r <= h")
  float r2_min;             // smallest float r2 with sqrtf(r2) >= EPSILON
  float diffuse_mix;        // dt / 750  (ompsph.hpp:203)
  // fast-math folded factors
  float sp_rho;             // SP * RHO_RECIP
  float p6_over_dq;         // P6 / P6dq
  float inv_rho;            // 1 / RHO
  // scene wells (ompsph.hpp:141-148): centre.xyz | force, device array
  uint32_t n_wells;
  const float4 *wells;
};

struct McConst {
  float resolution, isolevel, particle_size, particle_influence;
  float step;       // h / resolution
  float threshold;  // h * scale
  float r2_hit;     // largest squared distance whose glm::fastSqrt is < threshold (mc.cu)
  uint32_t sample[3];
  uint32_t march[3];
  uint64_t lattice_n, march_n;
};

// ---- strict single-precision helpers: never contracted into FMA, IEEE division / sqrt ------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
// glm::min / glm::max semantics (oracle/ref/glm_shim, glm 0.9.9.8 func_common): min(a,b) = (b<a)?b:a
__device__ __forceinline__ float glm_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float glm_max(float a, float b) { return (a < b) ? b : a; }

// float -> size_t as the reference's host code performs it (truncate, two's complement for negatives)
__device__ __forceinline__ uint32_t cell_coord(float v) { return (uint32_t)(unsigned long long)__float2ll_rz(v); }

// Predicted velocity / position and Morton key of one particle: ompsph.hpp:140-153, sph.hpp:198-201.
// Strict arithmetic: the key must be bit-exact.
__device__ __forceinline__ void predict(const StepConst &c, const float4 pos_mass, const float4 vel, float v_out[3],
                                        float ps_out[3], uint32_t &key) {
  const float p[3] = {pos_mass.x, pos_mass.y, pos_mass.z};
  const float v[3] = {vel.x, vel.y, vel.z};
  uint32_t cc[3];
  float force[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) force[a] = fmul(pos_mass.w, c.force[a]);
  // wells — ompsph.hpp:141-148: within 75 units, force += clamp((normalize(centre - pos) * well.force * mass) / dist^2,
  // -10, 10); glm::distance/normalize/clamp as in glm 0.9.9.8 (dot = (x*x + y*y) + z*z, inversesqrt = 1 / sqrt)
  for (uint32_t w = 0; w < c.n_wells; ++w) {
    const float4 well = __ldg(c.wells + w);
    const float d[3] = {fsub(well.x, p[0]), fsub(well.y, p[1]), fsub(well.z, p[2])};
    const float dot = fadd(fadd(fmul(d[0], d[0]), fmul(d[1], d[1])), fmul(d[2], d[2]));
    const float dist = fsqrt(dot);
    if (dist < 75.0f) {
      const float inv = fdiv(1.0f, fsqrt(dot)), dd = fmul(dist, dist);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float fw = fdiv(fmul(fmul(fmul(d[a], inv), well.w), pos_mass.w), dd);
        force[a] = fadd(force[a], glm_min(glm_max(fw, -10.0f), 10.0f));
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float f = force[a];
    v_out[a] = fadd(fmul(f, c.dt), v[a]);
    ps_out[a] = fadd(fmul(v_out[a], c.dt), fdiv(p[a], c.scale));
    cc[a] = cell_coord(fdiv(fsub(ps_out[a], c.min_extent[a]), c.h));
  }
  key = morton3(cc[0], cc[1], cc[2]);
}

// 128-bit read-only loads/stores
__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// number of particles a kernel covers: by value, or from device memory on the slab path
__device__ __forceinline__ uint32_t count_of(const StepConst &c) { return c.n_dyn ? __ldg(c.n_dyn) : c.n; }

// Which particles one launch of a solver pass processes.  Thread t takes particle
//     first + t          for t < count           (a contiguous range of the sorted array), then
//     idx[t - count]     for t - count < n_idx   (a compacted index list: ring-1 ghosts, boundary / interior particles)
// On the slab path count / n_idx are read from device memory (count_dev / n_idx_dev) and `bound` sizes the grid.
// `skip` (slab path: the interior delta pass) drops the particles of the range whose skip[t] != 0 — the range stays in
// place, so a warp's particles keep their 128-byte alignment (a compacted interior list shifted every warp off its lines).
struct Sel {
  uint32_t first = 0, count = 0, n_idx = 0;
  const uint32_t *count_dev = nullptr, *idx = nullptr, *n_idx_dev = nullptr, *skip = nullptr;
  uint32_t bound = 0;  // host-side upper bound of count + n_idx
};
__device__ __forceinline__ bool sel_particle(const Sel &s, uint32_t t, uint32_t &a) {
  const uint32_t count = s.count_dev ? __ldg(s.count_dev) : s.count;
  if (t < count) {
    a = s.first + t;
    return !(s.skip && __ldg(s.skip + t) != 0u);
  }
  if (!s.idx) return false;
  const uint32_t k = t - count;
  if (k >= (s.n_idx_dev ? __ldg(s.n_idx_dev) : s.n_idx)) return false;
  a = __ldg(s.idx + k);
  return true;
}
inline Sel sel_range(uint32_t first, uint32_t count) {
  Sel s;
  s.first = first; s.count = count; s.bound = count;
  return s;
}

// ---- device buffer with geometric growth ------------------------------------------------------------
template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  bool borrowed = false;  // points into memory owned elsewhere (the slab arena, dist.cu): never freed or re-allocated here
  cudaError_t reserve(size_t n, bool keep = false, cudaStream_t s = 0) {
    if (n <= cap) return cudaSuccess;
    if (borrowed) return cudaErrorMemoryAllocation;  // the arena's capacity is fixed between plan steps
    size_t want = n + n / 8 + 256;
    T *q = nullptr;
    cudaError_t e = cudaMalloc(&q, want * sizeof(T));
    if (e != cudaSuccess) return e;
    if (keep && p && cap) {
      e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return e;
      cudaStreamSynchronize(s);
    }
    if (p) cudaFree(p);
    p = q;
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p && !borrowed) cudaFree(p);
    p = nullptr;
    cap = 0;
    borrowed = false;
  }
};

}  // namespace pbf

// ---- scene dynamics of advance() (scene.cu) -------------------------------------------------------------------
struct pbf_scene_state {
  std::vector<pbf_well> wells;
  std::vector<pbf_source> sources;
  std::vector<pbf_drain> drains;
  std::vector<pbf_query> queries;
  pbf::DevBuf<float4> d_wells, d_drains, d_queries;  // centre.xyz | force / width / 0
  pbf::DevBuf<uint2> d_ranges;                       // per query: first particle, count (Z-sorted order)
  uint2 *h_ranges = nullptr;                         // pinned mirror, valid after a stream sync
  size_t h_ranges_cap = 0;
  uint32_t answered = 0;                             // queries answered by the last step
  bool empty() const { return wells.empty() && sources.empty() && drains.empty() && queries.empty(); }
};

// ---- the context -----------------------------------------------------------------------------------
struct pbf_ctx {
  int device = 0;
  float h = 0.1f;
  uint32_t flags = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t side_stream = nullptr;  // colour diffusion runs here, beside the solver iterations
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::string err;
  uint64_t launches = 0;
  int sm_count = 148;
  const uint32_t *diffuse_local_range = nullptr;  // slab path: {first, count} of the local array (device memory), see diffuse_tiled.cu
  int diffuse_blocks_per_sm = 0;  // diffuse_tiled.cu: resident blocks per SM on THIS context's device (0 = not set up yet)

  uint64_t n = 0;  // resident particles
  bool have_state = false;
  // resident SoA state; `cur` = index of the set holding the live (Z-sorted after a step) particles
  pbf::DevBuf<float4> pos[2], vel[2], col[2];
  pbf::DevBuf<unsigned long long> ids[2];
  int cur = 0;      // pos/vel/ids live set
  int cur_col = 0;  // colour live set
  pbf::DevBuf<float4> pstar[2];
  pbf::DevBuf<uint32_t> key_in, key_a, key_b, idx_a, idx_b;
  uint32_t *keys_sorted = nullptr, *perm = nullptr;  // point into the ping-pong buffers after the sort
  pbf::DevBuf<uint32_t> sort_hist, sort_tmp;
  pbf::DevBuf<uint32_t> table;
  pbf::DevBuf<uint32_t> scan_tmp;
  pbf::DevBuf<uint32_t> cand_count, nbr_count;
  pbf::DevBuf<uint32_t> list_hits;  // PBF_TAP_LIST_HITS: the production list's hit counts after the first lambda pass
  pbf::DevBuf<float> rho;
  pbf::DevBuf<pbf_particle> aos;  // staging for the drop-in path
  pbf_particle *host_pinned = nullptr;
  size_t host_pinned_cap = 0;
  void *pin_base = nullptr;  // caller array page-locked under PBF_FLAG_PIN_HOST, and its size
  size_t pin_bytes = 0;
  // marching cubes
  pbf::DevBuf<float4> mc_pn, mc_c;
  const float4 *mc_lattice_pn = nullptr, *mc_lattice_c = nullptr;  // the lattice of the last surface: mc_pn / mc_c, or the slab arena's
  pbf::DevBuf<uint32_t> mc_count, mc_offset;
  pbf::DevBuf<float> mesh_vs, mesh_ns, mesh_cs;
  uint32_t *mc_total_dev = nullptr;   // device word: total triangles
  uint32_t *mc_total_host = nullptr;  // pinned
  uint64_t n_triangles = 0;
  int *flag_dev = nullptr, *flag_host = nullptr;  // "saw a non-Fluid particle" (device word + pinned mirror)
  // tiled diffusion: queue of occupied 4x4x4 cell blocks
  pbf::DevBuf<uint32_t> blk_list, blk_info;  // blk_info[0] = number of occupied blocks, [1] = work counter
  // per-iteration neighbour list: nl[k * nl_stride + particle] = k-th in-radius candidate, nl_count[particle] = hits
  pbf::DevBuf<uint32_t> nl, nl_count;
  uint32_t nl_stride = 0;
  int list_cap = 192;     // hits kept per particle in the neighbour list: kListWide (or kListMax via pbf_debug_set_list_capacity)
  uint32_t nl_cap = 96;   // the capacity the current list was written with (neighbour_list.cu)

  pbf_scene_state scene;

  pbf_grid_info grid{};
  pbf::StepConst sc{};
  pbf::McConst mc{};
  bool mc_valid = false;

  // multi-GPU (dist.cu): slab state, or nullptr on a single-device context
  struct pbf_dist_state *dist = nullptr;

  // profiling (PBF_FLAG_PROFILE)
  static constexpr int kMaxEv = 16384;
  cudaEvent_t ev[kMaxEv];
  int ev_phase[kMaxEv];
  uint64_t ev_launch0[kMaxEv];
  int ev_used = 0;
  bool ev_created = false;
  pbf_profile prof{};
  uint32_t prof_mask = 0xFFFFFFFFu;  // families timed under PBF_FLAG_PROFILE (bit = PBF_PH_*), pbf_profile_set_mask
};

namespace pbf {

// error plumbing --------------------------------------------------------------------------------------
int fail(pbf_ctx *ctx, int code, const char *what, const char *detail);
#define PBF_CUDA(ctx, call)                                                                \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) return pbf::fail((ctx), PBF_ERR_CUDA, #call, cudaGetErrorString(_e)); \
  } while (0)
#define PBF_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != PBF_OK) return _rc; \
  } while (0)

// host-side maths shared with the oracle's definitions --------------------------------------------------
void host_grid(float h, const pbf_params &p, pbf_grid_info &g);
void host_step_const(float h, const pbf_params &p, const pbf_grid_info &g, uint32_t n, StepConst &sc);

// phase profiling ---------------------------------------------------------------------------------------
int prof_begin(pbf_ctx *ctx, int phase, cudaStream_t stream);  // -> event slot (or -1: not timed)
void prof_end(pbf_ctx *ctx, int slot, cudaStream_t stream);
struct PhaseScope {
  pbf_ctx *ctx;
  int phase;
  int slot;
  cudaStream_t stream;  // the stream the scope's work is enqueued on (default: the context's main stream)
  PhaseScope(pbf_ctx *c, int phase, cudaStream_t on = nullptr);
  ~PhaseScope();
};

// kernel launchers (each returns PBF_OK or a negative status) ----------------------------------------------
int launch_unpack_aos(pbf_ctx *ctx, const pbf_particle *aos, uint64_t n, float4 *pos, float4 *vel, float4 *col,
                      unsigned long long *ids, int *bad_type_flag);
int launch_pack_aos(pbf_ctx *ctx, pbf_particle *aos, uint64_t n, const float4 *pos, const float4 *vel, const float4 *col,
                    const unsigned long long *ids);
int launch_predict_key(pbf_ctx *ctx, const float4 *pos, const float4 *vel, uint32_t *keys);
// Stable LSD radix sort of (key, index) pairs over all 30 key bits; on return ctx->keys_sorted / ctx->perm are set.
// vals_in == nullptr sorts (key, 0..n-1); otherwise the given values travel with the keys (multi-GPU merge, dist.cu).
// n_dev != nullptr: the pair count lives in device memory and n is its upper bound (slab path).
int radix_sort_pairs(pbf_ctx *ctx, const uint32_t *keys_in, uint32_t n, const uint32_t *vals_in = nullptr,
                     const uint32_t *n_dev = nullptr);
int launch_reorder(pbf_ctx *ctx, const uint32_t *perm, const float4 *pos_in, const float4 *vel_in, const float4 *col_in,
                   const unsigned long long *ids_in, float4 *pos_out, float4 *vel_out, float4 *col_out,
                   unsigned long long *ids_out, float4 *pstar_out);
// table[z] = first index in [0, n) whose key >= z; with range_dev = {first, count} in device memory the search runs over
// keys_sorted[first .. first + count) and the table holds absolute indices (slab path: the local array starts at `first`).
int launch_cell_table(pbf_ctx *ctx, const uint32_t *keys_sorted, uint32_t *table, const uint32_t *range_dev = nullptr);
int launch_neighbour_counts(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *pstar,
                            uint32_t *cand, uint32_t *nbr);
int launch_diffuse(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *col_in, float4 *col_out);
int launch_lambda(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *pos_mass,
                  const float4 *pstar_in, float4 *pstar_lambda_out, float *rho_out);
int launch_delta(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *pstar_lambda_in,
                 float4 *pstar_out);
// global-memory neighbour kernels over the sorted range [first, first+count)
int launch_lambda_global(pbf_ctx *ctx, uint32_t first, uint32_t count, const uint32_t *keys_sorted,
                         const uint32_t *table, const float4 *pos_mass, const float4 *pstar_in, float4 *pstar_out,
                         float *rho_out);
int launch_delta_global(pbf_ctx *ctx, uint32_t first, uint32_t count, const uint32_t *keys_sorted,
                        const uint32_t *table, const float4 *pstar_in, float4 *pstar_out);
// neighbour-list kernels (neighbour_list.cu): the production lambda/delta passes
constexpr uint32_t kListMax = 96;    // list depth above 22 M particles per device (32-bit list indexing), and the A/B depth of the tests
constexpr uint32_t kListWide = 192;  // production depth; a particle with more hits takes the one-pass path in both passes
int launch_lambda_list(pbf_ctx *ctx, const Sel &sel, const uint32_t *keys_sorted, const uint32_t *table,
                       const float4 *pos_mass, const float4 *pstar_in, float4 *pstar_out, float *rho_out);
int launch_delta_list(pbf_ctx *ctx, const Sel &sel, const uint32_t *keys_sorted, const uint32_t *table,
                      const float4 *pstar_in, float4 *pstar_out);
// The production lambda / delta pass over the selected particles (neighbour-list kernels, or the one-pass global kernels
// under PBF_FLAG_GLOBAL_NEIGHBOURS, which take contiguous ranges only).  list_particles = size of the sorted array the
// neighbour list spans (its row stride); 0 = ctx->sc.n.
int solver_lambda(pbf_ctx *ctx, const Sel &sel, const float4 *pstar_in, float4 *pstar_out, float *rho_out);
int solver_delta(pbf_ctx *ctx, const Sel &sel, const float4 *pstar_in, float4 *pstar_out);
// shared-memory tiled colour diffusion (diffuse_tiled.cu).  own_range_dev = {first, count} in device memory: only cell
// blocks holding particles of that index range are processed (slab path: ghosts get their colours from their owners).
int launch_diffuse_tiled(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *col_in,
                         float4 *col_out, const uint32_t *own_range_dev = nullptr);
int launch_finalise(pbf_ctx *ctx, const float4 *pstar, float4 *pos, float4 *vel);
// opt-in extension after finalise (xsph.cu): XSPH viscosity / vorticity confinement per ctx->flags
int launch_xsph_vorticity(pbf_ctx *ctx, const uint32_t *keys_sorted, const uint32_t *table, const float4 *pstar, float4 *vel,
                          float4 *scratch_omega, float4 *scratch_vel);
int exclusive_scan_u32(pbf_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *total_out_dev);
int exclusive_scan_rows_u32(pbf_ctx *ctx, const uint32_t *in, uint32_t *out, uint32_t rows, uint32_t row_len, uint32_t *totals_dev);
int mc_run(pbf_ctx *ctx, const pbf_params &p, const uint32_t *table, const float4 *pos, const float4 *col);
// the same in three parts (slab path: every rank evaluates the lattice points it owns into rank 0's lattice, rank 0 extracts)
int mc_prepare(pbf_ctx *ctx, const pbf_params &p);
int mc_field(pbf_ctx *ctx, const uint32_t *table, const float4 *pos, const float4 *col, float4 *PN, float4 *LC, uint32_t key_lo,
             uint32_t key_hi);
int mc_extract(pbf_ctx *ctx, const float4 *PN, const float4 *LC);

void dist_release(pbf_ctx *ctx);  // dist.cu
int dist_refresh_counts(pbf_ctx *ctx);  // slab rank: wait for the last step and refresh ctx->n from the device-side counts
// host <-> device plumbing of the drop-in calls (context.cu), shared with the multi-device drop-in call (dist.cu)
int upload_device(pbf_ctx *ctx, const pbf_particle *xs, uint64_t n);  // async H2D + unpack on ctx->stream; sets ctx->n
void host_pin(pbf_ctx *ctx, void *p, size_t bytes);                   // PBF_FLAG_PIN_HOST
void host_unpin(pbf_ctx *ctx);
// scene dynamics (scene.cu)
int scene_set(pbf_ctx *ctx, const pbf_scene *scene);
int scene_edit_particles(pbf_ctx *ctx, const pbf_params &p);  // sources, then drains, on the resident particles
int scene_answer_queries(pbf_ctx *ctx);                       // after the cell table of a step
void scene_emit(float h, float scale, const std::vector<pbf_source> &sources, std::vector<pbf_particle> &out);
void scene_release(pbf_ctx *ctx);

inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

#define PBF_LAUNCH_CHECK(ctx)                                                                               \
  do {                                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                                    \
    if (_e != cudaSuccess) return pbf::fail((ctx), PBF_ERR_CUDA, "kernel launch", cudaGetErrorString(_e)); \
    (ctx)->launches++;                                                                                      \
  } while (0)

}  // namespace pbf
