// context.cu — the C ABI (include/pbf_cuda.h): context lifetime, the per-step schedule, host<->device plumbing,
// parity taps and per-kernel-family CUDA-event profiling.
//
// One step = ompsph.hpp:128-271 re-scheduled for the GPU:
//   predict_key -> radix sort (3 passes) -> reorder -> cell_table -> [counts tap] -> diffuse
//   -> iteration x { lambda, delta } -> finalise -> [marching cubes]
// State lives in HBM as SoA float4 (pos|mass, vel, colour, pStar|lambda) + u64 ids + u32 keys; the drop-in entry
// point converts from/to the reference's 56-byte AoS records on the device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges are no-ops unless a profiler is attached

#include "common.cuh"

namespace pbf {

static std::mutex g_err_mu;
static std::string g_create_err;

int fail(pbf_ctx *ctx, int code, const char *what, const char *detail) {
  std::string m = std::string(what ? what : "") + (detail ? std::string(": ") + detail : std::string());
  if (ctx) ctx->err = m;
  else {
    std::lock_guard<std::mutex> l(g_err_mu);
    g_create_err = m;
  }
  return code;
}

// ---- host maths, written exactly as the reference's host code forms these values -------------------------------
static inline uint64_t to_index(float v) { return (uint64_t)(int64_t)v; }  // sph.hpp:199 (float -> size_t)

void host_grid(float h, const pbf_params &p, pbf_grid_info &g) {
  std::memset(&g, 0, sizeof(g));
  const float padding = h * 2;  // ompsph.hpp:132
  for (int a = 0; a < 3; ++a) {
    const float mn = (p.min_bound[a] / p.scale) - padding;  // ompsph.hpp:133
    const float mx = (p.max_bound[a] / p.scale) + padding;  // ompsph.hpp:134
    g.min_extent[a] = mn;
    g.extent[a] = (uint32_t)to_index((mx - mn) / h);         // ompsph.hpp:135
  }
  g.grid_table_n = morton3(g.extent[0], g.extent[1], g.extent[2]);  // sph.hpp:240
  uint32_t bits = 0;
  if (g.grid_table_n > 1) for (uint32_t v = g.grid_table_n - 1; v; v >>= 1) ++bits;
  g.key_bits = bits;
  g.radix_passes = 3;  // all 30 key bits, 10 per pass (sort_scan.cu)
  if (p.surface_enabled)  // ompsph.hpp:283-284
    for (int a = 0; a < 3; ++a)
      g.sample_size[a] = (uint32_t)to_index(std::floor((float)g.extent[a] * p.surface.resolution)) + 1u;
}

void host_step_const(float h, const pbf_params &p, const pbf_grid_info &g, uint32_t n, StepConst &sc) {
  std::memset(&sc, 0, sizeof(sc));
  sc.h = h;
  sc.h2 = h * h;
  sc.scale = p.scale;
  sc.dt = p.dt;
  sc.inv_dt = 1.0f / p.dt;  // ompsph.hpp:261
  for (int a = 0; a < 3; ++a) {
    sc.force[a] = p.constant_force[a];
    sc.min_bound[a] = p.min_bound[a];
    sc.max_bound[a] = p.max_bound[a];
    sc.min_extent[a] = g.min_extent[a];
    sc.extent[a] = g.extent[a];
  }
  sc.G = g.grid_table_n;
  sc.n = n;
  // sph.hpp:251-253: std::pow(float,int) promotes to double; the result is rounded to float on return
  const float pi = std::acos(-1.0f);
  sc.P6 = (float)((double)315.0f / ((double)(64.0f * pi) * std::pow((double)h, 9.0)));
  sc.SP = (float)(-((double)45.0f / ((double)pi * std::pow((double)h, 6.0))));
  {  // ompsph.hpp:213 with poly6Kernel :67-69
    const float r = kCorrDeltaQ * h;
    const float d = (h * h) - r * r;
    sc.P6dq = sc.P6 * (d * d * d);
  }
  // neighbour-set thresholds on r^2, equivalent to the reference's tests on r = sqrt(r2) (sqrt is monotone)
  float x = h * h;
  while (std::sqrt(std::nextafter(x, INFINITY)) <= h) x = std::nextafter(x, INFINITY);
  while (std::sqrt(x) > h) x = std::nextafter(x, 0.0f);
  sc.r2_max = x;
  x = kEPSILON * kEPSILON;
  while (std::sqrt(x) < kEPSILON) x = std::nextafter(x, INFINITY);
  while (std::sqrt(std::nextafter(x, 0.0f)) >= kEPSILON) x = std::nextafter(x, 0.0f);
  sc.r2_min = x;
  sc.diffuse_mix = p.dt / 750.0f;  // ompsph.hpp:203
  sc.sp_rho = sc.SP * kRHO_RECIP;
  sc.p6_over_dq = sc.P6 / sc.P6dq;
  sc.inv_rho = 1.0f / kRHO;
}

// ---- profiling ---------------------------------------------------------------------------------------------------
// NVTX range names = the reference's stopwatch phases (ompsph.hpp:89,130,157,161,188,209,252,279,359,399,479), so a
// timeline of this backend reads like the reference's own "advance" breakdown.
static const char *const kPhaseNames[PBF_PH_COUNT] = {
    "advect+copy (predict_key)", "sortz (radix sort)", "sortz (reorder)", "gridtable", "sph-diffuse", "sph-lambda", "sph-delta",
    "sph-finalise", "mc-field", "mc_psum", "gpu_mc", "write back", "halo", "slab-setup", "slab-barrier", "slab-iterations"};

// A timed span that does not fit a C++ scope (dist.cu: the pre-iteration part of a slab step): returns the event slot, or -1.
int prof_begin(pbf_ctx *ctx, int ph, cudaStream_t stream) {
  if (!(ctx->flags & PBF_FLAG_PROFILE) || !((ctx->prof_mask >> ph) & 1u)) return -1;
  if (!ctx->ev_created) {
    for (int i = 0; i < pbf_ctx::kMaxEv; ++i) cudaEventCreate(&ctx->ev[i]);
    ctx->ev_created = true;
  }
  if (ctx->ev_used + 2 > pbf_ctx::kMaxEv) return -1;  // full: this span goes untimed until the next read
  const int slot = ctx->ev_used;
  ctx->ev_used += 2;
  ctx->ev_phase[slot] = ph;
  ctx->ev_launch0[slot] = ctx->launches;
  cudaEventRecord(ctx->ev[slot], stream);
  return slot;
}
void prof_end(pbf_ctx *ctx, int slot, cudaStream_t stream) {
  if (slot < 0) return;
  cudaEventRecord(ctx->ev[slot + 1], stream);
  ctx->prof.launches[ctx->ev_phase[slot]] += ctx->launches - ctx->ev_launch0[slot];
}

PhaseScope::PhaseScope(pbf_ctx *c, int ph, cudaStream_t on) : ctx(c), phase(ph), slot(-1), stream(on ? on : c->stream) {
  nvtxRangePushA(kPhaseNames[ph]);
  slot = prof_begin(ctx, ph, stream);
}
PhaseScope::~PhaseScope() {
  nvtxRangePop();
  prof_end(ctx, slot, stream);
}

static void profile_collect(pbf_ctx *ctx) {
  if (ctx->ev_used == 0) return;
  cudaStreamSynchronize(ctx->stream);
  for (int s = 0; s < ctx->ev_used; s += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[s], ctx->ev[s + 1]) == cudaSuccess) ctx->prof.ms[ctx->ev_phase[s]] += ms;
  }
  ctx->ev_used = 0;
}

// ---- the step ------------------------------------------------------------------------------------------------------
int solver_lambda(pbf_ctx *ctx, const Sel &sel, const float4 *pstar_in, float4 *pstar_out, float *rho_out) {
  PhaseScope ps(ctx, PBF_PH_LAMBDA);
  const float4 *pos_mass = ctx->pos[ctx->cur].p;
  if (ctx->flags & PBF_FLAG_GLOBAL_NEIGHBOURS) {
    if (sel.idx || sel.count_dev) return fail(ctx, PBF_ERR_STATE, "PBF_FLAG_GLOBAL_NEIGHBOURS", "contiguous host-sized ranges only (not on the slab path)");
    return launch_lambda_global(ctx, sel.first, sel.count, ctx->keys_sorted, ctx->table.p, pos_mass, pstar_in, pstar_out, rho_out);
  }
  return launch_lambda_list(ctx, sel, ctx->keys_sorted, ctx->table.p, pos_mass, pstar_in, pstar_out, rho_out);
}

int solver_delta(pbf_ctx *ctx, const Sel &sel, const float4 *pstar_in, float4 *pstar_out) {
  PhaseScope ps(ctx, PBF_PH_DELTA);
  if (ctx->flags & PBF_FLAG_GLOBAL_NEIGHBOURS) {
    if (sel.idx || sel.count_dev) return fail(ctx, PBF_ERR_STATE, "PBF_FLAG_GLOBAL_NEIGHBOURS", "contiguous host-sized ranges only (not on the slab path)");
    return launch_delta_global(ctx, sel.first, sel.count, ctx->keys_sorted, ctx->table.p, pstar_in, pstar_out);
  }
  return launch_delta_list(ctx, sel, ctx->keys_sorted, ctx->table.p, pstar_in, pstar_out);
}

static int validate(pbf_ctx *ctx, const pbf_params *p) {
  if (!p) return fail(ctx, PBF_ERR_INVALID, "params", "NULL");
  if (!(p->scale > 0.f) || !(p->dt > 0.f)) return fail(ctx, PBF_ERR_INVALID, "params", "scale and dt must be > 0");
  pbf_grid_info g;
  host_grid(ctx->h, *p, g);
  for (int a = 0; a < 3; ++a)
    if (g.extent[a] == 0 || g.extent[a] > 1023)
      return fail(ctx, PBF_ERR_INVALID, "grid", "extent must be 1..1023 cells per axis (10-bit Morton, curves.h:73)");
  return PBF_OK;
}

struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static int step_device(pbf_ctx *ctx, const pbf_params &p) {
  NvtxRange whole("advance");  // ompsph.hpp:89
  if (!ctx->scene.empty()) {
    NvtxRange r("source+drain");  // ompsph.hpp:91
    PBF_TRY(validate(ctx, &p));
    PBF_TRY(scene_edit_particles(ctx, p));  // sources, then drains (ompsph.hpp:91-120)
  }
  ctx->scene.answered = 0;
  const uint64_t n64 = ctx->n;
  if (n64 == 0) return PBF_OK;  // "Particles depleted" (ompsph.hpp:122-126)
  if (n64 >= 0xFFFFFFF0ull) return fail(ctx, PBF_ERR_INVALID, "n", "more than 2^32 particles on one device");
  const uint32_t n = (uint32_t)n64;
  PBF_TRY(validate(ctx, &p));
  host_grid(ctx->h, p, ctx->grid);
  ctx->grid.n_particles = n;
  host_step_const(ctx->h, p, ctx->grid, n, ctx->sc);
  ctx->sc.n_wells = (uint32_t)ctx->scene.wells.size();
  ctx->sc.wells = ctx->scene.d_wells.p;
  const int o = ctx->cur ^ 1, oc = ctx->cur_col ^ 1;
  PBF_CUDA(ctx, ctx->pos[o].reserve(n));
  PBF_CUDA(ctx, ctx->vel[o].reserve(n));
  PBF_CUDA(ctx, ctx->ids[o].reserve(n));
  PBF_CUDA(ctx, ctx->col[oc].reserve(n));
  PBF_CUDA(ctx, ctx->pstar[0].reserve(n));
  PBF_CUDA(ctx, ctx->pstar[1].reserve(n));
  PBF_CUDA(ctx, ctx->key_in.reserve(n));
  PBF_CUDA(ctx, ctx->table.reserve((size_t)ctx->sc.G + 1));
  PBF_CUDA(ctx, ctx->rho.reserve(n));

  PBF_TRY(launch_predict_key(ctx, ctx->pos[ctx->cur].p, ctx->vel[ctx->cur].p, ctx->key_in.p));
  PBF_TRY(radix_sort_pairs(ctx, ctx->key_in.p, n));
  PBF_TRY(launch_reorder(ctx, ctx->perm, ctx->pos[ctx->cur].p, ctx->vel[ctx->cur].p, ctx->col[ctx->cur_col].p,
                         ctx->ids[ctx->cur].p, ctx->pos[o].p, ctx->vel[o].p, ctx->col[oc].p, ctx->ids[o].p,
                         ctx->pstar[0].p));
  ctx->cur = o;
  ctx->cur_col = oc;
  PBF_TRY(launch_cell_table(ctx, ctx->keys_sorted, ctx->table.p));
  PBF_TRY(scene_answer_queries(ctx));  // ompsph.hpp:167-186
  const bool tiled = !(ctx->flags & PBF_FLAG_GLOBAL_NEIGHBOURS);
  if (ctx->flags & PBF_FLAG_DEBUG_COUNTS) {
    PBF_CUDA(ctx, ctx->cand_count.reserve(n));
    PBF_CUDA(ctx, ctx->nbr_count.reserve(n));
    PBF_TRY(launch_neighbour_counts(ctx, ctx->keys_sorted, ctx->table.p, ctx->pstar[0].p, ctx->cand_count.p,
                                    ctx->nbr_count.p));
  }
  PBF_CUDA(ctx, ctx->col[ctx->cur_col ^ 1].reserve(n));
  // Colour diffusion feeds nothing inside the step (ompsph.hpp:189-206 touches colours only): it runs on a side stream
  // beside the solver iterations — a shared-memory, latency-bound kernel under L1/issue-bound ones — and joins before
  // finalise, i.e. before anything can read the colours.
  {
    PBF_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    PBF_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = ctx->side_stream;
    const int rc = tiled ? launch_diffuse_tiled(ctx, ctx->keys_sorted, ctx->table.p, ctx->col[ctx->cur_col].p,
                                                ctx->col[ctx->cur_col ^ 1].p)
                         : launch_diffuse(ctx, ctx->keys_sorted, ctx->table.p, ctx->col[ctx->cur_col].p,
                                          ctx->col[ctx->cur_col ^ 1].p);
    ctx->stream = main_stream;
    PBF_TRY(rc);
    PBF_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
  }
  ctx->cur_col ^= 1;
  for (uint64_t it = 0; it < p.iteration; ++it) {
    PBF_TRY(solver_lambda(ctx, sel_range(0, n), ctx->pstar[0].p, ctx->pstar[1].p, it + 1 == p.iteration ? ctx->rho.p : nullptr));
    if (it == 0 && (ctx->flags & PBF_FLAG_DEBUG_COUNTS) && tiled) {  // PBF_TAP_LIST_HITS
      PBF_CUDA(ctx, ctx->list_hits.reserve(n));
      PBF_CUDA(ctx, cudaMemcpyAsync(ctx->list_hits.p, ctx->nl_count.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    PBF_TRY(solver_delta(ctx, sel_range(0, n), ctx->pstar[1].p, ctx->pstar[0].p));
  }
  PBF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  PBF_TRY(launch_finalise(ctx, ctx->pstar[0].p, ctx->pos[ctx->cur].p, ctx->vel[ctx->cur].p));
  // opt-in extension (no reference backend has it): scratch = the spare position / velocity set of the reorder
  PBF_TRY(launch_xsph_vorticity(ctx, ctx->keys_sorted, ctx->table.p, ctx->pstar[0].p, ctx->vel[ctx->cur].p,
                                ctx->pos[ctx->cur ^ 1].p, ctx->vel[ctx->cur ^ 1].p));
  ctx->n_triangles = 0;
  ctx->mc_valid = false;
  if (p.surface_enabled)
    PBF_TRY(mc_run(ctx, p, ctx->table.p, ctx->pos[ctx->cur].p, ctx->col[ctx->cur_col].p));
  ctx->prof.steps++;
  return PBF_OK;
}

int upload_device(pbf_ctx *ctx, const pbf_particle *xs, uint64_t n) {
  ctx->n = 0;
  ctx->have_state = false;
  if (n >= 0xFFFFFFF0ull) return fail(ctx, PBF_ERR_INVALID, "n", "more than 2^32 particles on one device");
  PBF_CUDA(ctx, ctx->aos.reserve(n + 1));
  PBF_CUDA(ctx, ctx->pos[ctx->cur].reserve(n));
  PBF_CUDA(ctx, ctx->vel[ctx->cur].reserve(n));
  PBF_CUDA(ctx, ctx->col[ctx->cur_col].reserve(n));
  PBF_CUDA(ctx, ctx->ids[ctx->cur].reserve(n));
  if (n) {
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->aos.p, xs, n * sizeof(pbf_particle), cudaMemcpyHostToDevice, ctx->stream));
    PBF_CUDA(ctx, cudaMemsetAsync(ctx->flag_dev, 0, sizeof(int), ctx->stream));
    PBF_TRY(launch_unpack_aos(ctx, ctx->aos.p, n, ctx->pos[ctx->cur].p, ctx->vel[ctx->cur].p, ctx->col[ctx->cur_col].p,
                              ctx->ids[ctx->cur].p, ctx->flag_dev));
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->flag_host, ctx->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  }
  ctx->n = n;
  ctx->have_state = true;
  return PBF_OK;
}

// PBF_FLAG_PIN_HOST: page-lock the caller's array so that the H2D / D2H of the drop-in call are DMA transfers.
void host_unpin(pbf_ctx *ctx) {
  if (!ctx->pin_base) return;
  cudaHostUnregister(ctx->pin_base);
  cudaGetLastError();  // a caller that freed the array first has broken the contract; nothing to unwind here
  ctx->pin_base = nullptr;
  ctx->pin_bytes = 0;
}

void host_pin(pbf_ctx *ctx, void *p, size_t bytes) {
  if (ctx->pin_base == p && bytes <= ctx->pin_bytes) return;  // the array we already hold
  host_unpin(ctx);
  if (!p || bytes < (1u << 20)) return;  // small arrays: registration costs more than it saves
  if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess) {
    ctx->pin_base = p;
    ctx->pin_bytes = bytes;
  } else {
    cudaGetLastError();  // already page-locked by the caller, or not registrable: plain copies work regardless
  }
}

static int download_device(pbf_ctx *ctx, pbf_particle *xs, uint64_t n) {
  if (n == 0) return PBF_OK;
  PBF_CUDA(ctx, ctx->aos.reserve(n + 1));
  PBF_TRY(launch_pack_aos(ctx, ctx->aos.p, n, ctx->pos[ctx->cur].p, ctx->vel[ctx->cur].p, ctx->col[ctx->cur_col].p,
                          ctx->ids[ctx->cur].p));
  PBF_CUDA(ctx, cudaMemcpyAsync(xs, ctx->aos.p, n * sizeof(pbf_particle), cudaMemcpyDeviceToHost, ctx->stream));
  return PBF_OK;
}

}  // namespace pbf

using namespace pbf;

#define PBF_ENTER(ctx)                                                        \
  if (!(ctx)) return pbf::fail(nullptr, PBF_ERR_INVALID, "ctx", "NULL");      \
  PBF_CUDA((ctx), cudaSetDevice((ctx)->device)) /* callers may be on any thread (visualise.cpp:85-109) */

extern "C" {

int pbf_abi_version(void) { return PBF_ABI_VERSION; }

int pbf_create(pbf_ctx **out, float h, int device) {
  if (!out) return fail(nullptr, PBF_ERR_INVALID, "pbf_create", "out is NULL");
  *out = nullptr;
  if (!(h > 0.f)) return fail(nullptr, PBF_ERR_INVALID, "pbf_create", "h must be > 0");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, PBF_ERR_CUDA, "pbf_create: no CUDA device (this backend has no CPU fallback)",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= count) return fail(nullptr, PBF_ERR_INVALID, "pbf_create", "device ordinal out of range");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, PBF_ERR_CUDA, "cudaSetDevice", cudaGetErrorString(e));
  pbf_ctx *ctx = new pbf_ctx();
  ctx->device = device;
  ctx->h = h;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->flag_dev, 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaHostAlloc(&ctx->flag_host, 4 * sizeof(int), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->mc_total_dev, 4 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaHostAlloc(&ctx->mc_total_host, 4 * sizeof(uint32_t), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    fail(nullptr, PBF_ERR_CUDA, "pbf_create", cudaGetErrorString(e));
    delete ctx;
    return PBF_ERR_CUDA;
  }
  ctx->own_stream = true;
  *out = ctx;
  return PBF_OK;
}

int pbf_unpin_host(pbf_ctx *ctx) {
  PBF_ENTER(ctx);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  host_unpin(ctx);
  return PBF_OK;
}

void pbf_destroy(pbf_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  host_unpin(ctx);
  dist_release(ctx);
  scene_release(ctx);
  for (int i = 0; i < 2; ++i) {
    ctx->pos[i].release(); ctx->vel[i].release(); ctx->col[i].release(); ctx->ids[i].release(); ctx->pstar[i].release();
  }
  ctx->key_in.release(); ctx->key_a.release(); ctx->key_b.release(); ctx->idx_a.release(); ctx->idx_b.release();
  ctx->sort_hist.release(); ctx->sort_tmp.release(); ctx->table.release(); ctx->scan_tmp.release();
  ctx->cand_count.release(); ctx->nbr_count.release(); ctx->list_hits.release(); ctx->rho.release(); ctx->aos.release();
  ctx->mc_pn.release(); ctx->mc_c.release(); ctx->mc_count.release(); ctx->mc_offset.release();
  ctx->mesh_vs.release(); ctx->mesh_ns.release(); ctx->mesh_cs.release();
  ctx->blk_list.release(); ctx->blk_info.release(); ctx->nl.release(); ctx->nl_count.release();
  if (ctx->flag_dev) cudaFree(ctx->flag_dev);
  if (ctx->flag_host) cudaFreeHost(ctx->flag_host);
  if (ctx->mc_total_dev) cudaFree(ctx->mc_total_dev);
  if (ctx->mc_total_host) cudaFreeHost(ctx->mc_total_host);
  if (ctx->ev_created) for (int i = 0; i < pbf_ctx::kMaxEv; ++i) cudaEventDestroy(ctx->ev[i]);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  delete ctx;
}

const char *pbf_last_error(const pbf_ctx *ctx) {
  if (ctx) return ctx->err.c_str();
  std::lock_guard<std::mutex> l(g_err_mu);
  static thread_local std::string copy;
  copy = g_create_err;
  return copy.c_str();
}

int pbf_set_flags(pbf_ctx *ctx, uint32_t flags) {
  PBF_ENTER(ctx);
  ctx->flags = flags;
  return PBF_OK;
}

int pbf_set_stream(pbf_ctx *ctx, void *cuda_stream) {
  PBF_ENTER(ctx);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return PBF_OK;
}

int pbf_advance_scene_host(pbf_ctx *ctx, const pbf_params *params, const pbf_scene *scene, pbf_particle *xs, uint64_t n,
                           uint64_t capacity, uint64_t *n_out, uint64_t *n_mesh_vertices) {
  PBF_ENTER(ctx);
  if (n_mesh_vertices) *n_mesh_vertices = 0;
  if (n_out) *n_out = n;
  if (ctx->dist) return fail(ctx, PBF_ERR_STATE, "pbf_advance_host", "this context is a slab rank: use pbf_dist_step");
  if (!ctx->scene.empty() || scene) PBF_TRY(scene_set(ctx, scene));
  if (n && !xs) return fail(ctx, PBF_ERR_INVALID, "xs", "NULL");
  PBF_TRY(validate(ctx, params));
  if (!ctx->scene.sources.empty()) {  // the emitted particles must fit the caller's array
    std::vector<pbf_particle> fresh;
    scene_emit(ctx->h, params->scale, ctx->scene.sources, fresh);
    if (n + fresh.size() > capacity) return fail(ctx, PBF_ERR_CAPACITY, "xs", "capacity too small for the particles the sources emit");
    if (!xs && !fresh.empty()) return fail(ctx, PBF_ERR_INVALID, "xs", "NULL");
  }
  if (ctx->flags & PBF_FLAG_PIN_HOST) host_pin(ctx, xs, (size_t)capacity * sizeof(pbf_particle));
  PBF_TRY(upload_device(ctx, xs, n));
  PBF_TRY(step_device(ctx, *params));
  // the type check result is known only now; the particle array is left untouched on failure
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n && *ctx->flag_host) {
    ctx->n = 0; ctx->have_state = false;
    return fail(ctx, PBF_ERR_INVALID, "xs", "Obstacle particles are not supported (the reference OMP backend drops them)");
  }
  if (n_out) *n_out = ctx->n;
  if (ctx->n == 0) return PBF_OK;  // ompsph.hpp:122-126: nothing to do / particles depleted
  PBF_TRY(download_device(ctx, xs, ctx->n));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_mesh_vertices) *n_mesh_vertices = ctx->n_triangles * 3;
  return PBF_OK;
}

int pbf_advance_host(pbf_ctx *ctx, const pbf_params *params, pbf_particle *xs, uint64_t n, uint64_t *n_mesh_vertices) {
  return pbf_advance_scene_host(ctx, params, nullptr, xs, n, n, nullptr, n_mesh_vertices);
}

int pbf_set_scene(pbf_ctx *ctx, const pbf_scene *scene) {
  PBF_ENTER(ctx);
  if (ctx->dist && scene && (scene->n_wells || scene->n_sources || scene->n_drains || scene->n_queries))
    return fail(ctx, PBF_ERR_STATE, "pbf_set_scene", "scene dynamics are not available on the slab path");
  return scene_set(ctx, scene);
}

int pbf_query_result(pbf_ctx *ctx, uint32_t index, uint64_t *ids, uint64_t capacity, uint64_t *count) {
  PBF_ENTER(ctx);
  if (count) *count = 0;
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (index >= ctx->scene.answered) return fail(ctx, PBF_ERR_INVALID, "index", "no such query in the last step");
  const uint2 r = ctx->scene.h_ranges[index];
  if (count) *count = r.y;
  if (r.y == 0) return PBF_OK;
  if (!ids || capacity < r.y) return fail(ctx, PBF_ERR_CAPACITY, "pbf_query_result", "ids too small");
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "id width");
  PBF_CUDA(ctx, cudaMemcpyAsync(ids, ctx->ids[ctx->cur].p + r.x, (size_t)r.y * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PBF_OK;
}

int pbf_mesh_download(pbf_ctx *ctx, float *vs, float *ns, float *cs, uint64_t capacity_vertices) {
  PBF_ENTER(ctx);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const uint64_t nv = ctx->n_triangles * 3;
  if (nv > capacity_vertices) return fail(ctx, PBF_ERR_CAPACITY, "pbf_mesh_download", "capacity_vertices too small");
  if (nv == 0) return PBF_OK;
  if (vs) PBF_CUDA(ctx, cudaMemcpyAsync(vs, ctx->mesh_vs.p, nv * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (ns) PBF_CUDA(ctx, cudaMemcpyAsync(ns, ctx->mesh_ns.p, nv * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (cs) PBF_CUDA(ctx, cudaMemcpyAsync(cs, ctx->mesh_cs.p, nv * 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PBF_OK;
}

int pbf_mesh_device(pbf_ctx *ctx, const float **vs, const float **ns, const float **cs, uint64_t *n_vertices) {
  PBF_ENTER(ctx);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const uint64_t nv = ctx->n_triangles * 3;
  if (n_vertices) *n_vertices = nv;
  if (vs) *vs = nv ? ctx->mesh_vs.p : nullptr;
  if (ns) *ns = nv ? ctx->mesh_ns.p : nullptr;
  if (cs) *cs = nv ? ctx->mesh_cs.p : nullptr;
  return PBF_OK;
}

int pbf_upload(pbf_ctx *ctx, const pbf_particle *xs, uint64_t n) {
  PBF_ENTER(ctx);
  if (n && !xs) return fail(ctx, PBF_ERR_INVALID, "xs", "NULL");
  PBF_TRY(upload_device(ctx, xs, n));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n && *ctx->flag_host) {
    ctx->n = 0; ctx->have_state = false;
    return fail(ctx, PBF_ERR_INVALID, "xs", "Obstacle particles are not supported (the reference OMP backend drops them)");
  }
  return PBF_OK;
}

int pbf_step(pbf_ctx *ctx, const pbf_params *params) {
  PBF_ENTER(ctx);
  if (!ctx->have_state) return fail(ctx, PBF_ERR_STATE, "pbf_step", "no resident particles: call pbf_upload first");
  if (!params) return fail(ctx, PBF_ERR_INVALID, "params", "NULL");
  if (ctx->dist) return fail(ctx, PBF_ERR_STATE, "pbf_step", "this context is a slab rank: use pbf_dist_step");
  return step_device(ctx, *params);
}

int pbf_sync(pbf_ctx *ctx) {
  PBF_ENTER(ctx);
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->mc_valid) ctx->n_triangles = ctx->mc_total_host[0];
  return PBF_OK;
}

int pbf_download(pbf_ctx *ctx, pbf_particle *xs, uint64_t capacity, uint64_t *n_out) {
  PBF_ENTER(ctx);
  if (!ctx->have_state) return fail(ctx, PBF_ERR_STATE, "pbf_download", "no resident particles");
  if (ctx->dist) return fail(ctx, PBF_ERR_STATE, "pbf_download", "this context is a slab rank: use pbf_dist_download");
  if (n_out) *n_out = ctx->n;
  if (capacity < ctx->n) return fail(ctx, PBF_ERR_CAPACITY, "pbf_download", "capacity too small");
  if (ctx->n && !xs) return fail(ctx, PBF_ERR_INVALID, "xs", "NULL");
  PBF_TRY(download_device(ctx, xs, ctx->n));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PBF_OK;
}

int pbf_particle_count(pbf_ctx *ctx, uint64_t *n_out) {
  if (!ctx || !n_out) return fail(ctx, PBF_ERR_INVALID, "pbf_particle_count", "NULL");
  if (ctx->dist) PBF_TRY(dist_refresh_counts(ctx));  // slab rank: the count lives on the device between plan steps
  *n_out = ctx->n;
  return PBF_OK;
}

int pbf_device_state(pbf_ctx *ctx, void **pos4, void **vel4, void **col4, void **ids) {
  PBF_ENTER(ctx);
  if (!ctx->have_state) return fail(ctx, PBF_ERR_STATE, "pbf_device_state", "no resident particles");
  if (pos4) *pos4 = ctx->pos[ctx->cur].p;
  if (vel4) *vel4 = ctx->vel[ctx->cur].p;
  if (col4) *col4 = ctx->col[ctx->cur_col].p;
  if (ids) *ids = ctx->ids[ctx->cur].p;
  return PBF_OK;
}

int pbf_grid(pbf_ctx *ctx, pbf_grid_info *out) {
  PBF_ENTER(ctx);
  if (!out) return fail(ctx, PBF_ERR_INVALID, "out", "NULL");
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = ctx->grid;
  if (ctx->mc_valid) ctx->n_triangles = ctx->mc_total_host[0];
  out->n_triangles = (uint32_t)ctx->n_triangles;
  return PBF_OK;
}

int pbf_debug_read(pbf_ctx *ctx, int tap, void *dst, uint64_t dst_bytes) {
  PBF_ENTER(ctx);
  if (!dst) return fail(ctx, PBF_ERR_INVALID, "dst", "NULL");
  if (ctx->dist && tap != PBF_TAP_MC_FIELD && tap != PBF_TAP_MC_COLOUR)  // the surface lattice is not a particle array
    return fail(ctx, PBF_ERR_STATE, "pbf_debug_read", "taps are single-device (a slab rank's arrays are laid out by the arena)");
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const uint64_t n = ctx->n;
  const void *src = nullptr;
  uint64_t bytes = 0;
  bool strided_w = false;
  switch (tap) {
    case PBF_TAP_KEYS_INPUT: src = ctx->key_in.p; bytes = n * 4; break;
    case PBF_TAP_PERM: src = ctx->perm; bytes = n * 4; break;
    case PBF_TAP_KEYS_SORTED: src = ctx->keys_sorted; bytes = n * 4; break;
    case PBF_TAP_CELL_TABLE: src = ctx->table.p; bytes = (uint64_t)ctx->grid.grid_table_n * 4; break;
    case PBF_TAP_CAND_COUNT: src = ctx->cand_count.p; bytes = n * 4; break;
    case PBF_TAP_NBR_COUNT: src = ctx->nbr_count.p; bytes = n * 4; break;
    case PBF_TAP_LIST_HITS: src = ctx->list_hits.p; bytes = n * 4; break;
    case PBF_TAP_LAMBDA: src = ctx->pstar[1].p; bytes = n * 4; strided_w = true; break;
    case PBF_TAP_RHO: src = ctx->rho.p; bytes = n * 4; break;
    case PBF_TAP_IDS: src = ctx->ids[ctx->cur].p; bytes = n * 8; break;
    case PBF_TAP_MC_FIELD: src = ctx->mc_lattice_pn; bytes = ctx->mc_valid ? ctx->mc.lattice_n * 16 : 0; break;
    case PBF_TAP_MC_COLOUR: src = ctx->mc_lattice_c; bytes = ctx->mc_valid ? ctx->mc.lattice_n * 16 : 0; break;
    default: return fail(ctx, PBF_ERR_INVALID, "tap", "unknown");
  }
  if (!src || bytes == 0) return fail(ctx, PBF_ERR_STATE, "pbf_debug_read", "tap not available (no step run, or flag not set)");
  if (dst_bytes < bytes) return fail(ctx, PBF_ERR_CAPACITY, "pbf_debug_read", "dst too small");
  if (strided_w)
    PBF_CUDA(ctx, cudaMemcpy2DAsync(dst, 4, (const char *)src + 12, 16, 4, n, cudaMemcpyDeviceToHost, ctx->stream));
  else
    PBF_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PBF_OK;
}

int pbf_debug_set_list_capacity(pbf_ctx *ctx, uint32_t hits) {
  PBF_ENTER(ctx);
  if (hits != kListMax && hits != kListWide) return fail(ctx, PBF_ERR_INVALID, "pbf_debug_set_list_capacity", "96 or 192");
  ctx->list_cap = (int)hits;
  return PBF_OK;
}

int pbf_debug_scan_u32(pbf_ctx *ctx, const uint32_t *in, uint64_t n, uint32_t *out, uint32_t *total_out) {
  PBF_ENTER(ctx);
  if ((!in || !out) && n) return fail(ctx, PBF_ERR_INVALID, "pbf_debug_scan_u32", "NULL array");
  DevBuf<uint32_t> buf;
  PBF_CUDA(ctx, buf.reserve(n + 1));
  PBF_CUDA(ctx, cudaMemcpyAsync(buf.p, in, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = exclusive_scan_u32(ctx, buf.p, buf.p, n, buf.p + n);
  if (rc == PBF_OK) {
    cudaError_t e = cudaMemcpyAsync(out, buf.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && total_out) e = cudaMemcpyAsync(total_out, buf.p + n, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, PBF_ERR_CUDA, "pbf_debug_scan_u32", cudaGetErrorString(e));
  }
  buf.release();
  return rc;
}

int pbf_debug_sort_pairs(pbf_ctx *ctx, const uint32_t *keys_in, uint32_t n, uint32_t *keys_out, uint32_t *perm_out) {
  PBF_ENTER(ctx);
  if ((!keys_in || !keys_out || !perm_out) && n) return fail(ctx, PBF_ERR_INVALID, "pbf_debug_sort_pairs", "NULL array");
  if (n == 0) return PBF_OK;
  DevBuf<uint32_t> buf;
  PBF_CUDA(ctx, buf.reserve(n));
  PBF_CUDA(ctx, cudaMemcpyAsync(buf.p, keys_in, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = radix_sort_pairs(ctx, buf.p, n);
  if (rc == PBF_OK) {
    cudaError_t e = cudaMemcpyAsync(keys_out, ctx->keys_sorted, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(perm_out, ctx->perm, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, PBF_ERR_CUDA, "pbf_debug_sort_pairs", cudaGetErrorString(e));
  }
  buf.release();
  return rc;
}

int pbf_profile_reset(pbf_ctx *ctx) {
  PBF_ENTER(ctx);
  profile_collect(ctx);
  std::memset(&ctx->prof, 0, sizeof(ctx->prof));
  return PBF_OK;
}

int pbf_profile_set_mask(pbf_ctx *ctx, uint32_t family_mask) {
  PBF_ENTER(ctx);
  ctx->prof_mask = family_mask;
  return PBF_OK;
}

int pbf_profile_read(pbf_ctx *ctx, pbf_profile *out) {
  PBF_ENTER(ctx);
  if (!out) return fail(ctx, PBF_ERR_INVALID, "out", "NULL");
  profile_collect(ctx);
  *out = ctx->prof;
  return PBF_OK;
}

uint64_t pbf_launch_count(const pbf_ctx *ctx) { return ctx ? ctx->launches : 0; }

void *pbf_host_alloc(uint64_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void pbf_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---- host-only helpers ----------------------------------------------------------------------------------------------
int pbf_host_grid(float h, const pbf_params *params, pbf_grid_info *out) {
  if (!params || !out || !(h > 0.f) || !(params->scale > 0.f)) return PBF_ERR_INVALID;
  host_grid(h, *params, *out);
  return PBF_OK;
}

uint32_t pbf_host_morton_encode(uint32_t x, uint32_t y, uint32_t z) { return morton3(x, y, z); }
void pbf_host_morton_decode(uint32_t key, uint32_t xyz[3]) {
  xyz[0] = compact10(key);
  xyz[1] = compact10(key >> 1);
  xyz[2] = compact10(key >> 2);
}

// applyMotionSinXCosZ — sph.hpp:147-158: the box slides by (300 sin(f/20), 0, 90 cos(f/20)).  std::sin/cos on a
// float argument are the float overloads, the x offset is a float product, the z offset is a float product
// widened to double for the `* 0.3`, then narrowed.
void pbf_host_apply_motion(const pbf_params *in, uint64_t frame, pbf_params *out) {
  const float offset_scale = 300.f, offset_rate = 20.f;
  const float ox = float(std::sin(float(frame) / offset_rate) * offset_scale);
  const float oz = float(std::cos(float(frame) / offset_rate) * offset_scale * 0.3);
  *out = *in;
  out->min_bound[0] += ox; out->max_bound[0] += ox;
  out->min_bound[1] += 0.f; out->max_bound[1] += 0.f;
  out->min_bound[2] += oz; out->max_bound[2] += oz;
}

void pbf_host_constants(float h, float out[5]) {
  pbf_params p{};
  p.dt = 1.f; p.scale = 1.f;
  p.max_bound[0] = p.max_bound[1] = p.max_bound[2] = 1.f;
  pbf_grid_info g;
  host_grid(h, p, g);
  StepConst sc;
  host_step_const(h, p, g, 0, sc);
  out[0] = sc.P6; out[1] = sc.SP; out[2] = sc.P6dq; out[3] = sc.r2_max; out[4] = sc.r2_min;
}

}  // extern "C"
