/*
 * pbf_cuda.h — C ABI of the B200-native (sm_100a) Position-Based-Fluids step.
 *
 * This is the drop-in boundary for ONE path of UoB-HPC/pbf-sph: sph::Solver::advance
 * (reference src/sph.hpp:119-125), i.e. the per-step PBF solve that the reference implements in
 * src/omp/ompsph.hpp:85-485 (OpenMP), src/ocl/oclsph.cpp:315-513 (OpenCL) and src/sycl/syclsph.hpp.
 * The C++ adaptor in include/pbf/cudasph.hpp wraps these entry points behind the reference's own
 * `sph::Solver<size_t,float,V>` interface; INTEGRATION.md shows the `case Impl::CUDA` a maintainer adds
 * beside src/benchmark.cpp:151.
 *
 * Plain pointers and sizes only.  Every function returns PBF_OK (0) or a negative pbf_status; the
 * message is available from pbf_last_error().  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with PBF_ERR_CUDA.
 */
#ifndef PBF_CUDA_H
#define PBF_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBF_ABI_VERSION 2

typedef enum pbf_status {
  PBF_OK = 0,
  PBF_ERR_INVALID = -1,     /* bad argument (NULL, Obstacle particle, grid extent > 1023 cells, ...) */
  PBF_ERR_CUDA = -2,        /* CUDA runtime/driver error, or no device */
  PBF_ERR_NCCL = -3,        /* NCCL error */
  PBF_ERR_STATE = -4,       /* call order (e.g. step before upload) */
  PBF_ERR_CAPACITY = -5     /* a caller-provided buffer is too small */
} pbf_status;

/* sph::Type — src/sph.hpp:15.  Only Fluid is accepted (SURVEY note O: the OMP backend drops Obstacles). */
enum { PBF_TYPE_FLUID = 0, PBF_TYPE_OBSTACLE = 1 };

/* sph::Particle<size_t,float,glm::vec> — src/sph.hpp:36-54.  56 bytes, 8-aligned:
 * id@0 type@8 mass@12 position@16 velocity@28 colour@40.  This is what crosses advance(). */
typedef struct pbf_particle {
  uint64_t id;
  uint8_t type;
  uint8_t _pad[3];
  float mass;
  float position[3];
  float velocity[3];
  float colour[4];
} pbf_particle;

/* sph::McParams<float> — src/sph.hpp:82-95 */
typedef struct pbf_mc_params {
  float resolution;
  float isolevel;
  float particle_size;
  float particle_influence;
} pbf_mc_params;

/* sph::SphParams<size_t,float,V> — src/sph.hpp:97-103.  `h` is omitted on purpose: no reference solver
 * reads SphParams::h, the smoothing length is the solver's constructor argument (ompsph.hpp:80-83). */
typedef struct pbf_params {
  float dt;
  float scale;
  uint64_t iteration;
  float constant_force[3];
  float min_bound[3];
  float max_bound[3];
  int32_t wait;            /* SphParams::wait — "synchronise"; the host-vector contract always syncs */
  int32_t surface_enabled; /* std::optional<McParams>::has_value() */
  pbf_mc_params surface;
} pbf_params;

/* sph::Scene — src/sph.hpp:56-80: the per-call scene dynamics of advance().  Positions are in the caller's
 * (unscaled) units, like pbf_particle::position. */
typedef struct pbf_well {   /* sph::Well   sph.hpp:56-60: radial force on particles closer than 75 units (ompsph.hpp:141-148) */
  uint64_t tag;
  float centre[3];
  float force;
} pbf_well;
typedef struct pbf_source { /* sph::Source sph.hpp:62-67: emits a floor(sqrt(rate)) x ceil(sqrt(rate)) sheet per call (ompsph.hpp:93-104) */
  uint64_t tag;             /* becomes the id of every emitted particle */
  float centre[3];
  float velocity[3];
  float colour[4];
  float rate;
} pbf_source;
typedef struct pbf_drain {  /* sph::Drain  sph.hpp:69-73: removes fluid closer than `width` to the centre (ompsph.hpp:106-118) */
  uint64_t tag;
  float centre[3];
  float width, depth;       /* depth is unused by every reference backend */
} pbf_drain;
typedef struct pbf_query {  /* sph::Query  sph.hpp:22-25: ids of the fluid particles in the cell that holds `point` (ompsph.hpp:167-186) */
  uint64_t id;
  float point[3];
} pbf_query;
typedef struct pbf_scene {
  const pbf_well *wells;     uint32_t n_wells;
  const pbf_source *sources; uint32_t n_sources;
  const pbf_drain *drains;   uint32_t n_drains;
  const pbf_query *queries;  uint32_t n_queries;
} pbf_scene;
#define PBF_MAX_WELLS 16

/* Grid derived once per step — ompsph.hpp:132-135 and sph.hpp:238-241. */
typedef struct pbf_grid_info {
  float min_extent[3];
  uint32_t extent[3];      /* cells per axis */
  uint32_t grid_table_n;   /* G = morton(extent) = size of the cell table */
  uint32_t key_bits;       /* bit_length(G-1) */
  uint32_t radix_passes;   /* LSD passes actually run */
  uint64_t n_particles;
  uint32_t sample_size[3]; /* MC lattice points per axis (0 when surface disabled) */
  uint32_t n_triangles;    /* triangles emitted by the last step (0 when surface disabled) */
} pbf_grid_info;

/* Behaviour flags (pbf_set_flags).  Defaults: 0. */
enum {
  PBF_FLAG_STRICT_FP = 1u << 0,    /* no FMA contraction, IEEE div/sqrt in the solver kernels: follows the oracle op-for-op */
  PBF_FLAG_DEBUG_COUNTS = 1u << 1, /* also run the neighbour-count tap each step (PBF_TAP_CAND_COUNT/NBR_COUNT) */
  PBF_FLAG_PROFILE = 1u << 2,      /* record CUDA events around every kernel family (pbf_profile_read) */
  PBF_FLAG_GLOBAL_NEIGHBOURS = 1u << 3, /* use the one-pass global-memory neighbour kernels (A/B testing) */
  /* Extensions, default off: NO reference backend has them (sph_constants.h:13-14 only declares C and
   * VORTICITY_EPSILON; finalise is ompsph.hpp:256-264).  Applied to the velocities after finalise, defined in
   * csrc/xsph.cu and pinned by oracle/pbf_oracle.c; not available on the slab path. */
  PBF_FLAG_XSPH = 1u << 4,      /* v_i += C * sum_j (v_j - v_i) poly6(r_ij) */
  PBF_FLAG_VORTICITY = 1u << 5, /* v_i += dt * VORTICITY_EPSILON * (N x omega_i) */
  /* Drop-in path only.  The reference's callers hand advance() a std::vector, i.e. PAGEABLE memory, which the CUDA
   * runtime copies through its own staging buffers at a fraction of the PCIe rate.  With this flag pbf_advance_host /
   * pbf_advance_scene_host page-lock the caller's array the first time they see it (cudaHostRegister over
   * [xs, xs + capacity)) and keep it locked while the same array comes back, so the 56 MB each way of a 1 M-particle
   * step move by DMA.  CONTRACT: while the context holds an array the caller must not free or reallocate it — call
   * pbf_unpin_host first (the C++ adaptor does so before it resizes the vector, and pbf_destroy does it last).
   * Memory that cannot be registered (already pinned, or the driver refuses) is simply copied as before. */
  PBF_FLAG_PIN_HOST = 1u << 6
};

/* Debug taps, all in SORTED particle order unless stated (pbf_debug_read). */
typedef enum pbf_tap {
  PBF_TAP_KEYS_INPUT = 0,  /* u32[n]  Morton key per particle, INPUT order (ompsph.hpp:152) */
  PBF_TAP_PERM = 1,        /* u32[n]  sorted position -> input index (stable sort, ompsph.hpp:158) */
  PBF_TAP_KEYS_SORTED = 2, /* u32[n] */
  PBF_TAP_CELL_TABLE = 3,  /* u32[G]  sph.hpp:238-250 */
  PBF_TAP_CAND_COUNT = 4,  /* u32[n]  27-cell candidates incl. self, on the predicted positions */
  PBF_TAP_NBR_COUNT = 5,   /* u32[n]  candidates with r <= h incl. self */
  PBF_TAP_LAMBDA = 6,      /* f32[n]  lambda of the LAST solver iteration (ompsph.hpp:231) */
  PBF_TAP_RHO = 7,         /* f32[n]  density of the LAST solver iteration (ompsph.hpp:227) */
  PBF_TAP_IDS = 8,         /* u64[n] */
  PBF_TAP_MC_FIELD = 9,    /* f32[4*L] lattice (value, normal.xyz), index3d order (ompsph.hpp:350-354) */
  PBF_TAP_MC_COLOUR = 10,  /* f32[4*L] lattice colour (ompsph.hpp:355) */
  /* u32[n]  in-radius candidates incl. self as the PRODUCTION search counted them: the hit count of the neighbour list
   * the lambda pass of the FIRST solver iteration wrote (csrc/neighbour_list.cu), i.e. on the same predicted positions as
   * PBF_TAP_NBR_COUNT.  Needs PBF_FLAG_DEBUG_COUNTS; not available under PBF_FLAG_GLOBAL_NEIGHBOURS (no list). */
  PBF_TAP_LIST_HITS = 11
} pbf_tap;

/* Kernel families timed under PBF_FLAG_PROFILE.  ms are accumulated since the last pbf_profile_reset. */
enum {
  PBF_PH_PREDICT_KEY = 0,
  PBF_PH_SORT = 1,
  PBF_PH_REORDER = 2,
  PBF_PH_CELL_TABLE = 3,
  PBF_PH_DIFFUSE = 4,
  PBF_PH_LAMBDA = 5,
  PBF_PH_DELTA = 6,
  PBF_PH_FINALISE = 7,
  PBF_PH_MC_FIELD = 8,
  PBF_PH_MC_COUNT_SCAN = 9,
  PBF_PH_MC_EMIT = 10,
  PBF_PH_PACK = 11,
  PBF_PH_HALO = 12,            /* slab path: bookkeeping and exchange kernels (masks, counts, scans, pushes) */
  PBF_PH_SLAB_SETUP = 13,      /* slab path, a SPAN: step start -> first solver iteration (phases A-E with their barriers) */
  PBF_PH_SLAB_BARRIER = 14,    /* slab path: the cross-rank barriers (waiting for the slowest rank included) */
  PBF_PH_SLAB_ITERATIONS = 15, /* slab path, a SPAN: the solver iterations with their halo exchanges, up to finalise */
  PBF_PH_COUNT = 16
};
typedef struct pbf_profile {
  double ms[PBF_PH_COUNT];
  uint64_t launches[PBF_PH_COUNT];
  uint64_t steps;
} pbf_profile;

typedef struct pbf_ctx pbf_ctx;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Replaces the backend constructors the drivers call: omp_impl::Solver(h) benchmark.cpp:160-163,
 * ocl_impl::Solver(h, kernelPath, includes, device) oclsph.hpp:145-148.  `device` = CUDA ordinal. */
int pbf_create(pbf_ctx **out, float h, int device);
void pbf_destroy(pbf_ctx *ctx);
/* Message of the last failure on `ctx` (or of the last failed pbf_create when ctx == NULL). */
const char *pbf_last_error(const pbf_ctx *ctx);
int pbf_abi_version(void);
int pbf_set_flags(pbf_ctx *ctx, uint32_t flags);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream) instead of the context's own. */
int pbf_set_stream(pbf_ctx *ctx, void *cuda_stream);

/* Release the page-lock PBF_FLAG_PIN_HOST holds on the caller's particle array (no-op when none is held). */
int pbf_unpin_host(pbf_ctx *ctx);

/* ---- drop-in path: sph::Solver::advance (sph.hpp:122-124) ------------------------------------ */
/* H2D of xs, one PBF step, D2H.  On return xs[0..n) holds the advanced particles in Z-SORTED order
 * (ompsph.hpp:479-481).  *n_mesh_vertices (may be NULL) receives 3*triangles when params->surface_enabled;
 * fetch the mesh with pbf_mesh_download.  n == 0 is PBF_OK and does nothing (ompsph.hpp:122-126). */
int pbf_advance_host(pbf_ctx *ctx, const pbf_params *params, pbf_particle *xs, uint64_t n,
                     uint64_t *n_mesh_vertices);
/* The same with a Scene (sph.hpp:75-80), i.e. the whole of advance(config, scene, xs): sources append particles
 * (so xs needs `capacity` >= n + emitted), drains remove them, wells add their force to the prediction, queries are
 * answered from the step's cell table.  *n_out = particles after the call.  scene == NULL is the empty scene.
 * Returns PBF_ERR_CAPACITY (xs untouched) when the emitted particles do not fit. */
int pbf_advance_scene_host(pbf_ctx *ctx, const pbf_params *params, const pbf_scene *scene, pbf_particle *xs,
                           uint64_t n, uint64_t capacity, uint64_t *n_out, uint64_t *n_mesh_vertices);
/* sph::QueryResult::neighbours (sph.hpp:27-31) of query `index` of the last step's scene, in Z-sorted order.
 * *count receives the number of ids (also when ids == NULL or capacity is too small -> PBF_ERR_CAPACITY). */
int pbf_query_result(pbf_ctx *ctx, uint32_t index, uint64_t *ids, uint64_t capacity, uint64_t *count);
/* sph::ColouredMesh (sph.hpp:105-112): vs/ns = 3 floats per vertex, cs = 4 floats per vertex,
 * non-indexed, triangles ordered by marching-cube index.  Each pointer may be NULL to skip it. */
int pbf_mesh_download(pbf_ctx *ctx, float *vs, float *ns, float *cs, uint64_t capacity_vertices);

/* Device-side hand-off of the same mesh (for a renderer that maps CUDA memory, e.g. CUDA-GL interop in a `visualise`
 * driver, visualise.cpp:29-197): device pointers to vs/ns (3 floats per vertex) and cs (4 floats per vertex), valid
 * until the next step on this context.  The stream is synchronised before returning. */
int pbf_mesh_device(pbf_ctx *ctx, const float **vs, const float **ns, const float **cs, uint64_t *n_vertices);

/* ---- resident path (no host round trip between steps) ------------------------------------------ */
int pbf_upload(pbf_ctx *ctx, const pbf_particle *xs, uint64_t n);
int pbf_step(pbf_ctx *ctx, const pbf_params *params);          /* enqueue one step (asynchronous) */
/* Scene of the following pbf_step calls (copied; NULL = empty).  Every step applies it the way every advance()
 * call does: sources emit, drains remove (one host synchronisation per step while there are drains), wells pull,
 * queries are answered (pbf_query_result after pbf_sync). */
int pbf_set_scene(pbf_ctx *ctx, const pbf_scene *scene);
int pbf_sync(pbf_ctx *ctx);
int pbf_download(pbf_ctx *ctx, pbf_particle *xs, uint64_t capacity, uint64_t *n_out);
int pbf_particle_count(pbf_ctx *ctx, uint64_t *n_out);
/* Device pointers of the resident SoA state (float4 pos|mass, float4 vel, float4 colour, u64 id). */
int pbf_device_state(pbf_ctx *ctx, void **pos4, void **vel4, void **col4, void **ids);

/* ---- introspection for parity tests and the bench ------------------------------------------------ */
int pbf_grid(pbf_ctx *ctx, pbf_grid_info *out);
int pbf_debug_read(pbf_ctx *ctx, int tap, void *dst, uint64_t dst_bytes);
/* Parity tests only: depth of the per-iteration neighbour list (hits kept per particle; 192 by default, 96 is the other
 * compiled depth).  A particle with more hits than the depth takes the one-pass 27-cell walk in both solver passes; the
 * tests use the shallow list to drive that path on moderately dense clumps. */
int pbf_debug_set_list_capacity(pbf_ctx *ctx, uint32_t hits);
/* Parity tests only: the device exclusive prefix sum (csrc/sort_scan.cu; every compaction of the step — marching-cubes
 * triangle offsets ompsph.hpp:359-397, drains, the slab path's send lists — goes through it) and the stable radix sort of
 * (key, index) pairs over key bits [0, 30) (the std::sort of ompsph.hpp:158), on HOST arrays: in -> device -> out. */
int pbf_debug_scan_u32(pbf_ctx *ctx, const uint32_t *in, uint64_t n, uint32_t *out, uint32_t *total_out);
int pbf_debug_sort_pairs(pbf_ctx *ctx, const uint32_t *keys_in, uint32_t n, uint32_t *keys_out, uint32_t *perm_out);
int pbf_profile_reset(pbf_ctx *ctx);
int pbf_profile_read(pbf_ctx *ctx, pbf_profile *out);
/* Families timed under PBF_FLAG_PROFILE: bit f = PBF_PH_f (default: all).  Every timed launch costs two event records
 * on the stream (~30 per step with all families on, which stretches a 1.8 ms step by a fifth); bench.py times the
 * dominant family alone (8 records per step) for the roofline's launch duration. */
int pbf_profile_set_mask(pbf_ctx *ctx, uint32_t family_mask);
/* Number of kernel launches issued by this context since creation. */
uint64_t pbf_launch_count(const pbf_ctx *ctx);

/* ---- multi-GPU: Z-curve slab decomposition, one rank per GPU ----------------------------------------- */
/* The reference is single-device; this is the north-star extension.  Rank r owns the
 * Morton key range [split[r], split[r+1]); per step: migrants to their owners, then a 2-cell ghost layer whose
 * pStar is refreshed once per solver iteration.  Every exchange is a kernel that stores into the peer's memory (each
 * rank's arrays live in one arena that all peers map) followed by a flag barrier on the stream; all counts stay on the
 * device, so an ordinary step never waits for the host.  Two ways to form the group:
 *   NCCL   one process per GPU: rank 0 makes the id (pbf_dist_unique_id), the launcher distributes it
 *          (bench.py: torch.distributed.broadcast), every rank calls pbf_dist_init.  NCCL carries the bootstrap (CUDA IPC
 *          handles of the arenas) and the key-histogram all-reduce of the plan steps; the data moves by peer stores.
 *   LOCAL  all ranks are contexts of ONE process (any devices, may share a device): pbf_dist_init_local;
 *          peers are addressed directly, barriers are CUDA events.  pbf_dist_step on any member steps the whole group.
 * Marching cubes (params->surface_enabled): every rank evaluates the lattice points whose cell it owns — from its own
 * particles and ring-1 ghosts, after one more push of the ghosts' final positions and diffused colours — straight into
 * rank 0's lattice; rank 0 counts, scans and emits.  The mesh is the single-device mesh, bit for bit, and lives on rank 0
 * (pbf_mesh_download / pbf_mesh_device on rank 0's context; the other ranks report no triangles).
 * Scene dynamics (sources, drains, wells, queries) and the XSPH / vorticity extension are single-device only. */
#define PBF_NCCL_ID_BYTES 128
int pbf_dist_unique_id(uint8_t id[PBF_NCCL_ID_BYTES]);
int pbf_dist_init(pbf_ctx *ctx, const uint8_t id[PBF_NCCL_ID_BYTES], int rank, int world);
int pbf_dist_init_local(pbf_ctx **ctxs, int world);
/* This rank's share of the particles — ANY subset; the first step migrates every particle to its owner. */
int pbf_dist_upload(pbf_ctx *ctx, const pbf_particle *xs, uint64_t n);
/* Collective: one PBF step of the whole fluid (asynchronous on the context's stream between exchanges). */
int pbf_dist_step(pbf_ctx *ctx, const pbf_params *params);
/* Owned particles of this rank, Z-sorted (concatenating ranks 0..world-1 gives the global Z order). */
int pbf_dist_download(pbf_ctx *ctx, pbf_particle *xs, uint64_t capacity, uint64_t *n_out);
/* sph::Solver::advance (sph.hpp:119-125) on a LOCAL group — what sph::cuda_impl::Solver(h, {dev0, dev1, ...}) calls:
 * the caller's array is cut into one block per rank, uploaded, stepped by the whole group and returned in the global
 * Z-sorted order (ompsph.hpp:479-481), exactly as pbf_advance_host does on one device (same particles, same order,
 * bit-identical values).  `ctx` is any member of the group; PBF_FLAG_PIN_HOST is read from rank 0's flags. */
int pbf_dist_advance_host(pbf_ctx *ctx, const pbf_params *params, pbf_particle *xs, uint64_t n,
                          uint64_t *n_mesh_vertices);
/* Re-plan the key splits from a global key histogram every `steps` steps (default 4; 0 = only at the first step). */
int pbf_dist_set_replan(pbf_ctx *ctx, uint32_t steps);
typedef struct pbf_dist_stats {
  uint64_t owned, ghosts, migrants_out, migrants_in; /* of the last step */
  uint64_t halo_bytes_per_iteration;                 /* pStar bytes this rank SENDS per solver iteration */
  uint32_t key_lo, key_hi;                           /* owned Morton range */
  uint32_t ghost_ring1;                              /* ghosts whose lambda is computed locally */
  uint32_t boundary;                                 /* owned particles some other rank holds as ghosts */
  uint32_t plan_steps;                               /* plan steps so far (scheduled, after uploads, and early ones) */
  uint32_t early_plans;                              /* ... of which forced by a capacity watermark */
  uint32_t capacity_owned, capacity_ghosts;          /* particles the arena holds: owned block, each ghost block */
} pbf_dist_stats;
int pbf_dist_stats_read(pbf_ctx *ctx, pbf_dist_stats *out);

/* Pinned host memory for callers that want truly asynchronous H2D/D2H on the drop-in path. */
void *pbf_host_alloc(uint64_t bytes);
void pbf_host_free(void *p);

/* ---- host-only helpers (no GPU needed; used by the CPU tests of the host logic) ------------------ */
/* Grid set-up exactly as ompsph.hpp:132-135 / sph.hpp:240. */
int pbf_host_grid(float h, const pbf_params *params, pbf_grid_info *out);
/* Work weights of the load-balance histogram (host-only; pbf_dist_step applies it before pbf_host_plan_splits): a
 * particle's cost grows with the local density, because the neighbour search tests every particle of its 27 cells.
 * Fitted on one device (dam-1m: 1.89 us per particle-step at 6.3 particles per cell, 2.08 us at 7.4): weight of bucket b
 * = count[b] * (4.4 * 2^shift + count[b]), i.e. count * (4.4 + particles per cell) up to a constant factor; the LAST
 * bucket collects every key >= G (particles outside the grid, no density there) and is weighted count * 4.4 * 2^shift. */
int pbf_host_work_weights(const uint32_t *bucket_hist, uint32_t n_buckets, uint32_t shift, uint64_t *weights);
/* Split the key space [0,G) into `world` contiguous ranges of ~equal particle count from a histogram
 * over coarse key buckets (bucket b covers keys [b<<shift, (b+1)<<shift)).  splits has world+1 entries. */
int pbf_host_plan_splits(const uint64_t *bucket_hist, uint32_t n_buckets, uint32_t shift, int world,
                         uint32_t *splits);
/* Solver constants as the host side forms them: {poly6Factor, spikyKernelFactor, poly6(0.3h), r2_max, r2_min}
 * (sph.hpp:251-253, ompsph.hpp:211-213). */
void pbf_host_constants(float h, float out[5]);
/* applyMotionSinXCosZ — src/sph.hpp:147-158: the moving wall of the stock benchmark scene. */
void pbf_host_apply_motion(const pbf_params *in, uint64_t frame, pbf_params *out);
/* 10-bit-per-axis Morton encode/decode — src/curves.h:46-88. */
uint32_t pbf_host_morton_encode(uint32_t x, uint32_t y, uint32_t z);
void pbf_host_morton_decode(uint32_t key, uint32_t xyz[3]);

#ifdef __cplusplus
}
#endif

#ifdef __cplusplus
static_assert(sizeof(pbf_particle) == 56, "must match sizeof(sph::Particle<size_t,float,glm::vec>)");
#else
_Static_assert(sizeof(pbf_particle) == 56, "must match sizeof(sph::Particle<size_t,float,glm::vec>)");
#endif

#endif /* PBF_CUDA_H */
