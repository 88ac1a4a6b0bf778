/*
 * pbf_oracle.c — CPU restatement of the reference's per-step PBF solve.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (pbf_sph_b200/csrc -> libpbf_cuda.so) never links or calls it.
 *
 * What it restates: omp_impl::Solver<size_t,float>::advance, /root/reference/src/omp/ompsph.hpp:85-485
 * (with src/sph.hpp:198-253, src/curves.h:17-88, src/sph_constants.h:5-16, src/utils.hpp:73-85), in plain
 * C99, single precision, with every float expression written in the reference's evaluation order and
 * compiled WITHOUT contraction or fast-math (-O2 -ffp-contract=off), so results are bit-reproducible.
 *
 * Pinning (see tests/test_oracle_vs_reference.py, tests/golden/README.md): in PBF_ORACLE_GAUSS_SEIDEL
 * mode this file reproduces, bit for bit, the UNMODIFIED reference OpenMP backend run on one thread
 * (built by oracle/Makefile from the sources under /root/reference into oracle/_ref/); golden vectors
 * produced by that reference build are committed under tests/golden/.
 *
 * Two modes:
 *   PBF_ORACLE_GAUSS_SEIDEL  the reference as shipped when run on ONE thread: the delta pass overwrites
 *                            pStar in place while later particles read it (ompsph.hpp:235-248, read :239,
 *                            write :247) and diffuse overwrites colour in place (:189-206).  Serial.
 *   default (Jacobi)         the parity oracle for the GPU: the delta pass writes a second pStar buffer
 *                            (as a data-parallel device must), diffuse writes a second colour buffer (as
 *                            the reference's own OpenCL backend does, oclsph_kernel.h:67-93), the sort is
 *                            stable.  Order-independent across threads => OpenMP-parallel and
 *                            deterministic.  SURVEY.md F2/F3 explain why this is the only well-defined
 *                            parity target.
 */
#include "../include/pbf_cuda.h"
#include "../include/pbf/mc_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PBF_ORACLE_GAUSS_SEIDEL 1u /* in-place delta + in-place diffuse, serial (== reference on 1 thread) */
#define PBF_ORACLE_SKIP_DIFFUSE 2u
#define PBF_ORACLE_XSPH 4u      /* extension: XSPH viscosity after finalise (not in any reference backend) */
#define PBF_ORACLE_VORTICITY 8u /* extension: vorticity confinement after finalise */

/* src/sph_constants.h:5-16 */
static const float VD = 0.49f;
static const float RHO = 6378.0f;
#define RHO_RECIP (1.f / RHO)
static const float EPSILON = 0.00000001f;
static const float CFM_EPSILON = 600.0f;
static const float CorrDeltaQ = 0.3f;
static const float CorrK = 0.0001f;
static const float CorrN = 4.f;
static const float C_XSPH = 0.00001f;            /* sph_constants.h:13 `C` (declared, never used by the reference) */
static const float VORTICITY_EPSILON = 0.0005f;  /* sph_constants.h:14 (declared, never used by the reference) */

typedef struct pbf_oracle_io {
  /* optional taps; NULL = not wanted.  All sized by the caller. */
  uint32_t *keys_input;  /* [n] */
  uint32_t *perm;        /* [n] */
  uint32_t *keys_sorted; /* [n] */
  uint32_t *cell_table;  /* [cell_table_cap] */
  uint64_t cell_table_cap;
  uint32_t *cand_count;  /* [n] */
  uint32_t *nbr_count;   /* [n] */
  float *lambda;         /* [n] last iteration */
  float *rho;            /* [n] last iteration */
  float *mc_field;       /* [4*mc_lattice_cap] */
  float *mc_colour;      /* [4*mc_lattice_cap] */
  uint64_t mc_lattice_cap;
  float *mesh_vs, *mesh_ns, *mesh_cs;
  uint64_t mesh_cap_vertices;
  /* optional input: sorted position -> input index; overrides the oracle's own stable sort */
  const uint32_t *forced_perm;
  /* results */
  pbf_grid_info grid;
  uint64_t n_vertices;
  /* scene dynamics that act INSIDE the step: wells (ompsph.hpp:141-148) and queries (:167-186); NULL = empty scene.
   * Sources and drains edit the particle list before the step: pbf_oracle_scene_edit. */
  const pbf_scene *scene;
  uint32_t *query_first; /* [scene->n_queries] first sorted index of the query's cell */
  uint32_t *query_count; /* [scene->n_queries] particles in it (all Fluid: Obstacles are rejected) */
} pbf_oracle_io;

/* ---- curves.h:46-88 ------------------------------------------------------------------------------ */
static inline uint64_t spread10(uint64_t v) {
  v = (v | (v << 16)) & 0x030000FFull;
  v = (v | (v << 8)) & 0x0300F00Full;
  v = (v | (v << 4)) & 0x030C30C3ull;
  v = (v | (v << 2)) & 0x09249249ull;
  return v;
}
static inline uint64_t morton3(uint64_t x, uint64_t y, uint64_t z) { /* curves.h:72-88 */
  return spread10(x) | (spread10(y) << 1) | (spread10(z) << 2);
}
static inline uint64_t gather10(uint64_t v) { /* curves.h:46-59: bits 0,3,6,...,27 -> 0..9 */
  uint64_t r = 0;
  for (int b = 0; b < 10; ++b) r |= (v & (1ull << (3 * b))) >> (2 * b);
  return r;
}
static inline uint64_t demorton(uint64_t key, int axis) { /* curves.h:61-65 */
  return gather10((key >> axis) & 0x9249249ull);
}

uint32_t pbf_oracle_morton_encode(uint32_t x, uint32_t y, uint32_t z) { return (uint32_t)morton3(x, y, z); }
void pbf_oracle_morton_decode(uint32_t key, uint32_t xyz[3]) {
  for (int a = 0; a < 3; ++a) xyz[a] = (uint32_t)demorton(key, a);
}

/* float -> size_t as x86-64 does it for the values that occur (cvttss2si): sph.hpp:199-200 */
static inline uint64_t to_index(float v) { return (uint64_t)(int64_t)v; }

/* sph.hpp:198-201 */
static inline uint64_t key_at(float x, float y, float z, float h) {
  return morton3(to_index(x / h), to_index(y / h), to_index(z / h));
}

/* sph.hpp:251-253 — std::pow(float,int) promotes to double, so the factors are formed in double and
 * rounded to float on return. */
static float pi_f(void) { return acosf(-1.0f); }
static float poly6_factor(float h) { return (float)((double)315.0f / ((double)(64.0f * pi_f()) * pow((double)h, 9.0))); }
static float spiky_factor(float h) { return (float)(-((double)45.0f / ((double)pi_f() * pow((double)h, 6.0)))); }

/* ompsph.hpp:67-69 */
static inline float poly6(float r, float factor, float h) {
  if (r <= h) {
    const float d = (h * h) - r * r;
    return factor * (d * d * d);
  }
  return 0.f;
}
/* ompsph.hpp:71-75: scalar part of the spiky gradient; gradient = (x - y) * s */
static inline int spiky_active(float r, float h) { return r >= EPSILON && r <= h; }
static inline float spiky_scalar(float r, float h, float factor) { return factor * (((h - r) * (h - r)) / r); }

static inline float dist3(const float *a, const float *b) { /* glm::distance = length(b-a), dot = (x*x+y*y)+z*z */
  const float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  return sqrtf((dx * dx + dy * dy) + dz * dz);
}
static inline float fmin_glm(float a, float b) { return (b < a) ? b : a; }
static inline float fmax_glm(float a, float b) { return (a < b) ? b : a; }
static inline float fast_sqrt_glm(float x) { return 1.0f / (1.0f / sqrtf(x)); } /* gtx/fast_square_root highp */

/* 27 neighbour keys in the reference's order — sph.hpp:215-236: x fastest, then y, then z, each -1,0,+1;
 * x-1 at 0 wraps (size_t underflow) and is masked by the Morton spread to 1023. */
static inline void neighbour_keys(uint64_t key, uint64_t out[27]) {
  const uint64_t x = demorton(key, 0), y = demorton(key, 1), z = demorton(key, 2);
  int t = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) out[t++] = morton3(x + (uint64_t)(int64_t)dx, y + (uint64_t)(int64_t)dy, z + (uint64_t)(int64_t)dz);
}

/* cell range per sph.hpp:203-213: offsets >= G skipped; the last cell G-1 is always empty */
static inline void cell_range(const uint32_t *table, uint64_t G, uint64_t o, uint32_t *start, uint32_t *end) {
  if (o >= G) { *start = *end = 0; return; }
  *start = table[o];
  *end = (o + 1) < G ? table[o + 1] : *start;
}

static int cmp_u64(const void *a, const void *b) {
  const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return (x > y) - (x < y);
}

int pbf_oracle_grid(float h, const pbf_params *p, pbf_grid_info *g) {
  /* ompsph.hpp:132-135 */
  const float padding = h * 2;
  memset(g, 0, sizeof(*g));
  for (int a = 0; a < 3; ++a) {
    const float mn = (p->min_bound[a] / p->scale) - padding;
    const float mx = (p->max_bound[a] / p->scale) + padding;
    g->min_extent[a] = mn;
    g->extent[a] = (uint32_t)to_index((mx - mn) / h);
  }
  g->grid_table_n = (uint32_t)morton3(g->extent[0], g->extent[1], g->extent[2]); /* sph.hpp:240 */
  uint32_t bits = 0;
  if (g->grid_table_n > 1) { uint32_t v = g->grid_table_n - 1; while (v) { ++bits; v >>= 1; } }
  g->key_bits = bits;
  g->radix_passes = (bits + 7) / 8;
  if (p->surface_enabled) /* ompsph.hpp:283-284 */
    for (int a = 0; a < 3; ++a) g->sample_size[a] = (uint32_t)to_index(floorf((float)g->extent[a] * p->surface.resolution)) + 1u;
  return 0;
}

int pbf_oracle_step(float h, const pbf_params *p, pbf_particle *xs, uint64_t n, uint32_t mode, pbf_oracle_io *io) {
  pbf_oracle_io local;
  if (!io) { memset(&local, 0, sizeof(local)); io = &local; }
  const int gs = (mode & PBF_ORACLE_GAUSS_SEIDEL) != 0;
  pbf_oracle_grid(h, p, &io->grid);
  io->grid.n_particles = n;
  io->n_vertices = 0;
  if (n == 0) return 0; /* ompsph.hpp:122-126 */
  for (uint64_t i = 0; i < n; ++i)
    if (xs[i].type != PBF_TYPE_FLUID) return PBF_ERR_INVALID;

  const float scale = p->scale, dt = p->dt;
  const float *minE = io->grid.min_extent;
  const uint64_t G = io->grid.grid_table_n;
  const uint32_t *ext = io->grid.extent;

  /* ---- advect + key: ompsph.hpp:137-154 ---- */
  const uint32_t n_wells = io->scene ? io->scene->n_wells : 0u;
  float *vel_in = malloc(sizeof(float) * 3 * n), *pstar_in = malloc(sizeof(float) * 3 * n);
  uint64_t *composite = malloc(sizeof(uint64_t) * n);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    const pbf_particle *q = &xs[i];
    float key_arg[3], combined[3];
    for (int a = 0; a < 3; ++a) combined[a] = q->mass * p->constant_force[a];
    for (uint32_t w = 0; w < n_wells; ++w) { /* ompsph.hpp:141-148 */
      const pbf_well *well = &io->scene->wells[w];
      const float dist = dist3(q->position, well->centre); /* glm::distance(p.position, well.centre) */
      if (dist < 75.0f) {
        float d[3];
        for (int a = 0; a < 3; ++a) d[a] = well->centre[a] - q->position[a];
        /* glm::normalize(v) = v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x) */
        const float inv = 1.0f / sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
        for (int a = 0; a < 3; ++a) {
          const float fw = (((d[a] * inv) * well->force) * q->mass) / (dist * dist);
          combined[a] += fmin_glm(fmax_glm(fw, -10.0f), 10.0f); /* glm::clamp = min(max(x, lo), hi) */
        }
      }
    }
    for (int a = 0; a < 3; ++a) {
      const float force = combined[a];
      const float v = force * dt + q->velocity[a];
      const float ps = (v * dt) + (q->position[a] / scale);
      vel_in[3 * i + a] = v;
      pstar_in[3 * i + a] = ps;
      key_arg[a] = ps - minE[a];
    }
    const uint64_t key = key_at(key_arg[0], key_arg[1], key_arg[2], h);
    composite[i] = (key << 32) | (uint64_t)i;
    if (io->keys_input) io->keys_input[i] = (uint32_t)key;
  }

  /* ---- sort by key: ompsph.hpp:158.  Stable (ties keep input order) unless a permutation is forced. ---- */
  uint32_t *perm = malloc(sizeof(uint32_t) * n);
  if (io->forced_perm) {
    memcpy(perm, io->forced_perm, sizeof(uint32_t) * n);
  } else {
    qsort(composite, n, sizeof(uint64_t), cmp_u64);
    for (uint64_t i = 0; i < n; ++i) perm[i] = (uint32_t)(composite[i] & 0xFFFFFFFFu);
  }
  /* sorted SoA work set (the reference sorts 96-byte AoS records, sph.hpp:255-261) */
  uint32_t *key = malloc(sizeof(uint32_t) * n);
  float *pos = malloc(sizeof(float) * 3 * n), *vel = malloc(sizeof(float) * 3 * n), *col = malloc(sizeof(float) * 4 * n);
  float *pstar = malloc(sizeof(float) * 3 * n), *pstar2 = malloc(sizeof(float) * 3 * n), *col2 = malloc(sizeof(float) * 4 * n);
  float *mass = malloc(sizeof(float) * n), *lambda = malloc(sizeof(float) * n), *rho_out = malloc(sizeof(float) * n);
  uint64_t *id = malloc(sizeof(uint64_t) * n);
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t s = perm[i];
    const pbf_particle *q = &xs[s];
    float ka[3];
    for (int a = 0; a < 3; ++a) {
      pos[3 * i + a] = q->position[a];
      vel[3 * i + a] = vel_in[3 * s + a];
      pstar[3 * i + a] = pstar_in[3 * s + a];
      ka[a] = pstar[3 * i + a] - minE[a];
    }
    for (int a = 0; a < 4; ++a) col[4 * i + a] = q->colour[a];
    mass[i] = q->mass;
    id[i] = q->id;
    key[i] = (uint32_t)key_at(ka[0], ka[1], ka[2], h);
    lambda[i] = 0.f;
    rho_out[i] = 0.f;
  }
  free(vel_in); free(pstar_in); free(composite);
  if (io->perm) memcpy(io->perm, perm, sizeof(uint32_t) * n);
  if (io->keys_sorted) memcpy(io->keys_sorted, key, sizeof(uint32_t) * n);

  /* ---- cell table: sph.hpp:238-250 ---- */
  uint32_t *table = malloc(sizeof(uint32_t) * (G ? G : 1));
  {
    uint64_t gi = 0;
    for (uint64_t z = 0; z < G; ++z) {
      table[z] = (uint32_t)gi;
      while (gi != n && key[gi] == z) gi++;
    }
  }
  if (io->cell_table) {
    if (io->cell_table_cap < G) return PBF_ERR_CAPACITY;
    memcpy(io->cell_table, table, sizeof(uint32_t) * G);
  }

  /* ---- queries: ompsph.hpp:167-186 (one cell per query; the neighbours are the ids of that sorted range) ---- */
  if (io->scene && io->scene->n_queries && io->query_first && io->query_count)
    for (uint32_t qi = 0; qi < io->scene->n_queries; ++qi) {
      const pbf_query *qq = &io->scene->queries[qi];
      float r[3];
      for (int a = 0; a < 3; ++a) r[a] = (qq->point[a] / scale) - minE[a];
      const uint64_t z = key_at(r[0], r[1], r[2], h);
      io->query_first[qi] = 0;
      io->query_count[qi] = 0;
      if (z < G && z + 1 < G) {
        io->query_first[qi] = table[z];
        io->query_count[qi] = table[z + 1] - table[z];
      }
    }

  /* ---- tap: candidate / in-radius counts on the predicted positions (what the first lambda pass sees) ---- */
  if (io->cand_count || io->nbr_count) {
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t a = 0; a < (int64_t)n; ++a) {
      uint64_t nk[27];
      neighbour_keys(key[a], nk);
      uint32_t cand = 0, nbr = 0;
      for (int c = 0; c < 27; ++c) {
        uint32_t s, e;
        cell_range(table, G, nk[c], &s, &e);
        for (uint32_t b = s; b < e; ++b) {
          ++cand;
          if (dist3(&pstar[3 * a], &pstar[3 * b]) <= h) ++nbr;
        }
      }
      if (io->cand_count) io->cand_count[a] = cand;
      if (io->nbr_count) io->nbr_count[a] = nbr;
    }
  }

  /* ---- diffuse: ompsph.hpp:189-206 ---- */
  if (!(mode & PBF_ORACLE_SKIP_DIFFUSE)) {
    float *dst = gs ? col : col2;
    const float mixf = dt / 750.0f;
#pragma omp parallel for schedule(dynamic, 256) if (!gs)
    for (int64_t a = 0; a < (int64_t)n; ++a) {
      uint64_t nk[27];
      neighbour_keys(key[a], nk);
      int nn = 0;
      float mx[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < 27; ++c) {
        uint32_t s, e;
        cell_range(table, G, nk[c], &s, &e);
        for (uint32_t b = s; b < e; ++b) {
          for (int k = 0; k < 4; ++k) mx[k] += col[4 * b + k];
          nn++;
        }
      }
      for (int k = 0; k < 4; ++k) {
        float out = col[4 * a + k];
        if (nn != 0) {
          /* glm::mix(x, y, t) = x*(1-t) + y*t with y = (mixture / n) * 1.33 */
          const float y = (mx[k] / (float)nn) * 1.33f;
          out = col[4 * a + k] * (1.0f - mixf) + y * mixf;
          out = fmin_glm(fmax_glm(out, 0.03f), 1.0f);
        }
        dst[4 * a + k] = out;
      }
    }
    if (!gs) { float *t = col; col = col2; col2 = t; }
  }

  /* ---- solver iterations: ompsph.hpp:211-249 ---- */
  const float P6 = poly6_factor(h), SP = spiky_factor(h);
  const float P6dq = poly6(CorrDeltaQ * h, P6, h);
  for (uint64_t itr = 0; itr < p->iteration; ++itr) {
    /* lambda: ompsph.hpp:217-232 */
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t a = 0; a < (int64_t)n; ++a) {
      uint64_t nk[27];
      neighbour_keys(key[a], nk);
      float g[3] = {0.f, 0.f, 0.f};
      float rho = 0.f;
      const float *pa = &pstar[3 * a];
      for (int c = 0; c < 27; ++c) {
        uint32_t s, e;
        cell_range(table, G, nk[c], &s, &e);
        for (uint32_t b = s; b < e; ++b) {
          const float *pb = &pstar[3 * b];
          const float r = dist3(pa, pb);
          if (spiky_active(r, h)) {
            const float sc = spiky_scalar(r, h, SP);
            for (int k = 0; k < 3; ++k) g[k] += ((pa[k] - pb[k]) * sc) * RHO_RECIP;
          }
          rho += mass[a] * poly6(r, P6, h);
        }
      }
      const float norm2 = (g[0] * g[0] + g[1] * g[1]) + g[2] * g[2];
      const float Ci = (rho / RHO - 1.0f);
      lambda[a] = -Ci / (norm2 + CFM_EPSILON);
      rho_out[a] = rho;
    }
    /* delta + clamp: ompsph.hpp:235-248 */
    float *dst = gs ? pstar : pstar2;
#pragma omp parallel for schedule(dynamic, 256) if (!gs)
    for (int64_t a = 0; a < (int64_t)n; ++a) {
      uint64_t nk[27];
      neighbour_keys(key[a], nk);
      float d[3] = {0.f, 0.f, 0.f};
      const float *pa = &pstar[3 * a];
      for (int c = 0; c < 27; ++c) {
        uint32_t s, e;
        cell_range(table, G, nk[c], &s, &e);
        for (uint32_t b = s; b < e; ++b) {
          const float *pb = &pstar[3 * b];
          const float r = dist3(pa, pb);
          const float corr = (-CorrK) * powf(poly6(r, P6, h) / P6dq, CorrN);
          const float factor = (lambda[a] + lambda[b] + corr) / RHO;
          if (spiky_active(r, h)) {
            const float sc = spiky_scalar(r, h, SP);
            for (int k = 0; k < 3; ++k) d[k] += ((pa[k] - pb[k]) * sc) * factor;
          }
        }
      }
      for (int k = 0; k < 3; ++k) {
        float ps = (pa[k] + d[k]) * scale;
        ps = fmin_glm(p->max_bound[k], fmax_glm(p->min_bound[k], ps));
        dst[3 * a + k] = ps / scale;
      }
    }
    if (!gs) { float *t = pstar; pstar = pstar2; pstar2 = t; }
  }
  if (io->lambda) memcpy(io->lambda, lambda, sizeof(float) * n);
  if (io->rho) memcpy(io->rho, rho_out, sizeof(float) * n);

  /* ---- finalise: ompsph.hpp:256-264 ---- */
  const float inv_dt = 1.0f / dt;
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < (int64_t)n; ++a)
    for (int k = 0; k < 3; ++k) {
      const float dx = pstar[3 * a + k] - pos[3 * a + k] / scale;
      pos[3 * a + k] = pstar[3 * a + k] * scale;
      vel[3 * a + k] = (dx * inv_dt + vel[3 * a + k]) * VD;
    }

  /* ---- extension (opt-in; the reference declares the constants but has neither term) ----------------
   * Macklin & Mueller 2013 §5 on the step's final state: positions pStar (scaled units), the velocities finalise has
   * just produced, the step's cell table, W = poly6Kernel and grad W = spikyKernelGradient (ompsph.hpp:67-75), sums
   * in the 27-cell visiting order, Jacobi (all sums read the post-finalise velocities):
   *   omega_i = (1/RHO) sum_j (v_j - v_i) x grad_{p_j} W(p_i - p_j),   grad_{p_j} W = -spiky(p_i, p_j)
   *             (1/RHO = the particle volume m/rho_0 at unit mass: the SPH curl estimate)
   *   eta_i   = sum_j |omega_j| spiky(p_i, p_j),  N = eta / |eta|  (no force when |eta| < EPSILON)
   *   v_i    += dt * VORTICITY_EPSILON * (N x omega_i)  +  C * sum_j (v_j - v_i) poly6(|p_i - p_j|) */
  if (mode & (PBF_ORACLE_XSPH | PBF_ORACLE_VORTICITY)) {
    const int do_x = (mode & PBF_ORACLE_XSPH) != 0, do_v = (mode & PBF_ORACLE_VORTICITY) != 0;
    float *omega = calloc(4 * n, sizeof(float)), *vnew = malloc(sizeof(float) * 3 * n);
    if (do_v) {
#pragma omp parallel for schedule(dynamic, 256)
      for (int64_t a = 0; a < (int64_t)n; ++a) {
        uint64_t nk[27];
        neighbour_keys(key[a], nk);
        float w[3] = {0.f, 0.f, 0.f};
        const float *pa = &pstar[3 * a], *va = &vel[3 * a];
        for (int c = 0; c < 27; ++c) {
          uint32_t s0, e0;
          cell_range(table, G, nk[c], &s0, &e0);
          for (uint32_t b = s0; b < e0; ++b) {
            const float *pb = &pstar[3 * b];
            const float r = dist3(pa, pb);
            if (!spiky_active(r, h)) continue;
            const float sc = spiky_scalar(r, h, SP);
            const float g[3] = {-((pa[0] - pb[0]) * sc), -((pa[1] - pb[1]) * sc), -((pa[2] - pb[2]) * sc)};
            const float d[3] = {vel[3 * b] - va[0], vel[3 * b + 1] - va[1], vel[3 * b + 2] - va[2]};
            w[0] += d[1] * g[2] - d[2] * g[1];
            w[1] += d[2] * g[0] - d[0] * g[2];
            w[2] += d[0] * g[1] - d[1] * g[0];
          }
        }
        for (int k = 0; k < 3; ++k) w[k] *= RHO_RECIP;
        omega[4 * a] = w[0]; omega[4 * a + 1] = w[1]; omega[4 * a + 2] = w[2];
        omega[4 * a + 3] = sqrtf((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
      }
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t a = 0; a < (int64_t)n; ++a) {
      uint64_t nk[27];
      neighbour_keys(key[a], nk);
      float eta[3] = {0.f, 0.f, 0.f}, xs_sum[3] = {0.f, 0.f, 0.f};
      const float *pa = &pstar[3 * a], *va = &vel[3 * a];
      for (int c = 0; c < 27; ++c) {
        uint32_t s0, e0;
        cell_range(table, G, nk[c], &s0, &e0);
        for (uint32_t b = s0; b < e0; ++b) {
          const float *pb = &pstar[3 * b];
          const float r = dist3(pa, pb);
          if (r > h) continue;
          if (do_x) {
            const float wgt = poly6(r, P6, h);
            for (int k = 0; k < 3; ++k) xs_sum[k] += (vel[3 * b + k] - va[k]) * wgt;
          }
          if (do_v && r >= EPSILON) {
            const float sc = spiky_scalar(r, h, SP) * omega[4 * b + 3];
            for (int k = 0; k < 3; ++k) eta[k] += (pa[k] - pb[k]) * sc;
          }
        }
      }
      float out[3] = {va[0], va[1], va[2]};
      if (do_v) {
        const float len = sqrtf((eta[0] * eta[0] + eta[1] * eta[1]) + eta[2] * eta[2]);
        if (len >= EPSILON) {
          const float N[3] = {eta[0] / len, eta[1] / len, eta[2] / len};
          const float *w = &omega[4 * a];
          const float f[3] = {N[1] * w[2] - N[2] * w[1], N[2] * w[0] - N[0] * w[2], N[0] * w[1] - N[1] * w[0]};
          for (int k = 0; k < 3; ++k) out[k] += (dt * VORTICITY_EPSILON) * f[k];
        }
      }
      if (do_x)
        for (int k = 0; k < 3; ++k) out[k] += C_XSPH * xs_sum[k];
      for (int k = 0; k < 3; ++k) vnew[3 * a + k] = out[k];
    }
    memcpy(vel, vnew, sizeof(float) * 3 * n);
    free(omega); free(vnew);
  }

  /* ---- marching cubes: ompsph.hpp:277-477 ---- */
  if (p->surface_enabled) {
    static const uint64_t TRI[256] = PBF_MC_TRI_WORDS_INIT;
    static const int CORNER[8][3] = PBF_MC_CORNER_OFFSETS_INIT;
    static const int EDGE[12][2] = PBF_MC_EDGE_CORNERS_INIT;
    const float res = p->surface.resolution, iso = p->surface.isolevel;
    const float psize = p->surface.particle_size, pinf = p->surface.particle_influence;
    const uint64_t sx = io->grid.sample_size[0], sy = io->grid.sample_size[1], sz = io->grid.sample_size[2];
    const uint64_t L = sx * sy * sz;
    float *PN = calloc(4 * L, sizeof(float)), *LC = calloc(4 * L, sizeof(float));
    const float step = h / res;
    const float threshold = h * scale * 1;
    /* field: ompsph.hpp:288-356 */
#pragma omp parallel for collapse(2) schedule(dynamic, 16)
    for (int64_t x = 0; x < (int64_t)sx; ++x)
      for (int64_t y = 0; y < (int64_t)sy; ++y)
        for (int64_t z = 0; z < (int64_t)sz; ++z) {
          const float lp[3] = {(float)x, (float)y, (float)z};
          float a[3];
          for (int k = 0; k < 3; ++k) a[k] = (minE[k] + (lp[k] * step)) * scale;
          const uint64_t zi = morton3(to_index(lp[0] / res), to_index(lp[1] / res), to_index(lp[2] / res));
          const int64_t c0[3] = {(int64_t)demorton(zi, 0), (int64_t)demorton(zi, 1), (int64_t)demorton(zi, 2)};
          if ((uint64_t)c0[0] == ext[0] && (uint64_t)c0[1] == ext[1] && (uint64_t)c0[2] == ext[2]) continue;
          int64_t lo[3], hi[3];
          for (int k = 0; k < 3; ++k) { /* glm::clamp(int) — ompsph.hpp:306-311 */
            const int64_t top = (int64_t)ext[k] - 1;
            int64_t l = c0[k] - 1, r = c0[k] + 1;
            l = l < 0 ? 0 : l; l = l > top ? top : l;   /* min(max(x,0),top) */
            r = r < 0 ? 0 : r; r = r > top ? top : r;
            lo[k] = l; hi[k] = r;
          }
          const int64_t cx[3] = {lo[0], c0[0], hi[0]}, cy[3] = {lo[1], c0[1], hi[1]}, cz[3] = {lo[2], c0[2], hi[2]};
          float v = 0.f, nrm[3] = {0.f, 0.f, 0.f}, cc[4] = {0.f, 0.f, 0.f, 0.f};
          uint64_t nn = 0;
          for (int kz = 0; kz < 3; ++kz)
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx) { /* order of ompsph.hpp:313-326 */
                const uint64_t o = morton3((uint64_t)cx[kx], (uint64_t)cy[ky], (uint64_t)cz[kz]);
                uint32_t s, e;
                cell_range(table, G, o, &s, &e);
                for (uint32_t b = s; b < e; ++b) {
                  const float *pb = &pos[3 * b];
                  /* fastDistance(pb, a) = fastLength(a - pb) */
                  const float ex = a[0] - pb[0], ey = a[1] - pb[1], ez = a[2] - pb[2];
                  if (fast_sqrt_glm((ex * ex + ey * ey) + ez * ez) < threshold) {
                    const float l[3] = {pb[0] - a[0], pb[1] - a[1], pb[2] - a[2]};
                    const float len = fast_sqrt_glm((l[0] * l[0] + l[1] * l[1]) + l[2] * l[2]);
                    const float den = powf(len, pinf);
                    v += (psize / den);
                    const float w = (-pinf) * psize;
                    for (int k = 0; k < 3; ++k) nrm[k] += w * (l[k] / den);
                    for (int k = 0; k < 4; ++k) cc[k] += col[4 * b + k];
                    nn++;
                  }
                }
              }
          { /* fastNormalize = v * (1/sqrt(dot)) */
            const float inv = 1.0f / sqrtf((nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2]);
            for (int k = 0; k < 3; ++k) nrm[k] = nrm[k] * inv;
          }
          const uint64_t idx = (uint64_t)x * sy * sz + (uint64_t)y * sz + (uint64_t)z; /* curves.h:17-19 */
          PN[4 * idx + 0] = v; PN[4 * idx + 1] = nrm[0]; PN[4 * idx + 2] = nrm[1]; PN[4 * idx + 3] = nrm[2];
          for (int k = 0; k < 4; ++k) LC[4 * idx + k] = cc[k] / (float)nn;
        }
    if (io->mc_field || io->mc_colour) {
      if (io->mc_lattice_cap < L) return PBF_ERR_CAPACITY;
      if (io->mc_field) memcpy(io->mc_field, PN, sizeof(float) * 4 * L);
      if (io->mc_colour) memcpy(io->mc_colour, LC, sizeof(float) * 4 * L);
    }
    /* count + emit, in cube-index order: ompsph.hpp:365-474 (the reference's atomic slot counter makes its
     * order thread-dependent; one thread visits cubes in index order, which is what is restated here) */
    const uint64_t mx = sx - 1, my = sy - 1, mz = sz - 1, MV = mx * my * mz;
    uint64_t ntri = 0;
    for (int pass = 0; pass < 2; ++pass) {
      uint64_t slot = 0;
      for (uint64_t i = 0; i < MV; ++i) {
        const uint64_t cxi = i / (my * mz), cyi = (i - cxi * my * mz) / mz, czi = i - cxi * my * mz - cyi * mz; /* utils.hpp:73-79 */
        float val[8];
        uint64_t li[8];
        uint32_t ci = 0;
        for (int c = 0; c < 8; ++c) {
          li[c] = (cxi + CORNER[c][0]) * sy * sz + (cyi + CORNER[c][1]) * sz + (czi + CORNER[c][2]);
          val[c] = PN[4 * li[c]];
          if (val[c] < iso) ci |= 1u << c;
        }
        const uint64_t row = TRI[ci];
        const uint32_t emask = pbf_mc_edge_mask(row);
        const uint32_t nv = emask == 0 ? 0u : pbf_mc_num_verts(row);
        if (pass == 0) { ntri += nv / 3; continue; }
        if (nv == 0) continue;
        float ts[12][3], ns[12][3], cs[12][4];
        for (int e = 0; e < 12; ++e) {
          if (!(emask & (1u << e))) continue;
          const int from = EDGE[e][0], to = EDGE[e][1];
          const float t = (iso - val[from]) / (val[to] - val[from]); /* utils.hpp:85 */
          float of[3], ot[3];
          const float cf[3] = {(float)(cxi + CORNER[from][0]), (float)(cyi + CORNER[from][1]), (float)(czi + CORNER[from][2])};
          const float ct[3] = {(float)(cxi + CORNER[to][0]), (float)(cyi + CORNER[to][1]), (float)(czi + CORNER[to][2])};
          for (int k = 0; k < 3; ++k) {
            of[k] = (minE[k] + (cf[k] * step)) * scale; /* ompsph.hpp:424 */
            ot[k] = (minE[k] + (ct[k] * step)) * scale;
            ts[e][k] = of[k] * (1.0f - t) + ot[k] * t;
            ns[e][k] = PN[4 * li[from] + 1 + k] * (1.0f - t) + PN[4 * li[to] + 1 + k] * t;
          }
          for (int k = 0; k < 4; ++k) cs[e][k] = LC[4 * li[from] + k] * (1.0f - t) + LC[4 * li[to] + k] * t;
        }
        for (uint32_t t = 0; t < nv; ++t) {
          const uint32_t e = (uint32_t)((row >> (4 * t)) & 0xF);
          const uint64_t vtx = slot * 3 + (t % 3);
          if (vtx < io->mesh_cap_vertices) {
            if (io->mesh_vs) memcpy(&io->mesh_vs[3 * vtx], ts[e], sizeof(float) * 3);
            if (io->mesh_ns) memcpy(&io->mesh_ns[3 * vtx], ns[e], sizeof(float) * 3);
            if (io->mesh_cs) memcpy(&io->mesh_cs[4 * vtx], cs[e], sizeof(float) * 4);
          }
          if (t % 3 == 2) slot++;
        }
      }
    }
    io->grid.n_triangles = (uint32_t)ntri;
    io->n_vertices = ntri * 3;
    free(PN); free(LC);
  }

  /* ---- write back in sorted order: ompsph.hpp:479-481 ---- */
  for (uint64_t i = 0; i < n; ++i) {
    pbf_particle *q = &xs[i];
    memset(q, 0, sizeof(*q));
    q->id = id[i];
    q->type = PBF_TYPE_FLUID;
    q->mass = mass[i];
    for (int a = 0; a < 3; ++a) { q->position[a] = pos[3 * i + a]; q->velocity[a] = vel[3 * i + a]; }
    for (int a = 0; a < 4; ++a) q->colour[a] = col[4 * i + a];
  }
  free(perm); free(key); free(pos); free(vel); free(col); free(pstar); free(pstar2); free(col2);
  free(mass); free(lambda); free(rho_out); free(id); free(table);
  return 0;
}

/* Solver constants as the oracle forms them (so tests can compare the host side of the product bit for bit). */
/* Sources then drains on the caller's particle list — ompsph.hpp:91-120.  Returns the new count, or a negative
 * pbf_status when the emitted particles do not fit `cap`. */
int64_t pbf_oracle_scene_edit(float h, const pbf_params *p, const pbf_scene *scene, pbf_particle *xs, uint64_t n,
                              uint64_t cap) {
  if (!scene) return (int64_t)n;
  const float spacing = (h * p->scale / 2);
  for (uint32_t si = 0; si < scene->n_sources; ++si) {
    const pbf_source *src = &scene->sources[si];
    const float size = sqrtf(src->rate);
    const uint64_t width = (uint64_t)floorf(size), depth = (uint64_t)ceilf(size);
    /* offset = centre - (V3(width, 0, depth) * 0.5 * spacing) */
    const float off[3] = {src->centre[0] - (((float)width * 0.5f) * spacing), src->centre[1] - ((0.0f * 0.5f) * spacing),
                          src->centre[2] - (((float)depth * 0.5f) * spacing)};
    for (uint64_t x = 0; x < width; ++x)
      for (uint64_t z = 0; z < depth; ++z) {
        if (n >= cap) return PBF_ERR_CAPACITY;
        pbf_particle *q = &xs[n++];
        memset(q, 0, sizeof(*q));
        q->id = src->tag;
        q->type = PBF_TYPE_FLUID;
        q->mass = 1.0f;
        q->position[0] = off[0] + ((float)x * spacing);
        q->position[1] = off[1] + (0.0f * spacing);
        q->position[2] = off[2] + ((float)z * spacing);
        for (int a = 0; a < 3; ++a) q->velocity[a] = src->velocity[a];
        for (int a = 0; a < 4; ++a) q->colour[a] = src->colour[a];
      }
  }
  uint64_t kept = 0;
  for (uint64_t i = 0; i < n; ++i) { /* std::remove_if keeps the survivors' order */
    int drained = 0;
    if (xs[i].type != PBF_TYPE_OBSTACLE)
      for (uint32_t di = 0; di < scene->n_drains && !drained; ++di)
        if (dist3(scene->drains[di].centre, xs[i].position) < scene->drains[di].width) drained = 1;
    if (!drained) xs[kept++] = xs[i];
  }
  return (int64_t)kept;
}

void pbf_oracle_constants(float h, float out[3]) {
  out[0] = poly6_factor(h);
  out[1] = spiky_factor(h);
  out[2] = poly6(CorrDeltaQ * h, out[0], h);
}

#ifdef _OPENMP
#include <omp.h>
void pbf_oracle_set_threads(int n) { omp_set_num_threads(n); }
int pbf_oracle_max_threads(void) { return omp_get_max_threads(); }
#else
void pbf_oracle_set_threads(int n) { (void)n; }
int pbf_oracle_max_threads(void) { return 1; }
#endif
