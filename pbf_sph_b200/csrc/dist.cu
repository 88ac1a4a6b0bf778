// dist.cu — multi-GPU PBF step: Z-curve slab decomposition, one rank per GPU (include/pbf_cuda.h "multi-GPU").
//
// The reference is single-device; this is the north-star extension.  Rank r owns the particles
// whose Morton key (curves.h:72-88, computed from the PREDICTED position exactly as ompsph.hpp:152) lies in
// [split[r], split[r+1]).  Keys and the cell table are computed once per step and reused by every solver iteration
// (ompsph.hpp:215-249), so ownership, the ghost-cell sets and the send lists are static within a step.
//
// B200 / NVSwitch form: every exchange is a KERNEL THAT STORES INTO THE PEER'S MEMORY (each rank's particle arrays live
// in one arena that all peers map: directly inside one process, through CUDA IPC between processes), followed by a
// cross-rank barrier on the stream.  All counts — migrants, owned particles, ghosts, send lists — stay in device
// memory (`SlabDyn`); kernels read them there and grids are sized by the arena's capacities, so an ordinary step
// never waits for the host: the host runs up to two steps ahead and launch latency disappears behind the solver kernels.
// The host synchronises only on PLAN steps (after an upload, every `replan` steps, and EARLY when some rank has filled a
// capacity beyond its watermark — every rank derives that verdict from the count matrices, SlabDyn::hot), where it
// re-plans the key splits from a global key histogram and re-sizes the arenas from the exact counts.
//
//   A  predict_key on the held particles; (plan steps) global key histogram -> new splits;
//      classify every particle by the rank owning its key; count row -> every peer          [barrier 1]
//   B  stable destination-major list of the leaving particles; PUSH them (pos, vel, colour, id, key) behind the held
//      particles of the destination's input arrays                                          [barrier 2]
//   C  ONE stable radix sort of [arrivals from lower ranks | kept | arrivals from higher ranks] — the order that
//      reproduces the single-device stable sort when the ranks' inputs are consecutive blocks of one array;
//      ghost masks: a cell is sent to every rank owning a cell within Chebyshev distance 2 (delta needs lambda of
//      ring-1 ghosts, whose lambda needs ring 2); ghost count row -> every peer              [barrier 3]
//   D  reorder (gather + predict) into the local arrays [ghosts below | owned | ghosts above] — globally key-sorted
//      because ranks own ascending key ranges; the owned block starts at a FIXED offset; PUSH the ghost payload
//      (pStar|mass, colour, key) into the destination's local arrays                        [barrier 4]
//   E  cell table over the local array, roles -> compact lists (ring-1 ghosts, boundary), diffuse on the owned cell blocks
//      (side stream); per iteration: lambda on owned + ring-1 ghosts, then side by side: delta on the interior in place
//      (boundary particles skipped by mask) ‖ on the high-priority comm stream delta on the boundary list, PUSH of the
//      boundary pStar (16 B per ghost) into the destination's halo inbox, barrier; inbox -> ghost slots; finally
//      finalise (owned).
//   F  (params.surface_enabled) marching cubes: final positions + diffused colours of the send list -> the destinations'
//      ghost slots [two barriers]; every rank evaluates the lattice points whose cell it owns into RANK 0's lattice
//      [barrier]; rank 0 counts, scans, emits.
//
// Groups: NCCL (one process per GPU; NCCL carries the arena handles and the plan-step reductions; the data moves by peer
// stores and the barriers are a flag kernel over peer memory) and LOCAL (every rank is a context of this process;
// barriers are CUDA events between the ranks' streams — this is what the 1-GPU parity tests and the multi-device
// sph::Solver drive).  Uploads are collective: every rank uploads between the same two steps.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; the symbols are resolved with dlsym

#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "cells.cuh"
#include "common.cuh"

using namespace pbf;

namespace {

// ------------------------------------------------------------------------------------------------- NCCL binding
struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string err;
  bool load() {
    if (handle) return true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) { err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define PBF_SYM(field, sym)                                                      \
  field = reinterpret_cast<decltype(field)>(dlsym(handle, #sym));                \
  if (!field) { err = "libnccl lacks " #sym; handle = nullptr; return false; }
    PBF_SYM(GetUniqueId, ncclGetUniqueId)
    PBF_SYM(CommInitRank, ncclCommInitRank)
    PBF_SYM(CommDestroy, ncclCommDestroy)
    PBF_SYM(AllGather, ncclAllGather)
    PBF_SYM(AllReduce, ncclAllReduce)
    PBF_SYM(GetErrorString, ncclGetErrorString)
#undef PBF_SYM
    return true;
  }
};
NcclApi g_nccl;

#define PBF_NCCL(ctx, call)                                                                         \
  do {                                                                                              \
    ncclResult_t _r = (call);                                                                       \
    if (_r != ncclSuccess) return pbf::fail((ctx), PBF_ERR_NCCL, #call, g_nccl.GetErrorString(_r)); \
  } while (0)

constexpr uint32_t kKeyEnd = 1u << 30;   // one past the largest 30-bit Morton key
constexpr int kMaxWorld = 32;            // ghost masks are 32-bit
constexpr uint32_t kHistBits = 16;       // load-balance histogram: at most 65 536 coarse key buckets
constexpr int kBlk = 256;
constexpr int kMaxTimedIters = 16;       // lambda launches timed on a measurement step (load-balance feedback)

// ------------------------------------------------------------------------------------------------- device-side state
// Everything the host used to read back.  One instance per rank, in device memory; a pinned mirror is refreshed at the
// end of every step (asynchronously) for the statistics, the download and the capacity checks of the plan steps.
struct SlabDyn {
  uint32_t n_in;         // particles this rank holds when the step starts (the previous step's n_own)
  uint32_t n_keep, total_out, total_in, in_lo;  // migration: staying, leaving, arriving (from lower ranks: in_lo)
  uint32_t n_own;        // owned after migration
  uint32_t n_glo, n_ghi, n_local, first_local;  // ghosts below / above; local array = [first_local, first_local + n_local)
  uint32_t n_send;       // owned particles other ranks hold as ghosts, counted once per destination
  uint32_t n_ring1, n_boundary;
  uint32_t any_outside;  // some particle (on any rank) is predicted outside the grid (key >= G)
  uint32_t overflow;     // bit 0: held / owned particles, bit 1: ghost slots, bit 2: send lists — a capacity was exceeded; bit 3: barrier timeout
  uint32_t want[4];      // the UNCLAMPED needs of this rank on this step: ghosts below, ghosts above, send-list entries, held + arrivals
  uint32_t hot;          // some rank of the group fills a capacity beyond its watermark: re-plan early (the same value on every rank)
  uint32_t own_range[2];    // {own_off, n_own}
  uint32_t local_range[2];  // {first_local, n_local}
  uint32_t mig_send_off[kMaxWorld + 1];  // destination-major leave list: block of destination q
  uint32_t mig_dst[kMaxWorld];           // ... lands at this index of q's input arrays (relative to q's own_off)
  uint32_t mig_k2_dst[kMaxWorld];        // ... and its keys at this index of q's merge-key array
  uint32_t gh_send_off[kMaxWorld + 1];   // destination-major ghost send list: block of destination q
  uint32_t gh_dst[kMaxWorld];            // ... lands at this ABSOLUTE index of q's local arrays
  uint32_t halo_dst[kMaxWorld];          // ... and, per iteration, at this slot of q's halo inbox
};

// The arena of one rank: every array a peer may store into, at offsets that are the same on every rank (capacities are
// agreed on plan steps).  Particle arrays hold cap_local elements; the owned block starts at element own_off = cap_g.
struct ArenaLayout {
  uint32_t cap_g = 0, cap_own = 0, cap_local = 0, own_off = 0;  // cap_local = cap_g + cap_own + cap_g
  uint32_t world = 1, row_words = 0;
  size_t pos[2]{}, vel[2]{}, col[2]{}, pstar0 = 0, ids[2]{}, keys_local = 0, k2 = 0, halo[2]{}, rows[2]{}, flags = 0, bytes = 0;
  uint64_t cap_lattice = 0;       // marching cubes: lattice points the two arrays below hold (0 until a surface is asked for)
  size_t mc_pn = 0, mc_lc = 0;    // the surface lattice (field | normal, colour): every rank stores its points into rank 0's
};

struct PeerTable {
  char *base[kMaxWorld];
};

template <typename T> __host__ __device__ __forceinline__ T *arena_ptr(char *base, size_t off) {
  return reinterpret_cast<T *>(base + off);
}

}  // namespace

struct pbf_dist_state {
  int rank = 0, world = 1;
  bool local_mode = false;
  std::shared_ptr<std::vector<pbf_ctx *>> group;  // LOCAL: every rank's context; NCCL: just this one
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;             // halo push overlapped with the interior delta pass
  cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr;
  cudaEvent_t ev_bar[2] = {nullptr, nullptr};     // LOCAL barriers: this rank's arrival on the main / comm stream
  uint64_t step_index = 0;
  uint32_t replan_every = 8;
  uint32_t hist_shift = 0, hist_buckets = 0;
  std::vector<uint32_t> splits;                   // world + 1 key boundaries
  // arena
  char *arena = nullptr;
  ArenaLayout lay;
  PeerTable peers{};                              // every rank's arena as this rank addresses it
  std::vector<void *> ipc_open;                   // mappings opened with cudaIpcOpenMemHandle
  uint32_t *bar_word = nullptr;                   // NCCL all-reduce operand (arena re-builds)
  uint32_t bar_epoch[2] = {0, 0};                 // barriers passed on the main / comm stream since the arena was built
  // fresh upload, staged outside the arena until the next step agrees on capacities
  DevBuf<float4> up_pos, up_vel, up_col;
  DevBuf<unsigned long long> up_ids;
  uint64_t n_up = 0;
  bool fresh = false;
  // device scratch
  SlabDyn *dyn = nullptr, *h_dyn = nullptr;       // device / pinned mirror (valid after a stream sync)
  DevBuf<uint32_t> d_splits, d_row, d_hist, d_scratch, row_tot;  // row_tot: totals of the row-wise scans (kMaxWorld words)
  DevBuf<uint32_t> mask, send_idx, leave_idx, blk_cnt, v2, role, role_cnt, ring1_idx, bnd_idx;
  DevBuf<float4> pstar1;
  bool diffuse_pending = false;
  uint32_t cnt_nblk = 0;                          // tile stride of blk_cnt as its last count pass wrote it (the arena may grow before the scatter)
  uint64_t want_lattice = 0;                      // lattice points the next arena build must hold (marching cubes)
  uint32_t send_plan = 0;                         // send-list entries the last plan sized for (the same on every rank)
  // early re-plan: the `hot` verdict of every step comes back through a small ring of pinned words, read two steps later
  // (all ranks read the verdict of the SAME step, so they still agree on which steps are plan steps without talking)
  static constexpr int kHotRing = 4;
  uint32_t *h_hot = nullptr;
  cudaEvent_t ev_step[kHotRing] = {};
  uint64_t last_plan_step = 0;
  uint32_t n_plans = 0, n_early_plans = 0;        // statistics (pbf_dist_stats)
  uint32_t kept_nblk = 0;                         // tiles of the kept counts classify_count_kernel wrote into role_cnt (phase A -> phase C)
  uint32_t *h_pinned = nullptr;                   // plan steps: the two count matrices
  std::vector<uint64_t> last_counts;              // pbf_dist_advance_host: particles every rank returned last time
  // feedback for the load balance: lambda-pass time of this rank on the step before a plan step
  cudaEvent_t ev_lam[2 * kMaxTimedIters] = {};
  int n_timed = 0;
  double busy_ms = 0.0;
  std::vector<double> rate;                       // per rank: correction of the work model (1 = as modelled)
  std::vector<uint64_t> last_weights;             // bucket weights of the last plan
  void release_arena() {
    for (void *p : ipc_open) cudaIpcCloseMemHandle(p);
    ipc_open.clear();
    if (arena) cudaFree(arena);
    arena = nullptr;
  }
  void release() {
    release_arena();
    up_pos.release(); up_vel.release(); up_col.release(); up_ids.release();
    d_splits.release(); d_row.release(); d_hist.release(); d_scratch.release(); row_tot.release();
    mask.release(); send_idx.release(); leave_idx.release(); blk_cnt.release(); v2.release(); role.release();
    role_cnt.release(); ring1_idx.release(); bnd_idx.release(); pstar1.release();
    if (dyn) cudaFree(dyn);
    if (h_dyn) cudaFreeHost(h_dyn);
    if (bar_word) cudaFree(bar_word);
    if (h_pinned) cudaFreeHost(h_pinned);
    if (h_hot) cudaFreeHost(h_hot);
    for (cudaEvent_t e : ev_step)
      if (e) cudaEventDestroy(e);
    if (comm_stream) cudaStreamDestroy(comm_stream);
    for (cudaEvent_t e : {ev_boundary, ev_halo, ev_bar[0], ev_bar[1]})
      if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_lam)
      if (e) cudaEventDestroy(e);
  }
};

namespace {

using D = pbf_dist_state;

// ------------------------------------------------------------------------------------------------- kernels
__global__ void key_hist_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ n_dev, uint32_t shift,
                                uint32_t n_buckets, uint32_t *__restrict__ hist) {
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  if (i >= __ldg(n_dev)) return;
  const uint32_t b = min(__ldg(keys + i) >> shift, n_buckets - 1u);
  atomicAdd(hist + b, 1u);
}

__device__ __forceinline__ int owner_of(const uint32_t *splits, int world, uint32_t key) {
  int o = 0;
  for (int d = 1; d < world; ++d) o += (key >= splits[d]) ? 1 : 0;
  return o;
}

// OR of `m` over the block (all threads get it).  Most tiles send nothing anywhere: the loops over the destinations below
// run over the set bits of this only.
__device__ __forceinline__ uint32_t block_or(uint32_t m) {
  __shared__ uint32_t acc;
  if (threadIdx.x == 0) acc = 0u;
  __syncthreads();
  const uint32_t w = __reduce_or_sync(0xFFFFFFFFu, m);
  if ((threadIdx.x & 31u) == 0u && w) atomicOr(&acc, w);
  __syncthreads();
  const uint32_t r = acc;
  __syncthreads();  // `acc` may be reset by the next call
  return r;
}

// Destination of every held particle after predict_key: mask[i] = 1 << owner when the owner is another rank, else 0
// (the same mask format as the ghost lists, so ghost_scatter_kernel builds the leave lists), and in the same pass the
// per-tile counts the two compactions of the migration need: cnt[d * nblk + blk] = particles of the tile leaving for rank d
// (totals into row[d]) and kept[blk] = particles of the tile that stay.  *n_outside counts particles predicted outside the
// grid (key >= G), which the last rank owns.
__global__ void classify_count_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ n_dev,
                                      const uint32_t *__restrict__ splits_g, int rank, int world, uint32_t G, uint32_t nblk,
                                      uint32_t *__restrict__ mask, uint32_t *__restrict__ cnt, uint32_t *__restrict__ kept,
                                      uint32_t *__restrict__ row, uint32_t *__restrict__ n_outside) {
  __shared__ uint32_t splits[kMaxWorld + 1];
  if (threadIdx.x <= (unsigned)world) splits[threadIdx.x] = splits_g[threadIdx.x];
  __syncthreads();
  const uint32_t n = __ldg(n_dev);
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t key = i < n ? __ldg(keys + i) : 0u;
  const int o = owner_of(splits, world, key);
  const uint32_t m = (i < n && o != rank) ? 1u << o : 0u;
  if (i < n) mask[i] = m;
  const int outside = __syncthreads_count(i < n && key >= G);
  if (threadIdx.x == 0 && outside) atomicAdd(n_outside, (uint32_t)outside);
  const int stay = __syncthreads_count(i < n && m == 0u);
  if (threadIdx.x == 0) kept[blockIdx.x] = (uint32_t)stay;
  if ((int)threadIdx.x < world) cnt[threadIdx.x * nblk + blockIdx.x] = 0u;
  for (uint32_t todo = block_or(m); todo; todo &= todo - 1u) {  // most tiles lose nobody
    const int d = __ffs(todo) - 1;
    const int c = __syncthreads_count((m >> d) & 1u);
    if (threadIdx.x == 0) {
      cnt[(uint32_t)d * nblk + blockIdx.x] = (uint32_t)c;
      atomicAdd(row + d, (uint32_t)c);
    }
  }
}

// Ghost destinations of every owned particle: bit d of mask[i] = rank d needs particle i as a ghost, i.e. owns a cell
// within Chebyshev distance 2 of the particle's cell.  The search is per CELL and warp-cooperative: the first particle of
// each cell is its leader; for every leader in the warp the 32 lanes split the 125 neighbour cells between them and
// OR-reduce the owners; the leader then writes the answer for its whole cell.
__global__ void ghost_mask_kernel(const uint32_t *__restrict__ keys, const SlabDyn *__restrict__ dyn, uint32_t G, int rank,
                                  int world, const uint32_t *__restrict__ splits_g, uint32_t *__restrict__ mask) {
  __shared__ uint32_t splits[kMaxWorld + 1];
  if (threadIdx.x <= (unsigned)world) splits[threadIdx.x] = splits_g[threadIdx.x];
  __syncthreads();
  const uint32_t n = dyn->n_own;
  const bool any_outside = dyn->any_outside != 0u;
  if (blockIdx.x * kBlk >= n) return;
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  const unsigned lane = threadIdx.x & 31;
  const uint32_t key = i < n ? __ldg(keys + i) : 0xFFFFFFFFu;
  const bool leader = i < n && (i == 0 || __ldg(keys + i - 1) != key);
  const uint32_t lo = splits[rank], hi = splits[rank + 1];
  // Most cells are deep inside the slab.  The Morton key is monotone in every coordinate, so the keys of the 5 x 5 x 5
  // box lie between those of its two extreme corners: when both corners are ours, every cell of the box is, and the
  // 125-cell search is not needed (exact, not a heuristic; boxes that wrap around the 10-bit grid take the search).
  bool search = leader && key < G;
  if (search) {
    const uint32_t x = compact10(key), y = compact10(key >> 1), z = compact10(key >> 2);
    if (x >= 2u && y >= 2u && z >= 2u && x <= 1021u && y <= 1021u && z <= 1021u) {
      const uint32_t kmin = morton3(x - 2u, y - 2u, z - 2u), kmax = morton3(x + 2u, y + 2u, z + 2u);
      search = !(kmin >= lo && kmax < hi);
    }
  }
  unsigned leaders = __ballot_sync(0xFFFFFFFFu, search);
  uint32_t mine = 0;
  while (leaders) {
    const int src = __ffs(leaders) - 1;
    leaders &= leaders - 1;
    const uint32_t k = __shfl_sync(0xFFFFFFFFu, key, src);
    uint32_t m = 0;
    if (k < G) {  // a particle outside the grid is in no cell (sph.hpp:203-213): nobody can see it
      const uint32_t x = compact10(k), y = compact10(k >> 1), z = compact10(k >> 2);
      for (uint32_t q = lane; q < 125u; q += 32u) {
        // offsets -2..+2 with the reference's 10-bit wrap-around (0 - 1 -> 1023, 1023 + 1 -> 0; sph.hpp:221, curves.h:73)
        const uint32_t nk = morton3((x + q % 5u - 2u) & 1023u, (y + (q / 5u) % 5u - 2u) & 1023u, (z + q / 25u - 2u) & 1023u);
        // a particle predicted outside the grid (key >= G) still walks its 27 cells as `a` (ompsph.hpp:217-232):
        // cells >= G matter only while such particles exist
        if (nk >= G && !any_outside) continue;
        if (nk < lo || nk >= hi) m |= 1u << owner_of(splits, world, nk);
      }
    }
    m = __reduce_or_sync(0xFFFFFFFFu, m);
    if ((int)lane == src) mine = m;
  }
  if (leader)
    for (uint32_t j = i; j < n && __ldg(keys + j) == key; ++j) mask[j] = mine;
}

// Per 256-particle tile and destination: how many particles go there (cnt[d * nblk + blk], zero for the tiles of the
// capacity-sized grid that lie beyond the count); totals into row[d].
__global__ void ghost_count_kernel(const uint32_t *__restrict__ mask, const uint32_t *__restrict__ n_dev, int world,
                                   uint32_t nblk, uint32_t *__restrict__ cnt, uint32_t *__restrict__ row) {
  const uint32_t n = __ldg(n_dev);
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t m = i < n ? __ldg(mask + i) : 0u;
  if ((int)threadIdx.x < world) cnt[threadIdx.x * nblk + blockIdx.x] = 0u;
  for (uint32_t todo = block_or(m); todo; todo &= todo - 1u) {  // block_or's barriers order the zeroes before the counts
    const int d = __ffs(todo) - 1;
    const int c = __syncthreads_count((m >> d) & 1u);
    if (threadIdx.x == 0) {
      cnt[(uint32_t)d * nblk + blockIdx.x] = (uint32_t)c;
      atomicAdd(row + d, (uint32_t)c);
    }
  }
}

// Stable scatter of the send lists: list[start of destination d + offs[d][blk] + rank within the tile] = i,
// destination-major.  offs = the per-destination rows of tile counts, each scanned on its own; row_tot[d] = row d's total.
__global__ void ghost_scatter_kernel(const uint32_t *__restrict__ mask, const uint32_t *__restrict__ n_dev, int world,
                                     uint32_t nblk, const uint32_t *__restrict__ offs, const uint32_t *__restrict__ row_tot,
                                     uint32_t *__restrict__ list, uint32_t list_cap) {
  __shared__ uint32_t wsum[kBlk / 32];
  __shared__ uint32_t row_base[kMaxWorld];
  const uint32_t n = __ldg(n_dev);
  if (blockIdx.x * kBlk >= n) return;
  if (threadIdx.x < 32) {  // exclusive sums of the (at most 32) row totals
    const uint32_t t = (int)threadIdx.x < world ? __ldg(row_tot + threadIdx.x) : 0u;
    uint32_t incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if ((int)threadIdx.x >= o) incl += up;
    }
    row_base[threadIdx.x] = incl - t;
  }
  __syncthreads();
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t m = i < n ? __ldg(mask + i) : 0u;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t todo = block_or(m); todo; todo &= todo - 1u) {
    const int d = __ffs(todo) - 1;
    const bool bit = (m >> d) & 1u;
    const unsigned b = __ballot_sync(0xFFFFFFFFu, bit);
    if (lane == 0) wsum[warp] = __popc(b);
    __syncthreads();
    uint32_t base = row_base[d] + __ldg(offs + (uint32_t)d * nblk + blockIdx.x);
    for (unsigned w = 0; w < warp; ++w) base += wsum[w];
    const uint32_t slot = base + __popc(b & ((1u << lane) - 1u));
    if (bit && slot < list_cap) list[slot] = i;
    __syncthreads();
  }
}

// This rank's count row -> row `me` of a count matrix in EVERY rank's arena (peer stores).
// `held` (migration row only): the number of particles this rank holds, which travels as the row's last word.
__global__ void push_row_kernel(PeerTable peers, size_t rows_off, int me, int world, uint32_t row_words,
                                const uint32_t *__restrict__ row, const uint32_t *__restrict__ held) {
  for (uint32_t t = threadIdx.x; t < (uint32_t)world * row_words; t += blockDim.x) {
    const uint32_t w = t % row_words, q = t / row_words;
    arena_ptr<uint32_t>(peers.base[q], rows_off)[(size_t)me * row_words + w] = (held && w == (uint32_t)world + 1u) ? __ldg(held) : row[w];
  }
}

// After barrier 1: the migration matrix M (row s, column t = particles going from s to t; column W = particles outside
// the grid; column W + 1 = particles s holds) -> everything phases B / C need, for this rank.  One thread.
__global__ void plan_migration_kernel(const uint32_t *__restrict__ M_g, int me, int W, uint32_t row_words, ArenaLayout lay,
                                      SlabDyn *__restrict__ dyn) {
  __shared__ uint32_t M[kMaxWorld * (kMaxWorld + 2)];  // the matrix once, coalesced: the walks below are O(W^2) dependent loads
  for (uint32_t t = threadIdx.x; t < (uint32_t)W * row_words; t += blockDim.x) M[t] = M_g[t];
  __syncthreads();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t total_out = 0, total_in = 0, in_lo = 0, outside = 0;
  for (int s = 0; s < W; ++s) {
    outside += M[s * row_words + W];
    if (s == me) continue;
    total_in += M[s * row_words + me];
    total_out += M[me * row_words + s];
    if (s < me) in_lo += M[s * row_words + me];
  }
  const uint32_t n_in = M[me * row_words + W + 1];
  dyn->n_in = n_in;
  dyn->total_out = total_out;
  dyn->total_in = total_in;
  dyn->in_lo = in_lo;
  dyn->n_keep = n_in - total_out;
  dyn->n_own = n_in - total_out + total_in;
  dyn->any_outside = outside != 0u;
  if ((uint64_t)n_in + total_in > (uint64_t)lay.cap_local - lay.own_off || dyn->n_own > lay.cap_own) atomicOr(&dyn->overflow, 1u);
  // my leavers, destination-major, and where each block lands at its destination: arrivals are appended behind the
  // destination's held particles in SOURCE-RANK order (the merge order of phase C)
  uint32_t soff = 0;
  for (int q = 0; q < W; ++q) {
    dyn->mig_send_off[q] = soff;
    if (q == me) { dyn->mig_dst[q] = 0; dyn->mig_k2_dst[q] = 0; continue; }
    soff += M[me * row_words + q];
    const uint32_t q_in = M[q * row_words + W + 1];
    uint32_t before = 0, q_out = 0;  // arrivals at q from ranks below me; particles leaving q
    for (int s = 0; s < W; ++s) {
      if (s == q) continue;
      if (s < me) before += M[s * row_words + q];
      q_out += M[q * row_words + s];
    }
    dyn->mig_dst[q] = q_in + before;
    // q's merge keys: [from lower ranks | kept = q_in - q_out | from higher ranks]
    dyn->mig_k2_dst[q] = me < q ? before : before + (q_in - q_out);
  }
  dyn->mig_send_off[W] = soff;
}

// leaving particles, destination-major (leave_idx from ghost_scatter_kernel): raw state + key, stored straight into the
// destination's input arrays / merge-key array
__global__ void push_migrants_kernel(const SlabDyn *__restrict__ dyn, PeerTable peers, ArenaLayout lay, int cur, int cur_col,
                                     int world, const uint32_t *__restrict__ leave_idx, const float4 *__restrict__ pos,
                                     const float4 *__restrict__ vel, const float4 *__restrict__ col,
                                     const unsigned long long *__restrict__ ids, const uint32_t *__restrict__ keys) {
  const uint32_t j = blockIdx.x * kBlk + threadIdx.x;
  if (j >= dyn->total_out) return;
  int q = 0;
  while (q + 1 < world && j >= dyn->mig_send_off[q + 1]) ++q;
  const uint32_t k = j - dyn->mig_send_off[q];
  const uint32_t s = __ldg(leave_idx + j);
  char *base = peers.base[q];
  const size_t at = (size_t)lay.own_off + dyn->mig_dst[q] + k;
  if (at >= lay.cap_local || (size_t)dyn->mig_k2_dst[q] + k >= lay.cap_local) return;  // the destination flags the overflow
  arena_ptr<float4>(base, lay.pos[cur])[at] = ldg4(pos + s);
  arena_ptr<float4>(base, lay.vel[cur])[at] = ldg4(vel + s);
  arena_ptr<float4>(base, lay.col[cur_col])[at] = ldg4(col + s);
  arena_ptr<unsigned long long>(base, lay.ids[cur])[at] = __ldg(ids + s);
  arena_ptr<uint32_t>(base, lay.k2)[dyn->mig_k2_dst[q] + k] = __ldg(keys + s);
}

// Merge input of phase C: keys k2 = [arrivals from lower ranks (stored by their senders) | kept, in input order |
// arrivals from higher ranks (stored by their senders)], values v2 = index of each entry in the input arrays.  One
// kernel: thread t stably compacts held particle t into the kept block (per-tile offsets `offs` = the scanned `kept`
// counts of classify_count_kernel) and fills the value of merge slot t when that slot belongs to an arrival.
__global__ void merge_scatter_kernel(const SlabDyn *__restrict__ dyn, const uint32_t *__restrict__ mask,
                                     const uint32_t *__restrict__ offs, const uint32_t *__restrict__ key_in,
                                     uint32_t *__restrict__ k2, uint32_t *__restrict__ v2) {
  __shared__ uint32_t wsum[kBlk / 32];
  const uint32_t t = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t in_lo = dyn->in_lo, n_keep = dyn->n_keep, n_in = dyn->n_in, n_own = dyn->n_own;
  if (t < n_own) {  // arrivals sit behind the held particles of the input arrays, in source-rank order
    if (t < in_lo) v2[t] = n_in + t;
    else if (t >= in_lo + n_keep) v2[t] = n_in + (t - n_keep);
  }
  const bool bit = t < n_in && __ldg(mask + t) == 0u;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(0xFFFFFFFFu, bit);
  if (lane == 0) wsum[warp] = __popc(b);
  __syncthreads();
  if (!bit) return;
  uint32_t slot = in_lo + __ldg(offs + blockIdx.x) + __popc(b & ((1u << lane) - 1u));
  for (unsigned w = 0; w < warp; ++w) slot += wsum[w];
  v2[slot] = t;
  k2[slot] = __ldg(key_in + t);
}

// After barrier 3: the ghost matrix GC (row s, column t = owned particles of s that t holds as ghosts) and the
// migration matrix (for every rank's n_own) -> local layout and the destinations of this rank's ghost blocks.
__global__ void plan_ghosts_kernel(const uint32_t *__restrict__ GC_g, const uint32_t *__restrict__ M_g, int me, int W,
                                   uint32_t row_words, ArenaLayout lay, uint32_t send_cap, uint32_t send_plan,
                                   SlabDyn *__restrict__ dyn) {
  // both matrices into shared memory first (one coalesced sweep by the warp): thread 0's O(W^2) walks below would
  // otherwise be a chain of dependent global loads (15 us at 8 ranks)
  __shared__ uint32_t sGC[kMaxWorld * (kMaxWorld + 2)], sM[kMaxWorld * (kMaxWorld + 2)];
  for (uint32_t t = threadIdx.x; t < (uint32_t)W * row_words; t += blockDim.x) { sGC[t] = GC_g[t]; sM[t] = M_g[t]; }
  __syncthreads();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint32_t *GC = sGC, *M = sM;
  auto own_of = [&](int q) {
    uint32_t n = M[q * row_words + W + 1];
    for (int t = 0; t < W; ++t)
      if (t != q) n = n - M[q * row_words + t] + M[t * row_words + q];
    return n;
  };
  uint32_t n_send = 0, n_glo = 0, n_ghi = 0;
  for (int q = 0; q < W; ++q) {
    if (q == me) continue;
    n_send += GC[me * row_words + q];
    if (q < me) n_glo += GC[q * row_words + me]; else n_ghi += GC[q * row_words + me];
  }
  dyn->want[0] = n_glo; dyn->want[1] = n_ghi; dyn->want[2] = n_send;
  // Watermarks, for EVERY rank of the group from the two matrices every rank holds (so all ranks reach the same verdict
  // without talking): ghosts or send lists beyond 3/4, owned particles beyond 7/8 of the capacity the last plan gave them.
  uint32_t hot = 0;
  for (int q = 0; q < W; ++q) {
    uint64_t lo = 0, hi = 0, snd = 0;
    for (int t = 0; t < W; ++t) {
      if (t == q) continue;
      snd += GC[q * row_words + t];
      if (t < q) lo += GC[t * row_words + q]; else hi += GC[t * row_words + q];
    }
    if (4 * lo > 3ull * lay.cap_g || 4 * hi > 3ull * lay.cap_g || 4 * snd > 3ull * send_plan || 8ull * own_of(q) > 7ull * lay.cap_own) hot = 1;
  }
  dyn->hot = hot;
  uint32_t ov = 0;
  if (n_send > send_cap) { n_send = send_cap; ov |= 4u; }
  if (n_glo > lay.cap_g || n_ghi > lay.cap_g || (uint64_t)lay.own_off + dyn->n_own + n_ghi > lay.cap_local) ov |= 2u;
  n_glo = min(n_glo, lay.cap_g);
  n_ghi = min(n_ghi, lay.cap_g);
  dyn->n_send = n_send;
  dyn->n_glo = n_glo;
  dyn->n_ghi = n_ghi;
  dyn->n_local = n_glo + dyn->n_own + n_ghi;
  dyn->first_local = lay.own_off - n_glo;
  dyn->local_range[0] = dyn->first_local;
  dyn->local_range[1] = dyn->n_local;
  dyn->own_range[0] = lay.own_off;
  dyn->own_range[1] = dyn->n_own;
  uint32_t soff = 0;
  for (int q = 0; q < W; ++q) {
    dyn->gh_send_off[q] = soff;
    if (q == me) { dyn->gh_dst[q] = 0; dyn->halo_dst[q] = 0; continue; }
    soff += GC[me * row_words + q];
    // ghosts at q: below q's owned block the ranks s < q in rank order, above it the ranks s > q in rank order
    uint32_t q_glo = 0, before = 0;
    for (int s = 0; s < q; ++s) q_glo += GC[s * row_words + q];
    q_glo = min(q_glo, lay.cap_g);
    if (me < q) {
      for (int s = 0; s < me; ++s) before += GC[s * row_words + q];
      dyn->gh_dst[q] = lay.own_off - q_glo + before;
      dyn->halo_dst[q] = lay.cap_g - q_glo + before;  // inbox slots [0, cap_g): ghosts below, right-aligned
    } else {
      for (int s = q + 1; s < me; ++s) before += GC[s * row_words + q];
      dyn->gh_dst[q] = lay.own_off + own_of(q) + before;
      dyn->halo_dst[q] = lay.cap_g + before;          // inbox slots [cap_g, 2 cap_g): ghosts above
    }
  }
  dyn->gh_send_off[W] = soff;
  if (ov) atomicOr(&dyn->overflow, ov);
}

// ghost payload of the send list -> the destination's local arrays: pStar with the mass riding in the (still unused)
// lambda slot, colour, key
__global__ void push_ghosts_kernel(const SlabDyn *__restrict__ dyn, PeerTable peers, ArenaLayout lay, int cur_col, int world,
                                   const uint32_t *__restrict__ send_idx, const float4 *__restrict__ pstar,
                                   const float4 *__restrict__ pos_mass, const float4 *__restrict__ col,
                                   const uint32_t *__restrict__ keys_sorted) {
  const uint32_t j = blockIdx.x * kBlk + threadIdx.x;
  if (j >= dyn->n_send) return;
  int q = 0;
  while (q + 1 < world && j >= dyn->gh_send_off[q + 1]) ++q;
  const size_t at = (size_t)dyn->gh_dst[q] + (j - dyn->gh_send_off[q]);
  if (at >= lay.cap_local) return;
  const uint32_t i = __ldg(send_idx + j);  // index among the owned particles
  float4 p = ldg4(pstar + lay.own_off + i);
  p.w = __ldg(&pos_mass[lay.own_off + i].w);
  char *base = peers.base[q];
  arena_ptr<float4>(base, lay.pstar0)[at] = p;
  arena_ptr<float4>(base, lay.col[cur_col])[at] = ldg4(col + lay.own_off + i);
  arena_ptr<uint32_t>(base, lay.keys_local)[at] = __ldg(keys_sorted + i);
}

// per-iteration halo: pStar of the send list -> the destination's halo inbox of this iteration's parity
__global__ void push_halo_kernel(const SlabDyn *__restrict__ dyn, PeerTable peers, ArenaLayout lay, int parity, int world,
                                 const uint32_t *__restrict__ send_idx, const float4 *__restrict__ pstar) {
  const uint32_t j = blockIdx.x * kBlk + threadIdx.x;
  if (j >= dyn->n_send) return;
  int q = 0;
  while (q + 1 < world && j >= dyn->gh_send_off[q + 1]) ++q;
  const uint32_t slot = dyn->halo_dst[q] + (j - dyn->gh_send_off[q]);
  if (slot >= 2u * lay.cap_g) return;
  arena_ptr<float4>(peers.base[q], lay.halo[parity])[slot] = ldg4(pstar + lay.own_off + __ldg(send_idx + j));
}

// halo inbox -> the ghost slots of pStar (after the iteration's barrier)
__global__ void unpack_halo_kernel(const SlabDyn *__restrict__ dyn, ArenaLayout lay, const float4 *__restrict__ inbox,
                                   float4 *__restrict__ pstar) {
  const uint32_t t = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t n_glo = dyn->n_glo, n_ghi = dyn->n_ghi;
  if (t < n_glo) pstar[lay.own_off - n_glo + t] = inbox[lay.cap_g - n_glo + t];
  else if (t - n_glo < n_ghi) pstar[lay.own_off + dyn->n_own + (t - n_glo)] = inbox[lay.cap_g + (t - n_glo)];
}

// received ghosts: move the mass from pStar.w into pos.w (the lambda pass reads its own particle's mass there); also
// plants the key sentinels just outside the local array (loops that walk a cell stop there)
__global__ void ghost_fix_kernel(const SlabDyn *__restrict__ dyn, ArenaLayout lay, float4 *__restrict__ pstar,
                                 float4 *__restrict__ pos, uint32_t *__restrict__ keys_local) {
  const uint32_t t = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t n_glo = dyn->n_glo, n_ghi = dyn->n_ghi, n_own = dyn->n_own;
  if (t == 0) {
    if (dyn->first_local > 0) keys_local[dyn->first_local - 1] = 0xFFFFFFFFu;
    if (dyn->first_local + dyn->n_local < lay.cap_local) keys_local[dyn->first_local + dyn->n_local] = 0xFFFFFFFFu;
  }
  uint32_t i;
  if (t < n_glo) i = lay.own_off - n_glo + t;
  else if (t - n_glo < n_ghi) i = lay.own_off + n_own + (t - n_glo);
  else return;
  float4 p = pstar[i];
  pos[i] = make_float4(0.f, 0.f, 0.f, p.w);
  p.w = 0.f;
  pstar[i] = p;
}

// Marching cubes on the slab path: the owners' FINAL positions (pStar * scale, the finalise arithmetic, with the mass) and
// DIFFUSED colours of the send list -> the ghost slots of the destination's position / colour arrays.  Runs after a barrier
// behind every rank's finalise (nobody reads those slots any more in this step).
__global__ void push_surface_kernel(const SlabDyn *__restrict__ dyn, PeerTable peers, ArenaLayout lay, int cur, int cur_col, int world,
                                    float scale, const uint32_t *__restrict__ send_idx, const float4 *__restrict__ pstar,
                                    const float4 *__restrict__ pos_mass, const float4 *__restrict__ col) {
  const uint32_t j = blockIdx.x * kBlk + threadIdx.x;
  if (j >= dyn->n_send) return;
  int q = 0;
  while (q + 1 < world && j >= dyn->gh_send_off[q + 1]) ++q;
  const size_t at = (size_t)dyn->gh_dst[q] + (j - dyn->gh_send_off[q]);
  if (at >= lay.cap_local) return;
  const uint32_t i = lay.own_off + __ldg(send_idx + j);
  const float4 p = ldg4(pstar + i);
  char *base = peers.base[q];
  arena_ptr<float4>(base, lay.pos[cur])[at] = make_float4(fmul(p.x, scale), fmul(p.y, scale), fmul(p.z, scale), __ldg(&pos_mass[i].w));
  arena_ptr<float4>(base, lay.col[cur_col])[at] = ldg4(col + i);
}

// Role of every particle of the local array for the solver passes (static within a step), as per-tile counts for the
// two compact lists the passes run over (the interior pass runs over the owned range in place, skipping the boundary):
//   ring-1 ghosts   a ghost one of whose 27 cells is ours: lambda is computed here (besides the owned particles)
//   boundary        owned, and some other rank holds it as a ghost: its delta pass runs first so the halo can leave
//   interior        owned, nobody else needs it: its delta pass overlaps the halo push
constexpr uint32_t kRoleRing1 = 1u, kRoleBoundary = 2u, kRoleInterior = 4u;
__global__ void roles_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ owned_mask,
                             const SlabDyn *__restrict__ dyn, ArenaLayout lay, uint32_t G, uint32_t lo, uint32_t hi,
                             uint32_t nblk, uint32_t *__restrict__ role, uint32_t *__restrict__ cnt) {
  const uint32_t t = blockIdx.x * kBlk + threadIdx.x;  // index into the local array, from first_local
  const uint32_t n_local = dyn->n_local, first = dyn->first_local, n_own = dyn->n_own;
  const bool any_outside = dyn->any_outside != 0u;
  uint32_t f = 0;
  if (t < n_local) {
    const uint32_t i = first + t;
    if (i >= lay.own_off && i < lay.own_off + n_own) {
      f = __ldg(owned_mask + (i - lay.own_off)) != 0u ? kRoleBoundary : kRoleInterior;
    } else {
      const uint32_t key = __ldg(keys + i);
      const uint32_t kx = key & kAxisMask, ky = (key >> 1) & kAxisMask, kz = (key >> 2) & kAxisMask;
      const uint32_t ax[3] = {dilated_dec(kx), kx, dilated_inc(kx)};
      const uint32_t ay[3] = {dilated_dec(ky), ky, dilated_inc(ky)};
      const uint32_t az[3] = {dilated_dec(kz), kz, dilated_inc(kz)};
      bool ring1 = false;
      for (int z = 0; z < 3; ++z)
        for (int y = 0; y < 3; ++y)
          for (int x = 0; x < 3; ++x) {
            const uint32_t nk = (az[z] << 2) | (ay[y] << 1) | ax[x];
            if (nk >= G && !any_outside) continue;
            ring1 |= nk >= lo && nk < hi;
          }
      f = ring1 ? kRoleRing1 : 0u;
    }
    role[t] = f;
  }
  for (uint32_t k = 0; k < 2u; ++k) {
    const int c = __syncthreads_count((f >> k) & 1u);
    if (threadIdx.x == 0) cnt[k * nblk + blockIdx.x] = (uint32_t)c;
  }
}
// the two lists (absolute indices, ascending) from the scanned tile counts; list k starts at offs[k * nblk]
__global__ void role_lists_kernel(const uint32_t *__restrict__ role, SlabDyn *__restrict__ dyn, uint32_t nblk,
                                  const uint32_t *__restrict__ offs, const uint32_t *__restrict__ row_tot,
                                  uint32_t *__restrict__ ring1, uint32_t *__restrict__ bnd, uint32_t cap_ring1,
                                  uint32_t cap_own) {
  __shared__ uint32_t wsum[kBlk / 32];
  const uint32_t t = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t n_local = dyn->n_local;
  if (t == 0) {
    const uint32_t r1 = __ldg(row_tot);
    dyn->n_ring1 = min(r1, cap_ring1);
    dyn->n_boundary = min(__ldg(row_tot + 1), cap_own);
    if (r1 > cap_ring1) atomicOr(&dyn->overflow, 2u);
  }
  if (blockIdx.x * kBlk >= n_local) return;
  const uint32_t f = t < n_local ? __ldg(role + t) : 0u;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t k = 0; k < 2u; ++k) {
    const bool bit = (f >> k) & 1u;
    const unsigned b = __ballot_sync(0xFFFFFFFFu, bit);
    if (lane == 0) wsum[warp] = __popc(b);
    __syncthreads();
    uint32_t base = __ldg(offs + k * nblk + blockIdx.x);  // the rows are scanned one by one
    for (unsigned w = 0; w < warp; ++w) base += wsum[w];
    const uint32_t slot = base + __popc(b & ((1u << lane) - 1u));
    uint32_t *out = k == 0 ? ring1 : bnd;
    if (bit && slot < (k == 0 ? cap_ring1 : cap_own)) out[slot] = dyn->first_local + t;
    __syncthreads();
  }
}

__global__ void set_count_kernel(SlabDyn *__restrict__ dyn, uint32_t n_in) {
  if (threadIdx.x == 0 && blockIdx.x == 0) dyn->n_in = n_in;
}
// end of a step: the next one starts from this step's owned particles
__global__ void next_step_kernel(SlabDyn *__restrict__ dyn) {
  if (threadIdx.x == 0 && blockIdx.x == 0) dyn->n_in = dyn->n_own;
}

// Cross-rank barrier on a stream, between processes, without a collective library call: lane q of one warp stores this
// rank's epoch into flag [me] of rank q's arena (after a system-scope fence: every store the earlier kernels of this stream
// made into peer memory is ordered before the flag), then spins until rank q's epoch has arrived in its own arena.  Epochs
// only grow between two arena builds.  ~3 us over NVLink against ~14 us for a one-word ncclAllReduce; seven per step.
// A peer that never arrives (it failed) must not hang the GPU: after ~4 s the wait gives up and flags the step (bit 3).
__global__ void flag_barrier_kernel(PeerTable peers, size_t flags_off, int me, int world, uint32_t epoch, SlabDyn *dyn) {
  const int q = (int)threadIdx.x;
  if (q < world && q != me) {
    __threadfence_system();
    uint32_t *remote = arena_ptr<uint32_t>(peers.base[q], flags_off) + me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const uint32_t *mine = arena_ptr<uint32_t>(peers.base[me], flags_off) + q;
    unsigned long long t0 = 0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const bool dead = (*(volatile uint32_t *)&dyn->overflow & 8u) != 0u;  // a peer already failed to arrive: never wait again
    for (; !dead;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 4000000000ull) { atomicOr(&dyn->overflow, 8u); break; }
    }
  }
  __syncwarp();
  __threadfence_system();
}

// ------------------------------------------------------------------------------------------------- cross-rank plumbing
int sync_all(std::vector<pbf_ctx *> &L) {
  for (pbf_ctx *c : L) {
    PBF_CUDA(c, cudaSetDevice(c->device));
    PBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->dist->comm_stream) PBF_CUDA(c, cudaStreamSynchronize(c->dist->comm_stream));
  }
  return PBF_OK;
}

// Barrier between the ranks' streams (which = 0 main, 1 comm): nothing enqueued after it on any rank runs before
// everything enqueued before it on every rank has finished — peer stores included.  No host synchronisation.
int barrier(std::vector<pbf_ctx *> &L, int which) {
  D *d0 = L[0]->dist;
  if (d0->world == 1) return PBF_OK;
  if (!d0->local_mode) {
    pbf_ctx *c = L[0];
    cudaStream_t st = which ? d0->comm_stream : c->stream;
    PhaseScope ps(c, PBF_PH_SLAB_BARRIER, st);
    const uint32_t epoch = ++d0->bar_epoch[which];
    flag_barrier_kernel<<<1, 32, 0, st>>>(d0->peers, d0->lay.flags + (size_t)which * kMaxWorld * 4, d0->rank, d0->world, epoch, d0->dyn);
    PBF_LAUNCH_CHECK(c);
    return PBF_OK;
  }
  for (pbf_ctx *c : L) {
    PBF_CUDA(c, cudaSetDevice(c->device));
    PBF_CUDA(c, cudaEventRecord(c->dist->ev_bar[which], which ? c->dist->comm_stream : c->stream));
  }
  for (pbf_ctx *c : L) {
    PBF_CUDA(c, cudaSetDevice(c->device));
    for (pbf_ctx *s : L)
      if (s != c) PBF_CUDA(c, cudaStreamWaitEvent(which ? c->dist->comm_stream : c->stream, s->dist->ev_bar[which], 0));
  }
  return PBF_OK;
}

template <typename FB> int all_reduce_sum_u32(std::vector<pbf_ctx *> &L, FB buf, size_t count) {
  D *d0 = L[0]->dist;
  if (!d0->local_mode) {
    pbf_ctx *c = L[0];
    PBF_NCCL(c, g_nccl.AllReduce(buf(c), buf(c), count, ncclUint32, ncclSum, d0->comm, c->stream));
    return PBF_OK;
  }
  PBF_TRY(sync_all(L));
  std::vector<uint32_t> acc(count, 0), tmp(count);
  for (pbf_ctx *s : L) {
    PBF_CUDA(s, cudaSetDevice(s->device));
    PBF_CUDA(s, cudaMemcpy(tmp.data(), buf(s), count * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < count; ++i) acc[i] += tmp[i];
  }
  for (pbf_ctx *s : L) {
    PBF_CUDA(s, cudaSetDevice(s->device));
    PBF_CUDA(s, cudaMemcpy(buf(s), acc.data(), count * 4, cudaMemcpyHostToDevice));
  }
  return PBF_OK;
}

// One host value per rank -> every rank's vector (plan steps only; synchronous).  `mine` holds one entry per LOCAL
// context (all of them in LOCAL mode, this rank's in NCCL mode).
int all_gather_host(std::vector<pbf_ctx *> &L, const std::vector<uint64_t> &mine, std::vector<uint64_t> &out) {
  D *d0 = L[0]->dist;
  const int W = d0->world;
  out.assign(W, 0);
  if (d0->local_mode) {
    for (int r = 0; r < W; ++r) out[r] = mine[r];
    return PBF_OK;
  }
  pbf_ctx *c = L[0];
  PBF_CUDA(c, cudaSetDevice(c->device));
  std::vector<uint32_t> buf(2 * W, 0);
  buf[2 * d0->rank] = (uint32_t)(mine[0] & 0xFFFFFFFFull);
  buf[2 * d0->rank + 1] = (uint32_t)(mine[0] >> 32);
  PBF_CUDA(c, cudaMemcpyAsync(d0->d_scratch.p, buf.data(), 2 * W * 4, cudaMemcpyHostToDevice, c->stream));
  PBF_NCCL(c, g_nccl.AllReduce(d0->d_scratch.p, d0->d_scratch.p, 2 * W, ncclUint32, ncclSum, d0->comm, c->stream));
  PBF_CUDA(c, cudaMemcpyAsync(buf.data(), d0->d_scratch.p, 2 * W * 4, cudaMemcpyDeviceToHost, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int r = 0; r < W; ++r) out[r] = (uint64_t)buf[2 * r] | ((uint64_t)buf[2 * r + 1] << 32);
  return PBF_OK;
}

// ------------------------------------------------------------------------------------------------- arena
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

ArenaLayout make_layout(uint32_t cap_g, uint32_t cap_own, int world, uint64_t cap_lattice) {
  ArenaLayout l;
  // Capacities are multiples of 256 elements: the owned block starts at element cap_g of every array, and the solver
  // passes index the neighbour-list rows, pStar and the keys by that position — a warp's 32 consecutive particles must
  // start on a 128-byte line (an odd cap_g made every coalesced access of the iteration kernels straddle two lines).
  cap_g = (uint32_t)align_up(cap_g, 256);
  cap_own = (uint32_t)align_up(cap_own, 256);
  l.cap_g = cap_g; l.cap_own = cap_own; l.cap_local = cap_g + cap_own + cap_g; l.own_off = cap_g;
  l.world = (uint32_t)world; l.row_words = (uint32_t)world + 2u;
  size_t at = 0;
  auto take = [&](size_t bytes) { const size_t o = at; at = align_up(at + bytes, 256); return o; };
  for (int i = 0; i < 2; ++i) l.pos[i] = take((size_t)l.cap_local * 16);
  for (int i = 0; i < 2; ++i) l.vel[i] = take((size_t)l.cap_local * 16);
  for (int i = 0; i < 2; ++i) l.col[i] = take((size_t)l.cap_local * 16);
  l.pstar0 = take((size_t)l.cap_local * 16);
  for (int i = 0; i < 2; ++i) l.ids[i] = take((size_t)l.cap_local * 8);
  l.keys_local = take((size_t)l.cap_local * 4);
  l.k2 = take((size_t)l.cap_local * 4);
  l.cap_lattice = cap_lattice;
  l.mc_pn = take((size_t)cap_lattice * 16);
  l.mc_lc = take((size_t)cap_lattice * 16);
  for (int i = 0; i < 2; ++i) l.halo[i] = take((size_t)2 * cap_g * 16);
  for (int i = 0; i < 2; ++i) l.rows[i] = take((size_t)world * l.row_words * 4);
  l.flags = take((size_t)2 * kMaxWorld * 4);  // barrier flags: [main | comm stream][source rank], zeroed with the rows
  l.bytes = at;
  return l;
}

template <typename T> void borrow(DevBuf<T> &b, void *p, size_t cap) {
  b.release();
  b.p = static_cast<T *>(p);
  b.cap = cap;
  b.borrowed = true;
}

// (Re)allocate every rank's arena for the given capacities, keeping the `live` held particles of the input arrays (at the
// old own_off) where there is an old arena, and re-establish the peer mappings.  Collective and synchronous.
int build_arenas(std::vector<pbf_ctx *> &L, uint32_t cap_g, uint32_t cap_own, const std::vector<uint64_t> &live) {
  D *d0 = L[0]->dist;
  const int W = d0->world;
  PBF_TRY(sync_all(L));
  for (size_t r = 0; r < L.size(); ++r) {
    pbf_ctx *c = L[r];
    D *d = c->dist;
    PBF_CUDA(c, cudaSetDevice(c->device));
    const ArenaLayout nl = make_layout(cap_g, cap_own, W, std::max(d->lay.cap_lattice, d->want_lattice));
    char *fresh = nullptr;
    PBF_CUDA(c, cudaMalloc(&fresh, nl.bytes));
    PBF_CUDA(c, cudaMemsetAsync(fresh + nl.rows[0], 0, nl.bytes - nl.rows[0], c->stream));
    const uint64_t keep = live.empty() ? 0 : live[r];
    if (d->arena && keep) {
      const ArenaLayout &ol = d->lay;
      if (keep > (uint64_t)nl.cap_local - nl.own_off) return fail(c, PBF_ERR_CAPACITY, "slab arena", "new capacity below the live particles");
      for (const auto &a : {std::make_pair(nl.pos[c->cur], ol.pos[c->cur]), std::make_pair(nl.vel[c->cur], ol.vel[c->cur]),
                            std::make_pair(nl.col[c->cur_col], ol.col[c->cur_col])})
        PBF_CUDA(c, cudaMemcpyAsync(fresh + a.first + (size_t)nl.own_off * 16, d->arena + a.second + (size_t)ol.own_off * 16, keep * 16,
                                    cudaMemcpyDeviceToDevice, c->stream));
      PBF_CUDA(c, cudaMemcpyAsync(fresh + nl.ids[c->cur] + (size_t)nl.own_off * 8, d->arena + ol.ids[c->cur] + (size_t)ol.own_off * 8,
                                  keep * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    // the sorted keys of the step in flight live in the owned block of the local key array (key_a is borrowed from the
    // arena): a re-grow between the sort and the reorder (plan steps, exact ghost counts) must carry them over
    const bool keys_in_arena = d->arena && c->key_a.borrowed && c->keys_sorted == c->key_a.p;
    if (keys_in_arena) {
      const ArenaLayout &ol = d->lay;
      const size_t n_keys = std::min<size_t>(ol.cap_local - ol.own_off, nl.cap_local - nl.own_off);
      PBF_CUDA(c, cudaMemcpyAsync(fresh + nl.keys_local + (size_t)nl.own_off * 4, d->arena + ol.keys_local + (size_t)ol.own_off * 4,
                                  n_keys * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    PBF_CUDA(c, cudaStreamSynchronize(c->stream));
    d->release_arena();
    d->arena = fresh;
    d->lay = nl;
    d->bar_epoch[0] = d->bar_epoch[1] = 0;
    for (int i = 0; i < 2; ++i) {
      borrow(c->pos[i], fresh + nl.pos[i], nl.cap_local);
      borrow(c->vel[i], fresh + nl.vel[i], nl.cap_local);
      borrow(c->col[i], fresh + nl.col[i], nl.cap_local);
      borrow(c->ids[i], fresh + nl.ids[i], nl.cap_local);
    }
    borrow(c->pstar[0], fresh + nl.pstar0, nl.cap_local);
    // the radix sort's last pass lands in key_a: make that the owned block of the local key array (no copy in phase D)
    borrow(c->key_a, fresh + nl.keys_local + (size_t)nl.own_off * 4, nl.cap_local - nl.own_off);
    if (keys_in_arena) c->keys_sorted = c->key_a.p;
    // buffers sized by the capacities, never re-allocated in mid-step (those that may hold live data keep it)
    PBF_CUDA(c, d->pstar1.reserve(nl.cap_local));
    borrow(c->pstar[1], d->pstar1.p, nl.cap_local);
    PBF_CUDA(c, c->key_in.reserve(nl.cap_local, true, c->stream));
    PBF_CUDA(c, d->mask.reserve(nl.cap_local, true, c->stream));
    PBF_CUDA(c, d->blk_cnt.reserve((size_t)W * div_up(nl.cap_local, kBlk) + 8, true, c->stream));
    PBF_CUDA(c, d->v2.reserve(nl.cap_local));
    PBF_CUDA(c, d->leave_idx.reserve(nl.cap_local));
    PBF_CUDA(c, d->send_idx.reserve((size_t)4 * nl.cap_g + 1024));
    PBF_CUDA(c, d->role.reserve(nl.cap_local));
    PBF_CUDA(c, d->ring1_idx.reserve((size_t)2 * nl.cap_g + 1));
    PBF_CUDA(c, d->bnd_idx.reserve(nl.cap_own + 1));
    PBF_CUDA(c, d->role_cnt.reserve((size_t)3 * div_up(nl.cap_local, kBlk) + 8, true, c->stream));  // holds the kept counts across a re-grow
    PBF_CUDA(c, c->rho.reserve(nl.cap_local));
  }
  // peer tables
  if (d0->local_mode) {
    for (pbf_ctx *c : L) {
      PBF_CUDA(c, cudaSetDevice(c->device));
      for (pbf_ctx *s : L) {
        c->dist->peers.base[s->dist->rank] = s->dist->arena;
        if (s->device != c->device) {
          int can = 0;
          PBF_CUDA(c, cudaDeviceCanAccessPeer(&can, c->device, s->device));
          if (!can) return fail(c, PBF_ERR_CUDA, "slab arena", "the devices of this group cannot address each other's memory");
          const cudaError_t e = cudaDeviceEnablePeerAccess(s->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, PBF_ERR_CUDA, "cudaDeviceEnablePeerAccess", cudaGetErrorString(e));
          cudaGetLastError();
        }
      }
    }
    return PBF_OK;
  }
  // one process per GPU: the arena handles travel through NCCL, every peer maps every arena (CUDA IPC)
  pbf_ctx *c = L[0];
  D *d = c->dist;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t mine;
  PBF_CUDA(c, cudaIpcGetMemHandle(&mine, d->arena));
  std::vector<cudaIpcMemHandle_t> all(W);
  PBF_CUDA(c, cudaMemcpyAsync(d->d_scratch.p + (size_t)16 * W, &mine, 64, cudaMemcpyHostToDevice, c->stream));
  PBF_NCCL(c, g_nccl.AllGather(d->d_scratch.p + (size_t)16 * W, d->d_scratch.p, 16, ncclUint32, d->comm, c->stream));
  PBF_CUDA(c, cudaMemcpyAsync(all.data(), d->d_scratch.p, (size_t)64 * W, cudaMemcpyDeviceToHost, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int q = 0; q < W; ++q) {
    if (q == d->rank) { d->peers.base[q] = d->arena; continue; }
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(c, PBF_ERR_CUDA, "cudaIpcOpenMemHandle (peer arena)", cudaGetErrorString(e));
    d->ipc_open.push_back(p);
    d->peers.base[q] = static_cast<char *>(p);
  }
  // nobody goes on (and possibly re-allocates again) before every peer holds the new mappings
  PBF_NCCL(c, g_nccl.AllReduce(d->bar_word, d->bar_word, 1, ncclUint32, ncclSum, d->comm, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));
  return PBF_OK;
}

uint32_t headroom(uint64_t n) { return (uint32_t)std::min<uint64_t>(0xFFFF0000ull, (n + n / 2 + 4096 + 255) / 256 * 256); }

// ------------------------------------------------------------------------------------------------- host planning
void plan_splits(const uint64_t *hist, uint32_t n_buckets, uint32_t shift, int world, uint32_t *splits) {
  uint64_t total = 0;
  for (uint32_t b = 0; b < n_buckets; ++b) total += hist[b];
  splits[0] = 0;
  uint64_t run = 0;
  uint32_t b = 0;
  for (int r = 1; r < world; ++r) {
    const uint64_t target = (total * (uint64_t)r + world - 1) / world;  // ceil(r * total / world)
    while (b < n_buckets && run < target) run += hist[b++];
    const uint64_t key = (uint64_t)b << shift;
    splits[r] = (uint32_t)(key > kKeyEnd ? kKeyEnd : key);
    if (splits[r] < splits[r - 1]) splits[r] = splits[r - 1];
  }
  splits[world] = kKeyEnd;
}

int owner_host(const std::vector<uint32_t> &splits, int world, uint64_t key) {
  int o = 0;
  for (int q = 1; q < world; ++q) o += key >= splits[q] ? 1 : 0;
  return o;
}

// ------------------------------------------------------------------------------------------------- the phases
void set_grid_constants(pbf_ctx *c, const pbf_params &p) {
  const ArenaLayout &l = c->dist->lay;
  host_step_const(c->h, p, c->grid, l.cap_local - l.own_off, c->sc);
}

// Start of a step on one rank: grid constants, (first step after an upload) the staged particles into the arena,
// predict_key, (re-plan steps) the key histogram.
int phase_a(pbf_ctx *c, const pbf_params &p, bool replan) {
  D *d = c->dist;
  PBF_CUDA(c, cudaSetDevice(c->device));
  const ArenaLayout &l = d->lay;
  set_grid_constants(c, p);
  PBF_CUDA(c, c->table.reserve((size_t)c->sc.G + 1));
  if (d->fresh) {  // the upload waited outside the arena until the capacities were agreed
    const size_t n = d->n_up;
    if (n) {
      PBF_CUDA(c, cudaMemcpyAsync(c->pos[c->cur].p + l.own_off, d->up_pos.p, n * 16, cudaMemcpyDeviceToDevice, c->stream));
      PBF_CUDA(c, cudaMemcpyAsync(c->vel[c->cur].p + l.own_off, d->up_vel.p, n * 16, cudaMemcpyDeviceToDevice, c->stream));
      PBF_CUDA(c, cudaMemcpyAsync(c->col[c->cur_col].p + l.own_off, d->up_col.p, n * 16, cudaMemcpyDeviceToDevice, c->stream));
      PBF_CUDA(c, cudaMemcpyAsync(c->ids[c->cur].p + l.own_off, d->up_ids.p, n * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    set_count_kernel<<<1, 32, 0, c->stream>>>(d->dyn, (uint32_t)n);
    PBF_LAUNCH_CHECK(c);
    d->fresh = false;
  }
  c->sc.n_dyn = &d->dyn->n_in;
  PBF_TRY(launch_predict_key(c, c->pos[c->cur].p + l.own_off, c->vel[c->cur].p + l.own_off, c->key_in.p));
  if (replan) {
    const uint32_t bits = c->grid.key_bits;
    d->hist_shift = bits > kHistBits ? bits - kHistBits : 0;
    d->hist_buckets = ((c->grid.grid_table_n - 1) >> d->hist_shift) + 2;  // last bucket: everything >= G
    PBF_CUDA(c, d->d_hist.reserve(d->hist_buckets + 64));
    PBF_CUDA(c, cudaMemsetAsync(d->d_hist.p, 0, d->hist_buckets * 4, c->stream));
    PhaseScope ps(c, PBF_PH_HALO);
    key_hist_kernel<<<div_up(c->sc.n, kBlk), kBlk, 0, c->stream>>>(c->key_in.p, &d->dyn->n_in, d->hist_shift, d->hist_buckets, d->d_hist.p);
    PBF_LAUNCH_CHECK(c);
  }
  return PBF_OK;
}

// New key splits from the all-reduced histogram.  Work, not particle counts, is balanced: a particle's cost grows with the
// local density (the search tests every particle of its 27 cells; pbf_host_work_weights), and that model is corrected by
// what the ranks MEASURED on the step before: `rate[r]` = lambda-pass milliseconds per unit of modelled work on rank r,
// relative to the mean, accumulated over the plans — a rank that was slower than the model said (more ghosts, denser
// water) hands keys to its neighbours.
int phase_plan(pbf_ctx *c, const std::vector<double> &rate) {
  D *d = c->dist;
  PBF_CUDA(c, cudaSetDevice(c->device));
  std::vector<uint32_t> h32(d->hist_buckets);
  PBF_CUDA(c, cudaMemcpyAsync(h32.data(), d->d_hist.p, d->hist_buckets * 4, cudaMemcpyDeviceToHost, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<uint64_t> h64(d->hist_buckets);
  pbf_host_work_weights(h32.data(), d->hist_buckets, d->hist_shift, h64.data());
  if (!d->splits.empty() && rate.size() == (size_t)d->world)
    for (uint32_t b = 0; b < d->hist_buckets; ++b)
      h64[b] = (uint64_t)((double)h64[b] * rate[owner_host(d->splits, d->world, std::min<uint64_t>((uint64_t)b << d->hist_shift, kKeyEnd - 1))] + 0.5);
  d->last_weights = h64;
  d->splits.assign(d->world + 1, 0);
  plan_splits(h64.data(), d->hist_buckets, d->hist_shift, d->world, d->splits.data());
  PBF_CUDA(c, cudaMemcpyAsync(d->d_splits.p, d->splits.data(), (d->world + 1) * 4, cudaMemcpyHostToDevice, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));  // the host vector may change before the copy would run
  return PBF_OK;
}

// classify the held particles by destination rank; count row (destinations | outside | held) -> every peer
int phase_a2(pbf_ctx *c) {
  D *d = c->dist;
  const int W = d->world;
  const ArenaLayout &l = d->lay;
  PBF_CUDA(c, cudaSetDevice(c->device));
  PhaseScope ps(c, PBF_PH_HALO);
  PBF_CUDA(c, cudaMemsetAsync(d->d_row.p, 0, (W + 2) * 4, c->stream));
  const uint32_t nblk = d->cnt_nblk = d->kept_nblk = div_up(l.cap_local - l.own_off, kBlk);
  classify_count_kernel<<<nblk, kBlk, 0, c->stream>>>(c->key_in.p, &d->dyn->n_in, d->d_splits.p, d->rank, W, c->sc.G, nblk, d->mask.p,
                                                      d->blk_cnt.p, d->role_cnt.p, d->d_row.p, d->d_row.p + W);
  PBF_LAUNCH_CHECK(c);
  push_row_kernel<<<1, 256, 0, c->stream>>>(d->peers, l.rows[0], d->rank, W, l.row_words, d->d_row.p, &d->dyn->n_in);
  PBF_LAUNCH_CHECK(c);
  return PBF_OK;
}

// after barrier 1: plan the migration on the device, list and push the leavers
int phase_b(pbf_ctx *c) {
  D *d = c->dist;
  const int W = d->world;
  const ArenaLayout &l = d->lay;
  PBF_CUDA(c, cudaSetDevice(c->device));
  PhaseScope ps(c, PBF_PH_HALO);
  plan_migration_kernel<<<1, 32, 0, c->stream>>>(arena_ptr<uint32_t>(d->arena, l.rows[0]), d->rank, W, l.row_words, l, d->dyn);
  PBF_LAUNCH_CHECK(c);
  if (W > 1) {
    const uint32_t nblk = d->cnt_nblk;  // as counted in phase A (a plan step may have grown the arena since)
    PBF_TRY(exclusive_scan_rows_u32(c, d->blk_cnt.p, d->blk_cnt.p, (uint32_t)W, nblk, d->row_tot.p));
    ghost_scatter_kernel<<<nblk, kBlk, 0, c->stream>>>(d->mask.p, &d->dyn->n_in, W, nblk, d->blk_cnt.p, d->row_tot.p, d->leave_idx.p, (uint32_t)d->leave_idx.cap);
    PBF_LAUNCH_CHECK(c);
    push_migrants_kernel<<<nblk, kBlk, 0, c->stream>>>(d->dyn, d->peers, l, c->cur, c->cur_col, W, d->leave_idx.p, c->pos[c->cur].p + l.own_off,
                                                       c->vel[c->cur].p + l.own_off, c->col[c->cur_col].p + l.own_off,
                                                       c->ids[c->cur].p + l.own_off, c->key_in.p);
    PBF_LAUNCH_CHECK(c);
  }
  return PBF_OK;
}

// after barrier 2: ONE stable sort of [arrivals from lower ranks | kept | arrivals from higher ranks]; then the ghost
// destinations of every owned particle and their count row -> every peer
int phase_c(pbf_ctx *c) {
  D *d = c->dist;
  const int W = d->world, r = d->rank;
  const ArenaLayout &l = d->lay;
  PBF_CUDA(c, cudaSetDevice(c->device));
  const uint32_t n_cap = l.cap_local - l.own_off;
  const uint32_t nblk = d->kept_nblk;  // as counted in phase A (a plan step may have grown the arena since)
  uint32_t *k2 = arena_ptr<uint32_t>(d->arena, l.k2);
  {
    PhaseScope ps(c, PBF_PH_HALO);
    // kept = mask 0, in input order: per-tile counts from phase A (role_cnt), scanned here, scattered into the merge input
    PBF_TRY(exclusive_scan_u32(c, d->role_cnt.p, d->role_cnt.p, nblk, nullptr));
    // the grid covers the CURRENT capacity: after a re-grow in this step the arrivals may reach beyond the tiles phase A
    // counted (those extra tiles hold no kept particle and read no offset)
    merge_scatter_kernel<<<std::max(nblk, div_up(n_cap, kBlk)), kBlk, 0, c->stream>>>(d->dyn, d->mask.p, d->role_cnt.p, c->key_in.p, k2, d->v2.p);
    PBF_LAUNCH_CHECK(c);
  }
  PBF_TRY(radix_sort_pairs(c, k2, n_cap, d->v2.p, &d->dyn->n_own));
  {
    PhaseScope ps(c, PBF_PH_HALO);
    PBF_CUDA(c, cudaMemsetAsync(d->d_row.p, 0, (W + 2) * 4, c->stream));
    const uint32_t oblk = d->cnt_nblk = div_up(l.cap_own, kBlk);
    ghost_mask_kernel<<<oblk, kBlk, 0, c->stream>>>(c->keys_sorted, d->dyn, c->sc.G, r, W, d->d_splits.p, d->mask.p);
    PBF_LAUNCH_CHECK(c);
    ghost_count_kernel<<<oblk, kBlk, 0, c->stream>>>(d->mask.p, &d->dyn->n_own, W, oblk, d->blk_cnt.p, d->d_row.p);
    PBF_LAUNCH_CHECK(c);
    push_row_kernel<<<1, 256, 0, c->stream>>>(d->peers, l.rows[1], r, W, l.row_words, d->d_row.p, nullptr);
    PBF_LAUNCH_CHECK(c);
  }
  return PBF_OK;
}

// after barrier 3: local layout on the device, reorder the owned particles into it, push the ghost payload
int phase_d(pbf_ctx *c) {
  D *d = c->dist;
  const int W = d->world, r = d->rank;
  const ArenaLayout &l = d->lay;
  PBF_CUDA(c, cudaSetDevice(c->device));
  const int o = c->cur ^ 1, oc = c->cur_col ^ 1;
  uint32_t *keys_local = arena_ptr<uint32_t>(d->arena, l.keys_local);
  {
    PhaseScope ps(c, PBF_PH_HALO);
    plan_ghosts_kernel<<<1, 32, 0, c->stream>>>(arena_ptr<uint32_t>(d->arena, l.rows[1]), arena_ptr<uint32_t>(d->arena, l.rows[0]), r, W,
                                                l.row_words, l, (uint32_t)d->send_idx.cap, d->send_plan, d->dyn);
    PBF_LAUNCH_CHECK(c);
  }
  c->sc.n = l.cap_own;
  c->sc.n_dyn = &d->dyn->n_own;
  PBF_TRY(launch_reorder(c, c->perm, c->pos[c->cur].p + l.own_off, c->vel[c->cur].p + l.own_off, c->col[c->cur_col].p + l.own_off,
                         c->ids[c->cur].p + l.own_off, c->pos[o].p + l.own_off, c->vel[o].p + l.own_off, c->col[oc].p + l.own_off,
                         c->ids[o].p + l.own_off, c->pstar[0].p + l.own_off));
  {
    PhaseScope ps(c, PBF_PH_HALO);
    if (W > 1) {
      const uint32_t oblk = d->cnt_nblk;  // as counted in phase C
      PBF_TRY(exclusive_scan_rows_u32(c, d->blk_cnt.p, d->blk_cnt.p, (uint32_t)W, oblk, d->row_tot.p));
      ghost_scatter_kernel<<<oblk, kBlk, 0, c->stream>>>(d->mask.p, &d->dyn->n_own, W, oblk, d->blk_cnt.p, d->row_tot.p, d->send_idx.p, (uint32_t)d->send_idx.cap);
      PBF_LAUNCH_CHECK(c);
      push_ghosts_kernel<<<div_up(d->send_idx.cap, kBlk), kBlk, 0, c->stream>>>(d->dyn, d->peers, l, oc, W, d->send_idx.p, c->pstar[0].p,
                                                                                c->pos[o].p, c->col[oc].p, c->keys_sorted);
      PBF_LAUNCH_CHECK(c);
    }
  }
  c->cur = o;
  c->cur_col = oc;
  return PBF_OK;
}

// after barrier 4: ghosts are in place -> cell table, role lists, diffuse
int phase_e(pbf_ctx *c) {
  D *d = c->dist;
  const ArenaLayout &l = d->lay;
  PBF_CUDA(c, cudaSetDevice(c->device));
  uint32_t *keys_local = arena_ptr<uint32_t>(d->arena, l.keys_local);
  c->sc.n = l.cap_local;
  c->sc.n_dyn = nullptr;
  c->keys_sorted = keys_local;
  const uint32_t nblk = div_up(l.cap_local, kBlk);
  {
    PhaseScope ps(c, PBF_PH_HALO);
    ghost_fix_kernel<<<div_up((uint64_t)2 * l.cap_g, kBlk) + 1, kBlk, 0, c->stream>>>(d->dyn, l, c->pstar[0].p, c->pos[c->cur].p, keys_local);
    PBF_LAUNCH_CHECK(c);
  }
  PBF_TRY(launch_cell_table(c, keys_local, c->table.p, d->dyn->local_range));
  {
    PhaseScope ps(c, PBF_PH_HALO);
    roles_kernel<<<nblk, kBlk, 0, c->stream>>>(keys_local, d->mask.p, d->dyn, l, c->sc.G, d->splits[d->rank], d->splits[d->rank + 1], nblk,
                                               d->role.p, d->role_cnt.p);
    PBF_LAUNCH_CHECK(c);
    PBF_TRY(exclusive_scan_rows_u32(c, d->role_cnt.p, d->role_cnt.p, 2u, nblk, d->row_tot.p));
    role_lists_kernel<<<nblk, kBlk, 0, c->stream>>>(d->role.p, d->dyn, nblk, d->role_cnt.p, d->row_tot.p, d->ring1_idx.p, d->bnd_idx.p,
                                                    2u * l.cap_g, l.cap_own);
    PBF_LAUNCH_CHECK(c);
  }
  {  // colour diffusion beside the solver iterations (as in the single-device step); group_step joins before finalise
    PBF_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
    PBF_CUDA(c, cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
    cudaStream_t main_stream = c->stream;
    c->stream = c->side_stream;
    c->diffuse_local_range = d->dyn->local_range;
    const int rc = launch_diffuse_tiled(c, keys_local, c->table.p, c->col[c->cur_col].p, c->col[c->cur_col ^ 1].p, d->dyn->own_range);
    c->diffuse_local_range = nullptr;
    c->stream = main_stream;
    PBF_TRY(rc);
    PBF_CUDA(c, cudaEventRecord(c->ev_join, c->side_stream));
    d->diffuse_pending = true;
  }
  c->cur_col ^= 1;
  return PBF_OK;
}

// Growth decision of a plan step, the same on every rank (same inputs, same arithmetic).
// A capacity grows as soon as less than the watermarks of plan_ghosts_kernel would be left free (owned: the need plus a
// quarter; ghosts: plus two fifths), and then to 1.5 x the need: right after a plan step no rank is `hot`.  0 = no opinion.
bool need_growth(const ArenaLayout &l, uint64_t want_own, uint64_t want_g, uint32_t &cap_g, uint32_t &cap_own) {
  cap_g = l.cap_g;
  cap_own = l.cap_own;
  bool grow = false;
  if (want_own && want_own + want_own / 4 > l.cap_own) { cap_own = headroom(want_own); grow = true; }
  if (want_g && want_g + 2 * want_g / 5 > l.cap_g) { cap_g = headroom(want_g); grow = true; }
  return grow;
}

int read_matrix(pbf_ctx *c, int which, uint32_t *host) {
  D *d = c->dist;
  PBF_CUDA(c, cudaSetDevice(c->device));
  PBF_CUDA(c, cudaMemcpyAsync(host, d->arena + d->lay.rows[which], (size_t)d->world * d->lay.row_words * 4, cudaMemcpyDeviceToHost, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));
  return PBF_OK;
}

// re-allocate the arenas in mid-step (plan steps only) without losing the count matrices the peers have stored there
int regrow(std::vector<pbf_ctx *> &L, uint32_t cap_g, uint32_t cap_own, const std::vector<uint64_t> &live, int matrices,
           const pbf_params &p) {
  const int W = L[0]->dist->world;
  const size_t words = (size_t)W * (W + 2);
  std::vector<std::vector<uint32_t>> keep(L.size());
  for (size_t r = 0; r < L.size(); ++r) {
    keep[r].resize(2 * words);
    for (int m = 0; m < matrices; ++m) PBF_TRY(read_matrix(L[r], m, keep[r].data() + m * words));
  }
  PBF_TRY(build_arenas(L, cap_g, cap_own, live));
  for (size_t r = 0; r < L.size(); ++r) {
    pbf_ctx *c = L[r];
    PBF_CUDA(c, cudaSetDevice(c->device));
    for (int m = 0; m < matrices; ++m)
      PBF_CUDA(c, cudaMemcpy(c->dist->arena + c->dist->lay.rows[m], keep[r].data() + m * words, words * 4, cudaMemcpyHostToDevice));
    set_grid_constants(c, p);
    c->sc.n_dyn = &c->dist->dyn->n_in;
  }
  return PBF_OK;
}

// Marching cubes of the whole fluid (ompsph.hpp:277-477) on the slab path.  A lattice point is evaluated by the rank that
// owns its cell (clamped into the grid): that rank holds every particle of the 27 cells the point looks at — its own and
// ring-1 ghosts — in the single-device order, so the sums are the single-device sums.  Ghosts need their FINAL positions
// and DIFFUSED colours first (one more 32-byte-per-ghost push).  All ranks store their points straight into rank 0's
// lattice (peer memory); rank 0 counts, scans and emits the triangles: the mesh is the single-device mesh, bit for bit.
int slab_surface(std::vector<pbf_ctx *> &L, const pbf_params &p) {
  D *d0 = L[0]->dist;
  const int W = d0->world;
  if (W > 1) {
    PBF_TRY(barrier(L, 0));  // every rank is past its iterations, its diffusion and its finalise
    for (pbf_ctx *c : L) {
      D *d = c->dist;
      const ArenaLayout &l = d->lay;
      PBF_CUDA(c, cudaSetDevice(c->device));
      PhaseScope ps(c, PBF_PH_HALO);
      push_surface_kernel<<<div_up(d->send_idx.cap, kBlk), kBlk, 0, c->stream>>>(d->dyn, d->peers, l, c->cur, c->cur_col, W, p.scale, d->send_idx.p,
                                                                                c->pstar[0].p, c->pos[c->cur].p, c->col[c->cur_col].p);
      PBF_LAUNCH_CHECK(c);
    }
    PBF_TRY(barrier(L, 0));
  }
  for (pbf_ctx *c : L) {
    D *d = c->dist;
    const ArenaLayout &l = d->lay;
    PBF_CUDA(c, cudaSetDevice(c->device));
    if (c->mc.lattice_n > l.cap_lattice) return fail(c, PBF_ERR_STATE, "slab surface", "lattice larger than the arena's (plan step missed)");
    char *root = d->peers.base[0];  // rank 0's arena as this rank addresses it
    const uint32_t lo = W > 1 ? d->splits[d->rank] : 0u, hi = W > 1 ? d->splits[d->rank + 1] : 0xFFFFFFFFu;
    c->sc.n = l.cap_local;
    c->sc.n_dyn = nullptr;
    PBF_TRY(mc_field(c, c->table.p, c->pos[c->cur].p, c->col[c->cur_col].p, arena_ptr<float4>(root, l.mc_pn), arena_ptr<float4>(root, l.mc_lc), lo, hi));
  }
  if (W > 1) PBF_TRY(barrier(L, 0));
  for (pbf_ctx *c : L) {
    D *d = c->dist;
    PBF_CUDA(c, cudaSetDevice(c->device));
    c->n_triangles = 0;
    c->mc_valid = false;
    if (d->rank != 0) continue;
    c->mc_lattice_pn = arena_ptr<float4>(d->arena, d->lay.mc_pn);
    c->mc_lattice_c = arena_ptr<float4>(d->arena, d->lay.mc_lc);
    PBF_TRY(mc_extract(c, c->mc_lattice_pn, c->mc_lattice_c));
  }
  return PBF_OK;
}

int group_step(std::vector<pbf_ctx *> &L, const pbf_params &p) {
  D *d0 = L[0]->dist;
  const int W = d0->world;
  for (pbf_ctx *c : L) {
    if (c->flags & PBF_FLAG_GLOBAL_NEIGHBOURS) return fail(c, PBF_ERR_STATE, "pbf_dist_step", "PBF_FLAG_GLOBAL_NEIGHBOURS is single-device only");
    host_grid(c->h, p, c->grid);
    for (int a = 0; a < 3; ++a)
      if (c->grid.extent[a] == 0 || c->grid.extent[a] > 1023)
        return fail(c, PBF_ERR_INVALID, "grid", "extent must be 1..1023 cells per axis (10-bit Morton, curves.h:73)");
    if (p.iteration > 0x7FFFFFFFull) return fail(c, PBF_ERR_INVALID, "params", "iteration");
  }
  bool any_fresh = false;
  for (pbf_ctx *c : L) any_fresh |= c->dist->fresh;
  // marching cubes: the lattice lives in the arena (rank 0's is the one everybody stores into); every rank derives the
  // same size from the same parameters, so a larger lattice makes this a plan step on every rank without talking
  bool lattice_grow = false;
  if (p.surface_enabled) {
    for (pbf_ctx *c : L) {
      PBF_TRY(mc_prepare(c, p));
      if (c->mc.lattice_n > c->dist->lay.cap_lattice) {
        c->dist->want_lattice = c->mc.lattice_n + c->mc.lattice_n / 4;
        lattice_grow = true;
      }
    }
  }
  // NCCL ranks agree on "is this a plan step" without talking: uploads are collective by contract and the schedule is a
  // function of the step index.
  // ... and of the `hot` verdict of the step two back (identical on every rank; a verdict older than the last plan is stale).
  bool hot = false;
  if (d0->step_index >= 2 && d0->step_index - 2 > d0->last_plan_step) {
    for (pbf_ctx *c : L) {
      D *d = c->dist;
      const int slot = (int)((d->step_index - 2) % D::kHotRing);
      PBF_CUDA(c, cudaSetDevice(c->device));
      PBF_CUDA(c, cudaEventSynchronize(d->ev_step[slot]));
      hot |= d->h_hot[slot] != 0u;
    }
  }
  const bool replan = d0->splits.empty() || hot || (d0->replan_every && d0->step_index % d0->replan_every == 0);
  const bool plan = replan || any_fresh || !d0->arena || lattice_grow;
  const bool measure = d0->replan_every && (d0->step_index + 1) % d0->replan_every == 0 && p.iteration <= (uint64_t)kMaxTimedIters;

  if (plan) {
    // ---- capacities for the particles that are about to enter the arenas ----------------------------------------
    std::vector<uint64_t> held(L.size()), all_held;
    for (size_t r = 0; r < L.size(); ++r) {
      D *d = L[r]->dist;
      if (d->fresh) held[r] = d->n_up;
      else {
        PBF_CUDA(L[r], cudaSetDevice(L[r]->device));
        PBF_CUDA(L[r], cudaStreamSynchronize(L[r]->stream));
        held[r] = d->h_dyn->n_own;
      }
    }
    PBF_TRY(all_gather_host(L, held, all_held));
    uint64_t max_held = 0, total = 0;
    for (uint64_t h : all_held) { max_held = std::max(max_held, h); total += h; }
    const uint64_t want_own = std::max<uint64_t>(max_held, (total + W - 1) / W) + 256;
    const uint64_t want_g = d0->arena ? 0 : std::max<uint64_t>(1024, want_own / 3);
    uint32_t cap_g, cap_own;
    if (need_growth(d0->lay, want_own, want_g, cap_g, cap_own) || !d0->arena || lattice_grow) {
      std::vector<uint64_t> live(L.size(), 0);
      for (size_t r = 0; r < L.size(); ++r) live[r] = L[r]->dist->fresh ? 0 : held[r];
      PBF_TRY(build_arenas(L, std::max(cap_g, 1u), std::max(cap_own, 1u), live));
    }
  }
  std::vector<int> span(L.size(), -1);
  for (size_t r = 0; r < L.size(); ++r) {
    PBF_CUDA(L[r], cudaSetDevice(L[r]->device));
    span[r] = prof_begin(L[r], PBF_PH_SLAB_SETUP, L[r]->stream);
  }
  for (pbf_ctx *c : L) PBF_TRY(phase_a(c, p, replan));
  if (replan) {
    PBF_TRY(all_reduce_sum_u32(L, [](pbf_ctx *c) { return c->dist->d_hist.p; }, L[0]->dist->hist_buckets));
    // what every rank measured on the step before: lambda-pass milliseconds against the work the last plan gave it
    bool shared_device = false;  // ranks that share a GPU time each other's kernels: no feedback from such a group
    for (size_t a = 0; a < L.size(); ++a)
      for (size_t b = a + 1; b < L.size(); ++b) shared_device |= L[a]->device == L[b]->device;
    if (!d0->splits.empty() && !d0->last_weights.empty() && !shared_device) {
      std::vector<uint64_t> mine(L.size()), all;
      for (size_t r = 0; r < L.size(); ++r) mine[r] = (uint64_t)(L[r]->dist->busy_ms * 1e6);
      PBF_TRY(all_gather_host(L, mine, all));
      std::vector<double> work(W, 0.0), rate(W, 1.0);
      for (uint32_t b = 0; b < d0->last_weights.size(); ++b)
        work[owner_host(d0->splits, W, std::min<uint64_t>((uint64_t)b << d0->hist_shift, kKeyEnd - 1))] += (double)d0->last_weights[b];
      double mean = 0.0;
      int cnt = 0;
      for (int r = 0; r < W; ++r)
        if (all[r] > 0 && work[r] > 0) { rate[r] = (double)all[r] / work[r]; mean += rate[r]; ++cnt; }
      if (cnt == W && mean > 0) {
        mean /= W;
        for (int r = 0; r < W; ++r) {
          const double prev = d0->rate.size() == (size_t)W ? d0->rate[r] : 1.0;
          rate[r] = std::min(1.7, std::max(0.6, prev * std::pow(rate[r] / mean, 0.75)));  // damped: regions shift with the splits
        }
        for (pbf_ctx *c : L) c->dist->rate = rate;
      }
    }
    for (pbf_ctx *c : L) { c->dist->busy_ms = 0.0; PBF_TRY(phase_plan(c, c->dist->rate)); }
  }
  for (pbf_ctx *c : L) PBF_TRY(phase_a2(c));
  PBF_TRY(barrier(L, 0));
  if (plan) {
    // exact migration counts: make room for the arrivals before anybody stores them
    uint64_t want_own = 0;
    std::vector<uint64_t> live(L.size(), 0);
    for (size_t r = 0; r < L.size(); ++r) {
      pbf_ctx *c = L[r];
      uint32_t *M = c->dist->h_pinned;
      PBF_TRY(read_matrix(c, 0, M));
      const uint32_t RW = c->dist->lay.row_words;
      for (int q = 0; q < W; ++q) {
        uint64_t in = M[q * RW + W + 1], own = in;
        for (int t = 0; t < W; ++t)
          if (t != q) { in += M[t * RW + q]; own = own - M[q * RW + t] + M[t * RW + q]; }
        want_own = std::max(want_own, std::max(in, own));
      }
      live[r] = M[c->dist->rank * RW + W + 1];
    }
    uint32_t cap_g, cap_own;
    if (need_growth(d0->lay, want_own + 256, 0, cap_g, cap_own)) PBF_TRY(regrow(L, cap_g, cap_own, live, 1, p));
  }
  for (pbf_ctx *c : L) PBF_TRY(phase_b(c));
  PBF_TRY(barrier(L, 0));
  for (pbf_ctx *c : L) PBF_TRY(phase_c(c));
  PBF_TRY(barrier(L, 0));
  if (plan) {
    // exact ghost counts: make room before anybody stores them (the owned particles still sit in the INPUT arrays)
    uint64_t want_g = 0, want_send = 0;
    std::vector<uint64_t> live(L.size(), 0);
    for (size_t r = 0; r < L.size(); ++r) {
      pbf_ctx *c = L[r];
      uint32_t *M = c->dist->h_pinned, *GC = c->dist->h_pinned + (size_t)(W + 1) * (W + 2);
      PBF_TRY(read_matrix(c, 1, GC));
      const uint32_t RW = c->dist->lay.row_words;
      for (int q = 0; q < W; ++q) {
        uint64_t lo = 0, hi = 0, send = 0;
        for (int s = 0; s < W; ++s) {
          if (s < q) lo += GC[s * RW + q];
          if (s > q) hi += GC[s * RW + q];
          if (s != q) send += GC[q * RW + s];
        }
        want_g = std::max(want_g, std::max(lo, hi));
        want_send = std::max(want_send, send);
      }
      uint64_t in = M[c->dist->rank * RW + W + 1];
      for (int t = 0; t < W; ++t)
        if (t != c->dist->rank) in += M[t * RW + c->dist->rank];
      live[r] = in;
    }
    uint32_t cap_g, cap_own;
    if (need_growth(d0->lay, 0, want_g + 256, cap_g, cap_own)) PBF_TRY(regrow(L, cap_g, cap_own, live, 2, p));
    for (pbf_ctx *c : L) {  // send lists: local buffers, every rank sized for the largest list of the group
      PBF_CUDA(c, cudaSetDevice(c->device));
      if (want_send + 2 * want_send / 5 + 256 > c->dist->send_plan) c->dist->send_plan = headroom(want_send);
      if (c->dist->send_plan > c->dist->send_idx.cap) {
        PBF_CUDA(c, cudaStreamSynchronize(c->stream));
        PBF_CUDA(c, c->dist->send_idx.reserve(c->dist->send_plan));
      }
    }
  }
  for (pbf_ctx *c : L) PBF_TRY(phase_d(c));
  PBF_TRY(barrier(L, 0));
  for (pbf_ctx *c : L) PBF_TRY(phase_e(c));
  for (size_t r = 0; r < L.size(); ++r) {
    PBF_CUDA(L[r], cudaSetDevice(L[r]->device));
    prof_end(L[r], span[r], L[r]->stream);
    span[r] = prof_begin(L[r], PBF_PH_SLAB_ITERATIONS, L[r]->stream);
  }

  for (uint64_t it = 0; it < p.iteration; ++it) {
    const bool exchange = W > 1 && it + 1 < p.iteration;
    const int parity = (int)(it & 1);
    for (pbf_ctx *c : L) {
      D *d = c->dist;
      const ArenaLayout &l = d->lay;
      PBF_CUDA(c, cudaSetDevice(c->device));
      float *rho = it + 1 == p.iteration ? c->rho.p : nullptr;
      Sel lam, bnd, inter;  // owned particles + ring-1 ghosts; boundary list; interior list
      lam.first = l.own_off; lam.count_dev = &d->dyn->n_own; lam.idx = d->ring1_idx.p; lam.n_idx_dev = &d->dyn->n_ring1;
      lam.bound = l.cap_own + 2u * l.cap_g;
      bnd.idx = d->bnd_idx.p; bnd.n_idx_dev = &d->dyn->n_boundary; bnd.bound = l.cap_own;
      inter.first = l.own_off; inter.count_dev = &d->dyn->n_own; inter.skip = d->mask.p; inter.bound = l.cap_own;  // mask != 0: boundary
      if (measure) PBF_CUDA(c, cudaEventRecord(d->ev_lam[2 * it], c->stream));
      PBF_TRY(solver_lambda(c, lam, c->pstar[0].p, c->pstar[1].p, rho));
      if (measure) PBF_CUDA(c, cudaEventRecord(d->ev_lam[2 * it + 1], c->stream));
      if (exchange) {
        // The boundary particles' delta pass, their halo push and the barrier run on the (high-priority) comm stream
        // BESIDE the interior pass: both passes read pStar[1] and write disjoint particles of pStar[0].
        PBF_CUDA(c, cudaEventRecord(d->ev_boundary, c->stream));
        PBF_CUDA(c, cudaStreamWaitEvent(d->comm_stream, d->ev_boundary, 0));
        cudaStream_t main_stream = c->stream;
        c->stream = d->comm_stream;
        const int rc = solver_delta(c, bnd, c->pstar[1].p, c->pstar[0].p);
        c->stream = main_stream;
        PBF_TRY(rc);
        PhaseScope ps(c, PBF_PH_HALO, d->comm_stream);
        push_halo_kernel<<<div_up(d->send_idx.cap, kBlk), kBlk, 0, d->comm_stream>>>(d->dyn, d->peers, l, parity, W, d->send_idx.p, c->pstar[0].p);
        PBF_LAUNCH_CHECK(c);
        PBF_TRY(solver_delta(c, inter, c->pstar[1].p, c->pstar[0].p));
      } else {
        Sel all;  // last iteration: nothing leaves, one pass over the owned range
        all.first = l.own_off; all.count_dev = &d->dyn->n_own; all.bound = l.cap_own;
        PBF_TRY(solver_delta(c, all, c->pstar[1].p, c->pstar[0].p));
      }
    }
    if (exchange) {
      PBF_TRY(barrier(L, 1));
      for (pbf_ctx *c : L) {
        D *d = c->dist;
        const ArenaLayout &l = d->lay;
        PBF_CUDA(c, cudaSetDevice(c->device));
        PBF_CUDA(c, cudaEventRecord(d->ev_halo, d->comm_stream));
        PBF_CUDA(c, cudaStreamWaitEvent(c->stream, d->ev_halo, 0));
        PhaseScope ps(c, PBF_PH_HALO);
        unpack_halo_kernel<<<div_up((uint64_t)2 * l.cap_g, kBlk), kBlk, 0, c->stream>>>(d->dyn, l, arena_ptr<float4>(d->arena, l.halo[parity]), c->pstar[0].p);
        PBF_LAUNCH_CHECK(c);
      }
    }
  }
  for (size_t r = 0; r < L.size(); ++r) {
    PBF_CUDA(L[r], cudaSetDevice(L[r]->device));
    prof_end(L[r], span[r], L[r]->stream);
  }
  for (pbf_ctx *c : L) {
    D *d = c->dist;
    const ArenaLayout &l = d->lay;
    PBF_CUDA(c, cudaSetDevice(c->device));
    if (d->diffuse_pending) {
      PBF_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
      d->diffuse_pending = false;
    }
    c->sc.n = l.cap_own;
    c->sc.n_dyn = &d->dyn->n_own;
    PBF_TRY(launch_finalise(c, c->pstar[0].p + l.own_off, c->pos[c->cur].p + l.own_off, c->vel[c->cur].p + l.own_off));
    c->sc.n = l.cap_local;
    c->sc.n_dyn = nullptr;
    c->n_triangles = 0;
    c->mc_valid = false;
  }
  if (p.surface_enabled) PBF_TRY(slab_surface(L, p));
  for (pbf_ctx *c : L) {
    D *d = c->dist;
    PBF_CUDA(c, cudaSetDevice(c->device));
    // the next step starts from this step's owned particles; the mirror serves statistics, downloads and plan steps
    next_step_kernel<<<1, 32, 0, c->stream>>>(d->dyn);
    PBF_LAUNCH_CHECK(c);
    PBF_CUDA(c, cudaMemcpyAsync(d->h_dyn, d->dyn, sizeof(SlabDyn), cudaMemcpyDeviceToHost, c->stream));
    {
      const int slot = (int)(d->step_index % D::kHotRing);
      PBF_CUDA(c, cudaMemcpyAsync(d->h_hot + slot, &d->dyn->hot, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      PBF_CUDA(c, cudaEventRecord(d->ev_step[slot], c->stream));
      if (plan) {
        d->last_plan_step = d->step_index;
        d->n_plans++;
        if (hot) d->n_early_plans++;
      }
    }
    c->have_state = true;
    c->prof.steps++;
    d->step_index++;
    d->n_timed = measure ? (int)p.iteration : 0;
  }
  // Load-balance feedback: on the step before a re-plan the lambda launches of every rank are timed (reading the events
  // waits for the step, which the plan step would do a moment later anyway).
  if (measure) {
    for (pbf_ctx *c : L) {
      D *d = c->dist;
      PBF_CUDA(c, cudaSetDevice(c->device));
      PBF_CUDA(c, cudaStreamSynchronize(c->stream));
      double sum = 0.0;
      for (int k = 0; k < d->n_timed; ++k) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, d->ev_lam[2 * k], d->ev_lam[2 * k + 1]) == cudaSuccess) sum += ms;
      }
      d->busy_ms = sum;
    }
  }
  return PBF_OK;
}

int dist_alloc(pbf_ctx *c, int rank, int world) {
  if (c->dist) return fail(c, PBF_ERR_STATE, "pbf_dist_init", "already initialised");
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(c, PBF_ERR_INVALID, "pbf_dist_init", "rank/world (world <= 32)");
  D *d = new D();
  d->rank = rank;
  d->world = world;
  c->dist = d;
  PBF_CUDA(c, cudaSetDevice(c->device));
  PBF_CUDA(c, d->d_splits.reserve(world + 2));
  PBF_CUDA(c, d->d_row.reserve(world + 4));
  PBF_CUDA(c, d->row_tot.reserve(kMaxWorld));
  PBF_CUDA(c, d->d_scratch.reserve((size_t)16 * (world + 1) + 64));
  PBF_CUDA(c, cudaMalloc(&d->dyn, sizeof(SlabDyn)));
  PBF_CUDA(c, cudaMemset(d->dyn, 0, sizeof(SlabDyn)));
  PBF_CUDA(c, cudaHostAlloc(&d->h_dyn, sizeof(SlabDyn), cudaHostAllocDefault));
  std::memset(d->h_dyn, 0, sizeof(SlabDyn));
  PBF_CUDA(c, cudaHostAlloc(&d->h_hot, D::kHotRing * sizeof(uint32_t), cudaHostAllocDefault));
  std::memset(d->h_hot, 0, D::kHotRing * sizeof(uint32_t));
  for (cudaEvent_t &e : d->ev_step) PBF_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  PBF_CUDA(c, cudaMalloc(&d->bar_word, 64));
  PBF_CUDA(c, cudaMemset(d->bar_word, 0, 64));
  PBF_CUDA(c, cudaHostAlloc(&d->h_pinned, ((size_t)2 * (world + 1) * (world + 2) + 64) * 4, cudaHostAllocDefault));
  {  // the comm stream's small kernels (boundary delta, halo push, barrier) must not queue behind the interior pass's blocks
    int lo = 0, hi = 0;
    PBF_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    PBF_CUDA(c, cudaStreamCreateWithPriority(&d->comm_stream, cudaStreamNonBlocking, hi));
  }
  PBF_CUDA(c, cudaEventCreateWithFlags(&d->ev_boundary, cudaEventDisableTiming));
  PBF_CUDA(c, cudaEventCreateWithFlags(&d->ev_halo, cudaEventDisableTiming));
  PBF_CUDA(c, cudaEventCreateWithFlags(&d->ev_bar[0], cudaEventDisableTiming));
  PBF_CUDA(c, cudaEventCreateWithFlags(&d->ev_bar[1], cudaEventDisableTiming));
  for (cudaEvent_t &e : d->ev_lam) PBF_CUDA(c, cudaEventCreate(&e));
  // every rank flips its double buffers in lockstep: peers address each other's "current" arrays by parity
  c->cur = 0;
  c->cur_col = 0;
  return PBF_OK;
}

}  // namespace

namespace pbf {

// the mirror of the device-side counts, after waiting for the step that wrote it
int dist_refresh_counts(pbf_ctx *c) {
  D *d = c->dist;
  PBF_CUDA(c, cudaSetDevice(c->device));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));
  if (d->fresh) { c->n = d->n_up; return PBF_OK; }
  if (d->h_dyn->overflow) {
    const SlabDyn &h = *d->h_dyn;
    char msg[512];
    std::snprintf(msg, sizeof msg,
                  "%s (rank %d, bits 0x%x: ghosts below/above %u/%u of %u, send list %u of %zu, owned %u of %u, ring-1 %u); "
                  "a capacity set by the last plan step was exceeded",
                  h.overflow & 8u   ? "a peer rank did not reach a barrier within 4 s"
                  : h.overflow & 1u ? "more particles arrived between two plan steps than the arena holds"
                                    : "more ghosts between two plan steps than the arena holds",
                  d->rank, h.overflow, h.want[0], h.want[1], d->lay.cap_g, h.want[2], d->send_idx.cap, h.n_own, d->lay.cap_own, h.n_ring1);
    return fail(c, PBF_ERR_CAPACITY, "slab arena", msg);
  }
  c->n = d->h_dyn->n_own;
  return PBF_OK;
}

void dist_release(pbf_ctx *ctx) {
  if (!ctx->dist) return;
  D *d = ctx->dist;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (d->comm_stream) cudaStreamSynchronize(d->comm_stream);
  // the context's buffers point into the arena: forget them before it goes
  for (int i = 0; i < 2; ++i) {
    for (DevBuf<float4> *b : {&ctx->pos[i], &ctx->vel[i], &ctx->col[i], &ctx->pstar[i]})
      if (b->borrowed) b->release();
    if (ctx->ids[i].borrowed) ctx->ids[i].release();
  }
  if (ctx->key_a.borrowed) ctx->key_a.release();
  if (d->comm && g_nccl.handle) g_nccl.CommDestroy(d->comm);
  d->release();
  delete d;
  ctx->dist = nullptr;
}
}  // namespace pbf

extern "C" {

int pbf_dist_unique_id(uint8_t *id) {
  if (!id) return PBF_ERR_INVALID;
  if (!g_nccl.load()) return fail(nullptr, PBF_ERR_NCCL, "pbf_dist_unique_id", g_nccl.err.c_str());
  static_assert(sizeof(ncclUniqueId) == PBF_NCCL_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u;
  ncclResult_t r = g_nccl.GetUniqueId(&u);
  if (r != ncclSuccess) return fail(nullptr, PBF_ERR_NCCL, "ncclGetUniqueId", g_nccl.GetErrorString(r));
  memcpy(id, &u, sizeof(u));
  return PBF_OK;
}

int pbf_dist_init(pbf_ctx *ctx, const uint8_t *id, int rank, int world) {
  if (!ctx || !id) return fail(ctx, PBF_ERR_INVALID, "pbf_dist_init", "NULL");
  if (!g_nccl.load()) return fail(ctx, PBF_ERR_NCCL, "pbf_dist_init", g_nccl.err.c_str());
  PBF_TRY(dist_alloc(ctx, rank, world));
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  PBF_NCCL(ctx, g_nccl.CommInitRank(&ctx->dist->comm, world, u, rank));
  ctx->dist->group = std::make_shared<std::vector<pbf_ctx *>>(1, ctx);
  return PBF_OK;
}

int pbf_dist_init_local(pbf_ctx **ctxs, int world) {
  if (!ctxs || world < 1) return fail(nullptr, PBF_ERR_INVALID, "pbf_dist_init_local", "NULL / world");
  auto group = std::make_shared<std::vector<pbf_ctx *>>(ctxs, ctxs + world);
  for (int r = 0; r < world; ++r) {
    if (!ctxs[r]) return fail(nullptr, PBF_ERR_INVALID, "pbf_dist_init_local", "NULL context");
    PBF_TRY(dist_alloc(ctxs[r], r, world));
    ctxs[r]->dist->local_mode = true;
    ctxs[r]->dist->group = group;
  }
  return PBF_OK;
}

int pbf_dist_set_replan(pbf_ctx *ctx, uint32_t steps) {
  if (!ctx || !ctx->dist) return fail(ctx, PBF_ERR_STATE, "pbf_dist_set_replan", "pbf_dist_init first");
  for (pbf_ctx *c : *ctx->dist->group) c->dist->replan_every = steps;
  return PBF_OK;
}

// The particles wait in staging buffers outside the arena: the next step (a plan step) agrees on the capacities with
// the other ranks first.  `keep_plan` (pbf_dist_advance_host) keeps the key splits and the step count of the group.
static int dist_upload(pbf_ctx *ctx, const pbf_particle *xs, uint64_t n, bool keep_plan, bool wait) {
  D *d = ctx->dist;
  PBF_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n && !xs) return fail(ctx, PBF_ERR_INVALID, "xs", "NULL");
  if (n >= 0xFFFFFFF0ull) return fail(ctx, PBF_ERR_INVALID, "n", "more than 2^32 particles on one device");
  if (!keep_plan) {
    d->splits.clear();
    d->rate.clear();
    d->last_weights.clear();
    d->step_index = 0;
  }
  PBF_CUDA(ctx, ctx->aos.reserve(n + 1));
  PBF_CUDA(ctx, d->up_pos.reserve(n + 1));
  PBF_CUDA(ctx, d->up_vel.reserve(n + 1));
  PBF_CUDA(ctx, d->up_col.reserve(n + 1));
  PBF_CUDA(ctx, d->up_ids.reserve(n + 1));
  if (n) {
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->aos.p, xs, n * sizeof(pbf_particle), cudaMemcpyHostToDevice, ctx->stream));
    PBF_CUDA(ctx, cudaMemsetAsync(ctx->flag_dev, 0, sizeof(int), ctx->stream));
    PBF_TRY(launch_unpack_aos(ctx, ctx->aos.p, n, d->up_pos.p, d->up_vel.p, d->up_col.p, d->up_ids.p, ctx->flag_dev));
    PBF_CUDA(ctx, cudaMemcpyAsync(ctx->flag_host, ctx->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  }
  d->n_up = n;
  d->fresh = true;
  ctx->n = n;
  ctx->have_state = true;
  if (wait) {
    PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n && *ctx->flag_host) {
      ctx->n = 0; ctx->have_state = false; d->fresh = false; d->n_up = 0;
      return fail(ctx, PBF_ERR_INVALID, "xs", "Obstacle particles are not supported (the reference OMP backend drops them)");
    }
  }
  return PBF_OK;
}

int pbf_dist_upload(pbf_ctx *ctx, const pbf_particle *xs, uint64_t n) {
  if (!ctx || !ctx->dist) return fail(ctx, PBF_ERR_STATE, "pbf_dist_upload", "pbf_dist_init first");
  return dist_upload(ctx, xs, n, false, true);
}

int pbf_dist_step(pbf_ctx *ctx, const pbf_params *params) {
  if (!ctx || !ctx->dist) return fail(ctx, PBF_ERR_STATE, "pbf_dist_step", "pbf_dist_init first");
  if (!params) return fail(ctx, PBF_ERR_INVALID, "params", "NULL");
  if (!(params->scale > 0.f) || !(params->dt > 0.f)) return fail(ctx, PBF_ERR_INVALID, "params", "scale and dt must be > 0");
  std::vector<pbf_ctx *> &L = *ctx->dist->group;
  for (pbf_ctx *c : L) {
    if (!c->have_state) return fail(ctx, PBF_ERR_STATE, "pbf_dist_step", "pbf_dist_upload on every rank first");
    if ((c->flags & (PBF_FLAG_XSPH | PBF_FLAG_VORTICITY)) || !c->scene.empty())
      return fail(ctx, PBF_ERR_STATE, "pbf_dist_step", "scene dynamics and the XSPH/vorticity extensions are single-device only");
  }
  const int rc = group_step(L, *params);
  if (rc != PBF_OK && ctx->err.empty())
    for (pbf_ctx *c : L)
      if (!c->err.empty()) { ctx->err = c->err; break; }
  return rc;
}

int pbf_dist_download(pbf_ctx *ctx, pbf_particle *xs, uint64_t capacity, uint64_t *n_out) {
  if (!ctx || !ctx->dist) return fail(ctx, PBF_ERR_STATE, "pbf_dist_download", "pbf_dist_init first");
  PBF_TRY(dist_refresh_counts(ctx));
  D *d = ctx->dist;
  if (n_out) *n_out = ctx->n;
  if (capacity < ctx->n) return fail(ctx, PBF_ERR_CAPACITY, "pbf_dist_download", "capacity too small");
  if (ctx->n == 0) return PBF_OK;
  if (!xs) return fail(ctx, PBF_ERR_INVALID, "xs", "NULL");
  PBF_CUDA(ctx, ctx->aos.reserve(ctx->n + 1));
  if (d->fresh) {
    PBF_TRY(launch_pack_aos(ctx, ctx->aos.p, ctx->n, d->up_pos.p, d->up_vel.p, d->up_col.p, d->up_ids.p));
  } else {
    const uint32_t off = d->lay.own_off;
    PBF_TRY(launch_pack_aos(ctx, ctx->aos.p, ctx->n, ctx->pos[ctx->cur].p + off, ctx->vel[ctx->cur].p + off,
                            ctx->col[ctx->cur_col].p + off, ctx->ids[ctx->cur].p + off));
  }
  PBF_CUDA(ctx, cudaMemcpyAsync(xs, ctx->aos.p, ctx->n * sizeof(pbf_particle), cudaMemcpyDeviceToHost, ctx->stream));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PBF_OK;
}

// sph::Solver::advance over a LOCAL group (the multi-device form of pbf_advance_host): the caller's array is cut into
// one contiguous block per rank (the previous call's per-rank counts when the particle count is unchanged, so that
// a returning array lands on the ranks that own it; equal blocks otherwise — the first step migrates anyway), uploaded
// on every rank's stream at once, stepped, and read back in rank order, which is the global Z order (ompsph.hpp:479-481).
int pbf_dist_advance_host(pbf_ctx *ctx, const pbf_params *params, pbf_particle *xs, uint64_t n, uint64_t *n_mesh_vertices) {
  if (n_mesh_vertices) *n_mesh_vertices = 0;
  if (!ctx || !ctx->dist) return fail(ctx, PBF_ERR_STATE, "pbf_dist_advance_host", "pbf_dist_init_local first");
  if (!ctx->dist->local_mode)
    return fail(ctx, PBF_ERR_STATE, "pbf_dist_advance_host", "one-process groups only (pbf_dist_init_local); NCCL ranks use pbf_dist_upload / step / download");
  if (!params) return fail(ctx, PBF_ERR_INVALID, "params", "NULL");
  if (n == 0) return PBF_OK;  // ompsph.hpp:122-126
  if (!xs) return fail(ctx, PBF_ERR_INVALID, "xs", "NULL");
  std::vector<pbf_ctx *> &L = *ctx->dist->group;
  const size_t W = L.size();
  D *d0 = L[0]->dist;
  std::vector<uint64_t> cnt(W);
  uint64_t had = 0;
  for (uint64_t c : d0->last_counts) had += c;
  if (d0->last_counts.size() == W && had == n) cnt = d0->last_counts;
  else for (size_t r = 0; r < W; ++r) cnt[r] = (n * (r + 1)) / W - (n * r) / W;
  if (L[0]->flags & PBF_FLAG_PIN_HOST) {
    PBF_CUDA(L[0], cudaSetDevice(L[0]->device));
    host_pin(L[0], xs, (size_t)n * sizeof(pbf_particle));
  }
  uint64_t off = 0;
  for (size_t r = 0; r < W; ++r) {  // every rank's H2D is enqueued before any is waited for
    PBF_TRY(dist_upload(L[r], xs + off, cnt[r], true, false));
    off += cnt[r];
  }
  for (pbf_ctx *c : L) {
    PBF_CUDA(c, cudaSetDevice(c->device));
    PBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->n && *c->flag_host) {
      for (pbf_ctx *q : L) { q->n = 0; q->have_state = false; q->dist->fresh = false; }
      return fail(ctx, PBF_ERR_INVALID, "xs", "Obstacle particles are not supported (the reference OMP backend drops them)");
    }
  }
  PBF_TRY(pbf_dist_step(ctx, params));
  off = 0;
  for (size_t r = 0; r < W; ++r) {
    pbf_ctx *c = L[r];
    const int rc = dist_refresh_counts(c);
    if (rc != PBF_OK) { if (ctx != c) ctx->err = c->err; return rc; }
    cnt[r] = c->n;
    if (off + c->n > n) return fail(ctx, PBF_ERR_STATE, "pbf_dist_advance_host", "particle count changed");
    if (c->n) {
      const uint32_t o = c->dist->lay.own_off;
      PBF_CUDA(c, c->aos.reserve(c->n + 1));
      PBF_TRY(launch_pack_aos(c, c->aos.p, c->n, c->pos[c->cur].p + o, c->vel[c->cur].p + o, c->col[c->cur_col].p + o,
                              c->ids[c->cur].p + o));
      PBF_CUDA(c, cudaMemcpyAsync(xs + off, c->aos.p, c->n * sizeof(pbf_particle), cudaMemcpyDeviceToHost, c->stream));
    }
    off += c->n;
  }
  if (off != n) return fail(ctx, PBF_ERR_STATE, "pbf_dist_advance_host", "particle count changed");
  PBF_TRY(sync_all(L));
  d0->last_counts = cnt;
  if (n_mesh_vertices) *n_mesh_vertices = L[0]->n_triangles * 3;  // marching cubes: rank 0 holds the mesh (pbf_mesh_download)
  return PBF_OK;
}

int pbf_dist_stats_read(pbf_ctx *ctx, pbf_dist_stats *out) {
  if (!ctx || !ctx->dist || !out) return fail(ctx, PBF_ERR_STATE, "pbf_dist_stats_read", "pbf_dist_init first");
  PBF_TRY(dist_refresh_counts(ctx));
  D *d = ctx->dist;
  const SlabDyn &h = *d->h_dyn;
  pbf_dist_stats s{};
  s.owned = h.n_own;
  s.ghosts = h.n_glo + h.n_ghi;
  s.migrants_out = h.total_out;
  s.migrants_in = h.total_in;
  s.halo_bytes_per_iteration = (uint64_t)h.n_send * 16;
  if (!d->splits.empty()) { s.key_lo = d->splits[d->rank]; s.key_hi = d->splits[d->rank + 1]; }
  s.ghost_ring1 = h.n_ring1;
  s.boundary = h.n_boundary;
  s.plan_steps = d->n_plans;
  s.early_plans = d->n_early_plans;
  s.capacity_owned = d->lay.cap_own;
  s.capacity_ghosts = d->lay.cap_g;
  *out = s;
  return PBF_OK;
}

int pbf_host_work_weights(const uint32_t *bucket_hist, uint32_t n_buckets, uint32_t shift, uint64_t *weights) {
  if (!bucket_hist || !weights || n_buckets == 0 || shift > 30) return PBF_ERR_INVALID;
  const double slots = (double)(1ull << shift);  // key slots (cells) per bucket
  for (uint32_t b = 0; b < n_buckets; ++b) {
    const bool beyond_grid = b + 1 == n_buckets;  // the last bucket collects every key >= G: no density there
    weights[b] = (uint64_t)bucket_hist[b] * (uint64_t)(4.4 * slots + (beyond_grid ? 0.0 : (double)bucket_hist[b]));
  }
  return PBF_OK;
}

int pbf_host_plan_splits(const uint64_t *bucket_hist, uint32_t n_buckets, uint32_t shift, int world, uint32_t *splits) {
  if (!bucket_hist || !splits || world < 1 || n_buckets == 0 || shift > 29) return PBF_ERR_INVALID;
  plan_splits(bucket_hist, n_buckets, shift, world, splits);
  return PBF_OK;
}

}  // extern "C"
