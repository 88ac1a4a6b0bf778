// dist.cu — multi-GPU PBF step: Z-curve slab decomposition, one rank per GPU (include/pbf_cuda.h "multi-GPU").
//
// The reference is single-device; this is the north-star extension.  Rank r owns the particles
// whose Morton key (curves.h:72-88, computed from the PREDICTED position exactly as ompsph.hpp:152) lies in
// [split[r], split[r+1]).  Keys and the cell table are computed once per step and reused by every solver iteration
// (ompsph.hpp:215-249), so ownership, the ghost-cell sets and the send lists are static within a step:
//
//   A  predict_key on the owned particles; (every `replan` steps) global key histogram -> new splits;
//      classify every particle by the rank owning its key
//   x1 all-gather of the per-destination counts
//   B  stable destination-major list of the leaving particles, pack; x2 all-to-all of the migrants
//      (pos, vel, colour, id, key)
//   C  ONE stable radix sort of [arrivals from lower ranks | kept | arrivals from higher ranks] — the order that
//      reproduces the single-device stable sort when the ranks' inputs are consecutive blocks of one array;
//      ghost masks: a cell is sent to every rank owning a cell within Chebyshev distance 2 (delta needs lambda of
//      ring-1 ghosts, whose lambda needs ring 2)
//   x3 all-gather of the ghost counts
//   D  reorder (gather + predict) into the local arrays [ghosts below | owned | ghosts above], which are globally
//      key-sorted because ranks own ascending key ranges; pack the ghost payload
//   x4 all-to-all of the ghosts (pStar|mass, colour, key), received in place
//   E  cell table over the local array, diffuse, then per iteration: lambda (owned + ring-1 ghosts), delta (owned),
//      x5 pStar of the ghost set (16 B each) — the once-per-iteration halo exchange; finally finalise (owned).
//
// Overlap: the delta pass runs first on the BOUNDARY particles (those some other rank holds as ghosts); their pack +
// exchange is issued on a second stream and proceeds while the interior delta pass runs on the compute stream.
//
// Transports: NCCL (ncclSend/ncclRecv groups, one process per GPU; the library is dlopen'ed so single-GPU users do
// not need it) and LOCAL (every rank is a context of this process; exchanges are device-to-device copies — this is
// what the 1-GPU parity tests drive, and it also serves one-process-many-GPUs callers).
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; the symbols are resolved with dlsym

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
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
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
    PBF_SYM(Send, ncclSend)
    PBF_SYM(Recv, ncclRecv)
    PBF_SYM(GroupStart, ncclGroupStart)
    PBF_SYM(GroupEnd, ncclGroupEnd)
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

// One logical message of an all-to-all: element offsets / counts per peer into a source and a destination array.
struct Msg {
  const char *src = nullptr;
  char *dst = nullptr;
  size_t elem = 0;
  std::vector<uint64_t> send_off, send_cnt, recv_off, recv_cnt;
  void init(int world, const void *s, void *d, size_t e) {
    src = static_cast<const char *>(s); dst = static_cast<char *>(d); elem = e;
    send_off.assign(world, 0); send_cnt.assign(world, 0); recv_off.assign(world, 0); recv_cnt.assign(world, 0);
  }
};

constexpr uint32_t kKeyEnd = 1u << 30;   // one past the largest 30-bit Morton key
constexpr int kMaxWorld = 32;            // ghost masks are 32-bit
constexpr uint32_t kHistBits = 16;       // load-balance histogram: at most 65 536 coarse key buckets

}  // namespace

struct pbf_dist_state {
  int rank = 0, world = 1;
  bool local_mode = false;
  std::shared_ptr<std::vector<pbf_ctx *>> group;  // LOCAL: every rank's context; NCCL: just this one
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;             // halo exchange overlapped with the interior delta pass
  cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr;
  uint64_t step_index = 0;
  uint32_t replan_every = 4;
  uint32_t hist_shift = 0, hist_buckets = 0;
  std::vector<uint32_t> splits;                   // world + 1 key boundaries
  uint32_t own_off = 0;                           // first owned particle in the local arrays (= ghosts below)
  uint32_t n_in = 0, n_keep = 0, n_own = 0, n_glo = 0, n_ghi = 0, n_local = 0, n_send = 0;
  uint32_t total_in = 0, total_out = 0, in_lo = 0;
  bool any_migrants = false, any_ghosts = false, any_outside = false;
  // device scratch
  DevBuf<uint32_t> d_splits, d_row, d_all, d_hist;
  DevBuf<uint32_t> mask, send_idx, blk_cnt, k2, v2, keys_local, role, sub_cnt;
  bool diffuse_pending = false;  // colour diffusion of this step is still running on the side stream
  DevBuf<float4> sb_a, sb_b, sb_c;
  DevBuf<unsigned long long> sb_id;
  DevBuf<uint32_t> sb_key;
  uint32_t *h_pinned = nullptr;  // (world + 1) * (world + 2) + 64 words
  std::vector<uint64_t> last_counts;  // pbf_dist_advance_host: particles every rank returned from the previous call
  std::vector<Msg> msgs;
  pbf_dist_stats stats{};
  void release() {
    d_splits.release(); d_row.release(); d_all.release(); d_hist.release();
    mask.release(); send_idx.release(); blk_cnt.release(); k2.release(); v2.release(); keys_local.release();
    role.release(); sub_cnt.release();
    sb_a.release(); sb_b.release(); sb_c.release(); sb_id.release(); sb_key.release();
    if (h_pinned) cudaFreeHost(h_pinned);
    if (comm_stream) cudaStreamDestroy(comm_stream);
    if (ev_boundary) cudaEventDestroy(ev_boundary);
    if (ev_halo) cudaEventDestroy(ev_halo);
  }
};

namespace {

using D = pbf_dist_state;
constexpr int kBlk = 256;

// ------------------------------------------------------------------------------------------------- kernels
__global__ void key_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t shift, uint32_t n_buckets,
                                uint32_t *__restrict__ hist) {
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  if (i >= n) return;
  const uint32_t b = min(__ldg(keys + i) >> shift, n_buckets - 1u);
  atomicAdd(hist + b, 1u);
}

// Destination of every owned particle after predict_key: mask[i] = 1 << owner when the owner is another rank, else 0
// (the same mask format as the ghost lists, so ghost_count_kernel / ghost_scatter_kernel build the leave lists);
// *n_outside counts particles predicted outside the grid (key >= G), which the last rank owns.
__global__ void classify_kernel(const uint32_t *__restrict__ keys, uint32_t n, const uint32_t *__restrict__ splits_g,
                                int rank, int world, uint32_t G, uint32_t *__restrict__ mask,
                                uint32_t *__restrict__ n_outside) {
  __shared__ uint32_t splits[kMaxWorld + 1];
  if (threadIdx.x <= (unsigned)world) splits[threadIdx.x] = splits_g[threadIdx.x];
  __syncthreads();
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t key = i < n ? __ldg(keys + i) : 0u;
  int o = 0;
  for (int d = 1; d < world; ++d) o += (key >= splits[d]) ? 1 : 0;
  if (i < n) mask[i] = o == rank ? 0u : 1u << o;
  const int outside = __syncthreads_count(i < n && key >= G);
  if (threadIdx.x == 0 && outside) atomicAdd(n_outside, (uint32_t)outside);
}

// leaving particles, destination-major (leave_idx from ghost_scatter_kernel), raw state + key
__global__ void pack_migrants_kernel(uint32_t n_leave, const uint32_t *__restrict__ leave_idx,
                                     const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                     const float4 *__restrict__ col, const unsigned long long *__restrict__ ids,
                                     const uint32_t *__restrict__ keys, float4 *__restrict__ o_pos,
                                     float4 *__restrict__ o_vel, float4 *__restrict__ o_col,
                                     unsigned long long *__restrict__ o_ids, uint32_t *__restrict__ o_keys) {
  const uint32_t j = blockIdx.x * kBlk + threadIdx.x;
  if (j >= n_leave) return;
  const uint32_t s = __ldg(leave_idx + j);
  o_pos[j] = ldg4(pos + s);
  o_vel[j] = ldg4(vel + s);
  o_col[j] = ldg4(col + s);
  o_ids[j] = __ldg(ids + s);
  o_keys[j] = __ldg(keys + s);
}

__global__ void gather_u32_kernel(uint32_t n, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ src,
                                  uint32_t *__restrict__ out) {
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  if (i < n) out[i] = __ldg(src + __ldg(idx + i));
}

__global__ void iota_kernel(uint32_t *__restrict__ out, uint32_t n, uint32_t first) {
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  if (i < n) out[i] = first + i;
}

__device__ __forceinline__ int owner_of(const uint32_t *splits, int world, uint32_t key) {
  int o = 0;
  for (int d = 1; d < world; ++d) o += (key >= splits[d]) ? 1 : 0;
  return o;
}

// Ghost destinations of every owned particle: bit d of mask[i] = rank d needs particle i as a ghost, i.e. owns a cell
// within Chebyshev distance 2 of the particle's cell.  The search is per CELL and warp-cooperative: the first particle of
// each cell is its leader; for every leader in the warp the 32 lanes split the 125 neighbour cells between them and
// OR-reduce the owners; the leader then writes the answer for its whole cell.
__global__ void ghost_mask_kernel(const uint32_t *__restrict__ keys, uint32_t n, uint32_t G, int rank, int world,
                                  const uint32_t *__restrict__ splits_g, int any_outside, uint32_t *__restrict__ mask) {
  __shared__ uint32_t splits[kMaxWorld + 1];
  if (threadIdx.x <= (unsigned)world) splits[threadIdx.x] = splits_g[threadIdx.x];
  __syncthreads();
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

// Role of every particle of the local array for the solver passes (static within a step):
//   kRoleLambda    lambda is computed here: owned particles and RING-1 ghosts (a ghost one of whose 27 cells is ours)
//   kRoleBoundary  owned, and some other rank holds it as a ghost: its delta pass runs first so the halo can leave
//   kRoleInterior  owned, nobody else needs it: its delta pass overlaps the halo exchange
// counts[0] += ring-1 ghosts, counts[1] += boundary particles (statistics only).
constexpr uint32_t kRoleLambda = 1u, kRoleBoundary = 2u, kRoleInterior = 4u;
__global__ void roles_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ owned_mask, uint32_t n_glo,
                             uint32_t n_own, uint32_t n_local, uint32_t G, uint32_t lo, uint32_t hi, int any_outside,
                             uint32_t *__restrict__ role, uint32_t *__restrict__ counts) {
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  uint32_t f = 0;
  bool ring1 = false, boundary = false;
  if (i < n_local) {
    if (i >= n_glo && i < n_glo + n_own) {
      boundary = __ldg(owned_mask + (i - n_glo)) != 0u;
      f = kRoleLambda | (boundary ? kRoleBoundary : kRoleInterior);
    } else {
      const uint32_t key = __ldg(keys + i);
      const uint32_t kx = key & kAxisMask, ky = (key >> 1) & kAxisMask, kz = (key >> 2) & kAxisMask;
      const uint32_t ax[3] = {dilated_dec(kx), kx, dilated_inc(kx)};
      const uint32_t ay[3] = {dilated_dec(ky), ky, dilated_inc(ky)};
      const uint32_t az[3] = {dilated_dec(kz), kz, dilated_inc(kz)};
      for (int z = 0; z < 3; ++z)
        for (int y = 0; y < 3; ++y)
          for (int x = 0; x < 3; ++x) {
            const uint32_t nk = (az[z] << 2) | (ay[y] << 1) | ax[x];
            if (nk >= G && !any_outside) continue;
            ring1 |= nk >= lo && nk < hi;
          }
      f = ring1 ? kRoleLambda : 0u;
    }
    role[i] = f;
  }
  const int c_ring1 = __syncthreads_count(ring1), c_boundary = __syncthreads_count(boundary);
  if (threadIdx.x == 0) {
    if (c_ring1) atomicAdd(counts, (uint32_t)c_ring1);
    if (c_boundary) atomicAdd(counts + 1, (uint32_t)c_boundary);
  }
}

// Per 256-particle tile and destination: how many particles go there (cnt[d * nblk + blk]); totals into row[d].
__global__ void ghost_count_kernel(const uint32_t *__restrict__ mask, uint32_t n, int world, uint32_t nblk,
                                   uint32_t *__restrict__ cnt, uint32_t *__restrict__ row) {
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t m = i < n ? __ldg(mask + i) : 0u;
  for (int d = 0; d < world; ++d) {
    const int c = __syncthreads_count((m >> d) & 1u);
    if (threadIdx.x == 0) {
      cnt[(uint32_t)d * nblk + blockIdx.x] = (uint32_t)c;
      if (c) atomicAdd(row + d, (uint32_t)c);
    }
  }
}

// Stable scatter of the send lists: send_idx[offs[d][blk] + rank within the tile] = i, destination-major.
__global__ void ghost_scatter_kernel(const uint32_t *__restrict__ mask, uint32_t n, int world, uint32_t nblk,
                                     const uint32_t *__restrict__ offs, uint32_t *__restrict__ send_idx) {
  __shared__ uint32_t wsum[kBlk / 32];
  const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
  const uint32_t m = i < n ? __ldg(mask + i) : 0u;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = 0; d < world; ++d) {
    const bool bit = (m >> d) & 1u;
    const unsigned b = __ballot_sync(0xFFFFFFFFu, bit);
    if (lane == 0) wsum[warp] = __popc(b);
    __syncthreads();
    uint32_t base = __ldg(offs + (uint32_t)d * nblk + blockIdx.x);
    for (unsigned w = 0; w < warp; ++w) base += wsum[w];
    if (bit) send_idx[base + __popc(b & ((1u << lane) - 1u))] = i;
    __syncthreads();
  }
}

// Generic stable compaction of "flag[i] != 0" (same tile scheme, one destination): used for the boundary / interior /
// lambda subsets.  want = 1 selects set flags, want = 0 selects clear flags; indices are offset by `first`.
__global__ void flag_count_kernel(const uint32_t *__restrict__ flag, uint32_t first, uint32_t n, uint32_t want,
                                  uint32_t *__restrict__ cnt) {
  const uint32_t t = blockIdx.x * kBlk + threadIdx.x;
  const bool bit = t < n && ((__ldg(flag + first + t) != 0u) == (want != 0u));
  const int c = __syncthreads_count(bit);
  if (threadIdx.x == 0) cnt[blockIdx.x] = (uint32_t)c;
}
__global__ void flag_scatter_kernel(const uint32_t *__restrict__ flag, uint32_t first, uint32_t n, uint32_t want,
                                    const uint32_t *__restrict__ offs, uint32_t *__restrict__ out) {
  __shared__ uint32_t wsum[kBlk / 32];
  const uint32_t t = blockIdx.x * kBlk + threadIdx.x;
  const bool bit = t < n && ((__ldg(flag + first + t) != 0u) == (want != 0u));
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(0xFFFFFFFFu, bit);
  if (lane == 0) wsum[warp] = __popc(b);
  __syncthreads();
  uint32_t base = __ldg(offs + blockIdx.x);
  for (unsigned w = 0; w < warp; ++w) base += wsum[w];
  if (bit) out[base + __popc(b & ((1u << lane) - 1u))] = first + t;
}

__global__ void pack_ghosts_kernel(uint32_t n_send, uint32_t own_off, const uint32_t *__restrict__ send_idx,
                                   const float4 *__restrict__ pstar, const float4 *__restrict__ pos_mass,
                                   const float4 *__restrict__ col, const uint32_t *__restrict__ keys_sorted,
                                   float4 *__restrict__ o_pstar, float4 *__restrict__ o_col, uint32_t *__restrict__ o_key) {
  const uint32_t j = blockIdx.x * kBlk + threadIdx.x;
  if (j >= n_send) return;
  const uint32_t i = __ldg(send_idx + j);
  float4 p = ldg4(pstar + own_off + i);
  p.w = __ldg(&pos_mass[own_off + i].w);  // the mass rides in the (still unused) lambda slot
  o_pstar[j] = p;
  o_col[j] = ldg4(col + own_off + i);
  o_key[j] = __ldg(keys_sorted + i);
}

__global__ void pack_pstar_kernel(uint32_t n_send, uint32_t own_off, const uint32_t *__restrict__ send_idx,
                                  const float4 *__restrict__ pstar, float4 *__restrict__ out) {
  const uint32_t j = blockIdx.x * kBlk + threadIdx.x;
  if (j < n_send) out[j] = ldg4(pstar + own_off + __ldg(send_idx + j));
}

// received ghosts: move the mass from pStar.w into pos.w (the lambda pass reads its own particle's mass there)
__global__ void ghost_fix_kernel(uint32_t n_glo, uint32_t n_own, uint32_t n_local, float4 *__restrict__ pstar,
                                 float4 *__restrict__ pos) {
  const uint32_t t = blockIdx.x * kBlk + threadIdx.x;
  if (t >= n_local - n_own) return;
  const uint32_t i = t < n_glo ? t : t + n_own;
  float4 p = pstar[i];
  pos[i] = make_float4(0.f, 0.f, 0.f, p.w);
  p.w = 0.f;
  pstar[i] = p;
}

// ------------------------------------------------------------------------------------------------- transport
int sync_all(std::vector<pbf_ctx *> &L) {
  for (pbf_ctx *c : L) {
    PBF_CUDA(c, cudaSetDevice(c->device));
    PBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->dist->comm_stream) PBF_CUDA(c, cudaStreamSynchronize(c->dist->comm_stream));
  }
  return PBF_OK;
}

// recv[r * count .. ) on every rank = send of rank r
template <typename FS, typename FR> int all_gather_u32(std::vector<pbf_ctx *> &L, FS send, FR recv, size_t count) {
  D *d0 = L[0]->dist;
  if (!d0->local_mode) {
    pbf_ctx *c = L[0];
    PBF_NCCL(c, g_nccl.AllGather(send(c), recv(c), count, ncclUint32, d0->comm, c->stream));
    return PBF_OK;
  }
  PBF_TRY(sync_all(L));
  for (pbf_ctx *r : L) {
    PBF_CUDA(r, cudaSetDevice(r->device));
    for (pbf_ctx *s : L)
      PBF_CUDA(r, cudaMemcpyAsync(recv(r) + (size_t)s->dist->rank * count, send(s), count * 4, cudaMemcpyDefault, r->stream));
  }
  return sync_all(L);
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

// The all-to-all of every rank's dist->msgs (the same number of messages on every rank), on `which` stream
// (0 = compute stream, 1 = comm stream).
int all_to_all(std::vector<pbf_ctx *> &L, int which) {
  D *d0 = L[0]->dist;
  if (!d0->local_mode) {
    pbf_ctx *c = L[0];
    cudaStream_t st = which ? d0->comm_stream : c->stream;
    PhaseScope ps(c, PBF_PH_HALO);
    PBF_NCCL(c, g_nccl.GroupStart());
    for (const Msg &m : d0->msgs)
      for (int p = 0; p < d0->world; ++p) {
        if (p == d0->rank) continue;
        if (m.send_cnt[p])
          PBF_NCCL(c, g_nccl.Send(m.src + m.send_off[p] * m.elem, m.send_cnt[p] * m.elem, ncclChar, p, d0->comm, st));
        if (m.recv_cnt[p])
          PBF_NCCL(c, g_nccl.Recv(m.dst + m.recv_off[p] * m.elem, m.recv_cnt[p] * m.elem, ncclChar, p, d0->comm, st));
      }
    PBF_NCCL(c, g_nccl.GroupEnd());
    return PBF_OK;
  }
  PBF_TRY(sync_all(L));
  for (pbf_ctx *r : L) {
    PBF_CUDA(r, cudaSetDevice(r->device));
    D *dr = r->dist;
    for (size_t k = 0; k < dr->msgs.size(); ++k)
      for (pbf_ctx *s : L) {
        if (s == r) continue;
        const Msg &mr = dr->msgs[k], &ms = s->dist->msgs[k];
        const uint64_t n = mr.recv_cnt[s->dist->rank];
        if (n != ms.send_cnt[dr->rank]) return fail(r, PBF_ERR_STATE, "all_to_all", "send/recv counts disagree");
        if (n)
          PBF_CUDA(r, cudaMemcpyAsync(mr.dst + mr.recv_off[s->dist->rank] * mr.elem,
                                      ms.src + ms.send_off[dr->rank] * ms.elem, n * mr.elem, cudaMemcpyDefault,
                                      which ? dr->comm_stream : r->stream));
      }
  }
  return sync_all(L);
}

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

// stable compaction of the indices first+t with (flag[first+t] != 0) == want into out[]; with slot > 0 the count is
// also copied to c->mc_total_host[slot] (slot 1..3; read it after synchronising the stream)
int compact_flags(pbf_ctx *c, const uint32_t *flag, uint32_t first, uint32_t n, uint32_t want, uint32_t *out, int slot) {
  D *d = c->dist;
  if (slot > 0) c->mc_total_host[slot] = 0;
  if (n == 0) return PBF_OK;
  const uint32_t nblk = div_up(n, kBlk);
  PBF_CUDA(c, d->sub_cnt.reserve(nblk + 4));
  flag_count_kernel<<<nblk, kBlk, 0, c->stream>>>(flag, first, n, want, d->sub_cnt.p);
  PBF_LAUNCH_CHECK(c);
  PBF_TRY(exclusive_scan_u32(c, d->sub_cnt.p, d->sub_cnt.p, nblk, slot > 0 ? c->mc_total_dev + slot : nullptr));
  flag_scatter_kernel<<<nblk, kBlk, 0, c->stream>>>(flag, first, n, want, d->sub_cnt.p, out);
  PBF_LAUNCH_CHECK(c);
  if (slot > 0)
    PBF_CUDA(c, cudaMemcpyAsync(c->mc_total_host + slot, c->mc_total_dev + slot, 4, cudaMemcpyDeviceToHost, c->stream));
  return PBF_OK;
}

// ------------------------------------------------------------------------------------------------- the phases
int phase_a(pbf_ctx *c, const pbf_params &p, bool replan) {
  D *d = c->dist;
  PBF_CUDA(c, cudaSetDevice(c->device));
  if (p.surface_enabled) return fail(c, PBF_ERR_INVALID, "pbf_dist_step", "marching cubes is not available on the slab path");
  d->n_in = (uint32_t)c->n;
  host_grid(c->h, p, c->grid);
  for (int a = 0; a < 3; ++a)
    if (c->grid.extent[a] == 0 || c->grid.extent[a] > 1023)
      return fail(c, PBF_ERR_INVALID, "grid", "extent must be 1..1023 cells per axis (10-bit Morton, curves.h:73)");
  host_step_const(c->h, p, c->grid, d->n_in, c->sc);
  PBF_CUDA(c, c->key_in.reserve(d->n_in + 1));
  if (d->n_in)
    PBF_TRY(launch_predict_key(c, c->pos[c->cur].p + d->own_off, c->vel[c->cur].p + d->own_off, c->key_in.p));
  if (replan) {
    const uint32_t bits = c->grid.key_bits;
    d->hist_shift = bits > kHistBits ? bits - kHistBits : 0;
    d->hist_buckets = ((c->grid.grid_table_n - 1) >> d->hist_shift) + 2;  // last bucket: everything >= G
    PBF_CUDA(c, d->d_hist.reserve(d->hist_buckets));
    PBF_CUDA(c, cudaMemsetAsync(d->d_hist.p, 0, d->hist_buckets * 4, c->stream));
    if (d->n_in) {
      PhaseScope ps(c, PBF_PH_HALO);
      key_hist_kernel<<<div_up(d->n_in, kBlk), kBlk, 0, c->stream>>>(c->key_in.p, d->n_in, d->hist_shift, d->hist_buckets, d->d_hist.p);
      PBF_LAUNCH_CHECK(c);
    }
  }
  return PBF_OK;
}

int phase_plan(pbf_ctx *c) {
  D *d = c->dist;
  PBF_CUDA(c, cudaSetDevice(c->device));
  std::vector<uint32_t> h32(d->hist_buckets);
  PBF_CUDA(c, cudaMemcpyAsync(h32.data(), d->d_hist.p, d->hist_buckets * 4, cudaMemcpyDeviceToHost, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));
  // Balance WORK, not particle counts: a particle's cost grows with the local density, because the neighbour search
  // tests every particle of its 27 cells.  Fit on one device (dam-1m: 1.89 us per particle-step at 6.3 particles per
  // cell, 2.08 us at 7.4): cost ~ 4.4 + density, with the bucket's particles per key slot as the density.
  std::vector<uint64_t> h64(d->hist_buckets);
  pbf_host_work_weights(h32.data(), d->hist_buckets, d->hist_shift, h64.data());
  d->splits.assign(d->world + 1, 0);
  plan_splits(h64.data(), d->hist_buckets, d->hist_shift, d->world, d->splits.data());
  PBF_CUDA(c, cudaMemcpyAsync(d->d_splits.p, d->splits.data(), (d->world + 1) * 4, cudaMemcpyHostToDevice, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));  // the host vector may change before the copy would run
  return PBF_OK;
}

// classify the owned particles by destination rank; counts per destination into d_row (x1 gathers the rows)
int phase_a2(pbf_ctx *c) {
  D *d = c->dist;
  const int W = d->world;
  PBF_CUDA(c, cudaSetDevice(c->device));
  PBF_CUDA(c, cudaMemsetAsync(d->d_row.p, 0, (W + 1) * 4, c->stream));
  if (d->n_in == 0) return PBF_OK;
  PhaseScope ps(c, PBF_PH_HALO);
  const uint32_t nblk = div_up(d->n_in, kBlk);
  PBF_CUDA(c, d->mask.reserve(d->n_in));
  PBF_CUDA(c, d->blk_cnt.reserve((size_t)W * nblk + 4));
  classify_kernel<<<nblk, kBlk, 0, c->stream>>>(c->key_in.p, d->n_in, d->d_splits.p, d->rank, W, c->sc.G, d->mask.p, d->d_row.p + W);
  PBF_LAUNCH_CHECK(c);
  ghost_count_kernel<<<nblk, kBlk, 0, c->stream>>>(d->mask.p, d->n_in, W, nblk, d->blk_cnt.p, d->d_row.p);
  PBF_LAUNCH_CHECK(c);
  return PBF_OK;
}

// after x1: read the migration matrix (row s, column t = particles going from s to t; the diagonal is unused, column W =
// particles outside the grid), pack the leaving particles, describe the all-to-all
int phase_b(pbf_ctx *c) {
  D *d = c->dist;
  const int W = d->world, r = d->rank, RW = W + 1;
  PBF_CUDA(c, cudaSetDevice(c->device));
  uint32_t *M = d->h_pinned;
  PBF_CUDA(c, cudaMemcpyAsync(M, d->d_all.p, (size_t)W * RW * 4, cudaMemcpyDeviceToHost, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));
  d->total_out = 0;
  d->total_in = 0;
  d->in_lo = 0;
  uint64_t off_diag = 0, outside = 0;
  for (int s = 0; s < W; ++s) {
    outside += M[s * RW + W];
    for (int t = 0; t < W; ++t)
      if (s != t) off_diag += M[s * RW + t];
    if (s != r) {
      d->total_in += M[s * RW + r];
      d->total_out += M[r * RW + s];
      if (s < r) d->in_lo += M[s * RW + r];
    }
  }
  d->n_keep = d->n_in - d->total_out;
  d->any_migrants = off_diag != 0;
  d->any_outside = outside != 0;
  d->n_own = d->n_keep + d->total_in;
  d->msgs.clear();
  if (!d->any_migrants) return PBF_OK;
  // room for the arrivals behind the owned particles of the input arrays
  const size_t need = (size_t)d->own_off + d->n_in + d->total_in;
  PBF_CUDA(c, c->pos[c->cur].reserve(need, true, c->stream));
  PBF_CUDA(c, c->vel[c->cur].reserve(need, true, c->stream));
  PBF_CUDA(c, c->col[c->cur_col].reserve(need, true, c->stream));
  PBF_CUDA(c, c->ids[c->cur].reserve(need, true, c->stream));
  PBF_CUDA(c, d->sb_a.reserve(d->total_out + 1));
  PBF_CUDA(c, d->sb_b.reserve(d->total_out + 1));
  PBF_CUDA(c, d->sb_c.reserve(d->total_out + 1));
  PBF_CUDA(c, d->sb_id.reserve(d->total_out + 1));
  PBF_CUDA(c, d->sb_key.reserve(d->total_out + 1));
  PBF_CUDA(c, d->send_idx.reserve(d->total_out + 1));
  PBF_CUDA(c, d->k2.reserve(d->n_own + 1));
  PBF_CUDA(c, d->v2.reserve(d->n_own + 1));
  float4 *pin = c->pos[c->cur].p + d->own_off, *vin = c->vel[c->cur].p + d->own_off, *cin = c->col[c->cur_col].p + d->own_off;
  unsigned long long *iin = c->ids[c->cur].p + d->own_off;
  if (d->total_out) {
    PhaseScope ps(c, PBF_PH_HALO);
    const uint32_t nblk = div_up(d->n_in, kBlk);
    PBF_TRY(exclusive_scan_u32(c, d->blk_cnt.p, d->blk_cnt.p, (uint64_t)W * nblk, nullptr));
    ghost_scatter_kernel<<<nblk, kBlk, 0, c->stream>>>(d->mask.p, d->n_in, W, nblk, d->blk_cnt.p, d->send_idx.p);
    PBF_LAUNCH_CHECK(c);
    pack_migrants_kernel<<<div_up(d->total_out, kBlk), kBlk, 0, c->stream>>>(d->total_out, d->send_idx.p, pin, vin, cin, iin, c->key_in.p,
                                                                          d->sb_a.p, d->sb_b.p, d->sb_c.p, d->sb_id.p, d->sb_key.p);
    PBF_LAUNCH_CHECK(c);
  }
  d->msgs.resize(5);
  d->msgs[0].init(W, d->sb_a.p, pin + d->n_in, 16);
  d->msgs[1].init(W, d->sb_b.p, vin + d->n_in, 16);
  d->msgs[2].init(W, d->sb_c.p, cin + d->n_in, 16);
  d->msgs[3].init(W, d->sb_id.p, iin + d->n_in, 8);
  d->msgs[4].init(W, d->sb_key.p, d->k2.p, 4);  // keys arrive in merge order: [from lower ranks | kept | from higher ranks]
  // Arrival order = source-rank order, with the kept particles between the lower and the higher ranks: when the
  // ranks' inputs are consecutive blocks of one array (dist.py shard()), the stable sort of that sequence reproduces
  // the single-GPU stable order exactly, cell by cell.
  uint64_t soff = 0, roff = 0;
  for (int q = 0; q < W; ++q) {
    if (q == r) continue;
    const uint64_t sc = M[r * RW + q], rc = M[q * RW + r];
    for (int k = 0; k < 5; ++k) {
      Msg &m = d->msgs[k];
      m.send_cnt[q] = sc;
      m.send_off[q] = soff;
      m.recv_cnt[q] = rc;
      m.recv_off[q] = (k == 4 && q > r) ? roff + d->n_keep : roff;
    }
    soff += sc;
    roff += rc;
  }
  return PBF_OK;
}

// after x2: ONE stable sort of [arrivals from lower ranks | kept | arrivals from higher ranks]; then the ghost
// destinations of every owned particle
int phase_c(pbf_ctx *c) {
  D *d = c->dist;
  const int W = d->world, r = d->rank;
  PBF_CUDA(c, cudaSetDevice(c->device));
  c->sc.n = d->n_own;
  if (d->any_migrants) {
    const uint32_t in_hi = d->total_in - d->in_lo;
    if (d->n_keep) {  // kept = mask 0, in input order; their keys gathered behind the lower ranks' arrivals
      PBF_TRY(compact_flags(c, d->mask.p, 0, d->n_in, 0, d->v2.p + d->in_lo, 0));
      gather_u32_kernel<<<div_up(d->n_keep, kBlk), kBlk, 0, c->stream>>>(d->n_keep, d->v2.p + d->in_lo, c->key_in.p, d->k2.p + d->in_lo);
      PBF_LAUNCH_CHECK(c);
    }
    if (d->in_lo) {  // the arrivals sit behind the owned particles of the input arrays, in source-rank order
      iota_kernel<<<div_up(d->in_lo, kBlk), kBlk, 0, c->stream>>>(d->v2.p, d->in_lo, d->n_in);
      PBF_LAUNCH_CHECK(c);
    }
    if (in_hi) {
      iota_kernel<<<div_up(in_hi, kBlk), kBlk, 0, c->stream>>>(d->v2.p + d->in_lo + d->n_keep, in_hi, d->n_in + d->in_lo);
      PBF_LAUNCH_CHECK(c);
    }
    PBF_TRY(radix_sort_pairs(c, d->k2.p, d->n_own, d->v2.p));
  } else {
    PBF_TRY(radix_sort_pairs(c, c->key_in.p, d->n_in));
  }
  c->n = d->n_own;
  PBF_CUDA(c, cudaMemsetAsync(d->d_row.p, 0, (W + 1) * 4, c->stream));
  if (d->n_own) {
    PhaseScope ps(c, PBF_PH_HALO);
    const uint32_t nblk = div_up(d->n_own, kBlk);
    PBF_CUDA(c, d->mask.reserve(d->n_own));
    PBF_CUDA(c, d->blk_cnt.reserve((size_t)W * nblk + 4));
    ghost_mask_kernel<<<nblk, kBlk, 0, c->stream>>>(c->keys_sorted, d->n_own, c->sc.G, r, W, d->d_splits.p, d->any_outside ? 1 : 0, d->mask.p);
    PBF_LAUNCH_CHECK(c);
    ghost_count_kernel<<<nblk, kBlk, 0, c->stream>>>(d->mask.p, d->n_own, W, nblk, d->blk_cnt.p, d->d_row.p);
    PBF_LAUNCH_CHECK(c);
  }
  return PBF_OK;
}

// after x3: lay out the local arrays, reorder the owned particles into them, pack the ghost payload
int phase_d(pbf_ctx *c) {
  D *d = c->dist;
  const int W = d->world, r = d->rank, RW = W + 1;
  PBF_CUDA(c, cudaSetDevice(c->device));
  uint32_t *GC = d->h_pinned;
  PBF_CUDA(c, cudaMemcpyAsync(GC, d->d_all.p, (size_t)W * RW * 4, cudaMemcpyDeviceToHost, c->stream));
  PBF_CUDA(c, cudaStreamSynchronize(c->stream));
  d->n_send = 0; d->n_glo = 0; d->n_ghi = 0;
  uint64_t all = 0;
  for (int s = 0; s < W; ++s)
    for (int t = 0; t < W; ++t) all += GC[s * RW + t];
  d->any_ghosts = all != 0;
  for (int q = 0; q < W; ++q) {
    d->n_send += GC[r * RW + q];
    if (q < r) d->n_glo += GC[q * RW + r];
    if (q > r) d->n_ghi += GC[q * RW + r];
  }
  d->n_local = d->n_glo + d->n_own + d->n_ghi;
  const uint32_t nl = d->n_local;
  if ((uint64_t)d->n_glo + d->n_own + d->n_ghi >= 0xFFFFFFF0ull) return fail(c, PBF_ERR_INVALID, "n", "too many local particles");
  const int o = c->cur ^ 1, oc = c->cur_col ^ 1;
  PBF_CUDA(c, c->pos[o].reserve(nl + 1));
  PBF_CUDA(c, c->vel[o].reserve(nl + 1));
  PBF_CUDA(c, c->ids[o].reserve(nl + 1));
  PBF_CUDA(c, c->col[oc].reserve(nl + 1));
  PBF_CUDA(c, c->pstar[0].reserve(nl + 1));
  PBF_CUDA(c, c->pstar[1].reserve(nl + 1));
  PBF_CUDA(c, d->keys_local.reserve(nl + 1));
  PBF_CUDA(c, c->table.reserve((size_t)c->sc.G + 1));
  PBF_CUDA(c, c->rho.reserve(nl + 1));
  PBF_CUDA(c, d->send_idx.reserve(d->n_send + 1));
  PBF_CUDA(c, d->sb_a.reserve(d->n_send + 1));
  PBF_CUDA(c, d->sb_b.reserve(d->n_send + 1));
  PBF_CUDA(c, d->sb_key.reserve(d->n_send + 1));
  if (d->n_own) {
    c->sc.n = d->n_own;
    PBF_TRY(launch_reorder(c, c->perm, c->pos[c->cur].p + d->own_off, c->vel[c->cur].p + d->own_off,
                           c->col[c->cur_col].p + d->own_off, c->ids[c->cur].p + d->own_off, c->pos[o].p + d->n_glo,
                           c->vel[o].p + d->n_glo, c->col[oc].p + d->n_glo, c->ids[o].p + d->n_glo, c->pstar[0].p + d->n_glo));
    PBF_CUDA(c, cudaMemcpyAsync(d->keys_local.p + d->n_glo, c->keys_sorted, (size_t)d->n_own * 4, cudaMemcpyDeviceToDevice, c->stream));
  }
  c->cur = o;
  c->cur_col = oc;
  d->own_off = d->n_glo;
  d->msgs.clear();
  if (!d->any_ghosts) return PBF_OK;
  if (d->n_send) {
    PhaseScope ps(c, PBF_PH_HALO);
    const uint32_t nblk = div_up(d->n_own, kBlk);
    PBF_TRY(exclusive_scan_u32(c, d->blk_cnt.p, d->blk_cnt.p, (uint64_t)W * nblk, nullptr));
    ghost_scatter_kernel<<<nblk, kBlk, 0, c->stream>>>(d->mask.p, d->n_own, W, nblk, d->blk_cnt.p, d->send_idx.p);
    PBF_LAUNCH_CHECK(c);
    pack_ghosts_kernel<<<div_up(d->n_send, kBlk), kBlk, 0, c->stream>>>(d->n_send, d->own_off, d->send_idx.p, c->pstar[0].p, c->pos[o].p,
                                                                       c->col[oc].p, c->keys_sorted, d->sb_a.p, d->sb_b.p, d->sb_key.p);
    PBF_LAUNCH_CHECK(c);
  }
  d->msgs.resize(3);
  d->msgs[0].init(W, d->sb_a.p, c->pstar[0].p, 16);
  d->msgs[1].init(W, d->sb_b.p, c->col[oc].p, 16);
  d->msgs[2].init(W, d->sb_key.p, d->keys_local.p, 4);
  uint64_t soff = 0, rlo = 0, rhi = (uint64_t)d->n_glo + d->n_own;
  for (int q = 0; q < W; ++q) {
    if (q == r) continue;
    const uint64_t sc = GC[r * RW + q], rc = GC[q * RW + r];
    for (Msg &m : d->msgs) {
      m.send_cnt[q] = sc; m.send_off[q] = soff;
      m.recv_cnt[q] = rc; m.recv_off[q] = q < r ? rlo : rhi;
    }
    soff += sc;
    if (q < r) rlo += rc; else rhi += rc;
  }
  return PBF_OK;
}

// after x4: ghosts are in place -> cell table, subsets, diffuse
int phase_e(pbf_ctx *c) {
  D *d = c->dist;
  PBF_CUDA(c, cudaSetDevice(c->device));
  const uint32_t n_gh = d->n_glo + d->n_ghi;
  c->sc.n = d->n_local;
  c->grid.n_particles = d->n_local;
  c->keys_sorted = d->keys_local.p;
  if (d->n_local == 0) return PBF_OK;
  if (n_gh) {
    PhaseScope ps(c, PBF_PH_HALO);
    ghost_fix_kernel<<<div_up(n_gh, kBlk), kBlk, 0, c->stream>>>(d->n_glo, d->n_own, d->n_local, c->pstar[0].p, c->pos[c->cur].p);
    PBF_LAUNCH_CHECK(c);
  }
  PBF_TRY(launch_cell_table(c, c->keys_sorted, c->table.p));
  {
    PhaseScope ps(c, PBF_PH_HALO);
    PBF_CUDA(c, d->role.reserve(d->n_local));
    PBF_CUDA(c, cudaMemsetAsync(c->mc_total_dev + 1, 0, 8, c->stream));
    roles_kernel<<<div_up(d->n_local, kBlk), kBlk, 0, c->stream>>>(c->keys_sorted, d->mask.p, d->n_glo, d->n_own, d->n_local, c->sc.G,
                                                                  d->splits[d->rank], d->splits[d->rank + 1], d->any_outside ? 1 : 0,
                                                                  d->role.p, c->mc_total_dev + 1);
    PBF_LAUNCH_CHECK(c);
    // statistics only: read after the next synchronisation (pbf_dist_stats_read)
    PBF_CUDA(c, cudaMemcpyAsync(c->mc_total_host + 1, c->mc_total_dev + 1, 8, cudaMemcpyDeviceToHost, c->stream));
  }
  if (c->flags & PBF_FLAG_DEBUG_COUNTS) {
    PBF_CUDA(c, c->cand_count.reserve(d->n_local));
    PBF_CUDA(c, c->nbr_count.reserve(d->n_local));
    PBF_TRY(launch_neighbour_counts(c, c->keys_sorted, c->table.p, c->pstar[0].p, c->cand_count.p, c->nbr_count.p));
  }
  PBF_CUDA(c, c->col[c->cur_col ^ 1].reserve(d->n_local + 1));
  {  // colour diffusion beside the solver iterations (as in the single-device step); group_step joins before finalise
    PBF_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
    PBF_CUDA(c, cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
    cudaStream_t main_stream = c->stream;
    c->stream = c->side_stream;
    const int rc = launch_diffuse_tiled(c, c->keys_sorted, c->table.p, c->col[c->cur_col].p, c->col[c->cur_col ^ 1].p);
    c->stream = main_stream;
    PBF_TRY(rc);
    PBF_CUDA(c, cudaEventRecord(c->ev_join, c->side_stream));
    d->diffuse_pending = true;
  }
  c->cur_col ^= 1;
  return PBF_OK;
}

int group_step(std::vector<pbf_ctx *> &L, const pbf_params &p) {
  D *d0 = L[0]->dist;
  const int W = d0->world, RW = W + 1;
  const bool replan = d0->splits.empty() || (d0->replan_every && d0->step_index % d0->replan_every == 0);
  for (pbf_ctx *c : L) PBF_TRY(phase_a(c, p, replan));
  if (replan) {
    PBF_TRY(all_reduce_sum_u32(L, [](pbf_ctx *c) { return c->dist->d_hist.p; }, L[0]->dist->hist_buckets));
    for (pbf_ctx *c : L) PBF_TRY(phase_plan(c));
  }
  for (pbf_ctx *c : L) PBF_TRY(phase_a2(c));
  PBF_TRY(all_gather_u32(L, [](pbf_ctx *c) { return c->dist->d_row.p; }, [](pbf_ctx *c) { return c->dist->d_all.p; }, RW));
  for (pbf_ctx *c : L) PBF_TRY(phase_b(c));
  if (d0->any_migrants) PBF_TRY(all_to_all(L, 0));
  for (pbf_ctx *c : L) PBF_TRY(phase_c(c));
  PBF_TRY(all_gather_u32(L, [](pbf_ctx *c) { return c->dist->d_row.p; }, [](pbf_ctx *c) { return c->dist->d_all.p; }, RW));
  for (pbf_ctx *c : L) PBF_TRY(phase_d(c));
  if (d0->any_ghosts) PBF_TRY(all_to_all(L, 0));
  for (pbf_ctx *c : L) PBF_TRY(phase_e(c));

  // per-iteration halo message: pStar of the send list -> the ghost ranges of pstar[0]
  for (pbf_ctx *c : L) {
    D *d = c->dist;
    if (!d->any_ghosts) continue;
    Msg keep = d->msgs[0];
    d->msgs.assign(1, keep);
    d->msgs[0].src = reinterpret_cast<const char *>(d->sb_a.p);
    d->msgs[0].dst = reinterpret_cast<char *>(c->pstar[0].p);
  }
  for (uint64_t it = 0; it < p.iteration; ++it) {
    const bool exchange = d0->any_ghosts && it + 1 < p.iteration;
    for (pbf_ctx *c : L) {
      D *d = c->dist;
      PBF_CUDA(c, cudaSetDevice(c->device));
      if (d->n_local == 0) continue;
      c->sc.n = d->n_local;
      float *rho = it + 1 == p.iteration ? c->rho.p : nullptr;
      const bool have_ghosts = d->n_glo + d->n_ghi != 0;
      PBF_TRY(solver_lambda(c, 0, d->n_local, c->pstar[0].p, c->pstar[1].p, rho, have_ghosts ? d->role.p : nullptr, kRoleLambda));
      if (d->n_own == 0) continue;
      if (exchange && d->n_send) {
        // boundary particles first; their halo goes out on the comm stream while the interior pass runs
        PBF_TRY(solver_delta(c, d->own_off, d->n_own, c->pstar[1].p, c->pstar[0].p, d->role.p, kRoleBoundary));
        PBF_CUDA(c, cudaEventRecord(d->ev_boundary, c->stream));
        PBF_CUDA(c, cudaStreamWaitEvent(d->comm_stream, d->ev_boundary, 0));
        pack_pstar_kernel<<<div_up(d->n_send, kBlk), kBlk, 0, d->comm_stream>>>(d->n_send, d->own_off, d->send_idx.p, c->pstar[0].p, d->sb_a.p);
        PBF_LAUNCH_CHECK(c);
        PBF_TRY(solver_delta(c, d->own_off, d->n_own, c->pstar[1].p, c->pstar[0].p, d->role.p, kRoleInterior));
      } else {
        // (role: ghosts are part of the local array; only owned particles move)
        PBF_TRY(solver_delta(c, d->own_off, d->n_own, c->pstar[1].p, c->pstar[0].p, have_ghosts ? d->role.p : nullptr,
                             kRoleBoundary | kRoleInterior));
        if (exchange) {  // a rank that sends nothing still takes part in the exchange
          PBF_CUDA(c, cudaEventRecord(d->ev_boundary, c->stream));
          PBF_CUDA(c, cudaStreamWaitEvent(d->comm_stream, d->ev_boundary, 0));
        }
      }
    }
    if (exchange) {
      // Receiving overwrites ghost entries of pstar[0] that this iteration's interior delta pass does not read
      // (delta reads pstar[1]); the next lambda pass waits for the exchange.
      PBF_TRY(all_to_all(L, 1));
      for (pbf_ctx *c : L) {
        D *d = c->dist;
        PBF_CUDA(c, cudaSetDevice(c->device));
        PBF_CUDA(c, cudaEventRecord(d->ev_halo, d->comm_stream));
        PBF_CUDA(c, cudaStreamWaitEvent(c->stream, d->ev_halo, 0));
      }
    }
  }
  for (pbf_ctx *c : L) {
    D *d = c->dist;
    PBF_CUDA(c, cudaSetDevice(c->device));
    if (d->diffuse_pending) {
      PBF_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
      d->diffuse_pending = false;
    }
    if (d->n_own) {
      c->sc.n = d->n_own;
      PBF_TRY(launch_finalise(c, c->pstar[0].p + d->own_off, c->pos[c->cur].p + d->own_off, c->vel[c->cur].p + d->own_off));
    }
    c->sc.n = d->n_local;
    c->n = d->n_own;
    c->have_state = true;
    c->prof.steps++;
    d->step_index++;
    d->stats.owned = d->n_own;
    d->stats.ghosts = d->n_glo + d->n_ghi;
    d->stats.migrants_out = d->total_out;
    d->stats.migrants_in = d->total_in;
    d->stats.halo_bytes_per_iteration = (uint64_t)d->n_send * 16;
    d->stats.key_lo = d->splits[d->rank];
    d->stats.key_hi = d->splits[d->rank + 1];
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
  PBF_CUDA(c, d->d_row.reserve(world + 2));
  PBF_CUDA(c, d->d_all.reserve((size_t)world * (world + 1) + 2));
  PBF_CUDA(c, cudaHostAlloc(&d->h_pinned, ((size_t)(world + 1) * (world + 2) + 64) * 4, cudaHostAllocDefault));
  PBF_CUDA(c, cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking));
  PBF_CUDA(c, cudaEventCreateWithFlags(&d->ev_boundary, cudaEventDisableTiming));
  PBF_CUDA(c, cudaEventCreateWithFlags(&d->ev_halo, cudaEventDisableTiming));
  return PBF_OK;
}

}  // namespace

namespace pbf {
void dist_release(pbf_ctx *ctx) {
  if (!ctx->dist) return;
  D *d = ctx->dist;
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

int pbf_dist_upload(pbf_ctx *ctx, const pbf_particle *xs, uint64_t n) {
  if (!ctx || !ctx->dist) return fail(ctx, PBF_ERR_STATE, "pbf_dist_upload", "pbf_dist_init first");
  ctx->dist->own_off = 0;
  ctx->dist->splits.clear();
  ctx->dist->step_index = 0;
  return pbf_upload(ctx, xs, n);
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
  PBF_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_out) *n_out = ctx->n;
  if (capacity < ctx->n) return fail(ctx, PBF_ERR_CAPACITY, "pbf_dist_download", "capacity too small");
  if (ctx->n == 0) return PBF_OK;
  if (!xs) return fail(ctx, PBF_ERR_INVALID, "xs", "NULL");
  const uint32_t off = ctx->dist->own_off;
  PBF_CUDA(ctx, ctx->aos.reserve(ctx->n + 1));
  PBF_TRY(launch_pack_aos(ctx, ctx->aos.p, ctx->n, ctx->pos[ctx->cur].p + off, ctx->vel[ctx->cur].p + off,
                          ctx->col[ctx->cur_col].p + off, ctx->ids[ctx->cur].p + off));
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
  for (size_t r = 0; r < W; ++r) {
    pbf_ctx *c = L[r];
    PBF_CUDA(c, cudaSetDevice(c->device));
    c->dist->own_off = 0;
    PBF_TRY(upload_device(c, xs + off, cnt[r]));
    off += cnt[r];
  }
  for (pbf_ctx *c : L) {
    PBF_CUDA(c, cudaSetDevice(c->device));
    PBF_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->n && *c->flag_host) {
      for (pbf_ctx *q : L) { q->n = 0; q->have_state = false; }
      return fail(ctx, PBF_ERR_INVALID, "xs", "Obstacle particles are not supported (the reference OMP backend drops them)");
    }
  }
  PBF_TRY(pbf_dist_step(ctx, params));
  off = 0;
  for (size_t r = 0; r < W; ++r) {  // every rank's D2H is enqueued before any is waited for
    pbf_ctx *c = L[r];
    PBF_CUDA(c, cudaSetDevice(c->device));
    cnt[r] = c->n;
    if (off + c->n > n) return fail(ctx, PBF_ERR_STATE, "pbf_dist_advance_host", "particle count changed");
    if (c->n) {
      const uint32_t o = c->dist->own_off;
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
  return PBF_OK;
}

int pbf_dist_stats_read(pbf_ctx *ctx, pbf_dist_stats *out) {
  if (!ctx || !ctx->dist || !out) return fail(ctx, PBF_ERR_STATE, "pbf_dist_stats_read", "pbf_dist_init first");
  PBF_CUDA(ctx, cudaSetDevice(ctx->device));
  PBF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = ctx->dist->stats;
  out->ghost_ring1 = ctx->mc_total_host[1];
  out->boundary = ctx->mc_total_host[2];
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
