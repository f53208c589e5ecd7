// Peer-memory arena for the z-slab decomposition: field storage that every rank of the box can address over
// NVLink (CUDA IPC), the halo exchange as ONE kernel of direct stores into the neighbours' halo planes, and a
// device-side barrier. No NCCL call, no host synchronisation, no pack / unpack buffers on this path.
//
// Layout of a rank's arena (the same on every rank, so an offset names the same object everywhere):
//   [0, HEADER)        flags[src rank][slot] (32-bit epochs, written by rank `src` with remote stores and polled
//                      locally), then the block counter of the halo kernel
//   [HEADER, bytes)    field storage handed out by the host layer (sopht_b200/parallel/peer.py)
//
// Halo exchange protocol, per z neighbour (epoch e is a host-side counter, identical on all ranks because every
// rank issues the same sequence of exchanges):
//   1. ready: store e into the neighbour's flags[me][READY]. The kernel runs in stream order, so this says
//      "every kernel of mine that read or wrote my halo planes before this exchange has finished".
//   2. every CTA polls its own flags[neighbour][READY] >= e, then stores this rank's boundary planes straight
//      into the neighbour's halo planes (16-byte vector stores over NVLink).
//   3. done: __threadfence_system(); the last CTA stores e into the neighbour's flags[me][DONE] and then polls
//      its own flags[neighbour][DONE] >= e - when the kernel retires, this rank's halo planes are complete.
// ref: the reference has no distributed path; what is exchanged is the ghost ring its stencil kernels read
// (SURVEY.md 8e).
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace sopht {
namespace {

constexpr int MAX_RANKS = 8;
constexpr int SLOTS = 4;  // READY, DONE, BARRIER, spare
constexpr int SLOT_READY = 0, SLOT_DONE = 1, SLOT_BARRIER = 2;
constexpr size_t HEADER = 4096;

struct Header {
  uint32_t flags[MAX_RANKS][SLOTS];
  uint32_t counter;
  uint32_t error;  // 0, or 1 + the rank a poll gave up on (sticky; read by sopht_peer_arena_status)
};
static_assert(sizeof(Header) <= HEADER, "arena header");

__device__ __forceinline__ void signal(uint32_t* remote, uint32_t epoch) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
}
// Poll a local flag until the peer has stored `epoch` into it. The spin is bounded (budget_ns of %globaltimer): a rank
// that died, or that issued its exchanges in a different order, must not hang every neighbour's GPU for good - the
// kernel gives up, records which peer it was waiting for in the header and lets the stream drain; the host layer
// turns the sticky error word into an exception (sopht_peer_arena_status).
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_for(const uint32_t* local, uint32_t epoch, Header* mine, int peer_rank,
                                         uint64_t budget_ns) {
  uint32_t v;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local) : "memory");
    if ((int32_t)(v - epoch) >= 0) return;  // wrap-safe
    if ((++spins & 1023u) == 0) {
      const uint64_t now = global_ns();
      if (!t0) t0 = now;
      if (now - t0 > budget_ns) {
        atomicCAS(&mine->error, 0u, 1u + (uint32_t)peer_rank);
        return;
      }
    }
  } while (true);
}

struct HaloField {
  int64_t offset;        // bytes from the arena base to element (0, 0, 0, 0) of the local array
  int64_t comp_stride;   // bytes between components
  int ncomp;
};
struct HaloArgs {
  HaloField f[4];
  int nfields;
  int nz_local, halo;
  int64_t plane_bytes;   // ny * nx * elem (a multiple of 16)
  char* self;
  char* lo;              // low / high z neighbour's arena base (nullptr: global boundary)
  char* hi;
  int rank, rank_lo, rank_hi;
  uint32_t epoch;
  uint64_t budget_ns;
};

__global__ void __launch_bounds__(256) halo_push_kernel(HaloArgs a) {
  Header* mine = reinterpret_cast<Header*>(a.self);
  Header* hlo = reinterpret_cast<Header*>(a.lo);
  Header* hhi = reinterpret_cast<Header*>(a.hi);
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      if (hlo) signal(&hlo->flags[a.rank][SLOT_READY], a.epoch);
      if (hhi) signal(&hhi->flags[a.rank][SLOT_READY], a.epoch);
    }
    if (hlo) wait_for(&mine->flags[a.rank_lo][SLOT_READY], a.epoch, mine, a.rank_lo, a.budget_ns);
    if (hhi) wait_for(&mine->flags[a.rank_hi][SLOT_READY], a.epoch, mine, a.rank_hi, a.budget_ns);
  }
  __syncthreads();
  // my first owned planes [h, 2h) -> low neighbour's planes [n + h, n + 2h); my last owned planes [n, n + h) ->
  // high neighbour's planes [0, h)
  const int64_t slab = (int64_t)a.halo * a.plane_bytes;  // bytes per (field, component, side)
  const int64_t vec_per_slab = slab / 16;
  int total_comp = 0;
  for (int q = 0; q < a.nfields; ++q) total_comp += a.f[q].ncomp;
  const int64_t total = vec_per_slab * total_comp * 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i % vec_per_slab;
    int64_t r = i / vec_per_slab;
    const int side = (int)(r & 1);
    r >>= 1;
    int q = 0, c = (int)r;
    while (c >= a.f[q].ncomp) c -= a.f[q++].ncomp;
    char* dst_base = side ? a.hi : a.lo;
    if (!dst_base) continue;
    const int64_t comp = a.f[q].offset + (int64_t)c * a.f[q].comp_stride;
    const int64_t src_plane = side ? a.nz_local : a.halo;
    const int64_t dst_plane = side ? 0 : a.nz_local + a.halo;
    const uint4 val = *reinterpret_cast<const uint4*>(a.self + comp + src_plane * a.plane_bytes + v * 16);
    *reinterpret_cast<uint4*>(dst_base + comp + dst_plane * a.plane_bytes + v * 16) = val;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(&mine->counter, 1u);
    if (t == gridDim.x - 1) {
      mine->counter = 0;  // next launch (stream-ordered) starts from zero
      if (hlo) signal(&hlo->flags[a.rank][SLOT_DONE], a.epoch);
      if (hhi) signal(&hhi->flags[a.rank][SLOT_DONE], a.epoch);
      if (hlo) wait_for(&mine->flags[a.rank_lo][SLOT_DONE], a.epoch, mine, a.rank_lo, a.budget_ns);
      if (hhi) wait_for(&mine->flags[a.rank_hi][SLOT_DONE], a.epoch, mine, a.rank_hi, a.budget_ns);
    }
  }
}

struct BarrierArgs {
  char* peer[MAX_RANKS];
  int nranks, rank;
  uint32_t epoch;
  uint64_t budget_ns;
};
// all-ranks barrier in stream order: everything every rank enqueued before it (including its stores into
// other ranks' memory) is complete and visible when it retires
__global__ void peer_barrier_kernel(BarrierArgs a) {
  const int q = threadIdx.x;
  __threadfence_system();
  if (q < a.nranks && q != a.rank) {
    signal(&reinterpret_cast<Header*>(a.peer[q])->flags[a.rank][SLOT_BARRIER], a.epoch);
    Header* mine = reinterpret_cast<Header*>(a.peer[a.rank]);
    wait_for(&mine->flags[q][SLOT_BARRIER], a.epoch, mine, q, a.budget_ns);
  }
}

}  // namespace
}  // namespace sopht

using namespace sopht;

struct sopht_peer_arena {
  char* base = nullptr;
  size_t bytes = 0;
  int nranks = 1, rank = 0;
  char* peer[MAX_RANKS] = {};
  bool opened = false;
  bool ring = false;  // periodic z direction: rank 0's low neighbour is rank P - 1 and vice versa
  uint32_t halo_epoch = 0, barrier_epoch = 0;
  uint64_t budget_ns = 30ull * 1000000000ull;  // SOPHT_PEER_TIMEOUT_S
};

extern "C" {

int sopht_peer_arena_create(sopht_peer_arena_t* handle, size_t payload_bytes, int nranks, int rank,
                            unsigned char* ipc_handle_out) {
  if (!handle || !ipc_handle_out) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null argument", __func__);
  if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks)
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: 1 <= nranks <= %d and 0 <= rank < nranks", __func__, MAX_RANKS);
  auto* h = new sopht_peer_arena();
  if (const char* v = getenv("SOPHT_PEER_TIMEOUT_S"))
    if (atof(v) > 0) h->budget_ns = (uint64_t)(atof(v) * 1e9);
  h->bytes = HEADER + ((payload_bytes + 255) / 256) * 256;
  h->nranks = nranks, h->rank = rank;
  if (cudaMalloc(&h->base, h->bytes) != cudaSuccess) {
    delete h;
    SOPHT_FAIL(SOPHT_ERR_ALLOC, "%s: out of device memory (%zu bytes)", __func__, payload_bytes);
  }
  if (cudaMemset(h->base, 0, h->bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(h->base);
    delete h;
    SOPHT_FAIL(SOPHT_ERR_CUDA, "%s: clearing the arena failed", __func__);
  }
  h->peer[rank] = h->base;
  cudaIpcMemHandle_t ipc;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (nranks > 1) {
    cudaError_t e = cudaIpcGetMemHandle(&ipc, h->base);
    if (e != cudaSuccess) {
      cudaFree(h->base);
      delete h;
      SOPHT_FAIL(SOPHT_ERR_CUDA, "%s: cudaIpcGetMemHandle failed: %s", __func__, cudaGetErrorString(e));
    }
    memcpy(ipc_handle_out, &ipc, 64);
  } else {
    memset(ipc_handle_out, 0, 64);
    h->opened = true;
  }
  *handle = h;
  return SOPHT_OK;
}

int sopht_peer_arena_open(sopht_peer_arena_t h, const unsigned char* all_ipc_handles) {
  if (!h || !all_ipc_handles) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle or buffer", __func__);
  for (int q = 0; q < h->nranks; ++q) {
    if (q == h->rank) continue;
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, all_ipc_handles + (size_t)q * 64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {  // leave the arena as it was before the call: no half-open peer table
      for (int r = 0; r < h->nranks; ++r)
        if (r != h->rank && h->peer[r]) {
          cudaIpcCloseMemHandle(h->peer[r]);
          h->peer[r] = nullptr;
        }
      SOPHT_FAIL(SOPHT_ERR_CUDA, "%s: cudaIpcOpenMemHandle(rank %d) failed: %s", __func__, q, cudaGetErrorString(e));
    }
    h->peer[q] = reinterpret_cast<char*>(p);
  }
  h->opened = true;
  return SOPHT_OK;
}

/* 0 = healthy; otherwise a device-side poll of this arena timed out (SOPHT_PEER_TIMEOUT_S, default 30 s) waiting for
 * rank (return value - 1): that rank died or issued a different sequence of exchanges / barriers. Synchronises the
 * device (reads one word back). */
int sopht_peer_arena_status(sopht_peer_arena_t h, int* stalled_rank_out) {
  if (!h) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle", __func__);
  uint32_t err = 0;
  SOPHT_CUDA(cudaMemcpy(&err, h->base + offsetof(Header, error), sizeof(err), cudaMemcpyDeviceToHost));
  if (stalled_rank_out) *stalled_rank_out = err ? (int)err - 1 : -1;
  if (err) SOPHT_FAIL(SOPHT_ERR_CUDA, "%s: a peer exchange timed out waiting for rank %d", __func__, (int)err - 1);
  return SOPHT_OK;
}

int sopht_peer_arena_set_periodic(sopht_peer_arena_t h, int periodic_z) {
  if (!h) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle", __func__);
  h->ring = periodic_z != 0;
  return SOPHT_OK;
}

void* sopht_peer_arena_payload(sopht_peer_arena_t h) { return h ? h->base + HEADER : nullptr; }

int sopht_peer_halo_exchange(sopht_peer_arena_t h, int nfields, const int64_t* payload_offsets_bytes,
                             const int64_t* comp_stride_bytes, const int* ncomp, int nz_local, int halo,
                             int64_t plane_bytes, void* stream) {
  if (!h || !h->opened) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: arena not opened", __func__);
  if (nfields < 1 || nfields > 4 || !payload_offsets_bytes || !comp_stride_bytes || !ncomp)
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: 1..4 fields", __func__);
  if (plane_bytes <= 0 || plane_bytes % 16 || halo < 1 || nz_local < halo)
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: planes must be a multiple of 16 bytes and nz_local >= halo >= 1", __func__);
  if (h->nranks == 1) return SOPHT_OK;
  HaloArgs a{};
  a.nfields = nfields;
  int total_comp = 0;
  for (int q = 0; q < nfields; ++q) {
    const int64_t end = payload_offsets_bytes[q] + comp_stride_bytes[q] * (ncomp[q] - 1) +
                        (int64_t)(nz_local + 2 * halo) * plane_bytes;
    if (payload_offsets_bytes[q] < 0 || payload_offsets_bytes[q] % 16 || comp_stride_bytes[q] % 16 ||
        ncomp[q] < 1 || (size_t)end + HEADER > h->bytes)
      SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field %d does not lie inside the arena / is not 16-byte aligned", __func__, q);
    a.f[q] = HaloField{(int64_t)HEADER + payload_offsets_bytes[q], comp_stride_bytes[q], ncomp[q]};
    total_comp += ncomp[q];
  }
  a.nz_local = nz_local, a.halo = halo, a.plane_bytes = plane_bytes;
  a.self = h->base;
  a.rank = h->rank, a.rank_lo = h->rank - 1, a.rank_hi = h->rank + 1;
  if (h->ring) {
    a.rank_lo = (h->rank + h->nranks - 1) % h->nranks;
    a.rank_hi = (h->rank + 1) % h->nranks;
  }
  a.lo = a.rank_lo >= 0 ? h->peer[a.rank_lo] : nullptr;
  a.hi = a.rank_hi < h->nranks ? h->peer[a.rank_hi] : nullptr;
  a.epoch = ++h->halo_epoch;
  a.budget_ns = h->budget_ns;
  const int64_t vecs = (int64_t)halo * plane_bytes / 16 * total_comp * 2;
  int blocks = (int)((vecs + 256 * 8 - 1) / (256 * 8));
  if (blocks > 64) blocks = 64;  // far below one CTA per SM: every CTA is resident, the polls cannot starve
  if (blocks < 1) blocks = 1;
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("comm.halo_push", st);
  halo_push_kernel<<<blocks, 256, 0, st>>>(a);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_peer_barrier(sopht_peer_arena_t h, void* stream) {
  if (!h || !h->opened) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: arena not opened", __func__);
  if (h->nranks == 1) return SOPHT_OK;
  BarrierArgs a{};
  for (int q = 0; q < h->nranks; ++q) a.peer[q] = h->peer[q];
  a.nranks = h->nranks, a.rank = h->rank;
  a.epoch = ++h->barrier_epoch;
  a.budget_ns = h->budget_ns;
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("comm.peer_barrier", st);
  peer_barrier_kernel<<<1, 32, 0, st>>>(a);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_peer_arena_destroy(sopht_peer_arena_t h) {
  if (!h) return SOPHT_OK;
  cudaDeviceSynchronize();
  for (int q = 0; q < h->nranks; ++q)
    if (q != h->rank && h->peer[q]) cudaIpcCloseMemHandle(h->peer[q]);
  cudaFree(h->base);
  delete h;
  return SOPHT_OK;
}

}  // extern "C"
