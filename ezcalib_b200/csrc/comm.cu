// comm.cu -- the multi-GPU exchange of the LM solve, owned by the library: a one-shot all-reduce over NVLink / NVSwitch peer
// memory instead of two NCCL launches per LM iteration issued through a host callback.
//
// What is exchanged: per iteration every rank holds two tiny packets in HBM -- the Schur-reduced partial
// system of its views (<= 170 doubles) and the candidate's H_cc / cost / model-cost sums (149 doubles).  Each rank owns a
// MAILBOX in its own device memory: 2 (sequence parity) x world slots of 256 doubles + a flag.  One kernel per rank
//   1. stores its packet into slot [rank] of EVERY rank's mailbox (plain P2P stores through the peer mapping),
//   2. fences (system scope) and publishes the call's sequence number in the slot flags,
//   3. spins until all `world` slots of its own mailbox carry that sequence number,
//   4. sums the slots in FIXED RANK ORDER into the caller's buffer.
// Latency: one NVLink store round + one poll; the result is bit-identical on every rank and independent of arrival
// order (1-rank and N-rank solves of the same split are reproducible).  Two parity buffers are enough: a rank can only
// start call n after every peer published call n-1, i.e. after every peer finished reading call n-2.
#include "ezc_common.cuh"

namespace ezc {

constexpr int kSlotDoubles = 256;
struct __align__(16) CommSlot {
  double data[kSlotDoubles];
  unsigned long long flag;
  unsigned long long pad;
};

struct Comm {
  int rank = 0, world = 1;
  CommSlot* local = nullptr;            // [2][world], this rank's mailbox
  CommSlot* peers[64] = {};             // mailbox base of every rank as mapped into this process (peers[rank] == local)
  bool opened[64] = {};                 // mapped through cudaIpcOpenMemHandle (must be closed)
  CommSlot** d_peers = nullptr;         // device copy of peers[]
  unsigned long long seq = 0;           // calls issued so far
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) comm_allreduce_kernel(CommSlot* const* __restrict__ peers, CommSlot* local, int rank, int world,
                                                              unsigned long long seq, double* __restrict__ buf, int count) {
  const int par = (int)(seq & 1);
  const int t = threadIdx.x;
  // 1. my packet into slot [rank] of every mailbox
  const double v = t < count ? buf[t] : 0.0;
  for (int p = 0; p < world; ++p) {
    CommSlot* dst = peers[p] + (size_t)par * world + rank;
    if (t < count) dst->data[t] = v;
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish
  if (t < world) st_release_sys(&(peers[t] + (size_t)par * world + rank)->flag, seq);
  // 3. wait for every rank's packet in my own mailbox
  if (t < world) {
    const unsigned long long* f = &(local + (size_t)par * world + t)->flag;
    while (ld_acquire_sys(f) != seq) { __nanosleep(20); }
  }
  __syncthreads();
  // 4. fixed rank order
  if (t < count) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += __ldcv(&(local + (size_t)par * world + r)->data[t]);
    buf[t] = s;
  }
}

static int32_t comm_allreduce_cb(void* user, void* device_buf, int32_t count, int32_t op, void* stream) {
  Comm* c = static_cast<Comm*>(user);
  if (op != 0 || count > kSlotDoubles || count < 0) return 2;   // sums of <= 256 doubles only (all the solver needs)
  ++c->seq;
  comm_allreduce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(c->d_peers, c->local, c->rank, c->world, c->seq,
                                                                          static_cast<double*>(device_buf), count);
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}

static void comm_free(void* p) {
  Comm* c = static_cast<Comm*>(p);
  if (!c) return;
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r] && c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
  if (c->d_peers) cudaFree(c->d_peers);
  if (c->local) cudaFree(c->local);
  delete c;
}

static Comm* comm_of(ezc_context* ctx, bool create) {
  if (!ctx->comm && create) {
    ctx->comm = new Comm();
    ctx->comm_free = comm_free;
  }
  return static_cast<Comm*>(ctx->comm);
}

}  // namespace ezc

using namespace ezc;

extern "C" {

ezc_status ezc_comm_mailbox(ezc_context* ctx, int32_t world, void** mailbox_out, uint8_t* ipc_handle_out) {
  EZC_CHECK(ctx && world >= 1 && world <= 64, EZC_ERR_INVALID, "ezc_comm_mailbox: bad argument (world 1..64)");
  EZC_CUDA(cudaSetDevice(ctx->device));
  Comm* c = comm_of(ctx, true);
  if (!c->local || c->world != world) {
    if (c->local) { cudaStreamSynchronize(ctx->stream); cudaFree(c->local); c->local = nullptr; }
    const size_t bytes = sizeof(CommSlot) * 2 * (size_t)world;
    EZC_CUDA(cudaMalloc(&c->local, bytes));          // a plain allocation: cudaIpcGetMemHandle wants the base of one
    EZC_CUDA(cudaMemset(c->local, 0, bytes));
    c->world = world;
    c->seq = 0;
  }
  if (mailbox_out) *mailbox_out = c->local;
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    EZC_CUDA(cudaIpcGetMemHandle(&h, c->local));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(ipc_handle_out, &h, sizeof(h));
  }
  return EZC_OK;
}

static ezc_status comm_finish(ezc_context* ctx, Comm* c, int32_t rank, int32_t world) {
  c->rank = rank;
  if (!c->d_peers) EZC_CUDA(cudaMalloc(&c->d_peers, sizeof(CommSlot*) * 64));
  EZC_CUDA(cudaMemcpy(c->d_peers, c->peers, sizeof(CommSlot*) * 64, cudaMemcpyHostToDevice));
  ctx->allreduce = comm_allreduce_cb;
  ctx->allreduce_user = c;
  ctx->rank = rank;
  ctx->world = world;
  return EZC_OK;
}

ezc_status ezc_comm_attach(ezc_context* ctx, int32_t rank, int32_t world, void* const* mailboxes) {
  EZC_CHECK(ctx && mailboxes && world >= 1 && world <= 64 && rank >= 0 && rank < world, EZC_ERR_INVALID, "ezc_comm_attach: bad argument");
  EZC_CUDA(cudaSetDevice(ctx->device));
  Comm* c = comm_of(ctx, false);
  EZC_CHECK(c && c->local && c->world == world, EZC_ERR_INVALID, "ezc_comm_attach: call ezc_comm_mailbox(ctx, world, ...) first");
  EZC_CHECK(mailboxes[rank] == c->local, EZC_ERR_INVALID, "ezc_comm_attach: mailboxes[rank] is not this context's mailbox");
  for (int r = 0; r < world; ++r) { c->peers[r] = static_cast<CommSlot*>(mailboxes[r]); c->opened[r] = false; }
  return comm_finish(ctx, c, rank, world);
}

ezc_status ezc_comm_init(ezc_context* ctx, int32_t rank, int32_t world, const uint8_t* all_ipc_handles) {
  EZC_CHECK(ctx && all_ipc_handles && world >= 1 && world <= 64 && rank >= 0 && rank < world, EZC_ERR_INVALID, "ezc_comm_init: bad argument");
  EZC_CUDA(cudaSetDevice(ctx->device));
  Comm* c = comm_of(ctx, false);
  EZC_CHECK(c && c->local && c->world == world, EZC_ERR_INVALID, "ezc_comm_init: call ezc_comm_mailbox(ctx, world, ...) first");
  for (int r = 0; r < world; ++r) {
    if (r == rank) { c->peers[r] = c->local; c->opened[r] = false; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, all_ipc_handles + (size_t)r * 64, 64);
    void* p = nullptr;
    EZC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peers[r] = static_cast<CommSlot*>(p);
    c->opened[r] = true;
  }
  return comm_finish(ctx, c, rank, world);
}

}  // extern "C"
