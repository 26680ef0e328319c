// Ghost-row exchange over NVLink peer memory: one kernel per assembly instead of an NCCL send/recv group,
// two zero-fill launches and two accumulate launches (design B).
//
// Domain decomposition as the reference sees it (one owner per node, ghosts numbered last and grouped by
// owner): after a rank has assembled its own cells into the rows of all its local nodes, the partial sums in
// its ghost rows are one contiguous slice of `values` per neighbour.  The owner PULLS that slice straight out
// of the neighbour's HBM (CUDA IPC mapping of the neighbour's `values`, loads travel over NVLink/NVSwitch)
// and adds it through its precomputed slot list; flags in peer memory order the three steps:
//
//   signal   flags[READY][me] on every neighbour := epoch     (my ghost rows are complete: the assembly kernel
//                                                               ended before this kernel started)
//   pull     wait flags[READY][q] == epoch, values[slots[i]] += peer_values[first + i]
//   ack      flags[PULLED][me] on q := epoch once all my blocks for q are done
//   zero     wait flags[PULLED][q] == epoch, then my slice for q is zeroed (the reference's ghost rows are
//            zero: isOwn gates, modules/testlab/CsrGpuBiliAssembly.cc:273,351)
//
// Every wait depends on remote progress only (never on another local block), the grid is small enough to be
// resident at once, and every wait gives up after AFB_P2P_TIMEOUT_NS with an error flag instead of hanging.
// This replaces what the reference delegates to the solver's parallel matrix assembly (HYPRE IJ off-processor
// values); the NCCL path of arcanefem_b200/distributed.py stays as the portable fallback.
#include <algorithm>
#include <cstring>
#include <vector>

#include <unistd.h>

#include "afb_internal.h"

namespace afb {

constexpr int P2P_MAX_RANK = 64;
constexpr int P2P_BLOCKS_PER_PEER = 48;  // x peers (2 for slabs) stays below one block per SM: all blocks resident at once; fewer per peer
                                         // when many peers would not fit the device together (every wait needs the whole grid resident)
constexpr int P2P_THREADS = 512;
constexpr int P2P_UNROLL = 8;           // remote loads in flight per thread: a pull is latency-bound (NVLink round trip), not bandwidth-bound
constexpr unsigned long long AFB_P2P_TIMEOUT_NS = 4000000000ull;

struct PeerDev {
  const double* peer_values;   // the neighbour's `values` (IPC mapping)
  uint32_t* peer_flags;        // the neighbour's flag block (IPC mapping)
  const int64_t* slots;        // my value slot of every double of the neighbour's slice
  long long pull_first, pull_n;  // the neighbour's slice for me, in its `values`
  long long send_first, send_n;  // my slice for the neighbour, in my `values`
  int rank;
  int pad;
};

struct P2PState {
  bool connected = false;
  uint32_t epoch = 0;
  int my_rank = 0;
  uint32_t* flags = nullptr;   // [2][P2P_MAX_RANK] + error word, cudaMalloc (exported)
  int* counters = nullptr;     // two per peer: blocks done, blocks that gave up waiting for READY
  unsigned long long* wait_ns = nullptr; // [2]: ns the kernels spent waiting for the neighbours' READY / PULLED flags (first block of each neighbour)
  long long stats_epoch0 = 0;  // epoch at the last afb_p2p_wait_stats
  int blocks_per_peer = P2P_BLOCKS_PER_PEER;
  uint32_t* h_err = nullptr;   // sticky error word in mapped pinned host memory: read without a stream synchronisation
  uint32_t* d_err = nullptr;   // its device address
  PeerDev* d_peers = nullptr;
  std::vector<void*> opened;
  int nb_peer = 0;
  const void* values_base = nullptr;
  // asynchronous form: the kernel runs on a side stream (highest priority) so that work which does not touch `values`
  // -- the next BuildMatrix of a time loop -- proceeds on the context stream while the ranks synchronise
  cudaStream_t side = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  bool inflight = false;
};

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ double ld_peer_f64(const double* p)
{
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// returns false on timeout
__device__ __forceinline__ bool wait_epoch(const uint32_t* flag, uint32_t epoch)
{
  const unsigned long long t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
    if (globaltimer_ns() - t0 > AFB_P2P_TIMEOUT_NS) return false;
    __nanosleep(20);
  }
  return true;
}

__global__ void __launch_bounds__(P2P_THREADS)
k_p2p_exchange(const PeerDev* __restrict__ peers, double* __restrict__ values, uint32_t* __restrict__ flags, int my_rank, uint32_t epoch, int* __restrict__ counters,
               int bpp, uint32_t* __restrict__ host_err, unsigned long long* __restrict__ wait_ns)
{
  __shared__ int s_ok;
  const int p = blockIdx.x / bpp, bb = blockIdx.x % bpp;
  const PeerDev P = peers[p];
  uint32_t* err = flags + 2 * P2P_MAX_RANK;
  auto fail = [&](uint32_t code) {
    atomicExch(err, code);
    *reinterpret_cast<volatile uint32_t*>(host_err) = code;
    __threadfence_system();
  };
  if (threadIdx.x == 0) {
    if (bb == 0) {
      __threadfence_system();
      st_release_sys(P.peer_flags + my_rank, epoch); // READY
    }
    const unsigned long long tw = globaltimer_ns();
    s_ok = wait_epoch(flags + P.rank, epoch) ? 1 : 0;
    if (bb == 0) atomicAdd(wait_ns, globaltimer_ns() - tw); // rank skew, as seen from this rank
    if (!s_ok) {
      fail(1u);
      atomicAdd(counters + 2 * p + 1, 1);
    }
  }
  __syncthreads();
  if (s_ok) {
    const long long stride = (long long)bpp * P2P_THREADS;
    const double* src = P.peer_values + P.pull_first;
    long long i = (long long)bb * P2P_THREADS + threadIdx.x;
    for (; i + (P2P_UNROLL - 1) * stride < P.pull_n; i += P2P_UNROLL * stride) {
      double a[P2P_UNROLL];
      long long sl[P2P_UNROLL];
#pragma unroll
      for (int q = 0; q < P2P_UNROLL; ++q) a[q] = ld_peer_f64(src + i + q * stride);
#pragma unroll
      for (int q = 0; q < P2P_UNROLL; ++q) sl[q] = P.slots[i + q * stride];
#pragma unroll
      for (int q = 0; q < P2P_UNROLL; ++q) atomicAdd(values + sl[q], a[q]);
    }
    for (; i < P.pull_n; i += stride) atomicAdd(values + P.slots[i], ld_peer_f64(src + i));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(counters + 2 * p, 1) == bpp - 1) { // last block of this neighbour
      const int gave_up = atomicExch(counters + 2 * p + 1, 0);
      counters[2 * p] = 0;
      __threadfence_system();
      // PULLED only if every block pulled its share: after a READY timeout the neighbour must keep its slice
      // (its own wait for PULLED then times out into status 2 instead of zeroing contributions nobody took)
      if (gave_up == 0) st_release_sys(P.peer_flags + P2P_MAX_RANK + my_rank, epoch);
    }
    const unsigned long long tw = globaltimer_ns();
    s_ok = wait_epoch(flags + P2P_MAX_RANK + P.rank, epoch) ? 1 : 0;
    if (bb == 0) atomicAdd(wait_ns + 1, globaltimer_ns() - tw);
    if (!s_ok) fail(2u);
  }
  __syncthreads();
  if (s_ok) {
    double* dst = values + P.send_first;
    for (long long i = (long long)bb * P2P_THREADS + threadIdx.x; i < P.send_n; i += (long long)bpp * P2P_THREADS) dst[i] = 0.0;
  }
}

static P2PState* state_of(afb_ctx* ctx)
{
  if (!ctx->p2p) ctx->p2p = new P2PState();
  return static_cast<P2PState*>(ctx->p2p);
}

int p2p_export(afb_ctx* ctx, void* values_handle, void* flags_handle)
{
  AFB_REQUIRE(ctx->values.p && ctx->values.owned, AFB_ERR_INVALID, "afb_p2p_export: no library-owned values array (build the pattern first)");
  P2PState* S = state_of(ctx);
  if (!S->flags) {
    AFB_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->flags), sizeof(uint32_t) * (2 * P2P_MAX_RANK + 2)));
    AFB_CUDA(cudaMemset(S->flags, 0, sizeof(uint32_t) * (2 * P2P_MAX_RANK + 2)));
    AFB_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->wait_ns), 2 * sizeof(unsigned long long)));
    AFB_CUDA(cudaMemset(S->wait_ns, 0, 2 * sizeof(unsigned long long)));
    AFB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&S->h_err), sizeof(uint32_t), cudaHostAllocMapped));
    *S->h_err = 0u;
    AFB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&S->d_err), S->h_err, 0));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == AFB_P2P_HANDLE_BYTES, "IPC handle size");
  AFB_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(values_handle), ctx->values.p));
  AFB_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(flags_handle), S->flags));
  S->values_base = ctx->values.p;
  return AFB_OK;
}

// extended descriptor for peers that may live in the same process (several contexts driven by one host program:
// afb_mgpu_*, tests/cpp): CUDA IPC handles cannot be opened by the process that created them, raw pointers can be used
int p2p_export_ex(afb_ctx* ctx, P2PEndpoint* ep)
{
  memset(ep, 0, sizeof(*ep));
  AFB_TRY(p2p_export(ctx, ep->values_handle, ep->flags_handle));
  P2PState* S = state_of(ctx);
  ep->pid = (uint64_t)getpid();
  ep->values_ptr = (uint64_t)(uintptr_t)ctx->values.p;
  ep->flags_ptr = (uint64_t)(uintptr_t)S->flags;
  ep->device = ctx->device;
  return AFB_OK;
}

int p2p_disconnect(afb_ctx* ctx)
{
  if (!ctx->p2p) return AFB_OK;
  P2PState* S = static_cast<P2PState*>(ctx->p2p);
  cudaStreamSynchronize(ctx->stream);
  if (S->side) cudaStreamSynchronize(S->side);
  S->inflight = false;
  for (void* p : S->opened) cudaIpcCloseMemHandle(p);
  S->opened.clear();
  if (S->d_peers) cudaFree(S->d_peers);
  if (S->counters) cudaFree(S->counters);
  S->d_peers = nullptr;
  S->counters = nullptr;
  S->connected = false;
  S->nb_peer = 0;
  return AFB_OK;
}

void p2p_destroy(afb_ctx* ctx)
{
  if (!ctx->p2p) return;
  p2p_disconnect(ctx);
  P2PState* S = static_cast<P2PState*>(ctx->p2p);
  if (S->flags) cudaFree(S->flags);
  if (S->h_err) cudaFreeHost(S->h_err);
  if (S->ev_ready) cudaEventDestroy(S->ev_ready);
  if (S->ev_done) cudaEventDestroy(S->ev_done);
  if (S->side) cudaStreamDestroy(S->side);
  delete S;
  ctx->p2p = nullptr;
}

int p2p_connect(afb_ctx* ctx, int my_rank, int nb_peer, const int32_t* peer_rank, const void* values_handles, const void* flags_handles, const int64_t* pull_first,
                const int64_t* pull_count, const int64_t* const* slots, const int64_t* send_first, const int64_t* send_count)
{
  return p2p_connect_ex(ctx, my_rank, nb_peer, peer_rank, values_handles, flags_handles, nullptr, pull_first, pull_count, slots, send_first, send_count);
}

int p2p_connect_ex(afb_ctx* ctx, int my_rank, int nb_peer, const int32_t* peer_rank, const void* values_handles, const void* flags_handles, const P2PEndpoint* endpoints,
                   const int64_t* pull_first, const int64_t* pull_count, const int64_t* const* slots, const int64_t* send_first, const int64_t* send_count)
{
  P2PState* S = state_of(ctx);
  AFB_REQUIRE(S->flags && S->values_base == ctx->values.p, AFB_ERR_INVALID, "afb_p2p_connect: call afb_p2p_export first (and again after the values array moved)");
  AFB_REQUIRE(my_rank >= 0 && my_rank < P2P_MAX_RANK && nb_peer >= 0 && nb_peer <= P2P_MAX_RANK, AFB_ERR_INVALID, "afb_p2p_connect: rank/peer count out of range (max %d)", P2P_MAX_RANK);
  AFB_TRY(p2p_disconnect(ctx));
  std::vector<PeerDev> h((size_t)nb_peer);
  const cudaIpcMemHandle_t* vh = static_cast<const cudaIpcMemHandle_t*>(values_handles);
  const cudaIpcMemHandle_t* fh = static_cast<const cudaIpcMemHandle_t*>(flags_handles);
  for (int k = 0; k < nb_peer; ++k) {
    AFB_REQUIRE(peer_rank[k] >= 0 && peer_rank[k] < P2P_MAX_RANK && peer_rank[k] != my_rank, AFB_ERR_INVALID, "afb_p2p_connect: bad peer rank %d", peer_rank[k]);
    void *pv = nullptr, *pf = nullptr;
    if (endpoints && endpoints[k].pid == (uint64_t)getpid()) { // same process: the peer's arrays are directly addressable
      if (endpoints[k].device != ctx->device) {
        cudaError_t pe = cudaDeviceEnablePeerAccess(endpoints[k].device, 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
        else AFB_CUDA(pe);
      }
      pv = reinterpret_cast<void*>((uintptr_t)endpoints[k].values_ptr);
      pf = reinterpret_cast<void*>((uintptr_t)endpoints[k].flags_ptr);
    }
    else {
      const cudaIpcMemHandle_t* hv = endpoints ? reinterpret_cast<const cudaIpcMemHandle_t*>(endpoints[k].values_handle) : vh + k;
      const cudaIpcMemHandle_t* hf = endpoints ? reinterpret_cast<const cudaIpcMemHandle_t*>(endpoints[k].flags_handle) : fh + k;
      AFB_CUDA(cudaIpcOpenMemHandle(&pv, *hv, cudaIpcMemLazyEnablePeerAccess));
      S->opened.push_back(pv);
      AFB_CUDA(cudaIpcOpenMemHandle(&pf, *hf, cudaIpcMemLazyEnablePeerAccess));
      S->opened.push_back(pf);
    }
    h[k].peer_values = static_cast<const double*>(pv);
    h[k].peer_flags = static_cast<uint32_t*>(pf);
    h[k].slots = slots[k];
    h[k].pull_first = pull_first[k];
    h[k].pull_n = pull_count[k];
    h[k].send_first = send_first[k];
    h[k].send_n = send_count[k];
    h[k].rank = peer_rank[k];
    h[k].pad = 0;
    AFB_REQUIRE(send_first[k] >= 0 && send_first[k] + send_count[k] <= (int64_t)ctx->nnz * ctx->b * ctx->b, AFB_ERR_INVALID, "afb_p2p_connect: send slice outside values");
  }
  if (nb_peer > 0) {
    AFB_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->d_peers), sizeof(PeerDev) * (size_t)nb_peer));
    AFB_CUDA(cudaMemcpy(S->d_peers, h.data(), sizeof(PeerDev) * (size_t)nb_peer, cudaMemcpyHostToDevice));
    AFB_CUDA(cudaMalloc(reinterpret_cast<void**>(&S->counters), sizeof(int) * 2 * (size_t)nb_peer));
    AFB_CUDA(cudaMemset(S->counters, 0, sizeof(int) * 2 * (size_t)nb_peer));
    // the kernel spin-waits on remote progress: its whole grid must be resident at once
    int per_sm = 0;
    AFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_p2p_exchange, P2P_THREADS, 0));
    const int resident = per_sm * ctx->sm_count;
    S->blocks_per_peer = std::min(P2P_BLOCKS_PER_PEER, resident / nb_peer);
    AFB_REQUIRE(S->blocks_per_peer >= 1, AFB_ERR_INVALID, "afb_p2p_connect: %d peers exceed the %d exchange blocks the device keeps resident", nb_peer, resident);
  }
  S->my_rank = my_rank;
  S->nb_peer = nb_peer;
  S->connected = true;
  return AFB_OK;
}

int p2p_wait(afb_ctx* ctx)
{
  P2PState* S = ctx->p2p ? static_cast<P2PState*>(ctx->p2p) : nullptr;
  if (!S) return AFB_OK;
  if (S->inflight) {
    AFB_CUDA(cudaStreamWaitEvent(ctx->stream, S->ev_done, 0));
    S->inflight = false;
  }
  // sticky error word of the exchanges that have completed so far (mapped host memory: no synchronisation);
  // afb_p2p_status reads it after draining the stream and clears it
  if (S->h_err && *reinterpret_cast<volatile uint32_t*>(S->h_err) != 0u)
    AFB_REQUIRE(false, AFB_ERR_CUDA, "ghost-row exchange timed out (status %u: %s); contributions of that step are incomplete -- afb_p2p_status reads and clears the error",
                *S->h_err, *S->h_err == 1u ? "a neighbour's rows never became ready" : "a neighbour never acknowledged the pull");
  return AFB_OK;
}

int p2p_exchange(afb_ctx* ctx, int async)
{
  P2PState* S = ctx->p2p ? static_cast<P2PState*>(ctx->p2p) : nullptr;
  AFB_REQUIRE(S && S->connected, AFB_ERR_INVALID, "afb_p2p_exchange: not connected");
  AFB_REQUIRE(S->values_base == ctx->values.p, AFB_ERR_INVALID, "afb_p2p_exchange: the values array moved since afb_p2p_export (re-export and re-connect)");
  AFB_TRY(p2p_wait(ctx)); // one exchange in flight at a time
  S->epoch++;
  if (S->nb_peer == 0) return AFB_OK;
  cudaStream_t st = ctx->stream;
  if (async) {
    if (!S->side) {
      int lo = 0, hi = 0;
      AFB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      AFB_CUDA(cudaStreamCreateWithPriority(&S->side, cudaStreamNonBlocking, hi));
      AFB_CUDA(cudaEventCreateWithFlags(&S->ev_ready, cudaEventDisableTiming));
      AFB_CUDA(cudaEventCreateWithFlags(&S->ev_done, cudaEventDisableTiming));
    }
    AFB_CUDA(cudaEventRecord(S->ev_ready, ctx->stream));
    AFB_CUDA(cudaStreamWaitEvent(S->side, S->ev_ready, 0));
    st = S->side;
  }
  k_p2p_exchange<<<S->nb_peer * S->blocks_per_peer, P2P_THREADS, 0, st>>>(S->d_peers, ctx->values.as<double>(), S->flags, S->my_rank, S->epoch, S->counters,
                                                                          S->blocks_per_peer, S->d_err, S->wait_ns);
  AFB_LAUNCH_CHECK(ctx);
  if (async) {
    AFB_CUDA(cudaEventRecord(S->ev_done, S->side));
    S->inflight = true;
  }
  return AFB_OK;
}

// time the exchange kernels since the last call spent waiting for the neighbours (synchronises; read and clear)
int p2p_wait_stats(afb_ctx* ctx, double* ready_wait_us, double* pulled_wait_us, int64_t* nb_exchange)
{
  P2PState* S = ctx->p2p ? static_cast<P2PState*>(ctx->p2p) : nullptr;
  *ready_wait_us = *pulled_wait_us = 0.0;
  *nb_exchange = 0;
  if (!S || !S->wait_ns) return AFB_OK;
  if (S->inflight) {
    AFB_CUDA(cudaStreamWaitEvent(ctx->stream, S->ev_done, 0));
    S->inflight = false;
  }
  unsigned long long w[2] = { 0ull, 0ull };
  AFB_CUDA(cudaMemcpyAsync(w, S->wait_ns, sizeof(w), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaMemsetAsync(S->wait_ns, 0, sizeof(w), ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  *ready_wait_us = 1e-3 * (double)w[0];
  *pulled_wait_us = 1e-3 * (double)w[1];
  *nb_exchange = (int64_t)S->epoch - S->stats_epoch0;
  S->stats_epoch0 = (long long)S->epoch;
  return AFB_OK;
}

// 0 = fine, 1 = a neighbour's rows never became ready, 2 = a neighbour never acknowledged the pull (after the stream drained)
int p2p_status(afb_ctx* ctx, int* status)
{
  P2PState* S = ctx->p2p ? static_cast<P2PState*>(ctx->p2p) : nullptr;
  *status = 0;
  if (!S || !S->flags) return AFB_OK;
  uint32_t e = 0;
  if (S->inflight) {
    AFB_CUDA(cudaStreamWaitEvent(ctx->stream, S->ev_done, 0));
    S->inflight = false;
  }
  AFB_CUDA(cudaMemcpyAsync(&e, S->flags + 2 * P2P_MAX_RANK, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaMemsetAsync(S->flags + 2 * P2P_MAX_RANK, 0, sizeof(uint32_t), ctx->stream)); // read-and-clear
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (S->h_err) *S->h_err = 0u;
  *status = (int)e;
  return AFB_OK;
}

} // namespace afb

