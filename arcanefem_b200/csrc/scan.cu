// Exclusive prefix sum of int32 (row_index / node->cell offsets).
// Replaces Arcane's Accelerator::Scanner (cub::DeviceScan) calls of the reference:
// modules/testlab/CsrGpuBiliAssembly.cc:124-126, femutils/BSRFormat.cc:571-572,921-923.
// Reduce-then-scan: one pass of block sums, one single-block scan of the sums, one pass
// that rescans each tile with its offset.  2 reads + 1 write of n ints; the inputs here
// are <= 68 MB, i.e. L2 resident between the passes on B200 (126 MB L2).
#include "afb_internal.h"

namespace afb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_inclusive_scan(int v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// exclusive scan of one value per thread over the block; returns exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* total)
{
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = warp_inclusive_scan(v);
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int nw = (blockDim.x + 31) >> 5;
    int s = lane < nw ? warp_sums[lane] : 0;
    int si = warp_inclusive_scan(s);
    warp_sums[lane] = si - s; // exclusive
    if (lane == 31) *total = si;
  }
  __syncthreads();
  int off = warp_sums[wid];
  __syncthreads();
  return off + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ sums)
{
  __shared__ int total;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += in[i];
  }
  (void)block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(int32_t* sums, int nblk, int32_t* grand_total)
{
  __shared__ int total;
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nblk ? sums[i] : 0;
    int ex = block_exclusive_scan(v, &total);
    int c = carry;
    if (i < nblk) sums[i] = ex + c;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ sums)
{
  __shared__ int total;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    v[k] = i < n ? in[i] : 0;
    s += v[k];
  }
  int ex = block_exclusive_scan(s, &total) + sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
}

// ---- single pass: chained scan with decoupled look-back (1 read + 1 write of n ints, one launch) -------------
// Tile state word: epoch << 34 | flag << 32 | value; flag 1 = the tile's own sum, 2 = inclusive prefix up to and
// including the tile.  The epoch (one per call) makes the words of earlier calls read as "not ready", so the
// state array is never cleared; tiles are handed out by a monotonic ticket counter (a tile only waits for tiles
// that already run).
constexpr unsigned long long SC_FLAG_SUM = 1ull, SC_FLAG_PREFIX = 2ull;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_chained(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, unsigned long long* __restrict__ state, unsigned* __restrict__ ticket,
               unsigned ticket_base, unsigned epoch, int nblk)
{
  __shared__ int total;
  __shared__ unsigned s_tile;
  __shared__ int s_prefix;
  pdl_enter();
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u) - ticket_base;
  __syncthreads();
  const int tile = (int)s_tile;
  const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
  if (base + SCAN_ITEMS <= n) {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k += 4) {
      const int4 q = *reinterpret_cast<const int4*>(in + base + k);
      v[k] = q.x; v[k + 1] = q.y; v[k + 2] = q.z; v[k + 3] = q.w;
    }
  }
  else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = base + k < n ? in[base + k] : 0;
  }
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) s += v[k];
  const int ex = block_exclusive_scan(s, &total);
  const unsigned long long tag = (unsigned long long)epoch << 34;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    volatile unsigned long long* st = state;
    int prefix = 0;
    if (tile == 0) {
      if (lane == 0) st[0] = tag | (SC_FLAG_PREFIX << 32) | (unsigned)total;
    }
    else {
      if (lane == 0) st[tile] = tag | (SC_FLAG_SUM << 32) | (unsigned)total;
      int look = tile - 1;
      while (true) { // 32 predecessors per round
        const int j = look - lane;
        unsigned long long w = 0;
        bool ready = j < 0;
        while (!ready) {
          w = st[j];
          ready = (w >> 34) == (unsigned long long)epoch && ((w >> 32) & 3ull) != 0ull;
        }
        const bool is_prefix = j >= 0 && ((w >> 32) & 3ull) == SC_FLAG_PREFIX;
        const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
        const int stop = pm ? __ffs(pm) - 1 : 32; // nearest predecessor holding an inclusive prefix
        int part = (j >= 0 && lane <= stop) ? (int)(unsigned)(w & 0xFFFFFFFFull) : 0;
#pragma unroll
        for (int d2 = 16; d2 > 0; d2 >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d2);
        prefix += part;
        if (pm || look - 32 < 0) break;
        look -= 32;
      }
      if (lane == 0) st[tile] = tag | (SC_FLAG_PREFIX << 32) | (unsigned)(prefix + total);
    }
    if (lane == 0) {
      s_prefix = prefix;
      if (tile == nblk - 1) out[n] = prefix + total;
    }
  }
  __syncthreads();
  int run = s_prefix + ex;
  if (base + SCAN_ITEMS <= n) {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k += 4) {
      int4 q;
      q.x = run; run += v[k];
      q.y = run; run += v[k + 1];
      q.z = run; run += v[k + 2];
      q.w = run; run += v[k + 3];
      *reinterpret_cast<int4*>(out + base + k) = q;
    }
  }
  else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
      if (base + k < n) out[base + k] = run;
      run += v[k];
    }
  }
}

int exclusive_scan_i32(afb_ctx* ctx, const int32_t* in, int32_t* out, int64_t n)
{
  if (n <= 0) {
    AFB_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), ctx->stream));
    return AFB_OK;
  }
  const int nblk = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (aligned) {
    const size_t need = sizeof(unsigned long long) * ((size_t)nblk + 2);
    if (ctx->scan_state.cap < need || !ctx->scan_state.p) { // fresh state words: epoch 0 never matches (epochs start at 1)
      AFB_TRY(ctx->scan_state.reserve(need));
      AFB_CUDA(cudaMemsetAsync(ctx->scan_state.p, 0, ctx->scan_state.cap, ctx->stream));
      ctx->scan_tickets = 0;
      ctx->scan_epoch = 0;
    }
    unsigned long long* state = ctx->scan_state.as<unsigned long long>() + 1;
    unsigned* ticket = ctx->scan_state.as<unsigned>();
    ctx->scan_epoch = ctx->scan_epoch % 0x3FFFFFFFu + 1u;
    AFB_CUDA(launch_pdl(k_scan_chained, nblk, SCAN_THREADS, 0, ctx->stream, in, out, n, state, ticket, ctx->scan_tickets, ctx->scan_epoch, nblk));
    AFB_LAUNCH_CHECK(ctx);
    ctx->scan_tickets += (unsigned)nblk;
    return AFB_OK;
  }
  AFB_TRY(ctx->tmp_scan.reserve(sizeof(int32_t) * (size_t)(nblk + 1)));
  int32_t* sums = ctx->tmp_scan.as<int32_t>();
  k_scan_reduce<<<nblk, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums);
  AFB_LAUNCH_CHECK(ctx);
  k_scan_sums<<<1, 1024, 0, ctx->stream>>>(sums, nblk, out + n);
  AFB_LAUNCH_CHECK(ctx);
  k_scan_apply<<<nblk, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, sums);
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

} // namespace afb
