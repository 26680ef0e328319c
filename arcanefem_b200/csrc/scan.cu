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

int exclusive_scan_i32(afb_ctx* ctx, const int32_t* in, int32_t* out, int64_t n)
{
  if (n <= 0) {
    AFB_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), ctx->stream));
    return AFB_OK;
  }
  int nblk = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
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
