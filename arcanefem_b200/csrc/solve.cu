// Jacobi-preconditioned conjugate gradient on the assembled matrix: the solve the reference hands to
// HYPRE / PETSc (femutils/HypreDoFLinearSystem.cc:461-520: PCG + preconditioner on the IJ matrix built from the same
// CSR arrays), kept in-repo so that the reference's golden solution files can be checked end to end on the GPU box.
// It works on the arrays exactly as assembled -- CSR (b = 1) or BSR in either value layout -- and on the RHS vector of
// the context; symmetric positive definite systems only (Poisson, elasticity with penalty / eliminated Dirichlet rows;
// not the bilaplacian saddle-point system).
//
// Per iteration three launches, no host round trip: the step lengths are computed on the device from dot products the
// kernels accumulate (fp64 atomics), the convergence test (preconditioned residual, sqrt(r.z)) is read back every
// PCG_CHECK iterations.
//   k_spmv_dot   q = A p, pq += p.q        8 lanes per block row, coalesced over 4 consecutive rows per warp
//   k_pcg_update alpha = rho/pq;  x += alpha p;  r -= alpha q;  rho' += r . (Dinv r)
//   k_pcg_dir    beta = rho'/rho;  p = Dinv r + beta p
#include <algorithm>
#include <cmath>
#include <type_traits>

#include "element.cuh"

namespace afb {

constexpr int PCG_CHECK = 16;
constexpr int SPMV_LANES = 8;

// scalars: rho[3] (rotating: old / new / being zeroed), pq -- each spread over PCG_SPREAD addresses (one fp64 atomic per
// block, blocks hashed over the slots: a single address would serialise ~5e4 atomics per kernel in one L2 slice)
constexpr int PCG_SPREAD = 16;
struct PcgScalars {
  double rho[3][PCG_SPREAD];
  double pq[PCG_SPREAD];
};

__device__ __forceinline__ double spread_sum(const double* __restrict__ v)
{
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < PCG_SPREAD; ++i) s += v[i];
  return s;
}

// block-wide sum of `part`, added once per block to slot blockIdx % PCG_SPREAD of `out`
__device__ __forceinline__ void block_accumulate(double part, double* __restrict__ out)
{
  __shared__ double s_part[8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_part[w];
    if (s != 0.0) atomicAdd(out + (blockIdx.x % PCG_SPREAD), s);
  }
}

template <int B, int LAYOUT>
__global__ void __launch_bounds__(256) k_spmv_dot(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, const double* __restrict__ values,
                                                   const double* __restrict__ x, double* __restrict__ y, int32_t nb_row, double* __restrict__ dot_out)
{
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t r = (int32_t)(g / SPMV_LANES);
  const int q = (int)(g % SPMV_LANES);
  double acc[B];
#pragma unroll
  for (int i = 0; i < B; ++i) acc[i] = 0.0;
  if (r < nb_row) {
    const int rb = __ldg(rows + r), re = __ldg(rows + r + 1), nz = re - rb;
    for (int p = rb + q; p < re; p += SPMV_LANES) {
      const int32_t c = __ldg(cols + p);
      double xv[B];
#pragma unroll
      for (int j = 0; j < B; ++j) xv[j] = __ldg(x + (int64_t)c * B + j);
#pragma unroll
      for (int i = 0; i < B; ++i)
#pragma unroll
        for (int j = 0; j < B; ++j) acc[i] = fma(__ldg(values + value_index<B, LAYOUT>(rb, nz, p, i, j)), xv[j], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < B; ++i) {
#pragma unroll
    for (int d = SPMV_LANES / 2; d > 0; d >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], d);
  }
  double part = 0.0;
  if (r < nb_row && q == 0) {
#pragma unroll
    for (int i = 0; i < B; ++i) {
      y[(int64_t)r * B + i] = acc[i];
      part = fma(acc[i], __ldg(x + (int64_t)r * B + i), part);
    }
  }
  if (dot_out) block_accumulate(part, dot_out);
}

// inverse diagonal (Jacobi): dinv[dof] = 1 / A[dof,dof]
template <int B, int LAYOUT>
__global__ void __launch_bounds__(256) k_inverse_diagonal(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, const double* __restrict__ values,
                                                           int32_t nb_row, double* __restrict__ dinv, int* __restrict__ bad)
{
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nb_row) return;
  const int rb = rows[r], re = rows[r + 1];
  const int p = re > rb ? find_col(cols, rb, re, r) : rb;
  if (re <= rb || cols[p] != r) {
    atomicExch(bad, 1);
    return;
  }
#pragma unroll
  for (int i = 0; i < B; ++i) {
    const double d = values[value_index<B, LAYOUT>(rb, re - rb, p, i, i)];
    if (!(d > 0.0)) atomicExch(bad, 2);
    dinv[(int64_t)r * B + i] = 1.0 / d;
  }
}

// initial guess x0 = Dinv b: rows carrying a Dirichlet penalty (diagonal 1e30, modules/testlab/FemModule.cc:728-790) are then
// satisfied from the start, so that the initial residual -- the reference of the relative stopping test -- has the scale of
// the free rows instead of the penalty's
__global__ void __launch_bounds__(256) k_pcg_guess(const double* __restrict__ b, const double* __restrict__ dinv, double* __restrict__ x, int64_t n)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = dinv[i] * b[i];
}

// r = b - q (q = A x0), z = Dinv r, p = z, rho[0] = r.z
__global__ void __launch_bounds__(256) k_pcg_init(const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv, double* __restrict__ r,
                                                   double* __restrict__ p, int64_t n, PcgScalars* __restrict__ S)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double part = 0.0;
  if (i < n) {
    const double ri = b[i] - q[i], zi = dinv[i] * ri;
    r[i] = ri;
    p[i] = zi;
    part = ri * zi;
  }
  block_accumulate(part, S->rho[0]);
}

__global__ void __launch_bounds__(256) k_pcg_update(const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ dinv, double* __restrict__ x,
                                                     double* __restrict__ r, int64_t n, PcgScalars* __restrict__ S, int k)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double rho = spread_sum(S->rho[k % 3]), pq = spread_sum(S->pq);
  const double alpha = pq != 0.0 ? rho / pq : 0.0;
  double part = 0.0;
  if (i < n) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    part = ri * (dinv[i] * ri);
  }
  block_accumulate(part, S->rho[(k + 1) % 3]);
}

__global__ void __launch_bounds__(256) k_pcg_dir(const double* __restrict__ r, const double* __restrict__ dinv, double* __restrict__ p, int64_t n,
                                                  PcgScalars* __restrict__ S, int k)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double rho = spread_sum(S->rho[k % 3]), rho_new = spread_sum(S->rho[(k + 1) % 3]);
  const double beta = rho != 0.0 ? rho_new / rho : 0.0;
  if (i < n) p[i] = fma(beta, p[i], dinv[i] * r[i]);
  if (i < PCG_SPREAD) { // nobody reads these two any more: pq was consumed by k_pcg_update, rho[(k+2)%3] was the "old" of iteration k-1
    S->pq[i] = 0.0;
    S->rho[(k + 2) % 3][i] = 0.0;
  }
}

template <class F> static int by_layout(int b, int layout, F f)
{
  const bool blk = layout == AFB_LAYOUT_PER_BLOCK;
  switch (b) {
  case 1: return f(std::integral_constant<int, 1>(), std::integral_constant<int, AFB_LAYOUT_PER_BLOCK>());
  case 2: return blk ? f(std::integral_constant<int, 2>(), std::integral_constant<int, AFB_LAYOUT_PER_BLOCK>()) : f(std::integral_constant<int, 2>(), std::integral_constant<int, AFB_LAYOUT_PER_ROW>());
  default: return blk ? f(std::integral_constant<int, 3>(), std::integral_constant<int, AFB_LAYOUT_PER_BLOCK>()) : f(std::integral_constant<int, 3>(), std::integral_constant<int, AFB_LAYOUT_PER_ROW>());
  }
}

int spmv(afb_ctx* ctx, const double* x, double* y, double* dot_out)
{
  const int32_t nb_row = ctx->nb_node;
  if (nb_row == 0) return AFB_OK;
  const int grid = grid_for((int64_t)nb_row * SPMV_LANES, 256);
  return by_layout(ctx->b, ctx->layout, [&](auto B, auto L) {
    k_spmv_dot<decltype(B)::value, decltype(L)::value><<<grid, 256, 0, ctx->stream>>>(ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), ctx->values.as<double>(), x, y, nb_row, dot_out);
    AFB_LAUNCH_CHECK(ctx);
    return AFB_OK;
  });
}

int solve_pcg(afb_ctx* ctx, double rtol, double atol, int max_iter, double* x_out, int mem_space, int* iterations, double* residual)
{
  const int64_t n = (int64_t)ctx->nb_node * ctx->b;
  if (iterations) *iterations = 0;
  if (residual) *residual = 0.0;
  if (n == 0) return AFB_OK;
  cudaStream_t st = ctx->stream;
  AFB_TRY(ctx->solver_work.reserve(sizeof(double) * (size_t)(5 * n) + sizeof(PcgScalars) + 64));
  double* x = ctx->solver_work.as<double>();
  double *r = x + n, *p = r + n, *q = p + n, *dinv = q + n;
  PcgScalars* S = reinterpret_cast<PcgScalars*>(dinv + n);
  int* bad = reinterpret_cast<int*>(S + 1);
  AFB_CUDA(cudaMemsetAsync(S, 0, sizeof(PcgScalars) + sizeof(int), st));
  AFB_TRY(by_layout(ctx->b, ctx->layout, [&](auto B, auto L) {
    k_inverse_diagonal<decltype(B)::value, decltype(L)::value><<<grid_for(ctx->nb_node, 256), 256, 0, st>>>(ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), ctx->values.as<double>(), ctx->nb_node, dinv, bad);
    AFB_LAUNCH_CHECK(ctx);
    return AFB_OK;
  }));
  const int grid = grid_for(n, 256);
  k_pcg_guess<<<grid, 256, 0, st>>>(ctx->rhs.as<double>(), dinv, x, n);
  AFB_LAUNCH_CHECK(ctx);
  AFB_TRY(spmv(ctx, x, q, nullptr));
  k_pcg_init<<<grid, 256, 0, st>>>(ctx->rhs.as<double>(), q, dinv, r, p, n, S);
  AFB_LAUNCH_CHECK(ctx);
  struct { PcgScalars s; int bad; } h;
  AFB_CUDA(cudaMemcpyAsync(&h, S, sizeof(PcgScalars) + sizeof(int), cudaMemcpyDeviceToHost, st));
  AFB_CUDA(cudaStreamSynchronize(st));
  AFB_REQUIRE(h.bad == 0, AFB_ERR_INVALID, h.bad == 1 ? "afb_solve_pcg: a row has no diagonal entry" : "afb_solve_pcg: non-positive diagonal (the Jacobi-PCG needs an SPD matrix)");
  auto host_sum = [](const double* v) { double t = 0.0; for (int i = 0; i < PCG_SPREAD; ++i) t += v[i]; return t; };
  const double rho0 = host_sum(h.s.rho[0]);
  const double target = std::max(rtol * sqrt(rho0 > 0.0 ? rho0 : 0.0), atol);
  double res = sqrt(rho0 > 0.0 ? rho0 : 0.0);
  int k = 0;
  while (res > target && k < max_iter) {
    const int stop = std::min(max_iter, k + PCG_CHECK);
    for (; k < stop; ++k) {
      AFB_TRY(spmv(ctx, p, q, S->pq));
      k_pcg_update<<<grid, 256, 0, st>>>(p, q, dinv, x, r, n, S, k);
      AFB_LAUNCH_CHECK(ctx);
      k_pcg_dir<<<grid, 256, 0, st>>>(r, dinv, p, n, S, k);
      AFB_LAUNCH_CHECK(ctx);
    }
    AFB_CUDA(cudaMemcpyAsync(&h.s, S, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaStreamSynchronize(st));
    const double rho = host_sum(h.s.rho[k % 3]);
    AFB_REQUIRE(rho == rho, AFB_ERR_CUDA, "afb_solve_pcg: the iteration broke down (NaN residual after %d iterations)", k);
    res = sqrt(rho > 0.0 ? rho : 0.0);
  }
  if (iterations) *iterations = k;
  if (residual) *residual = res;
  if (x_out) AFB_CUDA(cudaMemcpyAsync(x_out, x, sizeof(double) * (size_t)n, mem_space == AFB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  AFB_CUDA(cudaStreamSynchronize(st));
  AFB_REQUIRE(res <= target, AFB_ERR_INVALID, "afb_solve_pcg: not converged after %d iterations (preconditioned residual %.3e, target %.3e)", k, res, target);
  return AFB_OK;
}

} // namespace afb
