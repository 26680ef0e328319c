// Scratch micro-benchmark (not product code): feasibility of the tiled-gather executor.
// Synthetic plan with the statistics of the C2 box: tiles of 216 rows, 2058 cells, 3240 entries.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int T = 512;
constexpr int NW = T / 32;

struct TileDesc { int cell_off, C, unit_off, nunit, cs; };

__device__ __forceinline__ void tet_k(const double* __restrict__ coords, int4 nd, double (&K)[10])
{
  const double* p0 = coords + 3 * (int64_t)nd.x; const double* p1 = coords + 3 * (int64_t)nd.y;
  const double* p2 = coords + 3 * (int64_t)nd.z; const double* p3 = coords + 3 * (int64_t)nd.w;
  double x0 = __ldg(p0), y0 = __ldg(p0 + 1), z0 = __ldg(p0 + 2);
  double e1x = __ldg(p1) - x0, e1y = __ldg(p1 + 1) - y0, e1z = __ldg(p1 + 2) - z0;
  double e2x = __ldg(p2) - x0, e2y = __ldg(p2 + 1) - y0, e2z = __ldg(p2 + 2) - z0;
  double e3x = __ldg(p3) - x0, e3y = __ldg(p3 + 1) - y0, e3z = __ldg(p3 + 2) - z0;
  double c1x = e2y * e3z - e2z * e3y, c1y = e2z * e3x - e2x * e3z, c1z = e2x * e3y - e2y * e3x;
  double c2x = e3y * e1z - e3z * e1y, c2y = e3z * e1x - e3x * e1z, c2z = e3x * e1y - e3y * e1x;
  double c3x = e1y * e2z - e1z * e2y, c3y = e1z * e2x - e1x * e2z, c3z = e1x * e2y - e1y * e2x;
  double c0x = -(c1x + c2x + c3x), c0y = -(c1y + c2y + c3y), c0z = -(c1z + c2z + c3z);
  double det = fabs(e1x * c1x + e1y * c1y + e1z * c1z);
  double s = 1.0 / (6.0 * det);
  K[0] = (c0x * c0x + c0y * c0y + c0z * c0z) * s; K[1] = (c0x * c1x + c0y * c1y + c0z * c1z) * s;
  K[2] = (c0x * c2x + c0y * c2y + c0z * c2z) * s; K[3] = (c0x * c3x + c0y * c3y + c0z * c3z) * s;
  K[4] = (c1x * c1x + c1y * c1y + c1z * c1z) * s; K[5] = (c1x * c2x + c1y * c2y + c1z * c2z) * s;
  K[6] = (c1x * c3x + c1y * c3y + c1z * c3z) * s; K[7] = (c2x * c2x + c2y * c2y + c2z * c2z) * s;
  K[8] = (c2x * c3x + c2y * c3y + c2z * c3z) * s; K[9] = (c3x * c3x + c3y * c3y + c3z * c3z) * s;
}

template <int MODE>  // 0 = both, 1 = phase A only, 2 = phase B only
__global__ void __launch_bounds__(T, 1)
k_tiled(const TileDesc* __restrict__ tiles, int ntile, const double* __restrict__ coords, const int4* __restrict__ tcells,
        const uint32_t* __restrict__ unit_base, const uint16_t* __restrict__ unit_len, const uint16_t* __restrict__ lists,
        const uint32_t* __restrict__ gpos, double* __restrict__ values)
{
  extern __shared__ double Kc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
    const TileDesc d = tiles[t];
    if (MODE != 2) {
      for (int lc = threadIdx.x; lc < d.C; lc += T) {
        int4 nd = __ldg(tcells + d.cell_off + lc);
        double K[10];
        tet_k(coords, nd, K);
#pragma unroll
        for (int p = 0; p < 10; ++p) Kc[p * d.cs + lc] = K[p];
      }
    }
    if (threadIdx.x == 0) Kc[10 * d.cs] = 0.0;
    __syncthreads();
    if (MODE != 1) {
      for (int u = warp; u < d.nunit; u += NW) {
        const uint32_t base = __ldg(unit_base + d.unit_off + u);
        const int len = __ldg(unit_len + d.unit_off + u);
        const uint16_t* l = lists + base + lane;
        double acc0 = 0.0, acc1 = 0.0;
        int k = 0;
        for (; k + 4 <= len; k += 4) {
          uint16_t i0 = __ldg(l + (k + 0) * 32), i1 = __ldg(l + (k + 1) * 32), i2 = __ldg(l + (k + 2) * 32), i3 = __ldg(l + (k + 3) * 32);
          acc0 += Kc[i0]; acc1 += Kc[i1]; acc0 += Kc[i2]; acc1 += Kc[i3];
        }
        for (; k < len; ++k) acc0 += Kc[__ldg(l + k * 32)];
        const uint32_t g = __ldg(gpos + (size_t)(d.unit_off + u) * 32 + lane);
        if (g != 0xFFFFFFFFu) values[g] = acc0 + acc1;
      }
    }
    __syncthreads();
  }
}

int main(int argc, char** argv)
{
  const int n = 120, m = n + 1;
  const int nb_node = m * m * m;
  const int R = 216, C = 2058, E = 3240;
  const int ntile = nb_node / R;
  const int ndiag_units = (R + 31) / 32, noff_units = (E - R + 31) / 32;
  const int nunit = ndiag_units + noff_units;
  const int cs = C | 1;  // odd stride
  std::vector<double> coords(3 * (size_t)nb_node);
  for (auto& c : coords) c = rand() / (double)RAND_MAX;
  std::vector<TileDesc> tiles(ntile);
  std::vector<int4> tcells((size_t)ntile * C);
  std::vector<uint32_t> ubase((size_t)ntile * nunit), gpos((size_t)ntile * nunit * 32);
  std::vector<uint16_t> ulen((size_t)ntile * nunit);
  size_t list_total = 0;
  for (int t = 0; t < ntile; ++t)
    for (int u = 0; u < nunit; ++u) {
      int len = u < ndiag_units ? 24 : 6;
      ubase[(size_t)t * nunit + u] = (uint32_t)list_total;
      ulen[(size_t)t * nunit + u] = (uint16_t)len;
      list_total += (size_t)len * 32;
    }
  std::vector<uint16_t> lists(list_total);
  for (auto& v : lists) v = (uint16_t)(rand() % (10 * cs));
  size_t nnz = (size_t)ntile * E;
  for (int t = 0; t < ntile; ++t) {
    tiles[t] = { t * C, C, t * nunit, nunit, cs };
    for (int c = 0; c < C; ++c) {
      // nodes within a window of the tile's footprint (brick of ~8^3 around), clamped
      int base = (int)(((int64_t)t * R) % (nb_node - 3 * m * m));
      int4 nd;
      nd.x = base + rand() % 8 + m * (rand() % 8) + m * m * (rand() % 3);
      nd.y = nd.x + 1; nd.z = nd.x + m; nd.w = nd.x + m * m + (rand() % 2);
      tcells[(size_t)t * C + c] = nd;
    }
    for (int e = 0; e < nunit * 32; ++e) gpos[(size_t)t * nunit * 32 + e] = e < E ? (uint32_t)((size_t)t * E + e) : 0xFFFFFFFFu;
  }
  printf("ntile %d, cells(tile-instances) %zu, contributions %zu (%.1f per tet of 10.368M), nnz %zu\n", ntile, (size_t)ntile * C, list_total, list_total / 10.368e6, nnz);
  double *d_coords, *d_values; TileDesc* d_tiles; int4* d_tcells; uint32_t *d_ubase, *d_gpos; uint16_t *d_ulen, *d_lists;
  CK(cudaMalloc(&d_coords, coords.size() * 8)); CK(cudaMalloc(&d_values, nnz * 8)); CK(cudaMalloc(&d_tiles, tiles.size() * sizeof(TileDesc)));
  CK(cudaMalloc(&d_tcells, tcells.size() * 16)); CK(cudaMalloc(&d_ubase, ubase.size() * 4)); CK(cudaMalloc(&d_gpos, gpos.size() * 4));
  CK(cudaMalloc(&d_ulen, ulen.size() * 2)); CK(cudaMalloc(&d_lists, lists.size() * 2));
  CK(cudaMemcpy(d_coords, coords.data(), coords.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_tcells, tcells.data(), tcells.size() * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ubase, ubase.data(), ubase.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_gpos, gpos.data(), gpos.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ulen, ulen.data(), ulen.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_lists, lists.data(), lists.size() * 2, cudaMemcpyHostToDevice));
  size_t smem = (size_t)(10 * cs + 1) * 8;
  CK(cudaFuncSetAttribute(k_tiled<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_tiled<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_tiled<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = 148;
  for (int mode = 0; mode < 3; ++mode) {
    for (int it = 0; it < 6; ++it) {
      cudaEventRecord(e0);
      if (mode == 0) k_tiled<0><<<grid, T, smem>>>(d_tiles, ntile, d_coords, d_tcells, d_ubase, d_ulen, d_lists, d_gpos, d_values);
      if (mode == 1) k_tiled<1><<<grid, T, smem>>>(d_tiles, ntile, d_coords, d_tcells, d_ubase, d_ulen, d_lists, d_gpos, d_values);
      if (mode == 2) k_tiled<2><<<grid, T, smem>>>(d_tiles, ntile, d_coords, d_tcells, d_ubase, d_ulen, d_lists, d_gpos, d_values);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 3) printf("mode %d: %.3f ms  (%.1f G tets/s algorithmic @10.368M)\n", mode, ms, 10.368e6 / ms / 1e6);
    }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
