// kml_p2g_cell.cuh - cell lists for the cell-centric ULMPM kernels (3-D cubic B-splines, the BASELINE headline
// configuration).
//
// Re-binning (counting sort, no library sort): every particle's cell key is its stencil base node
// (i0,j0,k0) = (int)((x-lo)/h - 1) per axis (src/ulmpm.cpp:206-208), so all particles of a cell share the same
// 4x4x4 nodes.  k_cell_count histograms the keys (the atomic's return value is the particle's rank inside its
// cell), an exclusive scan turns counts into offsets and k_cell_fill writes the particle order.  The consumers
// are kml_p2g_cell3.cuh (scatter) and kml_gather_cell2.cuh (gathers).
#pragma once
#include "kml_kernels.cuh"
#include <cub/device/device_scan.cuh>
#include <cstring>

namespace kml {

struct CellLists {
  bool valid = false;
  long long ncells = 0, cap_np = 0;
  int *cell_of = nullptr;   // [np] cell key of each particle
  int *rank = nullptr;      // [np] rank of the particle inside its cell
  int *start = nullptr;     // [ncells + 1] counts, then exclusive offsets
  int *order = nullptr;     // [np] particle ids grouped by cell
  void *scan_tmp = nullptr; size_t scan_bytes = 0;
  int *disorder = nullptr;  // device: particles whose cell-sorted slot is more than DISORDER_FAR entries away from their storage slot
  int *h_disorder = nullptr; // pinned copy, refreshed asynchronously by every build
  long long far_count() const { long long t = 0; if (h_disorder) for (int i = 0; i < 32; i++) t += h_disorder[i * 32]; return t; }
  void far_reset() { if (h_disorder) for (int i = 0; i < 32; i++) h_disorder[i * 32] = 0; }
  int build(const SolidDev &s, const GridDev &g, long long capacity, cudaStream_t st, int *nlaunch);
  void release() {
    cudaFree(cell_of); cudaFree(rank); cudaFree(start); cudaFree(order); cudaFree(scan_tmp); cudaFree(disorder); if (h_disorder) cudaFreeHost(h_disorder);
    cell_of = rank = start = order = nullptr; scan_tmp = nullptr; disorder = nullptr; h_disorder = nullptr; valid = false; ncells = cap_np = 0; scan_bytes = 0;
  }
};

inline bool cell_p2g_supported(int dimension, int shape) { return dimension == 3 && shape == KML_SHAPE_CUBIC_SPLINE; }

__device__ __forceinline__ int cell_axis(double xp, double lo, double ih, int n, int goff) { // LOCAL stencil base
  int i0 = (int)__dsub_rn(__dmul_rn(__dsub_rn(xp, lo), ih), 1.0) - goff; // identical to axis_weights<cubic>
  return min(max(i0, 0), n - 1);
}

__global__ void k_cell_count(SolidDev s, GridDev g, int *cell_of, int *rank, int *count) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= s.np) return;
  const int i0 = cell_axis(s.x[0][ip], g.lo[0], g.inv_cellsize, g.n[0], g.goff0);
  const int j0 = cell_axis(s.x[1][ip], g.lo[1], g.inv_cellsize, g.n[1], 0);
  const int k0 = cell_axis(s.x[2][ip], g.lo[2], g.inv_cellsize, g.n[2], 0);
  const int key = (i0 * g.n[1] + j0) * g.n[2] + k0;
  cell_of[ip] = key;
  rank[ip] = atomicAdd(&count[key], 1);
}
constexpr int DISORDER_FAR = 64, DISORDER_SLOTS = 32;
__global__ void k_cell_fill(long long np, const int *cell_of, const int *rank, const int *start, int *order, int *disorder) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  bool far = false;
  if (ip < np) {
    const int dst = start[cell_of[ip]] + rank[ip];
    order[dst] = (int)ip;
    far = abs(dst - (int)ip) > DISORDER_FAR; // this particle's state is not where the cell-sorted kernels stream
  }
  // one atomic per block, spread over DISORDER_SLOTS counters in different 128-byte lines (a single address serialises in L2:
  // 3 M warp atomics cost 1.1 ms at 100 M shuffled particles)
  const int n = __syncthreads_count(far);
  if (threadIdx.x == 0 && n) atomicAdd(&disorder[(blockIdx.x % DISORDER_SLOTS) * 32], n);
}

inline int CellLists::build(const SolidDev &s, const GridDev &g, long long capacity, cudaStream_t st, int *nlaunch) {
  *nlaunch = 0; valid = false;
  if (s.np >= (1ll << 31) || g.nn >= (1ll << 31)) return 0; // 32-bit particle ids / keys
  if (g.nn != ncells || s.np > cap_np) {
    release();
    // sized for the solid's capacity (the room left for migration): the lists are never re-allocated inside a step as a slab gains particles
    ncells = g.nn; cap_np = s.np > capacity ? s.np : capacity;
    if (cudaMalloc(&cell_of, sizeof(int) * cap_np) || cudaMalloc(&rank, sizeof(int) * cap_np) || cudaMalloc(&order, sizeof(int) * cap_np) ||
        cudaMalloc(&start, sizeof(int) * (ncells + 1))) return 1;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, start, start, (int)(ncells + 1), st);
    if (cudaMalloc(&scan_tmp, scan_bytes)) return 1;
    if (cudaMalloc(&disorder, sizeof(int) * 32 * DISORDER_SLOTS) || cudaMallocHost(&h_disorder, sizeof(int) * 32 * DISORDER_SLOTS)) return 1;
    memset(h_disorder, 0, sizeof(int) * 32 * DISORDER_SLOTS);
  }
  if (cudaMemsetAsync(start, 0, sizeof(int) * (ncells + 1), st) || cudaMemsetAsync(disorder, 0, sizeof(int) * 32 * DISORDER_SLOTS, st)) return 1;
  const unsigned nb = (unsigned)((std::max<long long>(s.np, 1) + 255) / 256);
  k_cell_count<<<nb, 256, 0, st>>>(s, g, cell_of, rank, start);
  if (cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, start, start, (int)(ncells + 1), st)) return 1;
  k_cell_fill<<<nb, 256, 0, st>>>(s.np, cell_of, rank, start, order, disorder);
  if (cudaMemcpyAsync(h_disorder, disorder, sizeof(int) * 32 * DISORDER_SLOTS, cudaMemcpyDeviceToHost, st)) return 1;
  *nlaunch = 4;
  if (cudaGetLastError() != cudaSuccess) return 1;
  valid = true; return 0;
}

// one cubic-spline (value, derivative) pair of node i for a particle at xp; same expressions as axis_weights
// (i = LOCAL node index of n, global index i + goff of gn)
__device__ __forceinline__ void cubic_node(double xp, double lo, double h, double ih, int i, int n, int goff, int gn, double &w, double &dw) {
  if (i < 0 || i >= n) { w = 0; dw = 0; return; }
  const double xn = __dadd_rn(lo, __dmul_rn((double)(i + goff), h));
  const double r = __dmul_rn(__dsub_rn(xp, xn), ih);
  Basis<KML_SHAPE_CUBIC_SPLINE>::eval(r, node_type<KML_SHAPE_CUBIC_SPLINE>(i + goff, gn), ih, w, dw);
}

} // namespace kml
