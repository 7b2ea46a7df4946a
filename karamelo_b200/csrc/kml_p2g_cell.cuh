// kml_p2g_cell.cuh - cell-centric particle-to-grid for ULMPM with 3-D cubic B-splines
// (the BASELINE headline configuration).
//
// Re-binning (counting sort, no library): every particle's cell key is its stencil base node
// (i0,j0,k0) = (int)((x-lo)/h - 1) per axis (src/ulmpm.cpp:206-208), so all particles of a cell
// share the same 4x4x4 nodes.  k_cell_count histograms the keys (the atomic's return value is the
// particle's rank inside its cell), an exclusive scan turns counts into offsets and k_cell_fill
// writes the particle order.
//
// P2G: a group of 8 lanes owns one column of cells (i0,j0, k-segment).  Lane (a, b-pair) keeps the
// sums of its 2x4 stencil nodes in registers while it walks the particles of a cell (particle data
// arrives by broadcast loads, weights are recomputed per lane), then walks to the next cell along k:
// the stencil slides by one node plane, so only the completed plane is added to the grid (fp64 RED)
// and the other three stay in registers.  Per particle this issues ~14-16 atomics instead of the
// 448 of the one-thread-per-particle scatter, with the same arithmetic as the reference's node sums
// (src/solid.cpp:317-335, :337-390, :482-522) in a different summation order.
#pragma once
#include "kml_kernels.cuh"
#include <cub/device/device_scan.cuh>

namespace kml {

struct CellLists {
  bool valid = false;
  long long ncells = 0, cap_np = 0;
  int *cell_of = nullptr;   // [np] cell key of each particle
  int *rank = nullptr;      // [np] rank of the particle inside its cell
  int *start = nullptr;     // [ncells + 1] counts, then exclusive offsets
  int *order = nullptr;     // [np] particle ids grouped by cell
  void *scan_tmp = nullptr; size_t scan_bytes = 0;
  int build(const SolidDev &s, const GridDev &g, cudaStream_t st, int *nlaunch);
  void release() {
    cudaFree(cell_of); cudaFree(rank); cudaFree(start); cudaFree(order); cudaFree(scan_tmp);
    cell_of = rank = start = order = nullptr; scan_tmp = nullptr; valid = false; ncells = cap_np = 0; scan_bytes = 0;
  }
};

inline bool cell_p2g_supported(int dimension, int shape) { return dimension == 3 && shape == KML_SHAPE_CUBIC_SPLINE; }

__device__ __forceinline__ int cell_axis(double xp, double lo, double ih, int n) {
  int i0 = (int)__dsub_rn(__dmul_rn(__dsub_rn(xp, lo), ih), 1.0); // identical to axis_weights<cubic>
  return min(max(i0, 0), n - 1);
}

__global__ void k_cell_count(SolidDev s, GridDev g, int *cell_of, int *rank, int *count) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= s.np) return;
  const int i0 = cell_axis(s.x[0][ip], g.lo[0], g.inv_cellsize, g.n[0]);
  const int j0 = cell_axis(s.x[1][ip], g.lo[1], g.inv_cellsize, g.n[1]);
  const int k0 = cell_axis(s.x[2][ip], g.lo[2], g.inv_cellsize, g.n[2]);
  const int key = (i0 * g.n[1] + j0) * g.n[2] + k0;
  cell_of[ip] = key;
  rank[ip] = atomicAdd(&count[key], 1);
}
__global__ void k_cell_fill(long long np, const int *cell_of, const int *rank, const int *start, int *order) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  order[start[cell_of[ip]] + rank[ip]] = (int)ip;
}

inline int CellLists::build(const SolidDev &s, const GridDev &g, cudaStream_t st, int *nlaunch) {
  *nlaunch = 0; valid = false;
  if (s.np >= (1ll << 31) || g.nn >= (1ll << 31)) return 0; // 32-bit particle ids / keys
  if (g.nn != ncells || s.np > cap_np) {
    release();
    ncells = g.nn; cap_np = s.np;
    if (cudaMalloc(&cell_of, sizeof(int) * cap_np) || cudaMalloc(&rank, sizeof(int) * cap_np) || cudaMalloc(&order, sizeof(int) * cap_np) ||
        cudaMalloc(&start, sizeof(int) * (ncells + 1))) return 1;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, start, start, (int)(ncells + 1), st);
    if (cudaMalloc(&scan_tmp, scan_bytes)) return 1;
  }
  if (cudaMemsetAsync(start, 0, sizeof(int) * (ncells + 1), st)) return 1;
  const unsigned nb = (unsigned)((s.np + 255) / 256);
  k_cell_count<<<nb, 256, 0, st>>>(s, g, cell_of, rank, start);
  if (cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, start, start, (int)(ncells + 1), st)) return 1;
  k_cell_fill<<<nb, 256, 0, st>>>(s.np, cell_of, rank, start, order);
  *nlaunch = 4;
  if (cudaGetLastError() != cudaSuccess) return 1;
  valid = true; return 0;
}

// one cubic-spline (value, derivative) pair of node i for a particle at xp; same expressions as axis_weights
__device__ __forceinline__ void cubic_node(double xp, double lo, double h, double ih, int i, int n, double &w, double &dw) {
  if (i < 0 || i >= n) { w = 0; dw = 0; return; }
  const double xn = __dadd_rn(lo, __dmul_rn((double)i, h));
  const double r = __dmul_rn(__dsub_rn(xp, xn), ih);
  Basis<KML_SHAPE_CUBIC_SPLINE>::eval(r, node_type<KML_SHAPE_CUBIC_SPLINE>(i, n), ih, w, dw);
}

// FULL: mass + momentum + internal force (7 sums per node); !FULL: momentum only (MUSL re-projection)
template <bool FULL> struct CellAcc { static constexpr int Q = FULL ? 7 : 3; };

template <bool FULL, bool MASS>
__global__ void __launch_bounds__(128) k_p2g_cell(SolidDev s, GridDev g, const int *__restrict__ start, const int *__restrict__ order, int seglen, int nseg) {
  constexpr int Q = CellAcc<FULL>::Q;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long group = gid >> 3;
  const int lane8 = (int)(gid & 7), a = lane8 >> 1, b0 = (lane8 & 1) * 2;
  const long long ncol = (long long)g.n[0] * g.n[1];
  const long long col = group / nseg; const int seg = (int)(group % nseg);
  if (col >= ncol) return;
  const int i0 = (int)(col / g.n[1]), j0 = (int)(col % g.n[1]);
  const int kbeg = seg * seglen, kend = min(kbeg + seglen, g.n[2]);
  if (kbeg >= kend) return;
  const int ni = i0 + a;                       // this lane's node row
  const bool row_ok = ni < g.n[0];
  const long long cellbase = col * g.n[2];
  // quick exit: no particle in the whole segment
  if (start[cellbase + kend] == start[cellbase + kbeg]) return;

  double acc[2][4][Q];
#pragma unroll
  for (int b = 0; b < 2; b++)
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int q = 0; q < Q; q++) acc[b][c][q] = 0.0;

  // add plane `slot` (node plane kk) to the grid and clear it
  auto emit = [&](int slot, int kk) {
    if (row_ok && kk < g.n[2]) {
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int nj = j0 + b0 + b;
        if (nj < g.n[1]) {
          const long long node = ((long long)ni * g.n[1] + nj) * g.n[2] + kk;
          if (FULL) {
            if (MASS && acc[b][slot][0] != 0.0) atomicAdd(&g.mass[node], acc[b][slot][0]);
#pragma unroll
            for (int d = 0; d < 3; d++) {
              if (acc[b][slot][1 + d] != 0.0) atomicAdd(&g.v[d][node], acc[b][slot][1 + d]);
              if (acc[b][slot][4 + d] != 0.0) atomicAdd(&g.f[d][node], acc[b][slot][4 + d]);
            }
          } else {
#pragma unroll
            for (int d = 0; d < 3; d++) if (acc[b][slot][d] != 0.0) atomicAdd(&g.v[d][node], acc[b][slot][d]);
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < 2; b++)
#pragma unroll
      for (int q = 0; q < Q; q++) acc[b][slot][q] = 0.0;
  };

  // process cell kk with plane c living in register slot (c + R) & 3
#define KML_CELL_STEP(R)                                                                                         \
  {                                                                                                              \
    const int kk = k + R;                                                                                        \
    if (kk < kend) {                                                                                             \
      const int pbeg = start[cellbase + kk], pend = start[cellbase + kk + 1];                                    \
      for (int p = pbeg; p < pend; p++) {                                                                        \
        const int ip = order[p];                                                                                 \
        const double px = s.x[0][ip], py = s.x[1][ip], pz = s.x[2][ip];                                          \
        double wx, dwx, wy[2], dwy[2], wz[4], dwz[4];                                                            \
        cubic_node(px, g.lo[0], g.h, g.inv_cellsize, ni, g.n[0], wx, dwx);                                       \
        _Pragma("unroll") for (int b = 0; b < 2; b++) cubic_node(py, g.lo[1], g.h, g.inv_cellsize, j0 + b0 + b, g.n[1], wy[b], dwy[b]); \
        _Pragma("unroll") for (int c = 0; c < 4; c++) cubic_node(pz, g.lo[2], g.h, g.inv_cellsize, kk + c, g.n[2], wz[c], dwz[c]);     \
        const double m = s.mass[ip];                                                                             \
        const double v0 = s.v[0][ip], v1 = s.v[1][ip], v2 = s.v[2][ip];                                          \
        double A[9];                                                                                             \
        if (FULL) {                                                                                              \
          load_sym(s.sig, ip, A);                                                                                \
          const double vol = s.vol[ip];                                                                          \
          _Pragma("unroll") for (int e = 0; e < 9; e++) A[e] *= vol;                                             \
        }                                                                                                        \
        _Pragma("unroll") for (int b = 0; b < 2; b++) {                                                          \
          const double gxy = wx * wy[b];                                                                         \
          const double mm = gxy * m;                                                                             \
          const double M0 = mm * v0, M1 = mm * v1, M2 = mm * v2;                                                 \
          double P0 = 0, P1 = 0, P2 = 0, Q0 = 0, Q1 = 0, Q2 = 0;                                                 \
          if (FULL) {                                                                                            \
            const double gx = dwx * wy[b], gy = wx * dwy[b];                                                     \
            P0 = -(A[0] * gx + A[1] * gy); P1 = -(A[3] * gx + A[4] * gy); P2 = -(A[6] * gx + A[7] * gy);         \
            Q0 = -(A[2] * gxy); Q1 = -(A[5] * gxy); Q2 = -(A[8] * gxy);                                          \
          }                                                                                                      \
          _Pragma("unroll") for (int c = 0; c < 4; c++) {                                                        \
            double *ac = acc[b][(c + R) & 3];                                                                    \
            if (FULL) {                                                                                          \
              ac[0] += mm * wz[c];                                                                               \
              ac[1] += M0 * wz[c]; ac[2] += M1 * wz[c]; ac[3] += M2 * wz[c];                                     \
              ac[4] += P0 * wz[c] + Q0 * dwz[c]; ac[5] += P1 * wz[c] + Q1 * dwz[c]; ac[6] += P2 * wz[c] + Q2 * dwz[c]; \
            } else {                                                                                             \
              ac[0] += M0 * wz[c]; ac[1] += M1 * wz[c]; ac[2] += M2 * wz[c];                                     \
            }                                                                                                    \
          }                                                                                                      \
        }                                                                                                        \
      }                                                                                                          \
      emit(R, kk);                                                                                               \
    }                                                                                                            \
  }

  int k = kbeg;
  for (; k < kend; k += 4) { KML_CELL_STEP(0) KML_CELL_STEP(1) KML_CELL_STEP(2) KML_CELL_STEP(3) }
#undef KML_CELL_STEP
  // the three node planes above the last cell of the segment: plane c of cell (kend-1) sits in slot (c + r) & 3
  // with r = (kend - 1 - kbeg) & 3
  const int r = (kend - 1 - kbeg) & 3;
#pragma unroll
  for (int c = 1; c < 4; c++) {
    const int slot = (c + r) & 3;
    // slot is a runtime value here: select through a small switch so the accumulators stay in registers
    switch (slot) { case 0: emit(0, kend - 1 + c); break; case 1: emit(1, kend - 1 + c); break; case 2: emit(2, kend - 1 + c); break; default: emit(3, kend - 1 + c); break; }
  }
}

inline int cell_p2g_launch(const SolidDev &s, const GridDev &g, const CellLists &cl, int what, cudaStream_t st, int *nlaunch) {
  *nlaunch = 0;
  const bool full = (what & P2G_FORCE) != 0;
  // returns 0 = launched, -1 = combination not covered (caller uses the atomic kernel), 1 = CUDA error
  if (what & (P2G_MB | P2G_TEMP | P2G_HEAT)) return -1;
  if (full && !(what & P2G_MOM)) return -1;
  if (!full && (what & P2G_MASS)) return -1; // mass-only / mass+momentum passes (USF) use the atomic kernel
  if (!full && !(what & P2G_MOM)) return -1;
  const int seglen = 32;
  const int nseg = (g.n[2] + seglen - 1) / seglen;
  const long long ngroups = (long long)g.n[0] * g.n[1] * nseg;
  const unsigned nb = (unsigned)((ngroups * 8 + 127) / 128);
  if (full) {
    if (what & P2G_MASS) k_p2g_cell<true, true><<<nb, 128, 0, st>>>(s, g, cl.start, cl.order, seglen, nseg);
    else k_p2g_cell<true, false><<<nb, 128, 0, st>>>(s, g, cl.start, cl.order, seglen, nseg);
  } else k_p2g_cell<false, false><<<nb, 128, 0, st>>>(s, g, cl.start, cl.order, seglen, nseg);
  *nlaunch = 1;
  return cudaGetLastError() != cudaSuccess;
}

} // namespace kml
