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
// (i = LOCAL node index of n, global index i + goff of gn)
__device__ __forceinline__ void cubic_node(double xp, double lo, double h, double ih, int i, int n, int goff, int gn, double &w, double &dw) {
  if (i < 0 || i >= n) { w = 0; dw = 0; return; }
  const double xn = __dadd_rn(lo, __dmul_rn((double)(i + goff), h));
  const double r = __dmul_rn(__dsub_rn(xp, xn), ih);
  Basis<KML_SHAPE_CUBIC_SPLINE>::eval(r, node_type<KML_SHAPE_CUBIC_SPLINE>(i + goff, gn), ih, w, dw);
}

// FULL: mass + momentum + internal force (7 sums per node); !FULL: momentum only (MUSL re-projection)
template <bool FULL> struct CellAcc { static constexpr int Q = FULL ? 7 : 3; };

// Staged particle record in shared memory (doubles), written once per particle by the staging lanes:
//   [0..7]   x axis: (w0,dw0, w1,dw1, w2,dw2, w3,dw3)      [8..15] y axis likewise
//   [16..23] z axis: (w0,w1,w2,w3, dw0,dw1,dw2,dw3)
//   [24..27] m, m*vx, m*vy, m*vz                             [28..33] vol*sigma (xx,yy,zz,xy,xz,yz)
constexpr int CELL_REC = 36;                       // padded to a multiple of 2 doubles (LDS.128 alignment)
constexpr int CELL_CHUNK = 8;                      // particles staged per round
constexpr int CELL_GROUP_STRIDE = CELL_REC * CELL_CHUNK + 2; // +16 B: the two half-warps of a warp use different banks
constexpr int CELL_GROUPS_PER_BLOCK = 8;           // 128 threads = 8 groups of 16 lanes

// A group of 16 lanes (a,b) owns one column segment of cells; lane (a,b) accumulates the 4 nodes (i0+a, j0+b, k..k+3).
template <bool FULL, bool MASS>
__global__ void __launch_bounds__(128, 4) k_p2g_cell(SolidDev s, GridDev g, const int *__restrict__ start, const int *__restrict__ order, int seglen, int nseg) {
  constexpr int Q = CellAcc<FULL>::Q;
  __shared__ __align__(16) double stage[CELL_GROUPS_PER_BLOCK * CELL_GROUP_STRIDE];
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long group = gid >> 4;
  const int lane16 = threadIdx.x & 15, a = lane16 >> 2, b = lane16 & 3;
  const unsigned halfmask = 0xFFFFu << (threadIdx.x & 16);
  double *rec0 = stage + (threadIdx.x >> 4) * CELL_GROUP_STRIDE;
  const long long ncol = (long long)g.n[0] * g.n[1];
  const long long col = group / nseg; const int seg = (int)(group % nseg);
  if (col >= ncol) return;
  const int i0 = (int)(col / g.n[1]), j0 = (int)(col % g.n[1]);
  const int kbeg = seg * seglen, kend = min(kbeg + seglen, g.n[2]);
  if (kbeg >= kend) return;
  const int ni = i0 + a, nj = j0 + b;
  const bool col_ok = ni < g.n[0] && nj < g.n[1];
  const long long cellbase = col * g.n[2];
  if (start[cellbase + kend] == start[cellbase + kbeg]) return; // no particle in the whole segment (uniform per group)

  double acc[4][Q];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int q = 0; q < Q; q++) acc[c][q] = 0.0;

  auto emit = [&](int slot, int kk) { // add node plane kk (register slot `slot`) to the grid and clear it
    if (col_ok && kk < g.n[2]) {
      const long long node = ((long long)ni * g.n[1] + nj) * g.n[2] + kk;
      double4 *rec = &g.nv[node];
      if (FULL) {
        if (MASS && acc[slot][0] != 0.0) atomicAdd(&rec->w, acc[slot][0]);
#pragma unroll
        for (int d = 0; d < 3; d++) {
          if (acc[slot][1 + d] != 0.0) atomicAdd(comp_ptr(rec, d), acc[slot][1 + d]);
          if (acc[slot][4 + d] != 0.0) atomicAdd(&g.f[d][node], acc[slot][4 + d]);
        }
      } else {
#pragma unroll
        for (int d = 0; d < 3; d++) if (acc[slot][d] != 0.0) atomicAdd(comp_ptr(rec, d), acc[slot][d]);
      }
    }
#pragma unroll
    for (int q = 0; q < Q; q++) acc[slot][q] = 0.0;
  };

  // stage up to CELL_CHUNK particles [pbeg, pbeg+n) of cell kk: lanes 0-7 evaluate the 12 (w,dw) pairs and m*v of
  // one particle each, lanes 8-15 its vol*sigma
  auto stage_chunk = [&](int pbeg, int n, int kk) {
    const int q = lane16 & 7;
    if (q < n) {
      const int ip = order[pbeg + q];
      double *r = rec0 + q * CELL_REC;
      if (lane16 < 8) {
        const double px = s.x[0][ip], py = s.x[1][ip], pz = s.x[2][ip];
        const double m = s.mass[ip];
        const double v0 = s.v[0][ip], v1 = s.v[1][ip], v2 = s.v[2][ip];
#pragma unroll
        for (int t = 0; t < 4; t++) {
          double w, dw;
          cubic_node(px, g.lo[0], g.h, g.inv_cellsize, i0 + t, g.n[0], g.goff0, g.gn0, w, dw); r[2 * t] = w; r[2 * t + 1] = dw;
          cubic_node(py, g.lo[1], g.h, g.inv_cellsize, j0 + t, g.n[1], 0, g.n[1], w, dw); r[8 + 2 * t] = w; r[8 + 2 * t + 1] = dw;
          cubic_node(pz, g.lo[2], g.h, g.inv_cellsize, kk + t, g.n[2], 0, g.n[2], w, dw); r[16 + t] = w; r[20 + t] = dw;
        }
        r[24] = m; r[25] = m * v0; r[26] = m * v1; r[27] = m * v2;
      } else if (FULL) {
        const double vol = s.vol[ip];
#pragma unroll
        for (int e = 0; e < 6; e++) r[28 + e] = vol * s.sig[e][ip];
      }
    }
  };

  // accumulate the n staged particles into the 4 node planes; plane c lives in register slot (c + R) & 3
#define KML_CELL_ACCUM(R, n)                                                                                      \
  for (int q = 0; q < (n); q++) {                                                                                 \
    const double *r = rec0 + q * CELL_REC;                                                                        \
    const double2 X = *(const double2 *)(r + 2 * a), Y = *(const double2 *)(r + 8 + 2 * b);                        \
    const double2 Z01 = *(const double2 *)(r + 16), Z23 = *(const double2 *)(r + 18);                              \
    const double2 D01 = *(const double2 *)(r + 20), D23 = *(const double2 *)(r + 22);                              \
    const double2 MM = *(const double2 *)(r + 24), MV = *(const double2 *)(r + 26);                                \
    const double wz[4] = {Z01.x, Z01.y, Z23.x, Z23.y}, dwz[4] = {D01.x, D01.y, D23.x, D23.y};                      \
    const double gxy = X.x * Y.x;                                                                                  \
    const double M0 = gxy * MM.y, M1 = gxy * MV.x, M2 = gxy * MV.y;                                                \
    if (FULL) {                                                                                                    \
      const double2 A01 = *(const double2 *)(r + 28), A23 = *(const double2 *)(r + 30), A45 = *(const double2 *)(r + 32); \
      const double gx = X.y * Y.x, gy = X.x * Y.y, mm = gxy * MM.x;                                                \
      /* A = (xx,yy,zz,xy,xz,yz): f_x = -(xx gx + xy gy) wz - xz gxy dwz, ... */                                   \
      const double P0 = -(A01.x * gx + A23.y * gy), P1 = -(A23.y * gx + A01.y * gy), P2 = -(A45.x * gx + A45.y * gy); \
      const double Q0 = -(A45.x * gxy), Q1 = -(A45.y * gxy), Q2 = -(A23.x * gxy);                                  \
      _Pragma("unroll") for (int c = 0; c < 4; c++) {                                                              \
        double *ac = acc[(c + R) & 3];                                                                             \
        ac[0] += mm * wz[c];                                                                                       \
        ac[1] += M0 * wz[c]; ac[2] += M1 * wz[c]; ac[3] += M2 * wz[c];                                             \
        ac[4] += P0 * wz[c] + Q0 * dwz[c]; ac[5] += P1 * wz[c] + Q1 * dwz[c]; ac[6] += P2 * wz[c] + Q2 * dwz[c];    \
      }                                                                                                            \
    } else {                                                                                                       \
      _Pragma("unroll") for (int c = 0; c < 4; c++) {                                                              \
        double *ac = acc[(c + R) & 3];                                                                             \
        ac[0] += M0 * wz[c]; ac[1] += M1 * wz[c]; ac[2] += M2 * wz[c];                                             \
      }                                                                                                            \
    }                                                                                                              \
  }

  // walk the cells of the segment; after a cell its lowest node plane is complete: add it to the grid and slide the
  // register window down by one plane (21 register moves per cell, one copy of the loop body in the I-cache)
  for (int kk = kbeg; kk < kend; kk++) {
    const int pbeg = start[cellbase + kk], pend = start[cellbase + kk + 1];
    for (int p = pbeg; p < pend; p += CELL_CHUNK) {
      const int n = min(CELL_CHUNK, pend - p);
      stage_chunk(p, n, kk);
      __syncwarp(halfmask);
      KML_CELL_ACCUM(0, n)
      __syncwarp(halfmask);
    }
    emit(0, kk);
#pragma unroll
    for (int q = 0; q < Q; q++) { acc[0][q] = acc[1][q]; acc[1][q] = acc[2][q]; acc[2][q] = acc[3][q]; acc[3][q] = 0.0; }
  }
#undef KML_CELL_ACCUM
  // the three node planes above the last cell of the segment
  emit(0, kend); emit(1, kend + 1); emit(2, kend + 2);
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
  const unsigned nb = (unsigned)((ngroups * 16 + 127) / 128);
  if (full) {
    if (what & P2G_MASS) k_p2g_cell<true, true><<<nb, 128, 0, st>>>(s, g, cl.start, cl.order, seglen, nseg);
    else k_p2g_cell<true, false><<<nb, 128, 0, st>>>(s, g, cl.start, cl.order, seglen, nseg);
  } else k_p2g_cell<false, false><<<nb, 128, 0, st>>>(s, g, cl.start, cl.order, seglen, nseg);
  *nlaunch = 1;
  return cudaGetLastError() != cudaSuccess;
}

} // namespace kml
