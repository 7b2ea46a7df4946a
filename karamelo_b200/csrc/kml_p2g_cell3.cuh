// kml_p2g_cell3.cuh - cell-centric particle-to-grid, third generation (3-D cubic B-splines, ULMPM).
//
// Same decomposition as kml_p2g_cell.cuh (a lane group walks a column of cells along k and keeps a
// sliding 4-plane window of node sums in registers; only completed planes go to the grid as fp64 RED),
// rebuilt around the instruction budget of the FP64 pipe, which co-bounds this stage:
//   * every lane stages ONE particle per round (all 24 weights, m*v, vol*sigma) - no half-idle staging;
//   * interior columns evaluate the cubic B-spline pieces branch-free (node a of the stencil always lies in
//     the same interval of src/basis_functions.h:40-104); boundary columns keep the reference's branches;
//   * the raw data of the next round is loaded before the current round is accumulated (register prefetch);
//   * a ballot marks the particles that open a new cell plane, so the inner loop has no per-particle lookup;
//   * a lane may own NB = 2 node columns (8-lane groups): half the shared-memory traffic per FP64 FMA;
//   * no FP64 compares in the emit path (an integer dirty mask tracks non-empty planes).
// Arithmetic per (particle, node) is that of src/solid.cpp:317-335, :337-390, :482-522; the summation
// order differs (tolerance 1e-10, tests/test_parity_gpu.py).
#pragma once
#include "kml_p2g_cell.cuh"

namespace kml {

// Shared-memory loads through an explicit 32-bit shared-space address.  A generic pointer into __shared__ makes the compiler rebuild the
// shared window base (S2R SR_CgaCtaId + LEA) in front of the loads - inside the particle loop when registers are tight, where that S2R
// sits on the critical path of every iteration.
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds_d2(unsigned addr) {
  double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ double lds_d1(unsigned addr) {
  double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v;
}

// the four cubic B-spline pieces of an interior stencil (ntype 0 everywhere): node a sees r in [1-a, 2-a)
__device__ __forceinline__ void cubic_piece(int a, double r, double ih, double &w, double &dw) {
  if (a == 0) { w = ((-1.0 / 6.0 * r + 1) * r - 2) * r + 4.0 / 3.0; dw = ih * ((-0.5 * r + 2) * r - 2); }
  else if (a == 1) { w = (0.5 * r - 1) * r * r + 2.0 / 3.0; dw = ih * (3.0 / 2.0 * r - 2) * r; }
  else if (a == 2) { w = (-0.5 * r - 1) * r * r + 2.0 / 3.0; dw = ih * (-3.0 / 2.0 * r - 2) * r; }
  else { w = ((1.0 / 6.0 * r + 1) * r + 2) * r + 4.0 / 3.0; dw = ih * ((0.5 * r + 2) * r + 2); }
}

// Neighbour membership (kml_device.cuh weight_tiny): the branch-free pieces use FMAs; a tiny weight of one of the two outer nodes is
// recomputed with the reference's operation sequence, and no weight means no gradient.
__device__ __forceinline__ void cubic_membership(double xp, double lo, double h, double ih, int ig, bool upper, double &w, double &dw) { // ig: GLOBAL node index
  if (weight_tiny(w)) {
    const double xn = __dadd_rn(lo, __dmul_rn((double)ig, h));
    const double r = __dmul_rn(__dsub_rn(xp, xn), ih);
    w = upper ? horner3_unfused(1.0 / 6.0, 1.0, 2.0, 4.0 / 3.0, r) : horner3_unfused(-1.0 / 6.0, 1.0, -2.0, 4.0 / 3.0, r);
    if (w == 0.0) dw = 0.0;
  }
}

// (w, dw) of the 4 stencil nodes i0..i0+3 (LOCAL indices) of one axis; interior = all four nodes exist and have ntype 0
__device__ __forceinline__ void cubic_axis4(double xp, double lo, double h, double ih, int i0, int n, int goff, int gn, bool interior, double (&w)[4], double (&dw)[4]) {
  if (interior) {
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double xn = __dadd_rn(lo, __dmul_rn((double)(i0 + a + goff), h));
      const double r = __dmul_rn(__dsub_rn(xp, xn), ih);
      cubic_piece(a, r, ih, w[a], dw[a]);
    }
    cubic_membership(xp, lo, h, ih, i0 + goff, false, w[0], dw[0]);    // node 0 sees r in [1, 2), node 3 r in [-2, -1); the inner two carry at least 1/6
    cubic_membership(xp, lo, h, ih, i0 + 3 + goff, true, w[3], dw[3]);
  } else {
#pragma unroll
    for (int a = 0; a < 4; a++) cubic_node(xp, lo, h, ih, i0 + a, n, goff, gn, w[a], dw[a]); // Basis<>::eval: membership included
  }
}
__device__ __forceinline__ bool cubic_interior(int i0, int n, int goff, int gn) {
  const int ig = i0 + goff;
  return i0 >= 0 && i0 + 3 < n && ig >= 2 && ig + 3 <= gn - 3;
}

// Staged particle record (doubles).  FULL:  [0..7] x (w,dw)x4  [8..15] y (w,dw)x4  [16..19] z w  [20..23] z dw
//                                           [24..27] m, m*vx, m*vy, m*vz  [28..33] vol*sigma (xx,yy,zz,xy,xz,yz)  [34] cell plane k
//                                    !FULL: [0..3] x w  [4..7] y w  [8..11] z w  [12..14] m*v  [15] cell plane k
// The record stride is chosen so that the 16-byte stores of consecutive lanes fall into different banks (stride mod 128 B = 48 / 16):
// with a power-of-two stride every STS of the staging pass is a 16-way bank conflict (ncu: l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st
// 42 % of peak in the momentum pass with a 128-byte stride; the padding took that pass from 0.54 to 0.42 ms at 7 M particles).
template <bool FULL> struct Rec3 { static constexpr int N = FULL ? 38 : 18; static constexpr int K = FULL ? 34 : 15; };

template <bool FULL, bool MASS, int NB>
__global__ void __launch_bounds__(128, FULL ? (NB == 2 ? 2 : 3) : (NB == 4 ? 3 : 4))
k_p2g_cell3(SolidDev s, GridDev g, const int *__restrict__ start, const int *__restrict__ order, int seglen, int nseg) {
  constexpr int Q = FULL ? 7 : 3;
  constexpr int GL = 16 / NB;          // lanes per group = particles per staging round
  constexpr int GPB = 128 / GL;        // groups per block
  constexpr int REC = Rec3<FULL>::N;
  constexpr int GSTRIDE = GL * REC + (NB == 1 ? 8 : (NB == 2 ? 4 : 2)); // the groups of one warp start at different bank offsets (64 / 32 / 16 B apart)
  __shared__ __align__(16) double stage[GPB * GSTRIDE];

  const int lg = threadIdx.x % GL;
  int a = lg / (4 / NB), b0 = (lg % (4 / NB)) * NB;
  asm volatile("" : "+r"(a), "+r"(b0)); // lane constants stay in registers (no S2R + shifts inside the particle loop)
  const unsigned gmask = (GL == 32 ? 0xFFFFFFFFu : ((1u << GL) - 1u)) << ((threadIdx.x & 31) / GL * GL);
  double *rec0 = stage + (threadIdx.x / GL) * GSTRIDE;
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / GL;
  const long long ncol = (long long)g.n[0] * g.n[1];
  // consecutive groups = consecutive columns (j fastest) of the SAME segment: the groups of a warp carry equal work
  // and the planes a block emits overlap in L2
  const int seg = (int)(group / ncol); const long long col = group % ncol;
  if (seg >= nseg) return;
  const int i0 = (int)(col / g.n[1]), j0 = (int)(col % g.n[1]);
  const int kbeg = seg * seglen, kend = min(kbeg + seglen, g.n[2]);
  if (kbeg >= kend) return;
  const long long cellbase = col * g.n[2];
  const int pbeg = start[cellbase + kbeg], pend = start[cellbase + kend];
  if (pbeg == pend) return; // no particle in the whole segment (uniform per group)

  const int ni = i0 + a;
  const bool int_x = cubic_interior(i0, g.n[0], g.goff0, g.gn0), int_y = cubic_interior(j0, g.n[1], 0, g.n[1]);
  const double h = g.h, ih = g.inv_cellsize;

  double acc[NB][4][Q];
#pragma unroll
  for (int e = 0; e < NB; e++)
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int q = 0; q < Q; q++) acc[e][c][q] = 0.0;

  auto emit0 = [&](int kk) { // add plane kk (window slot 0) to the grid
    if (kk >= g.n[2] || ni >= g.n[0]) return;
#pragma unroll
    for (int e = 0; e < NB; e++) {
      const int nj = j0 + b0 + e;
      if (nj >= g.n[1]) continue;
      const long long node = ((long long)ni * g.n[1] + nj) * g.n[2] + kk;
      double4 *rec = &g.nv[node];
      if (FULL) {
        if (MASS) atomicAdd(&rec->w, acc[e][0][0]);
#pragma unroll
        for (int d = 0; d < 3; d++) { atomicAdd(comp_ptr(rec, d), acc[e][0][1 + d]); atomicAdd(&g.f[d][node], acc[e][0][4 + d]); }
      } else {
#pragma unroll
        for (int d = 0; d < 3; d++) atomicAdd(comp_ptr(rec, d), acc[e][0][d]);
      }
    }
  };
  auto slide = [&]() {
#pragma unroll
    for (int e = 0; e < NB; e++)
#pragma unroll
      for (int q = 0; q < Q; q++) { acc[e][0][q] = acc[e][1][q]; acc[e][1][q] = acc[e][2][q]; acc[e][2][q] = acc[e][3][q]; acc[e][3][q] = 0.0; }
  };

  // raw particle data of one staging round, one particle per lane
  double rx = 0, ry = 0, rz = 0, rm = 0, rv0 = 0, rv1 = 0, rv2 = 0, rvol = 0, rs[6] = {0, 0, 0, 0, 0, 0};
  auto load_raw = [&](int ip) { // ip: the particle's storage index, fetched from order[] one round earlier (no address stall here)
    rx = s.x[0][ip]; ry = s.x[1][ip]; rz = s.x[2][ip]; rm = s.mass[ip];
    rv0 = s.v[0][ip]; rv1 = s.v[1][ip]; rv2 = s.v[2][ip];
    if (FULL) {
      rvol = s.vol[ip];
#pragma unroll
      for (int e = 0; e < 6; e++) rs[e] = s.sig[e][ip];
    }
  };
  auto stage_raw = [&]() -> int { // weights + products of the lane's particle -> its record; returns the particle's cell plane
    double *r = rec0 + lg * REC;
    const int k0 = cell_axis(rz, g.lo[2], ih, g.n[2], 0);
    const bool int_z = cubic_interior(k0, g.n[2], 0, g.n[2]);
    double w[4], dw[4];
    cubic_axis4(rx, g.lo[0], h, ih, i0, g.n[0], g.goff0, g.gn0, int_x, w, dw);
    if (FULL) {
#pragma unroll
      for (int t = 0; t < 4; t++) *(double2 *)(r + 2 * t) = make_double2(w[t], dw[t]);
    } else { *(double2 *)(r + 0) = make_double2(w[0], w[1]); *(double2 *)(r + 2) = make_double2(w[2], w[3]); }
    cubic_axis4(ry, g.lo[1], h, ih, j0, g.n[1], 0, g.n[1], int_y, w, dw);
    if (FULL) {
#pragma unroll
      for (int t = 0; t < 4; t++) *(double2 *)(r + 8 + 2 * t) = make_double2(w[t], dw[t]);
    } else { *(double2 *)(r + 4) = make_double2(w[0], w[1]); *(double2 *)(r + 6) = make_double2(w[2], w[3]); }
    cubic_axis4(rz, g.lo[2], h, ih, k0, g.n[2], 0, g.n[2], int_z, w, dw);
    if (FULL) {
      *(double2 *)(r + 16) = make_double2(w[0], w[1]); *(double2 *)(r + 18) = make_double2(w[2], w[3]);
      *(double2 *)(r + 20) = make_double2(dw[0], dw[1]); *(double2 *)(r + 22) = make_double2(dw[2], dw[3]);
      *(double2 *)(r + 24) = make_double2(rm, rm * rv0); *(double2 *)(r + 26) = make_double2(rm * rv1, rm * rv2);
      *(double2 *)(r + 28) = make_double2(rvol * rs[0], rvol * rs[1]); *(double2 *)(r + 30) = make_double2(rvol * rs[2], rvol * rs[3]);
      *(double2 *)(r + 32) = make_double2(rvol * rs[4], rvol * rs[5]);
      *(double2 *)(r + 34) = make_double2(__longlong_as_double((long long)k0), 0.0);
    } else {
      *(double2 *)(r + 8) = make_double2(w[0], w[1]); *(double2 *)(r + 10) = make_double2(w[2], w[3]);
      *(double2 *)(r + 12) = make_double2(rm * rv0, rm * rv1);
      *(double2 *)(r + 14) = make_double2(rm * rv2, __longlong_as_double((long long)k0));
    }
    return k0;
  };

  // register copy of one staged record, as this lane needs it
  struct RecR { double2 X, Y[NB], Z01, Z23, D01, D23, MM, MV, A01, A23, A45; };
  const unsigned rec0s = smem_addr(rec0);              // the group's records as a shared-space address
  const unsigned offx = FULL ? 16u * a : 8u * a;       // byte offsets of this lane's x / y weights inside a record
  const unsigned offy = FULL ? 64u + 16u * b0 : 32u + 8u * b0;
  auto rec_load = [&](RecR &R, unsigned r) { // r: shared-space byte address of the record
    if (FULL) {
      R.X = lds_d2(r + offx);
#pragma unroll
      for (int e = 0; e < NB; e++) R.Y[e] = lds_d2(r + offy + 16u * e);
      R.Z01 = lds_d2(r + 128); R.Z23 = lds_d2(r + 144); R.D01 = lds_d2(r + 160); R.D23 = lds_d2(r + 176);
      R.MM = lds_d2(r + 192); R.MV = lds_d2(r + 208);
      R.A01 = lds_d2(r + 224); R.A23 = lds_d2(r + 240); R.A45 = lds_d2(r + 256);
    } else {
      R.X.x = lds_d1(r + offx);
#pragma unroll
      for (int e = 0; e < NB; e++) R.Y[e].x = lds_d1(r + offy + 8u * e);
      R.Z01 = lds_d2(r + 64); R.Z23 = lds_d2(r + 80);
      R.MM = lds_d2(r + 96); R.MV.x = lds_d1(r + 112);
    }
  };
  auto rec_accumulate = [&](const RecR &R) {
    const double wz[4] = {R.Z01.x, R.Z01.y, R.Z23.x, R.Z23.y};
    if (FULL) {
      const double dwz[4] = {R.D01.x, R.D01.y, R.D23.x, R.D23.y};
#pragma unroll
      for (int e = 0; e < NB; e++) {
        const double gxy = R.X.x * R.Y[e].x, gx = R.X.y * R.Y[e].x, gy = R.X.x * R.Y[e].y;
        const double mm = gxy * R.MM.x, M0 = gxy * R.MM.y, M1 = gxy * R.MV.x, M2 = gxy * R.MV.y;
        // A = (xx,yy,zz,xy,xz,yz): f_x = -(xx gx + xy gy) wz - xz gxy dwz, ...
        const double P0 = -(R.A01.x * gx + R.A23.y * gy), P1 = -(R.A23.y * gx + R.A01.y * gy), P2 = -(R.A45.x * gx + R.A45.y * gy);
        const double Q0 = -(R.A45.x * gxy), Q1 = -(R.A45.y * gxy), Q2 = -(R.A23.x * gxy);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          double *ac = acc[e][c];
          ac[0] += mm * wz[c];
          ac[1] += M0 * wz[c]; ac[2] += M1 * wz[c]; ac[3] += M2 * wz[c];
          ac[4] = fma(Q0, dwz[c], fma(P0, wz[c], ac[4])); // two FMAs per component (a sum of products would cost three FP64 instructions)
          ac[5] = fma(Q1, dwz[c], fma(P1, wz[c], ac[5])); ac[6] = fma(Q2, dwz[c], fma(P2, wz[c], ac[6]));
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < NB; e++) {
        const double gxy = R.X.x * R.Y[e].x;
        const double M0 = gxy * R.MM.x, M1 = gxy * R.MM.y, M2 = gxy * R.MV.x;
#pragma unroll
        for (int c = 0; c < 4; c++) { double *ac = acc[e][c]; ac[0] += M0 * wz[c]; ac[1] += M1 * wz[c]; ac[2] += M2 * wz[c]; }
      }
    }
  };

  const int gshift = (threadIdx.x & 31) / GL * GL;
  int kcur = kbeg, dirty = 0;
  int p = pbeg;
  // order[] is read TWO staging rounds ahead and the raw particle data ONE round ahead: ncu put 7 % of this kernel's stall samples on the
  // address computation that waited for order[p] right before its dependent loads
  int ip_nxt = p + lg < pend ? order[p + lg] : -1;
  if (ip_nxt >= 0) load_raw(ip_nxt);
  ip_nxt = p + GL + lg < pend ? order[p + GL + lg] : -1;
  while (p < pend) {
    const int n = min(GL, pend - p);
    const int k0 = lg < n ? stage_raw() : 0x7fffffff;
    // bit q of bm: particle q of this round lies in another cell plane than its predecessor (no per-particle lookup in the loop)
    int kprev = __shfl_up_sync(gmask, k0, 1, GL);
    if (lg == 0) kprev = kcur;
    const unsigned bm = __ballot_sync(gmask, lg < n && k0 != kprev) >> gshift;
    __syncwarp(gmask);
    const int pn = p + n;
    if (ip_nxt >= 0) load_raw(ip_nxt); // in flight while this round is accumulated
    ip_nxt = pn + GL + lg < pend ? order[pn + GL + lg] : -1;
    auto boundary = [&](int q) { // group-uniform: slide the window up to the particle's cell, emitting completed planes
      if ((bm >> q) & 1) {
        const int kq = (int)__double_as_longlong(lds_d1(rec0s + (unsigned)((q * REC + Rec3<FULL>::K) * 8))); // explicit shared-space address (see smem_addr)
        while (kcur < kq && dirty) { if (dirty & 1) emit0(kcur); slide(); dirty >>= 1; kcur++; }
        kcur = kq;
      }
      dirty = 0xF;
    };
    for (int q = 0; q < n; q++) {
      RecR cur; rec_load(cur, rec0s + (unsigned)(q * REC * 8));
      boundary(q); rec_accumulate(cur);
    }
    __syncwarp(gmask);
    p = pn;
  }
  // the node planes still in the window
  while (dirty) { if (dirty & 1) emit0(kcur); slide(); dirty >>= 1; kcur++; }
}

// returns 0 = launched, -1 = combination not covered (caller uses the atomic kernel), 1 = CUDA error
// segments of (almost) equal length, at most `target` cells long
inline void cell_segments(int n2, int target, int *seglen, int *nseg) {
  *nseg = (n2 + target - 1) / target; *seglen = (n2 + *nseg - 1) / *nseg;
}

// returns 0 = launched, -1 = combination not covered (caller uses the atomic kernel), 1 = CUDA error
template <bool FULL, bool MASS, int NB>
inline int cell_p2g3_launch_one(const SolidDev &s, const GridDev &g, const CellLists &cl, int seg_target, cudaStream_t st) {
  constexpr int GL = 16 / NB;
  int seglen, nseg; cell_segments(g.n[2], seg_target, &seglen, &nseg);
  const long long ngroups = (long long)g.n[0] * g.n[1] * nseg;
  const long long nb = (ngroups * GL + 127) / 128;
  if (nb >= (1ll << 31)) return -1;
  k_p2g_cell3<FULL, MASS, NB><<<(unsigned)nb, 128, 0, st>>>(s, g, cl.start, cl.order, seglen, nseg);
  return cudaGetLastError() != cudaSuccess;
}

// nb_full / nb_mom: node columns per lane for the full pass (1 | 2) and the momentum-only pass (1 | 2 | 4)
inline int cell_p2g3_launch(const SolidDev &s, const GridDev &g, const CellLists &cl, int what, int nb_full, int nb_mom, int seg_target, cudaStream_t st, int *nlaunch) {
  *nlaunch = 0;
  const bool full = (what & P2G_FORCE) != 0;
  if (what & (P2G_MB | P2G_TEMP | P2G_HEAT)) return -1;
  if (full && !(what & P2G_MOM)) return -1;
  if (!full && (what & P2G_MASS)) return -1; // mass-only / mass+momentum passes (USF) use the atomic kernel
  if (!full && !(what & P2G_MOM)) return -1;
  int rc;
#define KML_P2G3(F, M, N) cell_p2g3_launch_one<F, M, N>(s, g, cl, seg_target, st)
  if (full) {
    if (what & P2G_MASS) rc = nb_full == 2 ? KML_P2G3(true, true, 2) : KML_P2G3(true, true, 1);
    else rc = nb_full == 2 ? KML_P2G3(true, false, 2) : KML_P2G3(true, false, 1);
  } else {
    rc = nb_mom == 4 ? KML_P2G3(false, false, 4) : (nb_mom == 2 ? KML_P2G3(false, false, 2) : KML_P2G3(false, false, 1));
  }
#undef KML_P2G3
  if (rc == 0) *nlaunch = 1;
  return rc;
}

} // namespace kml
