// kml_p2g_cell4.cuh - cell-centric particle-to-grid, fourth generation (3-D cubic B-splines, ULMPM).
//
// Same decomposition and arithmetic as k_p2g_cell3 (kml_p2g_cell3.cuh: 16-lane groups walk a column of cells along k with a sliding
// 4-plane window of node sums in registers, completed planes leave as fp64 RED).  What changed is where the memory side lives.  ncu on
// k_p2g_cell3 at 100 M particles (profiles/r2b_*): 164 registers -> 3 blocks of 128 threads per SM, 17.8 % warps active, FP64 pipe 51 %,
// issue slots 47 %; per warp and particle ~720 cycles of which the FP64 block is 313 and the rest is exposed latency (record loads,
// window slide, staging) that only other warps can cover.  28 of the 164 registers held the NEXT round's raw particle data across the
// whole accumulate loop.  Here that prefetch goes through shared memory instead:
//   * every lane streams the 14 (7 for the momentum pass) raw doubles of its next particle into a private shared-memory slot with
//     8-byte cp.async (SASS LDGSTS) while the current round is accumulated - no register is live across the loop for it;
//   * the staging pass reads the slot one axis at a time, so its temporaries stay small;
//   * node addresses of the emit path are a lane constant plus the plane index.
// The kernel fits 128 registers: 4 blocks per SM (4 warps per scheduler instead of 3).
#pragma once
#include "kml_p2g_cell3.cuh"

namespace kml {

template <bool FULL, bool MASS, int NB, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_p2g_cell4(SolidDev s, GridDev g, const int *__restrict__ start, const int *__restrict__ order, int seglen, int nseg) {
  constexpr int Q = FULL ? 7 : 3;
  constexpr int GL = 16 / NB;          // lanes per group = particles per staging round
  constexpr int GPB = 128 / GL;        // groups per block
  constexpr int REC = Rec3<FULL>::N;
  constexpr int GSTRIDE = GL * REC + (NB == 1 ? 8 : (NB == 2 ? 4 : 2));
  constexpr int NRAW = FULL ? 14 : 7;  // x y z m vx vy vz [vol sig0..5]
  extern __shared__ __align__(16) double smem4[];
  double *stage = smem4;                       // [GPB][GSTRIDE] staged records
  double *raw = smem4 + GPB * GSTRIDE;         // [NRAW][128] raw data of every lane's next particle

  const int lg = threadIdx.x % GL;
  int a = lg / (4 / NB), b0 = (lg % (4 / NB)) * NB;
  asm volatile("" : "+r"(a), "+r"(b0));
  const unsigned gmask = (GL == 32 ? 0xFFFFFFFFu : ((1u << GL) - 1u)) << ((threadIdx.x & 31) / GL * GL);
  double *rec0 = stage + (threadIdx.x / GL) * GSTRIDE;
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / GL;
  const long long ncol = (long long)g.n[0] * g.n[1];
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
  // emit path: node index of plane 0 of this lane's column(s); negative = the column does not exist on this grid
  long long node0[NB];
#pragma unroll
  for (int e = 0; e < NB; e++) node0[e] = (ni < g.n[0] && j0 + b0 + e < g.n[1]) ? ((long long)ni * g.n[1] + (j0 + b0 + e)) * g.n[2] : -1;

  double acc[NB][4][Q];
#pragma unroll
  for (int e = 0; e < NB; e++)
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int q = 0; q < Q; q++) acc[e][c][q] = 0.0;

  auto emit0 = [&](int kk) { // add plane kk (window slot 0) to the grid
    if (kk >= g.n[2]) return;
#pragma unroll
    for (int e = 0; e < NB; e++) {
      if (node0[e] < 0) continue;
      const long long node = node0[e] + kk;
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

  double *rawl = raw + threadIdx.x; // this lane's slot: component f at rawl[f * 128]
  auto issue_raw = [&](int ip) {    // asynchronous: global -> this lane's slot
    cp_async8(rawl + 0 * 128, s.x[0] + ip); cp_async8(rawl + 1 * 128, s.x[1] + ip); cp_async8(rawl + 2 * 128, s.x[2] + ip);
    cp_async8(rawl + 3 * 128, s.mass + ip);
    cp_async8(rawl + 4 * 128, s.v[0] + ip); cp_async8(rawl + 5 * 128, s.v[1] + ip); cp_async8(rawl + 6 * 128, s.v[2] + ip);
    if (FULL) {
      cp_async8(rawl + 7 * 128, s.vol + ip);
#pragma unroll
      for (int e = 0; e < 6; e++) cp_async8(rawl + (8 + e) * 128, s.sig[e] + ip);
    }
  };
  auto stage_raw = [&]() -> int { // weights + products of the lane's particle (read from its slot) -> its record; returns the particle's cell plane
    double *r = rec0 + lg * REC;
    double w[4], dw[4];
    cubic_axis4(rawl[0 * 128], g.lo[0], h, ih, i0, g.n[0], g.goff0, g.gn0, int_x, w, dw);
    if (FULL) {
#pragma unroll
      for (int t = 0; t < 4; t++) *(double2 *)(r + 2 * t) = make_double2(w[t], dw[t]);
    } else { *(double2 *)(r + 0) = make_double2(w[0], w[1]); *(double2 *)(r + 2) = make_double2(w[2], w[3]); }
    cubic_axis4(rawl[1 * 128], g.lo[1], h, ih, j0, g.n[1], 0, g.n[1], int_y, w, dw);
    if (FULL) {
#pragma unroll
      for (int t = 0; t < 4; t++) *(double2 *)(r + 8 + 2 * t) = make_double2(w[t], dw[t]);
    } else { *(double2 *)(r + 4) = make_double2(w[0], w[1]); *(double2 *)(r + 6) = make_double2(w[2], w[3]); }
    const double rz = rawl[2 * 128];
    const int k0 = cell_axis(rz, g.lo[2], ih, g.n[2], 0);
    const bool int_z = cubic_interior(k0, g.n[2], 0, g.n[2]);
    cubic_axis4(rz, g.lo[2], h, ih, k0, g.n[2], 0, g.n[2], int_z, w, dw);
    const double rm = rawl[3 * 128], rv0 = rawl[4 * 128], rv1 = rawl[5 * 128], rv2 = rawl[6 * 128];
    if (FULL) {
      *(double2 *)(r + 16) = make_double2(w[0], w[1]); *(double2 *)(r + 18) = make_double2(w[2], w[3]);
      *(double2 *)(r + 20) = make_double2(dw[0], dw[1]); *(double2 *)(r + 22) = make_double2(dw[2], dw[3]);
      *(double2 *)(r + 24) = make_double2(rm, rm * rv0); *(double2 *)(r + 26) = make_double2(rm * rv1, rm * rv2);
      const double rvol = rawl[7 * 128];
      *(double2 *)(r + 28) = make_double2(rvol * rawl[8 * 128], rvol * rawl[9 * 128]);
      *(double2 *)(r + 30) = make_double2(rvol * rawl[10 * 128], rvol * rawl[11 * 128]);
      *(double2 *)(r + 32) = make_double2(rvol * rawl[12 * 128], rvol * rawl[13 * 128]);
      *(double2 *)(r + 34) = make_double2(__longlong_as_double((long long)k0), 0.0);
    } else {
      *(double2 *)(r + 8) = make_double2(w[0], w[1]); *(double2 *)(r + 10) = make_double2(w[2], w[3]);
      *(double2 *)(r + 12) = make_double2(rm * rv0, rm * rv1);
      *(double2 *)(r + 14) = make_double2(rm * rv2, __longlong_as_double((long long)k0));
    }
    return k0;
  };

  struct RecR { double2 X, Y[NB], Z01, Z23, D01, D23, MM, MV, A01, A23, A45; };
  const unsigned rec0s = smem_addr(rec0);
  const unsigned offx = FULL ? 16u * a : 8u * a;
  const unsigned offy = FULL ? 64u + 16u * b0 : 32u + 8u * b0;
  auto rec_load = [&](RecR &R, unsigned r) {
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
        const double P0 = -(R.A01.x * gx + R.A23.y * gy), P1 = -(R.A23.y * gx + R.A01.y * gy), P2 = -(R.A45.x * gx + R.A45.y * gy);
        const double Q0 = -(R.A45.x * gxy), Q1 = -(R.A45.y * gxy), Q2 = -(R.A23.x * gxy);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          double *ac = acc[e][c];
          ac[0] += mm * wz[c];
          ac[1] += M0 * wz[c]; ac[2] += M1 * wz[c]; ac[3] += M2 * wz[c];
          ac[4] = fma(Q0, dwz[c], fma(P0, wz[c], ac[4]));
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
  // order[] runs two staging rounds ahead, the raw data (asynchronous, into the slot) one round ahead
  int ip_nxt = p + lg < pend ? order[p + lg] : -1;
  if (ip_nxt >= 0) issue_raw(ip_nxt);
  cp_async_commit();
  ip_nxt = p + GL + lg < pend ? order[p + GL + lg] : -1;
  while (p < pend) {
    const int n = min(GL, pend - p);
    cp_async_wait_all(); // this lane's slot holds its particle of this round (only this lane reads it: no barrier)
    const int k0 = lg < n ? stage_raw() : 0x7fffffff;
    int kprev = __shfl_up_sync(gmask, k0, 1, GL);
    if (lg == 0) kprev = kcur;
    const unsigned bm = __ballot_sync(gmask, lg < n && k0 != kprev) >> gshift;
    __syncwarp(gmask);
    const int pn = p + n;
    if (ip_nxt >= 0) issue_raw(ip_nxt); // the slot has been read: stream the next round's particle while this round is accumulated
    cp_async_commit();
    ip_nxt = pn + GL + lg < pend ? order[pn + GL + lg] : -1;
    auto boundary = [&](int q) {
      if ((bm >> q) & 1) {
        const int kq = (int)__double_as_longlong(rec0[q * REC + Rec3<FULL>::K]);
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
  while (dirty) { if (dirty & 1) emit0(kcur); slide(); dirty >>= 1; kcur++; }
}

// returns 0 = launched, -1 = combination not covered (caller falls back), 1 = CUDA error
template <bool FULL, bool MASS, int NB, int MINB>
inline int cell_p2g4_launch_one(const SolidDev &s, const GridDev &g, const CellLists &cl, int seg_target, cudaStream_t st) {
  constexpr int GL = 16 / NB;
  constexpr int GSTRIDE = GL * Rec3<FULL>::N + (NB == 1 ? 8 : (NB == 2 ? 4 : 2));
  constexpr size_t smem = sizeof(double) * ((128 / GL) * GSTRIDE + (FULL ? 14 : 7) * 128);
  int seglen, nseg; cell_segments(g.n[2], seg_target, &seglen, &nseg);
  const long long ngroups = (long long)g.n[0] * g.n[1] * nseg;
  const long long nb = (ngroups * GL + 127) / 128;
  if (nb >= (1ll << 31)) return -1;
  auto kern = k_p2g_cell4<FULL, MASS, NB, MINB>;
  static bool attr_done = false; // per instantiation
  if (!attr_done) { if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; attr_done = true; }
  kern<<<(unsigned)nb, 128, smem, st>>>(s, g, cl.start, cl.order, seglen, nseg);
  return cudaGetLastError() != cudaSuccess;
}

// the full pass (mass + momentum + force) with one node column per lane at 4 blocks per SM, the momentum pass with two columns per lane
inline int cell_p2g4_launch(const SolidDev &s, const GridDev &g, const CellLists &cl, int what, int nb_mom, int minb_full, int seg_target, cudaStream_t st, int *nlaunch) {
  *nlaunch = 0;
  const bool full = (what & P2G_FORCE) != 0;
  if (what & (P2G_MB | P2G_TEMP | P2G_HEAT)) return -1;
  if (full && !(what & P2G_MOM)) return -1;
  if (!full && (what & P2G_MASS)) return -1;
  if (!full && !(what & P2G_MOM)) return -1;
  int rc;
  if (full) {
    if (what & P2G_MASS) rc = minb_full == 3 ? cell_p2g4_launch_one<true, true, 1, 3>(s, g, cl, seg_target, st) : cell_p2g4_launch_one<true, true, 1, 4>(s, g, cl, seg_target, st);
    else rc = minb_full == 3 ? cell_p2g4_launch_one<true, false, 1, 3>(s, g, cl, seg_target, st) : cell_p2g4_launch_one<true, false, 1, 4>(s, g, cl, seg_target, st);
  } else {
    rc = nb_mom == 1 ? cell_p2g4_launch_one<false, false, 1, 5>(s, g, cl, seg_target, st) : cell_p2g4_launch_one<false, false, 2, 4>(s, g, cl, seg_target, st);
  }
  if (rc == 0) *nlaunch = 1;
  return rc;
}

} // namespace kml
