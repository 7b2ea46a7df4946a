// kml_gather_cell.cuh - cell-run gather kernels for ULMPM with 3-D cubic B-splines: grid-to-particle
// (+advance) and velocity gradient + F + stress.
//
// A block owns one column segment of cells (i0, j0, [kbeg,kend)).  The 4 x 4 x (len+3) node records
// its particles can touch are staged once in shared memory with coalesced 32-byte loads along k; the
// particles of the segment (contiguous in the cell-sorted order) are then processed one per thread and
// every one of their 64 node reads is an LDS.128 pair instead of an L1/L2 round trip.  Arithmetic per
// particle is the same as the generic kernels (k_g2p / k_stress) - same (i,j,k) summation order.
#pragma once
#include "kml_p2g_cell3.cuh"

namespace kml {

template <bool STRESS>
__global__ void __launch_bounds__(128) k_gather_cell(SolidDev s, GridDev g, StepParams sp, StressParams tp, kml_material mat, const int *__restrict__ start,
                                                     const int *__restrict__ order, const int *__restrict__ cell_of, int seglen, int nseg) {
  extern __shared__ __align__(16) double4 tile[]; // STRESS: [16][TLEN] of the gathered field; G2P: nvu tile then nv tile
  const int TLEN = seglen + 3;
  const long long col = blockIdx.x / nseg; const int seg = (int)(blockIdx.x % nseg);
  const int i0 = (int)(col / g.n[1]), j0 = (int)(col % g.n[1]);
  const int kbeg = seg * seglen, kend = min(kbeg + seglen, g.n[2]);
  const long long cellbase = col * g.n[2];
  const int pbeg = start[cellbase + kbeg], pend = start[cellbase + kend];
  if (pbeg == pend) return; // block-uniform

  const double4 *__restrict__ src0 = STRESS ? (tp.doublemapping ? g.nv : g.nvu) : g.nvu;
  for (int e = threadIdx.x; e < 16 * TLEN; e += blockDim.x) {
    const int row = e / TLEN, t = e - row * TLEN;
    const int ni = i0 + (row >> 2), nj = j0 + (row & 3), nk = kbeg + t;
    double4 r0 = make_double4(0, 0, 0, 0), r1 = r0;
    if (ni < g.n[0] && nj < g.n[1] && nk < g.n[2]) {
      const long long node = ((long long)ni * g.n[1] + nj) * g.n[2] + nk;
      r0 = ldg4(&src0[node]);
      if (!STRESS) r1 = ldg4(&g.nv[node]);
    }
    tile[e] = r0;
    if (!STRESS) tile[16 * TLEN + e] = r1;
  }
  __syncthreads();

  double wave = 0, hr = 1.0;
  for (int p = pbeg + threadIdx.x; p < pend; p += blockDim.x) {
    const int ip = order[p];
    const int koff = (int)(cell_of[ip] - cellbase) - kbeg; // the particle's cell inside the segment
    const double px = s.x[0][ip], py = s.x[1][ip], pz = s.x[2][ip];
    double wx[4], dwx[4], wy[4], dwy[4], wz[4], dwz[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
      cubic_node(px, g.lo[0], g.h, g.inv_cellsize, i0 + t, g.n[0], g.goff0, g.gn0, wx[t], dwx[t]);
      cubic_node(py, g.lo[1], g.h, g.inv_cellsize, j0 + t, g.n[1], 0, g.n[1], wy[t], dwy[t]);
      cubic_node(pz, g.lo[2], g.h, g.inv_cellsize, kbeg + koff + t, g.n[2], 0, g.n[2], wz[t], dwz[t]);
    }
    if (STRESS) {
      double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      const double qv[3] = {0, 0, 0};
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const double gx = dwx[a] * wy[b], gy = wx[a] * dwy[b], gxy = wx[a] * wy[b];
          const double4 *row = tile + (a * 4 + b) * TLEN + koff;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const double4 rec = row[c];
            const double wfd0 = gx * wz[c], wfd1 = gy * wz[c], wfd2 = gxy * dwz[c];
            L[0] += rec.x * wfd0; L[1] += rec.x * wfd1; L[2] += rec.x * wfd2;
            L[3] += rec.y * wfd0; L[4] += rec.y * wfd1; L[5] += rec.y * wfd2;
            L[6] += rec.z * wfd0; L[7] += rec.z * wfd1; L[8] += rec.z * wfd2;
          }
        }
      PState ps; ps.load(s, mat, sp, ip);
      particle_stress<false>(s, g, sp, mat, ip, ps, L, qv, wave, hr);
    } else {
      double vu[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const double gxy = wx[a] * wy[b];
          const double4 *rowu = tile + (a * 4 + b) * TLEN + koff;
          const double4 *rowv = rowu + 16 * TLEN;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const double4 ru = rowu[c], rv = rowv[c];
            const double wf = gxy * wz[c];
            vu[0] += wf * ru.x; acc[0] += wf * (ru.x - rv.x);
            vu[1] += wf * ru.y; acc[1] += wf * (ru.y - rv.y);
            vu[2] += wf * ru.z; acc[2] += wf * (ru.z - rv.z);
          }
        }
      particle_advance<false>(s, sp, ip, vu, acc, 0.0);
    }
  }
  if (STRESS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wave = fmax(wave, __shfl_xor_sync(0xffffffffu, wave, o));
    if ((threadIdx.x & 31) == 0 && wave > 0) atomic_max_pos(tp.max_wave, wave);
  }
}

// returns 0 = launched, -1 = not covered, 1 = CUDA error
inline int cell_gather_launch(bool stress, const SolidDev &s, const GridDev &g, const StepParams &sp, const StressParams &tp, const kml_material &mat,
                              const CellLists &cl, cudaStream_t st) {
  if (sp.axisymmetric || sp.temp || !cl.valid) return -1;
  const int seglen = 32;
  const int nseg = (g.n[2] + seglen - 1) / seglen;
  const long long nblocks = (long long)g.n[0] * g.n[1] * nseg;
  if (nblocks >= (1ll << 31)) return -1;
  const size_t smem = sizeof(double4) * 16 * (seglen + 3) * (stress ? 1 : 2);
  if (stress) k_gather_cell<true><<<(unsigned)nblocks, 128, smem, st>>>(s, g, sp, tp, mat, cl.start, cl.order, cl.cell_of, seglen, nseg);
  else k_gather_cell<false><<<(unsigned)nblocks, 128, smem, st>>>(s, g, sp, tp, mat, cl.start, cl.order, cl.cell_of, seglen, nseg);
  return cudaGetLastError() != cudaSuccess;
}

} // namespace kml
