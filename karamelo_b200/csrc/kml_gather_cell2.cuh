// kml_gather_cell2.cuh - cell-run gather kernels (3-D cubic B-splines, ULMPM): grid-to-particle (+advance) and
// velocity gradient + F + stress.
//
// A block owns one column segment of cells (i0, j0, [kbeg,kend)).  The 4 x 4 x (len+3) node records its particles can
// touch are staged once in shared memory (the rows are contiguous along k in global memory); the particles of the
// segment (contiguous in the cell-sorted order) are then processed one per thread and every one of their 64 node
// reads is a shared-memory load instead of an L1/L2 round trip.  Built around the FP64 instruction budget
// (DESIGN.md section 3):
//   * interior columns evaluate the B-spline pieces branch-free (cubic_axis4);
//   * the velocity gradient is sum-factorised: contract the node values along k with (wz, dwz), then along
//     j with (wy, dwy), then along i with (wx, dwx) - 564 FP64 instructions per particle instead of 816;
//   * G2P reads packed 48-byte tile records {v_update, v_update - v}: three LDS.128 per node, and the
//     subtraction is done once per node when the tile is filled instead of once per (particle, node);
//   * the first particle's index and position are fetched before the tile is staged and the next particle's
//     while the current one is processed.
// Arithmetic: src/solid.cpp:576-635,786-796 (G2P + advance), :860-936 (gradient), :1155-1438 (F, stress).
#pragma once
#include "kml_p2g_cell3.cuh"
#include <cstdint>

namespace kml {

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) ------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "KML_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra KML_MBAR_DONE;\n"
      "bra KML_MBAR_WAIT;\n"
      "KML_MBAR_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy (1-D TMA); bytes a multiple of 16, both addresses 16-byte aligned; completes on the mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Packed gather records.  nvd is indexed on a PADDED grid (n1 + 3) x (n2 + NVD_PADK) with zero planes behind every axis, so a tile row
// (i, j, kbeg .. kbeg + seglen + 2) is always one in-bounds contiguous range, zero where no node exists.
constexpr int NVD_PADK = KML_NVD_PADK; // the longest segment (96 cells) + the stencil span + slack
static_assert(NVD_PADK >= 96 + 3, "tile rows of the longest segment must stay inside the padded grid");
__host__ __device__ __forceinline__ long long nvd_index(const GridDev &g, int i, int j, int k) { return ((long long)i * (g.n[1] + 3) + j) * (g.n[2] + NVD_PADK) + k; }
inline size_t nvd_doubles(const GridDev &g) { return (size_t)(g.n[0] + 3) * (g.n[1] + 3) * (g.n[2] + NVD_PADK) * 6; }


// ---- G2P + advance ------------------------------------------------------------------------------------------------
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_g2p_cell(SolidDev s, GridDev g, StepParams sp, const int *__restrict__ start, const int *__restrict__ order, int seglen, int nseg) {
  extern __shared__ __align__(16) double tile2[]; // [16][TLEN] x 6 doubles {v_update, v_update - v}
  const int TLEN = seglen + 3;
  const long long col = blockIdx.x / nseg; const int seg = (int)(blockIdx.x % nseg);
  const int i0 = (int)(col / g.n[1]), j0 = (int)(col % g.n[1]);
  const int kbeg = seg * seglen, kend = min(kbeg + seglen, g.n[2]);
  const long long cellbase = col * g.n[2];
  const int pbeg = start[cellbase + kbeg], pend = start[cellbase + kend];
  if (pbeg == pend) return; // block-uniform

  // first particle of this thread: index and position are in flight while the tile is staged
  int p = pbeg + threadIdx.x;
  int ip = p < pend ? order[p] : -1;
  int ipn = p + THREADS < pend ? order[p + THREADS] : -1; // order[] runs two particles ahead, positions one (no address stall in the loop)
  double px = 0, py = 0, pz = 0;
  if (ip >= 0) { px = s.x[0][ip]; py = s.x[1][ip]; pz = s.x[2][ip]; }

  for (int e = threadIdx.x; e < 16 * TLEN; e += THREADS) {
    const int row = e / TLEN, t = e - row * TLEN;
    const int ni = i0 + (row >> 2), nj = j0 + (row & 3), nk = kbeg + t;
    double4 r0 = make_double4(0, 0, 0, 0), r1 = r0;
    if (ni < g.n[0] && nj < g.n[1] && nk < g.n[2]) {
      const long long node = ((long long)ni * g.n[1] + nj) * g.n[2] + nk;
      r0 = ldg4(&g.nvu[node]); r1 = ldg4(&g.nv[node]);
    }
    double *d = tile2 + (size_t)e * 6;
    *(double2 *)d = make_double2(r0.x, r0.y); *(double2 *)(d + 2) = make_double2(r0.z, r0.x - r1.x);
    *(double2 *)(d + 4) = make_double2(r0.y - r1.y, r0.z - r1.z);
  }
  __syncthreads();

  const bool int_x = cubic_interior(i0, g.n[0], g.goff0, g.gn0), int_y = cubic_interior(j0, g.n[1], 0, g.n[1]);
  const double h = g.h, ih = g.inv_cellsize;
  unsigned tile_s = smem_addr(tile2);
  asm volatile("" : "+r"(tile_s)); // opaque: otherwise the window base is rematerialised (S2R + LEA) at the top of every particle
  while (ip >= 0) {
    // next particle of this thread
    const int pn = p + THREADS;
    const int ipnn = pn + THREADS < pend ? order[pn + THREADS] : -1;
    double nx = 0, ny = 0, nz = 0;
    if (ipn >= 0) { nx = s.x[0][ipn]; ny = s.x[1][ipn]; nz = s.x[2][ipn]; }
    double vold[3]; // in flight during the gather
    vold[0] = s.v[0][ip]; vold[1] = s.v[1][ip]; vold[2] = s.v[2][ip];

    const int k0 = cell_axis(pz, g.lo[2], ih, g.n[2], 0);
    const int koff = k0 - kbeg; // the particle's cell inside the segment
    double wx[4], wy[4], wz[4], dw_[4];
    cubic_axis4(px, g.lo[0], h, ih, i0, g.n[0], g.goff0, g.gn0, int_x, wx, dw_);
    cubic_axis4(py, g.lo[1], h, ih, j0, g.n[1], 0, g.n[1], int_y, wy, dw_);
    cubic_axis4(pz, g.lo[2], h, ih, k0, g.n[2], 0, g.n[2], cubic_interior(k0, g.n[2], 0, g.n[2]), wz, dw_);
    double vu[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const double gxy = wx[a] * wy[b];
        const unsigned row = tile_s + (unsigned)(((a * 4 + b) * TLEN + koff) * 48); // explicit shared-space address (see smem_addr)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const double2 r01 = lds_d2(row + 48 * c), r23 = lds_d2(row + 48 * c + 16), r45 = lds_d2(row + 48 * c + 32);
          const double wf = gxy * wz[c];
          vu[0] = fma(wf, r01.x, vu[0]); vu[1] = fma(wf, r01.y, vu[1]); vu[2] = fma(wf, r23.x, vu[2]);
          acc[0] = fma(wf, r23.y, acc[0]); acc[1] = fma(wf, r45.x, acc[1]); acc[2] = fma(wf, r45.y, acc[2]);
        }
      }
    particle_advance<false, false>(s, sp, ip, vu, acc, 0.0, vold); // kept accelerations (kml_keep_particle_acceleration) go through k_g2p
    p = pn; ip = ipn; ipn = ipnn; px = nx; py = ny; pz = nz;
  }
}

// ---- velocity gradient + F + stress: asynchronous staging ------------------------------------------------------
// Same tile / thread-per-particle scheme, but nothing on the memory side goes through registers: the node tile is
// filled with 16-byte cp.async copies, and every thread streams the 29-31 state doubles of its NEXT particle into a
// private shared-memory slot (8-byte cp.async) while it runs the constitutive update of the current one.  The register file only holds
// the weights, the gather accumulators and the constitutive update, so the kernel fits 168 registers (3 x 128 or
// 6 x 64 threads per SM); prefetching the state into registers instead costs 60 more and halves the occupancy.
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_stress_cell(SolidDev s, GridDev g, StepParams sp, StressParams tp, kml_material mat, const int *__restrict__ start, const int *__restrict__ order,
               int seglen, int nseg) {
  extern __shared__ __align__(16) double smem3[]; // [16][TLEN] double4 node tile, then [PSTATE_SLOTS][THREADS] particle state
  const int TLEN = seglen + 3;
  double *tile = smem3;
  double *state = smem3 + (size_t)16 * TLEN * 4;
  const int tid = threadIdx.x;
  const long long col = blockIdx.x / nseg; const int seg = (int)(blockIdx.x % nseg);
  const int i0 = (int)(col / g.n[1]), j0 = (int)(col % g.n[1]);
  const int kbeg = seg * seglen, kend = min(kbeg + seglen, g.n[2]);
  const long long cellbase = col * g.n[2];
  const int pbeg = start[cellbase + kbeg], pend = start[cellbase + kend];
  if (pbeg == pend) return; // block-uniform

  int p = pbeg + tid;
  int ip = p < pend ? order[p] : -1;
  int ipn = p + THREADS < pend ? order[p + THREADS] : -1;
  // node tile: 16 rows, contiguous along k in global memory
  const double4 *__restrict__ src0 = tp.doublemapping ? g.nv : g.nvu;
  for (int e = tid; e < 16 * TLEN; e += THREADS) {
    const int row = e / TLEN, t = e - row * TLEN;
    const int ni = i0 + (row >> 2), nj = j0 + (row & 3), nk = kbeg + t;
    double *d = tile + (size_t)e * 4;
    if (ni < g.n[0] && nj < g.n[1] && nk < g.n[2]) {
      const double4 *src = &src0[((long long)ni * g.n[1] + nj) * g.n[2] + nk];
      cp_async16(d, src); cp_async16(d + 2, (const double *)src + 2);
    } else { *(double2 *)d = make_double2(0.0, 0.0); *(double2 *)(d + 2) = make_double2(0.0, 0.0); }
  }
  double px = 0, py = 0, pz = 0;
  if (ip >= 0) { px = s.x[0][ip]; py = s.x[1][ip]; pz = s.x[2][ip]; pstate_issue_async(state + tid, THREADS, s, mat, sp, ip); }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

  const bool int_x = cubic_interior(i0, g.n[0], g.goff0, g.gn0), int_y = cubic_interior(j0, g.n[1], 0, g.n[1]);
  const double h = g.h, ih = g.inv_cellsize;
  double wave = 0, hr = 1.0;
  while (ip >= 0) {
    // next particle of this thread: position to registers now, state to shared memory once the slot is read
    const int pn = p + THREADS;
    double nx = 0, ny = 0, nz = 0;
    if (ipn >= 0) { nx = s.x[0][ipn]; ny = s.x[1][ipn]; nz = s.x[2][ipn]; }
    const int ipnn = pn + THREADS < pend ? order[pn + THREADS] : -1;

    const int k0 = cell_axis(pz, g.lo[2], ih, g.n[2], 0);
    const int koff = k0 - kbeg;
    const bool int_z = cubic_interior(k0, g.n[2], 0, g.n[2]);
    double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    {
      double wx[4], dwx[4], wy[4], dwy[4], wz[4], dwz[4];
      cubic_axis4(px, g.lo[0], h, ih, i0, g.n[0], g.goff0, g.gn0, int_x, wx, dwx);
      cubic_axis4(py, g.lo[1], h, ih, j0, g.n[1], 0, g.n[1], int_y, wy, dwy);
      cubic_axis4(pz, g.lo[2], h, ih, k0, g.n[2], 0, g.n[2], int_z, wz, dwz);
#pragma unroll 1
      for (int a = 0; a < 4; a++) { // rolled: the hot loop body stays small (instruction cache)
        const double wxa = a == 0 ? wx[0] : (a == 1 ? wx[1] : (a == 2 ? wx[2] : wx[3]));
        const double dwxa = a == 0 ? dwx[0] : (a == 1 ? dwx[1] : (a == 2 ? dwx[2] : dwx[3]));
        double A1[3] = {0, 0, 0}, A2[3] = {0, 0, 0}, B1[3] = {0, 0, 0};
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const double *row = tile + ((size_t)(a * 4 + b) * TLEN + koff) * 4;
          double A[3] = {0, 0, 0}, B[3] = {0, 0, 0};
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const double2 vxy = *(const double2 *)(row + 4 * c); const double vz = row[4 * c + 2];
            A[0] = fma(wz[c], vxy.x, A[0]); A[1] = fma(wz[c], vxy.y, A[1]); A[2] = fma(wz[c], vz, A[2]);
            B[0] = fma(dwz[c], vxy.x, B[0]); B[1] = fma(dwz[c], vxy.y, B[1]); B[2] = fma(dwz[c], vz, B[2]);
          }
#pragma unroll
          for (int d = 0; d < 3; d++) { A1[d] = fma(wy[b], A[d], A1[d]); A2[d] = fma(dwy[b], A[d], A2[d]); B1[d] = fma(wy[b], B[d], B1[d]); }
        }
#pragma unroll
        for (int d = 0; d < 3; d++) { L[3 * d] = fma(dwxa, A1[d], L[3 * d]); L[3 * d + 1] = fma(wxa, A2[d], L[3 * d + 1]); L[3 * d + 2] = fma(wxa, B1[d], L[3 * d + 2]); }
      }
    }
    PState ps;
    pstate_from_smem(ps, state + tid, THREADS, mat, sp);
    if (ipn >= 0) pstate_issue_async(state + tid, THREADS, s, mat, sp, ipn); // the slot is free again: stream the next particle's state
    cp_async_commit();
    const double qv[3] = {0, 0, 0};
    double wave_p = 0; // particle_stress ASSIGNS the particle's wave speed: keep the maximum over this thread's particles
    particle_stress<false>(s, g, sp, mat, ip, ps, L, qv, wave_p, hr);
    wave = fmax(wave, wave_p);
    cp_async_wait_all(); // the next particle's state has had the constitutive update to arrive
    p = pn; ip = ipn; ipn = ipnn; px = nx; py = ny; pz = nz;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wave = fmax(wave, __shfl_xor_sync(0xffffffffu, wave, o));
  if ((threadIdx.x & 31) == 0 && wave > 0) atomic_max_pos(tp.max_wave, wave);
}

// measurement knobs (environment, see kml.cu): cells per segment and threads per block
struct GatherTune { int seg_target = 64, seg_g2p = 32, seg_stress = 24, threads = 64, g2p_threads = 64; }; // seg_target: P2G / re-projection

// returns 0 = launched, -1 = not covered, 1 = CUDA error
inline int cell_gather_launch(bool stress, const SolidDev &s, const GridDev &g, const StepParams &sp, const StressParams &tp, const kml_material &mat,
                              const CellLists &cl, cudaStream_t st, const GatherTune &tune) {
  if (sp.axisymmetric || sp.temp || !cl.valid) return -1;
  // G2P: segments of (almost) equal length.  Stress: segments of EXACTLY seg_stress cells (the last one shorter) - measured at 100 M
  // particles: 24 cells 10.6 ms, 15 cells 11.0, 20 cells 12.0, 30 cells (the equalised split of 210 planes) 12.4, 12 cells 13.4; a short
  // tile leaves more of the SM's 256 KB to L1 for the 55 streamed state arrays, and 24 x 8 particles are 3 per thread on the usual lattice.
  int seglen, nseg;
  if (stress) { seglen = tune.seg_stress; nseg = (g.n[2] + seglen - 1) / seglen; }
  else cell_segments(g.n[2], tune.seg_g2p, &seglen, &nseg);
  const long long nblocks = (long long)g.n[0] * g.n[1] * nseg;
  if (nblocks >= (1ll << 31)) return -1;
  const size_t tile = sizeof(double) * 16 * (size_t)(seglen + 3) * (stress ? 4 : 6);
  const size_t smem = tile + (stress ? sizeof(double) * PSTATE_SLOTS * (size_t)tune.threads : 0);
#define KML_GATHER_LAUNCH(KERN, THREADS, ...)                                                                                       \
  do {                                                                                                                             \
    auto kern = KERN;                                                                                                              \
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
    kern<<<(unsigned)nblocks, THREADS, smem, st>>>(__VA_ARGS__);                                                                   \
  } while (0)
  if (stress) {
    if (tune.threads == 64) KML_GATHER_LAUNCH((k_stress_cell<64, 6>), 64, s, g, sp, tp, mat, cl.start, cl.order, seglen, nseg);
    else KML_GATHER_LAUNCH((k_stress_cell<128, 3>), 128, s, g, sp, tp, mat, cl.start, cl.order, seglen, nseg);
  } else {
    if (tune.g2p_threads == 64) KML_GATHER_LAUNCH((k_g2p_cell<64, 8>), 64, s, g, sp, cl.start, cl.order, seglen, nseg);
    else KML_GATHER_LAUNCH((k_g2p_cell<128, 4>), 128, s, g, sp, cl.start, cl.order, seglen, nseg);
  }
#undef KML_GATHER_LAUNCH
  return cudaGetLastError() != cudaSuccess;
}

} // namespace kml
