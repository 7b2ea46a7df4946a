// kml_gather_cell3.cuh - grid-to-particle (+advance) as a persistent, TMA-fed kernel (3-D cubic B-splines, ULMPM).
//
// k_g2p_cell (kml_gather_cell2.cuh) launches one block per column segment; ncu's per-instruction samples put 20 % of its
// stall time in the block prologue (start[] -> order[] -> x, node records through registers into the tile) and 35 % on
// `long_scoreboard` overall (profiles/).  Here:
//   * the node values the gather reads are PACKED once per step by the grid kernel: nvd[node] = {v_update, v_update - v}
//     (48 B, in an array padded with zero planes so that every tile row is in bounds and contiguous);
//   * a tile = 16 rows x (seglen + 3) records; one elected thread fetches the 16 rows with cp.async.bulk (TMA, SASS
//     UBLKCP) onto an mbarrier - no registers, no per-thread address arithmetic, no zero-fill branches;
//   * blocks are persistent (grid = SMs x resident blocks) and walk the column segments with stride gridDim.x; the tile
//     of the NEXT non-empty segment is requested before the current one is gathered (two tiles in flight, full / empty
//     mbarrier pairs), so no warp waits for a tile fill after the first one;
//   * warps of a block run decoupled: a warp that finishes its particles of segment m goes on to m + 1 (only the
//     producer waits for the slowest warp, on the `empty` barrier of the buffer it is about to refill).
// Arithmetic per (particle, node): src/solid.cpp:576-635, :786-796 - identical to k_g2p_cell.
#pragma once
#include "kml_gather_cell2.cuh"
#include <cstdint>

namespace kml {

#ifdef KML_MISC_KERNELS
__global__ void k_grid_pack_g2p(GridDev g, double *nvd) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.nn) return;
  const int k = (int)(n % g.n[2]); const long long t = n / g.n[2]; const int j = (int)(t % g.n[1]), i = (int)(t / g.n[1]);
  const double4 u = g.nvu[n], v = g.nv[n];
  double *d = nvd + nvd_index(g, i, j, k) * 6;
  *(double2 *)d = make_double2(u.x, u.y); *(double2 *)(d + 2) = make_double2(u.z, u.x - v.x); *(double2 *)(d + 4) = make_double2(u.y - v.y, u.z - v.z);
}
#endif

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_g2p_cell_tma(SolidDev s, GridDev g, StepParams sp, const double *__restrict__ nvd, const int *__restrict__ start, const int *__restrict__ order, int seglen, int nseg,
               int nitems) {
  extern __shared__ __align__(128) unsigned char smem_g3[]; // [2][16][TLEN] records of 48 B, then full[2], empty[2] mbarriers
  constexpr int NWARPS = THREADS / 32;
  const int TLEN = seglen + 3;
  const unsigned row_bytes = (unsigned)TLEN * 48u, tile_bytes = 16u * row_bytes;
  const unsigned smem0 = smem_addr(smem_g3);
  const unsigned bars = smem0 + 2u * tile_bytes; // full[0], full[1], empty[0], empty[1]
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(bars + 0, 1); mbar_init(bars + 8, 1); mbar_init(bars + 16, NWARPS); mbar_init(bars + 24, NWARPS);
    mbar_fence_init();
  }
  __syncthreads();

  // segment bounds of an item (uniform over the block): item = col * nseg + seg
  auto item_range = [&](int item, int &pb, int &pe) { // 32-bit: the cell lists exist only for grids below 2^31 nodes
    const int col = item / nseg; const int seg = item - col * nseg;
    const int kbeg = seg * seglen, kend = min(kbeg + seglen, g.n[2]);
    const int cellbase = col * g.n[2];
    pb = __ldg(&start[cellbase + kbeg]); pe = __ldg(&start[cellbase + kend]);
  };
  // producer (thread 0): request the tile of `item` into buffer b
  auto issue = [&](int item, int b) {
    const int col = item / nseg; const int seg = item - col * nseg;
    const int i0 = col / g.n[1], j0 = col % g.n[1], kbeg = seg * seglen;
    mbar_arrive_expect_tx(bars + 8u * b, tile_bytes);
    const unsigned dst = smem0 + (unsigned)b * tile_bytes;
#pragma unroll 1
    for (int r = 0; r < 16; r++)
      bulk_g2s(dst + (unsigned)r * row_bytes, nvd + nvd_index(g, i0 + (r >> 2), j0 + (r & 3), kbeg) * 6, row_bytes, bars + 8u * b);
  };

  const double h = g.h, ih = g.inv_cellsize;
  // producer cursor: the next non-empty item whose tile has not been requested yet, and how many have been requested
  int pitem = blockIdx.x; int pm = 0;
  auto produce_next = [&]() { // thread 0 only
    while (pitem < nitems) {
      int pb, pe; item_range(pitem, pb, pe);
      const int it = pitem; pitem += gridDim.x;
      if (pb == pe) continue;
      const int b = pm & 1, u = pm >> 1;
      if (u > 0) mbar_wait(bars + 16 + 8u * b, (unsigned)((u - 1) & 1)); // every warp is done with the previous use of this buffer
      issue(it, b); pm++;
      return;
    }
  };
  if (tid == 0) produce_next(); // tile of the first non-empty item

  int m = 0; // non-empty items this block has consumed
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    int pbeg, pend; item_range(item, pbeg, pend);
    if (pbeg == pend) continue; // block-uniform
    const int col = item / nseg; const int seg = item - col * nseg;
    const int i0 = col / g.n[1], j0 = col % g.n[1], kbeg = seg * seglen;

    // this thread's first particle: index and position in flight while the tile lands
    int p = pbeg + tid;
    int ip = p < pend ? order[p] : -1;
    int ipn = p + THREADS < pend ? order[p + THREADS] : -1;
    double px = 0, py = 0, pz = 0;
    if (ip >= 0) { px = s.x[0][ip]; py = s.x[1][ip]; pz = s.x[2][ip]; }
    if (tid == 0) produce_next(); // the tile after this one (its buffer was released by item m - 1)

    const int b = m & 1;
    mbar_wait(bars + 8u * b, (unsigned)((m >> 1) & 1));
    unsigned tile_s = smem0 + (unsigned)b * tile_bytes;
    asm volatile("" : "+r"(tile_s));
    const bool int_x = cubic_interior(i0, g.n[0], g.goff0, g.gn0), int_y = cubic_interior(j0, g.n[1], 0, g.n[1]);
    while (ip >= 0) {
      const int pn = p + THREADS;
      const int ipnn = pn + THREADS < pend ? order[pn + THREADS] : -1;
      double nx = 0, ny = 0, nz = 0;
      if (ipn >= 0) { nx = s.x[0][ipn]; ny = s.x[1][ipn]; nz = s.x[2][ipn]; }
      double vold[3];
      vold[0] = s.v[0][ip]; vold[1] = s.v[1][ip]; vold[2] = s.v[2][ip];

      const int k0 = cell_axis(pz, g.lo[2], ih, g.n[2], 0);
      const int koff = k0 - kbeg;
      double wx[4], wy[4], wz[4], dw_[4];
      cubic_axis4(px, g.lo[0], h, ih, i0, g.n[0], g.goff0, g.gn0, int_x, wx, dw_);
      cubic_axis4(py, g.lo[1], h, ih, j0, g.n[1], 0, g.n[1], int_y, wy, dw_);
      cubic_axis4(pz, g.lo[2], h, ih, k0, g.n[2], 0, g.n[2], cubic_interior(k0, g.n[2], 0, g.n[2]), wz, dw_);
      double vu[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int bb = 0; bb < 4; bb++) {
          const double gxy = wx[a] * wy[bb];
          const unsigned row = tile_s + (unsigned)(((a * 4 + bb) * TLEN + koff) * 48);
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const double2 r01 = lds_d2(row + 48 * c), r23 = lds_d2(row + 48 * c + 16), r45 = lds_d2(row + 48 * c + 32);
            const double wf = gxy * wz[c];
            vu[0] = fma(wf, r01.x, vu[0]); vu[1] = fma(wf, r01.y, vu[1]); vu[2] = fma(wf, r23.x, vu[2]);
            acc[0] = fma(wf, r23.y, acc[0]); acc[1] = fma(wf, r45.x, acc[1]); acc[2] = fma(wf, r45.y, acc[2]);
          }
        }
      particle_advance<false, false>(s, sp, ip, vu, acc, 0.0, vold);
      p = pn; ip = ipn; ipn = ipnn; px = nx; py = ny; pz = nz;
    }
    // this warp no longer reads buffer b
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(bars + 16 + 8u * b);
    m++;
  }
}

// One block per column segment like k_g2p_cell, but the tile arrives as 16 bulk copies (TMA, SASS UBLKCP) of the packed records onto one
// mbarrier.  k_g2p_cell fills its tile through registers in a loop ptxas unrolls by two (4 LDG -> 3 STS per pair of records): 9 records
// per thread are 4-5 serialised L2 round trips before the block's barrier, and ncu's samples put ~40 % of that kernel's warp time in the
// prologue.  Here the fill is one latency, in parallel with the start[] -> order[] -> x chain of the first particle; no second buffer, so
// the occupancy (8 blocks of 64 threads) is that of k_g2p_cell.
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_g2p_cell_bulk(SolidDev s, GridDev g, StepParams sp, const double *__restrict__ nvd, const int *__restrict__ start, const int *__restrict__ order, int seglen, int nseg) {
  extern __shared__ __align__(128) unsigned char smem_g4[]; // [16][TLEN] records of 48 B, then the mbarrier
  const int TLEN = seglen + 3;
  const unsigned row_bytes = (unsigned)TLEN * 48u, tile_bytes = 16u * row_bytes;
  const unsigned smem0 = smem_addr(smem_g4), bar = smem0 + tile_bytes;
  const int tid = threadIdx.x;
  const long long col = blockIdx.x / nseg; const int seg = (int)(blockIdx.x % nseg);
  const int i0 = (int)(col / g.n[1]), j0 = (int)(col % g.n[1]);
  const int kbeg = seg * seglen, kend = min(kbeg + seglen, g.n[2]);
  const long long cellbase = col * g.n[2];
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  const int pbeg = start[cellbase + kbeg], pend = start[cellbase + kend];
  if (pbeg == pend) return; // block-uniform
  if (tid == 0) { // the tile does not depend on the particles: request it before anything else waits
    mbar_arrive_expect_tx(bar, tile_bytes);
#pragma unroll 1
    for (int r = 0; r < 16; r++) bulk_g2s(smem0 + (unsigned)r * row_bytes, nvd + nvd_index(g, i0 + (r >> 2), j0 + (r & 3), kbeg) * 6, row_bytes, bar);
  }
  int p = pbeg + tid;
  int ip = p < pend ? order[p] : -1;
  int ipn = p + THREADS < pend ? order[p + THREADS] : -1;
  double px = 0, py = 0, pz = 0;
  if (ip >= 0) { px = s.x[0][ip]; py = s.x[1][ip]; pz = s.x[2][ip]; }
  __syncthreads(); // the barrier object is initialised for every thread that is about to wait on it
  mbar_wait(bar, 0);

  const bool int_x = cubic_interior(i0, g.n[0], g.goff0, g.gn0), int_y = cubic_interior(j0, g.n[1], 0, g.n[1]);
  const double h = g.h, ih = g.inv_cellsize;
  unsigned tile_s = smem0;
  asm volatile("" : "+r"(tile_s));
  while (ip >= 0) {
    const int pn = p + THREADS;
    const int ipnn = pn + THREADS < pend ? order[pn + THREADS] : -1;
    double nx = 0, ny = 0, nz = 0;
    if (ipn >= 0) { nx = s.x[0][ipn]; ny = s.x[1][ipn]; nz = s.x[2][ipn]; }
    double vold[3];
    vold[0] = s.v[0][ip]; vold[1] = s.v[1][ip]; vold[2] = s.v[2][ip];
    const int k0 = cell_axis(pz, g.lo[2], ih, g.n[2], 0);
    const int koff = k0 - kbeg;
    double wx[4], wy[4], wz[4], dw_[4];
    cubic_axis4(px, g.lo[0], h, ih, i0, g.n[0], g.goff0, g.gn0, int_x, wx, dw_);
    cubic_axis4(py, g.lo[1], h, ih, j0, g.n[1], 0, g.n[1], int_y, wy, dw_);
    cubic_axis4(pz, g.lo[2], h, ih, k0, g.n[2], 0, g.n[2], cubic_interior(k0, g.n[2], 0, g.n[2]), wz, dw_);
    double vu[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int bb = 0; bb < 4; bb++) {
        const double gxy = wx[a] * wy[bb];
        const unsigned row = tile_s + (unsigned)(((a * 4 + bb) * TLEN + koff) * 48);
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const double2 r01 = lds_d2(row + 48 * c), r23 = lds_d2(row + 48 * c + 16), r45 = lds_d2(row + 48 * c + 32);
          const double wf = gxy * wz[c];
          vu[0] = fma(wf, r01.x, vu[0]); vu[1] = fma(wf, r01.y, vu[1]); vu[2] = fma(wf, r23.x, vu[2]);
          acc[0] = fma(wf, r23.y, acc[0]); acc[1] = fma(wf, r45.x, acc[1]); acc[2] = fma(wf, r45.y, acc[2]);
        }
      }
    particle_advance<false, false>(s, sp, ip, vu, acc, 0.0, vold);
    p = pn; ip = ipn; ipn = ipnn; px = nx; py = ny; pz = nz;
  }
}

inline int cell_g2p_bulk_launch(const SolidDev &s, const GridDev &g, const StepParams &sp, const double *nvd, const CellLists &cl, cudaStream_t st, int seg_target, int threads) {
  int seglen, nseg; cell_segments(g.n[2], seg_target, &seglen, &nseg);
  const long long nblocks = (long long)g.n[0] * g.n[1] * nseg;
  if (nblocks >= (1ll << 31)) return 1;
  const size_t smem = (size_t)16 * (seglen + 3) * 48 + 16;
#define KML_G4_LAUNCH(T, B)                                                                                                              \
  do {                                                                                                                                   \
    auto kern = k_g2p_cell_bulk<T, B>;                                                                                                   \
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1; \
    kern<<<(unsigned)nblocks, T, smem, st>>>(s, g, sp, nvd, cl.start, cl.order, seglen, nseg);                                           \
  } while (0)
  if (threads == 64) KML_G4_LAUNCH(64, 8); else KML_G4_LAUNCH(128, 4);
#undef KML_G4_LAUNCH
  return cudaGetLastError() != cudaSuccess;
}

// returns 0 = launched, 1 = CUDA error
inline int cell_g2p_tma_launch(const SolidDev &s, const GridDev &g, const StepParams &sp, const double *nvd, const CellLists &cl, cudaStream_t st, int seg_target, int threads,
                               int blocks_per_sm, int nsm) {
  int seglen, nseg; cell_segments(g.n[2], seg_target, &seglen, &nseg);
  const long long nitems = (long long)g.n[0] * g.n[1] * nseg;
  if (nitems >= (1ll << 31) - 148 * 16) return 1;
  const size_t smem = (size_t)2 * 16 * (seglen + 3) * 48 + 64;
  const unsigned grid = (unsigned)std::min<long long>(nitems, (long long)nsm * blocks_per_sm);
#define KML_G3_LAUNCH(T, B)                                                                                                              \
  do {                                                                                                                                   \
    auto kern = k_g2p_cell_tma<T, B>;                                                                                                    \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;                     \
    kern<<<grid, T, smem, st>>>(s, g, sp, nvd, cl.start, cl.order, seglen, nseg, (int)nitems);                                                \
  } while (0)
  if (threads == 64) KML_G3_LAUNCH(64, 8); else if (blocks_per_sm == 3) KML_G3_LAUNCH(128, 3); else KML_G3_LAUNCH(128, 4);
#undef KML_G3_LAUNCH
  return cudaGetLastError() != cudaSuccess;
}

} // namespace kml
