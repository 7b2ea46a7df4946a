// kml_comm.cuh - slab decomposition over several GPUs: halo sums of the shared node planes and
// particle migration with NCCL over NVLink.  Replaces Grid::reduce_mass_ghost_nodes /
// reduce_ghost_nodes (reference src/grid.cpp:477-621, :881-1132) and ULMPM::exchange_particles
// (src/ulmpm.cpp:565-667) + Solid::pack/unpack_particle (src/solid.cpp:1612-1808).
//
// The grid is cut along x (the slowest node index), so the planes shared with a neighbour are one
// contiguous range of every node array: no pack kernel is needed for the halo - the partial sums are
// sent straight from the arrays, received into scratch and added.  Both neighbours send their partial
// sums and add what they receive (a + b == b + a in IEEE arithmetic), so both hold identical totals
// after ONE exchange, where the reference needs a reduce and a broadcast-back.
#pragma once
#include "kml_kernels.cuh"
#include "kml_nccl.h"

namespace kml {

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  double *halo_recv = nullptr; size_t halo_bytes = 0;      // scratch for the received planes
  int *mig_cnt = nullptr;                                   // device: sendL, sendR, recvL, recvR, nhole, nfill
  int *mig_list = nullptr; int mig_cap = 0;                 // device: [2][mig_cap] leaving particle ids, then holes / fillers
  int *mig_flag = nullptr;                                  // device: [2 * mig_cap] leaver flags of the vacated tail
  double *mig_send = nullptr, *mig_recv = nullptr; size_t mig_bytes = 0;
  int *h_cnt = nullptr;                                     // pinned
};

__global__ void k_halo_add_nv(double4 *dst, const double4 *src, long long n, int add_mass) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 a = dst[i]; const double4 b = src[i];
  a.x += b.x; a.y += b.y; a.z += b.z; if (add_mass) a.w += b.w;
  dst[i] = a;
}
__global__ void k_halo_add(double *dst, const double *src, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// ---- migration ---------------------------------------------------------------------------------
// dest of every particle from the GLOBAL stencil base of its (new) position; leavers are appended to a list
__global__ void k_mig_mark(const double *x0, long long np, double lo, double ih, int shape_linear, int base_lo, int base_hi, int rank, int nranks,
                           int *cnt, int *list, int cap) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const double t = __dmul_rn(__dsub_rn(x0[ip], lo), ih);
  const int b = shape_linear ? (int)t : (int)__dsub_rn(t, 1.0);
  int side = -1;
  if (b < base_lo && rank > 0) side = 0;
  else if (b >= base_hi && rank < nranks - 1) side = 1;
  if (side >= 0) { const int slot = atomicAdd(&cnt[side], 1); if (slot < cap) list[side * cap + slot] = (int)ip; }
}
// rows = [narr + 2][n]: the double arrays, then ptag (bit pattern) and mask
__global__ void k_mig_pack(double *rows, const int *list, int n, double *const base, long long cap, int narr, const long long *ptag, const int *mask) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int ip = list[j];
  for (int a = 0; a < narr; a++) rows[(long long)a * n + j] = base[a * cap + ip];
  rows[(long long)narr * n + j] = __longlong_as_double(ptag[ip]);
  rows[(long long)(narr + 1) * n + j] = (double)mask[ip];
}
__global__ void k_mig_unpack(const double *rows, int n, long long at, double *base, long long cap, int narr, long long *ptag, int *mask) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const long long ip = at + j;
  for (int a = 0; a < narr; a++) base[a * cap + ip] = rows[(long long)a * n + j];
  ptag[ip] = __double_as_longlong(rows[(long long)narr * n + j]);
  mask[ip] = (int)rows[(long long)(narr + 1) * n + j];
}
// holes = leavers below np_new, fillers = stayers at or above np_new (equal counts)
__global__ void k_mig_holes(const int *list, int cap, int nL, int nR, long long np_new, int *cnt, int *holes) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nL + nR) return;
  const int ip = j < nL ? list[j] : list[cap + (j - nL)];
  if (ip < np_new) holes[atomicAdd(&cnt[4], 1)] = ip;
}
__global__ void k_mig_flag(const int *list, int cap, int nL, int nR, int *flag, long long np_new) { // flag[ip - np_new] = 1 for leavers in the tail
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nL + nR) return;
  const int ip = j < nL ? list[j] : list[cap + (j - nL)];
  if (ip >= np_new) flag[ip - np_new] = 1;
}
__global__ void k_mig_fillers(const int *flag, long long np_new, int ntail, int *cnt, int *fillers) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ntail) return;
  if (!flag[j]) fillers[atomicAdd(&cnt[5], 1)] = (int)(np_new + j);
}
__global__ void k_mig_move(const int *holes, const int *fillers, int n, double *base, long long cap, int narr, long long *ptag, int *mask) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int h = holes[j], f = fillers[j];
  for (int a = 0; a < narr; a++) base[a * cap + h] = base[a * cap + f];
  ptag[h] = ptag[f]; mask[h] = mask[f];
}

} // namespace kml
