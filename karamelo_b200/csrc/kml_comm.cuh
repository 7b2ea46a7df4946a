// kml_comm.cuh - slab decomposition over several GPUs: halo sums of the shared node planes and
// particle migration with NCCL over NVLink.  Replaces Grid::reduce_mass_ghost_nodes /
// reduce_ghost_nodes (reference src/grid.cpp:477-621, :881-1132) and ULMPM::exchange_particles
// (src/ulmpm.cpp:565-667) + Solid::pack/unpack_particle (src/solid.cpp:1612-1808).
//
// The grid is cut along x (the slowest node index), so the planes shared with a neighbour are one
// contiguous range of every node array.  All exchanged fields are packed into ONE message per side
// (one pack kernel, one grouped send/recv pair per neighbour, one add kernel): the exchange is latency-
// bound (9 MB per side at 100 M particles), so the number of NCCL operations matters more than the two
// extra passes over the planes.  Both neighbours send their partial sums and add what they receive
// (a + b == b + a in IEEE arithmetic), so both hold identical totals after ONE exchange, where the
// reference needs a reduce and a broadcast-back.
#pragma once
#include "kml_kernels.cuh"
#include "kml_nccl.h"

namespace kml {

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  double *halo_buf = nullptr; size_t halo_bytes = 0;       // scratch: [sendL | sendR | recvL | recvR] packed planes
  int *mig_cnt = nullptr;                                   // device: sendL, sendR, recvL, recvR, nhole, nfill
  int *mig_list = nullptr; int mig_cap = 0;                 // device: [2][mig_cap] leaving particle ids, then holes / fillers
  int *mig_flag = nullptr;                                  // device: [2 * mig_cap] leaver flags of the vacated tail
  double *mig_send = nullptr, *mig_recv = nullptr; size_t mig_bytes = 0;
  int *h_cnt = nullptr;                                     // pinned
  // peer-memory halo (see k_halo_push): this rank's landing zone [flags | from-left x 2 | from-right x 2] and the neighbours' zones mapped with CUDA IPC
  int peer_state = 0;                                       // 0 = not tried, 1 = in use, -1 = unavailable (NCCL send/recv is used)
  unsigned char *zone = nullptr; size_t zone_slot = 0;      // slot = bytes of one landing buffer
  unsigned char *zone_left = nullptr, *zone_right = nullptr; // the neighbours' zones
  unsigned long long seq = 0;                                // exchanges done so far
  unsigned *push_done = nullptr;                             // device: blocks of the running push kernel that have finished
};
constexpr size_t HALO_ZONE_HDR = 256; // two 8-byte sequence flags (from the left at 0, from the right at 128), each in its own line

// Shared planes of every exchanged field -> one contiguous message per side, and back (adding).  A field is `width`
// doubles per node; the planes shared with the left neighbour start at node 0, those shared with the right one at
// node (n0 - nsh) * plane.  Message layout: field after field, each cnt * width doubles.
struct HaloFields { double *ptr[12]; int width[12]; int skip_w[12]; int n; }; // skip_w: do not add component 3 (the mass of a momentum-only pass)
__global__ void k_halo_pack(HaloFields hf, long long cnt, long long top_off_nodes, double *outL, double *outR, int left, int right) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; // node within the shared planes
  if (i >= cnt) return;
  long long off = 0;
  for (int f = 0; f < hf.n; f++) {
    const int w = hf.width[f];
    for (int k = 0; k < w; k++) {
      if (left) outL[off + i * w + k] = hf.ptr[f][i * w + k];
      if (right) outR[off + i * w + k] = hf.ptr[f][(top_off_nodes + i) * w + k];
    }
    off += cnt * w;
  }
}
__global__ void k_halo_add(HaloFields hf, long long cnt, long long top_off_nodes, const double *inL, const double *inR, int left, int right) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  long long off = 0;
  for (int f = 0; f < hf.n; f++) {
    const int w = hf.width[f];
    for (int k = 0; k < w; k++) {
      if (k == 3 && hf.skip_w[f]) continue;
      if (left) hf.ptr[f][i * w + k] += inL[off + i * w + k];
      if (right) hf.ptr[f][(top_off_nodes + i) * w + k] += inR[off + i * w + k];
    }
    off += cnt * w;
  }
}

// ---- halo sums over NVLink peer memory ------------------------------------------------------------------------------------------------
// The NCCL version above costs a pack kernel, one grouped send/recv pair per neighbour (two proxied launches, ~30 us each way before the
// first byte moves) and an add kernel per exchange, twice per step: 0.35-0.6 ms of a 6 ms step on 8 GPUs for messages NVLink moves in 20 us.
// Here ONE kernel packs the shared planes straight into the neighbour's landing buffer (stores to peer memory mapped with CUDA IPC - NVSwitch
// gives every GPU full bandwidth to every peer) and raises a sequence flag there; the add kernel of the receiver waits for that flag and
// sums from its own memory.  No collective call, no proxy thread, no host involvement on the data path.
//   * two landing buffers per side, used alternately: a neighbour can be at most one exchange ahead (it needs my push of exchange e + 1
//     before it can push e + 2, and my push of e + 1 is ordered after my add of e on my stream), so buffer e mod 2 is free again;
//   * flags are monotonic exchange counters written with a system-scope release after every block's stores were fenced; the receiver
//     polls with volatile loads and reads the payload with ld.global.cg (the lines may sit stale in its L1 from two exchanges ago).
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) { unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__global__ void k_halo_push(HaloFields hf, long long cnt, long long top_off_nodes, double *remoteL, double *remoteR, unsigned long long *flagL, unsigned long long *flagR,
                            unsigned long long seq, unsigned *done) { // remoteL: landing buffer "from the right" in the LEFT neighbour's zone (null: no neighbour)
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) {
    long long off = 0;
    for (int f = 0; f < hf.n; f++) {
      const int w = hf.width[f];
      for (int k = 0; k < w; k++) {
        if (remoteL) remoteL[off + i * w + k] = hf.ptr[f][i * w + k];
        if (remoteR) remoteR[off + i * w + k] = hf.ptr[f][(top_off_nodes + i) * w + k];
      }
      off += cnt * w;
    }
  }
  __threadfence_system(); // this thread's stores are ordered before whatever follows, for every observer
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) { // the last block: every other block's stores are fenced and counted
      *done = 0;
      __threadfence_system();
      if (flagL) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flagL), "l"(seq) : "memory");
      if (flagR) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flagR), "l"(seq) : "memory");
    }
  }
}
__device__ __forceinline__ unsigned long long global_timer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
constexpr unsigned long long HALO_TIMEOUT_NS = 30ull * 1000 * 1000 * 1000; // a neighbour that has not delivered after 30 s is gone: raise bit 6 of the error word instead of spinning for ever
__global__ void k_halo_wait_add(HaloFields hf, long long cnt, long long top_off_nodes, const double *inL, const double *inR, const unsigned long long *flagL,
                                const unsigned long long *flagR, unsigned long long seq, unsigned *err) { // inL: my landing buffer "from the left" (null: no neighbour)
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_timer_ns();
    if (inL) while (ld_volatile_u64(flagL) < seq) { __nanosleep(100); if (global_timer_ns() - t0 > HALO_TIMEOUT_NS) { atomicOr(err, 64u); break; } }
    if (inR) while (ld_volatile_u64(flagR) < seq) { __nanosleep(100); if (global_timer_ns() - t0 > HALO_TIMEOUT_NS) { atomicOr(err, 64u); break; } }
    __threadfence_system();
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  long long off = 0;
  for (int f = 0; f < hf.n; f++) {
    const int w = hf.width[f];
    for (int k = 0; k < w; k++) {
      if (k == 3 && hf.skip_w[f]) continue;
      if (inL) hf.ptr[f][i * w + k] += __ldcg(inL + off + i * w + k);
      if (inR) hf.ptr[f][(top_off_nodes + i) * w + k] += __ldcg(inR + off + i * w + k);
    }
    off += cnt * w;
  }
}

// Grid::reduce_rigid_ghost_nodes (src/grid.cpp:746-879): a node is rigid if ANY rank's rigid particle reaches it - the flags of the shared
// planes are OR-ed with the neighbours' (same message layout as the sums, one int per node)
__global__ void k_halo_pack_flag(const int *a, long long cnt, long long top_off_nodes, int *outL, int *outR, int left, int right) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  if (left) outL[i] = a[i];
  if (right) outR[i] = a[top_off_nodes + i];
}
__global__ void k_halo_or_flag(int *a, long long cnt, long long top_off_nodes, const int *inL, const int *inR, int left, int right) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  if (left) a[i] |= inL[i];
  if (right) a[top_off_nodes + i] |= inR[i];
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
// rows = [narr + 2][n]: the double arrays, then ptag (bit pattern) and mask.  The wire order is canonical (current positions x first, then
// the scratch xn): the positions trade buffer slots after every step and a rank that re-ordered its particles physically starts over at
// slot 0, so sender and receiver need not agree on where x lives (xs = buffer slot of x: 0 or 3).
__device__ __forceinline__ int mig_slot(int a, int xs) { return a < 3 ? xs + a : (a < 6 ? (3 - xs) + (a - 3) : a); }
__global__ void k_mig_pack(double *rows, const int *list, int n, double *const base, long long cap, int narr, int xs, const long long *ptag, const int *mask) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int ip = list[j];
  for (int a = 0; a < narr; a++) rows[(long long)a * n + j] = base[mig_slot(a, xs) * cap + ip];
  rows[(long long)narr * n + j] = __longlong_as_double(ptag[ip]);
  rows[(long long)(narr + 1) * n + j] = (double)mask[ip];
}
__global__ void k_mig_unpack(const double *rows, int n, long long at, double *base, long long cap, int narr, int xs, long long *ptag, int *mask) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const long long ip = at + j;
  for (int a = 0; a < narr; a++) base[mig_slot(a, xs) * cap + ip] = rows[(long long)a * n + j];
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
// the hole / filler counts stay on the device (cnt[4], cnt[5]; equal by construction - a mismatch raises bit 5 of the error word)
__global__ void k_mig_move(const int *holes, const int *fillers, const int *cnt, unsigned *flags, double *base, long long cap, int narr, long long *ptag, int *mask) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(cnt[4], cnt[5]);
  if (j == 0 && cnt[4] != cnt[5]) atomicOr(flags, 32u);
  if (j >= n) return;
  const int h = holes[j], f = fillers[j];
  for (int a = 0; a < narr; a++) base[a * cap + h] = base[a * cap + f];
  ptag[h] = ptag[f]; mask[h] = mask[f];
}

} // namespace kml
