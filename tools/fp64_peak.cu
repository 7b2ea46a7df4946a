// fp64_peak.cu - measured DFMA issue rate of one B200 (the FP64 roof of DESIGN.md section 3).
// Every thread runs ILP independent DFMA chains; blocks fill every SM with WARPS warps per sub-partition.
// Prints DFMA per clock per SM (from the SM clock the kernel itself reads) and T DFMA/s.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o karamelo_b200/lib/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double *out, int iters, long long *clocks) {
  double a[ILP], x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = i + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = fma(a[i], x, y);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

// Mixed issue: every DFMA is followed by NI independent integer IMADs (and NL shared-memory loads per 8 DFMAs).  If a DFMA occupied only
// its own pipe, the integer work would hide in the pipe's second cycle and the DFMA rate would not move until NI > 1; what the B200
// does is the measurement behind the "issue slots" column of DESIGN.md section 3.
template <int ILP, int NI, int NL>
__global__ void k_mixed(double *out, int iters, long long *clocks) {
  __shared__ double sh[256];
  sh[threadIdx.x & 255] = threadIdx.x;
  __syncthreads();
  double a[ILP], x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999, ld = 0;
  int k[ILP * (NI > 0 ? NI : 1)];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < ILP * (NI > 0 ? NI : 1); i++) k[i] = i + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      a[i] = fma(a[i], x, y);
#pragma unroll
      for (int j = 0; j < NI; j++) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(k[i * NI + j]) : "r"(blockDim.x), "r"(it)); // one IMAD, not foldable
    }
#pragma unroll
    for (int j = 0; j < NL; j++) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(sh) + 8u * ((threadIdx.x + j * 32) & 255))); ld += v; }
  }
  long long t1 = clock64();
  double s = ld;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
#pragma unroll
  for (int i = 0; i < ILP * (NI > 0 ? NI : 1); i++) s += k[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}
template <int ILP, int NI, int NL> void run_mixed(int nsm, int threads, int blocks_per_sm, int iters) {
  const int nb = nsm * blocks_per_sm;
  double *out; long long *clk; cudaMalloc(&out, sizeof(double) * nb * threads); cudaMalloc(&clk, sizeof(long long) * nb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_mixed<ILP, NI, NL><<<nb, threads>>>(out, iters / 8, clk); cudaDeviceSynchronize();
  cudaEventRecord(e0); k_mixed<ILP, NI, NL><<<nb, threads>>>(out, iters, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = (double)nb * threads * ILP * iters;
  printf("mixed: %d IMAD per DFMA, %d LDS per %d DFMA, %d warps/sub-partition: %.2f T DFMA/s, %.2f T warp-instr/s issued (%.3f ms)\n", NI, NL, ILP,
         threads * blocks_per_sm / 128, dfma / (ms * 1e-3) * 1e-12, dfma * (1 + NI + (double)NL / ILP + 2.0 / ILP) / 32 / (ms * 1e-3) * 1e-12, ms);
  cudaFree(out); cudaFree(clk);
}

// Three REGISTER operands per DFMA, as in the kernels' accumulate loops (acc[i] += b[j] * c[k], 28 accumulators, 4 + 4 multiplicands): the
// chains above feed two of the three operands from a uniform register and a reuse cache, which costs the register file one read per DFMA
// instead of three.  NL: LDS.128 per 28 DFMAs whose results replace the multiplicands (as the record loads of P2G do).
template <int NL>
__global__ void k_dfma3(double *out, const double *in, int iters, long long *clocks) {
  __shared__ __align__(16) double sh[512];
  sh[threadIdx.x & 511] = in[threadIdx.x & 63]; sh[(threadIdx.x + 256) & 511] = in[(threadIdx.x + 7) & 63];
  __syncthreads();
  double a[28], b[4], c[4];
#pragma unroll
  for (int i = 0; i < 28; i++) a[i] = i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; i++) { b[i] = in[i] + 1e-9 * threadIdx.x; c[i] = in[4 + i]; }
  const unsigned base = (unsigned)__cvta_generic_to_shared(sh) + 16u * (threadIdx.x & 7);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 28; i++) a[i] = fma(b[i & 3], c[(i >> 2) & 3], a[i]);
#pragma unroll
    for (int j = 0; j < NL; j++) { // the loaded values become multiplicands of the next iteration
      double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + 128u * j));
      if (j & 1) { c[(j >> 1) & 3] = v.x; b[(j >> 1) & 3] = v.y; } else { b[(j >> 1) & 3] = v.x; c[(j >> 1) & 3] = v.y; }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 28; i++) s += a[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}
template <int NL> void run_dfma3(int nsm, int threads, int blocks_per_sm, int iters) {
  const int nb = nsm * blocks_per_sm;
  double *out, *in; long long *clk; cudaMalloc(&out, sizeof(double) * nb * threads); cudaMalloc(&clk, sizeof(long long) * nb); cudaMalloc(&in, 64 * sizeof(double));
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1.0 + 1e-7 * i; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma3<NL><<<nb, threads>>>(out, in, iters / 8, clk); cudaDeviceSynchronize();
  cudaEventRecord(e0); k_dfma3<NL><<<nb, threads>>>(out, in, iters, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = (double)nb * threads * 28 * iters;
  printf("3 register operands per DFMA, 28 accumulators, %d LDS.128 per 28 DFMA, %d warps/sub-partition: %.2f T DFMA/s (%.3f ms)\n", NL, threads * blocks_per_sm / 128,
         dfma / (ms * 1e-3) * 1e-12, ms);
  cudaFree(out); cudaFree(clk); cudaFree(in);
}

template <int ILP> void run(int nsm, int threads, int blocks_per_sm, int iters) {
  const int nb = nsm * blocks_per_sm;
  double *out; long long *clk; cudaMalloc(&out, sizeof(double) * nb * threads); cudaMalloc(&clk, sizeof(long long) * nb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma<ILP><<<nb, threads>>>(out, iters / 8, clk); cudaDeviceSynchronize();
  cudaEventRecord(e0); k_dfma<ILP><<<nb, threads>>>(out, iters, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long *h = new long long[nb]; cudaMemcpy(h, clk, sizeof(long long) * nb, cudaMemcpyDeviceToHost);
  double cavg = 0; for (int i = 0; i < nb; i++) cavg += h[i]; cavg /= nb;
  const double dfma = (double)nb * threads * ILP * iters;
  printf("ILP %2d threads/block %4d blocks/SM %d: %.2f DFMA/clk/SM (block clocks), %.2f T DFMA/s (events, %.3f ms)\n", ILP, threads, blocks_per_sm,
         (double)threads * blocks_per_sm * ILP * iters / cavg, dfma / (ms * 1e-3) * 1e-12, ms);
  delete[] h; cudaFree(out); cudaFree(clk);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s: %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  const int nsm = p.multiProcessorCount, iters = 1 << 15;
  run<8>(nsm, 128, 1, iters);   // 1 warp per sub-partition
  run<8>(nsm, 128, 3, iters);   // 3 warps per sub-partition (the occupancy of k_p2g_cell3)
  run<8>(nsm, 256, 4, iters);   // 8 warps per sub-partition
  run<16>(nsm, 128, 3, iters);
  run<2>(nsm, 1024, 2, iters);  // 16 warps per sub-partition, little ILP
  // what shares the issue port with the FP64 pipe (3 warps per sub-partition, the occupancy of the P2G / stress kernels)
  run_mixed<8, 0, 0>(nsm, 128, 3, iters);
  run_mixed<8, 1, 0>(nsm, 128, 3, iters);
  run_mixed<8, 2, 0>(nsm, 128, 3, iters);
  run_mixed<8, 0, 2>(nsm, 128, 3, iters);
  run_mixed<8, 1, 2>(nsm, 128, 3, iters);
  run_mixed<8, 1, 0>(nsm, 128, 4, iters);
  run_dfma3<0>(nsm, 128, 1, iters / 4);
  run_dfma3<0>(nsm, 128, 3, iters / 4);
  run_dfma3<4>(nsm, 128, 3, iters / 4);
  run_dfma3<8>(nsm, 128, 3, iters / 4);
  run_dfma3<8>(nsm, 128, 4, iters / 4);
  return 0;
}
