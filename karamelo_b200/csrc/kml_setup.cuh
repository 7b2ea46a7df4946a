// kml_setup.cuh - set-up on the device: the particle lattice of Solid::populate
// (src/solid.cpp:1810-2336), region predicates (src/region_block.cpp, src/region_cylinder.cpp, src/region_sphere.cpp), particle
// group masks (Group::assign, src/group.cpp:65-238) and per-particle script expressions as postfix programs.
//
// Every floating-point operation that decides an integer (is the point inside the region? which slab?) or a stored position is written
// with explicitly rounded intrinsics in the order of the host expression, so that the device result is bit-identical to the host path
// (and to the reference): no fused multiply-add may sneak into `a * a + b * b`.
#pragma once
#include "kml_kernels.cuh"
#include <cub/device/device_scan.cuh>

namespace kml {

__device__ __forceinline__ int region_inside(const kml_region &r, double x, double y, double z) {
  if (r.style == KML_REGION_BLOCK) return x >= r.p[0] && x <= r.p[1] && y >= r.p[2] && y <= r.p[3] && z >= r.p[4] && z <= r.p[5];
  if (r.style == KML_REGION_CYLINDER) {
    double a, b, t; // axis x: (y, z | x), y: (x, z | y), z: (x, y | z) - src/region_cylinder.cpp:140-161
    if (r.axis == 0) { a = y; b = z; t = x; } else if (r.axis == 1) { a = x; b = z; t = y; } else { a = x; b = y; t = z; }
    const double da = __dsub_rn(a, r.p[0]), db = __dsub_rn(b, r.p[1]);
    const double dSq = __dadd_rn(__dmul_rn(da, da), __dmul_rn(db, db));
    return t >= r.p[3] && t <= r.p[4] && dSq <= r.p[2];
  }
  const double dx = __dsub_rn(x, r.p[0]), dy = __dsub_rn(y, r.p[1]), dz = __dsub_rn(z, r.p[2]);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) <= r.p[3];
}
__device__ __forceinline__ int region_match(const kml_region &r, double x, double y, double z) { // Region::match, src/region.cpp:64-72
  const int in = region_inside(r, x, y, z);
  return r.interior ? in : !in;
}

// lattice point L = ((i nsub1 + j) nsub2 + k) nip + q -> position; returns whether it becomes a particle (of this slab when slab_filter)
__device__ __forceinline__ bool lattice_point(const kml_lattice &l, const kml_region &r, long long L, bool slab_filter, double *x, int *base_out) {
  const int q = (int)(L % l.nip); long long c = L / l.nip;
  const int k = (int)(c % l.nsub[2]); c /= l.nsub[2];
  const int j = (int)(c % l.nsub[1]); const int i = (int)(c / l.nsub[1]);
  // boundlo + delta * (noffsetlo + i + 0.5 + ip), src/solid.cpp:2163-2180
  x[0] = __dadd_rn(l.boundlo[0], __dmul_rn(l.delta, __dadd_rn(__dadd_rn((double)(l.noffsetlo[0] + i), 0.5), l.ip[3 * q + 0])));
  x[1] = __dadd_rn(l.boundlo[1], __dmul_rn(l.delta, __dadd_rn(__dadd_rn((double)(l.noffsetlo[1] + j), 0.5), l.ip[3 * q + 1])));
  x[2] = l.dim == 3 ? __dadd_rn(l.boundlo[2], __dmul_rn(l.delta, __dadd_rn(__dadd_rn((double)(l.noffsetlo[2] + k), 0.5), l.ip[3 * q + 2]))) : 0.0;
  const bool in_sub = !(x[0] < l.sublo[0] || x[0] > l.subhi[0] || x[1] < l.sublo[1] || x[1] > l.subhi[1] || x[2] < l.sublo[2] || x[2] > l.subhi[2]);
  if (!in_sub || region_inside(r, x[0], x[1], x[2]) != 1) return false; // src/solid.cpp:2183 (inside == 1, not match)
  int base = 0;
  if (l.slab) {
    const double t = __dmul_rn(__dsub_rn(x[0], l.slab_lo), l.slab_ih);
    base = l.slab_linear ? (int)t : (int)__dsub_rn(t, 1.0);
  }
  if (base_out) *base_out = base;
  return !slab_filter || !l.slab || (base >= l.base_lo && base < l.base_hi);
}

// accepted points per clamped stencil base (bin 0 without slabs)
__global__ void k_lattice_hist(kml_lattice l, kml_region r, long long ntot, unsigned long long *hist, int nbins) {
  const long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double x[3]; int base = 0;
  const bool ok = L < ntot && lattice_point(l, r, L, false, x, &base);
  const int bin = ok ? min(max(base, 0), nbins - 1) : -1;
  const unsigned peers = __match_any_sync(0xffffffffu, bin); // lanes of a warp are consecutive lattice points: 1-2 bins
  if (ok && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned long long)__popc(peers));
}

// block-wise order-preserving compaction: counts per block of 1024 lattice points, then (after an exclusive scan) the fill
constexpr int LATTICE_BLOCK = 1024;
__global__ void __launch_bounds__(LATTICE_BLOCK) k_lattice_count(kml_lattice l, kml_region r, long long ntot, long long *counts) {
  const long long L = (long long)blockIdx.x * LATTICE_BLOCK + threadIdx.x;
  double x[3];
  const bool ok = L < ntot && lattice_point(l, r, L, true, x, nullptr);
  const int n = __syncthreads_count(ok);
  if (threadIdx.x == 0) counts[blockIdx.x] = n;
}
__global__ void __launch_bounds__(LATTICE_BLOCK) k_lattice_fill(kml_lattice l, kml_region r, long long ntot, const long long *offsets, SolidDev s, long long tag0) {
  __shared__ int wsum[LATTICE_BLOCK / 32];
  const long long L = (long long)blockIdx.x * LATTICE_BLOCK + threadIdx.x;
  double x[3];
  const bool ok = L < ntot && lattice_point(l, r, L, true, x, nullptr);
  const unsigned b = __ballot_sync(0xffffffffu, ok);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wsum[w] = __popc(b);
  __syncthreads();
  int before = 0;
  for (int i = 0; i < w; i++) before += wsum[i];
  if (!ok) return;
  const long long ip = offsets[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
  if (ip >= s.np) return; // cannot happen (np was derived from the same predicate); never write out of bounds
  for (int d = 0; d < 3; d++) { s.x[d][ip] = x[d]; s.x0[d][ip] = x[d]; }
  s.ptag[ip] = tag0 + ip; s.mask[ip] = 1;
  double m = l.mass, v = l.vol;
  if (l.axisymmetric) { m = __dmul_rn(l.mass, x[0]); v = __ddiv_rn(m, l.rho0); } // src/solid.cpp:2292-2298
  s.mass[ip] = m; s.vol0[ip] = v; s.vol[ip] = v; s.rho0[ip] = l.rho0;
  if (l.set_T) s.T[ip] = l.T0;
  s.F[0][ip] = 1.0; s.F[4][ip] = 1.0; s.F[8][ip] = 1.0;
  if (s.R[0]) { s.R[0][ip] = 1.0; s.R[4][ip] = 1.0; s.R[8][ip] = 1.0; }
}

// initial values of Solid::populate for solids whose particles are uploaded by the host (src/solid.cpp:2283-2321)
__global__ void k_solid_init(SolidDev s, double rho0) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= s.np) return;
  s.F[0][ip] = 1.0; s.F[4][ip] = 1.0; s.F[8][ip] = 1.0;
  if (s.R[0]) { s.R[0][ip] = 1.0; s.R[4][ip] = 1.0; s.R[8][ip] = 1.0; }
  s.rho0[ip] = rho0; s.mask[ip] = 1;
}

// Group::assign (particles), src/group.cpp:140-181
__global__ void k_group_assign(SolidDev s, kml_region r, int bit, unsigned long long *count) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = ip < s.np && region_match(r, s.x0[0][ip], s.x0[1][ip], s.x0[2][ip]);
  if (in) s.mask[ip] |= bit;
  const unsigned b = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (unsigned long long)__popc(b));
}

__global__ void k_sum(const double *a, long long n, double *out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0;
  for (; i < n; i += (long long)gridDim.x * blockDim.x) v += a[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(out, v);
}

// ---- per-particle expressions -------------------------------------------------------------------------------------------
// Postfix evaluation with a small register stack.  Arithmetic is explicitly rounded (no contraction across operations).
__device__ __forceinline__ double expr_eval(const kml_expr &e, const double *pv) {
  double st[16]; int sp = 0;
  for (int i = 0; i < e.n; i++) {
    const int op = e.op[i];
    if (op == KML_X_CONST) { st[sp++] = e.val[i]; continue; }
    if (op == KML_X_VAR) { st[sp++] = pv[(int)e.val[i]]; continue; }
    if (op == KML_X_NEG) { st[sp - 1] = -st[sp - 1]; continue; }
    if (op == KML_X_NOT) { st[sp - 1] = (double)!st[sp - 1]; continue; }
    if (op >= KML_X_EXP && op <= KML_X_LOG) {
      const double a = st[sp - 1];
      st[sp - 1] = op == KML_X_EXP ? exp(a) : op == KML_X_SQRT ? sqrt(a) : op == KML_X_COS ? cos(a) : op == KML_X_SIN ? sin(a) : op == KML_X_TAN ? tan(a) : log(a);
      continue;
    }
    const double b = st[--sp], a = st[sp - 1];
    double r;
    switch (op) {
    case KML_X_ADD: r = __dadd_rn(a, b); break; case KML_X_SUB: r = __dsub_rn(a, b); break;
    case KML_X_MUL: r = __dmul_rn(a, b); break; case KML_X_DIV: r = __ddiv_rn(a, b); break;
    case KML_X_POW: r = pow(a, b); break; case KML_X_ATAN2: r = atan2(a, b); break;
    case KML_X_GT: r = a > b; break; case KML_X_GE: r = a >= b; break; case KML_X_LT: r = a < b; break;
    case KML_X_LE: r = a <= b; break; case KML_X_EQ: r = a == b; break; default: r = a != b; break;
    }
    st[sp - 1] = r;
  }
  return st[0];
}
struct ExprSet { kml_expr e[3]; };
// components of a particle vector field = expressions of (x, y, z, x0, y0, z0) on the particles of a group
__global__ void k_set_particles_expr(SolidDev s, int groupbit, double *f0, double *f1, double *f2, int set_mask, const ExprSet *prog, int current_is_xn) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= s.np || !(s.mask[ip] & groupbit)) return;
  double pv[6];
  for (int d = 0; d < 3; d++) { pv[d] = current_is_xn ? s.xn[d][ip] : s.x[d][ip]; pv[3 + d] = s.x0[d][ip]; }
  double *f[3] = {f0, f1, f2};
  for (int d = 0; d < 3; d++) if (set_mask & (1 << d)) f[d][ip] = expr_eval(prog->e[d], pv);
}

} // namespace kml
