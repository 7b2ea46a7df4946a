// kml_cpdi.cuh - convected particle domain interpolation (ulcpdi / tlcpdi, 2-D, styles R4 and Q4).
//
// Reference: ULCPDI / TLCPDI::compute_grid_weight_functions_and_gradients (src/ulcpdi.cpp:111-393, src/tlcpdi.cpp:98-351),
// Solid::compute_particle_velocities_and_positions + compute_particle_acceleration (src/solid.cpp:696-784),
// Solid::update_particle_domain (src/solid.cpp:2338-2352), the Q4 volume of Solid::update_deformation_gradient
// (src/solid.cpp:1188-1201).
//
// A CPDI particle does not see a tensor-product stencil: its nodes are the union of the stencils of the four corners
// of its domain, with weights built from the corners' shape functions.  The kernels therefore work on an explicit
// per-particle list (node, wf, wfd, and for Q4 the four corner weights), stored entry-major ([entry][particle]) so
// that consecutive threads read consecutive addresses.  UL rebuilds the list every step; TL builds it once from the
// reference configuration and reuses it (north_star: "TLMPM/CPDI reuse reference-configuration weights cached once").
#pragma once
#include "kml_kernels.cuh"

namespace kml {

constexpr int CPDI_MAXN = 64; // 4 corners x at most 4 x 4 nodes each

struct CpdiDev {
  int style;            // 0 = R4, 1 = Q4
  int maxn; long long cap;
  int *n;               // [np] entries of each particle
  int *node;            // [maxn][cap]
  double *wf, *wfd[2];  // [maxn][cap]
  double *wfc[4];       // [maxn][cap] corner weights (Q4)
  double *rp[2][2], *rp0[2][2]; // R4 domain vectors r1, r2 (x, y)
  double *xpc[4][2], *xpc0[4][2]; // Q4 corner positions (x, y)
};

// stencil base of a corner along one axis, from domain->boxlo (both methods: src/tlcpdi.cpp:200-222, src/ulcpdi.cpp:223-244)
template <int SHAPE> __device__ __forceinline__ int cpdi_base(double xc, double lo, double ih, bool has_axis, int &m) {
  const double t = __dmul_rn(__dsub_rn(xc, lo), ih);
  if (SHAPE == KML_SHAPE_LINEAR) { m = 2; return (int)t; }
  if (SHAPE == KML_SHAPE_BERNSTEIN) { m = 3; int i0 = 2 * (int)t; if (has_axis && i0 >= 1 && (i0 % 2 != 0)) i0--; return i0; }
  m = 4; return (int)__dsub_rn(t, 1.0);
}

template <int SHAPE, bool TL>
__global__ void __launch_bounds__(64) k_cpdi_weights(SolidDev s, GridDev g, CpdiDev cp, double boxlo0, double boxlo1, unsigned *flags) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= s.np) return;
  const double px = TL ? s.x0[0][ip] : s.x[0][ip], py = TL ? s.x0[1][ip] : s.x[1][ip];
  const int nx = g.n[0], ny = g.n[1];
  const long long nnodes = (long long)nx * ny;
  const double ih = g.inv_cellsize;
  double xc[4][2];
  double r1[2] = {0, 0}, r2[2] = {0, 0};
  if (cp.style == 0) {
    r1[0] = cp.rp[0][0][ip]; r1[1] = cp.rp[0][1][ip]; r2[0] = cp.rp[1][0][ip]; r2[1] = cp.rp[1][1][ip];
#pragma unroll
    for (int d = 0; d < 2; d++) {
      const double p = d == 0 ? px : py;
      xc[0][d] = p - r1[d] - r2[d]; xc[1][d] = p + r1[d] - r2[d]; xc[2][d] = p + r1[d] + r2[d]; xc[3][d] = p - r1[d] + r2[d];
    }
  } else {
#pragma unroll
    for (int ic = 0; ic < 4; ic++) { xc[ic][0] = cp.xpc[ic][0][ip]; xc[ic][1] = cp.xpc[ic][1][ip]; }
  }
  // candidate nodes: union of the corners' stencils, kept sorted and unique (src/ulcpdi.cpp:283-286)
  int cand[CPDI_MAXN]; int nc_ = 0;
  for (int ic = 0; ic < 4; ic++) {
    int m, m2;
    const int i0 = cpdi_base<SHAPE>(xc[ic][0], boxlo0, ih, true, m), j0 = cpdi_base<SHAPE>(xc[ic][1], boxlo1, ih, true, m2);
    for (int i = i0; i < i0 + m; i++)
      for (int j = j0; j < j0 + m; j++) {
        const long long n = (long long)ny * i + j; // tag of a 2-D grid (nz = 1), src/ulcpdi.cpp:262
        if (n < 0 || n >= nnodes) continue;
        int pos = 0; while (pos < nc_ && cand[pos] < (int)n) pos++;
        if (pos < nc_ && cand[pos] == (int)n) continue;
        if (nc_ == CPDI_MAXN) { atomicOr(flags, 16u); continue; }
        for (int q = nc_; q > pos; q--) cand[q] = cand[q - 1];
        cand[pos] = (int)n; nc_++;
      }
  }
  const double vol = s.vol[ip];
  const double inv_Vp = 1.0 / vol;
  double a = 0, b = 0, alpha_over_Vp = 0, sixVp = 0;
  if (cp.style == 1) {
    a = (xc[3][0] - xc[0][0]) * (xc[1][1] - xc[2][1]) - (xc[1][0] - xc[2][0]) * (xc[3][1] - xc[0][1]);
    b = (xc[2][0] - xc[3][0]) * (xc[0][1] - xc[1][1]) - (xc[0][0] - xc[1][0]) * (xc[2][1] - xc[3][1]);
    alpha_over_Vp = 0.0417 * inv_Vp; sixVp = 6 * vol;
  }
  int cnt = 0;
  for (int q = 0; q < nc_; q++) {
    const int in = cand[q];
    const int i = in / ny, j = in - i * ny;
    const double xn0 = __dadd_rn(g.lo[0], __dmul_rn((double)(i + g.goff0), g.h)), xn1 = __dadd_rn(g.lo[1], __dmul_rn((double)j, g.h));
    const int nt0 = node_type<SHAPE>(i + g.goff0, g.gn0), nt1 = node_type<SHAPE>(j, ny);
    double wfc[4], wf = 0;
#pragma unroll
    for (int ic = 0; ic < 4; ic++) {
      double p0, p1, d_;
      Basis<SHAPE>::eval(__dmul_rn(__dsub_rn(xc[ic][0], xn0), ih), nt0, ih, p0, d_);
      Basis<SHAPE>::eval(__dmul_rn(__dsub_rn(xc[ic][1], xn1), ih), nt1, ih, p1, d_);
      wfc[ic] = p0 * p1;
      if (cp.style == 0 && wfc[ic] > 1.0e-12) wf += wfc[ic];
    }
    if (cp.style == 0) wf *= 0.25;
    else wf = alpha_over_Vp * ((sixVp - a - b) * wfc[0] + (sixVp - a + b) * wfc[1] + (sixVp + a + b) * wfc[2] + (sixVp + a - b) * wfc[3]);
    if (!(wf > 1.0e-12)) continue;
    double w0, w1;
    if (cp.style == 0) {
      w0 = ((wfc[0] - wfc[2]) * (r1[1] - r2[1]) + (wfc[1] - wfc[3]) * (r1[1] + r2[1])) * inv_Vp;
      w1 = ((wfc[0] - wfc[2]) * (r2[0] - r1[0]) - (wfc[1] - wfc[3]) * (r1[0] + r2[0])) * inv_Vp;
    } else {
      w0 = (wfc[0] * (xc[1][1] - xc[3][1]) + wfc[1] * (xc[2][1] - xc[0][1]) + wfc[2] * (xc[3][1] - xc[1][1]) + wfc[3] * (xc[0][1] - xc[2][1])) * (0.5 * inv_Vp);
      w1 = (wfc[0] * (xc[3][0] - xc[1][0]) + wfc[1] * (xc[0][0] - xc[2][0]) + wfc[2] * (xc[1][0] - xc[3][0]) + wfc[3] * (xc[2][0] - xc[0][0])) * (0.5 * inv_Vp);
    }
    if (cnt == cp.maxn) { atomicOr(flags, 16u); break; }
    const long long e = (long long)cnt * cp.cap + ip;
    cp.node[e] = in; cp.wf[e] = wf; cp.wfd[0][e] = w0; cp.wfd[1][e] = w1;
    if (cp.style == 1) {
#pragma unroll
      for (int ic = 0; ic < 4; ic++) cp.wfc[ic][e] = wfc[ic];
    }
    cnt++;
  }
  cp.n[ip] = cnt;
}

// P2G over the lists: mass, momentum, internal force (UL: vol sigma grad w; TL: vol0PK1 grad0 w), body force.
template <bool TL>
__global__ void __launch_bounds__(128) k_cpdi_p2g(SolidDev s, GridDev g, CpdiDev cp, int what) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= s.np) return;
  const double m = s.mass[ip];
  double mv[2] = {0, 0}, A[4] = {0, 0, 0, 0}, mbp[2] = {0, 0};
  if (what & P2G_MOM) { mv[0] = s.v[0][ip]; mv[1] = s.v[1][ip]; }
  if (what & P2G_FORCE) {
    if (TL) { A[0] = s.pk1[0][ip]; A[1] = s.pk1[1][ip]; A[2] = s.pk1[3][ip]; A[3] = s.pk1[4][ip]; }
    else { const double vol = s.vol[ip]; A[0] = vol * s.sig[0][ip]; A[1] = vol * s.sig[3][ip]; A[2] = A[1]; A[3] = vol * s.sig[1][ip]; }
  }
  if (what & P2G_MB) { mbp[0] = s.mbp[0][ip]; mbp[1] = s.mbp[1][ip]; }
  const int n = cp.n[ip];
  for (int j = 0; j < n; j++) {
    const long long e = (long long)j * cp.cap + ip;
    const int node = cp.node[e]; const double wf = cp.wf[e];
    if (what & P2G_MASS) atomicAdd(&g.nv[node].w, wf * m);
    if (what & P2G_MOM) { const double wm = wf * m; atomicAdd(&g.nv[node].x, wm * mv[0]); atomicAdd(&g.nv[node].y, wm * mv[1]); }
    if (what & P2G_FORCE) {
      const double w0 = cp.wfd[0][e], w1 = cp.wfd[1][e];
      atomicAdd(&g.f[0][node], -(A[0] * w0 + A[1] * w1)); atomicAdd(&g.f[1][node], -(A[2] * w0 + A[3] * w1));
    }
    if (what & P2G_MB) { atomicAdd(&g.mb[0][node], wf * mbp[0]); atomicAdd(&g.mb[1][node], wf * mbp[1]); }
  }
}

// G2P: v~_p, x_p += dt w v~_I node by node, Q4 corners, a_p = sum w (v~_I - v_I) / dt, then the PIC/FLIP blend
// (src/solid.cpp:696-784, :786-796).  UL writes the advanced position to xn like every other G2P of the engine.
template <bool TL>
__global__ void __launch_bounds__(128) k_cpdi_g2p(SolidDev s, GridDev g, CpdiDev cp, StepParams sp) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= s.np) return;
  double vu[2] = {0, 0}, a[2] = {0, 0}, x[2] = {s.x[0][ip], s.x[1][ip]}, vc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  const int n = cp.n[ip];
  for (int j = 0; j < n; j++) {
    const long long e = (long long)j * cp.cap + ip;
    const int node = cp.node[e]; const double wf = cp.wf[e];
    const double4 ru = ldg4(&g.nvu[node]), rv = ldg4(&g.nv[node]);
    vu[0] += wf * ru.x; vu[1] += wf * ru.y;
    x[0] += sp.dt * wf * ru.x; x[1] += sp.dt * wf * ru.y;
    a[0] += wf * (ru.x - rv.x); a[1] += wf * (ru.y - rv.y);
    if (cp.style == 1) {
#pragma unroll
      for (int ic = 0; ic < 4; ic++) { const double w = cp.wfc[ic][e]; vc[ic][0] += w * ru.x; vc[ic][1] += w * ru.y; }
    }
  }
  if (cp.style == 1) {
#pragma unroll
    for (int ic = 0; ic < 4; ic++) { cp.xpc[ic][0][ip] += sp.dt * vc[ic][0]; cp.xpc[ic][1][ip] += sp.dt * vc[ic][1]; }
  }
  const double inv_dt = 1.0 / sp.dt;
#pragma unroll
  for (int d = 0; d < 2; d++) {
    const double ad = a[d] * inv_dt;
    if (s.acc[0]) { s.acc[d][ip] = ad; s.vup[d][ip] = vu[d]; }
    s.v[d][ip] = (1 - sp.alpha) * vu[d] + sp.alpha * (s.v[d][ip] + sp.dt * ad);
    if (TL) s.x[d][ip] = x[d]; else s.xn[d][ip] = x[d];
  }
  if (!TL) {
    s.xn[2][ip] = s.x[2][ip];
    const bool in = x[0] >= sp.boxlo[0] && x[0] <= sp.boxhi[0] && x[1] >= sp.boxlo[1] && x[1] <= sp.boxhi[1];
    if (!in) atomicOr(sp.flags, 1u);
  }
}

// velocity gradient over the lists + F + stress (particle_stress), Q4 volume from the corner polygon, R4 domain update
template <bool TL>
__global__ void __launch_bounds__(128) k_cpdi_stress(SolidDev s, GridDev g, CpdiDev cp, StepParams sp, StressParams tp, kml_material mat) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double wave = 0, hr = 1.0;
  if (ip < s.np) {
    const double4 *__restrict__ gv = tp.doublemapping ? g.nv : g.nvu;
    double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int n = cp.n[ip];
    for (int j = 0; j < n; j++) {
      const long long e = (long long)j * cp.cap + ip;
      const double4 rec = ldg4(&gv[cp.node[e]]);
      const double w0 = cp.wfd[0][e], w1 = cp.wfd[1][e];
      L[0] += rec.x * w0; L[1] += rec.x * w1; L[3] += rec.y * w0; L[4] += rec.y * w1;
    }
    double vol_cpdi = -1.0;
    if (cp.style == 1) { // area of the corner polygon, src/solid.cpp:1190-1197 (corners already advanced by G2P)
      double q[4][2];
#pragma unroll
      for (int ic = 0; ic < 4; ic++) { q[ic][0] = cp.xpc[ic][0][ip]; q[ic][1] = cp.xpc[ic][1][ip]; }
      vol_cpdi = 0.5 * (q[0][0] * q[1][1] - q[1][0] * q[0][1] + q[1][0] * q[2][1] - q[2][0] * q[1][1] + q[2][0] * q[3][1] - q[3][0] * q[2][1] +
                        q[3][0] * q[0][1] - q[0][0] * q[3][1]);
    }
    PState ps; ps.load(s, mat, sp, ip);
    const double qv[3] = {0, 0, 0};
    particle_stress<TL>(s, g, sp, mat, ip, ps, L, qv, wave, hr, vol_cpdi);
    if (!TL && cp.style == 0) { // Solid::update_particle_domain: r = F r0 (ULCPDI only, src/ulcpdi.cpp:486-490)
      const double F00 = s.F[0][ip], F01 = s.F[1][ip], F10 = s.F[3][ip], F11 = s.F[4][ip];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const double r0x = cp.rp0[k][0][ip], r0y = cp.rp0[k][1][ip];
        cp.rp[k][0][ip] = F00 * r0x + F01 * r0y; cp.rp[k][1][ip] = F10 * r0x + F11 * r0y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { wave = fmax(wave, __shfl_xor_sync(0xffffffffu, wave, o)); if (TL) hr = fmin(hr, __shfl_xor_sync(0xffffffffu, hr, o)); }
  if ((threadIdx.x & 31) == 0) { if (wave > 0) atomic_max_pos(tp.max_wave, wave); if (TL && hr < 1.0) atomic_min_pos(tp.min_h_ratio, hr < 0 ? 0.0 : hr); }
}

} // namespace kml
