// kml_kernels.cuh - the per-stage kernels of the MPM step (baseline, all dimensions /
// shape functions / UL + TL).  Weights are recomputed in registers from the step-start
// positions by every consumer (no stored neighbour lists; the reference rebuilds
// vector<vector<>> lists every step, src/ulmpm.cpp:88-337); TL uses the reference
// configuration x0 (src/tlmpm.cpp:87-340), which is equivalent to its one-time cache.
#pragma once
#include "kml_device.cuh"

namespace kml {

struct SolidDev {
  long long np;
  double *x[3], *xn[3], *x0[3], *v[3], *mbp[3], *q[3];
  double *sig[6], *eel[6], *F[9];
  double *vol0, *vol, *rho0, *mass, *eps, *epsdot, *dmg, *dmgi, *ien, *T, *gamma;
  double *pk1[9], *R[9]; // TL only
  double *Lst[9];        // APIC family only: the velocity gradient of the previous step (UL: L, TL: Fdot), src/solid.cpp:392-426
  long long *ptag; int *mask;
  double *acc[3], *vup[3]; // only after kml_keep_particle_acceleration: a_p and v_update_p of the last grid_to_points (src/solid.cpp:576-635)
};

struct StepParams {
  double dt, alpha;        // alpha = PIC_FLIP
  double boxlo[3], boxhi[3];
  int axisymmetric, temp;
  double inv_tav;          // stress update: signal_velocity / (1000 cellsize) of the solid being updated
  // APIC family (Update::SubMethodType APIC / MLS / AFLIP / ASFLIP): affine momentum transfer with the diagonal inertia
  // tensor Di of Solid::compute_inertia_tensor (src/solid.cpp:1440-1478)
  int apic, mls, asflip;
  // ext fills the padding after asflip: growing this struct by even 8 bytes makes ptxas spill in k_g2p_cell (128 registers), +7 % on that stage.
  // bit 0: gradient-enhanced momentum projection v_p + L_p (x_I - x_p) (Method::ge, src/solid.cpp:369-371).
  // bits 1-2: rigid bodies (material(..., rigid), src/material.h:49): 0 = no rigid solid in this run, 1 = there are rigid solids and the
  // solid being processed is deformable, 2 = the solid being processed is rigid.  Nodes inside the stencil of a rigid particle carry
  // Grid::rigid (src/ulmpm.cpp:267-268, src/tlmpm.cpp:283); the flag is never cleared (src/grid.cpp:248).
  int ext;
  __host__ __device__ bool ge() const { return (ext & 1) != 0; }
  __host__ __device__ int rigid_mode() const { return ext >> 1; }
  double Di[3];
  unsigned *flags;         // device error word
};

enum { P2G_MASS = 1, P2G_MOM = 2, P2G_FORCE = 4, P2G_MB = 8, P2G_TEMP = 16, P2G_HEAT = 32,
       P2G_POSMOVED = 64 /* UL: explicit particle positions are the ones advanced by G2P (MUSL re-projection) */,
       P2G_MARK_RIGID = 128 /* only set Grid::rigid on the nodes this (rigid) solid's particles reach (wf != 0) */ };

// symmetric index helper: (xx,yy,zz,xy,xz,yz)
__device__ __forceinline__ void load_sym(double *const *a, long long i, double *m) {
  double xx = a[0][i], yy = a[1][i], zz = a[2][i], xy = a[3][i], xz = a[4][i], yz = a[5][i];
  m[0] = xx; m[1] = xy; m[2] = xz; m[3] = xy; m[4] = yy; m[5] = yz; m[6] = xz; m[7] = yz; m[8] = zz;
}
__device__ __forceinline__ void store_sym(double *const *a, long long i, const double *m) {
  a[0][i] = m[0]; a[1][i] = m[4]; a[2][i] = m[8]; a[3][i] = m[1]; a[4][i] = m[2]; a[5][i] = m[5];
}

template <int DIM, int SHAPE, bool TL> struct Stencil {
  static constexpr int SPAN = StencilSpan<SHAPE, TL>::value;
  int i0[3];
  double w[3][SPAN], dw[3][SPAN];
  __device__ __forceinline__ void build(const GridDev &g, double px, double py, double pz) {
    axis_weights<SHAPE, TL, SPAN>(px, g.lo[0], g.h, g.inv_cellsize, g.n[0], g.goff0, g.gn0, i0[0], w[0], dw[0]);
    if (DIM >= 2) axis_weights<SHAPE, TL, SPAN>(py, g.lo[1], g.h, g.inv_cellsize, g.n[1], 0, g.n[1], i0[1], w[1], dw[1]);
    else { i0[1] = 0; w[1][0] = 1; dw[1][0] = 0; }
    if (DIM == 3) axis_weights<SHAPE, TL, SPAN>(pz, g.lo[2], g.h, g.inv_cellsize, g.n[2], 0, g.n[2], i0[2], w[2], dw[2]);
    else { i0[2] = 0; w[2][0] = 1; dw[2][0] = 0; }
  }
};

// iterate the stencil in the reference's (i,j,k) order; body(node, wf, wfd0, wfd1, wfd2)
#define KML_FOR_STENCIL(st, g, ...)                                                                    \
  _Pragma("unroll") for (int sa_ = 0; sa_ < decltype(st)::SPAN; sa_++) {                               \
    const double wx = st.w[0][sa_], dwx = st.dw[0][sa_];                                               \
    if (wx == 0.0) continue;                                                                           \
    const long long ni = (long long)(st.i0[0] + sa_) * g.n[1];                                         \
    _Pragma("unroll") for (int sb_ = 0; sb_ < (DIM >= 2 ? decltype(st)::SPAN : 1); sb_++) {            \
      const double wy = st.w[1][sb_], dwy = st.dw[1][sb_];                                             \
      if (wy == 0.0) continue;                                                                         \
      const long long nij = (ni + (DIM >= 2 ? st.i0[1] + sb_ : 0)) * g.n[2];                           \
      _Pragma("unroll") for (int sc_ = 0; sc_ < (DIM == 3 ? decltype(st)::SPAN : 1); sc_++) {          \
        const double wz = st.w[2][sc_], dwz = st.dw[2][sc_];                                           \
        if (wz == 0.0) continue;                                                                       \
        const long long node = nij + (DIM == 3 ? st.i0[2] + sc_ : 0);                                  \
        const double wf = (DIM == 1) ? wx : ((DIM == 2) ? wx * wy : wx * wy * wz);                     \
        const double wfd0 = (DIM == 1) ? dwx : ((DIM == 2) ? dwx * wy : dwx * wy * wz);                \
        const double wfd1 = (DIM == 1) ? 0.0 : ((DIM == 2) ? wx * dwy : wx * dwy * wz);                \
        const double wfd2 = (DIM == 3) ? wx * wy * dwz : 0.0;                                          \
        __VA_ARGS__                                                                                    \
      }                                                                                                \
    }                                                                                                  \
  }

// ---- P2G scatter (baseline: one thread per particle, fp64 RED atomics) ----------------------
// Solid::compute_mass_nodes src/solid.cpp:317-335, compute_velocity_nodes :337-390 (momentum; the
// division by the node mass happens in the grid kernel), compute_external_and_internal_forces_nodes_UL
// :482-522, compute_external_forces_nodes :428-450, compute_internal_forces_nodes_TL :452-480,
// thermal P2G src/solid.cpp:2743-2796.
template <int DIM, int SHAPE, bool TL>
__global__ void __launch_bounds__(128) k_p2g(SolidDev s, GridDev g, StepParams sp, int what) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= s.np) return;
  const double px = TL ? s.x0[0][ip] : s.x[0][ip], py = TL ? s.x0[1][ip] : s.x[1][ip], pz = TL ? s.x0[2][ip] : s.x[2][ip];
  Stencil<DIM, SHAPE, TL> st; st.build(g, px, py, pz);
  if (what & P2G_MARK_RIGID) {
    KML_FOR_STENCIL(st, g, { g.rigid[node] = 1; (void)wf; (void)wfd0; (void)wfd1; (void)wfd2; })
    return;
  }
  const double m = s.mass[ip];
  double mv[3] = {0, 0, 0}, A[9], mbp[3] = {0, 0, 0}, hoop = 0, mT = 0, gam = 0, qv[3] = {0, 0, 0};
  if (what & P2G_MOM) { mv[0] = s.v[0][ip]; mv[1] = s.v[1][ip]; mv[2] = s.v[2][ip]; }
  if (what & P2G_FORCE) {
    if (TL) {
#pragma unroll
      for (int i = 0; i < 9; i++) A[i] = s.pk1[i][ip];
      if (sp.axisymmetric) hoop = A[8] / s.x0[0][ip];
    } else {
      load_sym(s.sig, ip, A);
      const double vol = s.vol[ip];
      if (sp.axisymmetric) hoop = vol * (A[8] / s.x[0][ip]);
#pragma unroll
      for (int i = 0; i < 9; i++) A[i] *= vol;
    }
  }
  if (what & P2G_MB) { mbp[0] = s.mbp[0][ip]; mbp[1] = s.mbp[1][ip]; mbp[2] = s.mbp[2][ip]; }
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, pp[3] = {px, py, pz};
  const bool affine = sp.apic || sp.ge();
  if (affine) { // compute_velocity_nodes_APIC / compute_external_and_internal_forces_nodes_UL_MLS use the particle position explicitly
    if ((what & P2G_MOM) && s.Lst[0]) {
#pragma unroll
      for (int i = 0; i < 9; i++) C[i] = s.Lst[i][ip];
    }
    if (!TL && (what & P2G_POSMOVED)) { pp[0] = s.xn[0][ip]; pp[1] = s.xn[1][ip]; pp[2] = s.xn[2][ip]; }
  }
  if (what & P2G_TEMP) mT = m * s.T[ip];
  if (what & P2G_HEAT) { gam = s.gamma[ip]; qv[0] = s.q[0][ip]; qv[1] = s.q[1][ip]; qv[2] = s.q[2][ip]; }

  KML_FOR_STENCIL(st, g, {
    // rigid nodes take mass and momentum from rigid solids only (src/solid.cpp:326,355,412), no body force (:438,:493-:512)
    // and, TL, no internal force (:460)
    const bool nrigid = sp.rigid_mode() != 0 && g.rigid[node] != 0;
    const bool take_mv = !(nrigid && sp.rigid_mode() != 2);
    if ((what & P2G_MASS) && take_mv) atomicAdd(&g.nv[node].w, wf * m);
    double dxn[3] = {0, 0, 0}; // x_I - x_p (APIC family)
    if (affine) {
      dxn[0] = __dadd_rn(g.lo[0], __dmul_rn((double)(st.i0[0] + sa_ + g.goff0), g.h)) - pp[0];
      if (DIM >= 2) dxn[1] = __dadd_rn(g.lo[1], __dmul_rn((double)(st.i0[1] + sb_), g.h)) - pp[1];
      if (DIM == 3) dxn[2] = __dadd_rn(g.lo[2], __dmul_rn((double)(st.i0[2] + sc_), g.h)) - pp[2];
    }
    if ((what & P2G_MOM) && take_mv) {
      const double wm = wf * m;
      if (affine) { // src/solid.cpp:392-426 (APIC family) and :369-371 (gradient-enhanced): (w m) (v + C (x_I - x_p))
        atomicAdd(&g.nv[node].x, wm * (mv[0] + (C[0] * dxn[0] + C[1] * dxn[1] + C[2] * dxn[2])));
        if (DIM >= 2) atomicAdd(&g.nv[node].y, wm * (mv[1] + (C[3] * dxn[0] + C[4] * dxn[1] + C[5] * dxn[2])));
        if (DIM == 3) atomicAdd(&g.nv[node].z, wm * (mv[2] + (C[6] * dxn[0] + C[7] * dxn[1] + C[8] * dxn[2])));
      } else {
        atomicAdd(&g.nv[node].x, wm * mv[0]);
        if (DIM >= 2) atomicAdd(&g.nv[node].y, wm * mv[1]);
        if (DIM == 3) atomicAdd(&g.nv[node].z, wm * mv[2]);
      }
    }
    if ((what & P2G_FORCE) && sp.mls && !TL) { // src/solid.cpp:524-574: f_I -= vol w (sigma Di (x_I - x_p))
      const double e0 = sp.Di[0] * dxn[0], e1 = sp.Di[1] * dxn[1], e2 = sp.Di[2] * dxn[2];
      atomicAdd(&g.f[0][node], -(wf * (A[0] * e0 + A[1] * e1 + A[2] * e2)));
      if (DIM >= 2) atomicAdd(&g.f[1][node], -(wf * (A[3] * e0 + A[4] * e1 + A[5] * e2)));
      if (DIM == 3) atomicAdd(&g.f[2][node], -(wf * (A[6] * e0 + A[7] * e1 + A[8] * e2)));
    } else if ((what & P2G_FORCE) && !(TL && nrigid)) {
      double f0 = -(A[0] * wfd0 + A[1] * wfd1 + A[2] * wfd2);
      if (sp.axisymmetric) f0 -= hoop * wf;
      atomicAdd(&g.f[0][node], f0);
      if (DIM >= 2) atomicAdd(&g.f[1][node], -(A[3] * wfd0 + A[4] * wfd1 + A[5] * wfd2));
      if (DIM == 3) atomicAdd(&g.f[2][node], -(A[6] * wfd0 + A[7] * wfd1 + A[8] * wfd2));
    }
    if ((what & P2G_MB) && !nrigid) {
      atomicAdd(&g.mb[0][node], wf * mbp[0]);
      if (DIM >= 2) atomicAdd(&g.mb[1][node], wf * mbp[1]);
      if (DIM == 3) atomicAdd(&g.mb[2][node], wf * mbp[2]);
    }
    if (what & P2G_TEMP) atomicAdd(&g.T[node], wf * mT);
    if (what & P2G_HEAT) {
      atomicAdd(&g.Qext[node], wf * gam);
      atomicAdd(&g.Qint[node], wfd0 * qv[0] + wfd1 * qv[1] + wfd2 * qv[2]);
    }
  })
}

#define KML_NVD_PADK 102 /* kml_gather_cell3.cuh NVD_PADK */
#ifdef KML_MISC_KERNELS  // non-template kernels: compiled once, in kml.cu
// ---- grid kernels ----------------------------------------------------------------------------
// normalise momentum -> velocity (the "/ grid->mass[in]" of src/solid.cpp:378, :2761) and
// Grid::update_grid_velocities / update_grid_temperature (src/grid.cpp:448-466, :1354-1362)
// nvd (optional): packed gather records {v_update, v_update - v} on the zero-padded grid of kml_gather_cell3.cuh, written in the same pass
__global__ void k_grid_update(GridDev g, double dt, int normalize, int update, int temp, int normalize_T, int rigid_aware, double *nvd) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nn) return;
  double4 rec = g.nv[i];
  const double m = rec.w;
  const bool free_node = !(rigid_aware && g.rigid[i] != 0); // rigid nodes keep v_update = v (src/grid.cpp:455-461)
  double v[3] = {rec.x, rec.y, rec.z};
  if (normalize) {
#pragma unroll
    for (int d = 0; d < 3; d++) v[d] = (m > 0) ? v[d] / m : 0.0;
    rec.x = v[0]; rec.y = v[1]; rec.z = v[2];
    g.nv[i] = rec;
  }
  double T = 0;
  if (temp) { T = g.T[i]; if (normalize_T) { T = (m > 0) ? T / m : 0.0; g.T[i] = T; } }
  if (update) {
    double4 u;
    u.x = (m != 0 && free_node) ? v[0] + dt * (g.f[0][i] + g.mb[0][i]) / m : v[0];
    u.y = (m != 0 && free_node) ? v[1] + dt * (g.f[1][i] + g.mb[1][i]) / m : v[1];
    u.z = (m != 0 && free_node) ? v[2] + dt * (g.f[2][i] + g.mb[2][i]) / m : v[2];
    u.w = temp ? ((m != 0) ? T + dt * (g.Qint[i] + g.Qext[i]) / m : T) : 0.0;
    g.nvu[i] = u;
    if (nvd) {
      const int k = (int)(i % g.n[2]); const long long t = i / g.n[2]; const int j = (int)(t % g.n[1]), ii = (int)(t / g.n[1]);
      double *d = nvd + (((long long)ii * (g.n[1] + 3) + j) * (g.n[2] + KML_NVD_PADK) + k) * 6;
      *(double2 *)d = make_double2(u.x, u.y); *(double2 *)(d + 2) = make_double2(u.z, u.x - v[0]); *(double2 *)(d + 4) = make_double2(u.y - v[1], u.z - v[2]);
    }
  }
}

// Grid::update_grid_positions, src/grid.cpp:468-474 (TL)
__global__ void k_grid_positions(GridDev g, double dt) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nn) return;
  const double4 rec = g.nv[i];
  g.x[0][i] += dt * rec.x; g.x[1][i] += dt * rec.y; g.x[2][i] += dt * rec.z;
}

// zero the velocity (momentum) part of the node records, keeping the mass (MUSL re-projection, TL passes)
__global__ void k_grid_zero_v(GridDev g, int zero_mass) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nn) return;
  // write-only: the mass (component w) is left alone instead of being read and written back
  double *r = (double *)&g.nv[i];
  *(double2 *)r = make_double2(0.0, 0.0);
  if (zero_mass) *(double2 *)(r + 2) = make_double2(0.0, 0.0); else r[2] = 0.0;
}

#endif // KML_MISC_KERNELS

// tail of G2P for one particle: a_p, x_p += dt v~_p, FLIP/PIC blend (src/solid.cpp:613-616, :786-796)
template <bool TL, bool KEEP = true>
__device__ __forceinline__ void particle_advance(const SolidDev &s, const StepParams &sp, long long ip, const double *vu, const double *a, double Tp,
                                                 const double *vold = nullptr) { // vold: the particle's velocity if the caller already loaded it
  const double inv_dt = 1.0 / sp.dt;
  double xnew[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double ad = a[d] * inv_dt;
    if (KEEP) { if (s.acc[0]) { s.acc[d][ip] = ad; s.vup[d][ip] = vu[d]; } }
    const double xo = s.x[d][ip];
    const double vnew = (1 - sp.alpha) * vu[d] + sp.alpha * ((vold ? vold[d] : s.v[d][ip]) + sp.dt * ad);
    // ASFLIP (UL): the position is advanced with the blended particle velocity, not with v~ (src/solid.cpp:637-694, :786-796)
    xnew[d] = (!TL && sp.asflip) ? xo + sp.dt * vnew : xo + sp.dt * vu[d];
    s.v[d][ip] = vnew;
    if (TL) s.x[d][ip] = xnew[d]; else s.xn[d][ip] = xnew[d];
  }
  if (sp.temp) s.T[ip] = Tp;
  if (!TL) { // Domain::inside, src/domain.cpp:176-185 -> error flag instead of abort (src/solid.cpp:617-627)
    bool in = xnew[0] >= sp.boxlo[0] && xnew[0] <= sp.boxhi[0] && xnew[1] >= sp.boxlo[1] && xnew[1] <= sp.boxhi[1] && xnew[2] >= sp.boxlo[2] && xnew[2] <= sp.boxhi[2];
    if (!in) atomicOr(sp.flags, 1u);
  }
}

// ---- G2P + advance ---------------------------------------------------------------------------
// Solid::compute_particle_accelerations_velocities_and_positions src/solid.cpp:576-635 fused with
// Solid::update_particle_velocities src/solid.cpp:786-796 and update_particle_temperature :2798-2808.
template <int DIM, int SHAPE, bool TL>
__global__ void __launch_bounds__(128) k_g2p(SolidDev s, GridDev g, StepParams sp) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= s.np) return;
  const double px = TL ? s.x0[0][ip] : s.x[0][ip], py = TL ? s.x0[1][ip] : s.x[1][ip], pz = TL ? s.x0[2][ip] : s.x[2][ip];
  Stencil<DIM, SHAPE, TL> st; st.build(g, px, py, pz);
  double vu[3] = {0, 0, 0}, a[3] = {0, 0, 0}, Tp = 0;
  KML_FOR_STENCIL(st, g, {
    const double4 ru = ldg4(&g.nvu[node]);
    const double4 rv = ldg4(&g.nv[node]);
    vu[0] += wf * ru.x; a[0] += wf * (ru.x - rv.x);
    if (DIM >= 2) { vu[1] += wf * ru.y; a[1] += wf * (ru.y - rv.y); }
    if (DIM == 3) { vu[2] += wf * ru.z; a[2] += wf * (ru.z - rv.z); }
    if (sp.temp) Tp += wf * ru.w;
    (void)wfd0; (void)wfd1; (void)wfd2;
  })
  if (sp.rigid_mode() == 2) a[0] = a[1] = a[2] = 0.0; // rigid particles: Solid::compute_particle_acceleration leaves a = 0 (src/solid.cpp:767-784)
  particle_advance<TL>(s, sp, ip, vu, a, Tp);
}

// ---- velocity gradient + deformation gradient + stress, fused ---------------------------------
// Solid::compute_rate_deformation_gradient_UL/_TL src/solid.cpp:798-936, update_deformation_gradient
// :1155-1244, update_stress :1246-1438 (+ update_heat_flux :2810-2839).  One thread per particle;
// L, D, Finv, J, rho never touch HBM.
struct StressParams {
  int doublemapping;    // gather from grid v (1) or v_update (0)
  int moved;            // axisymmetric hoop term uses the moved position (MUSL) or the step-start one (USL/USF)
  double *max_wave;     // device scalar: max_p (c_p + |v|_inf)
  double *min_h_ratio;  // device scalar (TL)
};

// Everything after the velocity-gradient gather for one particle: F update, J, vol, D (R for TL), the
// constitutive update and the particle's CFL wave speed (src/solid.cpp:1155-1438, :2810-2839).
// L is the gathered velocity gradient (Fdot for TL); qv the gathered -grad(T).
// The particle state the stress update reads; loading it is separate from the update so that callers can put the
// loads in flight before the velocity-gradient gather.
struct PState {
  double F[9], sig[6], eel[6], vol0, rho0, eps, epsdot, dmg, dmgi, T, v[3];
  __device__ __forceinline__ void load(const SolidDev &s, const kml_material &mat, const StepParams &sp, long long ip) {
#pragma unroll
    for (int i = 0; i < 9; i++) F[i] = s.F[i][ip];
#pragma unroll
    for (int i = 0; i < 6; i++) { sig[i] = s.sig[i][ip]; eel[i] = s.eel[i][ip]; }
    vol0 = s.vol0[ip]; rho0 = s.rho0[ip]; dmg = s.dmg[ip];
#pragma unroll
    for (int i = 0; i < 3; i++) v[i] = s.v[i][ip];
    eps = epsdot = dmgi = T = 0.0;
    if (mat.type == KML_MAT_EOS_STRENGTH) {
      eps = s.eps[ip]; epsdot = s.epsdot[ip];
      if (mat.damage_type != KML_DAMAGE_NONE) dmgi = s.dmgi[ip];
      if (mat.cp != 0 || sp.temp) T = s.T[ip];
    }
  }
};
// cp.async (LDGSTS): global -> shared copies that bypass the register file and complete asynchronously
// (the "memory" clobbers matter: without them the compiler may sink ordinary shared-memory loads of a slot below the
//  cp.async that refills it)
__device__ __forceinline__ void cp_async8(double *smem, const double *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// PState staged through shared memory: slot f of thread t lives at buf[f * stride + t] (conflict-free LDS.64)
constexpr int PSTATE_SLOTS = 32;
__device__ __forceinline__ void pstate_issue_async(double *buf, int stride, const SolidDev &s, const kml_material &mat, const StepParams &sp, long long ip) {
  int f = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) cp_async8(buf + (f++) * stride, s.F[i] + ip);
#pragma unroll
  for (int i = 0; i < 6; i++) cp_async8(buf + (f++) * stride, s.sig[i] + ip);
#pragma unroll
  for (int i = 0; i < 6; i++) cp_async8(buf + (f++) * stride, s.eel[i] + ip);
  cp_async8(buf + (f++) * stride, s.vol0 + ip); cp_async8(buf + (f++) * stride, s.rho0 + ip); cp_async8(buf + (f++) * stride, s.dmg + ip);
#pragma unroll
  for (int i = 0; i < 3; i++) cp_async8(buf + (f++) * stride, s.v[i] + ip);
  if (mat.type == KML_MAT_EOS_STRENGTH) {
    cp_async8(buf + 27 * stride, s.eps + ip); cp_async8(buf + 28 * stride, s.epsdot + ip);
    if (mat.damage_type != KML_DAMAGE_NONE) cp_async8(buf + 29 * stride, s.dmgi + ip);
    if (mat.cp != 0 || sp.temp) cp_async8(buf + 30 * stride, s.T + ip);
  }
}
__device__ __forceinline__ void pstate_from_smem(PState &ps, const double *buf, int stride, const kml_material &mat, const StepParams &sp) {
  int f = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) ps.F[i] = buf[(f++) * stride];
#pragma unroll
  for (int i = 0; i < 6; i++) ps.sig[i] = buf[(f++) * stride];
#pragma unroll
  for (int i = 0; i < 6; i++) ps.eel[i] = buf[(f++) * stride];
  ps.vol0 = buf[(f++) * stride]; ps.rho0 = buf[(f++) * stride]; ps.dmg = buf[(f++) * stride];
#pragma unroll
  for (int i = 0; i < 3; i++) ps.v[i] = buf[(f++) * stride];
  ps.eps = ps.epsdot = ps.dmgi = ps.T = 0.0;
  if (mat.type == KML_MAT_EOS_STRENGTH) {
    ps.eps = buf[27 * stride]; ps.epsdot = buf[28 * stride];
    if (mat.damage_type != KML_DAMAGE_NONE) ps.dmgi = buf[29 * stride];
    if (mat.cp != 0 || sp.temp) ps.T = buf[30 * stride];
  }
}

__device__ __forceinline__ void sym_to_full(const double *a, double *m) { // (xx,yy,zz,xy,xz,yz) -> row-major 3x3
  m[0] = a[0]; m[1] = a[3]; m[2] = a[4]; m[3] = a[3]; m[4] = a[1]; m[5] = a[5]; m[6] = a[4]; m[7] = a[5]; m[8] = a[2];
}

template <bool TL>
__device__ __forceinline__ void particle_stress(const SolidDev &s, const GridDev &g, const StepParams &sp, const kml_material &mat, long long ip,
                                                const PState &ps, double *L, const double *qv, double &wave, double &hr,
                                                double vol_cpdi = -1.0) { // vol_cpdi >= 0: CPDI-Q4 volume (area of the corner polygon)
  const double dt = sp.dt;
  double F[9], Fn[9];
#pragma unroll
  for (int i = 0; i < 9; i++) F[i] = ps.F[i];
  if (TL) {
#pragma unroll
    for (int i = 0; i < 9; i++) Fn[i] = F[i] + dt * L[i]; // here L holds Fdot
  } else {
    double IL[9];
#pragma unroll
    for (int i = 0; i < 9; i++) IL[i] = dt * L[i];
    IL[0] += 1; IL[4] += 1; IL[8] += 1;
    mul3(IL, F, Fn);
  }
#pragma unroll
  for (int i = 0; i < 9; i++) s.F[i][ip] = Fn[i];
  double Finv[9]; double iJ = inv3(Fn, Finv);
  const double vol0 = ps.vol0;
  double J, vol;
  if (vol_cpdi >= 0.0) { vol = vol_cpdi; J = vol / vol0; iJ = 1.0 / J; } // src/solid.cpp:1188-1201
  else { J = det3(Fn); vol = J * vol0; }
  s.vol[ip] = vol;
  const double damage_old = ps.dmg;
  if (J <= 0.0 && damage_old < 1.0) atomicOr(sp.flags, 2u);
  const double rho = ps.rho0 * iJ; // 1/J from the inverse: one FP64 division less per particle
  double D[9], R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (mat.type != KML_MAT_NEO_HOOKEAN) {
    if (TL) {
      if (!poldec3(Fn, R)) atomicOr(sp.flags, 8u);
      double Lt[9], S[9], T1[9];
      mul3(L, Finv, Lt);                     // L = Fdot Finv
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) S[3 * a + b] = Lt[3 * a + b] + Lt[3 * b + a];
      mul3_at(R, S, T1); mul3(T1, R, D);     // R^T (L + L^T) R
#pragma unroll
      for (int i = 0; i < 9; i++) D[i] *= 0.5;
#pragma unroll
      for (int i = 0; i < 9; i++) s.R[i][ip] = R[i];
    } else {
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) D[3 * a + b] = 0.5 * (L[3 * a + b] + L[3 * b + a]);
    }
  }
  double sig[9], eel[9];
  sym_to_full(ps.sig, sig); sym_to_full(ps.eel, eel);
  double damage = damage_old;
  if (mat.type == KML_MAT_LINEAR) {
    const double tr = dt * D[0] + dt * D[4] + dt * D[8];
#pragma unroll
    for (int i = 0; i < 9; i++) { const double inc = dt * D[i]; eel[i] += inc; sig[i] += 2 * mat.G * inc; }
    const double lt = mat.lambda * tr;
    sig[0] += lt; sig[4] += lt; sig[8] += lt;
  } else if (mat.type == KML_MAT_NEO_HOOKEAN) {
    double PK1[9]; const double lJ = mat.lambda * log(J);
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int b = 0; b < 3; b++) PK1[3 * a + b] = mat.G * (Fn[3 * a + b] - Finv[3 * b + a]) + lJ * Finv[3 * b + a];
    if (TL) {
#pragma unroll
      for (int i = 0; i < 9; i++) s.pk1[i][ip] = vol0 * PK1[i];
    }
    double FP[9]; mul3_bt(Fn, PK1, FP);
#pragma unroll
    for (int i = 0; i < 9; i++) sig[i] = iJ * FP[i];
    double C[9]; mul3_at(Fn, Fn, C);
    C[0] -= 1; C[4] -= 1; C[8] -= 1;
#pragma unroll
    for (int i = 0; i < 9; i++) eel[i] = 0.5 * C[i];
  } else { // EOS + strength (+ damage, + plastic-work heating): src/solid.cpp:1292-1374
    const bool thermal = mat.cp != 0;
    const double T = thermal ? ps.T : 0.0;
    const double trD = D[0] + D[4] + D[8];
    double ien;
    double pH = eos_pressure(mat, ien, J, rho, damage, trD, g.cellsize, T);
    s.ien[ip] = ien;
    if (thermal) pH += mat.tmp_alpha * (mat.tmp_T0 - T);
    double sdev[9], dep;
    double eps = ps.eps, epsdot = ps.epsdot;
    strength_dev(mat, dt, sig, D, sdev, dep, eps, epsdot, damage, T);
    eps += dep;
    const double itav = sp.inv_tav; // 1 / (1000 cellsize / signal_velocity), src/solid.cpp:1323-1327, uniform per launch
    epsdot -= epsdot * dt * itav;
    epsdot += dep * itav;
    epsdot = (0.0 > epsdot) ? 0.0 : epsdot;
    s.eps[ip] = eps; s.epsdot[ip] = epsdot;
    if (mat.damage_type != KML_DAMAGE_NONE) {
      double di = ps.dmgi;
      damage_jc(mat, di, damage, pH, sdev, epsdot, dep, sp.temp ? ps.T : 0.0);
      s.dmgi[ip] = di; s.dmg[ip] = damage;
    }
    if (thermal) {
      const double flow = KML_SQRT_3_OVER_2 * frob3(sdev);
      double gam = (T < mat.tmp_Tm) ? mat.tmp_chi * flow * epsdot : 0.0;
      gam *= (TL ? vol0 : vol) * mat.invcp;
      s.gamma[ip] = gam;
    }
    const double pf = (damage == 0 || pH >= 0) ? -pH : -pH * (1.0 - damage);
    const double te = (dt * trD + (eel[0] + eel[4] + eel[8])) * (1.0 / 3.0);
    const double iGd = 1.0 / ((damage > 1e-10) ? mat.G * (1 - damage) : mat.G); // one reciprocal instead of nine FP64 divisions
#pragma unroll
    for (int i = 0; i < 9; i++) { sig[i] = sdev[i]; eel[i] = sdev[i] * iGd; }
    sig[0] += pf; sig[4] += pf; sig[8] += pf;
    eel[0] += te; eel[4] += te; eel[8] += te;
  }
  store_sym(s.sig, ip, sig); store_sym(s.eel, ip, eel);
  if (TL && mat.type != KML_MAT_NEO_HOOKEAN) { // vol0PK1 = vol0 J (R sigma R^T) F^-T
    double T1[9], T2[9], P[9];
    mul3(R, sig, T1); mul3_bt(T1, R, T2); mul3_bt(T2, Finv, P);
    const double c = vol0 * J;
#pragma unroll
    for (int i = 0; i < 9; i++) s.pk1[i][ip] = c * P[i];
  }
  if (sp.temp) {
    const double c = (TL ? vol0 : vol) * mat.invcp * mat.kappa;
#pragma unroll
    for (int b = 0; b < 3; b++) s.q[b][ip] = qv[b] * c;
  }
  if (!(damage >= 1.0)) { // wave speed for the CFL limit, src/solid.cpp:1378-1385
    const double vx = fabs(ps.v[0]), vy = fabs(ps.v[1]), vz = fabs(ps.v[2]);
    wave = sqrt((mat.K + KML_FOUR_THIRD * mat.G) / rho) + fmax(fmax(vx, vy), vz);
    if (isnan(wave)) { atomicOr(sp.flags, 4u); wave = 0; }
    if (TL) { double e; if (eig3_min_abs_real(Fn, e)) hr = fmin(hr, e); }
  }
}

template <int DIM, int SHAPE, bool TL>
__global__ void __launch_bounds__(128) k_stress(SolidDev s, GridDev g, StepParams sp, StressParams tp, kml_material mat) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double wave = 0, hr = 1.0;
  if (ip < s.np) {
    const double px = TL ? s.x0[0][ip] : s.x[0][ip], py = TL ? s.x0[1][ip] : s.x[1][ip], pz = TL ? s.x0[2][ip] : s.x[2][ip];
    Stencil<DIM, SHAPE, TL> st; st.build(g, px, py, pz);
    const double4 *__restrict__ gv = tp.doublemapping ? g.nv : g.nvu;
    double L[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, hoop = 0, qv[3] = {0, 0, 0};
    double pp[3] = {px, py, pz}; // explicit particle position of the APIC gradient (src/solid.cpp:1006-1153)
    if (sp.apic && !TL && tp.moved) { pp[0] = s.xn[0][ip]; pp[1] = s.xn[1][ip]; pp[2] = s.xn[2][ip]; }
    KML_FOR_STENCIL(st, g, {
      double wfd[3] = {wfd0, wfd1, wfd2};
      if (sp.apic) { // L += v_I (x) (x_I - x_p) w, scaled by Di afterwards
        wfd[0] = (__dadd_rn(g.lo[0], __dmul_rn((double)(st.i0[0] + sa_ + g.goff0), g.h)) - pp[0]) * wf;
        wfd[1] = DIM >= 2 ? (__dadd_rn(g.lo[1], __dmul_rn((double)(st.i0[1] + sb_), g.h)) - pp[1]) * wf : 0.0;
        wfd[2] = DIM == 3 ? (__dadd_rn(g.lo[2], __dmul_rn((double)(st.i0[2] + sc_), g.h)) - pp[2]) * wf : 0.0;
      }
      const double4 rec = ldg4(&gv[node]);
      const double vn[3] = {rec.x, rec.y, rec.z};
#pragma unroll
      for (int a = 0; a < DIM; a++) {
#pragma unroll
        for (int b = 0; b < DIM; b++) L[3 * a + b] += vn[a] * wfd[b];
      }
      if (DIM == 2 && sp.axisymmetric) hoop += vn[0] * wf;
      if (sp.temp) {
        const double Tn = tp.doublemapping ? g.T[node] : rec.w;
#pragma unroll
        for (int b = 0; b < 3; b++) qv[b] -= wfd[b] * Tn;
      }
    })
    if (DIM == 2 && sp.axisymmetric) {
      const double xr = TL ? s.x0[0][ip] : (tp.moved ? s.xn[0][ip] : s.x[0][ip]);
      L[8] += hoop / xr; // the reference divides term by term; same value up to rounding
    }
    if (sp.apic) { // L *= Di (src/solid.cpp:1077,1151), kept for the next step's affine momentum transfer
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) L[3 * a + b] *= sp.Di[b];
    }
    if (s.Lst[0]) { // APIC family and gradient-enhanced projection read it back in the next momentum pass
#pragma unroll
      for (int i = 0; i < 9; i++) s.Lst[i][ip] = L[i];
    }
    PState ps; ps.load(s, mat, sp, ip);
    particle_stress<TL>(s, g, sp, mat, ip, ps, L, qv, wave, hr);
  }
  // block reduction -> one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { wave = fmax(wave, __shfl_xor_sync(0xffffffffu, wave, o)); if (TL) hr = fmin(hr, __shfl_xor_sync(0xffffffffu, hr, o)); }
  if ((threadIdx.x & 31) == 0) { if (wave > 0) atomic_max_pos(tp.max_wave, wave); if (TL && hr < 1.0) atomic_min_pos(tp.min_h_ratio, hr < 0 ? 0.0 : hr); }
}

#ifdef KML_MISC_KERNELS
// ---- fixes -----------------------------------------------------------------------------------
// FixVelocityNodes, src/fix_velocity_nodes.cpp:130-268
__global__ void k_fix_velocity_nodes(GridDev g, int groupbit, int set_mask, double v0, double v1, double v2, double p0, double p1, double p2,
                                     int which, double inv_dt, double *ftot) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double f[3] = {0, 0, 0};
  if (i < g.nn && (g.mask[i] & groupbit)) {
    const double v[3] = {v0, v1, v2}, p[3] = {p0, p1, p2};
    double4 rv = g.nv[i];
    if (which == 0) {
      double4 ru = g.nvu[i];
      const double c = inv_dt * rv.w;
#pragma unroll
      for (int d = 0; d < 3; d++) if (set_mask & (1 << d)) { f[d] = c * (v[d] - comp(ru, d)); comp(ru, d) = v[d]; comp(rv, d) = p[d]; }
      g.nvu[i] = ru;
    } else {
#pragma unroll
      for (int d = 0; d < 3; d++) if (set_mask & (1 << d)) comp(rv, d) = v[d];
    }
    g.nv[i] = rv;
  }
  if (which == 0) {
    const int plane = (int)(min(i, g.nn - 1) / ((long long)g.n[1] * g.n[2]));
    const bool own = plane >= g.own_lo && plane < g.own_hi; // shared slab planes are counted by their owner only
#pragma unroll
    for (int d = 0; d < 3; d++) {
      double x = own ? f[d] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(&ftot[d], x);
    }
  }
}
// FixVelocityParticles, src/fix_velocity_particles.cpp:131-300 (particle-independent values).  xold: the positions at the start of the
// step - UL keeps them in x until the next weight evaluation (G2P writes xn), TL remembers them in the otherwise unused xn.
__global__ void k_fix_velocity_particles(SolidDev s, int groupbit, int set_mask, double v0, double v1, double v2, int which, int is_TL, double dt, double *ftot) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double f[3] = {0, 0, 0};
  if (ip < s.np && (s.mask[ip] & groupbit)) {
    const double v[3] = {v0, v1, v2};
    if (which == 0) {
#pragma unroll
      for (int d = 0; d < 3; d++) { if (is_TL) s.xn[d][ip] = s.x[d][ip]; if (set_mask & (1 << d)) s.v[d][ip] = v[d]; }
    } else {
      const double c = (1.0 / dt) * s.mass[ip];
#pragma unroll
      for (int d = 0; d < 3; d++) if (set_mask & (1 << d)) {
        f[d] = c * (v[d] - s.v[d][ip]); s.v[d][ip] = v[d];
        if (is_TL) s.x[d][ip] = s.xn[d][ip] + dt * v[d]; else s.xn[d][ip] = s.x[d][ip] + dt * v[d];
      }
    }
  }
  if (which == 1) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      double x = f[d];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(&ftot[d], x);
    }
  }
}
// FixTemperatureNodes, src/fix_temperature_nodes.cpp:74-146
__global__ void k_fix_temperature_nodes(GridDev g, int groupbit, double T, double Tprev, int which) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nn || !(g.mask[i] & groupbit)) return;
  if (which == 0) { g.nvu[i].w = T; g.T[i] = Tprev; } else g.T[i] = T;
}
// FixTemperatureParticles, src/fix_temperature_particles.cpp:92-181 (particle-independent value)
__global__ void k_fix_temperature_particles(SolidDev s, int groupbit, double T) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip < s.np && (s.mask[ip] & groupbit)) s.T[ip] = T;
}
// FixBodyforce, src/fix_body_force.cpp:106-180
__global__ void k_fix_body_force(GridDev g, int groupbit, int set_mask, double f0, double f1, double f2, double *ftot) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double f[3] = {0, 0, 0};
  if (i < g.nn) {
    const double m = g.nv[i].w;
    if (m > 0 && (g.mask[i] & groupbit)) {
      const double fv[3] = {f0, f1, f2};
#pragma unroll
      for (int d = 0; d < 3; d++) if (set_mask & (1 << d)) { f[d] = fv[d] * m; g.mb[d][i] += f[d]; }
    }
  }
  const int plane = (int)(min(i, g.nn - 1) / ((long long)g.n[1] * g.n[2]));
  const bool own = plane >= g.own_lo && plane < g.own_hi;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double x = own ? f[d] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(&ftot[d], x);
  }
}

// FixForceNodes, src/fix_force_nodes.cpp:96-193: count the massive nodes of the group, then share the force among them
__global__ void k_fix_force_count(GridDev g, int groupbit, int *count) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // a decomposed grid counts every node once: on the rank that owns its plane (the counts are then summed over the ranks)
  const int plane = (int)(min(i, g.nn - 1) / ((long long)g.n[1] * g.n[2]));
  const bool in = i < g.nn && plane >= g.own_lo && plane < g.own_hi && g.nv[i].w > 0 && (g.mask[i] & groupbit);
  const unsigned b = __ballot_sync(0xffffffffu, in);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}
__global__ void k_fix_force_apply(GridDev g, int groupbit, int set_mask, double f0, double f1, double f2, const int *count, double *ftot) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double f[3] = {0, 0, 0};
  if (i < g.nn && g.nv[i].w > 0 && (g.mask[i] & groupbit)) {
    const double n = (double)*count; const double fv[3] = {f0, f1, f2};
#pragma unroll
    for (int d = 0; d < 3; d++) if (set_mask & (1 << d)) { f[d] = fv[d] / n; g.mb[d][i] += f[d]; }
  }
  const int plane = (int)(min(i, g.nn - 1) / ((long long)g.n[1] * g.n[2]));
  const bool own = plane >= g.own_lo && plane < g.own_hi;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double x = own ? f[d] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(&ftot[d], x);
  }
}

// FixContactHertz / FixContactMinPenetration, src/fix_contact_hertz.cpp:84-201,
// src/fix_contact_min_penetration.cpp:88-258.  All-pairs like the reference, tiled through shared
// memory: block = 128 particles of solid 1, loops over solid 2 in tiles.  The three screens of the
// reference are applied in order so the pair set is identical.
struct ContactParams { int dim, hertz, axisymmetric, temp; double Estar, max_cellsize, mu, dt, alpha, invcp1, invcp2; };

// One particle pair of the contact fixes (src/fix_contact_hertz.cpp:105-192, src/fix_contact_min_penetration.cpp:112-252): the reference's three
// screens in its order, then the force.  Particle 1's share is accumulated in registers (f1, g1), particle 2's goes out as atomics.
struct ContactP1 { double x, y, z, vol, m, vx, vy, vz, Rp; long long i; };
__device__ __forceinline__ double contact_radius(const ContactParams &cp, double vol, double x) {
  if (cp.dim == 2) return 0.5 * sqrt(cp.axisymmetric && !cp.hertz ? vol / x : vol);
  return cp.hertz ? 0.5 * pow(vol, 0.333333333) : 0.5 * cbrt(vol);
}
__device__ __forceinline__ void contact_pair(const SolidDev &s1, const SolidDev &s2, const ContactParams &cp, const ContactP1 &p1, long long j, double x2, double y2, double z2,
                                             double vol2, double m2, double vx2, double vy2, double vz2, double (&f1)[3], double &g1, double (&ft)[3]) {
  const double mc = cp.max_cellsize;
  const double dx = x2 - p1.x, dy = y2 - p1.y, dz = z2 - p1.z;
  bool near = dx < mc && dy < mc && dx > -mc && dy > -mc;
  if (cp.dim == 3 || cp.hertz) near = near && dz < mc && dz > -mc;
  if (!near) return;
  const double Rp1 = p1.Rp, Rp2 = contact_radius(cp, vol2, x2);
  const double Rp = Rp1 + Rp2;
  bool near2 = dx < Rp && dy < Rp && dx > -Rp && dy > -Rp;
  if (cp.dim == 3 || cp.hertz) near2 = near2 && dz < Rp && dz > -Rp;
  if (!near2) return;
  const double r = sqrt(dx * dx + dy * dy + dz * dz);
  if (!(r < Rp)) return;
  double f[3], g2 = 0;
  if (cp.hertz) {
    const double p = Rp - r;
    const double fmag = (cp.dim == 2 ? 0.25 * M_PI : 1.333333333) * cp.Estar * sqrt(Rp1 * Rp2 / (Rp1 + Rp2) * p * p * p);
    f[0] = fmag * dx / r; f[1] = fmag * dy / r; f[2] = fmag * dz / r;
    // s1 -= f ; s2 += f
    f1[0] -= f[0]; f1[1] -= f[1]; f1[2] -= f[2];
    atomicAdd(&s2.mbp[0][j], f[0]); atomicAdd(&s2.mbp[1][j], f[1]); atomicAdd(&s2.mbp[2][j], f[2]);
  } else {
    const double inv_r = 1.0 / r;
    const double m1 = p1.m;
    const double fmag = m1 * m2 / ((m1 + m2) * cp.dt * cp.dt) * (1 - Rp * inv_r);
    f[0] = fmag * dx; f[1] = fmag * dy; f[2] = fmag * dz;
    if (cp.mu != 0) {
      const double dvx = vx2 - p1.vx, dvy = vy2 - p1.vy, dvz = vz2 - p1.vz;
      const double dd = (dvx * dx + dvy * dy + dvz * dz) * inv_r * inv_r;
      double vt[3] = {dvx - dd * dx, dvy - dd * dy, dvz - dd * dz};
      const double vtn = sqrt(vt[0] * vt[0] + vt[1] * vt[1] + vt[2] * vt[2]);
      if (vtn != 0) {
        const double ffric = cp.mu * fmag * r;
        f[0] -= ffric * (vt[0] / vtn); f[1] -= ffric * (vt[1] / vtn); f[2] -= ffric * (vt[2] / vtn);
        if (cp.temp) {
          if (cp.dim == 2) { const double gm = ffric * vtn * cp.dt; g1 += cp.alpha * s1.vol0[p1.i] * cp.invcp1 * gm; g2 = (1.0 - cp.alpha) * s2.vol0[j] * cp.invcp2 * gm; }
          else { const double gm = cp.alpha * ffric * vtn * cp.dt; g1 += s1.vol0[p1.i] * cp.invcp1 * gm; g2 = s2.vol0[j] * cp.invcp2 * gm; }
        }
      }
    }
    // s1 += f ; s2 -= f
    f1[0] += f[0]; f1[1] += f[1]; f1[2] += f[2];
    atomicAdd(&s2.mbp[0][j], -f[0]); atomicAdd(&s2.mbp[1][j], -f[1]); atomicAdd(&s2.mbp[2][j], -f[2]);
    if (g2 != 0) atomicAdd(&s2.gamma[j], g2);
  }
  ft[0] += f[0]; ft[1] += f[1]; ft[2] += f[2];
}
__device__ __forceinline__ void contact_finish(const SolidDev &s1, long long i1, bool act, const double (&f1)[3], double g1, const double (&ft)[3], double *ftot) {
  if (act) {
    if (f1[0] != 0 || f1[1] != 0 || f1[2] != 0) { s1.mbp[0][i1] += f1[0]; s1.mbp[1][i1] += f1[1]; s1.mbp[2][i1] += f1[2]; }
    if (g1 != 0) s1.gamma[i1] += g1;
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double x = ft[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x != 0.0) atomicAdd(&ftot[d], x);
  }
}

// all pairs like the reference, tiled through shared memory (small bodies: C4 has 812 x 812)
__global__ void __launch_bounds__(128) k_contact(SolidDev s1, SolidDev s2, ContactParams cp, double *ftot) {
  __shared__ double sx[128], sy[128], sz[128], svol[128], sm[128], svx[128], svy[128], svz[128];
  const long long i1 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i1 < s1.np;
  ContactP1 p1{0, 0, 0, 1, 1, 0, 0, 0, 0, i1};
  if (act) { p1.x = s1.x[0][i1]; p1.y = s1.x[1][i1]; p1.z = s1.x[2][i1]; p1.vol = s1.vol[i1]; p1.m = s1.mass[i1]; p1.vx = s1.v[0][i1]; p1.vy = s1.v[1][i1]; p1.vz = s1.v[2][i1]; }
  p1.Rp = contact_radius(cp, p1.vol, p1.x);
  double f1[3] = {0, 0, 0}, g1 = 0, ft[3] = {0, 0, 0};
  for (long long base = 0; base < s2.np; base += 128) {
    const long long j = base + threadIdx.x;
    __syncthreads();
    if (j < s2.np) { sx[threadIdx.x] = s2.x[0][j]; sy[threadIdx.x] = s2.x[1][j]; sz[threadIdx.x] = s2.x[2][j]; svol[threadIdx.x] = s2.vol[j]; sm[threadIdx.x] = s2.mass[j];
      svx[threadIdx.x] = s2.v[0][j]; svy[threadIdx.x] = s2.v[1][j]; svz[threadIdx.x] = s2.v[2][j]; }
    __syncthreads();
    const int cnt = (int)min((long long)128, s2.np - base);
    if (!act) continue;
    for (int t = 0; t < cnt; t++) contact_pair(s1, s2, cp, p1, base + t, sx[t], sy[t], sz[t], svol[t], sm[t], svx[t], svy[t], svz[t], f1, g1, ft);
  }
  contact_finish(s1, i1, act, f1, g1, ft, ftot);
}

// Large bodies: the first screen of the reference (|dx_i| < max_cellsize) confines a particle's partners to the 3^dim bins of edge max_cellsize
// around its own, so solid 2 is binned (count, scan, fill - like the cell lists of the MPM kernels) and every particle of solid 1 visits those
// bins only: O(N1 x particles per bin neighbourhood) pair tests instead of the reference's O(N1 x N2) sweep, same screens, same pair set.
struct ContactBins { double lo[3]; double inv; int n[3]; int dim3; }; // dim3: bin along z as well (3-D, and Hertz which screens z in 2-D too)
__device__ __forceinline__ int contact_bin_axis(double x, double lo, double inv, int n) { const int b = (int)floor((x - lo) * inv); return min(max(b, 0), n - 1); }
__global__ void k_contact_bin_count(SolidDev s2, ContactBins cb, int *cell_of, int *rank, int *count) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= s2.np) return;
  const int bx = contact_bin_axis(s2.x[0][j], cb.lo[0], cb.inv, cb.n[0]), by = contact_bin_axis(s2.x[1][j], cb.lo[1], cb.inv, cb.n[1]);
  const int bz = cb.dim3 ? contact_bin_axis(s2.x[2][j], cb.lo[2], cb.inv, cb.n[2]) : 0;
  const int key = (bx * cb.n[1] + by) * cb.n[2] + bz;
  cell_of[j] = key; rank[j] = atomicAdd(&count[key], 1);
}
__global__ void k_contact_bin_fill(long long np, const int *cell_of, const int *rank, const int *start, int *order) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < np) order[start[cell_of[j]] + rank[j]] = (int)j;
}
__global__ void __launch_bounds__(128) k_contact_bins(SolidDev s1, SolidDev s2, ContactParams cp, ContactBins cb, const int *__restrict__ start, const int *__restrict__ order, double *ftot) {
  const long long i1 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i1 < s1.np;
  ContactP1 p1{0, 0, 0, 1, 1, 0, 0, 0, 0, i1};
  double f1[3] = {0, 0, 0}, g1 = 0, ft[3] = {0, 0, 0};
  if (act) {
    p1.x = s1.x[0][i1]; p1.y = s1.x[1][i1]; p1.z = s1.x[2][i1]; p1.vol = s1.vol[i1]; p1.m = s1.mass[i1]; p1.vx = s1.v[0][i1]; p1.vy = s1.v[1][i1]; p1.vz = s1.v[2][i1];
    p1.Rp = contact_radius(cp, p1.vol, p1.x);
    // partners lie within one bin edge of particle 1's bin (positions outside the box are clamped into the edge bins on both sides)
    const int bx = contact_bin_axis(p1.x, cb.lo[0], cb.inv, cb.n[0]), by = contact_bin_axis(p1.y, cb.lo[1], cb.inv, cb.n[1]), bz = cb.dim3 ? contact_bin_axis(p1.z, cb.lo[2], cb.inv, cb.n[2]) : 0;
    for (int ix = max(bx - 1, 0); ix <= min(bx + 1, cb.n[0] - 1); ix++)
      for (int iy = max(by - 1, 0); iy <= min(by + 1, cb.n[1] - 1); iy++) {
        const int z0 = cb.dim3 ? max(bz - 1, 0) : 0, z1 = cb.dim3 ? min(bz + 1, cb.n[2] - 1) : 0;
        if (z0 > z1) continue;
        const int k0 = (ix * cb.n[1] + iy) * cb.n[2];
        for (int q = start[k0 + z0]; q < start[k0 + z1 + 1]; q++) { // the bins of a z column are contiguous
          const int j = order[q];
          contact_pair(s1, s2, cp, p1, j, s2.x[0][j], s2.x[1][j], s2.x[2][j], s2.vol[j], s2.mass[j], s2.v[0][j], s2.v[1][j], s2.v[2][j], f1, g1, ft);
        }
      }
  }
  contact_finish(s1, i1, act, f1, g1, ft, ftot);
}

// ---- reductions for computes -----------------------------------------------------------------
// ComputeKineticEnergy src/compute_kinetic_energy.cpp:62-102, ComputeStrainEnergy src/compute_strain_energy.cpp:64-117
__global__ void k_energy(SolidDev s, int groupbit, int kinetic, double *out) {
  long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0;
  if (ip < s.np && (s.mask[ip] & groupbit)) {
    if (kinetic) { const double a = s.v[0][ip], b = s.v[1][ip], c = s.v[2][ip]; e = 0.5 * s.mass[ip] * (a * a + b * b + c * c); }
    else {
      double sg[9], el[9]; load_sym(s.sig, ip, sg); load_sym(s.eel, ip, el);
      double acc = 0;
#pragma unroll
      for (int i = 0; i < 9; i++) acc += sg[i] * el[i];
      e = 0.5 * s.vol[ip] * acc;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0 && e != 0.0) atomicAdd(out, e);
}

#endif // KML_MISC_KERNELS

} // namespace kml
