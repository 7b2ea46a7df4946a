/* kml.h - C ABI of the B200-native MPM time-step engine (libkml.so).
 *
 * This is the drop-in boundary for Karamelo's per-step hot path.  The
 * reference has no FFI; its boundary is the abstract C++ class `Method`
 * (reference src/method.h:25-54) called by the schemes (src/usl.cpp:43-93,
 * src/musl.cpp:43-94, src/usf.cpp:43-93), the constitutive virtuals
 * (src/strength.h:38-45, src/eos.h:38-40, src/damage.h:35-41,
 * src/temperature.h:24-27) and the Fix hooks that touch particle / node state
 * between stages (src/fix.h:42-63).  Every entry point below names the
 * reference interface it replaces.
 *
 * Conventions
 *   - plain C, opaque handle, no C++/torch types in any signature;
 *   - every function returns 0 on success, non-zero on error; the message is
 *     available from kml_last_error() (the reference calls error->one/all,
 *     src/error.cpp:33-76, which aborts; a host wrapper turns non-zero into that);
 *   - host arrays are "array of rows": vectors [n][3], matrices [n][9]
 *     row-major; the library transposes into its device SoA layout;
 *   - all arithmetic is IEEE fp64; indices int32, particle tags int64
 *     (src/mpmtype.h:14-18);
 *   - stage calls are asynchronous on the context's CUDA stream; calls that
 *     return data (download, adjust_dt, reductions) synchronise.
 */
#ifndef KML_H
#define KML_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kml_ctx kml_ctx;

/* Update::ShapeFunctions, src/update.h:70-75 */
enum { KML_SHAPE_LINEAR = 0, KML_SHAPE_CUBIC_SPLINE = 1, KML_SHAPE_QUADRATIC_SPLINE = 2, KML_SHAPE_BERNSTEIN = 3 };
/* Update::SubMethodType, src/update.h:62-69 */
enum { KML_SUB_PIC = 0, KML_SUB_FLIP = 1, KML_SUB_APIC = 2, KML_SUB_AFLIP = 3, KML_SUB_ASFLIP = 4, KML_SUB_MLS = 5 };
/* Material::constitutive_model, src/material.h */
enum { KML_MAT_LINEAR = 0, KML_MAT_NEO_HOOKEAN = 1, KML_MAT_EOS_STRENGTH = 2, KML_MAT_RIGID = 3 };
enum { KML_EOS_NONE = 0, KML_EOS_LINEAR = 1, KML_EOS_SHOCK = 2, KML_EOS_FLUID = 3 };
enum { KML_STRENGTH_NONE = 0, KML_STRENGTH_LINEAR = 1, KML_STRENGTH_PLASTIC = 2, KML_STRENGTH_JOHNSON_COOK = 3,
       KML_STRENGTH_SWIFT = 4, KML_STRENGTH_FLUID = 5 };
enum { KML_DAMAGE_NONE = 0, KML_DAMAGE_JOHNSON_COOK = 1 };
enum { KML_TEMPERATURE_NONE = 0, KML_TEMPERATURE_PLASTIC_WORK = 1 };

/* Method flags + Domain box: src/method.h:50-53, src/domain.h (boxlo/boxhi, dimension, axisymmetric). */
typedef struct kml_config {
  int dimension;      /* 1, 2, 3 */
  int is_TL;          /* tlmpm / tlcpdi */
  int is_CPDI;        /* ulcpdi / tlcpdi (2-D only in the reference) */
  int cpdi_style;     /* 0 = R4, 1 = Q4 (Method::style) */
  int shape_function; /* KML_SHAPE_* */
  int sub_method;     /* KML_SUB_* */
  double PIC_FLIP;    /* Update::PIC_FLIP, src/update.cpp:50,133-167 */
  int axisymmetric;
  int temp;           /* thermo-mechanical */
  int ge;             /* gradient-enhanced */
  double boxlo[3], boxhi[3];
  int device;         /* CUDA device ordinal */
  int rank, nranks;   /* slab (x) decomposition; 0,1 for a single GPU */
} kml_config;

/* Background grid.  Nodes sit at lo + i*h, tag = nz*ny*i + nz*j + k, i -> j -> k
 * loops, ntype by the formulas of src/grid.cpp:218-264 (h is already cellsize/2
 * for Bernstein, src/grid.cpp:82-85).  UL: one grid shared by all solids
 * (src/domain.cpp:496-551); TL: one per solid (src/solid.cpp:92-99). */
typedef struct kml_grid_desc {
  double lo[3];
  double h;        /* node spacing */
  double cellsize; /* Grid::cellsize: inv_cellsize = 1/cellsize scales r and the derivatives */
  int n[3];        /* nx, ny, nz of THIS rank's grid (1 in unused dimensions) */
  /* Slab decomposition along x (all zero = not decomposed).  lo[] stays the GLOBAL origin; local node
   * plane i is global plane i + goff of gn (node positions, tags and ntype are those of the global grid,
   * src/grid.cpp:236-251).  Planes [own_lo, own_hi) are owned by this rank; the planes above own_hi are
   * shared with (and owned by) the right neighbour and are summed after every scatter. */
  int goff, gn, own_lo, own_hi;
  /* particles whose GLOBAL stencil base lies in [base_lo, base_hi) belong to this rank */
  int base_lo, base_hi;
} kml_grid_desc;

/* Mat record + functor parameter blocks: src/material.h:28-56, src/material.cpp:720-778,
 * src/eos_*.cpp, src/strength_*.cpp, src/damage_jc.cpp, src/temperature_plastic_work.cpp. */
typedef struct kml_material {
  int type;   /* KML_MAT_* */
  int rigid;
  double rho0, E, nu, G, K, lambda, signal_velocity;
  double cp, invcp, kappa;
  int eos_type;
  /* linear: K ; shock: K, c0, S, Gamma, cv, Tr, Q1, Q2 (alpha = cv*rho0, e0 = 0) ; fluid: K, Gamma */
  double eos_K, eos_c0, eos_S, eos_Gamma, eos_cv, eos_Tr, eos_Q1, eos_Q2;
  int strength_type;
  /* linear/fluid: G ; plastic: G, A(yield) ; JC: G, A, B, n, epsdot0, C, m, Tr, Tm ; swift: G, A, B, C, n */
  double str_G, str_A, str_B, str_n, str_epsdot0, str_C, str_m, str_Tr, str_Tm;
  int damage_type;
  double dmg_d1, dmg_d2, dmg_d3, dmg_d4, dmg_d5, dmg_epsdot0, dmg_Tr, dmg_Tm;
  int temperature_type;
  double tmp_chi, tmp_cp, tmp_kappa, tmp_alpha, tmp_T0, tmp_Tm;
} kml_material;

typedef struct kml_solid_desc {
  int64_t np;       /* particles owned by this rank */
  int64_t capacity; /* >= np; room for migration (0 = np) */
  int grid;         /* grid id from kml_grid_create */
  int np_per_cell;  /* Solid::np_per_cell (src/solid.cpp:2027): selects the APIC inertia tensor of linear shape functions (src/solid.cpp:1453-1460); 0 = 2 */
  kml_material mat;
} kml_solid_desc;

/* Particle fields, Solid state of src/solid.h:50-112. */
enum {
  KML_P_PTAG = 0,    /* int64 [np]      */
  KML_P_X,           /* double [np][3]  */
  KML_P_X0,          /* double [np][3]  */
  KML_P_V,           /* double [np][3]  */
  KML_P_V_UPDATE,    /* double [np][3]  (download only, valid after grid_to_points) */
  KML_P_A,           /* double [np][3]  (download only) */
  KML_P_MBP,         /* double [np][3]  */
  KML_P_F,           /* double [np][3]  f_p = a_p m_p (download only) */
  KML_P_SIGMA,       /* double [np][9]  */
  KML_P_STRAIN_EL,   /* double [np][9]  */
  KML_P_VOL0PK1,     /* double [np][9]  (TL) */
  KML_P_FDEF,        /* double [np][9]  deformation gradient F */
  KML_P_R,           /* double [np][9]  (TL, download only) */
  KML_P_J,           /* double [np]     (download only: det F) */
  KML_P_VOL0,        /* double [np]     */
  KML_P_VOL,         /* double [np]     */
  KML_P_RHO0,        /* double [np]     */
  KML_P_RHO,         /* double [np]     (download only) */
  KML_P_MASS,        /* double [np]     */
  KML_P_EFF_PLASTIC_STRAIN,      /* double [np] */
  KML_P_EFF_PLASTIC_STRAIN_RATE, /* double [np] */
  KML_P_DAMAGE,      /* double [np]     */
  KML_P_DAMAGE_INIT, /* double [np]     */
  KML_P_IENERGY,     /* double [np]     */
  KML_P_MASK,        /* int32 [np]      */
  KML_P_T,           /* double [np]     */
  KML_P_GAMMA,       /* double [np]     */
  KML_P_Q,           /* double [np][3]  */
  KML_P_RP,          /* double [np][dim][3] CPDI-R4 domain vectors */
  KML_P_RP0,
  KML_P_XPC,         /* double [np][nc][3]  CPDI-Q4 corners */
  KML_P_XPC0,
  KML_P_NFIELDS
};

/* Node fields, Grid state of src/grid.h:81-96. */
enum {
  KML_N_X0 = 0,   /* double [nn][3] (download only) */
  KML_N_X,        /* double [nn][3] */
  KML_N_V,        /* double [nn][3] */
  KML_N_V_UPDATE, /* double [nn][3] */
  KML_N_MB,       /* double [nn][3] */
  KML_N_F,        /* double [nn][3] */
  KML_N_MASS,     /* double [nn]    */
  KML_N_MASK,     /* int32 [nn]     */
  KML_N_NTYPE,    /* int32 [nn][3]  (download only) */
  KML_N_RIGID,    /* int32 [nn]     */
  KML_N_T,        /* double [nn]    */
  KML_N_T_UPDATE, /* double [nn]    */
  KML_N_QEXT,     /* double [nn]    */
  KML_N_QINT,     /* double [nn]    */
  KML_N_NFIELDS
};

/* ---- life cycle -------------------------------------------------------------------- */
const char *kml_last_error(void);
const char *kml_backend(void); /* "cuda-sm_100a" for libkml.so */
int kml_create(const kml_config *cfg, kml_ctx **out);
int kml_destroy(kml_ctx *ctx);
int kml_synchronize(kml_ctx *ctx);
/* region(block) may widen Domain::boxlo/boxhi after the domain exists (src/region_block.cpp:62-118). */
int kml_set_domain_box(kml_ctx *ctx, const double lo[3], const double hi[3]);

/* Grid::init, src/grid.cpp:68-264 (node creation, ntype, tags). */
int kml_grid_create(kml_ctx *ctx, const kml_grid_desc *desc, int *grid_id);
int kml_grid_nnodes(kml_ctx *ctx, int grid_id, int64_t *nn);
int kml_grid_upload(kml_ctx *ctx, int grid_id, int field, const void *src);
int kml_grid_download(kml_ctx *ctx, int grid_id, int field, void *dst);

/* Solid::Solid / Solid::grow, src/solid.cpp:59-168,240-315: allocate device state. */
int kml_solid_create(kml_ctx *ctx, const kml_solid_desc *desc, int *solid_id);
int kml_solid_np(kml_ctx *ctx, int solid_id, int64_t *np);
/* Changes whenever the solid's particle set or order may have changed behind a caller's back (upload of tags / reference positions /
 * masks, delete_particles, migration between slabs, populate): hosts that mirror ptag / x0 / mask compare it before using the mirror. */
int kml_solid_generation(kml_ctx *ctx, int solid_id, uint64_t *generation);
/* The reference stores v_update, a and f = a m of every particle during grid_to_points (src/solid.cpp:576-635); the engine derives them in
 * registers and drops them.  After this call they are kept (24 + 24 bytes per particle and step) so that KML_P_V_UPDATE / KML_P_A / KML_P_F
 * can be downloaded (Group::internal_force / external_force, src/group.cpp:340-470; dump fields).  Must be called before the first step. */
int kml_keep_particle_acceleration(kml_ctx *ctx);
int kml_solid_upload(kml_ctx *ctx, int solid_id, int field, const void *src);
int kml_solid_download(kml_ctx *ctx, int solid_id, int field, void *dst);
/* DeleteParticles::command, src/delete_particles.cpp:51-80: dlist[np] flags the particles to remove.  For k ascending a flagged particle
 * is overwritten by the solid's last particle (Solid::copy_particle(np - 1, k), src/solid.cpp:1555-1589) and np shrinks by one; the
 * particle moved in is examined next - the surviving particles end up in exactly the reference's order. */
int kml_solid_delete_particles(kml_ctx *ctx, int solid_id, const int *dlist);
/* Device-resident state for callers that generate particles on the GPU
 * (synthetic blocks): device pointer of the SoA component `comp` of `field`. */
int kml_solid_device_ptr(kml_ctx *ctx, int solid_id, int field, int comp, void **dptr);

/* ---- set-up on the device ------------------------------------------------------------------
 * Region::inside of src/region_block.cpp:141-148, src/region_cylinder.cpp:140-161, src/region_sphere.cpp:120-127.
 *   block: p = xlo, xhi, ylo, yhi, zlo, zhi;  cylinder (axis 0 x | 1 y | 2 z): p = c1, c2, R^2, lo, hi;  sphere: p = c1, c2, c3, R^2.
 * interior = 0 inverts the test in Region::match (src/region.cpp:64-72). */
enum { KML_REGION_BLOCK = 0, KML_REGION_CYLINDER = 1, KML_REGION_SPHERE = 2 };
typedef struct kml_region { int style, interior, axis, pad_; double p[6]; } kml_region;
/* The particle lattice of Solid::populate (src/solid.cpp:1927-2127, :2155-2183): cells i -> j -> k of the solid's bounding box, nip
 * sub-points per cell at offsets ip (in cell units), position boundlo + delta (noffsetlo + i + 0.5 + ip); a point becomes a particle
 * iff it lies in the sub-domain and in the region.  Particles are numbered in lattice order (src/solid.cpp:2322). */
typedef struct kml_lattice {
  double boundlo[3], delta;
  int noffsetlo[3], nsub[3];
  int nip, dim;
  double ip[3 * 64];            /* nip <= 64 (4 particles per cell and direction) */
  double sublo[3], subhi[3];    /* Domain::inside_subdomain */
  double mass, vol, T0;         /* per particle; axisymmetric: mass x0[0], vol = mass / rho0 (src/solid.cpp:2292-2298) */
  int axisymmetric, set_T;
  double rho0;
  int64_t tag_first;            /* tag of the first particle of the whole (undecomposed) lattice */
  /* slab decomposition: keep the points whose global stencil base (int)((x - slab_lo) slab_ih [- 1]) lies in [base_lo, base_hi) */
  int slab, slab_linear, base_lo, base_hi;
  double slab_lo, slab_ih;
} kml_lattice;
/* 1 if this library can populate / assign groups / evaluate per-particle expressions itself (the CUDA engine), 0 otherwise. */
int kml_has_device_setup(void);
/* hist[b] = number of lattice points that become particles with clamped global stencil base b (0 <= b < nbins; every point goes to
 * bin 0 when lat->slab == 0).  The host derives the particle count, the slab cuts and the tag offsets from it. */
int kml_lattice_histogram(kml_ctx *ctx, const kml_lattice *lat, const kml_region *reg, int64_t *hist, int nbins);
/* Solid::populate, src/solid.cpp:2283-2336: fills a solid created with desc.np = the number of accepted points (of this slab):
 * ptag, x = x0, mass, vol0 = vol, T, F = R = I, mask = 1, everything else 0.  tag_offset = accepted points of lower slabs. */
int kml_solid_populate(kml_ctx *ctx, int solid_id, const kml_lattice *lat, const kml_region *reg, int64_t tag_offset);
/* Group::assign for a particle group, src/group.cpp:140-181: mask |= bit where the region matches the REFERENCE position. */
int kml_solid_group_assign(kml_ctx *ctx, int solid_id, const kml_region *reg, int bit, int64_t *count);
/* sum over the particles of component comp of a field (Solid::init totals, src/solid.cpp:170-195) */
int kml_solid_sum(kml_ctx *ctx, int solid_id, int field, int comp, double *sum);

/* Per-particle expressions of the script language, compiled by the host from the token stream of its parser into a postfix
 * program (the reference re-parses the expression string for every particle, e.g. src/fix_initial_velocity_particles.cpp:99-161).
 * Operands: constants, the particle variables x, y, z, x0, y0, z0.  +, -, *, /, sqrt are IEEE-exact on the device; pow, exp, log,
 * sin, cos, tan, atan2 are the CUDA math library's (<= 2 ulp from the host's). */
enum { KML_X_CONST = 0, KML_X_VAR /* val = 0..5: x y z x0 y0 z0 */, KML_X_ADD, KML_X_SUB, KML_X_MUL, KML_X_DIV, KML_X_POW, KML_X_NEG, KML_X_NOT,
       KML_X_GT, KML_X_GE, KML_X_LT, KML_X_LE, KML_X_EQ, KML_X_NE, KML_X_EXP, KML_X_SQRT, KML_X_COS, KML_X_SIN, KML_X_TAN, KML_X_LOG, KML_X_ATAN2 };
#define KML_EXPR_MAX 96
typedef struct kml_expr { int n; int op[KML_EXPR_MAX]; double val[KML_EXPR_MAX]; } kml_expr;
/* FixInitialVelocityParticles::initial_integrate (src/fix_initial_velocity_particles.cpp:99-161) and the other fixes that SET a
 * particle vector per component: field (KML_P_V) component d of every particle of the group = prog[d](x, y, z, x0, y0, z0)
 * for the components in set_mask.  solid = -1: every solid. */
int kml_fix_set_particles_expr(kml_ctx *ctx, int solid, int groupbit, int field, int set_mask, const kml_expr prog[3]);

/* Update::dt (src/update.h:29); Update::set_dt / ULMPM::adjust_dt write it. */
int kml_set_dt(kml_ctx *ctx, double dt);
int kml_get_dt(kml_ctx *ctx, double *dt);

/* ---- Method stage API, one entry per virtual of src/method.h:33-48 ---------------------- */
/* ULMPM/TLMPM/ULCPDI/TLCPDI::compute_grid_weight_functions_and_gradients
 * (src/ulmpm.cpp:88-337, src/tlmpm.cpp:87-340, src/ulcpdi.cpp:111-393, src/tlcpdi.cpp:98-351). */
int kml_compute_grid_weight_functions_and_gradients(kml_ctx *ctx);
/* ::reset, src/ulmpm.cpp:553-563 */
int kml_reset(kml_ctx *ctx);
/* ::particles_to_grid (+USF halves), src/ulmpm.cpp:339-431, src/tlmpm.cpp:342-408 */
int kml_particles_to_grid(kml_ctx *ctx);
int kml_particles_to_grid_USF_1(kml_ctx *ctx);
int kml_particles_to_grid_USF_2(kml_ctx *ctx);
/* ::update_grid_state, src/ulmpm.cpp:434-438 -> Grid::update_grid_velocities/_temperature src/grid.cpp:448-466,1354-1362 */
int kml_update_grid_state(kml_ctx *ctx);
/* ::grid_to_points, src/ulmpm.cpp:440-462 -> Solid::compute_particle_accelerations_velocities_and_positions src/solid.cpp:576-635 */
int kml_grid_to_points(kml_ctx *ctx);
/* ::advance_particles, src/ulmpm.cpp:464-474 -> Solid::update_particle_velocities src/solid.cpp:786-796 */
int kml_advance_particles(kml_ctx *ctx);
/* ::velocities_to_grid, src/ulmpm.cpp:476-496 */
int kml_velocities_to_grid(kml_ctx *ctx);
/* ::update_grid_positions, src/ulmpm.h:47 (no-op), src/tlmpm.cpp:455-460 */
int kml_update_grid_positions(kml_ctx *ctx);
/* ::compute_rate_deformation_gradient, src/ulmpm.cpp:498-505 -> src/solid.cpp:798-936,1006-1153 */
int kml_compute_rate_deformation_gradient(kml_ctx *ctx, int doublemapping);
/* ::update_deformation_gradient -> src/solid.cpp:1155-1244 */
int kml_update_deformation_gradient(kml_ctx *ctx);
/* ::update_stress -> src/solid.cpp:1246-1438 (+ update_heat_flux src/solid.cpp:2810-2839) */
int kml_update_stress(kml_ctx *ctx, int doublemapping);
/* ::adjust_dt, src/ulmpm.cpp:525-551: dt = min_solids(dtCFL) * dt_factor. Returns the new dt. */
int kml_adjust_dt(kml_ctx *ctx, double dt_factor, double *dt_out);
/* ::exchange_particles, src/ulmpm.cpp:565-667 (slab neighbours over NCCL; no-op on one rank) */
int kml_exchange_particles(kml_ctx *ctx);

/* ---- Fix hooks that run between stages -------------------------------------------------- */
/* FixVelocityNodes::post_update_grid_state / post_velocities_to_grid, src/fix_velocity_nodes.cpp:130-268.
 * which = 0: set v_update = v and v = vprev on masked nodes, ftot += mass*(v - v_update_old)/dt
 * which = 1: set v = v on masked nodes.  set_mask bit d = component d is set. solid = -1: all grids. */
int kml_fix_velocity_nodes(kml_ctx *ctx, int solid, int groupbit, int set_mask, const double v[3],
                           const double vprev[3], int which, double ftot[3]);
/* FixBodyforce::post_particles_to_grid with a constant force, src/fix_body_force.cpp:106-180 */
int kml_fix_body_force(kml_ctx *ctx, int solid, int groupbit, int set_mask, const double f[3], double ftot[3]);
/* FixForceNodes::post_particles_to_grid, src/fix_force_nodes.cpp:96-193: the force f is shared equally by the n nodes of the
 * group that carry mass (mb_I += f / n). solid = -1: every solid's grid, n counted per grid. */
int kml_fix_force_nodes(kml_ctx *ctx, int solid, int groupbit, int set_mask, const double f[3], double ftot[3]);
/* FixVelocityParticles::initial_integrate / post_advance_particles, src/fix_velocity_particles.cpp:131-300, for values that do not
 * depend on the particle (x, y, z, x0, y0, z0 absent from the expressions; the host evaluates v(t) and v(t - dt)).
 * which = 0 (before the step): v = vprev on the group's particles, their positions are remembered;
 * which = 1 (after advance_particles): ftot += m (v - v_p) / dt, v_p = v, x_p = x_remembered + dt v.  solid = -1: every solid. */
int kml_fix_velocity_particles(kml_ctx *ctx, int solid, int groupbit, int set_mask, const double v[3], const double vprev[3], int which,
                               double ftot[3]);
/* FixTemperatureNodes::post_update_grid_state / post_velocities_to_grid, src/fix_temperature_nodes.cpp:74-146.
 * which = 0: T_update = T, T = Tprev on masked nodes; which = 1: T = T.  solid = -1: all grids. */
int kml_fix_temperature_nodes(kml_ctx *ctx, int solid, int groupbit, double T, double Tprev, int which);
/* FixTemperatureParticles::initial_integrate / post_advance_particles with a particle-independent value,
 * src/fix_temperature_particles.cpp:92-181: T_p = T on the group's particles.  solid = -1: every solid. */
int kml_fix_temperature_particles(kml_ctx *ctx, int solid, int groupbit, double T);
/* FixContactHertz::initial_integrate, src/fix_contact_hertz.cpp:84-201 */
int kml_fix_contact_hertz(kml_ctx *ctx, int solid1, int solid2, double ftot[3]);
/* FixContactMinPenetration::initial_integrate, src/fix_contact_min_penetration.cpp:88-258 */
int kml_fix_contact_min_penetration(kml_ctx *ctx, int solid1, int solid2, double mu, double ftot[3]);

/* ---- reductions used by computes / log (src/compute_kinetic_energy.cpp:62-102,
 *      src/compute_strain_energy.cpp:60-117) ---------------------------------------------- */
int kml_compute_kinetic_energy(kml_ctx *ctx, int solid, int groupbit, double *ek);
int kml_compute_strain_energy(kml_ctx *ctx, int solid, int groupbit, double *es);

/* Device error word, the invariants of SURVEY section 4: bit0 particle left the domain
 * (src/solid.cpp:617-627), bit1 J <= 0 (src/solid.cpp:1208-1215), bit2 dtCFL NaN/0
 * (src/ulmpm.cpp:535-544), bit3 polar decomposition failed (src/solid.cpp:1229-1236); engine-side conditions: bit4 CPDI neighbour list
 * overflow, bit5 particle migration bookkeeping, bit6 halo exchange timed out (a neighbour rank never delivered its shared planes). */
int kml_error_flags(kml_ctx *ctx, unsigned *flags); /* on a decomposed run: collective (call on every rank), returns the union */

/* ---- slab decomposition over several GPUs (replaces Grid::reduce_ghost_nodes
 *      src/grid.cpp:477-621,881-1132 and ULMPM::exchange_particles) ------------------------ */
/* 128-byte NCCL unique id created on rank 0 and passed to every rank. */
int kml_comm_unique_id(void *id128);
int kml_comm_init(kml_ctx *ctx, const void *id128);
/* In-place sum over the ranks of n <= 16 host doubles (the MPI_Allreduce(SUM) of the reference's set-up commands, e.g.
 * src/delete_particles.cpp:76-78, src/solid.cpp:2302-2310); collective; a no-op on one rank. */
int kml_comm_sum(kml_ctx *ctx, double *vals, int n);

/* ---- measurement ------------------------------------------------------------------------- */
/* Per-stage device time (ms) accumulated with CUDA events on the context's stream since the last
 * reset: stage ids KML_STAGE_*.  Enabled with kml_profile(ctx, 1).  The events are only recorded while the steps run and
 * read here (one synchronisation), so the profile can stay on inside a timed region. */
enum { KML_STAGE_REBIN = 0, KML_STAGE_P2G, KML_STAGE_GRID, KML_STAGE_G2P, KML_STAGE_V2G, KML_STAGE_STRESS,
       KML_STAGE_CONTACT, KML_STAGE_OTHER, KML_STAGE_HALO /* ghost-node sums between slabs */, KML_STAGE_MIGRATE /* exchange_particles */,
       KML_STAGE_DT /* adjust_dt: reduction + readback */, KML_STAGE_COUNT };
int kml_profile(kml_ctx *ctx, int enable);
int kml_stage_times(kml_ctx *ctx, double ms[KML_STAGE_COUNT], int64_t launches[KML_STAGE_COUNT], int reset);
/* HOST wall-clock time (ms) spent inside the calls of each stage since the last reset (launch overhead, waits on events / NCCL): the
 * difference between a step's wall time and the device stage sum is looked for here. */
int kml_stage_host_times(kml_ctx *ctx, double ms[KML_STAGE_COUNT], int reset);
/* CUDA-event bracket on the context's stream: start records an event, stop records a second one,
 * synchronises and returns the elapsed device time in milliseconds. */
int kml_timer_start(kml_ctx *ctx);
int kml_timer_stop(kml_ctx *ctx, double *ms);

#ifdef __cplusplus
}
#endif
#endif
