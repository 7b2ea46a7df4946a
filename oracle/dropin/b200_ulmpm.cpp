// b200_ulmpm.cpp - the drop-in, compiled INTO the reference: an implementation of the reference's own `class ULMPM`
// (declared in /root/reference/src/ulmpm.h, registered as method style "ulmpm" by src/style_method.h -> src/update.cpp:119-124)
// whose stage functions forward to the C ABI of include/kml.h.  Linking the unmodified reference objects WITHOUT their
// ulmpm.o and WITH this file + libkml.so gives a `karamelo` binary that runs every unmodified ULMPM script
// (`method(ulmpm, ...)`) on the B200 engine: input parsing, regions, Solid::populate, groups, fixes, computes, dumps, log and
// restart files remain the reference's own code (oracle/Makefile target `ref_b200`, tests/test_dropin.py).
//
// TEST / INTEGRATION ARTEFACT: it lives under oracle/ because it needs /root/reference to compile; the product
// (karamelo_b200/) never links it.  No reference code is copied: only its headers are included.
//
// State ownership.  The reference keeps all state in host std::vectors that fixes, computes and dumps touch directly
// between the stages (src/modify.cpp:233-302, src/output.cpp:128-199).  This binding keeps those vectors coherent the simple
// way: before a stage it uploads the fields a hook may have modified since the last stage, after a stage it downloads the
// fields the stage produced ("slow, correct" - INTEGRATION.md section 1 lists the device-side fix kernels a maintainer
// would switch to, hook by hook, to drop these copies).
#include "ulmpm.h"
#include "domain.h"
#include "error.h"
#include "grid.h"
#include "input.h"
#include "material.h"
#include "eos.h"
#include "strength.h"
#include "damage.h"
#include "temperature.h"
#include "solid.h"
#include "universe.h"
#include "update.h"
#include "var.h"
#include "kml.h" // this repo's include/kml.h
#include <cstdio>
#include <fstream>
#include <map>
#include <unistd.h>

using namespace std;

namespace {
// ULMPM's members are fixed by the reference header; the binding's own state hangs off the object's address
struct B200State {
  kml_ctx *ctx = nullptr; int grid = -1; vector<int> solid; bool ready = false;
  vector<double> buf3, buf9;
};
map<const ULMPM *, B200State> g_state;

// parameter block of an EOS / strength / damage / temperature object: the classes keep their parameters protected, but every one of
// them serialises them through its public write_restart (src/eos_shock.cpp:135-146, src/strength_jc.cpp:186-196, ...)
template <class T> vector<double> params_of(T *obj) {
  char name[] = "/tmp/kml_dropin_XXXXXX"; const int fd = mkstemp(name); if (fd >= 0) close(fd);
  { ofstream of(name, ios::binary); obj->write_restart(&of); }
  vector<double> p; { ifstream in(name, ios::binary); double v; while (in.read(reinterpret_cast<char *>(&v), sizeof v)) p.push_back(v); }
  remove(name);
  return p;
}
} // namespace

#define ST() B200State &st = g_state[this]
#define CK(call) do { if (call) error->one(FLERR, string("kml: ") + kml_last_error() + "\n"); } while (0)

ULMPM::ULMPM(MPM *mpm) : Method(mpm) {
  update_Di = 1; rigid_solids = 0; apic = false;
  update->PIC_FLIP = 0.99; // src/ulmpm.cpp:37
  is_TL = false; is_CPDI = false; ge = false; temp = false;
  basis_function = nullptr; derivative_basis_function = nullptr;
}
ULMPM::~ULMPM() { auto it = g_state.find(this); if (it != g_state.end()) { if (it->second.ctx) kml_destroy(it->second.ctx); g_state.erase(it); } }

void ULMPM::setup(vector<string> args) { // the sub-method flags of src/ulmpm.cpp:50-86; the shape functions live in the kernels
  if (args.size() > 0) error->all(FLERR, "Illegal modify_method command: too many arguments.\n");
  const Update::SubMethodType sm = update->sub_method_type;
  if (sm == Update::SubMethodType::APIC || sm == Update::SubMethodType::MLS) { apic = true; update->PIC_FLIP = 0; }
  else if (sm == Update::SubMethodType::ASFLIP || sm == Update::SubMethodType::AFLIP) apic = true;
}

// ---- first step: mirror the object graph the script built into the engine ---------------------------------------------
static kml_material material_of(Mat *m, Error *error) {
  kml_material k{}; k.rho0 = m->rho0; k.E = m->E; k.nu = m->nu; k.G = m->G; k.K = m->K; k.lambda = m->lambda; k.signal_velocity = m->signal_velocity;
  k.cp = m->cp; k.invcp = m->invcp; k.kappa = m->kappa; k.rigid = m->rigid;
  switch (m->type) { // Material::constitutive_model, src/material.h:114-119
  case 0: k.type = KML_MAT_RIGID; return k;
  case 1: k.type = KML_MAT_LINEAR; return k;
  case 2: k.type = KML_MAT_NEO_HOOKEAN; return k;
  }
  k.type = KML_MAT_EOS_STRENGTH;
  { const vector<double> p = params_of(m->eos); const string &s = m->eos->style; // rho0, K, then the style's own block
    k.eos_K = p.at(1);
    if (s == "linear") k.eos_type = KML_EOS_LINEAR;
    else if (s == "shock") { k.eos_type = KML_EOS_SHOCK; k.eos_c0 = p.at(2); k.eos_S = p.at(3); k.eos_Gamma = p.at(4); k.eos_Tr = p.at(5); k.eos_cv = p.at(6); k.eos_Q1 = p.at(7); k.eos_Q2 = p.at(8); }
    else if (s == "fluid") { k.eos_type = KML_EOS_FLUID; k.eos_Gamma = p.at(2); }
    else error->one(FLERR, "b200 drop-in: unknown EOS style " + s + "\n"); }
  { const vector<double> p = params_of(m->strength); const string &s = m->strength->style;
    k.str_G = p.at(0);
    if (s == "linear") k.strength_type = KML_STRENGTH_LINEAR;
    else if (s == "fluid") k.strength_type = KML_STRENGTH_FLUID;
    else if (s == "plastic") { k.strength_type = KML_STRENGTH_PLASTIC; k.str_A = p.at(1); }
    else if (s == "johnson_cook") { k.strength_type = KML_STRENGTH_JOHNSON_COOK; k.str_A = p.at(1); k.str_B = p.at(2); k.str_n = p.at(3); k.str_m = p.at(4); k.str_epsdot0 = p.at(5); k.str_C = p.at(6); k.str_Tr = p.at(7); k.str_Tm = p.at(8); }
    else if (s == "swift") { k.strength_type = KML_STRENGTH_SWIFT; k.str_A = p.at(1); k.str_B = p.at(2); k.str_C = p.at(3); k.str_n = p.at(4); }
    else error->one(FLERR, "b200 drop-in: unknown strength style " + s + "\n"); }
  if (m->damage) { const vector<double> p = params_of(m->damage);
    k.damage_type = KML_DAMAGE_JOHNSON_COOK; k.dmg_d1 = p.at(0); k.dmg_d2 = p.at(1); k.dmg_d3 = p.at(2); k.dmg_d4 = p.at(3); k.dmg_d5 = p.at(4); k.dmg_epsdot0 = p.at(5); k.dmg_Tr = p.at(6); k.dmg_Tm = p.at(7); }
  if (m->temp) { const vector<double> p = params_of(m->temp);
    k.temperature_type = KML_TEMPERATURE_PLASTIC_WORK; k.tmp_chi = p.at(0); k.tmp_kappa = p.at(1); k.tmp_cp = p.at(2); k.tmp_alpha = p.at(3); k.tmp_T0 = p.at(4); k.tmp_Tm = p.at(5); }
  return k;
}

// vector<Eigen::Matrix3d> (column-major) <-> the ABI's row-major [np][9]
static void mats_out(const vector<Eigen::Matrix3d> &m, int n, vector<double> &b) { b.resize(9 * (size_t)n); for (int i = 0; i < n; i++) for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) b[9 * (size_t)i + 3 * r + c] = m[i](r, c); }
static void mats_in(vector<Eigen::Matrix3d> &m, int n, const vector<double> &b) { for (int i = 0; i < n; i++) for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) m[i](r, c) = b[9 * (size_t)i + 3 * r + c]; }
static void vecs_out(const vector<Eigen::Vector3d> &v, int n, vector<double> &b) { b.resize(3 * (size_t)n); for (int i = 0; i < n; i++) for (int d = 0; d < 3; d++) b[3 * (size_t)i + d] = v[i][d]; }
static void vecs_in(vector<Eigen::Vector3d> &v, int n, const vector<double> &b) { for (int i = 0; i < n; i++) for (int d = 0; d < 3; d++) v[i][d] = b[3 * (size_t)i + d]; }

namespace {
struct Sync { // field transfers between the reference's vectors and the engine
  Error *error; B200State &st; Domain *domain;
  void up3(int is, int f, const vector<Eigen::Vector3d> &v) { Solid *S = domain->solids[is]; vecs_out(v, S->np_local, st.buf3); if (kml_solid_upload(st.ctx, st.solid[is], f, st.buf3.data())) error->one(FLERR, kml_last_error()); }
  void dn3(int is, int f, vector<Eigen::Vector3d> &v) { Solid *S = domain->solids[is]; st.buf3.resize(3 * (size_t)S->np_local); if (kml_solid_download(st.ctx, st.solid[is], f, st.buf3.data())) error->one(FLERR, kml_last_error()); vecs_in(v, S->np_local, st.buf3); }
  void up9(int is, int f, const vector<Eigen::Matrix3d> &m) { Solid *S = domain->solids[is]; mats_out(m, S->np_local, st.buf9); if (kml_solid_upload(st.ctx, st.solid[is], f, st.buf9.data())) error->one(FLERR, kml_last_error()); }
  void dn9(int is, int f, vector<Eigen::Matrix3d> &m) { Solid *S = domain->solids[is]; st.buf9.resize(9 * (size_t)S->np_local); if (kml_solid_download(st.ctx, st.solid[is], f, st.buf9.data())) error->one(FLERR, kml_last_error()); mats_in(m, S->np_local, st.buf9); }
  void up1(int is, int f, const vector<double> &v) { if (kml_solid_upload(st.ctx, st.solid[is], f, v.data())) error->one(FLERR, kml_last_error()); }
  void dn1(int is, int f, vector<double> &v) { if (kml_solid_download(st.ctx, st.solid[is], f, v.data())) error->one(FLERR, kml_last_error()); }
  void gup3(int f, const vector<Eigen::Vector3d> &v) { Grid *g = domain->grid; vecs_out(v, (int)g->nnodes_local, st.buf3); if (kml_grid_upload(st.ctx, st.grid, f, st.buf3.data())) error->one(FLERR, kml_last_error()); }
  void gdn3(int f, vector<Eigen::Vector3d> &v) { Grid *g = domain->grid; st.buf3.resize(3 * (size_t)g->nnodes_local); if (kml_grid_download(st.ctx, st.grid, f, st.buf3.data())) error->one(FLERR, kml_last_error()); vecs_in(v, (int)g->nnodes_local, st.buf3); }
  void gdn1(int f, vector<double> &v) { if (kml_grid_download(st.ctx, st.grid, f, v.data())) error->one(FLERR, kml_last_error()); }
  void gup1(int f, const vector<double> &v) { if (kml_grid_upload(st.ctx, st.grid, f, v.data())) error->one(FLERR, kml_last_error()); }
};
} // namespace
#define SYNC() Sync sy{error, st, domain}

static void first_upload(ULMPM *self, B200State &st, Domain *domain, Update *update, Universe *universe, Error *error, bool temp, bool ge) {
  if (universe->nprocs != 1) error->all(FLERR, "b200 drop-in: one rank per GPU; this build runs single-rank (multi-GPU goes through kml_comm_init)\n");
  kml_config cfg{}; cfg.dimension = domain->dimension; cfg.is_TL = 0; cfg.is_CPDI = 0;
  switch (update->shape_function) { // Update::ShapeFunctions, src/update.h:70-75
  case Update::ShapeFunctions::LINEAR: cfg.shape_function = KML_SHAPE_LINEAR; break;
  case Update::ShapeFunctions::CUBIC_SPLINE: cfg.shape_function = KML_SHAPE_CUBIC_SPLINE; break;
  case Update::ShapeFunctions::QUADRATIC_SPLINE: cfg.shape_function = KML_SHAPE_QUADRATIC_SPLINE; break;
  default: cfg.shape_function = KML_SHAPE_BERNSTEIN; break;
  }
  switch (update->sub_method_type) { // Update::SubMethodType, src/update.h:62-69
  case Update::SubMethodType::PIC: cfg.sub_method = KML_SUB_PIC; break; case Update::SubMethodType::FLIP: cfg.sub_method = KML_SUB_FLIP; break;
  case Update::SubMethodType::APIC: cfg.sub_method = KML_SUB_APIC; break; case Update::SubMethodType::AFLIP: cfg.sub_method = KML_SUB_AFLIP; break;
  case Update::SubMethodType::ASFLIP: cfg.sub_method = KML_SUB_ASFLIP; break; default: cfg.sub_method = KML_SUB_MLS; break;
  }
  cfg.PIC_FLIP = update->PIC_FLIP; cfg.axisymmetric = domain->axisymmetric; cfg.temp = temp; cfg.ge = ge;
  for (int d = 0; d < 3; d++) { cfg.boxlo[d] = domain->boxlo[d]; cfg.boxhi[d] = domain->boxhi[d]; }
  cfg.device = 0; cfg.rank = 0; cfg.nranks = 1;
  if (kml_create(&cfg, &st.ctx)) error->one(FLERR, string("kml_create: ") + kml_last_error() + "\n");
  Grid *g = domain->grid;
  kml_grid_desc gd{}; for (int d = 0; d < 3; d++) gd.lo[d] = domain->boxlo[d];
  gd.cellsize = g->cellsize; gd.h = cfg.shape_function == KML_SHAPE_BERNSTEIN ? g->cellsize / 2 : g->cellsize; // src/grid.cpp:82-85
  gd.n[0] = g->nx_global; gd.n[1] = g->ny_global; gd.n[2] = g->nz_global;
  if (kml_grid_create(st.ctx, &gd, &st.grid)) error->one(FLERR, kml_last_error());
  { int64_t nn = 0; kml_grid_nnodes(st.ctx, st.grid, &nn); if (nn != (int64_t)g->nnodes_local) error->one(FLERR, "b200 drop-in: node count differs from the reference grid\n"); }
  if (kml_grid_upload(st.ctx, st.grid, KML_N_MASK, g->mask.data())) error->one(FLERR, kml_last_error());
  Sync sy{error, st, domain};
  for (size_t is = 0; is < domain->solids.size(); is++) {
    Solid *S = domain->solids[is];
    kml_solid_desc sd{}; sd.np = S->np_local; sd.capacity = S->np_local; sd.grid = st.grid; sd.np_per_cell = S->np_per_cell; sd.mat = material_of(S->mat, error);
    int id = -1; if (kml_solid_create(st.ctx, &sd, &id)) error->one(FLERR, kml_last_error());
    st.solid.push_back(id);
    static_assert(sizeof(tagint) == 8, "particle tags are 64-bit");
    if (kml_solid_upload(st.ctx, id, KML_P_PTAG, S->ptag.data()) || kml_solid_upload(st.ctx, id, KML_P_MASK, S->mask.data())) error->one(FLERR, kml_last_error());
    sy.up3(is, KML_P_X, S->x); sy.up3(is, KML_P_X0, S->x0); sy.up3(is, KML_P_V, S->v);
    sy.up9(is, KML_P_SIGMA, S->sigma); sy.up9(is, KML_P_STRAIN_EL, S->strain_el); sy.up9(is, KML_P_FDEF, S->F);
    sy.up1(is, KML_P_VOL0, S->vol0); sy.up1(is, KML_P_VOL, S->vol); sy.up1(is, KML_P_RHO0, S->rho0); sy.up1(is, KML_P_MASS, S->mass);
    sy.up1(is, KML_P_EFF_PLASTIC_STRAIN, S->eff_plastic_strain); sy.up1(is, KML_P_EFF_PLASTIC_STRAIN_RATE, S->eff_plastic_strain_rate);
    sy.up1(is, KML_P_DAMAGE, S->damage); sy.up1(is, KML_P_DAMAGE_INIT, S->damage_init);
    if (temp) sy.up1(is, KML_P_T, S->T);
  }
  if (kml_keep_particle_acceleration(st.ctx)) error->one(FLERR, kml_last_error()); // the reference stores a, f, v_update (src/solid.cpp:576-635)
  if (kml_set_dt(st.ctx, update->dt)) error->one(FLERR, kml_last_error());
  st.ready = true; (void)self;
}

void ULMPM::compute_grid_weight_functions_and_gradients() {
  ST();
  if (!st.ready) first_upload(this, st, domain, update, universe, error, temp, ge);
  SYNC();
  // fixes of the previous step's tail (post_advance_particles, final_integrate) and set-up commands may have edited particles
  for (size_t is = 0; is < domain->solids.size(); is++) { sy.up3(is, KML_P_X, domain->solids[is]->x); sy.up3(is, KML_P_V, domain->solids[is]->v); }
  CK(kml_set_dt(st.ctx, update->dt));
  CK(kml_compute_grid_weight_functions_and_gradients(st.ctx));
}
void ULMPM::reset() { // src/ulmpm.cpp:553-563
  ST(); CK(kml_reset(st.ctx));
  for (Solid *S : domain->solids) { S->dtCFL = 1.0e22; for (int ip = 0; ip < S->np_local; ip++) S->mbp[ip] = Eigen::Vector3d(); }
}
void ULMPM::particles_to_grid() {
  ST(); SYNC();
  for (size_t is = 0; is < domain->solids.size(); is++) { Solid *S = domain->solids[is]; sy.up3(is, KML_P_V, S->v); sy.up3(is, KML_P_MBP, S->mbp); if (temp) { sy.up1(is, KML_P_T, S->T); sy.up1(is, KML_P_GAMMA, S->gamma); } } // initial_integrate hooks
  CK(kml_particles_to_grid(st.ctx));
  Grid *g = domain->grid; sy.gdn1(KML_N_MASS, g->mass); sy.gdn3(KML_N_V, g->v); sy.gdn3(KML_N_F, g->f); sy.gdn3(KML_N_MB, g->mb);
}
void ULMPM::particles_to_grid_USF_1() { ST(); SYNC(); for (size_t is = 0; is < domain->solids.size(); is++) sy.up3(is, KML_P_V, domain->solids[is]->v); CK(kml_particles_to_grid_USF_1(st.ctx)); sy.gdn1(KML_N_MASS, domain->grid->mass); sy.gdn3(KML_N_V, domain->grid->v); }
void ULMPM::particles_to_grid_USF_2() { ST(); SYNC(); sy.gup3(KML_N_V, domain->grid->v); for (size_t is = 0; is < domain->solids.size(); is++) sy.up3(is, KML_P_MBP, domain->solids[is]->mbp); CK(kml_particles_to_grid_USF_2(st.ctx)); sy.gdn3(KML_N_F, domain->grid->f); sy.gdn3(KML_N_MB, domain->grid->mb); }
void ULMPM::update_grid_state() {
  ST(); SYNC();
  sy.gup3(KML_N_MB, domain->grid->mb); sy.gup3(KML_N_F, domain->grid->f); // post_particles_to_grid hooks (body force, force_nodes)
  CK(kml_update_grid_state(st.ctx));
  sy.gdn3(KML_N_V_UPDATE, domain->grid->v_update);
  if (temp) { sy.gdn1(KML_N_T, domain->grid->T); sy.gdn1(KML_N_T_UPDATE, domain->grid->T_update); }
}
void ULMPM::grid_to_points() {
  ST(); SYNC();
  sy.gup3(KML_N_V_UPDATE, domain->grid->v_update); sy.gup3(KML_N_V, domain->grid->v); // post_update_grid_state hooks (velocity_nodes)
  if (temp) { sy.gup1(KML_N_T, domain->grid->T); sy.gup1(KML_N_T_UPDATE, domain->grid->T_update); } // (temperature_nodes)
  CK(kml_grid_to_points(st.ctx));
}
void ULMPM::advance_particles() {
  ST(); SYNC();
  CK(kml_advance_particles(st.ctx));
  for (size_t is = 0; is < domain->solids.size(); is++) {
    Solid *S = domain->solids[is];
    sy.dn3(is, KML_P_X, S->x); sy.dn3(is, KML_P_V, S->v); sy.dn3(is, KML_P_V_UPDATE, S->v_update); sy.dn3(is, KML_P_A, S->a); sy.dn3(is, KML_P_F, S->f);
    if (temp) sy.dn1(is, KML_P_T, S->T);
  }
}
void ULMPM::velocities_to_grid() {
  ST(); SYNC();
  for (size_t is = 0; is < domain->solids.size(); is++) { sy.up3(is, KML_P_V, domain->solids[is]->v); sy.up3(is, KML_P_X, domain->solids[is]->x); if (temp) sy.up1(is, KML_P_T, domain->solids[is]->T); } // post_advance_particles hooks
  CK(kml_velocities_to_grid(st.ctx));
  sy.gdn3(KML_N_V, domain->grid->v);
  if (temp) sy.gdn1(KML_N_T, domain->grid->T);
}
void ULMPM::compute_rate_deformation_gradient(bool doublemapping) {
  ST(); SYNC();
  if (doublemapping) sy.gup3(KML_N_V, domain->grid->v); else sy.gup3(KML_N_V_UPDATE, domain->grid->v_update); // post_velocities_to_grid / post_update_grid_state hooks
  if (temp) { if (doublemapping) sy.gup1(KML_N_T, domain->grid->T); else sy.gup1(KML_N_T_UPDATE, domain->grid->T_update); }
  CK(kml_compute_rate_deformation_gradient(st.ctx, doublemapping));
}
void ULMPM::update_deformation_gradient() { ST(); CK(kml_update_deformation_gradient(st.ctx)); }
void ULMPM::update_stress(bool doublemapping) {
  ST(); SYNC();
  CK(kml_update_stress(st.ctx, doublemapping));
  for (size_t is = 0; is < domain->solids.size(); is++) { // what dumps, computes and the restart writer read
    Solid *S = domain->solids[is];
    sy.dn9(is, KML_P_SIGMA, S->sigma); sy.dn9(is, KML_P_STRAIN_EL, S->strain_el); sy.dn9(is, KML_P_FDEF, S->F);
    sy.dn1(is, KML_P_VOL, S->vol); sy.dn1(is, KML_P_J, S->J); sy.dn1(is, KML_P_RHO, S->rho);
    sy.dn1(is, KML_P_EFF_PLASTIC_STRAIN, S->eff_plastic_strain); sy.dn1(is, KML_P_EFF_PLASTIC_STRAIN_RATE, S->eff_plastic_strain_rate);
    sy.dn1(is, KML_P_DAMAGE, S->damage); sy.dn1(is, KML_P_DAMAGE_INIT, S->damage_init); sy.dn1(is, KML_P_IENERGY, S->ienergy);
    if (temp) { sy.dn1(is, KML_P_T, S->T); sy.dn1(is, KML_P_GAMMA, S->gamma); sy.dn3(is, KML_P_Q, S->q); }
  }
}
void ULMPM::adjust_dt() { // src/ulmpm.cpp:525-551
  if (update->dt_constant) return;
  ST(); double dt = 0;
  CK(kml_adjust_dt(st.ctx, update->dt_factor, &dt));
  update->dt = dt; (*input->vars)["dt"] = Var("dt", update->dt);
}
void ULMPM::exchange_particles() { ST(); CK(kml_exchange_particles(st.ctx)); }
