// sim.cpp - script commands that build the simulation (setup path).
// Integer-deciding arithmetic is restated from the reference, cited inline.
#include "sim.h"
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <iostream>

namespace kmlh {

#define BIG 1.0e20 /* src/region_block.cpp:24 */

static double dmax(double a, double b) { return a > b ? a : b; } // MAX/MIN macros, src/pointers.h:23-24
static double dmin(double a, double b) { return a < b ? a : b; }

Sim::Sim() : input(this) {
  for (int i = 0; i < MAX_GROUP; i++) { gnames[i] = ""; gbitmask[i] = 1 << i; gpon[i] = "all"; gsolid[i] = -1; gregion[i] = -1; }
  gnames[0] = "all";
  register_commands();
}
Sim::~Sim() { if (ctx) kml_destroy(ctx); }

void Sim::check(int rc) const { if (rc) fatal(std::string("kml: ") + kml_last_error()); }

int Sim::find_region(const std::string &n) const { for (size_t i = 0; i < regions.size(); i++) if (regions[i]->id == n) return (int)i; return -1; }
int Sim::find_solid(const std::string &n) const { for (size_t i = 0; i < solids.size(); i++) if (solids[i]->id == n) return (int)i; return -1; }
int Sim::find_material(const std::string &n) const { for (size_t i = 0; i < materials.size(); i++) if (materials[i].id == n) return (int)i; return -1; }
int Sim::find_group(const std::string &n) const { for (int i = 0; i < MAX_GROUP; i++) if (gnames[i] == n) return i; return -1; }

void Sim::register_commands() {
  auto &c = input.commands;
  c["method"] = [this](std::vector<std::string> &a) { return cmd_method(a); };
  c["scheme"] = [this](std::vector<std::string> &a) { return cmd_scheme(a); };
  c["dimension"] = [this](std::vector<std::string> &a) { return cmd_dimension(a); };
  c["axisymmetric"] = [this](std::vector<std::string> &a) { // Domain::set_axisymmetric, src/domain.cpp:557-569
    if (a.size() != 1) fatal("Error: axisymmetric takes one argument.\n");
    if (a[0] == "true") axisymmetric = true; else if (a[0] == "false") axisymmetric = false;
    return Var(0);
  };
  c["region"] = [this](std::vector<std::string> &a) { return cmd_region(a); };
  c["eos"] = [this](std::vector<std::string> &a) { return cmd_eos(a); };
  c["strength"] = [this](std::vector<std::string> &a) { return cmd_strength(a); };
  c["damage"] = [this](std::vector<std::string> &a) { return cmd_damage(a); };
  c["temperature"] = [this](std::vector<std::string> &a) { return cmd_temperature(a); };
  c["material"] = [this](std::vector<std::string> &a) { return cmd_material(a); };
  c["solid"] = [this](std::vector<std::string> &a) { return cmd_solid(a); };
  c["group"] = [this](std::vector<std::string> &a) { return cmd_group(a); };
  c["fix"] = [this](std::vector<std::string> &a) { return cmd_fix(a); };
  c["delete_fix"] = [this](std::vector<std::string> &a) {
    for (size_t i = 0; i < fixes.size(); i++) if (fixes[i]->id == a[0]) { fixes.erase(fixes.begin() + i); return Var(0); }
    fatal("Error: fix " + a[0] + " not found.\n");
  };
  c["compute"] = [this](std::vector<std::string> &a) { return cmd_compute(a); };
  c["dump"] = [this](std::vector<std::string> &a) { return cmd_dump(a); };
  c["dt_factor"] = [this](std::vector<std::string> &a) { // Update::set_dt_factor, src/update.cpp:62-67
    if (a.size() != 1) fatal("Illegal dt_factor command.\n");
    dt_factor = input.parsev(a[0]); return Var(0);
  };
  c["set_dt"] = [this](std::vector<std::string> &a) { // Update::set_dt, src/update.cpp:69-76
    if (a.size() != 1) fatal("Illegal set_dt command.\n");
    dt = input.parsev(a[0]); dt_constant = true; input.vars["dt"] = Var("dt", dt);
    if (ctx) check(kml_set_dt(ctx, dt));
    return Var(0);
  };
  c["set_output"] = [this](std::vector<std::string> &a) { every_log = (int)(double)input.parsev(a[0]); return Var(0); }; // Output::set_log
  c["log"] = [this](std::vector<std::string> &a) { every_log = (int)(double)input.parsev(a[0]); return Var(0); };
  c["log_modify"] = [this](std::vector<std::string> &a) { // Log::modify: log_modify(custom, step, dt, time, vars...)
    if (a.empty() || a[0] != "custom") fatal("Unknown log style.\n");
    log_fields.assign(a.begin() + 1, a.end()); return Var(0);
  };
  c["restart"] = [this](std::vector<std::string> &a) { // Output::create_restart, src/output.cpp:323-350
    if (a.size() < 1) fatal("Illegal restart command: too few arguments.\n");
    if (a.size() > 2) fatal("Illegal restart command: too many arguments.\n");
    restart_every = (int)(double)input.parsev(a[0]);
    if (restart_every == 0) { next_restart = 0; return Var(0); }
    if (a.size() < 2) fatal("Illegal restart command: too few arguments.\n");
    restart_name = a[1];
    next_restart = (ntimestep / restart_every) * restart_every + restart_every;
    return Var(0);
  };
  c["read_restart"] = [this](std::vector<std::string> &a) { // ReadRestart::command, src/read_restart.cpp:36-83
    if (a.size() < 1) fatal("Illegal read command.\n");
    read_restart(a[0]); return Var(0);
  };
  c["write_restart"] = [](std::vector<std::string> &a) { // WriteRestart::command only records the file name (src/write_restart.cpp:36-48);
    if (a.size() < 1) fatal("Illegal write command.\n");  // files are written by restart(N, file), through Output::write
    return Var(0);
  };
  c["run"] = [this](std::vector<std::string> &a) { return cmd_run(a, 0); };
  c["run_time"] = [this](std::vector<std::string> &a) { return cmd_run(a, 1); };
  c["run_until"] = [this](std::vector<std::string> &a) { return cmd_run(a, 2); };
  c["run_while"] = [this](std::vector<std::string> &a) { return cmd_run(a, 3); };
  // CentreOfMass / InternalForce commands (src/centre_of_mass.cpp, src/internal_force.cpp -> Group::xcm / internal_force,
  // src/group.cpp:262-408): lazy variables - the returned Var carries the call as its equation and is re-evaluated,
  // i.e. the state is downloaded again, every time a log line or an expression needs it.
  auto group_sum = [this](std::vector<std::string> &a, bool com) -> Var {
    if (a.size() < 2) fatal("Illegal run command");
    const int ig = find_group(a[0]);
    if (ig == -1) fatal("Error: could not find group named: " + a[0] + "\n");
    int dir; if (a[1] == "x") dir = 0; else if (a[1] == "y") dir = 1; else if (a[1] == "z") dir = 2;
    else fatal("Error: directions should be either x,y or z: " + a[1] + " not understood.\n");
    if (gsolid[ig] == -1) fatal("Error: " + std::string(com ? "xcm" : "internal_force") + " needs a group restricted to one solid (the reference indexes solids[-1] for \"all\").\n");
    SolidH &s = *solids[gsolid[ig]]; const int bit = gbitmask[ig];
    double num = 0, den = 0;
    if (gpon[ig] == "particles") {
      int64_t np = 0; check(kml_solid_np(ctx, s.dev, &np));
      std::vector<double> v(3 * np), m(np); std::vector<int> mask(np);
      if (!com) check(kml_keep_particle_acceleration(ctx)); // f_p = a_p m_p is not state of the engine unless asked for (before the first step)
      check(kml_solid_download(ctx, s.dev, com ? KML_P_X : KML_P_F, v.data()));
      check(kml_solid_download(ctx, s.dev, KML_P_MASK, mask.data()));
      if (com) check(kml_solid_download(ctx, s.dev, KML_P_MASS, m.data()));
      for (int64_t i = 0; i < np; i++) if (mask[i] & bit) { if (com) { num += v[3 * i + dir] * m[i]; den += m[i]; } else num += v[3 * i + dir]; }
    } else {
      GridH &g = *s.grid; std::vector<double> v(3 * g.nnodes), m(g.nnodes);
      check(kml_grid_download(ctx, g.id, com ? KML_N_X : KML_N_F, v.data()));
      if (com) check(kml_grid_download(ctx, g.id, KML_N_MASS, m.data()));
      for (int64_t i = 0; i < g.nnodes; i++) if (g.mask[i] & bit) { if (com) { num += v[3 * i + dir] * m[i]; den += m[i]; } else num += v[3 * i + dir]; }
    }
    const double val = com ? (den ? num / den : 0.0) : num;
    return Var(std::string(com ? "xcm(" : "internal_force(") + a[0] + "," + a[1] + ")", val);
  };
  c["xcm"] = [group_sum](std::vector<std::string> &a) { return group_sum(a, true); };
  c["internal_force"] = [group_sum](std::vector<std::string> &a) { return group_sum(a, false); };
  // ExtForce command (src/external_force.cpp -> Group::external_force, src/group.cpp:410-462): the sum of the particle forces
  // f_p = a_p m_p of a particle group restricted to one solid, as a lazy variable
  c["external_force"] = [this](std::vector<std::string> &a) -> Var {
    if (a.size() < 2) fatal("Illegal run command");
    const int ig = find_group(a[0]);
    if (ig == -1) fatal("Error: could not find group named: " + a[0] + "\n");
    int dir; if (a[1] == "x") dir = 0; else if (a[1] == "y") dir = 1; else if (a[1] == "z") dir = 2;
    else fatal("Error: directions should be either x,y or z: " + a[1] + " not understood.\n");
    if (gpon[ig] == "nodes") fatal("Error: cannot calculate the external forces applied to the node group " + a[0] + ".\n");
    if (gsolid[ig] == -1) fatal("Error: external_force needs a group restricted to one solid (the reference indexes solids[-1] for \"all\").\n");
    SolidH &s = *solids[gsolid[ig]]; const int bit = gbitmask[ig];
    int64_t np = 0; check(kml_solid_np(ctx, s.dev, &np));
    std::vector<double> f(3 * np); std::vector<int> mask(np);
    check(kml_keep_particle_acceleration(ctx));
    check(kml_solid_download(ctx, s.dev, KML_P_F, f.data())); check(kml_solid_download(ctx, s.dev, KML_P_MASK, mask.data()));
    double sum = 0; for (int64_t i = 0; i < np; i++) if (mask[i] & bit) sum += f[3 * i + dir];
    return Var("external_force(" + a[0] + "," + a[1] + ")", sum);
  };
  c["delete_compute"] = [this](std::vector<std::string> &a) { // Modify::delete_compute, src/modify.cpp:218-231
    for (size_t i = 0; i < computes.size(); i++) if (computes[i]->id == a[0]) { computes.erase(computes.begin() + i); return Var(0); }
    fatal("Could not find compute ID to delete.\n");
  };
  // TranslateParticles (src/translate_particles.cpp:28-120): translate_particles(solid | all, region, region-ID, dx, dy, dz) moves the
  // particles currently inside the region, reference positions included
  c["translate_particles"] = [this](std::vector<std::string> &a) -> Var {
    if (a.size() < 6) fatal("Error: not enough arguments.\nUsage: translate_particles(solid-ID, region, region-ID, delx, dely, delz)\n");
    const int isolid = find_solid(a[0]);
    if (isolid < 0 && a[0] != "all") fatal("Error: solid " + a[0] + " unknown.\n");
    if (a[1] != "region") fatal("Error: use of illegal keyword for translate_particles command: " + a[1] + "\n");
    const int ir = find_region(a[2]);
    if (ir < 0) fatal("Error: region " + a[2] + " unknown.\n");
    Var del[3]; bool set[3];
    for (int d = 0; d < 3; d++) { del[d] = input.parsev(a[3 + d]); set[d] = !(del[d].is_constant() && std::fabs(del[d].result()) <= 1.0e-12); }
    for (size_t is = 0; is < solids.size(); is++) {
      if (isolid >= 0 && (int)is != isolid) continue;
      SolidH &S = *solids[is]; sync(S);
      std::vector<double> x(3 * S.np), x0(3 * S.np);
      check(kml_solid_download(ctx, S.dev, KML_P_X, x.data())); check(kml_solid_download(ctx, S.dev, KML_P_X0, x0.data()));
      for (int64_t ip = 0; ip < S.np; ip++) {
        if (regions[ir]->inside(x[3 * ip], x[3 * ip + 1], x[3 * ip + 2]) != 1) continue;
        input.vars["x0"] = Var("x0", x0[3 * ip]); input.vars["y0"] = Var("y0", x0[3 * ip + 1]); input.vars["z0"] = Var("z0", x0[3 * ip + 2]);
        input.vars["x"] = Var("x", x[3 * ip]); input.vars["y"] = Var("y", x[3 * ip + 1]); input.vars["z"] = Var("z", x[3 * ip + 2]);
        for (int d = 0; d < 3; d++) if (set[d]) { const double v = del[d].result(&input); x0[3 * ip + d] += v; x[3 * ip + d] += v; S.x0[ip][d] = x0[3 * ip + d]; }
      }
      check(kml_solid_upload(ctx, S.dev, KML_P_X, x.data())); check(kml_solid_upload(ctx, S.dev, KML_P_X0, x0.data()));
      mirrors_current(S);
    }
    return Var(0);
  };
  // DeleteParticles (src/delete_particles.cpp:30-113): delete_particles(solid | all, region, region-ID) removes the particles whose
  // REFERENCE position lies in the region; survivors keep the reference's swap-with-last order (kml_solid_delete_particles)
  c["delete_particles"] = [this](std::vector<std::string> &a) -> Var {
    if (a.size() < 3) fatal("Error: not enough arguments.\nUsage: delete_particles(solid-ID, region, region-ID)\n");
    const int isolid = find_solid(a[0]);
    if (isolid < 0 && a[0] != "all") fatal("Error: solid " + a[0] + " unknown.\n");
    if (a[1] != "region") fatal("Error: use of illegal keyword for delete_particles command: " + a[1] + "\n");
    const int ir = find_region(a[2]);
    if (ir < 0) fatal("Error: region " + a[2] + " unknown.\n");
    for (size_t is = 0; is < solids.size(); is++) {
      if (isolid >= 0 && (int)is != isolid) continue;
      SolidH &S = *solids[is]; sync(S);
      std::vector<int> dl(S.np);
      for (int64_t ip = 0; ip < S.np; ip++) dl[ip] = regions[ir]->inside(S.x0[ip][0], S.x0[ip][1], S.x0[ip][2]) == 1;
      check(kml_solid_delete_particles(ctx, S.dev, dl.data()));
      int64_t n = S.np, k = 0; // the same compaction on the host copies
      while (k < n) {
        if (dl[k]) { S.x0[k] = S.x0[n - 1]; S.mask[k] = S.mask[n - 1]; S.ptag[k] = S.ptag[n - 1]; dl[k] = dl[n - 1]; n--; } else k++;
      }
      S.np = n; S.x0.resize(n); S.mask.resize(n); S.ptag.resize(n); mirrors_current(S);
      std::vector<double> vol(n), mass(n);
      check(kml_solid_download(ctx, S.dev, KML_P_VOL, vol.data())); check(kml_solid_download(ctx, S.dev, KML_P_MASS, mass.data()));
      S.vtot = S.mtot = 0; for (int64_t i = 0; i < n; i++) { S.vtot += vol[i]; S.mtot += mass[i]; }
      double red[3] = {(double)n, S.vtot, S.mtot}; // decomposed run: every rank deleted its own particles; the totals are sums over the slabs
      check(kml_comm_sum(ctx, red, 3));
      np_total = (int64_t)red[0]; // sic: the reference stores this solid's remaining count (domain->np_total = np_local_reduced, src/delete_particles.cpp:78)
      S.vtot = red[1]; S.mtot = red[2];
      if (!quiet) std::cout << "Solid " << S.id << " new total volume = " << S.vtot << std::endl;
    }
    return Var(0);
  };
  c["plot"] = [](std::vector<std::string> &) { return Var(0); };
  c["save_plot"] = [](std::vector<std::string> &) { return Var(0); };
}

// Update::create_method, src/update.cpp:108-196 (+ ULMPM/TLMPM ctor and setup: src/ulmpm.cpp:32-86, src/tlmpm.cpp:34-86)
Var Sim::cmd_method(std::vector<std::string> &a) {
  if (a.size() < 3) fatal("Illegal method command: not enough arguments.\n");
  size_t n = 0;
  method_type = a[0];
  PIC_FLIP = 0.99; temp = false; ge = false; is_TL = is_CPDI = false;
  if (method_type == "ulmpm") {}
  else if (method_type == "tlmpm") is_TL = true;
  else if (method_type == "ulcpdi") is_CPDI = true;
  else if (method_type == "tlcpdi") { is_TL = true; is_CPDI = true; }
  else fatal("Illegal method style.\n");
  bool isFLIP = false;
  n++;
  static const std::map<std::string, int> subs{{"PIC", KML_SUB_PIC}, {"FLIP", KML_SUB_FLIP}, {"APIC", KML_SUB_APIC}, {"AFLIP", KML_SUB_AFLIP}, {"ASFLIP", KML_SUB_ASFLIP}, {"MLS", KML_SUB_MLS}};
  auto si = subs.find(a[n]);
  if (si == subs.end()) fatal("Error: method type " + a[n] + " not understood. Expect: PIC, FLIP, APIC, AFLIP, or ASFLIP\n");
  sub_method = si->second;
  if (sub_method == KML_SUB_PIC) PIC_FLIP = 0;
  else if (sub_method == KML_SUB_FLIP || sub_method == KML_SUB_AFLIP || sub_method == KML_SUB_ASFLIP) {
    isFLIP = true;
    if (a.size() < 4) fatal("Illegal modify_method command: not enough arguments.\n");
  }
  n++;
  static const std::map<std::string, int> shapes{{"linear", KML_SHAPE_LINEAR}, {"cubic-spline", KML_SHAPE_CUBIC_SPLINE}, {"quadratic-spline", KML_SHAPE_QUADRATIC_SPLINE}, {"Bernstein-quadratic", KML_SHAPE_BERNSTEIN}};
  if (a.size() > n + isFLIP) {
    auto sh = shapes.find(a[n]);
    if (sh == shapes.end()) fatal("Illegal method_method argument: form function of type " + a[n] + " is unknown.\n");
    shape_function = sh->second; n++;
  }
  if (isFLIP) { PIC_FLIP = input.parsev(a[n]); n++; }
  if (a.size() >= n + 1) {
    if (a[n] == "thermo-mechanical") temp = true;
    else if (a[n] == "mechanical") temp = false;
    else fatal("Illegal modify_method command: keyword " + a[n] + " unknown. Expected \"thermo-mechanical\" or \"mechanical\".\n");
  }
  n++;
  if (a.size() >= n + 1 && a[n] == "gradient-enhanced") { ge = true; n++; }
  std::vector<std::string> extra; if (n < a.size()) extra.assign(a.begin() + n, a.end());
  additional_args = extra;
  if (is_CPDI) { // TLCPDI/ULCPDI::setup: style R4 / Q4
    if (!extra.empty()) { if (extra[0] == "R4") cpdi_style = 0; else if (extra[0] == "Q4") cpdi_style = 1; else fatal("Unknown CPDI style " + extra[0]); }
  } else if (!extra.empty()) fatal("Illegal modify_method command: too many arguments.\n");
  // Method::setup: APIC/MLS force PIC_FLIP = 0 (src/ulmpm.cpp:79-85, src/tlmpm.cpp:83-85)
  if (!is_TL && (sub_method == KML_SUB_APIC || sub_method == KML_SUB_MLS)) PIC_FLIP = 0;
  if (is_TL && sub_method == KML_SUB_APIC) PIC_FLIP = 0;
  // ULCPDI::advance_particles blends with its OWN member `FLIP` (src/ulcpdi.cpp:468, declared src/ulcpdi.h:31), which no code
  // ever assigns: the ratio given in the script is ignored and the freshly allocated Method object holds 0.0, i.e. the
  // reference's ULCPDI is pure PIC whatever the script says (verified against the reference build: bit-identical with 0).
  PIC_FLIP_script = PIC_FLIP; // Update::PIC_FLIP as the reference keeps (and writes to restart files) it
  if (method_type == "ulcpdi") PIC_FLIP = 0;
  method_set = true;
  return Var(0);
}

Var Sim::cmd_scheme(std::vector<std::string> &a) { // Update::create_scheme, src/update.cpp:83-102
  if (a.size() < 1) fatal("Illegal scheme command: not enough arguments.\n");
  if (a[0] != "usl" && a[0] != "musl" && a[0] != "usf") fatal("Illegal scheme style.\n");
  scheme_style = a[0]; return Var(0);
}

void Sim::ensure_ctx() {
  if (ctx) return;
  kml_config cfg; memset(&cfg, 0, sizeof cfg);
  cfg.dimension = dimension; cfg.is_TL = is_TL; cfg.is_CPDI = is_CPDI; cfg.cpdi_style = cpdi_style;
  cfg.shape_function = shape_function; cfg.sub_method = sub_method; cfg.PIC_FLIP = PIC_FLIP;
  cfg.axisymmetric = axisymmetric; cfg.temp = temp; cfg.ge = ge;
  for (int d = 0; d < 3; d++) { cfg.boxlo[d] = boxlo[d]; cfg.boxhi[d] = boxhi[d]; }
  cfg.device = device; cfg.rank = rank; cfg.nranks = nranks;
  check(kml_create(&cfg, &ctx));
  check(kml_set_dt(ctx, dt));
  if (nranks > 1) {
    if (is_TL) fatal("multi-GPU runs cover ULMPM (slab decomposition of the background grid)\n");
    if (nccl_id.size() == 128) check(kml_comm_init(ctx, nccl_id.data())); // absent in host-only (CPU) tests of the partition
  }
}

// Grid::init on one rank, src/grid.cpp:68-264 (node counts) - the device creates the nodes themselves.
void Sim::init_grid(GridH &g, const double *solidlo, const double *solidhi) {
  double h = g.cellsize;
  if (shape_function == KML_SHAPE_BERNSTEIN) h /= 2; // src/grid.cpp:82-85
  const double *boundlo = is_TL ? solidlo : boxlo, *boundhi = is_TL ? solidhi : boxhi;
  double Loffsetlo[3], Loffsethi_[3]; int noffsetlo[3], noffsethi_[3];
  for (int d = 0; d < 3; d++) {
    Loffsetlo[d] = dmax(0.0, sublo[d] - boundlo[d]);
    Loffsethi_[d] = dmax(0.0, dmin(subhi[d], boundhi[d]) - boundlo[d]);
    noffsetlo[d] = (int)ceil(Loffsetlo[d] / h);
    noffsethi_[d] = (int)ceil(Loffsethi_[d] / h);
  }
  double Lx = solidhi[0] - solidlo[0];
  g.nx_global = ((int)(Lx / h)) + 1;
  while (g.nx_global * h <= Lx + 0.5 * h) g.nx_global++;
  if (dimension >= 2) { double Ly = solidhi[1] - solidlo[1]; g.ny_global = ((int)Ly) / h + 1; while (g.ny_global * h <= Ly + 0.5 * h) g.ny_global++; } // cast binds to Ly, src/grid.cpp:158
  else g.ny_global = 1;
  if (dimension == 3) { double Lz = solidhi[2] - solidlo[2]; g.nz_global = ((int)Lz) / h + 1; while (g.nz_global * h <= Lz + 0.5 * h) g.nz_global++; }
  else g.nz_global = 1;
  int nx = std::max(0, noffsethi_[0] - noffsetlo[0]);
  int ny = dimension >= 2 ? std::max(0, noffsethi_[1] - noffsetlo[1]) : 1;
  int nz = dimension >= 3 ? std::max(0, noffsethi_[2] - noffsetlo[2]) : 1;
  while (boundlo[0] + h * (noffsetlo[0] + nx - 0.5) < dmin(subhi[0], boundhi[0])) nx++;
  while (boundlo[1] + h * (noffsetlo[1] + ny - 0.5) < dmin(subhi[1], boundhi[1])) ny++;
  while (boundlo[2] + h * (noffsetlo[2] + nz - 0.5) < dmin(subhi[2], boundhi[2])) nz++;
  if (nx != g.nx_global || ny != g.ny_global || nz != g.nz_global || noffsetlo[0] || noffsetlo[1] || noffsetlo[2])
    fatal("Grid::init: local node counts (" + std::to_string(nx) + "," + std::to_string(ny) + "," + std::to_string(nz) + ") differ from the global ones (" +
          std::to_string(g.nx_global) + "," + std::to_string(g.ny_global) + "," + std::to_string(g.nz_global) + "); the reference indexes out of bounds here.\n");
  g.desc.lo[0] = boundlo[0]; g.desc.lo[1] = dimension >= 2 ? boundlo[1] : 0; g.desc.lo[2] = dimension == 3 ? boundlo[2] : 0;
  g.desc.h = h; g.desc.cellsize = g.cellsize; g.desc.n[0] = nx; g.desc.n[1] = ny; g.desc.n[2] = nz;
  g.nnodes = (int64_t)nx * ny * nz;
  ensure_ctx();
  if (nranks > 1 && !is_TL) { grid_pending = true; return; } // the slab is cut when the first solid's particles are known
  check(kml_grid_create(ctx, &g.desc, &g.id));
  g.mask.assign(g.nnodes, 1);
}

// This rank's slab of the background grid: node planes [base_lo, base_hi + span - 1) of the global grid.
void Sim::create_device_grid(GridH &g, int base_lo, int base_hi) {
  const int span = shape_function == KML_SHAPE_LINEAR ? 2 : 4;
  const int nxg = g.nx_global;
  const int p0 = std::max(base_lo, 0), p1 = std::min(base_hi + span - 1, nxg);
  if (base_hi - base_lo < span - 1 && rank < nranks - 1) fatal("slab decomposition: a slab is thinner than the stencil overlap; use fewer GPUs\n");
  g.desc.n[0] = p1 - p0; g.desc.goff = p0; g.desc.gn = nxg;
  g.desc.own_lo = 0; g.desc.own_hi = rank == nranks - 1 ? p1 - p0 : std::max(0, base_hi - p0);
  g.desc.base_lo = base_lo; g.desc.base_hi = base_hi;
  g.nnodes = (int64_t)g.desc.n[0] * g.desc.n[1] * g.desc.n[2];
  check(kml_grid_create(ctx, &g.desc, &g.id));
  g.mask.assign(g.nnodes, 1);
  grid_pending = false;
}

// Domain::set_dimension, src/domain.cpp:496-551 (+ set_local_box on a 1x1x1 proc grid, src/domain.cpp:366-394)
Var Sim::cmd_dimension(std::vector<std::string> &a) {
  if (!method_set) fatal("Error: a method should be defined before calling dimension()!\n");
  if (a.empty()) fatal("Error: dimension did not receive enough arguments\n");
  size_t m = 0;
  int dim = (int)(double)input.parsev(a[m]);
  if (dim != 1 && dim != 2 && dim != 3) fatal("Error: dimension argument: " + a[m] + "\n.");
  dimension = dim;
  // Nargs_dimension = {1: 4, 2: 6, 3: 8} (src/domain.h:93-94); TL takes the cell size too but ignores it
  if (a.size() < (size_t)(2 * dim + 2)) fatal("Error: not enough arguments.\n");
  if (a.size() > (size_t)(2 * dim + 2)) fatal("Error: too many arguments.\n");
  boxlo[0] = input.parsev(a[++m]); boxhi[0] = input.parsev(a[++m]);
  if (dim > 1) { boxlo[1] = input.parsev(a[++m]); boxhi[1] = input.parsev(a[++m]); }
  if (dim == 3) { boxlo[2] = input.parsev(a[++m]); boxhi[2] = input.parsev(a[++m]); }
  double cs = 0;
  if (!is_TL) { cs = input.parsev(a[++m]); if (cs < 0) fatal("Error: cellsize negative!\n"); }
  // Universe::set_proc_grid stops every 1-D run (src/universe.cpp:645-647); the engine's 1-D kernels stay reachable through the C ABI
  if (dim == 1) fatal("New partitioning not supported for dimensions 1 yet!\n");
  for (int d = 0; d < 3; d++) {
    if (d < dim) { double h = (boxhi[d] - boxlo[d]) / 1; sublo[d] = 0 * h + boxlo[d]; subhi[d] = sublo[d] + h; }
    else sublo[d] = subhi[d] = 0;
  }
  if (!is_TL) { grid.reset(new GridH()); grid->cellsize = cs; init_grid(*grid, boxlo, boxhi); }
  created = true;
  return Var(0);
}

namespace {
template <class T> void put(std::ostream &os, const T &v) { os.write(reinterpret_cast<const char *>(&v), sizeof(T)); }
struct Block : Region {
  int inside(double x, double y, double z) const override { return x >= lim[0] && x <= lim[1] && y >= lim[2] && y <= lim[3] && z >= lim[4] && z <= lim[5]; }
  void write_restart(std::ostream &os) const override { for (int k = 0; k < 6; k++) put(os, lim[k]); } // src/region_block.cpp:170-177
  bool to_kml(kml_region &r) const override { r.style = KML_REGION_BLOCK; r.interior = interior; r.axis = 0; r.pad_ = 0; for (int k = 0; k < 6; k++) r.p[k] = lim[k]; return true; }
};
struct Cylinder : Region {
  char axis = 'z'; double c1 = 0, c2 = 0, R = 0, RSq = 0, lo = 0, hi = 0;
  int inside(double x, double y, double z) const override { // src/region_cylinder.cpp:140-161
    double dSq;
    if (axis == 'x') { dSq = (y - c1) * (y - c1) + (z - c2) * (z - c2); return x >= lo && x <= hi && dSq <= RSq; }
    if (axis == 'y') { dSq = (x - c1) * (x - c1) + (z - c2) * (z - c2); return y >= lo && y <= hi && dSq <= RSq; }
    dSq = (x - c1) * (x - c1) + (y - c2) * (y - c2); return z >= lo && z <= hi && dSq <= RSq;
  }
  void write_restart(std::ostream &os) const override { // src/region_cylinder.cpp:184-197
    put(os, c1); put(os, c2); put(os, R); put(os, lo); put(os, hi); put(os, axis); for (int k = 0; k < 6; k++) put(os, lim[k]);
  }
  bool to_kml(kml_region &r) const override {
    r.style = KML_REGION_CYLINDER; r.interior = interior; r.axis = axis == 'x' ? 0 : (axis == 'y' ? 1 : 2); r.pad_ = 0;
    r.p[0] = c1; r.p[1] = c2; r.p[2] = RSq; r.p[3] = lo; r.p[4] = hi; r.p[5] = 0; return true;
  }
};
struct Sphere : Region {
  double c1 = 0, c2 = 0, c3 = 0, R = 0, RSq = 0;
  int inside(double x, double y, double z) const override { return (x - c1) * (x - c1) + (y - c2) * (y - c2) + (z - c3) * (z - c3) <= RSq; }
  void write_restart(std::ostream &os) const override { put(os, c1); put(os, c2); put(os, c3); put(os, R); for (int k = 0; k < 6; k++) put(os, lim[k]); } // src/region_sphere.cpp:145-156
  bool to_kml(kml_region &r) const override {
    r.style = KML_REGION_SPHERE; r.interior = interior; r.axis = 0; r.pad_ = 0; r.p[0] = c1; r.p[1] = c2; r.p[2] = c3; r.p[3] = RSq; r.p[4] = r.p[5] = 0; return true;
  }
};
} // namespace

namespace {
// Union / Intersection / Difference of earlier regions (src/region_union.cpp, src/region_intersection.cpp, src/region_difference.cpp)
struct Composite : Region {
  int kind = 0; // 0 union, 1 intersection, 2 difference
  std::vector<int> ir; const std::vector<std::unique_ptr<Region>> *all = nullptr;
  int inside(double x, double y, double z) const override {
    if (kind == 0) { for (int i : ir) if ((*all)[i]->inside(x, y, z) == 1) return 1; return 0; }
    if (kind == 1) { for (int i : ir) if ((*all)[i]->inside(x, y, z) == 0) return 0; return 1; }
    return (*all)[ir[0]]->inside(x, y, z) == 1 && (*all)[ir[1]]->inside(x, y, z) == 0;
  }
  void write_restart(std::ostream &os) const override { // src/region_union.cpp:98-110 (the three styles share the layout)
    const size_t n = ir.size(); put(os, n); for (int i : ir) put(os, i); for (int k = 0; k < 6; k++) put(os, lim[k]);
  }
};
} // namespace

// Domain::add_region + Block_/Cylinder/Sphere constructors (src/domain.cpp:76-98, src/region_block.cpp:28-139,
// src/region_cylinder.cpp:30-134, src/region_sphere.cpp:30-118)
Var Sim::cmd_region(std::vector<std::string> &a) {
  if (a.size() < 3) fatal("Error: not enough arguments.\n");
  if (find_region(a[0]) >= 0) fatal("Error: reuse of region ID.\n");
  const int dim = dimension;
  auto is_inf = [](const std::string &s) { return s == "INF" || s == "-INF" || s == "+INF" || s == "EDGE"; };
  auto opts = [&](Region &r, size_t from) { for (size_t i = from; i < a.size(); i++) if (a[i] == "exterior") r.interior = 0; }; // src/region.cpp:30-62
  std::unique_ptr<Region> reg;
  if (a[1] == "block") {
    size_t need = dim == 3 ? 8 : (dim == 2 ? 6 : 4);
    if (a.size() < need) fatal("Error: region command not enough arguments.\n");
    auto b = new Block(); reg.reset(b); opts(*b, need);
    for (int d = 0; d < dim; d++) {
      const std::string &slo = a[2 + 2 * d], &shi = a[3 + 2 * d];
      if (is_inf(slo)) { if (regions.empty()) fatal("Cannot use region INF or EDGE when box does not exist.\n"); b->lim[2 * d] = -BIG; }
      else { b->lim[2 * d] = input.parsev(slo); if (boxlo[d] > b->lim[2 * d]) boxlo[d] = b->lim[2 * d]; }
      if (is_inf(shi)) { if (regions.empty()) fatal("Cannot use region INF or EDGE when box does not exist.\n"); b->lim[2 * d + 1] = BIG; }
      else { b->lim[2 * d + 1] = input.parsev(shi); if (boxhi[d] < b->lim[2 * d + 1]) boxhi[d] = b->lim[2 * d + 1]; }
    }
    if (b->lim[0] > b->lim[1] || b->lim[2] > b->lim[3] || b->lim[4] > b->lim[5]) fatal("Illegal region block command.\n");
  } else if (a[1] == "cylinder") {
    auto c = new Cylinder(); reg.reset(c);
    if (dim == 3) {
      if (a.size() < 8) fatal("Error: not enough arguments.\n");
      opts(*c, 8);
      if (a[2] == "x" || a[2] == "y" || a[2] == "z") c->axis = a[2][0]; else fatal("Error: region cylinder axis not understood, expect x, y, or z, received " + a[2] + ".\n");
      c->c1 = input.parsev(a[3]); c->c2 = input.parsev(a[4]); c->R = input.parsev(a[5]);
      if (a[6] == "INF" || a[6] == "EDGE") { if (regions.empty()) fatal("Cannot use region INF or EDGE when box does not exist.\n"); c->lo = -BIG; } else c->lo = input.parsev(a[6]);
      if (a[7] == "INF" || a[7] == "EDGE") { if (regions.empty()) fatal("Cannot use region INF or EDGE when box does not exist.\n"); c->hi = BIG; } else c->hi = input.parsev(a[7]);
    } else {
      if (a.size() < 5) fatal("Error: not enough arguments.\n");
      opts(*c, 5);
      c->axis = 'z'; c->c1 = input.parsev(a[2]); c->c2 = input.parsev(a[3]); c->R = input.parsev(a[4]); c->lo = c->hi = 0;
    }
    c->RSq = c->R * c->R;
    if (c->lo > c->hi) fatal("Illegal region cylinder command: low is higher than high.\n");
    double *l = c->lim;
    if (c->axis == 'x') { l[0] = c->lo; l[1] = c->hi; l[2] = c->c1 - c->R; l[3] = c->c1 + c->R; l[4] = c->c2 - c->R; l[5] = c->c2 + c->R; }
    else if (c->axis == 'y') { l[0] = c->c1 - c->R; l[1] = c->c1 + c->R; l[2] = c->lo; l[3] = c->hi; l[4] = c->c2 - c->R; l[5] = c->c2 + c->R; }
    else { l[0] = c->c1 - c->R; l[1] = c->c1 + c->R; l[2] = c->c2 - c->R; l[3] = c->c2 + c->R; l[4] = c->lo; l[5] = c->hi; }
  } else if (a[1] == "sphere") {
    auto s = new Sphere(); reg.reset(s);
    if (a.size() != (size_t)(dim + 3)) fatal("Error: region sphere: wrong number of arguments.\n");
    size_t i = 2;
    s->c1 = input.parsev(a[i++]); if (dim >= 2) s->c2 = input.parsev(a[i++]); if (dim == 3) s->c3 = input.parsev(a[i++]);
    double R = input.parsev(a[i++]);
    if (R < 0) fatal("Error: R cannot be negative.\n");
    s->R = R; s->RSq = R * R;
    double *l = s->lim; l[0] = s->c1 - R; l[1] = s->c1 + R; l[2] = s->c2 - R; l[3] = s->c2 + R; l[4] = s->c3 - R; l[5] = s->c3 + R;
    if (method_type == "tlmpm") {
      for (int d = 0; d < (dim == 3 ? 3 : 2); d++) { if (boxlo[d] > l[2 * d]) boxlo[d] = l[2 * d]; if (boxhi[d] < l[2 * d + 1]) boxhi[d] = l[2 * d + 1]; }
    } else {
      for (int d = 0; d < (dim == 3 ? 3 : 2); d++) if (boxlo[d] > l[2 * d]) l[2 * d] = boxlo[d];
    }
  } else if (a[1] == "union" || a[1] == "intersection" || a[1] == "difference") {
    auto c = new Composite(); reg.reset(c); c->all = &regions; c->kind = a[1] == "union" ? 0 : (a[1] == "intersection" ? 1 : 2);
    if (a.size() < 4) fatal("Error: region_" + a[1] + " command not enough arguments\n");
    if (c->kind == 2 && a.size() > 4) fatal("Error: region_difference command too many arguments\n");
    double *l = c->lim;
    for (int k = 0; k < 6; k++) l[k] = (c->kind == 0) == (k % 2 == 0) ? BIG : -BIG; // union starts from an empty box, intersection from everything
    for (size_t i = 2; i < a.size(); i++) {
      const int r = find_region(a[i]);
      if (r == -1) fatal("Error: region " + a[i] + " does not exist.\n");
      c->ir.push_back(r);
      const double *m = regions[r]->lim;
      if (c->kind == 2) { if (i == 2) for (int k = 0; k < 6; k++) l[k] = m[k]; continue; } // the box of the first region
      for (int k = 0; k < 6; k++) { const bool lo = k % 2 == 0; if (c->kind == 0 ? (lo ? l[k] > m[k] : l[k] < m[k]) : (lo ? l[k] < m[k] : l[k] > m[k])) l[k] = m[k]; }
    }
  } else fatal("Unknown region style " + a[1] + "\n");
  reg->id = a[0]; reg->style = a[1];
  regions.push_back(std::move(reg));
  if (ctx) check(kml_set_domain_box(ctx, boxlo, boxhi));
  return Var(0);
}

// ReadRestart::command (src/read_restart.cpp:36-83) and the read_restart members it calls: Update (src/update.cpp:279-349), Domain
// (src/domain.cpp:635-721), regions, Material (src/material.cpp:499-612), Solid (src/solid.cpp:2889-2957), Group (src/group.cpp:508-600),
// Modify (src/modify.cpp:335-366).  Objects whose constructors derive values (materials, the domain, groups) are rebuilt through the same
// command code as a script would use, with every number handed over as a constant variable so that no digit passes through the parser.
namespace {
template <class T> T rget(std::istream &is) { T v{}; is.read(reinterpret_cast<char *>(&v), sizeof(T)); return v; }
std::string rget_str(std::istream &is) { const size_t n = rget<size_t>(is); if (n > (1u << 20)) fatal("read_restart: corrupt string length\n"); std::string t(n, '\0'); is.read(&t[0], (std::streamsize)n); return t; }
} // namespace

void Sim::read_restart(const std::string &pattern) {
  std::string fn = pattern; const size_t star = fn.find('*');
  if (star != std::string::npos) fn = fn.substr(0, star) + (nranks > 1 ? "proc-" + std::to_string(rank) + "." : "") + fn.substr(star + 1);
  if (!quiet && rank == 0) std::cout << "read " << fn << std::endl;
  std::ifstream is(fn, std::ios::in | std::ios::binary);
  if (!is) fatal("Error: cannot read in file: " + fn + ".\n");
  if (nranks > 1) fatal("read_restart on a decomposed run is not supported\n");
  int nvar = 0;
  auto num = [&](double v) { const std::string name = "__rst" + std::to_string(nvar++); input.vars[name] = Var(v); return name; }; // exact constant
  auto call = [&](const char *cmd, std::vector<std::string> a) { input.commands.at(cmd)(a); };
  // header
  for (int flag = rget<int>(is); flag >= 0 && is; flag = rget<int>(is)) {
    if (flag == 0) { const std::string v = rget_str(is); if (!quiet) std::cout << "version = " << v << std::endl; }
    else if (flag == 1) dimension = rget<int>(is);
    else if (flag == 2) { const int np = rget<int>(is); if (np != nranks) fatal("Restart file written for " + std::to_string(np) + " CPUs.\n"); }
  }
  // Update
  const std::string mtype = rget_str(is); const bool temp_ = rget<bool>(is); const std::string scheme = rget_str(is);
  const int sub = rget<int>(is); const double pf = rget<double>(is); const int shape = rget<int>(is);
  std::vector<std::string> extra(rget<size_t>(is)); for (auto &e : extra) e = rget_str(is);
  const double atime_ = rget<double>(is); const int64_t nt = rget<int64_t>(is); const double dt_ = rget<double>(is), dtf = rget<double>(is);
  const bool dtc = rget<bool>(is);
  static const char *subn[] = {"PIC", "FLIP", "APIC", "AFLIP", "ASFLIP", "MLS"}, *shn[] = {"linear", "cubic-spline", "quadratic-spline", "Bernstein-quadratic"};
  if (sub < 0 || sub > 5 || shape < 0 || shape > 3) fatal("read_restart: unknown method enumerators\n");
  { std::vector<std::string> a{mtype, subn[sub], shn[shape]};
    if (sub == KML_SUB_FLIP || sub == KML_SUB_AFLIP || sub == KML_SUB_ASFLIP) a.push_back(num(pf));
    a.push_back(temp_ ? "thermo-mechanical" : "mechanical");
    a.insert(a.end(), extra.begin(), extra.end());
    call("method", a); }
  call("scheme", {scheme});
  // Domain
  double lo[3], hi[3];
  for (double &v : lo) v = rget<double>(is);
  for (double &v : hi) v = rget<double>(is);
  for (int k = 0; k < 6; k++) rget<double>(is); // sublo / subhi: one rank, recomputed
  const bool axi = rget<bool>(is); const int64_t np_total_ = rget<int64_t>(is);
  if (axi) call("axisymmetric", {"true"});
  { std::vector<std::string> a{std::to_string(dimension)};
    for (int d = 0; d < dimension; d++) { a.push_back(num(lo[d])); a.push_back(num(hi[d])); }
    a.push_back(is_TL ? num(0.0) : num(rget<double>(is))); // TL: dimension() takes and ignores a cell size
    call("dimension", a); }
  for (int n = rget<int>(is), i = 0; i < n; i++) { // regions: restart constructors do not touch the box (src/region_block.cpp:35-42)
    const std::string id = rget_str(is), style = rget_str(is);
    std::unique_ptr<Region> reg;
    if (style == "block") { auto b = new Block(); reg.reset(b); for (double &v : b->lim) v = rget<double>(is); }
    else if (style == "cylinder") {
      auto c = new Cylinder(); reg.reset(c);
      c->c1 = rget<double>(is); c->c2 = rget<double>(is); c->R = rget<double>(is); c->lo = rget<double>(is); c->hi = rget<double>(is); c->axis = rget<char>(is);
      for (double &v : c->lim) v = rget<double>(is);
      c->RSq = c->R * c->R;
    } else if (style == "sphere") {
      auto sp = new Sphere(); reg.reset(sp);
      sp->c1 = rget<double>(is); sp->c2 = rget<double>(is); sp->c3 = rget<double>(is); sp->R = rget<double>(is); for (double &v : sp->lim) v = rget<double>(is);
      sp->RSq = sp->R * sp->R;
    } else if (style == "union" || style == "intersection" || style == "difference") {
      auto c = new Composite(); reg.reset(c); c->all = &regions; c->kind = style == "union" ? 0 : (style == "intersection" ? 1 : 2);
      c->ir.resize(rget<size_t>(is)); for (int &r : c->ir) r = rget<int>(is);
      for (double &v : c->lim) v = rget<double>(is);
    } else fatal("read_restart: region style " + style + " is not supported\n");
    reg->id = id; reg->style = style; regions.push_back(std::move(reg));
  }
  // Material
  for (int n = rget<int>(is), i = 0; i < n; i++) {
    const std::string id = rget_str(is), style = rget_str(is);
    const double rho0 = rget<double>(is), K = rget<double>(is);
    if (style == "linear") call("eos", {id, style, num(rho0), num(K)});
    else if (style == "shock") { const double c0 = rget<double>(is), S = rget<double>(is), G = rget<double>(is), Tr = rget<double>(is), cv = rget<double>(is), Q1 = rget<double>(is), Q2 = rget<double>(is);
      call("eos", {id, style, num(rho0), num(K), num(c0), num(S), num(G), num(cv), num(Tr), num(Q1), num(Q2)}); }
    else if (style == "fluid") call("eos", {id, style, num(rho0), num(K), num(rget<double>(is))});
    else fatal("read_restart: EOS style " + style + " unknown\n");
  }
  for (int n = rget<int>(is), i = 0; i < n; i++) {
    const std::string id = rget_str(is), style = rget_str(is); const double G = rget<double>(is);
    if (style == "linear" || style == "fluid") call("strength", {id, style, num(G)});
    else if (style == "plastic") call("strength", {id, style, num(G), num(rget<double>(is))});
    else if (style == "johnson_cook") { const double A = rget<double>(is), B = rget<double>(is), n_ = rget<double>(is), m = rget<double>(is), e0 = rget<double>(is), C = rget<double>(is), Tr = rget<double>(is), Tm = rget<double>(is);
      call("strength", {id, style, num(G), num(A), num(B), num(n_), num(e0), num(C), num(m), num(Tr), num(Tm)}); }
    else if (style == "swift") { const double A = rget<double>(is), B = rget<double>(is), C = rget<double>(is), n_ = rget<double>(is); call("strength", {id, style, num(G), num(A), num(B), num(C), num(n_)}); }
    else fatal("read_restart: strength style " + style + " unknown\n");
  }
  for (int n = rget<int>(is), i = 0; i < n; i++) {
    const std::string id = rget_str(is), style = rget_str(is); std::vector<std::string> a{id, style};
    for (int k = 0; k < 8; k++) a.push_back(num(rget<double>(is))); // d1..d5, epsdot0, Tr, Tm
    call("damage", a);
  }
  for (int n = rget<int>(is), i = 0; i < n; i++) {
    const std::string id = rget_str(is), style = rget_str(is);
    const double chi = rget<double>(is), kappa = rget<double>(is), cp = rget<double>(is), alpha = rget<double>(is), T0 = rget<double>(is), Tm = rget<double>(is);
    call("temperature", {id, style, num(chi), num(cp), num(kappa), num(alpha), num(T0), num(Tm)});
  }
  for (int n = rget<int>(is), i = 0; i < n; i++) {
    const std::string id = rget_str(is); const int type = rget<int>(is);
    if (type == 3) {
      const int ie = rget<int>(is), ist = rget<int>(is), id_ = rget<int>(is), it = rget<int>(is);
      std::vector<std::string> a{id, "eos-strength", eoss.at(ie).id, strengths.at(ist).id};
      if (id_ >= 0) a.push_back(damages.at(id_).id);
      if (it >= 0) a.push_back(temperatures.at(it).id);
      call("material", a);
    } else if (type == 1 || type == 2) {
      const double rho0 = rget<double>(is), E = rget<double>(is), nu = rget<double>(is), cp = rget<double>(is), kappa = rget<double>(is);
      std::vector<std::string> a{id, type == 1 ? "linear" : "neo-hookean", num(rho0), num(E), num(nu)};
      if (cp != 0 || kappa != 0) { a.push_back(num(cp)); a.push_back(num(kappa)); }
      call("material", a);
    } else fatal("read_restart: the reference's writer stores nothing for a rigid material, its reader expects a density (src/material.cpp:478-484, :598-601): such files cannot be read\n");
  }
  if (rget<int>(is) != -2) fatal("Error: unexpected end to Material::read_restart(). Number of read entities unexpected.\n");
  // solids
  for (int n = rget<int>(is), i = 0; i < n; i++) {
    std::unique_ptr<SolidH> sp(new SolidH()); SolidH &S = *sp; S.id = rget_str(is);
    if (is_TL) { S.own_grid.reset(new GridH()); S.grid = S.own_grid.get(); } else S.grid = grid.get();
    for (double &v : S.solidlo) v = rget<double>(is);
    for (double &v : S.solidhi) v = rget<double>(is);
    for (int k = 0; k < 6; k++) rget<double>(is); // solidsublo / solidsubhi
    S.np_created = rget<int64_t>(is); S.np = rget<int>(is); const int nc = rget<int>(is); (void)nc;
    S.mat = rget<int>(is); const double cs = rget<double>(is);
    if (is_CPDI) fatal("read_restart: the layout holds no CPDI particle domains (src/solid.cpp:2889-2957)\n");
    if (S.mat < 0 || S.mat >= (int)materials.size()) fatal("read_restart: bad material index\n");
    if (is_TL) { S.grid->cellsize = cs; init_grid(*S.grid, S.solidlo, S.solidhi); }
    const int64_t np = S.np;
    std::vector<double> x(3 * np), v(3 * np), sig(9 * np), eel(9 * np), pk1(is_TL ? 9 * np : 0), F(9 * np), vol0(np), vol(np), rho0(np), mass(np), ep(np), epd(np), dmg(np), dmgi(np), T(temp ? np : 0), ie(np);
    S.x0.resize(np); S.mask.resize(np); S.ptag.resize(np);
    auto get_mat = [&](double *m) { for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) m[3 * r + c] = rget<double>(is); }; // Eigen column-major -> rows
    for (int64_t ip = 0; ip < np; ip++) {
      S.ptag[ip] = rget<int64_t>(is);
      for (int d = 0; d < 3; d++) S.x0[ip][d] = rget<double>(is);
      for (int d = 0; d < 3; d++) x[3 * ip + d] = rget<double>(is);
      for (int d = 0; d < 3; d++) v[3 * ip + d] = rget<double>(is);
      get_mat(&sig[9 * ip]); get_mat(&eel[9 * ip]); if (is_TL) get_mat(&pk1[9 * ip]); get_mat(&F[9 * ip]);
      const double J = rget<double>(is); vol0[ip] = rget<double>(is); vol[ip] = J * vol0[ip]; rho0[ip] = rget<double>(is); mass[ip] = rho0[ip] * vol0[ip];
      ep[ip] = rget<double>(is); epd[ip] = rget<double>(is); dmg[ip] = rget<double>(is); dmgi[ip] = rget<double>(is);
      if (temp) T[ip] = rget<double>(is);
      ie[ip] = rget<double>(is); S.mask[ip] = rget<int>(is);
    }
    if (!is) fatal("read_restart: unexpected end of file in solid " + S.id + "\n");
    kml_solid_desc d; memset(&d, 0, sizeof d); d.np = np; d.capacity = np; d.grid = S.grid->id; d.mat = materials[S.mat].km;
    check(kml_solid_create(ctx, &d, &S.dev));
    check(kml_solid_upload(ctx, S.dev, KML_P_PTAG, S.ptag.data())); check(kml_solid_upload(ctx, S.dev, KML_P_MASK, S.mask.data()));
    check(kml_solid_upload(ctx, S.dev, KML_P_X, x.data())); check(kml_solid_upload(ctx, S.dev, KML_P_X0, S.x0.data())); check(kml_solid_upload(ctx, S.dev, KML_P_V, v.data()));
    check(kml_solid_upload(ctx, S.dev, KML_P_SIGMA, sig.data())); check(kml_solid_upload(ctx, S.dev, KML_P_STRAIN_EL, eel.data()));
    if (is_TL) check(kml_solid_upload(ctx, S.dev, KML_P_VOL0PK1, pk1.data()));
    check(kml_solid_upload(ctx, S.dev, KML_P_FDEF, F.data()));
    check(kml_solid_upload(ctx, S.dev, KML_P_VOL0, vol0.data())); check(kml_solid_upload(ctx, S.dev, KML_P_VOL, vol.data()));
    check(kml_solid_upload(ctx, S.dev, KML_P_RHO0, rho0.data())); check(kml_solid_upload(ctx, S.dev, KML_P_MASS, mass.data()));
    check(kml_solid_upload(ctx, S.dev, KML_P_EFF_PLASTIC_STRAIN, ep.data())); check(kml_solid_upload(ctx, S.dev, KML_P_EFF_PLASTIC_STRAIN_RATE, epd.data()));
    check(kml_solid_upload(ctx, S.dev, KML_P_DAMAGE, dmg.data())); check(kml_solid_upload(ctx, S.dev, KML_P_DAMAGE_INIT, dmgi.data()));
    if (temp) check(kml_solid_upload(ctx, S.dev, KML_P_T, T.data()));
    check(kml_solid_upload(ctx, S.dev, KML_P_IENERGY, ie.data()));
    mirrors_current(S);
    S.vtot = S.mtot = 0; for (int64_t ip = 0; ip < np; ip++) { S.vtot += vol[ip]; S.mtot += mass[ip]; }
    solids.push_back(std::move(sp));
  }
  np_total = np_total_;
  // groups: re-assigned from region and reference position like Group::read_restart does
  const int ng = rget<int>(is);
  for (int ig = 1; ig < ng; ig++) {
    const std::string name = rget_str(is); const int bit = rget<int>(is); const bool nodes = rget<bool>(is); const int isolid = rget<int>(is), ireg = rget<int>(is); (void)bit;
    if (ireg < 0 || ireg >= (int)regions.size()) fatal("Error: could not find region with ID " + std::to_string(ireg) + ".\n");
    std::vector<std::string> a{name, nodes ? "nodes" : "particles", "region", regions[ireg]->id};
    if (isolid == -1) a.push_back("all"); else { a.push_back("solid"); a.push_back(solids.at(isolid)->id); }
    call("group", a);
  }
  // fixes
  for (size_t n = rget<size_t>(is), i = 0; i < n; i++) {
    const std::string id = rget_str(is), style = rget_str(is); const int ig = rget<int>(is);
    fixes.push_back(fix_from_restart(id, style, ig, is));
  }
  if (!is) fatal("read_restart: unexpected end of file\n");
  // Update::read_restart: time, step, dt
  restarted_TL = is_TL; // Domain::np_local is not restored (src/domain.cpp:635-721): TLMPM refuses the next run (src/tlmpm.cpp:91-99)
  atime = atime_; ntimestep = nt; atimestep = nt; dt = dt_; dt_factor = dtf; dt_constant = dtc;
  input.vars["time"] = Var("time", atime); input.vars["timestep"] = Var("timestep", (double)ntimestep); input.vars["dt"] = Var("dt", dt);
  check(kml_set_dt(ctx, dt));
}

// Material::add_EOS + EOS ctors (src/material.cpp:94-123, src/eos_linear.cpp, src/eos_shock.cpp:30-82, src/eos_fluid.cpp)
Var Sim::cmd_eos(std::vector<std::string> &a) {
  if (a.size() < 3) fatal("Error: not enough arguments.\n");
  for (auto &e : eoss) if (e.id == a[0]) fatal("Error: reuse of EOS ID.\n");
  EOSH e{}; e.id = a[0];
  auto P = [&](size_t i) { return (double)input.parsev(a[i]); };
  if (a[1] == "linear") { if (a.size() < 4) fatal("Error: not enough arguments.\n"); e.type = KML_EOS_LINEAR; e.rho0 = P(2); e.K = P(3); }
  else if (a[1] == "shock") {
    if (a.size() < 11) fatal("Error: not enough arguments.\n");
    e.type = KML_EOS_SHOCK; e.rho0 = P(2); e.K = P(3); e.c0 = P(4); e.S = P(5); e.Gamma = P(6); e.cv = P(7); e.Tr = P(8); e.Q1 = P(9); e.Q2 = P(10);
  } else if (a[1] == "fluid") { if (a.size() < 5) fatal("Error: not enough arguments.\n"); e.type = KML_EOS_FLUID; e.rho0 = P(2); e.K = P(3); e.Gamma = P(4); }
  else fatal("Unknown EOS style " + a[1] + "\n");
  eoss.push_back(e); return Var(0);
}
Var Sim::cmd_strength(std::vector<std::string> &a) {
  if (a.size() < 3) fatal("Error: too few arguments for the strength command.\n");
  for (auto &s : strengths) if (s.id == a[0]) fatal("Error: reuse of strength ID.\n");
  StrengthH s{}; s.id = a[0];
  auto P = [&](size_t i) { return (double)input.parsev(a[i]); };
  if (a[1] == "linear") { s.type = KML_STRENGTH_LINEAR; s.G = P(2); }
  else if (a[1] == "fluid") { s.type = KML_STRENGTH_FLUID; s.G = P(2); }
  else if (a[1] == "plastic") { if (a.size() < 4) fatal("Error: too few arguments for the strength command.\n"); s.type = KML_STRENGTH_PLASTIC; s.G = P(2); s.A = P(3); }
  else if (a[1] == "johnson_cook") {
    if (a.size() < 11) fatal("Error: too few arguments for the strength command.\n");
    s.type = KML_STRENGTH_JOHNSON_COOK; s.G = P(2); s.A = P(3); s.B = P(4); s.n = P(5); s.epsdot0 = P(6); s.C = P(7); s.m = P(8); s.Tr = P(9); s.Tm = P(10);
    if (s.Tr == s.Tm) fatal("Error: reference temperature Tr equals melting temperature Tm.\n");
  } else if (a[1] == "swift") {
    if (a.size() < 7) fatal("Error: too few arguments for the strength command.\n");
    s.type = KML_STRENGTH_SWIFT; s.G = P(2); s.A = P(3); s.B = P(4); s.C = P(5); s.n = P(6);
  } else fatal("Unknown strength style " + a[1] + "\n");
  strengths.push_back(s); return Var(0);
}
Var Sim::cmd_damage(std::vector<std::string> &a) {
  if (a.size() < 10 || a[1] != "damage_johnson_cook") fatal("Error: damage command: unknown style or too few arguments.\n");
  DamageH d{}; d.id = a[0]; d.type = KML_DAMAGE_JOHNSON_COOK;
  auto P = [&](size_t i) { return (double)input.parsev(a[i]); };
  d.d1 = P(2); d.d2 = P(3); d.d3 = P(4); d.d4 = P(5); d.d5 = P(6); d.epsdot0 = P(7); d.Tr = P(8); d.Tm = P(9);
  if (d.Tr == d.Tm) fatal("Error: reference temperature Tr equals melting temperature Tm.\n");
  damages.push_back(d); return Var(0);
}
Var Sim::cmd_temperature(std::vector<std::string> &a) {
  if (a.size() != 8 || a[1] != "plastic_work") fatal("Error: temperature command: unknown style or wrong number of arguments.\n");
  TemperatureH t{}; t.id = a[0]; t.type = KML_TEMPERATURE_PLASTIC_WORK;
  auto P = [&](size_t i) { return (double)input.parsev(a[i]); };
  t.chi = P(2); t.cp = P(3); t.kappa = P(4); t.alpha = P(5); t.T0 = P(6); t.Tm = P(7);
  temperatures.push_back(t); return Var(0);
}

// Material::add_material + Mat constructors, src/material.cpp:250-372, :720-778
Var Sim::cmd_material(std::vector<std::string> &a) {
  if (a.size() < 2) fatal("Error: material command not enough arguments\n");
  if (find_material(a[0]) >= 0) fatal("Error: reuse of material ID.\n");
  MaterialH M; M.id = a[0]; kml_material &m = M.km; memset(&m, 0, sizeof m);
  auto P = [&](size_t i) { return (double)input.parsev(a[i]); };
  if (a[1] == "linear" || a[1] == "neo-hookean") {
    if (a.size() < 5) fatal("Error: not enough arguments.\n");
    m.type = a[1] == "linear" ? KML_MAT_LINEAR : KML_MAT_NEO_HOOKEAN;
    double cp = 0, kappa = 0;
    if (a.size() >= 7) { cp = P(5); kappa = P(6); } // the reference reads args[6] already when 6 are given (src/material.cpp:284-287): UB there, an error here
    else if (a.size() == 6) fatal("material(linear): cp given without kappa (the reference reads past the argument list here).\n");
    m.rho0 = P(2); m.E = P(3); m.nu = P(4);
    m.G = m.E / (2 * (1 + m.nu));
    m.lambda = m.E * m.nu / ((1 + m.nu) * (1 - 2 * m.nu));
    m.K = m.E / (3 * (1 - 2 * m.nu));
    m.signal_velocity = sqrt(m.K / m.rho0);
    m.cp = cp; m.invcp = cp != 0 ? 1.0 / cp : 0; m.kappa = kappa;
  } else if (a[1] == "rigid") {
    if (a.size() < 3) fatal("Error: not enough arguments.\n");
    m.type = KML_MAT_RIGID; m.rigid = 1; m.rho0 = P(2);
  } else if (a[1] == "eos-strength") {
    if (a.size() < 4) fatal("Error: not enough arguments.\n");
    const EOSH *e = nullptr; for (auto &x : eoss) if (x.id == a[2]) e = &x;
    if (!e) fatal("Error: could not find EOS named: " + a[2] + ".\n");
    const StrengthH *s = nullptr; for (auto &x : strengths) if (x.id == a[3]) s = &x;
    if (!s) fatal("Error: could not find strength named: " + a[3] + ".\n");
    const DamageH *d = nullptr; const TemperatureH *t = nullptr;
    if (a.size() > 4) {
      for (auto &x : damages) if (x.id == a[4]) d = &x;
      if (!d) {
        for (auto &x : temperatures) if (x.id == a[4]) t = &x;
        if (!t) fatal("Error: could not find damage named: " + a[4] + ".\n");
        if (a.size() > 5) fatal("Error: the last argument of the material command should be the temperature\n");
      } else if (a.size() == 6) {
        for (auto &x : temperatures) if (x.id == a[5]) t = &x;
        if (!t) fatal("Error: could not find temperature named: " + a[5] + ".\n");
      }
    }
    m.type = KML_MAT_EOS_STRENGTH;
    M.ieos = (int)(e - eoss.data()); M.istrength = (int)(s - strengths.data()); M.idamage = d ? (int)(d - damages.data()) : -1;
    M.itemperature = t ? (int)(t - temperatures.data()) : -1;
    m.eos_type = e->type; m.eos_K = e->K; m.eos_c0 = e->c0; m.eos_S = e->S; m.eos_Gamma = e->Gamma; m.eos_cv = e->cv; m.eos_Tr = e->Tr; m.eos_Q1 = e->Q1; m.eos_Q2 = e->Q2;
    m.strength_type = s->type; m.str_G = s->G; m.str_A = s->A; m.str_B = s->B; m.str_n = s->n; m.str_epsdot0 = s->epsdot0; m.str_C = s->C; m.str_m = s->m; m.str_Tr = s->Tr; m.str_Tm = s->Tm;
    if (d) { m.damage_type = d->type; m.dmg_d1 = d->d1; m.dmg_d2 = d->d2; m.dmg_d3 = d->d3; m.dmg_d4 = d->d4; m.dmg_d5 = d->d5; m.dmg_epsdot0 = d->epsdot0; m.dmg_Tr = d->Tr; m.dmg_Tm = d->Tm; }
    if (t) { m.temperature_type = t->type; m.tmp_chi = t->chi; m.tmp_cp = t->cp; m.tmp_kappa = t->kappa; m.tmp_alpha = t->alpha; m.tmp_T0 = t->T0; m.tmp_Tm = t->Tm; }
    // Mat::Mat(eos, strength, damage, temp), src/material.cpp:720-745
    m.rho0 = e->rho0; m.K = e->K; m.G = s->G;
    m.E = 9 * m.K * m.G / (3 * m.K + m.G);
    m.nu = (3 * m.K - 2 * m.G) / (2 * (3 * m.K + m.G));
    m.lambda = m.K - 2 * m.G / 3;
    m.signal_velocity = sqrt((m.lambda + 2 * m.G) / m.rho0);
    if (t) { m.cp = t->cp; m.invcp = 1.0 / m.cp; m.kappa = t->kappa; }
  } else fatal("Error, keyword " + a[1] + " unknown!\n");
  materials.push_back(M);
  if (!quiet) std::cout << "Creating new mat with ID: " << a[0] << " (G=" << m.G << ", K=" << m.K << ", signal velocity=" << m.signal_velocity << ")\n";
  return Var(0);
}

void Sim::sync(SolidH &S) {
  uint64_t g = 0; int64_t n = 0;
  check(kml_solid_generation(ctx, S.dev, &g)); check(kml_solid_np(ctx, S.dev, &n));
  if (g == S.mirror_gen && n == S.np && (int64_t)S.x0.size() == n) return;
  S.np = n; S.x0.resize(n); S.mask.resize(n); S.ptag.resize(n);
  if (n > 0) {
    check(kml_solid_download(ctx, S.dev, KML_P_PTAG, S.ptag.data())); check(kml_solid_download(ctx, S.dev, KML_P_X0, S.x0.data()));
    check(kml_solid_download(ctx, S.dev, KML_P_MASK, S.mask.data()));
  }
  S.mirror_gen = g;
}
void Sim::mirrors_current(SolidH &S) { check(kml_solid_generation(ctx, S.dev, &S.mirror_gen)); check(kml_solid_np(ctx, S.dev, &S.np)); }

// Domain::add_solid -> Solid::Solid / options / populate / init (src/solid.cpp:59-238, :1810-2336)
Var Sim::cmd_solid(std::vector<std::string> &a) {
  if (!method_set) fatal("Error: a method should be defined before creating a solid!\n");
  if (a.size() < 2) fatal("Error: solid command not enough arguments.\n");
  if (find_solid(a[0]) >= 0) fatal("Error: reuse of solid ID.\n");
  if (a[1] != "region") fatal("solid(" + a[1] + "): only solid(..., region, ...) is supported by this build.\n");
  if (a.size() < 7) fatal("Error: not enough arguments.\n");
  std::unique_ptr<SolidH> s(new SolidH()); s->id = a[0];
  if (is_TL) { s->own_grid.reset(new GridH()); s->grid = s->own_grid.get(); } else s->grid = grid.get();
  // Solid::options, src/solid.cpp:197-238
  s->mat = find_material(a[4]);
  if (s->mat < 0) fatal("Error: could not find material named " + a[4] + "\n");
  s->grid->cellsize = input.parsev(a[5]); // Grid::setup, src/grid.cpp:415-420
  if (!is_TL && s->grid->cellsize != s->grid->desc.cellsize) fatal("solid(): cell size differs from the one given to dimension(); the reference silently changes the weight scaling here.\n");
  s->T0 = input.parsev(a[6]);
  populate(*s, a);
  solids.push_back(std::move(s));
  return Var(0);
}

// Room for the particles that migrate in: a solid of up to 4 M particles may end up entirely on one slab (a body crossing the box), a large one
// is balanced by construction and keeps 1/8 of head room.
static int64_t slab_capacity(int64_t np_local, int64_t np_global) {
  return np_global <= 4000000 ? np_global + 4096 : np_local + np_local / 8 + 4096;
}

void Sim::populate(SolidH &s, std::vector<std::string> &a) {
  int iregion = find_region(a[2]);
  if (iregion == -1) fatal("Error: region ID " + a[2] + " not does not exist.\n");
  if (!created) fatal("The domain must be created before any solids can (create_domain(...)).");
  const Region &reg = *regions[iregion];
  for (int d = 0; d < 3; d++) { s.solidlo[d] = reg.lim[2 * d]; s.solidhi[d] = reg.lim[2 * d + 1]; }
  const double delta = s.grid->cellsize;
  const double *boundlo, *boundhi;
  if (is_TL) { init_grid(*s.grid, s.solidlo, s.solidhi); boundlo = s.solidlo; boundhi = s.solidhi; }
  else { boundlo = boxlo; boundhi = boxhi; }
  int noffsetlo[3], noffsethi[3];
  for (int d = 0; d < 3; d++) {
    double Llo = dmax(0.0, sublo[d] - boundlo[d]);
    double Lhi = dmax(0.0, dmin(subhi[d], boundhi[d]) - boundlo[d]);
    noffsetlo[d] = (int)floor(Llo / delta);
    noffsethi[d] = (int)ceil(Lhi / delta);
  }
  int nsub[3];
  nsub[0] = std::max(0, noffsethi[0] - noffsetlo[0]);
  nsub[1] = dimension >= 2 ? std::max(0, noffsethi[1] - noffsetlo[1]) : 1;
  nsub[2] = dimension >= 3 ? std::max(0, noffsethi[2] - noffsetlo[2]) : 1;
  for (int d = 0; d < 3; d++)
    while (boundlo[d] + delta * (noffsetlo[d] + nsub[d] - 0.5) < dmin(subhi[d], boundhi[d])) nsub[d]++;

  double vol_ = dimension == 1 ? delta : (dimension == 2 ? delta * delta : delta * delta * delta);
  const kml_material &mat = materials[s.mat].km;
  double mass_ = mat.rho0 * vol_;
  s.np_per_cell = (int)(double)input.parsev(a[3]);
  const int ppc = s.np_per_cell;
  double xi; int nip = 1; std::vector<double> ip;
  const int nc = is_CPDI ? (1 << dimension) : 0;
  if (ppc == 1) { nip = 1; ip = {0, 0, 0}; }
  else if (ppc == 2) {
    nip = dimension == 1 ? 2 : (dimension == 2 ? 4 : 8); xi = 0.25;
    ip = {-xi, -xi, -xi, -xi, xi, -xi, xi, -xi, -xi, xi, xi, -xi, -xi, -xi, xi, -xi, xi, xi, xi, -xi, xi, xi, xi, xi};
  } else if (ppc == 3) {
    xi = nc == 0 ? 0.7746 / 2 : 1.0 / 3.0;
    nip = dimension == 1 ? 3 : (dimension == 2 ? 9 : 27);
    const double t[3] = {-xi, 0, xi};
    for (int c = 0; c < 3; c++) for (int aa = 0; aa < 3; aa++) for (int b = 0; b < 3; b++) { ip.push_back(t[aa]); ip.push_back(t[b]); ip.push_back(t[c]); } // order of src/solid.cpp:2076-2102
  } else {
    nip = dimension == 1 ? ppc : (dimension == 2 ? ppc * ppc : ppc * ppc * ppc);
    double d = 1.0 / ppc;
    for (int k = 0; k < ppc; k++) for (int i = 0; i < ppc; i++) for (int j = 0; j < ppc; j++) { ip.push_back((i + 0.5) * d - 0.5); ip.push_back((j + 0.5) * d - 0.5); ip.push_back((k + 0.5) * d - 0.5); }
  }
  mass_ /= (double)nip; vol_ /= (double)nip;
  // CPDI half-length of the particle domain, src/solid.cpp:2025-2027 (lp = delta), :2033,2053,2067,2105 (scaled by the lattice)
  double lp = delta;
  if (ppc == 1) lp *= 0.5; else if (ppc == 2) lp *= 0.25; else if (ppc == 3) lp *= 1.0 / 6.0; else lp *= 1.0 / (2 * ppc);

  // ---- populate on the device: the lattice is a pure function of (i, j, k, q); nothing of size np exists on the host
  kml_region kreg;
  if (kml_has_device_setup() && !is_CPDI && nip <= 64 && reg.to_kml(kreg) && !getenv("KML_HOST_POPULATE")) {
    kml_lattice L; memset(&L, 0, sizeof L);
    for (int d = 0; d < 3; d++) { L.boundlo[d] = boundlo[d]; L.noffsetlo[d] = noffsetlo[d]; L.nsub[d] = nsub[d]; L.sublo[d] = sublo[d]; L.subhi[d] = subhi[d]; }
    L.delta = delta; L.nip = nip; L.dim = dimension;
    for (int q = 0; q < 3 * nip; q++) L.ip[q] = ip[q];
    L.mass = mass_; L.vol = vol_; L.T0 = s.T0; L.set_T = temp ? 1 : 0; L.axisymmetric = axisymmetric ? 1 : 0; L.rho0 = mat.rho0; L.tag_first = np_total + 1;
    const bool slab = nranks > 1 && !is_TL;
    GridH &g = *s.grid;
    int nbins = 1;
    if (slab) {
      L.slab = 1; L.slab_linear = shape_function == KML_SHAPE_LINEAR; L.slab_lo = boxlo[0]; L.slab_ih = 1.0 / g.cellsize; L.base_lo = INT_MIN; L.base_hi = INT_MAX; nbins = g.nx_global + 1;
      // tags must be contiguous per slab: every sub-point of a lattice column shares one stencil base, non-decreasing along x
      int last = INT_MIN;
      for (int i = 0; i < nsub[0]; i++) for (int q = 0; q < nip; q++) {
        const double x = boundlo[0] + delta * (noffsetlo[0] + i + 0.5 + ip[3 * q]); const double t = (x - boxlo[0]) * L.slab_ih;
        const int b = std::min(std::max(L.slab_linear ? (int)t : (int)(t - 1.0), 0), g.nx_global);
        if (q > 0 ? b != last : b < last) fatal("slab decomposition: the particle lattice is not ordered along x\n");
        last = b;
      }
    }
    std::vector<int64_t> hist(nbins, 0);
    check(kml_lattice_histogram(ctx, &L, &kreg, hist.data(), nbins));
    int64_t np_global = 0, tag_offset = 0, np_local = 0;
    for (int64_t h : hist) np_global += h;
    if (slab) {
      int base_lo, base_hi;
      if (grid_pending) {
        std::vector<int> cut(nranks + 1, 0); cut[nranks] = g.nx_global + 4;
        int64_t cum = 0; int r = 1;
        for (int b = 0; b <= g.nx_global && r < nranks; b++) { cum += hist[b]; while (r < nranks && cum >= (np_global * r) / nranks) cut[r++] = b + 1; }
        base_lo = rank == 0 ? -4 : cut[rank]; base_hi = cut[rank + 1];
        create_device_grid(g, base_lo, base_hi);
      } else { base_lo = g.desc.base_lo; base_hi = g.desc.base_hi; }
      for (int b = 0; b < nbins; b++) { if (b < base_lo) tag_offset += hist[b]; else if (b < base_hi) np_local += hist[b]; }
      L.base_lo = base_lo; L.base_hi = base_hi;
    } else np_local = np_global;
    if (np_global == 0) fatal("Error: solid does not have any particles.\n"); // a slab of a decomposed run may hold none of them (yet)
    s.np = np_local; s.np_created = np_global; s.mirror_gen = 0; s.x0.clear(); s.mask.clear(); s.ptag.clear();
    kml_solid_desc d; memset(&d, 0, sizeof d); d.np = s.np; d.capacity = nranks > 1 ? slab_capacity(s.np, np_global) : s.np; d.grid = s.grid->id; d.np_per_cell = s.np_per_cell; d.mat = mat;
    check(kml_solid_create(ctx, &d, &s.dev));
    check(kml_solid_populate(ctx, s.dev, &L, &kreg, tag_offset));
    np_total += np_global; np_global_last = np_global; tag_offset_last = tag_offset;
    // Solid::init totals, src/solid.cpp:170-195: the reference adds particle after particle; every particle of a plane lattice carries the
    // same volume and mass, so the same sequence of additions is repeated here (device reduction for axisymmetry and beyond 10^7 particles)
    if (axisymmetric || s.np > 10000000) { check(kml_solid_sum(ctx, s.dev, KML_P_VOL, 0, &s.vtot)); check(kml_solid_sum(ctx, s.dev, KML_P_MASS, 0, &s.mtot)); }
    else { s.vtot = s.mtot = 0; for (int64_t i = 0; i < s.np; i++) { s.vtot += vol_; s.mtot += mass_; } }
    if (!quiet) std::cout << "Solid " << s.id << ": np=" << s.np << " total volume = " << s.vtot << " total mass = " << s.mtot << " grid " << s.grid->desc.n[0] << "x" << s.grid->desc.n[1] << "x" << s.grid->desc.n[2] << std::endl;
    return;
  }

  s.x0.clear();
  const int dim = dimension;
  auto lattice = [&](auto &&accept) {
    for (int i = 0; i < nsub[0]; i++) for (int j = 0; j < nsub[1]; j++) for (int k = 0; k < nsub[2]; k++) for (int q = 0; q < nip; q++) {
      std::array<double, 3> x;
      x[0] = boundlo[0] + delta * (noffsetlo[0] + i + 0.5 + ip[3 * q + 0]);
      x[1] = boundlo[1] + delta * (noffsetlo[1] + j + 0.5 + ip[3 * q + 1]);
      x[2] = dim == 3 ? boundlo[2] + delta * (noffsetlo[2] + k + 0.5 + ip[3 * q + 2]) : 0;
      bool in_sub = !(x[0] < sublo[0] || x[0] > subhi[0] || x[1] < sublo[1] || x[1] > subhi[1] || x[2] < sublo[2] || x[2] > subhi[2]); // Domain::inside_subdomain
      if (in_sub && reg.inside(x[0], x[1], x[2]) == 1) accept(x);
    }
  };
  int64_t tag_offset = 0, np_global = 0;
  if (nranks > 1 && !is_TL) {
    // slab ownership by the GLOBAL stencil base of the particle (what the kernels use); cuts balance the particle count
    GridH &g = *s.grid;
    const double ih = 1.0 / g.cellsize; const bool lin = shape_function == KML_SHAPE_LINEAR;
    auto base_of = [&](double x) { double t = (x - boxlo[0]) * ih; return lin ? (int)t : (int)(t - 1.0); };
    std::vector<int64_t> hist(g.nx_global + 1, 0);
    int last = -1; bool monotone = true;
    lattice([&](const std::array<double, 3> &x) { int b = std::min(std::max(base_of(x[0]), 0), g.nx_global); if (b < last) monotone = false; last = b; hist[b]++; np_global++; });
    if (!monotone) fatal("slab decomposition: the particle lattice is not ordered along x\n");
    int base_lo, base_hi;
    if (grid_pending) {
      std::vector<int> cut(nranks + 1, 0); cut[nranks] = g.nx_global + 4;
      int64_t cum = 0; int r = 1;
      for (int b = 0; b <= g.nx_global && r < nranks; b++) { cum += hist[b]; while (r < nranks && cum >= (np_global * r) / nranks) cut[r++] = b + 1; }
      base_lo = rank == 0 ? -4 : cut[rank]; base_hi = cut[rank + 1];
      create_device_grid(g, base_lo, base_hi);
    } else { base_lo = g.desc.base_lo; base_hi = g.desc.base_hi; }
    for (int b = 0; b < std::min(std::max(base_lo, 0), g.nx_global + 1); b++) tag_offset += hist[b];
    lattice([&](const std::array<double, 3> &x) { int b = base_of(x[0]); if (b >= base_lo && b < base_hi) s.x0.push_back(x); });
  } else {
    lattice([&](const std::array<double, 3> &x) { s.x0.push_back(x); });
    np_global = (int64_t)s.x0.size();
  }
  s.np = (int64_t)s.x0.size(); s.np_created = np_global;
  if (np_global == 0) fatal("Error: solid does not have any particles.\n"); // a slab of a decomposed run may hold none of them (yet)
  s.mask.assign(s.np, 1);
  s.ptag.resize(s.np);
  std::vector<double> mass(s.np), vol(s.np), T(s.np, s.T0);
  for (int64_t i = 0; i < s.np; i++) {
    if (axisymmetric) { mass[i] = mass_ * s.x0[i][0]; vol[i] = mass[i] / mat.rho0; }
    else { mass[i] = mass_; vol[i] = vol_; }
    s.ptag[i] = tag_offset + i + 1 + np_total; // src/solid.cpp:2322 (tag_offset: particles of lower slabs)
  }
  np_total += np_global; // domain->np_total += np, src/solid.cpp:2334
  np_global_last = np_global; tag_offset_last = tag_offset;

  // upload; everything not set here starts at the values of src/solid.cpp:2283-2321 (F = R = I, J = 1, mask = 1, rest 0)
  kml_solid_desc d; memset(&d, 0, sizeof d); d.np = s.np; d.capacity = nranks > 1 ? slab_capacity(s.np, np_global) : s.np; d.grid = s.grid->id; d.np_per_cell = s.np_per_cell; d.mat = mat;
  check(kml_solid_create(ctx, &d, &s.dev));
  check(kml_solid_upload(ctx, s.dev, KML_P_PTAG, s.ptag.data()));
  check(kml_solid_upload(ctx, s.dev, KML_P_X, s.x0.data()));
  check(kml_solid_upload(ctx, s.dev, KML_P_X0, s.x0.data()));
  check(kml_solid_upload(ctx, s.dev, KML_P_MASS, mass.data()));
  check(kml_solid_upload(ctx, s.dev, KML_P_VOL0, vol.data()));
  check(kml_solid_upload(ctx, s.dev, KML_P_VOL, vol.data()));
  if (temp) check(kml_solid_upload(ctx, s.dev, KML_P_T, T.data()));
  if (is_CPDI) { // particle domains, src/solid.cpp:2186-2247: R4 domain vectors / Q4 corner positions (2-D)
    if (dimension != 2) fatal("Error: ULCPDI is only 2D....\n");
    if (cpdi_style == 0) {
      std::vector<std::array<double, 3>> rp(2 * s.np);
      for (int64_t i = 0; i < s.np; i++) { rp[2 * i] = {lp, 0, 0}; rp[2 * i + 1] = {0, lp, 0}; }
      check(kml_solid_upload(ctx, s.dev, KML_P_RP0, rp.data())); check(kml_solid_upload(ctx, s.dev, KML_P_RP, rp.data()));
    } else {
      std::vector<std::array<double, 3>> xc(4 * s.np);
      for (int64_t i = 0; i < s.np; i++) {
        const double x = s.x0[i][0], y = s.x0[i][1];
        xc[4 * i] = {x - lp, y - lp, 0}; xc[4 * i + 1] = {x + lp, y - lp, 0}; xc[4 * i + 2] = {x + lp, y + lp, 0}; xc[4 * i + 3] = {x - lp, y + lp, 0};
      }
      check(kml_solid_upload(ctx, s.dev, KML_P_XPC0, xc.data())); check(kml_solid_upload(ctx, s.dev, KML_P_XPC, xc.data()));
    }
  }
  mirrors_current(s);
  // Solid::init totals, src/solid.cpp:170-195
  s.vtot = s.mtot = 0; for (int64_t i = 0; i < s.np; i++) { s.vtot += vol[i]; s.mtot += mass[i]; }
  if (!quiet) std::cout << "Solid " << s.id << ": np=" << s.np << " total volume = " << s.vtot << " total mass = " << s.mtot << " grid " << s.grid->desc.n[0] << "x" << s.grid->desc.n[1] << "x" << s.grid->desc.n[2] << std::endl;
}

// Group::assign, src/group.cpp:65-238
Var Sim::cmd_group(std::vector<std::string> &a) {
  if (a.size() < 5) fatal("Error: too few arguments for group: requires at least 4 arguments.\n");
  int ig = find_group(a[0]);
  if (ig == -1) {
    if (ngroup == MAX_GROUP) fatal("Too many groups.\n");
    for (int i = 0; i < MAX_GROUP; i++) if (gnames[i].empty()) { ig = i; break; }
    gnames[ig] = a[0]; ngroup++;
  }
  const int bit = gbitmask[ig];
  if (a[1] == "particles") gpon[ig] = "particles"; else if (a[1] == "nodes") gpon[ig] = "nodes";
  else fatal("Error: do not understand keyword " + a[1] + ", \"particles\" or \"nodes\" expected.\n");
  if (a[2] != "region") fatal("Error: unknown keyword in group command: " + a[2] + ".\n");
  gregion[ig] = find_region(a[3]);
  if (gregion[ig] == -1) fatal("Error: could not find region " + a[3] + ".\n");
  const Region &reg = *regions[gregion[ig]];
  std::vector<int> targets;
  if (a[4] == "all") { gsolid[ig] = -1; for (size_t i = 0; i < solids.size(); i++) targets.push_back((int)i); }
  else if (a[4] == "solid") {
    for (size_t i = 5; i < a.size(); i++) { gsolid[ig] = find_solid(a[i]); if (gsolid[ig] == -1) fatal("Error: cannot find solid with ID " + a[i] + ".\n"); targets.push_back(gsolid[ig]); }
  } else fatal("Error: unknown keyword in group command: " + a[3] + ".\n");
  for (int is : targets) {
    SolidH &s = *solids[is]; int n = 0;
    kml_region kreg;
    if (gpon[ig] == "particles" && kml_has_device_setup() && reg.to_kml(kreg) && !getenv("KML_HOST_POPULATE")) {
      int64_t cnt = 0; check(kml_solid_group_assign(ctx, s.dev, &kreg, bit, &cnt)); n = (int)cnt; // Group::assign on the device (the mirrors follow lazily)
    } else if (gpon[ig] == "particles") {
      sync(s);
      for (int64_t ip = 0; ip < s.np; ip++) if (reg.match(s.x0[ip][0], s.x0[ip][1], s.x0[ip][2])) { s.mask[ip] |= bit; n++; }
      check(kml_solid_upload(ctx, s.dev, KML_P_MASK, s.mask.data())); mirrors_current(s);
    } else {
      GridH &g = *s.grid; const kml_grid_desc &d = g.desc; int64_t l = 0;
      for (int i = 0; i < d.n[0]; i++) for (int j = 0; j < d.n[1]; j++) for (int k = 0; k < d.n[2]; k++, l++) {
        double x = d.lo[0] + (i + d.goff) * d.h, y = dimension >= 2 ? d.lo[1] + j * d.h : 0, z = dimension == 3 ? d.lo[2] + k * d.h : 0; // node positions, src/grid.cpp:222-227
        if (reg.match(x, y, z)) { g.mask[l] |= bit; n++; }
      }
      check(kml_grid_upload(ctx, g.id, KML_N_MASK, g.mask.data()));
    }
    if (!quiet) std::cout << n << " " << gpon[ig] << " from solid " << s.id << " found" << std::endl;
  }
  return Var(0);
}

} // namespace kmlh
