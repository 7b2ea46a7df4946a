// scheme.cpp - the per-step stage order (host-side orchestration, as in the reference),
// the fix / compute commands that sit between stages, run commands and log/dump output.
#include <algorithm>
#include "sim.h"
#include <cctype>
#include <climits>
#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>
#include <zlib.h>

namespace kmlh {

// ------------------------------------------------------------------ fixes -------------------
namespace {
// ---- restart files in the reference's binary layout (src/write_restart.cpp and the write_restart members it calls) ----
template <class T> void rput(std::ostream &os, const T &v) { os.write(reinterpret_cast<const char *>(&v), sizeof(T)); }
void rput_str(std::ostream &os, const std::string &t) { const size_t n = t.size(); rput(os, n); os.write(t.data(), (std::streamsize)n); }
void rput_var(std::ostream &os, const Var &v) { // Var::write_to_restart, src/var.cpp:312-319
  rput_str(os, v.eq()); const double val = v.result(); rput(os, val); const bool c = v.is_constant(); rput(os, c);
}
void rput_sets(std::ostream &os, const bool *set, const Var *val, const Var *prev) { // the xset/yset/zset + Var blocks of the node / particle fixes
  for (int d = 0; d < 3; d++) rput(os, set[d]);
  for (int d = 0; d < 3; d++) if (set[d]) { rput_var(os, val[d]); if (prev) rput_var(os, prev[d]); }
}


// FixInitialVelocityParticles, reference src/fix_initial_velocity_particles.cpp:36-161.
// Per-particle expressions are evaluated by the script interpreter exactly as the reference
// does (vars x,y,z,x0,y0,z0 set per particle, expression re-parsed), once, at step 1.
struct FixInitialVelocityParticles : Fix {
  bool set[3] = {false, false, false}; Var val[3];
  bool needs_time(const Sim &s) const override { return s.ntimestep < 1; } // acts in step 1 only
  void initial_integrate(Sim &s) override {
    if (s.ntimestep != 1) return;
    const int solid = s.gsolid[igroup];
    if (kml_has_device_setup()) { // the expressions as postfix programs, one kernel for the whole group
      kml_expr prog[3]; int m = 0; bool ok = true;
      for (int d = 0; d < 3; d++) { prog[d].n = 0; if (set[d]) { m |= 1 << d; ok = ok && s.input.compile(val[d], &prog[d]); } }
      if (ok) { s.check(kml_fix_set_particles_expr(s.ctx, solid == -1 ? -1 : s.solids[solid]->dev, groupbit, KML_P_V, m, prog)); return; }
    }
    for (size_t is = 0; is < s.solids.size(); is++) {
      if (solid != -1 && (int)is != solid) continue;
      SolidH &S = *s.solids[is]; s.sync(S);
      std::vector<std::array<double, 3>> x(S.np), v(S.np);
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_X, x.data()));
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_V, v.data()));
      bool fast = true; // constant expressions need no per-particle re-parse
      for (int d = 0; d < 3; d++) if (set[d] && !val[d].is_constant()) fast = false;
      for (int64_t ip = 0; ip < S.np; ip++) {
        if (!(S.mask[ip] & groupbit)) continue;
        if (!fast) {
          s.input.vars["x"] = Var("x", x[ip][0]); s.input.vars["y"] = Var("y", x[ip][1]); s.input.vars["z"] = Var("z", x[ip][2]);
          s.input.vars["x0"] = Var("x0", S.x0[ip][0]); s.input.vars["y0"] = Var("y0", S.x0[ip][1]); s.input.vars["z0"] = Var("z0", S.x0[ip][2]);
        }
        for (int d = 0; d < 3; d++) if (set[d]) v[ip][d] = val[d].result(&s.input);
      }
      s.check(kml_solid_upload(s.ctx, S.dev, KML_P_V, v.data()));
    }
  }
};

// Per-particle expressions: the reference publishes x, y, z, x0, y0, z0 of the particle before every evaluation
// (e.g. src/fix_velocity_particles.cpp:143-148).  Constant expressions skip that.
static inline void publish_particle(Sim &s, const double *x, const std::array<double, 3> &x0) {
  s.input.vars["x"] = Var("x", x[0]); s.input.vars["y"] = Var("y", x[1]); s.input.vars["z"] = Var("z", x[2]);
  s.input.vars["x0"] = Var("x0", x0[0]); s.input.vars["y0"] = Var("y0", x0[1]); s.input.vars["z0"] = Var("z0", x0[2]);
}
// does the expression mention x, y, z, x0, y0 or z0 as a whole word?  (Expressions that do not can be evaluated once per step and
// applied by a device kernel; the others are evaluated per particle on the host, like the reference does for every particle.)
static bool particle_dependent(const Var &v) {
  if (v.is_constant()) return false;
  const std::string &e = v.eq();
  for (size_t i = 0; i < e.size();) {
    if (std::isalpha((unsigned char)e[i]) || e[i] == '_') {
      size_t j = i; while (j < e.size() && (std::isalnum((unsigned char)e[j]) || e[j] == '_')) j++;
      const std::string w = e.substr(i, j - i);
      if (w == "x" || w == "y" || w == "z" || w == "x0" || w == "y0" || w == "z0") return true;
      i = j;
    } else i++;
  }
  return false;
}
template <class F> static void for_group_solids(Sim &s, int igroup, F f) { // "solid == -1: every solid" loops of the reference's fixes
  const int solid = s.gsolid[igroup];
  for (size_t is = 0; is < s.solids.size(); is++) if (solid == -1 || (int)is == solid) { s.sync(*s.solids[is]); f(*s.solids[is]); }
}

// FixVelocityParticles, reference src/fix_velocity_particles.cpp:30-300: v = v(t - dt) before the step, v = v(t) and
// x = x_old + dt v(t) after advance_particles; the reaction m (v(t) - v_advanced) / dt is published as <id>_x/_y/_z.
// (The reference also writes v_update at initial_integrate; only Solid::compute_velocity_nodes of a rigid solid reads it,
// into the scratch copy it keeps in grid->mb, which nothing consumes - src/solid.cpp:366-381, src/grid.cpp:455-461.)
struct FixVelocityParticles : Fix {
  void write_restart(std::ostream &os) const override { rput_sets(os, set, val, prev); } // src/fix_velocity_particles.cpp:303-320
  bool set[3] = {false, false, false}; Var val[3], prev[3];
  std::vector<std::vector<double>> xold; // per solid, rows [np][3]
  bool on_device() const { for (int d = 0; d < 3; d++) if (set[d] && (particle_dependent(val[d]) || particle_dependent(prev[d]))) return false; return true; }
  int dev_solid(Sim &s) const { return s.gsolid[igroup] == -1 ? -1 : s.solids[s.gsolid[igroup]]->dev; }
  void initial_integrate(Sim &s) override {
    if (on_device()) { // kernel path: the values are the same for every particle of the group
      double v[3] = {0, 0, 0}, vp[3] = {0, 0, 0}; int m = 0;
      for (int d = 0; d < 3; d++) if (set[d]) { m |= 1 << d; v[d] = val[d].result(&s.input); vp[d] = prev[d].result(&s.input); }
      s.check(kml_fix_velocity_particles(s.ctx, dev_solid(s), groupbit, m, v, vp, 0, nullptr));
      return;
    }
    xold.assign(s.solids.size(), {});
    const int solid = s.gsolid[igroup];
    for (size_t is = 0; is < s.solids.size(); is++) {
      if (solid != -1 && (int)is != solid) continue;
      SolidH &S = *s.solids[is]; s.sync(S);
      std::vector<double> &x = xold[is]; x.resize(3 * S.np); std::vector<double> v(3 * S.np);
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_X, x.data()));
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_V, v.data()));
      bool fast = true; for (int d = 0; d < 3; d++) if (set[d] && !prev[d].is_constant()) fast = false;
      for (int64_t ip = 0; ip < S.np; ip++) {
        if (!(S.mask[ip] & groupbit)) continue;
        if (!fast) publish_particle(s, &x[3 * ip], S.x0[ip]);
        for (int d = 0; d < 3; d++) if (set[d]) v[3 * ip + d] = prev[d].result(&s.input);
      }
      s.check(kml_solid_upload(s.ctx, S.dev, KML_P_V, v.data()));
    }
  }
  void post_advance_particles(Sim &s) override {
    double ftot[3] = {0, 0, 0}; const double inv_dt = 1.0 / s.dt;
    if (on_device()) {
      double v[3] = {0, 0, 0}; int m = 0;
      for (int d = 0; d < 3; d++) if (set[d]) { m |= 1 << d; v[d] = val[d].result(&s.input); }
      s.check(kml_fix_velocity_particles(s.ctx, dev_solid(s), groupbit, m, v, v, 1, ftot));
      s.input.vars[id + "_x"] = Var(id + "_x", ftot[0]); s.input.vars[id + "_y"] = Var(id + "_y", ftot[1]); s.input.vars[id + "_z"] = Var(id + "_z", ftot[2]);
      return;
    }
    const int solid = s.gsolid[igroup];
    for (size_t is = 0; is < s.solids.size(); is++) {
      if (solid != -1 && (int)is != solid) continue;
      SolidH &S = *s.solids[is]; s.sync(S);
      std::vector<double> x(3 * S.np), v(3 * S.np), mass(S.np); const std::vector<double> &xo = xold[is];
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_X, x.data()));
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_V, v.data()));
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_MASS, mass.data()));
      bool fast = true; for (int d = 0; d < 3; d++) if (set[d] && !val[d].is_constant()) fast = false;
      for (int64_t ip = 0; ip < S.np; ip++) {
        if (!(S.mask[ip] & groupbit)) continue;
        if (!fast) publish_particle(s, &xo[3 * ip], S.x0[ip]);
        double Dv[3] = {0, 0, 0};
        for (int d = 0; d < 3; d++) if (set[d]) {
          const double vd = val[d].result(&s.input);
          Dv[d] = vd - v[3 * ip + d]; v[3 * ip + d] = vd; x[3 * ip + d] = xo[3 * ip + d] + s.dt * vd;
        }
        for (int d = 0; d < 3; d++) ftot[d] += (inv_dt * mass[ip]) * Dv[d];
      }
      s.check(kml_solid_upload(s.ctx, S.dev, KML_P_V, v.data()));
      s.check(kml_solid_upload(s.ctx, S.dev, KML_P_X, x.data()));
    }
    s.input.vars[id + "_x"] = Var(id + "_x", ftot[0]); s.input.vars[id + "_y"] = Var(id + "_y", ftot[1]); s.input.vars[id + "_z"] = Var(id + "_z", ftot[2]);
  }
};

// FixTemperatureParticles, reference src/fix_temperature_particles.cpp:92-181: T = T(t - dt) before the step, T = T(t) after
// advance_particles
struct FixTemperatureParticles : Fix {
  void write_restart(std::ostream &os) const override { rput_var(os, val); rput_var(os, prev); } // src/fix_temperature_particles.cpp:184-187
  Var val, prev;
  void apply(Sim &s, Var &e) {
    if (!particle_dependent(e)) { // kernel path
      s.check(kml_fix_temperature_particles(s.ctx, s.gsolid[igroup] == -1 ? -1 : s.solids[s.gsolid[igroup]]->dev, groupbit, e.result(&s.input)));
      return;
    }
    for_group_solids(s, igroup, [&](SolidH &S) {
      std::vector<double> x(3 * S.np), T(S.np);
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_T, T.data()));
      const bool fast = e.is_constant();
      if (!fast) s.check(kml_solid_download(s.ctx, S.dev, KML_P_X, x.data()));
      for (int64_t ip = 0; ip < S.np; ip++) {
        if (!(S.mask[ip] & groupbit)) continue;
        if (!fast) publish_particle(s, &x[3 * ip], S.x0[ip]);
        T[ip] = e.result(&s.input);
      }
      s.check(kml_solid_upload(s.ctx, S.dev, KML_P_T, T.data()));
    });
  }
  void initial_integrate(Sim &s) override { apply(s, prev); }
  void post_advance_particles(Sim &s) override { apply(s, val); }
};

// FixInitialStress, reference src/fix_initial_stress.cpp:96-162: at step 1, sigma components (xx, yy, zz, yz, xz, xy) of the
// group's particles; total-Lagrangian: vol0PK1 = vol0 sigma
struct FixInitialStress : Fix {
  bool set[6] = {false, false, false, false, false, false}; Var val[6];
  void initial_integrate(Sim &s) override {
    if (s.ntimestep != 1) return;
    static const int A[6] = {0, 1, 2, 1, 0, 0}, B[6] = {0, 1, 2, 2, 2, 1};
    for_group_solids(s, igroup, [&](SolidH &S) {
      std::vector<double> x(3 * S.np), sig(9 * S.np), pk1, vol0;
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_X, x.data()));
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_SIGMA, sig.data()));
      if (s.is_TL) { pk1.resize(9 * S.np); vol0.resize(S.np); s.check(kml_solid_download(s.ctx, S.dev, KML_P_VOL0PK1, pk1.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_VOL0, vol0.data())); }
      bool fast = true; for (int k = 0; k < 6; k++) if (set[k] && !val[k].is_constant()) fast = false;
      for (int64_t ip = 0; ip < S.np; ip++) {
        if (!(S.mask[ip] & groupbit)) continue;
        if (!fast) publish_particle(s, &x[3 * ip], S.x0[ip]);
        for (int k = 0; k < 6; k++) if (set[k]) { const double v = val[k].result(&s.input); sig[9 * ip + 3 * A[k] + B[k]] = v; sig[9 * ip + 3 * B[k] + A[k]] = v; }
        if (s.is_TL) for (int k = 0; k < 9; k++) pk1[9 * ip + k] = vol0[ip] * sig[9 * ip + k];
      }
      s.check(kml_solid_upload(s.ctx, S.dev, KML_P_SIGMA, sig.data()));
      if (s.is_TL) s.check(kml_solid_upload(s.ctx, S.dev, KML_P_VOL0PK1, pk1.data()));
    });
  }
};

// Node fixes work on every solid's grid (TL) or on the shared background grid (UL), like the "solid == -1" loops of the reference
template <class F> static void for_group_grids(Sim &s, int igroup, F f) {
  if (!s.is_TL) { if (s.grid) f(*s.grid); return; }
  for_group_solids(s, igroup, [&](SolidH &S) { f(*S.grid); });
}

// FixTemperatureNodes, reference src/fix_temperature_nodes.cpp:74-146: T_update = T(t), T = T(t - dt) after the grid update;
// T = T(t) after the MUSL re-projection
struct FixTemperatureNodes : Fix {
  void write_restart(std::ostream &os) const override { rput_var(os, val); rput_var(os, prev); } // src/fix_temperature_nodes.cpp:149-152
  Var val, prev;
  int dev_solid(Sim &s) const { return s.gsolid[igroup] == -1 ? -1 : s.solids[s.gsolid[igroup]]->dev; }
  void post_update_grid_state(Sim &s) override {
    s.check(kml_fix_temperature_nodes(s.ctx, dev_solid(s), groupbit, val.result(&s.input), prev.result(&s.input), 0));
  }
  void post_velocities_to_grid(Sim &s) override {
    s.check(kml_fix_temperature_nodes(s.ctx, dev_solid(s), groupbit, val.result(&s.input), 0.0, 1));
  }
};

// FixInitialVelocityNodes, reference src/fix_initial_velocity_nodes.cpp:103-226: at step 1, v_update after the grid update and v
// after the MUSL re-projection take the given (x0, y0, z0)-dependent values on the group's nodes
struct FixInitialVelocityNodes : Fix {
  bool set[3] = {false, false, false}; Var val[3];
  void apply(Sim &s, int field) {
    if (s.ntimestep != 1) return;
    for_group_grids(s, igroup, [&](GridH &g) {
      std::vector<double> v(3 * g.nnodes), x0(3 * g.nnodes);
      s.check(kml_grid_download(s.ctx, g.id, field, v.data())); s.check(kml_grid_download(s.ctx, g.id, KML_N_X0, x0.data()));
      for (int64_t i = 0; i < g.nnodes; i++) {
        if (!(g.mask[i] & groupbit)) continue;
        s.input.vars["x0"] = Var("x0", x0[3 * i]); s.input.vars["y0"] = Var("y0", x0[3 * i + 1]); s.input.vars["z0"] = Var("z0", x0[3 * i + 2]);
        for (int d = 0; d < 3; d++) if (set[d]) v[3 * i + d] = val[d].result(&s.input);
      }
      s.check(kml_grid_upload(s.ctx, g.id, field, v.data()));
    });
  }
  void post_update_grid_state(Sim &s) override { apply(s, KML_N_V_UPDATE); }
  void post_velocities_to_grid(Sim &s) override { apply(s, KML_N_V); }
};

// FixVelocityNodes, reference src/fix_velocity_nodes.cpp:36-268
struct FixVelocityNodes : Fix {
  void write_restart(std::ostream &os) const override { rput_sets(os, set, val, prev); } // src/fix_velocity_nodes.cpp:270-292
  bool set[3] = {false, false, false}; Var val[3], prev[3];
  void apply(Sim &s, int which) {
    double v[3] = {0, 0, 0}, vp[3] = {0, 0, 0}, ftot[3]; int m = 0;
    for (int d = 0; d < 3; d++) if (set[d]) { m |= 1 << d; v[d] = val[d].result(&s.input); if (which == 0) vp[d] = prev[d].result(&s.input); }
    s.check(kml_fix_velocity_nodes(s.ctx, s.gsolid[igroup], groupbit, m, v, vp, which, which == 0 ? ftot : nullptr));
    if (which == 0) { s.input.vars[id + "_x"] = Var(id + "_x", ftot[0]); s.input.vars[id + "_y"] = Var(id + "_y", ftot[1]); s.input.vars[id + "_z"] = Var(id + "_z", ftot[2]); }
  }
  void post_update_grid_state(Sim &s) override { apply(s, 0); }
  void post_velocities_to_grid(Sim &s) override { apply(s, 1); }
};

// FixBodyforce, reference src/fix_body_force.cpp:36-180 (components that do not depend on x0,y0,z0)
struct FixBodyForce : Fix {
  void write_restart(std::ostream &os) const override { rput_sets(os, set, val, nullptr); } // src/fix_body_force.cpp:183-195
  bool set[3] = {false, false, false}; Var val[3];
  void post_particles_to_grid(Sim &s) override {
    double f[3] = {0, 0, 0}, ftot[3]; int m = 0;
    for (int d = 0; d < 3; d++) if (set[d]) { m |= 1 << d; f[d] = val[d].result(&s.input); }
    s.check(kml_fix_body_force(s.ctx, s.gsolid[igroup], groupbit, m, f, ftot));
    const char *sfx[3] = {"_x", "_y", "_z"};
    for (int d = 0; d < 3; d++) if (set[d]) s.input.vars[id + sfx[d]] = Var(id + sfx[d], ftot[d]);
  }
};

// FixForceNodes, reference src/fix_force_nodes.cpp:32-193
struct FixForceNodes : Fix {
  void write_restart(std::ostream &os) const override { rput_sets(os, set, val, nullptr); } // src/fix_force_nodes.cpp:196-230
  bool set[3] = {false, false, false}; Var val[3];
  void post_particles_to_grid(Sim &s) override {
    double f[3] = {0, 0, 0}, ftot[3]; int m = 0;
    for (int d = 0; d < 3; d++) if (set[d]) { m |= 1 << d; f[d] = val[d].result(&s.input); }
    s.check(kml_fix_force_nodes(s.ctx, s.gsolid[igroup], groupbit, m, f, ftot));
    const char *sfx[3] = {"_x", "_y", "_z"};
    for (int d = 0; d < 3; d++) if (set[d]) s.input.vars[id + sfx[d]] = Var(id + sfx[d], ftot[d]);
  }
};

// FixKineticEnergy / FixStrainEnergy, reference src/fix_kinetic_energy.cpp:66-110, src/fix_strain_energy.cpp: the energy of the
// group on output steps, published as <id>_s
struct FixEnergy : Fix {
  bool kinetic = true;
  void final_integrate(Sim &s) override {
    bool due = s.ntimestep == s.next_log || s.ntimestep == s.nsteps;
    for (auto &d : s.dumps) due = due || d.next == s.ntimestep;
    if (!due) return;
    double e = 0; const int solid = s.gsolid[igroup] == -1 ? -1 : s.solids[s.gsolid[igroup]]->dev;
    if (kinetic) s.check(kml_compute_kinetic_energy(s.ctx, solid, groupbit, &e)); else s.check(kml_compute_strain_energy(s.ctx, solid, groupbit, &e));
    s.input.vars[id + "_s"] = Var(id + "_s", e);
  }
};

// FixChecksolution, reference src/fix_check_solution.cpp:104-200: on output steps, the volume-weighted L2 distance between the particle
// displacements x - x0 and the given u(x0, y0, z0, time): <id>_s = sqrt(sum / vtot) now, <id>_x and <id>_y the time integrals of the error
// and of |u|^2, <id>_z = sqrt(<id>_x / <id>_y)
struct FixCheckSolution : Fix {
  bool set[3] = {false, false, false}; Var val[3];
  void write_restart(std::ostream &os) const override { rput_sets(os, set, val, nullptr); } // src/fix_check_solution.cpp:202-240
  void final_integrate(Sim &s) override {
    bool due = s.ntimestep == s.next_log || s.ntimestep == s.nsteps;
    for (auto &d : s.dumps) due = due || d.next == s.ntimestep;
    if (!due) return;
    double err[3] = {0, 0, 0}, uth[3] = {0, 0, 0}, vtot = 0; // (the reference leaves vtot uninitialised when the group names one solid)
    for_group_solids(s, igroup, [&](SolidH &S) {
      vtot += S.vtot;
      std::vector<double> x(3 * S.np), vol0(S.np);
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_X, x.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_VOL0, vol0.data()));
      for (int64_t ip = 0; ip < S.np; ip++) {
        if (!(S.mask[ip] & groupbit)) continue;
        s.input.vars["x0"] = Var("x0", S.x0[ip][0]); s.input.vars["y0"] = Var("y0", S.x0[ip][1]); s.input.vars["z0"] = Var("z0", S.x0[ip][2]);
        for (int d = 0; d < 3; d++) if (set[d]) {
          const double u = val[d].result(&s.input), e = u - (x[3 * ip + d] - S.x0[ip][d]);
          err[d] += vol0[ip] * (e * e); uth[d] += vol0[ip] * u * u;
        }
      }
    });
    const double e = err[0] + err[1] + err[2], u = uth[0] + uth[1] + uth[2];
    s.input.vars[id + "_s"] = Var(id + "_s", sqrt(e / vtot));
    s.input.vars[id + "_x"] = Var(id + "_x", s.input.vars[id + "_x"].result() + s.dt * e);
    s.input.vars[id + "_y"] = Var(id + "_y", s.input.vars[id + "_y"].result() + s.dt * u);
    s.input.vars[id + "_z"] = Var(id + "_z", sqrt(s.input.vars[id + "_x"].result() / s.input.vars[id + "_y"].result()));
  }
};

// FixCuttingTool, reference src/fix_cutting_tool.cpp:118-285 (2-D): a wedge with its tip at (x_t, y_t) and edges through A and B pushes
// the particles inside it out through the nearer edge with the penalty force K G p (1 - damage) n, added to the particle body force
struct FixCuttingTool : Fix {
  double K = 0; Var v[10]; // x_t, y_t, z_t, vt_x, vt_y, vt_z, x_A, y_A, x_B, y_B
  void write_restart(std::ostream &os) const override { rput(os, K); for (const Var &x : v) rput_var(os, x); } // src/fix_cutting_tool.cpp:287-300
  void initial_integrate(Sim &s) override {
    if (s.dimension == 3) fatal("fix_cuttingtool not supported in 3D\n");
    double q[10]; for (int i = 0; i < 10; i++) q[i] = v[i].result(&s.input);
    const double xt[2] = {q[0], q[1]}, xA[2] = {q[6], q[7]}, xB[2] = {q[8], q[9]};
    double l1[4] = {xA[1] - xt[1], -xA[0] + xt[0], xt[1] * xA[0] - xt[0] * xA[1], 0}, l2[4] = {xB[1] - xt[1], -xB[0] + xt[0], xt[1] * xB[0] - xt[0] * xB[1], 0};
    l1[3] = 1.0 / sqrt(l1[0] * l1[0] + l1[1] * l1[1]); l2[3] = 1.0 / sqrt(l2[0] * l2[0] + l2[1] * l2[1]);
    if (l1[0] * xB[0] + l1[1] * xB[1] + l1[2] < 0) for (int k = 0; k < 3; k++) l1[k] *= -1;
    if (l2[0] * xA[0] + l2[1] * xA[1] + l2[2] < 0) for (int k = 0; k < 3; k++) l2[k] *= -1;
    const double n1[2] = {l1[0] * -l1[3], l1[1] * -l1[3]}, n2[2] = {l2[0] * -l2[3], l2[1] * -l2[3]};
    double ftot[3] = {0, 0, 0};
    // with a group restricted to one solid the reference adds the line's y coefficient where the constant belongs (:222-223); kept
    const bool one = s.gsolid[igroup] != -1; const double k1 = one ? l1[1] : l1[2], k2 = one ? l2[1] : l2[2];
    for_group_solids(s, igroup, [&](SolidH &S) {
      std::vector<double> x(3 * S.np), mass(S.np), dmg(S.np), mbp(3 * S.np);
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_X, x.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_MASS, mass.data()));
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_DAMAGE, dmg.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_MBP, mbp.data()));
      const double G = s.materials[S.mat].km.G;
      for (int64_t ip = 0; ip < S.np; ip++) {
        if (!(mass[ip] > 0) || !(S.mask[ip] & groupbit)) continue;
        const double c1 = l1[0] * x[3 * ip] + l1[1] * x[3 * ip + 1] + k1, c2 = l2[0] * x[3 * ip] + l2[1] * x[3 * ip + 1] + k2;
        if (!(c1 >= 0 && c2 >= 0)) continue;
        const double p1 = fabs(c1 * l1[3]), p2 = fabs(c2 * l2[3]);
        const double p = p1 < p2 ? p1 : p2; const double *n = p1 < p2 ? n1 : n2;
        const double fmag = K * G * p * (1.0 - dmg[ip]);
        for (int d = 0; d < 2; d++) { const double f = fmag * n[d]; mbp[3 * ip + d] += f; ftot[d] += f; }
      }
      s.check(kml_solid_upload(s.ctx, S.dev, KML_P_MBP, mbp.data()));
    });
    s.input.vars[id + "_x"] = Var(id + "_x", ftot[0]); s.input.vars[id + "_y"] = Var(id + "_y", ftot[1]); s.input.vars[id + "_z"] = Var(id + "_z", ftot[2]);
  }
};

// FixContactHertz / FixContactMinPenetration, reference src/fix_contact_hertz.cpp, src/fix_contact_min_penetration.cpp
struct FixContact : Fix {
  void write_restart(std::ostream &os) const override { rput(os, solid1); rput(os, solid2); if (!hertz) rput(os, mu); } // src/fix_contact_hertz.cpp:204-207, src/fix_contact_min_penetration.cpp:261-265
  int solid1 = -1, solid2 = -1; double mu = 0; bool hertz = true;
  void initial_integrate(Sim &s) override {
    double ftot[3];
    if (hertz) s.check(kml_fix_contact_hertz(s.ctx, s.solids[solid1]->dev, s.solids[solid2]->dev, ftot));
    else s.check(kml_fix_contact_min_penetration(s.ctx, s.solids[solid1]->dev, s.solids[solid2]->dev, mu, ftot));
    s.input.vars[id + "_x"] = Var(id + "_x", ftot[0]); s.input.vars[id + "_y"] = Var(id + "_y", ftot[1]); s.input.vars[id + "_z"] = Var(id + "_z", ftot[2]);
  }
};

struct ComputeEnergy : Compute {
  bool kinetic = true;
  void compute_value(Sim &s) override {
    double e = 0;
    if (kinetic) s.check(kml_compute_kinetic_energy(s.ctx, s.gsolid[igroup] == -1 ? -1 : s.solids[s.gsolid[igroup]]->dev, groupbit, &e));
    else s.check(kml_compute_strain_energy(s.ctx, s.gsolid[igroup] == -1 ? -1 : s.solids[s.gsolid[igroup]]->dev, groupbit, &e));
    s.input.vars[id] = Var(id, e);
  }
};
// "if (update->ntimestep != output->next && update->ntimestep != update->nsteps) return;" at the top of every compute_value
// (src/compute_max_plastic_strain.cpp:64-65): nothing is evaluated at step 0 of a run
static bool output_due(const Sim &s) {
  bool due = s.ntimestep == s.next_log || s.ntimestep == s.nsteps;
  for (auto &d : s.dumps) due = due || d.next == s.ntimestep;
  return due;
}
// ComputeMaxPlasticStrain, reference src/compute_max_plastic_strain.cpp:62-106: <id>_Epmax, <id>_Tmax over the group
struct ComputeMaxPlasticStrain : Compute {
  void compute_value(Sim &s) override {
    if (!output_due(s)) return;
    double epmax = 0, tmax = 0;
    for_group_solids(s, igroup, [&](SolidH &S) {
      std::vector<double> ep(S.np), T(S.np);
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_EFF_PLASTIC_STRAIN, ep.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_T, T.data()));
      for (int64_t ip = 0; ip < S.np; ip++) if (S.mask[ip] & groupbit) { epmax = std::max(epmax, ep[ip]); tmax = std::max(tmax, T[ip]); }
    });
    s.input.vars[id + "_Epmax"] = Var(id + "_Epmax", epmax); s.input.vars[id + "_Tmax"] = Var(id + "_Tmax", tmax);
  }
};
// ComputeAverageVelocity, reference src/compute_average_velocity.cpp:60-117: <id>_x/_y/_z = sum(v) / n over the group
struct ComputeAverageVelocity : Compute {
  void compute_value(Sim &s) override {
    if (!output_due(s)) return;
    double sum[3] = {0, 0, 0}; int n = 0;
    for_group_solids(s, igroup, [&](SolidH &S) {
      std::vector<double> v(3 * S.np);
      s.check(kml_solid_download(s.ctx, S.dev, KML_P_V, v.data()));
      for (int64_t ip = 0; ip < S.np; ip++) if (S.mask[ip] & groupbit) { for (int d = 0; d < 3; d++) sum[d] += v[3 * ip + d]; n++; }
    });
    const char *sfx[3] = {"_x", "_y", "_z"};
    for (int d = 0; d < 3; d++) s.input.vars[id + sfx[d]] = Var(id + sfx[d], sum[d] / n);
  }
};
} // namespace

// Modify::read_restart, src/modify.cpp:335-366: the fix is created through its "restart" constructor (group from the stored index) and
// then reads what its write_restart wrote - flags and Vars as (equation, value, constant), src/var.cpp:322-331
std::unique_ptr<Fix> Sim::fix_from_restart(const std::string &id, const std::string &style, int ig, std::istream &is) {
  auto get = [&](auto &v) { is.read(reinterpret_cast<char *>(&v), sizeof v); };
  auto get_var = [&]() { size_t n = 0; get(n); if (n > (1u << 20)) fatal("read_restart: corrupt Var\n"); std::string eq(n, '\0'); is.read(&eq[0], (std::streamsize)n); double v = 0; bool c = false; get(v); get(c); return Var(eq, v, c); };
  auto get_sets = [&](bool *set, Var *val, Var *prev) { for (int d = 0; d < 3; d++) get(set[d]); for (int d = 0; d < 3; d++) if (set[d]) { val[d] = get_var(); if (prev) prev[d] = get_var(); } };
  std::unique_ptr<Fix> fix;
  if (style == "velocity_nodes") { auto f = new FixVelocityNodes(); fix.reset(f); get_sets(f->set, f->val, f->prev); f->mask = POST_UPDATE_GRID_STATE | POST_VELOCITIES_TO_GRID; }
  else if (style == "velocity_particles") { auto f = new FixVelocityParticles(); fix.reset(f); get_sets(f->set, f->val, f->prev); f->mask = INITIAL_INTEGRATE | POST_ADVANCE_PARTICLES; }
  else if (style == "body_force") { auto f = new FixBodyForce(); fix.reset(f); get_sets(f->set, f->val, nullptr); f->mask = POST_PARTICLES_TO_GRID; }
  else if (style == "force_nodes") { auto f = new FixForceNodes(); fix.reset(f); get_sets(f->set, f->val, nullptr); f->mask = POST_PARTICLES_TO_GRID; }
  else if (style == "temperature_nodes") { auto f = new FixTemperatureNodes(); fix.reset(f); f->val = get_var(); f->prev = get_var(); f->mask = POST_UPDATE_GRID_STATE | POST_VELOCITIES_TO_GRID; }
  else if (style == "temperature_particles") { auto f = new FixTemperatureParticles(); fix.reset(f); f->val = get_var(); f->prev = get_var(); f->mask = INITIAL_INTEGRATE | POST_ADVANCE_PARTICLES; }
  else if (style == "kinetic_energy" || style == "strain_energy") { auto f = new FixEnergy(); fix.reset(f); f->kinetic = style == "kinetic_energy"; f->mask = FINAL_INTEGRATE; }
  else if (style == "contact/hertz" || style == "contact/minimize_penetration") {
    auto f = new FixContact(); fix.reset(f); f->hertz = style == "contact/hertz"; get(f->solid1); get(f->solid2); if (!f->hertz) get(f->mu); f->mask = INITIAL_INTEGRATE;
  }
  else if (style == "check_solution") { auto f = new FixCheckSolution(); fix.reset(f); get_sets(f->set, f->val, nullptr); f->mask = FINAL_INTEGRATE; }
  else if (style == "cuttingtool") { auto f = new FixCuttingTool(); fix.reset(f); get(f->K); for (Var &x : f->v) x = get_var(); f->mask = INITIAL_INTEGRATE; }
  else if (style == "initial_velocity_particles") { fix.reset(new FixInitialVelocityParticles()); fix->mask = INITIAL_INTEGRATE; }   // nothing stored: these act at step 1 only
  else if (style == "initial_stress") { fix.reset(new FixInitialStress()); fix->mask = INITIAL_INTEGRATE; }
  else if (style == "initial_velocity_nodes") { fix.reset(new FixInitialVelocityNodes()); fix->mask = POST_UPDATE_GRID_STATE | POST_VELOCITIES_TO_GRID; }
  else fatal("read_restart: fix style " + style + " is not supported\n");
  fix->id = id; fix->style = style; fix->igroup = ig; fix->groupbit = ig >= 0 ? gbitmask[ig] : 0;
  for (const char *sfx : {"_x", "_y", "_z", "_s"}) if (!input.vars.count(id + sfx)) input.vars[id + sfx] = Var(id + sfx, 0.0);
  return fix;
}

// Modify::add_fix, reference src/modify.cpp:95-140 (fix(ID, style, group-ID, args...))
Var Sim::cmd_fix(std::vector<std::string> &a) {
  if (a.size() < 3) fatal("Error: too few arguments for the fix command.\n");
  for (auto &f : fixes) if (f->id == a[0]) fatal("Error: reuse of fix ID.\n");
  const std::string &style = a[1];
  std::unique_ptr<Fix> fix;
  auto group_of = [&](Fix &f) {
    f.igroup = find_group(a[2]);
    if (f.igroup == -1) fatal("Error: could not find group ID " + a[2] + "\n");
    f.groupbit = gbitmask[f.igroup];
  };
  // the value one step earlier: SpecialFunc::replace_all(input->parsev(arg).str(), "time", "(time - dt)"),
  // src/fix_velocity_particles.cpp:84-88, src/fix_temperature_nodes.cpp:50-52
  auto time_shifted = [&](const std::string &arg) {
    std::string e = input.parsev(arg).str(); size_t pos = 0;
    while ((pos = e.find("time", pos)) != std::string::npos) { e.replace(pos, 4, "(time - dt)"); pos += 11; }
    return e;
  };
  if (style == "initial_velocity_particles") {
    auto f = new FixInitialVelocityParticles(); fix.reset(f); group_of(*f);
    if (a.size() != 6) fatal("Error: fix initial_velocity_particles: wrong number of arguments.\n");
    if (gpon[f->igroup] != "particles" && gpon[f->igroup] != "all") fatal("fix_initial_velocity_particles needs to be given a group of particles.\n");
    for (int d = 0; d < 3; d++) if (a[3 + d] != "NULL") { f->val[d] = input.parsev(a[3 + d]); f->set[d] = true; }
    f->mask = INITIAL_INTEGRATE;
  } else if (style == "velocity_nodes") {
    auto f = new FixVelocityNodes(); fix.reset(f); group_of(*f);
    if (a.size() < (size_t)(3 + dimension)) fatal("Error: too few arguments for fix_velocity_nodes.\n");
    if (gpon[f->igroup] != "nodes") fatal("fix_velocity_nodes needs to be given a group of nodes.\n");
    for (int d = 0; d < dimension; d++) {
      if (a[3 + d] == "NULL") continue;
      f->val[d] = input.parsev(a[3 + d]); f->set[d] = true;
      std::string previous = a[3 + d]; // "time" -> "time - dt", src/fix_velocity_nodes.cpp:66-80
      size_t pos = 0;
      while ((pos = previous.find("time", pos)) != std::string::npos) { previous.replace(pos, 4, "time - dt"); pos += 9; }
      f->prev[d] = input.parsev(previous);
    }
    f->mask = POST_UPDATE_GRID_STATE | POST_VELOCITIES_TO_GRID;
  } else if (style == "velocity_particles") {
    auto f = new FixVelocityParticles(); fix.reset(f); group_of(*f);
    if (a.size() < (size_t)(3 + dimension)) fatal("Error: too few arguments for fix_velocity_nodes.\n");
    if (gpon[f->igroup] != "particles") fatal("fix_velocity_nodes needs to be given a group of nodes" + gpon[f->igroup] + ", " + a[2] + " is a group of " + gpon[f->igroup] + ".\n");
    for (int d = 0; d < dimension; d++) {
      if (a[3 + d] == "NULL") continue;
      f->val[d] = input.parsev(a[3 + d]); f->set[d] = true;
      f->prev[d] = input.parsev(time_shifted(a[3 + d]));
    }
    f->mask = INITIAL_INTEGRATE | POST_ADVANCE_PARTICLES;
  } else if (style == "temperature_particles" || style == "temperature_nodes") {
    if (a.size() < 4) fatal("Error: not enough arguments.\nUsage: fix(fix-ID, temperature_nodes, group-ID, T)\n");
    const bool nodes = style == "temperature_nodes";
    Var val = input.parsev(a[3]), prev = input.parsev(time_shifted(a[3]));
    if (nodes) {
      auto f = new FixTemperatureNodes(); fix.reset(f); group_of(*f);
      if (gpon[f->igroup] != "nodes") fatal("fix_temperature_nodes needs to be given a group of nodes" + gpon[f->igroup] + ", " + a[2] + " is a group of " + gpon[f->igroup] + ".\n");
      f->val = val; f->prev = prev; f->mask = POST_UPDATE_GRID_STATE | POST_VELOCITIES_TO_GRID;
    } else {
      auto f = new FixTemperatureParticles(); fix.reset(f); group_of(*f);
      if (gpon[f->igroup] != "particles") fatal("fix_temperature_nodes needs to be given a group of nodes" + gpon[f->igroup] + ", " + a[2] + " is a group of " + gpon[f->igroup] + ".\n");
      f->val = val; f->prev = prev; f->mask = INITIAL_INTEGRATE | POST_ADVANCE_PARTICLES;
    }
  } else if (style == "initial_stress") {
    auto f = new FixInitialStress(); fix.reset(f); group_of(*f);
    if (a.size() < 9) fatal("Error: not enough arguments.\nUsage: fix(fix-ID, initial_stress, group-ID, sigma_xx, sigma_yy, sigma_zz, sigma_yz, sigma_xz, sigma_xy)\n");
    if (a.size() > 9) fatal("Error: too many arguments.\nUsage: fix(fix-ID, initial_stress, group-ID, sigma_xx, sigma_yy, sigma_zz, sigma_yz, sigma_xz, sigma_xy)\n");
    if (gpon[f->igroup] != "particles" && gpon[f->igroup] != "all") fatal("fix_initial_stress needs to be given a group of particles" + gpon[f->igroup] + ", " + a[2] + " is a group of " + gpon[f->igroup] + ".\n");
    for (int k = 0; k < 6; k++) if (a[3 + k] != "NULL") { f->val[k] = input.parsev(a[3 + k]); f->set[k] = true; }
    f->mask = INITIAL_INTEGRATE;
  } else if (style == "initial_velocity_nodes") {
    auto f = new FixInitialVelocityNodes(); fix.reset(f); group_of(*f);
    if (a.size() < 6) fatal("Error: too few arguments for fix_initial_velocity_nodes: requires at least 6 arguments. " + std::to_string(a.size()) + " received.\n");
    if (gpon[f->igroup] != "nodes") fatal("fix_initial_velocity_nodes needs to be given a group of nodes" + gpon[f->igroup] + ", " + a[2] + " is a group of " + gpon[f->igroup] + ".\n");
    for (int d = 0; d < 3; d++) if (a[3 + d] != "NULL") { f->val[d] = input.parsev(a[3 + d]); f->set[d] = true; }
    f->mask = POST_UPDATE_GRID_STATE | POST_VELOCITIES_TO_GRID;
  } else if (style == "cuttingtool") {
    auto f = new FixCuttingTool(); fix.reset(f); group_of(*f);
    if (a.size() < 14) fatal("Error: not enough arguments.\nUsage: fix(fix-ID, cuttingtool, group, K, x_tip, y_tip, z_tip, vx_tip, vy_tip, vz_tip, xA, yA, xB, yB)\n");
    if (gpon[f->igroup] != "particles" && gpon[f->igroup] != "all") fatal("fix_cuttingtool needs to be given a group of nodes" + gpon[f->igroup] + ", " + a[2] + " is a group of " + gpon[f->igroup] + ".\n");
    f->K = input.parsev(a[3]).result(&input);
    for (int i = 0; i < 10; i++) f->v[i] = input.parsev(a[4 + i]);
    f->mask = INITIAL_INTEGRATE;
  } else if (style == "check_solution") {
    auto f = new FixCheckSolution(); fix.reset(f); group_of(*f);
    if (a.size() < (size_t)(3 + dimension)) fatal("Error: too few arguments for fix_check_solution: requires at least " + std::to_string(3 + dimension) + " arguments. " + std::to_string(a.size()) + " received.\n");
    if (gpon[f->igroup] != "nodes" && gpon[f->igroup] != "all") fatal("_check_solution needs to be given a group of nodes" + gpon[f->igroup] + ", " + a[2] + " is a group of " + gpon[f->igroup] + ".\n");
    for (int d = 0; d < dimension; d++) if (a[3 + d] != "NULL") { f->val[d] = input.parsev(a[3 + d]); f->set[d] = true; }
    f->mask = FINAL_INTEGRATE;
  } else if (style == "body_force") {
    auto f = new FixBodyForce(); fix.reset(f); group_of(*f);
    if (a.size() < (size_t)(3 + dimension)) fatal("Error: too few arguments for fix_body_force.\n");
    for (int d = 0; d < dimension; d++) if (a[3 + d] != "NULL") {
      f->val[d] = input.parsev(a[3 + d]); f->set[d] = true;
      const std::string &e = f->val[d].eq();
      if (!f->val[d].is_constant() && (e.find("x0") != std::string::npos || e.find("y0") != std::string::npos || e.find("z0") != std::string::npos))
        fatal("fix body_force: position-dependent body forces are not supported by this build.\n");
    }
    f->mask = POST_PARTICLES_TO_GRID;
  } else if (style == "force_nodes") {
    auto f = new FixForceNodes(); fix.reset(f); group_of(*f);
    if (a.size() < 6) fatal("Error: too few arguments for fix_force_nodes: requires at least 6 arguments. " + std::to_string(a.size()) + " received.\n");
    if (gpon[f->igroup] != "nodes") fatal("fix_force_nodes needs to be given a group of nodes" + gpon[f->igroup] + ", " + a[2] + " is a group of " + gpon[f->igroup] + ".\n");
    for (int d = 0; d < 3; d++) if (a[3 + d] != "NULL") { f->val[d] = input.parsev(a[3 + d]); f->set[d] = true; }
    f->mask = POST_PARTICLES_TO_GRID;
  } else if (style == "kinetic_energy" || style == "strain_energy") {
    auto f = new FixEnergy(); fix.reset(f); group_of(*f);
    f->kinetic = style == "kinetic_energy";
    if (gpon[f->igroup] != "particles" && gpon[f->igroup] != "all") fatal("fix_" + style + " needs to be given a group of particles" + gpon[f->igroup] + ", " + a[2] + " is a group of " + gpon[f->igroup] + ".\n");
    input.vars[a[0] + "_s"] = Var(a[0] + "_s", 0.0);
    f->mask = FINAL_INTEGRATE;
  } else if (style == "contact/hertz" || style == "contact/minimize_penetration") {
    auto f = new FixContact(); fix.reset(f);
    f->hertz = style == "contact/hertz";
    if (a.size() < (size_t)(f->hertz ? 4 : 5)) fatal("Error: not enough arguments.\n");
    f->solid1 = find_solid(a[2]); if (f->solid1 < 0) fatal("Error: solid " + a[2] + " unknown.\n");
    f->solid2 = find_solid(a[3]); if (f->solid2 < 0) fatal("Error: solid " + a[3] + " unknown.\n");
    if (!f->hertz) f->mu = input.parsev(a[4]);
    f->igroup = find_group(a[2]); // Fix::Fix looks the third argument up as a group (src/fix.cpp:36): a solid name gives -1, and that is what restart files hold
    f->mask = INITIAL_INTEGRATE;
  } else fatal("fix style " + style + " is outside the hot path covered by this build (see DESIGN.md).\n");
  fix->id = a[0]; fix->style = style;
  for (const char *sfx : {"_x", "_y", "_z", "_s"}) // Fix::Fix publishes the fix's outputs as variables from the start, src/fix.cpp:41-44
    if (!input.vars.count(a[0] + sfx)) input.vars[a[0] + sfx] = Var(a[0] + sfx, 0.0);
  fixes.push_back(std::move(fix));
  return Var(0);
}

Var Sim::cmd_compute(std::vector<std::string> &a) { // Modify::add_compute
  if (a.size() < 3) fatal("Error: too few arguments for the compute command.\n");
  Compute *c = nullptr;
  if (a[1] == "kinetic_energy" || a[1] == "strain_energy") { auto e = new ComputeEnergy(); e->kinetic = a[1] == "kinetic_energy"; c = e; input.vars[a[0]] = Var(a[0], 0); }
  else if (a[1] == "max_plastic_strain") { c = new ComputeMaxPlasticStrain(); input.vars[a[0] + "_Epmax"] = Var(a[0] + "_Epmax", 0); input.vars[a[0] + "_Tmax"] = Var(a[0] + "_Tmax", 0); }
  else if (a[1] == "average_velocity") { c = new ComputeAverageVelocity(); for (const char *sfx : {"_x", "_y", "_z"}) input.vars[a[0] + sfx] = Var(a[0] + sfx, 0); }
  else fatal("compute style " + a[1] + " is outside the hot path covered by this build.\n");
  c->id = a[0]; c->style = a[1];
  c->igroup = find_group(a[2]); if (c->igroup == -1) fatal("Error: could not find group ID " + a[2] + "\n");
  c->groupbit = gbitmask[c->igroup];
  if (a[1] != "kinetic_energy" && a[1] != "strain_energy" && gpon[c->igroup] != "particles" && gpon[c->igroup] != "all")
    fatal("compute_" + a[1] + " needs to be given a group of particles" + gpon[c->igroup] + ", " + a[2] + " is a group of " + gpon[c->igroup] + ".\n");
  computes.emplace_back(c);
  return Var(0);
}

void Sim::hooks(int which) {
  for (auto &f : fixes) {
    if (!(f->mask & which)) continue;
    switch (which) {
    case INITIAL_INTEGRATE: f->initial_integrate(*this); break;
    case POST_PARTICLES_TO_GRID: f->post_particles_to_grid(*this); break;
    case POST_UPDATE_GRID_STATE: f->post_update_grid_state(*this); break;
    case POST_ADVANCE_PARTICLES: f->post_advance_particles(*this); break;
    case POST_VELOCITIES_TO_GRID: f->post_velocities_to_grid(*this); break;
    case FINAL_INTEGRATE: f->final_integrate(*this); break;
    default: break;
    }
  }
}

// ------------------------------------------------------------------ output ------------------
Var Sim::cmd_dump(std::vector<std::string> &a) { // Output::add_dump: dump(ID, group, style, N, file, fields...)
  if (a.size() < 5) fatal("Error: too few arguments for dump.\n");
  Dump d; d.id = a[0]; d.igroup = find_group(a[1]); d.style = a[2]; d.every = (int)(double)input.parsev(a[3]); d.filename = a[4];
  d.fields.assign(a.begin() + 5, a.end());
  if (d.every <= 0) fatal("Error every_dump = 0 does not make sense.\n");
  dumps.push_back(d); return Var(0);
}

// DumpParticleGz / DumpGridGz (src/dump_particle_gz.cpp, src/dump_grid_gz.cpp) write the same text through gzstream
static void emit_dump(const std::string &fn, const std::string &text, bool gz) {
  if (!gz) {
    std::ofstream f(fn, std::ios::out | std::ios::binary);
    if (!f) fatal("Error: cannot write in file: " + fn + ".\n");
    f << text; return;
  }
  gzFile f = gzopen(fn.c_str(), "wb");
  if (!f) fatal("Error: cannot write in file: " + fn + ".\n");
  size_t off = 0;
  while (off < text.size()) { const unsigned n = (unsigned)std::min<size_t>(text.size() - off, 1u << 30); if (gzwrite(f, text.data() + off, n) <= 0) { gzclose(f); fatal("Error: cannot write in file: " + fn + ".\n"); } off += n; }
  gzclose(f);
}

static void write_particle_dump(Sim &s, const Dump &d) { // DumpParticle::write, reference src/dump_particle.cpp:52-161
  std::string fn = d.filename; size_t star = fn.find('*');
  if (star != std::string::npos) fn = fn.substr(0, star) + std::to_string(s.ntimestep) + fn.substr(star + 1);
  std::ostringstream os;
  int64_t total = 0; for (auto &S : s.solids) { s.sync(*S); total += S->np; }
  os << "ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n" << total << "\nITEM: BOX BOUNDS sm sm sm\n";
  for (int k = 0; k < 3; k++) os << s.boxlo[k] << " " << s.boxhi[k] << "\n";
  os << "ITEM: ATOMS id type tag ";
  for (auto &f : d.fields) os << f << " ";
  os << "\n";
  for (size_t is = 0; is < s.solids.size(); is++) {
    SolidH &S = *s.solids[is]; const int64_t n = S.np;
    std::vector<int64_t> tag(n); std::vector<double> x(3 * n), v(3 * n), sig(9 * n), R(9 * n), vol(n), mass(n), dmg(n), dmgi(n), ep(n), epdot(n), T(n), ie(n), rho(n);
    s.check(kml_solid_download(s.ctx, S.dev, KML_P_PTAG, tag.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_X, x.data()));
    s.check(kml_solid_download(s.ctx, S.dev, KML_P_V, v.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_SIGMA, sig.data()));
    s.check(kml_solid_download(s.ctx, S.dev, KML_P_VOL, vol.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_MASS, mass.data()));
    s.check(kml_solid_download(s.ctx, S.dev, KML_P_DAMAGE, dmg.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_DAMAGE_INIT, dmgi.data()));
    s.check(kml_solid_download(s.ctx, S.dev, KML_P_EFF_PLASTIC_STRAIN, ep.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_EFF_PLASTIC_STRAIN_RATE, epdot.data()));
    s.check(kml_solid_download(s.ctx, S.dev, KML_P_IENERGY, ie.data())); s.check(kml_solid_download(s.ctx, S.dev, KML_P_RHO, rho.data()));
    if (s.temp) s.check(kml_solid_download(s.ctx, S.dev, KML_P_T, T.data()));
    if (s.is_TL) s.check(kml_solid_download(s.ctx, S.dev, KML_P_R, R.data()));
    for (int64_t i = 0; i < n; i++) {
      double sg[9];
      if (s.is_TL) { // sigma_ = R sigma R^T
        const double *r = &R[9 * i], *q = &sig[9 * i]; double t[9];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) t[3 * a + b] = r[3 * a] * q[b] + r[3 * a + 1] * q[3 + b] + r[3 * a + 2] * q[6 + b];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) sg[3 * a + b] = t[3 * a] * r[3 * b] + t[3 * a + 1] * r[3 * b + 1] + t[3 * a + 2] * r[3 * b + 2];
      } else memcpy(sg, &sig[9 * i], sizeof sg);
      os << tag[i] << " " << is + 1 << " " << tag[i] << " ";
      for (auto &f : d.fields) {
        if (f == "x") os << x[3 * i]; else if (f == "y") os << x[3 * i + 1]; else if (f == "z") os << x[3 * i + 2];
        else if (f == "x0") os << S.x0[i][0]; else if (f == "y0") os << S.x0[i][1]; else if (f == "z0") os << S.x0[i][2];
        else if (f == "vx") os << v[3 * i]; else if (f == "vy") os << v[3 * i + 1]; else if (f == "vz") os << v[3 * i + 2];
        else if (f == "s11") os << sg[0]; else if (f == "s22") os << sg[4]; else if (f == "s33") os << sg[8];
        else if (f == "s12") os << sg[1]; else if (f == "s13") os << sg[2]; else if (f == "s23") os << sg[5];
        else if (f == "seq") os << sqrt(3. / 2.) * [&] { double tr = (sg[0] + sg[4] + sg[8]) / 3, q = 0; for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { double e = sg[3 * a + b] - (a == b ? tr : 0); q += e * e; } return sqrt(q); }();
        else if (f == "damage") os << dmg[i]; else if (f == "damage_init") os << dmgi[i]; else if (f == "volume") os << vol[i];
        else if (f == "mass") os << mass[i]; else if (f == "ep") os << ep[i]; else if (f == "epdot") os << epdot[i];
        else if (f == "ienergy") os << ie[i]; else if (f == "T") os << T[i]; else if (f == "rho") os << rho[i];
        else os << 0;
        os << " ";
      }
      os << "\n";
    }
  }
  emit_dump(fn, os.str(), d.style == "particle/gz");
}

static void write_grid_dump(Sim &s, const Dump &d) { // DumpGrid::write, reference src/dump_grid.cpp:50-150
  std::string fn = d.filename; size_t star = fn.find('*');
  if (star != std::string::npos) fn = fn.substr(0, star) + (s.nranks > 1 ? "proc-" + std::to_string(s.rank) + "." : "") + std::to_string(s.ntimestep) + fn.substr(star + 1);
  std::ostringstream os;
  std::vector<GridH *> grids;
  for (auto &S : s.solids) if (std::find(grids.begin(), grids.end(), S->grid) == grids.end()) grids.push_back(S->grid);
  int64_t total = 0; for (auto g : grids) total += g->nnodes;
  os << "ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n" << total << "\nITEM: BOX BOUNDS sm sm sm\n";
  for (int k = 0; k < 3; k++) os << s.boxlo[k] << " " << s.boxhi[k] << "\n";
  const bool gz = d.style == "grid/gz"; // the gz variant has no tag column (src/dump_grid_gz.cpp:99,108)
  os << (gz ? "ITEM: ATOMS id type " : "ITEM: ATOMS id type tag ");
  for (auto &f : d.fields) os << f << " ";
  os << "\n";
  int igrid = 0;
  for (auto g : grids) {
    const int64_t n = g->nnodes; const kml_grid_desc &gd = g->desc;
    std::vector<double> x(3 * n), v(3 * n), mb(3 * n), mass(n), T(n); std::vector<int> ntype(3 * n), rigid(n);
    s.check(kml_grid_download(s.ctx, g->id, KML_N_X, x.data())); s.check(kml_grid_download(s.ctx, g->id, KML_N_V, v.data()));
    s.check(kml_grid_download(s.ctx, g->id, KML_N_MB, mb.data())); s.check(kml_grid_download(s.ctx, g->id, KML_N_MASS, mass.data()));
    s.check(kml_grid_download(s.ctx, g->id, KML_N_NTYPE, ntype.data())); s.check(kml_grid_download(s.ctx, g->id, KML_N_RIGID, rigid.data()));
    if (s.temp) s.check(kml_grid_download(s.ctx, g->id, KML_N_T, T.data()));
    for (int64_t i = 0; i < n; i++) {
      const int64_t tag = i + (int64_t)gd.goff * gd.n[1] * gd.n[2]; // ntag = nz ny i + nz j + k of the global grid, src/grid.cpp:251
      os << tag << " " << igrid + 1 << " ";
      if (!gz) os << tag << " ";
      for (auto &f : d.fields) {
        if (f == "x") os << x[3 * i]; else if (f == "y") os << x[3 * i + 1]; else if (f == "z") os << x[3 * i + 2];
        else if (f == "vx") os << v[3 * i]; else if (f == "vy") os << v[3 * i + 1]; else if (f == "vz") os << v[3 * i + 2];
        else if (f == "bx") os << mb[3 * i]; else if (f == "by") os << mb[3 * i + 1]; else if (f == "bz") os << mb[3 * i + 2];
        else if (f == "mass") os << mass[i]; else if (f == "mask") os << g->mask[i];
        else if (f == "ntypex") os << ntype[3 * i]; else if (f == "ntypey") os << ntype[3 * i + 1]; else if (f == "ntypez") os << ntype[3 * i + 2];
        else if (f == "rigid") os << rigid[i]; else if (f == "T") os << T[i];
        os << " ";
      }
      os << "\n";
    }
    igrid++;
  }
  emit_dump(fn, os.str(), d.style == "grid/gz");
}

void Sim::output_setup() { // Output::setup, src/output.cpp:64-126
  for (auto &d : dumps) d.next = (ntimestep / d.every) * d.every + d.every;
  if (!quiet) {
    std::string hdr;
    for (auto &f : log_fields) hdr += (f == "step" ? "Step" : (f == "time" ? "Time" : f)) + "\t";
    std::cout << hdr << "\n";
    if (logfile.is_open()) logfile << hdr << "\n";
  }
  if (every_log) { next_log = (ntimestep / every_log) * every_log + every_log; if (laststep != 0) next_log = std::min(next_log, laststep); }
  else next_log = laststep;
}

// WriteRestart::write, src/write_restart.cpp:51-88: header, Update, Domain (box, regions, materials, solids), Group, Modify - every block in
// the byte layout of the reference's own write_restart members, so that read_restart(file) of the reference accepts the file.  The
// version string is ours unless KML_RESTART_VERSION names another one (the byte-for-byte test compares with the reference build's).
void Sim::write_restart(const std::string &pattern) {
  std::string fn = pattern; const size_t star = fn.find('*');
  if (star != std::string::npos) fn = fn.substr(0, star) + (nranks > 1 ? "proc-" + std::to_string(rank) + "." : "") + std::to_string(ntimestep) + fn.substr(star + 1);
  else fn = pattern + "proc-" + std::to_string(rank) + "."; // src/write_restart.cpp:64-65
  if (!quiet && rank == 0) std::cout << "write " << fn << std::endl;
  std::ofstream os(fn, std::ios::out | std::ios::binary);
  if (!os) fatal("Error: cannot write in file: " + fn + ".\n");
  // header (rank 0), src/write_restart.cpp:92-101
  if (rank == 0) {
    const char *ver = getenv("KML_RESTART_VERSION");
    int flag = 0; rput(os, flag); rput_str(os, ver && *ver ? ver : "karamelo-b200");
    flag = 1; rput(os, flag); rput(os, dimension);
    flag = 2; rput(os, flag); rput(os, nranks);
    flag = -1; rput(os, flag);
  }
  // Update::write_restart, src/update.cpp:220-275
  rput_str(os, method_type); { const bool t = temp; rput(os, t); }
  rput_str(os, scheme_style); rput(os, sub_method); rput(os, PIC_FLIP_script); rput(os, shape_function);
  { const size_t n = additional_args.size(); rput(os, n); for (auto &a : additional_args) rput_str(os, a); }
  rput(os, atime); rput(os, ntimestep); rput(os, dt); rput(os, dt_factor); { const bool c = dt_constant; rput(os, c); }
  // Domain::write_restart, src/domain.cpp:574-631
  for (int d = 0; d < 3; d++) rput(os, boxlo[d]);
  for (int d = 0; d < 3; d++) rput(os, boxhi[d]);
  for (int d = 0; d < 3; d++) rput(os, sublo[d]);
  for (int d = 0; d < 3; d++) rput(os, subhi[d]);
  { const bool ax = axisymmetric; rput(os, ax); } rput(os, np_total);
  if (!is_TL) rput(os, grid->cellsize);
  { const int n = (int)regions.size(); rput(os, n); }
  for (auto &r : regions) { rput_str(os, r->id); rput_str(os, r->style); r->write_restart(os); }
  // Material::write_restart, src/material.cpp:384-497
  static const char *eos_style[] = {"", "linear", "shock", "fluid"}, *str_style[] = {"", "linear", "plastic", "johnson_cook", "swift", "fluid"};
  { const int n = (int)eoss.size(); rput(os, n); }
  for (auto &e : eoss) {
    rput_str(os, e.id); rput_str(os, eos_style[e.type]); rput(os, e.rho0); rput(os, e.K);
    if (e.type == KML_EOS_SHOCK) { rput(os, e.c0); rput(os, e.S); rput(os, e.Gamma); rput(os, e.Tr); rput(os, e.cv); rput(os, e.Q1); rput(os, e.Q2); }
    else if (e.type == KML_EOS_FLUID) rput(os, e.Gamma);
  }
  { const int n = (int)strengths.size(); rput(os, n); }
  for (auto &t : strengths) {
    rput_str(os, t.id); rput_str(os, str_style[t.type]); rput(os, t.G);
    if (t.type == KML_STRENGTH_PLASTIC) rput(os, t.A);
    else if (t.type == KML_STRENGTH_JOHNSON_COOK) { rput(os, t.A); rput(os, t.B); rput(os, t.n); rput(os, t.m); rput(os, t.epsdot0); rput(os, t.C); rput(os, t.Tr); rput(os, t.Tm); }
    else if (t.type == KML_STRENGTH_SWIFT) { rput(os, t.A); rput(os, t.B); rput(os, t.C); rput(os, t.n); }
  }
  { const int n = (int)damages.size(); rput(os, n); }
  for (auto &d : damages) { rput_str(os, d.id); rput_str(os, "damage_johnson_cook"); rput(os, d.d1); rput(os, d.d2); rput(os, d.d3); rput(os, d.d4); rput(os, d.d5); rput(os, d.epsdot0); rput(os, d.Tr); rput(os, d.Tm); }
  { const int n = (int)temperatures.size(); rput(os, n); }
  for (auto &t : temperatures) { rput_str(os, t.id); rput_str(os, "plastic_work"); rput(os, t.chi); rput(os, t.kappa); rput(os, t.cp); rput(os, t.alpha); rput(os, t.T0); rput(os, t.Tm); }
  { const int n = (int)materials.size(); rput(os, n); }
  for (auto &m : materials) {
    rput_str(os, m.id);
    static const int ref_type[] = {1, 2, 3, 0}; // Material::constitutive_model: RIGID 0, LINEAR 1, NEO_HOOKEAN 2, SHOCK 3 (src/material.h:114-119)
    const int type = ref_type[m.km.type]; rput(os, type);
    if (type == 3) { rput(os, m.ieos); rput(os, m.istrength); rput(os, m.idamage); rput(os, m.itemperature); }
    else if (type == 1 || type == 2) { rput(os, m.km.rho0); rput(os, m.km.E); rput(os, m.km.nu); rput(os, m.km.cp); rput(os, m.km.kappa); }
  }
  { const int flag = -2; rput(os, flag); }
  // solids: Domain::write_restart + Solid::write_restart, src/solid.cpp:2841-2887 (matrices in Eigen's column-major order)
  { const int n = (int)solids.size(); rput(os, n); }
  for (auto &Sp : solids) {
    SolidH &S = *Sp; sync(S); const int64_t n = S.np;
    rput_str(os, S.id);
    for (int d = 0; d < 3; d++) rput(os, S.solidlo[d]);
    for (int d = 0; d < 3; d++) rput(os, S.solidhi[d]);
    for (int d = 0; d < 3; d++) { const double v = std::max(S.solidlo[d], sublo[d]); rput(os, v); } // solidsublo / solidsubhi, src/solid.cpp:1840-1846
    for (int d = 0; d < 3; d++) { const double v = std::min(S.solidhi[d], subhi[d]); rput(os, v); }
    rput(os, S.np_created); { const int nl = (int)n; rput(os, nl); } { const int nc = is_CPDI ? 1 << dimension : 0; rput(os, nc); } // Solid::nc (src/solid.cpp:83-88); the particle domains themselves are not in the layout
    rput(os, S.mat); rput(os, S.grid->cellsize);
    std::vector<int64_t> tag(n); std::vector<int> mask(n);
    std::vector<double> x0(3 * n), x(3 * n), v(3 * n), sig(9 * n), eel(9 * n), pk1, F(9 * n), J(n), vol0(n), rho0(n), ep(n), epdot(n), dmg(n), dmgi(n), T, ie(n);
    check(kml_solid_download(ctx, S.dev, KML_P_PTAG, tag.data())); check(kml_solid_download(ctx, S.dev, KML_P_MASK, mask.data()));
    check(kml_solid_download(ctx, S.dev, KML_P_X0, x0.data())); check(kml_solid_download(ctx, S.dev, KML_P_X, x.data())); check(kml_solid_download(ctx, S.dev, KML_P_V, v.data()));
    check(kml_solid_download(ctx, S.dev, KML_P_SIGMA, sig.data())); check(kml_solid_download(ctx, S.dev, KML_P_STRAIN_EL, eel.data()));
    if (is_TL) { pk1.resize(9 * n); check(kml_solid_download(ctx, S.dev, KML_P_VOL0PK1, pk1.data())); }
    check(kml_solid_download(ctx, S.dev, KML_P_FDEF, F.data())); check(kml_solid_download(ctx, S.dev, KML_P_J, J.data()));
    check(kml_solid_download(ctx, S.dev, KML_P_VOL0, vol0.data())); check(kml_solid_download(ctx, S.dev, KML_P_RHO0, rho0.data()));
    check(kml_solid_download(ctx, S.dev, KML_P_EFF_PLASTIC_STRAIN, ep.data())); check(kml_solid_download(ctx, S.dev, KML_P_EFF_PLASTIC_STRAIN_RATE, epdot.data()));
    check(kml_solid_download(ctx, S.dev, KML_P_DAMAGE, dmg.data())); check(kml_solid_download(ctx, S.dev, KML_P_DAMAGE_INIT, dmgi.data()));
    if (temp) { T.resize(n); check(kml_solid_download(ctx, S.dev, KML_P_T, T.data())); }
    check(kml_solid_download(ctx, S.dev, KML_P_IENERGY, ie.data()));
    auto put_mat = [&](const double *m) { for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) rput(os, m[3 * r + c]); };
    for (int64_t i = 0; i < n; i++) {
      rput(os, tag[i]); os.write((const char *)&x0[3 * i], 24); os.write((const char *)&x[3 * i], 24); os.write((const char *)&v[3 * i], 24);
      put_mat(&sig[9 * i]); put_mat(&eel[9 * i]); if (is_TL) put_mat(&pk1[9 * i]); put_mat(&F[9 * i]);
      rput(os, J[i]); rput(os, vol0[i]); rput(os, rho0[i]); rput(os, ep[i]); rput(os, epdot[i]); rput(os, dmg[i]); rput(os, dmgi[i]);
      if (temp) rput(os, T[i]);
      rput(os, ie[i]); rput(os, mask[i]);
    }
  }
  // Group::write_restart, src/group.cpp:470-504
  rput(os, ngroup);
  for (int ig = 1; ig < ngroup; ig++) {
    rput_str(os, gnames[ig]); rput(os, gbitmask[ig]); const bool p_or_n = gpon[ig] != "particles"; rput(os, p_or_n); rput(os, gsolid[ig]); rput(os, gregion[ig]);
  }
  // Modify::write_restart, src/modify.cpp:312-331
  { const size_t n = fixes.size(); rput(os, n); }
  for (auto &f : fixes) { rput_str(os, f->id); rput_str(os, f->style); rput(os, f->igroup); f->write_restart(os); }
}

void Sim::output_write(int64_t step) { // Output::write, src/output.cpp:128-199
  for (auto &d : dumps)
    if (d.next == step) { if (d.style == "particle" || d.style == "particle/gz") write_particle_dump(*this, d); else if (d.style == "grid" || d.style == "grid/gz") write_grid_dump(*this, d); d.next += d.every; }
  if (restart_every && next_restart == step) { write_restart(restart_name); next_restart += restart_every; } // src/output.cpp:152-155
  if (next_log == step || step == 0) {
    for (auto &c : computes) c->compute_value(*this); // Modify::run_computes
    if (!quiet) {
      std::ostringstream o;
      for (auto &f : log_fields) {
        if (f == "step") o << ntimestep; else if (f == "dt") o << dt; else if (f == "time") o << atime;
        else { auto it = input.vars.find(f); if (it == input.vars.end()) fatal("Error: unknown log keyword " + f + ".\n"); o << it->second.result(&input); }
        o << "\t";
      }
      o << "\n"; std::cout << o.str(); if (logfile.is_open()) logfile << o.str();
    }
    if (next_log == step) next_log += every_log ? every_log : 0;
  }
}

// ------------------------------------------------------------------ schemes -----------------
// USL::run src/usl.cpp:34-94, MUSL::run src/musl.cpp:34-95, USF::run src/usf.cpp:34-94
void Sim::run(Var condition) {
  const bool usl = scheme_style == "usl", musl = scheme_style == "musl", usf = scheme_style == "usf";
  output_write(ntimestep);
  while ((bool)condition.result(&input)) {
    ntimestep++; input.vars["timestep"] = Var("timestep", (double)ntimestep); // Update::update_timestep
    check(kml_compute_grid_weight_functions_and_gradients(ctx));
    check(kml_reset(ctx));
    hooks(INITIAL_INTEGRATE);
    if (!usf) {
      check(kml_particles_to_grid(ctx));
      hooks(POST_PARTICLES_TO_GRID);
      check(kml_update_grid_state(ctx));
      hooks(POST_UPDATE_GRID_STATE);
      if (usl) check(kml_compute_rate_deformation_gradient(ctx, 0));
      check(kml_grid_to_points(ctx));
      hooks(POST_GRID_TO_POINT);
      check(kml_advance_particles(ctx));
      hooks(POST_ADVANCE_PARTICLES);
      if (musl) { check(kml_velocities_to_grid(ctx)); hooks(POST_VELOCITIES_TO_GRID); }
      check(kml_update_grid_positions(ctx));
      if (musl) check(kml_compute_rate_deformation_gradient(ctx, 1));
      check(kml_update_deformation_gradient(ctx));
      check(kml_update_stress(ctx, musl ? 1 : 0));
    } else {
      check(kml_particles_to_grid_USF_1(ctx));
      hooks(POST_UPDATE_GRID_STATE);
      check(kml_compute_rate_deformation_gradient(ctx, 1));
      check(kml_update_deformation_gradient(ctx));
      check(kml_update_stress(ctx, 1));
      check(kml_particles_to_grid_USF_2(ctx));
      hooks(POST_PARTICLES_TO_GRID);
      check(kml_update_grid_state(ctx));
      hooks(POST_UPDATE_GRID_STATE);
      check(kml_grid_to_points(ctx));
      hooks(POST_GRID_TO_POINT);
      check(kml_advance_particles(ctx));
      hooks(POST_ADVANCE_PARTICLES);
      check(kml_update_grid_positions(ctx));
    }
    check(kml_exchange_particles(ctx));
    if (dt_stale) { check(kml_get_dt(ctx, &dt)); input.vars["dt"] = Var("dt", dt); dt_stale = false; } // resolved long ago by this step's grid update: no stall
    atime += (ntimestep - atimestep) * dt; atimestep = ntimestep; input.vars["time"] = Var("time", atime); // Update::update_time
    if (!dt_constant) { // Method::adjust_dt
      // The engine only enqueues the reduction; the value is read back when somebody needs it (kml_get_dt below, or the engine's own grid
      // update of the next step).  Nothing on the host needs it before the next step's stages unless a fix or the output evaluates
      // expressions at the start of the step - then it is fetched here like the reference does.
      const bool due_next = ntimestep == next_log || ntimestep == nsteps || (restart_every && next_restart == ntimestep);
      bool need_now = due_next || !dumps.empty() || maxtime != -1;
      for (auto &f : fixes) if (f->needs_time(*this)) need_now = true;
      check(kml_adjust_dt(ctx, dt_factor, need_now ? &dt : nullptr));
      if (need_now) input.vars["dt"] = Var("dt", dt); else dt_stale = true;
    }
    else { // set_dt: adjust_dt (which returns the device error word) is skipped, so ask for it - the reference stops inside the failing step
      unsigned fl = 0; check(kml_error_flags(ctx, &fl));
      if (fl) fatal("device error flags set: " + std::to_string(fl) + " (bit0 particle left the domain, bit1 J<=0, bit2 dtCFL invalid, bit3 polar decomposition failed)\n");
    }
    hooks(FINAL_INTEGRATE);
    if (maxtime != -1 && atime > maxtime) { nsteps = ntimestep; output_write(ntimestep); break; }
    bool due = ntimestep == next_log || ntimestep == nsteps || (restart_every && next_restart == ntimestep);
    for (auto &d : dumps) due = due || d.next == ntimestep;
    if (due) output_write(ntimestep);
  }
  if (dt_stale) { check(kml_get_dt(ctx, &dt)); input.vars["dt"] = Var("dt", dt); dt_stale = false; }
  unsigned flags = 0; check(kml_error_flags(ctx, &flags));
  if (flags) fatal("device error flags set: " + std::to_string(flags) + " (bit0 particle left the domain, bit1 J<=0, bit2 dtCFL invalid, bit3 polar decomposition failed)\n");
}

// Run / RunTime / RunUntil / RunWhile ::command, src/run.cpp:32-57, src/run_time.cpp:32-56, src/run_until.cpp, src/run_while.cpp
Var Sim::cmd_run(std::vector<std::string> &a, int kind) {
  if (a.size() < 1) fatal("Illegal run command.\n");
  if (!method_set || !ctx) fatal("Error: no method was defined!\n");
  if (restarted_TL) fatal("Bad domain decomposition, some CPUs (at least CPU #0) do not have any particles attached to.\nTry to increase or decrease the number of CPUs (using prime numbers might help).\n");
  check(kml_set_dt(ctx, dt));
  // Run::command calls scheme->setup() (-> Output::setup) BEFORE it updates laststep (src/run.cpp:41-53): the log interval of a
  // second run is therefore clipped by the PREVIOUS run's last step, which is why the reference repeats that step's row
  output_setup();
  maxtime = -1;
  firststep = ntimestep;
  Var cond;
  if (kind == 0) {
    int n = (int)(double)input.parsev(a[0]);
    nsteps = n; laststep = firststep + n;
    cond = Var("timestep<" + std::to_string(laststep), ntimestep < laststep);
  } else if (kind == 1) {
    double mt = (double)input.parsev(a[0]) + atime;
    nsteps = INT_MAX; maxtime = mt; laststep = INT_MAX;
    cond = Var("time<" + std::to_string(mt), atime < mt);
  } else {
    nsteps = INT_MAX; laststep = INT_MAX;
    Var c = input.parsev(a[0]);
    cond = kind == 2 ? Var("!(" + c.str() + ")", !c.result(), false) : Var(c.str(), c.result(), false);
  }
  run(cond);
  return Var(0);
}

} // namespace kmlh
