// capi.cpp - C entry points of the host driver (include/kml_host.h).
#include "../../include/kml_host.h"
#include "sim.h"
#include <stdexcept>

using namespace kmlh;
struct kmlh_sim { Sim sim; };
static std::string g_herr;
#define GUARD(body) try { body; return 0; } catch (const std::exception &e) { g_herr = e.what(); return 1; }

extern "C" {
const char *kmlh_last_error(void) { return g_herr.c_str(); }
int kmlh_create(kmlh_sim **out) { GUARD(*out = new kmlh_sim()) }
int kmlh_destroy(kmlh_sim *s) { delete s; return 0; }
int kmlh_set_quiet(kmlh_sim *s, int q) { s->sim.quiet = q != 0; s->sim.input.echo = !q; return 0; }
int kmlh_set_device(kmlh_sim *s, int d) { s->sim.device = d; return 0; }
int kmlh_set_ranks(kmlh_sim *s, int rank, int nranks, const void *nccl_id128) {
  s->sim.rank = rank; s->sim.nranks = nranks;
  if (nccl_id128) s->sim.nccl_id.assign((const unsigned char *)nccl_id128, (const unsigned char *)nccl_id128 + 128);
  if (rank != 0) { s->sim.quiet = true; s->sim.input.echo = false; }
  return 0;
}
int kmlh_slab_info(kmlh_sim *s, int i, int64_t info[8]) {
  if (i < 0 || i >= (int)s->sim.solids.size()) { g_herr = "bad solid index"; return 1; }
  SolidH &S = *s->sim.solids[i]; const kml_grid_desc &d = S.grid->desc;
  info[0] = d.base_lo; info[1] = d.base_hi; info[2] = d.goff; info[3] = d.n[0]; info[4] = d.own_lo; info[5] = d.own_hi;
  info[6] = s->sim.np_global_last; info[7] = s->sim.tag_offset_last;
  return 0;
}
int kmlh_run_file(kmlh_sim *s, const char *path) { GUARD(s->sim.input.file(path)) }
int kmlh_run_line(kmlh_sim *s, const char *line) { GUARD(s->sim.input.line(line)) }
int kmlh_get_var(kmlh_sim *s, const char *name, double *value) {
  GUARD(auto it = s->sim.input.vars.find(name); if (it == s->sim.input.vars.end()) fatal(std::string("unknown variable ") + name); *value = it->second.result(&s->sim.input))
}
int kmlh_nsolids(kmlh_sim *s, int *n) { *n = (int)s->sim.solids.size(); return 0; }
int kmlh_solid_info(kmlh_sim *s, int i, int64_t *np, int *solid_id, int *grid_id, int n[3]) {
  if (i < 0 || i >= (int)s->sim.solids.size()) { g_herr = "bad solid index"; return 1; }
  SolidH &S = *s->sim.solids[i];
  *np = S.np; *solid_id = S.dev; *grid_id = S.grid->id;
  for (int d = 0; d < 3; d++) n[d] = S.grid->desc.n[d];
  return 0;
}
int kmlh_state(kmlh_sim *s, int64_t *nt, double *t, double *dt) { *nt = s->sim.ntimestep; *t = s->sim.atime; *dt = s->sim.dt; return 0; }
kml_ctx *kmlh_ctx(kmlh_sim *s) { return s->sim.ctx; }
int kmlh_eval(kmlh_sim *s, const char *expr, double *value) { GUARD(Var v = s->sim.input.parsev(expr); *value = v.result(&s->sim.input)) }
int kmlh_set_particle_var(kmlh_sim *s, const char *name, double value) { GUARD(s->sim.input.vars[name] = Var(name, value)) }
int kmlh_compile_expr(kmlh_sim *s, const char *expr, int *n, int *ops, double *vals) {
  GUARD(kml_expr e; Var v = s->sim.input.parsev(expr); if (!s->sim.input.compile(v, &e)) fatal("expression does not fit a device program");
        *n = e.n; for (int i = 0; i < e.n; i++) { ops[i] = e.op[i]; vals[i] = e.val[i]; })
}
int kmlh_apply_initial_fixes(kmlh_sim *s) { // the INITIAL_INTEGRATE hooks as the first step would run them (Scheme::run: ntimestep = 1)
  GUARD(const int64_t keep = s->sim.ntimestep; s->sim.ntimestep = 1; try { s->sim.hooks(INITIAL_INTEGRATE); } catch (...) { s->sim.ntimestep = keep; throw; } s->sim.ntimestep = keep)
}
}
