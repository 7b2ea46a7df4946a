/* kml_host.h - C entry points of the host driver (libkml_host.so): the Karamelo
 * input-script front end + scheme orchestration that drives libkml.so.
 * Replaces the reference's `karamelo -i file` process entry (reference src/main.cpp:21-32,
 * src/mpm.cpp:31-135, src/input.cpp:101-142) for embedding; the CLI `kml` wraps it. */
#ifndef KML_HOST_H
#define KML_HOST_H
#include "kml.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct kmlh_sim kmlh_sim;
const char *kmlh_last_error(void);
int kmlh_create(kmlh_sim **out);
int kmlh_destroy(kmlh_sim *s);
int kmlh_set_quiet(kmlh_sim *s, int quiet);
int kmlh_set_device(kmlh_sim *s, int device);
/* one process per GPU: this process is rank `rank` of `nranks` slabs; nccl_id128 = the 128-byte id from
 * kml_comm_unique_id() of rank 0 (NULL for host-only tests of the partition).  Call before the script. */
int kmlh_set_ranks(kmlh_sim *s, int rank, int nranks, const void *nccl_id128);
/* slab of solid i: base_lo, base_hi, goff, local nx, own_lo, own_hi, global particle count, tag offset */
int kmlh_slab_info(kmlh_sim *s, int i, int64_t info[8]);
int kmlh_run_file(kmlh_sim *s, const char *path);   /* Input::file */
int kmlh_run_line(kmlh_sim *s, const char *line);   /* one script line through Input::parsev */
int kmlh_get_var(kmlh_sim *s, const char *name, double *value);
int kmlh_nsolids(kmlh_sim *s, int *n);
/* np, device solid id, device grid id and node counts of solid i */
int kmlh_solid_info(kmlh_sim *s, int i, int64_t *np, int *solid_id, int *grid_id, int n[3]);
int kmlh_state(kmlh_sim *s, int64_t *ntimestep, double *time, double *dt);
kml_ctx *kmlh_ctx(kmlh_sim *s);
/* Script expressions: evaluate one (Input::parsev + Var::result), publish a per-particle variable (x, y, z, x0, y0, z0) the way the fixes do
 * before every evaluation, and compile an expression into the postfix program of include/kml.h (kml_expr). */
int kmlh_eval(kmlh_sim *s, const char *expr, double *value);
int kmlh_set_particle_var(kmlh_sim *s, const char *name, double value);
int kmlh_compile_expr(kmlh_sim *s, const char *expr, int *n, int *ops, double *vals);
/* Runs the fixes' initial_integrate hooks as step 1 would (initial_velocity_particles, initial_stress ...) without taking a step. */
int kmlh_apply_initial_fixes(kmlh_sim *s);
#ifdef __cplusplus
}
#endif
#endif
