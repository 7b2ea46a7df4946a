// kml.cu - libkml.so: device state + the C ABI of include/kml.h on top of the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a (see karamelo_b200/Makefile).
#define KML_MISC_KERNELS
#include "kml_launch.h"
#include "kml_gather_cell3.cuh"
#include "kml_comm.cuh"
#include "kml_cpdi.cuh"
#include "kml_setup.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace kml;

static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return 1; }
enum { KML_RED_BITS = 64, KML_RED_N = KML_RED_BITS + 16 }; // d_red: [2 i] max wave speed, [2 i + 1] min_h_ratio of solid i; [KML_RED_BITS + b] bit b of the error word; [KML_RED_BITS + 8] a rank asks for a physical permute
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

struct kml_ctx;
extern "C" { static int resolve_dt(kml_ctx *c); }
struct Grid {
  kml_grid_desc d; GridDev g; double *buf = nullptr; int *ibuf = nullptr;
  bool v_is_momentum = false, T_is_weighted = false;
  // packed gather records {v_update, v_update - v} on a zero-padded grid for the TMA-fed G2P kernel (kml_gather_cell3.cuh); valid = they
  // reflect the current nv / nvu (written by k_grid_update, invalidated by everything else that touches node velocities)
  double *nvd = nullptr; bool nvd_valid = false;
};
struct Solid {
  kml_solid_desc d; SolidDev s; double *buf = nullptr; long long *lbuf = nullptr; int *ibuf = nullptr; long long cap = 0;
  CellLists cl;           // cell lists of the cell-centric kernels (UL, 3-D cubic splines): one set per solid, several solids share the grid
  bool moved = false;     // xn holds the positions after grid_to_points (UL)
  bool mbp_nonzero = false;
  bool rigid = false;     // Mat::rigid (src/material.h:49)
  CpdiDev cp{}; double *cpbuf = nullptr; int *cpibuf = nullptr; // CPDI neighbour lists and particle domains
  double *red = nullptr;  // device: [0] max wave speed, [1] min_h_ratio
  double dtCFL = 1.0e22;
  unsigned long long gen = 1; // kml_solid_generation
  double *accbuf = nullptr;   // kml_keep_particle_acceleration: a_p, v_update_p
  // physical re-ordering: second set of buffers, swapped with the current one by permute_solid
  int nd = 0; double *buf2 = nullptr; long long *lbuf2 = nullptr; int *ibuf2 = nullptr; bool no_permute = false;
  long long permutes = 0, last_permute_step = -1000000;
  // Cost-based policy (amortised rebuild): the stress kernel - the stage that streams the most state - is timed with an event pair every
  // step; t_clean is its time right after a permute, excess the time lost to disorder since then, permute_ms the measured cost of the
  // last permute.  A permute is due when excess >= permute_ms (for a linearly growing loss that is the optimal period).
  cudaEvent_t ev_s[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}; int ev_s_next = 0; bool ev_s_valid[2] = {false, false}; long long ev_s_step[2] = {0, 0}, clean_step = -1;
  cudaEvent_t ev_p[2] = {nullptr, nullptr}; bool ev_p_valid = false;
  double t_clean = -1, excess_ms = 0, permute_ms = -1;
};

struct kml_ctx {
  kml_config c; int dev = 0; cudaStream_t stream = nullptr;
  std::vector<Grid *> grids; std::vector<Solid *> solids;
  double dt = 1e-16;
  unsigned *d_flags = nullptr; double *d_scratch = nullptr; // scratch: small reduction outputs
  double *h_pinned = nullptr;                                // pinned readback buffer
  void *d_stage = nullptr; size_t stage_bytes = 0;           // upload / download staging (rows <-> SoA)
  bool tl_mass_done = false, tl_wf_done = false;
  bool has_rigid = false; // some solid is rigid (ULMPM::rigid_solids, src/ulmpm.cpp:98-99)
  bool apic = false; // affine transfer: TL: APIC; UL: APIC, MLS, AFLIP, ASFLIP (src/ulmpm.cpp:79-85, src/tlmpm.cpp:83-85)
  bool keep_acc = false; long long steps_started = 0; // kml_keep_particle_acceleration
  double *d_red = nullptr, *h_red = nullptr; cudaEvent_t ev_dt = nullptr; bool dt_pending = false; double dt_factor = 1.0; // deferred adjust_dt (resolve_dt)
  bool pending_g2p = false; int pending_grad = -1; bool pending_F = false; bool grad_moved = false;
  struct ContactScratch { int *start = nullptr, *cell_of = nullptr, *rank = nullptr, *order = nullptr; void *scan_tmp = nullptr; size_t scan_bytes = 0; long long nbins_cap = 0, np_cap = 0; } contact; // bins of the cell-list contact kernel
  bool permute_want = false, permute_go = false, dt_collective = false; long long last_collective_permute = -10; // decomposed runs: a rank's wish travels with the dt all-reduce and all ranks re-order in the same step (no rank waits for another's permute)
  double permute_frac = 0.05; int permute_min_steps = 2, permute_every = 0; // KML_PERMUTE_FRAC (negative: never), KML_PERMUTE_MIN_STEPS
  int g2p_tma = 4; int nsm = 148; // KML_G2P_TMA: 0 = tile through registers, 1 / 3 = persistent TMA-fed kernel, 4 = one block per segment with a bulk-copied tile (default)
  bool use_cell_p2g = true; int cell_mask = 7; int p2g_nb = 1, v2g_nb = 2; GatherTune gtune; // cell_mask (KML_CELL_MASK): 1 = P2G, 2 = G2P, 4 = stress use the cell kernels // measurement switches: KML_P2G=atomic, KML_P2G_NB, KML_V2G_NB, KML_SEGLEN, KML_GATHER_THREADS
  Comm comm;
  // profiling
  // Per-stage device time: event pairs are recorded around every stage and only READ in kml_stage_times (one synchronisation for the
  // whole timed region), so profiling does not serialise the host against the device and the stage sum stays inside the step time.
  bool profile = false; cudaEvent_t evA = nullptr, evB = nullptr; double ms[KML_STAGE_COUNT]; long long launches[KML_STAGE_COUNT]; double host_ms[KML_STAGE_COUNT] = {0};
  struct Pair { int stage; cudaEvent_t a, b; };
  std::vector<Pair> ev_pending; std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t ev_get() { if (ev_pool.empty()) { cudaEvent_t e; cudaEventCreate(&e); return e; } cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
};

namespace {
struct StageTimer {
  kml_ctx *c; int stage; cudaEvent_t a = nullptr; std::chrono::steady_clock::time_point h0;
  StageTimer(kml_ctx *c_, int st) : c(c_), stage(st) { if (c->profile) { a = c->ev_get(); cudaEventRecord(a, c->stream); h0 = std::chrono::steady_clock::now(); } }
  void stop() {
    if (a) {
      cudaEvent_t b = c->ev_get(); cudaEventRecord(b, c->stream); c->ev_pending.push_back({stage, a, b}); a = nullptr;
      c->host_ms[stage] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count(); // host time spent inside the call (launches, waits)
    }
  }
  ~StageTimer() { stop(); }
};
StepParams step_params(kml_ctx *c) {
  StepParams sp; sp.dt = c->dt; sp.alpha = c->c.PIC_FLIP;
  for (int d = 0; d < 3; d++) { sp.boxlo[d] = c->c.boxlo[d]; sp.boxhi[d] = c->c.boxhi[d]; }
  sp.axisymmetric = c->c.axisymmetric; sp.temp = c->c.temp; sp.inv_tav = 0.0; sp.flags = c->d_flags;
  sp.apic = c->apic; sp.mls = !c->c.is_TL && c->c.sub_method == KML_SUB_MLS; sp.asflip = !c->c.is_TL && c->c.sub_method == KML_SUB_ASFLIP;
  sp.Di[0] = sp.Di[1] = sp.Di[2] = 1.0; sp.ext = c->c.ge ? 1 : 0;
  return sp;
}

// Solid::compute_inertia_tensor, src/solid.cpp:1440-1478: Di = k / cellsize^2 on the active dimensions (linear TL: 16/4 with one
// particle per cell, 16/3 with two; other lattices are refused by kml_solid_create like the reference does)
void fill_inertia(kml_ctx *c, const Grid *G, const Solid *S, StepParams &sp) {
  if (!c->apic) return;
  const double cs = 1.0 / (G->d.cellsize * G->d.cellsize);
  const double klin = S->d.np_per_cell == 1 ? 16.0 / 4.0 : 16.0 / 3.0;
  const double k = c->c.shape_function == KML_SHAPE_LINEAR ? klin : (c->c.shape_function == KML_SHAPE_CUBIC_SPLINE ? 3.0 : 4.0);
  for (int d = 0; d < 3; d++) sp.Di[d] = d < c->c.dimension ? k * cs : 1.0;
}

// kernel dispatch on (dimension, TL); the shape function is switched inside each launcher (kml_launch.h)
#define KML_DISPATCH(FAMILY, ...)                                                                                          \
  do {                                                                                                                     \
    const int sh_ = c->c.shape_function;                                                                                   \
    if (c->c.is_TL) {                                                                                                      \
      if (c->c.dimension == 1) launch_##FAMILY##_d1_tl1(sh_, __VA_ARGS__);                                                 \
      else if (c->c.dimension == 2) launch_##FAMILY##_d2_tl1(sh_, __VA_ARGS__);                                            \
      else launch_##FAMILY##_d3_tl1(sh_, __VA_ARGS__);                                                                     \
    } else {                                                                                                               \
      if (c->c.dimension == 1) launch_##FAMILY##_d1_tl0(sh_, __VA_ARGS__);                                                 \
      else if (c->c.dimension == 2) launch_##FAMILY##_d2_tl0(sh_, __VA_ARGS__);                                            \
      else launch_##FAMILY##_d3_tl0(sh_, __VA_ARGS__);                                                                     \
    }                                                                                                                      \
  } while (0)

struct RowMap { double *comp[9]; int col[9]; int ncols, ncomp; };
__global__ void k_rows_to_soa(const double *rows, RowMap rm, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < rm.ncomp; k++) rm.comp[k][i] = rows[i * rm.ncols + rm.col[k]];
}
__global__ void k_soa_to_rows(double *rows, RowMap rm, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < rm.ncomp; k++) rows[i * rm.ncols + rm.col[k]] = rm.comp[k][i];
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}
int stage_reserve(kml_ctx *c, size_t bytes) {
  if (bytes <= c->stage_bytes) return 0;
  if (c->d_stage) cudaFree(c->d_stage);
  c->d_stage = nullptr; c->stage_bytes = 0;
  cudaError_t e = cudaMalloc(&c->d_stage, bytes);
  if (e != cudaSuccess) return fail(std::string("staging buffer: ") + cudaGetErrorString(e));
  c->stage_bytes = bytes; return 0;
}
} // namespace

namespace {
// the source index of every surviving slot after the reference's swap-with-last compaction (src/delete_particles.cpp:56-66)
std::vector<int> delete_order(const int *dlist, long long np) {
  std::vector<int> src(np), dl(dlist, dlist + np);
  for (long long i = 0; i < np; i++) src[i] = (int)i;
  long long n = np, k = 0;
  while (k < n) { if (dl[k]) { src[k] = src[n - 1]; dl[k] = dl[n - 1]; n--; } else k++; }
  src.resize(n); return src;
}
template <class T> __global__ void k_gather_rows(T *dst, const T *src, const int *idx, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
} // namespace

extern "C" {

const char *kml_last_error(void) { return g_err.c_str(); }
const char *kml_backend(void) { return "cuda-sm_100a"; }

int kml_create(const kml_config *cfg, kml_ctx **out) {
  if (cfg->is_CPDI && cfg->dimension != 2) return fail("Error: ULCPDI is only 2D....\n"); // src/ulcpdi.cpp:115-118, src/tlcpdi.cpp:102-104
  if (cfg->is_CPDI && (cfg->axisymmetric || cfg->temp)) return fail("kml: CPDI with axisymmetry / thermo-mechanical coupling is not implemented in the CUDA engine");
  if (cfg->ge && (cfg->is_TL || cfg->is_CPDI)) return fail("kml: gradient-enhanced projection is implemented for ulmpm only in the CUDA engine");
  const bool apic_ = cfg->is_TL ? cfg->sub_method == KML_SUB_APIC
                                : (cfg->sub_method == KML_SUB_APIC || cfg->sub_method == KML_SUB_MLS || cfg->sub_method == KML_SUB_AFLIP || cfg->sub_method == KML_SUB_ASFLIP);
  if (cfg->is_TL && cfg->sub_method != KML_SUB_PIC && cfg->sub_method != KML_SUB_FLIP && cfg->sub_method != KML_SUB_APIC)
    return fail("kml: total-Lagrangian methods take PIC, FLIP or APIC");
  if (apic_) { // Solid::compute_inertia_tensor, src/solid.cpp:1440-1478
    if (cfg->is_CPDI) return fail("kml: APIC with CPDI is not implemented in the CUDA engine");
    if (cfg->shape_function == KML_SHAPE_BERNSTEIN) return fail("Shape function not supported for APIC.");
    if (cfg->shape_function == KML_SHAPE_LINEAR && !cfg->is_TL) return fail("Shape function not supported for APIC and ULMPM.");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("kml: no CUDA device available - the engine has no CPU fallback");
  kml_ctx *c = new kml_ctx(); c->c = *cfg; c->dev = cfg->device; c->apic = apic_;
  CU(cudaSetDevice(c->dev));
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaMalloc(&c->d_flags, sizeof(unsigned))); CU(cudaMemset(c->d_flags, 0, sizeof(unsigned)));
  CU(cudaMalloc(&c->d_scratch, 64 * sizeof(double)));
  CU(cudaMallocHost(&c->h_pinned, 64 * sizeof(double)));
  CU(cudaMalloc(&c->d_red, KML_RED_N * sizeof(double))); CU(cudaMemset(c->d_red, 0, KML_RED_N * sizeof(double)));
  CU(cudaMallocHost(&c->h_red, KML_RED_N * sizeof(double))); CU(cudaEventCreateWithFlags(&c->ev_dt, cudaEventDisableTiming));
  CU(cudaEventCreate(&c->evA)); CU(cudaEventCreate(&c->evB));
  memset(c->ms, 0, sizeof c->ms); memset(c->launches, 0, sizeof c->launches);
  const char *e = getenv("KML_P2G");
  if (e && !strcmp(e, "atomic")) c->use_cell_p2g = false;
  auto env_int = [](const char *name, int dflt) { const char *v = getenv(name); return v && *v ? atoi(v) : dflt; };
  c->p2g_nb = env_int("KML_P2G_NB", 1) == 2 ? 2 : 1;
  c->cell_mask = env_int("KML_CELL_MASK", 7);
  { const char *v = getenv("KML_PERMUTE_FRAC"); if (v && *v) c->permute_frac = atof(v); }
  c->permute_min_steps = env_int("KML_PERMUTE_MIN_STEPS", 2);
  c->permute_every = env_int("KML_PERMUTE_EVERY", 0);
  c->g2p_tma = env_int("KML_G2P_TMA", 4);
  { cudaDeviceProp pr; if (cudaGetDeviceProperties(&pr, c->dev) == cudaSuccess) c->nsm = pr.multiProcessorCount; }
  { const int v = env_int("KML_V2G_NB", 2); c->v2g_nb = (v == 1 || v == 4) ? v : 2; }
  // cells per column segment, per kernel family (measured at 100 M particles: the stress kernel wants shorter segments than the other three)
  auto seg_env = [&](const char *name, int dflt) { return std::min(std::max(env_int(name, env_int("KML_SEGLEN", dflt)), 8), 96); };
  c->gtune.seg_target = seg_env("KML_SEGLEN_P2G", 64); // 64 (and 96) against 32 at 100 M particles: P2G 11.64 vs 12.01 ms, re-projection 4.91 vs 5.04 (gpurun_out/ab_r2w.log)
  c->gtune.seg_g2p = seg_env("KML_SEGLEN_G2P", 32);
  c->gtune.seg_stress = seg_env("KML_SEGLEN_STRESS", 24);
  c->gtune.threads = env_int("KML_GATHER_THREADS", 64) == 128 ? 128 : 64;
  c->gtune.g2p_threads = env_int("KML_G2P_THREADS", c->gtune.threads) == 128 ? 128 : 64;
  *out = c; return 0;
}

int kml_destroy(kml_ctx *c) {
  if (!c) return 0;
  cudaSetDevice(c->dev); cudaStreamSynchronize(c->stream);
  for (auto g : c->grids) { cudaFree(g->buf); cudaFree(g->ibuf); cudaFree(g->nvd); delete g; }
  for (auto s : c->solids) { cudaFree(s->cpbuf); cudaFree(s->cpibuf); cudaFree(s->buf); cudaFree(s->lbuf); cudaFree(s->ibuf); cudaFree(s->accbuf); cudaFree(s->buf2); cudaFree(s->lbuf2); cudaFree(s->ibuf2); s->cl.release();
    for (int a = 0; a < 2; a++) { for (int b = 0; b < 2; b++) if (s->ev_s[a][b]) cudaEventDestroy(s->ev_s[a][b]); if (s->ev_p[a]) cudaEventDestroy(s->ev_p[a]); }
    delete s; }
  if (c->comm.comm) {
    nccl().CommDestroy(c->comm.comm);
    if (c->comm.zone_left) cudaIpcCloseMemHandle(c->comm.zone_left); if (c->comm.zone_right) cudaIpcCloseMemHandle(c->comm.zone_right); cudaFree(c->comm.zone); cudaFree(c->comm.push_done);
    cudaFree(c->comm.halo_buf); cudaFree(c->comm.mig_cnt); cudaFree(c->comm.mig_list); cudaFree(c->comm.mig_flag);
    cudaFree(c->comm.mig_send); cudaFree(c->comm.mig_recv); cudaFreeHost(c->comm.h_cnt);
  }
  cudaFree(c->contact.start); cudaFree(c->contact.cell_of); cudaFree(c->contact.rank); cudaFree(c->contact.order); cudaFree(c->contact.scan_tmp);
  cudaFree(c->d_flags); cudaFree(c->d_scratch); cudaFreeHost(c->h_pinned); cudaFree(c->d_stage); cudaFree(c->d_red); cudaFreeHost(c->h_red); if (c->ev_dt) cudaEventDestroy(c->ev_dt);
  for (auto &p : c->ev_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  cudaEventDestroy(c->evA); cudaEventDestroy(c->evB); cudaStreamDestroy(c->stream);
  delete c; return 0;
}
int kml_synchronize(kml_ctx *c) { CU(cudaSetDevice(c->dev)); CU(cudaStreamSynchronize(c->stream)); return 0; }
int kml_set_domain_box(kml_ctx *c, const double lo[3], const double hi[3]) { for (int d = 0; d < 3; d++) { c->c.boxlo[d] = lo[d]; c->c.boxhi[d] = hi[d]; } return 0; }

// ---- grids ------------------------------------------------------------------------------------
static const int GRID_NDBL = 4 + 4 + 3 + 3 + 3 + 3; // nv(4) nvu(4) f mb T Qext Qint x

int kml_grid_create(kml_ctx *c, const kml_grid_desc *d, int *gid) {
  CU(cudaSetDevice(c->dev));
  Grid *G = new Grid(); G->d = *d; GridDev &g = G->g;
  struct Guard { Grid *G; bool keep = false; ~Guard() { if (!keep) { cudaFree(G->buf); cudaFree(G->ibuf); delete G; } } } guard{G}; // a failed CU(...) below returns early
  for (int k = 0; k < 3; k++) { g.lo[k] = d->lo[k]; g.n[k] = d->n[k]; }
  g.h = d->h; g.cellsize = d->cellsize; g.inv_cellsize = 1.0 / d->cellsize;
  g.goff0 = d->goff; g.gn0 = d->gn > 0 ? d->gn : d->n[0];
  g.own_lo = d->gn > 0 ? d->own_lo : 0; g.own_hi = d->gn > 0 ? d->own_hi : d->n[0];
  g.nn = (long long)d->n[0] * d->n[1] * d->n[2];
  const long long nn = g.nn, stride = (nn + 31) / 32 * 32;
  CU(cudaMalloc(&G->buf, sizeof(double) * stride * GRID_NDBL)); CU(cudaMemsetAsync(G->buf, 0, sizeof(double) * stride * GRID_NDBL, c->stream));
  CU(cudaMalloc(&G->ibuf, sizeof(int) * stride * 2));
  double *p = G->buf; auto take = [&](int n = 1) { double *r = p; p += stride * n; return r; };
  g.nv = (double4 *)take(4); g.nvu = (double4 *)take(4);
  for (int k = 0; k < 3; k++) g.f[k] = take(); for (int k = 0; k < 3; k++) g.mb[k] = take();
  g.T = take(); g.Qext = take(); g.Qint = take(); for (int k = 0; k < 3; k++) g.x[k] = take();
  g.mask = G->ibuf; g.rigid = G->ibuf + stride;
  std::vector<int> ones(nn, 1);
  CU(cudaMemcpyAsync(g.mask, ones.data(), sizeof(int) * nn, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(g.rigid, 0, sizeof(int) * nn, c->stream));
  // node positions x = x0 (only TL moves them)
  std::vector<double> xs(nn);
  for (int k = 0; k < 3; k++) {
    long long l = 0;
    for (int i = 0; i < d->n[0]; i++) for (int j = 0; j < d->n[1]; j++) for (int kk = 0; kk < d->n[2]; kk++, l++) {
      int idx = k == 0 ? i + g.goff0 : (k == 1 ? j : kk);
      xs[l] = (k < c->c.dimension) ? d->lo[k] + idx * d->h : 0.0;
    }
    CU(cudaMemcpyAsync(g.x[k], xs.data(), sizeof(double) * nn, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  guard.keep = true;
  c->grids.push_back(G); *gid = (int)c->grids.size() - 1; return 0;
}
int kml_grid_nnodes(kml_ctx *c, int gid, int64_t *nn) { *nn = c->grids[gid]->g.nn; return 0; }

static int grid_normalize_if_needed(kml_ctx *c, Grid *G) {
  if (!G->v_is_momentum && !G->T_is_weighted) return 0;
  k_grid_update<<<nblocks(G->g.nn, 256), 256, 0, c->stream>>>(G->g, 0.0 /* dt is not used without the update */, G->v_is_momentum, 0, c->c.temp, G->T_is_weighted, 0, nullptr);
  G->v_is_momentum = false; G->T_is_weighted = false; G->nvd_valid = false; c->launches[KML_STAGE_GRID]++;
  return check_launch("k_grid_update(normalize)");
}

// component pointers + element stride (in doubles) of a node field
static int grid_field(kml_ctx *c, Grid *G, int field, double **comp, int *ncomp, int *stride, int **icomp) {
  GridDev &g = G->g; *icomp = nullptr; *ncomp = 1; *stride = 1;
  double *nv = (double *)g.nv, *nvu = (double *)g.nvu;
  switch (field) {
  case KML_N_X: for (int k = 0; k < 3; k++) comp[k] = g.x[k]; *ncomp = 3; break;
  case KML_N_V: for (int k = 0; k < 3; k++) comp[k] = nv + k; *ncomp = 3; *stride = 4; break;
  case KML_N_V_UPDATE: for (int k = 0; k < 3; k++) comp[k] = nvu + k; *ncomp = 3; *stride = 4; break;
  case KML_N_MB: for (int k = 0; k < 3; k++) comp[k] = g.mb[k]; *ncomp = 3; break;
  case KML_N_F: for (int k = 0; k < 3; k++) comp[k] = g.f[k]; *ncomp = 3; break;
  case KML_N_MASS: comp[0] = nv + 3; *stride = 4; break;
  case KML_N_T: comp[0] = g.T; break;
  case KML_N_T_UPDATE: comp[0] = nvu + 3; *stride = 4; break;
  case KML_N_QEXT: comp[0] = g.Qext; break; case KML_N_QINT: comp[0] = g.Qint; break;
  case KML_N_MASK: *icomp = g.mask; break; case KML_N_RIGID: *icomp = g.rigid; break;
  default: return fail("grid field not supported");
  }
  return 0;
}

int kml_grid_upload(kml_ctx *c, int gid, int field, const void *src) {
  CU(cudaSetDevice(c->dev));
  Grid *G = c->grids[gid]; const long long nn = G->g.nn;
  double *comp[3]; int nc, stride; int *ic;
  if (grid_field(c, G, field, comp, &nc, &stride, &ic)) return 1;
  if (ic) { CU(cudaMemcpyAsync(ic, src, sizeof(int) * nn, cudaMemcpyHostToDevice, c->stream)); CU(cudaStreamSynchronize(c->stream)); return 0; }
  if (field == KML_N_V) G->v_is_momentum = false;
  if (field == KML_N_T) G->T_is_weighted = false;
  G->nvd_valid = false;
  const double *s = (const double *)src;
  for (int k = 0; k < nc; k++) { // strided 2-D copy: host column k of [nn][nc] -> device component
    CU(cudaMemcpy2DAsync(comp[k], sizeof(double) * stride, s + k, sizeof(double) * nc, sizeof(double), nn, cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
int kml_grid_download(kml_ctx *c, int gid, int field, void *dst) {
  CU(cudaSetDevice(c->dev));
  Grid *G = c->grids[gid]; const long long nn = G->g.nn; const kml_grid_desc &d = G->d;
  if (field == KML_N_X0 || field == KML_N_NTYPE) {
    long long l = 0;
    for (int i = 0; i < d.n[0]; i++) for (int j = 0; j < d.n[1]; j++) for (int k = 0; k < d.n[2]; k++, l++) {
      int idx[3] = {i + G->g.goff0, j, k};
      for (int a = 0; a < 3; a++) {
        if (field == KML_N_X0) ((double *)dst)[3 * l + a] = a < c->c.dimension ? d.lo[a] + idx[a] * d.h : 0.0;
        else {
          int nt = 0, n = a == 0 ? G->g.gn0 : d.n[a], ii = idx[a];
          if (c->c.shape_function == KML_SHAPE_BERNSTEIN) nt = ii % 2;
          else if (c->c.shape_function != KML_SHAPE_LINEAR) nt = std::min(2, ii) - std::min(n - 1 - ii, 2);
          ((int *)dst)[3 * l + a] = nt;
        }
      }
    }
    return 0;
  }
  if (field == KML_N_V || field == KML_N_T) if (grid_normalize_if_needed(c, G)) return 1;
  double *comp[3]; int nc, stride; int *ic;
  if (grid_field(c, G, field, comp, &nc, &stride, &ic)) return 1;
  if (ic) { CU(cudaMemcpyAsync(dst, ic, sizeof(int) * nn, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream)); return 0; }
  double *o = (double *)dst;
  for (int k = 0; k < nc; k++)
    CU(cudaMemcpy2DAsync(o + k, sizeof(double) * nc, comp[k], sizeof(double) * stride, sizeof(double), nn, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- solids -----------------------------------------------------------------------------------
static const int SOLID_NDBL_UL = 3 * 6 + 6 + 6 + 9 + 11;
static const int SOLID_NDBL_TL = SOLID_NDBL_UL + 18;

static int alloc_acc(kml_ctx *c, Solid *S);
// SoA component pointers of the solid's current buffers (the physical permute swaps buffers)
static void bind_pointers(kml_ctx *c, Solid *S) {
  SolidDev &s = S->s; const long long cap = S->cap;
  double *p = S->buf; auto take = [&]() { double *r = p; p += cap; return r; };
  for (int k = 0; k < 3; k++) s.x[k] = take(); for (int k = 0; k < 3; k++) s.xn[k] = take(); for (int k = 0; k < 3; k++) s.x0[k] = take();
  for (int k = 0; k < 3; k++) s.v[k] = take(); for (int k = 0; k < 3; k++) s.mbp[k] = take(); for (int k = 0; k < 3; k++) s.q[k] = take();
  for (int k = 0; k < 6; k++) s.sig[k] = take(); for (int k = 0; k < 6; k++) s.eel[k] = take(); for (int k = 0; k < 9; k++) s.F[k] = take();
  s.vol0 = take(); s.vol = take(); s.rho0 = take(); s.mass = take(); s.eps = take(); s.epsdot = take(); s.dmg = take(); s.dmgi = take();
  s.ien = take(); s.T = take(); s.gamma = take();
  if (c->c.is_TL) { for (int k = 0; k < 9; k++) s.pk1[k] = take(); for (int k = 0; k < 9; k++) s.R[k] = take(); }
  else { for (int k = 0; k < 9; k++) { s.pk1[k] = nullptr; s.R[k] = nullptr; } }
  for (int k = 0; k < 9; k++) s.Lst[k] = (c->apic || c->c.ge) ? take() : nullptr;
  s.ptag = S->lbuf; s.mask = S->ibuf;
}

// Physical re-ordering (SURVEY section 8d S0'): particle state gathered into cell order through the re-bin's order[], into the solid's second
// buffer, which then becomes current.  The cell-sorted kernels read 31-55 SoA streams per particle through order[]; while order[] is close to
// the identity those streams coalesce, but mixing (and, on a decomposed run, migration: arrivals are appended, holes are filled from the tail)
// turns them into sector-granular gathers.  xn (scratch between grid_to_points and the next weight evaluation) is not copied.
// xslot: the buffer slot (0 or 3) that holds the CURRENT positions - x and xn trade places after every step (kml_compute_grid_weight_...);
// they land in slot 0 of the destination, whose pointers are bound in the canonical order
__global__ void k_permute(const double *__restrict__ src, double *__restrict__ dst, long long cap, int nd, int xslot, const int *__restrict__ order, long long np,
                          const long long *__restrict__ tsrc, long long *__restrict__ tdst, const int *__restrict__ msrc, int *__restrict__ mdst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const long long j = order[i];
  for (int a = 0; a < nd; a++) { if (a >= 3 && a < 6) continue; const int sa = a < 3 ? xslot + a : a; dst[a * cap + i] = src[sa * cap + j]; }
  tdst[i] = tsrc[j]; mdst[i] = msrc[j];
}
__global__ void k_iota(int *a, long long n) { const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = (int)i; }

// the second buffer of the double-buffered permute, allocated at the FIRST re-bin (a 26-52 GB cudaMalloc takes tens of milliseconds once peer
// access is enabled; it must not land in the middle of a run)
static int permute_reserve(kml_ctx *c, Solid *S) {
  if (S->no_permute || S->accbuf || S->cpbuf) return 0;
  if (!S->buf2) {
    const size_t bytes = sizeof(double) * S->cap * S->nd;
    if (cudaMalloc(&S->buf2, bytes) != cudaSuccess || cudaMalloc(&S->lbuf2, sizeof(long long) * S->cap) != cudaSuccess || cudaMalloc(&S->ibuf2, sizeof(int) * S->cap) != cudaSuccess) {
      cudaGetLastError(); cudaFree(S->buf2); cudaFree(S->lbuf2); cudaFree(S->ibuf2); S->buf2 = nullptr; S->lbuf2 = nullptr; S->ibuf2 = nullptr;
      S->no_permute = true; return 0; // not enough memory for the second buffer: keep the index-only order
    }
    CU(cudaMemsetAsync(S->buf2, 0, bytes, c->stream));
  }
  return 0;
}
static int permute_solid(kml_ctx *c, Solid *S, Grid *G) {
  if (permute_reserve(c, S)) return 1;
  if (S->no_permute || S->accbuf || S->cpbuf || !S->buf2) return 0;
  const long long np = S->s.np;
  const int xslot = (int)((S->s.x[0] - S->buf) / S->cap); // 0, or 3 after an odd number of x <-> xn swaps
  if (xslot != 0 && xslot != 3) return fail("permute_solid: unexpected position slot");
  if (!S->ev_p[0]) { CU(cudaEventCreate(&S->ev_p[0])); CU(cudaEventCreate(&S->ev_p[1])); }
  CU(cudaEventRecord(S->ev_p[0], c->stream));
  k_permute<<<nblocks(np, 128), 128, 0, c->stream>>>(S->buf, S->buf2, S->cap, S->nd, xslot, S->cl.order, np, S->lbuf, S->lbuf2, S->ibuf, S->ibuf2);
  k_iota<<<nblocks(np, 256), 256, 0, c->stream>>>(S->cl.order, np);
  CU(cudaEventRecord(S->ev_p[1], c->stream)); S->ev_p_valid = true;
  if (check_launch("k_permute")) return 1;
  std::swap(S->buf, S->buf2); std::swap(S->lbuf, S->lbuf2); std::swap(S->ibuf, S->ibuf2);
  bind_pointers(c, S); S->gen++; S->permutes++;
  c->launches[KML_STAGE_REBIN] += 2;
  return 0;
}

int kml_solid_create(kml_ctx *c, const kml_solid_desc *d, int *sid) {
  CU(cudaSetDevice(c->dev));
  const bool rigid_ = d->mat.rigid || d->mat.type == KML_MAT_RIGID;
  if (rigid_ && c->c.is_CPDI) return fail("kml: rigid solids with CPDI are not implemented in the CUDA engine");
  if (c->apic && c->c.shape_function == KML_SHAPE_LINEAR && d->np_per_cell != 0 && d->np_per_cell != 1 && d->np_per_cell != 2)
    return fail("Number of particle per cell not supported with linear shape functions and APIC."); // src/solid.cpp:1453-1460
  Solid *S = new Solid(); S->d = *d; S->rigid = rigid_; if (rigid_) c->has_rigid = true; SolidDev &s = S->s;
  struct Guard { Solid *S; bool keep = false; ~Guard() { if (!keep) { cudaFree(S->buf); cudaFree(S->lbuf); cudaFree(S->ibuf); cudaFree(S->cpbuf); cudaFree(S->cpibuf); delete S; } } } guard{S}; // a failed CU(...) below returns early
  for (int k = 0; k < 3; k++) { s.acc[k] = nullptr; s.vup[k] = nullptr; }
  s.np = d->np; S->cap = std::max<long long>(d->capacity, d->np);
  long long cap = (S->cap + 31) / 32 * 32;
  { const char *pad = getenv("KML_CAP_PAD"); if (pad && *pad) cap += atoll(pad) / 32 * 32; } // experiment: de-alias the SoA component stride
  S->cap = cap;
  const int nd = (c->c.is_TL ? SOLID_NDBL_TL : SOLID_NDBL_UL) + ((c->apic || c->c.ge) ? 9 : 0);
  CU(cudaMalloc(&S->buf, sizeof(double) * cap * nd)); CU(cudaMemsetAsync(S->buf, 0, sizeof(double) * cap * nd, c->stream));
  CU(cudaMalloc(&S->lbuf, sizeof(long long) * cap)); CU(cudaMemsetAsync(S->lbuf, 0, sizeof(long long) * cap, c->stream));
  CU(cudaMalloc(&S->ibuf, sizeof(int) * cap));
  if (2 * (c->solids.size() + 1) > (size_t)KML_RED_BITS) return fail("too many solids");
  S->red = c->d_red + 2 * c->solids.size(); // slots of the context's reduction buffer (one all-reduce for all solids)
  S->nd = nd; bind_pointers(c, S);
  // initial values of Solid::populate, src/solid.cpp:2283-2321: F = R = I, rho0 = mat.rho0, mask = 1
  if (d->np > 0) { k_solid_init<<<nblocks(d->np, 256), 256, 0, c->stream>>>(s, d->mat.rho0); if (check_launch("k_solid_init")) return 1; }
  if (c->c.is_CPDI) {
    CpdiDev &cp = S->cp; cp.style = c->c.cpdi_style; cp.cap = cap; cp.maxn = c->c.shape_function == KML_SHAPE_LINEAR ? 16 : CPDI_MAXN;
    const size_t per = (size_t)cp.maxn * cap;
    CU(cudaMalloc(&S->cpibuf, sizeof(int) * (cap + per))); CU(cudaMemsetAsync(S->cpibuf, 0, sizeof(int) * (cap + per), c->stream));
    CU(cudaMalloc(&S->cpbuf, sizeof(double) * (7 * per + 24 * cap))); CU(cudaMemsetAsync(S->cpbuf, 0, sizeof(double) * (7 * per + 24 * cap), c->stream));
    cp.n = S->cpibuf; cp.node = S->cpibuf + cap;
    double *q = S->cpbuf; auto takep = [&](size_t n) { double *r = q; q += n; return r; };
    cp.wf = takep(per); cp.wfd[0] = takep(per); cp.wfd[1] = takep(per);
    for (int k = 0; k < 4; k++) cp.wfc[k] = takep(per);
    for (int k = 0; k < 2; k++) for (int d2 = 0; d2 < 2; d2++) { cp.rp[k][d2] = takep(cap); cp.rp0[k][d2] = takep(cap); }
    for (int k = 0; k < 4; k++) for (int d2 = 0; d2 < 2; d2++) { cp.xpc[k][d2] = takep(cap); cp.xpc0[k][d2] = takep(cap); }
  }
  if (c->keep_acc && alloc_acc(c, S)) return 1;
  CU(cudaStreamSynchronize(c->stream));
  guard.keep = true;
  c->solids.push_back(S); *sid = (int)c->solids.size() - 1; return 0;
}
int kml_solid_np(kml_ctx *c, int sid, int64_t *np) { *np = c->solids[sid]->s.np; return 0; }
int kml_solid_generation(kml_ctx *c, int sid, uint64_t *gen) { *gen = c->solids[sid]->gen; return 0; }

static int alloc_acc(kml_ctx *c, Solid *S) {
  if (S->accbuf) return 0;
  CU(cudaMalloc(&S->accbuf, sizeof(double) * 6 * S->cap)); CU(cudaMemsetAsync(S->accbuf, 0, sizeof(double) * 6 * S->cap, c->stream));
  for (int k = 0; k < 3; k++) { S->s.acc[k] = S->accbuf + (size_t)k * S->cap; S->s.vup[k] = S->accbuf + (size_t)(3 + k) * S->cap; }
  return 0;
}
int kml_keep_particle_acceleration(kml_ctx *c) {
  CU(cudaSetDevice(c->dev));
  if (c->keep_acc) return 0;
  if (c->steps_started > 0) return fail("kml: particle accelerations / forces (internal_force, external_force, dump fields) must be requested before the first step");
  if (c->c.nranks > 1) return fail("kml: stored particle accelerations are single-GPU in the CUDA engine (they do not migrate between slabs)");
  c->keep_acc = true;
  for (Solid *S : c->solids) if (alloc_acc(c, S)) return 1;
  return 0;
}

// component pointers of a particle field; sym = stored as 6 symmetric components
static int solid_field(kml_ctx *c, Solid *S, int field, double **comp, int *ncomp, bool *sym) {
  SolidDev &s = S->s; *sym = false; *ncomp = 1;
  auto v3 = [&](double *const *a) { for (int k = 0; k < 3; k++) comp[k] = a[k]; *ncomp = 3; };
  auto m9 = [&](double *const *a) { for (int k = 0; k < 9; k++) comp[k] = a[k]; *ncomp = 9; };
  switch (field) {
  case KML_P_X: v3((!c->c.is_TL && S->moved) ? s.xn : s.x); break;
  case KML_P_X0: v3(s.x0); break; case KML_P_V: v3(s.v); break; case KML_P_MBP: v3(s.mbp); break; case KML_P_Q: v3(s.q); break;
  case KML_P_SIGMA: for (int k = 0; k < 6; k++) comp[k] = s.sig[k]; *ncomp = 6; *sym = true; break;
  case KML_P_STRAIN_EL: for (int k = 0; k < 6; k++) comp[k] = s.eel[k]; *ncomp = 6; *sym = true; break;
  case KML_P_FDEF: m9(s.F); break;
  case KML_P_VOL0PK1: if (!c->c.is_TL) return fail("vol0PK1 exists only for total-Lagrangian methods"); m9(s.pk1); break;
  case KML_P_R: if (!c->c.is_TL) return fail("R exists only for total-Lagrangian methods"); m9(s.R); break;
  case KML_P_VOL0: comp[0] = s.vol0; break; case KML_P_VOL: comp[0] = s.vol; break; case KML_P_RHO0: comp[0] = s.rho0; break;
  case KML_P_MASS: comp[0] = s.mass; break; case KML_P_EFF_PLASTIC_STRAIN: comp[0] = s.eps; break;
  case KML_P_EFF_PLASTIC_STRAIN_RATE: comp[0] = s.epsdot; break; case KML_P_DAMAGE: comp[0] = s.dmg; break;
  case KML_P_DAMAGE_INIT: comp[0] = s.dmgi; break; case KML_P_IENERGY: comp[0] = s.ien; break; case KML_P_T: comp[0] = s.T; break;
  case KML_P_GAMMA: comp[0] = s.gamma; break;
  case KML_P_A: case KML_P_V_UPDATE:
    if (!s.acc[0]) return fail("kml: KML_P_A / KML_P_V_UPDATE / KML_P_F are kept only after kml_keep_particle_acceleration (called before the first step)");
    v3(field == KML_P_A ? s.acc : s.vup); break;
  default: return fail("particle field " + std::to_string(field) + " is not materialised by the CUDA engine (derived in-kernel)");
  }
  return 0;
}
static const int SYM_OF[9] = {0, 3, 4, 3, 1, 5, 4, 5, 2}; // row-major (a,b) -> (xx,yy,zz,xy,xz,yz)

// CPDI domain fields: host rows [np][2][3] (RP, RP0) / [np][4][3] (XPC, XPC0); the device keeps the x, y components
static bool cpdi_rowmap(Solid *S, int field, RowMap &rm) {
  CpdiDev &cp = S->cp; if (!S->cpbuf) return false;
  const bool r = field == KML_P_RP || field == KML_P_RP0, x = field == KML_P_XPC || field == KML_P_XPC0;
  if (!r && !x) return false;
  const int nv = r ? 2 : 4; rm.ncols = 3 * nv; rm.ncomp = 2 * nv;
  for (int k = 0; k < nv; k++) for (int d = 0; d < 2; d++) {
    rm.comp[2 * k + d] = field == KML_P_RP ? cp.rp[k][d] : (field == KML_P_RP0 ? cp.rp0[k][d] : (field == KML_P_XPC ? cp.xpc[k][d] : cp.xpc0[k][d]));
    rm.col[2 * k + d] = 3 * k + d;
  }
  return true;
}

int kml_solid_upload(kml_ctx *c, int sid, int field, const void *src) {
  CU(cudaSetDevice(c->dev));
  Solid *S = c->solids[sid]; const long long np = S->s.np;
  if (field == KML_P_PTAG || field == KML_P_MASK || field == KML_P_X0) S->gen++;
  if (field == KML_P_A || field == KML_P_V_UPDATE || field == KML_P_F || field == KML_P_J || field == KML_P_RHO || field == KML_P_R) return fail("particle field " + std::to_string(field) + " is download-only");
  if (field == KML_P_PTAG) { CU(cudaMemcpyAsync(S->s.ptag, src, sizeof(long long) * np, cudaMemcpyHostToDevice, c->stream)); CU(cudaStreamSynchronize(c->stream)); return 0; }
  if (field == KML_P_MASK) { CU(cudaMemcpyAsync(S->s.mask, src, sizeof(int) * np, cudaMemcpyHostToDevice, c->stream)); CU(cudaStreamSynchronize(c->stream)); return 0; }
  // KML_P_X between advance_particles and the next weight evaluation lands in the advanced positions (xn, see solid_field): the
  // rest of the step keeps the step-start weights, like the reference's cached lists (fix velocity_particles, src/fix_velocity_particles.cpp:228-300)
  if (field == KML_P_MBP) S->mbp_nonzero = true;
  { RowMap rm;
    if (cpdi_rowmap(S, field, rm)) {
      if (stage_reserve(c, sizeof(double) * np * rm.ncols)) return 1;
      CU(cudaMemcpyAsync(c->d_stage, src, sizeof(double) * np * rm.ncols, cudaMemcpyHostToDevice, c->stream));
      k_rows_to_soa<<<nblocks(np, 256), 256, 0, c->stream>>>((const double *)c->d_stage, rm, np);
      if (check_launch("k_rows_to_soa")) return 1;
      CU(cudaStreamSynchronize(c->stream)); return 0;
    } }
  double *comp[9]; int nc; bool sym;
  if (solid_field(c, S, field, comp, &nc, &sym)) return 1;
  // rows [np][ncols] on the host -> SoA components on the device: one H2D copy + a transposition kernel
  const int ncols = sym ? 9 : nc;
  if (stage_reserve(c, sizeof(double) * np * ncols)) return 1;
  CU(cudaMemcpyAsync(c->d_stage, src, sizeof(double) * np * ncols, cudaMemcpyHostToDevice, c->stream));
  RowMap rm; rm.ncols = ncols; rm.ncomp = nc;
  static const int RM_OF_SYM[6] = {0, 4, 8, 1, 2, 5};
  for (int k = 0; k < nc; k++) { rm.comp[k] = comp[k]; rm.col[k] = sym ? RM_OF_SYM[k] : k; }
  k_rows_to_soa<<<nblocks(np, 256), 256, 0, c->stream>>>((const double *)c->d_stage, rm, np);
  if (check_launch("k_rows_to_soa")) return 1;
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int kml_solid_download(kml_ctx *c, int sid, int field, void *dst) {
  CU(cudaSetDevice(c->dev));
  Solid *S = c->solids[sid]; const long long np = S->s.np;
  if (field == KML_P_PTAG) { CU(cudaMemcpyAsync(dst, S->s.ptag, sizeof(long long) * np, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream)); return 0; }
  if (field == KML_P_MASK) { CU(cudaMemcpyAsync(dst, S->s.mask, sizeof(int) * np, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream)); return 0; }
  if (field == KML_P_F) { // f_p = a_p m_p, src/solid.cpp:615
    std::vector<double> a(3 * np), m(np);
    if (kml_solid_download(c, sid, KML_P_A, a.data()) || kml_solid_download(c, sid, KML_P_MASS, m.data())) return 1;
    for (long long i = 0; i < np; i++) for (int d = 0; d < 3; d++) ((double *)dst)[3 * i + d] = a[3 * i + d] * m[i];
    return 0;
  }
  if (field == KML_P_J || field == KML_P_RHO) { // derived: J = det F, rho = rho0 / J (src/solid.cpp:1201-1217)
    std::vector<double> F(9 * np), r0(np);
    if (kml_solid_download(c, sid, KML_P_FDEF, F.data()) || kml_solid_download(c, sid, KML_P_RHO0, r0.data())) return 1;
    for (long long i = 0; i < np; i++) {
      const double *m = &F[9 * i];
      double J = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
      ((double *)dst)[i] = field == KML_P_J ? J : r0[i] / J;
    }
    return 0;
  }
  { RowMap rm;
    if (cpdi_rowmap(S, field, rm)) {
      if (stage_reserve(c, sizeof(double) * np * rm.ncols)) return 1;
      CU(cudaMemsetAsync(c->d_stage, 0, sizeof(double) * np * rm.ncols, c->stream)); // the z components
      k_soa_to_rows<<<nblocks(np, 256), 256, 0, c->stream>>>((double *)c->d_stage, rm, np);
      if (check_launch("k_soa_to_rows")) return 1;
      CU(cudaMemcpyAsync(dst, c->d_stage, sizeof(double) * np * rm.ncols, cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream)); return 0;
    } }
  double *comp[9]; int nc; bool sym;
  if (solid_field(c, S, field, comp, &nc, &sym)) return 1;
  const int ncols = sym ? 9 : nc;
  if (stage_reserve(c, sizeof(double) * np * ncols)) return 1;
  RowMap rm; rm.ncols = ncols; rm.ncomp = ncols;
  for (int k = 0; k < ncols; k++) { rm.comp[k] = comp[sym ? SYM_OF[k] : k]; rm.col[k] = k; }
  k_soa_to_rows<<<nblocks(np, 256), 256, 0, c->stream>>>((double *)c->d_stage, rm, np);
  if (check_launch("k_soa_to_rows")) return 1;
  CU(cudaMemcpyAsync(dst, c->d_stage, sizeof(double) * np * ncols, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int kml_solid_delete_particles(kml_ctx *c, int sid, const int *dlist) {
  CU(cudaSetDevice(c->dev));
  Solid *S = c->solids[sid]; const long long np = S->s.np;
  if (c->c.is_CPDI) return fail("kml: delete_particles with CPDI is not supported (the reference does not move the particle domains either, src/solid.cpp:1590-1610)");
  if (!c->c.is_TL && S->moved) { for (int k = 0; k < 3; k++) std::swap(S->s.x[k], S->s.xn[k]); S->moved = false; }
  const std::vector<int> src = delete_order(dlist, np); const long long n = (long long)src.size();
  if (n == np) return 0;
  int *d_idx = nullptr;
  CU(cudaMalloc(&d_idx, sizeof(int) * std::max<long long>(n, 1)));
  CU(cudaMemcpyAsync(d_idx, src.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
  if (stage_reserve(c, sizeof(double) * std::max<long long>(n, 1))) { cudaFree(d_idx); return 1; }
  const int nd = (c->c.is_TL ? SOLID_NDBL_TL : SOLID_NDBL_UL) + ((c->apic || c->c.ge) ? 9 : 0);
  if (n > 0) {
    for (int a = 0; a < nd; a++) { // every double array of the solid, then tag and mask
      double *arr = S->buf + (size_t)a * S->cap;
      k_gather_rows<double><<<nblocks(n, 256), 256, 0, c->stream>>>((double *)c->d_stage, arr, d_idx, n);
      CU(cudaMemcpyAsync(arr, c->d_stage, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    }
    for (int a = 0; S->accbuf && a < 6; a++) {
      double *arr = S->accbuf + (size_t)a * S->cap;
      k_gather_rows<double><<<nblocks(n, 256), 256, 0, c->stream>>>((double *)c->d_stage, arr, d_idx, n);
      CU(cudaMemcpyAsync(arr, c->d_stage, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    }
    k_gather_rows<long long><<<nblocks(n, 256), 256, 0, c->stream>>>((long long *)c->d_stage, S->s.ptag, d_idx, n);
    CU(cudaMemcpyAsync(S->s.ptag, c->d_stage, sizeof(long long) * n, cudaMemcpyDeviceToDevice, c->stream));
    k_gather_rows<int><<<nblocks(n, 256), 256, 0, c->stream>>>((int *)c->d_stage, S->s.mask, d_idx, n);
    CU(cudaMemcpyAsync(S->s.mask, c->d_stage, sizeof(int) * n, cudaMemcpyDeviceToDevice, c->stream));
  }
  const int rc = check_launch("k_gather_rows");
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(d_idx);
  S->s.np = n; S->d.np = n; S->gen++;
  c->tl_mass_done = false; // TL: node masses are computed once from the particle set (update_mass_nodes, src/tlmpm.cpp:345-351)
  return rc;
}

int kml_solid_device_ptr(kml_ctx *c, int sid, int field, int comp_idx, void **dptr) {
  Solid *S = c->solids[sid];
  if (field == KML_P_PTAG) { *dptr = S->s.ptag; return 0; }
  if (field == KML_P_MASK) { *dptr = S->s.mask; return 0; }
  double *comp[9]; int nc; bool sym;
  if (solid_field(c, S, field, comp, &nc, &sym)) return 1;
  if (comp_idx < 0 || comp_idx >= nc) return fail("component out of range");
  *dptr = comp[comp_idx]; return 0;
}

// ---- set-up on the device ---------------------------------------------------------------------------
int kml_has_device_setup(void) { return 1; }

static long long lattice_points(const kml_lattice *l) { return (long long)l->nsub[0] * l->nsub[1] * l->nsub[2] * l->nip; }

int kml_lattice_histogram(kml_ctx *c, const kml_lattice *lat, const kml_region *reg, int64_t *hist, int nbins) {
  CU(cudaSetDevice(c->dev));
  if (lat->nip > 64 || nbins < 1) return fail("kml_lattice_histogram: bad arguments");
  const long long ntot = lattice_points(lat);
  unsigned long long *d_hist = nullptr;
  CU(cudaMalloc(&d_hist, sizeof(unsigned long long) * nbins)); CU(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nbins, c->stream));
  if (ntot > 0) k_lattice_hist<<<nblocks(ntot, 256), 256, 0, c->stream>>>(*lat, *reg, ntot, d_hist, nbins);
  const int rc = check_launch("k_lattice_hist");
  static_assert(sizeof(int64_t) == sizeof(unsigned long long), "64-bit histogram");
  if (!rc) { CU(cudaMemcpyAsync(hist, d_hist, sizeof(int64_t) * nbins, cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream)); }
  cudaFree(d_hist);
  return rc;
}

int kml_solid_populate(kml_ctx *c, int sid, const kml_lattice *lat, const kml_region *reg, int64_t tag_offset) {
  CU(cudaSetDevice(c->dev));
  Solid *S = c->solids[sid];
  if (c->c.is_CPDI) return fail("kml_solid_populate: CPDI particle domains are set up by the host");
  const long long ntot = lattice_points(lat);
  if (ntot <= 0) return fail("kml_solid_populate: empty lattice");
  const long long nb = (ntot + LATTICE_BLOCK - 1) / LATTICE_BLOCK;
  long long *d_cnt = nullptr; void *d_tmp = nullptr; size_t tmp_bytes = 0;
  CU(cudaMalloc(&d_cnt, sizeof(long long) * (nb + 1)));
  CU(cudaMemsetAsync(d_cnt + nb, 0, sizeof(long long), c->stream));
  k_lattice_count<<<(unsigned)nb, LATTICE_BLOCK, 0, c->stream>>>(*lat, *reg, ntot, d_cnt);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_cnt, (int)(nb + 1), c->stream);
  if (cudaMalloc(&d_tmp, tmp_bytes) != cudaSuccess) { cudaFree(d_cnt); return fail("kml_solid_populate: out of memory"); }
  cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, d_cnt, (int)(nb + 1), c->stream);
  long long total = 0;
  CU(cudaMemcpyAsync(&total, d_cnt + nb, sizeof(long long), cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
  int rc = 0;
  if (total != S->s.np) rc = fail("kml_solid_populate: the lattice yields " + std::to_string(total) + " particles, the solid was created for " + std::to_string(S->s.np));
  else {
    k_lattice_fill<<<(unsigned)nb, LATTICE_BLOCK, 0, c->stream>>>(*lat, *reg, ntot, d_cnt, S->s, (long long)(lat->tag_first + tag_offset));
    rc = check_launch("k_lattice_fill");
    CU(cudaStreamSynchronize(c->stream));
  }
  cudaFree(d_cnt); cudaFree(d_tmp);
  S->gen++;
  return rc;
}

int kml_solid_group_assign(kml_ctx *c, int sid, const kml_region *reg, int bit, int64_t *count) {
  CU(cudaSetDevice(c->dev));
  Solid *S = c->solids[sid];
  unsigned long long *cnt = (unsigned long long *)(c->d_scratch + 24);
  CU(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c->stream));
  if (S->s.np > 0) k_group_assign<<<nblocks(S->s.np, 256), 256, 0, c->stream>>>(S->s, *reg, bit, cnt);
  if (check_launch("k_group_assign")) return 1;
  CU(cudaMemcpyAsync(c->h_pinned + 56, cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
  unsigned long long n; memcpy(&n, c->h_pinned + 56, sizeof n);
  if (count) *count = (int64_t)n;
  S->gen++;
  return 0;
}

int kml_solid_sum(kml_ctx *c, int sid, int field, int comp_idx, double *sum) {
  CU(cudaSetDevice(c->dev));
  Solid *S = c->solids[sid];
  double *comp[9]; int nc; bool sym;
  if (solid_field(c, S, field, comp, &nc, &sym)) return 1;
  if (comp_idx < 0 || comp_idx >= nc) return fail("component out of range");
  CU(cudaMemsetAsync(c->d_scratch + 8, 0, sizeof(double), c->stream));
  if (S->s.np > 0) k_sum<<<std::min<unsigned>(nblocks(S->s.np, 256), 148 * 8), 256, 0, c->stream>>>(comp[comp_idx], S->s.np, c->d_scratch + 8);
  if (check_launch("k_sum")) return 1;
  CU(cudaMemcpyAsync(c->h_pinned + 40, c->d_scratch + 8, sizeof(double), cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream));
  *sum = c->h_pinned[40]; return 0;
}

int kml_fix_set_particles_expr(kml_ctx *c, int solid, int groupbit, int field, int set_mask, const kml_expr prog[3]) {
  CU(cudaSetDevice(c->dev));
  for (int d = 0; d < 3; d++) if ((set_mask & (1 << d)) && (prog[d].n < 1 || prog[d].n > KML_EXPR_MAX)) return fail("kml_fix_set_particles_expr: bad program");
  ExprSet *d_prog = nullptr;
  CU(cudaMalloc(&d_prog, sizeof(ExprSet)));
  CU(cudaMemcpyAsync(d_prog, prog, sizeof(ExprSet), cudaMemcpyHostToDevice, c->stream));
  int rc = 0;
  for (size_t is = 0; is < c->solids.size() && !rc; is++) {
    if (solid != -1 && (int)is != solid) continue;
    Solid *S = c->solids[is];
    if (S->s.np == 0) continue;
    double *comp[9]; int nc; bool sym;
    if (solid_field(c, S, field, comp, &nc, &sym) || nc != 3) { rc = fail("kml_fix_set_particles_expr: the field must be a particle vector"); break; }
    k_set_particles_expr<<<nblocks(S->s.np, 128), 128, 0, c->stream>>>(S->s, groupbit, comp[0], comp[1], comp[2], set_mask, d_prog, (!c->c.is_TL && S->moved) ? 1 : 0);
    c->launches[KML_STAGE_OTHER]++;
    rc = check_launch("k_set_particles_expr");
  }
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(d_prog);
  return rc;
}

int kml_set_dt(kml_ctx *c, double dt) { if (resolve_dt(c)) return 1; c->dt = dt; return 0; }
int kml_get_dt(kml_ctx *c, double *dt) { if (resolve_dt(c)) return 1; *dt = c->dt; return 0; }

// ---- stages -----------------------------------------------------------------------------------
#define NC(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail(std::string(#call) + ": " + nccl().GetErrorString(r_)); } while (0)
static int halo_or_rigid(kml_ctx *c, Grid *G, int stage);
static std::vector<Grid *> active_grids(kml_ctx *c);

int kml_compute_grid_weight_functions_and_gradients(kml_ctx *c) {
  CU(cudaSetDevice(c->dev));
  c->steps_started++;
  // Weights are functions of the step-start positions: make the positions advanced by
  // the previous step current.  UL re-bins the particles by cell for the cell-centric P2G.
  if (!c->c.is_TL)
    for (Solid *S : c->solids) if (S->moved) { for (int k = 0; k < 3; k++) std::swap(S->s.x[k], S->s.xn[k]); S->moved = false; }
  if (c->c.is_CPDI) { // explicit per-particle lists: every step for UL, once for TL (update_wf, src/tlcpdi.cpp:100,351)
    if (c->c.is_TL && c->tl_wf_done) return 0;
    StageTimer t(c, KML_STAGE_REBIN);
    for (Solid *S : c->solids) {
      Grid *G = c->grids[S->d.grid];
      const unsigned nb = nblocks(S->s.np, 64);
#define KML_CPDI_W(SH) do { if (c->c.is_TL) k_cpdi_weights<SH, true><<<nb, 64, 0, c->stream>>>(S->s, G->g, S->cp, c->c.boxlo[0], c->c.boxlo[1], c->d_flags); \
                            else k_cpdi_weights<SH, false><<<nb, 64, 0, c->stream>>>(S->s, G->g, S->cp, c->c.boxlo[0], c->c.boxlo[1], c->d_flags); } while (0)
      switch (c->c.shape_function) {
      case KML_SHAPE_LINEAR: KML_CPDI_W(KML_SHAPE_LINEAR); break;
      case KML_SHAPE_CUBIC_SPLINE: KML_CPDI_W(KML_SHAPE_CUBIC_SPLINE); break;
      case KML_SHAPE_QUADRATIC_SPLINE: KML_CPDI_W(KML_SHAPE_QUADRATIC_SPLINE); break;
      default: KML_CPDI_W(KML_SHAPE_BERNSTEIN); break;
      }
#undef KML_CPDI_W
      c->launches[KML_STAGE_REBIN]++;
      if (check_launch("k_cpdi_weights")) return 1;
    }
    c->tl_wf_done = true;
    return 0;
  }
  if (c->has_rigid && !(c->c.is_TL && c->tl_wf_done)) { // Grid::rigid, set where a rigid particle's weight is non-zero (src/ulmpm.cpp:263-268, src/tlmpm.cpp:279-283)
    StageTimer t(c, KML_STAGE_REBIN);
    StepParams sp = step_params(c);
    for (Solid *S : c->solids) {
      if (!S->rigid || S->s.np == 0) continue;
      Grid *G = c->grids[S->d.grid];
      KML_DISPATCH(p2g, S->s, G->g, sp, P2G_MARK_RIGID, c->stream);
      c->launches[KML_STAGE_REBIN]++;
      if (check_launch("k_p2g(mark rigid)")) return 1;
    }
    if (c->comm.nranks > 1) for (Grid *G : active_grids(c)) if (halo_or_rigid(c, G, KML_STAGE_REBIN)) return 1;
    if (c->c.is_TL) c->tl_wf_done = true;
  }
  if (!c->c.is_TL) {
    if (c->use_cell_p2g && !c->apic && !c->c.ge && !c->has_rigid && cell_p2g_supported(c->c.dimension, c->c.shape_function)) {
      StageTimer t(c, KML_STAGE_REBIN);
      const bool go_now = c->permute_go; // decomposed runs: the ranks agreed (through the dt all-reduce) to re-order in this step
      for (Solid *S : c->solids) {
        Grid *G = c->grids[S->d.grid];
        int nl = 0;
        if (S->cl.build(S->s, G->g, S->cap, c->stream, &nl)) return fail(std::string("cell list build: ") + cudaGetErrorString(cudaGetLastError()));
        c->launches[KML_STAGE_REBIN] += nl;
        // physical re-ordering when too many particles sit far from their cell-sorted position (the count of the PREVIOUS re-bin, read
        // without a synchronisation; the first re-bin waits for its own)
        if (c->permute_frac >= 0 && S->cl.valid) {
          if (c->steps_started == 1) { if (permute_reserve(c, S)) return 1; CU(cudaStreamSynchronize(c->stream)); }
          const long long far = S->cl.far_count();
          // measurements of earlier steps, read without waiting (an event that has not completed yet is looked at next step)
          float t = 0;
          if (S->ev_p_valid && cudaEventQuery(S->ev_p[1]) == cudaSuccess && cudaEventElapsedTime(&t, S->ev_p[0], S->ev_p[1]) == cudaSuccess) { S->permute_ms = t; S->ev_p_valid = false; }
          for (int k = 0; k < 2; k++)
            if (S->ev_s_valid[k] && cudaEventQuery(S->ev_s[k][1]) == cudaSuccess && cudaEventElapsedTime(&t, S->ev_s[k][0], S->ev_s[k][1]) == cudaSuccess) {
              S->ev_s_valid[k] = false;
              if (S->ev_s_step[k] < S->clean_step) continue;                                       // a step from before the last permute
              if (S->ev_s_step[k] == S->clean_step) { S->t_clean = t; S->excess_ms = 0; }          // the step of the permute itself: the clean time
              else if (S->t_clean > 0) S->excess_ms += std::max(0.0, (double)t - S->t_clean);
            }
          cudaGetLastError();
          bool due;
          if (S->permute_ms < 0 || (S->t_clean < 0 && S->clean_step < 0)) due = far > c->permute_frac * (double)S->s.np; // nothing measured yet: the disorder threshold
          else if (S->t_clean < 0) due = false;                                                               // the clean time of the last permute has not been read yet
          else due = S->excess_ms >= S->permute_ms && far > 0.005 * (double)S->s.np;                          // amortised rebuild
          if (c->permute_every > 0) due = c->steps_started - S->last_permute_step >= c->permute_every && far > 0; // KML_PERMUTE_EVERY: fixed period (measurements)
          if (c->comm.nranks > 1 && c->dt_collective && c->permute_every <= 0) { // collective decision (one step late): ask now, act when the all-reduce said so
            if (due && c->steps_started - S->last_permute_step >= c->permute_min_steps && c->steps_started > 1) c->permute_want = true;
            due = go_now || (c->steps_started == 1 && due);
          }
          if (due && c->steps_started - S->last_permute_step >= (c->comm.nranks > 1 && c->dt_collective ? 0 : c->permute_min_steps)) {
            if (getenv("KML_DEBUG")) fprintf(stderr, "[kml rank %d] step %lld: physical permute, %lld of %lld particles far from their cell-sorted slot (stress %.3f ms clean, %.3f ms lost since, permute %.3f ms)\n",
                                             c->c.rank, c->steps_started, far, (long long)S->s.np, S->t_clean, S->excess_ms, S->permute_ms);
            if (permute_solid(c, S, G)) return 1;
            S->last_permute_step = c->steps_started; S->cl.far_reset(); S->clean_step = c->steps_started; S->t_clean = -1; S->excess_ms = 0;
          }
        }
      }
      if (go_now) { c->permute_go = false; c->permute_want = false; c->last_collective_permute = c->steps_started; }
    }
  }
  return 0;
}

int kml_reset(kml_ctx *c) { // ULMPM::reset: mbp = 0, dtCFL = 1e22
  CU(cudaSetDevice(c->dev));
  for (Solid *S : c->solids) {
    S->dtCFL = 1.0e22;
    if (S->mbp_nonzero) { for (int k = 0; k < 3; k++) CU(cudaMemsetAsync(S->s.mbp[k], 0, sizeof(double) * S->s.np, c->stream)); S->mbp_nonzero = false; }
    const double init[2] = {0.0, 1.0};
    CU(cudaMemcpyAsync(S->red, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
  }
  return 0;
}

static std::vector<Grid *> active_grids(kml_ctx *c) {
  std::vector<Grid *> r;
  for (Solid *S : c->solids) { Grid *G = c->grids[S->d.grid]; if (std::find(r.begin(), r.end(), G) == r.end()) r.push_back(G); }
  return r;
}


// One-time, collective: every rank allocates its landing zone, the neighbours trade CUDA IPC handles (64-byte blobs through ncclSend/Recv) and map
// each other's zones; the peer path is used only if EVERY rank succeeded (KML_HALO=nccl forces the send/recv path).
static int halo_peer_setup(kml_ctx *c, size_t slot_bytes) {
  Comm &cm = c->comm; cm.peer_state = -1;
  const char *mode = getenv("KML_HALO");
  const bool want = !(mode && !strcmp(mode, "nccl"));
  const int left = cm.rank > 0, right = cm.rank < cm.nranks - 1;
  slot_bytes = (slot_bytes + 255) / 256 * 256;
  bool ok = want;
  cudaIpcMemHandle_t mine, hl, hr; memset(&mine, 0, sizeof mine); memset(&hl, 0, sizeof hl); memset(&hr, 0, sizeof hr);
  if (ok) ok = cudaMalloc(&cm.zone, HALO_ZONE_HDR + 4 * slot_bytes) == cudaSuccess && cudaMemsetAsync(cm.zone, 0, HALO_ZONE_HDR, c->stream) == cudaSuccess &&
               cudaMalloc(&cm.push_done, sizeof(unsigned)) == cudaSuccess && cudaMemsetAsync(cm.push_done, 0, sizeof(unsigned), c->stream) == cudaSuccess &&
               cudaIpcGetMemHandle(&mine, cm.zone) == cudaSuccess;
  cudaGetLastError();
  unsigned char *d_h = nullptr; // [mine | from left | from right]
  CU(cudaMalloc(&d_h, 3 * sizeof mine));
  CU(cudaMemcpyAsync(d_h, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
  NC(nccl().GroupStart());
  if (left) { NC(nccl().Send(d_h, sizeof mine, ncclChar, cm.rank - 1, cm.comm, c->stream)); NC(nccl().Recv(d_h + sizeof mine, sizeof mine, ncclChar, cm.rank - 1, cm.comm, c->stream)); }
  if (right) { NC(nccl().Send(d_h, sizeof mine, ncclChar, cm.rank + 1, cm.comm, c->stream)); NC(nccl().Recv(d_h + 2 * sizeof mine, sizeof mine, ncclChar, cm.rank + 1, cm.comm, c->stream)); }
  NC(nccl().GroupEnd());
  CU(cudaMemcpyAsync(&hl, d_h + sizeof mine, sizeof mine, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&hr, d_h + 2 * sizeof mine, sizeof mine, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(d_h);
  if (ok && left) ok = cudaIpcOpenMemHandle((void **)&cm.zone_left, hl, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  if (ok && right) ok = cudaIpcOpenMemHandle((void **)&cm.zone_right, hr, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  cudaGetLastError();
  double *flag = c->d_scratch + 40; const double mine_ok = ok ? 1.0 : 0.0;
  CU(cudaMemcpyAsync(flag, &mine_ok, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NC(nccl().AllReduce(flag, flag, 1, ncclDouble, ncclMin, c->comm.comm, c->stream));
  double all_ok = 0; CU(cudaMemcpyAsync(&all_ok, flag, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (all_ok == 1.0) { cm.peer_state = 1; cm.zone_slot = slot_bytes; cm.seq = 0; }
  else {
    if (cm.zone_left) cudaIpcCloseMemHandle(cm.zone_left); if (cm.zone_right) cudaIpcCloseMemHandle(cm.zone_right);
    cudaFree(cm.zone); cm.zone = cm.zone_left = cm.zone_right = nullptr; cudaGetLastError();
  }
  if (getenv("KML_DEBUG")) fprintf(stderr, "[kml rank %d] halo exchange: %s\n", cm.rank, cm.peer_state == 1 ? "NVLink peer memory (CUDA IPC)" : "NCCL send/recv");
  return 0;
}

// Sum of the node planes shared with the slab neighbours (see kml_comm.cuh).  Local planes [0, nsh) are shared
// with the left neighbour, [n0 - nsh, n0) with the right one, nsh = stencil span - 1.
static int halo_sum(kml_ctx *c, Grid *G, int what, int stage) {
  Comm &cm = c->comm; GridDev &g = G->g;
  const int span = c->c.shape_function == KML_SHAPE_LINEAR ? 2 : 4;
  const int nsh = span - 1;
  const long long plane = (long long)g.n[1] * g.n[2], cnt = plane * nsh;
  HaloFields hf; hf.n = 0;
  auto add = [&](double *p, int w, int skip_w) { hf.ptr[hf.n] = p; hf.width[hf.n] = w; hf.skip_w[hf.n] = skip_w; hf.n++; };
  if (what & P2G_MOM) add((double *)g.nv, 4, (what & P2G_MASS) ? 0 : 1);
  if (what & P2G_FORCE) for (int d = 0; d < 3; d++) add(g.f[d], 1, 0);
  if (what & P2G_MB) for (int d = 0; d < 3; d++) add(g.mb[d], 1, 0);
  if (what & P2G_TEMP) add(g.T, 1, 0);
  if (what & P2G_HEAT) { add(g.Qext, 1, 0); add(g.Qint, 1, 0); }
  if (hf.n == 0) return 0;
  size_t total = 0; for (int f = 0; f < hf.n; f++) total += (size_t)cnt * hf.width[f];
  const int left = cm.rank > 0, right = cm.rank < cm.nranks - 1;
  if (cm.peer_state == 0 && halo_peer_setup(c, sizeof(double) * (size_t)cnt * 13)) return 1; // 13 doubles per node: every field a pass can exchange
  if (cm.peer_state == 1 && total * sizeof(double) <= cm.zone_slot) { // pack straight into the neighbours' memory over NVLink, add from ours
    const unsigned long long seq = ++cm.seq; const size_t b = seq & 1, sl = cm.zone_slot;
    double *remL = left ? (double *)(cm.zone_left + HALO_ZONE_HDR + (2 + b) * sl) : nullptr;  // the left neighbour's "from the right" buffer
    double *remR = right ? (double *)(cm.zone_right + HALO_ZONE_HDR + (0 + b) * sl) : nullptr; // the right neighbour's "from the left" buffer
    unsigned long long *fL = left ? (unsigned long long *)(cm.zone_left + 128) : nullptr, *fR = right ? (unsigned long long *)(cm.zone_right + 0) : nullptr;
    const long long top_ = (long long)(g.n[0] - nsh) * plane;
    k_halo_push<<<nblocks(cnt, 256), 256, 0, c->stream>>>(hf, cnt, top_, remL, remR, fL, fR, seq, cm.push_done);
    const double *inL = left ? (const double *)(cm.zone + HALO_ZONE_HDR + (0 + b) * sl) : nullptr, *inR = right ? (const double *)(cm.zone + HALO_ZONE_HDR + (2 + b) * sl) : nullptr;
    k_halo_wait_add<<<nblocks(cnt, 256), 256, 0, c->stream>>>(hf, cnt, top_, inL, inR, (const unsigned long long *)(cm.zone + 0), (const unsigned long long *)(cm.zone + 128), seq, c->d_flags);
    c->launches[stage] += 2;
    return check_launch("halo_sum (peer memory)");
  }
  const size_t need = total * 4 * sizeof(double);
  if (need > cm.halo_bytes) { cudaFree(cm.halo_buf); cm.halo_buf = nullptr; CU(cudaMalloc(&cm.halo_buf, need)); cm.halo_bytes = need; }
  double *sl = cm.halo_buf, *sr = sl + total, *rl = sr + total, *rr = rl + total;
  const long long top = (long long)(g.n[0] - nsh) * plane;
  k_halo_pack<<<nblocks(cnt, 256), 256, 0, c->stream>>>(hf, cnt, top, sl, sr, left, right);
  NC(nccl().GroupStart());
  if (left) { NC(nccl().Send(sl, total, ncclDouble, cm.rank - 1, cm.comm, c->stream)); NC(nccl().Recv(rl, total, ncclDouble, cm.rank - 1, cm.comm, c->stream)); }
  if (right) { NC(nccl().Send(sr, total, ncclDouble, cm.rank + 1, cm.comm, c->stream)); NC(nccl().Recv(rr, total, ncclDouble, cm.rank + 1, cm.comm, c->stream)); }
  NC(nccl().GroupEnd());
  k_halo_add<<<nblocks(cnt, 256), 256, 0, c->stream>>>(hf, cnt, top, rl, rr, left, right);
  c->launches[stage] += 2;
  return check_launch("halo_sum");
}

// Grid::reduce_rigid_ghost_nodes (src/grid.cpp:746-879): OR of the rigid flags on the planes shared with the slab neighbours
static int halo_or_rigid(kml_ctx *c, Grid *G, int stage) {
  Comm &cm = c->comm; GridDev &g = G->g;
  const int nsh = (c->c.shape_function == KML_SHAPE_LINEAR ? 2 : 4) - 1;
  const long long plane = (long long)g.n[1] * g.n[2], cnt = plane * nsh;
  const size_t need = (size_t)cnt * 4 * sizeof(int);
  if (need > cm.halo_bytes) { cudaFree(cm.halo_buf); cm.halo_buf = nullptr; CU(cudaMalloc(&cm.halo_buf, need)); cm.halo_bytes = need; }
  const int left = cm.rank > 0, right = cm.rank < cm.nranks - 1;
  int *sl = (int *)cm.halo_buf, *sr = sl + cnt, *rl = sr + cnt, *rr = rl + cnt;
  const long long top = (long long)(g.n[0] - nsh) * plane;
  k_halo_pack_flag<<<nblocks(cnt, 256), 256, 0, c->stream>>>(g.rigid, cnt, top, sl, sr, left, right);
  NC(nccl().GroupStart());
  if (left) { NC(nccl().Send(sl, cnt, ncclInt, cm.rank - 1, cm.comm, c->stream)); NC(nccl().Recv(rl, cnt, ncclInt, cm.rank - 1, cm.comm, c->stream)); }
  if (right) { NC(nccl().Send(sr, cnt, ncclInt, cm.rank + 1, cm.comm, c->stream)); NC(nccl().Recv(rr, cnt, ncclInt, cm.rank + 1, cm.comm, c->stream)); }
  NC(nccl().GroupEnd());
  k_halo_or_flag<<<nblocks(cnt, 256), 256, 0, c->stream>>>(g.rigid, cnt, top, rl, rr, left, right);
  c->launches[stage] += 2;
  return check_launch("halo_or_rigid");
}

static int p2g_launch(kml_ctx *c, int what_in, int stage) {
  StageTimer t(c, stage);
  StepParams sp = step_params(c);
  const bool TL = c->c.is_TL;
  // zero the node accumulators touched by this pass (the "reset" branches of src/solid.cpp:317-574)
  for (size_t is = 0; is < c->solids.size(); is++) {
    Solid *S = c->solids[is]; Grid *G = c->grids[S->d.grid]; GridDev &g = G->g;
    const bool reset = TL ? true : (is == 0);
    int what = what_in;
    if (TL && (what & P2G_MASS)) { if (c->tl_mass_done) what &= ~P2G_MASS; }
    if ((what & P2G_MB) && !S->mbp_nonzero) what &= ~P2G_MB;
    const size_t nb = sizeof(double) * g.nn;
    if (reset) {
      if ((what & P2G_MASS) && (what & P2G_MOM)) CU(cudaMemsetAsync(g.nv, 0, sizeof(double4) * g.nn, c->stream));
      else if (what & P2G_MOM) { k_grid_zero_v<<<nblocks(g.nn, 256), 256, 0, c->stream>>>(g, 0); c->launches[stage]++; }
      else if (what & P2G_MASS) return fail("mass-only P2G pass is not used by any scheme");
      if (what_in & P2G_FORCE) for (int k = 0; k < 3; k++) { CU(cudaMemsetAsync(g.f[k], 0, nb, c->stream)); CU(cudaMemsetAsync(g.mb[k], 0, nb, c->stream)); }
      if (what & P2G_TEMP) CU(cudaMemsetAsync(g.T, 0, nb, c->stream));
      if (what & P2G_HEAT) { CU(cudaMemsetAsync(g.Qext, 0, nb, c->stream)); CU(cudaMemsetAsync(g.Qint, 0, nb, c->stream)); }
    }
    if (what == 0) continue;
    bool done = false;
    fill_inertia(c, G, S, sp);
    sp.ext = (sp.ext & 1) | ((c->has_rigid ? (S->rigid ? 2 : 1) : 0) << 1);
    if (!TL && !c->apic && !c->c.ge && !c->has_rigid && c->use_cell_p2g && (c->cell_mask & 1) && S->cl.valid && cell_p2g_supported(c->c.dimension, c->c.shape_function) && !sp.axisymmetric &&
        !(what & (P2G_TEMP | P2G_HEAT))) {
      int nl = 0;
      const int rc = cell_p2g3_launch(S->s, g, S->cl, what, c->p2g_nb, c->v2g_nb, c->gtune.seg_target, c->stream, &nl); // -1: combination not covered -> atomic kernel
      if (rc > 0) return fail("cell p2g launch failed");
      if (rc == 0) { c->launches[stage] += nl; done = true; }
    }
    if (!done && c->c.is_CPDI) {
      if (TL) k_cpdi_p2g<true><<<nblocks(S->s.np, 128), 128, 0, c->stream>>>(S->s, g, S->cp, what);
      else k_cpdi_p2g<false><<<nblocks(S->s.np, 128), 128, 0, c->stream>>>(S->s, g, S->cp, what);
      c->launches[stage]++; done = true;
    }
    if (!done) {
      KML_DISPATCH(p2g, S->s, g, sp, what, c->stream);
      c->launches[stage]++;
    }
    if (check_launch("k_p2g")) return 1;
    if (what & P2G_MOM) G->v_is_momentum = true;
    if (what & P2G_TEMP) G->T_is_weighted = true;
    G->nvd_valid = false;
  }
  if (TL && (what_in & P2G_MASS)) c->tl_mass_done = true;
  t.stop();
  if (c->comm.nranks > 1) { StageTimer th(c, KML_STAGE_HALO); for (Grid *G : active_grids(c)) if (halo_sum(c, G, what_in, KML_STAGE_HALO)) return 1; }
  return 0;
}

int kml_particles_to_grid(kml_ctx *c) {
  CU(cudaSetDevice(c->dev));
  int what = P2G_MASS | P2G_MOM | P2G_FORCE | P2G_MB | (c->c.temp ? (P2G_TEMP | P2G_HEAT) : 0);
  return p2g_launch(c, what, KML_STAGE_P2G);
}
int kml_particles_to_grid_USF_1(kml_ctx *c) {
  CU(cudaSetDevice(c->dev));
  return p2g_launch(c, P2G_MASS | P2G_MOM | (c->c.temp ? P2G_TEMP : 0), KML_STAGE_P2G);
}
int kml_particles_to_grid_USF_2(kml_ctx *c) {
  CU(cudaSetDevice(c->dev));
  return p2g_launch(c, P2G_FORCE | P2G_MB | (c->c.temp ? P2G_HEAT : 0), KML_STAGE_P2G);
}

int kml_update_grid_state(kml_ctx *c) {
  CU(cudaSetDevice(c->dev));
  if (resolve_dt(c)) return 1;
  StageTimer t(c, KML_STAGE_GRID);
  for (Grid *G : active_grids(c)) {
    double *nvd = nullptr;
    if (c->g2p_tma && !c->c.is_TL && cell_p2g_supported(c->c.dimension, c->c.shape_function)) {
      if (!G->nvd) { const size_t nb = sizeof(double) * nvd_doubles(G->g); CU(cudaMalloc(&G->nvd, nb)); CU(cudaMemsetAsync(G->nvd, 0, nb, c->stream)); }
      nvd = G->nvd;
    }
    k_grid_update<<<nblocks(G->g.nn, 256), 256, 0, c->stream>>>(G->g, c->dt, G->v_is_momentum, 1, c->c.temp, G->T_is_weighted, c->has_rigid, nvd);
    G->v_is_momentum = false; G->T_is_weighted = false; G->nvd_valid = nvd != nullptr; c->launches[KML_STAGE_GRID]++;
    if (check_launch("k_grid_update")) return 1;
  }
  return 0;
}

int kml_grid_to_points(kml_ctx *c) { c->pending_g2p = true; return 0; }

int kml_advance_particles(kml_ctx *c) {
  CU(cudaSetDevice(c->dev));
  if (resolve_dt(c)) return 1;
  if (!c->pending_g2p) return fail("advance_particles called without grid_to_points");
  c->pending_g2p = false;
  StageTimer t(c, KML_STAGE_G2P);
  StepParams sp = step_params(c);
  for (Solid *S : c->solids) {
    Grid *G = c->grids[S->d.grid];
    if (grid_normalize_if_needed(c, G)) return 1;
    int rc = -1;
    fill_inertia(c, G, S, sp);
    sp.ext = (sp.ext & 1) | ((c->has_rigid ? (S->rigid ? 2 : 1) : 0) << 1);
    if (!c->c.is_TL && !c->apic && !c->c.ge && !c->has_rigid && !c->keep_acc && c->use_cell_p2g && (c->cell_mask & 2) && S->cl.valid && cell_p2g_supported(c->c.dimension, c->c.shape_function)) {
      if (c->g2p_tma && G->nvd && !sp.axisymmetric && !sp.temp) {
        if (!G->nvd_valid) { k_grid_pack_g2p<<<nblocks(G->g.nn, 256), 256, 0, c->stream>>>(G->g, G->nvd); G->nvd_valid = true; c->launches[KML_STAGE_G2P]++; }
        if (c->g2p_tma == 4) rc = cell_g2p_bulk_launch(S->s, G->g, sp, G->nvd, S->cl, c->stream, c->gtune.seg_g2p, c->gtune.g2p_threads); // one block per segment, bulk-copied tile
        else rc = cell_g2p_tma_launch(S->s, G->g, sp, G->nvd, S->cl, c->stream, c->gtune.seg_g2p, c->gtune.g2p_threads, c->gtune.g2p_threads == 64 ? 8 : (c->g2p_tma == 3 ? 3 : 4), c->nsm);
        if (rc > 0) return fail(std::string("cell g2p (TMA) launch failed: ") + cudaGetErrorString(cudaGetLastError()));
      } else {
        StressParams none{}; rc = cell_gather_launch(false, S->s, G->g, sp, none, S->d.mat, S->cl, c->stream, c->gtune);
        if (rc > 0) return fail("cell g2p launch failed");
      }
    }
    if (c->c.is_CPDI) {
      if (c->c.is_TL) k_cpdi_g2p<true><<<nblocks(S->s.np, 128), 128, 0, c->stream>>>(S->s, G->g, S->cp, sp);
      else k_cpdi_g2p<false><<<nblocks(S->s.np, 128), 128, 0, c->stream>>>(S->s, G->g, S->cp, sp);
    } else if (rc < 0) KML_DISPATCH(g2p, S->s, G->g, sp, c->stream);
    c->launches[KML_STAGE_G2P]++;
    if (check_launch("k_g2p")) return 1;
    if (!c->c.is_TL) S->moved = true;
  }
  return 0;
}

int kml_velocities_to_grid(kml_ctx *c) {
  CU(cudaSetDevice(c->dev));
  if (c->pending_g2p) return fail("velocities_to_grid called between grid_to_points and advance_particles");
  if (p2g_launch(c, P2G_MOM | P2G_POSMOVED | (c->c.temp ? P2G_TEMP : 0), KML_STAGE_V2G)) return 1;
  // the reference divides by the node mass inside compute_velocity_nodes; fixes that follow
  // (post_velocities_to_grid) and the gradient gather need velocities, so normalise now
  StageTimer t(c, KML_STAGE_V2G);
  for (Grid *G : active_grids(c)) if (grid_normalize_if_needed(c, G)) return 1;
  return 0;
}

int kml_update_grid_positions(kml_ctx *c) {
  if (!c->c.is_TL) return 0;
  CU(cudaSetDevice(c->dev));
  if (resolve_dt(c)) return 1;
  StageTimer t(c, KML_STAGE_GRID);
  for (Grid *G : active_grids(c)) {
    if (grid_normalize_if_needed(c, G)) return 1;
    k_grid_positions<<<nblocks(G->g.nn, 256), 256, 0, c->stream>>>(G->g, c->dt); c->launches[KML_STAGE_GRID]++;
  }
  return check_launch("k_grid_positions");
}

int kml_compute_rate_deformation_gradient(kml_ctx *c, int doublemapping) {
  c->pending_grad = doublemapping ? 1 : 0;
  c->grad_moved = false;
  for (Solid *S : c->solids) if (S->moved) c->grad_moved = true;
  return 0;
}
int kml_update_deformation_gradient(kml_ctx *c) {
  if (c->pending_grad < 0) return fail("update_deformation_gradient called without compute_rate_deformation_gradient");
  c->pending_F = true; return 0;
}

int kml_update_stress(kml_ctx *c, int doublemapping) {
  CU(cudaSetDevice(c->dev));
  if (resolve_dt(c)) return 1;
  if (!c->pending_F || c->pending_grad < 0) return fail("update_stress called without compute_rate_deformation_gradient + update_deformation_gradient");
  StageTimer t(c, KML_STAGE_STRESS);
  StepParams sp = step_params(c);
  for (Solid *S : c->solids) {
    Grid *G = c->grids[S->d.grid];
    if (grid_normalize_if_needed(c, G)) return 1;
    if (S->rigid) continue; // src/solid.cpp:799,862,1157,1248: the gradient, F and stress updates return at once for a rigid material
    sp.inv_tav = S->d.mat.signal_velocity / (1000 * G->d.cellsize);
    StressParams tp; tp.doublemapping = c->pending_grad; tp.moved = c->grad_moved; tp.max_wave = S->red; tp.min_h_ratio = S->red + 1;
    (void)doublemapping; // heat flux uses the same nodal field choice as the gradient in every scheme (usl/musl/usf)
    int rc = -1;
    fill_inertia(c, G, S, sp);
    if (!c->c.is_TL && !c->apic && !c->c.ge && !c->has_rigid && c->use_cell_p2g && (c->cell_mask & 4) && S->cl.valid && cell_p2g_supported(c->c.dimension, c->c.shape_function)) {
      const int k = S->ev_s_next; S->ev_s_next ^= 1; // timed every step for the permute policy (kml_compute_grid_weight_...)
      if (!S->ev_s[k][0]) { CU(cudaEventCreate(&S->ev_s[k][0])); CU(cudaEventCreate(&S->ev_s[k][1])); }
      CU(cudaEventRecord(S->ev_s[k][0], c->stream));
      rc = cell_gather_launch(true, S->s, G->g, sp, tp, S->d.mat, S->cl, c->stream, c->gtune);
      if (rc > 0) return fail("cell stress launch failed");
      if (rc == 0) { CU(cudaEventRecord(S->ev_s[k][1], c->stream)); S->ev_s_valid[k] = true; S->ev_s_step[k] = c->steps_started; }
    }
    if (c->c.is_CPDI) {
      if (c->c.is_TL) k_cpdi_stress<true><<<nblocks(S->s.np, 128), 128, 0, c->stream>>>(S->s, G->g, S->cp, sp, tp, S->d.mat);
      else k_cpdi_stress<false><<<nblocks(S->s.np, 128), 128, 0, c->stream>>>(S->s, G->g, S->cp, sp, tp, S->d.mat);
    } else if (rc < 0) KML_DISPATCH(stress, S->s, G->g, sp, tp, S->d.mat, c->stream);
    c->launches[KML_STAGE_STRESS]++;
    if (check_launch("k_stress")) return 1;
  }
  c->pending_F = false; c->pending_grad = -1;
  return 0;
}

// bit b of the device error word -> slot b of a double array (1.0 / 0.0): a max all-reduce of the slots is the union of the words
__global__ void k_flag_bits(const unsigned *flags, double *bits, double want_permute) { if (threadIdx.x < 8) bits[threadIdx.x] = ((*flags >> threadIdx.x) & 1u) ? 1.0 : 0.0; if (threadIdx.x == 8) bits[8] = want_permute; }
__global__ void k_bits_flag(const double *bits, unsigned *flags) { unsigned f = 0; for (int b = 0; b < 8; b++) if (bits[b] != 0.0) f |= 1u << b; *flags = f; }

// The dt of the next step is needed first by the grid update of the next step, a re-bin and a scatter later.  adjust_dt therefore only
// ENQUEUES the reduction and the read-back; the value is resolved (one event wait, normally long satisfied) when something asks for it:
// a kernel that takes dt, kml_get_dt, or a caller that passes dt_out.  The device error word travels with it.
static int resolve_dt(kml_ctx *c) {
  if (!c->dt_pending) return 0;
  c->dt_pending = false;
  CU(cudaEventSynchronize(c->ev_dt));
  const int ns = (int)c->solids.size();
  unsigned flags = 0; for (int b = 0; b < 8; b++) if (c->h_red[KML_RED_BITS + b] != 0.0) flags |= 1u << b;
  // some rank asked: every rank re-orders at its next re-bin.  A request that was made before the permute of this or the previous step is stale
  // (the ranks around the threshold ask one after the other; the all-reduce delivers each request one step late).
  if (c->comm.nranks > 1 && c->h_red[KML_RED_BITS + 8] != 0.0 && c->steps_started - c->last_collective_permute > 1) c->permute_go = true;
  if (flags) return fail("device error flags " + std::to_string(flags) + " (1: particle left the domain, 2: J<=0, 4: NaN wave speed, 8: polar decomposition failed, 16: CPDI neighbour list overflow, 32: particle migration bookkeeping, 64: halo exchange timed out - a neighbour rank never delivered its planes)");
  double dtCFL = 1.0e22;
  for (int i = 0; i < ns; i++) { // src/solid.cpp:1429 then src/ulmpm.cpp:525-551
    Solid *S = c->solids[i]; Grid *G = c->grids[S->d.grid];
    const double wave = c->h_red[2 * i], hr = c->c.is_TL ? c->h_red[2 * i + 1] : 1.0;
    S->dtCFL = std::min(1.0e22, G->d.cellsize * hr / wave);
    dtCFL = std::min(dtCFL, S->dtCFL);
  }
  if (dtCFL == 0 || std::isnan(dtCFL)) return fail("dtCFL == 0 or NaN");
  c->dt = dtCFL * c->dt_factor;
  return 0;
}

int kml_adjust_dt(kml_ctx *c, double dt_factor, double *dt_out) {
  CU(cudaSetDevice(c->dev));
  if (resolve_dt(c)) return 1; // an adjust_dt nobody consumed (two calls in a row)
  {
    StageTimer t(c, KML_STAGE_DT);
    const int ns = (int)c->solids.size();
    if (2 * ns > KML_RED_BITS) return fail("too many solids");
    k_flag_bits<<<1, 32, 0, c->stream>>>(c->d_flags, c->d_red + KML_RED_BITS, c->permute_want ? 1.0 : 0.0);
    c->permute_want = false;
    if (c->comm.nranks > 1) { // MPI_Allreduce(MIN) of dtCFL in the reference (src/ulmpm.cpp:547) == max of the wave speeds here; ONE collective for all solids + the error word
      NC(nccl().AllReduce(c->d_red, c->d_red, KML_RED_N, ncclDouble, ncclMax, c->comm.comm, c->stream));
      k_bits_flag<<<1, 1, 0, c->stream>>>(c->d_red + KML_RED_BITS, c->d_flags);
    }
    CU(cudaMemcpyAsync(c->h_red, c->d_red, KML_RED_N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev_dt, c->stream));
    c->launches[KML_STAGE_DT] += 1;
  }
  c->dt_pending = true; c->dt_factor = dt_factor; c->dt_collective = true;
  if (dt_out) { if (resolve_dt(c)) return 1; *dt_out = c->dt; }
  return 0;
}

// ULMPM::exchange_particles (src/ulmpm.cpp:565-667): particles whose stencil base left this rank's slab move to the
// neighbour that owns it.  The positions advanced by grid_to_points become current here.
int kml_exchange_particles(kml_ctx *c) {
  Comm &cm = c->comm;
  if (cm.nranks <= 1) return 0;
  CU(cudaSetDevice(c->dev));
  StageTimer t(c, KML_STAGE_MIGRATE);
  for (Solid *S : c->solids) {
    Grid *G = c->grids[S->d.grid]; SolidDev &s = S->s;
    if (S->moved) { for (int k = 0; k < 3; k++) std::swap(s.x[k], s.xn[k]); S->moved = false; }
    const int xs = (int)((s.x[0] - S->buf) / S->cap); // buffer slot of the current positions (0 or 3), see k_mig_pack
    const int narr = SOLID_NDBL_UL + (s.Lst[0] ? 9 : 0); // the stored velocity gradient of the APIC family / gradient-enhanced projection follows the UL block
    auto reserve = [&](long long want) -> int { // migration scratch for `want` particles per direction (never shrinks)
      const int cap_mig = (int)std::min<long long>(want, 1 << 26);
      if (cap_mig <= cm.mig_cap && (size_t)(narr + 2) * 2 * cm.mig_cap * sizeof(double) <= cm.mig_bytes) return 0;
      const int cap_new = std::max(cap_mig, cm.mig_cap);
      cudaFree(cm.mig_list); cudaFree(cm.mig_send); cudaFree(cm.mig_recv); cudaFree(cm.mig_flag);
      CU(cudaMalloc(&cm.mig_list, sizeof(int) * 6 * (size_t)cap_new)); // leavers left | right (cap each), holes, fillers (2 cap each: up to nL + nR entries)
      CU(cudaMalloc(&cm.mig_flag, sizeof(int) * 2 * (size_t)cap_new));
      cm.mig_bytes = sizeof(double) * (size_t)(narr + 2) * 2 * cap_new;
      CU(cudaMalloc(&cm.mig_send, cm.mig_bytes)); CU(cudaMalloc(&cm.mig_recv, cm.mig_bytes));
      cm.mig_cap = cap_new;
      return 0;
    };
    auto mark = [&]() -> int {
      CU(cudaMemsetAsync(cm.mig_cnt, 0, 8 * sizeof(int), c->stream));
      k_mig_mark<<<nblocks(s.np, 256), 256, 0, c->stream>>>(s.x[0], s.np, G->g.lo[0], G->g.inv_cellsize, c->c.shape_function == KML_SHAPE_LINEAR,
                                                            G->d.base_lo, G->d.base_hi, cm.rank, cm.nranks, cm.mig_cnt, cm.mig_list, cm.mig_cap);
      return 0;
    };
    { const char *e = getenv("KML_MIG_CAP"); if (reserve(e && *e ? atoll(e) : std::max<long long>(s.np / 8, 1024))) return 1; } // KML_MIG_CAP: test knob (forces the growth path)
    if (mark()) return 1;
    const bool left = cm.rank > 0, right = cm.rank < cm.nranks - 1;
    NC(nccl().GroupStart());
    if (left) { NC(nccl().Send(cm.mig_cnt + 0, 1, ncclInt, cm.rank - 1, cm.comm, c->stream)); NC(nccl().Recv(cm.mig_cnt + 2, 1, ncclInt, cm.rank - 1, cm.comm, c->stream)); }
    if (right) { NC(nccl().Send(cm.mig_cnt + 1, 1, ncclInt, cm.rank + 1, cm.comm, c->stream)); NC(nccl().Recv(cm.mig_cnt + 3, 1, ncclInt, cm.rank + 1, cm.comm, c->stream)); }
    NC(nccl().GroupEnd());
    CU(cudaMemcpyAsync(cm.h_cnt, cm.mig_cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const int nL = cm.h_cnt[0], nR = cm.h_cnt[1], rL = cm.h_cnt[2], rR = cm.h_cnt[3];
    c->launches[KML_STAGE_MIGRATE]++;
    if (nL > cm.mig_cap || nR > cm.mig_cap || rL > cm.mig_cap || rR > cm.mig_cap) {
      // more migrants than the scratch holds (a body entering a slab): grow it and, if this rank's own lists were cut off, mark again.  Every rank
      // decides from counts it already has; the payload exchange below is sized by those counts, so no rank waits for another's decision.
      const bool relist = nL > cm.mig_cap || nR > cm.mig_cap;
      const long long need = std::max(std::max(nL, nR), std::max(rL, rR));
      if (reserve(need + need / 2 + 1024)) return 1;
      if (relist && mark()) return 1; // same particles, same counts; only the lists were truncated
    }
    if (nL + nR + rL + rR == 0) continue;
    const long long np_new = s.np - nL - nR;
    if (np_new + rL + rR > S->cap) return fail("particle capacity exceeded by migration");
    double *sendL = cm.mig_send, *sendR = cm.mig_send + (size_t)(narr + 2) * cm.mig_cap;
    double *recvL = cm.mig_recv, *recvR = cm.mig_recv + (size_t)(narr + 2) * cm.mig_cap;
    if (nL) k_mig_pack<<<nblocks(nL, 128), 128, 0, c->stream>>>(sendL, cm.mig_list, nL, S->buf, S->cap, narr, xs, s.ptag, s.mask);
    if (nR) k_mig_pack<<<nblocks(nR, 128), 128, 0, c->stream>>>(sendR, cm.mig_list + cm.mig_cap, nR, S->buf, S->cap, narr, xs, s.ptag, s.mask);
    NC(nccl().GroupStart());
    if (left) { if (nL) NC(nccl().Send(sendL, (size_t)(narr + 2) * nL, ncclDouble, cm.rank - 1, cm.comm, c->stream)); if (rL) NC(nccl().Recv(recvL, (size_t)(narr + 2) * rL, ncclDouble, cm.rank - 1, cm.comm, c->stream)); }
    if (right) { if (nR) NC(nccl().Send(sendR, (size_t)(narr + 2) * nR, ncclDouble, cm.rank + 1, cm.comm, c->stream)); if (rR) NC(nccl().Recv(recvR, (size_t)(narr + 2) * rR, ncclDouble, cm.rank + 1, cm.comm, c->stream)); }
    NC(nccl().GroupEnd());
    if (nL + nR) { // close the holes left by the leavers with stayers from the tail
      int *holes = cm.mig_list + 2 * (size_t)cm.mig_cap, *fillers = cm.mig_list + 4 * (size_t)cm.mig_cap;
      const int ntail = nL + nR; // the last ntail slots [np_new, np) are vacated
      CU(cudaMemsetAsync(cm.mig_flag, 0, sizeof(int) * ntail, c->stream));
      k_mig_flag<<<nblocks(ntail, 128), 128, 0, c->stream>>>(cm.mig_list, cm.mig_cap, nL, nR, cm.mig_flag, np_new);
      k_mig_holes<<<nblocks(ntail, 128), 128, 0, c->stream>>>(cm.mig_list, cm.mig_cap, nL, nR, np_new, cm.mig_cnt, holes);
      k_mig_fillers<<<nblocks(ntail, 128), 128, 0, c->stream>>>(cm.mig_flag, np_new, ntail, cm.mig_cnt, fillers);
      k_mig_move<<<nblocks(ntail, 128), 128, 0, c->stream>>>(holes, fillers, cm.mig_cnt, c->d_flags, S->buf, S->cap, narr, s.ptag, s.mask); // no read-back: at most ntail moves
    }
    if (rL) k_mig_unpack<<<nblocks(rL, 128), 128, 0, c->stream>>>(recvL, rL, np_new, S->buf, S->cap, narr, xs, s.ptag, s.mask);
    if (rR) k_mig_unpack<<<nblocks(rR, 128), 128, 0, c->stream>>>(recvR, rR, np_new + rL, S->buf, S->cap, narr, xs, s.ptag, s.mask);
    s.np = np_new + rL + rR; S->gen++;
    c->launches[KML_STAGE_MIGRATE] += 6;
    if (check_launch("migration")) return 1;
  }
  return 0;
}

// ---- fixes ------------------------------------------------------------------------------------
static int read_scratch3(kml_ctx *c, double out[3]) {
  if (c->comm.nranks > 1) NC(nccl().AllReduce(c->d_scratch, c->d_scratch, 3, ncclDouble, ncclSum, c->comm.comm, c->stream));
  CU(cudaMemcpyAsync(c->h_pinned + 32, c->d_scratch, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int d = 0; d < 3; d++) out[d] = c->h_pinned[32 + d];
  return 0;
}

int kml_fix_velocity_nodes(kml_ctx *c, int solid, int groupbit, int set_mask, const double v[3], const double vprev[3], int which, double ftot[3]) {
  CU(cudaSetDevice(c->dev));
  if (resolve_dt(c)) return 1;
  StageTimer t(c, KML_STAGE_GRID);
  if (which == 0) CU(cudaMemsetAsync(c->d_scratch, 0, 3 * sizeof(double), c->stream));
  std::vector<Grid *> gs;
  if (solid == -1) gs = active_grids(c); else gs.push_back(c->grids[c->solids[solid]->d.grid]);
  for (Grid *G : gs) {
    if (grid_normalize_if_needed(c, G)) return 1;
    G->nvd_valid = false;
    k_fix_velocity_nodes<<<nblocks(G->g.nn, 256), 256, 0, c->stream>>>(G->g, groupbit, set_mask, v[0], v[1], v[2], vprev ? vprev[0] : 0, vprev ? vprev[1] : 0,
                                                                        vprev ? vprev[2] : 0, which, 1.0 / c->dt, c->d_scratch);
    c->launches[KML_STAGE_GRID]++;
  }
  if (check_launch("k_fix_velocity_nodes")) return 1;
  if (which == 0 && ftot) return read_scratch3(c, ftot);
  return 0;
}

int kml_fix_velocity_particles(kml_ctx *c, int solid, int groupbit, int set_mask, const double v[3], const double vprev[3], int which, double ftot[3]) {
  CU(cudaSetDevice(c->dev));
  if (resolve_dt(c)) return 1;
  StageTimer t(c, KML_STAGE_OTHER);
  if (c->c.is_CPDI) return fail("kml: fix velocity_particles on the device is not implemented for CPDI (particle domains are not moved)");
  if (which == 1) CU(cudaMemsetAsync(c->d_scratch, 0, 3 * sizeof(double), c->stream));
  const double *val = which == 0 ? vprev : v;
  for (size_t is = 0; is < c->solids.size(); is++) {
    if (solid != -1 && (int)is != solid) continue;
    Solid *S = c->solids[is];
    if (S->s.np == 0) continue;
    if (which == 1 && !c->c.is_TL && !S->moved) return fail("fix velocity_particles (after the step) called before advance_particles");
    if (which == 0 && !c->c.is_TL && S->moved) return fail("fix velocity_particles (before the step) called after advance_particles");
    k_fix_velocity_particles<<<nblocks(S->s.np, 256), 256, 0, c->stream>>>(S->s, groupbit, set_mask, val[0], val[1], val[2], which, c->c.is_TL, c->dt, c->d_scratch);
    c->launches[KML_STAGE_OTHER]++;
  }
  if (check_launch("k_fix_velocity_particles")) return 1;
  if (which == 1 && ftot) return read_scratch3(c, ftot);
  return 0;
}

int kml_fix_temperature_nodes(kml_ctx *c, int solid, int groupbit, double T, double Tprev, int which) {
  CU(cudaSetDevice(c->dev));
  StageTimer t(c, KML_STAGE_GRID);
  std::vector<Grid *> gs;
  if (solid == -1) gs = active_grids(c); else gs.push_back(c->grids[c->solids[solid]->d.grid]);
  for (Grid *G : gs) {
    if (grid_normalize_if_needed(c, G)) return 1;
    k_fix_temperature_nodes<<<nblocks(G->g.nn, 256), 256, 0, c->stream>>>(G->g, groupbit, T, Tprev, which);
    c->launches[KML_STAGE_GRID]++;
  }
  return check_launch("k_fix_temperature_nodes");
}

int kml_fix_temperature_particles(kml_ctx *c, int solid, int groupbit, double T) {
  CU(cudaSetDevice(c->dev));
  StageTimer t(c, KML_STAGE_OTHER);
  for (size_t is = 0; is < c->solids.size(); is++) {
    if (solid != -1 && (int)is != solid) continue;
    Solid *S = c->solids[is];
    if (S->s.np == 0) continue;
    k_fix_temperature_particles<<<nblocks(S->s.np, 256), 256, 0, c->stream>>>(S->s, groupbit, T);
    c->launches[KML_STAGE_OTHER]++;
  }
  return check_launch("k_fix_temperature_particles");
}

int kml_fix_body_force(kml_ctx *c, int solid, int groupbit, int set_mask, const double f[3], double ftot[3]) {
  CU(cudaSetDevice(c->dev));
  StageTimer t(c, KML_STAGE_GRID);
  CU(cudaMemsetAsync(c->d_scratch, 0, 3 * sizeof(double), c->stream));
  std::vector<Grid *> gs;
  if (solid == -1) gs = active_grids(c); else gs.push_back(c->grids[c->solids[solid]->d.grid]);
  for (Grid *G : gs) { k_fix_body_force<<<nblocks(G->g.nn, 256), 256, 0, c->stream>>>(G->g, groupbit, set_mask, f[0], f[1], f[2], c->d_scratch); c->launches[KML_STAGE_GRID]++; }
  if (check_launch("k_fix_body_force")) return 1;
  if (ftot) return read_scratch3(c, ftot);
  return 0;
}

int kml_fix_force_nodes(kml_ctx *c, int solid, int groupbit, int set_mask, const double f[3], double ftot[3]) {
  CU(cudaSetDevice(c->dev));
  // Decomposed runs divide by the GLOBAL node count of the group, i.e. they reproduce the undecomposed run.  (The reference divides by each
  // rank's local + ghost count without reducing it, src/fix_force_nodes.cpp:128-150, so its total depends on the MPI decomposition.)
  StageTimer t(c, KML_STAGE_GRID);
  CU(cudaMemsetAsync(c->d_scratch, 0, 3 * sizeof(double), c->stream));
  std::vector<Grid *> gs;
  if (solid == -1) gs = active_grids(c); else gs.push_back(c->grids[c->solids[solid]->d.grid]);
  int *cnt = (int *)(c->d_scratch + 16);
  for (Grid *G : gs) {
    CU(cudaMemsetAsync(cnt, 0, sizeof(int), c->stream));
    k_fix_force_count<<<nblocks(G->g.nn, 256), 256, 0, c->stream>>>(G->g, groupbit, cnt);
    if (c->comm.nranks > 1) NC(nccl().AllReduce(cnt, cnt, 1, ncclInt, ncclSum, c->comm.comm, c->stream));
    k_fix_force_apply<<<nblocks(G->g.nn, 256), 256, 0, c->stream>>>(G->g, groupbit, set_mask, f[0], f[1], f[2], cnt, c->d_scratch);
    c->launches[KML_STAGE_GRID] += 2;
  }
  if (check_launch("k_fix_force_nodes")) return 1;
  if (ftot) return read_scratch3(c, ftot);
  return 0;
}

static int contact(kml_ctx *c, int s1, int s2, int hertz, double mu, double ftot[3]) {
  CU(cudaSetDevice(c->dev));
  if (resolve_dt(c)) return 1;
  StageTimer t(c, KML_STAGE_CONTACT);
  Solid *A = c->solids[s1], *B = c->solids[s2];
  if (c->c.dimension == 1) return 0;
  ContactParams cp; cp.dim = c->c.dimension; cp.hertz = hertz; cp.axisymmetric = c->c.axisymmetric; cp.temp = c->c.temp; cp.mu = mu; cp.dt = c->dt;
  const kml_material &m1 = A->d.mat, &m2 = B->d.mat;
  cp.Estar = 1.0 / ((1 - m1.nu * m1.nu) / m1.E + (1 - m2.nu * m2.nu) / m2.E);
  cp.max_cellsize = std::max(c->grids[A->d.grid]->d.cellsize, c->grids[B->d.grid]->d.cellsize);
  cp.alpha = m1.kappa / (m1.kappa + m2.kappa); cp.invcp1 = m1.invcp; cp.invcp2 = m2.invcp;
  CU(cudaMemsetAsync(c->d_scratch, 0, 3 * sizeof(double), c->stream));
  // contact uses the current positions: for UL these are the step-start positions (x)
  SolidDev a = A->s, b = B->s;
  // large bodies: bin solid 2 by the reference's first screen and visit 3^dim bins per particle (kml_kernels.cuh k_contact_bins); KML_CONTACT=bins | pairs forces a path
  const char *mode = getenv("KML_CONTACT");
  bool bins = mode ? !strcmp(mode, "bins") : (double)a.np * (double)b.np > 5.0e7;
  ContactBins cb; long long nbins = 1;
  if (bins) {
    cb.inv = 1.0 / cp.max_cellsize; cb.dim3 = c->c.dimension == 3;
    for (int d = 0; d < 3; d++) {
      cb.lo[d] = c->c.boxlo[d];
      cb.n[d] = (d == 2 && !cb.dim3) ? 1 : std::max(1, (int)std::ceil((c->c.boxhi[d] - c->c.boxlo[d]) * cb.inv) + 1);
      nbins *= cb.n[d];
    }
    if (nbins > (1ll << 27) || b.np >= (1ll << 31)) bins = false; // too fine a box for a dense bin table: the all-pairs sweep stays correct
  }
  if (bins) {
    kml_ctx::ContactScratch &w = c->contact;
    if (nbins + 1 > w.nbins_cap) { cudaFree(w.start); CU(cudaMalloc(&w.start, sizeof(int) * (nbins + 1))); w.nbins_cap = nbins + 1; cudaFree(w.scan_tmp); w.scan_tmp = nullptr; w.scan_bytes = 0; }
    if (b.np > w.np_cap) { cudaFree(w.cell_of); cudaFree(w.rank); cudaFree(w.order); const long long cap = b.np + b.np / 8 + 1024;
      CU(cudaMalloc(&w.cell_of, sizeof(int) * cap)); CU(cudaMalloc(&w.rank, sizeof(int) * cap)); CU(cudaMalloc(&w.order, sizeof(int) * cap)); w.np_cap = cap; }
    if (!w.scan_tmp) { cub::DeviceScan::ExclusiveSum(nullptr, w.scan_bytes, w.start, w.start, (int)(nbins + 1), c->stream); CU(cudaMalloc(&w.scan_tmp, w.scan_bytes)); }
    CU(cudaMemsetAsync(w.start, 0, sizeof(int) * (nbins + 1), c->stream));
    k_contact_bin_count<<<nblocks(b.np, 256), 256, 0, c->stream>>>(b, cb, w.cell_of, w.rank, w.start);
    if (cub::DeviceScan::ExclusiveSum(w.scan_tmp, w.scan_bytes, w.start, w.start, (int)(nbins + 1), c->stream) != cudaSuccess) return fail("contact bins: scan failed");
    k_contact_bin_fill<<<nblocks(b.np, 256), 256, 0, c->stream>>>(b.np, w.cell_of, w.rank, w.start, w.order);
    k_contact_bins<<<nblocks(a.np, 128), 128, 0, c->stream>>>(a, b, cp, cb, w.start, w.order, c->d_scratch);
    c->launches[KML_STAGE_CONTACT] += 4;
  } else {
    k_contact<<<nblocks(a.np, 128), 128, 0, c->stream>>>(a, b, cp, c->d_scratch);
    c->launches[KML_STAGE_CONTACT]++;
  }
  if (check_launch("k_contact")) return 1;
  A->mbp_nonzero = B->mbp_nonzero = true;
  if (ftot) return read_scratch3(c, ftot);
  return 0;
}
int kml_fix_contact_hertz(kml_ctx *c, int s1, int s2, double ftot[3]) { return contact(c, s1, s2, 1, 0.0, ftot); }
int kml_fix_contact_min_penetration(kml_ctx *c, int s1, int s2, double mu, double ftot[3]) { return contact(c, s1, s2, 0, mu, ftot); }

static int energy(kml_ctx *c, int solid, int groupbit, int kinetic, double *out) {
  CU(cudaSetDevice(c->dev));
  CU(cudaMemsetAsync(c->d_scratch + 8, 0, sizeof(double), c->stream));
  for (size_t i = 0; i < c->solids.size(); i++) {
    if (solid != -1 && (int)i != solid) continue;
    Solid *S = c->solids[i];
    k_energy<<<nblocks(S->s.np, 256), 256, 0, c->stream>>>(S->s, groupbit, kinetic, c->d_scratch + 8);
    c->launches[KML_STAGE_OTHER]++;
  }
  if (check_launch("k_energy")) return 1;
  if (c->comm.nranks > 1) NC(nccl().AllReduce(c->d_scratch + 8, c->d_scratch + 8, 1, ncclDouble, ncclSum, c->comm.comm, c->stream));
  CU(cudaMemcpyAsync(c->h_pinned + 40, c->d_scratch + 8, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  *out = c->h_pinned[40]; return 0;
}
int kml_compute_kinetic_energy(kml_ctx *c, int solid, int groupbit, double *ek) { return energy(c, solid, groupbit, 1, ek); }
int kml_compute_strain_energy(kml_ctx *c, int solid, int groupbit, double *es) { return energy(c, solid, groupbit, 0, es); }

int kml_error_flags(kml_ctx *c, unsigned *flags) { // collective on a decomposed run: every rank sees the UNION of every rank's bits
  CU(cudaSetDevice(c->dev));
  if (c->comm.nranks > 1) {
    double *bits = c->d_scratch + 40;
    k_flag_bits<<<1, 32, 0, c->stream>>>(c->d_flags, bits, 0.0); // d_scratch has 64 doubles: bits[8] is inside it
    NC(nccl().AllReduce(bits, bits, 8, ncclDouble, ncclMax, c->comm.comm, c->stream));
    k_bits_flag<<<1, 1, 0, c->stream>>>(bits, c->d_flags);
  }
  CU(cudaMemcpyAsync(c->h_pinned + 48, c->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  memcpy(flags, c->h_pinned + 48, sizeof(unsigned)); return 0;
}

int kml_comm_unique_id(void *id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  { const std::string why = nccl().load(); if (!why.empty()) return fail("kml: " + why); }
  ncclUniqueId id; NC(nccl().GetUniqueId(&id)); memcpy(id128, &id, sizeof id); return 0;
}
int kml_comm_init(kml_ctx *c, const void *id128) {
  CU(cudaSetDevice(c->dev));
  if (c->c.nranks <= 1) return 0;
  if (c->c.is_TL) return fail("kml: the slab decomposition covers ULMPM (TL solids own private grids; replicate them instead)");
  { const std::string why = nccl().load(); if (!why.empty()) return fail("kml: " + why); }
  ncclUniqueId id; memcpy(&id, id128, sizeof id);
  NC(nccl().CommInitRank(&c->comm.comm, c->c.nranks, id, c->c.rank));
  c->comm.rank = c->c.rank; c->comm.nranks = c->c.nranks;
  CU(cudaMalloc(&c->comm.mig_cnt, 8 * sizeof(int)));
  CU(cudaMallocHost(&c->comm.h_cnt, 8 * sizeof(int)));
  return 0;
}

int kml_comm_sum(kml_ctx *c, double *vals, int n) { // set-up commands only: one blocking round trip
  if (c->comm.nranks <= 1 || n <= 0) return 0;
  if (n > 16) return fail("kml_comm_sum: at most 16 values");
  CU(cudaSetDevice(c->dev));
  double *d = c->d_scratch + 40;
  CU(cudaMemcpyAsync(d, vals, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  NC(nccl().AllReduce(d, d, n, ncclDouble, ncclSum, c->comm.comm, c->stream));
  CU(cudaMemcpyAsync(vals, d, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int kml_profile(kml_ctx *c, int enable) { c->profile = enable != 0; return 0; }
int kml_timer_start(kml_ctx *c) { CU(cudaSetDevice(c->dev)); CU(cudaEventRecord(c->evA, c->stream)); return 0; }
int kml_timer_stop(kml_ctx *c, double *ms) {
  CU(cudaSetDevice(c->dev)); CU(cudaEventRecord(c->evB, c->stream)); CU(cudaEventSynchronize(c->evB));
  float t = 0; CU(cudaEventElapsedTime(&t, c->evA, c->evB)); *ms = t; return 0;
}
int kml_stage_times(kml_ctx *c, double ms[KML_STAGE_COUNT], int64_t launches[KML_STAGE_COUNT], int reset) {
  if (!c->ev_pending.empty()) {
    CU(cudaSetDevice(c->dev)); CU(cudaStreamSynchronize(c->stream));
    for (auto &p : c->ev_pending) { float t = 0; cudaEventElapsedTime(&t, p.a, p.b); c->ms[p.stage] += t; c->ev_pool.push_back(p.a); c->ev_pool.push_back(p.b); }
    c->ev_pending.clear();
  }
  for (int i = 0; i < KML_STAGE_COUNT; i++) { ms[i] = c->ms[i]; launches[i] = c->launches[i]; }
  if (reset) { memset(c->ms, 0, sizeof c->ms); memset(c->launches, 0, sizeof c->launches); }
  return 0;
}
int kml_stage_host_times(kml_ctx *c, double ms[KML_STAGE_COUNT], int reset) {
  for (int i = 0; i < KML_STAGE_COUNT; i++) ms[i] = c->host_ms[i];
  if (reset) memset(c->host_ms, 0, sizeof c->host_ms);
  return 0;
}

} // extern "C"
