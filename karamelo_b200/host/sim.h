// sim.h - host driver state: the object graph the Karamelo script commands build
// (reference src/mpm.h:41-72: Update, Domain, Material, Group, Modify, Output) reduced to
// what the MPM time-step path needs.  Setup-time arithmetic that decides integer counts
// (node counts, particle lattice, group masks) is restated from the reference with
// file:line citations; the per-step work is delegated to the C ABI in include/kml.h.
#pragma once
#include <istream>
#include <ostream>
#include "../../include/kml.h"
#include "input.h"
#include <array>
#include <cstdint>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

namespace kmlh {

// ---- regions (reference src/region_block.cpp, src/region_cylinder.cpp, src/region_sphere.cpp) ----
struct Region {
  std::string id, style;
  int interior = 1;
  double lim[6] = {0, 0, 0, 0, 0, 0};
  virtual ~Region() {}
  virtual int inside(double x, double y, double z) const = 0;
  int match(double x, double y, double z) const { return interior ? inside(x, y, z) : !inside(x, y, z); } // src/region.cpp:64-72
  virtual void write_restart(std::ostream &) const {} // Region::write_restart of the style (src/region_block.cpp:170-177, ...)
  virtual bool to_kml(kml_region &) const { return false; } // the predicate for the device kernels (block, cylinder, sphere)
};

struct MaterialH {
  std::string id;
  kml_material km;
  int ieos = -1, istrength = -1, idamage = -1, itemperature = -1; // indices into the EOS / strength / ... lists (restart files)
};
struct EOSH { std::string id; int type; double rho0, K, c0, S, Gamma, cv, Tr, Q1, Q2; };
struct StrengthH { std::string id; int type; double G, A, B, n, epsdot0, C, m, Tr, Tm; };
struct DamageH { std::string id; int type; double d1, d2, d3, d4, d5, epsdot0, Tr, Tm; };
struct TemperatureH { std::string id; int type; double chi, cp, kappa, alpha, T0, Tm; };

struct GridH {
  int id = -1;                 // device grid id
  double cellsize = 0;
  kml_grid_desc desc{};
  int nx_global = 0, ny_global = 0, nz_global = 0;
  int64_t nnodes = 0;
  std::vector<int> mask;       // host copy of node group masks
};

struct SolidH {
  std::string id;
  int dev = -1;                // device solid id
  int mat = -1;
  GridH *grid = nullptr;       // UL: the domain grid; TL: own grid
  std::unique_ptr<GridH> own_grid;
  int64_t np = 0;
  int64_t np_created = 0;      // Solid::np: set by populate, unchanged by delete_particles (written to restart files)
  int np_per_cell = 0;
  double solidlo[3], solidhi[3];
  double T0 = 0;
  double vtot = 0, mtot = 0;
  // Host mirrors of the particle attributes the script side reads per particle (reference positions for regions / expressions, group
  // masks, tags).  The engine may change the particle set behind the host's back (migration between slabs, uploads through the C ABI,
  // populate on the device): Sim::sync compares kml_solid_generation before every use and downloads again when it moved.
  std::vector<std::array<double, 3>> x0;
  std::vector<int> mask;
  std::vector<int64_t> ptag;
  uint64_t mirror_gen = 0; // 0 = mirrors never filled
};

enum HookMask { INITIAL_INTEGRATE = 1, POST_PARTICLES_TO_GRID = 2, POST_UPDATE_GRID_STATE = 4, POST_GRID_TO_POINT = 8,
                POST_ADVANCE_PARTICLES = 16, POST_VELOCITIES_TO_GRID = 32, FINAL_INTEGRATE = 64 }; // src/fix.h:55-63

class Sim;
struct Fix {
  std::string id, style;
  int igroup = 0, groupbit = 1, mask = 0;
  virtual ~Fix() {}
  virtual void initial_integrate(Sim &) {}
  virtual void post_particles_to_grid(Sim &) {}
  virtual void post_update_grid_state(Sim &) {}
  virtual void post_advance_particles(Sim &) {}
  virtual void post_velocities_to_grid(Sim &) {}
  virtual void final_integrate(Sim &) {}
  virtual void write_restart(std::ostream &) const {} // Fix::write_restart of the style (src/fix_velocity_nodes.cpp:270-292, ...)
  // does one of this fix's hooks evaluate script expressions (time, dt, ...) during the NEXT step?  The one-shot fixes answer no after step 1,
  // which lets the scheme leave the new dt on the device until the engine itself needs it (Sim::run).
  virtual bool needs_time(const Sim &) const { return true; }
};
struct Compute {
  std::string id, style;
  int igroup = 0, groupbit = 1;
  virtual ~Compute() {}
  virtual void compute_value(Sim &) = 0;
};

struct Dump { std::string id, style, filename; int igroup; int every; std::vector<std::string> fields; int64_t next = 0; };

class Sim {
public:
  Sim();
  ~Sim();
  Input input;

  // ---- Update (src/update.h) ----
  double dt = 1e-16, dt_factor = 0.9; bool dt_constant = false;       // src/update.cpp:38-40
  bool dt_stale = false; // the engine holds a newer dt than `dt` (deferred adjust_dt, see Sim::run)
  int64_t ntimestep = 0, atimestep = 0, firststep = 0, laststep = 0; double atime = 0, maxtime = -1; int64_t nsteps = 0;
  std::string method_type, scheme_style = "musl";                    // default scheme MUSL, src/update.cpp:42-45
  bool method_set = false, is_TL = false, is_CPDI = false, temp = false, ge = false; int cpdi_style = 0;
  int shape_function = KML_SHAPE_LINEAR, sub_method = KML_SUB_FLIP; double PIC_FLIP = 0.99, PIC_FLIP_script = 0.99;

  // ---- Domain ----
  int dimension = 0; double boxlo[3] = {0, 0, 0}, boxhi[3] = {0, 0, 0}, sublo[3] = {0, 0, 0}, subhi[3] = {0, 0, 0};
  bool axisymmetric = false, created = false; int64_t np_total = 0;
  std::vector<std::unique_ptr<Region>> regions;
  std::vector<std::unique_ptr<SolidH>> solids;
  std::unique_ptr<GridH> grid; // UL background grid

  // ---- Material ----
  std::vector<EOSH> eoss; std::vector<StrengthH> strengths; std::vector<DamageH> damages; std::vector<TemperatureH> temperatures;
  std::vector<MaterialH> materials;

  // ---- Group (src/group.cpp:28-48) ----
  static const int MAX_GROUP = 32;
  std::string gnames[MAX_GROUP], gpon[MAX_GROUP]; int gbitmask[MAX_GROUP], gsolid[MAX_GROUP], gregion[MAX_GROUP]; int ngroup = 1;

  // ---- Modify ----
  std::vector<std::unique_ptr<Fix>> fixes; std::vector<std::unique_ptr<Compute>> computes;

  // ---- Output ----
  int every_log = 1 /* src/output.cpp:48 */; std::vector<std::string> log_fields{"step", "dt", "time"}; int64_t next_log = 0;
  std::vector<Dump> dumps; std::ofstream logfile; bool quiet = false;
  int restart_every = 0; int64_t next_restart = 0; std::string restart_name; // Output::create_restart, src/output.cpp:323-350
  bool restarted_TL = false;                                                 // read_restart of a total-Lagrangian run (see Sim::read_restart)
  std::vector<std::string> additional_args;                                  // Update::additional_args (CPDI style), src/update.cpp:189-194
  void write_restart(const std::string &pattern);                            // WriteRestart::write, src/write_restart.cpp:51-88
  void read_restart(const std::string &pattern);                             // ReadRestart::command, src/read_restart.cpp:36-83
  std::unique_ptr<Fix> fix_from_restart(const std::string &id, const std::string &style, int igroup, std::istream &is); // Modify::read_restart

  // ---- device ----
  kml_ctx *ctx = nullptr; int device = 0;
  // ---- slab decomposition over several GPUs (one process per GPU; replaces Universe/Domain sub-boxes) ----
  int rank = 0, nranks = 1; std::vector<unsigned char> nccl_id; bool grid_pending = false;
  int64_t np_global_last = 0, tag_offset_last = 0; // of the solid created last (for tests / reports)
  void create_device_grid(GridH &g, int base_lo, int base_hi);

  // helpers
  int find_region(const std::string &n) const; int find_solid(const std::string &n) const; int find_material(const std::string &n) const;
  int find_group(const std::string &n) const;
  void check(int rc) const; // kml_* return code -> fatal(kml_last_error())
  void sync(SolidH &S);            // np + host mirrors (x0, mask, ptag) follow the engine's particle set
  void mirrors_current(SolidH &S); // the host just wrote the same change to the mirrors and to the device
  void ensure_ctx();
  void init_grid(GridH &g, const double *lo, const double *hi); // Grid::init, src/grid.cpp:68-264
  void run(Var condition);        // Scheme::run, src/usl.cpp / src/musl.cpp / src/usf.cpp
  void output_setup(); void output_write(int64_t step);
  void hooks(int which);

private:
  void register_commands();
  Var cmd_method(std::vector<std::string> &a); Var cmd_scheme(std::vector<std::string> &a); Var cmd_dimension(std::vector<std::string> &a);
  Var cmd_region(std::vector<std::string> &a); Var cmd_eos(std::vector<std::string> &a); Var cmd_strength(std::vector<std::string> &a);
  Var cmd_damage(std::vector<std::string> &a); Var cmd_temperature(std::vector<std::string> &a); Var cmd_material(std::vector<std::string> &a);
  Var cmd_solid(std::vector<std::string> &a); Var cmd_group(std::vector<std::string> &a); Var cmd_fix(std::vector<std::string> &a);
  Var cmd_compute(std::vector<std::string> &a); Var cmd_dump(std::vector<std::string> &a); Var cmd_run(std::vector<std::string> &a, int kind);
  void populate(SolidH &s, std::vector<std::string> &a); // Solid::populate, src/solid.cpp:1810-2336
};

} // namespace kmlh
