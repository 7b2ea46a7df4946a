/* Single-rank MPI stand-in used ONLY to compile the unmodified reference
 * sources (/root/reference/src) into oracle/_ref/karamelo_ref.
 *
 * TEST INFRASTRUCTURE - not part of the product.  MPI is not installed in
 * this image; with nprocs == 1 every Send/Recv loop in the reference is empty
 * (universe->sendnrecv is empty), Allreduce degenerates to a copy and Bcast to
 * a no-op, so the arithmetic of the reference is unchanged.
 */
#ifndef KML_ORACLE_SHIM_MPI_H
#define KML_ORACLE_SHIM_MPI_H

#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long MPI_Aint;
struct MPI_Status { int dummy; };

#define MPI_COMM_WORLD 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_SUCCESS 0

/* datatype handle = size in bytes (enough for a single-rank copy) */
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_FLOAT 1004 /* size 4, distinct handle */
#define MPI_DOUBLE 8
#define MPI_LONG_LONG 1008
#define MPI_INT64_T 2008
#define MPI_C_BOOL 2001
#define MPI_CXX_BOOL 3001

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_LOR 4
#define MPI_LAND 5

static inline int kml_shim_mpi_size(MPI_Datatype t) {
  if (t >= 100000) return t - 100000; /* derived struct type */
  return t % 1000;
}

static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int MPI_Comm_free(MPI_Comm *) { return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Abort(MPI_Comm, int code) { fflush(stdout); exit(code ? code : 1); return 0; }
static inline double MPI_Wtime() {
  using namespace std::chrono;
  return duration_cast<duration<double>>(steady_clock::now().time_since_epoch()).count();
}
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  if (s != r) memmove(r, s, (size_t)n * kml_shim_mpi_size(t));
  return 0;
}
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
  if (s != r) memmove(r, s, (size_t)n * kml_shim_mpi_size(t));
  return 0;
}
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) {
  fprintf(stderr, "mpi shim: MPI_Send called with a single rank\n"); abort(); return 0;
}
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) {
  fprintf(stderr, "mpi shim: MPI_Recv called with a single rank\n"); abort(); return 0;
}
static inline int MPI_Get_address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)(intptr_t)p; return 0; }
static inline int MPI_Type_create_struct(int n, const int *bl, const MPI_Aint *, const MPI_Datatype *ty, MPI_Datatype *out) {
  int sz = 0; for (int i = 0; i < n; i++) sz += bl[i] * kml_shim_mpi_size(ty[i]);
  *out = 100000 + sz; return 0;
}
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
static inline int MPI_Type_free(MPI_Datatype *) { return 0; }

#endif
