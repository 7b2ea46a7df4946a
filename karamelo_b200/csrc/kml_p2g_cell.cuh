// kml_p2g_cell.cuh - cell-centric particle-to-grid for ULMPM (3-D cubic B-splines).
//
// Re-binning: particles are counted per background cell (cell key = the stencil base node,
// src/ulmpm.cpp:206-208), an exclusive scan gives cell offsets and a fill pass writes the
// particle order.  The P2G kernel then walks whole cells: all particles of a cell share the
// same 4x4x4 node stencil, so their contributions are reduced in registers before a single
// update per node and cell leaves the SM.
#pragma once
#include "kml_kernels.cuh"

namespace kml {

struct CellLists {
  bool valid = false;
  long long ncells = 0, cap_np = 0;
  int nc[3] = {0, 0, 0};
  int *cell_of = nullptr;    // [np] cell index of each particle
  int *count = nullptr;      // [ncells + 1] particles per cell, then exclusive offsets
  int *cursor = nullptr;     // [ncells] fill cursors
  int *order = nullptr;      // [np] particle ids grouped by cell
  void *scan_tmp = nullptr; size_t scan_bytes = 0;
  int build(const SolidDev &s, const GridDev &g, cudaStream_t st, int *nlaunch);
  void release();
};

inline bool cell_p2g_supported(int, int) { return false; }
inline int CellLists::build(const SolidDev &, const GridDev &, cudaStream_t, int *nl) { *nl = 0; valid = false; return 0; }
inline void CellLists::release() {}
inline int cell_p2g_launch(const SolidDev &, const GridDev &, const CellLists &, int, cudaStream_t, int *nl) { *nl = 0; return 1; }

} // namespace kml
