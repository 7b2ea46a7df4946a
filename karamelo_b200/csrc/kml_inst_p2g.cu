// instantiates k_p2g for one (dimension, UL/TL); compiled with -DKML_DIM=.. -DKML_TL=..
#include "kml_launch.h"
namespace kml {
int KML_LAUNCHER(p2g)(int shape, const SolidDev &s, const GridDev &g, const StepParams &sp, int what, cudaStream_t st) {
  KML_SWITCH_SHAPE(k_p2g, <<<nblocks(s.np, 128), 128, 0, st>>>(s, g, sp, what))
  return (int)cudaGetLastError();
}
} // namespace kml
