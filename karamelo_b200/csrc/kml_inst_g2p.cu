// instantiates k_g2p for one (dimension, UL/TL); compiled with -DKML_DIM=.. -DKML_TL=..
#include "kml_launch.h"
namespace kml {
int KML_LAUNCHER(g2p)(int shape, const SolidDev &s, const GridDev &g, const StepParams &sp, cudaStream_t st) {
  KML_SWITCH_SHAPE(k_g2p, <<<nblocks(s.np, 128), 128, 0, st>>>(s, g, sp))
  return (int)cudaGetLastError();
}
} // namespace kml
