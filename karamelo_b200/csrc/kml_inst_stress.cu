// instantiates k_stress for one (dimension, UL/TL); compiled with -DKML_DIM=.. -DKML_TL=..
#include "kml_launch.h"
namespace kml {
int KML_LAUNCHER(stress)(int shape, const SolidDev &s, const GridDev &g, const StepParams &sp, const StressParams &tp, const kml_material &mat,
                         cudaStream_t st) {
  KML_SWITCH_SHAPE(k_stress, <<<nblocks(s.np, 128), 128, 0, st>>>(s, g, sp, tp, mat))
  return (int)cudaGetLastError();
}
} // namespace kml
