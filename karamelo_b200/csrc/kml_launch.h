// kml_launch.h - kernel launchers, one translation unit per (kernel family, dimension, UL/TL)
// so that the library builds in parallel (see karamelo_b200/Makefile).
#pragma once
#include "kml_kernels.cuh"
#include <algorithm>

namespace kml {
#define KML_DECL_LAUNCHERS(D, T)                                                                                                     \
  int launch_p2g_d##D##_tl##T(int shape, const SolidDev &s, const GridDev &g, const StepParams &sp, int what, cudaStream_t st);          \
  int launch_g2p_d##D##_tl##T(int shape, const SolidDev &s, const GridDev &g, const StepParams &sp, cudaStream_t st);                    \
  int launch_stress_d##D##_tl##T(int shape, const SolidDev &s, const GridDev &g, const StepParams &sp, const StressParams &tp,           \
                                 const kml_material &mat, cudaStream_t st);
KML_DECL_LAUNCHERS(1, 0) KML_DECL_LAUNCHERS(2, 0) KML_DECL_LAUNCHERS(3, 0)
KML_DECL_LAUNCHERS(1, 1) KML_DECL_LAUNCHERS(2, 1) KML_DECL_LAUNCHERS(3, 1)
#undef KML_DECL_LAUNCHERS

// at least one block: a solid may hold no particle on a rank of a decomposed run (every kernel checks its index against the count)
inline unsigned nblocks(long long n, int bs) { return (unsigned)((std::max<long long>(n, 1) + bs - 1) / bs); }
} // namespace kml

// expands to launch_<family>_d<KML_DIM>_tl<KML_TL>
#define KML_CAT5_(a, b, c, d, e) a##b##c##d##e
#define KML_CAT5(a, b, c, d, e) KML_CAT5_(a, b, c, d, e)
#define KML_LAUNCHER(family) KML_CAT5(launch_##family##_d, KML_DIM, _tl, KML_TL, )

#define KML_SWITCH_SHAPE(KERNEL, ...)                                                                      \
  switch (shape) {                                                                                         \
  case KML_SHAPE_LINEAR: KERNEL<KML_DIM, KML_SHAPE_LINEAR, (KML_TL != 0)> __VA_ARGS__; break;              \
  case KML_SHAPE_CUBIC_SPLINE: KERNEL<KML_DIM, KML_SHAPE_CUBIC_SPLINE, (KML_TL != 0)> __VA_ARGS__; break;   \
  case KML_SHAPE_QUADRATIC_SPLINE: KERNEL<KML_DIM, KML_SHAPE_QUADRATIC_SPLINE, (KML_TL != 0)> __VA_ARGS__; break; \
  default: KERNEL<KML_DIM, KML_SHAPE_BERNSTEIN, (KML_TL != 0)> __VA_ARGS__; break;                          \
  }
