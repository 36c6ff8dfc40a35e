// Per-lane history kernel: MODE 0 = k-eigenvalue transport for fixed-source problems (fission neutrons as secondaries); the noise modes: MODE 1 = power-iteration generation that may sample the noise source,
// MODE 2 = noise particles (transport.cuh, noise.cuh; see kernel_entry.h for why this is its own translation unit).
#define ABL_TABLES_GLOBAL 1  // this unit's kernels read the tables from global memory (detmath.cuh: ldt)
#include "kernel_entry.h"
namespace abl {
template <int MODE>
static TransportKernel lane_kernel_mode(int tracking) {
  switch (tracking) {
    case ABL_TRACK_SURFACE: return transport_kernel<ABL_TRACK_SURFACE, MODE>;
    case ABL_TRACK_DELTA: return transport_kernel<ABL_TRACK_DELTA, MODE>;
    default: return transport_kernel<ABL_TRACK_CARTER, MODE>;
  }
}
TransportKernel lane_kernel(int tracking, int mode) {
  return mode == 2 ? lane_kernel_mode<2>(tracking) : (mode == 1 ? lane_kernel_mode<1>(tracking) : lane_kernel_mode<0>(tracking));
}
}  // namespace abl
