// Implicit-leakage delta tracking (src/implicit_leakage_delta_tracker.cpp): per-lane history loop (transport.cuh) in all
// three modes -- 0 = k-eigenvalue generation, 1 = power-iteration generation of a noise run, 2 = noise particles.  Its own
// translation unit for the reason kernel_entry.h gives.
#define ABL_TABLES_GLOBAL 1  // this unit's kernels read the tables from global memory (detmath.cuh: ldt)
#include "kernel_entry.h"
namespace abl {
TransportKernel implicit_kernel(int mode) {
  switch (mode) {
    case 0: return transport_kernel<ABL_TRACK_IMPLICIT_LEAKAGE, 0>;
    case 1: return transport_kernel<ABL_TRACK_IMPLICIT_LEAKAGE, 1>;
    default: return transport_kernel<ABL_TRACK_IMPLICIT_LEAKAGE, 2>;
  }
}
}  // namespace abl
