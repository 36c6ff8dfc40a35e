// Staged history kernel, surface tracking, production build without the per-history trace (see kernel_entry.h for why this
// is its own translation unit).
#include "kernel_entry.h"
namespace abl {
TransportKernel history_kernel_surface(bool trace) {
  return trace ? history_kernel_traced(ABL_TRACK_SURFACE) : history_kernel<ABL_TRACK_SURFACE, false>;
}
}  // namespace abl
