// Staged history kernel, delta tracking, production build without the per-history trace (see kernel_entry.h for why this
// is its own translation unit).
#include "kernel_entry.h"
namespace abl {
TransportKernel history_kernel_delta(bool trace) {
  return trace ? history_kernel_traced(ABL_TRACK_DELTA) : history_kernel<ABL_TRACK_DELTA, false>;
}
}  // namespace abl
