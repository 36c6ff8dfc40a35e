// Staged history kernel, carter tracking, production build without the per-history trace (see kernel_entry.h for why this
// is its own translation unit).
#include "kernel_entry.h"
namespace abl {
TransportKernel history_kernel_carter(bool trace) {
  return trace ? history_kernel_traced(ABL_TRACK_CARTER) : history_kernel<ABL_TRACK_CARTER, false>;
}
}  // namespace abl
