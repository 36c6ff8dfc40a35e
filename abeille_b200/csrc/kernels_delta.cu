// Staged history kernel, delta tracking, production build without the per-history trace (see kernel_entry.h for why this
// is its own translation unit).
#include "kernel_entry.h"
namespace abl {
HistoryKernel history_kernel_delta(bool trace, bool tle) {
  if (trace) return history_kernel_traced(ABL_TRACK_DELTA);
  if (tle) return HK_THIS_UNIT((history_kernel<ABL_TRACK_DELTA, false, true>));
  return HK_THIS_UNIT((history_kernel<ABL_TRACK_DELTA, false, false>));
}
}  // namespace abl
