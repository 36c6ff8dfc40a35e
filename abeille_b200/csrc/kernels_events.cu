// Event-queue history kernel, delta tracking, production build without the per-history trace (see kernel_entry.h for why
// this is its own translation unit).
#include "kernel_entry.h"
namespace abl {
HistoryKernel event_kernel_delta(bool trace, bool tle) {
  if (trace) return event_kernel_traced(ABL_TRACK_DELTA);
  if (tle) return EQ_THIS_UNIT((event_kernel<ABL_TRACK_DELTA, false, true>));
  return EQ_THIS_UNIT((event_kernel<ABL_TRACK_DELTA, false, false>));
}
}  // namespace abl
