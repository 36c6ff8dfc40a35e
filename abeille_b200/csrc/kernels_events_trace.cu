// Event-queue history kernels with the per-history trace (flight / collision counts, event hash, final RNG state: what the
// parity tests compare with the oracle); see kernel_entry.h for why this is its own translation unit.
#include "kernel_entry.h"
namespace abl {
HistoryKernel event_kernel_traced(int tracking) {
  if (tracking == ABL_TRACK_DELTA) return EQ_THIS_UNIT((event_kernel<ABL_TRACK_DELTA, true, true>));
  return EQ_THIS_UNIT((event_kernel<ABL_TRACK_CARTER, true, true>));
}
}  // namespace abl
