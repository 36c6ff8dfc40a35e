// Staged history kernels with the per-history trace (flight / collision counts, event hash, final RNG state: what the
// parity tests compare with the oracle), all trackers (see kernel_entry.h for why this is its own translation unit).
#include "kernel_entry.h"
namespace abl {
HistoryKernel history_kernel_traced(int tracking) {
  switch (tracking) {
    case ABL_TRACK_SURFACE: return HK_THIS_UNIT((history_kernel<ABL_TRACK_SURFACE, true, true>));
    case ABL_TRACK_DELTA: return HK_THIS_UNIT((history_kernel<ABL_TRACK_DELTA, true, true>));
    default: return HK_THIS_UNIT((history_kernel<ABL_TRACK_CARTER, true, true>));
  }
}
}  // namespace abl
