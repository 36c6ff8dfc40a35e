// Staged history kernel, carter tracking, production build without the per-history trace (see kernel_entry.h for why this
// is its own translation unit).
#include "kernel_entry.h"
namespace abl {
HistoryKernel history_kernel_carter(bool trace, bool tle, bool fixed) {
  if (trace) return history_kernel_traced(ABL_TRACK_CARTER);
  if (fixed && !tle) return HK_THIS_UNIT_FIXED((history_kernel<ABL_TRACK_CARTER, false, false, HK_FIXED_NF, HK_FIXED_NP, HK_HIST>), HK_FIXED_NF, HK_FIXED_NP);
  if (fixed) return HK_THIS_UNIT_FIXED((history_kernel<ABL_TRACK_CARTER, false, true, HK_FIXED_NF, HK_FIXED_NP, HK_HIST>), HK_FIXED_NF, HK_FIXED_NP);
  if (tle) return HK_THIS_UNIT((history_kernel<ABL_TRACK_CARTER, false, true>));
  return HK_THIS_UNIT((history_kernel<ABL_TRACK_CARTER, false, false>));
}
}  // namespace abl
