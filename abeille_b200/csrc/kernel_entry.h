/* kernel_entry.h -- the history kernels, one translation unit per tracker.
 *
 * nvcc compiles each .cu as a whole program: the non-inlined device helpers (boundary search, track-length scorer,
 * fission banking ...) get ONE register allocation per translation unit, negotiated over all the kernels that call them.
 * With every instantiation in one file, adding the surface-tracking kernel changed the register allocation of the
 * delta-tracking kernel (96 -> 132 bytes of spills, 3 % slower on the bench workload) although not a line of its code had
 * changed.  So each tracker's kernels live in their own file and the host code gets them as function pointers.
 */
#pragma once
#include "events.cuh"

namespace abl {
typedef void (*TransportKernel)(const DevProblem, const RunArgs);
// a staged history kernel with the launch shape its translation unit was compiled for (HK_THREADS may differ per unit)
struct HistoryKernel {
  TransportKernel fn;
  int threads;           // threads per CTA (history threads + service warps)
  int hist_threads;      // threads that own a history
  unsigned fixed_bytes;  // shared memory before the per-history columns (HK_COLS_OFFSET)
  bool events;           // the event-queue kernel (events.cuh): hist_threads = the most slots a CTA can hold
  // a build with the column shape compiled in (history.cuh: NFC, NPC, SC): only valid for exactly this shape; 0 = any
  int fixed_nf, fixed_np, fixed_slots;
};
#define HK_THIS_UNIT(fn) HistoryKernel{fn, HK_THREADS, HK_HIST, HK_COLS_OFFSET, false, 0, 0, 0}
#define HK_THIS_UNIT_FIXED(fn, nf, np) HistoryKernel{fn, HK_THREADS, HK_HIST, HK_COLS_OFFSET, false, nf, np, HK_HIST}
#define EQ_THIS_UNIT(fn) HistoryKernel{fn, EQ_THREADS, EQ_MAX_SLOTS, EQ_COLS_OFFSET, true, 0, 0, 0}
// the nesting depth the fixed builds are compiled for: root lattice -> assembly lattice -> pin universe -> cell (C5G7)
#define HK_FIXED_NF 3
#define HK_FIXED_NP 4
// tle = the run scores track-length tallies this generation (a build without the scorer serves the others)
// fixed = the build with the column shape HK_FIXED_NF x HK_FIXED_NP x HK_HIST compiled in (untraced kernels only)
HistoryKernel history_kernel_delta(bool trace, bool tle, bool fixed = false);    // kernels_delta.cu
HistoryKernel history_kernel_carter(bool trace, bool tle, bool fixed = false);   // kernels_carter.cu
HistoryKernel history_kernel_surface(bool trace, bool tle, bool fixed = false);  // kernels_surface.cu
HistoryKernel history_kernel_traced(int tracking); // kernels_trace.cu
// the event-queue kernel (events.cuh), delta and carter tracking
HistoryKernel event_kernel_delta(bool trace, bool tle);   // kernels_events.cu
HistoryKernel event_kernel_carter(bool trace, bool tle);  // kernels_events_carter.cu
HistoryKernel event_kernel_traced(int tracking);          // kernels_events_trace.cu
TransportKernel lane_kernel(int tracking, int mode); // kernels_lane.cu: per-lane kernel of the noise modes (mode 1 | 2)
TransportKernel implicit_kernel(int mode);            // kernels_implicit.cu: implicit-leakage delta tracking, per-lane kernel (mode 0 | 1 | 2)
}  // namespace abl
