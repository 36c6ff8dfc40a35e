/* events.cuh -- the event-queue history kernel (delta and carter tracking, k-eigenvalue mode).
 *
 * Same arithmetic, same RNG consumption and same per-history outcomes as the staged kernel (history.cuh) and the per-lane
 * kernel (transport.cuh), which follow DeltaTracker::transport (src/delta_tracker.cpp:72-263) and CarterTracker::transport
 * (src/carter_tracker.cpp:92-294) of the reference.  What changes is WHO executes an event of a history.
 *
 * In the staged kernel a thread owns a history, and a warp walks through the stages of the loop with whatever lanes happen
 * to be in that stage (ncu: 17 of 32 lanes per instruction -- 45 % of the flights need a re-descent through the universe
 * tree, 83 % end in a real collision, and the lanes that do not wait).  Here the state of a history lives in a SLOT of the
 * CTA's shared memory (the same columns as history.cuh), a CTA holds more slots than it has threads, and the slots travel
 * through per-event queues:
 *
 *     REFILL  a finished slot takes the next bank index, loads the particle             -> LOCATE (full lookup)
 *     STEP    the flight arrives: real or virtual collision; Transporter::collision (scores, tally, fission,
 *             roulette, scatter); then the next flight: sample the distance, advance the cursor, re-validate
 *             its pads                                                                  -> STEP | LOCATE | REFILL
 *     LOCATE  (re-)descent through the universe tree; two queues: histories that only change cell inside their
 *             universe, and histories that go through the lattices                      -> STEP | BOUNDARY
 *     BOUNDARY a particle left the geometry: boundary search, leak or reflection        -> LOCATE | REFILL
 *     FISSION jobs (not slots): n_new sites of a collision are sampled and appended to the scratch bank
 *
 * 55 % of the flights go STEP -> STEP (the pads still hold after the move), 45 % STEP -> LOCATE -> STEP.
 *
 * A warp repeatedly picks the event queue that fills most of its lanes, takes one slot per lane, runs that event for all of
 * them, and pushes every slot to the queue of its next event.  No thread is bound to a history, no warp to a stage, and
 * there is no block-wide barrier in the loop.
 *
 * Queues are per LANE: slot s = 32 k + l is only ever handled by lane l (of any warp), so a warp's accesses to the slot
 * columns are conflict-free like the staged kernel's (lane l touches word l of every column row), and a queue is 32
 * independent small rings.  Push / pop are one shared-memory atomic on the lane's own counter each.
 */
#pragma once
#include "history.cuh"

namespace abl {

#ifndef EQ_THREADS
#define EQ_THREADS 512
#endif
#ifndef EQ_FQ
#define EQ_FQ 128  // fission-job ring entries (a full ring makes the collision bank its sites inline)
#endif
#ifndef EQ_STICKY
#define EQ_STICKY 24  // lanes the CTA's current event must still fill for a warp to stay with it
#endif
// a warp that finds fewer than EQ_MINFILL lanes of work naps EQ_NAP ns (at most EQ_PATIENCE times in a row) while the running
// events push more, instead of running a thin batch
#ifndef EQ_MINFILL
#define EQ_MINFILL 0
#endif
#ifndef EQ_PATIENCE
#define EQ_PATIENCE 2
#endif
#ifndef EQ_NAP
#define EQ_NAP 300
#endif
#ifndef EQ_REFILL_AT
#define EQ_REFILL_AT 33  // (33 = never: measured 132-166 ms at 16 / 8 / 4 against 115 ms -- thin REFILL and BOUNDARY batches cost more than the idle slots)
#endif
#define EQ_RING 32  // ring entries per (queue, lane) = the most slots a lane can own
#define EQ_MAX_SLOTS (32 * EQ_RING)
#define EQ_EMPTY 0xffu

// queues in the order the dispatcher breaks ties
enum { Q_STEP = 0, Q_LOC_TREE, Q_LOC_CELL, Q_BOUNDARY, Q_REFILL, Q_N, Q_FISSION = Q_N };
// what a LOCATE event is part of (history.cuh: PH_*)
enum { EP_FLIGHT = 0, EP_BIRTH, EP_REFLECTED, EP_RESURRECT };
// the slot's pending event.  LOCATE: need (first bad pad, 0 = full lookup) | EP_* << 8; STEP: 0 = the flight arrives at its
// tentative collision site, EV_MOVE_ONLY = a particle that starts a flight (birth, reflection, resurrection)
#define HI_EVT HI_BDONE
#define EV_MOVE_ONLY 0x10000
enum { X_COLLIDE = 100, X_MOVE = 101 };  // inside a STEP: what the slot does next (not queues)
enum { ST_TASKS = 0, ST_LANES = Q_N + 1, ST_IDLE = 2 * (Q_N + 1), ST_N };  // dispatcher statistics (RunArgs::eq_stats)

struct EQFixed {  // the fixed part of the dynamic shared memory; staged tables and the slot columns follow it
  FisJob fq[EQ_FQ];
  volatile unsigned fq_seq[EQ_FQ];  // cell i: == pos -> free for the job with sequence number pos; == pos + 1 -> holds it
  unsigned fq_head, fq_tail;
  unsigned char ring[Q_N][EQ_RING][32];  // k of slot 32 k + lane, EQ_EMPTY = not (yet) written
  int cnt[Q_N][32];                      // entries pushed and not yet claimed (may dip below 0 for a moment)
  unsigned head[Q_N][32], tail[Q_N][32];
  int live;  // slots that may still receive a history
  volatile int cur;  // the event most warps are running (see the dispatcher)
  volatile int abort;
  unsigned long long deadline_ns;
  unsigned rare[RC_N];
  double leak, leak_mig;
  double sd[EQ_THREADS / 32][3];
  unsigned long long wcnt[EQ_THREADS / 32][WC_N];
  unsigned long long stat[ST_N];
};
#define EQS (*reinterpret_cast<EQFixed*>(hk_shared_raw))
#define EQ_COLS_OFFSET ((unsigned)((sizeof(EQFixed) + 15) & ~size_t(15)))

// Orders a slot's column stores before the ring entry that publishes it (and the reader's column loads after the entry it
// took).  EQ_FENCE_SOFT makes it a compiler barrier only (an experiment: shared-memory accesses of a warp are performed in
// order by the hardware, but the PTX memory model does not promise it).
#ifdef EQ_FENCE_SOFT
#define EQ_FENCE() asm volatile("" ::: "memory")
#else
#define EQ_FENCE() __threadfence_block()
#endif

// ---- slot queues ---------------------------------------------------------------------------------------------------------
// (every store to the slot's columns happens before the push: the fence orders them before the ring entry becomes visible)
__device__ __forceinline__ void eq_push(EQFixed& S, int qi, int lane, int k) {
  const unsigned pos = atomicAdd(&S.tail[qi][lane], 1u) & (EQ_RING - 1);
  volatile unsigned char* cell = &S.ring[qi][pos][lane];
  while (*cell != EQ_EMPTY && !S.abort) {}  // (a lane owns at most EQ_RING slots, so the cell is free unless a reader is mid-way)
  EQ_FENCE();
  *cell = (unsigned char)k;
  atomicAdd(&S.cnt[qi][lane], 1);
}
__device__ __forceinline__ int eq_pop(EQFixed& S, int qi, int lane) {
  int k = -1;
  if (*(volatile int*)&S.cnt[qi][lane] > 0) {
    if (atomicSub(&S.cnt[qi][lane], 1) > 0) {
      const unsigned pos = atomicAdd(&S.head[qi][lane], 1u) & (EQ_RING - 1);
      volatile unsigned char* cell = &S.ring[qi][pos][lane];
      unsigned v;
      while ((v = *cell) == EQ_EMPTY && !S.abort) {}
      *cell = EQ_EMPTY;
      if (v != EQ_EMPTY) k = (int)v;
    } else {
      atomicAdd(&S.cnt[qi][lane], 1);
    }
  }
  return k;
}

// ---- fission jobs: a bounded multi-producer / multi-consumer ring with per-cell sequence numbers -------------------------------
__device__ __forceinline__ bool eq_try_post(EQFixed& S, const FisJob& j) {
  unsigned pos;
  for (;;) {
    pos = *(volatile unsigned*)&S.fq_head;
    const unsigned seq = S.fq_seq[pos % EQ_FQ];
    if (seq == pos) {
      if (atomicCAS(&S.fq_head, pos, pos + 1) == pos) break;
    } else if ((int)(seq - pos) < 0) {
      return false;  // the ring is full
    }
    if (S.abort) return true;
  }
  S.fq[pos % EQ_FQ] = j;
  __threadfence_block();
  S.fq_seq[pos % EQ_FQ] = pos + 1;
  return true;
}
struct EventPost {
  static __device__ __forceinline__ void post(const DevProblem& P, const RunArgs& A, const FisJob& j, int) {
    if (eq_try_post(EQS, j)) return;
    uint64_t rng = j.rng;  // ring full: bank the sites here
    const FissionTables ft{P.chi_cp, P.dg_cp, P.dg_off, P.gmid, P.G};
    bank_fission_sites<HK_MATH>(ft, A.sites, A.n_sites, A.site_capacity, rng, V3{j.x, j.y, j.z}, V3{j.ux, j.uy, j.uz}, j.w, j.parent,
                                j.daughter0, j.n_new, j.mg / P.G, j.mg, ldt(&P.nud[j.mg]) / ldt(&P.nu[j.mg]));
  }
  static __device__ __forceinline__ void sites(unsigned n) { atomicAdd(&EQS.rare[RC_SITES], n); }
};

// what one event of one slot adds to the warp's counters (summed over the lanes at the end of the event)
struct EvCount {
  unsigned real, virt, coll_scores, tl_bins;
};

// Particle::split (particle.hpp:165-173) after a flight that ended in a collision or a reflection (carter tracking)
template <bool TRACE>
__device__ __forceinline__ void eq_carter_split(const DevProblem& P, const RunArgs& A, const Cols& q, uint32_t gslot, uint32_t nslots) {
  const double w = HK_D(q, HD_W);
  if (fabs(w) >= P.wgt_split) {
    const int n_new = (int)ceil(fabs(w));
    if (n_new > 1) {
      Hist h;
      h.r = hk_ld3(q, HD_R);
      h.nsec = HK_I(q, HI_NSEC);
      h.w = w / (double)n_new;
      h.w2 = 0. / (double)n_new;
      HK_D(q, HD_W) = h.w;
      const V3 u = hk_ld3(q, HD_U);
      const double E = HK_D(q, HD_E);
      for (int np = 0; np < n_new - 1; np++)
        if (!push_secondary(A, h, u, E, h.w, h.w2, gslot, nslots)) {
          raise_error(A, ABL_ERR_BANK_OVERFLOW, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
          break;
        }
      HK_I(q, HI_NSEC) = h.nsec;
    }
  }
}

// the end of a particle: the next secondary of the history (Particle::resurect + Tracker restart, delta_tracker.cpp:197-229)
// or the end of the history.  Returns the slot's next queue.
template <int TRK, bool TRACE>
__device__ __forceinline__ int eq_particle_dead(const DevProblem& P, const RunArgs& A, const Cols& q, uint32_t gslot, uint32_t nslots) {
  if (TRK == ABL_TRACK_CARTER && HK_I(q, HI_NSEC) > 0) {
    Hist h;
    h.nsec = HK_I(q, HI_NSEC);
    pop_secondary(P, A, h, gslot, nslots);
    HK_I(q, HI_NSEC) = h.nsec;
    hk_st3(q, HD_R, h.r);
    hk_st3(q, HD_U, h.u);
    HK_D(q, HD_E) = h.E;
    HK_D(q, HD_W) = h.w;
    HK_I(q, HI_G) = h.g | (h.emid ? 0x100 : 0);
    HK_I(q, HI_TOK) = 0;
    HK_FR(q, 0, 0) = h.r.x;
    HK_FR(q, 1, 0) = h.r.y;
    HK_FR(q, 2, 0) = h.r.z;
    HK_I(q, HI_EVT) = 0 | (EP_RESURRECT << 8);
    return Q_LOC_TREE;
  }
  const uint32_t idx = (uint32_t)HK_I(q, HI_IDX);
  A.nfis[idx] = (uint32_t)HK_I(q, HI_DAU);  // sites produced = daughters numbered
  if (TRACE) {
    A.tr_flights[idx] = (uint32_t)HK_I(q, HI_NFL);
    A.tr_real[idx] = (uint32_t)HK_I(q, HI_NRE);
    A.tr_virtual[idx] = (uint32_t)HK_I(q, HI_NVI);
    A.tr_hash[idx] = HK_U(q, HD_PT + A.hk_np);
    A.tr_rng[idx] = HK_U(q, HD_RNG);
  }
  return Q_REFILL;
}

// The flight arrives at its tentative collision site inside cell ccell (delta_tracker.cpp:133,167-195,
// carter_tracker.cpp:189-207): track-length tallies from the pre-move position, then real or virtual collision.
template <int TRK, bool TRACE>
__device__ __forceinline__ int eq_arrive(const DevProblem& P, const RunArgs& A, const Cols& q, int ccell, bool tle, EvCount& ec,
                                         uint32_t gslot, uint32_t nslots) {
  const double d_coll = HK_D(q, HD_DC);
  const int g = HK_I(q, HI_G) & 0xff;
  if (tle) ec.tl_bins += score_flight_cols(tle_args(P), q, HK_I(q, HI_HMAT) * P.G + g, d_coll);
  V3 r = hk_ld3(q, HD_R);
  {
    const V3 u = hk_ld3(q, HD_U);
    r.x = r.x + d_coll * u.x;
    r.y = r.y + d_coll * u.y;
    r.z = r.z + d_coll * u.z;
  }
  hk_st3(q, HD_R, r);
  const int hmat = HK_I(q, HI_MAT);
  HK_I(q, HI_HMAT) = hmat;
  bool had_collision = false, alive = true;
  const double Esample = ldt(&P.smp[g]);
  const double Et = ldt(&P.Et[hmat * P.G + g]);
  const double real_frac = ldt(&P.real_frac[hmat * P.G + g]);  // Et / Esample
  uint64_t rng = HK_U(q, HD_RNG);
  if (TRK == ABL_TRACK_DELTA) {
    if (Et - Esample > 1.E-10) {
      raise_error(A, ABL_ERR_MAJORANT, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
      alive = false;
    } else if (HK_MATH::rand(rng) < real_frac) {
      had_collision = true;
    }
  } else {
    if (Esample >= Et) {
      if (HK_MATH::rand(rng) < real_frac) had_collision = true;
    } else {  // under-estimated sampling xs: signed-weight branch (carter_tracker.cpp:192-207)
      const double D = HK_MATH::div(Et, 2. * Et - Esample);
      const double F = HK_MATH::div(Et, D * Esample);
      const double w = HK_D(q, HD_W);
      if ((D - HK_MATH::rand(rng)) > 0.) {
        HK_D(q, HD_W) = w * F;
        had_collision = true;
      } else {
        HK_D(q, HD_W) = -w * F;
      }
    }
  }
  HK_U(q, HD_RNG) = rng;
  if (!alive) {
    if (TRK == ABL_TRACK_CARTER) HK_I(q, HI_NSEC) = 0;
    return eq_particle_dead<TRK, TRACE>(P, A, q, gslot, nslots);
  }
  note_col<TRACE>(q, A.hk_np, (had_collision ? 0x2000000000000000ULL : 0x1000000000000000ULL) | (uint64_t)(uint32_t)(ccell + 1));
  if (had_collision) return X_COLLIDE;
  ec.virt++;
  if (TRACE) HK_I(q, HI_NVI) = HK_I(q, HI_NVI) + 1;
  if (TRK == ABL_TRACK_CARTER) eq_carter_split<TRACE>(P, A, q, gslot, nslots);
  return X_MOVE;
}

template <int TRK, bool TRACE, bool TLE>
__global__ void __launch_bounds__(EQ_THREADS, 1) event_kernel(const DevProblem P, const RunArgs A) {
  EQFixed& S = EQS;
  const unsigned FULL = 0xffffffffu;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslots_cta = A.hk_slots, K = nslots_cta >> 5;  // slots of this CTA: K per lane
  // ---- set-up ----------------------------------------------------------------------------------------------------------
  for (int i = threadIdx.x; i < Q_N * EQ_RING * 32; i += EQ_THREADS) (&S.ring[0][0][0])[i] = EQ_EMPTY;
  for (int i = threadIdx.x; i < Q_N * 32; i += EQ_THREADS) {
    (&S.cnt[0][0])[i] = 0;
    (&S.head[0][0])[i] = 0;
    (&S.tail[0][0])[i] = 0;
  }
  for (int i = threadIdx.x; i < EQ_FQ; i += EQ_THREADS) S.fq_seq[i] = (unsigned)i;
  for (int i = threadIdx.x; i < (EQ_THREADS / 32) * WC_N; i += EQ_THREADS) (&S.wcnt[0][0])[i] = 0;
  for (int i = threadIdx.x; i < ST_N; i += EQ_THREADS) S.stat[i] = 0;
  if (threadIdx.x == 0) {
    S.fq_head = S.fq_tail = 0;
    S.live = nslots_cta;
    S.cur = Q_REFILL;
    S.abort = 0;
    S.deadline_ns = hk_now_ns() + A.timeout_ns;
    S.leak = S.leak_mig = 0.;
    for (int i = 0; i < RC_N; i++) S.rare[i] = 0;
  }
  if (A.hk_tables) {  // the table arena, staged behind the fixed part (the host already pointed P's tables at this copy)
    const uint4* src = reinterpret_cast<const uint4*>(A.arena);
    uint4* dst = reinterpret_cast<uint4*>(hk_shared_raw + EQ_COLS_OFFSET);
    for (int i = threadIdx.x; i < A.hk_tables / 16; i += EQ_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  if (A.hk_tables && threadIdx.x == 0 && (unsigned long long)(uintptr_t)(void*)hk_shared_raw != A.smem_generic_base) {
    raise_error(A, ABL_ERR_CUDA, 0);  // the shared window is not where the host assumed: the table pointers are wrong
    S.abort = 1;
  }
  // every slot starts in the REFILL queue
  if (threadIdx.x < 32) {
    for (int k = 0; k < K; k++) S.ring[Q_REFILL][k][lane] = (unsigned char)k;
    S.cnt[Q_REFILL][lane] = K;
    S.tail[Q_REFILL][lane] = (unsigned)K;
  }
  __syncthreads();

  const unsigned cols_base = EQ_COLS_OFFSET + (unsigned)A.hk_tables;
  const uint32_t nslots = gridDim.x * (uint32_t)nslots_cta;  // of the grid (the secondaries' stride)
  const uint64_t N = A.bank.n;
  const bool tle = TLE && A.converged && P.n_tl_tallies;
  const bool stats = A.eq_stats != nullptr;
  HAcc acc;
  acc.k_col = acc.k_abs = acc.mig = acc.k_trk = 0.;
  unsigned iter = 0;
  int patience = 0;

  // (every decision that steers the loop is made warp-uniform -- shuffle from lane 0 or a vote --: the lanes of a warp do not
  // read a shared variable at the same instant once they have diverged, and a warp that splits here would meet its own
  // full-mask votes with lanes missing)
  for (;;) {
    // watchdog: a loop that runs past the deadline is wound down with ABL_ERR_TIMEOUT instead of spinning for ever
    if ((++iter & 63u) == 0 && lane == 0 && hk_now_ns() > S.deadline_ns && !S.abort) {
      S.abort = 1;
      raise_error(A, ABL_ERR_TIMEOUT, 0);
    }
    if (__any_sync(FULL, S.abort != 0)) break;
    // ---- pick the event: the one the other warps are running while it still fills most of a warp -- warps that run the same
    // event at the same time share the instruction-cache lines they fetch (the loop body is several times the size of the
    // SM's instruction cache; without this clustering ncu showed 5.3 stall cycles per issue waiting for instructions) --
    // otherwise the event that fills most lanes, which becomes the CTA's current event
    int qi = -1, best = 0;
    const bool scavenge = (iter & 63u) == 32u;  // now and then the emptiest queue goes first, so that no slot waits for ever
    // finished slots first once half a warp of them waits: a slot in the REFILL (or BOUNDARY) queue is a slot that carries no
    // history, and the largest-queue rule below would leave them there (these queues grow at 1/45 of the rate of STEP)
    if ((iter & 3u) == 1u) {
      const int nr = __popc(__ballot_sync(FULL, *(volatile int*)&S.cnt[Q_REFILL][lane] > 0));
      const int nb = __popc(__ballot_sync(FULL, *(volatile int*)&S.cnt[Q_BOUNDARY][lane] > 0));
      if (nr >= EQ_REFILL_AT) {
        qi = Q_REFILL;
        best = nr;
      } else if (nb >= EQ_REFILL_AT) {
        qi = Q_BOUNDARY;
        best = nb;
      }
    }
    if (qi < 0 && !scavenge) {
      const int cur = __shfl_sync(FULL, S.cur, 0);
      const int n = __popc(__ballot_sync(FULL, *(volatile int*)&S.cnt[cur][lane] > 0));
      if (n >= EQ_STICKY) {
        qi = cur;
        best = n;
      }
    }
    if (qi < 0) {
#pragma unroll
      for (int c = 0; c < Q_N; c++) {
        const int n = __popc(__ballot_sync(FULL, *(volatile int*)&S.cnt[c][lane] > 0));
        if (scavenge ? (n > 0 && (qi < 0 || n < best)) : (n > best)) {
          best = n;
          qi = c;
        }
      }
      if (!scavenge && qi >= 0 && qi < Q_BOUNDARY && lane == 0) S.cur = qi;
    }
    if (qi < 0 || (iter & 3u) == 0) {  // (the job ring holds EQ_FQ jobs and an event posts about one: every 4th look is enough)
      const unsigned nj = *(volatile unsigned*)&S.fq_head - *(volatile unsigned*)&S.fq_tail;
      const int njobs = __shfl_sync(FULL, (int)min(nj, 32u), 0);
      if (njobs > 0 && (njobs >= 32 || njobs > best || qi < 0)) qi = Q_FISSION;
    }
    if (qi < 0) {
      const int fin = *(volatile int*)&S.live <= 0;
      if (__shfl_sync(FULL, fin, 0)) break;  // every slot is retired (so nothing can post a job any more)
      if (stats && lane == 0) atomicAdd(&S.stat[ST_IDLE], 1ULL);
      __nanosleep(200);
      continue;
    }
    if (qi == Q_FISSION) {
      // ---- fission jobs ---------------------------------------------------------------------------------------------------
      unsigned t = 0, n = 0;
      if (lane == 0) {
        for (;;) {
          t = *(volatile unsigned*)&S.fq_tail;
          n = min(32u, *(volatile unsigned*)&S.fq_head - t);
          if (n == 0 || atomicCAS(&S.fq_tail, t, t + n) == t) break;
        }
      }
      t = __shfl_sync(FULL, t, 0);
      n = __shfl_sync(FULL, n, 0);
      if (stats && lane == 0 && n) {
        atomicAdd(&S.stat[ST_TASKS + Q_FISSION], 1ULL);
        atomicAdd(&S.stat[ST_LANES + Q_FISSION], (unsigned long long)n);
      }
      if ((unsigned)lane < n) {
        const unsigned pos = t + lane, cell = pos % EQ_FQ;
        while (S.fq_seq[cell] != pos + 1 && !S.abort) {}
        __threadfence_block();
        const FisJob j = S.fq[cell];
        __threadfence_block();
        S.fq_seq[cell] = pos + EQ_FQ;
        if (!S.abort) {
          uint64_t rng = j.rng;
          const FissionTables ft{P.chi_cp, P.dg_cp, P.dg_off, P.gmid, P.G};
          bank_fission_sites<HK_MATH>(ft, A.sites, A.n_sites, A.site_capacity, rng, V3{j.x, j.y, j.z}, V3{j.ux, j.uy, j.uz}, j.w,
                                      j.parent, j.daughter0, j.n_new, j.mg / P.G, j.mg, ldt(&P.nud[j.mg]) / ldt(&P.nu[j.mg]));
        }
      }
      __syncwarp();
      continue;
    }
    // ---- take one slot per lane ------------------------------------------------------------------------------------------------
    if (EQ_MINFILL > 0 && best < EQ_MINFILL && patience < EQ_PATIENCE && !scavenge) {
      patience++;
      __nanosleep(EQ_NAP);
      continue;
    }
    patience = 0;
    const int k = eq_pop(S, qi, lane);
    const bool have = k >= 0;
    const unsigned act = __ballot_sync(FULL, have);
    if (act == 0) continue;
    EQ_FENCE();
    if (stats && lane == 0) {
      atomicAdd(&S.stat[ST_TASKS + qi], 1ULL);
      atomicAdd(&S.stat[ST_LANES + qi], (unsigned long long)__popc(act));
    }
    const int slot = (k << 5) | lane;
    const Cols q = make_cols_at(cols_base, have ? slot : lane, nslots_cta, A.hk_nf, A.hk_np, TRACE);
    const uint32_t gslot = blockIdx.x * (uint32_t)nslots_cta + (uint32_t)slot;
    int dest = -1;

    if (qi == Q_STEP) {
      // ---- STEP: arrive (real or virtual collision), collide, start the next flight ------------------------------------------------
      EvCount ec{0u, 0u, 0u, 0u};
      bool flew = false;
      int next = -1;
      if (have) {
        next = X_MOVE;
        if (!(HK_I(q, HI_EVT) & EV_MOVE_ONLY)) next = eq_arrive<TRK, TRACE>(P, A, q, HK_I(q, HI_CELL), tle, ec, gslot, nslots);
        if (next == X_COLLIDE) {  // Transporter::collision (transporter.cpp:60-93,269-312)
          bool alive = true;
          ICount ic{0u, 0u, 0u, 0u};
          collision_cols<HK_MATH, TRACE, EventPost>(P, A, q, hk_ld3(q, HD_R), HK_I(q, HI_HMAT), acc, ic, 0, alive);
          ec.real += ic.real;
          ec.coll_scores += ic.coll_scores;
          if (alive) {
            if (TRK == ABL_TRACK_CARTER) eq_carter_split<TRACE>(P, A, q, gslot, nslots);
            next = X_MOVE;
          } else {
            next = eq_particle_dead<TRK, TRACE>(P, A, q, gslot, nslots);
          }
        }
      }
      const unsigned movers = __ballot_sync(FULL, next == X_MOVE);  // the lanes that start a flight (they vote in the validation)
      flew = next == X_MOVE;
      if (flew) {
        // sample the flight, move the cursor, re-validate its pads (delta_tracker.cpp:105-118)
        uint64_t rng = HK_U(q, HD_RNG);
        const int g = HK_I(q, HI_G) & 0xff;
        const V3 u = hk_ld3(q, HD_U);
        const double d_coll = rng_exponential<HK_MATH>(rng, ldt(&P.smp[g]));
        HK_U(q, HD_RNG) = rng;
        HK_D(q, HD_DC) = d_coll;
        if (TRACE) HK_I(q, HI_NFL) = HK_I(q, HI_NFL) + 1;
        SCursor c;
        c.q = q;
        c.err = 0;
        const int pf = HK_I(q, HI_NPNF);
        c.np = pf & 0xff;
        c.nf = pf >> 8;
        cursor_move(c, d_coll, u);
        HK_I(q, HI_TOK) = 0;
        const int first_bad = cursor_validate(P, c, u);
        if (first_bad < c.np) {
          HK_I(q, HI_EVT) = first_bad | (EP_FLIGHT << 8);
          // a re-descent that starts at a cell universe stays inside it (the pin cell changed, the tile did not)
          dest = (first_bad > 0 && pad_type(pad_info(c, first_bad - 1)) == PAD_UNIVERSE) ? Q_LOC_CELL : Q_LOC_TREE;
        } else {
          HK_I(q, HI_EVT) = 0;
          dest = Q_STEP;
        }
      } else if (have) {
        dest = next;
      }
      if (have && dest >= 0) eq_push(S, dest, lane, k);
      __syncwarp();
      // the event's counters: flights by ballot; real / virtual collisions and tally scores summed in one packed reduction
      // (a lane adds at most 1, 1 and ABL_MAX_TALLIES: 10-bit fields cannot overflow over 32 lanes)
      const unsigned nf = __popc(__ballot_sync(FULL, flew));
      const unsigned packed = __reduce_add_sync(FULL, ec.real | (ec.virt << 10) | (ec.coll_scores << 20));
      unsigned nt = 0;
      if (tle) nt = __reduce_add_sync(FULL, ec.tl_bins);
      if (lane == 0) {
        S.wcnt[wid][WC_FLIGHTS] += nf;
        S.wcnt[wid][WC_REAL] += packed & 1023u;
        S.wcnt[wid][WC_VIRT] += (packed >> 10) & 1023u;
        S.wcnt[wid][WC_COLLSCORES] += packed >> 20;
        if (tle) S.wcnt[wid][WC_TLBINS] += nt;
      }
      continue;
    }
    unsigned tl_bins = 0;
    if (qi == Q_LOC_TREE || qi == Q_LOC_CELL) {
      // ---- LOCATE: (re-)descent through the universe tree, all lanes in step by universe type -------------------------------------
      if (have) {
        SCursor c;
        cursor_load(c, q);
        const V3 u = hk_ld3(q, HD_U);
        const int ev = HK_I(q, HI_EVT);
        const int need = ev & 0xff, part = ev >> 8;
        cursor_relocate_sync(P, c, need, u, act);
        cursor_store(c);
        if (c.err) raise_error(A, c.err, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
        const int ccell = c.cell;
        if (part == EP_FLIGHT) {
          if (ccell < 0) {
            dest = Q_BOUNDARY;  // left the geometry: the boundary is looked for from the pre-flight position
          } else {
            HK_I(q, HI_EVT) = 0;
            dest = Q_STEP;
          }
        } else if (ccell < 0) {
          if (part == EP_BIRTH) {  // lost at birth: warning + kill in the reference (delta_tracker.cpp:92-98)
            atomicAdd(&S.rare[RC_LOST], 1u);
          } else {
            raise_error(A, ABL_ERR_LOST, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
            if (TRK == ABL_TRACK_CARTER) HK_I(q, HI_NSEC) = 0;
          }
          dest = eq_particle_dead<TRK, TRACE>(P, A, q, gslot, nslots);
        } else {
          if (part == EP_REFLECTED) {
            note_col<TRACE>(q, A.hk_np, 0x4000000000000000ULL | (uint64_t)(uint32_t)(ccell + 1));
            if (TRK == ABL_TRACK_CARTER) eq_carter_split<TRACE>(P, A, q, gslot, nslots);
          } else {  // birth, resurrection: the MaterialHelper of the new particle
            HK_I(q, HI_HMAT) = c.mat;
          }
          HK_I(q, HI_EVT) = EV_MOVE_ONLY;
          dest = Q_STEP;
        }
      }
    } else if (qi == Q_REFILL) {
      // ---- REFILL: the next bank index (one aggregated atomic per warp) ----------------------------------------------------------
      if (have) {
        unsigned long long idx;
        {
          cg::coalesced_group grp = cg::coalesced_threads();
          unsigned long long base = 0;
          if (grp.thread_rank() == 0) base = atomicAdd(A.ticket, (unsigned long long)grp.size());
          idx = grp.shfl(base, 0) + grp.thread_rank();
        }
        if (idx >= N) {
          atomicSub(&S.live, 1);  // the bank is empty: the slot retires
        } else {
          // (a streamed bank -- A.avail -- : wait until the row has been copied; L2 loads, a cached L1 line could hold
          // neighbouring rows from before they arrived)
          if (A.avail != nullptr)
            while (idx >= *(const volatile unsigned long long*)A.avail && !S.abort) __nanosleep(256);
          const V3 r{__ldcg(&A.bank.x[idx]), __ldcg(&A.bank.y[idx]), __ldcg(&A.bank.z[idx])};
          const V3 u{__ldcg(&A.bank.ux[idx]), __ldcg(&A.bank.uy[idx]), __ldcg(&A.bank.uz[idx])};
          const double E = __ldcg(&A.bank.E[idx]);
          HK_I(q, HI_IDX) = (int)(uint32_t)idx;
          hk_st3(q, HD_R, r);
          hk_st3(q, HD_RB, r);
          hk_st3(q, HD_U, u);
          HK_D(q, HD_E) = E;
          HK_D(q, HD_W) = __ldcg(&A.bank.wgt[idx]);
          const int g = group_of(P, E);
          HK_I(q, HI_G) = g | ((g < P.G && E == group_mid(P, g)) ? 0x100 : 0);
          HK_U(q, HD_RNG) = __ldcg(&A.bank.id_c[idx]);  // pcg32 state: seeded by seed_streams_kernel / source sampling
          HK_I(q, HI_DAU) = 0;
          HK_I(q, HI_NSEC) = 0;
          HK_I(q, HI_TOK) = 0;
          HK_I(q, HI_NPNF) = 1 << 8;
          HK_I(q, HI_CELL) = -1;
          HK_I(q, HI_MAT) = -1;
          HK_I(q, HI_HMAT) = -1;
          if (TRACE) {
            HK_U(q, HD_PT + A.hk_np) = 1469598103934665603ULL;
            HK_I(q, HI_NFL) = 0;
            HK_I(q, HI_NRE) = 0;
            HK_I(q, HI_NVI) = 0;
          }
          HK_FR(q, 0, 0) = r.x;
          HK_FR(q, 1, 0) = r.y;
          HK_FR(q, 2, 0) = r.z;
          HK_I(q, HI_EVT) = 0 | (EP_BIRTH << 8);
          dest = Q_LOC_TREE;
        }
      }
    } else {
      // ---- BOUNDARY: Tracker::restart_get_current at the pre-flight position + get_boundary_condition
      // (delta_tracker.cpp:120-127), then leak (delta_tracker.cpp:233-238) or Tracker::do_reflection (tracker.hpp:314-360) --------
      if (have) {
        const V3 r = hk_ld3(q, HD_R), u = hk_ld3(q, HD_U);  // (the pre-flight position: the flight never happened)
        const GeoTables G = geo_tables(P);
        Cursor lc;
        lc.err = 0;
        lc.token = 0;
        cursor_restart_nl(G, lc, r, u);
        const Boundary b = cursor_boundary_condition_nl(G, lc, u);
        const uint32_t idx = (uint32_t)HK_I(q, HI_IDX);
        if (lc.err) raise_error(A, lc.err, A.bank.id_a[idx]);
        const double w = HK_D(q, HD_W);
        if (tle) {  // delta_tracker.cpp:133: scored from the pre-move position over min(d_coll, boundary distance)
          const int mg = HK_I(q, HI_HMAT) * P.G + (HK_I(q, HI_G) & 0xff);
          tl_bins += score_flight_cols(tle_args(P), q, mg, fmin(HK_D(q, HD_DC), b.distance));
        }
        atomicAdd(&S.rare[RC_BOUNDARY], 1u);
        if (b.btype == ABL_BC_VACUUM) {
          note_col<TRACE>(q, A.hk_np, 0x3000000000000000ULL | (uint64_t)(uint32_t)(lc.cell + 1));
          const V3 rb = hk_ld3(q, HD_RB);
          const V3 d{r.x + b.distance * u.x - rb.x, r.y + b.distance * u.y - rb.y, r.z + b.distance * u.z - rb.z};
          atomicAdd(&S.leak, w);
          atomicAdd(&S.leak_mig, leak_mig_score(w, d));
          dest = eq_particle_dead<TRK, TRACE>(P, A, q, gslot, nslots);
        } else if (b.btype == ABL_BC_REFLECTIVE && b.surface_index >= 0) {
          const Reflected rf = reflect_nl(G.surfaces, b.surface_index, r, u, b.distance);
          hk_st3(q, HD_R, rf.r);
          hk_st3(q, HD_U, rf.u);
          HK_I(q, HI_TOK) = 0;
          HK_FR(q, 0, 0) = rf.r.x;
          HK_FR(q, 1, 0) = rf.r.y;
          HK_FR(q, 2, 0) = rf.r.z;
          HK_I(q, HI_EVT) = 0 | (EP_REFLECTED << 8);
          dest = Q_LOC_TREE;
        } else {
          raise_error(A, ABL_ERR_LOST, A.bank.id_a[idx]);
          if (TRK == ABL_TRACK_CARTER) HK_I(q, HI_NSEC) = 0;
          dest = eq_particle_dead<TRK, TRACE>(P, A, q, gslot, nslots);
        }
      }
    }
    // ---- hand every slot to its next event -------------------------------------------------------------------------------------
    if (have && dest >= 0) eq_push(S, dest, lane, k);
    __syncwarp();
    if (tle && qi == Q_BOUNDARY) {
      const unsigned nt = __reduce_add_sync(FULL, tl_bins);
      if (lane == 0) S.wcnt[wid][WC_TLBINS] += nt;
    }
  }

  // ---- the scores: per lane in registers -> per warp -> one atomic per block -----------------------------------------------------------
  __syncwarp();
  {
    double s0 = acc.k_col, s1 = acc.k_abs, s2 = acc.mig;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(FULL, s0, o);
      s1 += __shfl_xor_sync(FULL, s1, o);
      s2 += __shfl_xor_sync(FULL, s2, o);
    }
    if (lane == 0) {
      S.sd[wid][0] = s0;
      S.sd[wid][1] = s1;
      S.sd[wid][2] = s2;
    }
  }
  constexpr int NW = EQ_THREADS / 32;
  __syncthreads();
  if (threadIdx.x < 5) {
    const int kk = threadIdx.x;  // k_col, k_abs, (k_trk: not scored by these trackers), leak, mig
    double v = kk == 3 ? S.leak : (kk == 4 ? S.leak_mig : 0.);
    if (kk < 2)
      for (int w = 0; w < NW; w++) v += S.sd[w][kk];
    if (kk == 4)
      for (int w = 0; w < NW; w++) v += S.sd[w][2];
    const int slot = kk < 3 ? kk : kk + 1;  // scores layout: k_col,k_abs,k_trk,k_tot(unused),leak,mig
    if (kk != 2) atomicAdd(&A.scores[slot], v);
  } else if (threadIdx.x >= 32 && threadIdx.x < 40) {
    // counters layout: flights, real, virtual, tl_bins, sites, boundary, lost, coll_scores
    const int kk = threadIdx.x - 32;
    unsigned long long v = kk == 4 ? S.rare[RC_SITES] : (kk == 5 ? S.rare[RC_BOUNDARY] : (kk == 6 ? S.rare[RC_LOST] : 0u));
    const int wc = kk == 0 ? WC_FLIGHTS : kk == 1 ? WC_REAL : kk == 2 ? WC_VIRT : kk == 3 ? WC_TLBINS : kk == 7 ? WC_COLLSCORES : -1;
    if (wc >= 0)
      for (int w = 0; w < NW; w++) v += S.wcnt[w][wc];
    atomicAdd(&A.counters[kk], v);
  } else if (stats && threadIdx.x >= 64 && threadIdx.x < 64 + ST_N) {
    atomicAdd(&A.eq_stats[threadIdx.x - 64], S.stat[threadIdx.x - 64]);
  }
}

}  // namespace abl
