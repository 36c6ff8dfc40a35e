/* history.cuh -- the delta- / carter-tracking history kernel (v1): warp-synchronous staged state machine.
 *
 * Same arithmetic, same RNG consumption and same per-history outcomes as transport_kernel (transport.cuh), which
 * follows DeltaTracker::transport (src/delta_tracker.cpp:72-263) and CarterTracker::transport
 * (src/carter_tracker.cpp:92-294) of the reference.  What changes is how the 32 lanes of a warp walk through it.
 *
 * v0 let every lane run the reference's control flow on its own; ncu showed 7.6 of 32 lanes active per issued
 * instruction and an instruction-cache bound kernel (228 KB of SASS), because the rare, long paths (history
 * start-up, geometry re-descent, boundary search, reflection, fission banking) were inlined at every call site
 * and executed by one or two lanes while the rest of the warp waited.
 *
 * v1 gives every lane a small phase variable and runs ONE loop body per warp whose stages appear exactly once
 * in the code:
 *
 *     R  refill    dead lanes take the next bank index (one aggregated atomic per warp)
 *     M  move      lanes in flight sample the distance, advance the geometry cursor, re-validate its pads
 *     L  locate    every lane that needs a (re-)descent through the universe tree does it here -- births,
 *                  tile / cell changes, the rewind of a lost particle, reflections, resurrected secondaries
 *     B  boundary  lost particles: boundary-condition search, leak or reflect
 *     T  track-length tally (one call site)
 *     C  collide   real / virtual decision, Transporter::collision
 *     E  end       secondaries, history epilogue
 *
 * A lane whose particle needs a rare path simply sits out the stages it cannot take part in for one or two
 * iterations (e.g. lost -> [L at the old position] -> B -> [L at the reflected position]); the common path
 * M -> L -> C runs with most lanes active.  __all_sync at the top of the loop is the reconvergence point.
 */
#pragma once
#include "transport.cuh"

namespace abl {

enum { PH_DEAD = 0, PH_FLIGHT, PH_BIRTH, PH_LOST, PH_REFLECTED, PH_RESURRECT };

// Tracker::get_current, first half (tracker.hpp:235-270): index of the first pad that no longer holds, or np
__device__ __forceinline__ int cursor_validate(const DevProblem& P, const Cursor& c, const V3& u) {
  int first_bad = c.np;
  for (int it = 0; it < c.np; it++) {
    const int info = c.pinfo[it];
    const int type = pad_type(info);
    if (type == PAD_CELL) {
      if (!cell_is_inside_fast(P, pad_index(info), frame_r(c, pad_frame(info)), u, c.token)) {
        first_bad = it;
        break;
      }
    } else if (type == PAD_LATTICE) {
      const Tile3 t3 = lattice_tile_nl(P.universes + pad_index(info), frame_r(c, pad_frame(info)), u);
      if (c.ptile[it][0] != t3.nx || c.ptile[it][1] != t3.ny || c.ptile[it][2] != t3.nz) {
        first_bad = it;
        break;
      }
    }
  }
  return first_bad;
}

// Tracker::get_current, second half (tracker.hpp:272-306) and Tracker::restart_get_current (tracker.hpp:63-74):
// re-descend from the pad above the first bad one; from == 0 is a full lookup from the root at frame 0.
__device__ __forceinline__ void cursor_relocate(const DevProblem& P, Cursor& c, int from, const V3& u) {
  int uni = P.root, f = 0;
  bool full = true;
  if (from > 0) {
    const int back = c.pinfo[from - 1];
    if (pad_type(back) != PAD_CELL) {  // (a lattice directly inside a cell: see cursor_get_current in geom.cuh)
      c.np = from - 1;
      uni = pad_index(back);
      f = pad_frame(back);
      full = false;
    }
  }
  if (full) {
    c.np = 0;
    c.nf = 1;
  }
  for (;;) {
    c.cell = descend<true>(P, c, uni, f, u);
    if (c.cell >= 0 || full) break;
    full = true;  // partial re-descent failed: full restart from the root
    c.np = 0;
    c.nf = 1;
    uni = P.root;
    f = 0;
  }
  c.mat = c.cell >= 0 ? __ldg(&P.cells[c.cell].material) : -1;
}

__device__ __noinline__ Boundary cursor_boundary_condition_nl(const GeoTables G, const Cursor& c, const V3 u) {
  return cursor_boundary_condition(G, c, u);
}

// Tracker::do_reflection (tracker.hpp:314-360), the arithmetic part: point on the surface and reflected direction
struct Reflected {
  V3 r, u;
};
__device__ __noinline__ Reflected reflect_nl(const abl_surface* __restrict__ surfaces, int surface_index, const V3 r, const V3 u, double distance) {
  GeoTables G{};
  G.surfaces = surfaces;
  const Surf s = load_surface(G, surface_index);
  Reflected o;
  o.r = V3{r.x + distance * u.x, r.y + distance * u.y, r.z + distance * u.z};
  const V3 n = surf_norm(s, o.r);
  const double f = 2. * dot3(u, n);
  o.u = make_direction(u.x - n.x * f, u.y - n.y * f, u.z - n.z * f);
  return o;
}

// every track-length tally (tallies.hpp:57-63); returns the number of bins scored
__device__ __noinline__ int score_flight_all_nl(const DevTally* __restrict__ tallies, int ntallies, const V3 r, const V3 u, double d,
                                                double E, double w, double w2, const MatXS mx) {
  int nb = 0;
  for (int t = 0; t < ntallies; t++) {
    if (__ldg(&tallies[t].estimator) != ABL_EST_TRACK_LENGTH) continue;
    const DevTally T = tallies[t];
    nb += score_flight(T, r, u, d, E, w, w2, mx);
  }
  return nb;
}

// One CTA per SM.  All its warps execute the loop in lock step (block-wide vote at the top, barriers between the
// stages): the SM instruction cache is much smaller than the loop body, and ncu showed the free-running version
// bound by instruction-cache misses (sm__icc hit rate 57 %, 60 % of stall samples "no instruction").  In lock step
// a line fetched for one warp is a hit (or a hit under miss) for the other fifteen.
#ifndef HK_THREADS
#define HK_THREADS 768
#endif
#ifndef HK_STAGE_SYNC
#define HK_STAGE_SYNC 1
#endif
#ifndef HK_MATH
#define HK_MATH CallMath
#endif
#ifndef HK_VOTE_EVERY
#define HK_VOTE_EVERY 1  // block-wide vote (and re-alignment of the warps) every this many iterations (power of two)
#endif
#if HK_STAGE_SYNC
#define HK_SYNC() __syncthreads()
#else
#define HK_SYNC() __syncwarp()
#endif

template <int TRK, bool TRACE>
__global__ void __launch_bounds__(HK_THREADS, 1) history_kernel(const DevProblem P, const RunArgs A) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const unsigned FULL = 0xffffffffu;
  Acc acc;
  acc.k_col = acc.k_abs = acc.k_trk = acc.leak = acc.mig = 0.;
  acc.flights = acc.real = acc.virt = acc.tl_bins = acc.sites = acc.boundary = acc.lost = acc.coll_scores = 0;
  Hist h;
  h.alive = false;
  h.nsec = 0;
  h.idx = 0;
  Cursor c;
  c.err = 0;
  c.np = 0;
  c.nf = 1;
  c.token = 0;
  c.cell = c.mat = -1;
  int phase = PH_DEAD;
  int need = -1;        // pending (re-)descent: first bad pad, 0 = full lookup from the root, -1 = none
  double d_coll = 0.;   // sampled flight distance (kept across the iterations of a boundary event)
  bool exhausted = false;
  const uint64_t N = A.bank.n;
  const bool tle = A.converged && P.n_tl_tallies;
  uint32_t iter = 0;

  for (;;) {
    // ---- R: refill ---------------------------------------------------------------------------------------------
    if (phase == PH_DEAD && !exhausted) {
      unsigned long long idx;
      {
        cg::coalesced_group grp = cg::coalesced_threads();
        unsigned long long base = 0;
        if (grp.thread_rank() == 0) base = atomicAdd(A.ticket, (unsigned long long)grp.size());
        idx = grp.shfl(base, 0) + grp.thread_rank();
      }
      if (idx >= N) {
        exhausted = true;
      } else {
        h.idx = (uint32_t)idx;
        h.r = {A.bank.x[idx], A.bank.y[idx], A.bank.z[idx]};
        h.u = {A.bank.ux[idx], A.bank.uy[idx], A.bank.uz[idx]};
        h.rb = h.r;
        h.E = A.bank.E[idx];
        h.w = A.bank.wgt[idx];
        h.w2 = 0.;
        h.g = group_of(P, h.E);
        h.emid = h.g < P.G && h.E == group_mid(P, h.g);
        h.rng = A.bank.id_c[idx];  // pcg32 state: seeded by seed_streams_kernel / source sampling
        h.hash = 1469598103934665603ULL;
        h.daughter = 0;
        h.n_flights = h.n_real = h.n_virtual = h.n_fis = 0;
        h.nsec = 0;
        h.alive = true;
        c.token = 0;
        c.fx[0] = h.r.x;
        c.fy[0] = h.r.y;
        c.fz[0] = h.r.z;
        need = 0;
        phase = PH_BIRTH;
      }
    }
    if ((iter++ & (HK_VOTE_EVERY - 1)) == 0) {
      if (__syncthreads_and(phase == PH_DEAD)) break;
    }

    // ---- M: sample the flight, move the cursor, re-validate its pads ----------------------------------------------------------
    if (phase == PH_FLIGHT) {
      d_coll = rng_exponential<HK_MATH>(h.rng, __ldg(&P.smp[h.g]));
      acc.flights++;
      if (TRACE) h.n_flights++;
      cursor_move(c, d_coll, h.u);
      const int first_bad = cursor_validate(P, c, h.u);
      if (first_bad < c.np) need = first_bad;
    }
    HK_SYNC();

    // ---- L: (re-)descent through the universe tree -----------------------------------------------------------------------------------
    if (need >= 0) {
      cursor_relocate(P, c, need, h.u);
      need = -1;
      if (c.err) {
        raise_error(A, c.err, A.bank.id_a[h.idx]);
        c.err = 0;
      }
    }
    HK_SYNC();

    // ---- B1: what did the move do? -------------------------------------------------------------------------------------------------------
    double tle_d = -1.;
    bool test_collision = false, flight_done = false;
    Boundary bound{ABL_INF, -1, ABL_BC_NORMAL, 0};
    if (phase == PH_FLIGHT) {
      if (c.cell < 0) {  // left the geometry: rewind to the pre-flight position and look for the boundary
        c.token = 0;
        c.fx[0] = h.r.x;
        c.fy[0] = h.r.y;
        c.fz[0] = h.r.z;
        need = 0;
        phase = PH_LOST;
      } else {
        tle_d = d_coll;
        test_collision = true;
      }
    } else if (phase == PH_LOST) {  // the cursor is back at the pre-flight position (delta_tracker.cpp:120-127)
      bound = cursor_boundary_condition_nl(geo_tables(P), c, h.u);
      tle_d = fmin(d_coll, bound.distance);
    }
    // ---- T: track-length mesh tallies, scored from the pre-move position -----------------------------------------------------------------
    if (tle && tle_d >= 0.) {
      const int mg = h.mat * P.G + h.g;
      const MatXS mx{__ldg(&P.Et[mg]), __ldg(&P.Ea[mg]), __ldg(&P.Ef[mg]), __ldg(&P.Es[mg])};
      acc.tl_bins += score_flight_all_nl(P.tally_dev, P.ntallies, h.r, h.u, tle_d, h.E, h.w, h.w2, mx);
    }

    // ---- B2: boundary events and the checks that follow a full lookup ---------------------------------------------------------------------------
    if (phase == PH_LOST && need < 0) {
      acc.boundary++;
      if (bound.btype == ABL_BC_VACUUM) {
        if (TRACE) note(h, 0x3000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
        leak(h, acc, bound);
        phase = PH_FLIGHT;
      } else if (bound.btype == ABL_BC_REFLECTIVE && bound.surface_index >= 0) {
        // Tracker::do_reflection (tracker.hpp:314-360); the full lookup happens in the next L stage
        const Reflected rf = reflect_nl(P.surfaces, bound.surface_index, h.r, h.u, bound.distance);
        h.u = rf.u;
        h.r = rf.r;
        c.token = 0;
        c.fx[0] = h.r.x;
        c.fy[0] = h.r.y;
        c.fz[0] = h.r.z;
        need = 0;
        phase = PH_REFLECTED;
      } else {
        raise_error(A, ABL_ERR_LOST, A.bank.id_a[h.idx]);
        h.alive = false;
        h.nsec = 0;
        phase = PH_FLIGHT;
      }
    } else if (phase == PH_REFLECTED && need < 0) {
      if (c.cell < 0) {
        raise_error(A, ABL_ERR_LOST, A.bank.id_a[h.idx]);
        h.alive = false;
        h.nsec = 0;
      } else if (TRACE) {
        note(h, 0x4000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
      }
      flight_done = true;
      phase = PH_FLIGHT;
    } else if (phase == PH_BIRTH) {
      if (c.cell < 0) {  // lost at birth: warning + kill in the reference (delta_tracker.cpp:92-98)
        acc.lost++;
        h.alive = false;
      } else {
        h.mat = c.mat;
      }
      phase = PH_FLIGHT;
    } else if (phase == PH_RESURRECT) {
      if (c.cell < 0) {
        raise_error(A, ABL_ERR_LOST, A.bank.id_a[h.idx]);
        h.alive = false;
        h.nsec = 0;
      } else {
        h.mat = c.mat;
      }
      phase = PH_FLIGHT;
    }
    HK_SYNC();

    // ---- C: arrive, real or virtual collision (delta_tracker.cpp:167-195, carter_tracker.cpp:189-207) ----------------------------------------
    if (test_collision) {
      bool had_collision = false;
      flight_done = true;
      h.r.x = h.r.x + d_coll * h.u.x;
      h.r.y = h.r.y + d_coll * h.u.y;
      h.r.z = h.r.z + d_coll * h.u.z;
      h.mat = c.mat;
      const double Esample = __ldg(&P.smp[h.g]);
      const double Et = __ldg(&P.Et[h.mat * P.G + h.g]);
      if (TRK == ABL_TRACK_DELTA) {
        if (Et - Esample > 1.E-10) {
          raise_error(A, ABL_ERR_MAJORANT, A.bank.id_a[h.idx]);
          h.alive = false;
          h.nsec = 0;
        } else if (HK_MATH::rand(h.rng) < HK_MATH::div(Et, Esample)) {
          had_collision = true;
        }
      } else {
        if (Esample >= Et) {
          if (HK_MATH::rand(h.rng) < HK_MATH::div(Et, Esample)) had_collision = true;
        } else {  // under-estimated sampling xs: signed-weight branch (carter_tracker.cpp:192-207)
          const double D = HK_MATH::div(Et, 2. * Et - Esample);
          const double F = HK_MATH::div(Et, D * Esample);
          if ((D - HK_MATH::rand(h.rng)) > 0.) {
            h.w = h.w * F;
            had_collision = true;
          } else {
            h.w = -h.w * F;
          }
        }
      }
      if (h.alive) {
        if (TRACE) note(h, (had_collision ? 0x2000000000000000ULL : 0x1000000000000000ULL) | (uint64_t)(uint32_t)(c.cell + 1));
        if (had_collision) {
          collision<false, HK_MATH>(P, A, h, acc);
        } else {
          acc.virt++;
          if (TRACE) h.n_virtual++;
        }
      }
    }
    if (TRK == ABL_TRACK_CARTER) {
      if (flight_done && h.alive && fabs(h.w) >= P.wgt_split) {  // Particle::split (particle.hpp:165-173)
        const int n_new = (int)ceil(fabs(h.w));
        if (n_new > 1) {
          h.w = h.w / (double)n_new;
          h.w2 = h.w2 / (double)n_new;
          for (int np = 0; np < n_new - 1; np++)
            if (!push_secondary(A, h, h.u, h.E, h.w, h.w2, tid, nthreads)) {
              raise_error(A, ABL_ERR_BANK_OVERFLOW, A.bank.id_a[h.idx]);
              break;
            }
        }
      }
    }
    HK_SYNC();

    // ---- E: secondaries, end of history -----------------------------------------------------------------------------------------------------------
    if (phase == PH_FLIGHT && !h.alive) {
      if (h.nsec > 0) {  // Particle::resurect + Tracker restart (delta_tracker.cpp:197-229)
        pop_secondary(P, A, h, tid, nthreads);
        c.token = 0;
        c.fx[0] = h.r.x;
        c.fy[0] = h.r.y;
        c.fz[0] = h.r.z;
        need = 0;
        phase = PH_RESURRECT;
      } else {
        A.nfis[h.idx] = h.n_fis;
        if (TRACE) {
          A.tr_flights[h.idx] = h.n_flights;
          A.tr_real[h.idx] = h.n_real;
          A.tr_virtual[h.idx] = h.n_virtual;
          A.tr_hash[h.idx] = h.hash;
          A.tr_rng[h.idx] = h.rng;
        }
        phase = PH_DEAD;
      }
    }
  }

  // ---- reduce the per-thread accumulators: warp shuffle, then one atomic per block ------------------------------------
  double dv[5] = {acc.k_col, acc.k_abs, acc.k_trk, acc.leak, acc.mig};
  unsigned long long cv[8] = {acc.flights, acc.real, acc.virt, acc.tl_bins, acc.sites, acc.boundary, acc.lost, acc.coll_scores};
  constexpr int NW = HK_THREADS / 32;
  __shared__ double sd[NW][5];
  __shared__ unsigned long long sc[NW][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 5; q++) {
    double v = dv[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    if (lane == 0) sd[wid][q] = v;
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    unsigned long long v = cv[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    if (lane == 0) sc[wid][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    const int q = threadIdx.x;
    double v = 0.;
    for (int w = 0; w < NW; w++) v += sd[w][q];
    const int slot = q < 3 ? q : q + 1;  // scores layout: k_col,k_abs,k_trk,k_tot(unused),leak,mig
    atomicAdd(&A.scores[slot], v);
  } else if (threadIdx.x >= 32 && threadIdx.x < 40) {
    const int q = threadIdx.x - 32;
    unsigned long long v = 0;
    for (int w = 0; w < NW; w++) v += sc[w][q];
    atomicAdd(&A.counters[q], v);
  }
}

// pcg32 state of every history of a bank: seed(seed); advance(stride * history_id) (particle.hpp:188-193)
__global__ void __launch_bounds__(256) seed_streams_kernel(const DevProblem P, const uint64_t* __restrict__ history_id, uint64_t n,
                                                           uint64_t* __restrict__ state) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    state[i] = pcg_advance(P.seed_state, P.stride * history_id[i], P.jump);
}

}  // namespace abl
