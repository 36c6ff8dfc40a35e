/* history.cuh -- the staged history kernel (delta, carter and surface tracking; k-eigenvalue mode).
 *
 * Same arithmetic, same RNG consumption and same per-history outcomes as the per-lane kernel (transport.cuh), which
 * follows DeltaTracker::transport (src/delta_tracker.cpp:72-263), CarterTracker::transport
 * (src/carter_tracker.cpp:92-294) and SurfaceTracker::transport (src/surface_tracker.cpp:40-219) of the reference.
 * What changes is how the lanes of a warp walk through it.
 *
 * Left to itself (every lane running the reference's control flow) the loop executed with 7.6 of 32 lanes active and
 * was bound by instruction-cache misses (228 KB of SASS): the rare, long paths -- history start-up, geometry
 * re-descent, boundary search, reflection, fission banking -- were inlined at every call site and executed by one or
 * two lanes while the rest of the warp waited.
 *
 * Here every lane carries a small phase variable and a warp runs ONE loop body whose stages appear exactly once:
 *
 *     R  refill    dead lanes take the next bank index (one aggregated atomic per warp); with a streamed input bank
 *                  (abl_transport) a lane waits here until its row has arrived
 *     M  move      lanes in flight sample the distance, advance the geometry cursor, re-validate its pads
 *                  (surface tracking: nearest-boundary search, track-length score, crossing / reflection / collision)
 *     L  locate    every lane that needs a (re-)descent through the universe tree does it here, all lanes of the warp
 *                  in step by universe type -- births, tile / cell changes, crossings, reflections, secondaries
 *     B  boundary  delta / carter: a lane that left the geometry posts a request to the service warp and parks
 *     T  track-length tally (delta / carter; one call site)
 *     C  collide   real / virtual decision, Transporter::collision (fission sites go to the service warp as jobs)
 *     E  end       secondaries, history epilogue
 *
 * The state of every history (cursor included) lives in shared-memory columns, each stage loads what it works on, and
 * the rare long events run on a service warp; see "CTA layout" below.
 * DESIGN.md section 3.1 has the measurements behind each of these choices.
 */
#pragma once
#include "transport.cuh"

namespace abl {

enum { PH_DEAD = 0, PH_FLIGHT, PH_BIRTH, PH_LOST, PH_REFLECTED, PH_RESURRECT, PH_CROSSED };

// Tracker::get_current, first half (tracker.hpp:235-270): index of the first pad that no longer holds, or np
template <class CUR, bool HEX = true>
__device__ __forceinline__ int cursor_validate(const DevProblem& P, const CUR& c, const V3& u) {
  int first_bad = c.np;
  for (int it = 0; it < c.np; it++) {
    const int info = pad_info(c, it);
    const int type = pad_type(info);
    if (type == PAD_CELL) {
      if (!cell_is_inside_fast(P, pad_index(info), frame_r(c, pad_frame(info)), u, c.token)) {
        first_bad = it;
        break;
      }
    } else if (type == PAD_LATTICE) {
      const Tile3 t3 = lattice_tile_nl<HEX>(P.universes + pad_index(info), frame_r(c, pad_frame(info)), u);
      if (!pad_tile_is(c, it, t3.nx, t3.ny, t3.nz)) {
        first_bad = it;
        break;
      }
    }
  }
  return first_bad;
}

// The same result with the pads visited by TYPE -- every lattice pad of every lane first, then the cell pads above the first
// lattice that failed -- so that the lanes of a warp run the tile test together and the cell test together whatever the
// depth of their pad stacks (a history in the reflector holds [lattice, universe, cell], one in a pin [lattice, lattice,
// universe, cell]: visited in index order the cell test of the one met the lattice test of the other).  The pads hold or
// fail independently of each other, so the first bad pad in index order is the smaller of the first bad lattice and the first
// bad cell below it.  mask = the lanes that call this together.
template <class CUR>
__device__ __forceinline__ int cursor_validate_by_type(const DevProblem& P, const CUR& c, const V3& u, unsigned mask) {
  int lim = c.np;
  for (int pass = 0; pass < 2; pass++) {
    const int want = pass == 0 ? PAD_LATTICE : PAD_CELL;
    int it = 0;
    for (;;) {
      while (it < lim && pad_type(pad_info(c, it)) != want) it++;
      const bool go = it < lim;
      if (!__any_sync(mask, go)) break;
      if (go) {
        const int info = pad_info(c, it);
        bool ok;
        if (pass == 0) {
          const Tile3 t3 = lattice_tile_nl(P.universes + pad_index(info), frame_r(c, pad_frame(info)), u);
          ok = pad_tile_is(c, it, t3.nx, t3.ny, t3.nz);
        } else {
          ok = cell_is_inside_fast(P, pad_index(info), frame_r(c, pad_frame(info)), u, c.token);
        }
        if (!ok) lim = it;  // nothing above a bad pad matters
        it++;
      }
    }
  }
  return lim;
}

// Tracker::get_current, second half (tracker.hpp:272-306) and Tracker::restart_get_current (tracker.hpp:63-74):
// re-descend from the pad above the first bad one; from == 0 is a full lookup from the root at frame 0.
template <class CUR>
__device__ __forceinline__ void cursor_relocate(const DevProblem& P, CUR& c, int from, const V3& u) {
  int uni = P.root, f = 0;
  bool full = true;
  if (from > 0) {
    const int back = pad_info(c, from - 1);
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
  c.mat = c.cell >= 0 ? ldt(&P.cells[c.cell].material) : -1;
}

// One step of Universe::get_cell through a lattice (rect_lattice.cpp:132-207): 0 = descend further (uni / f updated),
// -1 = no universe here (lost)
template <class CUR, bool HEX = true>
__device__ __forceinline__ int descend_lattice_step(const DevProblem& P, CUR& c, int& uni, int& f, const V3& u) {
  const abl_universe* U = P.universes + uni;
  const V3 r = frame_r(c, f);
  if (HEX && ldt(&U->type) == ABL_UNI_HEX) {  // HexLattice::get_cell, hex_lattice.cpp:140-202
    const HexStep hs = hex_lattice_step(U, P.tiles, r);
    c.nf = f + 1;
    if (hs.sub >= 0) {
      if (!push_pad(c, make_pad(PAD_LATTICE, 0, f, uni), hs.t.nx, hs.t.ny, hs.t.nz)) return -1;
      if (f + 1 >= ABL_MAX_FRAMES) {
        c.err = ABL_ERR_GEOMETRY;
        return -1;
      }
      set_frame(c, f + 1, hs.r_tile.x, hs.r_tile.y, hs.r_tile.z);
      f++;
      uni = hs.sub;
      return 0;
    }
    if (hs.outer >= 0) {
      if (!push_pad(c, make_pad(PAD_LATTICE, 1, f, uni), hs.t.nx, hs.t.ny, hs.t.nz)) return -1;
      uni = hs.outer;
      return 0;
    }
    push_pad(c, make_pad(PAD_LATTICE, 0, f, uni), hs.t.nx, hs.t.ny, hs.t.nz);
    return -1;
  }
  const Tile3 t3 = lattice_tile_nl<HEX>(U, r, u);
  Lat L;
  L.Nx = ldt(&U->N[0]); L.Ny = ldt(&U->N[1]); L.Nz = ldt(&U->N[2]);
  L.tile_offset = ldt(&U->tile_offset);
  L.outer = ldt(&U->outer);
  int sub = -1;
  if (tile_in_range(L, t3.nx, t3.ny, t3.nz)) sub = ldt(&P.tiles[L.tile_offset + t3.nz * (L.Nx * L.Ny) + t3.nx * L.Ny + t3.ny]);
  c.nf = f + 1;
  if (sub >= 0) {
    if (!push_pad(c, make_pad(PAD_LATTICE, 0, f, uni), t3.nx, t3.ny, t3.nz)) return -1;
    if (f + 1 >= ABL_MAX_FRAMES) {
      c.err = ABL_ERR_GEOMETRY;
      return -1;
    }
    L.Px = ldt(&U->P[0]); L.Py = ldt(&U->P[1]); L.Pz = ldt(&U->P[2]);
    L.Xl = ldt(&U->Xl[0]); L.Yl = ldt(&U->Xl[1]); L.Zl = ldt(&U->Xl[2]);
    const V3 ctr = tile_center(L, t3.nx, t3.ny, t3.nz);
    set_frame(c, f + 1, r.x - ctr.x, r.y - ctr.y, r.z - ctr.z);
    f++;
    uni = sub;
    return 0;
  }
  if (L.outer >= 0) {  // outside the lattice or an empty tile: the outer universe, un-shifted r
    if (!push_pad(c, make_pad(PAD_LATTICE, 1, f, uni), t3.nx, t3.ny, t3.nz)) return -1;
    uni = L.outer;
    return 0;
  }
  push_pad(c, make_pad(PAD_LATTICE, 0, f, uni), t3.nx, t3.ny, t3.nz);
  return -1;
}

// One step through a cell universe (cell_universe.cpp:72-109): 0 = descend into the fill universe, 1 = material cell
// found (cell), -1 = no cell here (lost)
template <class CUR>
__device__ __forceinline__ int descend_cells_step(const DevProblem& P, CUR& c, int& uni, int f, const V3& u, int& cell) {
  const abl_universe* U = P.universes + uni;
  const V3 r = frame_r(c, f);
  if (!push_pad(c, make_pad(PAD_UNIVERSE, 0, f, uni))) return -1;
  const int off = ldt(&U->cell_offset), n = ldt(&U->ncells);
  int found = -1;
  for (int k = 0; k < n; k++) {
    const int ci = ldt(&P.ucells[off + k]);
    if (cell_is_inside_fast(P, ci, r, u, c.token)) {
      found = ci;
      break;
    }
  }
  c.nf = f + 1;
  if (found < 0) return -1;
  if (!push_pad(c, make_pad(PAD_CELL, 0, f, found))) return -1;
  const int fill = ldt(&P.cells[found].fill_universe);
  if (fill < 0) {
    cell = found;
    return 1;
  }
  uni = fill;
  return 0;
}

// cursor_relocate for the lanes `mask` of a warp together (every lane of mask must call it).  The same per-lane
// operation sequence, but the lanes advance in step by universe TYPE: lattice steps while any lane stands at a lattice,
// then one cell-universe step for everybody.  Lanes re-descend from different depths (a changed pin cell: the cell
// universe only; a changed tile: the lattices above it first), and a plain per-lane loop made the warp run the
// cell-universe code once per distinct depth with a handful of lanes each (ncu: 5 of 32).
template <class CUR, bool HEX = true>
__device__ __forceinline__ void cursor_relocate_sync(const DevProblem& P, CUR& c, int from, const V3& u, unsigned mask) {
  int uni = P.root, f = 0;
  bool full = true;
  if (from > 0) {
    const int back = pad_info(c, from - 1);
    if (pad_type(back) != PAD_CELL) {
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
  bool active = true;
  int cell = -1;
  for (;;) {
    for (;;) {
      const bool at_lattice = active && ldt(&P.universes[uni].type) != ABL_UNI_CELLS;
      if (!__any_sync(mask, at_lattice)) break;
      int st = 0;
      if (at_lattice) st = descend_lattice_step<CUR, HEX>(P, c, uni, f, u);
      if (st < 0) {
        if (full) {
          active = false;
        } else {  // partial re-descent failed: full restart from the root
          full = true;
          c.np = 0;
          c.nf = 1;
          uni = P.root;
          f = 0;
        }
      }
    }
    if (!__any_sync(mask, active)) break;
    if (active) {
      const int st = descend_cells_step(P, c, uni, f, u, cell);
      if (st > 0) {
        active = false;
      } else if (st < 0) {
        if (full) {
          active = false;
        } else {
          full = true;
          c.np = 0;
          c.nf = 1;
          uni = P.root;
          f = 0;
        }
      }
    }
  }
  c.cell = cell;
  c.mat = cell >= 0 ? ldt(&P.cells[cell].material) : -1;
}


// Tracker::do_reflection (tracker.hpp:314-360), the arithmetic part: point on the surface and reflected direction
struct Reflected {
  V3 r, u;
};
static __device__ __noinline__ Reflected reflect_nl(const abl_surface* __restrict__ surfaces, int surface_index, const V3 r, const V3 u, double distance) {
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

// ---- CTA layout ---------------------------------------------------------------------------------------------------------
// One CTA of HK_THREADS threads per SM: HK_HIST history threads + HK_SERVICE_WARPS service warps.
//
// * The whole state of a history -- position, direction, weight, pcg32 state, geometry cursor -- lives in SHARED memory,
//   one column per thread ([field][thread]: conflict-free for any per-lane field index), not in registers.  Every stage of
//   the loop loads the handful of fields it works on and stores what it changed.  The loop is bound by the latency of
//   dependent fp64 chains (ncu of the register-resident version: issue slots 32 % busy with 4 warps per scheduler, 128
//   registers per thread filling the register file with 16 warps); with the state in shared memory a stage needs 64-80
//   registers and an SM holds 24-32 warps.  The column counts follow the problem: NF coordinate frames and NP pads are the
//   nesting depth of the geometry (C5G7: 3 and 4 -> 264 B per history), so deeper geometries simply hold fewer histories.
// * The history warps execute the staged loop in lock step (block-wide vote at the top): the SM instruction cache is much
//   smaller than the loop body, and in lock step a line fetched for one warp is a hit (or a hit under miss) for the others.
// * Rare, long events leave the lock-step instruction stream as JOBS for the service warps: (1) the boundary search of
//   a particle that left the geometry (full lookup at the pre-flight position, boundary-condition search, reflection),
//   request / response through the owner's idle cursor frames; (2) fission-site banking (sampling n_new sites and
//   appending them to the scratch bank), fire-and-forget: the owner skips its RNG stream ahead by the draws the sites
//   consume (a table jump) and goes on.  A job queue is single-consumer (history warp w posts to service warp
//   w % HK_SERVICE_WARPS) and lives in shared memory.
#ifndef HK_THREADS
#define HK_THREADS 768
#endif
#ifndef HK_MINBLOCKS
#define HK_MINBLOCKS 1
#endif
#ifndef HK_SERVICE_WARPS
#define HK_SERVICE_WARPS 1
#endif
#define HK_HIST (HK_THREADS - 32 * HK_SERVICE_WARPS)
#ifndef HK_STAGE_SYNC
#define HK_STAGE_SYNC 0
#endif
#ifndef HK_MATH
#define HK_MATH InlineMath
#endif
#ifndef HK_VOTE_EVERY
#define HK_VOTE_EVERY 1  // block-wide vote (and lock step) every n-th iteration
#endif
#ifndef HK_FQ
#define HK_FQ 64  // fission-job ring entries per service warp (a full ring makes the owner bank its sites inline)
#endif
#define HK_BQ 1024  // boundary-request ring entries per service warp (>= HK_HIST: one outstanding request per thread)
#define HK_BQ_EMPTY 0xffffu
static_assert(HK_HIST <= HK_BQ && HK_HIST < HK_BQ_EMPTY, "boundary ring must hold one request per history thread");
#define HK_MIN_FRAMES 3  // the boundary mailbox uses cursor frames 0..2

struct alignas(16) FisJob {  // 80 B
  double x, y, z, ux, uy, uz, w;
  uint64_t rng;
  uint32_t parent, daughter0;
  int32_t n_new, mg;
};

// Scores and event counters are summed per warp once per iteration (shuffle / ballot / redux) into shared memory, so no
// accumulator stays in registers across the loop; rare events add to CTA-wide shared-memory accumulators.
struct HAcc {  // scores of one loop iteration of one lane (summed per warp at the end of the iteration)
  double k_col, k_abs, mig;
  double k_trk;  // surface tracking only (surface_tracker.cpp:81-83)
};
struct ICount {  // events of one loop iteration of one lane
  unsigned real, virt, coll_scores, tl_bins;
};
enum { RC_SITES = 0, RC_BOUNDARY, RC_LOST, RC_N };
enum { WC_FLIGHTS = 0, WC_REAL, WC_VIRT, WC_TLBINS, WC_BOUNDARY, WC_COLLSCORES, WC_N };

struct HKFixed {  // the fixed part of the dynamic shared memory; the per-history columns follow it
  FisJob fq[HK_SERVICE_WARPS][HK_FQ];
  volatile unsigned fq_ready[HK_SERVICE_WARPS][HK_FQ];  // sequence number + 1 of the job whose payload is complete
  volatile unsigned short bq[HK_SERVICE_WARPS][HK_BQ];  // history-thread index of a requester, HK_BQ_EMPTY = not yet written
  unsigned bq_head[HK_SERVICE_WARPS], fq_head[HK_SERVICE_WARPS];
  volatile unsigned bq_tail[HK_SERVICE_WARPS], fq_tail[HK_SERVICE_WARPS];
  volatile int done;
  int groups_done;
  unsigned long long deadline_ns;  // %globaltimer value at which the watchdog fires
  volatile int abort;  // the watchdog fired: every loop of the CTA winds down (a kernel must never spin forever)
  unsigned rare[RC_N];
  double leak, leak_mig;  // Tallies::score_leak and the migration-area term of a leak (delta_tracker.cpp:233-238)
  double sd[HK_THREADS / 32][5];
  unsigned long long wcnt[HK_THREADS / 32][WC_N];
};
extern __shared__ __align__(16) unsigned char hk_shared_raw[];
#define HKS (*reinterpret_cast<HKFixed*>(hk_shared_raw))
#define HK_COLS_OFFSET ((unsigned)((sizeof(HKFixed) + 15) & ~size_t(15)))

// ---- the per-history columns -------------------------------------------------------------------------------------------------
// 8-byte columns: frame[a][f] (3 NF), then r u rb (3 each), w, d_coll, E, rng, ptile[i] (NP), [trace: hash];
// 4-byte columns: pinfo[i] (NP), then the HI_* fields below.  S = slots (histories) of the CTA.
#ifdef HK_TEST_ALIAS_RB  // (occupancy experiment only: the birth position shares the position's columns -- wrong migration area)
enum { HD_R = 0, HD_U = 3, HD_RB = 0, HD_W = 6, HD_DC = 7, HD_E = 8, HD_RNG = 9, HD_PT = 10 };
#else
enum { HD_R = 0, HD_U = 3, HD_RB = 6, HD_W = 9, HD_DC = 10, HD_E = 11, HD_RNG = 12, HD_PT = 13 };
#endif
enum { HI_IDX = 0, HI_DAU, HI_TOK, HI_CELL, HI_MAT, HI_HMAT, HI_G, HI_NPNF, HI_BDONE, HI_NSEC, HI_N, HI_NFL = HI_N, HI_NRE, HI_NVI, HI_N_TRACE };
__host__ __device__ inline unsigned hk_cols8(int NF, int NP, bool trace) { return 3u * NF + HD_PT + NP + (trace ? 1u : 0u); }
__host__ __device__ inline unsigned hk_cols4(int NP, bool trace) { return (unsigned)NP + (trace ? HI_N_TRACE : HI_N); }
__host__ __device__ inline unsigned hk_slot_bytes(int NF, int NP, bool trace) { return 8u * hk_cols8(NF, NP, trace) + 4u * hk_cols4(NP, trace); }

struct Cols {  // byte offsets from hk_shared_raw of one thread's entries
  unsigned d0;  // frame[0][0]
  unsigned r0;  // r.x (the 8-byte fields HD_* follow at stride S8)
  unsigned i0;  // pinfo[0]
  unsigned j0;  // HI_IDX (the 4-byte fields HI_* follow at stride S4)
  unsigned S8, S4, F8;  // column strides: 8 S, 4 S, 8 S NF
};
// base = byte offset of the columns from hk_shared_raw (fixed part + staged tables), t = the history's slot
__device__ __forceinline__ Cols make_cols_at(unsigned base, int t, int S, int NF, int NP, bool trace) {
  Cols q;
  q.S8 = 8u * S;
  q.S4 = 4u * S;
  q.F8 = q.S8 * NF;
  q.d0 = base + 8u * t;
  q.r0 = q.d0 + 3u * q.F8;
  q.i0 = base + hk_cols8(NF, NP, trace) * q.S8 + 4u * t;
  q.j0 = q.i0 + q.S4 * NP;
  return q;
}
__device__ __forceinline__ Cols make_cols(int t, int S, int NF, int NP, bool trace, int tables) {
  return make_cols_at(HK_COLS_OFFSET + (unsigned)tables, t, S, NF, NP, trace);
}
#define HK_D(q, k) (*reinterpret_cast<double*>(hk_shared_raw + ((q).r0 + (unsigned)(k) * (q).S8)))
#define HK_U(q, k) (*reinterpret_cast<unsigned long long*>(hk_shared_raw + ((q).r0 + (unsigned)(k) * (q).S8)))
#define HK_I(q, k) (*reinterpret_cast<int*>(hk_shared_raw + ((q).j0 + (unsigned)(k) * (q).S4)))
#define HK_FR(q, a, f) (*reinterpret_cast<double*>(hk_shared_raw + ((q).d0 + (unsigned)(a) * (q).F8 + (unsigned)(f) * (q).S8)))
#define HK_PI(q, i) (*reinterpret_cast<int*>(hk_shared_raw + ((q).i0 + (unsigned)(i) * (q).S4)))
__device__ __forceinline__ V3 hk_ld3(const Cols& q, int k) { return V3{HK_D(q, k), HK_D(q, k + 1), HK_D(q, k + 2)}; }
__device__ __forceinline__ void hk_st3(const Cols& q, int k, const V3& v) {
  HK_D(q, k) = v.x;
  HK_D(q, k + 1) = v.y;
  HK_D(q, k + 2) = v.z;
}

// the shared-memory cursor: scalars in registers for the length of a stage, arrays in the thread's columns
struct SCursor {
  int token, cell, mat, np, nf, err;
  Cols q;
};
__device__ __forceinline__ void cursor_load(SCursor& c, const Cols& q) {
  c.q = q;
  c.err = 0;
  c.token = HK_I(q, HI_TOK);
  const int pf = HK_I(q, HI_NPNF);
  c.np = pf & 0xff;
  c.nf = pf >> 8;
  c.cell = HK_I(q, HI_CELL);
  c.mat = HK_I(q, HI_MAT);
}
__device__ __forceinline__ void cursor_store(const SCursor& c) {
  HK_I(c.q, HI_TOK) = c.token;
  HK_I(c.q, HI_NPNF) = c.np | (c.nf << 8);
  HK_I(c.q, HI_CELL) = c.cell;
  HK_I(c.q, HI_MAT) = c.mat;
}
__device__ __forceinline__ V3 frame_r(const SCursor& c, int f) { return {HK_FR(c.q, 0, f), HK_FR(c.q, 1, f), HK_FR(c.q, 2, f)}; }
__device__ __forceinline__ void set_frame(SCursor& c, int f, double x, double y, double z) {
  HK_FR(c.q, 0, f) = x;
  HK_FR(c.q, 1, f) = y;
  HK_FR(c.q, 2, f) = z;
}
__device__ __forceinline__ void shift_frame(SCursor& c, int f, double dx, double dy, double dz) {
  HK_FR(c.q, 0, f) = HK_FR(c.q, 0, f) + dx;
  HK_FR(c.q, 1, f) = HK_FR(c.q, 1, f) + dy;
  HK_FR(c.q, 2, f) = HK_FR(c.q, 2, f) + dz;
}
__device__ __forceinline__ int pad_info(const SCursor& c, int i) { return HK_PI(c.q, i); }
// tile indices are compared for equality only; 21 bits per axis (two's complement) cover +-10^6 tiles
__device__ __forceinline__ unsigned long long pack_tile(int nx, int ny, int nz) {
  return ((unsigned long long)((unsigned)nx & 0x1fffffu)) | ((unsigned long long)((unsigned)ny & 0x1fffffu) << 21) |
         ((unsigned long long)((unsigned)nz & 0x1fffffu) << 42);
}
__device__ __forceinline__ void store_pad(SCursor& c, int i, int info, int tx, int ty, int tz) {
  HK_PI(c.q, i) = info;
  HK_U(c.q, HD_PT + i) = pack_tile(tx, ty, tz);
}
__device__ __forceinline__ bool pad_tile_is(const SCursor& c, int i, int nx, int ny, int nz) {
  return HK_U(c.q, HD_PT + i) == pack_tile(nx, ny, nz);
}
__device__ __forceinline__ int unpack_tile_field(unsigned long long v) {
  const int x = (int)(v & 0x1fffffu);
  return (x ^ 0x100000) - 0x100000;  // sign-extend 21 bits
}
__device__ __forceinline__ Tile3 pad_tile3(const SCursor& c, int i) {
  const unsigned long long p = HK_U(c.q, HD_PT + i);
  return Tile3{unpack_tile_field(p), unpack_tile_field(p >> 21), unpack_tile_field(p >> 42)};
}
// Every track-length tally (tallies.hpp:57-63) over the first d of the flight that starts at the history's position
// (R / U columns: the pre-move position); returns the number of bins scored.  A real call that reads the history from its
// columns: a handful of arguments, so the call site keeps nothing alive around it.
struct TleArgs {
  const DevTally* tallies;
  const double *Et, *Ea, *Ef, *Es;
  int ntallies;
};
static __device__ __noinline__ int score_flight_cols(const TleArgs T, const Cols q, int mg, double d) {
  const MatXS mx{ldt(&T.Et[mg]), ldt(&T.Ea[mg]), ldt(&T.Ef[mg]), ldt(&T.Es[mg])};
  const V3 r = hk_ld3(q, HD_R), u = hk_ld3(q, HD_U);
  const double E = HK_D(q, HD_E), w = HK_D(q, HD_W);
  int nb = 0;
  for (int t = 0; t < T.ntallies; t++) {
    if (ldt(&T.tallies[t].estimator) != ABL_EST_TRACK_LENGTH) continue;
    const DevTally Y = T.tallies[t];
    nb += score_flight(Y, r, u, d, E, w, 0., mx);
  }
  return nb;
}
__device__ __forceinline__ TleArgs tle_args(const DevProblem& P) { return TleArgs{P.tally_dev, P.Et, P.Ea, P.Ef, P.Es, P.ntallies}; }

// Tracker::get_nearest_boundary (tracker.hpp:163-225) on the shared-memory cursor, one copy per kernel
static __device__ __noinline__ Boundary cursor_nearest_boundary_s(const GeoTables G, const SCursor c, const V3 u) {
  return cursor_nearest_boundary(G, c, u);
}
// ... with the boundary-condition search skipped when bc_floor allows it (geom.cuh: cursor_nearest_boundary_lazy)
static __device__ __noinline__ Boundary cursor_nearest_boundary_floor(const GeoTables G, const SCursor c, const V3 u, double bc_floor) {
  return cursor_nearest_boundary_lazy(G, c, u, bc_floor);
}
// a lower bound of the distance from the global position r to every boundary-condition surface, shrunk by the tolerances
// the comparison needs (|u| is 1 to rounding; candidates within BOUNDRY_TOL of each other are ties); -1: no bound
__device__ __forceinline__ double bc_floor_of(const DevProblem& P, const V3& r) {
  if (P.n_bc_planes == 0) return -1.;
  double lo = ABL_INF;
  for (int k = 0; k < P.n_bc_planes; k++) {
    const double rc = P.bc_axis[k] == 0 ? r.x : (P.bc_axis[k] == 1 ? r.y : r.z);
    lo = fmin(lo, fabs(P.bc_p0[k] - rc));
  }
  return lo * (1. - 1e-9) - 2. * ABL_BOUNDRY_TOL;
}

__device__ __forceinline__ unsigned long long hk_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// barrier 1: the history threads only (the service warps never join it)
__device__ __forceinline__ void hist_sync() { asm volatile("bar.sync 1, %0;" ::"n"(HK_HIST) : "memory"); }
// the vote of ONE group of history warps (HK_VOTE_GROUPS > 1: the history warps keep lock step within a group only -- fewer
// warps wait for the slowest one; the groups drift apart and share fewer instruction-cache lines)
#ifndef HK_VOTE_GROUPS
#define HK_VOTE_GROUPS 1
#endif
__device__ __forceinline__ bool hist_all_group(bool pred, int barrier_id, int nthreads) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.and.pred q, %2, %3, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"((int)pred), "r"(barrier_id), "r"(nthreads)
      : "memory");
  return r != 0;
}
__device__ __forceinline__ bool hist_all(bool pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.and.pred q, 1, %2, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"((int)pred), "n"(HK_HIST)
      : "memory");
  return r != 0;
}
#if HK_STAGE_SYNC
#define HK_SYNC() hist_sync()
#else
#define HK_SYNC() __syncwarp()
#endif

// ---- jobs: the service side -----------------------------------------------------------------------------------------------
// What the service warps read, passed by value: the service loop is a real function call (its register needs stay out of
// the history threads' allocation), and handing it the whole kernel parameter block would force a local-memory copy.
struct ServiceArgs {
  GeoTables G;
  FissionTables ft;
  const double *nud, *nu;
  Site* sites;
  unsigned long long* n_sites;
  uint64_t site_capacity;
  int slots, nf, np, trace, tables;
};

// boundary request of the history thread with columns q: Tracker::restart_get_current at the pre-flight position +
// get_boundary_condition (delta_tracker.cpp:120-127) + the arithmetic of do_reflection (tracker.hpp:314-360).
// Mailbox = the requester's idle cursor frames: request r -> frame 0, u -> frame 1; response: reflected r -> frame 0,
// reflected u -> frame 1, (distance, surface | type, error | cell) -> frame 2.
__device__ __forceinline__ void serve_boundary(const GeoTables& G, const Cols q) {
  const V3 r{HK_FR(q, 0, 0), HK_FR(q, 1, 0), HK_FR(q, 2, 0)};
  const V3 u{HK_FR(q, 0, 1), HK_FR(q, 1, 1), HK_FR(q, 2, 1)};
  Cursor lc;
  lc.err = 0;
  lc.token = 0;
  cursor_restart_nl(G, lc, r, u);
  const Boundary b = cursor_boundary_condition_nl(G, lc, u);
  if (b.btype == ABL_BC_REFLECTIVE && b.surface_index >= 0) {
    const Reflected rf = reflect_nl(G.surfaces, b.surface_index, r, u, b.distance);
    HK_FR(q, 0, 0) = rf.r.x; HK_FR(q, 1, 0) = rf.r.y; HK_FR(q, 2, 0) = rf.r.z;
    HK_FR(q, 0, 1) = rf.u.x; HK_FR(q, 1, 1) = rf.u.y; HK_FR(q, 2, 1) = rf.u.z;
  }
  HK_FR(q, 0, 2) = b.distance;
  HK_FR(q, 1, 2) = __hiloint2double(b.surface_index, b.btype);
  HK_FR(q, 2, 2) = __hiloint2double(lc.err, lc.cell);
  __threadfence_block();
  *(volatile int*)&HK_I(q, HI_BDONE) = 1;
}

template <class M>
__device__ __forceinline__ void serve_fission(const ServiceArgs& X, const FisJob j) {
  uint64_t rng = j.rng;
  const int mat = j.mg / X.ft.G;
  bank_fission_sites<M>(X.ft, X.sites, X.n_sites, X.site_capacity, rng, V3{j.x, j.y, j.z}, V3{j.ux, j.uy, j.uz}, j.w, j.parent, j.daughter0,
                        j.n_new, mat, j.mg, ldt(&X.nud[j.mg]) / ldt(&X.nu[j.mg]));
}

template <class M>
static __device__ __noinline__ void service_loop(const ServiceArgs X, int sw) {
  HKFixed& S = HKS;
  const int lane = threadIdx.x & 31;
  for (;;) {
    bool worked = false;
    {  // boundary requests (head is read by one lane: the batch size must be warp-uniform)
      const unsigned tail = S.bq_tail[sw];
      const unsigned head = __shfl_sync(0xffffffffu, *(volatile unsigned*)&S.bq_head[sw], 0);
      const unsigned n = min(32u, head - tail);
      if (n) {
        worked = true;
        if ((unsigned)lane < n) {
          const unsigned slot = (tail + lane) % HK_BQ;
          unsigned t;
          while ((t = S.bq[sw][slot]) == HK_BQ_EMPTY && !S.abort) {}
          if (t != HK_BQ_EMPTY) {
            S.bq[sw][slot] = HK_BQ_EMPTY;
            __threadfence_block();
            serve_boundary(X.G, make_cols((int)t, X.slots, X.nf, X.np, X.trace != 0, X.tables));
          }
        }
        __syncwarp();
        if (lane == 0) S.bq_tail[sw] = tail + n;
        __syncwarp();
      }
    }
    {  // fission jobs
      const unsigned tail = S.fq_tail[sw];
      const unsigned head = __shfl_sync(0xffffffffu, *(volatile unsigned*)&S.fq_head[sw], 0);
      const unsigned n = min(32u, head - tail);
      if (n) {
        worked = true;
        FisJob j;
        if ((unsigned)lane < n) {
          const unsigned seq = tail + lane, slot = seq % HK_FQ;
          while (S.fq_ready[sw][slot] != seq + 1 && !S.abort) {}
          __threadfence_block();
          j = S.fq[sw][slot];
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) S.fq_tail[sw] = tail + n;  // the slots may be reused: the payloads are in registers
        __syncwarp();
        if ((unsigned)lane < n && !S.abort) serve_fission<M>(X, j);
        __syncwarp();
      }
    }
    if (!worked) {
      // (done is set after the last history of the CTA has finished, so nothing can be posted after it reads 1)
      const int fin = S.abort || (S.done && *(volatile unsigned*)&S.bq_head[sw] == S.bq_tail[sw] && *(volatile unsigned*)&S.fq_head[sw] == S.fq_tail[sw]);
      if (__shfl_sync(0xffffffffu, fin, 0)) break;
      __nanosleep(256);
    }
  }
}

// ---- jobs: the history side -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void post_boundary_request(const Cols& q, int t, const V3& r, const V3& u, int sw) {
  HKFixed& S = HKS;
  HK_FR(q, 0, 0) = r.x; HK_FR(q, 1, 0) = r.y; HK_FR(q, 2, 0) = r.z;
  HK_FR(q, 0, 1) = u.x; HK_FR(q, 1, 1) = u.y; HK_FR(q, 2, 1) = u.z;
  __threadfence_block();
  const unsigned slot = atomicAdd(&S.bq_head[sw], 1u) % HK_BQ;
  S.bq[sw][slot] = (unsigned short)t;
}

// (a full ring makes the poster wait: the service warps never wait for a history warp, so the ring always drains)
__device__ __forceinline__ void post_fission_job(const FisJob& j, int sw) {
  HKFixed& S = HKS;
  unsigned seq;
  for (;;) {
    seq = *(volatile unsigned*)&S.fq_head[sw];
    if (S.abort) return;
    if (seq - S.fq_tail[sw] >= HK_FQ) continue;
    if (atomicCAS(&S.fq_head[sw], seq, seq + 1) == seq) break;
  }
  const unsigned slot = seq % HK_FQ;
  S.fq[sw][slot] = j;
  __threadfence_block();
  S.fq_ready[sw][slot] = seq + 1;
}

// the RNG draws n_new x MGNuclide::sample_fission consumes (mg_nuclide.cpp:504-543): [chi: 1 if G >= 2] + mu + phi +
// delayed test [+ 1 for the delayed family when the neutron is delayed and there are >= 2 families]; 2 engine steps each
template <class M>
__device__ __forceinline__ uint64_t skip_fission_draws(const DevProblem& P, uint64_t rng, int n_new, int mat, int mg) {
  const unsigned per_site = 2u * ((P.G >= 2 ? 1u : 0u) + 3u);
  const int ndg = ldt(&P.dg_off[mat + 1]) - ldt(&P.dg_off[mat]);
  if (ndg < 2) return pcg_advance(rng, (uint64_t)per_site * (unsigned)n_new, P.jump);
  const double P_delayed = ldt(&P.nud[mg]) / ldt(&P.nu[mg]);
  for (int i = 0; i < n_new; i++) {
    rng = pcg_advance(rng, per_site - 2u, P.jump);
    if (M::rand(rng) < P_delayed) rng = pcg_advance(rng, 2u, P.jump);
  }
  return rng;
}

enum { PH_WAIT = PH_LOST };  // a lost particle waits for the service warp's boundary response

// the per-history event hash of the traced kernels (transport.cuh: note)
template <bool TRACE>
__device__ __forceinline__ void note_col(const Cols& q, int NP, uint64_t v) {
  if (TRACE) HK_U(q, HD_PT + NP) = (HK_U(q, HD_PT + NP) ^ v) * 1099511628211ULL;
}

// Transporter::collision + branching_collision (transporter.cpp:60-93,269-312), k-eigenvalue branch: the same operation
// sequence as collision<> in transport.cuh on the history's shared-memory columns -- every field is loaded where it is
// first needed and stored when it is final, so the stage needs few registers -- with the fission sites handed to a service
// warp.  r = the collision site (already in the R column), hmat = the material there.
// POST: where the fission sites of a collision go -- post(P, A, job, sw) hands the job over (the staged kernel: to a service
// warp; the event kernel: to its job ring), sites(n) counts them.
struct ServicePost {
  static __device__ __forceinline__ void post(const DevProblem&, const RunArgs&, const FisJob& j, int sw) { post_fission_job(j, sw); }
  static __device__ __forceinline__ void sites(unsigned n) { atomicAdd(&HKS.rare[RC_SITES], n); }
};
template <class M, bool TRACE, class POST = ServicePost>
__device__ __forceinline__ void collision_cols(const DevProblem& P, const RunArgs& A, const Cols& q, const V3 r, int hmat, HAcc& acc,
                                               ICount& ic, int sw, bool& alive) {
  const int gi = HK_I(q, HI_G);
  const int g = gi & 0xff;
  const int mg = hmat * P.G + g;
  const double Et = ldt(&P.Et[mg]), Ea = ldt(&P.Ea[mg]), Ef = ldt(&P.Ef[mg]), nu = ldt(&P.nu[mg]);
  double w = HK_D(q, HD_W);
  ic.real++;
  if (TRACE) HK_I(q, HI_NRE) = HK_I(q, HI_NRE) + 1;
  if (A.converged && P.n_coll_tallies) {
    const MatXS mx{Et, Ea, Ef, ldt(&P.Es[mg])};
    for (int t = 0; t < P.ntallies; t++)
      if (P.tally[t].estimator == ABL_EST_COLLISION) {
        const int l = (gi >> 8) ? ldt(&P.tally_gbin[t * P.G + g]) : tally_energy_bin(P.tally[t], HK_D(q, HD_E));
        ic.coll_scores += score_collision_pre(P.tally[t], r, l, w, 0., mx, ldt(&P.inv_score[t * (P.M * P.G) + mg]));
      }
  }
  {
    const double k_col_scr = ddiv_pos<M>(w * (nu * Ef), Et);
    const V3 rb = hk_ld3(q, HD_RB);
    const V3 dr{r.x - rb.x, r.y - rb.y, r.z - rb.z};
    const double mig_dist = norm3<M>(dr);
    const double mig_area_scr = ddiv_pos<M>(w * Ea, Et) * mig_dist * mig_dist;
    acc.k_col += k_col_scr;
    acc.mig += mig_area_scr;
  }
  uint64_t rng = HK_U(q, HD_RNG);
  (void)M::rand(rng);  // MaterialHelper::sample_nuclide always draws (material_helper.hpp:189)
  const double k_abs_scr = ddiv_pos<M>(w * nu * Ef, Et);
  acc.k_abs += k_abs_scr;
  const int n_new = (int)floor(ddiv_pos<M>(fabs(k_abs_scr), A.k_col) + M::rand(rng));  // transporter.cpp:358-487
  if (n_new > 0) {
    const V3 u = hk_ld3(q, HD_U);
    const uint32_t idx = (uint32_t)HK_I(q, HI_IDX), daughter = (uint32_t)HK_I(q, HI_DAU);
    FisJob j;
    j.x = r.x; j.y = r.y; j.z = r.z;
    j.ux = u.x; j.uy = u.y; j.uz = u.z;
    j.w = w;
    j.rng = rng;
    j.parent = idx;
    j.daughter0 = daughter;
    j.n_new = n_new;
    j.mg = mg;
    POST::post(P, A, j, sw);
    rng = skip_fission_draws<M>(P, rng, n_new, hmat, mg);
    HK_I(q, HI_DAU) = (int)(daughter + (uint32_t)n_new);
    POST::sites((unsigned)n_new);
  }
  note_col<TRACE>(q, A.hk_np, 0x5000000000000000ULL | (uint64_t)(uint32_t)n_new);
  {
    const double surv = ldt(&P.surv_frac[mg]);  // 1 - Ea / Et: implicit capture (transporter.cpp:295-298)
    Hist hh;  // (wgt2 is zero throughout a k-eigenvalue run; its roulette still draws: transporter.cpp:47-54)
    hh.w = w * surv;
    hh.w2 = 0. * surv;
    hh.rng = rng;
    hh.alive = true;
    russian_roulette<false, M>(P, hh);
    w = hh.w;
    rng = hh.rng;
    alive = hh.alive;
  }
  int g_out = 0;
  if (alive) {  // MGNuclide::sample_scatter (mg_nuclide.cpp:442-461); the yield matrix is never applied
    int ei = 0;
    if (P.G >= 2) ei = rng_discrete<M>(rng, P.ps_cp + (size_t)mg * P.G, P.G);
    const double E_out = group_mid(P, ei);
    const double mu = sample_mu<M>(P, P.angle + (size_t)mg * P.G + ei, rng);
    const double phi = 2. * ABL_PI * M::rand(rng);
    hk_st3(q, HD_U, rotate_dir<M>(hk_ld3(q, HD_U), mu, phi));
    HK_D(q, HD_E) = E_out;
    HK_I(q, HI_G) = ei | 0x100;
    w = w * 1.;
    if (E_out < P.min_energy) alive = false;
    g_out = ei + 1;
  }
  HK_D(q, HD_W) = w;
  HK_U(q, HD_RNG) = rng;
  note_col<TRACE>(q, A.hk_np, 0x6000000000000000ULL | (alive ? (uint64_t)g_out : 0ULL));
}

// NFC / NPC / SC != 0: the column shape (frames, pads, slots) as compile-time constants -- the builds for the common nesting
// depth: every column access becomes [thread base + immediate] instead of a multiply-add on runtime strides (integer
// address arithmetic was 47 % of the executed instructions, IMAD alone 19 %).  0: taken from RunArgs.
template <int TRK, bool TRACE, bool TLE, int NFC = 0, int NPC = 0, int SC = 0>
__global__ void __launch_bounds__(HK_THREADS, HK_MINBLOCKS) history_kernel(const DevProblem P, const RunArgs A) {
  HKFixed& S = HKS;
  const int hk_slots = SC ? SC : A.hk_slots, hk_nf = NFC ? NFC : A.hk_nf, hk_np = NPC ? NPC : A.hk_np;
  constexpr bool HEX = NFC == 0;  // the fixed-shape builds are also the builds without the hexagonal-lattice branches
  const unsigned FULL = 0xffffffffu;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- set-up: queues empty, then everyone meets once ---------------------------------------------------------------------
  for (int i = threadIdx.x; i < HK_SERVICE_WARPS * HK_BQ; i += HK_THREADS) (&S.bq[0][0])[i] = HK_BQ_EMPTY;
  for (int i = threadIdx.x; i < HK_SERVICE_WARPS * HK_FQ; i += HK_THREADS) (&S.fq_ready[0][0])[i] = 0;
  for (int i = threadIdx.x; i < (HK_THREADS / 32) * WC_N; i += HK_THREADS) (&S.wcnt[0][0])[i] = 0;
  for (int i = threadIdx.x; i < (HK_THREADS / 32) * 5; i += HK_THREADS) (&S.sd[0][0])[i] = 0.;
  if (threadIdx.x < HK_SERVICE_WARPS) {
    S.bq_head[threadIdx.x] = S.fq_head[threadIdx.x] = 0;
    S.bq_tail[threadIdx.x] = S.fq_tail[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) {
    S.done = 0;
    S.groups_done = 0;
    S.abort = 0;
    S.deadline_ns = hk_now_ns() + A.timeout_ns;
    S.leak = S.leak_mig = 0.;
    for (int q = 0; q < RC_N; q++) S.rare[q] = 0;
  }
  // the table arena, staged behind the fixed part (the host already pointed P's tables at this copy)
  if (A.hk_tables) {
    const uint4* src = reinterpret_cast<const uint4*>(A.arena);
    uint4* dst = reinterpret_cast<uint4*>(hk_shared_raw + HK_COLS_OFFSET);
    for (int i = threadIdx.x; i < A.hk_tables / 16; i += HK_THREADS) dst[i] = src[i];
    if (threadIdx.x == 0 && (unsigned long long)(uintptr_t)(void*)hk_shared_raw != A.smem_generic_base) {
      raise_error(A, ABL_ERR_CUDA, 0);  // the shared window is not where the host assumed: the table pointers are wrong
      S.abort = 1;
    }
  }
  const Cols q = make_cols(threadIdx.x, hk_slots, hk_nf, hk_np, TRACE, A.hk_tables);
  if (threadIdx.x < hk_slots) *(volatile int*)&HK_I(q, HI_BDONE) = 0;
  __syncthreads();

  if (wid >= HK_HIST / 32) {
    ServiceArgs X;
    X.G = geo_tables(P);
    X.ft = FissionTables{P.chi_cp, P.dg_cp, P.dg_off, P.gmid, P.G};
    X.nud = P.nud;
    X.nu = P.nu;
    X.sites = A.sites;
    X.n_sites = A.n_sites;
    X.site_capacity = A.site_capacity;
    X.slots = A.hk_slots;
    X.nf = A.hk_nf;
    X.np = A.hk_np;
    X.trace = TRACE ? 1 : 0;
    X.tables = A.hk_tables;
    service_loop<HK_MATH>(X, wid - HK_HIST / 32);
  } else {
    const uint32_t tid = blockIdx.x * HK_HIST + threadIdx.x;  // index among the history threads of the grid
    const uint32_t nthreads = gridDim.x * HK_HIST;
    const int sw = wid % HK_SERVICE_WARPS;
    int phase = PH_DEAD;
    int need = -1;  // pending (re-)descent: first bad pad, 0 = full lookup from the root, -1 = none
    bool exhausted = (int)threadIdx.x >= hk_slots;
#if HK_VOTE_GROUPS > 1
    constexpr int NHW = HK_HIST / 32;  // history warps, split into HK_VOTE_GROUPS runs of consecutive warps
    const int vgrp = wid * HK_VOTE_GROUPS / NHW;
    const int vgrp_first = (vgrp * NHW + HK_VOTE_GROUPS - 1) / HK_VOTE_GROUPS;
    const int vgrp_threads = 32 * ((((vgrp + 1) * NHW + HK_VOTE_GROUPS - 1) / HK_VOTE_GROUPS) - vgrp_first);
#endif  // threads beyond the CTA's slots (deep geometries) own no history
    bool have_ticket = false;
    const uint64_t N = A.bank.n;
    // (TLE = false: a build without the track-length scorer, for runs that have no track-length tally to score)
    const bool tle = TLE && A.converged && P.n_tl_tallies;
    unsigned iter = 0;

    for (;;) {
      // ---- R: refill ---------------------------------------------------------------------------------------------
      // (with a streamed bank -- A.avail -- a lane that holds a ticket waits here until its row has been copied)
      if (phase == PH_DEAD && !exhausted) {
        if (!have_ticket) {
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
            have_ticket = true;
            HK_I(q, HI_IDX) = (int)(uint32_t)idx;
          }
        }
        if (have_ticket) {
          const uint32_t idx = (uint32_t)HK_I(q, HI_IDX);
          if (A.avail == nullptr || (unsigned long long)idx < *(const volatile unsigned long long*)A.avail) {
            have_ticket = false;
            // (L2 loads: with a streamed bank a cached L1 line could hold neighbouring rows from before they arrived)
            const V3 r{__ldcg(&A.bank.x[idx]), __ldcg(&A.bank.y[idx]), __ldcg(&A.bank.z[idx])};
            const V3 u{__ldcg(&A.bank.ux[idx]), __ldcg(&A.bank.uy[idx]), __ldcg(&A.bank.uz[idx])};
            const double E = __ldcg(&A.bank.E[idx]);
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
            need = 0;
            phase = PH_BIRTH;
          }
        }
      }
      // watchdog: a history loop that runs past the deadline (a bug, or a geometry that traps particles) is wound down with
      // ABL_ERR_TIMEOUT instead of spinning for ever -- every lane drops its history, the vote below then ends the CTA
      if ((++iter & 255u) == 0 && lane == 0 && hk_now_ns() > S.deadline_ns && !S.abort) {
        S.abort = 1;
        raise_error(A, ABL_ERR_TIMEOUT, 0);
      }
      if (S.abort) {
        phase = PH_DEAD;
        have_ticket = false;
        exhausted = true;
        need = -1;
      }
#if HK_VOTE_EVERY > 1
      if ((iter % HK_VOTE_EVERY) == 0)
#endif
#if HK_VOTE_GROUPS > 1
      if (hist_all_group(phase == PH_DEAD && !have_ticket, 1 + vgrp, vgrp_threads)) break;
#else
      if (hist_all(phase == PH_DEAD && !have_ticket)) break;
#endif

      HAcc acc;
      acc.k_col = acc.k_abs = acc.mig = acc.k_trk = 0.;
      bool alive = true;        // the history survived this iteration's events
      bool did_flight = false;  // a flight was sampled this iteration
      bool did_boundary = false;
      ICount ic{0u, 0u, 0u, 0u};
      // ---- M: sample the flight, move the cursor, re-validate its pads ----------------------------------------------------------
      bool collide_now = false;  // surface tracking: the flight ended in a collision inside the current cell
      if (TRK == ABL_TRACK_SURFACE) {
        // SurfaceTracker::transport loop body (surface_tracker.cpp:72-146): distance to collision against the nearest
        // boundary over all pads; the track-length estimators score the segment from the pre-move position
        if (phase == PH_FLIGHT) {
          SCursor c;
          cursor_load(c, q);
          uint64_t rng = HK_U(q, HD_RNG);
          const int g = HK_I(q, HI_G) & 0xff;
          const int mg = HK_I(q, HI_HMAT) * P.G + g;
          V3 r = hk_ld3(q, HD_R);
          V3 u = hk_ld3(q, HD_U);
          const double w = HK_D(q, HD_W);
          const double d_coll = rng_exponential<HK_MATH>(rng, ldt(&P.Et[mg]));
          HK_U(q, HD_RNG) = rng;
          const double bc_floor = bc_floor_of(P, r);
          const Boundary sb = P.n_bc_planes ? cursor_nearest_boundary_floor(geo_tables(P), c, u, bc_floor)
                                            : cursor_nearest_boundary_s(geo_tables(P), c, u);
          did_flight = true;
          if (TRACE) HK_I(q, HI_NFL) = HK_I(q, HI_NFL) + 1;
          const double d_min = fmin(d_coll, sb.distance);
          if (tle) ic.tl_bins += score_flight_cols(tle_args(P), q, mg, d_min);
          acc.k_trk += w * d_min * (ldt(&P.nu[mg]) * ldt(&P.Ef[mg]));
          if (sb.distance < d_coll || fabs(sb.distance - d_coll) < ABL_BOUNDRY_TOL) {
            did_boundary = true;
            if (sb.btype == ABL_BC_VACUUM) {
              if (TRACE) {
                Hist ht;
                ht.hash = HK_U(q, HD_PT + A.hk_np);
                note(ht, 0x3000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
                HK_U(q, HD_PT + A.hk_np) = ht.hash;
              }
              alive = false;  // leak (surface_tracker.cpp:95-99)
              const V3 rb = hk_ld3(q, HD_RB);
              const V3 d{r.x + sb.distance * u.x - rb.x, r.y + sb.distance * u.y - rb.y, r.z + sb.distance * u.z - rb.z};
              atomicAdd(&S.leak, w);
              atomicAdd(&S.leak_mig, leak_mig_score(w, d));
            } else if (sb.btype == ABL_BC_REFLECTIVE) {
              if (sb.surface_index < 0) {
                raise_error(A, ABL_ERR_LOST, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
                alive = false;
                if (TRK == ABL_TRACK_CARTER) HK_I(q, HI_NSEC) = 0;
              } else {  // Tracker::do_reflection (tracker.hpp:314-360); the full lookup happens in the L stage
                const Reflected rf = reflect_nl(P.surfaces, sb.surface_index, r, u, sb.distance);
                hk_st3(q, HD_U, rf.u);
                hk_st3(q, HD_R, rf.r);
                HK_I(q, HI_TOK) = 0;
                HK_FR(q, 0, 0) = rf.r.x;
                HK_FR(q, 1, 0) = rf.r.y;
                HK_FR(q, 2, 0) = rf.r.z;
                need = 0;
                phase = PH_REFLECTED;
              }
            } else {  // Tracker::cross_surface + get_current (tracker.hpp:227-231)
              cursor_move(c, sb.distance, u);
              c.token = -sb.token;
              HK_I(q, HI_TOK) = c.token;
              const int first_bad = cursor_validate(P, c, u);
              if (first_bad < c.np) need = first_bad;
              r.x = r.x + sb.distance * u.x;
              r.y = r.y + sb.distance * u.y;
              r.z = r.z + sb.distance * u.z;
              hk_st3(q, HD_R, r);
              phase = PH_CROSSED;
            }
          } else {
            r.x = r.x + d_coll * u.x;
            r.y = r.y + d_coll * u.y;
            r.z = r.z + d_coll * u.z;
            hk_st3(q, HD_R, r);
            cursor_move(c, d_coll, u);
            HK_I(q, HI_TOK) = 0;
            collide_now = true;
          }
        }
      } else if (phase == PH_FLIGHT) {
        uint64_t rng = HK_U(q, HD_RNG);
        const int g = HK_I(q, HI_G) & 0xff;
        const V3 u = hk_ld3(q, HD_U);
        const double d_coll = rng_exponential<HK_MATH>(rng, ldt(&P.smp[g]));
        HK_U(q, HD_RNG) = rng;
        HK_D(q, HD_DC) = d_coll;
        did_flight = true;
        if (TRACE) HK_I(q, HI_NFL) = HK_I(q, HI_NFL) + 1;
        SCursor c;
        c.q = q;
        c.err = 0;
        const int pf = HK_I(q, HI_NPNF);
        c.np = pf & 0xff;
        c.nf = pf >> 8;
        cursor_move(c, d_coll, u);
        HK_I(q, HI_TOK) = 0;
        const int first_bad = cursor_validate<SCursor, HEX>(P, c, u);
        if (first_bad < c.np) need = first_bad;
      }
      HK_SYNC();

      // ---- L: (re-)descent through the universe tree -----------------------------------------------------------------------------------
      const unsigned relocating = __ballot_sync(FULL, need >= 0);
      if (need >= 0) {
        SCursor c;
        cursor_load(c, q);
        const V3 u = hk_ld3(q, HD_U);
#ifdef HK_PER_LANE_RELOCATE
        cursor_relocate(P, c, need, u);
#else
        cursor_relocate_sync<SCursor, HEX>(P, c, need, u, relocating);
#endif
        need = -1;
        cursor_store(c);
        if (c.err) raise_error(A, c.err, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
      }
      HK_SYNC();

      // ---- after the move.  Three sections, each loading the columns it works on: the rare phase bookkeeping (boundary
      // events, births, reflections ...), the hot one (arrive, real or virtual collision), the end of a history. ----------------------
      const int ccell = phase != PH_DEAD ? HK_I(q, HI_CELL) : -1;
      bool flight_done = false;  // carter tracking: a flight ended in a collision or a reflection (the split test follows)

      // ---- B: boundary events and the checks that follow a full lookup -------------------------------------------------------------
      if (TRK != ABL_TRACK_SURFACE && phase == PH_FLIGHT && did_flight && ccell < 0) {
        // left the geometry: a service warp looks for the boundary from the pre-flight position
        post_boundary_request(q, threadIdx.x, hk_ld3(q, HD_R), hk_ld3(q, HD_U), sw);
        phase = PH_WAIT;
      } else if (TRK != ABL_TRACK_SURFACE && phase == PH_WAIT) {
        if (*(volatile int*)&HK_I(q, HI_BDONE)) {
          __threadfence_block();
          *(volatile int*)&HK_I(q, HI_BDONE) = 0;
          const uint32_t idx = (uint32_t)HK_I(q, HI_IDX);
          Boundary bound{ABL_INF, -1, ABL_BC_NORMAL, 0};
          bound.distance = HK_FR(q, 0, 2);
          const double pk = HK_FR(q, 1, 2), pk2 = HK_FR(q, 2, 2);
          bound.surface_index = __double2hiint(pk);
          bound.btype = __double2loint(pk);
          const int cell_before = __double2loint(pk2);
          const int err = __double2hiint(pk2);
          if (err) raise_error(A, err, A.bank.id_a[idx]);
          const V3 r = hk_ld3(q, HD_R), u = hk_ld3(q, HD_U);  // (the pre-flight position: the flight never happened)
          const double w = HK_D(q, HD_W);
          if (tle) {  // delta_tracker.cpp:133: scored from the pre-move position over min(d_coll, boundary distance)
            const int mg = HK_I(q, HI_HMAT) * P.G + (HK_I(q, HI_G) & 0xff);
            ic.tl_bins += score_flight_cols(tle_args(P), q, mg, fmin(HK_D(q, HD_DC), bound.distance));
          }
          atomicAdd(&S.rare[RC_BOUNDARY], 1u);
          if (bound.btype == ABL_BC_VACUUM) {
            note_col<TRACE>(q, A.hk_np, 0x3000000000000000ULL | (uint64_t)(uint32_t)(cell_before + 1));
            // leak (delta_tracker.cpp:233-238): score_leak(w), mig_area += w * |r_exit - r_birth|^2
            alive = false;
            const V3 rb = hk_ld3(q, HD_RB);
            const V3 d{r.x + bound.distance * u.x - rb.x, r.y + bound.distance * u.y - rb.y, r.z + bound.distance * u.z - rb.z};
            atomicAdd(&S.leak, w);
            atomicAdd(&S.leak_mig, leak_mig_score(w, d));
            phase = PH_FLIGHT;
          } else if (bound.btype == ABL_BC_REFLECTIVE && bound.surface_index >= 0) {
            // Tracker::do_reflection (tracker.hpp:314-360); the full lookup happens in the next L stage.  The response
            // already sits in frame 0 (the reflected position) and frame 1 (the reflected direction).
            hk_st3(q, HD_R, V3{HK_FR(q, 0, 0), HK_FR(q, 1, 0), HK_FR(q, 2, 0)});
            hk_st3(q, HD_U, V3{HK_FR(q, 0, 1), HK_FR(q, 1, 1), HK_FR(q, 2, 1)});
            HK_I(q, HI_TOK) = 0;
            need = 0;
            phase = PH_REFLECTED;
          } else {
            raise_error(A, ABL_ERR_LOST, A.bank.id_a[idx]);
            alive = false;
            if (TRK == ABL_TRACK_CARTER) HK_I(q, HI_NSEC) = 0;
            phase = PH_FLIGHT;
          }
        }
      } else if (phase == PH_REFLECTED && need < 0) {
        if (ccell < 0) {
          raise_error(A, ABL_ERR_LOST, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
          alive = false;
          if (TRK == ABL_TRACK_CARTER) HK_I(q, HI_NSEC) = 0;
        } else {
          note_col<TRACE>(q, A.hk_np, 0x4000000000000000ULL | (uint64_t)(uint32_t)(ccell + 1));
        }
        flight_done = true;
        phase = PH_FLIGHT;
      } else if (TRK == ABL_TRACK_SURFACE && phase == PH_CROSSED && need < 0) {
        if (ccell < 0) {
          raise_error(A, ABL_ERR_LOST, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
          alive = false;
        } else {
          HK_I(q, HI_HMAT) = HK_I(q, HI_MAT);
          note_col<TRACE>(q, A.hk_np, 0x7000000000000000ULL | (uint64_t)(uint32_t)(ccell + 1));
        }
        phase = PH_FLIGHT;
      } else if (phase == PH_BIRTH) {
        if (ccell < 0) {  // lost at birth: warning + kill in the reference (delta_tracker.cpp:92-98)
          atomicAdd(&S.rare[RC_LOST], 1u);
          alive = false;
        } else {
          HK_I(q, HI_HMAT) = HK_I(q, HI_MAT);
        }
        phase = PH_FLIGHT;
      } else if (phase == PH_RESURRECT) {
        if (ccell < 0) {
          raise_error(A, ABL_ERR_LOST, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
          alive = false;
          if (TRK == ABL_TRACK_CARTER) HK_I(q, HI_NSEC) = 0;
        } else {
          HK_I(q, HI_HMAT) = HK_I(q, HI_MAT);
        }
        phase = PH_FLIGHT;
      } else if (TRK == ABL_TRACK_SURFACE) {
        // ---- C (surface tracking): the flight ended in a collision inside the current cell ----------------------------------------
        if (collide_now && alive) {
          note_col<TRACE>(q, A.hk_np, 0x2000000000000000ULL | (uint64_t)(uint32_t)(ccell + 1));
          collision_cols<HK_MATH, TRACE>(P, A, q, hk_ld3(q, HD_R), HK_I(q, HI_HMAT), acc, ic, sw, alive);
        }
      } else if (phase == PH_FLIGHT && did_flight) {
        // ---- T + C (delta / carter): track-length tallies from the pre-move position, then arrive: real or virtual collision
        // (delta_tracker.cpp:133,167-195, carter_tracker.cpp:189-207) ----------------------------------------------------------------------
        const double d_coll = HK_D(q, HD_DC);
        const int g = HK_I(q, HI_G) & 0xff;
        if (tle) ic.tl_bins += score_flight_cols(tle_args(P), q, HK_I(q, HI_HMAT) * P.G + g, d_coll);
        V3 r = hk_ld3(q, HD_R);
        {
          const V3 u = hk_ld3(q, HD_U);
          r.x = r.x + d_coll * u.x;
          r.y = r.y + d_coll * u.y;
          r.z = r.z + d_coll * u.z;
        }
        hk_st3(q, HD_R, r);
        flight_done = true;
        const int hmat = HK_I(q, HI_MAT);
        HK_I(q, HI_HMAT) = hmat;
        bool had_collision = false;
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
        if (alive) {
          note_col<TRACE>(q, A.hk_np, (had_collision ? 0x2000000000000000ULL : 0x1000000000000000ULL) | (uint64_t)(uint32_t)(ccell + 1));
          if (had_collision) {
            collision_cols<HK_MATH, TRACE>(P, A, q, r, hmat, acc, ic, sw, alive);
          } else {
            ic.virt++;
            if (TRACE) HK_I(q, HI_NVI) = HK_I(q, HI_NVI) + 1;
          }
        } else if (TRK == ABL_TRACK_CARTER) {
          HK_I(q, HI_NSEC) = 0;
        }
      }
      HK_SYNC();

      // ---- E: splitting (carter), secondaries, end of history ----------------------------------------------------------------------------
      if (TRK == ABL_TRACK_CARTER && flight_done && alive) {
        const double w = HK_D(q, HD_W);
        if (fabs(w) >= P.wgt_split) {  // Particle::split (particle.hpp:165-173)
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
              if (!push_secondary(A, h, u, E, h.w, h.w2, tid, nthreads)) {
                raise_error(A, ABL_ERR_BANK_OVERFLOW, A.bank.id_a[(uint32_t)HK_I(q, HI_IDX)]);
                break;
              }
            HK_I(q, HI_NSEC) = h.nsec;
          }
        }
      }
      if (phase == PH_FLIGHT && !alive) {
        if (TRK == ABL_TRACK_CARTER && HK_I(q, HI_NSEC) > 0) {  // Particle::resurect + Tracker restart (delta_tracker.cpp:197-229)
          Hist h;
          h.nsec = HK_I(q, HI_NSEC);
          pop_secondary(P, A, h, tid, nthreads);
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
          need = 0;
          phase = PH_RESURRECT;
        } else {
          const uint32_t idx = (uint32_t)HK_I(q, HI_IDX);
          A.nfis[idx] = (uint32_t)HK_I(q, HI_DAU);  // sites produced = daughters numbered
          if (TRACE) {
            A.tr_flights[idx] = (uint32_t)HK_I(q, HI_NFL);
            A.tr_real[idx] = (uint32_t)HK_I(q, HI_NRE);
            A.tr_virtual[idx] = (uint32_t)HK_I(q, HI_NVI);
            A.tr_hash[idx] = HK_U(q, HD_PT + A.hk_np);
            A.tr_rng[idx] = HK_U(q, HD_RNG);
          }
          phase = PH_DEAD;
        }
      }
      HK_SYNC();

      // ---- event counters of this iteration: one shared-memory update per warp --------------------------------------------------------
      {
        const unsigned nf = __popc(__ballot_sync(FULL, did_flight));
        const unsigned nr = __reduce_add_sync(FULL, ic.real);
        const unsigned nv = __reduce_add_sync(FULL, ic.virt);
        const unsigned ns = __reduce_add_sync(FULL, ic.coll_scores);
        unsigned nt = 0, nb = 0;
        if (tle) nt = __reduce_add_sync(FULL, ic.tl_bins);
        if (TRK == ABL_TRACK_SURFACE) nb = __popc(__ballot_sync(FULL, did_boundary));
        double s0 = acc.k_col, s1 = acc.k_abs, s2 = acc.mig, s3 = acc.k_trk;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s0 += __shfl_xor_sync(FULL, s0, o);
          s1 += __shfl_xor_sync(FULL, s1, o);
          s2 += __shfl_xor_sync(FULL, s2, o);
          if (TRK == ABL_TRACK_SURFACE) s3 += __shfl_xor_sync(FULL, s3, o);
        }
        if (lane == 0) {
          S.sd[wid][0] += s0;
          S.sd[wid][1] += s1;
          S.sd[wid][4] += s2;
          if (TRK == ABL_TRACK_SURFACE) S.sd[wid][2] += s3;
          S.wcnt[wid][WC_FLIGHTS] += nf;
          S.wcnt[wid][WC_REAL] += nr;
          S.wcnt[wid][WC_VIRT] += nv;
          S.wcnt[wid][WC_COLLSCORES] += ns;
          if (tle) S.wcnt[wid][WC_TLBINS] += nt;
          if (TRK == ABL_TRACK_SURFACE) S.wcnt[wid][WC_BOUNDARY] += nb;
        }
      }
    }
    // every history of this CTA is finished and nothing more will be posted: release the service warps
#if HK_VOTE_GROUPS > 1
    if ((int)threadIdx.x == 32 * vgrp_first && atomicAdd(&S.groups_done, 1) == HK_VOTE_GROUPS - 1) S.done = 1;
#else
    hist_sync();
    if (threadIdx.x == 0) S.done = 1;
#endif
  }

  // ---- the per-warp sums: one atomic per block --------------------------------------------------------------------------
  constexpr int NW = HK_THREADS / 32;
  __syncthreads();
  if (threadIdx.x < 5) {
    const int k = threadIdx.x;
    double v = k == 3 ? S.leak : (k == 4 ? S.leak_mig : 0.);
    for (int w = 0; w < NW; w++) v += S.sd[w][k];
    const int slot = k < 3 ? k : k + 1;  // scores layout: k_col,k_abs,k_trk,k_tot(unused),leak,mig
    atomicAdd(&A.scores[slot], v);
  } else if (threadIdx.x >= 32 && threadIdx.x < 40) {
    // counters layout: flights, real, virtual, tl_bins, sites, boundary, lost, coll_scores
    const int k = threadIdx.x - 32;
    unsigned long long v = k == 4 ? S.rare[RC_SITES] : (k == 5 ? S.rare[RC_BOUNDARY] : (k == 6 ? S.rare[RC_LOST] : 0u));
    const int wc = k == 0 ? WC_FLIGHTS : k == 1 ? WC_REAL : k == 2 ? WC_VIRT : k == 3 ? WC_TLBINS : k == 5 ? WC_BOUNDARY : k == 7 ? WC_COLLSCORES : -1;
    if (wc >= 0)
      for (int w = 0; w < HK_HIST / 32; w++) v += S.wcnt[w][wc];
    atomicAdd(&A.counters[k], v);
  }
}

}  // namespace abl
