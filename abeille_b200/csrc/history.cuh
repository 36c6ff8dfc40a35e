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
 * The geometry cursors live in shared memory, the rare long events run on a service warp; see "CTA layout" below.
 * DESIGN.md section 3.1 has the measurements behind each of these choices.
 */
#pragma once
#include "transport.cuh"

namespace abl {

enum { PH_DEAD = 0, PH_FLIGHT, PH_BIRTH, PH_LOST, PH_REFLECTED, PH_RESURRECT, PH_CROSSED };

// Tracker::get_current, first half (tracker.hpp:235-270): index of the first pad that no longer holds, or np
template <class CUR>
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
      const Tile3 t3 = lattice_tile_nl(P.universes + pad_index(info), frame_r(c, pad_frame(info)), u);
      if (!pad_tile_is(c, it, t3.nx, t3.ny, t3.nz)) {
        first_bad = it;
        break;
      }
    }
  }
  return first_bad;
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
  c.mat = c.cell >= 0 ? __ldg(&P.cells[c.cell].material) : -1;
}

// One step of Universe::get_cell through a lattice (rect_lattice.cpp:132-207): 0 = descend further (uni / f updated),
// -1 = no universe here (lost)
template <class CUR>
__device__ __forceinline__ int descend_lattice_step(const DevProblem& P, CUR& c, int& uni, int& f, const V3& u) {
  const abl_universe* U = P.universes + uni;
  const V3 r = frame_r(c, f);
  const Tile3 t3 = lattice_tile_nl(U, r, u);
  Lat L;
  L.Nx = __ldg(&U->N[0]); L.Ny = __ldg(&U->N[1]); L.Nz = __ldg(&U->N[2]);
  L.tile_offset = __ldg(&U->tile_offset);
  L.outer = __ldg(&U->outer);
  int sub = -1;
  if (tile_in_range(L, t3.nx, t3.ny, t3.nz)) sub = __ldg(&P.tiles[L.tile_offset + t3.nz * (L.Nx * L.Ny) + t3.nx * L.Ny + t3.ny]);
  c.nf = f + 1;
  if (sub >= 0) {
    if (!push_pad(c, make_pad(PAD_LATTICE, 0, f, uni), t3.nx, t3.ny, t3.nz)) return -1;
    if (f + 1 >= ABL_MAX_FRAMES) {
      c.err = ABL_ERR_GEOMETRY;
      return -1;
    }
    L.Px = __ldg(&U->P[0]); L.Py = __ldg(&U->P[1]); L.Pz = __ldg(&U->P[2]);
    L.Xl = __ldg(&U->Xl[0]); L.Yl = __ldg(&U->Xl[1]); L.Zl = __ldg(&U->Xl[2]);
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
  const int off = __ldg(&U->cell_offset), n = __ldg(&U->ncells);
  int found = -1;
  for (int k = 0; k < n; k++) {
    const int ci = __ldg(&P.ucells[off + k]);
    if (cell_is_inside_fast(P, ci, r, u, c.token)) {
      found = ci;
      break;
    }
  }
  c.nf = f + 1;
  if (found < 0) return -1;
  if (!push_pad(c, make_pad(PAD_CELL, 0, f, found))) return -1;
  const int fill = __ldg(&P.cells[found].fill_universe);
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
template <class CUR>
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
      const bool at_lattice = active && __ldg(&P.universes[uni].type) != ABL_UNI_CELLS;
      if (!__any_sync(mask, at_lattice)) break;
      int st = 0;
      if (at_lattice) st = descend_lattice_step(P, c, uni, f, u);
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
  c.mat = cell >= 0 ? __ldg(&P.cells[cell].material) : -1;
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

// every track-length tally (tallies.hpp:57-63); returns the number of bins scored
static __device__ __noinline__ int score_flight_all_nl(const DevTally* __restrict__ tallies, int ntallies, const V3 r, const V3 u, double d,
                                                double E, double w, double w2, const MatXS mx) {
  int nb = 0;
  for (int t = 0; t < ntallies; t++) {
    if (__ldg(&tallies[t].estimator) != ABL_EST_TRACK_LENGTH) continue;
    const DevTally T = tallies[t];
    nb += score_flight(T, r, u, d, E, w, w2, mx);
  }
  return nb;
}

// ---- CTA layout ---------------------------------------------------------------------------------------------------------
// One CTA of HK_THREADS threads per SM: HK_HIST history threads + HK_SERVICE_WARPS service warps.
//
// * The history warps execute the staged loop in lock step (named barrier 1 between the stages, block-wide vote at the
//   top): the SM instruction cache is much smaller than the loop body, and ncu showed a free-running version bound by
//   instruction-cache misses (sm__icc hit rate 57 %, 60 % of stall samples "no instruction").  In lock step a line
//   fetched for one warp is a hit (or a hit under miss) for the others.
// * Their geometry cursors live in SHARED memory, laid out [slot][thread] (conflict-free for any per-lane slot index).
//   In local memory the cursor's dynamically indexed arrays did not fit L1 (768 threads x 816 B of stack) and ncu showed
//   83 GB of DRAM write-back per 10^7 histories from evicted stack lines, 24 % L1 misses on local loads and
//   "long scoreboard" as the second stall reason.
// * Rare, long events leave the lock-step instruction stream as JOBS for the service warps: (1) the boundary search of
//   a particle that left the geometry (full lookup at the pre-flight position, boundary-condition search, reflection:
//   ~5000 warp instructions that v1 executed with 1.2 of 32 lanes active while the other 23 warps waited at the next
//   stage barrier), request / response through the owner's idle cursor frames; (2) fission-site banking (sampling
//   n_new sites and appending them to the scratch bank), fire-and-forget: the owner skips its RNG stream ahead by the
//   draws the sites consume (a table jump) and goes on.  A job queue is single-consumer (history warp w posts to
//   service warp w % HK_SERVICE_WARPS) and lives in shared memory.
#ifndef HK_THREADS
#define HK_THREADS 512
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
#ifndef HK_FQ
#define HK_FQ 64  // fission-job ring entries per service warp (a full ring makes the owner bank its sites inline)
#endif
#define HK_BQ 1024  // boundary-request ring entries per service warp (>= HK_HIST: one outstanding request per thread)
#define HK_BQ_EMPTY 0xffffu
static_assert(HK_HIST <= HK_BQ && HK_HIST < HK_BQ_EMPTY, "boundary ring must hold one request per history thread");
static_assert(ABL_MAX_FRAMES >= 5, "the boundary mailbox uses cursor frames 0..4");

struct alignas(16) FisJob {  // 80 B
  double x, y, z, ux, uy, uz, w;
  uint64_t rng;
  uint32_t parent, daughter0;
  int32_t n_new, mg;
};

// per-thread accumulators of the history threads (registers): what every flight / collision touches.  Rare events
// (sites banked, boundary events, lost-at-birth, leakage) add to CTA-wide shared-memory accumulators instead.
struct HAcc {
  double k_col, k_abs, mig;
  double k_trk;       // surface tracking only (surface_tracker.cpp:81-83)
  uint32_t flights, real, virt, tl_bins, coll_scores;
  uint32_t boundary;  // surface tracking only: a boundary event every few flights (delta / carter: rare, counted in shared memory)
};
enum { RC_SITES = 0, RC_BOUNDARY, RC_LOST, RC_N };

struct HKShared {
  double frame[3][ABL_MAX_FRAMES][HK_HIST];        // r_local of every cursor frame (x | y | z planes)
  unsigned long long ptile[ABL_MAX_PADS][HK_HIST];  // lattice pads: tile found at descent, 21 bits per axis
  int pinfo[ABL_MAX_PADS][HK_HIST];
  double rb[3][HK_HIST];                            // birth position (read once per collision: migration area)
  volatile int bdone[HK_HIST];                      // boundary response ready
  FisJob fq[HK_SERVICE_WARPS][HK_FQ];
  volatile unsigned fq_ready[HK_SERVICE_WARPS][HK_FQ];  // sequence number + 1 of the job whose payload is complete
  volatile unsigned short bq[HK_SERVICE_WARPS][HK_BQ];  // history-thread index of a requester, HK_BQ_EMPTY = not yet written
  unsigned bq_head[HK_SERVICE_WARPS], fq_head[HK_SERVICE_WARPS];
  volatile unsigned bq_tail[HK_SERVICE_WARPS], fq_tail[HK_SERVICE_WARPS];
  volatile int done;
  unsigned rare[RC_N];
  double leak, leak_mig;  // Tallies::score_leak and the migration-area term of a leak (delta_tracker.cpp:233-238)
  double sd[HK_THREADS / 32][5];
  unsigned long long sc[HK_THREADS / 32][8];
};
extern __shared__ __align__(16) unsigned char hk_shared_raw[];
#define HKS (*reinterpret_cast<HKShared*>(hk_shared_raw))

// the shared-memory cursor: scalars in registers, arrays in HKShared column t
struct SCursor {
  int token, cell, mat, np, nf, err;
  int t;
};
__device__ __forceinline__ V3 frame_r(const SCursor& c, int f) { return {HKS.frame[0][f][c.t], HKS.frame[1][f][c.t], HKS.frame[2][f][c.t]}; }
__device__ __forceinline__ void set_frame(SCursor& c, int f, double x, double y, double z) {
  HKS.frame[0][f][c.t] = x;
  HKS.frame[1][f][c.t] = y;
  HKS.frame[2][f][c.t] = z;
}
__device__ __forceinline__ void shift_frame(SCursor& c, int f, double dx, double dy, double dz) {
  HKS.frame[0][f][c.t] = HKS.frame[0][f][c.t] + dx;
  HKS.frame[1][f][c.t] = HKS.frame[1][f][c.t] + dy;
  HKS.frame[2][f][c.t] = HKS.frame[2][f][c.t] + dz;
}
__device__ __forceinline__ int pad_info(const SCursor& c, int i) { return HKS.pinfo[i][c.t]; }
// tile indices are compared for equality only; 21 bits per axis (two's complement) cover +-10^6 tiles
__device__ __forceinline__ unsigned long long pack_tile(int nx, int ny, int nz) {
  return ((unsigned long long)((unsigned)nx & 0x1fffffu)) | ((unsigned long long)((unsigned)ny & 0x1fffffu) << 21) |
         ((unsigned long long)((unsigned)nz & 0x1fffffu) << 42);
}
__device__ __forceinline__ void store_pad(SCursor& c, int i, int info, int tx, int ty, int tz) {
  HKS.pinfo[i][c.t] = info;
  HKS.ptile[i][c.t] = pack_tile(tx, ty, tz);
}
__device__ __forceinline__ bool pad_tile_is(const SCursor& c, int i, int nx, int ny, int nz) {
  return HKS.ptile[i][c.t] == pack_tile(nx, ny, nz);
}

__device__ __forceinline__ int unpack_tile_field(unsigned long long v) {
  const int x = (int)(v & 0x1fffffu);
  return (x ^ 0x100000) - 0x100000;  // sign-extend 21 bits
}
__device__ __forceinline__ Tile3 pad_tile3(const SCursor& c, int i) {
  const unsigned long long p = HKS.ptile[i][c.t];
  return Tile3{unpack_tile_field(p), unpack_tile_field(p >> 21), unpack_tile_field(p >> 42)};
}
// Tracker::get_nearest_boundary (tracker.hpp:163-225) on the shared-memory cursor, one copy per kernel
static __device__ __noinline__ Boundary cursor_nearest_boundary_s(const GeoTables G, const SCursor c, const V3 u) {
  return cursor_nearest_boundary(G, c, u);
}

// barrier 1: the history threads only (the service warps never join it)
__device__ __forceinline__ void hist_sync() { asm volatile("bar.sync 1, %0;" ::"n"(HK_HIST) : "memory"); }
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
// boundary request of history thread t: Tracker::restart_get_current at the pre-flight position + get_boundary_condition
// (delta_tracker.cpp:120-127) + the arithmetic of do_reflection (tracker.hpp:314-360)
__device__ __forceinline__ void serve_boundary(const DevProblem& P, int t) {
  HKShared& S = HKS;
  const V3 r{S.frame[0][0][t], S.frame[1][0][t], S.frame[2][0][t]};
  const V3 u{S.frame[0][1][t], S.frame[1][1][t], S.frame[2][1][t]};
  Cursor lc;
  lc.err = 0;
  lc.token = 0;
  cursor_restart(P, lc, r, u);
  const Boundary b = cursor_boundary_condition_nl(geo_tables(P), lc, u);
  if (b.btype == ABL_BC_REFLECTIVE && b.surface_index >= 0) {
    const Reflected rf = reflect_nl(P.surfaces, b.surface_index, r, u, b.distance);
    S.frame[0][2][t] = rf.r.x; S.frame[1][2][t] = rf.r.y; S.frame[2][2][t] = rf.r.z;
    S.frame[0][3][t] = rf.u.x; S.frame[1][3][t] = rf.u.y; S.frame[2][3][t] = rf.u.z;
  }
  S.frame[0][4][t] = b.distance;
  S.frame[1][4][t] = __hiloint2double(b.surface_index, b.btype);
  S.frame[2][4][t] = __hiloint2double(lc.err, lc.cell);
  __threadfence_block();
  S.bdone[t] = 1;
}

template <class M>
__device__ __forceinline__ void serve_fission(const DevProblem& P, const RunArgs& A, const FisJob j) {
  const FissionTables ft{P.chi_cp, P.dg_cp, P.dg_off, P.gmid, P.G};
  uint64_t rng = j.rng;
  const int mat = j.mg / P.G;
  bank_fission_sites<M>(ft, A.sites, A.n_sites, A.site_capacity, rng, V3{j.x, j.y, j.z}, V3{j.ux, j.uy, j.uz}, j.w, j.parent, j.daughter0,
                        j.n_new, mat, j.mg, __ldg(&P.nud[j.mg]) / __ldg(&P.nu[j.mg]));
}

template <class M>
__device__ __forceinline__ void service_loop(const DevProblem& P, const RunArgs& A, int sw) {
  HKShared& S = HKS;
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
          while ((t = S.bq[sw][slot]) == HK_BQ_EMPTY) {}
          S.bq[sw][slot] = HK_BQ_EMPTY;
          __threadfence_block();
          serve_boundary(P, (int)t);
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
          while (S.fq_ready[sw][slot] != seq + 1) {}
          __threadfence_block();
          j = S.fq[sw][slot];
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) S.fq_tail[sw] = tail + n;  // the slots may be reused: the payloads are in registers
        __syncwarp();
        if ((unsigned)lane < n) serve_fission<M>(P, A, j);
        __syncwarp();
      }
    }
    if (!worked) {
      // (done is set after the last history of the CTA has finished, so nothing can be posted after it reads 1)
      const int fin = S.done && *(volatile unsigned*)&S.bq_head[sw] == S.bq_tail[sw] && *(volatile unsigned*)&S.fq_head[sw] == S.fq_tail[sw];
      if (__shfl_sync(0xffffffffu, fin, 0)) break;
      __nanosleep(256);
    }
  }
}

// ---- jobs: the history side -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void post_boundary_request(SCursor& c, const V3& r, const V3& u, int sw) {
  HKShared& S = HKS;
  set_frame(c, 0, r.x, r.y, r.z);
  set_frame(c, 1, u.x, u.y, u.z);
  __threadfence_block();
  const unsigned slot = atomicAdd(&S.bq_head[sw], 1u) % HK_BQ;
  S.bq[sw][slot] = (unsigned short)c.t;
}

// false when the ring is full (the caller then banks the sites itself)
__device__ __forceinline__ bool post_fission_job(const FisJob& j, int sw) {
  HKShared& S = HKS;
  unsigned seq;
  for (;;) {
    seq = *(volatile unsigned*)&S.fq_head[sw];
    if (seq - S.fq_tail[sw] >= HK_FQ) return false;
    if (atomicCAS(&S.fq_head[sw], seq, seq + 1) == seq) break;
  }
  const unsigned slot = seq % HK_FQ;
  S.fq[sw][slot] = j;
  __threadfence_block();
  S.fq_ready[sw][slot] = seq + 1;
  return true;
}

// the RNG draws n_new x MGNuclide::sample_fission consumes (mg_nuclide.cpp:504-543): [chi: 1 if G >= 2] + mu + phi +
// delayed test [+ 1 for the delayed family when the neutron is delayed and there are >= 2 families]; 2 engine steps each
template <class M>
__device__ __forceinline__ uint64_t skip_fission_draws(const DevProblem& P, uint64_t rng, int n_new, int mat, int mg) {
  const unsigned per_site = 2u * ((P.G >= 2 ? 1u : 0u) + 3u);
  const int ndg = __ldg(&P.dg_off[mat + 1]) - __ldg(&P.dg_off[mat]);
  if (ndg < 2) return pcg_advance(rng, (uint64_t)per_site * (unsigned)n_new, P.jump);
  const double P_delayed = __ldg(&P.nud[mg]) / __ldg(&P.nu[mg]);
  for (int i = 0; i < n_new; i++) {
    rng = pcg_advance(rng, per_site - 2u, P.jump);
    if (M::rand(rng) < P_delayed) rng = pcg_advance(rng, 2u, P.jump);
  }
  return rng;
}

// Transporter::collision + branching_collision (transporter.cpp:60-93,269-312), k-eigenvalue branch: the same operation
// sequence as collision<> in transport.cuh, with the fission sites handed to a service warp
template <class M, bool TRACE>
__device__ __forceinline__ void collision_hk(const DevProblem& P, const RunArgs& A, Hist& h, HAcc& acc, int sw, int ht) {
  const int mg = h.mat * P.G + h.g;
  const double Et = __ldg(&P.Et[mg]), Ea = __ldg(&P.Ea[mg]), Ef = __ldg(&P.Ef[mg]), nu = __ldg(&P.nu[mg]);
  acc.real++;
  if (TRACE) h.n_real++;
  if (A.converged && P.n_coll_tallies) {
    const MatXS mx{Et, Ea, Ef, __ldg(&P.Es[mg])};
    for (int t = 0; t < P.ntallies; t++)
      if (P.tally[t].estimator == ABL_EST_COLLISION) {
        const int l = h.emid ? __ldg(&P.tally_gbin[t * P.G + h.g]) : tally_energy_bin(P.tally[t], h.E);
        acc.coll_scores += score_collision_pre(P.tally[t], h.r, l, h.w, h.w2, mx, __ldg(&P.inv_score[t * (P.M * P.G) + mg]));
      }
  }
  {
    const double k_col_scr = ddiv_pos<M>(h.w * (nu * Ef), Et);
    const V3 dr{h.r.x - HKS.rb[0][ht], h.r.y - HKS.rb[1][ht], h.r.z - HKS.rb[2][ht]};
    const double mig_dist = norm3<M>(dr);
    const double mig_area_scr = ddiv_pos<M>(h.w * Ea, Et) * mig_dist * mig_dist;
    acc.k_col += k_col_scr;
    acc.mig += mig_area_scr;
  }
  (void)M::rand(h.rng);  // MaterialHelper::sample_nuclide always draws (material_helper.hpp:189)
  const double k_abs_scr = ddiv_pos<M>(h.w * nu * Ef, Et);
  acc.k_abs += k_abs_scr;
  const int n_new = (int)floor(ddiv_pos<M>(fabs(k_abs_scr), A.k_col) + M::rand(h.rng));  // transporter.cpp:358-487
  if (n_new > 0) {
    FisJob j;
    j.x = h.r.x; j.y = h.r.y; j.z = h.r.z;
    j.ux = h.u.x; j.uy = h.u.y; j.uz = h.u.z;
    j.w = h.w;
    j.rng = h.rng;
    j.parent = h.idx;
    j.daughter0 = h.daughter;
    j.n_new = n_new;
    j.mg = mg;
    if (post_fission_job(j, sw)) {
      h.rng = skip_fission_draws<M>(P, h.rng, n_new, h.mat, mg);
    } else {
      const FissionTables ft{P.chi_cp, P.dg_cp, P.dg_off, P.gmid, P.G};
      bank_fission_sites<M>(ft, A.sites, A.n_sites, A.site_capacity, h.rng, h.r, h.u, h.w, h.idx, h.daughter, n_new, h.mat, mg,
                            __ldg(&P.nud[mg]) / nu);
    }
    h.daughter += (uint32_t)n_new;
    atomicAdd(&HKS.rare[RC_SITES], (unsigned)n_new);
  }
  if (TRACE) note(h, 0x5000000000000000ULL | (uint64_t)(uint32_t)n_new);
  const double surv = __ldg(&P.surv_frac[mg]);  // 1 - Ea / Et: implicit capture (transporter.cpp:295-298)
  h.w = h.w * surv;
  h.w2 = h.w2 * surv;
  russian_roulette<false, M>(P, h);
  if (h.alive) {  // MGNuclide::sample_scatter (mg_nuclide.cpp:442-461); the yield matrix is never applied
    int ei = 0;
    if (P.G >= 2) ei = rng_discrete<M>(h.rng, P.ps_cp + (size_t)mg * P.G, P.G);
    const double E_out = group_mid(P, ei);
    const double mu = sample_mu<M>(P, P.angle + (size_t)mg * P.G + ei, h.rng);
    const double phi = 2. * ABL_PI * M::rand(h.rng);
    h.u = rotate_dir<M>(h.u, mu, phi);
    h.E = E_out;
    h.g = ei;
    h.emid = true;
    h.w = h.w * 1.;
    h.w2 = h.w2 * 1.;
    if (h.E < P.min_energy) h.alive = false;
  }
  if (TRACE) note(h, 0x6000000000000000ULL | (h.alive ? (uint64_t)(h.g + 1) : 0ULL));
}

enum { PH_WAIT = PH_LOST };  // v2: a lost particle waits for the service warp's boundary response

template <int TRK, bool TRACE>
__global__ void __launch_bounds__(HK_THREADS, 1) history_kernel(const DevProblem P, const RunArgs A) {
  HKShared& S = HKS;
  const unsigned FULL = 0xffffffffu;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- set-up: queues empty, then everyone meets once ---------------------------------------------------------------------
  for (int i = threadIdx.x; i < HK_SERVICE_WARPS * HK_BQ; i += HK_THREADS) (&S.bq[0][0])[i] = HK_BQ_EMPTY;
  for (int i = threadIdx.x; i < HK_SERVICE_WARPS * HK_FQ; i += HK_THREADS) (&S.fq_ready[0][0])[i] = 0;
  for (int i = threadIdx.x; i < HK_HIST; i += HK_THREADS) S.bdone[i] = 0;
  if (threadIdx.x < HK_SERVICE_WARPS) {
    S.bq_head[threadIdx.x] = S.fq_head[threadIdx.x] = 0;
    S.bq_tail[threadIdx.x] = S.fq_tail[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) {
    S.done = 0;
    S.leak = S.leak_mig = 0.;
    for (int q = 0; q < RC_N; q++) S.rare[q] = 0;
  }
  __syncthreads();

  HAcc acc;
  acc.k_col = acc.k_abs = acc.mig = acc.k_trk = 0.;
  acc.flights = acc.real = acc.virt = acc.tl_bins = acc.coll_scores = acc.boundary = 0;

  if (wid >= HK_HIST / 32) {
    service_loop<HK_MATH>(P, A, wid - HK_HIST / 32);
  } else {
    const uint32_t tid = blockIdx.x * HK_HIST + threadIdx.x;  // index among the history threads of the grid
    const uint32_t nthreads = gridDim.x * HK_HIST;
    const int sw = wid % HK_SERVICE_WARPS;
    Hist h;
    h.alive = false;
    h.nsec = 0;
    h.idx = 0;
    SCursor c;
    c.err = 0;
    c.np = 0;
    c.nf = 1;
    c.token = 0;
    c.cell = c.mat = -1;
    c.t = threadIdx.x;
    int phase = PH_DEAD;
    int need = -1;        // pending (re-)descent: first bad pad, 0 = full lookup from the root, -1 = none
    double d_coll = 0.;   // sampled flight distance (kept across the iterations of a boundary event)
    bool exhausted = false, have_ticket = false;
    const uint64_t N = A.bank.n;
    const bool tle = A.converged && P.n_tl_tallies;

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
            h.idx = (uint32_t)idx;
          }
        }
        if (have_ticket && (A.avail == nullptr || (unsigned long long)h.idx < *(const volatile unsigned long long*)A.avail)) {
          const uint32_t idx = h.idx;
          have_ticket = false;
          // (L2 loads: with a streamed bank a cached L1 line could hold neighbouring rows from before they arrived)
          h.r = {__ldcg(&A.bank.x[idx]), __ldcg(&A.bank.y[idx]), __ldcg(&A.bank.z[idx])};
          h.u = {__ldcg(&A.bank.ux[idx]), __ldcg(&A.bank.uy[idx]), __ldcg(&A.bank.uz[idx])};
          S.rb[0][c.t] = h.r.x;
          S.rb[1][c.t] = h.r.y;
          S.rb[2][c.t] = h.r.z;
          h.E = __ldcg(&A.bank.E[idx]);
          h.w = __ldcg(&A.bank.wgt[idx]);
          h.w2 = 0.;
          h.g = group_of(P, h.E);
          h.emid = h.g < P.G && h.E == group_mid(P, h.g);
          h.rng = __ldcg(&A.bank.id_c[idx]);  // pcg32 state: seeded by seed_streams_kernel / source sampling
          h.hash = 1469598103934665603ULL;
          h.daughter = 0;
          h.n_flights = h.n_real = h.n_virtual = 0;
          h.nsec = 0;
          h.alive = true;
          c.token = 0;
          set_frame(c, 0, h.r.x, h.r.y, h.r.z);
          need = 0;
          phase = PH_BIRTH;
        }
      }
      if (hist_all(phase == PH_DEAD && !have_ticket)) break;

      // ---- M: sample the flight, move the cursor, re-validate its pads ----------------------------------------------------------
      bool collide_now = false;  // surface tracking: the flight ended in a collision inside the current cell
      if (TRK == ABL_TRACK_SURFACE) {
        // SurfaceTracker::transport loop body (surface_tracker.cpp:72-146): distance to collision against the nearest
        // boundary over all pads; the track-length estimators score the segment from the pre-move position
        if (phase == PH_FLIGHT) {
          const int mg = h.mat * P.G + h.g;
          d_coll = rng_exponential<HK_MATH>(h.rng, __ldg(&P.Et[mg]));
          const Boundary sb = cursor_nearest_boundary_s(geo_tables(P), c, h.u);
          acc.flights++;
          if (TRACE) h.n_flights++;
          const double d_min = fmin(d_coll, sb.distance);
          if (tle) {
            const MatXS mx{__ldg(&P.Et[mg]), __ldg(&P.Ea[mg]), __ldg(&P.Ef[mg]), __ldg(&P.Es[mg])};
            acc.tl_bins += score_flight_all_nl(P.tally_dev, P.ntallies, h.r, h.u, d_min, h.E, h.w, h.w2, mx);
          }
          acc.k_trk += h.w * d_min * (__ldg(&P.nu[mg]) * __ldg(&P.Ef[mg]));
          if (sb.distance < d_coll || fabs(sb.distance - d_coll) < ABL_BOUNDRY_TOL) {
            acc.boundary++;
            if (sb.btype == ABL_BC_VACUUM) {
              if (TRACE) note(h, 0x3000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
              h.alive = false;  // leak (surface_tracker.cpp:95-99)
              const V3 d{h.r.x + sb.distance * h.u.x - S.rb[0][c.t], h.r.y + sb.distance * h.u.y - S.rb[1][c.t],
                         h.r.z + sb.distance * h.u.z - S.rb[2][c.t]};
              atomicAdd(&S.leak, h.w);
              atomicAdd(&S.leak_mig, leak_mig_score(h.w, d));
            } else if (sb.btype == ABL_BC_REFLECTIVE) {
              if (sb.surface_index < 0) {
                raise_error(A, ABL_ERR_LOST, A.bank.id_a[h.idx]);
                h.alive = false;
                h.nsec = 0;
              } else {  // Tracker::do_reflection (tracker.hpp:314-360); the full lookup happens in the L stage
                const Reflected rf = reflect_nl(P.surfaces, sb.surface_index, h.r, h.u, sb.distance);
                h.u = rf.u;
                h.r = rf.r;
                c.token = 0;
                set_frame(c, 0, h.r.x, h.r.y, h.r.z);
                need = 0;
                phase = PH_REFLECTED;
              }
            } else {  // Tracker::cross_surface + get_current (tracker.hpp:227-231)
              cursor_move(c, sb.distance, h.u);
              c.token = -sb.token;
              const int first_bad = cursor_validate(P, c, h.u);
              if (first_bad < c.np) need = first_bad;
              h.r.x = h.r.x + sb.distance * h.u.x;
              h.r.y = h.r.y + sb.distance * h.u.y;
              h.r.z = h.r.z + sb.distance * h.u.z;
              phase = PH_CROSSED;
            }
          } else {
            h.r.x = h.r.x + d_coll * h.u.x;
            h.r.y = h.r.y + d_coll * h.u.y;
            h.r.z = h.r.z + d_coll * h.u.z;
            cursor_move(c, d_coll, h.u);
            collide_now = true;
          }
        }
      } else if (phase == PH_FLIGHT) {
        d_coll = rng_exponential<HK_MATH>(h.rng, __ldg(&P.smp[h.g]));
        acc.flights++;
        if (TRACE) h.n_flights++;
        cursor_move(c, d_coll, h.u);
        const int first_bad = cursor_validate(P, c, h.u);
        if (first_bad < c.np) need = first_bad;
      }
      HK_SYNC();

      // ---- L: (re-)descent through the universe tree -----------------------------------------------------------------------------------
      const unsigned relocating = __ballot_sync(FULL, need >= 0);
      if (need >= 0) {
#ifdef HK_PER_LANE_RELOCATE
        cursor_relocate(P, c, need, h.u);
#else
        cursor_relocate_sync(P, c, need, h.u, relocating);
#endif
        need = -1;
        if (c.err) {
          raise_error(A, c.err, A.bank.id_a[h.idx]);
          c.err = 0;
        }
      }
      HK_SYNC();

      // ---- B1: what did the move do? -------------------------------------------------------------------------------------------------------
      double tle_d = -1.;
      bool test_collision = false, flight_done = false, answered = false;
      Boundary bound{ABL_INF, -1, ABL_BC_NORMAL, 0};
      int cell_before = -1;
      if (TRK == ABL_TRACK_SURFACE) {
        // (the boundary was found before the move: nothing to ask the service warp)
      } else if (phase == PH_FLIGHT) {
        if (c.cell < 0) {  // left the geometry: a service warp looks for the boundary from the pre-flight position
          post_boundary_request(c, h.r, h.u, sw);
          phase = PH_WAIT;
        } else {
          tle_d = d_coll;
          test_collision = true;
        }
      } else if (phase == PH_WAIT) {
        if (S.bdone[c.t]) {
          __threadfence_block();
          S.bdone[c.t] = 0;
          answered = true;
          bound.distance = S.frame[0][4][c.t];
          const double pk = S.frame[1][4][c.t], pk2 = S.frame[2][4][c.t];
          bound.surface_index = __double2hiint(pk);
          bound.btype = __double2loint(pk);
          cell_before = __double2loint(pk2);
          const int err = __double2hiint(pk2);
          if (err) raise_error(A, err, A.bank.id_a[h.idx]);
          tle_d = fmin(d_coll, bound.distance);  // delta_tracker.cpp:133: scored from the pre-move position
        }
      }
      // ---- T: track-length mesh tallies, scored from the pre-move position -----------------------------------------------------------------
      if (TRK != ABL_TRACK_SURFACE && tle && tle_d >= 0.) {
        const int mg = h.mat * P.G + h.g;
        const MatXS mx{__ldg(&P.Et[mg]), __ldg(&P.Ea[mg]), __ldg(&P.Ef[mg]), __ldg(&P.Es[mg])};
        acc.tl_bins += score_flight_all_nl(P.tally_dev, P.ntallies, h.r, h.u, tle_d, h.E, h.w, h.w2, mx);
      }

      // ---- B2: boundary events and the checks that follow a full lookup ---------------------------------------------------------------------------
      if (answered) {
        atomicAdd(&S.rare[RC_BOUNDARY], 1u);
        if (bound.btype == ABL_BC_VACUUM) {
          if (TRACE) note(h, 0x3000000000000000ULL | (uint64_t)(uint32_t)(cell_before + 1));
          {  // leak (delta_tracker.cpp:233-238): score_leak(w), mig_area += w * |r_exit - r_birth|^2
            h.alive = false;
            const V3 d{h.r.x + bound.distance * h.u.x - S.rb[0][c.t], h.r.y + bound.distance * h.u.y - S.rb[1][c.t],
                       h.r.z + bound.distance * h.u.z - S.rb[2][c.t]};
            atomicAdd(&S.leak, h.w);
            atomicAdd(&S.leak_mig, leak_mig_score(h.w, d));
          }
          phase = PH_FLIGHT;
        } else if (bound.btype == ABL_BC_REFLECTIVE && bound.surface_index >= 0) {
          // Tracker::do_reflection (tracker.hpp:314-360); the full lookup happens in the next L stage
          h.r = V3{S.frame[0][2][c.t], S.frame[1][2][c.t], S.frame[2][2][c.t]};
          h.u = V3{S.frame[0][3][c.t], S.frame[1][3][c.t], S.frame[2][3][c.t]};
          c.token = 0;
          set_frame(c, 0, h.r.x, h.r.y, h.r.z);
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
      } else if (TRK == ABL_TRACK_SURFACE && phase == PH_CROSSED && need < 0) {
        if (c.cell < 0) {
          raise_error(A, ABL_ERR_LOST, A.bank.id_a[h.idx]);
          h.alive = false;
          h.nsec = 0;
        } else {
          h.mat = c.mat;
          if (TRACE) note(h, 0x7000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
        }
        phase = PH_FLIGHT;
      } else if (phase == PH_BIRTH) {
        if (c.cell < 0) {  // lost at birth: warning + kill in the reference (delta_tracker.cpp:92-98)
          atomicAdd(&S.rare[RC_LOST], 1u);
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
      if (TRK == ABL_TRACK_SURFACE) {
        if (collide_now && h.alive) {
          if (TRACE) note(h, 0x2000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
          collision_hk<HK_MATH, TRACE>(P, A, h, acc, sw, c.t);
        }
      } else if (test_collision) {
        bool had_collision = false;
        flight_done = true;
        h.r.x = h.r.x + d_coll * h.u.x;
        h.r.y = h.r.y + d_coll * h.u.y;
        h.r.z = h.r.z + d_coll * h.u.z;
        h.mat = c.mat;
        const double Esample = __ldg(&P.smp[h.g]);
        const double Et = __ldg(&P.Et[h.mat * P.G + h.g]);
        const double real_frac = __ldg(&P.real_frac[h.mat * P.G + h.g]);  // Et / Esample
        if (TRK == ABL_TRACK_DELTA) {
          if (Et - Esample > 1.E-10) {
            raise_error(A, ABL_ERR_MAJORANT, A.bank.id_a[h.idx]);
            h.alive = false;
            h.nsec = 0;
          } else if (HK_MATH::rand(h.rng) < real_frac) {
            had_collision = true;
          }
        } else {
          if (Esample >= Et) {
            if (HK_MATH::rand(h.rng) < real_frac) had_collision = true;
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
            collision_hk<HK_MATH, TRACE>(P, A, h, acc, sw, c.t);
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
          set_frame(c, 0, h.r.x, h.r.y, h.r.z);
          need = 0;
          phase = PH_RESURRECT;
        } else {
          A.nfis[h.idx] = h.daughter;  // sites produced = daughters numbered
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
    // every history of this CTA is finished and nothing more will be posted: release the service warps
    hist_sync();
    if (threadIdx.x == 0) S.done = 1;
  }

  // ---- reduce the per-thread accumulators: warp shuffle, then one atomic per block ------------------------------------
  double dv[5] = {acc.k_col, acc.k_abs, TRK == ABL_TRACK_SURFACE ? acc.k_trk : 0., 0., acc.mig};
  unsigned long long cv[8] = {acc.flights, acc.real, acc.virt, acc.tl_bins, 0, TRK == ABL_TRACK_SURFACE ? acc.boundary : 0u, 0, acc.coll_scores};
  constexpr int NW = HK_THREADS / 32;
#pragma unroll
  for (int q = 0; q < 5; q++) {
    double v = dv[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    if (lane == 0) S.sd[wid][q] = v;
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    unsigned long long v = cv[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    if (lane == 0) S.sc[wid][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    const int q = threadIdx.x;
    double v = q == 3 ? S.leak : (q == 4 ? S.leak_mig : 0.);
    for (int w = 0; w < NW; w++) v += S.sd[w][q];
    const int slot = q < 3 ? q : q + 1;  // scores layout: k_col,k_abs,k_trk,k_tot(unused),leak,mig
    atomicAdd(&A.scores[slot], v);
  } else if (threadIdx.x >= 32 && threadIdx.x < 40) {
    const int q = threadIdx.x - 32;
    unsigned long long v = q == 4 ? S.rare[RC_SITES] : (q == 5 ? S.rare[RC_BOUNDARY] : (q == 6 ? S.rare[RC_LOST] : 0u));
    for (int w = 0; w < NW; w++) v += S.sc[w][q];
    atomicAdd(&A.counters[q], v);
  }
}

}  // namespace abl
