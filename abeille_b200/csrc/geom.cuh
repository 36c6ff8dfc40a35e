/* geom.cuh -- CSG geometry on flattened tables: surfaces, cells (RPN half-space logic), cell
 * universes, rectilinear lattices, and the per-history geometry cursor.
 *
 * Behaviour follows the reference (paths relative to the reference tree):
 *   surfaces   sign / distance / norm      src/xplane.cpp:32-56, src/plane.cpp:33-60, src/zcylinder.cpp:33-80,
 *                                          src/cylinder.cpp:60-110, src/sphere.cpp:33-80
 *   cells      is_inside, distances        src/cell.cpp:71-201
 *   universes  get_cell / boundary search  src/cell_universe.cpp:43-154, src/lattice.cpp:77-108
 *   lattices   tile math                   src/rect_lattice.cpp:33-52,132-310
 *   cursor     Tracker                     include/simulation/tracker.hpp:41-372
 * Design differences (results are bit-identical): no virtual dispatch, no id->index maps, no heap.
 * The cursor keeps the reference's stack of "lily pads", but pads that live in the same local
 * coordinate frame share one r_local (the reference stores and increments a copy per pad; copies in
 * one frame are always bitwise equal because they start equal and receive identical increments).
 * A frame changes only when a lattice tile is entered (r_local = r - tile_center).
 */
#pragma once
#include "tables.h"

namespace abl {

#define ABL_SURFACE_COINCIDENT 1E-12
#define ABL_BOUNDRY_TOL (500. * 1E-12)

struct V3 {
  double x, y, z;
};
__device__ __forceinline__ double dot3(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// Migration-area score of a leak, `p.wgt() * (r_leak - r_birth) * (r_leak - r_birth)` (src/delta_tracker.cpp:236-238,
// surface_tracker.cpp:91-92): C++ groups it as (w * d) . d -- the weight scales the vector first -- not w * (d . d).
__device__ __forceinline__ double leak_mig_score(double w, const V3& d) {
  return (d.x * w) * d.x + (d.y * w) * d.y + (d.z * w) * d.z;
}
template <class M = InlineMath>
__device__ __forceinline__ double norm3(const V3& a) { return M::sqrt_(a.x * a.x + a.y * a.y + a.z * a.z); }
// Direction constructors renormalise every time (include/utils/direction.hpp:37-43)
template <class M = InlineMath>
__device__ __forceinline__ V3 make_direction(double x, double y, double z) {
  const double m = M::sqrt_(x * x + y * y + z * z);
  return {M::div(x, m), M::div(y, m), M::div(z, m)};
}

// include/utils/direction.hpp:130-151
template <class M = InlineMath>
__device__ __forceinline__ V3 rotate_direction(const V3& u, double mu, double phi) {
  double sn, cs;
  det_sincos(phi, &sn, &cs);
  const double sqrt_mu = M::sqrt_(1. - mu * mu);
  const double sqrt_w = M::sqrt_(1. - u.z * u.z);
  double ux, uy, uz;
  if (sqrt_w > 1.E-10) {
    ux = mu * u.x + M::div(sqrt_mu * (u.x * u.z * cs - u.y * sn), sqrt_w);
    uy = mu * u.y + M::div(sqrt_mu * (u.y * u.z * cs + u.x * sn), sqrt_w);
    uz = mu * u.z - sqrt_mu * sqrt_w * cs;
  } else {
    const double sqrt_v = M::sqrt_(1. - u.y * u.y);
    ux = mu * u.x + M::div(sqrt_mu * (u.x * u.y * cs + u.z * sn), sqrt_v);
    uy = mu * u.y - sqrt_mu * sqrt_v * cs;
    uz = mu * u.z + M::div(sqrt_mu * (u.y * u.z * cs - u.x * sn), sqrt_v);
  }
  return make_direction<M>(ux, uy, uz);
}
// one shared copy for kernels that use CallMath (scatter in the hot loop, fission banking in the cold path)
static __device__ ABL_HOT_CALL V3 rotate_direction_call(const V3 u, double mu, double phi) { return rotate_direction<CallMath>(u, mu, phi); }
template <class M>
__device__ __forceinline__ V3 rotate_dir(const V3& u, double mu, double phi) { return rotate_direction<M>(u, mu, phi); }
template <>
__device__ __forceinline__ V3 rotate_dir<CallMath>(const V3& u, double mu, double phi) { return rotate_direction_call(u, mu, phi); }

// ---- surfaces ---------------------------------------------------------------------------------
struct Surf {  // register copy of one table row
  int type, bc;
  double p0, p1, p2, p3, p4, p5, p6;
};
template <class PT>
__device__ __forceinline__ Surf load_surface(const PT& P, int i) {
  const abl_surface* s = P.surfaces + i;
  Surf r;
  r.type = ldt(&s->type);
  r.bc = ldt(&s->bc);
  r.p0 = ldt(&s->p[0]);
  r.p1 = ldt(&s->p[1]);
  r.p2 = ldt(&s->p[2]);
  r.p3 = ldt(&s->p[3]);
  r.p4 = r.p5 = r.p6 = 0.;
  if (r.type == ABL_SURF_CYL) {
    r.p4 = ldt(&s->p[4]);
    r.p5 = ldt(&s->p[5]);
    r.p6 = ldt(&s->p[6]);
  }
  return r;
}

__device__ __forceinline__ double surf_eval(const Surf& s, const V3& r) {
  switch (s.type) {
    case ABL_SURF_XPLANE: return r.x - s.p0;
    case ABL_SURF_YPLANE: return r.y - s.p0;
    case ABL_SURF_ZPLANE: return r.z - s.p0;
    case ABL_SURF_PLANE: return s.p0 * r.x + s.p1 * r.y + s.p2 * r.z - s.p3;
    case ABL_SURF_XCYL: { const double y = r.y - s.p0, z = r.z - s.p1; return y * y + z * z - s.p2 * s.p2; }
    case ABL_SURF_YCYL: { const double x = r.x - s.p0, z = r.z - s.p1; return x * x + z * z - s.p2 * s.p2; }
    case ABL_SURF_ZCYL: { const double x = r.x - s.p0, y = r.y - s.p1; return y * y + x * x - s.p2 * s.p2; }
    case ABL_SURF_CYL: {
      const double x = r.x - s.p0, y = r.y - s.p1, z = r.z - s.p2;
      return s.p3 * x * x + s.p4 * y * y + s.p5 * z * z - s.p6 * s.p6;
    }
    default: {
      const double x = r.x - s.p0, y = r.y - s.p1, z = r.z - s.p2;
      return (x * x) + (y * y) + (z * z) - s.p3 * s.p3;
    }
  }
}

// Surface::norm: the gradient direction, normalised by the Direction constructor (direction.hpp:37-43).  One
// normalisation after the switch instead of one per case; for the axis planes the constructor yields (1,0,0) etc.
// exactly (sqrt(1) = 1, x / 1 = x), so they return at once.
static __device__ __noinline__ V3 surf_norm(const Surf& s, const V3& r) {
  double nx, ny, nz;
  switch (s.type) {
    case ABL_SURF_XPLANE: return V3{1., 0., 0.};
    case ABL_SURF_YPLANE: return V3{0., 1., 0.};
    case ABL_SURF_ZPLANE: return V3{0., 0., 1.};
    case ABL_SURF_PLANE: nx = s.p0; ny = s.p1; nz = s.p2; break;
    case ABL_SURF_XCYL: nx = 0.; ny = r.y - s.p0; nz = r.z - s.p1; break;
    case ABL_SURF_YCYL: nx = r.x - s.p0; ny = 0.; nz = r.z - s.p1; break;
    case ABL_SURF_ZCYL: nx = r.x - s.p0; ny = r.y - s.p1; nz = 0.; break;
    case ABL_SURF_CYL: nx = s.p3 * (r.x - s.p0); ny = s.p4 * (r.y - s.p1); nz = s.p5 * (r.z - s.p2); break;
    default: nx = r.x - s.p0; ny = r.y - s.p1; nz = r.z - s.p2; break;
  }
  return make_direction(nx, ny, nz);
}

// +1 / -1 : which side of the surface; on-surface ties broken by the flight direction
__device__ __forceinline__ int surf_sign(const Surf& s, const V3& r, const V3& u) {
  const double e = surf_eval(s, r);
  if (e > ABL_SURFACE_COINCIDENT) return 1;
  if (e < -ABL_SURFACE_COINCIDENT) return -1;
  const V3 n = surf_norm(s, r);
  if (dot3(u, n) > 0.) return 1;
  return -1;
}

__device__ __forceinline__ double quadric_distance(double a, double k, double c, bool on_surf) {
  const double quad = k * k - a * c;
  if (quad < 0.) return ABL_INF;
  if (on_surf || fabs(c) < ABL_SURFACE_COINCIDENT) {
    if (k >= 0.) return ABL_INF;
    return (-k + sqrt(quad)) / a;
  } else if (c < 0.) {
    return (-k + sqrt(quad)) / a;
  } else {
    const double d = (-k - sqrt(quad)) / a;
    if (d < 0.) return ABL_INF;
    return d;
  }
}

__device__ __forceinline__ double surf_distance(const Surf& s, const V3& r, const V3& u, bool on_surf) {
  switch (s.type) {
    case ABL_SURF_XPLANE:
    case ABL_SURF_YPLANE:
    case ABL_SURF_ZPLANE: {
      const double rc = s.type == ABL_SURF_XPLANE ? r.x : (s.type == ABL_SURF_YPLANE ? r.y : r.z);
      const double uc = s.type == ABL_SURF_XPLANE ? u.x : (s.type == ABL_SURF_YPLANE ? u.y : u.z);
      const double diff = s.p0 - rc;
      if (on_surf || fabs(diff) < ABL_SURFACE_COINCIDENT || uc == 0.) return ABL_INF;
      const double d = diff / uc;
      if (d < 0.) return ABL_INF;
      return d;
    }
    case ABL_SURF_PLANE: {
      const double num = s.p3 - s.p0 * r.x - s.p1 * r.y - s.p2 * r.z;
      const double denom = s.p0 * u.x + s.p1 * u.y + s.p2 * u.z;
      const double d = num / denom;
      if (on_surf || fabs(d) < ABL_SURFACE_COINCIDENT || denom == 0.) return ABL_INF;
      if (d < 0.) return ABL_INF;
      return d;
    }
    case ABL_SURF_XCYL: {
      const double a = u.y * u.y + u.z * u.z;
      if (a == 0.) return ABL_INF;
      const double y = r.y - s.p0, z = r.z - s.p1;
      const double k = y * u.y + z * u.z;
      const double c = y * y + z * z - s.p2 * s.p2;
      return quadric_distance(a, k, c, on_surf);
    }
    case ABL_SURF_YCYL: {
      const double a = u.x * u.x + u.z * u.z;
      if (a == 0.) return ABL_INF;
      const double x = r.x - s.p0, z = r.z - s.p1;
      const double k = x * u.x + z * u.z;
      const double c = x * x + z * z - s.p2 * s.p2;
      return quadric_distance(a, k, c, on_surf);
    }
    case ABL_SURF_ZCYL: {
      const double a = u.y * u.y + u.x * u.x;
      if (a == 0.) return ABL_INF;
      const double x = r.x - s.p0, y = r.y - s.p1;
      const double k = y * u.y + x * u.x;
      const double c = y * y + x * x - s.p2 * s.p2;
      return quadric_distance(a, k, c, on_surf);
    }
    case ABL_SURF_CYL: {
      const double a = s.p3 * u.x * u.x + s.p4 * u.y * u.y + s.p5 * u.z * u.z;
      if (a == 0.) return ABL_INF;
      const double x = r.x - s.p0, y = r.y - s.p1, z = r.z - s.p2;
      const double k = s.p3 * x * u.x + s.p4 * y * u.y + s.p5 * z * u.z;
      const double c = s.p3 * x * x + s.p4 * y * y + s.p5 * z * z - s.p6 * s.p6;
      return quadric_distance(a, k, c, on_surf);
    }
    default: {  // sphere: no division by a
      const double x = r.x - s.p0, y = r.y - s.p1, z = r.z - s.p2;
      const double k = x * u.x + y * u.y + z * u.z;
      const double c = x * x + y * y + z * z - s.p3 * s.p3;
      const double quad = k * k - c;
      if (quad < 0.) return ABL_INF;
      if (on_surf || fabs(c) < ABL_SURFACE_COINCIDENT) {
        if (k >= 0.) return ABL_INF;
        return -k + sqrt(quad);
      } else if (c < 0.) {
        return -k + sqrt(quad);
      } else {
        const double d = -k - sqrt(quad);
        if (d < 0.) return ABL_INF;
        return d;
      }
    }
  }
}

// the geometry tables alone (what a non-inlined helper needs; passing DevProblem by reference to a real
// function call would force a local-memory copy of the whole kernel parameter block)
struct GeoTables {
  const abl_surface* surfaces;
  const abl_cell* cells;
  const int32_t* rpn;
  const abl_universe* universes;
  const int32_t* ucells;
  const int32_t* tiles;
  int32_t root;
};
__device__ __forceinline__ GeoTables geo_tables(const DevProblem& P) {
  return GeoTables{P.surfaces, P.cells, P.rpn, P.universes, P.ucells, P.tiles, P.root};
}

// ---- cells -------------------------------------------------------------------------------------
__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }

template <class PT>
__device__ inline bool cell_is_inside(const PT& P, int ci, const V3& r, const V3& u, int on_surf) {
  const abl_cell* c = P.cells + ci;
  const int off = ldt(&c->rpn_offset), len = ldt(&c->rpn_len);
  if (ldt(&c->simple)) {
    for (int k = 0; k < len; k++) {
      const int token = ldt(&P.rpn[off + k]);
      if (token == on_surf) {
      } else if (-token == on_surf) {
        return false;
      } else {
        const Surf s = load_surface(P, iabs(token) - 1);
        const int sg = surf_sign(s, r, u);
        if ((sg > 0 && token < 0) || (sg < 0 && token > 0)) return false;
      }
    }
    return true;
  }
  // RPN evaluation; the boolean stack is a bit mask (host guarantees len <= 64)
  uint64_t stck = 0;
  int i_stck = -1;
  for (int k = 0; k < len; k++) {
    const int token = ldt(&P.rpn[off + k]);
    if (token == ABL_OP_UNION) {
      const bool v = ((stck >> (i_stck - 1)) & 1ULL) || ((stck >> i_stck) & 1ULL);
      i_stck--;
      stck = (stck & ~(1ULL << i_stck)) | ((uint64_t)v << i_stck);
    } else if (token == ABL_OP_INTERSECTION) {
      const bool v = ((stck >> (i_stck - 1)) & 1ULL) && ((stck >> i_stck) & 1ULL);
      i_stck--;
      stck = (stck & ~(1ULL << i_stck)) | ((uint64_t)v << i_stck);
    } else if (token == ABL_OP_COMPLEMENT) {
      stck ^= (1ULL << i_stck);
    } else {
      i_stck++;
      bool v;
      if (token == on_surf) {
        v = true;
      } else if (-token == on_surf) {
        v = false;
      } else {
        const Surf s = load_surface(P, iabs(token) - 1);
        const int sg = surf_sign(s, r, u);
        v = ((sg > 0 && token > 0) || (sg < 0 && token < 0));
      }
      stck = (stck & ~(1ULL << i_stck)) | ((uint64_t)v << i_stck);
    }
  }
  if (i_stck == 0) return (stck & 1ULL) != 0;
  return true;
}

// the generic evaluator as a real function call: keeps the cold path out of the hot loop's instruction footprint
static __device__ __noinline__ bool cell_is_inside_nl(const GeoTables G, int ci, const V3 r, const V3 u, int on_surf) {
  return cell_is_inside(G, ci, r, u, on_surf);
}

// Cell::is_inside through the compiled descriptor (tables.h: CellFast): 1 inside, 0 outside, -1 undecided --
// the particle sits on a surface or is within SURFACE_COINCIDENT of one (direction-dependent tie), or the
// region has no compiled form; the caller then runs the generic evaluator.  One shared copy per kernel.
static __device__ ABL_HOT_CALL int cell_fast_nl(const CellFast* __restrict__ cf, const V3 r, int on_surf) {
  const int kind = ldt(&cf->kind);
  if (on_surf != 0) return -1;
  if (kind == CF_BOX) {
    const double2 bx = ldt(reinterpret_cast<const double2*>(&cf->a[0]));
    const double2 by = ldt(reinterpret_cast<const double2*>(&cf->a[2]));
    const double2 bz = ldt(reinterpret_cast<const double2*>(&cf->a[4]));
    const double e0 = r.x - bx.x, e1 = r.x - bx.y, e2 = r.y - by.x, e3 = r.y - by.y, e4 = r.z - bz.x, e5 = r.z - bz.y;
    const double T = ABL_SURFACE_COINCIDENT;
    const bool tie = (fabs(e0) <= T) | (fabs(e1) <= T) | (fabs(e2) <= T) | (fabs(e3) <= T) | (fabs(e4) <= T) | (fabs(e5) <= T);
    const bool in = (e0 > T) & (e1 < -T) & (e2 > T) & (e3 < -T) & (e4 > T) & (e5 < -T);
    return tie ? -1 : (in ? 1 : 0);
  }
  if (kind == CF_ZCYL) {
    const double2 xy = ldt(reinterpret_cast<const double2*>(&cf->a[0]));
    const double r2 = ldt(&cf->a[2]);
    const double x = r.x - xy.x, y = r.y - xy.y;
    const double e = y * y + x * x - r2;
    const bool in = (ldt(&cf->sense) < 0) ? (e < 0.) : (e > 0.);
    return fabs(e) > ABL_SURFACE_COINCIDENT ? (in ? 1 : 0) : -1;
  }
  return -1;
}
__device__ __forceinline__ bool cell_is_inside_fast(const DevProblem& P, int ci, const V3& r, const V3& u, int on_surf) {
  const int q = cell_fast_nl(P.cellfast + ci, r, on_surf);
  if (q >= 0) return q != 0;
  return cell_is_inside_nl(geo_tables(P), ci, r, u, on_surf);
}

// nearest surface of the cell along u (cell.cpp:79-142); bc_only = distance_to_boundary_condition
template <class PT>
__device__ inline void cell_distance_impl(const PT& P, int ci, const V3& r, const V3& u, int on_surf, bool bc_only,
                                          double& min_dist, int& i_surf) {
  min_dist = ABL_INF;
  i_surf = 0;
  const abl_cell* c = P.cells + ci;
  if (bc_only && !ldt(&c->vac_or_refl)) return;
  const int off = ldt(&c->rpn_offset), len = ldt(&c->rpn_len);
  for (int k = 0; k < len; k++) {
    const int token = ldt(&P.rpn[off + k]);
    if (token >= ABL_OP_UNION) continue;
    const bool coincident = iabs(token) == iabs(on_surf);
    const Surf s = load_surface(P, iabs(token) - 1);
    if (bc_only && s.bc == ABL_BC_NORMAL) continue;
    const double d = surf_distance(s, r, u, coincident);
    if (d < min_dist) {
      // (against the initial min_dist = DBL_MAX the quotient is 1 - d/DBL_MAX >= 1e-14 for every d < DBL_MAX: no division)
      if (min_dist == ABL_INF || fabs(d - min_dist) / min_dist >= 1e-14) {
        min_dist = d;
        i_surf = -token;
      }
    }
  }
}

// one shared copy for the callers that work on the bare geometry tables (the per-lane kernel and the service warp reach
// this from three places; the surface-distance switch inside is several hundred instructions)
static __device__ __noinline__ void cell_distance_nl(const GeoTables G, int ci, const V3 r, const V3 u, int on_surf, bool bc_only,
                                              double* min_dist, int* i_surf) {
  double d;
  int is;
  cell_distance_impl(G, ci, r, u, on_surf, bc_only, d, is);
  *min_dist = d;
  *i_surf = is;
}
template <class PT>
__device__ __forceinline__ void cell_distance(const PT& P, int ci, const V3& r, const V3& u, int on_surf, bool bc_only,
                                              double& min_dist, int& i_surf) {
  cell_distance_impl(P, ci, r, u, on_surf, bc_only, min_dist, i_surf);
}
template <>
__device__ __forceinline__ void cell_distance<GeoTables>(const GeoTables& P, int ci, const V3& r, const V3& u, int on_surf, bool bc_only,
                                                         double& min_dist, int& i_surf) {
  cell_distance_nl(P, ci, r, u, on_surf, bc_only, &min_dist, &i_surf);
}

// ---- rectilinear lattice -------------------------------------------------------------------------
struct Lat {
  int Nx, Ny, Nz, tile_offset, outer;
  double Px, Py, Pz, Pxi, Pyi, Pzi, Xl, Yl, Zl;
};
__device__ __forceinline__ Lat load_lattice(const abl_universe* U) {
  Lat L;
  L.Nx = ldt(&U->N[0]); L.Ny = ldt(&U->N[1]); L.Nz = ldt(&U->N[2]);
  L.tile_offset = ldt(&U->tile_offset);
  L.outer = ldt(&U->outer);
  L.Px = ldt(&U->P[0]); L.Py = ldt(&U->P[1]); L.Pz = ldt(&U->P[2]);
  L.Pxi = ldt(&U->Pinv[0]); L.Pyi = ldt(&U->Pinv[1]); L.Pzi = ldt(&U->Pinv[2]);
  L.Xl = ldt(&U->Xl[0]); L.Yl = ldt(&U->Xl[1]); L.Zl = ldt(&U->Xl[2]);
  return L;
}
__device__ __forceinline__ V3 tile_center(const Lat& L, int nx, int ny, int nz) {  // rect_lattice.cpp:303-309
  return {((double)nx + 0.5) * L.Px + L.Xl, ((double)ny + 0.5) * L.Py + L.Yl, ((double)nz + 0.5) * L.Pz + L.Zl};
}
__device__ __forceinline__ void get_tile(const Lat& L, const V3& r, const V3& u, int& nx, int& ny, int& nz) {
  nx = (int)floor((r.x - L.Xl) * L.Pxi);  // rect_lattice.cpp:209-237
  ny = (int)floor((r.y - L.Yl) * L.Pyi);
  nz = (int)floor((r.z - L.Zl) * L.Pzi);
  const V3 rt = tile_center(L, nx, ny, nz);
  const double xl = rt.x - L.Px * 0.5;
  if (fabs(xl - r.x) < ABL_SURFACE_COINCIDENT && u.x < 0.) nx--;
  const double xh = rt.x + L.Px * 0.5;
  if (fabs(xh - r.x) < ABL_SURFACE_COINCIDENT && u.x >= 0.) nx++;
  const double yl = rt.y - L.Py * 0.5;
  if (fabs(yl - r.y) < ABL_SURFACE_COINCIDENT && u.y < 0.) ny--;
  const double yh = rt.y + L.Py * 0.5;
  if (fabs(yh - r.y) < ABL_SURFACE_COINCIDENT && u.y >= 0.) ny++;
  const double zl = rt.z - L.Pz * 0.5;
  if (fabs(zl - r.z) < ABL_SURFACE_COINCIDENT && u.z < 0.) nz--;
  const double zh = rt.z + L.Pz * 0.5;
  if (fabs(zh - r.z) < ABL_SURFACE_COINCIDENT && u.z >= 0.) nz++;
}
struct Tile3 {
  int nx, ny, nz;
};

// ---- hexagonal lattice (src/hex_lattice.cpp; table layout: abl_universe in include/abeille_b200.h) --------------------------
// The reference's quirks are kept: get_cell shifts the position by the lattice origin before it looks for the tile while the
// Tracker's pad validation and distance_to_tile_boundary use the pad's un-shifted position, and the z bin subtracts Z_o twice
// on the get_cell path (hex_lattice.cpp:89-96,283-288).  With the origin at zero none of that matters.
struct HexLat {
  int nrings, width, nz, top, tile_offset, outer;
  double pitch, pitch_z, X_o, Y_o, Z_o, cos_pi_6, sin_pi_6, cos_pi_3, sin_pi_3;
};
__device__ __forceinline__ HexLat load_hex(const abl_universe* U) {
  HexLat H;
  const int packed = ldt(&U->pad_);
  H.nrings = packed & 0xffff;
  H.top = packed >> 16;
  H.width = ldt(&U->N[0]);
  H.nz = ldt(&U->N[2]);
  H.tile_offset = ldt(&U->tile_offset);
  H.outer = ldt(&U->outer);
  H.pitch = ldt(&U->P[0]); H.sin_pi_3 = ldt(&U->P[1]); H.pitch_z = ldt(&U->P[2]);
  H.cos_pi_6 = ldt(&U->Pinv[0]); H.sin_pi_6 = ldt(&U->Pinv[1]); H.cos_pi_3 = ldt(&U->Pinv[2]);
  H.X_o = ldt(&U->Xl[0]); H.Y_o = ldt(&U->Xl[1]); H.Z_o = ldt(&U->Xl[2]);
  return H;
}
__device__ inline Tile3 hex_get_tile(const HexLat& H, const V3& p) {  // get_nearest_hex + get_tile, hex_lattice.cpp:238-288
  double q, r, det;
  if (H.top == 1) {
    det = -H.pitch * H.pitch * H.sin_pi_3;
    q = (H.pitch / det) * (0. * p.x - 1. * p.y);
    r = (H.pitch / det) * (-H.sin_pi_3 * p.x + H.cos_pi_3 * p.y);
  } else {
    det = -H.pitch * H.pitch * H.cos_pi_6;
    q = (H.pitch / det) * (H.sin_pi_6 * p.x - H.cos_pi_6 * p.y);
    r = (H.pitch / det) * (-1. * p.x - 0. * p.y);
  }
  const double x = q, z = r, y = -x - z;
  double rx = round(x), ry = round(y), rz = round(z);
  const double x_diff = fabs(rx - x), y_diff = fabs(ry - y), z_diff = fabs(rz - z);
  if (x_diff > y_diff && x_diff > z_diff) {
    rx = -ry - rz;
  } else if (y_diff > x_diff && y_diff > z_diff) {
    ry = -rx - rz;
  } else {
    rz = -rx - ry;
  }
  const double Z_low = H.Z_o - 0.5 * (double)H.nz * H.pitch_z;
  Tile3 t;
  t.nx = (int)rx;
  t.ny = (int)rz;
  t.nz = (int)floor((p.z - Z_low) / H.pitch_z);
  return t;
}
__device__ inline V3 hex_tile_center(const HexLat& H, int q_, int r_, int nz) {  // hex_lattice.cpp:290-331
  const double q = (double)q_, r = (double)r_;
  double x, y;
  if (H.top == 1) {
    x = H.pitch * (H.cos_pi_3 * q + 1. * r);
    y = H.pitch * (H.sin_pi_3 * q + 0. * r);
  } else {
    x = H.pitch * (0. * q + H.cos_pi_6 * r);
    y = H.pitch * (1. * q + H.sin_pi_6 * r);
  }
  const double Z_low = H.Z_o - 0.5 * (double)H.nz * H.pitch_z;
  return V3{x, y, ((double)nz + 0.5) * H.pitch_z + Z_low};
}
__device__ __forceinline__ int hex_ring(int x, int z) {  // hex_lattice.cpp:333-345
  const int y = -x - z;
  return max(max(iabs(x), iabs(y)), iabs(z));
}
__device__ __forceinline__ double hex_distance_to_line(const V3& r, const V3& u, double x1, double y1, double x2, double y2) {
  const double A = y2 - y1, B = x1 - x2;  // hex_lattice.cpp:437-455
  const double D = (x2 - x1) * y1 - (y2 - y1) * x1;
  const double num = D - A * r.x - B * r.y;
  const double denom = A * u.x + B * u.y;
  const double d = num / denom;
  if (d < 0.) return __longlong_as_double(0x7ff0000000000000LL);  // INFINITY
  return d;
}
static __device__ __noinline__ double hex_distance_to_tile_boundary(const abl_universe* __restrict__ U, const V3 r_local, const V3 u,
                                                                    const Tile3 tile) {  // hex_lattice.cpp:347-435
  const HexLat H = load_hex(U);
  const V3 center = hex_tile_center(H, tile.nx, tile.ny, tile.nz);
  const V3 r_tile{r_local.x - center.x, r_local.y - center.y, r_local.z - center.z};
  double d1, d2, d3, d4, d5, d6;
  if (H.top == 0) {
    const double x1 = 0., y1 = H.pitch / (2. * H.cos_pi_6), x2 = H.pitch / 2., y2 = H.pitch / (2. * H.sin_pi_6);
    d1 = hex_distance_to_line(r_tile, u, x1, y1, x2, y1);
    d2 = hex_distance_to_line(r_tile, u, x2, y2, x2, -y2);
    d3 = hex_distance_to_line(r_tile, u, x2, -y2, x1, -y1);
    d4 = hex_distance_to_line(r_tile, u, x1, -y1, -x2, -y2);
    d5 = hex_distance_to_line(r_tile, u, -x2, -y2, -x2, y2);
    d6 = hex_distance_to_line(r_tile, u, -x2, y2, x1, y1);
  } else {
    const double x1 = H.pitch / (2. * H.sin_pi_6), y1 = H.pitch / 2., x2 = H.pitch / (2. * H.cos_pi_6), y2 = 0.;
    d1 = hex_distance_to_line(r_tile, u, x1, y1, x2, y1);
    d2 = hex_distance_to_line(r_tile, u, x2, y2, x1, -y1);
    d3 = hex_distance_to_line(r_tile, u, x1, -y1, -x1, -y1);
    d4 = hex_distance_to_line(r_tile, u, -x1, -y1, -x2, -y2);
    d5 = hex_distance_to_line(r_tile, u, -x2, -y2, -x1, y1);
    d6 = hex_distance_to_line(r_tile, u, -x1, y1, x1, y1);
  }
  const double dzl = (-H.pitch_z * 0.5 - r_tile.z) / u.z;
  const double dzu = (H.pitch_z * 0.5 - r_tile.z) / u.z;
  double d = ABL_INF;
  if (d1 > 0. && d1 < d) d = d1;
  if (d2 > 0. && d2 < d) d = d2;
  if (d3 > 0. && d3 < d) d = d3;
  if (d4 > 0. && d4 < d) d = d4;
  if (d5 > 0. && d5 < d) d = d5;
  if (d6 > 0. && d6 < d) d = d6;
  if (dzl > 0. && dzl < d) d = dzl;
  if (dzu > 0. && dzu < d) d = dzu;
  return d;
}
// One step of HexLattice::get_cell (hex_lattice.cpp:140-202): the tile of r, the universe inside it (-1: none: the outer
// universe or nothing) and the position in the tile's frame
struct HexStep {
  Tile3 t;
  int sub, outer;
  V3 r_tile;
};
static __device__ __noinline__ HexStep hex_lattice_step(const abl_universe* __restrict__ U, const int32_t* __restrict__ tiles, const V3 r) {
  const HexLat H = load_hex(U);
  const V3 r_o{r.x - H.X_o, r.y - H.Y_o, r.z - H.Z_o};
  HexStep s;
  s.t = hex_get_tile(H, r_o);
  s.outer = H.outer;
  s.sub = -1;
  s.r_tile = r;
  if (hex_ring(s.t.nx, s.t.ny) < H.nrings && !(s.t.nz < 0 || s.t.nz >= H.nz)) {
    const int mid = H.width / 2;
    s.sub = ldt(&tiles[H.tile_offset + s.t.nz * (H.width * H.width) + (s.t.ny + mid) * H.width + (s.t.nx + mid)]);
    if (s.sub >= 0) {
      const V3 ctr = hex_tile_center(H, s.t.nx, s.t.ny, s.t.nz);
      s.r_tile = V3{r_o.x - ctr.x, r_o.y - ctr.y, r_o.z - ctr.z};
    }
  }
  return s;
}
static __device__ __noinline__ Tile3 hex_tile_nl(const abl_universe* __restrict__ U, const V3 p) { return hex_get_tile(load_hex(U), p); }

// Lattice::get_tile as one shared copy per kernel (used by the pad validation and by the descent)
// HEX = false: a build for problems without a hexagonal lattice (the host checks: DevProblem::has_hex); the hexagonal branch
// costs the rectilinear hot path registers (107 -> 113 ms per 1e7 histories on the bench workload when it is compiled in)
template <bool HEX = true>
static __device__ ABL_HOT_CALL Tile3 lattice_tile_nl(const abl_universe* __restrict__ U, const V3 r, const V3 u) {
  if (HEX && ldt(&U->type) == ABL_UNI_HEX) return hex_tile_nl(U, r);  // (the position as handed in: see the note on the reference's quirks)
  const Lat L = load_lattice(U);
  Tile3 t;
  get_tile(L, r, u, t.nx, t.ny, t.nz);
  return t;
}
__device__ __forceinline__ bool tile_in_range(const Lat& L, int nx, int ny, int nz) {
  return !((nx < 0 || nx >= L.Nx) || (ny < 0 || ny >= L.Ny) || (nz < 0 || nz >= L.Nz));
}
// u_inv = (1/u.x, 1/u.y, 1/u.z): the reference evaluates the three reciprocals in every call; they depend on the
// direction only, so the caller computes them once per boundary search
__device__ inline double distance_to_tile_boundary(const Lat& L, const V3& r_local, const V3& u_inv, int nx, int ny, int nz) {
  const V3 center = tile_center(L, nx, ny, nz);  // rect_lattice.cpp:239-282
  const double tx = r_local.x - center.x, ty = r_local.y - center.y, tz = r_local.z - center.z;
  double dist = ABL_INF;
  const double diff_xl = -L.Px * 0.5 - tx;
  const double diff_xh = L.Px * 0.5 - tx;
  const double diff_yl = -L.Py * 0.5 - ty;
  const double diff_yh = L.Py * 0.5 - ty;
  const double diff_zl = -L.Pz * 0.5 - tz;
  const double diff_zh = L.Pz * 0.5 - tz;
  const double ux_inv = u_inv.x, uy_inv = u_inv.y, uz_inv = u_inv.z;
  const double d_xl = diff_xl * ux_inv, d_xh = diff_xh * ux_inv;
  const double d_yl = diff_yl * uy_inv, d_yh = diff_yh * uy_inv;
  const double d_zl = diff_zl * uz_inv, d_zh = diff_zh * uz_inv;
  const double guard = 100 * ABL_SURFACE_COINCIDENT;
  if (d_xl > 0. && d_xl < dist && fabs(diff_xl) > guard) dist = d_xl;
  if (d_xh > 0. && d_xh < dist && fabs(diff_xh) > guard) dist = d_xh;
  if (d_yl > 0. && d_yl < dist && fabs(diff_yl) > guard) dist = d_yl;
  if (d_yh > 0. && d_yh < dist && fabs(diff_yh) > guard) dist = d_yh;
  if (d_zl > 0. && d_zl < dist && fabs(diff_zl) > guard) dist = d_zl;
  if (d_zh > 0. && d_zh < dist && fabs(diff_zh) > guard) dist = d_zh;
  return dist;
}

// ---- boundaries ------------------------------------------------------------------------------------
struct Boundary {  // include/geometry/boundary.hpp:33-42
  double distance;
  int surface_index;
  int btype;
  int token;
};

// candidate (d, i_surf) from a cell against the running nearest boundary (tracker.hpp:104-131,
// cell_universe.cpp:121-147): takes it when closer by more than BOUNDRY_TOL and a surface was found
template <class PT>
__device__ __forceinline__ void take_cell_candidate(const PT& P, double d, int i_surf, const V3& r, const V3& u,
                                                    Boundary& b) {
  if (d < b.distance && fabs(d - b.distance) > ABL_BOUNDRY_TOL) {
    const int tmp_token = iabs(i_surf);
    if (tmp_token) {
      b.token = tmp_token;
      b.distance = d;
      b.surface_index = tmp_token - 1;
      const Surf s = load_surface(P, b.surface_index);
      b.btype = s.bc;
      if (surf_sign(s, r, u) < 0) b.token *= -1;
    }
  }
}

// Universe::get_boundary_condition (cell_universe.cpp:111-154, lattice.cpp:77-92)
template <class PT>
__device__ inline Boundary universe_boundary_condition(const PT& P, int uni, const V3& r, const V3& u, int on_surf) {
  Boundary b{ABL_INF, -1, ABL_BC_VACUUM, 0};
  const abl_universe* U = P.universes + uni;
  while (ldt(&U->type) != ABL_UNI_CELLS) {  // lattices defer to their outer universe
    if (!ldt(&U->has_bc)) return b;
    U = P.universes + ldt(&U->outer);
  }
  if (ldt(&U->has_bc)) {
    const int off = ldt(&U->cell_offset), n = ldt(&U->ncells);
    for (int k = 0; k < n; k++) {
      const int ci = ldt(&P.ucells[off + k]);
      if (!ldt(&P.cells[ci].vac_or_refl)) continue;
      double d;
      int is;
      cell_distance(P, ci, r, u, on_surf, true, d, is);
      take_cell_candidate(P, d, is, r, u, b);
    }
  }
  return b;
}

// ---- the geometry cursor -----------------------------------------------------------------------------
enum { PAD_UNIVERSE = 0, PAD_LATTICE = 1, PAD_CELL = 2 };

struct Cursor {
  int token;     // surface the particle sits on: +-(surface index+1), 0 = none (tracker.hpp:367-370)
  int cell, mat; // current cell / material index, -1 = lost
  int np, nf;    // pads, frames in use
  int err;       // ABL_ERR_* raised by the cursor (stack overflow, malformed nesting)
  int pinfo[ABL_MAX_PADS];     // type | outside_flag<<2 | frame<<3 | index<<8
  int ptile[ABL_MAX_PADS][3];  // lattice pads: tile found at descent
  double fx[ABL_MAX_FRAMES], fy[ABL_MAX_FRAMES], fz[ABL_MAX_FRAMES];  // r_local of each frame; frame 0 = global r
};

__device__ __forceinline__ int pad_type(int info) { return info & 3; }
__device__ __forceinline__ int pad_flag(int info) { return (info >> 2) & 1; }
__device__ __forceinline__ int pad_frame(int info) { return (info >> 3) & 31; }
__device__ __forceinline__ int pad_index(int info) { return info >> 8; }
__device__ __forceinline__ int make_pad(int type, int flag, int frame, int index) {
  return type | (flag << 2) | (frame << 3) | (index << 8);
}
// Storage accessors.  The algorithms below (descent, move, validation) are written against these, so the same
// code drives the local-memory Cursor and the shared-memory cursor of the history kernel (history.cuh: SCursor).
__device__ __forceinline__ V3 frame_r(const Cursor& c, int f) { return {c.fx[f], c.fy[f], c.fz[f]}; }
__device__ __forceinline__ void set_frame(Cursor& c, int f, double x, double y, double z) {
  c.fx[f] = x;
  c.fy[f] = y;
  c.fz[f] = z;
}
__device__ __forceinline__ void shift_frame(Cursor& c, int f, double dx, double dy, double dz) {
  c.fx[f] = c.fx[f] + dx;
  c.fy[f] = c.fy[f] + dy;
  c.fz[f] = c.fz[f] + dz;
}
__device__ __forceinline__ int pad_info(const Cursor& c, int i) { return c.pinfo[i]; }
__device__ __forceinline__ void store_pad(Cursor& c, int i, int info, int tx, int ty, int tz) {
  c.pinfo[i] = info;
  c.ptile[i][0] = tx;
  c.ptile[i][1] = ty;
  c.ptile[i][2] = tz;
}
__device__ __forceinline__ Tile3 pad_tile3(const Cursor& c, int i) { return Tile3{c.ptile[i][0], c.ptile[i][1], c.ptile[i][2]}; }
__device__ __forceinline__ bool pad_tile_is(const Cursor& c, int i, int nx, int ny, int nz) {
  return c.ptile[i][0] == nx && c.ptile[i][1] == ny && c.ptile[i][2] == nz;
}

template <class CUR>
__device__ __forceinline__ bool push_pad(CUR& c, int info, int tx = 0, int ty = 0, int tz = 0) {
  if (c.np >= ABL_MAX_PADS) {
    c.err = ABL_ERR_GEOMETRY;
    return false;
  }
  store_pad(c, c.np, info, tx, ty, tz);
  c.np++;
  return true;
}

// Universe::get_cell(stack, r, u, on_surf) made iterative (cell_universe.cpp:72-109, rect_lattice.cpp:132-207).
// Descends from universe `uni` whose coordinates are frame f; returns the material cell or -1 (lost).
template <bool FAST = false, class CUR = Cursor, class PT = DevProblem>
__device__ inline int descend(const PT& P, CUR& c, int uni, int f, const V3& u) {
  for (;;) {
    const abl_universe* U = P.universes + uni;
    const V3 r = frame_r(c, f);
    if (ldt(&U->type) == ABL_UNI_CELLS) {
      if (!push_pad(c, make_pad(PAD_UNIVERSE, 0, f, uni))) return -1;
      const int off = ldt(&U->cell_offset), n = ldt(&U->ncells);
      int found = -1;
      for (int k = 0; k < n; k++) {
        const int ci = ldt(&P.ucells[off + k]);
        bool in;
        if constexpr (FAST) in = cell_is_inside_fast(P, ci, r, u, c.token);
        else in = cell_is_inside(P, ci, r, u, c.token);
        if (in) {
          found = ci;
          break;
        }
      }
      c.nf = f + 1;
      if (found < 0) return -1;
      if (!push_pad(c, make_pad(PAD_CELL, 0, f, found))) return -1;
      const int fill = ldt(&P.cells[found].fill_universe);
      if (fill < 0) return found;
      uni = fill;
      continue;
    }
    if (ldt(&U->type) == ABL_UNI_HEX) {
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
        continue;
      }
      if (hs.outer >= 0) {
        if (!push_pad(c, make_pad(PAD_LATTICE, 1, f, uni), hs.t.nx, hs.t.ny, hs.t.nz)) return -1;
        uni = hs.outer;
        continue;
      }
      push_pad(c, make_pad(PAD_LATTICE, 0, f, uni), hs.t.nx, hs.t.ny, hs.t.nz);
      return -1;
    }
    int nx, ny, nz;
    Lat L;
    if (FAST) {
      const Tile3 t3 = lattice_tile_nl(U, r, u);
      nx = t3.nx; ny = t3.ny; nz = t3.nz;
      L.Nx = ldt(&U->N[0]); L.Ny = ldt(&U->N[1]); L.Nz = ldt(&U->N[2]);
      L.tile_offset = ldt(&U->tile_offset);
      L.outer = ldt(&U->outer);
      L.Px = ldt(&U->P[0]); L.Py = ldt(&U->P[1]); L.Pz = ldt(&U->P[2]);
      L.Xl = ldt(&U->Xl[0]); L.Yl = ldt(&U->Xl[1]); L.Zl = ldt(&U->Xl[2]);
    } else {
      L = load_lattice(U);
      get_tile(L, r, u, nx, ny, nz);
    }
    int sub = -1;
    if (tile_in_range(L, nx, ny, nz)) sub = ldt(&P.tiles[L.tile_offset + nz * (L.Nx * L.Ny) + nx * L.Ny + ny]);
    c.nf = f + 1;
    if (sub >= 0) {
      if (!push_pad(c, make_pad(PAD_LATTICE, 0, f, uni), nx, ny, nz)) return -1;
      if (f + 1 >= ABL_MAX_FRAMES) {
        c.err = ABL_ERR_GEOMETRY;
        return -1;
      }
      const V3 ctr = tile_center(L, nx, ny, nz);
      set_frame(c, f + 1, r.x - ctr.x, r.y - ctr.y, r.z - ctr.z);
      f++;
      uni = sub;
      continue;
    }
    if (L.outer >= 0) {  // outside the lattice or an empty tile: the outer universe, un-shifted r
      if (!push_pad(c, make_pad(PAD_LATTICE, 1, f, uni), nx, ny, nz)) return -1;
      uni = L.outer;
      continue;
    }
    push_pad(c, make_pad(PAD_LATTICE, 0, f, uni), nx, ny, nz);
    return -1;
  }
}

// Tracker::restart_get_current (tracker.hpp:63-74): full lookup from the root at global position r
template <class PT>
__device__ inline void cursor_restart(const PT& P, Cursor& c, const V3& r, const V3& u) {
  c.np = 0;
  c.fx[0] = r.x;
  c.fy[0] = r.y;
  c.fz[0] = r.z;
  c.nf = 1;
  c.cell = descend<false, Cursor, PT>(P, c, P.root, 0, u);
  c.mat = c.cell >= 0 ? ldt(&P.cells[c.cell].material) : -1;
}

// Tracker::move (tracker.hpp:76-85): every frame advances by d*u, the surface token is dropped
template <class CUR>
__device__ __forceinline__ void cursor_move(CUR& c, double d, const V3& u) {
  const double dx = d * u.x, dy = d * u.y, dz = d * u.z;
#pragma unroll 1
  for (int f = 0; f < c.nf; f++) shift_frame(c, f, dx, dy, dz);
  c.token = 0;
}

// Tracker::get_current (tracker.hpp:235-306): re-validate the pads top-down, re-descend from the
// first one that no longer holds.  (check_tree() is always true here: the cursor's global position
// IS frame 0; the callers that reposition the particle call cursor_restart directly.)
template <class PT>
__device__ inline void cursor_get_current(const PT& P, Cursor& c, const V3& u) {
  int first_bad = c.np;
  for (int it = 0; it < c.np; it++) {
    const int info = c.pinfo[it];
    const int type = pad_type(info);
    if (type == PAD_CELL) {
      if (!cell_is_inside(P, pad_index(info), frame_r(c, pad_frame(info)), u, c.token)) {
        first_bad = it;
        break;
      }
    } else if (type == PAD_LATTICE) {
      const Tile3 tt = lattice_tile_nl(P.universes + pad_index(info), frame_r(c, pad_frame(info)), u);
      const int nx = tt.nx, ny = tt.ny, nz = tt.nz;
      if (c.ptile[it][0] != nx || c.ptile[it][1] != ny || c.ptile[it][2] != nz) {
        first_bad = it;
        break;
      }
    }
  }
  if (first_bad == c.np) return;
  // re-descend from the pad above the first bad one; a full lookup from the root when there is none, when that pad is a
  // cell (a lattice sitting directly inside a universe-filled cell: the reference looks the CELL id up in its
  // universe-id map here (tracker.hpp:287), which lands in an unrelated universe; no shipped deck has this nesting, and
  // the geometrically correct answer is a fresh lookup), or when the partial descent finds nothing
  int uni = P.root, f = 0;
  bool full = true;
  if (first_bad > 0) {
    const int back = c.pinfo[first_bad - 1];
    if (pad_type(back) != PAD_CELL) {
      c.np = first_bad - 1;
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
    c.cell = descend<false, Cursor, PT>(P, c, uni, f, u);
    if (c.cell >= 0 || full) break;
    full = true;
    c.np = 0;
    c.nf = 1;
    uni = P.root;
    f = 0;
  }
  c.mat = c.cell >= 0 ? ldt(&P.cells[c.cell].material) : -1;
}

// Tracker::get_boundary_condition (tracker.hpp:94-161)
template <class PT, class CUR>
__device__ inline Boundary cursor_boundary_condition(const PT& P, const CUR& c, const V3& u) {
  if (c.cell < 0) return universe_boundary_condition(P, P.root, frame_r(c, 0), u, c.token);
  Boundary b{ABL_INF, -1, ABL_BC_VACUUM, 0};
  for (int it = 0; it < c.np; it++) {
    const int info = pad_info(c, it);
    const V3 r = frame_r(c, pad_frame(info));
    if (pad_type(info) == PAD_CELL) {
      const int ci = pad_index(info);
      if (!ldt(&P.cells[ci].vac_or_refl)) continue;
      double d;
      int is;
      cell_distance(P, ci, r, u, c.token, true, d, is);
      take_cell_candidate(P, d, is, r, u, b);
    } else {
      const int ui = pad_index(info);
      if (ldt(&P.universes[ui].has_bc)) {
        const Boundary ub = universe_boundary_condition(P, ui, r, u, c.token);
        if (ub.distance < b.distance && fabs(ub.distance - b.distance) > ABL_BOUNDRY_TOL) b = ub;
      }
    }
  }
  return b;
}

// Tracker::get_nearest_boundary (tracker.hpp:163-225); the cursor is never lost when this is called
template <class PT, class CUR>
__device__ inline Boundary cursor_nearest_boundary(const PT& P, const CUR& c, const V3& u) {
  Boundary b = cursor_boundary_condition(P, c, u);
  V3 u_inv{0., 0., 0.};
  bool have_u_inv = false;
  for (int it = 0; it < c.np; it++) {
    const int info = pad_info(c, it);
    const int type = pad_type(info);
    const V3 r = frame_r(c, pad_frame(info));
    if (type == PAD_LATTICE) {
      const abl_universe* LU = P.universes + pad_index(info);
      const Tile3 t3 = pad_tile3(c, it);
      double d;
      if (ldt(&LU->type) == ABL_UNI_HEX) {
        d = hex_distance_to_tile_boundary(LU, r, u, t3);
      } else {
        const Lat L = load_lattice(LU);
        if (!have_u_inv) {
          u_inv = V3{1. / u.x, 1. / u.y, 1. / u.z};
          have_u_inv = true;
        }
        d = distance_to_tile_boundary(L, r, u_inv, t3.nx, t3.ny, t3.nz);
      }
      if (d < b.distance && fabs(d - b.distance) > ABL_BOUNDRY_TOL) {
        b.distance = d;
        b.btype = ABL_BC_NORMAL;
        b.surface_index = -1;
        b.token = 0;
      }
    } else if (type == PAD_CELL) {
      double d;
      int is;
      cell_distance(P, pad_index(info), r, u, c.token, false, d, is);
      if (d < b.distance && fabs(d - b.distance) > ABL_BOUNDRY_TOL) {
        b.distance = d;
        b.token = iabs(is);
        b.surface_index = b.token ? b.token - 1 : -1;
        if (b.surface_index >= 0) {
          const Surf s = load_surface(P, b.surface_index);
          b.btype = s.bc;
          if (surf_sign(s, r, u) < 0) b.token *= -1;
        } else {
          b.btype = ABL_BC_NORMAL;
        }
      }
    }
  }
  return b;
}

// Tracker::get_nearest_boundary when the boundary conditions are known to lie further than bc_floor (a certified lower
// bound of the distance to every vacuum / reflective surface, minus BOUNDRY_TOL: see DevProblem::bc_*).  The reference
// evaluates the boundary-condition search first and then compares every pad's candidate against it; while the running
// nearest boundary is still "some boundary condition beyond bc_floor", a candidate below bc_floor wins against it whatever
// its exact distance is, and from then on the comparisons are the reference's.  When a candidate is not below bc_floor before
// one was accepted (or there is none) the boundary-condition search is evaluated after all and the comparisons are replayed
// against its real result; the candidates' distances are computed once either way.  The result is the reference's, bit for
// bit; the six plane distances (six fp64 divisions) of the boundary-condition search -- up to three times per flight in the
// reflector -- are evaluated only for flights that end near a boundary condition.
template <class PT, class CUR>
__device__ inline Boundary cursor_nearest_boundary_lazy(const PT& P, const CUR& c, const V3& u, double bc_floor) {
  // every pad's candidate distance first (the expensive part, the same evaluations as the reference's)
  double cd[ABL_MAX_PADS];
  int cis[ABL_MAX_PADS];  // cell pads: the surface token the cell reported; lattice pads: 0
  V3 u_inv{0., 0., 0.};
  bool have_u_inv = false;
  for (int it = 0; it < c.np; it++) {
    const int info = pad_info(c, it);
    const int type = pad_type(info);
    cd[it] = ABL_INF;
    cis[it] = 0;
    if (type == PAD_LATTICE) {
      const abl_universe* LU = P.universes + pad_index(info);
      const Tile3 t3 = pad_tile3(c, it);
      if (ldt(&LU->type) == ABL_UNI_HEX) {
        cd[it] = hex_distance_to_tile_boundary(LU, frame_r(c, pad_frame(info)), u, t3);
      } else {
        const Lat L = load_lattice(LU);
        if (!have_u_inv) {
          u_inv = V3{1. / u.x, 1. / u.y, 1. / u.z};
          have_u_inv = true;
        }
        cd[it] = distance_to_tile_boundary(L, frame_r(c, pad_frame(info)), u_inv, t3.nx, t3.ny, t3.nz);
      }
    } else if (type == PAD_CELL) {
      cell_distance(P, pad_index(info), frame_r(c, pad_frame(info)), u, c.token, false, cd[it], cis[it]);
    }
  }
  // the reference's chain of comparisons with the boundary condition still symbolic ("beyond bc_floor")
  int win = -1;
  double best = ABL_INF;
  bool need_bc = false;
  for (int it = 0; it < c.np && !need_bc; it++) {
    if (pad_type(pad_info(c, it)) == PAD_UNIVERSE) continue;
    const double d = cd[it];
    if (win < 0) {
      if (d < bc_floor) {
        win = it;
        best = d;
      } else {
        need_bc = true;
      }
    } else if (d < best && fabs(d - best) > ABL_BOUNDRY_TOL) {
      win = it;
      best = d;
    }
  }
  Boundary b{ABL_INF, -1, ABL_BC_VACUUM, 0};
  if (need_bc || win < 0) {  // the boundary conditions can matter: evaluate them and replay the chain against the real value
    b = cursor_boundary_condition(P, c, u);
    win = -1;
    best = b.distance;
    for (int it = 0; it < c.np; it++) {
      if (pad_type(pad_info(c, it)) == PAD_UNIVERSE) continue;
      const double d = cd[it];
      if (d < best && fabs(d - best) > ABL_BOUNDRY_TOL) {
        win = it;
        best = d;
      }
    }
    if (win < 0) return b;
  }
  // the winner's description (what the reference fills in whenever a candidate is taken; only the last one survives)
  const int info = pad_info(c, win);
  b.distance = best;
  if (pad_type(info) == PAD_LATTICE) {
    b.btype = ABL_BC_NORMAL;
    b.surface_index = -1;
    b.token = 0;
  } else {
    b.token = iabs(cis[win]);
    b.surface_index = b.token ? b.token - 1 : -1;
    if (b.surface_index >= 0) {
      const Surf s = load_surface(P, b.surface_index);
      b.btype = s.bc;
      if (surf_sign(s, frame_r(c, pad_frame(info)), u) < 0) b.token *= -1;
    } else {
      b.btype = ABL_BC_NORMAL;
    }
  }
  return b;
}

// The cursor operations as real function calls on the geometry tables alone: the per-lane kernel (transport.cuh) calls
// each of them from several places, and inlined copies made it 230 KB of SASS -- ncu showed it waiting for instructions
// (21 stall cycles per issue "no instruction", instruction-cache hit rate 53 %).
static __device__ __noinline__ void cursor_restart_nl(const GeoTables G, Cursor& c, const V3 r, const V3 u) { cursor_restart(G, c, r, u); }
static __device__ __noinline__ void cursor_get_current_nl(const GeoTables G, Cursor& c, const V3 u) { cursor_get_current(G, c, u); }
static __device__ __noinline__ Boundary cursor_nearest_boundary_nl(const GeoTables G, const Cursor& c, const V3 u) {
  return cursor_nearest_boundary(G, c, u);
}
static __device__ __noinline__ Boundary cursor_boundary_condition_nl(const GeoTables G, const Cursor& c, const V3 u) {
  return cursor_boundary_condition(G, c, u);
}

}  // namespace abl
