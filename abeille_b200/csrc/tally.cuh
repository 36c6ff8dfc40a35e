/* tally.cuh -- mesh-tally scoring on the device.
 *
 * Follows the reference scorers:
 *   CollisionMeshTally::score_collision      src/collision_mesh_tally.cpp:33-110
 *   TrackLengthMeshTally::score_flight       src/track_length_mesh_tally.cpp:77-435
 *   SourceMeshTally::score_source            src/source_mesh_tally.cpp:30-78
 * tally_gen is one fp64 array [Ne,Nx,Ny,Nz] (C order) per tally in HBM; every score is one
 * RED.ADD.F64 (fire-and-forget atomic, resolved in L2).  For the meshes of the reference decks
 * (0.84 GB for c5g7) the array is far larger than L2, so the cost of a score is one 32 B sector
 * read-modify-write in DRAM; see DESIGN.md for the measured rates.
 */
#pragma once
#include "geom.cuh"

namespace abl {

struct MatXS {
  double Et, Ea, Ef, Es;
};

__device__ __forceinline__ int tally_energy_bin(const DevTally& t, double E) {  // "<= E <=", first match
  for (int e = 0; e < t.Ne; e++)
    if (ldt(&t.ebounds[e]) <= E && E <= ldt(&t.ebounds[e + 1])) return e;
  return -1;
}

__device__ __forceinline__ size_t tally_index(const DevTally& t, int e, int i, int j, int k) {
  return (((size_t)e * (size_t)t.Nx + (size_t)i) * (size_t)t.Ny + (size_t)j) * (size_t)t.Nz + (size_t)k;
}

__device__ __forceinline__ void red_add(double* p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ double quantity_factor(int quantity, double w, double w2, const MatXS& m) {
  switch (quantity) {
    case ABL_Q_FLUX: return w;
    case ABL_Q_ELASTIC: return w * m.Es;
    case ABL_Q_ABSORPTION: return w * m.Ea;
    case ABL_Q_FISSION: return w * m.Ef;
    case ABL_Q_TOTAL: return w * m.Et;
    case ABL_Q_MT: return w * 0.;
    case ABL_Q_REAL_FLUX: return w;
    case ABL_Q_IMAG_FLUX: return w2;
    default: return 1.;  // source quantities are not collision / track-length quantities
  }
}

// returns 1 when a bin was scored
// l = energy bin of the particle (tally_energy_bin, or the per-group table for mid-point energies)
// scr0 = 1 / (Et * net_weight) (collision_mesh_tally.cpp:35-40), evaluated by the caller or read from DevProblem::inv_score
__device__ inline int score_collision_pre(const DevTally& t, const V3& r, int l, double w, double w2, const MatXS& m, double scr0);
template <class M = InlineMath>
__device__ inline int score_collision(const DevTally& t, const V3& r, int l, double w, double w2, const MatXS& m) {
  return score_collision_pre(t, r, l, w, w2, m, M::div(1., m.Et * t.net_weight));
}
__device__ inline int score_collision_pre(const DevTally& t, const V3& r, int l, double w, double w2, const MatXS& m, double scr0) {
  double scr = scr0;
  const int i = (int)floor((r.x - t.lowx) * t.dx_inv);
  const int j = (int)floor((r.y - t.lowy) * t.dy_inv);
  const int k = (int)floor((r.z - t.lowz) * t.dz_inv);
  if (l == -1) return 0;
  if (i >= 0 && i < t.Nx && j >= 0 && j < t.Ny && k >= 0 && k < t.Nz) {
    if (t.quantity <= ABL_Q_IMAG_FLUX) scr *= quantity_factor(t.quantity, w, w2, m);
    red_add(t.gen + tally_index(t, l, i, j, k), scr);
    return 1;
  }
  return 0;
}

struct MeshPos {
  int i, j, k;
  int on0, on1, on2;
};

// track_length_mesh_tally.cpp:256-326
__device__ __forceinline__ void mesh_initialize_indices(const DevTally& t, const V3& r, const V3& u, MeshPos& p) {
  p.i = (int)floor((r.x - t.lowx) * t.dx_inv);
  p.j = (int)floor((r.y - t.lowy) * t.dy_inv);
  p.k = (int)floor((r.z - t.lowz) * t.dz_inv);
  p.on0 = p.on1 = p.on2 = 0;
  const double xc = t.lowx + p.i * t.dx + 0.5 * t.dx;
  const double yc = t.lowy + p.j * t.dy + 0.5 * t.dy;
  const double zc = t.lowz + p.k * t.dz + 0.5 * t.dz;
  const double xl = xc - 0.5 * t.dx, xh = xc + 0.5 * t.dx;
  const double yl = yc - 0.5 * t.dy, yh = yc + 0.5 * t.dy;
  const double zl = zc - 0.5 * t.dz, zh = zc + 0.5 * t.dz;
  if (fabs(xl - r.x) < ABL_SURFACE_COINCIDENT) {
    if (u.x < 0.) { p.i--; p.on0 = 1; } else { p.on0 = -1; }
  } else if (fabs(xh - r.x) < ABL_SURFACE_COINCIDENT) {
    if (u.x < 0.) { p.on0 = 1; } else { p.i++; p.on0 = -1; }
  }
  if (fabs(yl - r.y) < ABL_SURFACE_COINCIDENT) {
    if (u.y < 0.) { p.j--; p.on1 = 1; } else { p.on1 = -1; }
  } else if (fabs(yh - r.y) < ABL_SURFACE_COINCIDENT) {
    if (u.y < 0.) { p.on1 = 1; } else { p.j++; p.on1 = -1; }
  }
  if (fabs(zl - r.z) < ABL_SURFACE_COINCIDENT) {
    if (u.z < 0.) { p.k--; p.on2 = 1; } else { p.on2 = -1; }
  } else if (fabs(zh - r.z) < ABL_SURFACE_COINCIDENT) {
    if (u.z < 0.) { p.on2 = 1; } else { p.k++; p.on2 = -1; }
  }
}

// track_length_mesh_tally.cpp:182-254 ; moves r to the mesh entry point, shortens d_flight
__device__ inline bool mesh_find_entry_point(const DevTally& t, V3& r, const V3& u, double& d_flight) {
  const double ux_inv = 1. / u.x, uy_inv = 1. / u.y, uz_inv = 1. / u.z;
  double d_min = (t.lowx - r.x) * ux_inv;
  double d_max = (t.hix - r.x) * ux_inv;
  double tmp;
  if (d_min > d_max) { tmp = d_min; d_min = d_max; d_max = tmp; }
  double d_y_min = (t.lowy - r.y) * uy_inv;
  double d_y_max = (t.hiy - r.y) * uy_inv;
  if (d_y_min > d_y_max) { tmp = d_y_min; d_y_min = d_y_max; d_y_max = tmp; }
  if ((d_min > d_y_max) || (d_y_min > d_max)) return false;
  if (d_y_min > d_min) d_min = d_y_min;
  if (d_y_max < d_max) d_max = d_y_max;
  double d_z_min = (t.lowz - r.z) * uz_inv;
  double d_z_max = (t.hiz - r.z) * uz_inv;
  if (d_z_min > d_z_max) { tmp = d_z_min; d_z_min = d_z_max; d_z_max = tmp; }
  if ((d_min > d_z_max) || (d_z_min > d_max)) return false;
  if (d_z_min > d_min) d_min = d_z_min;
  if (d_z_max < d_max) d_max = d_z_max;
  if (d_max < d_min) { tmp = d_max; d_max = d_min; d_min = tmp; }
  if ((d_max < 0.) && (d_min < 0.)) return false;
  if (d_min < 0.) return false;
  r.x = r.x + d_min * u.x;
  r.y = r.y + d_min * u.y;
  r.z = r.z + d_min * u.z;
  d_flight -= d_min;
  return true;
}

// track_length_mesh_tally.cpp:328-393
__device__ __forceinline__ void mesh_distance_to_next(const DevTally& t, const V3& r, const V3& u, const MeshPos& p,
                                                       double ux_inv, double uy_inv, double uz_inv, double& dist, int& key) {
  const double xc = t.lowx + p.i * t.dx + 0.5 * t.dx;
  const double yc = t.lowy + p.j * t.dy + 0.5 * t.dy;
  const double zc = t.lowz + p.k * t.dz + 0.5 * t.dz;
  const double tx = r.x - xc, ty = r.y - yc, tz = r.z - zc;
  dist = ABL_INF;
  key = 0;
  const double diff_xl = -t.dx * 0.5 - tx;
  const double diff_xh = t.dx * 0.5 - tx;
  const double diff_yl = -t.dy * 0.5 - ty;
  const double diff_yh = t.dy * 0.5 - ty;
  const double diff_zl = -t.dz * 0.5 - tz;
  const double diff_zh = t.dz * 0.5 - tz;
  const double d_xl = diff_xl * ux_inv, d_xh = diff_xh * ux_inv;
  const double d_yl = diff_yl * uy_inv, d_yh = diff_yh * uy_inv;
  const double d_zl = diff_zl * uz_inv, d_zh = diff_zh * uz_inv;
  if (d_xl > 0. && d_xl < dist && p.on0 != -1) { dist = d_xl; key = -1; }
  if (d_xh > 0. && d_xh < dist && p.on0 != 1) { dist = d_xh; key = 1; }
  if (d_yl > 0. && d_yl < dist && p.on1 != -1) { dist = d_yl; key = -2; }
  if (d_yh > 0. && d_yh < dist && p.on1 != 1) { dist = d_yh; key = 2; }
  if (d_zl > 0. && d_zl < dist && p.on2 != -1) { dist = d_zl; key = -3; }
  if (d_zh > 0. && d_zh < dist && p.on2 != 1) { dist = d_zh; key = 3; }
}

// TrackLengthMeshTally::score_flight ; returns the number of bins scored
__device__ inline int score_flight(const DevTally& t, V3 r, const V3& u, double d, double E, double w, double w2,
                                   const MatXS& m) {
  MeshPos p;
  mesh_initialize_indices(t, r, u, p);
  const bool inside = (p.i >= 0 && p.i < t.Nx && p.j >= 0 && p.j < t.Ny && p.k >= 0 && p.k < t.Nz);
  if (!inside) {
    if (!mesh_find_entry_point(t, r, u, d)) return 0;
    mesh_initialize_indices(t, r, u, p);
  }
  double base = 1. / t.net_weight;  // track_length_mesh_tally.cpp:33-75
  if (t.quantity <= ABL_Q_IMAG_FLUX) base *= quantity_factor(t.quantity, w, w2, m);
  const int l = tally_energy_bin(t, E);
  if (l == -1) return 0;
  const double ux_inv = 1. / u.x, uy_inv = 1. / u.y, uz_inv = 1. / u.z;
  double distance_remaining = d;
  int nb = 0;
  while (distance_remaining > 0.) {
    double dn;
    int key;
    mesh_distance_to_next(t, r, u, p, ux_inv, uy_inv, uz_inv, dn, key);
    if (dn == ABL_INF) break;
    const double d_tile = fmin(dn, distance_remaining);
    if (p.i >= 0 && p.i < t.Nx && p.j >= 0 && p.j < t.Ny && p.k >= 0 && p.k < t.Nz) {
      red_add(t.gen + tally_index(t, l, p.i, p.j, p.k), d_tile * base);
      nb++;
    } else {
      return nb;
    }
    distance_remaining -= d_tile;
    if (distance_remaining <= 0.) break;
    r.x = r.x + d_tile * u.x;
    r.y = r.y + d_tile * u.y;
    r.z = r.z + d_tile * u.z;
    p.on0 = p.on1 = p.on2 = 0;  // update_indices, track_length_mesh_tally.cpp:395-435
    switch (key) {
      case -1: p.i--; p.on0 = 1; break;
      case 1: p.i++; p.on0 = -1; break;
      case -2: p.j--; p.on1 = 1; break;
      case 2: p.j++; p.on1 = -1; break;
      case -3: p.k--; p.on2 = 1; break;
      case 3: p.k++; p.on2 = -1; break;
      default: break;
    }
  }
  return nb;
}

}  // namespace abl
