/* orc_geom.h -- TEST INFRASTRUCTURE (oracle side).
 *
 * CPU restatement of the reference's CSG geometry and per-history geometry
 * cursor.  Each block cites the reference lines it follows (paths relative to
 * /root/reference).  The object model (Surface / Cell / Universe / lattice /
 * GeoLilyPad stack / Tracker) is kept on purpose so that rounding follows the
 * reference's incremental r_local updates (include/simulation/tracker.hpp:76-85).
 */
#ifndef ORC_GEOM_H
#define ORC_GEOM_H
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "orc_rng.h"

namespace orc {

constexpr double SURFACE_COINCIDENT = 1E-12;               // constants.hpp:62
constexpr double BOUNDRY_TOL = 500. * SURFACE_COINCIDENT;  // constants.hpp:63

// include/utils/vector.hpp:34-57
struct Vec {
  double x, y, z;
  double dot(const Vec& v) const { return x * v.x + y * v.y + z * v.z; }
  double norm() const { return std::sqrt(x * x + y * y + z * z); }
};
inline Vec operator+(const Vec& a, const Vec& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec operator-(const Vec& a, const Vec& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec operator*(const Vec& a, double d) { return {a.x * d, a.y * d, a.z * d}; }
inline Vec operator*(double d, const Vec& a) { return a * d; }

// include/utils/direction.hpp:37-43  (every Direction construction renormalises)
inline Vec make_direction(double x, double y, double z) {
  Vec v{x, y, z};
  double m = v.norm();
  return {v.x / m, v.y / m, v.z / m};
}
// include/utils/direction.hpp:44-64
inline Vec make_direction_mu_phi(double mu, double phi) {
  if (mu < -1.) mu = -1.; else if (mu > 1.) mu = 1.;
  if (phi < 0.) phi = 0.; else if (phi > 2 * PI) phi = 2 * PI;
  Vec v{std::sqrt(1. - mu * mu) * g_math.cos(phi), std::sqrt(1. - mu * mu) * g_math.sin(phi), mu};
  double m = v.norm();
  return {v.x / m, v.y / m, v.z / m};
}
// include/utils/direction.hpp:130-151
inline Vec rotate_direction(Vec u, double mu, double phi) {
  double cos = g_math.cos(phi);
  double sin = g_math.sin(phi);
  double sqrt_mu = std::sqrt(1. - mu * mu);
  double sqrt_w = std::sqrt(1. - u.z * u.z);
  double ux, uy, uz;
  if (sqrt_w > 1.E-10) {
    ux = mu * u.x + sqrt_mu * (u.x * u.z * cos - u.y * sin) / sqrt_w;
    uy = mu * u.y + sqrt_mu * (u.y * u.z * cos + u.x * sin) / sqrt_w;
    uz = mu * u.z - sqrt_mu * sqrt_w * cos;
  } else {
    double sqrt_v = std::sqrt(1. - u.y * u.y);
    ux = mu * u.x + sqrt_mu * (u.x * u.y * cos + u.z * sin) / sqrt_v;
    uy = mu * u.y - sqrt_mu * sqrt_v * cos;
    uz = mu * u.z + sqrt_mu * (u.y * u.z * cos - u.x * sin) / sqrt_v;
  }
  return make_direction(ux, uy, uz);
}

enum BoundaryType { BC_VACUUM = 0, BC_REFLECTIVE = 1, BC_NORMAL = 2 };  // surface.hpp:34
enum SurfType { S_XPLANE, S_YPLANE, S_ZPLANE, S_PLANE, S_XCYL, S_YCYL, S_ZCYL, S_CYL, S_SPHERE };

struct Surface {
  int type = S_XPLANE;
  int bc = BC_NORMAL;
  uint32_t id = 0;
  // parameters: planes x0|y0|z0 or A,B,C,D; cylinders centre + R; general
  // cylinder x0,y0,z0,(alpha,beta,gamma),R ; sphere x0,y0,z0,R
  double p[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  // src/cylinder.cpp:28-58 : normalise axis, alpha=1-u0^2 ...
  void finish_general_cylinder(double u0, double v0, double w0) {
    const double mag = std::sqrt(u0 * u0 + v0 * v0 + w0 * w0);
    u0 /= mag; v0 /= mag; w0 /= mag;
    p[3] = 1. - u0 * u0;
    p[4] = 1. - v0 * v0;
    p[5] = 1. - w0 * w0;
  }

  Vec norm(const Vec& r) const {
    switch (type) {
      case S_XPLANE: return make_direction(1., 0., 0.);              // xplane.cpp:56
      case S_YPLANE: return make_direction(0., 1., 0.);              // yplane.cpp
      case S_ZPLANE: return make_direction(0., 0., 1.);              // zplane.cpp
      case S_PLANE: return make_direction(p[0], p[1], p[2]);         // plane.cpp:60
      case S_XCYL: return make_direction(0., r.y - p[0], r.z - p[1]);  // xcylinder.cpp:76
      case S_YCYL: return make_direction(r.x - p[0], 0., r.z - p[1]);  // ycylinder.cpp:76
      case S_ZCYL: return make_direction(r.x - p[0], r.y - p[1], 0.);  // zcylinder.cpp:78
      case S_CYL:
        return make_direction(p[3] * (r.x - p[0]), p[4] * (r.y - p[1]), p[5] * (r.z - p[2]));
      default: return make_direction(r.x - p[0], r.y - p[1], r.z - p[2]);  // sphere.cpp:78
    }
  }

  double eval(const Vec& r) const {
    switch (type) {
      case S_XPLANE: return r.x - p[0];
      case S_YPLANE: return r.y - p[0];
      case S_ZPLANE: return r.z - p[0];
      case S_PLANE: return p[0] * r.x + p[1] * r.y + p[2] * r.z - p[3];  // plane.cpp:37
      case S_XCYL: { const double y = r.y - p[0], z = r.z - p[1]; return y * y + z * z - p[2] * p[2]; }
      case S_YCYL: { const double x = r.x - p[0], z = r.z - p[1]; return x * x + z * z - p[2] * p[2]; }
      case S_ZCYL: { const double x = r.x - p[0], y = r.y - p[1]; return y * y + x * x - p[2] * p[2]; }  // zcylinder.cpp:36
      case S_CYL: {
        const double x = r.x - p[0], y = r.y - p[1], z = r.z - p[2];
        return p[3] * x * x + p[4] * y * y + p[5] * z * z - p[6] * p[6];
      }
      default: {
        const double x = r.x - p[0], y = r.y - p[1], z = r.z - p[2];
        return (x * x) + (y * y) + (z * z) - p[3] * p[3];
      }
    }
  }

  // e.g. src/xplane.cpp:32-43, src/zcylinder.cpp:33-46
  int sign(const Vec& r, const Vec& u) const {
    const double e = eval(r);
    if (e > SURFACE_COINCIDENT) return 1;
    if (e < -SURFACE_COINCIDENT) return -1;
    if (u.dot(norm(r)) > 0.) return 1;
    return -1;
  }

  static double quadric(double a, double k, double c, bool on_surf) {
    // src/zcylinder.cpp:58-76 (a-form) -- sphere uses a == 1 without the division
    const double quad = k * k - a * c;
    if (quad < 0.) return INF;
    if (on_surf || std::abs(c) < SURFACE_COINCIDENT) {
      if (k >= 0.) return INF;
      return (-k + std::sqrt(quad)) / a;
    } else if (c < 0.) {
      return (-k + std::sqrt(quad)) / a;
    } else {
      const double d = (-k - std::sqrt(quad)) / a;
      if (d < 0.) return INF;
      return d;
    }
  }

  double distance(const Vec& r, const Vec& u, bool on_surf) const {
    switch (type) {
      case S_XPLANE: case S_YPLANE: case S_ZPLANE: {  // xplane.cpp:45-54
        const double rc = type == S_XPLANE ? r.x : (type == S_YPLANE ? r.y : r.z);
        const double uc = type == S_XPLANE ? u.x : (type == S_YPLANE ? u.y : u.z);
        const double diff = p[0] - rc;
        if (on_surf || std::abs(diff) < SURFACE_COINCIDENT || uc == 0.) return INF;
        if (diff / uc < 0.) return INF;
        return diff / uc;
      }
      case S_PLANE: {  // plane.cpp:47-58
        const double num = p[3] - p[0] * r.x - p[1] * r.y - p[2] * r.z;
        const double denom = p[0] * u.x + p[1] * u.y + p[2] * u.z;
        const double d = num / denom;
        if (on_surf || std::abs(d) < SURFACE_COINCIDENT || denom == 0.) return INF;
        if (d < 0.) return INF;
        return d;
      }
      case S_XCYL: {  // xcylinder.cpp:46-72
        const double a = u.y * u.y + u.z * u.z;
        if (a == 0.) return INF;
        const double y = r.y - p[0], z = r.z - p[1];
        const double k = y * u.y + z * u.z;
        const double c = y * y + z * z - p[2] * p[2];
        return quadric(a, k, c, on_surf);
      }
      case S_YCYL: {  // ycylinder.cpp:46-72
        const double a = u.x * u.x + u.z * u.z;
        if (a == 0.) return INF;
        const double x = r.x - p[0], z = r.z - p[1];
        const double k = x * u.x + z * u.z;
        const double c = x * x + z * z - p[2] * p[2];
        return quadric(a, k, c, on_surf);
      }
      case S_ZCYL: {  // zcylinder.cpp:48-76
        const double a = u.y * u.y + u.x * u.x;
        if (a == 0.) return INF;
        const double x = r.x - p[0], y = r.y - p[1];
        const double k = y * u.y + x * u.x;
        const double c = y * y + x * x - p[2] * p[2];
        return quadric(a, k, c, on_surf);
      }
      case S_CYL: {  // cylinder.cpp:76-104
        const double a = p[3] * u.x * u.x + p[4] * u.y * u.y + p[5] * u.z * u.z;
        if (a == 0.) return INF;
        const double x = r.x - p[0], y = r.y - p[1], z = r.z - p[2];
        const double k = p[3] * x * u.x + p[4] * y * u.y + p[5] * z * u.z;
        const double c = p[3] * x * x + p[4] * y * y + p[5] * z * z - p[6] * p[6];
        return quadric(a, k, c, on_surf);
      }
      default: {  // sphere.cpp:50-74 (no division by a)
        const double x = r.x - p[0], y = r.y - p[1], z = r.z - p[2];
        const double k = x * u.x + y * u.y + z * u.z;
        const double c = x * x + y * y + z * z - p[3] * p[3];
        const double quad = k * k - c;
        if (quad < 0.) return INF;
        if (on_surf || std::abs(c) < SURFACE_COINCIDENT) {
          if (k >= 0.) return INF;
          return -k + std::sqrt(quad);
        } else if (c < 0.) {
          return -k + std::sqrt(quad);
        } else {
          const double d = -k - std::sqrt(quad);
          if (d < 0.) return INF;
          return d;
        }
      }
    }
  }
};

// include/geometry/cell.hpp:43-49
enum OP : int32_t {
  L_PAR = INT32_MAX, R_PAR = INT32_MAX - 1, COMP = INT32_MAX - 2, INTR = INT32_MAX - 3, UNIN = INT32_MAX - 4
};

struct Geometry;

struct Cell {
  std::vector<int32_t> rpn;
  bool simple = true;
  bool vac_or_refl = false;
  bool fill_universe = false;
  int material = -1;   // index into materials
  int universe = -1;   // index into universes
  uint32_t id = 0;
};

enum UniType { U_CELLS, U_RECT, U_HEX };

struct Universe {
  int type = U_CELLS;
  uint32_t id = 0;
  bool has_bc = false;
  std::vector<uint32_t> cell_indices;  // U_CELLS
  // U_RECT (src/rect_lattice.cpp:33-52)
  uint32_t Nx = 0, Ny = 0, Nz = 0;
  double Px = 0, Py = 0, Pz = 0, Px_inv = 0, Py_inv = 0, Pz_inv = 0, Xl = 0, Yl = 0, Zl = 0;
  std::vector<int32_t> lattice_universes;  // universe index or -1
  int32_t outer_universe_index = -1;
  // U_HEX (src/hex_lattice.cpp:30-56; include/geometry/hex_lattice.hpp:58-66): Nz and lattice_universes as above
  uint32_t Nrings = 0, width = 0, mid_qr = 0;
  int top = 0;  // 0 pointy, 1 flat
  double pitch = 0, pitch_z = 0, X_o = 0, Y_o = 0, Z_o = 0;
  double cos_pi_6 = 0, sin_pi_6 = 0, cos_pi_3 = 0, sin_pi_3 = 0;  // std::cos(PI / 6.0) ... as the reference evaluates them
};

// include/geometry/geo_lily_pad.hpp:33-49
struct GeoLilyPad {
  enum PadType { PUniverse, PLattice, PCell };
  PadType type = PUniverse;
  int index = 0;  // universe index or cell index (the reference stores ids + std::map lookups)
  Vec r_local{0, 0, 0};
  std::array<int32_t, 3> tile{0, 0, 0};
  bool in_lattice_outside_universe = false;
};

struct Boundary {  // include/geometry/boundary.hpp:33-42
  double distance;
  int surface_index;
  int boundary_type;
  int32_t token = 0;
  Boundary(double d, int i, int b) : distance(d), surface_index(i), boundary_type(b) {}
};

struct Geometry {
  std::vector<Surface> surfaces;
  std::vector<Cell> cells;
  std::vector<Universe> universes;
  int root = -1;
  std::map<uint32_t, int> surface_id_to_indx, cell_id_to_indx, universe_id_to_indx;

  // ---- Cell (src/cell.cpp) ------------------------------------------------
  bool cell_is_inside(const Cell& c, const Vec& r, const Vec& u, int32_t on_surf) const {
    if (c.simple) {  // cell.cpp:144-158
      for (const int32_t& token : c.rpn) {
        if (token == on_surf) {
        } else if (-token == on_surf)
          return false;
        else {
          int sign = surfaces[static_cast<size_t>(std::abs(token) - 1)].sign(r, u);
          if ((sign > 0 && token < 0) || (sign < 0 && token > 0)) return false;
        }
      }
      return true;
    }
    // cell.cpp:160-201
    std::vector<bool> stck(c.rpn.size());
    int i_stck = -1;
    for (int32_t token : c.rpn) {
      if (token == OP::UNIN) {
        stck[i_stck - 1] = stck[i_stck - 1] || stck[i_stck];
        i_stck--;
      } else if (token == OP::INTR) {
        stck[i_stck - 1] = stck[i_stck - 1] && stck[i_stck];
        i_stck--;
      } else if (token == OP::COMP) {
        stck[i_stck] = !stck[i_stck];
      } else {
        i_stck++;
        if (token == on_surf) {
          stck[i_stck] = true;
        } else if (-token == on_surf) {
          stck[i_stck] = false;
        } else {
          int sign = surfaces[static_cast<size_t>(std::abs(token) - 1)].sign(r, u);
          stck[i_stck] = ((sign > 0 && token > 0) || (sign < 0 && token < 0));
        }
      }
    }
    if (i_stck == 0) return stck[0];
    return true;
  }

  // cell.cpp:79-142 ; bc_only selects distance_to_boundary_condition
  std::pair<double, int32_t> cell_distance(const Cell& c, const Vec& r, const Vec& u, int32_t on_surf,
                                           bool bc_only) const {
    if (bc_only && !c.vac_or_refl) return {INF, 0};
    double min_dist = INF;
    int32_t i_surf = 0;
    for (int32_t token : c.rpn) {
      if (token >= OP::UNIN) continue;
      bool coincident = std::abs(token) == std::abs(on_surf);
      const Surface& surf = surfaces[static_cast<size_t>(std::abs(token) - 1)];
      if (bc_only && surf.bc == BC_NORMAL) continue;
      double d = surf.distance(r, u, coincident);
      if (d < min_dist) {
        if (std::abs(d - min_dist) / min_dist >= 1e-14) {
          min_dist = d;
          i_surf = -token;
        }
      }
    }
    return {min_dist, i_surf};
  }

  // ---- HexLattice (src/hex_lattice.cpp) -------------------------------------
  static std::array<int32_t, 2> hex_nearest(const Universe& L, const Vec& p) {  // get_nearest_hex :238-281
    double q, r, det;
    if (L.top == 1) {
      det = -L.pitch * L.pitch * L.sin_pi_3;
      q = (L.pitch / det) * (0. * p.x - 1. * p.y);
      r = (L.pitch / det) * (-L.sin_pi_3 * p.x + L.cos_pi_3 * p.y);
    } else {
      det = -L.pitch * L.pitch * L.cos_pi_6;
      q = (L.pitch / det) * (L.sin_pi_6 * p.x - L.cos_pi_6 * p.y);
      r = (L.pitch / det) * (-1. * p.x - 0. * p.y);
    }
    double x = q, z = r, y = -x - z;
    double rx = std::round(x), ry = std::round(y), rz = std::round(z);
    double x_diff = std::abs(rx - x), y_diff = std::abs(ry - y), z_diff = std::abs(rz - z);
    if (x_diff > y_diff && x_diff > z_diff) {
      rx = -ry - rz;
    } else if (y_diff > x_diff && y_diff > z_diff) {
      ry = -rx - rz;
    } else {
      rz = -rx - ry;
    }
    return {static_cast<int32_t>(rx), static_cast<int32_t>(rz)};
  }
  static std::array<int32_t, 3> hex_get_tile(const Universe& L, const Vec& p) {  // get_tile :283-288
    auto qr = hex_nearest(L, p);
    double Z_low = L.Z_o - 0.5 * static_cast<double>(L.Nz) * L.pitch_z;
    int32_t nz = static_cast<int32_t>(std::floor((p.z - Z_low) / L.pitch_z));
    return {qr[0], qr[1], nz};
  }
  static Vec hex_tile_center(const Universe& L, int q_, int r_, int nz) {  // get_hex_center + tile_center :290-331
    double x, y;
    double q = static_cast<double>(q_), r = static_cast<double>(r_);
    if (L.top == 1) {
      x = L.pitch * (L.cos_pi_3 * q + 1. * r);
      y = L.pitch * (L.sin_pi_3 * q + 0. * r);
    } else {
      x = L.pitch * (0. * q + L.cos_pi_6 * r);
      y = L.pitch * (1. * q + L.sin_pi_6 * r);
    }
    double Z_low = L.Z_o - 0.5 * static_cast<double>(L.Nz) * L.pitch_z;
    double z = (static_cast<double>(nz) + 0.5) * L.pitch_z + Z_low;
    return {x, y, z};
  }
  static uint32_t hex_ring(int32_t x, int32_t z) {  // get_ring :333-345
    int32_t y = -x - z;
    uint32_t ax = static_cast<uint32_t>(std::abs(x)), ay = static_cast<uint32_t>(std::abs(y)), az = static_cast<uint32_t>(std::abs(z));
    return std::max(std::max(ax, ay), az);
  }
  static size_t hex_linear_index(const Universe& L, int32_t q_, int32_t r_, int32_t z) {  // :224-236
    uint32_t q = static_cast<uint32_t>(q_) + L.mid_qr, r = static_cast<uint32_t>(r_) + L.mid_qr;
    return static_cast<size_t>(static_cast<uint32_t>(z) * (L.width * L.width) + r * L.width + q);
  }
  static double hex_distance_to_line(const Vec& r, const Vec& u, double x1, double y1, double x2, double y2) {  // :437-455
    double A = y2 - y1, B = x1 - x2;
    double D = (x2 - x1) * y1 - (y2 - y1) * x1;
    double num = D - A * r.x - B * r.y;
    double denom = A * u.x + B * u.y;
    double d = num / denom;
    if (d < 0.) return INFINITY;
    return d;
  }
  static double hex_distance_to_tile_boundary(const Universe& L, const Vec& r_local, const Vec& u,
                                              const std::array<int32_t, 3>& tile) {  // :347-435
    Vec center = hex_tile_center(L, tile[0], tile[1], tile[2]);
    Vec r_tile = r_local - center;
    double d1, d2, d3, d4, d5, d6;
    if (L.top == 0) {
      double x1 = 0., y1 = L.pitch / (2. * L.cos_pi_6), x2 = L.pitch / 2., y2 = L.pitch / (2. * L.sin_pi_6);
      d1 = hex_distance_to_line(r_tile, u, x1, y1, x2, y1);
      d2 = hex_distance_to_line(r_tile, u, x2, y2, x2, -y2);
      d3 = hex_distance_to_line(r_tile, u, x2, -y2, x1, -y1);
      d4 = hex_distance_to_line(r_tile, u, x1, -y1, -x2, -y2);
      d5 = hex_distance_to_line(r_tile, u, -x2, -y2, -x2, y2);
      d6 = hex_distance_to_line(r_tile, u, -x2, y2, x1, y1);
    } else {
      double x1 = L.pitch / (2. * L.sin_pi_6), y1 = L.pitch / 2., x2 = L.pitch / (2. * L.cos_pi_6), y2 = 0.;
      d1 = hex_distance_to_line(r_tile, u, x1, y1, x2, y1);
      d2 = hex_distance_to_line(r_tile, u, x2, y2, x1, -y1);
      d3 = hex_distance_to_line(r_tile, u, x1, -y1, -x1, -y1);
      d4 = hex_distance_to_line(r_tile, u, -x1, -y1, -x2, -y2);
      d5 = hex_distance_to_line(r_tile, u, -x2, -y2, -x1, y1);
      d6 = hex_distance_to_line(r_tile, u, -x1, y1, x1, y1);
    }
    double dzl = (-L.pitch_z * 0.5 - r_tile.z) / u.z;
    double dzu = (L.pitch_z * 0.5 - r_tile.z) / u.z;
    double d = INF;
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
  static bool hex_is_inside(const Universe& L, const Vec& r) {  // :58-81
    Vec r_o{r.x - L.X_o, r.y - L.Y_o, r.z - L.Z_o};
    auto qr = hex_nearest(L, r_o);
    if (hex_ring(qr[0], qr[1]) >= L.Nrings) return false;
    double Z_low = L.Z_o - 0.5 * static_cast<double>(L.Nz) * L.pitch_z;
    int32_t nz = static_cast<int32_t>(std::floor((r.z - Z_low) / L.pitch_z));
    if (nz < 0 || nz >= static_cast<int32_t>(L.Nz)) return false;
    return true;
  }

  // ---- RectLattice (src/rect_lattice.cpp) ----------------------------------
  static Vec tile_center(const Universe& L, int nx, int ny, int nz) {  // :303-309
    double x = (static_cast<double>(nx) + 0.5) * L.Px + L.Xl;
    double y = (static_cast<double>(ny) + 0.5) * L.Py + L.Yl;
    double z = (static_cast<double>(nz) + 0.5) * L.Pz + L.Zl;
    return {x, y, z};
  }
  static std::array<int32_t, 3> get_tile(const Universe& L, const Vec& r, const Vec& u) {  // :209-237
    if (L.type == U_HEX) return hex_get_tile(L, r);  // (the position as handed in: Tracker passes the pad's un-shifted r_local)
    if (L.type != U_RECT) return {0, 0, 0};  // universe.cpp:31-34
    int32_t nx = static_cast<int32_t>(std::floor((r.x - L.Xl) * L.Px_inv));
    int32_t ny = static_cast<int32_t>(std::floor((r.y - L.Yl) * L.Py_inv));
    int32_t nz = static_cast<int32_t>(std::floor((r.z - L.Zl) * L.Pz_inv));
    Vec rt = tile_center(L, nx, ny, nz);
    double xl = rt.x - L.Px * 0.5;
    if (std::abs(xl - r.x) < SURFACE_COINCIDENT && u.x < 0.) nx--;
    double xh = rt.x + L.Px * 0.5;
    if (std::abs(xh - r.x) < SURFACE_COINCIDENT && u.x >= 0.) nx++;
    double yl = rt.y - L.Py * 0.5;
    if (std::abs(yl - r.y) < SURFACE_COINCIDENT && u.y < 0.) ny--;
    double yh = rt.y + L.Py * 0.5;
    if (std::abs(yh - r.y) < SURFACE_COINCIDENT && u.y >= 0.) ny++;
    double zl = rt.z - L.Pz * 0.5;
    if (std::abs(zl - r.z) < SURFACE_COINCIDENT && u.z < 0.) nz--;
    double zh = rt.z + L.Pz * 0.5;
    if (std::abs(zh - r.z) < SURFACE_COINCIDENT && u.z >= 0.) nz++;
    return {nx, ny, nz};
  }
  static double distance_to_tile_boundary(const Universe& L, const Vec& r_local, const Vec& u,
                                          const std::array<int32_t, 3>& tile) {  // :239-282
    if (L.type == U_HEX) return hex_distance_to_tile_boundary(L, r_local, u, tile);
    if (L.type != U_RECT) return INF;  // universe.cpp:36-40
    Vec center = tile_center(L, tile[0], tile[1], tile[2]);
    Vec r_tile = r_local - center;
    double dist = INF;
    const double diff_xl = -L.Px * 0.5 - r_tile.x;
    const double diff_xh = L.Px * 0.5 - r_tile.x;
    const double diff_yl = -L.Py * 0.5 - r_tile.y;
    const double diff_yh = L.Py * 0.5 - r_tile.y;
    const double diff_zl = -L.Pz * 0.5 - r_tile.z;
    const double diff_zh = L.Pz * 0.5 - r_tile.z;
    const double ux_inv = 1. / u.x;
    const double uy_inv = 1. / u.y;
    const double uz_inv = 1. / u.z;
    const double d_xl = diff_xl * ux_inv;
    const double d_xh = diff_xh * ux_inv;
    const double d_yl = diff_yl * uy_inv;
    const double d_yh = diff_yh * uy_inv;
    const double d_zl = diff_zl * uz_inv;
    const double d_zh = diff_zh * uz_inv;
    if (d_xl > 0. && d_xl < dist && std::abs(diff_xl) > 100 * SURFACE_COINCIDENT) dist = d_xl;
    if (d_xh > 0. && d_xh < dist && std::abs(diff_xh) > 100 * SURFACE_COINCIDENT) dist = d_xh;
    if (d_yl > 0. && d_yl < dist && std::abs(diff_yl) > 100 * SURFACE_COINCIDENT) dist = d_yl;
    if (d_yh > 0. && d_yh < dist && std::abs(diff_yh) > 100 * SURFACE_COINCIDENT) dist = d_yh;
    if (d_zl > 0. && d_zl < dist && std::abs(diff_zl) > 100 * SURFACE_COINCIDENT) dist = d_zl;
    if (d_zh > 0. && d_zh < dist && std::abs(diff_zh) > 100 * SURFACE_COINCIDENT) dist = d_zh;
    return dist;
  }

  // ---- Universe::get_cell(stack, ...) ---------------------------------------
  // returns cell index or -1 (lost).  cell_universe.cpp:72-109, rect_lattice.cpp:132-207
  int get_cell(int uni, std::vector<GeoLilyPad>& stack, Vec r, const Vec& u, int32_t on_surf) const {
    const Universe& U = universes[static_cast<size_t>(uni)];
    if (U.type == U_CELLS) {
      stack.push_back({GeoLilyPad::PUniverse, uni, r, {0, 0, 0}, false});
      for (size_t i = 0; i < U.cell_indices.size(); i++) {
        const uint32_t indx = U.cell_indices[i];
        const Cell& c = cells[indx];
        if (cell_is_inside(c, r, u, on_surf)) {
          stack.push_back({GeoLilyPad::PCell, static_cast<int>(indx), r, {0, 0, 0}, false});
          if (!c.fill_universe) return static_cast<int>(indx);
          return get_cell(c.universe, stack, r, u, on_surf);
        }
      }
      return -1;
    }
    if (U.type == U_HEX) {  // HexLattice::get_cell(stack, ...) hex_lattice.cpp:140-202
      Vec r_o{r.x - U.X_o, r.y - U.Y_o, r.z - U.Z_o};
      auto qrz = hex_get_tile(U, r_o);
      bool in = hex_ring(qrz[0], qrz[1]) < U.Nrings && !(qrz[2] < 0 || qrz[2] >= static_cast<int32_t>(U.Nz));
      size_t indx = 0;
      if (in) {
        indx = hex_linear_index(U, qrz[0], qrz[1], qrz[2]);
        in = U.lattice_universes[indx] != -1;
      }
      if (!in) {
        if (U.outer_universe_index == -1) {
          stack.push_back({GeoLilyPad::PLattice, uni, r, qrz, false});
          return -1;
        }
        stack.push_back({GeoLilyPad::PLattice, uni, r, qrz, true});
        return get_cell(U.outer_universe_index, stack, r, u, on_surf);
      }
      Vec r_tile = r_o - hex_tile_center(U, qrz[0], qrz[1], qrz[2]);
      stack.push_back({GeoLilyPad::PLattice, uni, r, qrz, false});
      return get_cell(U.lattice_universes[indx], stack, r_tile, u, on_surf);
    }
    auto tile = get_tile(U, r, u);
    int nx = tile[0], ny = tile[1], nz = tile[2];
    if ((nx < 0 || nx >= static_cast<int>(U.Nx)) || (ny < 0 || ny >= static_cast<int>(U.Ny)) ||
        (nz < 0 || nz >= static_cast<int>(U.Nz))) {
      if (U.outer_universe_index >= 0) {
        stack.push_back({GeoLilyPad::PLattice, uni, r, {nx, ny, nz}, true});
        return get_cell(U.outer_universe_index, stack, r, u, on_surf);
      }
      stack.push_back({GeoLilyPad::PLattice, uni, r, {nx, ny, nz}, false});
      return -1;
    }
    // linear_index: rect_lattice.cpp:293-301
    const size_t lin = static_cast<size_t>(static_cast<uint32_t>(nz) * (U.Nx * U.Ny) +
                                           static_cast<uint32_t>(nx) * U.Ny + static_cast<uint32_t>(ny));
    if (U.lattice_universes[lin] >= 0) {
      Vec r_local = r - tile_center(U, nx, ny, nz);
      stack.push_back({GeoLilyPad::PLattice, uni, r, {nx, ny, nz}, false});
      return get_cell(U.lattice_universes[lin], stack, r_local, u, on_surf);
    }
    if (U.outer_universe_index >= 0) {
      stack.push_back({GeoLilyPad::PLattice, uni, r, {nx, ny, nz}, true});
      return get_cell(U.outer_universe_index, stack, r, u, on_surf);
    }
    stack.push_back({GeoLilyPad::PLattice, uni, r, {nx, ny, nz}, false});
    return -1;
  }

  // Universe::get_boundary_condition : cell_universe.cpp:111-154, lattice.cpp:77-92
  Boundary universe_boundary_condition(int uni, const Vec& r, const Vec& u, int32_t on_surf) const {
    const Universe& U = universes[static_cast<size_t>(uni)];
    if (U.type != U_CELLS) {
      if (U.has_bc) return universe_boundary_condition(U.outer_universe_index, r, u, on_surf);
      Boundary b(INF, -1, BC_VACUUM);
      b.token = 0;
      return b;
    }
    double dist = INF;
    int btype = BC_VACUUM;
    int surface_index = -1;
    int32_t token = 0;
    if (U.has_bc) {
      for (auto indx : U.cell_indices) {
        const Cell& cell = cells[indx];
        if (!cell.vac_or_refl) continue;
        auto d_t = cell_distance(cell, r, u, on_surf, true);
        if (d_t.first < dist && std::abs(d_t.first - dist) > BOUNDRY_TOL) {
          double tmp_dist = d_t.first;
          int32_t tmp_token = std::abs(d_t.second);
          if (tmp_token) {
            token = tmp_token;
            dist = tmp_dist;
            surface_index = token - 1;
          } else {
            continue;
          }
          btype = surfaces[static_cast<size_t>(surface_index)].bc;
          if (surfaces[static_cast<size_t>(surface_index)].sign(r, u) < 0) token *= -1;
        }
      }
    }
    Boundary b(dist, surface_index, btype);
    b.token = token;
    return b;
  }

  // lost_get_boundary : cell_universe.cpp:172-230 (cells), lattice.cpp:94-108
  Boundary universe_lost_get_boundary(int uni, const Vec& r, const Vec& u, int32_t on_surf) const {
    const Universe& U = universes[static_cast<size_t>(uni)];
    if (U.type == U_HEX) {  // Lattice::lost_get_boundary with HexLattice::is_inside (lattice.cpp:94-108, hex_lattice.cpp:58-81)
      if (U.outer_universe_index >= 0 && !hex_is_inside(U, r)) return universe_lost_get_boundary(U.outer_universe_index, r, u, on_surf);
      Boundary b(hex_distance_to_tile_boundary(U, r, u, hex_get_tile(U, r)), -1, BC_NORMAL);
      b.token = 0;
      return b;
    }
    if (U.type != U_CELLS) {
      auto tile = get_tile(U, r, u);
      bool inside = !((tile[0] < 0 || tile[0] >= (int)U.Nx) || (tile[1] < 0 || tile[1] >= (int)U.Ny) ||
                      (tile[2] < 0 || tile[2] >= (int)U.Nz));
      if (inside) {
        size_t lin = (size_t)((uint32_t)tile[2] * (U.Nx * U.Ny) + (uint32_t)tile[0] * U.Ny + (uint32_t)tile[1]);
        inside = U.lattice_universes[lin] >= 0;
      }
      if (U.outer_universe_index >= 0 && !inside)
        return universe_lost_get_boundary(U.outer_universe_index, r, u, on_surf);
      Boundary b(distance_to_tile_boundary(U, r, u, tile), -1, BC_NORMAL);
      b.token = 0;
      return b;
    }
    double dist = INF;
    int btype = BC_VACUUM;
    int surface_index = -1;
    int32_t token = 0;
    for (auto indx : U.cell_indices) {
      const Cell& cell = cells[indx];
      auto d_t = cell_distance(cell, r, u, on_surf, false);
      if (d_t.first < dist && std::abs(d_t.first - dist) > BOUNDRY_TOL) {
        double tmp_dist = d_t.first;
        int32_t tmp_token = std::abs(d_t.second);
        if (tmp_token) {
          token = tmp_token;
          dist = tmp_dist;
          surface_index = token - 1;
        } else {
          continue;
        }
        btype = surfaces[static_cast<size_t>(surface_index)].bc;
        if (surfaces[static_cast<size_t>(surface_index)].sign(r, u) < 0) token *= -1;
      }
    }
    Boundary b(dist, surface_index, btype);
    b.token = token;
    return b;
  }

  // ---- construction helpers ---------------------------------------------------
  // src/cell.cpp:250-299
  static std::vector<int32_t> infix_to_rpn(const std::vector<int32_t>& infix) {
    std::vector<int32_t> rpn, stack;
    for (const auto& token : infix) {
      if (token < OP::UNIN) {
        rpn.push_back(token);
      } else if (token < OP::R_PAR) {
        while (stack.size() > 0) {
          int32_t op = stack.back();
          if (op < OP::R_PAR && ((token == OP::COMP && token < op) || (token != OP::COMP && token <= op))) {
            rpn.push_back(op);
            stack.pop_back();
          } else {
            break;
          }
        }
        stack.push_back(token);
      } else if (token == OP::L_PAR) {
        stack.push_back(token);
      } else {
        while (true) {
          if (stack.empty()) throw std::runtime_error("Mismatched parentheses in cell region definition.");
          if (stack.back() == OP::L_PAR) break;
          rpn.push_back(stack.back());
          stack.pop_back();
        }
        stack.pop_back();
      }
    }
    while (stack.size() > 0) {
      int32_t op = stack.back();
      if (op >= OP::R_PAR) throw std::runtime_error("Mismatched parentheses in cell region definition.");
      rpn.push_back(op);
      stack.pop_back();
    }
    return rpn;
  }

  // src/cell.cpp:301-361 (region string -> tokens) + :203-246 (check_for_bc, simplify)
  Cell make_cell(const std::string& region_str, uint32_t id) const {
    std::vector<int32_t> region;
    std::string temp;
    auto flush = [&]() {
      if (temp.size() > 0) {
        int32_t signed_id = std::stoi(temp);
        auto it = surface_id_to_indx.find(static_cast<uint32_t>(std::abs(signed_id)));
        if (it == surface_id_to_indx.end()) throw std::runtime_error("unknown surface id in region");
        int32_t indx = it->second + 1;
        if (signed_id < 0) indx *= -1;
        region.push_back(indx);
        temp = "";
      }
    };
    for (char c : region_str) {
      if (c == '&' || c == '(' || c == ')' || c == 'U' || c == '~') {
        flush();
        if (c == '&') region.push_back(OP::INTR);
        else if (c == '(') region.push_back(OP::L_PAR);
        else if (c == ')') region.push_back(OP::R_PAR);
        else if (c == 'U') region.push_back(OP::UNIN);
        else region.push_back(OP::COMP);
      } else if (c == '+' || c == '-' || (c >= '0' && c <= '9')) {
        temp += c;
      } else if (c != ' ') {
        throw std::runtime_error("Invalid character in cell region definition.");
      }
    }
    flush();
    Cell cell;
    cell.id = id;
    cell.rpn = infix_to_rpn(region);
    cell.simple = true;
    for (auto el : cell.rpn)
      if (el == OP::COMP || el == OP::UNIN) { cell.simple = false; break; }
    if (cell.simple) {
      std::vector<int32_t> kept;
      for (auto el : cell.rpn) if (el < OP::UNIN) kept.push_back(el);
      cell.rpn = kept;
    }
    for (int32_t token : cell.rpn) {
      if (token >= OP::UNIN) continue;
      int bc = surfaces[static_cast<size_t>(std::abs(token) - 1)].bc;
      if (bc == BC_VACUUM || bc == BC_REFLECTIVE) cell.vac_or_refl = true;
    }
    return cell;
  }

  void finalize_bc_flags() {
    // cell_universe.cpp:30-41 ; lattice.cpp:45-52 (outer universe decides)
    for (auto& U : universes) {
      if (U.type == U_CELLS) {
        U.has_bc = false;
        for (auto i : U.cell_indices) if (cells[i].vac_or_refl) { U.has_bc = true; break; }
      }
    }
    for (int pass = 0; pass < 8; pass++)
      for (auto& U : universes)
        if (U.type != U_CELLS) U.has_bc = U.outer_universe_index >= 0 && universes[(size_t)U.outer_universe_index].has_bc;
  }
};

// ---- Tracker (include/simulation/tracker.hpp) ------------------------------------
struct Tracker {
  const Geometry* geo;
  Vec r_, u_;
  std::vector<GeoLilyPad> tree;
  int current_cell = -1;
  int current_mat = -1;
  int32_t surface_token_ = 0;

  Tracker(const Geometry* g, Vec r, Vec u, int32_t token = 0) : geo(g), r_(r), u_(u), surface_token_(token) {
    tree.reserve(10);
    current_cell = geo->get_cell(geo->root, tree, r_, u_, surface_token_);
    if (current_cell >= 0) current_mat = geo->cells[(size_t)current_cell].material;
  }
  void set_r(Vec r) { r_ = r; surface_token_ = 0; }
  void set_u(Vec u) { u_ = u; }
  bool is_lost() const { return current_cell < 0; }

  void restart_get_current() {  // :63-74
    tree.clear();
    current_cell = geo->get_cell(geo->root, tree, r_, u_, surface_token_);
    if (current_cell >= 0) {
      if (geo->cells[(size_t)current_cell].fill_universe) throw std::runtime_error("Did not find a cell with a material.");
      current_mat = geo->cells[(size_t)current_cell].material;
    } else {
      current_mat = -1;
    }
  }
  void move(double d) {  // :76-85
    r_ = r_ + d * u_;
    for (auto& leaf : tree) leaf.r_local = leaf.r_local + d * u_;
    surface_token_ = 0;
  }
  bool check_tree() const {  // :308-312
    return r_.x == tree.front().r_local.x && r_.y == tree.front().r_local.y && r_.z == tree.front().r_local.z;
  }

  Boundary get_boundary_condition() const {  // :94-161
    if (!is_lost()) {
      double dist = INF;
      int btype = BC_VACUUM;
      int surface_index = -1;
      int32_t token = 0;
      for (const auto& pad : tree) {
        if (pad.type == GeoLilyPad::PCell) {
          const Cell& cell = geo->cells[(size_t)pad.index];
          if (!cell.vac_or_refl) continue;
          auto d_t = geo->cell_distance(cell, pad.r_local, u_, surface_token_, true);
          if (d_t.first < dist && std::abs(d_t.first - dist) > BOUNDRY_TOL) {
            double tmp_dist = d_t.first;
            int32_t tmp_token = std::abs(d_t.second);
            if (tmp_token) {
              dist = tmp_dist;
              token = tmp_token;
              surface_index = token - 1;
            } else {
              continue;
            }
            btype = geo->surfaces[(size_t)surface_index].bc;
            if (geo->surfaces[(size_t)surface_index].sign(pad.r_local, u_) < 0) token *= -1;
          }
        } else {
          const Universe& uni = geo->universes[(size_t)pad.index];
          if (uni.has_bc) {
            Boundary ub = geo->universe_boundary_condition(pad.index, pad.r_local, u_, surface_token_);
            if (ub.distance < dist && std::abs(ub.distance - dist) > BOUNDRY_TOL) {
              dist = ub.distance;
              token = ub.token;
              surface_index = ub.surface_index;
              btype = ub.boundary_type;
            }
          }
        }
      }
      Boundary b(dist, surface_index, btype);
      b.token = token;
      return b;
    }
    return geo->universe_boundary_condition(geo->root, r_, u_, surface_token_);
  }

  Boundary get_nearest_boundary() const {  // :163-225
    if (!is_lost()) {
      auto bound = get_boundary_condition();
      double dist = bound.distance;
      int btype = bound.boundary_type;
      int surface_index = bound.surface_index;
      int32_t token = bound.token;
      for (const auto& pad : tree) {
        if (pad.type == GeoLilyPad::PLattice) {
          double d = Geometry::distance_to_tile_boundary(geo->universes[(size_t)pad.index], pad.r_local, u_, pad.tile);
          if (d < dist && std::abs(d - dist) > BOUNDRY_TOL) {
            dist = d;
            btype = BC_NORMAL;
            surface_index = -1;
            token = 0;
          }
        } else if (pad.type == GeoLilyPad::PCell) {
          auto d_t = geo->cell_distance(geo->cells[(size_t)pad.index], pad.r_local, u_, surface_token_, false);
          if (d_t.first < dist && std::abs(d_t.first - dist) > BOUNDRY_TOL) {
            dist = d_t.first;
            token = std::abs(d_t.second);
            surface_index = token ? token - 1 : -1;
            if (surface_index >= 0) btype = geo->surfaces[(size_t)surface_index].bc;
            else btype = BC_NORMAL;
            if (surface_index >= 0 && geo->surfaces[(size_t)surface_index].sign(pad.r_local, u_) < 0) token *= -1;
          }
        }
      }
      Boundary b(dist, surface_index, btype);
      b.token = token;
      return b;
    }
    return geo->universe_lost_get_boundary(geo->root, r_, u_, surface_token_);
  }

  void cross_surface(const Boundary& d_t) {  // :227-231
    move(d_t.distance);
    surface_token_ = -d_t.token;
  }

  void get_current() {  // :235-306
    size_t first_bad = tree.size();
    if (!check_tree()) {
      restart_get_current();
      if (!check_tree()) throw std::runtime_error("BAD POSITIONS!");
    }
    for (size_t it = 0; it < tree.size(); it++) {
      if (tree[it].type == GeoLilyPad::PCell) {
        if (!geo->cell_is_inside(geo->cells[(size_t)tree[it].index], tree[it].r_local, u_, surface_token_)) {
          first_bad = it;
          break;
        }
      } else if (tree[it].type == GeoLilyPad::PLattice) {
        auto tile = Geometry::get_tile(geo->universes[(size_t)tree[it].index], tree[it].r_local, u_);
        if (tree[it].tile[0] != tile[0] || tree[it].tile[1] != tile[1] || tree[it].tile[2] != tile[2]) {
          first_bad = it;
          break;
        }
      }
    }
    if (first_bad != tree.size()) {
      size_t size = first_bad;
      if (size == 0) return restart_get_current();
      tree.resize(size);
      // NOTE: the reference maps tree.back().id through universe_id_to_indx even when the
      // pad is a Cell pad (tracker.hpp:287); no shipped deck nests a lattice inside a
      // universe-filled cell, the only case where that differs.  We require a universe pad.
      if (tree.back().type == GeoLilyPad::PCell) throw std::runtime_error("lattice nested in a universe-filled cell");
      int uni_indx = tree.back().index;
      Vec r_local = tree.back().r_local;
      tree.pop_back();
      current_cell = geo->get_cell(uni_indx, tree, r_local, u_, surface_token_);
      if (current_cell < 0) restart_get_current();
      if (current_cell >= 0) {
        if (geo->cells[(size_t)current_cell].fill_universe) throw std::runtime_error("Did not find a cell with a material.");
        current_mat = geo->cells[(size_t)current_cell].material;
      }
    }
  }
};

}  // namespace orc
#endif
