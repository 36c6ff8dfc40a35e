/* orc_tally.h -- TEST INFRASTRUCTURE (oracle side).
 *
 * Scalar generation scores and mesh tallies, restating
 *   src/tallies.cpp:100-203,268-280, src/mesh_tally.cpp:66-152,
 *   src/collision_mesh_tally.cpp:33-110, src/track_length_mesh_tally.cpp:34-435,
 *   src/source_mesh_tally.cpp:30-78.
 */
#ifndef ORC_TALLY_H
#define ORC_TALLY_H
#include <array>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "orc_phys.h"

namespace orc {

enum Estimator { EST_COLLISION = 0, EST_TRACK_LENGTH = 1, EST_SOURCE = 2 };
enum Quantity {
  Q_FLUX = 0, Q_TOTAL = 1, Q_ELASTIC = 2, Q_ABSORPTION = 3, Q_FISSION = 4, Q_MT = 5, Q_REAL_FLUX = 6,
  Q_IMG_FLUX = 7, Q_SOURCE = 8, Q_REAL_SOURCE = 9, Q_IMAG_SOURCE = 10
};

struct Counters {  // instrumentation the reference lacks
  uint64_t flights = 0, real_collisions = 0, virtual_collisions = 0, tl_bins = 0, fission_sites = 0,
           boundary_events = 0, lost_at_birth = 0, coll_scores = 0;
  void add(const Counters& o) {
    flights += o.flights; real_collisions += o.real_collisions; virtual_collisions += o.virtual_collisions;
    tl_bins += o.tl_bins; fission_sites += o.fission_sites; boundary_events += o.boundary_events;
    lost_at_birth += o.lost_at_birth; coll_scores += o.coll_scores;
  }
};

struct MeshTally {
  std::string name;
  int estimator = EST_COLLISION, quantity = Q_FLUX;
  bool noise_source = false;  // source tally attached to the noise source list
  Vec r_low{0, 0, 0}, r_hi{0, 0, 0};
  uint64_t Nx = 1, Ny = 1, Nz = 1;
  uint64_t g = 0;
  double dx = 0, dy = 0, dz = 0, dx_inv = 0, dy_inv = 0, dz_inv = 0, net_weight = 1.;
  std::vector<double> energy_bounds;
  std::vector<double> tally_gen, tally_avg, tally_var;

  void init() {  // mesh_tally.cpp:66-116
    dx = (r_hi.x - r_low.x) / static_cast<double>(Nx);
    dy = (r_hi.y - r_low.y) / static_cast<double>(Ny);
    dz = (r_hi.z - r_low.z) / static_cast<double>(Nz);
    dx_inv = 1. / dx;
    dy_inv = 1. / dy;
    dz_inv = 1. / dz;
    size_t n = (energy_bounds.size() - 1) * Nx * Ny * Nz;
    tally_gen.assign(n, 0.);
    tally_avg.assign(n, 0.);
    tally_var.assign(n, 0.);
  }
  size_t idx(uint64_t e, uint64_t i, uint64_t j, uint64_t k) const { return ((e * Nx + i) * Ny + j) * Nz + k; }
  void add(size_t ix, double v) {
#pragma omp atomic
    tally_gen[ix] += v;
  }
  int energy_bin(double E) const {  // collision_mesh_tally.cpp:50-56 (<= E <=, first match)
    for (size_t e = 0; e < energy_bounds.size() - 1; e++)
      if (energy_bounds[e] <= E && E <= energy_bounds[e + 1]) return static_cast<int>(e);
    return -1;
  }

  // xs values of the CURRENT material/energy that the quantities may need
  struct MatXS { double Et, Ea, Ef, Eel; };

  void score_collision(const Particle& p, const MatXS& m, Counters& cn) {  // collision_mesh_tally.cpp:33-110
    double Et = m.Et;
    double scr = 1. / (Et * net_weight);
    int i = static_cast<int>(std::floor((p.state.position.x - r_low.x) * dx_inv));
    int j = static_cast<int>(std::floor((p.state.position.y - r_low.y) * dy_inv));
    int k = static_cast<int>(std::floor((p.state.position.z - r_low.z) * dz_inv));
    int l = energy_bin(p.E());
    if (l == -1) return;
    if (i >= 0 && i < static_cast<int>(Nx) && j >= 0 && j < static_cast<int>(Ny) && k >= 0 && k < static_cast<int>(Nz)) {
      switch (quantity) {
        case Q_FLUX: scr *= p.wgt(); break;
        case Q_ELASTIC: scr *= p.wgt() * m.Eel; break;
        case Q_ABSORPTION: scr *= p.wgt() * m.Ea; break;
        case Q_FISSION: scr *= p.wgt() * m.Ef; break;
        case Q_TOTAL: scr *= p.wgt() * Et; break;
        case Q_MT: scr *= p.wgt() * 0.; break;
        case Q_REAL_FLUX: scr *= p.wgt(); break;
        case Q_IMG_FLUX: scr *= p.wgt2(); break;
        default: break;
      }
      add(idx((uint64_t)l, (uint64_t)i, (uint64_t)j, (uint64_t)k), scr);
      cn.coll_scores++;
    }
  }

  double base_score(const Particle& p, const MatXS& m) const {  // track_length_mesh_tally.cpp:33-75
    double b = 1. / net_weight;
    switch (quantity) {
      case Q_FLUX: b *= p.wgt(); break;
      case Q_ELASTIC: b *= p.wgt() * m.Eel; break;
      case Q_ABSORPTION: b *= p.wgt() * m.Ea; break;
      case Q_FISSION: b *= p.wgt() * m.Ef; break;
      case Q_TOTAL: b *= p.wgt() * m.Et; break;
      case Q_MT: b *= p.wgt() * 0.; break;
      case Q_REAL_FLUX: b *= p.wgt(); break;
      case Q_IMG_FLUX: b *= p.wgt2(); break;
      default: break;
    }
    return b;
  }

  bool find_entry_point(Vec& r, const Vec& u, double& d_flight) const {  // :182-254
    const double ux_inv = 1. / u.x, uy_inv = 1. / u.y, uz_inv = 1. / u.z;
    double d_min = (r_low.x - r.x) * ux_inv;
    double d_max = (r_hi.x - r.x) * ux_inv;
    if (d_min > d_max) std::swap(d_min, d_max);
    double d_y_min = (r_low.y - r.y) * uy_inv;
    double d_y_max = (r_hi.y - r.y) * uy_inv;
    if (d_y_min > d_y_max) std::swap(d_y_min, d_y_max);
    if ((d_min > d_y_max) || (d_y_min > d_max)) return false;
    if (d_y_min > d_min) d_min = d_y_min;
    if (d_y_max < d_max) d_max = d_y_max;
    double d_z_min = (r_low.z - r.z) * uz_inv;
    double d_z_max = (r_hi.z - r.z) * uz_inv;
    if (d_z_min > d_z_max) std::swap(d_z_min, d_z_max);
    if ((d_min > d_z_max) || (d_z_min > d_max)) return false;
    if (d_z_min > d_min) d_min = d_z_min;
    if (d_z_max < d_max) d_max = d_z_max;
    if (d_max < d_min) std::swap(d_max, d_min);
    if ((d_max < 0.) && (d_min < 0.)) return false;
    if (d_min < 0.) return false;
    r = r + d_min * u;
    d_flight -= d_min;
    return true;
  }

  void initialize_indices(const Vec& r, const Vec& u, int& i, int& j, int& k, std::array<int, 3>& on) const {  // :256-326
    i = static_cast<int>(std::floor((r.x - r_low.x) * dx_inv));
    j = static_cast<int>(std::floor((r.y - r_low.y) * dy_inv));
    k = static_cast<int>(std::floor((r.z - r_low.z) * dz_inv));
    on.fill(0);
    double xc = r_low.x + i * dx + 0.5 * dx;
    double yc = r_low.y + j * dy + 0.5 * dy;
    double zc = r_low.z + k * dz + 0.5 * dz;
    double xl = xc - 0.5 * dx, xh = xc + 0.5 * dx;
    double yl = yc - 0.5 * dy, yh = yc + 0.5 * dy;
    double zl = zc - 0.5 * dz, zh = zc + 0.5 * dz;
    if (std::abs(xl - r.x) < SURFACE_COINCIDENT) {
      if (u.x < 0.) { i--; on[0] = 1; } else { on[0] = -1; }
    } else if (std::abs(xh - r.x) < SURFACE_COINCIDENT) {
      if (u.x < 0.) { on[0] = 1; } else { i++; on[0] = -1; }
    }
    if (std::abs(yl - r.y) < SURFACE_COINCIDENT) {
      if (u.y < 0.) { j--; on[1] = 1; } else { on[1] = -1; }
    } else if (std::abs(yh - r.y) < SURFACE_COINCIDENT) {
      if (u.y < 0.) { on[1] = 1; } else { j++; on[1] = -1; }
    }
    if (std::abs(zl - r.z) < SURFACE_COINCIDENT) {
      if (u.z < 0.) { k--; on[2] = 1; } else { on[2] = -1; }
    } else if (std::abs(zh - r.z) < SURFACE_COINCIDENT) {
      if (u.z < 0.) { on[2] = 1; } else { k++; on[2] = -1; }
    }
  }

  std::pair<double, int> distance_to_next_index(const Vec& r, const Vec& u, int i, int j, int k,
                                                const std::array<int, 3>& on) const {  // :328-393
    double xc = r_low.x + i * dx + 0.5 * dx;
    double yc = r_low.y + j * dy + 0.5 * dy;
    double zc = r_low.z + k * dz + 0.5 * dz;
    Vec r_tile{r.x - xc, r.y - yc, r.z - zc};
    double dist = INF;
    int key = 0;
    const double diff_xl = -dx * 0.5 - r_tile.x;
    const double diff_xh = dx * 0.5 - r_tile.x;
    const double diff_yl = -dy * 0.5 - r_tile.y;
    const double diff_yh = dy * 0.5 - r_tile.y;
    const double diff_zl = -dz * 0.5 - r_tile.z;
    const double diff_zh = dz * 0.5 - r_tile.z;
    const double ux_inv = 1. / u.x, uy_inv = 1. / u.y, uz_inv = 1. / u.z;
    const double d_xl = diff_xl * ux_inv, d_xh = diff_xh * ux_inv;
    const double d_yl = diff_yl * uy_inv, d_yh = diff_yh * uy_inv;
    const double d_zl = diff_zl * uz_inv, d_zh = diff_zh * uz_inv;
    if (d_xl > 0. && d_xl < dist && on[0] != -1) { dist = d_xl; key = -1; }
    if (d_xh > 0. && d_xh < dist && on[0] != 1) { dist = d_xh; key = 1; }
    if (d_yl > 0. && d_yl < dist && on[1] != -1) { dist = d_yl; key = -2; }
    if (d_yh > 0. && d_yh < dist && on[1] != 1) { dist = d_yh; key = 2; }
    if (d_zl > 0. && d_zl < dist && on[2] != -1) { dist = d_zl; key = -3; }
    if (d_zh > 0. && d_zh < dist && on[2] != 1) { dist = d_zh; key = 3; }
    return {dist, key};
  }

  static void update_indices(int key, int& i, int& j, int& k, std::array<int, 3>& on) {  // :395-435
    on.fill(0);
    switch (key) {
      case -1: i--; on[0] = 1; break;
      case 1: i++; on[0] = -1; break;
      case -2: j--; on[1] = 1; break;
      case 2: j++; on[1] = -1; break;
      case -3: k--; on[2] = 1; break;
      case 3: k++; on[2] = -1; break;
      default: break;
    }
  }

  void score_flight(const Particle& p, double d, const MatXS& m, Counters& cn) {  // :77-180
    Vec r = p.state.position;
    Vec u = p.state.direction;
    int i = 0, j = 0, k = 0;
    std::array<int, 3> on{0, 0, 0};
    initialize_indices(r, u, i, j, k, on);
    bool inside = (i >= 0 && i < (int)Nx && j >= 0 && j < (int)Ny && k >= 0 && k < (int)Nz);
    if (!inside) {
      if (!find_entry_point(r, u, d)) return;
      initialize_indices(r, u, i, j, k, on);
      // (reference warns if still outside; the loop below then returns at once)
    }
    double base = base_score(p, m);
    int l = energy_bin(p.E());
    if (l == -1) return;
    uint64_t uE = (uint64_t)l;
    double distance_remaining = d;
    while (distance_remaining > 0.) {
      auto next_tile = distance_to_next_index(r, u, i, j, k, on);
      if (next_tile.first == INF) break;
      double d_tile = std::min(next_tile.first, distance_remaining);
      if (i >= 0 && i < (int)Nx && j >= 0 && j < (int)Ny && k >= 0 && k < (int)Nz) {
        add(idx(uE, (uint64_t)i, (uint64_t)j, (uint64_t)k), d_tile * base);
        cn.tl_bins++;
      } else {
        return;
      }
      distance_remaining -= d_tile;
      if (distance_remaining <= 0.) break;
      r = r + d_tile * u;
      update_indices(next_tile.second, i, j, k, on);
    }
  }

  void score_source(const BankedParticle& p) {  // source_mesh_tally.cpp:30-78  (note: division by dx, not *dx_inv)
    int i = static_cast<int>(std::floor((p.r.x - r_low.x) / dx));
    int j = static_cast<int>(std::floor((p.r.y - r_low.y) / dy));
    int k = static_cast<int>(std::floor((p.r.z - r_low.z) / dz));
    int l = energy_bin(p.E);
    if (l == -1) return;
    if (i >= 0 && i < (int)Nx && j >= 0 && j < (int)Ny && k >= 0 && k < (int)Nz) {
      double scr = 1. / net_weight;
      switch (quantity) {
        case Q_IMAG_SOURCE: scr *= p.wgt2; break;
        default: scr *= p.wgt; break;
      }
      add(idx((uint64_t)l, (uint64_t)i, (uint64_t)j, (uint64_t)k), scr);
    }
  }

  void record_generation(double multiplier) {  // mesh_tally.cpp:121-150
    g++;
    const double dg = static_cast<double>(g);
    for (size_t i = 0; i < tally_gen.size(); i++) {
      double old_avg = tally_avg[i];
      double val = tally_gen[i] * multiplier;
      double avg = old_avg + (val - old_avg) / dg;
      tally_avg[i] = avg;
      double var = tally_var[i];
      var = var + (((val - old_avg) * (val - avg) - (var)) / dg);
      tally_var[i] = var;
    }
  }
  void clear_generation() { std::fill(tally_gen.begin(), tally_gen.end(), 0.); }
};

struct ThreadLocalScores { double k_col = 0, k_abs = 0, k_trk = 0, k_tot = 0, leakage = 0, mig = 0; };

struct Tallies {  // src/tallies.cpp
  double total_weight = 1.;
  double k_col_score = 0, k_abs_score = 0, k_trk_score = 0, leak_score = 0, k_tot_score = 0, mig_area_score = 0;
  double k_col = 1., k_col_avg = 0, k_col_var = 0, k_abs = 1., k_abs_avg = 0, k_abs_var = 0, k_trk = 1.,
         k_trk_avg = 0, k_trk_var = 0, leak = 0, leak_avg = 0, leak_var = 0, k_tot = 1., k_tot_avg = 0,
         k_tot_var = 0, mig = 0, mig_avg = 0, mig_var = 0;
  double keff_ = 1.;
  int gen = 0;
  std::vector<double> k_col_vec, k_abs_vec, k_trk_vec, leak_vec, mig_vec;
  std::vector<MeshTally> mesh;

  void clear_generation() {  // :143-157
    k_col_score = k_abs_score = k_trk_score = k_tot_score = leak_score = mig_area_score = 0.;
    for (auto& t : mesh) t.clear_generation();
  }
  void calc_gen_values() {  // :159-181
    k_col = k_col_score / total_weight;
    k_abs = k_abs_score / total_weight;
    k_trk = k_trk_score / total_weight;
    leak = leak_score / total_weight;
    k_tot = k_tot_score / total_weight;
    mig = mig_area_score / total_weight;
    k_col_vec.push_back(k_col);
    k_abs_vec.push_back(k_abs);
    k_trk_vec.push_back(k_trk);
    leak_vec.push_back(leak);
    mig_vec.push_back(mig);
  }
  void update_avg_and_var(double x, double& x_avg, double& x_var) const {  // :268-280
    double dgen = static_cast<double>(gen);
    double x_avg_old = x_avg, x_var_old = x_var;
    x_avg = x_avg_old + (x - x_avg_old) / (dgen);
    if (gen > 1) x_var = x_var_old + ((x - x_avg_old) * (x - x_avg_old) / (dgen)) - ((x_var_old) / (dgen - 1.));
  }
  void record_generation(double multiplier = 1.) {  // :183-203
    gen++;
    update_avg_and_var(k_col, k_col_avg, k_col_var);
    update_avg_and_var(k_abs, k_abs_avg, k_abs_var);
    update_avg_and_var(k_trk, k_trk_avg, k_trk_var);
    update_avg_and_var(leak, leak_avg, leak_var);
    update_avg_and_var(k_tot, k_tot_avg, k_tot_var);
    update_avg_and_var(mig, mig_avg, mig_var);
    for (auto& t : mesh) t.record_generation(multiplier);
  }
  double err(double var) const { return std::sqrt(var / static_cast<double>(gen)); }  // tallies.hpp:113-115
};

}  // namespace orc
#endif
