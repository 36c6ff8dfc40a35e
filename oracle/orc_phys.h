/* orc_phys.h -- TEST INFRASTRUCTURE (oracle side).
 *
 * Multigroup materials, samplers and particle state, restating
 *   src/mg_nuclide.cpp:73-118,343-543, include/materials/material_helper.hpp:47-224,
 *   include/materials/mg_angle_distribution.hpp:45-101, src/legendre_distribution.cpp:62-155,
 *   include/simulation/particle.hpp:38-243.
 * In MG mode every material has exactly one nuclide with atoms_bcm = 1
 * (src/material.cpp:50-62), so MaterialHelper reduces to one table row.
 */
#ifndef ORC_PHYS_H
#define ORC_PHYS_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "orc_geom.h"
#include "orc_rng.h"

namespace orc {

constexpr double TOLERANCE = 0.0001;  // constants.hpp:58

// include/materials/mg_angle_distribution.hpp
struct AngleDist {
  std::vector<double> mu{-1., 1.}, pdf{0.5, 0.5}, cdf{0., 1.};

  double sample_mu(Pcg32& rng) const {  // :45-60
    const double xi = rng_rand(rng);
    size_t l = static_cast<size_t>(std::lower_bound(cdf.begin(), cdf.end(), xi) - cdf.begin());
    if (xi == cdf[l]) return mu[l];
    l--;
    if (pdf[l] == pdf[l + 1]) return mu[l] + ((xi - cdf[l]) / pdf[l]);  // histogram_interp :92-94
    double m = (pdf[l + 1] - pdf[l]) / (mu[l + 1] - mu[l]);                // linear_interp :96-101
    return mu[l] + (1. / m) * (std::sqrt(pdf[l] * pdf[l] + 2. * m * (xi - cdf[l])) - pdf[l]);
  }
  double pdf_at(double x) const {  // MGAngleDistribution::pdf (mg_angle_distribution.hpp:62-75)
    if (x < mu.front()) return pdf.front();
    if (x > mu.back()) return pdf.back();
    size_t l = static_cast<size_t>(std::lower_bound(mu.begin(), mu.end(), x) - mu.begin());
    if (x == mu[l]) return pdf[l];
    l--;
    const double m = (pdf[l + 1] - pdf[l]) / (mu[l + 1] - mu[l]);
    return m * (x - mu[l]) + pdf[l];
  }
};

// include/materials/legendre_distribution.hpp:155-200 + src/legendre_distribution.cpp:105-155
struct Legendre {
  std::vector<double> a{0.5};
  void set_moment(size_t l, double coeff) {  // hpp:92-103
    if (l == 0) return;
    if (a.size() < l + 1) a.resize(l + 1, 0.);
    a[l] = coeff * (2. * static_cast<double>(l) + 1.) / 2.;
  }
  static double legendre(unsigned n, double x) {
    switch (n) {
      case 0: return 1.;
      case 1: return x;
      case 2: return 0.5 * (3. * x * x - 1.);
      case 3: return 0.5 * (5. * x * x * x - 3. * x);
      case 4: { const double x_2 = x * x; return 0.125 * (35. * x_2 * x_2 - 30. * x_2 + 3.); }
      default: {
        const double x_2 = x * x, x_3 = x_2 * x, x_4 = x_3 * x;
        double p3 = 0.5 * (5. * x_3 - 3. * x);
        double p4 = 0.125 * (35. * x_4 - 30. * x_2 + 3.);
        unsigned l = 4;
        while (l < n) {
          std::swap(p3, p4);
          p4 = ((2. * l + 1.) * x * p3 - l * p4) / (l + 1.);
          l++;
        }
        return p4;
      }
    }
  }
  double pdf(double mu) const {
    double p = 0;
    for (unsigned l = 0; l < a.size(); l++) p += a[l] * legendre(l, mu);
    return p;
  }
  AngleDist linearize() const {
    std::vector<double> mu{-1., 1.}, p;
    p.push_back(pdf(-1.));
    p.push_back(pdf(1.));
    size_t i = 0;
    while (i < (mu.size() - 1)) {
      double mu_mid = 0.5 * (mu[i] + mu[i + 1]);
      double p_interp = 0.5 * (p[i] + p[i + 1]);
      double p_real = pdf(mu_mid);
      double rel_diff = std::abs(p_interp - p_real) / p_real;
      if (rel_diff > TOLERANCE) {
        mu.insert(mu.begin() + static_cast<long>(i) + 1, mu_mid);
        p.insert(p.begin() + static_cast<long>(i) + 1, p_real);
      } else {
        i++;
      }
    }
    std::vector<double> cdf(mu.size(), 0.);
    for (size_t k = 0; k < mu.size() - 1; k++) cdf[k + 1] = ((mu[k + 1] - mu[k]) * 0.5 * (p[k + 1] + p[k])) + cdf[k];
    const double norm = cdf.back();
    for (size_t k = 0; k < cdf.size(); k++) {
      p[k] /= norm;
      cdf[k] /= norm;
    }
    AngleDist d;
    d.mu = mu;
    d.pdf = p;
    d.cdf = cdf;
    return d;
  }
};

struct MicroXS {  // include/materials/nuclide.hpp:41-55
  double total = 0, fission = 0, absorption = 0, elastic = 0, nu_total = 0, nu_delayed = 0, noise_copy = 0;
  size_t energy_index = 0;
};

struct Material {
  uint32_t id = 0;
  size_t G = 0;
  std::vector<double> Et, Ea, Ef, Es, nu_p, nu_d, speeds;
  bool has_nu_d = false;
  std::vector<std::vector<double>> chi, Ps;               // normalised rows (mg_nuclide.cpp:73-118)
  std::vector<std::vector<double>> chi_cp, Ps_cp;         // libstdc++ discrete tables of those rows
  std::vector<std::vector<AngleDist>> angle;
  std::vector<double> P_delayed_group, decay_constants, dg_cp;
  bool fissile = false;

  void finish() {
    Es.assign(G, 0.);
    for (size_t i = 0; i < G; i++) {  // make_scatter_xs :93-118
      Es[i] = 0.;
      for (size_t o = 0; o < G; o++) Es[i] += Ps[i][o];
      for (size_t o = 0; o < G; o++) Ps[i][o] /= Es[i];
    }
    for (size_t i = 0; i < G; i++) {  // normalize_chi :73-91
      double chi_i = 0.;
      for (size_t o = 0; o < G; o++) chi_i += chi[i][o];
      for (size_t o = 0; o < G; o++) chi[i][o] /= chi_i;
    }
    chi_cp.resize(G);
    Ps_cp.resize(G);
    for (size_t i = 0; i < G; i++) {
      chi_cp[i] = discrete_table(chi[i].data(), G);
      Ps_cp[i] = discrete_table(Ps[i].data(), G);
    }
    dg_cp = discrete_table(P_delayed_group.data(), P_delayed_group.size());
    fissile = false;
    for (size_t i = 0; i < G; i++) if (Ef[i] != 0. && (nu_p[i] + (has_nu_d ? nu_d[i] : 0.)) != 0.) fissile = true;
  }

  MicroXS micro(size_t g) const {  // mg_nuclide.cpp:394-411
    MicroXS xs;
    xs.energy_index = g;
    xs.total = Et[g];
    xs.fission = Ef[g];
    xs.absorption = xs.fission + (Ea[g] - Ef[g]);
    xs.elastic = Es[g];
    double nu_tot = nu_p[g];
    if (has_nu_d) nu_tot += nu_d[g];
    xs.nu_total = nu_tot;
    xs.nu_delayed = has_nu_d ? nu_d[g] : 0.;
    xs.noise_copy = 0.;
    return xs;
  }
};

struct Settings {
  enum Mode { K_EIGENVALUE, NOISE, FIXED_SOURCE, BRANCHLESS } mode = K_EIGENVALUE;  // FIXED_SOURCE: src/fixed_source.cpp (fission neutrons are secondaries)
  bool branchless_material = true, branchless_splitting = false, branchless_combing = true;  // settings.cpp:87-89
  enum Tracking { SURFACE, DELTA, CARTER, IMPLICIT_LEAKAGE } tracking = SURFACE;
  uint32_t ngroups = 1;
  std::vector<double> energy_bounds;
  int nparticles = 100000, ngenerations = 120, nignored = 20, nskip = 10;
  double wgt_cutoff = 0.25, wgt_survival = 1.0, wgt_split = 2.0;
  uint64_t rng_seed = 19073486328125ULL, rng_stride = 152917ULL;
  double min_energy = 0., max_energy = 100000.;
  std::vector<double> sample_xs_ratio;
  bool regional_cancellation = false, regional_cancellation_noise = false;
  int n_cancel_noise_gens = INT32_MAX;
  bool inner_generations = true, normalize_noise_source = true;
  double w_noise = -1., eta = 1., keff = 1.;

  size_t group(double E) const {  // MGNuclide::energy_grid_index, mg_nuclide.cpp:382-392
    size_t i = 0;
    for (i = 0; i < energy_bounds.size() - 1; i++)
      if (energy_bounds[i] <= E && E < energy_bounds[i + 1]) break;
    return i;
  }
  double group_mid(size_t g) const { return 0.5 * (energy_bounds[g] + energy_bounds[g + 1]); }
};

// include/simulation/particle.hpp:38-57 (only the fields the hot path consumes downstream)
struct BankedParticle {
  Vec r, u;
  double E, wgt, wgt2;
  uint64_t parent_history_id, parent_daughter_id, family_id;
  // what the exact cancelators read (particle.hpp:52-57): where the parent was before the flight that ended in this fission,
  // and the sampling cross section of that flight
  Vec parents_previous_position{0, 0, 0};
  double Esmp_parent = 0.;
  // ... and what `type: exact` reads on top (exact_mg_cancelator.cpp:319-327): the parent's direction and energy before its last
  // scatter, its energy at the fission, and whether the collision before was a virtual one
  Vec parents_previous_direction{1, 0, 0};
  double parents_previous_previous_energy = 0., parents_previous_energy = 0.;
  bool parents_previous_was_virtual = false;
};

struct ParticleState { Vec position, direction; double energy, weight, weight2; };

struct Particle {  // particle.hpp:68-243
  ParticleState state;
  uint64_t history_id = 0, family_id = 0, secondary_id = 0, daughter_counter_ = 0;
  std::vector<ParticleState> secondaries;
  std::vector<BankedParticle> history_fission_bank, history_noise_bank;
  bool alive = true, reflected = false, previous_collision_virtual = false;
  Vec previous_position{0, 0, 0}, r_birth{0, 0, 0};
  double Esmp_ = 0.;  // sampling cross section of the current flight (delta_tracker.cpp:111, carter_tracker.cpp:130)
  Vec previous_direction{1, 0, 0};  // Direction() (direction.hpp:36); set_direction / set_energy keep the value they replace
  double previous_energy = 0.;
  void set_direction(Vec u) { previous_direction = state.direction; state.direction = u; }
  void set_energy(double E) { previous_energy = state.energy; state.energy = E; }
  Pcg32 rng;
  // instrumentation (not in the reference): per-history integer outcomes
  uint32_t n_flights = 0, n_real = 0, n_virtual = 0, n_fission = 0, n_boundary = 0;
  uint64_t hash = 1469598103934665603ULL;
  void note(uint64_t v) { hash = (hash ^ v) * 1099511628211ULL; }

  Particle(Vec r, Vec u, double E, double w, uint64_t id = 0) : state{r, u, E, w, 0.}, history_id(id), r_birth(r) {}
  Vec& r() { return state.position; }
  Vec& u() { return state.direction; }
  double E() const { return state.energy; }
  double wgt() const { return state.weight; }
  double wgt2() const { return state.weight2; }
  uint64_t daughter_counter() { return daughter_counter_++; }
  void set_position(Vec r) { previous_position = state.position; state.position = r; }
  void move(double dist) {  // :125-133
    if (!reflected) {
      previous_position = state.position;
      state.position = state.position + dist * state.direction;
    } else {
      state.position = state.position + dist * state.direction;
      reflected = false;
    }
  }
  void kill() { alive = false; }
  void make_secondary(Vec u, double E, double w, double w2 = 0.) { secondaries.push_back({state.position, u, E, w, w2}); }
  void split(int n_new) {  // :165-173
    if (n_new > 1) {
      state.weight = state.weight / static_cast<double>(n_new);
      state.weight2 = state.weight2 / static_cast<double>(n_new);
      for (int np = 0; np < n_new - 1; np++) make_secondary(state.direction, state.energy, state.weight, state.weight2);
    }
  }
  void resurect() {  // :175-186
    if (!alive && !secondaries.empty()) {
      alive = true;
      state = secondaries.back();
      secondaries.pop_back();
      secondary_id++;
    }
  }
  void initialize_rng(uint64_t seed, uint64_t stride) {  // :188-193
    rng.seed(seed);
    rng.advance(stride * history_id);
  }
};

}  // namespace orc
#endif
