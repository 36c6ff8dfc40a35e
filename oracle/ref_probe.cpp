/* oracle/ref_probe.cpp -- C entry points around the REFERENCE'S OWN code, for pinning the oracle (test infrastructure).
 *
 * The reference cannot be built as a whole offline (yaml-cpp, PapillonNDL, pcg-cpp, HighFive, NDArray ... are fetched by
 * its CMake), but a part of the hot path compiles from its own sources as they lie under /root/reference:
 *   surfaces     src/{x,y,z}plane.cpp, plane.cpp, {x,y,z}cylinder.cpp, cylinder.cpp, sphere.cpp, surface.cpp
 *   directions   include/utils/direction.hpp (Direction constructors, rotate_direction)
 *   RNG          include/utils/rng.hpp over the pcg32 engine (pcg header vendored by pyarrow, see ref_shim/pcg_random.hpp)
 *   angles       src/mg_angle_distribution.cpp, src/legendre_distribution.cpp (sample_mu, linearize)
 * oracle/Makefile compiles those files in place (nothing is copied) together with this driver into
 * oracle/_ref/libabeille_ref.so.  tests/test_reference_pins.py compares the oracle's restatement with it bit for bit on
 * seeded inputs and keeps golden vectors generated from it (tests/golden/ref_pins.json) for machines without the reference.
 */
#include <geometry/surfaces/cylinder.hpp>
#include <geometry/surfaces/plane.hpp>
#include <geometry/surfaces/sphere.hpp>
#include <geometry/surfaces/xcylinder.hpp>
#include <geometry/surfaces/xplane.hpp>
#include <geometry/surfaces/ycylinder.hpp>
#include <geometry/surfaces/yplane.hpp>
#include <geometry/surfaces/zcylinder.hpp>
#include <geometry/surfaces/zplane.hpp>
#include <materials/legendre_distribution.hpp>
#include <materials/mg_angle_distribution.hpp>
#include <materials/mg_nuclide.hpp>
#include <utils/direction.hpp>
#include <utils/error.hpp>
#include <utils/rng.hpp>
#include <utils/settings.hpp>

#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

// utils/error.hpp declares these; src/error.cpp needs MPI, so the driver supplies them
void fatal_error(std::string mssg, std::source_location loc) {
  throw std::runtime_error(mssg + " (" + loc.file_name() + ":" + std::to_string(loc.line()) + ")");
}
void warning(std::string mssg, std::source_location) { std::fprintf(stderr, "reference warning: %s\n", mssg.c_str()); }

// src/settings.cpp needs HighFive (utils/output.hpp); the driver defines the few settings the multigroup sources read,
// with the declarations of include/utils/settings.hpp
namespace settings {
uint32_t ngroups = 0;
SimulationMode mode = SimulationMode::K_EIGENVALUE;
pcg32 rng;
std::vector<double> energy_bounds;
bool chi_matrix = false;
bool use_virtual_collisions = true;
}  // namespace settings
// src/nuclide.cpp is one line of data definitions behind the same heavy includes
std::map<uint32_t, std::shared_ptr<Nuclide>> nuclides;
std::unordered_set<uint32_t> zaids_with_urr;
uint32_t Nuclide::id_counter = 0;

namespace {
// surface type codes of include/abeille_b200.h / oracle/orc_geom.h
std::unique_ptr<Surface> make(int type, const double* p) {
  const BoundaryType b = BoundaryType::Normal;
  switch (type) {
    case 0: return std::make_unique<XPlane>(p[0], b, 1, "");
    case 1: return std::make_unique<YPlane>(p[0], b, 1, "");
    case 2: return std::make_unique<ZPlane>(p[0], b, 1, "");
    case 3: return std::make_unique<Plane>(p[0], p[1], p[2], p[3], b, 1, "");
    case 4: return std::make_unique<XCylinder>(p[0], p[1], p[2], b, 1, "");
    case 5: return std::make_unique<YCylinder>(p[0], p[1], p[2], b, 1, "");
    case 6: return std::make_unique<ZCylinder>(p[0], p[1], p[2], b, 1, "");
    case 7: return std::make_unique<Cylinder>(p[0], p[1], p[2], p[3], p[4], p[5], p[6], b, 1, "");
    default: return std::make_unique<Sphere>(p[0], p[1], p[2], p[3], b, 1, "");
  }
}
}  // namespace

extern "C" {

// n evaluations of one surface: sign, distance (on_surf as given) and norm
int ref_surface(int type, const double* params, int n, const double* r3, const double* u3, const int* on_surf, int* sign,
                double* dist, double* norm3) {
  try {
    auto s = make(type, params);
    for (int i = 0; i < n; i++) {
      const Position r(r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]);
      const Direction u(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]);
      sign[i] = s->sign(r, u);
      dist[i] = s->distance(r, u, on_surf[i] != 0);
      const Direction nn = s->norm(r);
      norm3[3 * i] = nn.x(); norm3[3 * i + 1] = nn.y(); norm3[3 * i + 2] = nn.z();
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_surface: %s\n", e.what());
    return 1;
  }
}

// Direction(x,y,z) (normalising constructor) and rotate_direction(u, mu, phi)
void ref_direction(int n, const double* xyz, double* out) {
  for (int i = 0; i < n; i++) {
    const Direction d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    out[3 * i] = d.x(); out[3 * i + 1] = d.y(); out[3 * i + 2] = d.z();
  }
}
void ref_rotate_direction(int n, const double* u3, const double* mu, const double* phi, double* out) {
  for (int i = 0; i < n; i++) {
    const Direction u(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]);
    const Direction d = rotate_direction(u, mu[i], phi[i]);
    out[3 * i] = d.x(); out[3 * i + 1] = d.y(); out[3 * i + 2] = d.z();
  }
}

// the stream of history `id`: seed(seed); advance(stride * id) (particle.hpp:188-193), then n x RNG::rand
void ref_rng_rand(uint64_t seed, uint64_t stride, uint64_t id, int n, double* out) {
  pcg32 rng;
  rng.seed(seed);
  rng.advance(stride * id);
  for (int i = 0; i < n; i++) out[i] = RNG::rand(rng);
}
double ref_rng_exponential(uint64_t seed, uint64_t stride, uint64_t id, double lambda) {
  pcg32 rng;
  rng.seed(seed);
  rng.advance(stride * id);
  return RNG::exponential(rng, lambda);
}
// ndraws x RNG::discrete over the same weights, then one RNG::rand (shows how many engine steps were consumed)
double ref_rng_discrete(uint64_t seed, uint64_t stride, uint64_t id, const double* w, int nw, int ndraws, int* out) {
  pcg32 rng;
  rng.seed(seed);
  rng.advance(stride * id);
  const std::vector<double> weights(w, w + nw);
  for (int i = 0; i < ndraws; i++) out[i] = RNG::discrete(rng, weights);
  return RNG::rand(rng);
}

// MGAngleDistribution(mu, pdf, cdf)::sample_mu, n draws from the stream of history `id`
void ref_sample_mu(const double* mu, const double* pdf, const double* cdf, int npts, uint64_t seed, uint64_t stride, uint64_t id,
                   int n, double* out) {
  const MGAngleDistribution d(std::vector<double>(mu, mu + npts), std::vector<double>(pdf, pdf + npts),
                              std::vector<double>(cdf, cdf + npts));
  pcg32 rng;
  rng.seed(seed);
  rng.advance(stride * id);
  for (int i = 0; i < n; i++) out[i] = d.sample_mu(rng);
}
// LegendreDistribution(a)::linearize(): number of points (<= cap) and the tables
int ref_legendre_linearize(const double* a, int na, int cap, double* mu, double* pdf, double* cdf) {
  try {
    LegendreDistribution L(std::vector<double>(a, a + na));
    const MGAngleDistribution d = L.linearize();
    const int n = (int)d.mu().size();
    if (n > cap) return -n;
    std::memcpy(mu, d.mu().data(), n * sizeof(double));
    std::memcpy(pdf, d.pdf().data(), n * sizeof(double));
    std::memcpy(cdf, d.cdf().data(), n * sizeof(double));
    return n;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_legendre_linearize: %s\n", e.what());
    return 0;
  }
}

// MGNuclide built through its vector constructor (src/mg_nuclide.cpp:36-70) from the arrays make_mg_nuclide would have
// read from the YAML material (:579-922; the Legendre moments go through LegendreDistribution::set_moment and linearize()
// as at :555-575 and :700-710).  micro: per group total, fission, absorption, elastic, nu_total, nu_delayed of
// get_micro_xs.  Per history h < nhist (stream seed / stride / h) and draw d < ndraw, from group (h + d) % G and the
// direction of the previous draw: sample_scatter -> scat[4] = E, u; sample_fission(Pdelayed = nu_delayed / nu_total)
// -> fis[6] = E, u, delayed, lambda.
int ref_mg_nuclide(int G, const double* ebounds, const double* Et, const double* Ea, const double* Ef, const double* nu_p,
                   const double* nu_d, const double* chi, const double* Es, int nleg, const double* leg, int ndg,
                   const double* Pd, const double* lam, uint64_t seed, uint64_t stride, int nhist, int ndraw, double* micro,
                   double* scat, double* fis) {
  try {
    const std::size_t g = (std::size_t)G;
    settings::ngroups = (uint32_t)G;
    settings::energy_bounds.assign(ebounds, ebounds + G + 1);
    auto vec = [g](const double* p) { return p ? std::vector<double>(p, p + g) : std::vector<double>(); };
    auto mat = [g](const double* p) {
      std::vector<std::vector<double>> m(g);
      for (std::size_t i = 0; i < g; i++) m[i].assign(p + i * g, p + (i + 1) * g);
      return m;
    };
    std::vector<std::vector<LegendreDistribution>> legendre(g, std::vector<LegendreDistribution>(g));
    for (int l = 1; l <= nleg; l++)
      for (std::size_t i = 0; i < g; i++)
        for (std::size_t o = 0; o < g; o++) legendre[i][o].set_moment((std::size_t)l, leg[((std::size_t)(l - 1) * g + i) * g + o]);
    std::vector<std::vector<MGAngleDistribution>> angles(g, std::vector<MGAngleDistribution>(g));
    for (std::size_t i = 0; i < g; i++)
      for (std::size_t o = 0; o < g; o++) angles[i][o] = legendre[i][o].linearize();
    const std::vector<std::vector<double>> yields(g, std::vector<double>(g, 1.));
    const MGNuclide nuc(std::vector<double>(g, 1.), vec(Et), vec(Ea), vec(Ef), vec(nu_p), vec(nu_d), mat(chi), mat(Es), yields,
                        angles, std::vector<double>(Pd, Pd + ndg), std::vector<double>(lam, lam + ndg));
    for (std::size_t i = 0; i < g; i++) {
      const MicroXSs xs = nuc.get_micro_xs(0.5 * (ebounds[i] + ebounds[i + 1]));
      if (xs.energy_index != i) return 2;
      double* m = micro + 6 * i;
      m[0] = xs.total; m[1] = xs.fission; m[2] = xs.absorption; m[3] = xs.elastic; m[4] = xs.nu_total; m[5] = xs.nu_delayed;
    }
    for (int h = 0; h < nhist; h++) {
      pcg32 rng;
      rng.seed(seed);
      rng.advance(stride * (uint64_t)h);
      Direction u(0., 0., 1.);
      for (int d = 0; d < ndraw; d++) {
        const std::size_t gi = (std::size_t)(h + d) % g;
        const double E = 0.5 * (ebounds[gi] + ebounds[gi + 1]);
        const MicroXSs xs = nuc.get_micro_xs(E);
        const ScatterInfo si = nuc.sample_scatter(E, u, xs, rng);
        double* s = scat + 4 * ((std::size_t)h * ndraw + d);
        s[0] = si.energy; s[1] = si.direction.x(); s[2] = si.direction.y(); s[3] = si.direction.z();
        u = si.direction;
        const double Pdelayed = xs.nu_total > 0. ? xs.nu_delayed / xs.nu_total : 0.;
        const FissionInfo fi = nuc.sample_fission(E, u, xs.energy_index, Pdelayed, rng);
        double* f = fis + 6 * ((std::size_t)h * ndraw + d);
        f[0] = fi.energy; f[1] = fi.direction.x(); f[2] = fi.direction.y(); f[3] = fi.direction.z();
        f[4] = fi.delayed ? 1. : 0.; f[5] = fi.precursor_decay_constant;
      }
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_mg_nuclide: %s\n", e.what());
    return 1;
  }
}

}  // extern "C"
