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
#include <utils/direction.hpp>
#include <utils/error.hpp>
#include <utils/rng.hpp>

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

}  // extern "C"
