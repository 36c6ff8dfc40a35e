/* oracle/ref_probe.cpp -- C entry points around the REFERENCE'S OWN code, for pinning the oracle (test infrastructure).
 *
 * The reference cannot be built as a whole offline (yaml-cpp, PapillonNDL, pcg-cpp, HighFive, NDArray ... are fetched by
 * its CMake), but a part of the hot path compiles from its own sources as they lie under /root/reference:
 *   surfaces     src/{x,y,z}plane.cpp, plane.cpp, {x,y,z}cylinder.cpp, cylinder.cpp, sphere.cpp, surface.cpp
 *   directions   include/utils/direction.hpp (Direction constructors, rotate_direction)
 *   RNG          include/utils/rng.hpp, src/rng.cpp over the pcg32 engine (pcg header vendored by pyarrow, ref_shim/pcg_random.hpp)
 *   angles       src/mg_angle_distribution.cpp, src/legendre_distribution.cpp (sample_mu, linearize)
 *   nuclide      src/mg_nuclide.cpp (vector constructor, get_micro_xs, sample_scatter, sample_fission)
 *   geometry     src/cell.cpp, universe.cpp, cell_universe.cpp, lattice.cpp, rect_lattice.cpp, hex_lattice.cpp, geometry.cpp,
 *                particle.cpp and the header-only Tracker (include/simulation/tracker.hpp)
 *   transport    src/surface_tracker.cpp, delta_tracker.cpp, carter_tracker.cpp, transporter.cpp, material_helper.cpp,
 *                tallies.cpp, mesh_tally.cpp, collision_mesh_tally.cpp, track_length_mesh_tally.cpp, source_mesh_tally.cpp,
 *                noise_maker.cpp and the two noise sources, mpi.cpp (no-MPI build), output.cpp, header.cpp, logo.cpp:
 *                ref_transport / ref_transport_noise run the reference's own Transporter::transport on a bank
 * oracle/Makefile (target `ref`) compiles those files in place (nothing is copied) together with this driver into
 * oracle/_ref/libabeille_ref.so.  Stand-in headers under oracle/ref_shim/ replace what the reference's CMake downloads
 * (yaml-cpp, PapillonNDL, Boost.Unordered, HighFive, NDArray) and two reference headers that would drag the plotter and the
 * YAML parser in (utils/parser.hpp, plotting/slice_plot.hpp); each says what it stands in for and why it cannot change a
 * result.  Restated in this file because their translation units cannot be compiled: the multigroup branch of
 * make_majorant_xs (src/majorant.cpp:131-176), the settings globals (src/settings.cpp) and the assembly of materials and
 * geometry from a deck (the YAML factory functions).  ref_power_iteration runs the reference's PowerIterator (power_iterator.cpp,
 * simulation.cpp, entropy.cpp, source.cpp + distributions, cancelator.cpp, approximate_mesh_cancelator.cpp).  ref_noise_run runs the reference's
 * Noise driver (noise.cpp).  Not covered: exact cancelators, fixed-source / branchless drivers (out of scope).
 * oracle/ref_pins.py runs seeded cases through this library and through the oracle; tests/test_reference_pins.py compares
 * them bit for bit and keeps the reference's outputs as tests/golden/ref_pins.npz for machines without the reference.
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
#include <geometry/cell_universe.hpp>
#include <geometry/geometry.hpp>
#include <geometry/hex_lattice.hpp>
#include <simulation/fixed_source.hpp>
#include <simulation/modified_fixed_source.hpp>
#include <geometry/rect_lattice.hpp>
#include <materials/legendre_distribution.hpp>
#include <materials/mg_angle_distribution.hpp>
#include <materials/mg_nuclide.hpp>
#include <plotting/plotter.hpp>
#include <simulation/carter_tracker.hpp>
#include <simulation/delta_tracker.hpp>
#include <simulation/implicit_leakage_delta_tracker.hpp>
#include <simulation/flat_vibration_noise_source.hpp>
#include <simulation/approximate_mesh_cancelator.hpp>
#include <simulation/basic_exact_mg_cancelator.hpp>
#include <simulation/exact_mg_cancelator.hpp>
#include <sobol/sobol.hpp>
#include <simulation/box.hpp>
#include <simulation/entropy.hpp>
#include <simulation/cone.hpp>
#include <simulation/isotropic.hpp>
#include <simulation/maxwellian.hpp>
#include <simulation/mono_directional.hpp>
#include <simulation/mono_energetic.hpp>
#include <simulation/watt.hpp>
#include <simulation/noise.hpp>
#include <simulation/noise_maker.hpp>
#include <simulation/point.hpp>
#include <simulation/power_iterator.hpp>
#include <simulation/branchless_power_iterator.hpp>
#include <simulation/source.hpp>
#include <simulation/square_oscillation_noise_source.hpp>
#include <simulation/surface_tracker.hpp>
#include <simulation/tracker.hpp>
#include <utils/direction.hpp>
#include <utils/error.hpp>
#include <utils/majorant.hpp>
#include <utils/mpi.hpp>
#include <utils/output.hpp>
#include <utils/rng.hpp>
#include <utils/settings.hpp>

#include <omp.h>

#include "../integration/gpu_transporter.hpp"
#include "../integration/flatten_problem.hpp"

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
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
int nparticles = 100000;
int ngenerations = 120;
int nignored = 20;
int nskip = 10;
uint32_t ngroups = 0;
int n_cancel_noise_gens = INT32_MAX;
bool plotting_mode = false;
double min_energy = 0.;
double max_energy = 100000.;
SimulationMode mode = SimulationMode::K_EIGENVALUE;
TrackingMode tracking = TrackingMode::SURFACE_TRACKING;
EnergyMode energy_mode = EnergyMode::MG;
uint64_t rng_seed = 19073486328125;
uint64_t rng_stride = 152917;
pcg32 rng;
double wgt_cutoff = 0.25;
double wgt_survival = 1.0;
double wgt_split = 2.0;
double w_noise = -1.;
double eta = 1.;
double keff = 1.;
bool use_urr_ptables = false;  // src/parser.cpp switches the tables off in multigroup mode
bool converged = false;
bool regional_cancellation = false;
bool regional_cancellation_noise = false;
bool inner_generations = true;
bool normalize_noise_source = true;
bool rng_stride_warnings = false;
// settings::group (src/settings.cpp:95-104, not among the compiled sources: the globals of that file are defined here): index of
// the first group whose closed interval holds E, 0 when there is none
std::size_t group(double E) {
  for (std::size_t g = 0; g + 1 < energy_bounds.size(); g++)
    if (!(E < energy_bounds[g]) && !(E > energy_bounds[g + 1])) return g;
  return 0;
}
bool branchless_splitting = false;
bool branchless_combing = false;
bool branchless_material = true;
std::vector<double> energy_bounds;
std::vector<double> sample_xs_ratio;
bool chi_matrix = false;
bool use_virtual_collisions = true;
Timer alpha_omega_timer;
double max_time = INF;
bool pair_distance_sqrd = false;
bool families = false;
bool empty_entropy_bins = false;
bool load_source_file = false;
void initialize_global_rng() {  // src/settings.cpp
  rng.seed(rng_seed);
  rng.set_stream(2);
}
}  // namespace settings
// src/parser.cpp and src/plotter.cpp (the whole simulation layer) are not built: the id -> index maps they own are defined
// here, and the geometry is assembled by ref_geometry_load below
std::map<uint32_t, size_t> surface_id_to_indx;
std::map<uint32_t, size_t> cell_id_to_indx;
std::map<uint32_t, size_t> universe_id_to_indx;
void find_universe(const YAML::Node&, uint32_t) { throw std::runtime_error("YAML factory called in oracle/_ref"); }
namespace plotter {
std::map<uint32_t, Pixel> cell_id_to_color;
std::map<uint32_t, Pixel> material_id_to_color;
}  // namespace plotter
// src/material.cpp needs the plotter and HighFive; the material table it owns is defined here
std::map<uint32_t, std::shared_ptr<Material>> materials;

// src/majorant.cpp cannot be compiled (its continuous-energy branch needs PapillonNDL's thermal-scattering classes).
// This is a restatement of its MULTIGROUP branch (src/majorant.cpp:131-176), statement for statement.
std::pair<std::vector<double>, std::vector<double>> make_majorant_xs() {
  std::vector<double> egrid;
  egrid.push_back(settings::energy_bounds[0]);
  if (egrid.front() == 0.) egrid.front() = 1.E-11;
  for (size_t i = 1; i < settings::energy_bounds.size() - 1; i++) {
    egrid.push_back(settings::energy_bounds[i]);
    egrid.push_back(settings::energy_bounds[i]);
  }
  egrid.push_back(settings::energy_bounds.back());
  std::vector<double> maj_xs(egrid.size(), 0.);
  for (const auto& material : materials) {
    MaterialHelper mat(material.second.get(), 1.);
    for (uint32_t g = 0; g < settings::ngroups; g++) {
      size_t i = g * 2;
      double Eg = 0.5 * (egrid[i] + egrid[i + 1]);
      double xs = mat.Et(Eg);
      if (xs > maj_xs[i]) {
        maj_xs[i] = xs;
        maj_xs[i + 1] = xs;
      }
    }
  }
  return {egrid, maj_xs};
}

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

// ---- geometry: the reference's Surface / Cell / CellUniverse / RectLattice objects and its Tracker -----------------------
namespace {
struct CellDef { uint32_t id; bool fill_universe; uint32_t fill; std::string region; };
struct UniDef {
  uint32_t id; bool lattice; std::vector<uint32_t> cells;
  uint32_t shape[3]; double pitch[3], origin[3]; int32_t outer; std::vector<int32_t> tiles;
  bool hex = false; int top = 0;  // hexagonal lattice: shape = {nrings, nz}, pitch = {pitch, pitch_z}
};
struct GeoDeck {
  std::vector<CellDef> cells;
  std::vector<UniDef> unis;
  std::map<uint32_t, std::shared_ptr<Material>> materials;  // empty Material objects: the cursor only hands the pointer back
  std::map<const Material*, int> material_index;
} deck;

// what make_universe / find_universe / make_cell_universe / make_rect_lattice do with a YAML node (src/parser.cpp:295-339,
// src/cell_universe.cpp:292-341, src/rect_lattice.cpp:312-424), from the parsed records: same recursion, same order of
// geometry::universes
void need_universe(uint32_t id);
void build_universe(const UniDef& d) {
  if (universe_id_to_indx.count(d.id)) return;
  if (!d.lattice) {
    std::vector<uint32_t> cells;
    for (uint32_t cid : d.cells) {
      if (!cell_id_to_indx.count(cid)) throw std::runtime_error("Referenced cell id could not be found.");
      cells.push_back(static_cast<uint32_t>(cell_id_to_indx[cid]));
    }
    universe_id_to_indx[d.id] = geometry::universes.size();
    geometry::universes.push_back(std::make_shared<CellUniverse>(cells, d.id, ""));
    return;
  }
  std::vector<int32_t> uni_indicies;
  for (int32_t u_id : d.tiles) {
    if (u_id == -1) { uni_indicies.push_back(u_id); continue; }
    need_universe(static_cast<uint32_t>(u_id));
    uni_indicies.push_back(static_cast<int32_t>(universe_id_to_indx[static_cast<uint32_t>(u_id)]));
  }
  std::shared_ptr<Lattice> lat;
  if (d.hex)  // make_hex_lattice, src/hex_lattice.cpp:546-552
    lat = std::make_shared<HexLattice>(d.shape[0], d.shape[1], d.pitch[0], d.pitch[1], d.origin[0], d.origin[1], d.origin[2],
                                       d.top ? HexLattice::Top::Flat : HexLattice::Top::Pointy, d.id, "");
  else
    lat = std::make_shared<RectLattice>(d.shape[0], d.shape[1], d.shape[2], d.pitch[0], d.pitch[1], d.pitch[2], d.origin[0],
                                        d.origin[1], d.origin[2], d.id, "");
  lat->set_elements(uni_indicies);
  if (d.outer != -1) {
    need_universe(static_cast<uint32_t>(d.outer));
    lat->set_outisde_universe(static_cast<int32_t>(universe_id_to_indx[static_cast<uint32_t>(d.outer)]));
  }
  universe_id_to_indx[d.id] = geometry::universes.size();
  geometry::universes.push_back(lat);
}
void need_universe(uint32_t id) {
  if (universe_id_to_indx.count(id)) return;
  for (const auto& d : deck.unis)
    if (d.id == id) { build_universe(d); return; }
  throw std::runtime_error("Could not find universe.");
}

// the tokeniser of make_cell (src/cell.cpp:311-357), then the reference's own infix_to_rpn and Cell constructors
void build_cell(const CellDef& c) {
  std::vector<int32_t> region;
  std::string temp;
  auto flush = [&]() {
    if (temp.empty()) return;
    const int32_t signed_id = std::stoi(temp);
    int32_t indx = static_cast<int32_t>(surface_id_to_indx.at(static_cast<uint32_t>(std::abs(signed_id)))) + 1;
    if (signed_id < 0) indx *= -1;
    region.push_back(indx);
    temp.clear();
  };
  for (char ch : c.region) {
    if (ch == '&' || ch == '(' || ch == ')' || ch == 'U' || ch == '~') {
      flush();
      region.push_back(ch == '&' ? OP::INTR : ch == '(' ? OP::L_PAR : ch == ')' ? OP::R_PAR : ch == 'U' ? OP::UNIN : OP::COMP);
    } else if (ch == '+' || ch == '-' || (ch >= '0' && ch <= '9')) {
      temp += ch;
    } else if (ch != ' ') {
      throw std::runtime_error("Invalid character in cell region definition.");
    }
  }
  flush();
  region = infix_to_rpn(region);
  std::shared_ptr<Cell> cell;
  if (!c.fill_universe) {
    cell = std::make_shared<Cell>(region, deck.materials.at(c.fill), c.id, "");
  } else {
    need_universe(c.fill);
    cell = std::make_shared<Cell>(region, geometry::universes[universe_id_to_indx[c.fill]], c.id, "");
  }
  cell_id_to_indx[c.id] = geometry::cells.size();
  geometry::cells.push_back(cell);
}

std::unique_ptr<Surface> make_bc(int type, const double* p, BoundaryType b, uint32_t id) {
  switch (type) {
    case 0: return std::make_unique<XPlane>(p[0], b, id, "");
    case 1: return std::make_unique<YPlane>(p[0], b, id, "");
    case 2: return std::make_unique<ZPlane>(p[0], b, id, "");
    case 3: return std::make_unique<Plane>(p[0], p[1], p[2], p[3], b, id, "");
    case 4: return std::make_unique<XCylinder>(p[0], p[1], p[2], b, id, "");
    case 5: return std::make_unique<YCylinder>(p[0], p[1], p[2], b, id, "");
    case 6: return std::make_unique<ZCylinder>(p[0], p[1], p[2], b, id, "");
    case 7: return std::make_unique<Cylinder>(p[0], p[1], p[2], p[3], p[4], p[5], p[6], b, id, "");
    default: return std::make_unique<Sphere>(p[0], p[1], p[2], p[3], b, id, "");
  }
}

inline void put(double* o, Tracker& t) {
  o[0] = t.is_lost() ? -1. : static_cast<double>(t.cell()->id());
  o[1] = t.is_lost() || !t.material() ? -1. : static_cast<double>(deck.material_index.at(t.material()));
  o[2] = 0.;  // (cell instances are not on the multigroup hot path; the oracle does not restate them)
}
}  // namespace

extern "C" {

// The geometry section of the oracle's flat deck text (oracle/deck.py: "nmat"-independent lines nsurf ... root), i.e. the
// surfaces, cells, universes and root-universe entries of the YAML deck, assembled in the order make_geometry
// (src/parser.cpp:178-246) assembles them.  material ids are listed so that cells can be handed (empty) Material objects.
int ref_geometry_load(const char* text, int nmat, const int* material_ids) {
  try {
    geometry::surfaces.clear(); geometry::cells.clear(); geometry::universes.clear(); geometry::root_universe = nullptr;
    surface_id_to_indx.clear(); cell_id_to_indx.clear(); universe_id_to_indx.clear();
    deck = GeoDeck();
    for (int m = 0; m < nmat; m++) {
      auto mat = std::make_shared<Material>();
      deck.materials[static_cast<uint32_t>(material_ids[m])] = mat;
      deck.material_index[mat.get()] = m;
    }
    if (nmat < 0) {  // called by ref_problem_load: the cells get the real materials
      int m = 0;
      for (const auto& kv : materials) { deck.materials[kv.first] = kv.second; deck.material_index[kv.second.get()] = m++; }
    }
    static const std::map<std::string, int> types = {{"xplane", 0}, {"yplane", 1}, {"zplane", 2}, {"plane", 3}, {"xcylinder", 4},
                                                     {"ycylinder", 5}, {"zcylinder", 6}, {"cylinder", 7}, {"sphere", 8}};
    std::istringstream in(text);
    std::string key;
    uint32_t root = 0;
    while (in >> key) {
      if (key == "nsurf" || key == "ncell" || key == "nuni") { int n; in >> n; continue; }
      if (key == "surf") {
        uint32_t id; std::string type, bc; int np; double p[7] = {0, 0, 0, 0, 0, 0, 0};
        in >> id >> type >> bc >> np;
        for (int k = 0; k < np; k++) in >> p[k];
        const BoundaryType b = bc == "vacuum" ? BoundaryType::Vacuum : bc == "reflective" ? BoundaryType::Reflective : BoundaryType::Normal;
        surface_id_to_indx[id] = geometry::surfaces.size();
        geometry::surfaces.push_back(make_bc(types.at(type), p, b, id));
      } else if (key == "cell") {
        CellDef c; std::string kind;
        in >> c.id >> kind >> c.fill >> c.region;
        c.fill_universe = kind == "u";
        deck.cells.push_back(c);
      } else if (key == "uni") {
        UniDef u; std::string kind;
        in >> u.id >> kind;
        u.lattice = kind == "rect" || kind == "hex";
        u.hex = kind == "hex";
        if (!u.lattice) {
          size_t n; in >> n; u.cells.resize(n);
          for (auto& c : u.cells) in >> c;
        } else if (u.hex) {
          size_t n;
          in >> u.shape[0] >> u.shape[1] >> u.pitch[0] >> u.pitch[1] >> u.origin[0] >> u.origin[1] >> u.origin[2] >> u.top >> u.outer >> n;
          u.tiles.resize(n);
          for (auto& t : u.tiles) in >> t;
        } else {
          size_t n;
          in >> u.shape[0] >> u.shape[1] >> u.shape[2] >> u.pitch[0] >> u.pitch[1] >> u.pitch[2] >> u.origin[0] >> u.origin[1] >>
              u.origin[2] >> u.outer >> n;
          u.tiles.resize(n);
          for (auto& t : u.tiles) in >> t;
        }
        deck.unis.push_back(u);
      } else if (key == "root") {
        in >> root;
      } else {
        throw std::runtime_error("unknown key " + key);
      }
    }
    for (const auto& c : deck.cells) build_cell(c);
    for (const auto& u : deck.unis) build_universe(u);
    for (auto& uni : geometry::universes) uni->make_offset_map();
    geometry::root_universe = geometry::universes[universe_id_to_indx.at(root)];
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_geometry_load: %s\n", e.what());
    return 1;
  }
}

// Surface-tracking walk of n rays through the loaded geometry, the geometry calls of SurfaceTracker::transport
// (src/surface_tracker.cpp:60-140) without physics: Tracker(r, u); then up to nsteps times get_nearest_boundary() and
// Vacuum -> stop / Reflective -> Tracker::do_reflection / Normal -> cross_surface + get_current.
// out[ray][step][8] = cell id, material index, 0 (after the step; step 0 = at birth), boundary distance,
// surface index, boundary type, token, 1 (row written).  A lost ray ends its walk.
int ref_geometry_walk_surface(int n, const double* r3, const double* u3, int nsteps, double* out) {
  try {
    for (int i = 0; i < n; i++) {
      double* o = out + static_cast<size_t>(i) * (nsteps + 1) * 8;
      Particle p(Position(r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]), Direction(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]), 1., 1.);
      Tracker trkr(p.r(), p.u());
      put(o, trkr); o[7] = 1.;
      for (int s = 1; s <= nsteps && !trkr.is_lost(); s++) {
        o += 8;
        const Boundary b = trkr.get_nearest_boundary();
        o[3] = b.distance; o[4] = b.surface_index; o[5] = static_cast<double>(static_cast<int>(b.boundary_type)); o[6] = b.token;
        o[7] = 1.;
        if (b.boundary_type == BoundaryType::Vacuum) { o[0] = o[1] = -2.; break; }
        if (b.boundary_type == BoundaryType::Reflective) {
          trkr.do_reflection(p, b);
        } else {
          trkr.cross_surface(b);
          trkr.get_current();
          p.move(b.distance);
        }
        put(o, trkr);
      }
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_geometry_walk_surface: %s\n", e.what());
    return 1;
  }
}

// Delta-tracking walk, the geometry calls of DeltaTracker::transport (src/delta_tracker.cpp:113-160) without physics:
// Tracker(r, u); per step move(d[ray][step]) + get_current(); when that loses the particle, set_r(back) + get_current() +
// get_boundary_condition(), then Vacuum -> stop / Reflective -> Tracker::do_reflection.  Every second completed flight is
// followed by a direction change to unew[ray][step] (set_u), as after a collision.  Row layout as in the surface walk; the
// boundary columns are INF, -1, Normal, 0 for a flight that stayed inside.
int ref_geometry_walk_delta(int n, const double* r3, const double* u3, int nsteps, const double* d, const double* unew3,
                            double* out) {
  try {
    for (int i = 0; i < n; i++) {
      double* o = out + static_cast<size_t>(i) * (nsteps + 1) * 8;
      Particle p(Position(r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]), Direction(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]), 1., 1.);
      Tracker trkr(p.r(), p.u());
      put(o, trkr); o[7] = 1.;
      for (int s = 1; s <= nsteps && !trkr.is_lost(); s++) {
        o += 8;
        const size_t k = static_cast<size_t>(i) * nsteps + (s - 1);
        Boundary b(INF, -1, BoundaryType::Normal);
        bool crossed_boundary = false;
        trkr.move(d[k]);
        trkr.get_current();
        if (trkr.is_lost()) {
          trkr.set_r(p.r());
          trkr.get_current();
          b = trkr.get_boundary_condition();
          crossed_boundary = true;
        }
        o[3] = b.distance; o[4] = b.surface_index; o[5] = static_cast<double>(static_cast<int>(b.boundary_type)); o[6] = b.token;
        o[7] = 1.;
        if (crossed_boundary) {
          if (b.boundary_type == BoundaryType::Vacuum) { o[0] = o[1] = -2.; break; }
          if (b.boundary_type != BoundaryType::Reflective) throw std::runtime_error("Help me, how did I get here ?");
          trkr.do_reflection(p, b);
        } else {
          p.move(d[k]);
          if (s % 2 == 0) {
            p.set_direction(Direction(unew3[3 * k], unew3[3 * k + 1], unew3[3 * k + 2]));
            trkr.set_u(p.u());
          }
        }
        put(o, trkr);
      }
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_geometry_walk_delta: %s\n", e.what());
    return 1;
  }
}

}  // extern "C"

// ---- the whole hot path: the reference's SurfaceTracker / DeltaTracker / CarterTracker::transport ------------------------
namespace {
struct Tok {
  std::istringstream in;
  explicit Tok(const char* t) : in(t) {}
  std::string next() { std::string s; if (!(in >> s)) throw std::runtime_error("deck: unexpected end"); return s; }
  void expect(const char* k) { const std::string s = next(); if (s != k) throw std::runtime_error("deck: expected " + std::string(k) + ", got " + s); }
  double d() { return std::stod(next()); }
  long long ll() { return std::stoll(next()); }
  std::vector<double> dv(size_t n) { std::vector<double> v(n); for (auto& x : v) x = d(); return v; }
};
// A bank direction bit for bit: in the reference a particle's direction is a copy of the Direction object the source or the
// fission sampler made (src/power_iterator.cpp: Particle(p.r, p.u, ...)), never normalised a second time, and Direction has
// no constructor that skips the normalisation.
static_assert(sizeof(Direction) == 3 * sizeof(double), "Direction is three doubles");
Direction raw_direction(const double* u) {
  Direction d;
  std::memcpy(static_cast<void*>(&d), u, sizeof(Direction));
  return d;
}
std::shared_ptr<Tallies> g_tallies;
std::shared_ptr<Transporter> g_transporter;
std::shared_ptr<GPUTransporter> g_gpu_transporter;
// the generation array of a mesh tally is a protected member (include/simulation/mesh_tally.hpp:79); the probe's tallies are
// created as these derived types so that it can be read back.  Nothing is overridden.
struct CollisionTallyProbe : CollisionMeshTally {
  using CollisionMeshTally::CollisionMeshTally;
  const NDArray<double>& gen() const { return tally_gen; }
  static void forget_names() { used_tally_names.clear(); }
};
struct TrackLengthTallyProbe : TrackLengthMeshTally {
  using TrackLengthMeshTally::TrackLengthMeshTally;
  const NDArray<double>& gen() const { return tally_gen; }
};
std::vector<const NDArray<double>*> g_tally_gen;
std::vector<const MeshTally*> g_mesh_tallies;
std::unique_ptr<NoiseMaker> g_noise_maker;
int g_threads = 1;
double g_last_transport_seconds = 0.;
double g_last_simulation_seconds = 0.;
}  // namespace

extern "C" {

// Loads the oracle's flat deck text (oracle/deck.py) up to the root universe: settings, materials (MGNuclide through its
// vector constructor, with the fissile gate and the chi-row replication of make_mg_nuclide, src/mg_nuclide.cpp:716-852;
// Material() + add_component({1., nuclide}) as Material(nuc, id) does, src/material.cpp:51-63), geometry (as
// ref_geometry_load), then Tallies(nparticles) and the tracker the deck names.
int ref_problem_load(const char* text) {
  try {
    Output::set_output_filename("/nonexistent-dir/oracle_ref.h5");  // never opened: HighFive is a stand-in
    Tok tk(text);
    tk.expect("ORCDECK"); tk.ll();
    tk.expect("mode");
    {
      const std::string mode = tk.next();
      settings::mode = mode == "noise" ? settings::SimulationMode::NOISE
                       : mode == "mfs" ? settings::SimulationMode::MODIFIED_FIXED_SOURCE
                       : mode == "fs" ? settings::SimulationMode::FIXED_SOURCE
                       : mode == "branchless" ? settings::SimulationMode::BRANCHLESS_K_EIGENVALUE
                                       : settings::SimulationMode::K_EIGENVALUE;
      if (settings::mode == settings::SimulationMode::BRANCHLESS_K_EIGENVALUE) {  // parser.cpp:367-409
        settings::branchless_material = tk.ll() != 0;
        settings::branchless_splitting = tk.ll() != 0;
        settings::branchless_combing = tk.ll() != 0;
      }
    }
    tk.expect("tracking");
    const std::string trk = tk.next();
    settings::tracking = trk == "delta" ? settings::TrackingMode::DELTA_TRACKING
                         : trk == "carter" ? settings::TrackingMode::CARTER_TRACKING
                         : trk == "implicit" ? settings::TrackingMode::IMPLICIT_LEAKAGE_DELTA_TRACKING
                                             : settings::TrackingMode::SURFACE_TRACKING;
    tk.expect("ngroups");
    const size_t G = (size_t)tk.ll();
    settings::ngroups = (uint32_t)G;
    tk.expect("ebounds");
    settings::energy_bounds = tk.dv(G + 1);
    tk.expect("nparticles"); settings::nparticles = (int)tk.ll();
    tk.expect("ngenerations"); settings::ngenerations = (int)tk.ll();
    tk.expect("nignored"); settings::nignored = (int)tk.ll();
    tk.expect("nskip"); settings::nskip = (int)tk.ll();
    tk.expect("wgt"); settings::wgt_cutoff = tk.d(); settings::wgt_survival = tk.d(); settings::wgt_split = tk.d();
    tk.expect("seed"); settings::rng_seed = (uint64_t)std::stoull(tk.next());
    tk.expect("stride"); settings::rng_stride = (uint64_t)std::stoull(tk.next());
    settings::initialize_global_rng();  // the stream cancellation draws from (called once the seed is known)
    tk.expect("ratios");
    settings::sample_xs_ratio = tk.dv((size_t)tk.ll());
    tk.expect("cancel"); tk.ll(); tk.ll(); tk.ll();
    tk.expect("noise");
    settings::w_noise = tk.d(); settings::keff = tk.d();
    settings::inner_generations = tk.ll() != 0; settings::normalize_noise_source = tk.ll() != 0;
    settings::min_energy = 0.; settings::max_energy = 100000.;
    settings::converged = false;
    settings::chi_matrix = false; settings::use_virtual_collisions = true;

    materials.clear();
    nuclides.clear();
    std::vector<int> material_ids;
    tk.expect("nmat");
    const size_t M = (size_t)tk.ll();
    for (size_t m = 0; m < M; m++) {
      tk.expect("mat");
      const uint32_t id = (uint32_t)tk.ll();
      tk.expect("total"); auto Et = tk.dv(G);
      tk.expect("absorption"); auto Ea = tk.dv(G);
      tk.expect("fission"); auto Ef = tk.dv(G);
      tk.expect("nu_p"); auto nu_p = tk.dv(G);
      tk.expect("nu_d"); auto nu_d = tk.dv(G);
      tk.expect("speeds"); auto speeds = tk.dv(G);
      tk.expect("chi");
      const size_t nrows = (size_t)tk.ll();
      std::vector<std::vector<double>> chi(G, std::vector<double>(G, 0.)), rows(nrows);
      for (auto& r : rows) r = tk.dv(G);
      tk.expect("scatter");
      std::vector<std::vector<double>> Es(G);
      for (auto& r : Es) r = tk.dv(G);
      tk.expect("nleg");
      const size_t L = (size_t)tk.ll();
      std::vector<std::vector<LegendreDistribution>> legendre(G, std::vector<LegendreDistribution>(G));
      for (size_t l = 1; l <= L; l++) {
        tk.expect("P");
        const size_t order = (size_t)tk.ll();
        for (size_t i = 0; i < G; i++)
          for (size_t o = 0; o < G; o++) legendre[i][o].set_moment(order, tk.d());
      }
      std::vector<std::vector<MGAngleDistribution>> angles(G, std::vector<MGAngleDistribution>(G));
      for (size_t i = 0; i < G; i++)
        for (size_t o = 0; o < G; o++) angles[i][o] = legendre[i][o].linearize();
      tk.expect("ndg");
      const size_t ND = (size_t)tk.ll();
      auto Pd = tk.dv(ND);
      auto lam = tk.dv(ND);
      bool fissile = false;
      for (double v : Ef) if (v > 0.) fissile = true;
      if (!fissile) {
        nu_p.assign(G, 0.); nu_d.assign(G, 0.);
      } else if (nrows == G) {
        chi = rows;
        settings::chi_matrix = true;
      } else {
        for (size_t i = 0; i < G; i++) chi[i] = rows[0];
      }
      const std::vector<std::vector<double>> yields(G, std::vector<double>(G, 1.));
      auto nuc = std::make_shared<MGNuclide>(speeds, Et, Ea, Ef, nu_p, nu_d, chi, Es, yields, angles, Pd, lam);
      nuclides[nuc->id()] = nuc;  // as make_mg_nuclide does (src/mg_nuclide.cpp:919); the vibration sources look nuclides up
      auto mat = std::make_shared<Material>();
      mat->add_component({1., nuc});
      materials[id] = mat;
      material_ids.push_back((int)id);
    }
    for (const auto& mat : materials) {  // src/parser.cpp:157-168
      const double emax = mat.second->max_energy(), emin = mat.second->min_energy();
      if (emin > settings::min_energy) settings::min_energy = emin;
      if (emax < settings::max_energy) settings::max_energy = emax;
    }
    // geometry: the rest of the text up to and including "root"
    std::string rest, line;
    std::getline(tk.in, line);
    while (std::getline(tk.in, line)) {
      rest += line + "\n";
      if (line.rfind("root ", 0) == 0) break;
    }
    if (ref_geometry_load(rest.c_str(), -1, nullptr) != 0) return 1;

    g_tallies = std::make_shared<Tallies>(static_cast<double>(settings::nparticles));
    g_tallies->set_keff(settings::keff);  // src/parser.cpp:885
    // mesh tallies: "tally name estimator quantity noise_like nx ny nz low[3] hi[3] nE ebounds[nE]" lines of the deck text,
    // through the plain constructors (collision_mesh_tally.hpp:34-37, track_length_mesh_tally.hpp:34-37); the quantity code
    // is the position in MeshTally::Quantity.  Source-estimator tallies are not scored inside transport() and are skipped.
    g_tally_gen.clear();
    g_mesh_tallies.clear();
    CollisionTallyProbe::forget_names();
    g_noise_maker = std::make_unique<NoiseMaker>();
    while (std::getline(tk.in, line)) {
      if (line.rfind("sqosc ", 0) == 0 || line.rfind("flatvib ", 0) == 0) {
        // noise sources through their plain constructors (square_oscillation_noise_source.hpp:35-37,
        // flat_vibration_noise_source.hpp:40-43); materials of a vibration are given as positions in the deck's material list
        std::istringstream ns(line);
        std::string kind;
        double lo[3], hi[3], w0;
        ns >> kind >> lo[0] >> lo[1] >> lo[2] >> hi[0] >> hi[1] >> hi[2] >> w0;
        if (kind == "sqosc") {
          double et, ef, es;
          ns >> et >> ef >> es;
          g_noise_maker->add_noise_source(std::shared_ptr<OscillationNoiseSource>(std::make_shared<SquareOscillationNoiseSource>(
              Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]), et, ef, es, w0)));
        } else {
          int basis, ip, in;
          ns >> basis >> ip >> in;
          g_noise_maker->add_noise_source(std::shared_ptr<VibrationNoiseSource>(std::make_shared<FlatVibrationNoiseSource>(
              Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]), static_cast<FlatVibrationNoiseSource::Basis>(basis),
              materials.at((uint32_t)material_ids[(size_t)ip]), materials.at((uint32_t)material_ids[(size_t)in]), w0)));
        }
        continue;
      }
      if (line.rfind("tally ", 0) != 0) continue;
      std::istringstream ls(line);
      std::string key, name;
      int est, qty, noise_like;
      uint64_t nx, ny, nz;
      double lo[3], hi[3];
      size_t ne;
      ls >> key >> name >> est >> qty >> noise_like >> nx >> ny >> nz >> lo[0] >> lo[1] >> lo[2] >> hi[0] >> hi[1] >> hi[2] >> ne;
      std::vector<double> eb(ne);
      for (auto& e : eb) ls >> e;
      const auto q = static_cast<MeshTally::Quantity>(qty);
      if (est == 0) {
        auto t = std::make_shared<CollisionTallyProbe>(Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]), nx, ny, nz, eb, q, name);
        g_tallies->add_collision_mesh_tally(t);
        g_tally_gen.push_back(&t->gen());
        g_mesh_tallies.push_back(t.get());
      } else if (est == 1) {
        auto t = std::make_shared<TrackLengthTallyProbe>(Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]), nx, ny, nz, eb, q, name);
        g_tallies->add_track_length_mesh_tally(t);
        g_tally_gen.push_back(&t->gen());
        g_mesh_tallies.push_back(t.get());
      } else {  // source estimator (source_mesh_tally.hpp:36-40): scored by the drivers, not inside transport()
        auto t = std::make_shared<SourceMeshTally>(Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]), nx, ny, nz, eb,
                                                   static_cast<SourceMeshTally::Quantity>(qty - 8), name);
        if (t->noise_like_score()) g_tallies->add_noise_source_mesh_tally(t);
        else g_tallies->add_source_mesh_tally(t);
        g_tally_gen.push_back(nullptr);
        g_mesh_tallies.push_back(t.get());
      }
    }
    switch (settings::tracking) {
      case settings::TrackingMode::DELTA_TRACKING: g_transporter = std::make_shared<DeltaTracker>(g_tallies); break;
      case settings::TrackingMode::CARTER_TRACKING: g_transporter = std::make_shared<CarterTracker>(g_tallies); break;
      case settings::TrackingMode::IMPLICIT_LEAKAGE_DELTA_TRACKING:
        g_transporter = std::make_shared<ImplicitLeakageDeltaTracker>(g_tallies);
        break;
      default: g_transporter = std::make_shared<SurfaceTracker>(g_tallies); break;
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_problem_load: %s\n", e.what());
    return 1;
  }
}

// One Transporter::transport(bank) call of the reference, as PowerIterator::run makes it (src/power_iterator.cpp:371):
// particles Particle(r, u, E, wgt, history id) with set_family_id and initialize_rng(seed, stride) (:196-200), k_col of the previous
// generation in the tallies.  Out: the fission bank in the order transport() returns it (9 doubles + parent history id,
// parent daughter id, family id per site) and the generation values Tallies::calc_gen_values makes of the scores (k_col,
// k_abs, k_trk, k_tot, leakage, migration area).  ref_set_threads(1) (the default) accumulates the score sums in bank order.
// what the exact cancelators read from the fission bank of the last ref_transport (particle.hpp:52-57), [n][10]:
// parents_previous_position, Esmp_parent, parents_previous_direction, parents_previous_previous_energy, parents_previous_energy,
// parents_previous_was_virtual
static std::vector<BankedParticle> g_last_fission;
uint64_t ref_last_parents(double* out10n, uint64_t n) {
  const uint64_t m = std::min<uint64_t>(n, g_last_fission.size());
  for (uint64_t i = 0; i < m; i++) {
    const BankedParticle& f = g_last_fission[i];
    double* o = out10n + 10 * i;
    o[0] = f.parents_previous_position.x(); o[1] = f.parents_previous_position.y(); o[2] = f.parents_previous_position.z();
    o[3] = f.Esmp_parent;
    o[4] = f.parents_previous_direction.x(); o[5] = f.parents_previous_direction.y(); o[6] = f.parents_previous_direction.z();
    o[7] = f.parents_previous_previous_energy; o[8] = f.parents_previous_energy; o[9] = f.parents_previous_was_virtual ? 1. : 0.;
  }
  return m;
}
int ref_transport(uint64_t n, const double* r3, const double* u3, const double* E, const double* wgt, const uint64_t* hid,
                  const uint64_t* family, double k_col, int converged, uint64_t cap, double* out9, uint64_t* out_ids3, uint64_t* n_out, double* scores6) {
  try {
    omp_set_num_threads(g_threads);
    std::vector<Particle> bank;
    bank.reserve(n);
    for (uint64_t i = 0; i < n; i++) {
      bank.emplace_back(Position(r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]), raw_direction(u3 + 3 * i), E[i], wgt[i], hid[i]);
      bank.back().set_family_id(family[i]);
      bank.back().initialize_rng(settings::rng_seed, settings::rng_stride);
    }
    g_tallies->clear_generation();
    g_tallies->set_kcol(k_col);
    settings::converged = converged != 0;  // mesh tallies score only then (tallies.hpp:49-63)
    const double t0 = omp_get_wtime();
    const std::vector<BankedParticle> fis = g_transporter->transport(bank, false, nullptr, nullptr);
    g_last_transport_seconds = omp_get_wtime() - t0;
    g_tallies->calc_gen_values();  // score sums / total weight (src/tallies.cpp:159-181)
    scores6[0] = g_tallies->kcol(); scores6[1] = g_tallies->kabs(); scores6[2] = g_tallies->ktrk();
    scores6[3] = g_tallies->ktot(); scores6[4] = g_tallies->leakage(); scores6[5] = g_tallies->mig_area();
    *n_out = fis.size();
    g_last_fission = fis;
    for (uint64_t i = 0; i < fis.size() && i < cap; i++) {
      double* o = out9 + 9 * i;
      o[0] = fis[i].r.x(); o[1] = fis[i].r.y(); o[2] = fis[i].r.z();
      o[3] = fis[i].u.x(); o[4] = fis[i].u.y(); o[5] = fis[i].u.z();
      o[6] = fis[i].E; o[7] = fis[i].wgt; o[8] = fis[i].wgt2;
      out_ids3[3 * i] = fis[i].parent_history_id; out_ids3[3 * i + 1] = fis[i].parent_daughter_id; out_ids3[3 * i + 2] = fis[i].family_id;
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_transport: %s\n", e.what());
    return 1;
  }
}

// Transporter::transport(bank, noise, &noise_bank, &noise_maker) as PowerIterator::run (noise == 0, the noise source is
// sampled at the collisions: src/power_iterator.cpp:371, src/transporter.cpp:74-78) and Noise::run (noise != 0: complex
// weights wgt + i wgt2, src/noise.cpp:312-314) make it.  keff is what Tallies::keff() returns to the samplers.  Out: fission
// bank and noise-source bank, 9 doubles + 3 ids per particle each.
int ref_transport_noise(uint64_t n, const double* r3, const double* u3, const double* E, const double* wgt, const double* wgt2,
                        const uint64_t* hid, const uint64_t* family, double k_col, double keff, int noise, int sample_noise,
                        uint64_t cap, double* out9, uint64_t* out_ids3, uint64_t* n_out, double* noise9, uint64_t* noise_ids3,
                        uint64_t* n_noise, double* scores6) {
  try {
    omp_set_num_threads(g_threads);
    std::vector<Particle> bank;
    bank.reserve(n);
    for (uint64_t i = 0; i < n; i++) {
      bank.emplace_back(Position(r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]), raw_direction(u3 + 3 * i), E[i], wgt[i], wgt2[i], hid[i]);
      bank.back().set_family_id(family[i]);
      bank.back().initialize_rng(settings::rng_seed, settings::rng_stride);
    }
    g_tallies->clear_generation();
    g_tallies->set_kcol(k_col);
    g_tallies->set_keff(keff);
    settings::converged = false;
    std::vector<BankedParticle> nb;
    const std::vector<BankedParticle> fis =
        g_transporter->transport(bank, noise != 0, sample_noise ? &nb : nullptr, sample_noise ? g_noise_maker.get() : nullptr);
    g_tallies->calc_gen_values();
    scores6[0] = g_tallies->kcol(); scores6[1] = g_tallies->kabs(); scores6[2] = g_tallies->ktrk();
    scores6[3] = g_tallies->ktot(); scores6[4] = g_tallies->leakage(); scores6[5] = g_tallies->mig_area();
    auto put_bank = [cap](const std::vector<BankedParticle>& v, double* o9, uint64_t* oi) {
      for (uint64_t i = 0; i < v.size() && i < cap; i++) {
        double* o = o9 + 9 * i;
        o[0] = v[i].r.x(); o[1] = v[i].r.y(); o[2] = v[i].r.z();
        o[3] = v[i].u.x(); o[4] = v[i].u.y(); o[5] = v[i].u.z();
        o[6] = v[i].E; o[7] = v[i].wgt; o[8] = v[i].wgt2;
        oi[3 * i] = v[i].parent_history_id; oi[3 * i + 1] = v[i].parent_daughter_id; oi[3 * i + 2] = v[i].family_id;
      }
    };
    *n_out = fis.size();
    put_bank(fis, out9, out_ids3);
    *n_noise = nb.size();
    put_bank(nb, noise9, noise_ids3);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_transport_noise: %s\n", e.what());
    return 1;
  }
}

// average (which = 1) / error of the mean (which = 2: write_tally has converted the variance accumulator in place) of mesh
// tally t after ref_power_iteration (src/mesh_tally.cpp:121-150,195-197; private members, see ref_power_iteration)
void ref_tally_get_stat(int t, int which, double* out) {
  const NDArray<double>& a = which == 1 ? g_mesh_tallies[(size_t)t]->tally_avg : g_mesh_tallies[(size_t)t]->tally_var;
  for (size_t i = 0; i < a.size(); i++) out[i] = a[i];
}
// mesh tallies of the device after ref_power_iteration_gpu: number, size, and average (which = 1) / error of the mean (3)
int ref_gpu_ntallies() { return g_gpu_transporter ? g_gpu_transporter->ntallies() : 0; }
uint64_t ref_gpu_tally_get(int t, int which, double* out, uint64_t cap) {
  const std::vector<double> v = g_gpu_transporter->tally(t, which);
  for (uint64_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
  return v.size();
}
void ref_gpu_release() { g_gpu_transporter.reset(); g_transporter.reset(); }
// the reference's own simulation_timer of the last ref_power_iteration[_gpu]: the generation loop without initialisation
double ref_last_simulation_seconds() { return g_last_simulation_seconds; }
// OpenMP threads of the next ref_transport calls (1 for the bit-exact pins: score sums in bank order; all cores for timing)
void ref_set_threads(int n) { g_threads = n > 0 ? n : 1; }
// wall time of the transport() call alone inside the last ref_transport (bank construction and copies excluded)
double ref_last_transport_seconds() { return g_last_transport_seconds; }
// generation scores of mesh tally t after the last ref_transport (C order [E][x][y][z]); size 0 for a source tally
uint64_t ref_tally_size(int t) { return g_tally_gen[(size_t)t] ? g_tally_gen[(size_t)t]->size() : 0; }
int ref_ntallies() { return (int)g_tally_gen.size(); }
void ref_tally_get(int t, double* out) {
  const NDArray<double>& a = *g_tally_gen[(size_t)t];
  for (size_t i = 0; i < a.size(); i++) out[i] = a[i];
}

}  // extern "C"

namespace {
struct DriverParts {
  std::vector<std::shared_ptr<Source>> sources;
  std::shared_ptr<Cancelator> cancelator;
  std::vector<std::string> entropy_lines;
};
// sources, cancelator and entropy lines of the deck text through the reference's plain constructors
DriverParts driver_parts(const char* text) {
  DriverParts d;
  std::istringstream in(text);
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "src") {
      double w, lo[3], hi[3], E;
      int fissile_only;
      std::string kind, ekey;
      ls >> w >> fissile_only >> kind;
      std::shared_ptr<SpatialDistribution> sp;
      if (kind == "box") {
        ls >> lo[0] >> lo[1] >> lo[2] >> hi[0] >> hi[1] >> hi[2];
        sp = std::make_shared<Box>(Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]));
      } else {
        ls >> lo[0] >> lo[1] >> lo[2];
        sp = std::make_shared<Point>(Position(lo[0], lo[1], lo[2]));
      }
      ls >> ekey >> E;
      std::shared_ptr<EnergyDistribution> ed;
      std::string dkey, dkind;
      ls >> dkey;
      if (dkey == "maxwellian" || dkey == "watt") {
        double a, b;
        ls >> a;
        if (dkey == "watt") {
          ls >> b;
          ed = std::make_shared<Watt>(a, b);
        } else {
          ed = std::make_shared<Maxwellian>(a);
        }
        ls >> dkey;
      }
      if (!ed) ed = std::make_shared<MonoEnergetic>(E);
      ls >> dkind;
      std::shared_ptr<DirectionDistribution> dd = std::make_shared<Isotropic>();
      if (dkind == "mono" || dkind == "cone") {
        double x, y, z, aperture = 0.;
        ls >> x >> y >> z;
        if (dkind == "cone") {
          ls >> aperture;
          dd = std::make_shared<Cone>(Direction(x, y, z), aperture);
        } else {
          dd = std::make_shared<MonoDirectional>(Direction(x, y, z));
        }
      }
      d.sources.push_back(std::make_shared<Source>(sp, dd, ed, fissile_only != 0, w));
    } else if (key == "cancel") {
      int a, b, c;
      ls >> a >> b >> c;
      settings::regional_cancellation = a != 0;
      settings::regional_cancellation_noise = b != 0;
      settings::n_cancel_noise_gens = c;
    } else if (key == "cancelator") {
      int on;
      ls >> on;
      if (on == 2) {  // type: basic-exact (src/basic_exact_mg_cancelator.cpp:610-705): beta 0 zero, 1 minimum, 2 average-f, 3 average-g
        uint32_t nx, ny, nz, nsmp;
        double lo[3], hi[3];
        int beta, sobol;
        ls >> nx >> ny >> nz >> lo[0] >> lo[1] >> lo[2] >> hi[0] >> hi[1] >> hi[2] >> beta >> sobol >> nsmp;
        const BasicExactMGCancelator::BetaMode modes[4] = {BasicExactMGCancelator::BetaMode::Zero, BasicExactMGCancelator::BetaMode::Minimum,
                                                           BasicExactMGCancelator::BetaMode::OptAverageF, BasicExactMGCancelator::BetaMode::OptAverageGain};
        d.cancelator = std::make_shared<BasicExactMGCancelator>(Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]), nx, ny, nz,
                                                                modes[beta], sobol != 0, nsmp);
      } else if (on == 3) {  // type: exact (src/exact_mg_cancelator.cpp:594-686): n-samples, group bins (count, then groups per bin)
        size_t nx, ny, nz, nbins;
        double lo[3], hi[3];
        uint32_t nsmp;
        ls >> nx >> ny >> nz >> lo[0] >> lo[1] >> lo[2] >> hi[0] >> hi[1] >> hi[2] >> nsmp >> nbins;
        std::vector<std::vector<std::size_t>> group_bins(nbins);
        for (auto& b : group_bins) {
          size_t cnt;
          ls >> cnt;
          b.resize(cnt);
          for (auto& g : b) ls >> g;
        }
        d.cancelator = std::make_shared<ExactMGCancelator>(Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]),
                                                           std::array<std::size_t, 4>{nx, ny, nz, nbins}, group_bins, settings::chi_matrix,
                                                           settings::use_virtual_collisions, nsmp);
      } else if (on) {
        uint32_t nx, ny, nz;
        double lo[3], hi[3];
        size_t ne;
        ls >> nx >> ny >> nz >> lo[0] >> lo[1] >> lo[2] >> hi[0] >> hi[1] >> hi[2] >> ne;
        std::vector<double> eb(ne);
        for (auto& e : eb) ls >> e;
        if (ne)
          d.cancelator = std::make_shared<ApproximateMeshCancelator>(Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]), nx, ny, nz, eb);
        else
          d.cancelator = std::make_shared<ApproximateMeshCancelator>(Position(lo[0], lo[1], lo[2]), Position(hi[0], hi[1], hi[2]), nx, ny, nz);
      }
    } else if (key == "entropy") {
      d.entropy_lines.push_back(line);
    }
  }
  return d;
}
uint64_t g_last_bank_size = 0;
// settings: pair-distance-sqrd / families / empty-entropy-bins of the last run (src/power_iterator.cpp:283-297) and its final bank
std::vector<double> g_diag_r_sqrd, g_diag_families, g_diag_empty, g_last_bank_xyzw;
void set_entropy(Simulation& sim, const std::vector<std::string>& lines) {
  for (const auto& el : lines) {
    std::istringstream ls(el);
    std::string key;
    int on;
    ls >> key >> on;
    if (!on) continue;
    double lo[3], hi[3];
    uint32_t sh[3];
    ls >> lo[0] >> lo[1] >> lo[2] >> hi[0] >> hi[1] >> hi[2] >> sh[0] >> sh[1] >> sh[2];
    const Position low_r(lo[0], lo[1], lo[2]), hi_r(hi[0], hi[1], hi[2]);
    const std::array<uint32_t, 3> shp{sh[0], sh[1], sh[2]};
    sim.set_p_pre_entropy(std::make_shared<Entropy>(low_r, hi_r, shp, Entropy::Sign::Positive));
    sim.set_n_pre_entropy(std::make_shared<Entropy>(low_r, hi_r, shp, Entropy::Sign::Negative));
    sim.set_t_pre_entropy(std::make_shared<Entropy>(low_r, hi_r, shp, Entropy::Sign::Total));
    sim.set_p_post_entropy(std::make_shared<Entropy>(low_r, hi_r, shp, Entropy::Sign::Positive));
    sim.set_n_post_entropy(std::make_shared<Entropy>(low_r, hi_r, shp, Entropy::Sign::Negative));
    sim.set_t_post_entropy(std::make_shared<Entropy>(low_r, hi_r, shp, Entropy::Sign::Total));
  }
}
template <class Iterator>
void run_iterator(DriverParts& d, int ngen, double* kcol, double* ktrk, double* leak, double* mig, double* entropy) {
  std::shared_ptr<Iterator> pi;
  pi = d.cancelator ? std::make_shared<Iterator>(g_tallies, g_transporter, d.sources, d.cancelator)
                    : std::make_shared<Iterator>(g_tallies, g_transporter, d.sources);
  set_entropy(*pi, d.entropy_lines);
  pi->initialize();
  pi->run();
  g_last_simulation_seconds = pi->simulation_timer.elapsed_time();  // the generation loop (src/power_iterator.cpp:316-318,432)
  g_last_bank_size = pi->bank.size();
  g_diag_r_sqrd = pi->r_sqrd_vec;
  g_diag_families.assign(pi->families_vec.begin(), pi->families_vec.end());
  g_diag_empty = pi->empty_entropy_frac_vec;
  g_last_bank_xyzw.clear();
  for (const Particle& p : pi->bank)
    for (double v : {p.r().x(), p.r().y(), p.r().z(), p.wgt()}) g_last_bank_xyzw.push_back(v);
  const Tallies& T = *g_tallies;
  for (int g = 0; g < ngen; g++) {
    kcol[g] = T.k_col_vec[(size_t)g]; ktrk[g] = T.k_trk_vec[(size_t)g]; leak[g] = T.leak_vec[(size_t)g]; mig[g] = T.mig_vec[(size_t)g];
    entropy[g] = (size_t)g < pi->t_pre_entropy_vec.size() ? pi->t_pre_entropy_vec[(size_t)g] : 0.;
  }
}
}  // namespace

extern "C" {
// flatten_problem() on the deck's live objects, dumped in the format of ablh_dump_tables (abeille_b200/host/capi.cpp) so that the
// two flatteners -- this one from the reference's objects, that one from the YAML deck -- can be compared value for value
int ref_flatten_dump(const char* text, char* out, long long out_cap) {
  try {
    if (ref_problem_load(text) != 0) return 1;
    DriverParts d = driver_parts(text);
    abl_integration::FlatProblem F;
    abl_integration::flatten_problem(F, *g_tallies, d.cancelator.get());
    std::string s;
    char buf[64];
    auto dump = [&](const char* name, const std::vector<double>& v) {
      s += name;
      for (double x : v) {
        std::snprintf(buf, sizeof buf, " %.17g", x);
        s += buf;
      }
      s += "\n";
    };
    auto dumpi = [&](const char* name, const std::vector<int32_t>& v) {
      s += name;
      for (int32_t x : v) s += " " + std::to_string(x);
      s += "\n";
    };
    dump("Et", F.Et); dump("Ea", F.Ea); dump("Ef", F.Ef); dump("Es", F.Es); dump("nu", F.nu); dump("nud", F.nud);
    dump("chi_cdf", F.chi_cdf); dump("scatter_cdf", F.scatter_cdf); dump("amu", F.amu); dump("apdf", F.apdf); dump("acdf", F.acdf);
    dump("smp", F.smp);
    dumpi("rpn", std::vector<int32_t>(F.rpn.begin(), F.rpn.begin() + F.p.nrpn)); dumpi("universe_cells", F.universe_cells);
    dumpi("lattice_tiles", std::vector<int32_t>(F.lattice_tiles.begin(), F.lattice_tiles.begin() + F.p.n_lattice_tiles));
    std::vector<int32_t> ang;
    for (const auto& a : F.angle) { ang.push_back(a.offset); ang.push_back(a.n); }
    dumpi("angle", ang);
    // geometry records and tallies, which ablh_dump_tables does not print: surfaces (type, bc, parameters), cells, universes
    std::vector<double> geo;
    for (const auto& sf : F.surfaces) { geo.push_back(sf.type); geo.push_back(sf.bc); for (double v : sf.p) geo.push_back(v); }
    dump("surfaces", geo);
    std::vector<int32_t> cl;
    for (const auto& c : F.cells) { cl.push_back(c.rpn_offset); cl.push_back(c.rpn_len); cl.push_back(c.simple); cl.push_back(c.vac_or_refl); cl.push_back(c.fill_universe); cl.push_back(c.material); }
    dumpi("cells", cl);
    std::vector<double> un;
    for (const auto& u : F.universes) {
      for (double v : {(double)u.type, (double)u.has_bc, (double)u.cell_offset, (double)u.ncells, (double)u.N[0], (double)u.N[1], (double)u.N[2],
                       (double)u.tile_offset, (double)u.outer, u.P[0], u.P[1], u.P[2], u.Pinv[0], u.Pinv[1], u.Pinv[2], u.Xl[0], u.Xl[1], u.Xl[2]})
        un.push_back(v);
    }
    dump("universes", un);
    dumpi("root", {F.p.root_universe});
    if (static_cast<long long>(s.size()) + 1 > out_cap) return 2;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_flatten_dump: %s\n", e.what());
    return 1;
  }
}

// the reference's vendored Sobol sequence (vendor/sobol), the points BasicExactMGCancelator::sample_position_sobol uses
void ref_sobol_points(int n, double* out4n) {
  for (int i = 0; i < n; i++)
    for (unsigned d = 0; d < 4; d++) out4n[4 * i + d] = sobol::sample(static_cast<unsigned long long>(i), d);
}
// size of the source bank the last ref_power_iteration ended with (after combing, for a branchless deck)
uint64_t ref_last_bank_size() { return g_last_bank_size; }
// the optional diagnostics of the next ref_power_iteration (settings::pair_distance_sqrd, families, empty_entropy_bins) ...
void ref_set_diagnostics(int pair_distance, int families, int empty_entropy) {
  settings::pair_distance_sqrd = pair_distance != 0;
  settings::families = families != 0;
  settings::empty_entropy_bins = empty_entropy != 0;
}
// ... their per-generation series after it (each array holds ngen values when its option was on, else nothing is written;
// the three lengths are returned in n3) and the final bank, [n][4] = x y z wgt, normalised, in bank order
void ref_pi_diagnostics(double* r_sqrd, double* families, double* empty, uint64_t cap, uint64_t n3[3]) {
  const std::vector<double>* v[3] = {&g_diag_r_sqrd, &g_diag_families, &g_diag_empty};
  double* out[3] = {r_sqrd, families, empty};
  for (int k = 0; k < 3; k++) {
    n3[k] = v[k]->size();
    for (uint64_t i = 0; i < v[k]->size() && i < cap; i++) out[k][i] = (*v[k])[i];
  }
}
uint64_t ref_last_bank_get(double* xyzw, uint64_t cap_rows) {
  const uint64_t n = g_last_bank_xyzw.size() / 4;
  for (uint64_t i = 0; i < 4 * std::min(n, cap_rows); i++) xyzw[i] = g_last_bank_xyzw[i];
  return n;
}

// The reference's own PowerIterator::initialize() + run() (src/power_iterator.cpp:170-473) on the deck text: sources,
// entropy mesh and cancelator are built through their plain constructors from the "src", "entropy" and "cancelator" lines
// (src/source.cpp:92-140, src/parser.cpp:1008-1054, src/cancelator.cpp:32-78).  Out, per generation g < ngen: k_col, k_trk,
// leakage, migration area (Tallies' generation vectors), the total pre-cancellation entropy; summary = k_col avg / err,
// k_trk avg / err, leakage avg / err.  (This file is compiled with -fno-access-control to read the private per-generation
// vectors; access control does not enter the object layout.)
int ref_power_iteration_gpu(const char* text, const char* host_library, const char* yaml_deck, int device, int ngen, int nignored,
                            double* kcol, double* ktrk, double* leak, double* mig, double* entropy, double* summary);
int ref_power_iteration(const char* text, int ngen, int nignored, double* kcol, double* ktrk, double* leak, double* mig,
                        double* entropy, double* summary) {
  return ref_power_iteration_gpu(text, nullptr, nullptr, 0, ngen, nignored, kcol, ktrk, leak, mig, entropy, summary);
}
// The same, with the transporter replaced by GPUTransporter (integration/gpu_transporter.hpp) when host_library is given:
// the reference's PowerIterator::run() -- its own source sampling, entropy, cancellation, normalisation and statistics --
// drives the B200 backend through the C ABI.  yaml_deck is the YAML file of the same deck (the backend's host library
// flattens it); mesh tallies are then scored on the device only and the reference's Tallies object holds none.
int ref_power_iteration_gpu(const char* text, const char* host_library, const char* yaml_deck, int device, int ngen, int nignored,
                            double* kcol, double* ktrk, double* leak, double* mig, double* entropy, double* summary) {
  try {
    if (ref_problem_load(text) != 0) return 1;
    DriverParts d = driver_parts(text);
    if (host_library) {
      // yaml_deck empty: `host_library` is libabeille_b200.so itself and the problem tables come from the reference's live objects
      // (integration/flatten_problem.hpp) -- settings, geometry::, materials, the mesh tallies ref_problem_load gave the Tallies object
      abl_integration::FlatProblem flat;
      const bool from_objects = yaml_deck == nullptr || yaml_deck[0] == 0;
      if (from_objects) abl_integration::flatten_problem(flat, *g_tallies, d.cancelator.get());
      g_tallies = std::make_shared<Tallies>(static_cast<double>(settings::nparticles));
      g_tallies->set_keff(settings::keff);
      g_tally_gen.clear();
      g_mesh_tallies.clear();
      g_gpu_transporter = from_objects ? std::make_shared<GPUTransporter>(g_tallies, host_library, flat.p, device)
                                       : std::make_shared<GPUTransporter>(g_tallies, host_library, yaml_deck, device);
      g_transporter = g_gpu_transporter;
    }
    omp_set_num_threads(g_threads);
    settings::ngenerations = ngen;
    settings::nignored = nignored;
    // branchless-k-eigenvalue decks run the reference's BranchlessPowerIterator (src/branchless_power_iterator.cpp: the same loop
    // plus comb_particles), which draws from settings::rng as ref_problem_load left it (seeded, stream 2, no colour draws)
    if (settings::mode == settings::SimulationMode::BRANCHLESS_K_EIGENVALUE)
      run_iterator<BranchlessPowerIterator>(d, ngen, kcol, ktrk, leak, mig, entropy);
    else
      run_iterator<PowerIterator>(d, ngen, kcol, ktrk, leak, mig, entropy);
    if (host_library) g_gpu_transporter->finish();
    const Tallies& T = *g_tallies;
    summary[0] = T.kcol_avg(); summary[1] = T.kcol_err(); summary[2] = T.ktrk_avg(); summary[3] = T.ktrk_err();
    summary[4] = T.leakage_avg(); summary[5] = T.leakage_err();
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_power_iteration: %s\n", e.what());
    return 1;
  }
}

// The reference's own Noise::initialize() + run() (src/noise.cpp:211-559): nignored power-iteration generations, then
// nbatches noise batches of nskip - 1 plain generations, one generation that samples the noise source, and the noise
// simulation of that source (inner generations, regional cancellation of the noise fission banks).  Out: k_col of every
// power-iteration generation and 0 per noise batch (Tallies' generation vector); final_bank3 = size of the last bank, its
// first history id, the global history counter; the mesh tallies are read with ref_tally_get_stat.
// The reference's ModifiedFixedSource::run() (src/modified_fixed_source.cpp:59-141): per batch the source is sampled and
// transported, the fission bank becomes the next bank (weights kept, fresh history ids) until it is empty, then the
// generation values and mesh tallies are recorded.  Out: k_col, leakage and migration area of every batch, the number of
// histories transported; the mesh tallies are read with ref_tally_get_stat.
int ref_modified_fixed_source(const char* text, int nbatches, double* kcol, double* leak, double* mig, uint64_t* transported) {
  try {
    if (ref_problem_load(text) != 0) return 1;
    omp_set_num_threads(g_threads);
    settings::ngenerations = nbatches;
    DriverParts d = driver_parts(text);
    auto sim = std::make_shared<ModifiedFixedSource>(g_tallies, g_transporter, d.sources);
    sim->initialize();
    sim->run();
    const Tallies& T = *g_tallies;
    for (int g = 0; g < nbatches; g++) {
      kcol[g] = T.k_col_vec[(size_t)g]; leak[g] = T.leak_vec[(size_t)g]; mig[g] = T.mig_vec[(size_t)g];
    }
    *transported = sim->transported_histories;
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_modified_fixed_source: %s\n", e.what());
    return 1;
  }
}

// The reference's FixedSource::run() (src/fixed_source.cpp:81-175): every batch samples the source and transports it; fission
// neutrons are secondaries of the history that made them (transporter.cpp:460-463), so transport() returns an empty bank.
int ref_fixed_source(const char* text, int nbatches, double* kcol, double* leak, double* mig) {
  try {
    if (ref_problem_load(text) != 0) return 1;
    omp_set_num_threads(g_threads);
    settings::ngenerations = nbatches;
    DriverParts d = driver_parts(text);
    auto sim = std::make_shared<FixedSource>(g_tallies, g_transporter, d.sources);
    sim->initialize();
    sim->run();
    const Tallies& T = *g_tallies;
    for (int g = 0; g < nbatches; g++) {
      kcol[g] = T.k_col_vec[(size_t)g]; leak[g] = T.leak_vec[(size_t)g]; mig[g] = T.mig_vec[(size_t)g];
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_fixed_source: %s\n", e.what());
    return 1;
  }
}

int ref_noise_run_gpu(const char* text, const char* host_library, const char* yaml_deck, int device, int nbatches, int nignored, int nskip,
                      double* kcol, int* n_kcol, uint64_t* final_bank3);
int ref_noise_run(const char* text, int nbatches, int nignored, int nskip, double* kcol, int* n_kcol, uint64_t* final_bank3) {
  return ref_noise_run_gpu(text, nullptr, nullptr, 0, nbatches, nignored, nskip, kcol, n_kcol, final_bank3);
}
// ... with host_library != nullptr: the reference's own Noise::run() over GPUTransporter (integration/gpu_transporter.hpp) -- the
// power-iteration generations, the generations that sample the noise source and the noise particles' inner generations all go
// through abl_transport / abl_transport_noise; source normalisation, cancellation of the noise banks and bank hand-over stay the
// reference's
int ref_noise_run_gpu(const char* text, const char* host_library, const char* yaml_deck, int device, int nbatches, int nignored, int nskip,
                      double* kcol, int* n_kcol, uint64_t* final_bank3) {
  try {
    if (ref_problem_load(text) != 0) return 1;
    if (host_library) {
      abl_integration::FlatProblem flat;  // yaml_deck empty: the tables come from the reference's live objects, noise sources included
      const bool from_objects = yaml_deck == nullptr || yaml_deck[0] == 0;
      if (from_objects) {
        DriverParts dp = driver_parts(text);
        abl_integration::flatten_problem(flat, *g_tallies, dp.cancelator.get(), g_noise_maker.get());
      }
      g_tallies = std::make_shared<Tallies>(static_cast<double>(settings::nparticles));
      g_tallies->set_keff(settings::keff);
      g_tally_gen.clear();
      g_mesh_tallies.clear();
      g_gpu_transporter = from_objects ? std::make_shared<GPUTransporter>(g_tallies, host_library, flat.p, device)
                                       : std::make_shared<GPUTransporter>(g_tallies, host_library, yaml_deck, device);
      g_transporter = g_gpu_transporter;
    }
    omp_set_num_threads(g_threads);
    settings::ngenerations = nbatches;
    settings::nignored = nignored;
    settings::nskip = nskip;
    DriverParts d = driver_parts(text);
    std::shared_ptr<Noise> sim = d.cancelator ? std::make_shared<Noise>(g_tallies, g_transporter, d.sources, d.cancelator, *g_noise_maker)
                                              : std::make_shared<Noise>(g_tallies, g_transporter, d.sources, *g_noise_maker);
    set_entropy(*sim, d.entropy_lines);
    sim->initialize();
    sim->run();
    const Tallies& T = *g_tallies;
    final_bank3[0] = sim->bank.size(); final_bank3[1] = sim->bank.empty() ? 0 : sim->bank.front().history_id();
    final_bank3[2] = sim->global_histories_counter;
    *n_kcol = (int)T.k_col_vec.size();
    for (size_t g = 0; g < T.k_col_vec.size(); g++) kcol[g] = T.k_col_vec[g];
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_noise_run: %s\n", e.what());
    return 1;
  }
}

}  // extern "C"
