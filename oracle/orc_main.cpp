/* orc_main.cpp -- TEST INFRASTRUCTURE: CPU oracle of Abeille's MG transport hot path.
 *
 * This is a restatement (not a copy) of the reference's algorithm, used ONLY by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  The product (abeille_b200/) never links or calls it.
 *
 * PARITY PIN STATUS: PINNED against the reference's own compiled code for the k-eigenvalue hot path.  The reference
 * ships no tests or golden vectors and its CMake build needs the network, but its multigroup
 * translation units compile where they lie under /root/reference with stand-in headers for the downloaded libraries
 * (oracle/Makefile target `ref`, oracle/ref_probe.cpp, oracle/ref_shim/): oracle/_ref/libabeille_ref.so runs the
 * reference's own SurfaceTracker / DeltaTracker / CarterTracker::transport, and this oracle reproduces its fission bank,
 * generation values and mesh-tally scores bit for bit on 14 decks, as well as surfaces, directions, RNG helpers, angle
 * tables, MGNuclide sampling and the geometry cursor piece by piece (tests/test_reference_pins.py, golden vectors in
 * tests/golden/ref_pins.npz made by scripts/make_ref_pins.py).  Further pins: RNG known-answer vectors from pcg32 +
 * libstdc++ (tests/golden/rng_kat.json), the Sood analytic k values quoted in the reference's decks, k_col == k_abs.
 * Noise mode is pinned the same way (transport with complex weights; NoiseMaker::sample_noise_source with the square-
 * oscillation and flat-vibration sources), and so are whole k-eigenvalue simulations: the reference's PowerIterator::run with
 * source sampling, entropy, approximate mesh cancellation, normalisation and tally statistics.  The Noise driver (oracle/api.py run_noise) is
 * pinned against the reference's Noise::run the same way.
 *
 * Follows: src/delta_tracker.cpp:72-263, src/surface_tracker.cpp:40-219,
 * src/carter_tracker.cpp:53-294, src/transporter.cpp:35-93,269-487,
 * src/power_iterator.cpp:305-431,538-586,751-777, src/simulation.cpp:55-77,
 * src/source.cpp:44-90, src/box.cpp:37-42, src/isotropic.cpp:28-36,
 * src/entropy.cpp:32-93, src/majorant.cpp:133-176,
 * src/approximate_mesh_cancelator.cpp:97-190, src/noise.cpp, src/noise_maker.cpp,
 * src/square_oscillation_noise_source.cpp (noise mode).
 */
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <set>
#include <sstream>
#include <unordered_map>

#include "orc_detmath.h"
#include "orc_geom.h"
#include "orc_phys.h"
#include "orc_rng.h"
#include "orc_tally.h"

namespace orc {

static double libm_log(double x) { return std::log(x); }
static double libm_sin(double x) { return std::sin(x); }
static double libm_cos(double x) { return std::cos(x); }
static double libm_exp(double x) { return std::exp(x); }
static double libm_acos(double x) { return std::acos(x); }
MathFns g_math = {libm_log, libm_sin, libm_cos, libm_exp, libm_acos};

// ------------------------------------------------------------------------------------
struct Source {  // src/source.cpp, box.cpp, point.cpp, isotropic.cpp, mono_energetic.cpp
  double weight = 1.;
  bool fissile_only = false;
  bool is_box = true;
  Vec low{0, 0, 0}, hi{0, 0, 0};  // point: low == position
  double energy = 1.;
  int dir_kind = 0;               // 0 isotropic, 1 mono-directional (mono_directional.hpp:38), 2 cone (cone.cpp:31-42)
  Vec dir{0, 0, 1};               // normalised, as Direction(x, y, z) leaves it
  double cos_aperture = 1.;       // Cone::Cone stores std::cos(aperture)
  int en_kind = 0;                // 0 mono-energetic, 1 maxwellian (maxwellian.cpp:34-42), 2 watt (watt.cpp:42-47)
  double en_a = 0., en_b = 0.;
};

struct Cancelator {  // kind 1: ApproximateMeshCancelator, kind 2: BasicExactMGCancelator (beta 0 zero, 1 minimum, 2 average-f, 3 average-g)
  bool present = false;
  int kind = 1, beta = 0;
  bool sobol = true;
  uint32_t nsamples = 10;
  std::vector<std::vector<size_t>> group_bins;  // kind 3 (ExactMGCancelator): groups of every energy bin
  Vec low{0, 0, 0}, hi{0, 0, 0};
  uint32_t shape[4] = {1, 1, 1, 1};
  std::vector<double> energy_edges;
  double dx = 0, dy = 0, dz = 0;
};

struct EntropyMesh {  // src/entropy.cpp + include/simulation/entropy.hpp
  bool present = false;
  Vec low{0, 0, 0}, up{0, 0, 0};
  uint32_t shape[3] = {1, 1, 1};
  double dx = 0, dy = 0, dz = 0;
  std::vector<double> bins;
  double total_weight = 0.;
  void init() {
    bins.assign((size_t)shape[0] * shape[1] * shape[2], 0.);
    dx = (up.x - low.x) / static_cast<double>(shape[0]);
    dy = (up.y - low.y) / static_cast<double>(shape[1]);
    dz = (up.z - low.z) / static_cast<double>(shape[2]);
  }
  void zero() { std::fill(bins.begin(), bins.end(), 0.); total_weight = 0.; }
  void add_point(const Vec& r, double w) {  // sign == Total
    int32_t nx = static_cast<int32_t>(std::floor((r.x - low.x) / dx));
    int32_t ny = static_cast<int32_t>(std::floor((r.y - low.y) / dy));
    int32_t nz = static_cast<int32_t>(std::floor((r.z - low.z) / dz));
    if (nx >= 0 && nx < (int32_t)shape[0] && ny >= 0 && ny < (int32_t)shape[1] && nz >= 0 && nz < (int32_t)shape[2]) {
      total_weight += w;
      bins[(size_t)((shape[1] * shape[2]) * (uint32_t)nx + (shape[2]) * (uint32_t)ny + (uint32_t)nz)] += w;
    }
  }
  double calculate_entropy() const {
    double sum = 0.;
    for (const auto& b : bins) {
      double p = std::fabs(b) / total_weight;
      if (p > 1.0) {
      } else if (p != 0.) {
        sum -= p * std::log2(p);
      }
    }
    return sum;
  }
};

struct NoiseSource {  // src/square_oscillation_noise_source.cpp, src/flat_vibration_noise_source.cpp
  bool vibration = false;
  Vec low{0, 0, 0}, hi{0, 0, 0};
  double w0 = 0, eps_t = 0, eps_f = 0, eps_s = 0;  // oscillation
  int basis = 0, mat_pos = -1, mat_neg = -1;       // vibration: axis, material indices (== nuclide ids in MG)
  double x0 = 0, eps = 0;                          // vibration: interface position and half width (:66-82)
  bool is_inside(const Vec& r) const {
    return r.x > low.x && r.y > low.y && r.z > low.z && r.x < hi.x && r.y < hi.y && r.z < hi.z;
  }
};

// BasicExactMGCancelator's bins (include/simulation/basic_exact_mg_cancelator.hpp:48-90).  The reference keys the outer map with
// {i, j, k} hashed as std::hash<int>(k + Nz (j + Ny i)) and the inner one with the Material pointer; here the outer key is that
// integer (same hash values, same insertion sequence, same libstdc++: same iteration order) and the inner key the material index
// (identical for bins that hold one material; with several the reference's order follows heap addresses).
struct ExactCancelBin {
  struct Averages { double f = 0., f_inv = 0.; };
  double uniform_wgt = 0., uniform_wgt2 = 0., W = 0., W2 = 0., sum_c = 0., sum_c_wgt = 0., sum_c_wgt2 = 0.;
  uint64_t rng_seed_advance = 0;
  bool can_cancel = true;
  std::vector<BankedParticle*> particles;
  std::vector<Averages> averages;
};
using ExactBins = std::unordered_map<int, std::unordered_map<int, ExactCancelBin>>;

struct Problem {
  Settings st;
  Geometry geo;
  std::vector<Material> materials;
  std::map<uint32_t, int> material_id_to_indx;
  std::vector<Source> sources;
  Tallies tallies;
  Cancelator cancel;
  EntropyMesh entropy;
  std::vector<NoiseSource> noise_sources;
  std::vector<double> majorant;  // per group  (src/majorant.cpp:133-176)
  std::vector<double> sampling;  // carter: ratio*majorant (src/carter_tracker.cpp:60-75)
  bool converged = false;
  uint64_t histories_counter = 0, global_histories_counter = 0;
  Pcg32Stream global_rng;  // settings::rng
  std::vector<double> last_parent_info;  // x, y, z, Esmp per row of the fission bank of the last orc_transport call
  std::vector<double> last_parent_state;  // ... previous direction, previous previous energy, previous energy, was_virtual
  bool chi_matrix = false;  // settings::chi_matrix
  ExactBins exact_bins;    // BasicExactMGCancelator::bins (kept across generations: clear() keeps the bucket array)
  Counters counters;
  std::string error;
  // per-history trace of the LAST transport call (instrumentation)
  bool want_trace = false;
  std::vector<uint32_t> tr_flights, tr_real, tr_virtual, tr_fission;
  std::vector<uint64_t> tr_hash, tr_rng_state;

  void build_majorant() {
    size_t G = st.ngroups;
    majorant.assign(G, 0.);
    for (const auto& m : materials)
      for (size_t g = 0; g < G; g++) {
        double xs = 0. + 1. * m.Et[g];
        if (xs > majorant[g]) majorant[g] = xs;
      }
    sampling = majorant;
    if (st.tracking == Settings::CARTER)
      for (size_t g = 0; g < G; g++) sampling[g] = majorant[g] * st.sample_xs_ratio[g];
  }
};

// ---- deck reader -------------------------------------------------------------------------
struct Tok {
  std::vector<std::string> t;
  size_t i = 0;
  const std::string& next() {
    if (i >= t.size()) throw std::runtime_error("deck: unexpected end");
    return t[i++];
  }
  void expect(const char* s) {
    const std::string& v = next();
    if (v != s) throw std::runtime_error(std::string("deck: expected '") + s + "' got '" + v + "'");
  }
  double d() { return std::strtod(next().c_str(), nullptr); }
  long long ll() { return std::strtoll(next().c_str(), nullptr, 10); }
  unsigned long long ull() { return std::strtoull(next().c_str(), nullptr, 10); }
  std::vector<double> dv(size_t n) {
    std::vector<double> v(n);
    for (auto& x : v) x = d();
    return v;
  }
};

static int surf_type_from(const std::string& s) {
  if (s == "xplane") return S_XPLANE;
  if (s == "yplane") return S_YPLANE;
  if (s == "zplane") return S_ZPLANE;
  if (s == "plane") return S_PLANE;
  if (s == "xcylinder") return S_XCYL;
  if (s == "ycylinder") return S_YCYL;
  if (s == "zcylinder") return S_ZCYL;
  if (s == "cylinder") return S_CYL;
  if (s == "sphere") return S_SPHERE;
  throw std::runtime_error("deck: unknown surface type " + s);
}

static Problem* load_problem(const char* path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error(std::string("cannot open ") + path);
  Tok tk;
  std::string w;
  while (f >> w) tk.t.push_back(w);
  auto P = std::make_unique<Problem>();
  Settings& st = P->st;
  tk.expect("ORCDECK");
  tk.ll();
  tk.expect("mode");
  // ("mfs": modified-fixed-source.  Its transport is the k-eigenvalue one with n_new = floor(|k_abs_scr| + xi) -- no division by
  // k_col (transporter.cpp:381-386) --, which the driver obtains by setting k_col to 1: x / 1 is x exactly.)
  {
    const std::string mode = tk.next();
    st.mode = mode == "noise" ? Settings::NOISE : (mode == "fs" ? Settings::FIXED_SOURCE : (mode == "branchless" ? Settings::BRANCHLESS : Settings::K_EIGENVALUE));
    if (st.mode == Settings::BRANCHLESS) {  // branchless-k-eigenvalue: material, splitting, combing (parser.cpp:367-409)
      st.branchless_material = tk.ll() != 0;
      st.branchless_splitting = tk.ll() != 0;
      st.branchless_combing = tk.ll() != 0;
    }
  }
  tk.expect("tracking");
  {
    std::string t = tk.next();
    st.tracking = t == "delta" ? Settings::DELTA
                  : t == "carter" ? Settings::CARTER
                  : t == "implicit" ? Settings::IMPLICIT_LEAKAGE : Settings::SURFACE;
  }
  tk.expect("ngroups");
  st.ngroups = (uint32_t)tk.ll();
  const size_t G = st.ngroups;
  tk.expect("ebounds");
  st.energy_bounds = tk.dv(G + 1);
  // parser.cpp:158-168 with mg_nuclide.cpp:425-427: the materials' common energy range narrows the defaults
  if (st.energy_bounds.front() > st.min_energy) st.min_energy = st.energy_bounds.front();
  if (st.energy_bounds.back() < st.max_energy) st.max_energy = st.energy_bounds.back();
  tk.expect("nparticles"); st.nparticles = (int)tk.ll();
  tk.expect("ngenerations"); st.ngenerations = (int)tk.ll();
  tk.expect("nignored"); st.nignored = (int)tk.ll();
  tk.expect("nskip"); st.nskip = (int)tk.ll();
  tk.expect("wgt"); st.wgt_cutoff = tk.d(); st.wgt_survival = tk.d(); st.wgt_split = tk.d();
  tk.expect("seed"); st.rng_seed = tk.ull();
  tk.expect("stride"); st.rng_stride = tk.ull();
  tk.expect("ratios");
  { size_t n = (size_t)tk.ll(); st.sample_xs_ratio = tk.dv(n); }
  tk.expect("cancel");
  st.regional_cancellation = tk.ll() != 0;
  st.regional_cancellation_noise = tk.ll() != 0;
  st.n_cancel_noise_gens = (int)tk.ll();
  tk.expect("noise");
  st.w_noise = tk.d(); st.keff = tk.d(); st.inner_generations = tk.ll() != 0; st.normalize_noise_source = tk.ll() != 0;
  tk.expect("nmat");
  size_t M = (size_t)tk.ll();
  for (size_t m = 0; m < M; m++) {
    Material mat;
    mat.G = G;
    tk.expect("mat"); mat.id = (uint32_t)tk.ll();
    tk.expect("total"); mat.Et = tk.dv(G);
    tk.expect("absorption"); mat.Ea = tk.dv(G);
    tk.expect("fission"); mat.Ef = tk.dv(G);
    tk.expect("nu_p"); mat.nu_p = tk.dv(G);
    tk.expect("nu_d"); mat.nu_d = tk.dv(G);
    mat.has_nu_d = true;  // the reference's nu_delyd_ vector is never empty (mg_nuclide.cpp:737-738)
    tk.expect("speeds"); mat.speeds = tk.dv(G);
    tk.expect("chi");
    size_t nrows = (size_t)tk.ll();
    mat.chi.assign(G, std::vector<double>(G, 0.));
    {
      bool fissile_xs = false;
      for (double v : mat.Ef) if (v > 0.) fissile_xs = true;
      if (fissile_xs && nrows == G) P->chi_matrix = true;  // settings::chi_matrix (mg_nuclide.cpp:836-850; one group counts as a matrix)
    }
    if (nrows == 1) {
      auto row = tk.dv(G);
      for (size_t i = 0; i < G; i++) mat.chi[i] = row;  // mg_nuclide.cpp:836-850
    } else if (nrows == G) {
      for (size_t i = 0; i < G; i++) mat.chi[i] = tk.dv(G);
    }
    tk.expect("scatter");
    mat.Ps.resize(G);
    for (size_t i = 0; i < G; i++) mat.Ps[i] = tk.dv(G);
    tk.expect("nleg");
    size_t L = (size_t)tk.ll();
    std::vector<std::vector<Legendre>> leg(G, std::vector<Legendre>(G));
    for (size_t l = 1; l <= L; l++) {
      tk.expect("P");
      size_t order = (size_t)tk.ll();
      for (size_t i = 0; i < G; i++)
        for (size_t o = 0; o < G; o++) leg[i][o].set_moment(order, tk.d());
    }
    mat.angle.assign(G, std::vector<AngleDist>(G));
    for (size_t i = 0; i < G; i++)
      for (size_t o = 0; o < G; o++) mat.angle[i][o] = leg[i][o].linearize();
    tk.expect("ndg");
    size_t ND = (size_t)tk.ll();
    mat.P_delayed_group = tk.dv(ND);
    mat.decay_constants = tk.dv(ND);
    // fissile gate for nu/chi (mg_nuclide.cpp:718-733): non-fissile materials keep zero nu and chi
    bool fissile = false;
    for (double v : mat.Ef) if (v > 0.) fissile = true;
    if (!fissile) {
      mat.nu_p.assign(G, 0.);
      mat.nu_d.assign(G, 0.);
      mat.chi.assign(G, std::vector<double>(G, 0.));
    }
    mat.finish();
    P->material_id_to_indx[mat.id] = (int)P->materials.size();
    P->materials.push_back(mat);
  }
  Geometry& geo = P->geo;
  tk.expect("nsurf");
  size_t S = (size_t)tk.ll();
  for (size_t s = 0; s < S; s++) {
    Surface sf;
    tk.expect("surf");
    sf.id = (uint32_t)tk.ll();
    sf.type = surf_type_from(tk.next());
    std::string bc = tk.next();
    sf.bc = bc == "vacuum" ? BC_VACUUM : (bc == "reflective" ? BC_REFLECTIVE : BC_NORMAL);
    size_t np = (size_t)tk.ll();
    auto pv = tk.dv(np);
    if (sf.type == S_CYL) {
      sf.p[0] = pv[0]; sf.p[1] = pv[1]; sf.p[2] = pv[2]; sf.p[6] = pv[6];
      sf.finish_general_cylinder(pv[3], pv[4], pv[5]);
    } else {
      for (size_t i = 0; i < np; i++) sf.p[i] = pv[i];
    }
    geo.surface_id_to_indx[sf.id] = (int)geo.surfaces.size();
    geo.surfaces.push_back(sf);
  }
  tk.expect("ncell");
  size_t C = (size_t)tk.ll();
  std::vector<std::pair<bool, uint32_t>> cell_fill(C);
  for (size_t c = 0; c < C; c++) {
    tk.expect("cell");
    uint32_t id = (uint32_t)tk.ll();
    std::string kind = tk.next();
    uint32_t fill = (uint32_t)tk.ll();
    std::string region = tk.next();
    Cell cell = geo.make_cell(region, id);
    cell.fill_universe = kind == "u";
    cell_fill[c] = {cell.fill_universe, fill};
    if (!cell.fill_universe) {
      auto it = P->material_id_to_indx.find(fill);
      if (it == P->material_id_to_indx.end()) throw std::runtime_error("deck: unknown material id");
      cell.material = it->second;
    }
    geo.cell_id_to_indx[id] = (int)geo.cells.size();
    geo.cells.push_back(cell);
  }
  tk.expect("nuni");
  size_t U = (size_t)tk.ll();
  std::vector<std::vector<long long>> lat_ids(U);
  std::vector<long long> outer_ids(U, -1);
  for (size_t u = 0; u < U; u++) {
    Universe uni;
    tk.expect("uni");
    uni.id = (uint32_t)tk.ll();
    std::string kind = tk.next();
    if (kind == "cells") {
      uni.type = U_CELLS;
      size_t n = (size_t)tk.ll();
      for (size_t i = 0; i < n; i++) {
        uint32_t cid = (uint32_t)tk.ll();
        auto it = geo.cell_id_to_indx.find(cid);
        if (it == geo.cell_id_to_indx.end()) throw std::runtime_error("deck: unknown cell id");
        uni.cell_indices.push_back((uint32_t)it->second);
      }
    } else if (kind == "rect") {
      uni.type = U_RECT;
      uni.Nx = (uint32_t)tk.ll(); uni.Ny = (uint32_t)tk.ll(); uni.Nz = (uint32_t)tk.ll();
      uni.Px = tk.d(); uni.Py = tk.d(); uni.Pz = tk.d();
      uni.Px_inv = 1. / uni.Px; uni.Py_inv = 1. / uni.Py; uni.Pz_inv = 1. / uni.Pz;
      double xl = tk.d(), yl = tk.d(), zl = tk.d();
      uni.Xl = xl - static_cast<double>(uni.Nx) * 0.5 * uni.Px;  // rect_lattice.cpp:49-51
      uni.Yl = yl - static_cast<double>(uni.Ny) * 0.5 * uni.Py;
      uni.Zl = zl - static_cast<double>(uni.Nz) * 0.5 * uni.Pz;
      outer_ids[u] = tk.ll();
      size_t n = (size_t)tk.ll();
      for (size_t i = 0; i < n; i++) lat_ids[u].push_back(tk.ll());
    } else if (kind == "hex") {  // make_hex_lattice, src/hex_lattice.cpp:457-577: nrings nz pitch pitch_z origin top outer n ids
      uni.type = U_HEX;
      uni.Nrings = (uint32_t)tk.ll(); uni.Nz = (uint32_t)tk.ll();
      uni.pitch = tk.d(); uni.pitch_z = tk.d();
      uni.X_o = tk.d(); uni.Y_o = tk.d(); uni.Z_o = tk.d();
      uni.top = (int)tk.ll();
      uni.width = 2 * (uni.Nrings - 1) + 1;  // hex_lattice.cpp:53-55
      uni.mid_qr = uni.width / 2;
      const double PI_ = 3.14159265358979323846264338327950288;  // utils/constants.hpp; hex_lattice.hpp:60-63
      uni.cos_pi_6 = std::cos(PI_ / 6.0); uni.sin_pi_6 = std::sin(PI_ / 6.0);
      uni.cos_pi_3 = std::cos(PI_ / 3.0); uni.sin_pi_3 = std::sin(PI_ / 3.0);
      outer_ids[u] = tk.ll();
      size_t n = (size_t)tk.ll();
      for (size_t i = 0; i < n; i++) lat_ids[u].push_back(tk.ll());
    } else {
      throw std::runtime_error("deck: unsupported universe kind " + kind);
    }
    geo.universe_id_to_indx[uni.id] = (int)geo.universes.size();
    geo.universes.push_back(uni);
  }
  for (size_t u = 0; u < U; u++) {
    Universe& uni = geo.universes[u];
    if (uni.type == U_RECT) {
      for (long long id : lat_ids[u]) uni.lattice_universes.push_back(id < 0 ? -1 : geo.universe_id_to_indx.at((uint32_t)id));
      uni.outer_universe_index = outer_ids[u] < 0 ? -1 : geo.universe_id_to_indx.at((uint32_t)outer_ids[u]);
    } else if (uni.type == U_HEX) {  // HexLattice::set_elements, hex_lattice.cpp:204-222: the ids fill the hexagon row by row
      uint32_t nhex = 0;
      for (uint32_t r = 0; r < uni.Nrings; r++) nhex += r == 0 ? 1 : 6 * r;
      if (lat_ids[u].size() != (size_t)nhex * uni.Nz) throw std::runtime_error("deck: Improper number of universes for HexLattice.");
      uni.lattice_universes.assign((size_t)uni.width * uni.width * uni.Nz, -1);
      size_t indx = 0;
      for (uint32_t az = 0; az < uni.Nz; az++)
        for (uint32_t ar = 0; ar < uni.width; ar++)
          for (uint32_t aq = 0; aq < uni.width; aq++) {
            const int32_t q = static_cast<int32_t>(aq - uni.mid_qr), r = static_cast<int32_t>(ar - uni.mid_qr);
            if (Geometry::hex_ring(q, r) < uni.Nrings) {
              const long long id = lat_ids[u][indx++];
              uni.lattice_universes[Geometry::hex_linear_index(uni, q, r, (int32_t)az)] = id < 0 ? -1 : geo.universe_id_to_indx.at((uint32_t)id);
            }
          }
      uni.outer_universe_index = outer_ids[u] < 0 ? -1 : geo.universe_id_to_indx.at((uint32_t)outer_ids[u]);
    }
  }
  for (size_t c = 0; c < C; c++)
    if (cell_fill[c].first) geo.cells[c].universe = geo.universe_id_to_indx.at(cell_fill[c].second);
  geo.finalize_bc_flags();
  tk.expect("root");
  geo.root = geo.universe_id_to_indx.at((uint32_t)tk.ll());
  tk.expect("nsrc");
  size_t NS = (size_t)tk.ll();
  for (size_t s = 0; s < NS; s++) {
    Source src;
    tk.expect("src");
    src.weight = tk.d();
    src.fissile_only = tk.ll() != 0;
    std::string sp = tk.next();
    if (sp == "box") {
      src.is_box = true;
      src.low = {tk.d(), tk.d(), tk.d()};
      src.hi = {tk.d(), tk.d(), tk.d()};
    } else {
      src.is_box = false;
      src.low = {tk.d(), tk.d(), tk.d()};
    }
    tk.expect("energy");
    src.energy = tk.d();
    std::string dk = tk.next();
    if (dk == "maxwellian" || dk == "watt") {
      src.en_kind = dk == "watt" ? 2 : 1;
      src.en_a = tk.d();
      if (src.en_kind == 2) src.en_b = tk.d();
      dk = tk.next();
    }
    if (dk != "dir") throw std::runtime_error("expected dir, got " + dk);
    dk = tk.next();
    if (dk == "mono" || dk == "cone") {
      const double dx = tk.d(), dy = tk.d(), dz = tk.d();
      src.dir = make_direction(dx, dy, dz);
      src.dir_kind = 1;
      if (dk == "cone") {
        src.cos_aperture = std::cos(tk.d());  // the host's libm, once, like the reference's constructor
        src.dir_kind = 2;
      }
    } else if (dk != "iso") {
      throw std::runtime_error("unknown source direction kind " + dk);
    }
    P->sources.push_back(src);
  }
  tk.expect("ntally");
  size_t T = (size_t)tk.ll();
  P->tallies.total_weight = static_cast<double>(st.nparticles);  // parser.cpp:870-871
  for (size_t t = 0; t < T; t++) {
    MeshTally mt;
    tk.expect("tally");
    mt.name = tk.next();
    mt.estimator = (int)tk.ll();
    mt.quantity = (int)tk.ll();
    mt.noise_source = tk.ll() != 0;
    mt.Nx = (uint64_t)tk.ll(); mt.Ny = (uint64_t)tk.ll(); mt.Nz = (uint64_t)tk.ll();
    mt.r_low = {tk.d(), tk.d(), tk.d()};
    mt.r_hi = {tk.d(), tk.d(), tk.d()};
    size_t ne = (size_t)tk.ll();
    mt.energy_bounds = tk.dv(ne);
    mt.net_weight = P->tallies.total_weight;
    mt.init();
    P->tallies.mesh.push_back(std::move(mt));
  }
  P->tallies.keff_ = st.keff;
  tk.expect("cancelator");
  if (const long long ckind = tk.ll(); ckind != 0) {
    Cancelator& c = P->cancel;
    c.present = true;
    c.kind = (int)ckind;
    c.shape[0] = (uint32_t)tk.ll(); c.shape[1] = (uint32_t)tk.ll(); c.shape[2] = (uint32_t)tk.ll();
    c.low = {tk.d(), tk.d(), tk.d()};
    c.hi = {tk.d(), tk.d(), tk.d()};
    if (c.kind == 2) {
      c.beta = (int)tk.ll(); c.sobol = tk.ll() != 0; c.nsamples = (uint32_t)tk.ll();
    } else if (c.kind == 3) {
      c.nsamples = (uint32_t)tk.ll();
      c.group_bins.resize((size_t)tk.ll());
      for (auto& b : c.group_bins) {
        b.resize((size_t)tk.ll());
        for (auto& g : b) g = (size_t)tk.ll();
      }
      c.shape[3] = c.group_bins.empty() ? 1 : (uint32_t)c.group_bins.size();
    } else {
      size_t ne = (size_t)tk.ll();
      c.energy_edges = tk.dv(ne);
      c.shape[3] = ne >= 2 ? (uint32_t)(ne - 1) : 1;
    }
    c.dx = (c.hi.x - c.low.x) / static_cast<double>(c.shape[0]);
    c.dy = (c.hi.y - c.low.y) / static_cast<double>(c.shape[1]);
    c.dz = (c.hi.z - c.low.z) / static_cast<double>(c.shape[2]);
  }
  tk.expect("entropy");
  if (tk.ll() != 0) {
    EntropyMesh& e = P->entropy;
    e.present = true;
    e.low = {tk.d(), tk.d(), tk.d()};
    e.up = {tk.d(), tk.d(), tk.d()};
    e.shape[0] = (uint32_t)tk.ll(); e.shape[1] = (uint32_t)tk.ll(); e.shape[2] = (uint32_t)tk.ll();
    e.init();
  }
  tk.expect("nnoise");
  size_t NN = (size_t)tk.ll();
  for (size_t n = 0; n < NN; n++) {
    NoiseSource ns;
    const std::string kind = tk.next();
    ns.low = {tk.d(), tk.d(), tk.d()};
    ns.hi = {tk.d(), tk.d(), tk.d()};
    ns.w0 = tk.d();
    if (kind == "sqosc") {
      ns.eps_t = tk.d(); ns.eps_f = tk.d(); ns.eps_s = tk.d();
    } else if (kind == "flatvib") {
      ns.vibration = true;
      ns.basis = (int)tk.ll(); ns.mat_pos = (int)tk.ll(); ns.mat_neg = (int)tk.ll();
      const double lo = ns.basis == 0 ? ns.low.x : (ns.basis == 1 ? ns.low.y : ns.low.z);
      const double hi = ns.basis == 0 ? ns.hi.x : (ns.basis == 1 ? ns.hi.y : ns.hi.z);
      ns.x0 = 0.5 * (lo + hi);
      ns.eps = (hi - lo) / 2.;
    } else {
      throw std::runtime_error("unknown noise source kind " + kind);
    }
    P->noise_sources.push_back(ns);
  }
  P->build_majorant();
  return P.release();
}

// ---- material helper (MG, one nuclide, N = 1) -------------------------------------------------
struct Mat {  // include/materials/material_helper.hpp
  const Problem* P;
  int m;
  Mat(const Problem* p, int mi) : P(p), m(mi) {}
  const Material& mat() const { return P->materials[(size_t)m]; }
  double Ew(double E, bool noise) const {  // :65-84
    double Ew_ = 0.;
    if (noise) {
      size_t g = P->st.group(E);
      Ew_ += P->st.eta * P->st.w_noise / mat().speeds[g];
    }
    return Ew_;
  }
  double Et(double E, bool noise = false) const {  // :47-63
    double Et_ = 0.;
    Et_ += 1. * mat().micro(P->st.group(E)).total;
    if (noise) Et_ += Ew(E, noise);
    return Et_;
  }
  double Ea(double E) const { double v = 0.; v += 1. * mat().micro(P->st.group(E)).absorption; return v; }
  double Ef(double E) const { double v = 0.; v += 1. * mat().micro(P->st.group(E)).fission; return v; }
  double vEf(double E) const {
    double v = 0.;
    const MicroXS x = mat().micro(P->st.group(E));
    v += 1. * x.nu_total * x.fission;
    return v;
  }
  double Eelastic(double E) const { double v = 0.; v += 1. * mat().micro(P->st.group(E)).elastic; return v; }
  double Es(double E) const {  // :100-113
    double v = 0.;
    const MicroXS x = mat().micro(P->st.group(E));
    v += 1. * std::max(x.total - x.absorption, 0.);
    return v;
  }
  MicroXS sample_branchless_nuclide(double E, Pcg32& rng) const {  // :226-275: one nuclide, xi = rand * sum <= sum always picks it
    (void)rng_rand(rng);
    return mat().micro(P->st.group(E));
  }
  MeshTally::MatXS tally_xs(double E) const { return {Et(E), Ea(E), Ef(E), Eelastic(E)}; }
  MicroXS sample_nuclide(double E, Pcg32& rng, bool noise) const {  // :178-224
    const double Ew_ = Ew(E, noise);
    const double xi = rng_rand(rng);  // always drawn, even with one nuclide
    (void)xi;
    MicroXS micro = mat().micro(P->st.group(E));
    if (noise) {
      micro.noise_copy = Ew_ / (1. * 1.);
      micro.total += micro.noise_copy;
    }
    return micro;
  }
};

struct Ctx {
  Problem* P;
  bool noise;
  bool sample_noise_source;  // a NoiseMaker was passed
  std::vector<BankedParticle>* noise_bank;
  Counters cn;
  ThreadLocalScores ts;
};

// ---- samplers (src/mg_nuclide.cpp:442-543) -------------------------------------------------------
struct ScatterInfo { double energy; Vec direction; };
static ScatterInfo sample_scatter(const Problem& P, const Material& nuc, const Vec& u, size_t g, Pcg32& rng) {
  size_t ei = static_cast<size_t>(rng_discrete(rng, nuc.Ps_cp[g]));
  double E_out = 0.5 * (P.st.energy_bounds[ei] + P.st.energy_bounds[ei + 1]);
  double mu = nuc.angle[g][ei].sample_mu(rng);
  double phi = 2. * PI * rng_rand(rng);
  return {E_out, rotate_direction(u, mu, phi)};
}
struct FissionInfo { double energy; Vec direction; bool delayed; double lambda; };
template <class Engine>
static FissionInfo sample_fission(const Problem& P, const Material& nuc, const Vec& u, size_t g, double Pdelayed, Engine& rng) {
  size_t ei = static_cast<size_t>(rng_discrete(rng, nuc.chi_cp[g]));
  double E_out = 0.5 * (P.st.energy_bounds[ei] + P.st.energy_bounds[ei + 1]);
  double mu = 2. * rng_rand(rng) - 1.;
  double phi = 2. * PI * rng_rand(rng);
  FissionInfo info{E_out, rotate_direction(u, mu, phi), false, 0.};
  if (rng_rand(rng) < Pdelayed) {
    size_t dgrp = static_cast<size_t>(rng_discrete(rng, nuc.dg_cp));
    info.delayed = true;
    info.lambda = nuc.decay_constants[dgrp];
  }
  return info;
}

// ---- collision physics (src/transporter.cpp) ---------------------------------------------------------
static void russian_roulette(const Settings& st, Particle& p) {  // :35-58
  if (std::abs(p.wgt()) < st.wgt_cutoff) {
    double P_kill = 1.0 - (std::abs(p.wgt()) / st.wgt_survival);
    if (rng_rand(p.rng) < P_kill) p.state.weight = 0.;
    else p.state.weight = std::copysign(st.wgt_survival, p.wgt());
  }
  if (std::abs(p.wgt2()) < st.wgt_cutoff) {
    double P_kill = 1.0 - (std::abs(p.wgt2()) / st.wgt_survival);
    if (rng_rand(p.rng) < P_kill) p.state.weight2 = 0.;
    else p.state.weight2 = std::copysign(st.wgt_survival, p.wgt2());
  }
  if (p.wgt() == 0. && p.wgt2() == 0.) p.kill();
}

static void make_fission_neutrons(Ctx& cx, Particle& p, const MicroXS& microxs, const Material& nuc, bool noise) {  // :358-487
  const Settings& st = cx.P->st;
  double k_abs_scr = p.wgt() * microxs.nu_total * microxs.fission / microxs.total;
  int n_new = 0;
  if (st.mode == Settings::K_EIGENVALUE || (st.mode == Settings::NOISE && !noise)) {
    n_new = static_cast<int>(std::floor(std::abs(k_abs_scr) / cx.P->tallies.k_col + rng_rand(p.rng)));
  } else if (st.mode == Settings::FIXED_SOURCE) {  // transporter.cpp:374-379: no normalisation by k
    n_new = static_cast<int>(std::floor(std::abs(k_abs_scr) + rng_rand(p.rng)));
  } else {
    n_new = static_cast<int>(std::floor((microxs.nu_total * microxs.fission / (microxs.total * cx.P->tallies.keff_)) + rng_rand(p.rng)));
  }
  double P_delayed = microxs.nu_delayed / microxs.nu_total;
  for (int i = 0; i < n_new; i++) {
    auto finfo = sample_fission(*cx.P, nuc, p.u(), microxs.energy_index, P_delayed, p.rng);
    double wgt = p.wgt() > 0. ? 1. : -1.;
    double wgt2 = 0.;
    if (noise) {
      wgt = p.wgt();
      wgt2 = p.wgt2();
      if (finfo.delayed) {
        std::complex<double> wgt_cmpx{wgt, wgt2};
        double lambda = finfo.lambda;
        double denom = (lambda * lambda) + (st.w_noise * st.w_noise);
        std::complex<double> mult{lambda * lambda / denom, -lambda * st.w_noise / denom};
        wgt_cmpx *= mult;
        wgt = wgt_cmpx.real();
        wgt2 = wgt_cmpx.imag();
      }
    }
    BankedParticle fp{p.r(), finfo.direction, finfo.energy, wgt, wgt2, p.history_id, p.daughter_counter(), p.family_id};
    fp.parents_previous_position = p.previous_position;
    fp.Esmp_parent = p.Esmp_;
    fp.parents_previous_direction = p.previous_direction;
    fp.parents_previous_previous_energy = p.previous_energy;
    fp.parents_previous_energy = p.E();
    fp.parents_previous_was_virtual = p.previous_collision_virtual;
    if (st.mode == Settings::FIXED_SOURCE) {  // transporter.cpp:460-463: the fission neutron continues the history
      p.make_secondary(fp.u, fp.E, fp.wgt, fp.wgt2);
    } else if (st.mode == Settings::K_EIGENVALUE || !noise || st.inner_generations) {
      p.history_fission_bank.push_back(fp);
    } else {
      p.make_secondary(fp.u, fp.E, fp.wgt, fp.wgt2);
    }
    p.n_fission++;
    cx.cn.fission_sites++;
  }
  p.note(0x5000000000000000ULL | (uint64_t)(uint32_t)n_new);
}

// Transporter::branchless_collision_iso / _mat (transporter.cpp:104-267): every collision is either a scatter that carries the
// multiplicity m in its weight or a fission that banks ONE site of weight w*m and ends the particle.
static void branchless_collision(Ctx& cx, Particle& p, const Mat& mat) {
  const Settings& st = cx.P->st;
  const Material& nuc = mat.mat();
  MicroXS microxs;
  double Pscatter, m;
  bool scatter = false;
  if (st.branchless_material) {  // :183-267
    const double Es = mat.Es(p.E());
    const double vEf = mat.vEf(p.E());
    const double Et = mat.Et(p.E());
    Pscatter = Es / (vEf + Es);
    m = (vEf + Es) / Et;
    if (rng_rand(p.rng) < Pscatter) scatter = true;
    microxs = mat.sample_branchless_nuclide(p.E(), p.rng);
    const double m_i = (microxs.nu_total * microxs.fission + (microxs.total - microxs.absorption)) / microxs.total;
    const double k_abs_scr = (m / m_i) * p.wgt() * microxs.nu_total * microxs.fission / microxs.total;
    cx.ts.k_abs += k_abs_scr;
    russian_roulette(st, p);
    if (!p.alive) return;
  } else {  // :104-181
    microxs = mat.sample_nuclide(p.E(), p.rng, false);
    const double k_abs_scr = p.wgt() * microxs.nu_total * microxs.fission / microxs.total;
    cx.ts.k_abs += k_abs_scr;
    Pscatter = (microxs.total - microxs.absorption) / (microxs.nu_total * microxs.fission + (microxs.total - microxs.absorption));
    m = (microxs.nu_total * microxs.fission + (microxs.total - microxs.absorption)) / microxs.total;
    russian_roulette(st, p);
    if (!p.alive) return;
    if (rng_rand(p.rng) < Pscatter) scatter = true;
  }
  if (scatter) {
    ScatterInfo s = sample_scatter(*cx.P, nuc, p.u(), microxs.energy_index, p.rng);  // (yield == 1 in multi-group)
    p.set_energy(s.energy);
    p.set_direction(s.direction);
    p.state.weight = p.wgt() * m * 1.;
    p.state.weight2 = p.wgt2() * m * 1.;
    if (p.E() < st.min_energy) p.kill();
    if (st.branchless_splitting && p.alive && std::abs(p.wgt()) >= st.wgt_split) {
      // the material flavour rounds up, the isotope flavour down (:239-241 against :151-155)
      const int n_new = static_cast<int>(st.branchless_material ? std::ceil(std::abs(p.wgt())) : std::floor(std::abs(p.wgt())));
      p.split(n_new);
    }
  } else {
    const double P_delayed = microxs.nu_delayed / microxs.nu_total;
    auto finfo = sample_fission(*cx.P, nuc, p.u(), microxs.energy_index, P_delayed, p.rng);
    BankedParticle fp{p.r(), finfo.direction, finfo.energy, p.wgt() * m, p.wgt2() * m, p.history_id, p.daughter_counter(), p.family_id};
    fp.parents_previous_position = p.previous_position;
    fp.Esmp_parent = p.Esmp_;
    fp.parents_previous_direction = p.previous_direction;
    fp.parents_previous_previous_energy = p.previous_energy;
    fp.parents_previous_energy = p.E();
    fp.parents_previous_was_virtual = p.previous_collision_virtual;
    p.history_fission_bank.push_back(fp);
    p.n_fission++;
    cx.cn.fission_sites++;
    p.kill();
  }
}

// noise source sampling (src/noise_maker.cpp:277-445, square_oscillation_noise_source.cpp:76-170)
static void sample_noise_source(Ctx& cx, Particle& p, const Mat& mat, double keff, double w);

static void collision(Ctx& cx, Particle& p, const Mat& mat, bool noise) {  // :60-93 + :269-312
  Problem& P = *cx.P;
  const Settings& st = P.st;
  cx.cn.real_collisions++;
  p.n_real++;
  if (P.converged) {  // tallies.hpp:49-55
    MeshTally::MatXS mx = mat.tally_xs(p.E());
    for (auto& t : P.tallies.mesh)
      if (t.estimator == EST_COLLISION) t.score_collision(p, mx, cx.cn);
  }
  if (!noise) {
    double k_col_scr = p.wgt() * mat.vEf(p.E()) / mat.Et(p.E(), noise);
    double mig_dist = (p.r() - p.r_birth).norm();
    double mig_area_scr = p.wgt() * mat.Ea(p.E()) / mat.Et(p.E()) * mig_dist * mig_dist;
    cx.ts.k_col += k_col_scr;
    cx.ts.mig += mig_area_scr;
  }
  if (cx.sample_noise_source) sample_noise_source(cx, p, mat, P.tallies.keff_, st.w_noise);
  if (st.mode == Settings::BRANCHLESS) {  // :81-88
    if (noise) throw std::runtime_error("Cannot perform noise simulations with branchless collisions.");
    branchless_collision(cx, p, mat);
    p.note(0x6000000000000000ULL | (p.alive ? (uint64_t)(st.group(p.E()) + 1) : 0ULL));
    return;
  }

  // branching_collision :269-312
  MicroXS microxs = mat.sample_nuclide(p.E(), p.rng, noise);
  const Material& nuc = mat.mat();
  if (!noise) {
    double k_abs_scr = p.wgt() * microxs.nu_total * microxs.fission / microxs.total;
    cx.ts.k_abs += k_abs_scr;
  }
  make_fission_neutrons(cx, p, microxs, nuc, noise);
  if (noise) {  // make_noise_copy :349-356
    std::complex<double> weight_copy{p.wgt(), p.wgt2()};
    if (microxs.noise_copy / microxs.total + rng_rand(p.rng) >= 1.) {
      std::complex<double> yield{1., -1. / st.eta};
      weight_copy *= yield;
      p.make_secondary(p.u(), p.E(), weight_copy.real(), weight_copy.imag());
    }
  }
  p.state.weight = p.wgt() * (1. - (microxs.absorption + microxs.noise_copy) / microxs.total);
  p.state.weight2 = p.wgt2() * (1. - (microxs.absorption + microxs.noise_copy) / microxs.total);
  russian_roulette(st, p);
  if (p.alive) {
    ScatterInfo s = sample_scatter(P, nuc, p.u(), microxs.energy_index, p.rng);  // do_scatter :314-347 (yield == 1)
    p.set_direction(s.direction);
    p.set_energy(s.energy);
    p.state.weight = p.wgt() * 1.;
    p.state.weight2 = p.wgt2() * 1.;
    if (p.E() < st.min_energy) p.kill();
  }
  p.note(0x6000000000000000ULL | (p.alive ? (uint64_t)(st.group(p.E()) + 1) : 0ULL));
}

// Tracker::do_reflection (tracker.hpp:314-360)
static void do_reflection(Tracker& trkr, Particle& p, const Boundary& boundary) {
  const Geometry& geo = *trkr.geo;
  if (boundary.surface_index < 0) throw std::runtime_error("Bad surface index in Tracker::do_reflection");
  const Surface& surface = geo.surfaces[(size_t)boundary.surface_index];
  int32_t token = boundary.surface_index + 1;
  if (surface.sign(p.r(), p.u()) < 0) token *= -1;
  trkr.surface_token_ = token;
  Vec r_pre_refs = p.r();
  if (p.reflected) r_pre_refs = p.previous_position;
  Vec u = p.u();
  Vec r_on_surf = p.r() + boundary.distance * u;
  Vec n = surface.norm(r_on_surf);
  Vec new_dir = u - 2. * (u.dot(n)) * n;
  Vec u_new = make_direction(new_dir.x, new_dir.y, new_dir.z);
  double d = boundary.distance + (p.r() - r_pre_refs).norm();
  Vec r_prev = r_on_surf - d * u_new;
  p.set_position(r_on_surf);
  p.previous_position = r_prev;
  p.set_direction(u_new);
  p.reflected = true;
  trkr.set_r(p.r());  // NB: wipes the token again
  trkr.set_u(p.u());
  trkr.restart_get_current();
}

static void score_flight(Ctx& cx, const Particle& p, double d, const Mat& mat) {  // tallies.hpp:57-63
  Problem& P = *cx.P;
  if (!P.converged) return;
  bool any = false;
  for (auto& t : P.tallies.mesh) if (t.estimator == EST_TRACK_LENGTH) any = true;
  if (!any) return;
  MeshTally::MatXS mx = mat.tally_xs(p.E());
  for (auto& t : P.tallies.mesh)
    if (t.estimator == EST_TRACK_LENGTH) t.score_flight(p, d, mx, cx.cn);
}

static void leak(Ctx& cx, Particle& p, const Boundary& bound) {
  p.kill();
  cx.ts.leakage += p.wgt();
  Vec r_leak = p.r() + bound.distance * p.u();
  // `p.wgt() * (r_leak - r_birth) * (r_leak - r_birth)` groups as (w * d) . d (position.hpp:80-96), not w * (d . d):
  // found by running the reference's own transport() next to this one (tests/test_reference_pins.py)
  const Vec d = r_leak - p.r_birth;
  const Vec wd{d.x * p.wgt(), d.y * p.wgt(), d.z * p.wgt()};
  cx.ts.mig += wd.dot(d);
}

static void try_resurrect(Ctx& cx, Particle& p, Tracker& trkr, Mat& mat) {
  p.resurect();
  if (p.alive) {
    trkr.set_r(p.r());
    trkr.set_u(p.u());
    trkr.restart_get_current();
    if (trkr.is_lost()) throw std::runtime_error("Particle has become lost after resurection.");
    mat.m = trkr.current_mat;
  }
  (void)cx;
}

// ---- the three history loops ---------------------------------------------------------------------
static void history_delta_or_carter(Ctx& cx, Particle& p, bool carter) {
  Problem& P = *cx.P;
  const Settings& st = P.st;
  const bool noise = cx.noise;
  Tracker trkr(&P.geo, p.r(), p.u());
  if (trkr.is_lost()) {
    cx.cn.lost_at_birth++;
    p.kill();
  }
  Mat mat(&P, trkr.current_mat);
  while (p.alive) {
    bool had_collision = false, crossed_boundary = false;
    const size_t g = st.group(p.E());
    double Esample = (carter ? P.sampling[g] : P.majorant[g]) + mat.Ew(p.E(), noise);
    p.Esmp_ = Esample;  // "Sampling XS saved for cancellation"
    double d_coll = rng_exponential(p.rng, Esample);
    Boundary bound(INF, -1, BC_NORMAL);
    cx.cn.flights++;
    p.n_flights++;
    trkr.move(d_coll);
    trkr.get_current();
    if (trkr.is_lost()) {
      trkr.set_r(p.r());
      trkr.get_current();
      bound = trkr.get_boundary_condition();
      crossed_boundary = true;
    }
    score_flight(cx, p, std::min(d_coll, bound.distance), mat);
    if (crossed_boundary) {
      cx.cn.boundary_events++;
      p.n_boundary++;
      if (bound.boundary_type == BC_VACUUM) {
        p.note(0x3000000000000000ULL | (uint64_t)(uint32_t)(trkr.current_cell + 1));
        leak(cx, p, bound);
      } else if (bound.boundary_type == BC_REFLECTIVE) {
        do_reflection(trkr, p, bound);
        if (trkr.is_lost()) throw std::runtime_error("Particle has become lost after reflection.");
        p.note(0x4000000000000000ULL | (uint64_t)(uint32_t)(trkr.current_cell + 1));
        if (carter) bound = trkr.get_boundary_condition();
      } else {
        throw std::runtime_error("Help me, how did I get here ?");
      }
    } else {
      p.move(d_coll);
      mat.m = trkr.current_mat;
      double Et = mat.Et(p.E(), noise);
      if (!carter) {
        if (Et - Esample > 1.E-10) throw std::runtime_error("Total cross section excedeed majorant");
        if (rng_rand(p.rng) < (Et / Esample)) had_collision = true;
      } else {
        if (Esample >= Et) {
          if (rng_rand(p.rng) < (Et / Esample)) had_collision = true;
        } else {
          double D = Et / (2. * Et - Esample);
          double F = Et / (D * Esample);
          if ((D - rng_rand(p.rng)) > 0.) {
            p.state.weight = p.wgt() * F;
            had_collision = true;
          } else {
            p.state.weight = -p.wgt() * F;
          }
        }
      }
      p.note((had_collision ? 0x2000000000000000ULL : 0x1000000000000000ULL) | (uint64_t)(uint32_t)(trkr.current_cell + 1));
    }
    if (p.alive && had_collision) {
      collision(cx, p, mat, noise);
      trkr.set_u(p.u());
      p.previous_collision_virtual = false;
    } else if (p.alive) {
      if (!crossed_boundary) { cx.cn.virtual_collisions++; p.n_virtual++; }
      p.previous_collision_virtual = true;
    }
    if (carter && p.alive && std::abs(p.wgt()) >= st.wgt_split) {
      int n_new = static_cast<int>(std::ceil(std::abs(p.wgt())));
      p.split(n_new);
    }
    if (!p.alive) try_resurrect(cx, p, trkr, mat);
  }
}

// ImplicitLeakageDeltaTracker::transport (src/implicit_leakage_delta_tracker.cpp:73-263): delta tracking that asks for the
// boundary condition before every flight; towards a vacuum boundary the weight that would fly out is scored as leakage
// and the flight distance is sampled from the exponential truncated at the boundary.
static void history_implicit_leakage(Ctx& cx, Particle& p) {
  Problem& P = *cx.P;
  const Settings& st = P.st;
  const bool noise = cx.noise;
  Tracker trkr(&P.geo, p.r(), p.u());
  if (trkr.is_lost()) {
    cx.cn.lost_at_birth++;
    p.kill();
  }
  Mat mat(&P, trkr.current_mat);
  while (p.alive) {
    bool had_collision = false, crossed_boundary = false;
    const size_t g = st.group(p.E());
    const double Emajorant = P.majorant[g] + mat.Ew(p.E(), noise);
    Boundary bound = trkr.get_boundary_condition();
    cx.cn.flights++;
    p.n_flights++;
    double d_coll = 0.;
    if (bound.boundary_type == BC_VACUUM) {  // :117-153
      const double P_leak = g_math.exp(-Emajorant * bound.distance);
      const double P_no_leak = 1. - P_leak;
      const double wgt_leak = p.wgt() * P_leak, wgt2_leak = p.wgt2() * P_leak;
      const double wgt_collides = p.wgt() * P_no_leak, wgt2_collides = p.wgt2() * P_no_leak;
      cx.ts.leakage += wgt_leak;
      {
        Vec r_leak = p.r() + bound.distance * p.u();
        const Vec d = r_leak - p.r_birth;
        const Vec wd{d.x * wgt_leak, d.y * wgt_leak, d.z * wgt_leak};  // (w * d) . d, as in leak()
        cx.ts.mig += wd.dot(d);
      }
      d_coll = -g_math.log(1. - P_no_leak * rng_rand(p.rng)) / Emajorant;
      p.state.weight = wgt_leak;
      p.state.weight2 = wgt2_leak;
      score_flight(cx, p, bound.distance, mat);
      p.state.weight = wgt_collides;
      p.state.weight2 = wgt2_collides;
      score_flight(cx, p, d_coll, mat);
    } else {
      d_coll = rng_exponential(p.rng, Emajorant);
      score_flight(cx, p, std::min(d_coll, bound.distance), mat);
    }
    if (bound.distance < d_coll || std::abs(bound.distance - d_coll) < BOUNDRY_TOL) {  // :165-194
      crossed_boundary = true;
      cx.cn.boundary_events++;
      p.n_boundary++;
      if (bound.boundary_type == BC_VACUUM) {
        p.note(0x3000000000000000ULL | (uint64_t)(uint32_t)(trkr.current_cell + 1));
        leak(cx, p, bound);
      } else if (bound.boundary_type == BC_REFLECTIVE) {
        do_reflection(trkr, p, bound);
        if (trkr.is_lost()) throw std::runtime_error("Particle has become lost after reflection.");
        p.note(0x4000000000000000ULL | (uint64_t)(uint32_t)(trkr.current_cell + 1));
      } else {
        throw std::runtime_error("Help me, how did I get here ?");
      }
    } else {  // :195-231
      p.move(d_coll);
      trkr.move(d_coll);
      trkr.get_current();
      if (trkr.is_lost()) throw std::runtime_error("Particle has become lost after a flight.");
      mat.m = trkr.current_mat;
      double Et = mat.Et(p.E(), noise);
      if (Et - Emajorant > 1.E-10) throw std::runtime_error("Total cross section excedeed majorant");
      if (rng_rand(p.rng) < (Et / Emajorant)) had_collision = true;
      p.note((had_collision ? 0x2000000000000000ULL : 0x1000000000000000ULL) | (uint64_t)(uint32_t)(trkr.current_cell + 1));
    }
    if (p.alive && had_collision) {
      collision(cx, p, mat, noise);
      trkr.set_u(p.u());
      p.previous_collision_virtual = false;
    } else if (p.alive) {
      if (!crossed_boundary) { cx.cn.virtual_collisions++; p.n_virtual++; }
      p.previous_collision_virtual = true;
    }
    if (!p.alive) try_resurrect(cx, p, trkr, mat);
  }
}

static void history_surface(Ctx& cx, Particle& p) {
  Problem& P = *cx.P;
  const bool noise = cx.noise;
  Tracker trkr(&P.geo, p.r(), p.u());
  if (trkr.is_lost()) {
    cx.cn.lost_at_birth++;
    p.kill();
  }
  Mat mat(&P, trkr.current_mat);
  while (p.alive) {
    bool had_collision = false;
    double d_coll = rng_exponential(p.rng, mat.Et(p.E(), noise));
    auto bound = trkr.get_nearest_boundary();
    cx.cn.flights++;
    p.n_flights++;
    score_flight(cx, p, std::min(d_coll, bound.distance), mat);
    double k_trk_scr = p.wgt() * std::min(d_coll, bound.distance) * mat.vEf(p.E());
    cx.ts.k_trk += k_trk_scr;
    if (bound.distance < d_coll || std::abs(bound.distance - d_coll) < BOUNDRY_TOL) {
      cx.cn.boundary_events++;
      p.n_boundary++;
      if (bound.boundary_type == BC_VACUUM) {
        p.note(0x3000000000000000ULL | (uint64_t)(uint32_t)(trkr.current_cell + 1));
        leak(cx, p, bound);
      } else if (bound.boundary_type == BC_REFLECTIVE) {
        do_reflection(trkr, p, bound);
        if (trkr.is_lost()) throw std::runtime_error("Particle has become lost after reflection.");
        p.note(0x4000000000000000ULL | (uint64_t)(uint32_t)(trkr.current_cell + 1));
      } else {
        trkr.cross_surface(bound);
        trkr.get_current();
        p.move(bound.distance);
        if (trkr.is_lost()) throw std::runtime_error("Particle has become lost after crossing a surface.");
        mat.m = trkr.current_mat;
        p.note(0x7000000000000000ULL | (uint64_t)(uint32_t)(trkr.current_cell + 1));
      }
    } else {
      p.move(d_coll);
      trkr.move(d_coll);
      had_collision = true;
      p.note(0x2000000000000000ULL | (uint64_t)(uint32_t)(trkr.current_cell + 1));
    }
    if (p.alive && had_collision) {
      collision(cx, p, mat, noise);
      trkr.set_u(p.u());
    }
    if (!p.alive) try_resurrect(cx, p, trkr, mat);
  }
}

// Transporter::transport  -- returns fission bank in bank order, drains noise banks
static std::vector<BankedParticle> transport(Problem& P, std::vector<Particle>& bank, bool noise,
                                             std::vector<BankedParticle>* noise_bank, bool noise_maker) {
  const int nth = omp_get_max_threads();
  std::vector<ThreadLocalScores> tss((size_t)nth);
  std::vector<Counters> cns((size_t)nth);
  std::string err;
#pragma omp parallel
  {
    Ctx cx{&P, noise, noise_maker && noise_bank, noise_bank, Counters(), ThreadLocalScores()};
#pragma omp for schedule(dynamic)
    for (size_t n = 0; n < bank.size(); n++) {
      try {
        Particle& p = bank[n];
        if (P.st.tracking == Settings::SURFACE) history_surface(cx, p);
        else if (P.st.tracking == Settings::IMPLICIT_LEAKAGE) history_implicit_leakage(cx, p);
        else history_delta_or_carter(cx, p, P.st.tracking == Settings::CARTER);
      } catch (const std::exception& e) {
#pragma omp critical
        err = e.what();
      }
    }
    tss[(size_t)omp_get_thread_num()] = cx.ts;
    cns[(size_t)omp_get_thread_num()] = cx.cn;
  }
  if (!err.empty()) throw std::runtime_error(err);
  for (int t = 0; t < nth; t++) {  // tallies->score_k_col(...) etc (delta_tracker.cpp:233-238)
    P.tallies.k_col_score += tss[(size_t)t].k_col;
    P.tallies.k_abs_score += tss[(size_t)t].k_abs;
    P.tallies.k_trk_score += tss[(size_t)t].k_trk;
    P.tallies.k_tot_score += tss[(size_t)t].k_tot;
    P.tallies.leak_score += tss[(size_t)t].leakage;
    P.tallies.mig_area_score += tss[(size_t)t].mig;
    P.counters.add(cns[(size_t)t]);
  }
  if (P.want_trace) {
    size_t N = bank.size();
    P.tr_flights.resize(N); P.tr_real.resize(N); P.tr_virtual.resize(N); P.tr_fission.resize(N);
    P.tr_hash.resize(N); P.tr_rng_state.resize(N);
    for (size_t n = 0; n < N; n++) {
      P.tr_flights[n] = bank[n].n_flights; P.tr_real[n] = bank[n].n_real; P.tr_virtual[n] = bank[n].n_virtual;
      P.tr_fission[n] = bank[n].n_fission; P.tr_hash[n] = bank[n].hash; P.tr_rng_state[n] = bank[n].rng.state;
    }
  }
  std::vector<BankedParticle> fission_neutrons;
  for (auto& p : bank) {
    fission_neutrons.insert(fission_neutrons.end(), p.history_fission_bank.begin(), p.history_fission_bank.end());
    p.history_fission_bank.clear();
  }
  if (noise_bank && noise_maker)
    for (auto& p : bank) {
      noise_bank->insert(noise_bank->end(), p.history_noise_bank.begin(), p.history_noise_bank.end());
      p.history_noise_bank.clear();
    }
  bank.clear();
  return fission_neutrons;
}

// ---- noise source (config 5) -------------------------------------------------------------------
// Frequency gate shared by every SquareOscillationNoiseSource factor (square_oscillation_noise_source.cpp:85-170):
// the factor is eps * pi (times Sigma_t for dEt) when w is the source's fundamental +-w0 within 1 %, else 0.
static bool sqosc_on(const NoiseSource& ns, double w) {
  int32_t n = static_cast<int32_t>(std::round(w / ns.w0));
  double err = (n * ns.w0 - w) / w;
  return (n == 1 || n == -1) && std::abs(err) < 0.01;
}

// FlatVibrationNoiseSource::C_R / C_L (src/flat_vibration_noise_source.cpp:155-192) for the harmonics this restatement
// (n = 0 cannot pass the frequency gate: its relative error is 1).  C_L differs from C_R for n = 0 only.
static std::complex<double> vib_C(const NoiseSource& ns, uint32_t n, double x) {
  double rel_diff = (x - ns.x0) / ns.eps;
  if (rel_diff > 1.) rel_diff = 1.;
  else if (rel_diff < -1.) rel_diff = -1.;
  if (n == 1) return {0., -2. * std::sqrt(1. - (rel_diff * rel_diff))};
  if (n == 2) return {-2. * rel_diff * std::sqrt(1. - (rel_diff * rel_diff)), 0.};
  if (n == 0) throw std::runtime_error("oracle: the zeroth harmonic never passes the frequency gate of dEt / dN");
  // n >= 3 (:176-179): (2 / n) sin(n acos(rel_diff)) exp(-i n pi / 2); std::exp of the complex argument (-0, -n pi / 2) is
  // polar(exp(-0), -n pi / 2) = (cos, sin) of the imaginary part
  const double dn = static_cast<double>(n);
  const double a = (2. / dn) * g_math.sin(dn * g_math.acos(rel_diff));
  const double th = -dn * PI * 0.5;
  return a * std::complex<double>{g_math.cos(th), g_math.sin(th)};
}
// the frequency gate and harmonic of dEt / dN (:222-244,277-312): 0 = no component at this frequency
static int vib_harmonic(const NoiseSource& ns, double w) {
  int32_t n = static_cast<int32_t>(std::round(w / ns.w0));
  double err = ((n * ns.w0) - w) / w;
  if (std::abs(err) > 0.01) return 0;
  return n;
}
static double vib_get_x(const NoiseSource& ns, const Vec& r) { return ns.basis == 0 ? r.x : (ns.basis == 1 ? r.y : r.z); }

// NoiseMaker::sample_noise_source (src/noise_maker.cpp:277-501): the noise copy (:155-181), the vibration sources
// (:446-501 with the homogenised "fake" material of :106-153) and the oscillation sources (:293-445).  Noise particles go
// to the history's noise bank.  In MG every material is one nuclide with atoms_bcm = 1 and nuclide id == material index.
static void sample_noise_source(Ctx& cx, Particle& p, const Mat& mat, double keff, double w) {
  Problem& P = *cx.P;
  bool inside = false, inside_vib = false, inside_osc = false;  // NoiseMaker::is_inside :93-104
  for (const auto& ns : P.noise_sources)
    if (ns.is_inside(p.r())) {
      inside = true;
      (ns.vibration ? inside_vib : inside_osc) = true;
    }
  if (!inside) return;
  const size_t g = P.st.group(p.E());

  {  // sample_noise_copy :155-181; NoiseMaker::dEt :60-78: vibration sources first, then oscillation sources
    std::complex<double> dEt{0., 0.};
    for (const auto& ns : P.noise_sources) {
      if (!ns.vibration || !ns.is_inside(p.r())) continue;  // FlatVibrationNoiseSource::dEt :222-244
      const double x = vib_get_x(ns, p.r());
      const double D_Et = Mat(&P, ns.mat_neg).Et(p.E()) - Mat(&P, ns.mat_pos).Et(p.E());
      const int n = vib_harmonic(ns, w);
      if (n == 0) { dEt += std::complex<double>{0., 0.}; continue; }
      if (n < 0) throw std::runtime_error("oracle: negative noise frequency");
      dEt += D_Et * vib_C(ns, (uint32_t)n, x);
    }
    for (const auto& ns : P.noise_sources) {
      if (ns.vibration || !ns.is_inside(p.r())) continue;  // SquareOscillationNoiseSource::dEt :85-113: looks the
      Tracker trkr(&P.geo, p.r(), Vec{1., 0., 0.});         // material up again with a fresh Tracker, direction (1,0,0)
      if (trkr.current_mat < 0) throw std::runtime_error("No material found at the position of a noise-source sample.");
      Mat fm(&P, trkr.current_mat);
      double xs = fm.Et(p.E());
      if (sqosc_on(ns, w)) dEt += std::complex<double>{ns.eps_t * xs * PI, 0.};
      else dEt += std::complex<double>{0., 0.};
    }
    const std::complex<double> dEt_Et = dEt / mat.Et(p.E());
    std::complex<double> weight_copy{p.wgt(), p.wgt2()};
    weight_copy *= -dEt_Et;
    BankedParticle np{p.r(), p.u(), p.E(), weight_copy.real(), weight_copy.imag(), p.history_id, p.daughter_counter(), p.family_id};
    p.history_noise_bank.push_back(np);
  }

  if (inside_vib) {  // sample_vibration_noise_source :446-501
    // make_fake_material :106-153: union of the nuclides of the sources we are in (sorted ids), each with the mean over
    // those sources of the source's own concentration (the mean of the two materials' where both hold the nuclide)
    std::vector<int> nuclides;
    std::vector<const NoiseSource*> sources;
    for (const auto& ns : P.noise_sources)
      if (ns.vibration && ns.is_inside(p.r())) {
        sources.push_back(&ns);
        for (int m : {ns.mat_neg, ns.mat_pos})
          if (std::find(nuclides.begin(), nuclides.end(), m) == nuclides.end()) nuclides.push_back(m);
      }
    std::sort(nuclides.begin(), nuclides.end());
    std::vector<double> conc(nuclides.size(), 0.);
    const double num_sources = static_cast<double>(sources.size());
    for (size_t i = 0; i < nuclides.size(); i++) {
      double conc_sum = 0.;
      for (const NoiseSource* ns : sources) {
        if (nuclides[i] != ns->mat_neg && nuclides[i] != ns->mat_pos)
          throw std::runtime_error("overlapping vibration noise sources with different nuclides");  // map::at throws
        double c = 1.;                                       // nuclide_info_: atoms_bcm of the one material that holds it,
        if (ns->mat_neg == ns->mat_pos) c = (1. + 1.) / 2.;  // or the mean of the two (:91-104)
        conc_sum += c;
      }
      conc[i] = conc_sum / num_sources;
    }
    double Et_fake = 0.;  // MaterialHelper::Et of the fake material
    for (size_t i = 0; i < nuclides.size(); i++) Et_fake += conc[i] * P.materials[(size_t)nuclides[i]].micro(g).total;
    // fake_mat.sample_nuclide(E, rng) :178-224
    const double invs_Et = 1. / Et_fake;
    const double xi = rng_rand(p.rng);
    size_t pick = nuclides.size() - 1;
    double prob_sum = 0.;
    for (size_t i = 0; i < nuclides.size(); i++) {
      const double nuc_prob = invs_Et * conc[i] * P.materials[(size_t)nuclides[i]].micro(g).total;
      prob_sum += nuc_prob;
      if (xi <= prob_sum) { pick = i; break; }
    }
    const int nuclide_id = nuclides[pick];
    const Material& nuc = P.materials[(size_t)nuclide_id];
    const MicroXS microxs = nuc.micro(g);
    const double N = conc[pick];
    std::complex<double> dN{0., 0.};  // NoiseMaker::dN :80-91, FlatVibrationNoiseSource::dN :277-312
    for (const NoiseSource* ns : sources) {
      if (nuclide_id != ns->mat_neg && nuclide_id != ns->mat_pos) { dN += std::complex<double>{0., 0.}; continue; }
      const double N_neg = ns->mat_neg == nuclide_id ? 1. : 0., N_pos = ns->mat_pos == nuclide_id ? 1. : 0.;
      const double D_N = N_neg - N_pos;
      const int n = vib_harmonic(*ns, w);
      if (n == 0) { dN += std::complex<double>{0., 0.}; continue; }
      dN += D_N * vib_C(*ns, (uint32_t)n, vib_get_x(*ns, p.r()));
    }
    const std::complex<double> dN_N = dN / N;
    const double Etfake_Et = Et_fake / mat.Et(p.E());
    if (nuc.fissile) {  // sample_vibration_noise_fission :183-237
      const double k_abs = microxs.nu_total * microxs.fission / microxs.total;
      const int n_new = static_cast<int>(std::floor(k_abs / keff + rng_rand(p.rng)));
      const double P_delayed = microxs.nu_delayed / microxs.nu_total;
      for (int i = 0; i < n_new; i++) {
        auto finfo = sample_fission(P, nuc, p.u(), microxs.energy_index, P_delayed, p.rng);
        BankedParticle bnp{p.r(), finfo.direction, finfo.energy, p.wgt(), p.wgt2(), p.history_id, p.daughter_counter(), p.family_id};
        if (finfo.delayed) {
          std::complex<double> wgt_cmpx{bnp.wgt, bnp.wgt2};
          double lambda = finfo.lambda;
          double denom = (lambda * lambda) + (w * w);
          std::complex<double> mult{lambda * lambda / denom, -lambda * w / denom};
          wgt_cmpx *= mult;
          bnp.wgt = wgt_cmpx.real();
          bnp.wgt2 = wgt_cmpx.imag();
        }
        std::complex<double> bnp_wgt{bnp.wgt, bnp.wgt2};
        bnp_wgt *= dN_N;
        bnp_wgt *= Etfake_Et;
        bnp.wgt = bnp_wgt.real();
        bnp.wgt2 = bnp_wgt.imag();
        p.history_noise_bank.push_back(bnp);
      }
    }
    const double P_scatter = 1. - (microxs.absorption / microxs.total);
    {  // sample_vibration_noise_scatter :239-275
      ScatterInfo sinfo = sample_scatter(P, nuc, p.u(), microxs.energy_index, p.rng);
      BankedParticle p_noise{p.r(), sinfo.direction, sinfo.energy, 0., 0., p.history_id, p.daughter_counter(), p.family_id};
      std::complex<double> wgt{p.wgt(), p.wgt2()};
      wgt *= 1.;
      wgt *= P_scatter;
      wgt *= dN_N * Etfake_Et;
      p_noise.wgt = wgt.real();
      p_noise.wgt2 = wgt.imag();
      p.history_noise_bank.push_back(p_noise);
    }
  }

  if (!inside_osc) return;  // sample_oscillation_noise_source :293-323
  MicroXS microxs = mat.sample_nuclide(p.E(), p.rng, false);
  const Material& nuc = mat.mat();
  if (nuc.fissile) {  // sample_oscillation_noise_fission :383-445
    const double k_abs = microxs.nu_total * microxs.fission / microxs.total;
    const int n_new = static_cast<int>(std::floor(k_abs / keff + rng_rand(p.rng)));
    const double P_delayed = microxs.nu_delayed / microxs.nu_total;
    std::complex<double> dEf_Ef{0., 0.};
    for (const auto& ns : P.noise_sources)
      if (!ns.vibration && ns.is_inside(p.r())) dEf_Ef += sqosc_on(ns, w) ? std::complex<double>{ns.eps_f * PI, 0.} : std::complex<double>{0., 0.};
    for (int i = 0; i < n_new; i++) {
      auto finfo = sample_fission(P, nuc, p.u(), microxs.energy_index, P_delayed, p.rng);
      BankedParticle bnp{p.r(), finfo.direction, finfo.energy, p.wgt(), p.wgt2(), p.history_id, p.daughter_counter(), p.family_id};
      if (finfo.delayed) {
        std::complex<double> wgt_cmpx{bnp.wgt, bnp.wgt2};
        double lambda = finfo.lambda;
        double denom = (lambda * lambda) + (w * w);
        std::complex<double> mult{lambda * lambda / denom, -lambda * w / denom};
        wgt_cmpx *= mult;
        bnp.wgt = wgt_cmpx.real();
        bnp.wgt2 = wgt_cmpx.imag();
      }
      std::complex<double> fnp_weight{bnp.wgt, bnp.wgt2};
      fnp_weight *= dEf_Ef;
      bnp.wgt = fnp_weight.real();
      bnp.wgt2 = fnp_weight.imag();
      p.history_noise_bank.push_back(bnp);
    }
  }
  const double P_scatter = 1. - (microxs.absorption / microxs.total);
  {  // sample_oscillation_noise_scatter :325-381 (MG: mt == 2, yield == 1)
    ScatterInfo sinfo = sample_scatter(P, nuc, p.u(), microxs.energy_index, p.rng);
    BankedParticle p_noise{p.r(), sinfo.direction, sinfo.energy, 0., 0., p.history_id, p.daughter_counter(), p.family_id};
    std::complex<double> wgt{p.wgt(), p.wgt2()};
    wgt *= 1.;
    wgt *= P_scatter;
    std::complex<double> dE_E{0., 0.};
    for (const auto& ns : P.noise_sources)
      if (!ns.vibration && ns.is_inside(p.r())) dE_E += sqosc_on(ns, w) ? std::complex<double>{ns.eps_s * PI, 0.} : std::complex<double>{0., 0.};
    wgt *= dE_E;
    p_noise.wgt = wgt.real();
    p_noise.wgt2 = wgt.imag();
    p.history_noise_bank.push_back(p_noise);
  }
}

// ---- inter-generation steps --------------------------------------------------------------------
static void perform_exact_cancellation(Problem& P, std::vector<BankedParticle>& next_gen);
static void perform_regional_cancellation(Problem& P, std::vector<BankedParticle>& next_gen) {  // power_iterator.cpp:751-777
  const Cancelator& c = P.cancel;
  if (c.kind == 2 || c.kind == 3) return perform_exact_cancellation(P, next_gen);
  std::unordered_map<int, std::vector<BankedParticle*>> bins;
  for (auto& p : next_gen) {  // approximate_mesh_cancelator.cpp:97-145
    int i = static_cast<int>(std::floor((p.r.x - c.low.x) / c.dx));
    int j = static_cast<int>(std::floor((p.r.y - c.low.y) / c.dy));
    int k = static_cast<int>(std::floor((p.r.z - c.low.z) / c.dz));
    int l = -1;
    if (!c.energy_edges.empty()) {
      for (size_t e = 0; e < c.energy_edges.size() - 1; e++)
        if (c.energy_edges[e] <= p.E && p.E <= c.energy_edges[e + 1]) { l = (int)e; break; }
    } else {
      l = 0;
    }
    if (i < 0 || j < 0 || k < 0 || l < 0) continue;
    if (i >= (int)c.shape[0] || j >= (int)c.shape[1] || k >= (int)c.shape[2] || l >= (int)c.shape[3]) continue;
    int key = l + (int)c.shape[3] * (k + (int)c.shape[2] * (j + (int)c.shape[1] * i));
    bins[key].push_back(&p);
  }
  for (auto& kb : bins) {  // :147-190
    auto& bin = kb.second;
    if (bin.size() > 1) {
      bool pp1 = false, np1 = false, pp2 = false, np2 = false;
      double sum_wgt = 0., sum_wgt2 = 0.;
      for (const auto& p : bin) {
        if (p->wgt > 0.) pp1 = true; else if (p->wgt < 0.) np1 = true;
        sum_wgt += p->wgt;
        if (p->wgt2 > 0.) pp2 = true; else if (p->wgt2 < 0.) np2 = true;
        sum_wgt2 += p->wgt2;
      }
      double N = static_cast<double>(bin.size());
      double avg_wgt = sum_wgt / N, avg_wgt2 = sum_wgt2 / N;
      for (auto& p : bin) {
        if (pp1 && np1) p->wgt = avg_wgt;
        if (pp2 && np2) p->wgt2 = avg_wgt2;
      }
    }
  }
}

// ---- BasicExactMGCancelator (src/basic_exact_mg_cancelator.cpp) --------------------------------------------------------------
// Sobol points (vendor/sobol: Gruenschloss' 52-bit matrices of the Joe-Kuo direction numbers new-joe-kuo-6.21201).  The first three
// dimensions, generated from the published recurrence instead of the table: dimension 0 is the van der Corput sequence (m_i = 1),
// dimension 1 has the polynomial x + 1 with m_1 = 1, dimension 2 has x^2 + x + 1 with m = {1, 3}.  Checked against the
// reference's vendored table by tests/test_reference_pins.py.
struct SobolMatrices {  // (dimension 3: x^3 + x + 1 with m = {1, 3, 1}; ExactMGCancelator::sample_point draws the energy group from it)
  unsigned long long m[4][52];
  SobolMatrices() {
    unsigned long long d1[53], d2[53], d3[53];
    d1[1] = 1;
    for (int i = 2; i <= 52; i++) d1[i] = (2 * d1[i - 1]) ^ d1[i - 1];                       // s = 1, a = 0
    d2[1] = 1; d2[2] = 3;
    for (int i = 3; i <= 52; i++) d2[i] = (2 * d2[i - 1]) ^ (4 * d2[i - 2]) ^ d2[i - 2];     // s = 2, a_1 = 1
    d3[1] = 1; d3[2] = 3; d3[3] = 1;
    for (int i = 4; i <= 52; i++) d3[i] = (4 * d3[i - 2]) ^ (8 * d3[i - 3]) ^ d3[i - 3];     // s = 3, a_1 = 0, a_2 = 1
    for (int i = 1; i <= 52; i++) {
      m[0][i - 1] = 1ULL << (52 - i);
      m[1][i - 1] = d1[i] << (52 - i);
      m[2][i - 1] = d2[i] << (52 - i);
      m[3][i - 1] = d3[i] << (52 - i);
    }
  }
};
static double sobol_sample(unsigned long long index, unsigned dimension) {  // vendor/sobol/include/sobol/sobol.hpp:36-51, scramble = 0
  static const SobolMatrices M;
  unsigned long long result = 0;
  for (unsigned i = 0; index; index >>= 1, ++i)
    if (index & 1) result ^= M.m[dimension][i];
  return static_cast<double>(result) * (1.0 / (1ULL << 52));
}

struct ExactCancel {
  Problem& P;
  const Cancelator& c;
  static constexpr uint32_t N_MAX_POS = 100;
  explicit ExactCancel(Problem& p) : P(p), c(p.cancel) {}
  int get_material(const Vec& r) const {  // :178-196: geometry::get_cell(r, {1, 0, 0}); -1 where there is no cell
    Tracker t(&P.geo, r, Vec{1., 0., 0.});
    return t.is_lost() ? -1 : t.current_mat;
  }
  void ijk(int key, int& i, int& j, int& k) const {
    k = key % (int)c.shape[2];
    j = (key / (int)c.shape[2]) % (int)c.shape[1];
    i = key / ((int)c.shape[2] * (int)c.shape[1]);
  }
  bool add_particle(BankedParticle& p) {  // :68-108
    int i = static_cast<int>(std::floor((p.r.x - c.low.x) / c.dx));
    int j = static_cast<int>(std::floor((p.r.y - c.low.y) / c.dy));
    int k = static_cast<int>(std::floor((p.r.z - c.low.z) / c.dz));
    if (i < 0 || j < 0 || k < 0) return false;
    if (i >= (int)c.shape[0] || j >= (int)c.shape[1] || k >= (int)c.shape[2]) return false;
    const int key = k + (int)c.shape[2] * (j + (int)c.shape[1] * i);
    const int mat = get_material(p.r);
    if (P.exact_bins.find(key) == P.exact_bins.end()) P.exact_bins[key] = std::unordered_map<int, ExactCancelBin>();
    if (P.exact_bins[key].find(mat) == P.exact_bins[key].end()) P.exact_bins[key][mat] = ExactCancelBin();
    ExactCancelBin& b = P.exact_bins[key][mat];
    b.particles.push_back(&p);
    b.W += p.wgt;
    b.W2 += p.wgt2;
    return true;
  }
  template <class Engine>
  bool sample_position(int key, int mat, Engine& rng, Vec& out) const {  // :110-142
    int i, j, k;
    ijk(key, i, j, k);
    double Xl = c.low.x + i * c.dx, Yl = c.low.y + j * c.dy, Zl = c.low.z + k * c.dz;
    uint32_t N_TRIES = 0;
    bool position_sampled = false;
    while (N_TRIES < N_MAX_POS && !position_sampled) {
      double x = Xl + rng_rand(rng) * c.dx;
      double y = Yl + rng_rand(rng) * c.dy;
      double z = Zl + rng_rand(rng) * c.dz;
      out = Vec{x, y, z};
      if (get_material(out) == mat) position_sampled = true;
      N_TRIES++;
    }
    return position_sampled;
  }
  bool sample_position_sobol(int key, int mat, unsigned long long& idx, Vec& out) const {  // :144-176
    int i, j, k;
    ijk(key, i, j, k);
    double Xl = c.low.x + i * c.dx, Yl = c.low.y + j * c.dy, Zl = c.low.z + k * c.dz;
    uint32_t N_TRIES = 0;
    bool position_sampled = false;
    while (N_TRIES < N_MAX_POS && !position_sampled) {
      double x = Xl + sobol_sample(idx, 0) * c.dx;
      double y = Yl + sobol_sample(idx, 1) * c.dy;
      double z = Zl + sobol_sample(idx, 2) * c.dz;
      out = Vec{x, y, z};
      if (get_material(out) == mat) position_sampled = true;
      N_TRIES++;
      idx++;
    }
    return position_sampled;
  }
  static double get_f(const Vec& r, const Vec& r_parent, double Esmp) {  // :198-224 (std::pow(x, 2.) is x * x exactly)
    double d = std::sqrt((r.x - r_parent.x) * (r.x - r_parent.x) + (r.y - r_parent.y) * (r.y - r_parent.y) +
                         (r.z - r_parent.z) * (r.z - r_parent.z));
    return (1. / (d * d)) * g_math.exp(-Esmp * d);
  }
  double get_min_f(int key, const Vec& r_parent, double Esmp) const {  // :226-262
    int i, j, k;
    ijk(key, i, j, k);
    double Xl = c.low.x + i * c.dx, Xh = Xl + c.dx, Yl = c.low.y + j * c.dy, Yh = Yl + c.dy, Zl = c.low.z + k * c.dz, Zh = Zl + c.dz;
    double f1 = get_f({Xh, Yh, Zh}, r_parent, Esmp), f2 = get_f({Xh, Yh, Zl}, r_parent, Esmp), f3 = get_f({Xh, Yl, Zh}, r_parent, Esmp),
           f4 = get_f({Xh, Yl, Zl}, r_parent, Esmp), f5 = get_f({Xl, Yh, Zh}, r_parent, Esmp), f6 = get_f({Xl, Yh, Zl}, r_parent, Esmp),
           f7 = get_f({Xl, Yl, Zh}, r_parent, Esmp), f8 = get_f({Xl, Yl, Zl}, r_parent, Esmp);
    return std::min(std::min(std::min(f1, f2), std::min(f3, f4)), std::min(std::min(f5, f6), std::min(f7, f8)));
  }
  template <class Engine>
  void get_averages(int key, int mat, ExactCancelBin& bin, Engine* rng) const {  // :264-312 (rng) and :314-364 (sobol: rng == nullptr)
    bin.averages.resize(bin.particles.size());
    std::vector<Vec> r_smps;
    r_smps.reserve(c.nsamples);
    unsigned long long sobol_index = 0;
    for (size_t j = 0; j < c.nsamples; j++) {
      Vec r;
      const bool ok = rng ? sample_position(key, mat, *rng, r) : sample_position_sobol(key, mat, sobol_index, r);
      if (!ok) { bin.can_cancel = false; return; }
      r_smps.push_back(r);
    }
    for (size_t i = 0; i < bin.particles.size(); i++) {
      const Vec r_parent = bin.particles[i]->parents_previous_position;
      const double Esmp = bin.particles[i]->Esmp_parent;
      double sum_f = 0., sum_f_inv = 0.;
      for (const auto& r_smp : r_smps) {
        double f = get_f(r_smp, r_parent, Esmp);
        sum_f += f;
        sum_f_inv += 1. / f;
      }
      bin.averages[i].f = sum_f / static_cast<double>(c.nsamples);
      bin.averages[i].f_inv = sum_f_inv / static_cast<double>(c.nsamples);
    }
    if (c.beta == 3) {
      auto C = [](double f, double f_inv) { return 1. / (2. * f * f_inv - 1.); };
      double sum_c = 0.;
      for (size_t i = 0; i < bin.particles.size(); i++) sum_c += C(bin.averages[i].f, bin.averages[i].f_inv);
      double sum_c_wgt = 0., sum_c_wgt2 = 0.;
      for (size_t i = 0; i < bin.particles.size(); i++) {
        sum_c_wgt += C(bin.averages[i].f, bin.averages[i].f_inv) * bin.particles[i]->wgt;
        sum_c_wgt2 += C(bin.averages[i].f, bin.averages[i].f_inv) * bin.particles[i]->wgt2;
      }
      bin.sum_c = sum_c; bin.sum_c_wgt = sum_c_wgt; bin.sum_c_wgt2 = sum_c_wgt2;
    }
  }
  double get_beta(int key, const ExactCancelBin& bin, size_t i, const Vec& r_parent, double Esmp, double wgt, bool first_wgt) const {  // :366-406
    if (!bin.can_cancel) return 0.;
    switch (c.beta) {
      case 0: return 0.;
      case 1: return get_min_f(key, r_parent, Esmp);
      case 2: {
        double f = bin.averages[i].f;
        double N = static_cast<double>(bin.particles.size());
        double W = first_wgt ? bin.W : bin.W2;
        return f * (1. - ((W) / ((N + 1.) * wgt)));
      }
      default: {
        double sum_c_wgt = first_wgt ? bin.sum_c_wgt : bin.sum_c_wgt2;
        double S = sum_c_wgt / (1. + bin.sum_c);
        double f = bin.averages[i].f, f_inv = bin.averages[i].f_inv;
        return f * (1. / (2. * f * f_inv - 1.)) * (1. - (S / wgt));
      }
    }
  }
  void cancel_bin(int key, ExactCancelBin& bin, bool first_wgt) const {  // :408-455
    for (size_t i = 0; i < bin.particles.size(); i++) {
      const Vec r_parent = bin.particles[i]->parents_previous_position;
      const double Esmp = bin.particles[i]->Esmp_parent;
      double wgt = first_wgt ? bin.particles[i]->wgt : bin.particles[i]->wgt2;
      if (wgt == 0.) return;
      const double B = get_beta(key, bin, i, r_parent, Esmp, wgt, first_wgt);
      const double f = get_f(bin.particles[i]->r, r_parent, Esmp);
      const double P_p = (f - B) / f;
      const double P_u = B / f;
      if (std::isinf(P_u) || std::isinf(P_p) || std::isnan(P_u) || std::isnan(P_p)) return;
      if (first_wgt) {
        bin.uniform_wgt += bin.particles[i]->wgt * P_u;
        bin.particles[i]->wgt *= P_p;
      } else {
        bin.uniform_wgt2 += bin.particles[i]->wgt2 * P_u;
        bin.particles[i]->wgt2 *= P_p;
      }
    }
  }
  void perform_cancellation(Pcg32Stream& rng) {  // :457-555
    ExactBins& bins = P.exact_bins;
    if (c.beta == 0) return;
    if (bins.size() == 0) return;
    std::vector<std::pair<int, int>> keys;
    keys.reserve(bins.size());
    for (const auto& kb : bins)
      for (const auto& mb : kb.second) keys.push_back({kb.first, mb.first});
    uint64_t seed_advance = 0;
    if (c.beta != 1) {
      uint64_t max_rn_per_part = (uint64_t)c.nsamples * N_MAX_POS * (c.beta == 3 ? 2 : 1);
      bins[keys[0].first][keys[0].second].rng_seed_advance = 0;
      seed_advance = bins[keys[0].first][keys[0].second].particles.size() * max_rn_per_part;
      for (size_t i = 1; i < keys.size(); i++) {
        bins[keys[i].first][keys[i].second].rng_seed_advance = seed_advance;
        seed_advance += bins[keys[i].first][keys[i].second].particles.size() * max_rn_per_part;
      }
    }
    for (size_t i = 0; i < keys.size(); i++) {
      const int key = keys[i].first, mat = keys[i].second;
      ExactCancelBin& bin = bins[key][mat];
      Pcg32Stream rng_local = rng;
      rng_local.advance(bin.rng_seed_advance);
      if (bin.particles.size() > 1) {
        bool has_pos_w1 = false, has_neg_w1 = false, has_pos_w2 = false, has_neg_w2 = false;
        for (const auto& p : bin.particles) {
          if (p->wgt > 0.) has_pos_w1 = true; else if (p->wgt < 0.) has_neg_w1 = true;
          if (p->wgt2 > 0.) has_pos_w2 = true; else if (p->wgt2 < 0.) has_neg_w2 = true;
          if (has_pos_w1 && has_neg_w1 && has_pos_w2 && has_neg_w2) break;
        }
        if (((has_pos_w1 && has_neg_w1) || (has_pos_w2 && has_neg_w2)) && (c.beta == 2 || c.beta == 3))
          get_averages(key, mat, bin, c.sobol ? static_cast<Pcg32Stream*>(nullptr) : &rng_local);
        if (has_pos_w1 && has_neg_w1) cancel_bin(key, bin, true);
        if (has_pos_w2 && has_neg_w2) cancel_bin(key, bin, false);
        bin.particles.clear();
        bin.averages.clear();
        bin.sum_c = 0.; bin.sum_c_wgt = 0.; bin.sum_c_wgt2 = 0.;
      }
    }
    if (c.beta != 1) rng.advance(seed_advance);
  }
  std::vector<BankedParticle> get_new_particles(Pcg32Stream& rng) {  // :557-617
    if (c.beta == 0) return {};
    std::vector<BankedParticle> uniform_particles;
    for (auto& kb : P.exact_bins) {
      const int key = kb.first;
      for (auto& mb : kb.second) {
        const int mat = mb.first;
        ExactCancelBin& bin = mb.second;
        uint32_t N = static_cast<uint32_t>(std::ceil(std::max(std::abs(bin.uniform_wgt), std::abs(bin.uniform_wgt2))));
        if (N > 0) {
          double w = bin.uniform_wgt / N, w2 = bin.uniform_wgt2 / N;
          for (size_t i = 0; i < N; i++) {
            Vec r;
            if (!sample_position(key, mat, rng, r)) throw std::runtime_error("Couldn't sample position for uniform particle.");
            FissionInfo finfo = sample_fission(P, P.materials[(size_t)mat], Vec{0., 0., 1.}, 0, 0., rng);
            BankedParticle up{r, finfo.direction, finfo.energy, w, w2, 0, 0, 0};
            uniform_particles.push_back(up);
          }
        }
        bin.uniform_wgt = 0.;
        bin.uniform_wgt2 = 0.;
      }
    }
    return uniform_particles;
  }
};

// ---- ExactMGCancelator (src/exact_mg_cancelator.cpp; cancelator: {type: exact}) -----------------------------------------------
// As BasicExactMGCancelator with beta = average-g and Sobol points, for general multi-group physics: the density of a fission
// site also carries the parent's scattering-angle pdf towards the site and, with a chi matrix, the chi element; bins are keyed by
// mesh cell and energy bin.
struct ExactFullBin {
  struct Averages { double f = 0., f_inv = 0.; };
  double uniform_wgt = 0., uniform_wgt2 = 0., W = 0., W2 = 0., sum_c = 0., sum_c_wgt = 0., sum_c_wgt2 = 0.;
  bool can_cancel = true;
  std::vector<BankedParticle*> particles;
  std::vector<Averages> averages;
};
struct ExactFullKey { size_t i, j, k, e; };
struct ExactFullCancel {
  Problem& P;
  const Cancelator& c;
  // outer key: Key::hash_key() (exact_mg_cancelator.hpp:78-80), which std::hash<size_t> passes through; the key itself beside it
  std::unordered_map<size_t, std::unordered_map<int, ExactFullBin>>& bins;
  std::unordered_map<size_t, ExactFullKey>& keys;
  static constexpr uint32_t N_MAX_POS = 100;
  ExactFullCancel(Problem& p, std::unordered_map<size_t, std::unordered_map<int, ExactFullBin>>& b, std::unordered_map<size_t, ExactFullKey>& k)
      : P(p), c(p.cancel), bins(b), keys(k) {}
  static size_t group(const Settings& st, double E) {  // settings::group (src/settings.cpp:95-104): closed intervals, first hit
    if (st.energy_bounds.size() <= 1) return 0;
    for (size_t g = 0; g < st.energy_bounds.size() - 1; g++)
      if (st.energy_bounds[g] <= E && E <= st.energy_bounds[g + 1]) return g;
    return 0;
  }
  int get_material(const Vec& r) const {
    Tracker t(&P.geo, r, Vec{1., 0., 0.});
    return t.is_lost() ? -1 : t.current_mat;
  }
  size_t hash_key(const ExactFullKey& k) const { return k.e + c.shape[3] * (k.k + c.shape[2] * (k.j + c.shape[1] * k.i)); }
  bool get_key(const Vec& r, size_t g, ExactFullKey& key) const {  // :115-152
    if (r.x < c.low.x || r.x > c.hi.x || r.y < c.low.y || r.y > c.hi.y || r.z < c.low.z || r.z > c.hi.z) return false;
    key.i = static_cast<size_t>(std::floor((r.x - c.low.x) / c.dx));
    key.j = static_cast<size_t>(std::floor((r.y - c.low.y) / c.dy));
    key.k = static_cast<size_t>(std::floor((r.z - c.low.z) / c.dz));
    bool e_determined = false;
    size_t e = 0;
    for (e = 0; e < c.group_bins.size(); e++) {
      for (size_t q = 0; q < c.group_bins[e].size(); q++)
        if (g == c.group_bins[e][q]) { e_determined = true; break; }
      if (e_determined) break;
    }
    if (!e_determined && P.chi_matrix) return false;
    key.e = e;
    return true;
  }
  bool add_particle(BankedParticle& p) {  // :173-213 (USE_VIRTUAL_COLLISIONS is true: settings::use_virtual_collisions)
    ExactFullKey key{};
    if (!get_key(p.r, group(P.st, p.E), key)) return false;
    const size_t hk = hash_key(key);
    if (bins.find(hk) == bins.end()) { bins[hk] = std::unordered_map<int, ExactFullBin>(); keys[hk] = key; }
    const int mat = get_material(p.r);
    if (bins[hk].find(mat) == bins[hk].end()) bins[hk][mat] = ExactFullBin();
    ExactFullBin& b = bins[hk][mat];
    b.particles.push_back(&p);
    b.W += p.wgt;
    b.W2 += p.wgt2;
    return true;
  }
  double get_f(const Vec& r1, const Vec& u1, size_t g1, size_t g3, const Vec& r4, size_t g4, double Esmp, const Material& nuc) const {  // :215-235
    const double d = (r4 - r1).norm();
    const Vec u = make_direction(r4.x - r1.x, r4.y - r1.y, r4.z - r1.z);
    const double mu = u.dot(u1);
    const double pdf_mu = nuc.angle[g1][g3].pdf_at(mu);
    const double pdf_chi = P.chi_matrix ? nuc.chi[g3][g4] : 1.;
    return (pdf_mu * pdf_chi / (d * d)) * g_math.exp(-Esmp * d);
  }
  bool sample_point(const ExactFullKey& key, int mat, unsigned long long& idx, Vec& r, size_t& g) const {  // :249-294
    double Xl = c.low.x + static_cast<double>(key.i) * c.dx, Yl = c.low.y + static_cast<double>(key.j) * c.dy,
           Zl = c.low.z + static_cast<double>(key.k) * c.dz;
    uint32_t N_TRIES = 0;
    bool position_sampled = false;
    while (N_TRIES < N_MAX_POS && !position_sampled) {
      double x = Xl + sobol_sample(idx, 0) * c.dx;
      double y = Yl + sobol_sample(idx, 1) * c.dy;
      double z = Zl + sobol_sample(idx, 2) * c.dz;
      r = Vec{x, y, z};
      if (get_material(r) == mat) position_sampled = true;
      N_TRIES++;
      idx++;
    }
    if (!position_sampled) return false;
    g = 0;
    if (P.chi_matrix) {
      const double xi = sobol_sample(idx - 1, 3);
      const size_t g_index = static_cast<size_t>(std::floor(xi * static_cast<double>(c.group_bins[key.e].size())));
      g = c.group_bins[key.e][g_index];
    }
    return true;
  }
  void compute_averages(const ExactFullKey& key, int mat, ExactFullBin& bin) const {  // :296-364
    const Material& nuc = P.materials[(size_t)mat];
    bin.averages.resize(bin.particles.size());
    std::vector<std::pair<Vec, size_t>> smps;
    unsigned long long sobol_index = 0;
    for (size_t j = 0; j < c.nsamples; j++) {
      Vec r;
      size_t g;
      if (!sample_point(key, mat, sobol_index, r, g)) { bin.can_cancel = false; return; }
      smps.push_back({r, g});
    }
    for (size_t i = 0; i < bin.particles.size(); i++) {
      const BankedParticle& p = *bin.particles[i];
      const size_t g1 = group(P.st, p.parents_previous_previous_energy), g3 = group(P.st, p.parents_previous_energy);
      double sum_f = 0., sum_f_inv = 0.;
      for (const auto& sm : smps) {
        double f = get_f(p.parents_previous_position, p.parents_previous_direction, g1, g3, sm.first, sm.second, p.Esmp_parent, nuc);
        if (f == 0.) { bin.can_cancel = false; return; }
        sum_f += f;
        sum_f_inv += 1. / f;
      }
      bin.averages[i].f = sum_f / static_cast<double>(c.nsamples);
      bin.averages[i].f_inv = sum_f_inv / static_cast<double>(c.nsamples);
    }
    auto C = [](double f, double f_inv) { return 1. / (2. * f * f_inv - 1.); };
    double sum_c = 0.;
    for (size_t i = 0; i < bin.particles.size(); i++) sum_c += C(bin.averages[i].f, bin.averages[i].f_inv);
    double sum_c_wgt = 0., sum_c_wgt2 = 0.;
    for (size_t i = 0; i < bin.particles.size(); i++) {
      sum_c_wgt += C(bin.averages[i].f, bin.averages[i].f_inv) * bin.particles[i]->wgt;
      sum_c_wgt2 += C(bin.averages[i].f, bin.averages[i].f_inv) * bin.particles[i]->wgt2;
    }
    bin.sum_c = sum_c; bin.sum_c_wgt = sum_c_wgt; bin.sum_c_wgt2 = sum_c_wgt2;
  }
  double get_beta(const ExactFullBin& bin, size_t i, bool wgt_1) const {  // :237-247
    if (!bin.can_cancel) return 0.;
    const double wgt = wgt_1 ? bin.particles[i]->wgt : bin.particles[i]->wgt2;
    const double sum_c_wgt = wgt_1 ? bin.sum_c_wgt : bin.sum_c_wgt2;
    const double S = sum_c_wgt / (1. + bin.sum_c);
    const double f = bin.averages[i].f, f_inv = bin.averages[i].f_inv;
    return f * (1. / (2. * f * f_inv - 1.)) * (1. - (S / wgt));
  }
  void cancel_bin(ExactFullBin& bin, int mat, bool first_wgt) const {  // :366-401
    const Material& nuc = P.materials[(size_t)mat];
    for (size_t i = 0; i < bin.particles.size(); i++) {
      BankedParticle& p = *bin.particles[i];
      const size_t g1 = group(P.st, p.parents_previous_previous_energy), g3 = group(P.st, p.parents_previous_energy), g4 = group(P.st, p.E);
      const double B = get_beta(bin, i, first_wgt);
      const double f = get_f(p.parents_previous_position, p.parents_previous_direction, g1, g3, p.r, g4, p.Esmp_parent, nuc);
      const double P_p = (f - B) / f, P_u = B / f;
      if (std::isinf(P_u) || std::isinf(P_p) || std::isnan(P_u) || std::isnan(P_p)) return;
      if (first_wgt) { bin.uniform_wgt += p.wgt * P_u; p.wgt *= P_p; }
      else { bin.uniform_wgt2 += p.wgt2 * P_u; p.wgt2 *= P_p; }
    }
  }
  void perform_cancellation() {  // :403-476
    if (bins.size() == 0) return;
    for (auto& kb : bins)
      for (auto& mb : kb.second) {
        ExactFullBin& bin = mb.second;
        if (bin.particles.size() > 1) {
          bool p1 = false, n1 = false, p2 = false, n2 = false;
          for (const auto& p : bin.particles) {
            if (p->wgt > 0.) p1 = true; else if (p->wgt < 0.) n1 = true;
            if (p->wgt2 > 0.) p2 = true; else if (p->wgt2 < 0.) n2 = true;
            if (p1 && n1 && p2 && n2) break;
          }
          if ((p1 && n1) || (p2 && n2)) compute_averages(keys[kb.first], mb.first, bin);
          if (p1 && n1) cancel_bin(bin, mb.first, true);
          if (p2 && n2) cancel_bin(bin, mb.first, false);
          bin.particles.clear();
          bin.averages.clear();
          bin.sum_c = 0.; bin.sum_c_wgt = 0.; bin.sum_c_wgt2 = 0.;
        }
      }
  }
  std::vector<BankedParticle> get_new_particles(Pcg32Stream& rng) {  // :511-590
    std::vector<BankedParticle> uniform_particles;
    for (auto& kb : bins) {
      const ExactFullKey key = keys[kb.first];
      for (auto& mb : kb.second) {
        const int mat = mb.first;
        ExactFullBin& bin = mb.second;
        uint32_t N = static_cast<uint32_t>(std::ceil(std::max(std::abs(bin.uniform_wgt), std::abs(bin.uniform_wgt2))));
        if (N > 0) {
          double w = bin.uniform_wgt / N, w2 = bin.uniform_wgt2 / N;
          const Material& nuc = P.materials[(size_t)mat];
          double Xl = c.low.x + static_cast<double>(key.i) * c.dx, Yl = c.low.y + static_cast<double>(key.j) * c.dy,
                 Zl = c.low.z + static_cast<double>(key.k) * c.dz;
          for (size_t i = 0; i < N; i++) {
            Vec r{0, 0, 0};
            uint32_t N_TRIES = 0;
            bool position_sampled = false;
            while (N_TRIES < N_MAX_POS && !position_sampled) {  // sample_position :478-509
              double x = Xl + rng_rand(rng) * c.dx;
              double y = Yl + rng_rand(rng) * c.dy;
              double z = Zl + rng_rand(rng) * c.dz;
              r = Vec{x, y, z};
              if (get_material(r) == mat) position_sampled = true;
              N_TRIES++;
            }
            if (!position_sampled) throw std::runtime_error("Couldn't sample position for uniform particle.");
            size_t e_index = 0;
            if (P.chi_matrix) {
              const double xi_E = rng_rand(rng);
              size_t g_index = static_cast<size_t>(std::floor(xi_E * static_cast<double>(c.group_bins[key.e].size())));
              e_index = c.group_bins[key.e][g_index];
            } else {
              e_index = static_cast<size_t>(rng_discrete(rng, nuc.chi_cp[0]));
            }
            double E_smp = 0.5 * (P.st.energy_bounds[e_index] + P.st.energy_bounds[e_index + 1]);
            // Direction u_smp(2. * RNG::rand(rng) - 1., 2. * PI * RNG::rand(rng)) (:573): two draws inside one argument list, whose
            // order the language leaves open -- g++ evaluates the arguments from the right, so phi takes the first draw
            double phi = 2. * PI * rng_rand(rng);
            double mu = 2. * rng_rand(rng) - 1.;
            if (mu < -1.) mu = -1.; else if (mu > 1.) mu = 1.;
            if (phi < 0.) phi = 0.; else if (phi > 2 * PI) phi = 2 * PI;
            const Vec u_smp = make_direction(std::sqrt(1. - mu * mu) * g_math.cos(phi), std::sqrt(1. - mu * mu) * g_math.sin(phi), mu);
            BankedParticle up{r, u_smp, E_smp, w, w2, 0, 0, 0};  // (the reference leaves the three ids uninitialised)
            uniform_particles.push_back(up);
          }
        }
        bin.uniform_wgt = 0.;
        bin.uniform_wgt2 = 0.;
      }
    }
    return uniform_particles;
  }
};
static std::unordered_map<Problem*, std::pair<std::unordered_map<size_t, std::unordered_map<int, ExactFullBin>>, std::unordered_map<size_t, ExactFullKey>>> g_exact_full;

// PowerIterator::perform_regional_cancellation with an exact cancelator (power_iterator.cpp:751-777): the uniform particles are
// appended to the bank
static void perform_exact_cancellation(Problem& P, std::vector<BankedParticle>& next_gen) {
  if (P.cancel.kind == 3) {
    auto& st = g_exact_full[&P];
    ExactFullCancel ec(P, st.first, st.second);
    for (auto& p : next_gen) (void)ec.add_particle(p);
    ec.perform_cancellation();
    auto tmp = ec.get_new_particles(P.global_rng);
    next_gen.insert(next_gen.end(), tmp.begin(), tmp.end());
    st.first.clear();   // cancelator->clear(): the bucket array survives
    st.second.clear();
    return;
  }
  ExactCancel ec(P);
  for (auto& p : next_gen) (void)ec.add_particle(p);
  ec.perform_cancellation(P.global_rng);
  auto tmp = ec.get_new_particles(P.global_rng);
  next_gen.insert(next_gen.end(), tmp.begin(), tmp.end());
  P.exact_bins.clear();
}

struct GenStats { int Npos = 0, Nneg = 0, Ntot = 0, Nnet = 0; double Wpos = 0, Wneg = 0; };
static GenStats normalize_weights(Problem& P, std::vector<BankedParticle>& next_gen) {  // power_iterator.cpp:538-586
  GenStats s;
  double W = 0., W_neg = 0., W_pos = 0.;
  for (size_t i = 0; i < next_gen.size(); i++) {
    if (next_gen[i].wgt > 0.) { W_pos += next_gen[i].wgt; s.Npos++; }
    else { W_neg -= next_gen[i].wgt; s.Nneg++; }
  }
  W = W_pos - W_neg;
  s.Ntot = s.Npos + s.Nneg;
  s.Nnet = s.Npos - s.Nneg;
  double w_per_part = static_cast<double>(P.st.nparticles) / W;
  W_neg *= w_per_part;
  W_pos *= w_per_part;
  for (size_t i = 0; i < next_gen.size(); i++) next_gen[i].wgt *= w_per_part;
  s.Wpos = W_pos;
  s.Wneg = W_neg;
  return s;
}

// BranchlessPowerIterator::comb_particles (src/branchless_power_iterator.cpp:592-651), as written: the negative comb divides by
// Npos and copies POSITIVE particles (it only runs when negative weights exist).  std::shuffle is libstdc++'s, on the global engine.
static void comb_particles(Problem& P, std::vector<BankedParticle>& next_gen) {
  std::vector<BankedParticle> positive_particles, negative_particles;
  positive_particles.reserve(next_gen.size());
  negative_particles.reserve(next_gen.size() / 3);
  double Wpos = 0., Wneg = 0.;
  for (size_t i = 0; i < next_gen.size(); i++) {
    if (next_gen[i].wgt > 0.) { Wpos += next_gen[i].wgt; positive_particles.push_back(next_gen[i]); }
    else { Wneg += next_gen[i].wgt; negative_particles.push_back(next_gen[i]); }
  }
  next_gen.clear();
  size_t Npos = static_cast<size_t>(std::ceil(Wpos));
  size_t Nneg = static_cast<size_t>(std::ceil(std::abs(Wneg)));
  next_gen.reserve(Npos + Nneg);
  std::shuffle(positive_particles.begin(), positive_particles.end(), P.global_rng);
  double avg_pos_wgt = Wpos / static_cast<double>(Npos);
  double comb_pos = rng_rand(P.global_rng) * avg_pos_wgt;
  double current_particle = 0.;
  for (size_t i = 0; i < positive_particles.size(); i++) {
    current_particle += positive_particles[i].wgt;
    while (comb_pos < current_particle) {
      next_gen.push_back(positive_particles[i]);
      next_gen.back().wgt = avg_pos_wgt;
      comb_pos += avg_pos_wgt;
    }
  }
  std::shuffle(negative_particles.begin(), negative_particles.end(), P.global_rng);
  double avg_neg_wgt = std::abs(Wneg) / static_cast<double>(Npos);
  comb_pos = rng_rand(P.global_rng) * avg_neg_wgt;
  current_particle = 0.;
  for (size_t i = 0; i < negative_particles.size(); i++) {
    current_particle -= negative_particles[i].wgt;
    while (comb_pos < current_particle) {
      if (i >= positive_particles.size()) throw std::runtime_error("comb_particles: the reference reads past its positive buffer here");
      next_gen.push_back(positive_particles[i]);
      next_gen.back().wgt = -avg_neg_wgt;
      comb_pos += avg_neg_wgt;
    }
  }
  std::shuffle(next_gen.begin(), next_gen.end(), P.global_rng);
}

// Simulation::sample_sources + Source::generate_particle
static std::vector<Particle> sample_sources(Problem& P, size_t N) {
  std::vector<double> wgts;
  for (auto& s : P.sources) wgts.push_back(s.weight);
  auto src_cp = discrete_table(wgts.data(), wgts.size());
  std::vector<Particle> out;
  out.reserve(N);
  for (size_t i = 0; i < N; i++) {
    uint64_t history_id = P.histories_counter++;
    Pcg32 rng;
    rng.seed(P.st.rng_seed);
    rng.advance(P.st.rng_stride * history_id);
    size_t indx = (size_t)rng_discrete(rng, src_cp);
    const Source& S = P.sources[indx];
    Vec u = S.dir;  // mono-directional: the stored direction, no draw
    if (S.dir_kind == 0) {  // src/isotropic.cpp:28-36
      double mu = 2. * rng_rand(rng) - 1.;
      double phi = 2. * PI * rng_rand(rng);
      u = make_direction_mu_phi(mu, phi);
    } else if (S.dir_kind == 2) {  // src/cone.cpp:34-42
      double mu = (1. - S.cos_aperture) * rng_rand(rng) + S.cos_aperture;
      double phi = 2. * PI * rng_rand(rng);
      u = rotate_direction(S.dir, mu, phi);
    }
    double E = S.energy;  // mono-energetic: sampled twice, no draws (source.cpp:49-58)
    if (S.en_kind != 0) {
      auto sample_energy = [&]() {
        const double xi1 = rng_rand(rng), xi2 = rng_rand(rng), xi3 = rng_rand(rng);  // maxwellian.cpp:34-42
        const double c = g_math.cos(PI * xi3 / 2.);
        const double w = -S.en_a * (g_math.log(xi1) + g_math.log(xi2) * c * c);
        if (S.en_kind == 1) return w;
        return w + 0.25 * S.en_a * S.en_a * S.en_b + (2. * rng_rand(rng) - 1.) * std::sqrt(S.en_a * S.en_a * S.en_b * w);  // watt.cpp:42-47
      };
      E = sample_energy();
      int E_count = 0;
      do {  // source.cpp:50-58
        if (E_count > 200) throw std::runtime_error("Exceded 200 samplings of energy.");
        E = sample_energy();
        E_count++;
      } while (E <= P.st.min_energy || P.st.max_energy <= E);
    }
    auto sample_pos = [&]() -> Vec {
      if (!S.is_box) return S.low;
      double x = (S.hi.x - S.low.x) * rng_rand(rng) + S.low.x;  // box.cpp:37-42
      double y = (S.hi.y - S.low.y) * rng_rand(rng) + S.low.y;
      double z = (S.hi.z - S.low.z) * rng_rand(rng) + S.low.z;
      return {x, y, z};
    };
    Vec r = sample_pos();
    Tracker trkr(&P.geo, r, u);
    int guard = 0;
    while (trkr.is_lost()) {
      if (!S.is_box && ++guard > 1) throw std::runtime_error("point source outside geometry");
      r = sample_pos();
      trkr.set_r(r);
      trkr.restart_get_current();
    }
    if (S.fissile_only) {
      int count = 0;
      while (trkr.is_lost() || !P.materials[(size_t)trkr.current_mat].fissile) {
        if (count == 201) throw std::runtime_error("Exceded 200 samplings of position in fissile-only source.");
        r = sample_pos();
        trkr.set_r(r);
        trkr.restart_get_current();
        count++;
      }
    }
    Particle p(r, u, E, 1.0, history_id);
    p.rng = rng;
    p.rng.ndraw = 0;
    out.push_back(std::move(p));
  }
  return out;
}

}  // namespace orc

// =====================================================================================================
// C API (ctypes) -- test infrastructure only
// =====================================================================================================
using namespace orc;

extern "C" {

struct orc_bank {  // SoA view of a particle / fission bank, all arrays length n (caller-owned)
  uint64_t n;
  double *x, *y, *z, *ux, *uy, *uz, *E, *wgt, *wgt2;
  uint64_t *id_a;  // in: history id      out: parent_history_id
  uint64_t *id_b;  // in: family id       out: parent_daughter_id
  uint64_t *id_c;  // in: rng state (0-ptr => seed/stride/history id)   out: family id
};

const char* orc_last_error(void* h) { return static_cast<Problem*>(h)->error.c_str(); }

static int g_math_mode = 0;
int orc_get_math() { return g_math_mode; }
void orc_set_math(int mode) {
  g_math_mode = mode == 0 ? 0 : 1;
  if (mode == 0) g_math = {libm_log, libm_sin, libm_cos, libm_exp, libm_acos};
  else g_math = {orc_log, orc_sin, orc_cos, orc_exp, orc_acos};
}
void orc_set_threads(int n) { omp_set_num_threads(n); }
int orc_max_threads() { return omp_get_max_threads(); }

void* orc_load(const char* path, char* errbuf, int errlen) {
  try {
    return load_problem(path);
  } catch (const std::exception& e) {
    if (errbuf && errlen > 0) { std::strncpy(errbuf, e.what(), (size_t)errlen - 1); errbuf[errlen - 1] = 0; }
    return nullptr;
  }
}

void orc_set_nparticles(void* h, int n) {
  Problem& P = *static_cast<Problem*>(h);
  P.st.nparticles = n;
  P.tallies.total_weight = static_cast<double>(n);
  for (auto& t : P.tallies.mesh) t.net_weight = P.tallies.total_weight;
}
void orc_set_converged(void* h, int c) { static_cast<Problem*>(h)->converged = c != 0; }
void orc_set_kcol(void* h, double k) { static_cast<Problem*>(h)->tallies.k_col = k; }
void orc_set_trace(void* h, int t) { static_cast<Problem*>(h)->want_trace = t != 0; }
void orc_reset_counters(void* h) { static_cast<Problem*>(h)->counters = Counters(); }
void orc_get_counters(void* h, uint64_t* out8) {
  const Counters& c = static_cast<Problem*>(h)->counters;
  out8[0] = c.flights; out8[1] = c.real_collisions; out8[2] = c.virtual_collisions; out8[3] = c.tl_bins;
  out8[4] = c.fission_sites; out8[5] = c.boundary_events; out8[6] = c.lost_at_birth; out8[7] = c.coll_scores;
}
void orc_get_majorant(void* h, double* maj, double* smp) {
  Problem& P = *static_cast<Problem*>(h);
  for (size_t g = 0; g < P.st.ngroups; g++) { maj[g] = P.majorant[g]; smp[g] = P.sampling[g]; }
}
int orc_ngroups(void* h) { return (int)static_cast<Problem*>(h)->st.ngroups; }
int orc_nparticles(void* h) { return static_cast<Problem*>(h)->st.nparticles; }

// RNG known-answer helpers
void orc_rng_stream(uint64_t seed, uint64_t stride, uint64_t history_id, int n, uint32_t* out_u32) {
  Pcg32 g; g.seed(seed); g.advance(stride * history_id);
  for (int i = 0; i < n; i++) out_u32[i] = g.next();
}
// ---- probes of the pieces that tests/test_reference_pins.py compares with the reference's own code (oracle/ref_probe.cpp)
int orc_surface_probe(int type, const double* params, int n, const double* r3, const double* u3, const int* on_surf, int* sign,
                      double* dist, double* norm3) {
  Surface s;
  s.type = type;
  for (int k = 0; k < 7; k++) s.p[k] = params[k];
  if (type == S_CYL) {  // Cylinder(x0,y0,z0,u0,v0,w0,R): axis normalised into alpha, beta, gamma (cylinder.cpp:28-58)
    s.p[6] = params[6];
    s.finish_general_cylinder(params[3], params[4], params[5]);
  }
  for (int i = 0; i < n; i++) {
    const Vec r{r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]};
    const Vec u = make_direction(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]);
    sign[i] = s.sign(r, u);
    dist[i] = s.distance(r, u, on_surf[i] != 0);
    const Vec nn = s.norm(r);
    norm3[3 * i] = nn.x; norm3[3 * i + 1] = nn.y; norm3[3 * i + 2] = nn.z;
  }
  return 0;
}
void orc_direction_probe(int n, const double* xyz, double* out) {
  for (int i = 0; i < n; i++) {
    const Vec d = make_direction(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
  }
}
void orc_rotate_direction_probe(int n, const double* u3, const double* mu, const double* phi, double* out) {
  for (int i = 0; i < n; i++) {
    const Vec u = make_direction(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]);
    const Vec d = rotate_direction(u, mu[i], phi[i]);
    out[3 * i] = d.x; out[3 * i + 1] = d.y; out[3 * i + 2] = d.z;
  }
}
void orc_sample_mu_probe(const double* mu, const double* pdf, const double* cdf, int npts, uint64_t seed, uint64_t stride,
                         uint64_t id, int n, double* out) {
  AngleDist d;
  d.mu.assign(mu, mu + npts); d.pdf.assign(pdf, pdf + npts); d.cdf.assign(cdf, cdf + npts);
  Pcg32 rng;
  rng.seed(seed);
  rng.advance(stride * id);
  for (int i = 0; i < n; i++) out[i] = d.sample_mu(rng);
}
int orc_legendre_linearize_probe(const double* a, int na, int cap, double* mu, double* pdf, double* cdf) {
  Legendre L;  // LegendreDistribution(a): a[l-1] is the l-th moment, the 0th (= 1) is implied (legendre_distribution.cpp:41-60)
  for (int l = 1; l <= na; l++) L.set_moment((size_t)l, a[l - 1]);
  const AngleDist d = L.linearize();
  const int n = (int)d.mu.size();
  if (n > cap) return -n;
  std::memcpy(mu, d.mu.data(), n * sizeof(double));
  std::memcpy(pdf, d.pdf.data(), n * sizeof(double));
  std::memcpy(cdf, d.cdf.data(), n * sizeof(double));
  return n;
}

void orc_rng_rand(uint64_t seed, uint64_t stride, uint64_t history_id, int n, double* out) {
  Pcg32 g; g.seed(seed); g.advance(stride * history_id);
  for (int i = 0; i < n; i++) out[i] = rng_rand(g);
}
double orc_rng_exponential(uint64_t seed, uint64_t stride, uint64_t history_id, double lambda) {
  Pcg32 g; g.seed(seed); g.advance(stride * history_id);
  return rng_exponential(g, lambda);
}
int orc_rng_discrete(uint64_t seed, uint64_t stride, uint64_t history_id, const double* w, int nw, int ndraws, int* out) {
  Pcg32 g; g.seed(seed); g.advance(stride * history_id);
  auto cp = discrete_table(w, (size_t)nw);
  for (int i = 0; i < ndraws; i++) out[i] = rng_discrete(g, cp);
  return (int)g.ndraw;
}
// the counterpart of ref_mg_nuclide (oracle/ref_probe.cpp): same arrays in, same draws out
int orc_mg_nuclide_probe(int G, const double* ebounds, const double* Et, const double* Ea, const double* Ef, const double* nu_p,
                         const double* nu_d, const double* chi, const double* Es, int nleg, const double* leg, int ndg,
                         const double* Pd, const double* lam, uint64_t seed, uint64_t stride, int nhist, int ndraw,
                         double* micro, double* scat, double* fis) {
  const size_t g = (size_t)G;
  Problem P;
  P.st.ngroups = (uint32_t)G;
  P.st.energy_bounds.assign(ebounds, ebounds + G + 1);
  Material m;
  m.G = g;
  m.Et.assign(Et, Et + g); m.Ea.assign(Ea, Ea + g); m.Ef.assign(Ef, Ef + g); m.nu_p.assign(nu_p, nu_p + g);
  m.has_nu_d = nu_d != nullptr;
  if (nu_d) m.nu_d.assign(nu_d, nu_d + g);
  m.speeds.assign(g, 1.);
  m.chi.resize(g); m.Ps.resize(g);
  for (size_t i = 0; i < g; i++) { m.chi[i].assign(chi + i * g, chi + (i + 1) * g); m.Ps[i].assign(Es + i * g, Es + (i + 1) * g); }
  std::vector<std::vector<Legendre>> legendre(g, std::vector<Legendre>(g));
  for (int l = 1; l <= nleg; l++)
    for (size_t i = 0; i < g; i++)
      for (size_t o = 0; o < g; o++) legendre[i][o].set_moment((size_t)l, leg[((size_t)(l - 1) * g + i) * g + o]);
  m.angle.assign(g, std::vector<AngleDist>(g));
  for (size_t i = 0; i < g; i++)
    for (size_t o = 0; o < g; o++) m.angle[i][o] = legendre[i][o].linearize();
  m.P_delayed_group.assign(Pd, Pd + ndg);
  m.decay_constants.assign(lam, lam + ndg);
  m.finish();
  for (size_t i = 0; i < g; i++) {
    if (P.st.group(0.5 * (ebounds[i] + ebounds[i + 1])) != i) return 2;
    const MicroXS xs = m.micro(i);
    double* q = micro + 6 * i;
    q[0] = xs.total; q[1] = xs.fission; q[2] = xs.absorption; q[3] = xs.elastic; q[4] = xs.nu_total; q[5] = xs.nu_delayed;
  }
  for (int h = 0; h < nhist; h++) {
    Pcg32 rng; rng.seed(seed); rng.advance(stride * (uint64_t)h);
    Vec u = make_direction(0., 0., 1.);
    for (int d = 0; d < ndraw; d++) {
      const size_t gi = (size_t)(h + d) % g;
      const MicroXS xs = m.micro(gi);
      const ScatterInfo si = sample_scatter(P, m, u, gi, rng);
      double* s = scat + 4 * ((size_t)h * ndraw + d);
      s[0] = si.energy; s[1] = si.direction.x; s[2] = si.direction.y; s[3] = si.direction.z;
      u = si.direction;
      const double Pdelayed = xs.nu_total > 0. ? xs.nu_delayed / xs.nu_total : 0.;
      const FissionInfo fi = sample_fission(P, m, u, gi, Pdelayed, rng);
      double* f = fis + 6 * ((size_t)h * ndraw + d);
      f[0] = fi.energy; f[1] = fi.direction.x; f[2] = fi.direction.y; f[3] = fi.direction.z;
      f[4] = fi.delayed ? 1. : 0.; f[5] = fi.lambda;
    }
  }
  return 0;
}
// ndraws x RNG::discrete, then one RNG::rand: the value shows how many engine steps the draws consumed
double orc_rng_discrete_probe(uint64_t seed, uint64_t stride, uint64_t history_id, const double* w, int nw, int ndraws, int* out) {
  Pcg32 g; g.seed(seed); g.advance(stride * history_id);
  auto cp = discrete_table(w, (size_t)nw);
  for (int i = 0; i < ndraws; i++) out[i] = rng_discrete(g, cp);
  return rng_rand(g);
}
void orc_acos_eval(int n, const double* x, double* out) {
  for (int i = 0; i < n; i++) out[i] = g_math.acos(x[i]);
}
void orc_math_eval(int n, const double* x, double* lg, double* sn, double* cs) {
  for (int i = 0; i < n; i++) { lg[i] = g_math.log(x[i]); sn[i] = g_math.sin(x[i]); cs[i] = g_math.cos(x[i]); }
}

// geometry probe: cell index + material index for points (restart lookups), -1 when lost
void orc_find_cells(void* h, int n, const double* r3, const double* u3, int32_t* cell, int32_t* mat) {
  Problem& P = *static_cast<Problem*>(h);
  for (int i = 0; i < n; i++) {
    Tracker t(&P.geo, Vec{r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]}, Vec{u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]});
    cell[i] = t.current_cell;
    mat[i] = t.current_mat;
  }
}

// counterparts of ref_geometry_walk_surface / ref_geometry_walk_delta (oracle/ref_probe.cpp): same rays in, same rows out
static void walk_put(double* o, const Geometry& geo, const Tracker& t) {
  o[0] = t.is_lost() ? -1. : static_cast<double>(geo.cells[(size_t)t.current_cell].id);
  o[1] = t.is_lost() ? -1. : static_cast<double>(t.current_mat);
  o[2] = 0.;
}
int orc_geometry_walk_surface(void* h, int n, const double* r3, const double* u3, int nsteps, double* out) {
  Problem& P = *static_cast<Problem*>(h);
  try {
    for (int i = 0; i < n; i++) {
      double* o = out + static_cast<size_t>(i) * (nsteps + 1) * 8;
      Particle p(Vec{r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]}, make_direction(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]), 1., 1.);
      Tracker trkr(&P.geo, p.r(), p.u());
      walk_put(o, P.geo, trkr); o[7] = 1.;
      for (int s = 1; s <= nsteps && !trkr.is_lost(); s++) {
        o += 8;
        const Boundary b = trkr.get_nearest_boundary();
        o[3] = b.distance; o[4] = b.surface_index; o[5] = b.boundary_type; o[6] = b.token; o[7] = 1.;
        if (b.boundary_type == BC_VACUUM) { o[0] = o[1] = -2.; break; }
        if (b.boundary_type == BC_REFLECTIVE) {
          do_reflection(trkr, p, b);
        } else {
          trkr.cross_surface(b);
          trkr.get_current();
          p.move(b.distance);
        }
        walk_put(o, P.geo, trkr);
      }
    }
    return 0;
  } catch (const std::exception& e) { P.error = e.what(); return 1; }
}
int orc_geometry_walk_delta(void* h, int n, const double* r3, const double* u3, int nsteps, const double* d, const double* unew3,
                            double* out) {
  Problem& P = *static_cast<Problem*>(h);
  try {
    for (int i = 0; i < n; i++) {
      double* o = out + static_cast<size_t>(i) * (nsteps + 1) * 8;
      Particle p(Vec{r3[3 * i], r3[3 * i + 1], r3[3 * i + 2]}, make_direction(u3[3 * i], u3[3 * i + 1], u3[3 * i + 2]), 1., 1.);
      Tracker trkr(&P.geo, p.r(), p.u());
      walk_put(o, P.geo, trkr); o[7] = 1.;
      for (int s = 1; s <= nsteps && !trkr.is_lost(); s++) {
        o += 8;
        const size_t k = static_cast<size_t>(i) * nsteps + (s - 1);
        Boundary b(INF, -1, BC_NORMAL);
        bool crossed_boundary = false;
        trkr.move(d[k]);
        trkr.get_current();
        if (trkr.is_lost()) {
          trkr.set_r(p.r());
          trkr.get_current();
          b = trkr.get_boundary_condition();
          crossed_boundary = true;
        }
        o[3] = b.distance; o[4] = b.surface_index; o[5] = b.boundary_type; o[6] = b.token; o[7] = 1.;
        if (crossed_boundary) {
          if (b.boundary_type == BC_VACUUM) { o[0] = o[1] = -2.; break; }
          if (b.boundary_type != BC_REFLECTIVE) throw std::runtime_error("Help me, how did I get here ?");
          do_reflection(trkr, p, b);
        } else {
          p.move(d[k]);
          if (s % 2 == 0) {
            p.state.direction = make_direction(unew3[3 * k], unew3[3 * k + 1], unew3[3 * k + 2]);
            trkr.set_u(p.u());
          }
        }
        walk_put(o, P.geo, trkr);
      }
    }
    return 0;
  } catch (const std::exception& e) { P.error = e.what(); return 1; }
}

// Source sampling: n particles with history ids continuing P.histories_counter
int orc_sample_source(void* h, orc_bank* out) {
  Problem& P = *static_cast<Problem*>(h);
  try {
    auto v = sample_sources(P, (size_t)out->n);
    P.global_histories_counter = P.histories_counter;
    for (size_t i = 0; i < v.size(); i++) {
      out->x[i] = v[i].r().x; out->y[i] = v[i].r().y; out->z[i] = v[i].r().z;
      out->ux[i] = v[i].u().x; out->uy[i] = v[i].u().y; out->uz[i] = v[i].u().z;
      out->E[i] = v[i].E(); out->wgt[i] = v[i].wgt(); out->wgt2[i] = v[i].wgt2();
      out->id_a[i] = v[i].history_id; out->id_b[i] = v[i].history_id; out->id_c[i] = v[i].rng.state;
    }
    return 0;
  } catch (const std::exception& e) { P.error = e.what(); return 1; }
}
void orc_set_history_counter(void* h, uint64_t c) {
  Problem& P = *static_cast<Problem*>(h);
  P.histories_counter = c;
  P.global_histories_counter = c;
}

static std::vector<Particle> bank_from(const Problem& P, const orc_bank* in) {
  std::vector<Particle> bank;
  bank.reserve(in->n);
  for (uint64_t i = 0; i < in->n; i++) {
    Particle p(Vec{in->x[i], in->y[i], in->z[i]}, Vec{in->ux[i], in->uy[i], in->uz[i]}, in->E[i], in->wgt[i], in->id_a[i]);
    p.state.weight2 = in->wgt2 ? in->wgt2[i] : 0.;
    p.family_id = in->id_b ? in->id_b[i] : in->id_a[i];
    if (in->id_c) { p.rng.state = in->id_c[i]; p.rng.ndraw = 0; }
    else p.initialize_rng(P.st.rng_seed, P.st.rng_stride);
    bank.push_back(std::move(p));
  }
  return bank;
}
static void bank_to(const std::vector<BankedParticle>& v, orc_bank* out) {
  for (size_t i = 0; i < v.size() && i < out->n; i++) {
    out->x[i] = v[i].r.x; out->y[i] = v[i].r.y; out->z[i] = v[i].r.z;
    out->ux[i] = v[i].u.x; out->uy[i] = v[i].u.y; out->uz[i] = v[i].u.z;
    out->E[i] = v[i].E; out->wgt[i] = v[i].wgt; out->wgt2[i] = v[i].wgt2;
    out->id_a[i] = v[i].parent_history_id; out->id_b[i] = v[i].parent_daughter_id; out->id_c[i] = v[i].family_id;
  }
}

// One Transporter::transport call.  out->n is the capacity on entry; *n_out the true count.
// scores6 = k_col, k_abs, k_trk, k_tot, leak, mig (raw sums added by this call)
int orc_transport(void* h, const orc_bank* in, int noise, orc_bank* out, uint64_t* n_out, double* scores6) {
  Problem& P = *static_cast<Problem*>(h);
  try {
    auto bank = bank_from(P, in);
    Tallies& T = P.tallies;
    double b[6] = {T.k_col_score, T.k_abs_score, T.k_trk_score, T.k_tot_score, T.leak_score, T.mig_area_score};
    auto fis = transport(P, bank, noise != 0, nullptr, false);
    scores6[0] = T.k_col_score - b[0]; scores6[1] = T.k_abs_score - b[1]; scores6[2] = T.k_trk_score - b[2];
    scores6[3] = T.k_tot_score - b[3]; scores6[4] = T.leak_score - b[4]; scores6[5] = T.mig_area_score - b[5];
    *n_out = fis.size();
    bank_to(fis, out);
    P.last_parent_info.resize(4 * fis.size());
    P.last_parent_state.resize(6 * fis.size());
    for (size_t i = 0; i < fis.size(); i++) {
      P.last_parent_info[4 * i] = fis[i].parents_previous_position.x; P.last_parent_info[4 * i + 1] = fis[i].parents_previous_position.y;
      P.last_parent_info[4 * i + 2] = fis[i].parents_previous_position.z; P.last_parent_info[4 * i + 3] = fis[i].Esmp_parent;
      P.last_parent_state[6 * i] = fis[i].parents_previous_direction.x; P.last_parent_state[6 * i + 1] = fis[i].parents_previous_direction.y;
      P.last_parent_state[6 * i + 2] = fis[i].parents_previous_direction.z; P.last_parent_state[6 * i + 3] = fis[i].parents_previous_previous_energy;
      P.last_parent_state[6 * i + 4] = fis[i].parents_previous_energy; P.last_parent_state[6 * i + 5] = fis[i].parents_previous_was_virtual ? 1. : 0.;
    }
    return 0;
  } catch (const std::exception& e) { P.error = e.what(); return 1; }
}

// BankedParticle::parents_previous_position / Esmp_parent of the fission bank the last orc_transport returned ([n][4])
uint64_t orc_last_parent_info(void* h, double* out4n, uint64_t n) {
  Problem& P = *static_cast<Problem*>(h);
  const uint64_t m = std::min<uint64_t>(n, P.last_parent_info.size() / 4);
  for (uint64_t i = 0; i < 4 * m; i++) out4n[i] = P.last_parent_info[i];
  return m;
}

// ... parents_previous_direction, parents_previous_previous_energy, parents_previous_energy, parents_previous_was_virtual ([n][6])
uint64_t orc_last_parent_state(void* h, double* out6n, uint64_t n) {
  Problem& P = *static_cast<Problem*>(h);
  const uint64_t m = std::min<uint64_t>(n, P.last_parent_state.size() / 6);
  for (uint64_t i = 0; i < 6 * m; i++) out6n[i] = P.last_parent_state[i];
  return m;
}

// PowerIterator::perform_regional_cancellation with the exact cancelator on a bank + its parent info ([n][4]); rng2 = {state,
// increment} of settings::rng.  out->n = capacity on entry; the uniform particles follow the n input rows.
int orc_cancel_exact_state(void* h, const orc_bank* in, const double* parent4n, const double* state6n, orc_bank* out, uint64_t* n_out, uint64_t* rng2);
int orc_cancel_exact(void* h, const orc_bank* in, const double* parent4n, orc_bank* out, uint64_t* n_out, uint64_t* rng2) {
  return orc_cancel_exact_state(h, in, parent4n, nullptr, out, n_out, rng2);
}
// ... state6n ([n][6], may be NULL): parents_previous_direction, _previous_energy, _energy, _was_virtual, which `type: exact` reads
int orc_cancel_exact_state(void* h, const orc_bank* in, const double* parent4n, const double* state6n, orc_bank* out, uint64_t* n_out, uint64_t* rng2) {
  Problem& P = *static_cast<Problem*>(h);
  try {
    std::vector<BankedParticle> v(in->n);
    for (size_t i = 0; i < v.size(); i++) {
      v[i].r = {in->x[i], in->y[i], in->z[i]}; v[i].u = {in->ux[i], in->uy[i], in->uz[i]};
      v[i].E = in->E[i]; v[i].wgt = in->wgt[i]; v[i].wgt2 = in->wgt2[i];
      v[i].parent_history_id = in->id_a[i]; v[i].parent_daughter_id = in->id_b[i]; v[i].family_id = in->id_c[i];
      v[i].parents_previous_position = {parent4n[4 * i], parent4n[4 * i + 1], parent4n[4 * i + 2]};
      v[i].Esmp_parent = parent4n[4 * i + 3];
      if (state6n) {
        v[i].parents_previous_direction = {state6n[6 * i], state6n[6 * i + 1], state6n[6 * i + 2]};
        v[i].parents_previous_previous_energy = state6n[6 * i + 3];
        v[i].parents_previous_energy = state6n[6 * i + 4];
        v[i].parents_previous_was_virtual = state6n[6 * i + 5] != 0.;
      }
    }
    P.global_rng.state = rng2[0]; P.global_rng.inc = rng2[1];
    perform_exact_cancellation(P, v);
    rng2[0] = P.global_rng.state; rng2[1] = P.global_rng.inc;
    *n_out = v.size();
    bank_to(v, out);
    return 0;
  } catch (const std::exception& e) { P.error = e.what(); return 1; }
}

// Transporter::transport(bank, noise, &noise_bank, &noise_maker): as orc_transport, plus the noise-source bank when
// sample_noise != 0 (noise_out->n is its capacity on entry, *n_noise the true count)
int orc_transport_noise(void* h, const orc_bank* in, int noise, int sample_noise, orc_bank* out, uint64_t* n_out,
                        orc_bank* noise_out, uint64_t* n_noise, double* scores6) {
  Problem& P = *static_cast<Problem*>(h);
  try {
    auto bank = bank_from(P, in);
    Tallies& T = P.tallies;
    double b[6] = {T.k_col_score, T.k_abs_score, T.k_trk_score, T.k_tot_score, T.leak_score, T.mig_area_score};
    std::vector<BankedParticle> nb;
    auto fis = transport(P, bank, noise != 0, sample_noise ? &nb : nullptr, sample_noise != 0);
    scores6[0] = T.k_col_score - b[0]; scores6[1] = T.k_abs_score - b[1]; scores6[2] = T.k_trk_score - b[2];
    scores6[3] = T.k_tot_score - b[3]; scores6[4] = T.leak_score - b[4]; scores6[5] = T.mig_area_score - b[5];
    *n_out = fis.size();
    bank_to(fis, out);
    *n_noise = nb.size();
    if (noise_out) bank_to(nb, noise_out);
    return 0;
  } catch (const std::exception& e) { P.error = e.what(); return 1; }
}
void orc_set_keff(void* h, double keff) { static_cast<Problem*>(h)->tallies.keff_ = keff; }
// Tallies::score_noise_source / score_source over a bank (tallies.hpp:65-92)
void orc_score_source(void* h, const orc_bank* b, int noise_source) {
  Problem& P = *static_cast<Problem*>(h);
  for (uint64_t i = 0; i < b->n; i++) {
    BankedParticle p{Vec{b->x[i], b->y[i], b->z[i]}, Vec{b->ux[i], b->uy[i], b->uz[i]}, b->E[i], b->wgt[i], b->wgt2 ? b->wgt2[i] : 0., 0, 0, 0};
    for (auto& t : P.tallies.mesh)
      if (t.estimator == EST_SOURCE && (t.noise_source != 0) == (noise_source != 0)) t.score_source(p);
  }
}

// per-history trace of the last transport call (needs orc_set_trace(1))
void orc_get_trace(void* h, uint32_t* flights, uint32_t* real, uint32_t* virt, uint32_t* fis, uint64_t* hash, uint64_t* rng_state) {
  Problem& P = *static_cast<Problem*>(h);
  size_t N = P.tr_hash.size();
  std::memcpy(flights, P.tr_flights.data(), N * 4); std::memcpy(real, P.tr_real.data(), N * 4);
  std::memcpy(virt, P.tr_virtual.data(), N * 4); std::memcpy(fis, P.tr_fission.data(), N * 4);
  std::memcpy(hash, P.tr_hash.data(), N * 8); std::memcpy(rng_state, P.tr_rng_state.data(), N * 8);
}

// mesh tallies
int orc_tally_estimator(void* h, int t) { return (int)static_cast<Problem*>(h)->tallies.mesh[(size_t)t].estimator; }
int orc_ntallies(void* h) { return (int)static_cast<Problem*>(h)->tallies.mesh.size(); }
uint64_t orc_tally_size(void* h, int t) { return static_cast<Problem*>(h)->tallies.mesh[(size_t)t].tally_gen.size(); }
void orc_tally_shape(void* h, int t, uint64_t* shape4) {
  const MeshTally& m = static_cast<Problem*>(h)->tallies.mesh[(size_t)t];
  shape4[0] = m.energy_bounds.size() - 1; shape4[1] = m.Nx; shape4[2] = m.Ny; shape4[3] = m.Nz;
}
void orc_tally_get(void* h, int t, int which, double* out) {  // 0 gen, 1 avg, 2 var, 3 std
  const MeshTally& m = static_cast<Problem*>(h)->tallies.mesh[(size_t)t];
  const std::vector<double>& v = which == 0 ? m.tally_gen : (which == 1 ? m.tally_avg : m.tally_var);
  if (which == 3) {  // mesh_tally.cpp:195-197
    for (size_t i = 0; i < m.tally_var.size(); i++) out[i] = std::sqrt(m.tally_var[i] / static_cast<double>(m.g));
  } else {
    std::memcpy(out, v.data(), v.size() * 8);
  }
}
void orc_tallies_record(void* h, double mult) { static_cast<Problem*>(h)->tallies.record_generation(mult); }
void orc_tallies_clear(void* h) { static_cast<Problem*>(h)->tallies.clear_generation(); }
void orc_tallies_calc_gen(void* h, double* out6) {
  Tallies& T = static_cast<Problem*>(h)->tallies;
  T.calc_gen_values();
  out6[0] = T.k_col; out6[1] = T.k_abs; out6[2] = T.k_trk; out6[3] = T.k_tot; out6[4] = T.leak; out6[5] = T.mig;
}

// approximate cancellation + weight normalisation on a fission bank, in place
int orc_cancel_and_normalize(void* h, orc_bank* b, int do_cancel, double* stats6) {
  Problem& P = *static_cast<Problem*>(h);
  std::vector<BankedParticle> v(b->n);
  for (size_t i = 0; i < v.size(); i++) {
    v[i].r = {b->x[i], b->y[i], b->z[i]}; v[i].u = {b->ux[i], b->uy[i], b->uz[i]};
    v[i].E = b->E[i]; v[i].wgt = b->wgt[i]; v[i].wgt2 = b->wgt2[i];
    v[i].parent_history_id = b->id_a[i]; v[i].parent_daughter_id = b->id_b[i]; v[i].family_id = b->id_c[i];
  }
  if (do_cancel && P.cancel.present) perform_regional_cancellation(P, v);
  GenStats s = normalize_weights(P, v);
  stats6[0] = s.Npos; stats6[1] = s.Nneg; stats6[2] = s.Ntot; stats6[3] = s.Nnet; stats6[4] = s.Wpos; stats6[5] = s.Wneg;
  bank_to(v, b);
  return 0;
}

void orc_sobol_points(int n, double* out4n) {
  for (int i = 0; i < n; i++)
    for (unsigned d = 0; d < 4; d++) out4n[4 * i + d] = sobol_sample(static_cast<unsigned long long>(i), d);
}

// BranchlessPowerIterator::comb_particles alone.  rng2 = {state, increment} of settings::rng, updated; out->n = capacity on entry.
int orc_comb(void* h, const orc_bank* in, orc_bank* out, uint64_t* n_out, uint64_t* rng2) {
  Problem& P = *static_cast<Problem*>(h);
  try {
    std::vector<BankedParticle> v(in->n);
    for (size_t i = 0; i < v.size(); i++) {
      v[i].r = {in->x[i], in->y[i], in->z[i]}; v[i].u = {in->ux[i], in->uy[i], in->uz[i]};
      v[i].E = in->E[i]; v[i].wgt = in->wgt[i]; v[i].wgt2 = in->wgt2[i];
      v[i].parent_history_id = in->id_a[i]; v[i].parent_daughter_id = in->id_b[i]; v[i].family_id = in->id_c[i];
    }
    P.global_rng.state = rng2[0]; P.global_rng.inc = rng2[1];
    comb_particles(P, v);
    rng2[0] = P.global_rng.state; rng2[1] = P.global_rng.inc;
    *n_out = v.size();
    bank_to(v, out);
    return 0;
  } catch (const std::exception& e) { P.error = e.what(); return 1; }
}

// approximate cancellation alone (Noise::perform_regional_cancellation on a noise fission bank, noise.cpp:508-515)
int orc_cancel(void* h, orc_bank* b) {
  Problem& P = *static_cast<Problem*>(h);
  std::vector<BankedParticle> v(b->n);
  for (size_t i = 0; i < v.size(); i++) {
    v[i].r = {b->x[i], b->y[i], b->z[i]}; v[i].u = {b->ux[i], b->uy[i], b->uz[i]};
    v[i].E = b->E[i]; v[i].wgt = b->wgt[i]; v[i].wgt2 = b->wgt2[i];
    v[i].parent_history_id = b->id_a[i]; v[i].parent_daughter_id = b->id_b[i]; v[i].family_id = b->id_c[i];
  }
  if (P.cancel.present) perform_regional_cancellation(P, v);
  bank_to(v, b);
  return 0;
}

// Whole PowerIterator::run (k-eigenvalue). results: per generation kcol, ktrk, leak, mig, entropy, bank size
// summary[0..7] = kcol_avg, kcol_err, ktrk_avg, ktrk_err, leak_avg, leak_err, wall seconds, active particles
// ---- k-eigenvalue generation loop (PowerIterator::initialize / run, src/power_iterator.cpp:46-133,305-431) ----
// Stateful so that a caller (bench.py's CPU baseline legs) can time generation by generation.
struct PIState {
  std::vector<Particle> bank;
  int gen = 0, nignored = 0;
  double active_particles = 0;
  // settings: pair-distance-sqrd, families, empty-entropy-bins (power_iterator.cpp:283-297), per generation when `diagnostics` is on
  bool diagnostics = false;
  std::vector<double> r_sqrd, families, empty_frac;
};
static std::map<Problem*, PIState> g_pi;

static void pi_init(Problem& P, int nignored) {
  PIState& S = g_pi[&P];
  P.histories_counter = 0;
  P.global_histories_counter = 0;
  P.global_rng.seed_global(P.st.rng_seed);  // (no parser here: the colour draws of make_material / make_cell are not taken)
  S.bank = sample_sources(P, (size_t)P.st.nparticles);  // PowerIterator::initialize
  P.global_histories_counter += (uint64_t)P.st.nparticles;
  for (auto& p : S.bank) p.family_id = p.history_id;
  P.converged = (nignored == 0);
  if (P.entropy.present) P.entropy.zero();
  S.gen = 0;
  S.nignored = nignored;
  S.active_particles = 0;
  S.r_sqrd.clear(); S.families.clear(); S.empty_frac.clear();
}

// PowerIterator::compute_pair_dist_sqrd (power_iterator.cpp:637-663), one thread: the sum over all pairs in the reference's order
static double compute_pair_dist_sqrd(const std::vector<BankedParticle>& next_gen) {
  double Ntot = 0., r_sqr = 0.;
  for (size_t i = 0; i < next_gen.size(); i++) {
    Ntot += next_gen[i].wgt;
    for (size_t j = 0; j < next_gen.size(); j++) {
      const Vec r = next_gen[i].r - next_gen[j].r;
      r_sqr += r.dot(r) * next_gen[i].wgt * next_gen[j].wgt;
    }
  }
  r_sqr /= 2. * Ntot * Ntot;
  return r_sqr;
}

// one generation; out5 = k_col, k_trk, leak, mig, entropy
static void pi_generation(Problem& P, double* out5, uint64_t* nbank_in) {
  PIState& S = g_pi[&P];
  Tallies& T = P.tallies;
  std::vector<Particle>& bank = S.bank;
  const int g = ++S.gen;
  if (P.converged) S.active_particles += (double)bank.size();
  *nbank_in = bank.size();
  if (S.diagnostics) {  // the families that enter the generation (power_iterator.cpp:326-331)
    std::set<uint64_t> fam;
    for (const auto& p : bank) fam.insert(p.family_id);
    S.families.push_back((double)fam.size());
  }
  auto next_gen = transport(P, bank, false, nullptr, false);
  if (next_gen.empty()) throw std::runtime_error("No fission neutrons were produced.");
  if (P.entropy.present) for (auto& p : next_gen) P.entropy.add_point(p.r, p.wgt);
  if (S.diagnostics && P.entropy.present) {  // Entropy::calculate_empty_fraction (entropy.cpp:95-105) at power_iterator.cpp:613-615
    double num_empty_bins = 0.;
    for (double b : P.entropy.bins) if (b == 0.) num_empty_bins += 1.;
    S.empty_frac.push_back(num_empty_bins / static_cast<double>(P.entropy.bins.size()));
  }
  T.calc_gen_values();
  if (P.st.regional_cancellation && P.cancel.present) perform_regional_cancellation(P, next_gen);
  normalize_weights(P, next_gen);
  if (S.diagnostics) S.r_sqrd.push_back(compute_pair_dist_sqrd(next_gen));  // power_iterator.cpp:361-365
  if (P.st.mode == Settings::BRANCHLESS && P.st.branchless_combing) comb_particles(P, next_gen);  // branchless_power_iterator.cpp:361-363
  if (P.converged) {
    for (const auto& p : next_gen)
      for (auto& t : T.mesh) if (t.estimator == EST_SOURCE && !t.noise_source) t.score_source(p);
    T.record_generation();
  }
  T.clear_generation();
  out5[4] = P.entropy.present ? P.entropy.calculate_entropy() : 0.;
  bank.clear();
  P.histories_counter = P.global_histories_counter;
  bank.reserve(next_gen.size());
  for (auto& p : next_gen) {  // power_iterator.cpp:396-401
    Particle np(p.r, p.u, p.E, p.wgt, P.histories_counter++);
    np.initialize_rng(P.st.rng_seed, P.st.rng_stride);
    np.family_id = p.family_id;
    bank.push_back(std::move(np));
  }
  P.global_histories_counter += (uint64_t)next_gen.size();  // accumulate(node_nparticles); distribute_particles set it to bank.size() (simulation.cpp:128-135)
  out5[0] = T.k_col; out5[1] = T.k_trk; out5[2] = T.leak; out5[3] = T.mig;
  if (P.entropy.present) P.entropy.zero();
  if (g == S.nignored) P.converged = true;
}

void orc_free(void* h) {
  g_pi.erase(static_cast<Problem*>(h));
  delete static_cast<Problem*>(h);
}

// the optional diagnostics of the generation loop: on / off, and the series since pi_init (three arrays of up to cap values;
// returns the number of generations)
void orc_pi_set_diagnostics(void* h, int on) { g_pi[static_cast<Problem*>(h)].diagnostics = on != 0; }
int orc_pi_diagnostics(void* h, double* r_sqrd, double* families, double* empty_frac, int cap) {
  const PIState& S = g_pi[static_cast<Problem*>(h)];
  for (int g = 0; g < cap; g++) {
    if ((size_t)g < S.r_sqrd.size()) r_sqrd[g] = S.r_sqrd[(size_t)g];
    if ((size_t)g < S.families.size()) families[g] = S.families[(size_t)g];
    if ((size_t)g < S.empty_frac.size()) empty_frac[g] = S.empty_frac[(size_t)g];
  }
  return (int)S.families.size();
}

int orc_pi_init(void* h, int nignored) {
  Problem& P = *static_cast<Problem*>(h);
  try { pi_init(P, nignored); return 0; } catch (const std::exception& e) { P.error = e.what(); return 1; }
}

// runs ngen generations; out4 = seconds, particles entering transport, real collisions, last k_col
int orc_pi_run(void* h, int ngen, double* out4) {
  Problem& P = *static_cast<Problem*>(h);
  try {
    const uint64_t c0 = P.counters.real_collisions;
    double o5[5] = {0, 0, 0, 0, 0}, particles = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int g = 0; g < ngen; g++) {
      uint64_t nb = 0;
      pi_generation(P, o5, &nb);
      particles += (double)nb;
    }
    out4[0] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out4[1] = particles;
    out4[2] = (double)(P.counters.real_collisions - c0);
    out4[3] = o5[0];
    return 0;
  } catch (const std::exception& e) { P.error = e.what(); return 1; }
}

int orc_run_power_iteration(void* h, int ngen, int nignored, double* kcol, double* ktrk, double* leak, double* mig,
                            double* entropy, uint64_t* nbank, double* summary) {
  Problem& P = *static_cast<Problem*>(h);
  try {
    Tallies& T = P.tallies;
    pi_init(P, nignored);
    auto t0 = std::chrono::steady_clock::now();
    for (int g = 1; g <= ngen; g++) {
      double o5[5];
      pi_generation(P, o5, &nbank[g - 1]);
      kcol[g - 1] = o5[0]; ktrk[g - 1] = o5[1]; leak[g - 1] = o5[2]; mig[g - 1] = o5[3]; entropy[g - 1] = o5[4];
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    summary[0] = T.k_col_avg; summary[1] = T.gen > 0 ? T.err(T.k_col_var) : 0.;
    summary[2] = T.k_trk_avg; summary[3] = T.gen > 0 ? T.err(T.k_trk_var) : 0.;
    summary[4] = T.leak_avg; summary[5] = T.gen > 0 ? T.err(T.leak_var) : 0.;
    summary[6] = secs; summary[7] = g_pi[&P].active_particles;
    return 0;
  } catch (const std::exception& e) { P.error = e.what(); return 1; }
}

}  // extern "C"
