/* model.cpp -- YAML deck -> object model -> flat tables.  See model.hpp. */
#include "model.hpp"

#include <algorithm>
#include <cmath>
#include <functional>
#include <sstream>
#include <stdexcept>

namespace abeille {

using yaml_lite::Node;

namespace {

[[noreturn]] void fatal_error(const std::string& m) { throw std::runtime_error(m); }  // src/error.cpp:36-45

std::vector<double> doubles(const Node& n, size_t expect, const std::string& what) {
  if (!n || !n.IsSequence() || (expect && n.size() != expect)) fatal_error("Invalid " + what + " entry.");
  return n.as_doubles();
}

// ---- settings (src/parser.cpp:341-866) ------------------------------------------------------------
void make_settings(const Node& input, Settings& st) {
  const Node& s = input["settings"];
  if (!s || !s.IsMap()) fatal_error("Not settings specified in input file.");
  if (s["simulation"] && s["simulation"].IsScalar()) {
    const std::string sim = s["simulation"].as_string();
    if (sim == "k-eigenvalue") st.mode = ABL_MODE_K_EIGENVALUE;
    else if (sim == "noise") st.mode = ABL_MODE_NOISE;
    // modified-fixed-source (src/modified_fixed_source.cpp) transports with the k-eigenvalue kernels: its make_fission_neutrons
    // drops the division by k_col (transporter.cpp:381-386), which the driver (abeille_b200/fixed_source.py) gets with k_col = 1
    else if (sim == "modified-fixed-source") { st.mode = ABL_MODE_K_EIGENVALUE; st.fixed_source = true; }
    // fixed-source (src/fixed_source.cpp): fission neutrons are secondaries of their history; driver abeille_b200/fixed_source.py
    else if (sim == "fixed-source") { st.mode = ABL_MODE_FIXED_SOURCE; st.fixed_source = true; }
    else if (sim == "branchless-k-eigenvalue") {  // parser.cpp:355-409
      st.mode = ABL_MODE_BRANCHLESS;
      if (s["branchless-splitting"]) st.branchless_splitting = s["branchless-splitting"].as_bool();
      if (s["branchless-combing"]) st.branchless_combing = s["branchless-combing"].as_bool();
      if (s["branchless-material"]) st.branchless_material = s["branchless-material"].as_bool();
    }
    else fatal_error("Simulation mode \"" + sim + "\" is not provided by the B200 backend (k-eigenvalue, branchless-k-eigenvalue, noise, fixed-source, modified-fixed-source).");
  } else {
    fatal_error("No simulation type provided.");
  }
  if (s["transport"] && s["transport"].IsScalar()) {
    const std::string t = s["transport"].as_string();
    if (t == "delta-tracking") st.tracking = ABL_TRACK_DELTA;
    else if (t == "surface-tracking") st.tracking = ABL_TRACK_SURFACE;
    else if (t == "implicit-leakage-delta-tracking") st.tracking = ABL_TRACK_IMPLICIT_LEAKAGE;  // parser.cpp:415-418
    else if (t == "carter-tracking") {
      st.tracking = ABL_TRACK_CARTER;
      if (!input["sampling-xs-ratio"] || !input["sampling-xs-ratio"].IsSequence())
        fatal_error("Must provide the \"sampling-xs-ratio\" vector to use Carter Tracking.");
      st.sample_xs_ratio = input["sampling-xs-ratio"].as_doubles();
      for (double v : st.sample_xs_ratio)
        if (v <= 0.) fatal_error("Sampling XS ratios must be > 0.");
    } else {
      fatal_error("Invalid tracking method " + t + ".");
    }
  }
  if (!s["energy-mode"] || s["energy-mode"].as_string() != "multi-group")
    fatal_error("The B200 backend transports multi-group problems only (energy-mode: multi-group).");
  if (!s["ngroups"] || !s["ngroups"].IsScalar()) fatal_error("Number of groups for mutli-group mode not provided.");
  st.ngroups = static_cast<int>(s["ngroups"].as_int());
  if (st.ngroups < 1) fatal_error("Number of groups may not be negative.");
  if (!s["energy-bounds"] || !s["energy-bounds"].IsSequence()) fatal_error("No energy-bounds entry found in settings.");
  st.energy_bounds = s["energy-bounds"].as_doubles();
  if (static_cast<int>(st.energy_bounds.size()) != st.ngroups + 1)
    fatal_error("The number of energy-bounds must be equal to ngroups + 1.");
  if (!s["nparticles"]) fatal_error("Number of particles not specified in settings.");
  st.nparticles = static_cast<int>(s["nparticles"].as_int());
  if (!s["ngenerations"]) fatal_error("Number of generations not specified in settings.");
  st.ngenerations = static_cast<int>(s["ngenerations"].as_int());
  if (s["nignored"]) st.nignored = static_cast<int>(s["nignored"].as_int());
  else if ((st.mode == ABL_MODE_K_EIGENVALUE || st.mode == ABL_MODE_BRANCHLESS) && !st.fixed_source) fatal_error("Number of ignored generations not specified in settings.");
  else st.nignored = 0;
  if ((st.mode == ABL_MODE_K_EIGENVALUE || st.mode == ABL_MODE_BRANCHLESS) && !st.fixed_source && st.nignored >= st.ngenerations)
    fatal_error("Number of ignored generations is greater than or equal to the number of total generations.");
  if (s["nskip"]) st.nskip = static_cast<int>(s["nskip"].as_int());
  if (s["wgt-cutoff"]) {
    st.wgt_cutoff = s["wgt-cutoff"].as_double();
    if (st.wgt_cutoff < 0.) fatal_error("Roulette cutoff (\"wgt-cutoff\") must be >= 0.");
  }
  if (s["wgt-survival"]) {
    st.wgt_survival = s["wgt-survival"].as_double();
    if (st.wgt_survival <= 0.) fatal_error("Roulette survival weight (\"wgt-survival\") must be > 0.");
    if (st.wgt_survival <= st.wgt_cutoff) fatal_error("Roulette survival weight (\"wgt-survival\") must be > cutoff weight (\"wgt-cutoff\").");
  }
  if (s["wgt-split"]) {
    st.wgt_split = s["wgt-split"].as_double();
    if (st.wgt_split <= 0.) fatal_error("Splitting weight (\"wgt-split\") must be > 0.");
    if (st.wgt_split <= st.wgt_survival) fatal_error("Splitting weight (\"wgt-split\") must be > roulette survival weight (\"wgt-survival\").");
  }
  if (s["seed"]) st.rng_seed = s["seed"].as_uint64();
  if (s["stride"]) st.rng_stride = s["stride"].as_uint64();
  if (s["cancellation"]) st.regional_cancellation = s["cancellation"].as_bool();
  if (s["noise-cancellation"]) st.regional_cancellation_noise = s["noise-cancellation"].as_bool();
  if (s["cancel-noise-gens"]) st.n_cancel_noise_gens = static_cast<int>(s["cancel-noise-gens"].as_int());
  if (s["noise-angular-frequency"]) st.w_noise = s["noise-angular-frequency"].as_double();
  if (s["keff"]) st.keff = s["keff"].as_double();
  if (s["inner-generations"]) st.inner_generations = s["inner-generations"].as_bool();
  if (s["normalize-noise-source"]) st.normalize_noise_source = s["normalize-noise-source"].as_bool();
  if (s["max-run-time"] && s["max-run-time"].IsScalar()) st.max_time = s["max-run-time"].as_double() * 60.;  // minutes (parser.cpp:703-711)
  // optional diagnostics of the power iteration (src/parser.cpp:833-858)
  if (s["pair-distance-sqrd"] && s["pair-distance-sqrd"].IsScalar()) st.pair_distance_sqrd = s["pair-distance-sqrd"].as_bool();
  if (s["families"] && s["families"].IsScalar()) st.families = s["families"].as_bool();
  else if (s["families"]) fatal_error("The settings option \"families\" must be a single boolean value.");
  if (s["empty-entropy-bins"] && s["empty-entropy-bins"].IsScalar()) st.empty_entropy_bins = s["empty-entropy-bins"].as_bool();
}

// ---- materials (src/mg_nuclide.cpp:546-922, src/legendre_distribution.cpp) ----------------------------
struct LegendreDistribution {
  std::vector<double> a{0.5};
  void set_moment(size_t l, double coeff) {
    if (l == 0) return;
    a.resize(l + 1, 0.);
    a[l] = coeff * (2. * static_cast<double>(l) + 1.) / 2.;
  }
  static double legendre(unsigned n, double x) {
    switch (n) {
      case 0: return 1.;
      case 1: return x;
      case 2: return 0.5 * (3. * x * x - 1.);
      case 3: return 0.5 * (5. * x * x * x - 3. * x);
      case 4: {
        const double x_2 = x * x;
        return 0.125 * (35. * x_2 * x_2 - 30. * x_2 + 3.);
      }
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
  // bisect until the pdf is linearly interpolable to 1e-4, trapezoid cdf, renormalise
  AngleTable linearize() const {
    constexpr double TOLERANCE = 0.0001;  // constants.hpp:58
    std::vector<double> mu{-1., 1.}, p{pdf(-1.), pdf(1.)};
    size_t i = 0;
    while (i < (mu.size() - 1)) {
      const double mu_mid = 0.5 * (mu[i] + mu[i + 1]);
      const double p_interp = 0.5 * (p[i] + p[i + 1]);
      const double p_real = pdf(mu_mid);
      const double rel_diff = std::abs(p_interp - p_real) / p_real;
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
    return AngleTable{mu, p, cdf};
  }
};

MGNuclide make_mg_nuclide(const Node& mat, uint32_t id, const Settings& st) {
  const size_t G = static_cast<size_t>(st.ngroups);
  const std::string sid = std::to_string(id);
  MGNuclide n;
  n.id = id;
  if (mat["name"]) n.name = mat["name"].as_string();
  n.Et = doubles(mat["total"], G, "total xs (material " + sid + ")");
  n.Ea = doubles(mat["absorption"], G, "absorption xs (material " + sid + ")");
  if (!mat["scatter"] || !mat["scatter"].IsSequence() || mat["scatter"].size() != G)
    fatal_error("Invalid scatter matrix entry in material " + sid + ".");
  n.Ps.assign(G, std::vector<double>(G, 0.));
  for (size_t ei = 0; ei < G; ei++) {
    n.Ps[ei] = doubles(mat["scatter"][ei], G, "scatter matrix row (material " + sid + ")");
    for (double v : n.Ps[ei])
      if (v < 0.) fatal_error("Negative scattering component in material " + sid + ".");
  }
  std::vector<std::vector<LegendreDistribution>> leg(G, std::vector<LegendreDistribution>(G));
  for (size_t l = 1; l <= 5; l++) {
    const Node& m = mat["P" + std::to_string(l)];
    if (!m) continue;
    if (!m.IsSequence() || m.size() != G) fatal_error("Invalid P" + std::to_string(l) + " matrix entry in material " + sid + ".");
    for (size_t ei = 0; ei < G; ei++) {
      const std::vector<double> row = doubles(m[ei], G, "Legendre moment row (material " + sid + ")");
      for (size_t eo = 0; eo < G; eo++) leg[ei][eo].set_moment(l, row[eo]);
    }
  }
  n.angles.assign(G, std::vector<AngleTable>(G));
  for (size_t i = 0; i < G; i++)
    for (size_t o = 0; o < G; o++) n.angles[i][o] = leg[i][o].linearize();
  n.Ef.assign(G, 0.);
  bool fissile = false;
  if (mat["fission"]) {
    n.Ef = doubles(mat["fission"], G, "fission xs (material " + sid + ")");
    for (double v : n.Ef) {
      if (v < 0.) fatal_error("Negative fission xs in material " + sid + ".");
      if (v > 0.) fissile = true;
    }
  }
  n.nu_prmpt.assign(G, 0.);
  n.nu_delyd.assign(G, 0.);
  if (fissile) {
    if (mat["nu"]) {
      n.nu_prmpt = doubles(mat["nu"], G, "nu (material " + sid + ")");
    } else if (mat["nu_prompt"] && mat["nu_delayed"]) {
      n.nu_prmpt = doubles(mat["nu_prompt"], G, "nu_prompt (material " + sid + ")");
      n.nu_delyd = doubles(mat["nu_delayed"], G, "nu_delayed (material " + sid + ")");
    } else {
      fatal_error("No nu data is provided in material " + sid + ".");
    }
  }
  n.chi.assign(G, std::vector<double>(G, 0.));
  if (fissile) {
    const Node& c = mat["chi"];
    if (!c || !c.IsSequence() || (c.size() != 1 && c.size() != G)) fatal_error("Invalid chi matrix entry in material " + sid + ".");
    n.chi_is_matrix = c.size() == G;
    for (size_t ei = 0; ei < G; ei++) {
      n.chi[ei] = doubles(c[c.size() == G ? ei : 0], G, "chi row (material " + sid + ")");
      for (double v : n.chi[ei])
        if (v < 0.) fatal_error("chi is negative in material " + sid + ".");
    }
  }
  if (mat["delayed_groups"] && mat["delayed_groups"].IsMap()) {
    n.P_delayed_group = doubles(mat["delayed_groups"]["probabilities"], 0, "delayed_groups probabilities (material " + sid + ")");
    n.decay_constants = doubles(mat["delayed_groups"]["constants"], 0, "delayed_groups constants (material " + sid + ")");
    if (n.P_delayed_group.size() != n.decay_constants.size())
      fatal_error("In delayed_groups entry for material " + sid + ", probabilities and constants entries have different sizes.");
  } else if (mat["delayed_groups"]) {
    fatal_error("Invalid delayed_groups entry in material " + sid + ".");
  }
  n.speeds.assign(G, 1.);
  if (mat["group-speeds"]) n.speeds = doubles(mat["group-speeds"], G, "group-speeds (material " + sid + ")");
  else if (st.mode == ABL_MODE_NOISE) fatal_error("Missing group-speeds entry in material " + sid + ".");

  // MGNuclide constructor: make_scatter_xs, normalize_chi (src/mg_nuclide.cpp:73-118)
  n.Es.assign(G, 0.);
  for (size_t i = 0; i < G; i++) {
    n.Es[i] = 0.;
    for (size_t o = 0; o < G; o++) n.Es[i] += n.Ps[i][o];
    for (size_t o = 0; o < G; o++) n.Ps[i][o] /= n.Es[i];
  }
  for (size_t i = 0; i < G; i++) {
    double chi_i = 0.;
    for (size_t o = 0; o < G; o++) chi_i += n.chi[i][o];
    for (size_t o = 0; o < G; o++) n.chi[i][o] /= chi_i;
  }
  // check_xs (src/mg_nuclide.cpp:255-282)
  for (size_t i = 0; i < G; i++) {
    if (n.Ef[i] > n.Ea[i]) fatal_error("Ef > Ea in material " + sid + ", group " + std::to_string(i) + ".");
    const double diff = n.Et[i] - (n.Es[i] + n.Ea[i]);
    if (std::abs(diff) / n.Et[i] > 0.001) fatal_error("Es + Ea != Et in material " + sid + ", group " + std::to_string(i) + ".");
  }
  n.fissile = false;
  for (size_t i = 0; i < G; i++)
    if (n.Ef[i] != 0. && (n.nu_prmpt[i] + n.nu_delyd[i]) != 0.) n.fissile = true;
  return n;
}

// ---- surfaces (src/parser.cpp:248-293 + the surface factories) --------------------------------------------
Surface make_surface(const Node& s) {
  Surface out;
  if (!s["id"]) fatal_error("Surface must have a valid id.");
  out.id = static_cast<uint32_t>(s["id"].as_int());
  if (s["name"]) out.name = s["name"].as_string();
  if (!s["type"] || !s["type"].IsScalar()) fatal_error("Surface type not provided.");
  const std::string type = s["type"].as_string();
  auto get = [&](const char* key) {
    if (!s[key]) fatal_error("Surface " + std::to_string(out.id) + " (" + type + ") needs a \"" + key + "\" entry.");
    return s[key].as_double();
  };
  abl_surface& f = out.flat;
  f.bc = ABL_BC_NORMAL;
  if (s["boundary"]) {
    const std::string b = s["boundary"].as_string();
    if (b == "vacuum") f.bc = ABL_BC_VACUUM;
    else if (b == "reflective") f.bc = ABL_BC_REFLECTIVE;
    else if (b == "normal") f.bc = ABL_BC_NORMAL;
    else fatal_error("Unknown boundary type \"" + b + "\".");
  }
  if (type == "xplane") { f.type = ABL_SURF_XPLANE; f.p[0] = get("x0"); }
  else if (type == "yplane") { f.type = ABL_SURF_YPLANE; f.p[0] = get("y0"); }
  else if (type == "zplane") { f.type = ABL_SURF_ZPLANE; f.p[0] = get("z0"); }
  else if (type == "plane") { f.type = ABL_SURF_PLANE; f.p[0] = get("A"); f.p[1] = get("B"); f.p[2] = get("C"); f.p[3] = get("D"); }
  else if (type == "xcylinder") { f.type = ABL_SURF_XCYL; f.p[0] = get("y0"); f.p[1] = get("z0"); f.p[2] = get("r"); }
  else if (type == "ycylinder") { f.type = ABL_SURF_YCYL; f.p[0] = get("x0"); f.p[1] = get("z0"); f.p[2] = get("r"); }
  else if (type == "zcylinder") { f.type = ABL_SURF_ZCYL; f.p[0] = get("x0"); f.p[1] = get("y0"); f.p[2] = get("r"); }
  else if (type == "sphere") { f.type = ABL_SURF_SPHERE; f.p[0] = get("x0"); f.p[1] = get("y0"); f.p[2] = get("z0"); f.p[3] = get("r"); }
  else if (type == "cylinder") {  // src/cylinder.cpp:28-58
    f.type = ABL_SURF_CYL;
    f.p[0] = get("x0"); f.p[1] = get("y0"); f.p[2] = get("z0");
    double u0 = get("u0"), v0 = get("v0"), w0 = get("w0");
    const double mag = std::sqrt(u0 * u0 + v0 * v0 + w0 * w0);
    u0 /= mag; v0 /= mag; w0 /= mag;
    f.p[3] = 1. - u0 * u0;
    f.p[4] = 1. - v0 * v0;
    f.p[5] = 1. - w0 * w0;
    f.p[6] = get("r");
  } else {
    fatal_error("Surface type \"" + type + "\" is not known.");
  }
  return out;
}

// ---- cells (src/cell.cpp:203-361) -------------------------------------------------------------------------
constexpr int32_t OP_L_PAR = INT32_MAX, OP_R_PAR = INT32_MAX - 1, OP_COMP = ABL_OP_COMPLEMENT, OP_INTR = ABL_OP_INTERSECTION,
                  OP_UNIN = ABL_OP_UNION;

std::vector<int32_t> infix_to_rpn(const std::vector<int32_t>& infix) {  // shunting yard, cell.cpp:250-299
  std::vector<int32_t> rpn, stack;
  for (int32_t token : infix) {
    if (token < OP_UNIN) {
      rpn.push_back(token);
    } else if (token < OP_R_PAR) {
      while (!stack.empty()) {
        const int32_t op = stack.back();
        if (op < OP_R_PAR && ((token == OP_COMP && token < op) || (token != OP_COMP && token <= op))) {
          rpn.push_back(op);
          stack.pop_back();
        } else {
          break;
        }
      }
      stack.push_back(token);
    } else if (token == OP_L_PAR) {
      stack.push_back(token);
    } else {
      for (;;) {
        if (stack.empty()) fatal_error("Mismatched parentheses in cell region definition.");
        if (stack.back() == OP_L_PAR) break;
        rpn.push_back(stack.back());
        stack.pop_back();
      }
      stack.pop_back();
    }
  }
  while (!stack.empty()) {
    const int32_t op = stack.back();
    if (op >= OP_R_PAR) fatal_error("Mismatched parentheses in cell region definition.");
    rpn.push_back(op);
    stack.pop_back();
  }
  return rpn;
}

Cell make_cell(const Node& c, const Problem& P) {
  Cell cell;
  if (!c["id"]) fatal_error("Cell does not have a valid id.");
  cell.id = static_cast<uint32_t>(c["id"].as_int());
  if (c["name"]) cell.name = c["name"].as_string();
  if (!c["region"] || !c["region"].IsScalar()) fatal_error("Cell " + std::to_string(cell.id) + " has no valid region.");
  const std::string region_str = c["region"].as_string();
  std::vector<int32_t> region;
  std::string temp;
  auto flush = [&]() {
    if (temp.empty()) return;
    const int32_t signed_id = std::stoi(temp);
    const auto it = P.surface_id_to_indx.find(static_cast<uint32_t>(std::abs(signed_id)));
    if (it == P.surface_id_to_indx.end()) fatal_error("Could not find surface with id " + temp + " (cell " + std::to_string(cell.id) + ").");
    int32_t indx = it->second + 1;
    if (signed_id < 0) indx *= -1;
    region.push_back(indx);
    temp.clear();
  };
  for (char ch : region_str) {
    if (ch == '&' || ch == '(' || ch == ')' || ch == 'U' || ch == '~') {
      flush();
      region.push_back(ch == '&' ? OP_INTR : ch == '(' ? OP_L_PAR : ch == ')' ? OP_R_PAR : ch == 'U' ? OP_UNIN : OP_COMP);
    } else if (ch == '+' || ch == '-' || (ch >= '0' && ch <= '9')) {
      temp += ch;
    } else if (ch != ' ') {
      fatal_error("Invalid character in cell region definition.");
    }
  }
  flush();
  cell.rpn = infix_to_rpn(region);
  cell.simple = true;
  for (int32_t el : cell.rpn)
    if (el == OP_COMP || el == OP_UNIN) cell.simple = false;
  if (cell.simple) {
    std::vector<int32_t> kept;
    for (int32_t el : cell.rpn)
      if (el < OP_UNIN) kept.push_back(el);
    cell.rpn = kept;
  }
  for (int32_t token : cell.rpn) {
    if (token >= OP_UNIN) continue;
    const int bc = P.surfaces[static_cast<size_t>(std::abs(token) - 1)].flat.bc;
    if (bc == ABL_BC_VACUUM || bc == ABL_BC_REFLECTIVE) cell.vac_or_refl = true;
  }
  if (c["material"] && c["universe"]) fatal_error("Cell " + std::to_string(cell.id) + " has both a material and a universe.");
  if (c["material"]) {
    cell.fill_is_universe = false;
    cell.fill_id = static_cast<uint32_t>(c["material"].as_int());
    const auto it = P.material_id_to_indx.find(cell.fill_id);
    if (it == P.material_id_to_indx.end()) fatal_error("Could not find material with id " + std::to_string(cell.fill_id) + ".");
    cell.material_index = it->second;
  } else if (c["universe"]) {
    cell.fill_is_universe = true;
    cell.fill_id = static_cast<uint32_t>(c["universe"].as_int());
  } else {
    fatal_error("Cell " + std::to_string(cell.id) + " has neither a material nor a universe.");
  }
  return cell;
}

// ---- universes (src/parser.cpp:295-339, src/cell_universe.cpp, src/rect_lattice.cpp:312-424) ----------------
Universe make_universe(const Node& u, const Problem& P) {
  Universe uni;
  if (!u["id"]) fatal_error("Universe must have a valid id.");
  uni.id = static_cast<uint32_t>(u["id"].as_int());
  if (u["name"]) uni.name = u["name"].as_string();
  if (u["cells"]) {
    uni.type = ABL_UNI_CELLS;
    if (!u["cells"].IsSequence()) fatal_error("Invalid cells entry in universe " + std::to_string(uni.id) + ".");
    for (size_t i = 0; i < u["cells"].size(); i++) {
      const uint32_t cid = static_cast<uint32_t>(u["cells"][i].as_int());
      const auto it = P.cell_id_to_indx.find(cid);
      if (it == P.cell_id_to_indx.end()) fatal_error("Referenced cell id " + std::to_string(cid) + " could not be found.");
      uni.cell_indices.push_back(it->second);
    }
  } else if (u["pitch"]) {
    const std::string type = u["type"] ? u["type"].as_string() : std::string("");
    if (type == "hexagonal") {  // make_hex_lattice + HexLattice::HexLattice + set_elements (src/hex_lattice.cpp:30-56,204-222,457-577)
      uni.type = ABL_UNI_HEX;
      if (!u["shape"] || u["shape"].size() != 2) fatal_error("Lattice must have a valid shape.");
      const long long nrings = u["shape"][0].as_int(), nz = u["shape"][1].as_int();
      if (nrings < 1 || nz < 1 || nrings > 1000) fatal_error("Lattice must have a valid shape.");
      const std::vector<double> pitch = doubles(u["pitch"], 2, "lattice pitch");
      if (!u["origin"]) fatal_error("Lattice must have a valid origin.");
      const std::vector<double> origin = doubles(u["origin"], 3, "lattice origin");
      int top = 0;
      if (u["top"]) {
        const std::string t = u["top"].as_string();
        if (t == "pointy") top = 0;
        else if (t == "flat") top = 1;
        else fatal_error(" Uknown top for hexagonal lattice " + t + ".");
      }
      const int width = 2 * (static_cast<int>(nrings) - 1) + 1, mid = width / 2;
      uni.N = {width, width, static_cast<int>(nz)};
      uni.hex_rings = static_cast<int>(nrings);
      uni.hex_top = top;
      uni.P = {pitch[0], std::sin(3.14159265358979323846264338327950288 / 3.0), pitch[1]};
      uni.Pinv = {std::cos(3.14159265358979323846264338327950288 / 6.0), std::sin(3.14159265358979323846264338327950288 / 6.0),
                  std::cos(3.14159265358979323846264338327950288 / 3.0)};
      uni.Xl = {origin[0], origin[1], origin[2]};
      if (u["outer"]) uni.outer_id = u["outer"].as_int();
      const Node& ids = u["universes"];
      size_t nhex = 0;
      for (long long r = 0; r < nrings; r++) nhex += r == 0 ? 1 : static_cast<size_t>(6 * r);
      if (!ids || !ids.IsSequence() || ids.size() != nhex * static_cast<size_t>(nz)) fatal_error("Improper number of universes for HexLattice.");
      uni.tile_ids.assign(static_cast<size_t>(width) * static_cast<size_t>(width) * static_cast<size_t>(nz), -1);
      size_t indx = 0;
      for (int az = 0; az < nz; az++)
        for (int ar = 0; ar < width; ar++)
          for (int aq = 0; aq < width; aq++) {
            const int q = aq - mid, r = ar - mid, y = -q - r;
            const int ring = std::max(std::max(std::abs(q), std::abs(y)), std::abs(r));
            if (ring < nrings) uni.tile_ids[static_cast<size_t>(az) * static_cast<size_t>(width * width) + static_cast<size_t>(ar * width + aq)] = ids[indx++].as_int();
          }
      return uni;
    }
    if (type != "rectlinear") fatal_error("Lattice type \"" + type + "\" is not provided by the B200 backend (rectlinear and hexagonal).");
    uni.type = ABL_UNI_RECT;
    if (!u["shape"] || u["shape"].size() != 3) fatal_error("Lattice must have a valid shape.");
    for (int k = 0; k < 3; k++) uni.N[static_cast<size_t>(k)] = static_cast<int>(u["shape"][k].as_int());
    const std::vector<double> pitch = doubles(u["pitch"], 3, "lattice pitch");
    std::vector<double> origin{0., 0., 0.};
    if (u["origin"]) origin = doubles(u["origin"], 3, "lattice origin");
    for (size_t k = 0; k < 3; k++) {
      uni.P[k] = pitch[k];
      uni.Pinv[k] = 1. / pitch[k];
      uni.Xl[k] = origin[k] - static_cast<double>(uni.N[k]) * 0.5 * pitch[k];  // rect_lattice.cpp:49-51
    }
    if (u["outer"]) uni.outer_id = u["outer"].as_int();
    const Node& ids = u["universes"];
    const size_t nt = static_cast<size_t>(uni.N[0]) * static_cast<size_t>(uni.N[1]) * static_cast<size_t>(uni.N[2]);
    if (!ids || !ids.IsSequence() || ids.size() != nt) fatal_error("Lattice " + std::to_string(uni.id) + " has an invalid number of universes.");
    for (size_t i = 0; i < nt; i++) uni.tile_ids.push_back(ids[i].as_int());
  } else {
    fatal_error("Invalid universe definition.");  // src/parser.cpp:302-308 (neither `cells` nor `pitch`: e.g. the old `lattice: n` form)
  }
  return uni;
}

MeshTallySpec make_mesh_tally(const Node& t, const Settings& st) {
  MeshTallySpec m;
  abl_mesh_tally& f = m.flat;
  const std::vector<double> low = doubles(t["low"], 3, "tally low"), hi = doubles(t["hi"], 3, "tally hi");
  for (size_t k = 0; k < 3; k++) {
    f.low[k] = low[k];
    f.hi[k] = hi[k];
    f.N[k] = 1;
    if (low[k] >= hi[k]) fatal_error("Mesh tally low must be < hi.");
  }
  if (t["shape"]) {
    if (!t["shape"].IsSequence() || t["shape"].size() != 3) fatal_error("Invalid shape provided to mesh tally.");
    for (int k = 0; k < 3; k++) {
      const long long v = t["shape"][k].as_int();
      if (v < 1) fatal_error("Mesh tally shapes must be values >= 1.");
      f.N[k] = static_cast<int32_t>(v);
    }
  }
  if (!t["energy-bounds"] || !t["energy-bounds"].IsSequence()) fatal_error("No valid energy-bounds enetry provided to mesh tally.");
  m.energy_bounds = t["energy-bounds"].as_doubles();
  if (m.energy_bounds.size() < 2) fatal_error("Energy-bounds must have at least two entries.");
  f.n_energy_bins = static_cast<int32_t>(m.energy_bounds.size() - 1);
  if (!t["name"] || !t["name"].IsScalar()) fatal_error("No valid name provided to mesh tally.");
  m.name = t["name"].as_string();
  const std::string est = t["estimator"] ? t["estimator"].as_string() : std::string("collision");
  if (est == "collision") f.estimator = ABL_EST_COLLISION;
  else if (est == "track-length") f.estimator = ABL_EST_TRACK_LENGTH;
  else if (est == "source") f.estimator = ABL_EST_SOURCE;
  else fatal_error("Unknown estimator type of \"" + est + "\".");
  if (!t["quantity"] || !t["quantity"].IsScalar()) fatal_error("No tallied quantity given for mesh tally " + m.name + ".");
  const std::string q = t["quantity"].as_string();
  static const std::map<std::string, int> qmap = {
      {"flux", ABL_Q_FLUX}, {"total", ABL_Q_TOTAL}, {"elastic", ABL_Q_ELASTIC}, {"absorption", ABL_Q_ABSORPTION},
      {"fission", ABL_Q_FISSION}, {"mt", ABL_Q_MT}, {"real-flux", ABL_Q_REAL_FLUX}, {"imag-flux", ABL_Q_IMAG_FLUX},
      {"source", ABL_Q_SOURCE}, {"real-source", ABL_Q_REAL_SOURCE}, {"imag-source", ABL_Q_IMAG_SOURCE}};
  const auto it = qmap.find(q);
  if (it == qmap.end()) fatal_error("Unknown tally quantity \"" + q + "\".");
  f.quantity = it->second;
  m.quantity_str = q;
  m.estimator_str = est;
  if (f.quantity == ABL_Q_MT && t["mt"] && t["mt"].IsScalar()) m.mt = t["mt"].as_int();
  const bool source_q = f.quantity >= ABL_Q_SOURCE;
  if ((f.estimator == ABL_EST_SOURCE) != source_q) fatal_error("Quantity \"" + q + "\" is not valid for estimator \"" + est + "\".");
  if (f.estimator == ABL_EST_TRACK_LENGTH && f.quantity == ABL_Q_MT) fatal_error("Track-length tallies score flux-like quantities only.");
  f.noise_source = (f.quantity == ABL_Q_REAL_SOURCE || f.quantity == ABL_Q_IMAG_SOURCE) ? 1 : 0;
  f.net_weight = static_cast<double>(st.nparticles);  // tallies->total_weight, parser.cpp:870-871
  return m;
}

Source make_source(const Node& s) {  // src/source.cpp:92-140
  Source out;
  abl_source& f = out.flat;
  if (!s["spatial"] || !s["spatial"].IsMap()) fatal_error("No valid spatial distribution entry provided for source.");
  const Node& sp = s["spatial"];
  const std::string stype = sp["type"] ? sp["type"].as_string() : std::string("");
  if (stype == "box") {
    f.is_box = 1;
    const std::vector<double> low = doubles(sp["low"], 3, "box low"), hi = doubles(sp["hi"], 3, "box hi");
    for (size_t k = 0; k < 3; k++) { f.low[k] = low[k]; f.hi[k] = hi[k]; }
  } else if (stype == "point") {
    f.is_box = 0;
    const std::vector<double> pos = doubles(sp["position"], 3, "point position");
    for (size_t k = 0; k < 3; k++) { f.low[k] = pos[k]; f.hi[k] = pos[k]; }
  } else {
    fatal_error("Spatial distribution \"" + stype + "\" is not provided by the B200 backend (box, point).");
  }
  // src/direction_distribution.cpp:36-56
  if (!s["direction"] || !s["direction"].IsMap()) fatal_error("No valid direction distribution entry provided for source.");
  if (!s["direction"]["type"] || !s["direction"]["type"].IsScalar()) fatal_error("No valid type provided to direction distribution entry.");
  const std::string dtype = s["direction"]["type"].as_string();
  if (dtype == "isotropic") {
    f.direction_kind = ABL_DIR_ISOTROPIC;
  } else if (dtype == "mono-directional" || dtype == "cone") {  // src/mono_directional.cpp:28-42, src/cone.cpp:46-65
    const std::string what = dtype == "cone" ? "cone" : "mono-directional";
    const Node& dn = s["direction"]["direction"];
    if (!dn || !dn.IsSequence() || dn.size() != 3) fatal_error("No valid direction entry for " + what + " distribution.");
    const double x = dn[0].as_double(), y = dn[1].as_double(), z = dn[2].as_double();
    const double m = std::sqrt(x * x + y * y + z * z);  // Direction(x, y, z) renormalises (direction.hpp:37-42)
    f.dir[0] = x / m; f.dir[1] = y / m; f.dir[2] = z / m;
    f.direction_kind = ABL_DIR_MONO;
    if (dtype == "cone") {
      if (!s["direction"]["aperture"] || !s["direction"]["aperture"].IsScalar()) fatal_error("No valid aperture entry for cone distribution.");
      f.cos_aperture = std::cos(s["direction"]["aperture"].as_double());  // Cone::Cone keeps the cosine (cone.cpp:31-32)
      f.direction_kind = ABL_DIR_CONE;
    }
  } else {
    fatal_error("Invalid direction distribution type " + dtype + ".");
  }
  // src/energy_distribution.cpp:36-58
  if (!s["energy"] || !s["energy"].IsMap()) fatal_error("No valid energy distribution entry provided for source.");
  if (!s["energy"]["type"] || !s["energy"]["type"].IsScalar()) fatal_error("No valid type provided to energy distribution entry.");
  const std::string etype = s["energy"]["type"].as_string();
  const Node& en = s["energy"];
  if (etype == "mono-energetic") {  // src/mono_energetic.cpp
    if (!en["energy"] || !en["energy"].IsScalar()) fatal_error("No valid energy entry for mono-energetic distribution.");
    f.energy_kind = ABL_EN_MONO;
    f.energy = en["energy"].as_double();
  } else if (etype == "maxwellian") {  // src/maxwellian.cpp:44-55
    if (!en["a"] || !en["a"].IsScalar()) fatal_error("No valid \"a\" entry in maxwellian distribution.");
    f.energy_kind = ABL_EN_MAXWELLIAN;
    f.en_a = en["a"].as_double();
    if (f.en_a <= 0.) fatal_error("Maxwellian parameter a must be >= 0.");
  } else if (etype == "watt") {  // src/watt.cpp:49-66
    if (!en["a"] || !en["a"].IsScalar()) fatal_error("No valid \"a\" entry in watt distribution.");
    f.en_a = en["a"].as_double();
    if (f.en_a <= 0.) fatal_error("Watt parameter a must be >= 0.");
    if (!en["b"] || !en["b"].IsScalar()) fatal_error("No valid \"b\" entry in watt distribution.");
    f.en_b = en["b"].as_double();
    if (f.en_b <= 0.) fatal_error("Watt parameter b must be >= 0.");
    f.energy_kind = ABL_EN_WATT;
  } else if (etype == "tabulated") {
    fatal_error("Energy distribution \"tabulated\" (PapillonNDL's PCTable) is not provided by the B200 backend.");
  } else {
    fatal_error("Invalid energy distribution type " + etype + ".");
  }
  f.fissile_only = (s["fissile-only"] && s["fissile-only"].as_bool()) ? 1 : 0;  // SOURCE level only (source.cpp:104-110)
  if (!s["weight"] || !s["weight"].IsScalar()) fatal_error("No weight given to source.");
  f.weight = s["weight"].as_double();
  if (f.weight <= 0.) fatal_error("Source weight must be greater than zero.");
  return out;
}

MeshSpec make_mesh_spec(const Node& n, const char* what) {
  MeshSpec m;
  m.present = true;
  const std::vector<double> low = doubles(n["low"], 3, std::string(what) + " low"), hi = doubles(n["hi"], 3, std::string(what) + " hi");
  if (!n["shape"] || !n["shape"].IsSequence() || n["shape"].size() != 3) fatal_error(std::string("No valid shape provided for ") + what + ".");
  for (size_t k = 0; k < 3; k++) {
    m.low[k] = low[k];
    m.hi[k] = hi[k];
    m.N[k] = static_cast<int>(n["shape"][k].as_int());
    if (m.N[k] < 1) fatal_error(std::string(what) + " shape must be >= 1.");
  }
  return m;
}

}  // namespace

std::vector<double> discrete_table(const std::vector<double>& w) {  // bits/random.tcc:2655-2713
  std::vector<double> cp;
  if (w.size() < 2) return cp;
  double sum = 0.0;
  for (double v : w) sum += v;
  std::vector<double> p(w.size());
  for (size_t i = 0; i < w.size(); i++) p[i] = w[i] / sum;
  cp.resize(w.size());
  double acc = p[0];
  cp[0] = acc;
  for (size_t i = 1; i < w.size(); i++) {
    acc = acc + p[i];
    cp[i] = acc;
  }
  cp.back() = 1.0;
  return cp;
}

Problem Problem::from_yaml(const Node& input) {
  Problem P;
  if (!input.IsMap()) fatal_error("Input deck is not a YAML mapping.");
  make_settings(input, P.settings);
  // materials
  if (!input["materials"] || !input["materials"].IsSequence()) fatal_error("No materials are provided in input file.");
  for (size_t m = 0; m < input["materials"].size(); m++) {
    const Node& mat = input["materials"][m];
    if (!mat["id"]) fatal_error("Material is missing a valid id.");
    const uint32_t id = static_cast<uint32_t>(mat["id"].as_int());
    if (P.material_id_to_indx.count(id)) fatal_error("Material id " + std::to_string(id) + " appears more than once.");
    P.material_id_to_indx[id] = static_cast<int>(P.materials.size());
    P.materials.push_back(make_mg_nuclide(mat, id, P.settings));
    if (P.materials.back().chi_is_matrix) P.settings.chi_matrix = true;
  }
  P.settings.min_energy = P.settings.energy_bounds.front();  // parser.cpp:158-168, mg_nuclide.cpp:425-427
  P.settings.max_energy = P.settings.energy_bounds.back();
  // geometry
  if (!input["surfaces"] || !input["surfaces"].IsSequence()) fatal_error("No surfaces are provided in input file.");
  for (size_t s = 0; s < input["surfaces"].size(); s++) {
    Surface sf = make_surface(input["surfaces"][s]);
    if (P.surface_id_to_indx.count(sf.id)) fatal_error("The surface id " + std::to_string(sf.id) + " appears multiple times.");
    P.surface_id_to_indx[sf.id] = static_cast<int>(P.surfaces.size());
    P.surfaces.push_back(sf);
  }
  if (!input["cells"] || !input["cells"].IsSequence()) fatal_error("No cells are provided in input file.");
  for (size_t c = 0; c < input["cells"].size(); c++) {
    Cell cell = make_cell(input["cells"][c], P);
    if (P.cell_id_to_indx.count(cell.id)) fatal_error("The cell id " + std::to_string(cell.id) + " appears multiple times.");
    P.cell_id_to_indx[cell.id] = static_cast<int>(P.cells.size());
    P.cells.push_back(cell);
  }
  if (!input["universes"] || !input["universes"].IsSequence()) fatal_error("No universes are provided in input file.");
  for (size_t u = 0; u < input["universes"].size(); u++) {
    Universe uni = make_universe(input["universes"][u], P);
    if (P.universe_id_to_indx.count(uni.id)) fatal_error("The universe id " + std::to_string(uni.id) + " appears multiple times.");
    P.universe_id_to_indx[uni.id] = static_cast<int>(P.universes.size());
    P.universes.push_back(uni);
  }
  auto uni_index = [&](long long id, const std::string& who) {
    const auto it = P.universe_id_to_indx.find(static_cast<uint32_t>(id));
    if (id < 0 || it == P.universe_id_to_indx.end()) fatal_error("Could not find universe with id " + std::to_string(id) + " (" + who + ").");
    return it->second;
  };
  for (auto& U : P.universes) {
    if (U.type == ABL_UNI_CELLS) continue;
    for (long long id : U.tile_ids) U.tiles.push_back(id < 0 ? -1 : uni_index(id, "lattice " + std::to_string(U.id)));
    U.outer = U.outer_id < 0 ? -1 : uni_index(U.outer_id, "lattice " + std::to_string(U.id) + " outer");
  }
  for (auto& c : P.cells)
    if (c.fill_is_universe) c.universe_index = uni_index(c.fill_id, "cell " + std::to_string(c.id));
  // has_boundary_conditions: cell_universe.cpp:30-41, lattice.cpp:45-52 (the outer universe decides)
  for (auto& U : P.universes)
    if (U.type == ABL_UNI_CELLS)
      for (int ci : U.cell_indices)
        if (P.cells[static_cast<size_t>(ci)].vac_or_refl) U.has_bc = true;
  for (size_t pass = 0; pass < P.universes.size() + 1; pass++)
    for (auto& U : P.universes)
      if (U.type != ABL_UNI_CELLS) U.has_bc = U.outer >= 0 && P.universes[static_cast<size_t>(U.outer)].has_bc;
  if (!input["root-universe"]) fatal_error("No root-universe is specified in the input file.");
  P.root_universe = uni_index(input["root-universe"].as_int(), "root-universe");
  // tallies
  if (input["tallies"]) {
    if (!input["tallies"].IsSequence()) fatal_error("Tallies entry must be provided as a sequence.");
    for (size_t t = 0; t < input["tallies"].size(); t++) P.tallies.push_back(make_mesh_tally(input["tallies"][t], P.settings));
  }
  // cancelator (approximate only; src/cancelator.cpp:32-78)
  if (P.settings.regional_cancellation || P.settings.regional_cancellation_noise) {
    const Node& c = input["cancelator"];
    if (!c || !c.IsMap()) fatal_error("Regional cancelation is activated, but no cancelator entry is provided.");
    const std::string type = c["type"] ? c["type"].as_string() : std::string("");
    if (type == "basic-exact") {  // src/cancelator.cpp:42-57, src/basic_exact_mg_cancelator.cpp:610-705
      if (P.settings.tracking == ABL_TRACK_SURFACE) fatal_error("basic-exect cancelators may not be used with surface-tracking.");
      if (P.settings.tracking == ABL_TRACK_IMPLICIT_LEAKAGE)
        fatal_error("basic-exact cancelators need the sampling cross section of a flight: delta-tracking or carter-tracking.");
      P.cancelator = make_mesh_spec(c, "basic exact MG cancelator");
      P.cancelator.kind = ABL_CANCEL_BASIC_EXACT;
      if (!c["beta"] || !c["beta"].IsScalar()) fatal_error("No valid beta entry for basic exact MG cancelator.");
      const std::string beta = c["beta"].as_string();
      if (beta == "zero") P.cancelator.beta = ABL_BETA_ZERO;
      else if (beta == "minimum") P.cancelator.beta = ABL_BETA_MINIMUM;
      else if (beta == "average-f") P.cancelator.beta = ABL_BETA_AVERAGE_F;
      else if (beta == "average-g") P.cancelator.beta = ABL_BETA_AVERAGE_G;
      else fatal_error("Unkown beta entry \"" + beta + "\" for basic exact MG cancelator.");
      if (c["sobol"]) P.cancelator.sobol = c["sobol"].as_bool() ? 1 : 0;
      if (c["n-samples"]) P.cancelator.n_samples = static_cast<int>(c["n-samples"].as_int());
      if (P.cancelator.n_samples <= 0) fatal_error("n-samples must be greater than zero.");
    } else if (type == "exact") {  // src/cancelator.cpp:58-72, src/exact_mg_cancelator.cpp:594-686
      // the kernels keep what this cancelator reads from the bank, so the reference's own ExactMGCancelator runs over the GPU
      // transporter; this repo's drivers do not run it themselves (PowerIterator says so when it gets there)
      if (P.settings.tracking != ABL_TRACK_DELTA && P.settings.tracking != ABL_TRACK_CARTER)
        fatal_error("exact cancelators may not be used with surface-tracking.");
      P.cancelator = make_mesh_spec(c, "exact MG cancelator");
      P.cancelator.kind = ABL_CANCEL_EXACT;
      if (c["group-bins"] && c["group-bins"].IsSequence()) {
        for (size_t b = 0; b < c["group-bins"].size(); b++) {
          std::vector<int> bin;
          for (size_t q = 0; q < c["group-bins"][b].size(); q++) {
            const long long g = c["group-bins"][b][q].as_int();
            if (g < 0 || g >= P.settings.ngroups) fatal_error("Invalid group index in group-bins of the exact MG cancelator.");
            bin.push_back(static_cast<int>(g));
          }
          P.cancelator.group_bins.push_back(bin);
        }
      } else if (P.settings.chi_matrix) {
        fatal_error("Chi matrix is used, but no group_bins provided.\nImpossible to have exact cancellation.");
      }
      if (c["n-samples"]) P.cancelator.n_samples = static_cast<int>(c["n-samples"].as_int());
      if (P.cancelator.n_samples <= 0) fatal_error("n-samples must be greater than zero.");
    } else {
    if (type != "approximate") fatal_error("Cancelator type \"" + type + "\" is not provided by the B200 backend (approximate, basic-exact, exact).");
    P.cancelator = make_mesh_spec(c, "approximate mesh cancelator");
    P.cancelator.kind = ABL_CANCEL_APPROXIMATE;
    if (c["energy-bounds"]) {
      if (!c["energy-bounds"].IsSequence()) fatal_error("No valid energy-bounds entry for approximate mesh cancelator.");
      P.cancelator.energy_edges = c["energy-bounds"].as_doubles();
    }
    }
  }
  // sources
  if (input["sources"] && input["sources"].IsSequence())
    for (size_t s = 0; s < input["sources"].size(); s++) P.sources.push_back(make_source(input["sources"][s]));
  else
    fatal_error("No source specified for problem.");
  // Source::generate_particle redraws the energy until it lies inside (min_energy, max_energy) and gives up after 200 draws
  // (src/source.cpp:48-58); a mono-energetic source outside the range can only ever end there
  for (const Source& src : P.sources)
    if (src.flat.energy_kind == ABL_EN_MONO && (src.flat.energy <= P.settings.min_energy || P.settings.max_energy <= src.flat.energy)) fatal_error("Exceded 200 samplings of energy.");
  if (input["entropy"] && input["entropy"].IsMap()) P.entropy = make_mesh_spec(input["entropy"], "entropy mesh");
  // noise sources (src/noise_maker.cpp:39-58, src/square_oscillation_noise_source.cpp:38-83,177-250)
  if (input["noise-sources"] && input["noise-sources"].IsSequence())
    for (size_t n = 0; n < input["noise-sources"].size(); n++) {
      const Node& ns = input["noise-sources"][n];
      const std::string type = ns["type"] ? ns["type"].as_string() : std::string("");
      if (type != "square-oscillation" && type != "flat-vibration") fatal_error("Invalid noise source type " + type + ".");
      abl_noise_source f{};
      f.type = type == "flat-vibration" ? ABL_NOISE_FLAT_VIBRATION : ABL_NOISE_SQUARE_OSCILLATION;
      f.material_pos = f.material_neg = -1;
      const auto vec3 = [&](const char* key, double out[3]) {
        if (!ns[key] || !ns[key].IsSequence() || ns[key].size() != 3) fatal_error(std::string("No valid ") + key + " entry for oscillation noise source.");
        const std::vector<double> v = ns[key].as_doubles();
        for (int k = 0; k < 3; k++) out[k] = v[static_cast<size_t>(k)];
      };
      vec3("low", f.low);
      vec3("hi", f.hi);
      const auto scalar = [&](const char* key) {
        if (!ns[key]) fatal_error(std::string("No valid ") + key + " entry for oscillation noise source.");
        return ns[key].as_double();
      };
      f.angular_frequency = scalar("angular-frequency");
      if (f.low[0] >= f.hi[0] || f.low[1] >= f.hi[1] || f.low[2] >= f.hi[2]) fatal_error("Low is greater than or equal to hi in noise source.");
      if (f.angular_frequency <= 0.) fatal_error("Negative or zero frequency provided to noise source.");
      if (f.type == ABL_NOISE_SQUARE_OSCILLATION) {
        f.eps_total = scalar("epsilon-total");
        f.eps_fission = scalar("epsilon-fission");
        f.eps_scatter = scalar("epsilon-scatter");
        if (f.eps_total <= 0. || f.eps_fission <= 0. || f.eps_scatter <= 0.) fatal_error("Negative or zero epsilon provided to OscillationNoiseSource.");
      } else {  // src/flat_vibration_noise_source.cpp:314-407
        if (!ns["direction"]) fatal_error("No valid \"direction\" entry for flat vibration noise source.");
        const std::string dir = ns["direction"].as_string();
        if (dir == "x" || dir == "X") f.basis = 0;
        else if (dir == "y" || dir == "Y") f.basis = 1;
        else if (dir == "z" || dir == "Z") f.basis = 2;
        else fatal_error("Invalid direction for flat vibration noise source.");
        const auto material = [&](const char* key) {
          if (!ns[key]) fatal_error(std::string("No valid ") + key + " entry given for flat vibration noise source.");
          const auto it = P.material_id_to_indx.find(static_cast<uint32_t>(ns[key].as_int()));
          if (it == P.material_id_to_indx.end()) fatal_error(std::string(key) + " not found for flat vibration noise source.");
          return it->second;
        };
        f.material_pos = material("positive-material");
        f.material_neg = material("negative-material");
      }
      P.noise_sources.push_back(f);
    }
  if (P.settings.mode == ABL_MODE_NOISE && P.noise_sources.empty()) fatal_error("No noise source specified for noise problem.");
  // majorant: per group max over materials (src/majorant.cpp:133-176)
  const size_t G = static_cast<size_t>(P.settings.ngroups);
  P.majorant.assign(G, 0.);
  for (const auto& m : P.materials)
    for (size_t g = 0; g < G; g++) {
      const double xs = 0. + 1. * m.Et[g];
      if (xs > P.majorant[g]) P.majorant[g] = xs;
    }
  if (P.settings.tracking == ABL_TRACK_CARTER && P.settings.sample_xs_ratio.size() != G)
    fatal_error("The length of sampling-xs-ratio must be equal to ngroups.");
  if (P.max_stack_depth() > ABL_MAX_PADS) fatal_error("Geometry nesting is deeper than the backend's stack (ABL_MAX_PADS).");
  if (P.max_frame_depth() > ABL_MAX_FRAMES)
    fatal_error("Lattices are nested deeper than the backend's coordinate frames (ABL_MAX_FRAMES).");
  return P;
}

int Problem::max_stack_depth() const {
  std::function<int(int, int)> depth = [&](int uni, int guard) -> int {
    if (guard > 64) fatal_error("Universe nesting is recursive.");
    const Universe& U = universes[static_cast<size_t>(uni)];
    int best = 0;
    if (U.type == ABL_UNI_CELLS) {
      for (int ci : U.cell_indices) {
        const Cell& c = cells[static_cast<size_t>(ci)];
        best = std::max(best, c.fill_is_universe ? 2 + depth(c.universe_index, guard + 1) : 2);
      }
      return best;
    }
    for (int t : U.tiles)
      if (t >= 0) best = std::max(best, 1 + depth(t, guard + 1));
    if (U.outer >= 0) best = std::max(best, 1 + depth(U.outer, guard + 1));
    return std::max(best, 1);
  };
  return depth(root_universe, 0);
}

// local coordinate frames a particle can hold at once: the global one plus one per nested lattice tile entered
int Problem::max_frame_depth() const {
  std::function<int(int, int)> depth = [&](int uni, int guard) -> int {
    if (guard > 64) fatal_error("Universe nesting is recursive.");
    const Universe& U = universes[static_cast<size_t>(uni)];
    int best = 0;
    if (U.type == ABL_UNI_CELLS) {
      for (int ci : U.cell_indices) {
        const Cell& c = cells[static_cast<size_t>(ci)];
        if (c.fill_is_universe) best = std::max(best, depth(c.universe_index, guard + 1));
      }
      return best;
    }
    for (int t : U.tiles)
      if (t >= 0) best = std::max(best, 1 + depth(t, guard + 1));
    if (U.outer >= 0) best = std::max(best, depth(U.outer, guard + 1));
    return best;
  };
  return 1 + depth(root_universe, 0);
}

void Problem::flatten(FlatProblem& F) const {
  const Settings& st = settings;
  const size_t G = static_cast<size_t>(st.ngroups), M = materials.size();
  abl_problem& p = F.p;
  p = abl_problem{};
  p.mode = st.mode;
  p.branchless_flags = (st.branchless_material ? ABL_BRANCHLESS_MATERIAL : 0) | (st.branchless_splitting ? ABL_BRANCHLESS_SPLITTING : 0);
  p.tracking = st.tracking;
  p.ngroups = st.ngroups;
  p.inner_generations = st.inner_generations ? 1 : 0;
  p.energy_bounds = st.energy_bounds.data();
  p.wgt_cutoff = st.wgt_cutoff;
  p.wgt_survival = st.wgt_survival;
  p.wgt_split = st.wgt_split;
  p.min_energy = st.min_energy;
  p.rng_seed = st.rng_seed;
  p.rng_stride = st.rng_stride;
  p.w_noise = st.w_noise;
  p.eta = st.eta;
  p.keff = st.keff;
  // geometry
  F.surfaces.clear();
  for (const auto& s : surfaces) F.surfaces.push_back(s.flat);
  F.cells.clear();
  F.rpn.clear();
  for (const auto& c : cells) {
    abl_cell fc{};
    fc.rpn_offset = static_cast<int32_t>(F.rpn.size());
    fc.rpn_len = static_cast<int32_t>(c.rpn.size());
    F.rpn.insert(F.rpn.end(), c.rpn.begin(), c.rpn.end());
    fc.simple = c.simple;
    fc.vac_or_refl = c.vac_or_refl;
    fc.fill_universe = c.fill_is_universe ? c.universe_index : -1;
    fc.material = c.fill_is_universe ? -1 : c.material_index;
    F.cells.push_back(fc);
  }
  F.universes.clear();
  F.universe_cells.clear();
  F.lattice_tiles.clear();
  for (const auto& U : universes) {
    abl_universe fu{};
    fu.type = U.type;
    fu.has_bc = U.has_bc;
    fu.outer = -1;
    if (U.type == ABL_UNI_CELLS) {
      fu.cell_offset = static_cast<int32_t>(F.universe_cells.size());
      fu.ncells = static_cast<int32_t>(U.cell_indices.size());
      F.universe_cells.insert(F.universe_cells.end(), U.cell_indices.begin(), U.cell_indices.end());
    } else {
      for (size_t k = 0; k < 3; k++) {
        fu.N[k] = U.N[k];
        fu.P[k] = U.P[k];
        fu.Pinv[k] = U.Pinv[k];
        fu.Xl[k] = U.Xl[k];
      }
      if (U.type == ABL_UNI_HEX) fu.pad_ = U.hex_rings | (U.hex_top << 16);
      fu.tile_offset = static_cast<int32_t>(F.lattice_tiles.size());
      F.lattice_tiles.insert(F.lattice_tiles.end(), U.tiles.begin(), U.tiles.end());
      fu.outer = U.outer;
    }
    F.universes.push_back(fu);
  }
  p.nsurfaces = static_cast<int32_t>(F.surfaces.size());
  p.ncells = static_cast<int32_t>(F.cells.size());
  p.nrpn = static_cast<int32_t>(F.rpn.size());
  p.nuniverses = static_cast<int32_t>(F.universes.size());
  p.n_universe_cells = static_cast<int32_t>(F.universe_cells.size());
  p.n_lattice_tiles = static_cast<int32_t>(F.lattice_tiles.size());
  p.root_universe = root_universe;
  p.surfaces = F.surfaces.data();
  p.cells = F.cells.data();
  p.rpn = F.rpn.data();
  p.universes = F.universes.data();
  p.universe_cells = F.universe_cells.data();
  p.lattice_tiles = F.lattice_tiles.data();
  // materials -> [M*G] / [M*G*G] tables; micro xs as MGNuclide::get_micro_xs builds them (mg_nuclide.cpp:394-411)
  for (auto* v : {&F.Et, &F.Ea, &F.Ef, &F.Es, &F.nu, &F.nud, &F.speeds, &F.chi_cdf, &F.scatter_cdf, &F.amu, &F.apdf, &F.acdf, &F.dcdf,
                  &F.dlambda, &F.smp, &F.tally_eb})
    v->clear();
  F.angle.clear();
  F.delayed_offset.assign(1, 0);
  F.fissile.clear();
  for (size_t m = 0; m < M; m++) {
    const MGNuclide& n = materials[m];
    for (size_t g = 0; g < G; g++) {
      F.Et.push_back(n.Et[g]);
      F.Ef.push_back(n.Ef[g]);
      F.Ea.push_back(n.Ef[g] + (n.Ea[g] - n.Ef[g]));
      F.Es.push_back(n.Es[g]);
      F.nu.push_back(n.nu_prmpt[g] + n.nu_delyd[g]);
      F.nud.push_back(n.nu_delyd[g]);
      F.speeds.push_back(n.speeds[g]);
      std::vector<double> cc = discrete_table(n.chi[g]), sc = discrete_table(n.Ps[g]);
      cc.resize(G, 1.0);  // G == 1: no table is used (and no draw is made)
      sc.resize(G, 1.0);
      F.chi_cdf.insert(F.chi_cdf.end(), cc.begin(), cc.end());
      F.scatter_cdf.insert(F.scatter_cdf.end(), sc.begin(), sc.end());
      for (size_t o = 0; o < G; o++) {
        const AngleTable& a = n.angles[g][o];
        // identical tables (the isotropic default above all) are stored once
        int32_t off = -1;
        for (const auto& prev : F.angle) {
          if (static_cast<size_t>(prev.n) != a.mu.size()) continue;
          const size_t po = static_cast<size_t>(prev.offset);
          if (std::equal(a.mu.begin(), a.mu.end(), F.amu.begin() + static_cast<long>(po)) &&
              std::equal(a.pdf.begin(), a.pdf.end(), F.apdf.begin() + static_cast<long>(po)) &&
              std::equal(a.cdf.begin(), a.cdf.end(), F.acdf.begin() + static_cast<long>(po))) {
            off = prev.offset;
            break;
          }
        }
        if (off < 0) {
          off = static_cast<int32_t>(F.amu.size());
          F.amu.insert(F.amu.end(), a.mu.begin(), a.mu.end());
          F.apdf.insert(F.apdf.end(), a.pdf.begin(), a.pdf.end());
          F.acdf.insert(F.acdf.end(), a.cdf.begin(), a.cdf.end());
        }
        F.angle.push_back(abl_angle_table{off, static_cast<int32_t>(a.mu.size())});
      }
    }
    std::vector<double> dc = discrete_table(n.P_delayed_group);
    dc.resize(n.P_delayed_group.size(), 1.0);
    F.dcdf.insert(F.dcdf.end(), dc.begin(), dc.end());
    F.dlambda.insert(F.dlambda.end(), n.decay_constants.begin(), n.decay_constants.end());
    F.delayed_offset.push_back(static_cast<int32_t>(F.dcdf.size()));
    F.fissile.push_back(n.fissile ? 1 : 0);
  }
  if (F.dcdf.empty()) {  // keep the pointers non-null
    F.dcdf.push_back(1.0);
    F.dlambda.push_back(0.0);
  }
  p.nmaterials = static_cast<int32_t>(M);
  p.n_angle_points = static_cast<int32_t>(F.amu.size());
  p.xs_total = F.Et.data();
  p.xs_absorption = F.Ea.data();
  p.xs_fission = F.Ef.data();
  p.xs_elastic = F.Es.data();
  p.nu_total = F.nu.data();
  p.nu_delayed = F.nud.data();
  p.speeds = F.speeds.data();
  p.chi_cdf = F.chi_cdf.data();
  p.scatter_cdf = F.scatter_cdf.data();
  p.angle = F.angle.data();
  p.angle_mu = F.amu.data();
  p.angle_pdf = F.apdf.data();
  p.angle_cdf = F.acdf.data();
  p.delayed_offset = F.delayed_offset.data();
  p.delayed_cdf = F.dcdf.data();
  p.delayed_lambda = F.dlambda.data();
  p.fissile = F.fissile.data();
  // sampling xs: majorant (delta) or ratio * majorant (carter, carter_tracker.cpp:60-75)
  F.smp = majorant;
  if (st.tracking == ABL_TRACK_CARTER)
    for (size_t g = 0; g < G; g++) F.smp[g] = majorant[g] * st.sample_xs_ratio[g];
  p.sampling_xs = F.smp.data();
  // tallies
  F.tallies.clear();
  for (const auto& t : tallies) {
    abl_mesh_tally ft = t.flat;
    ft.ebounds_offset = static_cast<int32_t>(F.tally_eb.size());
    F.tally_eb.insert(F.tally_eb.end(), t.energy_bounds.begin(), t.energy_bounds.end());
    F.tallies.push_back(ft);
  }
  auto mesh3 = [&](const MeshSpec& m) {
    abl_mesh3 f{};
    f.present = m.present ? 1 : 0;
    for (size_t k = 0; k < 3; k++) {
      f.N[k] = m.N[k];
      f.low[k] = m.low[k];
      f.hi[k] = m.hi[k];
    }
    f.n_energy_edges = static_cast<int32_t>(m.energy_edges.size());
    f.eedges_offset = static_cast<int32_t>(F.tally_eb.size());
    F.tally_eb.insert(F.tally_eb.end(), m.energy_edges.begin(), m.energy_edges.end());
    return f;
  };
  p.entropy = mesh3(entropy);
  p.cancelator = mesh3(cancelator);
  p.cancelator.kind = cancelator.kind;
  p.cancelator.beta = cancelator.beta;
  p.cancelator.sobol = cancelator.sobol;
  p.cancelator.n_samples = cancelator.n_samples;
  F.chi_pdf.clear();
  for (const auto& n : materials)
    for (size_t g = 0; g < G; g++) F.chi_pdf.insert(F.chi_pdf.end(), n.chi[g].begin(), n.chi[g].end());
  F.exact_group_bins.assign(1, static_cast<int32_t>(cancelator.group_bins.size()));
  for (const auto& b : cancelator.group_bins) {
    F.exact_group_bins.push_back(static_cast<int32_t>(b.size()));
    F.exact_group_bins.insert(F.exact_group_bins.end(), b.begin(), b.end());
  }
  p.chi_pdf = F.chi_pdf.data();
  p.exact_group_bins = F.exact_group_bins.data();
  p.n_exact_group_bins = static_cast<int32_t>(F.exact_group_bins.size());
  p.chi_matrix = st.chi_matrix ? 1 : 0;
  if (F.tally_eb.empty()) F.tally_eb.push_back(0.);
  p.ntallies = static_cast<int32_t>(F.tallies.size());
  p.n_tally_energy_bounds = static_cast<int32_t>(F.tally_eb.size());
  p.tallies = F.tallies.data();
  p.tally_energy_bounds = F.tally_eb.data();
  F.sources.clear();
  for (const auto& s : sources) F.sources.push_back(s.flat);
  p.nsources = static_cast<int32_t>(F.sources.size());
  p.sources = F.sources.data();
  F.noise_sources = noise_sources;
  p.n_noise_sources = static_cast<int32_t>(F.noise_sources.size());
  p.noise_sources = F.noise_sources.data();
}

}  // namespace abeille
