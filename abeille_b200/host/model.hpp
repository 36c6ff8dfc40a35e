/* model.hpp -- host-side object model of an Abeille problem and its flattening into abl_problem.
 *
 * Keeps the reference's input schema and object names (Surface, Cell, Universe/RectLattice, MGNuclide,
 * Source, MeshTally descriptions, settings) -- reference src/parser.cpp:77-137 builds the same graph
 * from the same YAML keys.  Unlike the reference, the objects here carry data only: all geometry
 * and physics evaluation happens on the device, from the flat tables produced by Problem::flatten().
 */
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/abeille_b200.h"
#include "yaml_lite.hpp"

namespace abeille {

struct Settings {  // include/utils/settings.hpp:54-128, defaults src/settings.cpp:34-114
  int mode = ABL_MODE_K_EIGENVALUE;
  bool fixed_source = false;  // simulation: modified-fixed-source (transport is the k-eigenvalue one; no ignored generations)
  bool branchless_splitting = false, branchless_combing = true, branchless_material = true;  // settings.cpp:87-89
  bool chi_matrix = false;  // a fissile material gave one chi row per group (mg_nuclide.cpp:836-850)
  int tracking = ABL_TRACK_SURFACE;
  int ngroups = 0;
  std::vector<double> energy_bounds;
  int nparticles = 100000, ngenerations = 120, nignored = 20, nskip = 10;
  double wgt_cutoff = 0.25, wgt_survival = 1.0, wgt_split = 2.0;
  uint64_t rng_seed = 19073486328125ULL, rng_stride = 152917ULL;
  double min_energy = 0., max_energy = 100000.;
  std::vector<double> sample_xs_ratio;
  bool regional_cancellation = false, regional_cancellation_noise = false;
  int n_cancel_noise_gens = INT32_MAX;
  bool inner_generations = true, normalize_noise_source = true;
  bool pair_distance_sqrd = false, families = false, empty_entropy_bins = false;  // settings.cpp:74-76, parser.cpp:833-858
  double max_time = 1.e300;  // [s] settings: max-run-time, given in minutes (settings.cpp:46, parser.cpp:703-711)
  double w_noise = -1., eta = 1., keff = 1.;
};

struct Surface {  // include/geometry/surfaces/*.hpp
  uint32_t id = 0;
  std::string name;
  abl_surface flat{};
};

struct Cell {  // include/geometry/cell.hpp
  uint32_t id = 0;
  std::string name;
  std::vector<int32_t> rpn;
  bool simple = true, vac_or_refl = false, fill_is_universe = false;
  uint32_t fill_id = 0;
  int material_index = -1, universe_index = -1;
};

struct Universe {  // CellUniverse | RectLattice | HexLattice
  uint32_t id = 0;
  std::string name;
  int type = ABL_UNI_CELLS;
  bool has_bc = false;
  std::vector<int> cell_indices;
  std::array<int, 3> N{0, 0, 0};
  std::array<double, 3> P{0, 0, 0}, Pinv{0, 0, 0}, Xl{0, 0, 0};
  std::vector<long long> tile_ids;  // as read
  std::vector<int> tiles;           // universe indices
  long long outer_id = -1;
  int outer = -1;
  int hex_rings = 0, hex_top = 0;  // HexLattice: N = {width, width, nz}; tile_ids laid out by HexLattice::linear_index, -1 outside
};

struct AngleTable {  // MGAngleDistribution
  std::vector<double> mu{-1., 1.}, pdf{0.5, 0.5}, cdf{0., 1.};
};

struct MGNuclide {  // src/mg_nuclide.cpp:30-118,579-922 ; one per material in MG mode
  uint32_t id = 0;
  std::string name;
  std::vector<double> Et, Ea, Ef, Es, nu_prmpt, nu_delyd, speeds;
  std::vector<std::vector<double>> chi, Ps;
  std::vector<std::vector<AngleTable>> angles;
  std::vector<double> P_delayed_group, decay_constants;
  bool fissile = false;
  bool chi_is_matrix = false;  // one chi row per group was given (sets settings::chi_matrix, mg_nuclide.cpp:808-810)
};

struct Source {
  abl_source flat{};
};

struct MeshTallySpec {
  std::string name;
  abl_mesh_tally flat{};
  std::vector<double> energy_bounds;
  // what MeshTally::write_tally stores as attributes beside the arrays (src/mesh_tally.cpp:154-206)
  std::string quantity_str, estimator_str;
  long long mt = 0;
};

struct MeshSpec {
  bool present = false;
  std::array<int, 3> N{1, 1, 1};
  std::array<double, 3> low{0, 0, 0}, hi{0, 0, 0};
  std::vector<double> energy_edges;
  std::vector<std::vector<int>> group_bins;  // type: exact (exact_mg_cancelator.cpp:634-662)
  int sobol = 1, n_samples = 10;  // basic-exact defaults (basic_exact_mg_cancelator.cpp:670-686)
  int kind = 0, beta = 0;  // cancelator: ABL_CANCEL_*, ABL_BETA_* (src/cancelator.cpp:40-57, basic_exact_mg_cancelator.cpp:650-668)
};

// Owns the vectors an abl_problem points into
struct FlatProblem {
  abl_problem p{};
  std::vector<abl_surface> surfaces;
  std::vector<abl_cell> cells;
  std::vector<int32_t> rpn, universe_cells, lattice_tiles, delayed_offset, fissile, exact_group_bins;
  std::vector<double> chi_pdf;
  std::vector<abl_universe> universes;
  std::vector<double> Et, Ea, Ef, Es, nu, nud, speeds, chi_cdf, scatter_cdf, amu, apdf, acdf, dcdf, dlambda, smp, tally_eb;
  std::vector<abl_angle_table> angle;
  std::vector<abl_mesh_tally> tallies;
  std::vector<abl_source> sources;
  std::vector<abl_noise_source> noise_sources;
};

class Problem {
 public:
  Settings settings;
  std::vector<Surface> surfaces;
  std::vector<Cell> cells;
  std::vector<Universe> universes;
  std::vector<MGNuclide> materials;
  std::vector<Source> sources;
  std::vector<abl_noise_source> noise_sources;  // square-oscillation only (src/noise_maker.cpp:39-58)
  std::vector<MeshTallySpec> tallies;
  MeshSpec entropy, cancelator;
  int root_universe = -1;
  std::map<uint32_t, int> surface_id_to_indx, cell_id_to_indx, universe_id_to_indx, material_id_to_indx;
  std::vector<double> majorant;  // per group (src/majorant.cpp:133-176)
  std::vector<std::string> warnings;

  static Problem from_yaml(const yaml_lite::Node& input);  // parse_input_file (src/parser.cpp:77-137)
  void flatten(FlatProblem& out) const;
  int max_stack_depth() const;
  int max_frame_depth() const;
};

// libstdc++ std::discrete_distribution partial sums (what RNG::discrete builds on every call, rng.hpp:88-96)
std::vector<double> discrete_table(const std::vector<double>& w);

}  // namespace abeille
