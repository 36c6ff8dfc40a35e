/* flatten_problem.hpp -- abl_problem from the reference's LIVE objects.
 *
 * What INTEGRATION.md's GPUTransporter constructor needs when it is built inside the reference instead of from the YAML deck
 * (src/parser.cpp:889-911, make_transporter()): the tables of include/abeille_b200.h filled from
 *   settings::*                                   (include/utils/settings.hpp)
 *   geometry::surfaces / cells / universes / root (include/geometry/geometry.hpp; Surface subclasses, Cell, CellUniverse,
 *                                                  RectLattice, HexLattice)
 *   materials (one MGNuclide each, atoms_bcm = 1) (include/materials/material.hpp, mg_nuclide.hpp, mg_angle_distribution.hpp)
 *   the Tallies object's collision / track-length mesh tallies (include/simulation/tallies.hpp, mesh_tally.hpp)
 *   the cancelator, when it is one the kernels must keep the parents' data for (BasicExactMGCancelator)
 *   the NoiseMaker's square-oscillation and flat-vibration sources, in a noise run
 * Sources, the entropy mesh and the approximate cancelator mesh are not flattened: Transporter::transport does not use them (the
 * reference's drivers sample, bin and cancel on their side of the boundary).
 *
 * The reference keeps most of this in private members without accessors; this header reads them directly and is compiled with
 * -fno-access-control in oracle/_ref (a maintainer would add the accessors or a friend declaration).  Tested in two ways: the
 * tables equal, value for value, what this repo's host library flattens from the same deck (tests/test_reference_pins.py), and the
 * reference's own PowerIterator::run() over a GPUTransporter built from them reproduces its CPU results
 * (tests/test_gpu_reference_golden.py).
 */
#pragma once
#include <geometry/cell_universe.hpp>
#include <geometry/geometry.hpp>
#include <geometry/hex_lattice.hpp>
#include <geometry/rect_lattice.hpp>
#include <geometry/surfaces/all_surfaces.hpp>
#include <materials/material.hpp>
#include <materials/mg_nuclide.hpp>
#include <simulation/basic_exact_mg_cancelator.hpp>
#include <simulation/cancelator.hpp>
#include <simulation/exact_mg_cancelator.hpp>
#include <simulation/flat_vibration_noise_source.hpp>
#include <simulation/noise_maker.hpp>
#include <simulation/square_oscillation_noise_source.hpp>
#include <simulation/collision_mesh_tally.hpp>
#include <simulation/tallies.hpp>
#include <simulation/track_length_mesh_tally.hpp>
#include <utils/error.hpp>
#include <utils/settings.hpp>

#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <vector>

#include "../include/abeille_b200.h"

extern std::map<uint32_t, std::shared_ptr<Material>> materials;  // include/materials/material.hpp

namespace abl_integration {

struct FlatProblem {  // owns what the abl_problem points into
  abl_problem p{};
  std::vector<abl_surface> surfaces;
  std::vector<abl_cell> cells;
  std::vector<abl_universe> universes;
  std::vector<int32_t> rpn, universe_cells, lattice_tiles, delayed_offset, fissile, exact_group_bins;
  std::vector<double> chi_pdf;
  std::vector<double> energy_bounds, Et, Ea, Ef, Es, nu, nud, speeds, chi_cdf, scatter_cdf, amu, apdf, acdf, dcdf, dlambda, smp, tally_eb;
  std::vector<abl_angle_table> angle;
  std::vector<abl_mesh_tally> tallies;
  std::vector<abl_noise_source> noise_sources;
};

// std::discrete_distribution's table as RNG::discrete uses it (include/utils/rng.hpp:88-96; libstdc++ random.tcc:2655-2713):
// probabilities w / sum, partial sums, the last one set to 1; fewer than two weights: no table (and no draw)
inline std::vector<double> discrete_table(const std::vector<double>& w) {
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

inline abl_surface flatten_surface(const Surface& s) {
  abl_surface f{};
  f.bc = s.boundary() == BoundaryType::Vacuum ? ABL_BC_VACUUM : (s.boundary() == BoundaryType::Reflective ? ABL_BC_REFLECTIVE : ABL_BC_NORMAL);
  if (auto* a = dynamic_cast<const XPlane*>(&s)) { f.type = ABL_SURF_XPLANE; f.p[0] = a->x0; }
  else if (auto* b = dynamic_cast<const YPlane*>(&s)) { f.type = ABL_SURF_YPLANE; f.p[0] = b->y0; }
  else if (auto* c = dynamic_cast<const ZPlane*>(&s)) { f.type = ABL_SURF_ZPLANE; f.p[0] = c->z0; }
  else if (auto* d = dynamic_cast<const Plane*>(&s)) { f.type = ABL_SURF_PLANE; f.p[0] = d->A; f.p[1] = d->B; f.p[2] = d->C; f.p[3] = d->D; }
  else if (auto* e = dynamic_cast<const XCylinder*>(&s)) { f.type = ABL_SURF_XCYL; f.p[0] = e->y0; f.p[1] = e->z0; f.p[2] = e->R; }
  else if (auto* g = dynamic_cast<const YCylinder*>(&s)) { f.type = ABL_SURF_YCYL; f.p[0] = g->x0; f.p[1] = g->z0; f.p[2] = g->R; }
  else if (auto* h = dynamic_cast<const ZCylinder*>(&s)) { f.type = ABL_SURF_ZCYL; f.p[0] = h->x0; f.p[1] = h->y0; f.p[2] = h->R; }
  else if (auto* i = dynamic_cast<const Sphere*>(&s)) { f.type = ABL_SURF_SPHERE; f.p[0] = i->x0; f.p[1] = i->y0; f.p[2] = i->z0; f.p[3] = i->R; }
  else if (auto* j = dynamic_cast<const Cylinder*>(&s)) {
    f.type = ABL_SURF_CYL;
    f.p[0] = j->x0; f.p[1] = j->y0; f.p[2] = j->z0; f.p[3] = j->alpha; f.p[4] = j->beta; f.p[5] = j->gamma; f.p[6] = j->R;
  } else {
    fatal_error("flatten_problem: unknown Surface subclass");
  }
  return f;
}

// cancelator: only its kind matters to the transporter (an exact cancelator makes the kernels keep the parents' data)
// noise_maker: the NoiseMaker of a noise run (its square-oscillation and flat-vibration sources), else null
inline void flatten_problem(FlatProblem& F, const Tallies& tallies, const Cancelator* cancelator, const NoiseMaker* noise_maker = nullptr) {
  abl_problem& p = F.p;
  p = abl_problem{};
  const size_t G = settings::ngroups;
  switch (settings::mode) {
    case settings::SimulationMode::K_EIGENVALUE:
    case settings::SimulationMode::MODIFIED_FIXED_SOURCE: p.mode = ABL_MODE_K_EIGENVALUE; break;
    case settings::SimulationMode::FIXED_SOURCE: p.mode = ABL_MODE_FIXED_SOURCE; break;
    case settings::SimulationMode::BRANCHLESS_K_EIGENVALUE: p.mode = ABL_MODE_BRANCHLESS; break;
    default:
      p.mode = ABL_MODE_NOISE;
      if (!noise_maker) fatal_error("flatten_problem: a noise run needs its NoiseMaker");
  }
  p.branchless_flags = (settings::branchless_material ? ABL_BRANCHLESS_MATERIAL : 0) | (settings::branchless_splitting ? ABL_BRANCHLESS_SPLITTING : 0);
  switch (settings::tracking) {
    case settings::TrackingMode::SURFACE_TRACKING: p.tracking = ABL_TRACK_SURFACE; break;
    case settings::TrackingMode::DELTA_TRACKING: p.tracking = ABL_TRACK_DELTA; break;
    case settings::TrackingMode::CARTER_TRACKING: p.tracking = ABL_TRACK_CARTER; break;
    default: p.tracking = ABL_TRACK_IMPLICIT_LEAKAGE; break;
  }
  p.ngroups = static_cast<int32_t>(G);
  p.inner_generations = settings::inner_generations ? 1 : 0;
  F.energy_bounds = settings::energy_bounds;
  p.energy_bounds = F.energy_bounds.data();
  p.wgt_cutoff = settings::wgt_cutoff;
  p.wgt_survival = settings::wgt_survival;
  p.wgt_split = settings::wgt_split;
  p.min_energy = settings::min_energy;
  p.rng_seed = settings::rng_seed;
  p.rng_stride = settings::rng_stride;
  p.w_noise = settings::w_noise;
  p.eta = settings::eta;
  p.keff = settings::keff;

  // materials in the order of the global map (by id); cells refer to them by position in that order
  std::vector<Material*> mats;
  for (const auto& kv : materials) mats.push_back(kv.second.get());
  auto material_index = [&](const Material* m) {
    for (size_t i = 0; i < mats.size(); i++)
      if (mats[i] == m) return static_cast<int32_t>(i);
    fatal_error("flatten_problem: a cell is filled with a material that is not in the global materials map");
    return -1;
  };
  auto universe_index = [&](const Universe* u) {
    for (size_t i = 0; i < geometry::universes.size(); i++)
      if (geometry::universes[i].get() == u) return static_cast<int32_t>(i);
    fatal_error("flatten_problem: unknown universe");
    return -1;
  };

  // geometry
  for (const auto& s : geometry::surfaces) F.surfaces.push_back(flatten_surface(*s));
  for (const auto& c : geometry::cells) {
    abl_cell fc{};
    fc.rpn_offset = static_cast<int32_t>(F.rpn.size());
    fc.rpn_len = static_cast<int32_t>(c->rpn.size());
    for (int32_t t : c->rpn)  // operands are signed (surface index + 1); the operator codes are the reference's (cell.hpp:42-48)
      F.rpn.push_back(t == OP::COMP ? ABL_OP_COMPLEMENT : (t == OP::INTR ? ABL_OP_INTERSECTION : (t == OP::UNIN ? ABL_OP_UNION : t)));
    fc.simple = c->simple ? 1 : 0;
    fc.vac_or_refl = c->vacuum_or_reflective_ ? 1 : 0;
    const bool fill_universe = c->fill() == Cell::Fill::Universe;
    fc.fill_universe = fill_universe ? universe_index(c->universe()) : -1;
    fc.material = fill_universe ? -1 : material_index(c->material());
    F.cells.push_back(fc);
  }
  for (const auto& up : geometry::universes) {
    abl_universe fu{};
    fu.outer = -1;
    fu.has_bc = up->has_boundary_conditions() ? 1 : 0;
    if (auto* cu = dynamic_cast<const CellUniverse*>(up.get())) {
      fu.type = ABL_UNI_CELLS;
      fu.cell_offset = static_cast<int32_t>(F.universe_cells.size());
      fu.ncells = static_cast<int32_t>(cu->cell_indicies.size());
      for (uint32_t ci : cu->cell_indicies) F.universe_cells.push_back(static_cast<int32_t>(ci));
    } else if (auto* rl = dynamic_cast<const RectLattice*>(up.get())) {
      fu.type = ABL_UNI_RECT;
      fu.N[0] = static_cast<int32_t>(rl->Nx); fu.N[1] = static_cast<int32_t>(rl->Ny); fu.N[2] = static_cast<int32_t>(rl->Nz);
      fu.P[0] = rl->Px; fu.P[1] = rl->Py; fu.P[2] = rl->Pz;
      fu.Pinv[0] = rl->Px_inv; fu.Pinv[1] = rl->Py_inv; fu.Pinv[2] = rl->Pz_inv;
      fu.Xl[0] = rl->Xl; fu.Xl[1] = rl->Yl; fu.Xl[2] = rl->Zl;
      fu.tile_offset = static_cast<int32_t>(F.lattice_tiles.size());
      for (int32_t t : rl->lattice_universes) F.lattice_tiles.push_back(t);
      fu.outer = rl->outer_universe_index;
    } else if (auto* hl = dynamic_cast<const HexLattice*>(up.get())) {  // layout: include/abeille_b200.h, abl_universe
      fu.type = ABL_UNI_HEX;
      fu.N[0] = static_cast<int32_t>(hl->width); fu.N[1] = static_cast<int32_t>(hl->width); fu.N[2] = static_cast<int32_t>(hl->Nz);
      fu.pad_ = static_cast<int32_t>(hl->Nrings) | ((hl->top_ == HexLattice::Top::Flat ? 1 : 0) << 16);
      fu.P[0] = hl->pitch_; fu.P[1] = hl->sin_pi_3; fu.P[2] = hl->pitch_z_;
      fu.Pinv[0] = hl->cos_pi_6; fu.Pinv[1] = hl->sin_pi_6; fu.Pinv[2] = hl->cos_pi_3;
      fu.Xl[0] = hl->X_o; fu.Xl[1] = hl->Y_o; fu.Xl[2] = hl->Z_o;
      fu.tile_offset = static_cast<int32_t>(F.lattice_tiles.size());
      for (int32_t t : hl->lattice_universes) F.lattice_tiles.push_back(t);
      fu.outer = hl->outer_universe_index;
    } else {
      fatal_error("flatten_problem: unknown Universe subclass");
    }
    F.universes.push_back(fu);
  }
  p.nsurfaces = static_cast<int32_t>(F.surfaces.size());
  p.ncells = static_cast<int32_t>(F.cells.size());
  p.nrpn = static_cast<int32_t>(F.rpn.size());
  p.nuniverses = static_cast<int32_t>(F.universes.size());
  p.n_universe_cells = static_cast<int32_t>(F.universe_cells.size());
  p.n_lattice_tiles = static_cast<int32_t>(F.lattice_tiles.size());
  p.root_universe = universe_index(geometry::root_universe.get());
  if (F.rpn.empty()) F.rpn.push_back(0);
  if (F.lattice_tiles.empty()) F.lattice_tiles.push_back(-1);
  p.surfaces = F.surfaces.data();
  p.cells = F.cells.data();
  p.rpn = F.rpn.data();
  p.universes = F.universes.data();
  p.universe_cells = F.universe_cells.data();
  p.lattice_tiles = F.lattice_tiles.data();

  // materials: micro cross sections as MGNuclide::get_micro_xs builds them (src/mg_nuclide.cpp:394-411)
  F.delayed_offset.assign(1, 0);
  std::vector<double> majorant(G, 0.);
  for (Material* m : mats) {
    if (m->components().size() != 1) fatal_error("flatten_problem: multi-group materials hold one nuclide");
    const auto* n = dynamic_cast<const MGNuclide*>(m->components()[0].nuclide.get());
    if (!n) fatal_error("flatten_problem: the B200 backend transports multi-group problems only");
    for (size_t g = 0; g < G; g++) {
      F.Et.push_back(n->Et_[g]);
      F.Ef.push_back(n->Ef_[g]);
      F.Ea.push_back(n->Ef_[g] + (n->Ea_[g] - n->Ef_[g]));  // fission + disappearance, as get_micro_xs adds them
      F.Es.push_back(n->Es_[g]);
      F.nu.push_back(n->nu_prmpt_[g] + n->nu_delyd_[g]);
      F.nud.push_back(n->nu_delyd_[g]);
      F.speeds.push_back(n->group_speeds_.size() > g ? n->group_speeds_[g] : 0.);
      const double xs = 0. + 1. * n->Et_[g];  // src/majorant.cpp:133-176
      if (xs > majorant[g]) majorant[g] = xs;
      std::vector<double> cc = discrete_table(n->chi_[g]), sc = discrete_table(n->Ps_[g]);
      cc.resize(G, 1.0);
      sc.resize(G, 1.0);
      F.chi_cdf.insert(F.chi_cdf.end(), cc.begin(), cc.end());
      F.scatter_cdf.insert(F.scatter_cdf.end(), sc.begin(), sc.end());
      for (size_t o = 0; o < G; o++) {
        const MGAngleDistribution& a = n->angle_dists_[g][o];
        int32_t off = -1;  // identical tables (the isotropic default above all) are stored once
        for (const auto& prev : F.angle) {
          if (static_cast<size_t>(prev.n) != a.mu_.size()) continue;
          const long po = prev.offset;
          if (std::equal(a.mu_.begin(), a.mu_.end(), F.amu.begin() + po) && std::equal(a.pdf_.begin(), a.pdf_.end(), F.apdf.begin() + po) &&
              std::equal(a.cdf_.begin(), a.cdf_.end(), F.acdf.begin() + po)) {
            off = prev.offset;
            break;
          }
        }
        if (off < 0) {
          off = static_cast<int32_t>(F.amu.size());
          F.amu.insert(F.amu.end(), a.mu_.begin(), a.mu_.end());
          F.apdf.insert(F.apdf.end(), a.pdf_.begin(), a.pdf_.end());
          F.acdf.insert(F.acdf.end(), a.cdf_.begin(), a.cdf_.end());
        }
        F.angle.push_back(abl_angle_table{off, static_cast<int32_t>(a.mu_.size())});
      }
    }
    std::vector<double> dc = discrete_table(n->P_delayed_group);
    dc.resize(n->P_delayed_group.size(), 1.0);
    F.dcdf.insert(F.dcdf.end(), dc.begin(), dc.end());
    F.dlambda.insert(F.dlambda.end(), n->delayed_group_decay_constants.begin(), n->delayed_group_decay_constants.end());
    F.delayed_offset.push_back(static_cast<int32_t>(F.dcdf.size()));
    F.fissile.push_back(n->fissile_ ? 1 : 0);
  }
  if (F.dcdf.empty()) {
    F.dcdf.push_back(1.0);
    F.dlambda.push_back(0.0);
  }
  p.nmaterials = static_cast<int32_t>(mats.size());
  p.n_angle_points = static_cast<int32_t>(F.amu.size());
  p.xs_total = F.Et.data(); p.xs_absorption = F.Ea.data(); p.xs_fission = F.Ef.data(); p.xs_elastic = F.Es.data();
  p.nu_total = F.nu.data(); p.nu_delayed = F.nud.data(); p.speeds = F.speeds.data();
  p.chi_cdf = F.chi_cdf.data(); p.scatter_cdf = F.scatter_cdf.data();
  p.angle = F.angle.data(); p.angle_mu = F.amu.data(); p.angle_pdf = F.apdf.data(); p.angle_cdf = F.acdf.data();
  p.delayed_offset = F.delayed_offset.data(); p.delayed_cdf = F.dcdf.data(); p.delayed_lambda = F.dlambda.data();
  p.fissile = F.fissile.data();
  // sampling cross section: the majorant (delta tracking), or ratio x majorant (carter tracking, src/carter_tracker.cpp:60-75)
  F.smp = majorant;
  if (settings::tracking == settings::TrackingMode::CARTER_TRACKING) {
    if (settings::sample_xs_ratio.size() != G) fatal_error("The length of sampling-xs-ratio must be equal to ngroups.");
    for (size_t g = 0; g < G; g++) F.smp[g] = majorant[g] * settings::sample_xs_ratio[g];
  }
  p.sampling_xs = F.smp.data();

  // mesh tallies scored inside transport(): collision estimators, then track-length estimators
  auto add_tally = [&](const MeshTally& t, int estimator, int quantity) {
    abl_mesh_tally ft{};
    ft.estimator = estimator;
    ft.quantity = quantity;
    ft.noise_source = 0;
    ft.N[0] = static_cast<int32_t>(t.Nx); ft.N[1] = static_cast<int32_t>(t.Ny); ft.N[2] = static_cast<int32_t>(t.Nz);
    ft.n_energy_bins = static_cast<int32_t>(t.energy_bounds.size()) - 1;
    ft.ebounds_offset = static_cast<int32_t>(F.tally_eb.size());
    F.tally_eb.insert(F.tally_eb.end(), t.energy_bounds.begin(), t.energy_bounds.end());
    ft.low[0] = t.r_low.x(); ft.low[1] = t.r_low.y(); ft.low[2] = t.r_low.z();
    ft.hi[0] = t.r_hi.x(); ft.hi[1] = t.r_hi.y(); ft.hi[2] = t.r_hi.z();
    ft.net_weight = t.net_weight;
    F.tallies.push_back(ft);
  };
  for (const auto& t : tallies.collision_mesh_tallies_) add_tally(*t, ABL_EST_COLLISION, static_cast<int>(t->quantity));
  for (const auto& t : tallies.track_length_mesh_tallies_) add_tally(*t, ABL_EST_TRACK_LENGTH, static_cast<int>(t->quantity));
  if (F.tally_eb.empty()) F.tally_eb.push_back(0.);
  p.ntallies = static_cast<int32_t>(F.tallies.size());
  p.n_tally_energy_bounds = static_cast<int32_t>(F.tally_eb.size());
  p.tallies = F.tallies.data();
  p.tally_energy_bounds = F.tally_eb.data();

  // noise sources (src/noise_maker.cpp:39-58): vibrations, then oscillations, as the NoiseMaker holds them
  if (noise_maker) {
    for (const auto& vs : noise_maker->vibration_noise_sources_) {
      const auto* fv = dynamic_cast<const FlatVibrationNoiseSource*>(vs.get());
      if (!fv) fatal_error("flatten_problem: unknown VibrationNoiseSource subclass");
      abl_noise_source ns{};
      ns.type = ABL_NOISE_FLAT_VIBRATION;
      ns.low[0] = fv->low_.x(); ns.low[1] = fv->low_.y(); ns.low[2] = fv->low_.z();
      ns.hi[0] = fv->hi_.x(); ns.hi[1] = fv->hi_.y(); ns.hi[2] = fv->hi_.z();
      ns.angular_frequency = fv->w0_;
      ns.basis = static_cast<int32_t>(fv->basis_);
      ns.material_pos = material_index(fv->material_pos_.get());
      ns.material_neg = material_index(fv->material_neg_.get());
      F.noise_sources.push_back(ns);
    }
    for (const auto& os : noise_maker->oscillation_noise_sources_) {
      const auto* so = dynamic_cast<const SquareOscillationNoiseSource*>(os.get());
      if (!so) fatal_error("flatten_problem: unknown OscillationNoiseSource subclass");
      abl_noise_source ns{};
      ns.type = ABL_NOISE_SQUARE_OSCILLATION;
      ns.low[0] = so->low_.x(); ns.low[1] = so->low_.y(); ns.low[2] = so->low_.z();
      ns.hi[0] = so->hi_.x(); ns.hi[1] = so->hi_.y(); ns.hi[2] = so->hi_.z();
      ns.angular_frequency = so->w0_;
      ns.eps_total = so->eps_t_; ns.eps_fission = so->eps_f_; ns.eps_scatter = so->eps_s_;
      F.noise_sources.push_back(ns);
    }
    p.n_noise_sources = static_cast<int32_t>(F.noise_sources.size());
    p.noise_sources = F.noise_sources.data();
  }

  // the cancelator's kind (the mesh itself stays on the reference's side of the boundary)
  if (const auto* be = dynamic_cast<const BasicExactMGCancelator*>(cancelator)) {
    p.cancelator.present = 1;
    p.cancelator.kind = ABL_CANCEL_BASIC_EXACT;
    p.cancelator.N[0] = static_cast<int32_t>(be->hash_fn.shape[0]);
    p.cancelator.N[1] = static_cast<int32_t>(be->hash_fn.shape[1]);
    p.cancelator.N[2] = static_cast<int32_t>(be->hash_fn.shape[2]);
    p.cancelator.low[0] = be->r_low.x(); p.cancelator.low[1] = be->r_low.y(); p.cancelator.low[2] = be->r_low.z();
    p.cancelator.hi[0] = be->r_hi.x(); p.cancelator.hi[1] = be->r_hi.y(); p.cancelator.hi[2] = be->r_hi.z();
    p.cancelator.beta = static_cast<int32_t>(be->beta_mode);
    p.cancelator.sobol = be->use_sobol ? 1 : 0;
    p.cancelator.n_samples = static_cast<int32_t>(std::min<uint32_t>(be->N_SAMPLES, 64));
  } else if (const auto* ex = dynamic_cast<const ExactMGCancelator*>(cancelator)) {
    using Key = ExactMGCancelator::Key;  // (its mesh lives in static members)
    p.cancelator.present = 1;
    p.cancelator.kind = ABL_CANCEL_EXACT;
    for (int k = 0; k < 3; k++) p.cancelator.N[k] = static_cast<int32_t>(Key::shape[static_cast<size_t>(k)]);
    p.cancelator.low[0] = Key::r_low.x(); p.cancelator.low[1] = Key::r_low.y(); p.cancelator.low[2] = Key::r_low.z();
    p.cancelator.hi[0] = Key::r_hi.x(); p.cancelator.hi[1] = Key::r_hi.y(); p.cancelator.hi[2] = Key::r_hi.z();
    p.cancelator.n_samples = static_cast<int32_t>(std::min<uint32_t>(ex->N_SAMPLES, 64));
    F.exact_group_bins.assign(1, static_cast<int32_t>(Key::group_bins.size()));
    for (const auto& b : Key::group_bins) {
      F.exact_group_bins.push_back(static_cast<int32_t>(b.size()));
      for (std::size_t g : b) F.exact_group_bins.push_back(static_cast<int32_t>(g));
    }
    for (Material* m : mats) {
      const auto* n = static_cast<const MGNuclide*>(m->components()[0].nuclide.get());
      for (size_t g = 0; g < G; g++) F.chi_pdf.insert(F.chi_pdf.end(), n->chi_[g].begin(), n->chi_[g].end());
    }
    p.chi_pdf = F.chi_pdf.data();
    p.exact_group_bins = F.exact_group_bins.data();
    p.n_exact_group_bins = static_cast<int32_t>(F.exact_group_bins.size());
    p.chi_matrix = settings::chi_matrix ? 1 : 0;
  }
}

}  // namespace abl_integration
