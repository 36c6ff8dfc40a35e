/* tables.h -- device-side view of the flattened problem (immutable during a run).
 *
 * Everything the kernels read that is not particle state: CSG tables, multigroup cross sections,
 * sampling tables, tally mesh descriptors.  For the reference decks the whole set is ~12 KB
 * (C5G7: 7 surfaces, 8 cells, 10 universes, 582 lattice tiles, 7x7 groups), so it lives in L1/L2
 * and is read through the read-only path.  The struct is passed to kernels by value.
 */
#pragma once
#include <stdint.h>

#include "../../include/abeille_b200.h"
#include "rng.cuh"

namespace abl {

#define ABL_MAX_TALLIES 8
#define ABL_MAX_BC_PLANES 8

struct DevTally {
  int32_t estimator, quantity, noise_source;
  int32_t Nx, Ny, Nz, Ne;
  const double* ebounds;  // Ne+1
  double lowx, lowy, lowz, hix, hiy, hiz;
  double dx, dy, dz, dx_inv, dy_inv, dz_inv;  // mesh_tally.cpp:66-101
  double net_weight;
  double* gen;  // [Ne,Nx,Ny,Nz]
  double* avg;
  double* var;
  uint64_t size;
};

// Compiled form of a cell region for the common shapes (built once by abl_create).  The boolean it yields is
// the one Cell::is_inside (src/cell.cpp:144-201) yields; whenever a point is within SURFACE_COINCIDENT of one
// of the cell's surfaces -- where the reference breaks the tie with the flight direction -- or the particle
// sits on a surface (token != 0), the generic evaluator is used instead.
enum { CF_GENERIC = 0, CF_BOX = 1, CF_ZCYL = 2 };
struct alignas(16) CellFast {
  double a[6];     // CF_BOX: lo/hi per axis (x0 < x < x1 ...), +-inf when unbounded; CF_ZCYL: x0, y0, R*R
  int32_t kind;
  int32_t sense;   // CF_ZCYL: -1 inside the cylinder, +1 outside
  int32_t pad_[2];
};

struct DevNoiseSrc {  // a noise source region (box) with the run's frequency already matched against its own
  double low[3], hi[3];
  // square oscillation (square_oscillation_noise_source.cpp): gated amplitude factors
  double eps_t;               // dEt = eps_t * Sigma_t * pi
  double eps_f_pi, eps_s_pi;  // dEf/Ef = eps_f * pi, dEs/Es = eps_s * pi
  int32_t on;                 // the run's noise frequency is the source's fundamental (|n| == 1 within 1 %)
  // flat vibration (flat_vibration_noise_source.cpp): interface at x0 along `basis`, half width eps, materials on the
  // two sides, harmonic n = round(w / w0) of the run's frequency (0: no component, gate |n w0 - w| / w <= 0.01)
  int32_t vibration, basis, mat_pos, mat_neg, harmonic;
  double x0, eps;
};

struct DevMesh3 {
  int32_t present, Nx, Ny, Nz, Ne;
  int32_t kind, beta, sobol, nsamples;  // cancelator only: ABL_CANCEL_*, ABL_BETA_*, Sobol points or engine draws, points per bin
  const double* eedges;  // Ne+1 or null
  double lowx, lowy, lowz, hix, hiy, hiz, dx, dy, dz;
};

struct DevProblem {
  int32_t mode, tracking, G, inner_generations;
  const double* ebounds;  // G+1
  double wgt_cutoff, wgt_survival, wgt_split, min_energy;
  uint64_t seed_state;  // pcg_seed_state(rng_seed)
  uint64_t stride;
  double w_noise, eta;
  const JumpTable* jump;
  // geometry
  int32_t nsurfaces, ncells, nuniverses, root;
  const abl_surface* surfaces;
  const abl_cell* cells;
  const int32_t* rpn;
  const abl_universe* universes;
  const int32_t* ucells;
  const int32_t* tiles;
  const CellFast* cellfast;  // [ncells]
  // Boundary-condition surfaces when ALL of them are axis planes that are only ever evaluated in the global frame
  // (api.cu: boundary_planes): |p0 - r[axis]| of the nearest one is a lower bound of the distance to any boundary condition
  // in any direction, which lets the surface tracker skip the boundary-condition search of a flight that ends far inside
  // (geom.cuh: cursor_nearest_boundary_lazy).  n_bc_planes = 0: no such bound.
  int32_t exact_cancel;  // the problem has an exact cancelator: the per-lane kernel records what it reads (transport.cuh Hist::rprev / esmp)
  int32_t branchless;  // ABL_BRANCHLESS_* bits (mode == ABL_MODE_BRANCHLESS)
  int32_t has_hex;  // the geometry holds a hexagonal lattice (the fixed-shape kernel builds leave those branches out)
  int32_t n_bc_planes;
  int32_t bc_axis[ABL_MAX_BC_PLANES];
  double bc_p0[ABL_MAX_BC_PLANES];
  // materials [M*G]
  int32_t M;
  const double *Et, *Ea, *Ef, *Es, *nu, *nud, *speed;
  const double *chi_cp, *ps_cp;  // [M*G*G]
  const double* chi_pdf;         // [M*G*G] normalised chi rows (type: exact cancelator) or null
  const int32_t* egb;            // its energy bins: {number of bins, per bin: size, groups} or null
  int32_t chi_matrix;            // settings::chi_matrix
  const abl_angle_table* angle;  // [M*G*G]
  const double *amu, *apdf, *acdf;
  const int32_t* dg_off;  // [M+1]
  const double *dg_cp, *dg_lambda;
  const int32_t* fissile;
  const double* smp;  // [G] sampling xs (majorant or ratio*majorant)
  // quotients of table entries, evaluated once on the host with the same IEEE division the kernels would use
  // (bit-identical): they take three fp64 divisions out of every flight / collision of the history kernel
  const double* real_frac;  // [M*G] Et / smp[g]: probability that a tentative collision is real (delta_tracker.cpp:182)
  const double* surv_frac;  // [M*G] 1 - Ea / Et: implicit-capture weight factor (transporter.cpp:295-298)
  const double* inv_score;  // [ntallies*M*G] 1 / (Et * net_weight): the collision-estimator score (collision_mesh_tally.cpp:35-40)
  // tallies
  int32_t ntallies, n_coll_tallies, n_tl_tallies;
  DevTally tally[ABL_MAX_TALLIES];
  const DevTally* tally_dev;
  const int32_t* tally_gbin;  // [ntallies*G] energy bin of each tally for E = mid-point of group g (-1: none)
  const double* gmid;         // [G] group mid-points 0.5*(b[g]+b[g+1])  // the same descriptors in global memory (for non-inlined scorers)
  // sources
  int32_t nsources;
  const abl_source* sources;
  const double* source_cp;  // discrete table over source weights (nsources >= 2)
  DevMesh3 entropy, cancel;
  // noise sources
  int32_t n_noise_src;
  const DevNoiseSrc* noise_src;
};

}  // namespace abl
