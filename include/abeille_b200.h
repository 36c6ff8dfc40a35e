/* abeille_b200.h -- C ABI of the B200 transport backend (libabeille_b200.so).
 *
 * This is the drop-in boundary for Abeille's particle-transport hot path.  The one
 * reference interface it replaces is
 *
 *   class Transporter {
 *     virtual std::vector<BankedParticle> transport(std::vector<Particle>& bank, bool noise,
 *         std::vector<BankedParticle>* noise_bank, const NoiseMaker* noise_maker) = 0; };
 *                                   (reference include/simulation/transporter.hpp:39-47)
 *
 * together with the state that call reads implicitly from globals (settings::*, geometry::*,
 * materials, tallies->kcol(); reference src/transporter.cpp:371,399, include/utils/settings.hpp:54-128)
 * and the side effects it has on the Tallies object (src/tallies.cpp:100-140, src/mesh_tally.cpp:121-152).
 *
 * Plain C: pointers and sizes only.  Two flavours of every data-path entry point:
 *   abl_xxx         HOST buffers (what a reference-side adapter binds; copies are inside the call)
 *   abl_xxx_device  DEVICE buffers owned by the caller (e.g. torch tensors), asynchronous on `stream`
 * All functions return ABL_OK (0) or a negative abl_status; abl_last_error() gives the text.
 * A handle is bound to one CUDA device; calls on one handle must be serialised by the caller
 * (the reference's transport() is not re-entrant either).
 */
#ifndef ABEILLE_B200_H
#define ABEILLE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct abl_context* abl_handle;

typedef enum abl_status {
  ABL_OK = 0,
  ABL_ERR_INVALID = -1,       /* bad argument / inconsistent tables                                   */
  ABL_ERR_CUDA = -2,          /* CUDA runtime error (no device, out of memory, launch failure)        */
  ABL_ERR_BANK_OVERFLOW = -3, /* more fission sites / secondaries than the provided capacity          */
  ABL_ERR_LOST = -4,          /* fatal: particle lost after reflection / crossing / resurrection      */
  ABL_ERR_MAJORANT = -5,      /* fatal: total xs exceeded the majorant (delta_tracker.cpp:174-180)    */
  ABL_ERR_GEOMETRY = -6,      /* geometry nesting deeper than ABL_MAX_PADS / malformed tables         */
  ABL_ERR_UNSUPPORTED = -7,
  ABL_ERR_TIMEOUT = -8        /* a history kernel ran past its deadline (ABEILLE_B200_KERNEL_TIMEOUT_S, default 120 s) and was wound down */
} abl_status;

#define ABL_MAX_PADS 10   /* geometry stack depth (reference reserves 10: tracker.hpp:46) */
#define ABL_MAX_FRAMES 5  /* nested local coordinate frames (root + lattice levels)       */

/* ---- enums mirrored from the reference --------------------------------------------------------- */
enum { ABL_MODE_K_EIGENVALUE = 0, ABL_MODE_NOISE = 1,
       ABL_MODE_FIXED_SOURCE = 2,    /* settings.hpp; FIXED_SOURCE: fission neutrons continue their history as secondaries
                                        and transport returns an empty bank (src/transporter.cpp:374-379,460-463) */
       ABL_MODE_BRANCHLESS = 3 };    /* BRANCHLESS_K_EIGENVALUE: Transporter::branchless_collision_mat / _iso
                                        (src/transporter.cpp:95-267); the flavour is abl_problem::branchless_flags */
enum { ABL_BRANCHLESS_MATERIAL = 1,    /* settings::branchless_material (collision on the material, else on the isotope) */
       ABL_BRANCHLESS_SPLITTING = 2 }; /* settings::branchless_splitting (split when |wgt| >= wgt_split)                 */
enum { ABL_TRACK_SURFACE = 0, ABL_TRACK_DELTA = 1, ABL_TRACK_CARTER = 2,
       ABL_TRACK_IMPLICIT_LEAKAGE = 3 };  /* parser.cpp:408-420,889-911; 3 replaces ImplicitLeakageDeltaTracker::transport
                                             (src/implicit_leakage_delta_tracker.cpp:73-263) */
enum { ABL_BC_VACUUM = 0, ABL_BC_REFLECTIVE = 1, ABL_BC_NORMAL = 2 };           /* surface.hpp:34      */
enum {
  ABL_SURF_XPLANE = 0, ABL_SURF_YPLANE, ABL_SURF_ZPLANE, ABL_SURF_PLANE,
  ABL_SURF_XCYL, ABL_SURF_YCYL, ABL_SURF_ZCYL, ABL_SURF_CYL, ABL_SURF_SPHERE
};
enum { ABL_UNI_CELLS = 0, ABL_UNI_RECT = 1, ABL_UNI_HEX = 2 };  /* CellUniverse, RectLattice, HexLattice */
enum { ABL_EST_COLLISION = 0, ABL_EST_TRACK_LENGTH = 1, ABL_EST_SOURCE = 2 };
enum {
  ABL_Q_FLUX = 0, ABL_Q_TOTAL, ABL_Q_ELASTIC, ABL_Q_ABSORPTION, ABL_Q_FISSION, ABL_Q_MT,
  ABL_Q_REAL_FLUX, ABL_Q_IMAG_FLUX, ABL_Q_SOURCE, ABL_Q_REAL_SOURCE, ABL_Q_IMAG_SOURCE
};
/* RPN operator tokens of cell regions (cell.hpp:43-49); surface tokens are +-(surface index + 1) */
#define ABL_OP_UNION (INT32_MAX - 4)
#define ABL_OP_INTERSECTION (INT32_MAX - 3)
#define ABL_OP_COMPLEMENT (INT32_MAX - 2)

/* ---- flattened immutable problem tables (uploaded once by abl_create) ----------------------------- */
typedef struct abl_surface {
  int32_t type;  /* ABL_SURF_*                                                                        */
  int32_t bc;    /* ABL_BC_*                                                                          */
  double p[8];   /* planes: x0|y0|z0 or A,B,C,D; axis cylinders: c0,c1,R; sphere x0,y0,z0,R;
                    general cylinder: x0,y0,z0,1-u0^2,1-v0^2,1-w0^2,R (cylinder.cpp:28-58)            */
} abl_surface;

typedef struct abl_cell {
  int32_t rpn_offset, rpn_len; /* slice of abl_problem.rpn                                            */
  int32_t simple;              /* 1: pure intersection list (cell.cpp:144-158)                        */
  int32_t vac_or_refl;         /* touches a vacuum/reflective surface (cell.cpp:203-246)              */
  int32_t fill_universe;       /* universe index, or -1 when filled with a material                   */
  int32_t material;            /* material index, or -1                                               */
} abl_cell;

typedef struct abl_universe {
  int32_t type;                 /* ABL_UNI_*                                                          */
  int32_t has_bc;               /* cell_universe.cpp:30-41, lattice.cpp:45-52                         */
  int32_t cell_offset, ncells;  /* ABL_UNI_CELLS: slice of abl_problem.universe_cells (cell indices)  */
  int32_t N[3];                 /* ABL_UNI_RECT: tiles along x,y,z                                    */
  int32_t tile_offset;          /* slice of abl_problem.lattice_tiles: universe index or -1, linear
                                   index nz*Nx*Ny + nx*Ny + ny (rect_lattice.cpp:293-301)             */
  int32_t outer;                /* outer universe index or -1                                         */
  int32_t pad_;                 /* ABL_UNI_HEX: nrings | top << 16 (top: 0 pointy, 1 flat)            */
  double P[3], Pinv[3], Xl[3];  /* pitch, 1/pitch, lower corner origin - N*P/2 (rect_lattice.cpp:33-52) */
  /* ABL_UNI_HEX (src/hex_lattice.cpp:30-56): N = {width, width, nz} with width = 2 (nrings - 1) + 1, tiles in the
   * reference's linear_index order nz*width*width + (r + width/2)*width + (q + width/2), -1 outside the hexagon;
   * P[0] = pitch, P[2] = pitch_z, Xl = origin of the centre hexagon (X_o, Y_o, Z_o), and the four constants the reference
   * evaluates with std::cos / std::sin at construction (hex_lattice.hpp:60-63): Pinv[0] = cos(pi/6), Pinv[1] = sin(pi/6),
   * Pinv[2] = cos(pi/3), P[1] = sin(pi/3) */
} abl_universe;

typedef struct abl_angle_table { /* linearised mu distribution of one (g_in,g_out) pair               */
  int32_t offset, n;             /* slice of angle_mu / angle_pdf / angle_cdf                          */
} abl_angle_table;

typedef struct abl_mesh_tally {
  int32_t estimator, quantity;   /* ABL_EST_*, ABL_Q_*                                                 */
  int32_t noise_source;          /* source tally scored from the noise-source bank                    */
  int32_t N[3];
  int32_t n_energy_bins;
  int32_t ebounds_offset;        /* slice (n_energy_bins+1 values) of abl_problem.tally_energy_bounds  */
  double low[3], hi[3];
  double net_weight;             /* tallies->total_weight (parser.cpp:870-871)                         */
} abl_mesh_tally;

/* direction distribution of a source (src/direction_distribution.cpp:36-56) */
#define ABL_DIR_ISOTROPIC 0      /* src/isotropic.cpp:28-36: two draws                                 */
#define ABL_DIR_MONO 1           /* src/mono_directional.cpp, include/simulation/mono_directional.hpp:38: no draw */
#define ABL_DIR_CONE 2           /* src/cone.cpp:31-42: mu uniform in [cos(aperture), 1], phi, rotate_direction */

/* energy distribution of a source (src/energy_distribution.cpp:36-58); `tabulated` needs PapillonNDL's PCTable and is not provided */
#define ABL_EN_MONO 0            /* src/mono_energetic.cpp: no draw                                    */
#define ABL_EN_MAXWELLIAN 1      /* src/maxwellian.cpp:34-42: three draws, -a (log xi1 + log xi2 cos^2(pi xi3 / 2)) */
#define ABL_EN_WATT 2            /* src/watt.cpp:42-47: a Maxwellian(a) sample w and one more draw     */

typedef struct abl_source {      /* box | point; isotropic | mono-directional | cone; mono-energetic | maxwellian | watt (source.cpp:44-90) */
  double weight;
  int32_t fissile_only, is_box;
  double low[3], hi[3];          /* point: low == position                                             */
  double energy;
  int32_t direction_kind, pad_;  /* ABL_DIR_*                                                          */
  double dir[3];                 /* mono-directional / cone axis, NORMALISED as Direction(x, y, z) does (direction.hpp:37-42) */
  double cos_aperture;           /* cone: std::cos(aperture), taken once on the host (cone.cpp:31-32)  */
  int32_t energy_kind, pad2_;    /* ABL_EN_*; `energy` is the mono-energetic value                     */
  double en_a, en_b;             /* maxwellian: a; watt: a, b                                          */
} abl_source;

typedef struct abl_mesh3 {       /* entropy mesh (entropy.cpp) / approximate cancelator mesh           */
  int32_t present;
  int32_t N[3];
  double low[3], hi[3];
  int32_t n_energy_edges;        /* cancelator only; 0 = no energy binning                             */
  int32_t eedges_offset;         /* slice of abl_problem.tally_energy_bounds                           */
  int32_t kind;                  /* cancelator only: ABL_CANCEL_* (0 reads as approximate)             */
  int32_t beta;                  /* basic-exact cancelator: ABL_BETA_*                                 */
  int32_t sobol, n_samples;      /* ... average-f / average-g: Sobol points (else engine draws), points per bin (<= 64) */
} abl_mesh3;
enum { ABL_CANCEL_APPROXIMATE = 1, ABL_CANCEL_BASIC_EXACT = 2,
       ABL_CANCEL_EXACT = 3 };  /* src/cancelator.cpp:40-72.  EXACT: src/exact_mg_cancelator.cpp */
enum { ABL_BETA_ZERO = 0, ABL_BETA_MINIMUM = 1, ABL_BETA_AVERAGE_F = 2, ABL_BETA_AVERAGE_G = 3 };  /* BasicExactMGCancelator::BetaMode */

enum { ABL_NOISE_SQUARE_OSCILLATION = 0, ABL_NOISE_FLAT_VIBRATION = 1 };
typedef struct abl_noise_source { /* square_oscillation_noise_source.cpp:38-83, flat_vibration_noise_source.cpp:34-118 */
  double low[3], hi[3];
  double angular_frequency;
  double eps_total, eps_fission, eps_scatter; /* square oscillation                                           */
  int32_t type;                               /* ABL_NOISE_*                                                  */
  int32_t basis;                              /* flat vibration: 0 x, 1 y, 2 z (the direction of the motion)  */
  int32_t material_pos, material_neg;         /* flat vibration: material indices on the two sides            */
} abl_noise_source;

typedef struct abl_problem {
  int32_t mode, tracking, ngroups, inner_generations;
  const double* energy_bounds;  /* [ngroups+1] */
  double wgt_cutoff, wgt_survival, wgt_split, min_energy;
  uint64_t rng_seed, rng_stride;
  double w_noise, eta, keff;
  /* geometry */
  int32_t nsurfaces, ncells, nrpn, nuniverses, n_universe_cells, n_lattice_tiles, root_universe,
      branchless_flags;  /* ABL_BRANCHLESS_* bits; read when mode == ABL_MODE_BRANCHLESS */
  const abl_surface* surfaces;
  const abl_cell* cells;
  const int32_t* rpn;
  const abl_universe* universes;
  const int32_t* universe_cells;
  const int32_t* lattice_tiles;
  /* materials: one MGNuclide each, atoms_bcm = 1 (material.cpp:50-62); all arrays [nmaterials*ngroups] */
  int32_t nmaterials, n_angle_points;
  const double *xs_total, *xs_absorption, *xs_fission, *xs_elastic, *nu_total, *nu_delayed, *speeds;
  const double* chi_cdf;      /* [M*G*G] libstdc++ discrete_distribution partial sums of chi rows       */
  const double* scatter_cdf;  /* [M*G*G] same for the normalised scatter rows                          */
  const abl_angle_table* angle; /* [M*G*G] */
  const double *angle_mu, *angle_pdf, *angle_cdf; /* pools, n_angle_points each                       */
  const int32_t* delayed_offset; /* [M+1] slices of delayed_cdf / delayed_lambda                       */
  const double *delayed_cdf, *delayed_lambda;
  const int32_t* fissile;       /* [M] */
  const double* sampling_xs;    /* [G] majorant (delta) or ratio*majorant (carter); majorant.cpp:133-176 */
  /* tallies */
  int32_t ntallies, n_tally_energy_bounds;
  const abl_mesh_tally* tallies;
  const double* tally_energy_bounds;
  /* sources, entropy mesh, approximate cancelator */
  int32_t nsources, pad1_;
  const abl_source* sources;
  abl_mesh3 entropy;
  abl_mesh3 cancelator;
  /* noise sources (noise mode; noise_maker.cpp:39-58) */
  int32_t n_noise_sources, pad2_;
  const abl_noise_source* noise_sources;
  /* `type: exact` cancelator (src/exact_mg_cancelator.cpp): the normalised chi rows themselves ([M*G*G], MGNuclide::chi()), whether
   * any fissile material gave a chi matrix (settings::chi_matrix; one group counts), and the energy bins as a flat list
   * {number of bins, then per bin: its size, its groups}; NULL / 0 when the problem has no such cancelator.
   * Its mesh is `cancelator` (kind ABL_CANCEL_EXACT, n_samples). */
  const double* chi_pdf;
  const int32_t* exact_group_bins;
  int32_t n_exact_group_bins, chi_matrix;
} abl_problem;

/* ---- banks --------------------------------------------------------------------------------------
 * Structure-of-arrays view of a particle bank (input) or a fission-site bank (output).
 * Particle bank  (Particle, particle.hpp:68-243):        id_a = history id, id_b = family id,
 *                                                         id_c = pcg32 state or NULL (=> seed, advance(stride*history id))
 * Fission bank   (BankedParticle, particle.hpp:38-66):   id_a = parent_history_id, id_b = parent_daughter_id,
 *                                                         id_c = family_id
 * wgt2 may be NULL outside noise mode (treated as 0 on input, not written on output).               */
typedef struct abl_bank {
  uint64_t n; /* particles (input) or capacity (output) */
  double *x, *y, *z, *ux, *uy, *uz, *E, *wgt, *wgt2;
  uint64_t *id_a, *id_b, *id_c;
} abl_bank;

typedef struct abl_gen_params {
  double k_col;       /* tallies->kcol() of the previous generation (transporter.cpp:370-371)          */
  double keff;        /* noise mode (transporter.cpp:399)                                              */
  int32_t converged;  /* settings::converged: mesh tallies score only when set (tallies.hpp:49-63)     */
  int32_t noise;      /* transport(bank, noise=true)                                                   */
  int32_t trace;      /* keep per-history integer outcomes for abl_get_trace                           */
  int32_t sample_noise_source; /* transport(bank, false, &noise_bank, &noise_maker) (noise.cpp:312-314): noise
                                  mode only; the noise particles go to the bank given to abl_transport_noise_device */
} abl_gen_params;

/* scores[6] = raw sums k_col, k_abs, k_trk, k_tot, leakage, mig_area (tallies.cpp:100-140)           */
/* counters[8] = flights, real collisions, virtual collisions, track-length bins, fission sites,
 *               boundary events, lost at birth, collision-tally scores                                */
typedef struct abl_trace { /* per history of the last traced transport call, arrays of length n      */
  uint32_t *flights, *real, *virt, *fission;
  uint64_t *hash, *rng_state;
} abl_trace;

/* ---- lifetime ------------------------------------------------------------------------------------ */
int abl_create(const abl_problem* problem, int device, abl_handle* out);
void abl_destroy(abl_handle h);
const char* abl_last_error(abl_handle h); /* h may be NULL: error of the last failed abl_create        */
int abl_device_info(abl_handle h, int* sm_count, int* cc_major, int* cc_minor, uint64_t* kernel_launches);
/* device time (CUDA events on the launching stream) of the history kernel of the last transport call, and its grid */
int abl_last_transport_kernel(abl_handle h, float* milliseconds, int* grid_blocks, int* block_threads);

/* ---- Transporter::transport, host buffers (the reference-facing entry point) ------------------------ */
int abl_transport(abl_handle h, const abl_bank* bank, const abl_gen_params* params, abl_bank* fission_out,
                  uint64_t* n_fission, double scores[6], uint64_t counters[8]);
/* The same with the noise-source bank (transport(bank, false, &noise_bank, &noise_maker), noise.cpp:312-314): host
 * buffers in, fission bank and noise bank out (noise_out->n = capacity, *n_noise = count; wgt2 is written). */
int abl_transport_noise(abl_handle h, const abl_bank* bank, const abl_gen_params* params, abl_bank* fission_out,
                        uint64_t* n_fission, abl_bank* noise_out, uint64_t* n_noise, double scores[6], uint64_t counters[8]);
/* abl_transport in two halves, for a caller that normalises the fission bank before it uses it, as PowerIterator::run does
 * (normalize_weights and the hand-out of fresh history ids, src/power_iterator.cpp:397-399,538-569).  begin: host bank in,
 * kernels, the fission bank stays on the device; returns its size, the scores and weight_stats = {particles with positive
 * weight, with negative weight, sum of the positive weights, minus the sum of the negative ones} (what normalize_weights
 * sums; with MPI / several GPUs the caller adds these over the ranks).  finish: wgt *= weight_factor, id_a = first_history_id
 * + row, id_b = family id, on the device; then the bank is copied to fission_out (host; x y z ux uy uz E wgt id_a id_b).
 * Same bytes over PCIe as abl_transport, no pass over the bank on the host. */
int abl_transport_begin(abl_handle h, const abl_bank* bank, const abl_gen_params* params, uint64_t capacity,
                        uint64_t* n_fission, double scores[6], uint64_t counters[8], double weight_stats[4]);
int abl_transport_finish(abl_handle h, double weight_factor, uint64_t first_history_id, abl_bank* fission_out);
int abl_get_trace(abl_handle h, uint64_t n, abl_trace* out);

/* ---- Transporter::transport, device buffers (bank stays resident in HBM) ---------------------------- */
int abl_transport_device(abl_handle h, const abl_bank* bank_dev, const abl_gen_params* params,
                         abl_bank* fission_dev, uint64_t* n_fission, double scores[6], uint64_t counters[8],
                         void* stream);

/* Transporter::transport with the noise-source bank (noise.cpp:305-318): as abl_transport_device; when
 * params->sample_noise_source is set, every real collision inside a noise source also samples NoiseMaker::sample_noise_source
 * (noise_maker.cpp:277-445) and the noise particles are written to noise_dev in bank order (noise_dev->n = capacity on
 * entry, *n_noise = count).  noise_dev may be NULL when nothing is sampled. */
int abl_transport_noise_device(abl_handle h, const abl_bank* bank_dev, const abl_gen_params* params, abl_bank* fission_dev,
                               uint64_t* n_fission, abl_bank* noise_dev, uint64_t* n_noise, double scores[6],
                               uint64_t counters[8], void* stream);
/* sum over the bank of sqrt(wgt^2 + wgt2^2), and wgt /= d, wgt2 /= d (Noise::noise_simulation, noise.cpp:438-456) */
int abl_bank_weight_magnitude_device(abl_handle h, const abl_bank* bank_dev, double* sum, void* stream);
int abl_bank_divide_weights_device(abl_handle h, abl_bank* bank_dev, double divisor, void* stream);

/* ---- mesh tallies: MeshTally::record_generation / clear_generation / write (mesh_tally.cpp:121-206) - */
int abl_tally_count(abl_handle h);
int abl_tally_shape(abl_handle h, int tally, uint64_t shape4[4]); /* Ne, Nx, Ny, Nz */
int abl_tallies_record(abl_handle h, double multiplier);
int abl_tallies_clear(abl_handle h);
/* which: 0 = tally_gen, 1 = tally_avg, 2 = tally_var, 3 = std = sqrt(var/g) */
int abl_tally_fetch(abl_handle h, int tally, int which, double* out_host);
int abl_tally_device_ptr(abl_handle h, int tally, int which, double** out_dev, uint64_t* n);

/* ---- inter-generation bank pipeline on the device (power_iterator.cpp:341-404,538-586) ---------------- */
/* Source sampling (simulation.cpp:55-77, Source::generate_particle source.cpp:44-90): n particles with history ids first_id.., written
 * to bank_dev.  Per particle, on its own stream: the source by weight, the direction (abl_source.direction_kind), the energy once and
 * then again until it lies inside (min_energy, max_energy) -- ABL_ERR_INVALID after 201 redraws --, the position until it is inside the
 * geometry (and, fissile-only, inside a fissile material: ABL_ERR_INVALID after 201 attempts). */
int abl_sample_source_device(abl_handle h, uint64_t n, uint64_t first_history_id, abl_bank* bank_dev, void* stream);
/* stats[6] = Npos, Nneg, Wpos, Wneg (unnormalised), and after scaling Wpos', Wneg'                      */
int abl_bank_weight_stats_device(abl_handle h, const abl_bank* bank_dev, double stats[4], void* stream);
/* Weighted moments of a device bank about `origin`: moments = { sum w, sum w (x - ox), sum w (y - oy), sum w (z - oz), sum w |r - o|^2 }.
 * Replaces PowerIterator::compute_pair_dist_sqrd (src/power_iterator.cpp:637-663, settings: pair-distance-sqrd), a double sum over all
 * pairs of the normalised fission bank: sum_ij w_i w_j |r_i - r_j|^2 / (2 W^2) = sum_i w_i |r_i - c|^2 / W about the weighted centroid
 * c, i.e. one call about any origin for c and one about c.  The sums of several ranks add. */
int abl_bank_moments_device(abl_handle h, const abl_bank* bank_dev, const double origin[3], double moments[5], void* stream);
int abl_bank_scale_weights_device(abl_handle h, abl_bank* bank_dev, double factor, void* stream);
/* fission bank -> next particle bank: history ids first_id + i, family kept, rng from seed/stride       */
int abl_bank_to_particles_device(abl_handle h, abl_bank* bank_dev, uint64_t first_history_id, void* stream);
/* Shannon-entropy binning (entropy.cpp:32-60): bins_dev[Nx*Ny*Nz] += w, total_dev[0] += w               */
int abl_entropy_bin_device(abl_handle h, const abl_bank* bank_dev, double* bins_dev, double* total_dev, void* stream);
/* SourceMeshTally::score_source for every source tally (source_mesh_tally.cpp:30-78)                    */
int abl_score_source_device(abl_handle h, const abl_bank* bank_dev, int noise_source, void* stream);
/* ApproximateMeshCancelator (approximate_mesh_cancelator.cpp:97-190), weights replaced in place        */
int abl_cancel_device(abl_handle h, abl_bank* bank_dev, void* stream);
/* The same in two steps, for several GPUs: accumulate this GPU's bank into the dense bins, all-reduce (sum) the five bin
 * arrays across the GPUs (the reference gathers the whole bank on the master instead, power_iterator.cpp:751-777), apply
 * (which also re-zeroes the touched bins).  sums_dev = positive / negative sums of wgt, then of wgt2. */
int abl_cancel_accumulate_device(abl_handle h, const abl_bank* bank_dev, void* stream);
int abl_cancel_apply_device(abl_handle h, abl_bank* bank_dev, void* stream);
int abl_cancel_bins_device(abl_handle h, double* sums_dev[4], uint32_t** count_dev, uint64_t* nbins);

/* BasicExactMGCancelator (src/basic_exact_mg_cancelator.cpp; PowerIterator::perform_regional_cancellation, src/power_iterator.cpp:
 * 751-777) on the fission bank the LAST transport call of this handle produced: with an exact cancelator in the problem the
 * transport kernels keep, per fission site, the parent's previous position and the sampling cross section of its flight
 * (BankedParticle::parents_previous_position / Esmp_parent, particle.hpp:52-57) in a side table of the handle, in bank order.
 * Weights are reduced in place, the uniform particles the cancelled weight turns into are appended (bank_dev->n grows, at most
 * `capacity` rows), rng2 = {state, increment} of settings::rng is advanced as the reference advances it.  All four beta modes.
 * abl_parent_info_download hands the side table to a caller that runs the reference's own cancelator on the host.           */
int abl_cancel_exact_device(abl_handle h, abl_bank* bank_dev, uint64_t capacity, uint64_t rng2[2], void* stream);
int abl_parent_info_download(abl_handle h, uint64_t n, double* x, double* y, double* z, double* esmp);
/* ... and what `type: exact` reads on top (src/exact_mg_cancelator.cpp:319-327): parents_previous_direction, parents_previous_
 * previous_energy (the parent's energy before its last scatter), parents_previous_energy (its energy at the fission) and
 * parents_previous_was_virtual (0 / 1), so that the reference's ExactMGCancelator too runs unchanged over the adapter.     */
int abl_parent_state_download(abl_handle h, uint64_t n, double* ux, double* uy, double* uz, double* e_before_last_scatter, double* e_parent,
                              double* was_virtual);

/* ---- device memory helpers for callers that do not link a CUDA runtime themselves ------------------------ */
int abl_bank_alloc_device(abl_handle h, uint64_t capacity, abl_bank* out_dev); /* all 12 arrays, out_dev->n = capacity */
int abl_bank_free_device(abl_handle h, abl_bank* bank_dev);
int abl_bank_upload(abl_handle h, const abl_bank* host, abl_bank* bank_dev);   /* copies host->n entries, sets bank_dev->n */
int abl_bank_download(abl_handle h, const abl_bank* bank_dev, uint64_t n, abl_bank* host);
/* dst row k = src row rows_host[k] (every column), with wgt = wgts_host[k] when wgts_host is given: what a comb or any other
 * selection of the bank made on the host from the weights alone amounts to (BranchlessPowerIterator::comb_particles,
 * src/branchless_power_iterator.cpp:592-651); the bank itself stays in HBM.  src and dst must not overlap; sets dst_dev->n. */
int abl_bank_gather_device(abl_handle h, const abl_bank* src_dev, const uint32_t* rows_host, const double* wgts_host, uint64_t n,
                           abl_bank* dst_dev, void* stream);
int abl_device_alloc(abl_handle h, uint64_t bytes, void** out_dev);            /* zero-initialised */
int abl_device_free(abl_handle h, void* dev);
int abl_device_zero(abl_handle h, void* dev, uint64_t bytes, void* stream);
int abl_device_read(abl_handle h, void* dst_host, const void* src_dev, uint64_t bytes, void* stream); /* synchronises */

/* Fission-bank capacity that holds the sites `n_particles` histories of total |weight| `sum_abs_weight` can bank at
 * fission normalisation k_col (src/transporter.cpp:370-371): a history of weight w banks at most
 * w * max_{material,group}(nu Sigma_f / Sigma_a) / k_col sites on average (implicit capture: the collision weights form
 * a geometric series), so 1.25 x that bound plus a fixed slack covers the fluctuation.  The reference's vectors grow
 * without limit; callers size their output banks with this instead of a fixed multiple of the bank.   */
uint64_t abl_fission_capacity_hint(abl_handle h, uint64_t n_particles, double sum_abs_weight, double k_col);

/* ---- probes used by the parity tests ------------------------------------------------------------------ */
/* cell / material index (or -1) of n points, fresh lookup from the root universe (geometry.cpp:43-55)   */
int abl_find_cells(abl_handle h, uint64_t n, const double* r3, const double* u3, int32_t* cell, int32_t* material);
/* first n pcg32 outputs / RNG::rand values of a history stream (particle.hpp:188-193, rng.hpp:41)        */
int abl_rng_probe(abl_handle h, uint64_t history_id, int n, uint32_t* out_u32, double* out_rand);
/* device log / sin / cos used by the kernels                                                             */
int abl_math_probe(abl_handle h, int n, const double* x, double* lg, double* sn, double* cs);
/* Surface::sign / distance / norm of surface `surface_index` of the problem at n points (e.g. src/plane.cpp:33-60,
 * src/sphere.cpp:33-80, src/cylinder.cpp:60-110): u3 are unit directions, on_surf[i] != 0 = "the particle sits on this
 * surface" (the argument of Surface::distance)                                                            */
int abl_surface_probe(abl_handle h, int surface_index, uint64_t n, const double* r3, const double* u3, const int32_t* on_surf,
                      int32_t* sign, double* distance, double* norm3);
/* Replaces the per-group sampling cross section (the majorant of delta tracking, ratio * majorant of carter tracking:
 * src/majorant.cpp:158-173, src/carter_tracker.cpp:60-75) and the quotients derived from it.  A caller that changes
 * materials between runs uses it; the parity tests use it to provoke the reference's "total xs above the majorant" fatal
 * error (src/delta_tracker.cpp:174-180).                                                                 */
int abl_set_sampling_xs(abl_handle h, const double* sampling_xs, int ngroups);

#ifdef __cplusplus
}
#endif
#endif /* ABEILLE_B200_H */
