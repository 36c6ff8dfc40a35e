/* api.cu -- the C ABI of libabeille_b200.so (declared in include/abeille_b200.h).
 *
 * Owns the device-side copy of the flattened problem, the scratch buffers of the transport
 * kernel and the mesh-tally arrays.  No CPU fallback exists: without a CUDA device abl_create fails.
 */
#define ABL_TABLES_GLOBAL 1  // this unit's kernels read the tables from global memory (detmath.cuh: ldt)
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <functional>
#include <vector>

#include "bank_ops.cuh"
#include "kernel_entry.h"

using namespace abl;

namespace {

struct DevSmall {  // small per-call block, zeroed before every transport launch
  unsigned long long ticket, n_sites;
  double scores[6];
  unsigned long long counters[8];
  int error[4];
  double stats[4];
  uint32_t grand_total;
  uint32_t grand_total_noise;
  unsigned long long n_nsites;
  unsigned long long eq_stats[16];  // event kernel: tasks and lanes per event (ABEILLE_B200_EQ_STATS=1)
  unsigned long long avail;  // rows of a streamed input bank that have arrived (abl_transport)
  double moments[8];         // abl_bank_moments_device
};

thread_local std::string g_create_error;

}  // namespace

struct abl_context {
  int device = 0;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;
  DevProblem P{};
  std::vector<void*> allocs;  // immutable tables
  std::vector<uint64_t> tally_g;
  double k_site_max = 1.;       // max over (material, group) of nu Sigma_f / Sigma_a: sites per unit weight, bound
  std::vector<double> host_Et;  // [M*G] total cross sections (abl_set_sampling_xs rebuilds the quotient table from them)
  // scratch
  Site* sites = nullptr;
  uint64_t site_cap = 0;
  uint32_t *nfis = nullptr, *offsets = nullptr, *tile_sums = nullptr;
  uint32_t *tr_flights = nullptr, *tr_real = nullptr, *tr_virtual = nullptr;
  uint64_t *tr_hash = nullptr, *tr_rng = nullptr;
  uint64_t hist_cap = 0, trace_cap = 0, trace_n = 0;
  DevSmall* small_dev = nullptr;
  DevSmall* small_host = nullptr;  // pinned
  double* secondaries = nullptr;
  uint64_t sec_threads = 0;
  // noise mode: scratch noise particles, per-history counts / offsets, daughter ids of the scratch sites
  Site* nsites = nullptr;
  uint64_t nsite_cap = 0, did_cap = 0, ndid_cap = 0, nnoise_cap = 0;
  uint32_t *nnoise = nullptr, *noffsets = nullptr, *ntile_sums = nullptr, *site_did = nullptr, *nsite_did = nullptr;
  uint64_t pending_rows = 0;  // abl_transport_begin .. abl_transport_finish: rows of the fission bank waiting in stage_out
  uint32_t* site_inv = nullptr;  // bank row -> scratch site (place_sites_kernel)
  uint64_t inv_cap = 0;
  // exact cancelators: {parent's previous position, sampling xs} of the scratch sites, and of the rows of the last fission bank
  double* site_parent = nullptr;
  uint64_t site_parent_cap = 0;
  double* parent_info = nullptr;  // [ABL_PARENT_FIELDS][parent_cap] in bank order: x, y, z, Esmp, ux, uy, uz, E before the last scatter, E, was_virtual
  uint64_t parent_cap = 0, parent_n = 0;
  // BasicExactMGCancelator::bins.  Kept across calls and clear()ed like the reference's: the bucket array survives, and with it
  // the order the map is walked in (which orders the uniform particles).  Outer key k + Nz (j + Ny i) -- the value the reference's
  // KeyHash feeds std::hash<int> --, inner key the material index.
  struct ExactBin {
    double uniform_wgt = 0., uniform_wgt2 = 0., W = 0., W2 = 0.;
    std::vector<uint64_t> particles;
  };
  std::unordered_map<int, std::unordered_map<int, ExactBin>> exact_bins;
  std::vector<int32_t> exact_slot_of_key;  // dense key -> slot table of abl_cancel_exact_device (all -1 between calls)
  bool exact_full_ready = false;  // kind ABL_CANCEL_EXACT with its tables (chi rows, energy bins) on the device
  int nm_blocks_per_sm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  int implicit_blocks_per_sm[3] = {0, 0, 0};  // implicit-leakage delta tracking: per-lane kernel in modes 0 | 1 | 2
  // host-buffer entry point: the bank is copied in row chunks on its own stream while the history kernel runs
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_ids = nullptr, ev_zeroed = nullptr;
  unsigned long long* chunk_ends = nullptr;  // pinned
  bool streamed_input = false;
  CancelBins cancel{nullptr, nullptr, nullptr, nullptr, nullptr};
  // staging banks of the host-buffer API
  BankView stage_in{}, stage_out{};
  uint64_t stage_in_cap = 0, stage_out_cap = 0;
  double* probe_buf = nullptr;
  uint64_t probe_cap = 0;
  int blocks_per_sm[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // by tracker and kernel build (plain, track-length, traced)
  int hk_slots[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  int hk_fixed[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  int hk_tables[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // bytes of tables staged in shared memory (0: read from global memory)  // histories per CTA of the staged kernel (by tracker, trace)
  int geo_frames = 1, geo_pads = 2;  // nesting depth of the geometry: coordinate frames / stack pads a history can hold
  int smem_optin = 0, smem_total_sm = 0;  // shared memory a CTA may opt in to / an SM has
  // the small immutable tables live in one arena, so the staged history kernel can copy them into shared memory in one
  // sweep and run with its table pointers redirected there (stage_tables)
  unsigned char* arena = nullptr;
  size_t arena_used = 0, arena_cap = 0;
  unsigned long long smem_generic_base = 0;  // generic address of byte 0 of a kernel's dynamic shared memory
  uint64_t* rng_scratch = nullptr;  // pcg32 states seeded on the device when the caller passes id_c = NULL
  uint64_t rng_cap = 0;
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_kernel_ms = 0.f;
  int last_grid = 0, last_block = 128;
  cudaStream_t last_stream = nullptr;
  bool has_last_stream = false;
  std::string error;
};

namespace {

#define ABL_CUDA(h, call)                                                                         \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      (h)->error = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
      return ABL_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

// Calls on a handle share scratch buffers: when the caller switches streams, drain the previous one first.
cudaStream_t use_stream(abl_handle h, cudaStream_t s) {
  if (h->has_last_stream && h->last_stream != s) cudaStreamSynchronize(h->last_stream);
  h->last_stream = s;
  h->has_last_stream = true;
  return s;
}

int fail(abl_handle h, int code, const std::string& msg) {
  h->error = msg;
  return code;
}

constexpr size_t ARENA_CAP = 48 * 1024, ARENA_MAX_TABLE = 16 * 1024;

template <typename T>
int upload(abl_handle h, const T* src, size_t n, const T** dst) {
  *dst = nullptr;
  if (n == 0) return ABL_OK;
  if (!src) return fail(h, ABL_ERR_INVALID, "null table pointer");
  void* d = nullptr;
  const size_t bytes = n * sizeof(T), at = (h->arena_used + 15) & ~size_t(15);
  if (h->arena && bytes <= ARENA_MAX_TABLE && at + bytes <= h->arena_cap) {
    d = h->arena + at;
    h->arena_used = at + bytes;
  } else {
    ABL_CUDA(h, cudaMalloc(&d, bytes));
    h->allocs.push_back(d);
  }
  ABL_CUDA(h, cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
  *dst = static_cast<const T*>(d);
  return ABL_OK;
}

// libstdc++ discrete_distribution partial sums (bits/random.tcc:2655-2713)
std::vector<double> discrete_table(const std::vector<double>& w) {
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

int validate(abl_handle h, const abl_problem* p) {
  if (p->ngroups < 1 || !p->energy_bounds) return fail(h, ABL_ERR_INVALID, "ngroups / energy_bounds");
  if (p->tracking < ABL_TRACK_SURFACE || p->tracking > ABL_TRACK_IMPLICIT_LEAKAGE) return fail(h, ABL_ERR_INVALID, "tracking");
  if (p->mode != ABL_MODE_K_EIGENVALUE && p->mode != ABL_MODE_NOISE && p->mode != ABL_MODE_FIXED_SOURCE && p->mode != ABL_MODE_BRANCHLESS)
    return fail(h, ABL_ERR_UNSUPPORTED, "simulation modes on the device: k-eigenvalue, noise");
  if (p->mode == ABL_MODE_NOISE) {
    if (p->n_noise_sources < 1 || !p->noise_sources) return fail(h, ABL_ERR_INVALID, "noise mode without a noise source");
    if (!(p->w_noise > 0.)) return fail(h, ABL_ERR_INVALID, "noise mode needs a positive noise-angular-frequency");
  }
  if (p->cancelator.present && p->cancelator.kind == ABL_CANCEL_EXACT) {
    if (p->tracking != ABL_TRACK_DELTA && p->tracking != ABL_TRACK_CARTER)
      return fail(h, ABL_ERR_UNSUPPORTED, "exact cancelators need delta or carter tracking (src/cancelator.cpp:59-63)");
    if (p->mode == ABL_MODE_NOISE) return fail(h, ABL_ERR_UNSUPPORTED, "exact cancelators in noise mode (the noise kernels do not carry the parents' data)");
  }
  if (p->cancelator.present && p->cancelator.kind == ABL_CANCEL_BASIC_EXACT) {
    if (p->tracking != ABL_TRACK_DELTA && p->tracking != ABL_TRACK_CARTER)
      return fail(h, ABL_ERR_UNSUPPORTED, "basic-exact cancelators need delta or carter tracking (src/cancelator.cpp:43-47)");
    if (p->cancelator.beta < ABL_BETA_ZERO || p->cancelator.beta > ABL_BETA_AVERAGE_G) return fail(h, ABL_ERR_INVALID, "cancelator beta");
    if (p->mode == ABL_MODE_NOISE)
      return fail(h, ABL_ERR_UNSUPPORTED, "basic-exact cancelators in noise mode (the noise kernels do not carry the parents' data)");
    if (p->cancelator.beta >= ABL_BETA_AVERAGE_F && (p->cancelator.n_samples < 1 || p->cancelator.n_samples > ABL_EXACT_MAX_SAMPLES))
      return fail(h, ABL_ERR_UNSUPPORTED, "basic-exact cancelator: n-samples between 1 and 64");
  }
  if (p->ntallies > ABL_MAX_TALLIES) return fail(h, ABL_ERR_UNSUPPORTED, "more than ABL_MAX_TALLIES mesh tallies");
  if (p->root_universe < 0 || p->root_universe >= p->nuniverses) return fail(h, ABL_ERR_INVALID, "root universe");
  for (int c = 0; c < p->ncells; c++) {
    const abl_cell& cl = p->cells[c];
    if (cl.rpn_offset < 0 || cl.rpn_len < 0 || cl.rpn_offset + cl.rpn_len > p->nrpn) return fail(h, ABL_ERR_INVALID, "cell rpn slice");
    if (cl.rpn_len > 64) return fail(h, ABL_ERR_UNSUPPORTED, "cell region longer than 64 tokens");
    for (int k = 0; k < cl.rpn_len; k++) {
      const int32_t t = p->rpn[cl.rpn_offset + k];
      if (t >= ABL_OP_UNION) continue;
      const int s = t < 0 ? -t : t;
      if (s < 1 || s > p->nsurfaces) return fail(h, ABL_ERR_INVALID, "cell references an unknown surface");
    }
    if (cl.fill_universe >= p->nuniverses || (cl.fill_universe < 0 && (cl.material < 0 || cl.material >= p->nmaterials)))
      return fail(h, ABL_ERR_INVALID, "cell fill");
  }
  for (int u = 0; u < p->nuniverses; u++) {
    const abl_universe& U = p->universes[u];
    if (U.type == ABL_UNI_CELLS) {
      if (U.cell_offset < 0 || U.cell_offset + U.ncells > p->n_universe_cells) return fail(h, ABL_ERR_INVALID, "universe cell slice");
      for (int k = 0; k < U.ncells; k++)
        if (p->universe_cells[U.cell_offset + k] < 0 || p->universe_cells[U.cell_offset + k] >= p->ncells)
          return fail(h, ABL_ERR_INVALID, "universe cell index");
    } else if (U.type == ABL_UNI_RECT || U.type == ABL_UNI_HEX) {
      if (U.type == ABL_UNI_HEX && (U.N[0] != U.N[1] || U.N[0] != 2 * ((U.pad_ & 0xffff) - 1) + 1)) return fail(h, ABL_ERR_INVALID, "hexagonal lattice shape");
      const long long nt = (long long)U.N[0] * U.N[1] * U.N[2];
      if (nt <= 0 || U.tile_offset < 0 || U.tile_offset + nt > p->n_lattice_tiles) return fail(h, ABL_ERR_INVALID, "lattice tile slice");
      for (long long k = 0; k < nt; k++)
        if (p->lattice_tiles[U.tile_offset + k] >= p->nuniverses) return fail(h, ABL_ERR_INVALID, "lattice tile universe");
      if (U.outer >= p->nuniverses) return fail(h, ABL_ERR_INVALID, "lattice outer universe");
    } else {
      return fail(h, ABL_ERR_UNSUPPORTED, "universe type");
    }
  }
  if (p->tracking != ABL_TRACK_SURFACE && !p->sampling_xs) return fail(h, ABL_ERR_INVALID, "sampling_xs missing");
  // Source::generate_particle's energy rejection (src/source.cpp:48-58): a mono-energetic source outside (min, max) never passes
  for (int s = 0; s < p->nsources; s++)
    if (p->sources[s].energy_kind == ABL_EN_MONO && (p->sources[s].energy <= p->min_energy || p->energy_bounds[p->ngroups] <= p->sources[s].energy))
      return fail(h, ABL_ERR_INVALID, "source energy outside (min_energy, max_energy): Exceded 200 samplings of energy.");
  for (int s = 0; s < p->nsources; s++) {
    const abl_source& f = p->sources[s];
    if (f.direction_kind < ABL_DIR_ISOTROPIC || f.direction_kind > ABL_DIR_CONE) return fail(h, ABL_ERR_INVALID, "source direction kind");
    if (f.energy_kind < ABL_EN_MONO || f.energy_kind > ABL_EN_WATT) return fail(h, ABL_ERR_INVALID, "source energy kind");
    if (f.energy_kind != ABL_EN_MONO && !(f.en_a > 0.)) return fail(h, ABL_ERR_INVALID, "source energy parameter a must be > 0");
    if (f.energy_kind == ABL_EN_WATT && !(f.en_b > 0.)) return fail(h, ABL_ERR_INVALID, "source energy parameter b must be > 0");
  }
  return ABL_OK;
}

// Cell regions that are a pure intersection of axis-aligned half-spaces (a box) or a single z-cylinder sense
// get a compiled descriptor (tables.h: CellFast); everything else stays CF_GENERIC.
std::vector<CellFast> compile_cells(const abl_problem* p) {
  std::vector<CellFast> out((size_t)p->ncells);
  const double inf = std::numeric_limits<double>::infinity();
  for (int c = 0; c < p->ncells; c++) {
    CellFast& f = out[(size_t)c];
    std::memset(&f, 0, sizeof f);
    f.kind = CF_GENERIC;
    const abl_cell& cl = p->cells[c];
    if (!cl.simple || cl.rpn_len < 1) continue;
    bool all_planes = true;
    for (int k = 0; k < cl.rpn_len; k++) {
      const int32_t t = p->rpn[cl.rpn_offset + k];
      if (t >= ABL_OP_UNION) { all_planes = false; break; }
      const abl_surface& s = p->surfaces[(t < 0 ? -t : t) - 1];
      if (s.type != ABL_SURF_XPLANE && s.type != ABL_SURF_YPLANE && s.type != ABL_SURF_ZPLANE) all_planes = false;
    }
    if (all_planes) {
      f.kind = CF_BOX;
      for (int a = 0; a < 3; a++) { f.a[2 * a] = -inf; f.a[2 * a + 1] = inf; }
      for (int k = 0; k < cl.rpn_len; k++) {
        const int32_t t = p->rpn[cl.rpn_offset + k];
        const abl_surface& s = p->surfaces[(t < 0 ? -t : t) - 1];
        const int a = s.type - ABL_SURF_XPLANE;
        if (t > 0) { if (s.p[0] > f.a[2 * a]) f.a[2 * a] = s.p[0]; }          // r - p0 > 0
        else { if (s.p[0] < f.a[2 * a + 1]) f.a[2 * a + 1] = s.p[0]; }        // r - p0 < 0
      }
      continue;
    }
    if (cl.rpn_len == 1) {
      const int32_t t = p->rpn[cl.rpn_offset];
      const abl_surface& s = p->surfaces[(t < 0 ? -t : t) - 1];
      if (s.type == ABL_SURF_ZCYL) {
        f.kind = CF_ZCYL;
        f.sense = t < 0 ? -1 : 1;
        f.a[0] = s.p[0];
        f.a[1] = s.p[1];
        f.a[2] = s.p[2] * s.p[2];  // the same IEEE product the generic evaluator forms (zcylinder.cpp:33-47)
      }
    }
  }
  return out;
}

// Nesting depth of the geometry: the most pads (universe / lattice / cell entries of the Tracker's stack,
// tracker.hpp:41-50) and coordinate frames (the global one plus one per lattice tile entered) a history can hold.
// They size the per-history columns of the staged kernel's shared memory (history.cuh).
void geometry_depth(const abl_problem* p, int& frames, int& pads) {
  std::vector<int> pd((size_t)p->nuniverses, -1), fd((size_t)p->nuniverses, -1);
  std::function<void(int, int)> visit = [&](int u, int guard) {
    if (u < 0 || u >= p->nuniverses || pd[(size_t)u] >= 0 || guard > 64) return;
    const abl_universe& U = p->universes[u];
    int bp = 1, bf = 0;
    pd[(size_t)u] = 0;  // (marks the visit; recursion is rejected by validate())
    if (U.type == ABL_UNI_CELLS) {
      bp = 2;
      for (int k = 0; k < U.ncells; k++) {
        const int fill = p->cells[p->universe_cells[U.cell_offset + k]].fill_universe;
        if (fill < 0) continue;
        visit(fill, guard + 1);
        bp = std::max(bp, 2 + pd[(size_t)fill]);
        bf = std::max(bf, fd[(size_t)fill]);
      }
    } else {
      const int nt = U.N[0] * U.N[1] * U.N[2];
      for (int k = 0; k < nt; k++) {
        const int t = p->lattice_tiles[U.tile_offset + k];
        if (t < 0) continue;
        visit(t, guard + 1);
        bp = std::max(bp, 1 + pd[(size_t)t]);
        bf = std::max(bf, 1 + fd[(size_t)t]);
      }
      if (U.outer >= 0) {
        visit(U.outer, guard + 1);
        bp = std::max(bp, 1 + pd[(size_t)U.outer]);
        bf = std::max(bf, fd[(size_t)U.outer]);
      }
    }
    pd[(size_t)u] = bp;
    fd[(size_t)u] = bf;
  };
  visit(p->root_universe, 0);
  pads = std::min(std::max(pd[(size_t)p->root_universe], 2), ABL_MAX_PADS);
  frames = std::min(1 + std::max(fd[(size_t)p->root_universe], 0), ABL_MAX_FRAMES);
}

// The vacuum / reflective surfaces as axis planes in the global frame (DevProblem::bc_*), or n = 0 when the geometry does not
// allow it: every surface with a boundary condition must be an x / y / z plane, and no universe that holds a cell touching one
// may be reachable through a lattice tile (inside a tile positions are local: r - tile centre; cells reached through cell
// fills and a lattice's outer universe keep the global frame, rect_lattice.cpp:132-207).
int boundary_planes(const abl_problem* p, int32_t* axis, double* p0) {
  int n = 0;
  for (int i = 0; i < p->nsurfaces; i++) {
    const abl_surface& s = p->surfaces[i];
    if (s.bc == ABL_BC_NORMAL) continue;
    if (s.type != ABL_SURF_XPLANE && s.type != ABL_SURF_YPLANE && s.type != ABL_SURF_ZPLANE) return 0;
    if (n == ABL_MAX_BC_PLANES) return 0;
    axis[n] = s.type == ABL_SURF_XPLANE ? 0 : (s.type == ABL_SURF_YPLANE ? 1 : 2);
    p0[n] = s.p[0];
    n++;
  }
  if (n == 0) return 0;
  // universes reachable after at least one frame shift
  std::vector<char> shifted((size_t)p->nuniverses, 0), seen((size_t)p->nuniverses, 0);
  std::function<void(int, bool, int)> visit = [&](int u, bool sh, int guard) {
    if (u < 0 || u >= p->nuniverses || guard > 64) return;
    if (sh ? shifted[(size_t)u] : seen[(size_t)u]) return;
    (sh ? shifted : seen)[(size_t)u] = 1;
    const abl_universe& U = p->universes[u];
    if (U.type == ABL_UNI_CELLS) {
      for (int k = 0; k < U.ncells; k++) visit(p->cells[p->universe_cells[U.cell_offset + k]].fill_universe, sh, guard + 1);
    } else {
      const int nt = U.N[0] * U.N[1] * U.N[2];
      for (int k = 0; k < nt; k++) visit(p->lattice_tiles[U.tile_offset + k], true, guard + 1);
      visit(U.outer, sh, guard + 1);
    }
  };
  visit(p->root_universe, false, 0);
  for (int u = 0; u < p->nuniverses; u++) {
    if (!shifted[(size_t)u]) continue;
    const abl_universe& U = p->universes[u];
    if (U.type != ABL_UNI_CELLS) continue;
    for (int k = 0; k < U.ncells; k++)
      if (p->cells[p->universe_cells[U.cell_offset + k]].vac_or_refl) return 0;
  }
  return n;
}

DevMesh3 make_mesh3(const abl_mesh3& m, const double* tally_eb_dev) {
  DevMesh3 d{};
  d.present = m.present;
  d.Nx = m.N[0]; d.Ny = m.N[1]; d.Nz = m.N[2];
  d.lowx = m.low[0]; d.lowy = m.low[1]; d.lowz = m.low[2];
  d.hix = m.hi[0]; d.hiy = m.hi[1]; d.hiz = m.hi[2];
  if (m.present) {
    d.dx = (m.hi[0] - m.low[0]) / static_cast<double>(m.N[0]);
    d.dy = (m.hi[1] - m.low[1]) / static_cast<double>(m.N[1]);
    d.dz = (m.hi[2] - m.low[2]) / static_cast<double>(m.N[2]);
  }
  if (m.n_energy_edges >= 2) {
    d.Ne = m.n_energy_edges - 1;
    d.eedges = tally_eb_dev + m.eedges_offset;
  } else {
    d.Ne = 1;
    d.eedges = nullptr;
  }
  return d;
}

int grid_for(abl_handle h, uint64_t n, int threads) {
  uint64_t blocks = (n + threads - 1) / threads;
  const uint64_t cap = (uint64_t)h->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

BankView view_of(const abl_bank* b) {
  BankView v;
  v.n = b->n;
  v.x = b->x; v.y = b->y; v.z = b->z; v.ux = b->ux; v.uy = b->uy; v.uz = b->uz;
  v.E = b->E; v.wgt = b->wgt; v.wgt2 = b->wgt2;
  v.id_a = b->id_a; v.id_b = b->id_b; v.id_c = b->id_c;
  return v;
}

int ensure_history_scratch(abl_handle h, uint64_t n, bool trace) {
  if (n > h->hist_cap) {
    const uint64_t cap = n + n / 4 + 1024;
    if (h->nfis) cudaFree(h->nfis);
    if (h->offsets) cudaFree(h->offsets);
    if (h->tile_sums) cudaFree(h->tile_sums);
    h->nfis = h->offsets = h->tile_sums = nullptr;
    h->hist_cap = 0;
    ABL_CUDA(h, cudaMalloc(&h->nfis, cap * sizeof(uint32_t)));
    ABL_CUDA(h, cudaMalloc(&h->offsets, cap * sizeof(uint32_t)));
    ABL_CUDA(h, cudaMalloc(&h->tile_sums, (cap / ABL_SCAN_TILE + 2) * sizeof(uint32_t)));
    h->hist_cap = cap;
  }
  if (trace && n > h->trace_cap) {
    const uint64_t cap = n + 1024;
    for (void* p : {(void*)h->tr_flights, (void*)h->tr_real, (void*)h->tr_virtual, (void*)h->tr_hash, (void*)h->tr_rng})
      if (p) cudaFree(p);
    h->tr_flights = h->tr_real = h->tr_virtual = nullptr;  // (so that a failed allocation below leaves no dangling pointer)
    h->tr_hash = h->tr_rng = nullptr;
    h->trace_cap = 0;
    ABL_CUDA(h, cudaMalloc(&h->tr_flights, cap * 4));
    ABL_CUDA(h, cudaMalloc(&h->tr_real, cap * 4));
    ABL_CUDA(h, cudaMalloc(&h->tr_virtual, cap * 4));
    ABL_CUDA(h, cudaMalloc(&h->tr_hash, cap * 8));
    ABL_CUDA(h, cudaMalloc(&h->tr_rng, cap * 8));
    h->trace_cap = cap;
  }
  return ABL_OK;
}

int ensure_sites(abl_handle h, uint64_t cap) {
  if (cap > h->site_cap) {
    if (h->sites) cudaFree(h->sites);
    h->sites = nullptr;
    h->site_cap = 0;
    ABL_CUDA(h, cudaMalloc(&h->sites, cap * sizeof(Site)));
    h->site_cap = cap;
  }
  if (h->P.exact_cancel && cap > h->site_parent_cap) {
    if (h->site_parent) cudaFree(h->site_parent);
    h->site_parent = nullptr;
    h->site_parent_cap = 0;
    ABL_CUDA(h, cudaMalloc(&h->site_parent, cap * ABL_PARENT_FIELDS * sizeof(double)));
    h->site_parent_cap = cap;
  }
  return ABL_OK;
}

// rows of the exact cancelators' side table (x, y, z, Esmp columns of `cap` entries each); contents of the first `keep` rows survive
int ensure_parent_info(abl_handle h, uint64_t cap, uint64_t keep) {
  if (cap <= h->parent_cap) return ABL_OK;
  double* n = nullptr;
  ABL_CUDA(h, cudaMalloc(&n, cap * ABL_PARENT_FIELDS * sizeof(double)));
  ABL_CUDA(h, cudaMemset(n, 0, cap * ABL_PARENT_FIELDS * sizeof(double)));
  if (h->parent_info && keep)
    for (int q = 0; q < ABL_PARENT_FIELDS; q++)
      ABL_CUDA(h, cudaMemcpy(n + q * cap, h->parent_info + q * h->parent_cap, keep * sizeof(double), cudaMemcpyDeviceToDevice));
  if (h->parent_info) cudaFree(h->parent_info);
  h->parent_info = n;
  h->parent_cap = cap;
  return ABL_OK;
}

void free_bank(BankView& b) {
  for (void* p : {(void*)b.x, (void*)b.y, (void*)b.z, (void*)b.ux, (void*)b.uy, (void*)b.uz, (void*)b.E, (void*)b.wgt,
                  (void*)b.wgt2, (void*)b.id_a, (void*)b.id_b, (void*)b.id_c})
    if (p) cudaFree(p);
  b = BankView{};
}

int alloc_bank(abl_handle h, BankView& b, uint64_t cap) {
  free_bank(b);
  double** dp[9] = {&b.x, &b.y, &b.z, &b.ux, &b.uy, &b.uz, &b.E, &b.wgt, &b.wgt2};
  for (auto p : dp) ABL_CUDA(h, cudaMalloc(p, cap * sizeof(double)));
  uint64_t** up[3] = {&b.id_a, &b.id_b, &b.id_c};
  for (auto p : up) ABL_CUDA(h, cudaMalloc(p, cap * sizeof(uint64_t)));
  b.n = cap;
  return ABL_OK;
}

// The staged history kernel copies the table arena into its shared memory (at HK_COLS_OFFSET) and runs on a DevProblem
// whose table pointers point at that copy: generic addresses of the shared window, the same for every CTA.  Pointers
// outside the arena (tables too large for it, tally arrays) are left alone.
DevProblem stage_tables(abl_handle h, unsigned smem_offset) {
  DevProblem Q = h->P;
  const unsigned char *lo = h->arena, *hi = h->arena + h->arena_used;
  const long long delta = (long long)(h->smem_generic_base + smem_offset) - (long long)reinterpret_cast<uintptr_t>(lo);
  auto mv = [&](auto& ptr) {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(ptr);
    if (b >= lo && b < hi) ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(reinterpret_cast<uintptr_t>(b) + delta);
  };
  mv(Q.ebounds); mv(Q.jump); mv(Q.surfaces); mv(Q.cells); mv(Q.rpn); mv(Q.universes); mv(Q.ucells); mv(Q.tiles); mv(Q.cellfast);
  mv(Q.Et); mv(Q.Ea); mv(Q.Ef); mv(Q.Es); mv(Q.nu); mv(Q.nud); mv(Q.speed); mv(Q.chi_cp); mv(Q.ps_cp); mv(Q.angle);
  mv(Q.amu); mv(Q.apdf); mv(Q.acdf); mv(Q.dg_off); mv(Q.dg_cp); mv(Q.dg_lambda); mv(Q.fissile); mv(Q.smp);
  mv(Q.real_frac); mv(Q.surv_frac); mv(Q.inv_score); mv(Q.tally_gbin); mv(Q.gmid); mv(Q.tally_dev);
  for (int t = 0; t < Q.ntallies; t++) mv(Q.tally[t].ebounds);
  return Q;
}

template <int TRK, bool TRACE>
int launch_transport(abl_handle h, const RunArgs& A, uint64_t n, cudaStream_t s) {
  // the staged lock-step loop with the histories in shared-memory columns and service warps for the rare events
  // (history.cuh); one translation unit per tracker, each with its own launch shape
  const bool tle = A.converged && h->P.n_tl_tallies;
  // delta and carter tracking have two kernels: the staged lock-step kernel (history.cuh) and the event-queue kernel
  // (events.cuh, ABEILLE_B200_EVENTS=1).  Both are bit-exact; on the bench workload the staged kernel measures 117 ms per
  // 1e7 histories and the event kernel 126 ms, so the staged one is the default.  Surface tracking
  // has the staged kernel only.
  static const bool want_events = getenv("ABEILLE_B200_EVENTS") != nullptr;
  const bool events = TRK != ABL_TRACK_SURFACE && want_events;
  const int nf = h->geo_frames < HK_MIN_FRAMES ? HK_MIN_FRAMES : h->geo_frames, np = h->geo_pads;
  // the build with the column shape compiled in when the geometry has that nesting depth (and the shared memory holds a
  // history per thread: checked below, remembered in hk_fixed)
  int& slots = h->hk_slots[TRK][TRACE ? 2 : (tle ? 1 : 0)];
  int& fixed_state = h->hk_fixed[TRK][TRACE ? 2 : (tle ? 1 : 0)];  // 0 untried, 1 in use, -1 does not fit
  static const bool no_fixed = getenv("ABEILLE_B200_NO_FIXED_SHAPE") != nullptr;
  const bool fixed = !TRACE && !events && !no_fixed && nf == HK_FIXED_NF && np == HK_FIXED_NP && fixed_state >= 0 && !h->P.has_hex;
  auto pick = [&](bool fx) {
    return TRK == ABL_TRACK_SURFACE ? history_kernel_surface(TRACE, tle, fx)
           : events ? (TRK == ABL_TRACK_DELTA ? event_kernel_delta(TRACE, tle) : event_kernel_carter(TRACE, tle))
                    : (TRK == ABL_TRACK_DELTA ? history_kernel_delta(TRACE, tle, fx) : history_kernel_carter(TRACE, tle, fx));
  };
  HistoryKernel hk = pick(fixed);
  TransportKernel kern = hk.fn;
  const int threads = hk.threads;
  int& bps = h->blocks_per_sm[TRK][TRACE ? 2 : (tle ? 1 : 0)];
  const unsigned slot_bytes = hk_slot_bytes(nf, np, TRACE);
  if (bps == 0) {
    // as many histories per CTA as the shared memory holds (a multiple of 32, at most one per history thread); two CTAs
    // per SM when the kernel was built for that
    cudaFuncAttributes fa;
    ABL_CUDA(h, cudaFuncGetAttributes(&fa, kern));
    int want_blocks = 1;
    if (!hk.events) {
      int regs_limit = 65536 / (fa.numRegs * threads > 0 ? fa.numRegs * threads : 1);
      if (regs_limit >= 2 && 2 * threads <= 2048) want_blocks = 2;
    }
    const int room = (want_blocks == 2 ? (h->smem_total_sm / 2 - 1024) : h->smem_optin) - (int)hk.fixed_bytes - (int)fa.sharedSizeBytes;
    // the tables are staged in shared memory when that costs at most a quarter of the histories the CTA could hold
    const int tables = (int)((h->arena_used + 15) & ~size_t(15));
    int sl_plain = room / (int)slot_bytes, sl = (room - tables) / (int)slot_bytes;
    sl_plain -= sl_plain % 32;
    sl -= sl % 32;
    if (sl_plain > hk.hist_threads) sl_plain = hk.hist_threads;
    if (sl > hk.hist_threads) sl = hk.hist_threads;
    bool stage = h->smem_generic_base != 0 && fa.sharedSizeBytes == 0 && sl >= 32 && 4 * sl >= 3 * sl_plain;
    if (getenv("ABEILLE_B200_NO_SMEM_TABLES")) stage = false;
    if (!stage) sl = sl_plain;
    if (hk.events) {  // (tuning aid: fewer slots per CTA than the shared memory holds)
      const char* e = getenv("ABEILLE_B200_EQ_SLOTS");
      const int v = e ? atoi(e) : 0;
      if (v >= 32 && v < sl) sl = v - v % 32;
    }
    h->hk_tables[TRK][TRACE ? 2 : (tle ? 1 : 0)] = stage ? tables : 0;
    if (sl < 32) {
      h->error = "geometry nesting too deep for the staged history kernel's shared memory";
      return ABL_ERR_GEOMETRY;
    }
    if (fixed && sl != hk.fixed_slots) {  // the fixed-shape build needs exactly its slot count: use the general one
      fixed_state = -1;
      return launch_transport<TRK, TRACE>(h, A, n, s);
    }
    if (fixed) fixed_state = 1;
    slots = sl;
    const size_t smem = hk.fixed_bytes + (size_t)(stage ? tables : 0) + (size_t)slot_bytes * (size_t)sl;
    ABL_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    ABL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem));
    if (nb < 1) nb = 1;
    bps = nb;
  }
  const int tables = h->hk_tables[TRK][TRACE ? 2 : (tle ? 1 : 0)];
  const size_t smem = hk.fixed_bytes + (size_t)tables + (size_t)slot_bytes * (size_t)slots;
  const int worker_threads = slots;  // threads of a block that own histories
  uint64_t blocks = (uint64_t)h->sm_count * bps;
  const uint64_t need = (n + worker_threads - 1) / worker_threads;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  RunArgs B = A;
  B.hk_slots = slots;
  B.hk_nf = nf;
  B.hk_np = np;
  B.hk_tables = tables;
  B.arena = h->arena;
  B.smem_generic_base = h->smem_generic_base;
  static const bool eq_stats = getenv("ABEILLE_B200_EQ_STATS") != nullptr;
  B.eq_stats = hk.events && eq_stats ? h->small_dev->eq_stats : nullptr;
  {
    static const double timeout_s = [] {
      const char* e = getenv("ABEILLE_B200_KERNEL_TIMEOUT_S");
      const double v = e ? atof(e) : 0.;
      return v > 0. ? v : 120.;
    }();
    B.timeout_ns = (unsigned long long)(timeout_s * 1e9);
  }
  if (TRK == ABL_TRACK_CARTER) {
    // one LIFO per history in flight: per thread in the staged kernel, per slot in the event kernel
    const int owners = hk.events ? slots : threads;
    const uint64_t nthreads = blocks * owners;
    if (nthreads > h->sec_threads) {
      if (h->secondaries) cudaFree(h->secondaries);
      h->secondaries = nullptr;
      h->sec_threads = 0;
      const uint64_t cap = (uint64_t)h->sm_count * bps * owners;
      ABL_CUDA(h, cudaMalloc(&h->secondaries, cap * ABL_SEC_CAP * ABL_SEC_FIELDS * sizeof(double)));
      h->sec_threads = cap;
    }
    B.secondaries = h->secondaries;
  }
  if (B.bank.id_c == nullptr) {  // seed(seed); advance(stride * history id) for the whole bank
    if (n > h->rng_cap) {
      if (h->rng_scratch) cudaFree(h->rng_scratch);
      h->rng_scratch = nullptr;
      h->rng_cap = 0;
      ABL_CUDA(h, cudaMalloc(&h->rng_scratch, (n + n / 4 + 1024) * sizeof(uint64_t)));
      h->rng_cap = n + n / 4 + 1024;
    }
    seed_streams_kernel<<<grid_for(h, n, 256), 256, 0, s>>>(h->P, B.bank.id_a, n, h->rng_scratch);
    h->launches++;
    B.bank.id_c = h->rng_scratch;
  }
  cudaEventRecord(h->ev0, s);
  kern<<<(unsigned)blocks, threads, smem, s>>>(tables ? stage_tables(h, hk.fixed_bytes) : h->P, B);
  h->last_block = threads;
  cudaEventRecord(h->ev1, s);
  h->last_grid = (int)blocks;
  h->launches++;
  ABL_CUDA(h, cudaGetLastError());
  return ABL_OK;
}

// noise-mode runs (simulation: noise) use the per-lane history loop for every tracker: MODE 1 = power-iteration
// generation (may sample the noise source), MODE 2 = noise particles (noise.cuh)
template <int TRK, int MODE>
int launch_transport_nm(abl_handle h, const RunArgs& A, uint64_t n, cudaStream_t s) {
  const bool implicit = TRK == ABL_TRACK_IMPLICIT_LEAKAGE;
  TransportKernel kern = implicit ? implicit_kernel(MODE) : lane_kernel(TRK, MODE);
  const int threads = TK_THREADS;
  int& bps = implicit ? h->implicit_blocks_per_sm[MODE] : h->nm_blocks_per_sm[implicit ? 0 : TRK][MODE];
  if (bps == 0) {
    int nb = 0;
    ABL_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, 0));
    bps = nb < 1 ? 1 : nb;
  }
  uint64_t blocks = (uint64_t)h->sm_count * bps;
  const uint64_t need = (n + threads - 1) / threads;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  RunArgs B = A;
  if (MODE == 2 || TRK == ABL_TRACK_CARTER || h->P.mode == ABL_MODE_FIXED_SOURCE || h->P.mode == ABL_MODE_BRANCHLESS) {  // secondaries: noise copies, carter / branchless splitting, noise fission without inner generations, fixed-source fission
    const uint64_t cap = (uint64_t)h->sm_count * bps * threads;
    if (cap > h->sec_threads) {
      if (h->secondaries) cudaFree(h->secondaries);
      h->secondaries = nullptr;
      h->sec_threads = 0;
      ABL_CUDA(h, cudaMalloc(&h->secondaries, cap * ABL_SEC_CAP * ABL_SEC_FIELDS * sizeof(double)));
      h->sec_threads = cap;
    }
    B.secondaries = h->secondaries;
  }
  cudaEventRecord(h->ev0, s);
  kern<<<(unsigned)blocks, threads, 0, s>>>(h->P, B);
  cudaEventRecord(h->ev1, s);
  h->last_block = threads;
  h->last_grid = (int)blocks;
  h->launches++;
  ABL_CUDA(h, cudaGetLastError());
  return ABL_OK;
}
template <int MODE>
int launch_transport_nm(abl_handle h, const RunArgs& A, uint64_t n, cudaStream_t s) {
  switch (h->P.tracking) {
    case ABL_TRACK_SURFACE: return launch_transport_nm<ABL_TRACK_SURFACE, MODE>(h, A, n, s);
    case ABL_TRACK_DELTA: return launch_transport_nm<ABL_TRACK_DELTA, MODE>(h, A, n, s);
    case ABL_TRACK_IMPLICIT_LEAKAGE: return launch_transport_nm<ABL_TRACK_IMPLICIT_LEAKAGE, MODE>(h, A, n, s);
    default: return launch_transport_nm<ABL_TRACK_CARTER, MODE>(h, A, n, s);
  }
}

template <int TRK>
int launch_transport(abl_handle h, const RunArgs& A, uint64_t n, cudaStream_t s, bool trace) {
  return trace ? launch_transport<TRK, true>(h, A, n, s) : launch_transport<TRK, false>(h, A, n, s);
}

int status_from_device_error(abl_handle h, const DevSmall& sm) {
  if (sm.error[0] == 0) return ABL_OK;
  const uint64_t hid = (uint64_t)(uint32_t)sm.error[1] | ((uint64_t)(uint32_t)sm.error[2] << 32);
  char buf[160];
  const char* what = "device error";
  switch (sm.error[0]) {
    case ABL_ERR_LOST: what = "particle became lost after a reflection / crossing / resurrection"; break;
    case ABL_ERR_MAJORANT: what = "total cross section exceeded the majorant"; break;
    case ABL_ERR_GEOMETRY: what = "geometry nesting deeper than ABL_MAX_PADS / ABL_MAX_FRAMES"; break;
    case ABL_ERR_BANK_OVERFLOW: what = "secondary stack overflow (ABL_SEC_CAP)"; break;
    case ABL_ERR_TIMEOUT: what = "history kernel ran past its deadline and was wound down (ABEILLE_B200_KERNEL_TIMEOUT_S)"; break;
    case ABL_ERR_INVALID: what = "source sampling failed (point source outside the geometry, fissile-only rejection limit, or 200 samplings of energy exceeded)"; break;
  }
  snprintf(buf, sizeof buf, "%s (history %llu)", what, (unsigned long long)hid);
  h->error = buf;
  return sm.error[0];
}

int grow_u32(abl_handle h, uint32_t*& p, uint64_t& cap, uint64_t need) {
  if (need > cap) {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    ABL_CUDA(h, cudaMalloc(&p, need * sizeof(uint32_t)));
    cap = need;
  }
  return ABL_OK;
}

int transport_impl(abl_handle h, const BankView& in, const abl_gen_params* params, const BankView& out, uint64_t* n_fission,
                   double scores[6], uint64_t counters[8], cudaStream_t s, const BankView* noise_out = nullptr,
                   uint64_t* n_noise = nullptr) {
  const uint64_t N = in.n;
  if (N >= (1ull << 32)) return fail(h, ABL_ERR_UNSUPPORTED, "bank larger than 2^32-1 histories per device");
  const bool noise_mode = h->P.mode == ABL_MODE_NOISE;
  const bool sample_noise = params->sample_noise_source != 0;
  // a power-iteration generation of a noise run that does not sample the noise source is a plain k-eigenvalue generation:
  // it goes through the staged kernel; noise particles and sampling generations use the per-lane kernel (noise.cuh)
  // fixed-source problems (fission neutrons as secondaries of their history) run the per-lane kernel in its k-eigenvalue mode
  // and so do branchless collisions (their splitting needs the secondaries LIFO as well)
  // and problems with an exact cancelator (the per-lane kernel carries the parent's previous position and sampling xs)
  const bool fixed_source = h->P.mode == ABL_MODE_FIXED_SOURCE || h->P.mode == ABL_MODE_BRANCHLESS || (h->P.exact_cancel && !noise_mode);
  const bool lane_kernel_call = (noise_mode && (params->noise != 0 || sample_noise)) || fixed_source;
  if ((params->noise || sample_noise) && !noise_mode)
    return fail(h, ABL_ERR_INVALID, "noise transport / noise-source sampling needs a problem with simulation: noise");
  if (params->noise && sample_noise) return fail(h, ABL_ERR_INVALID, "the noise source is sampled in power-iteration generations only");
  if (sample_noise && (!noise_out || !n_noise)) return fail(h, ABL_ERR_INVALID, "noise-source sampling needs a noise bank");
  if (params->noise && !in.wgt2) return fail(h, ABL_ERR_INVALID, "noise transport needs the second weight (wgt2)");
  if (n_noise) *n_noise = 0;
  int rc = ensure_history_scratch(h, N, params->trace != 0);
  if (rc) return rc;
  rc = ensure_sites(h, out.n);
  if (rc) return rc;
  if (noise_mode) {
    if ((rc = grow_u32(h, h->site_did, h->did_cap, out.n)) != 0) return rc;
    if (sample_noise) {
      if (noise_out->n > h->nsite_cap) {
        if (h->nsites) cudaFree(h->nsites);
        h->nsites = nullptr;
        h->nsite_cap = 0;
        ABL_CUDA(h, cudaMalloc(&h->nsites, noise_out->n * sizeof(Site)));
        h->nsite_cap = noise_out->n;
      }
      if ((rc = grow_u32(h, h->nsite_did, h->ndid_cap, noise_out->n)) != 0) return rc;
      if (N > h->nnoise_cap) {
        const uint64_t cap = N + N / 4 + 1024;
        for (uint32_t** p : {&h->nnoise, &h->noffsets, &h->ntile_sums}) {
          if (*p) cudaFree(*p);
          *p = nullptr;
        }
        h->nnoise_cap = 0;
        ABL_CUDA(h, cudaMalloc(&h->nnoise, cap * sizeof(uint32_t)));
        ABL_CUDA(h, cudaMalloc(&h->noffsets, cap * sizeof(uint32_t)));
        ABL_CUDA(h, cudaMalloc(&h->ntile_sums, (cap / ABL_SCAN_TILE + 2) * sizeof(uint32_t)));
        h->nnoise_cap = cap;
      }
    }
  }
  // (a streamed input bank keeps its arrival counter, the last member, across this reset)
  ABL_CUDA(h, cudaMemsetAsync(h->small_dev, 0, h->streamed_input ? offsetof(DevSmall, avail) : sizeof(DevSmall), s));
  RunArgs A{};
  A.bank = in;
  A.ticket = &h->small_dev->ticket;
  A.sites = h->sites;
  A.n_sites = &h->small_dev->n_sites;
  A.site_capacity = out.n;
  A.nfis = h->nfis;
  if (params->trace) {
    A.tr_flights = h->tr_flights; A.tr_real = h->tr_real; A.tr_virtual = h->tr_virtual;
    A.tr_hash = h->tr_hash; A.tr_rng = h->tr_rng;
    h->trace_n = N;
  }
  A.scores = h->small_dev->scores;
  A.counters = h->small_dev->counters;
  A.error = h->small_dev->error;
  A.secondaries = nullptr;
  A.site_parent = h->P.exact_cancel ? h->site_parent : nullptr;
  A.k_col = params->k_col;
  A.keff = params->keff;
  A.converged = params->converged;
  A.avail = h->streamed_input ? &h->small_dev->avail : nullptr;
  A.sample_noise = sample_noise ? 1 : 0;
  A.site_did = h->site_did;
  if (sample_noise) {
    A.nsites = h->nsites;
    A.n_nsites = &h->small_dev->n_nsites;
    A.nsite_capacity = noise_out->n;
    A.nnoise = h->nnoise;
    A.nsite_did = h->nsite_did;
  }
  if (N > 0) {
    if (lane_kernel_call) {  // (the per-lane kernel seeds the streams itself when id_c is NULL)
      rc = fixed_source ? launch_transport_nm<0>(h, A, N, s) : (params->noise ? launch_transport_nm<2>(h, A, N, s) : launch_transport_nm<1>(h, A, N, s));
    } else if (h->P.tracking == ABL_TRACK_IMPLICIT_LEAKAGE) {  // k-eigenvalue generation through the per-lane kernel
      rc = launch_transport_nm<ABL_TRACK_IMPLICIT_LEAKAGE, 0>(h, A, N, s);
    } else {
      switch (h->P.tracking) {
        case ABL_TRACK_SURFACE: rc = launch_transport<ABL_TRACK_SURFACE>(h, A, N, s, params->trace != 0); break;
        case ABL_TRACK_DELTA: rc = launch_transport<ABL_TRACK_DELTA>(h, A, N, s, params->trace != 0); break;
        default: rc = launch_transport<ABL_TRACK_CARTER>(h, A, N, s, params->trace != 0); break;
      }
    }
    if (rc) return rc;
    // fission bank in the reference's order: offsets = exclusive scan of per-history counts
    const uint32_t ntiles = (uint32_t)((N + ABL_SCAN_TILE - 1) / ABL_SCAN_TILE);
    scan_tile_sums_kernel<<<ntiles, ABL_SCAN_THREADS, 0, s>>>(h->nfis, N, h->tile_sums);
    scan_top_kernel<<<1, ABL_SCAN_THREADS, 0, s>>>(h->tile_sums, ntiles, &h->small_dev->grand_total);
    scan_apply_kernel<<<ntiles, ABL_SCAN_THREADS, 0, s>>>(h->nfis, N, h->tile_sums, h->offsets);
    h->launches += 3;
    if (sample_noise) {  // the same for the noise bank
      scan_tile_sums_kernel<<<ntiles, ABL_SCAN_THREADS, 0, s>>>(h->nnoise, N, h->ntile_sums);
      scan_top_kernel<<<1, ABL_SCAN_THREADS, 0, s>>>(h->ntile_sums, ntiles, &h->small_dev->grand_total_noise);
      scan_apply_kernel<<<ntiles, ABL_SCAN_THREADS, 0, s>>>(h->nnoise, N, h->ntile_sums, h->noffsets);
      h->launches += 3;
    }
  }
  ABL_CUDA(h, cudaMemcpyAsync(h->small_host, h->small_dev, sizeof(DevSmall), cudaMemcpyDeviceToHost, s));
  ABL_CUDA(h, cudaStreamSynchronize(s));
  if (N > 0) cudaEventElapsedTime(&h->last_kernel_ms, h->ev0, h->ev1);
  const DevSmall& sm = *h->small_host;
  if (getenv("ABEILLE_B200_EQ_STATS") && sm.eq_stats[0] + sm.eq_stats[1]) {  // (development aid; order: events.cuh Q_*)
    static const char* names[] = {"step", "loc_tree", "loc_cell", "boundary", "refill", "fission"};
    fprintf(stderr, "event kernel:");
    for (int q = 0; q < 6; q++)
      fprintf(stderr, " %s %llu tasks x %.1f lanes;", names[q], sm.eq_stats[q], sm.eq_stats[q] ? (double)sm.eq_stats[6 + q] / sm.eq_stats[q] : 0.);
    fprintf(stderr, " idle polls %llu\n", sm.eq_stats[12]);
  }
  for (int i = 0; i < 6; i++) scores[i] = sm.scores[i];
  if (counters)
    for (int i = 0; i < 8; i++) counters[i] = sm.counters[i];
  *n_fission = sm.n_sites;
  h->parent_n = A.site_parent ? sm.n_sites : 0;
  rc = status_from_device_error(h, sm);
  if (rc) return rc;
  if (sm.n_sites > out.n) {
    char buf[128];
    snprintf(buf, sizeof buf, "fission bank overflow: %llu sites, capacity %llu", sm.n_sites, (unsigned long long)out.n);
    return fail(h, ABL_ERR_BANK_OVERFLOW, buf);
  }
  if (sm.n_sites > 0) {
    if ((rc = grow_u32(h, h->site_inv, h->inv_cap, out.n)) != 0) return rc;
    site_inverse_kernel<<<grid_for(h, sm.n_sites, 256), 256, 0, s>>>(h->sites, sm.n_sites, h->offsets, out.n, h->site_inv);
    place_sites_kernel<<<grid_for(h, sm.n_sites, 256), 256, 0, s>>>(h->sites, sm.n_sites, h->site_inv, in, out,
                                                                     lane_kernel_call ? h->site_did : nullptr);
    h->launches += 2;
    if (A.site_parent) {
      if ((rc = ensure_parent_info(h, out.n, 0)) != 0) return rc;
      place_parent_info_kernel<<<grid_for(h, sm.n_sites, 256), 256, 0, s>>>(h->site_parent, sm.n_sites, h->site_inv, h->parent_info, h->parent_cap);
      h->launches += 1;
    }
    ABL_CUDA(h, cudaGetLastError());
  }
  if (sample_noise) {
    *n_noise = sm.n_nsites;
    if (sm.n_nsites > noise_out->n) {
      char buf[128];
      snprintf(buf, sizeof buf, "noise bank overflow: %llu particles, capacity %llu", sm.n_nsites, (unsigned long long)noise_out->n);
      return fail(h, ABL_ERR_BANK_OVERFLOW, buf);
    }
    if (sm.n_nsites > 0) {
      if ((rc = grow_u32(h, h->site_inv, h->inv_cap, noise_out->n > out.n ? noise_out->n : out.n)) != 0) return rc;
      site_inverse_kernel<<<grid_for(h, sm.n_nsites, 256), 256, 0, s>>>(h->nsites, sm.n_nsites, h->noffsets, noise_out->n, h->site_inv);
      place_sites_kernel<<<grid_for(h, sm.n_nsites, 256), 256, 0, s>>>(h->nsites, sm.n_nsites, h->site_inv, in, *noise_out, h->nsite_did);
      h->launches += 2;
      ABL_CUDA(h, cudaGetLastError());
    }
  }
  return ABL_OK;
}

}  // namespace

extern "C" {

const char* abl_last_error(abl_handle h) { return h ? h->error.c_str() : g_create_error.c_str(); }

void abl_destroy(abl_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (void* p : h->allocs) cudaFree(p);
  for (int t = 0; t < h->P.ntallies; t++) {
    cudaFree(h->P.tally[t].gen);
    cudaFree(h->P.tally[t].avg);
    cudaFree(h->P.tally[t].var);
  }
  for (void* p : {(void*)h->sites, (void*)h->nfis, (void*)h->offsets, (void*)h->tile_sums, (void*)h->tr_flights, (void*)h->tr_real,
                  (void*)h->tr_virtual, (void*)h->tr_hash, (void*)h->tr_rng, (void*)h->small_dev, (void*)h->secondaries,
                  (void*)h->cancel.sum_pos, (void*)h->cancel.sum_neg, (void*)h->cancel.sum_pos2, (void*)h->cancel.sum_neg2,
                  (void*)h->cancel.count, (void*)h->probe_buf, (void*)h->rng_scratch,
                  (void*)h->nsites, (void*)h->nnoise, (void*)h->noffsets, (void*)h->ntile_sums, (void*)h->site_did, (void*)h->nsite_did, (void*)h->site_inv,
                  (void*)h->site_parent, (void*)h->parent_info})
    if (p) cudaFree(p);
  free_bank(h->stage_in);
  free_bank(h->stage_out);
  if (h->small_host) cudaFreeHost(h->small_host);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->ev_ids) cudaEventDestroy(h->ev_ids);
  if (h->ev_zeroed) cudaEventDestroy(h->ev_zeroed);
  if (h->chunk_ends) cudaFreeHost(h->chunk_ends);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int abl_create(const abl_problem* p, int device, abl_handle* out) {
  if (!p || !out) {
    g_create_error = "null argument";
    return ABL_ERR_INVALID;
  }
  *out = nullptr;
  abl_handle h = new abl_context();
  auto bail = [&](int rc) {
    g_create_error = h->error;
    abl_destroy(h);
    return rc;
  };
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    h->error = "no CUDA device: the B200 backend has no CPU fallback";
    return bail(ABL_ERR_CUDA);
  }
  if (device < 0 || device >= ndev) {
    h->error = "device index out of range";
    return bail(ABL_ERR_INVALID);
  }
  h->device = device;
  int rc = validate(h, p);
  if (rc) return bail(rc);
  auto CU = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
      h->error = std::string(what) + ": " + cudaGetErrorString(e);
      return true;
    }
    return false;
  };
  if (CU(cudaSetDevice(device), "cudaSetDevice")) return bail(ABL_ERR_CUDA);
  cudaDeviceProp prop;
  if (CU(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return bail(ABL_ERR_CUDA);
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  h->smem_optin = (int)prop.sharedMemPerBlockOptin;
  h->smem_total_sm = (int)prop.sharedMemPerMultiprocessor;
  geometry_depth(p, h->geo_frames, h->geo_pads);
  for (int u = 0; u < p->nuniverses; u++)
    if (p->universes[u].type == ABL_UNI_HEX) h->P.has_hex = 1;
  // (off by default: bit-exact, but a warp runs the boundary-condition search as soon as ONE lane needs it, and with the
  // reflector holding 40 % of the histories every warp does -- measured 226 ms against 190 ms per 2e6 histories on config 3)
  h->P.n_bc_planes = getenv("ABEILLE_B200_BC_BOUND") ? boundary_planes(p, h->P.bc_axis, h->P.bc_p0) : 0;
  if (CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return bail(ABL_ERR_CUDA);
  if (CU(cudaEventCreate(&h->ev0), "cudaEventCreate") || CU(cudaEventCreate(&h->ev1), "cudaEventCreate")) return bail(ABL_ERR_CUDA);
  if (CU(cudaMalloc(&h->small_dev, sizeof(DevSmall)), "cudaMalloc")) return bail(ABL_ERR_CUDA);
  if (CU(cudaMallocHost(&h->small_host, sizeof(DevSmall)), "cudaMallocHost")) return bail(ABL_ERR_CUDA);

  if (CU(cudaMalloc(&h->arena, ARENA_CAP), "cudaMalloc")) return bail(ABL_ERR_CUDA);
  h->allocs.push_back(h->arena);
  h->arena_cap = ARENA_CAP;
  if (CU(cudaMemset(h->arena, 0, ARENA_CAP), "cudaMemset")) return bail(ABL_ERR_CUDA);
  {
    smem_base_probe_kernel<<<1, 32, 64, h->stream>>>(reinterpret_cast<unsigned long long*>(h->small_dev));
    unsigned long long base = 0;
    if (CU(cudaMemcpyAsync(&base, h->small_dev, sizeof base, cudaMemcpyDeviceToHost, h->stream), "cudaMemcpyAsync") ||
        CU(cudaStreamSynchronize(h->stream), "smem_base_probe_kernel"))
      return bail(ABL_ERR_CUDA);
    h->smem_generic_base = base;
  }
  DevProblem& P = h->P;
  const int G = p->ngroups, M = p->nmaterials;
  P.mode = p->mode;
  P.branchless = p->mode == ABL_MODE_BRANCHLESS ? p->branchless_flags : 0;
  P.tracking = p->tracking;
  P.G = G;
  P.inner_generations = p->inner_generations;
  P.wgt_cutoff = p->wgt_cutoff;
  P.wgt_survival = p->wgt_survival;
  P.wgt_split = p->wgt_split;
  P.min_energy = p->min_energy;
  P.seed_state = pcg_seed_state(p->rng_seed);
  P.stride = p->rng_stride;
  P.w_noise = p->w_noise;
  P.eta = p->eta;
  P.nsurfaces = p->nsurfaces;
  P.ncells = p->ncells;
  P.nuniverses = p->nuniverses;
  P.root = p->root_universe;
  P.M = M;
#define UP(src, n, dst)                                  \
  if ((rc = upload(h, src, (size_t)(n), &dst)) != 0) return bail(rc)
  UP(p->energy_bounds, G + 1, P.ebounds);
  {
    JumpTable jt;  // LCG jump-ahead by 2^k steps
    jt.mult[0] = ABL_PCG_MULT;
    jt.plus[0] = ABL_PCG_INC;
    for (int k = 1; k < 64; k++) {
      jt.plus[k] = (jt.mult[k - 1] + 1) * jt.plus[k - 1];
      jt.mult[k] = jt.mult[k - 1] * jt.mult[k - 1];
    }
    UP(&jt, 1, P.jump);
  }
  UP(p->surfaces, p->nsurfaces, P.surfaces);
  UP(p->cells, p->ncells, P.cells);
  UP(p->rpn, p->nrpn, P.rpn);
  UP(p->universes, p->nuniverses, P.universes);
  UP(p->universe_cells, p->n_universe_cells, P.ucells);
  UP(p->lattice_tiles, p->n_lattice_tiles, P.tiles);
  {
    std::vector<CellFast> cf = compile_cells(p);
    UP(cf.data(), cf.size(), P.cellfast);
  }
  const size_t MG = (size_t)M * G, MGG = MG * G;
  UP(p->xs_total, MG, P.Et);
  UP(p->xs_absorption, MG, P.Ea);
  UP(p->xs_fission, MG, P.Ef);
  UP(p->xs_elastic, MG, P.Es);
  UP(p->nu_total, MG, P.nu);
  UP(p->nu_delayed, MG, P.nud);
  UP(p->speeds, MG, P.speed);
  UP(p->chi_cdf, MGG, P.chi_cp);
  UP(p->scatter_cdf, MGG, P.ps_cp);
  {
    // default isotropic tables {mu:[-1,1], pdf:[.5,.5], cdf:[0,1]} are marked n = -2: sample_mu has a closed form
    std::vector<abl_angle_table> at(p->angle, p->angle + MGG);
    for (auto& a : at) {
      if (a.n != 2 || a.offset < 0 || a.offset + 2 > p->n_angle_points) continue;
      const double *mu = p->angle_mu + a.offset, *pdf = p->angle_pdf + a.offset, *cdf = p->angle_cdf + a.offset;
      if (mu[0] == -1. && mu[1] == 1. && pdf[0] == 0.5 && pdf[1] == 0.5 && cdf[0] == 0. && cdf[1] == 1.) a.n = -2;
    }
    UP(at.data(), MGG, P.angle);
  }
  UP(p->angle_mu, p->n_angle_points, P.amu);
  UP(p->angle_pdf, p->n_angle_points, P.apdf);
  UP(p->angle_cdf, p->n_angle_points, P.acdf);
  UP(p->delayed_offset, M + 1, P.dg_off);
  {
    const int nd = p->delayed_offset ? p->delayed_offset[M] : 0;
    UP(p->delayed_cdf, nd, P.dg_cp);
    UP(p->delayed_lambda, nd, P.dg_lambda);
  }
  UP(p->fissile, M, P.fissile);
  if (p->sampling_xs) UP(p->sampling_xs, G, P.smp);
  const double* teb = nullptr;
  UP(p->tally_energy_bounds, p->n_tally_energy_bounds, teb);
  P.ntallies = p->ntallies;
  P.n_coll_tallies = P.n_tl_tallies = 0;
  h->tally_g.assign((size_t)p->ntallies, 0);
  for (int t = 0; t < p->ntallies; t++) {
    const abl_mesh_tally& mt = p->tallies[t];
    DevTally& d = P.tally[t];
    d.estimator = mt.estimator;
    d.quantity = mt.quantity;
    d.noise_source = mt.noise_source;
    d.Nx = mt.N[0]; d.Ny = mt.N[1]; d.Nz = mt.N[2];
    d.Ne = mt.n_energy_bins;
    d.ebounds = teb + mt.ebounds_offset;
    d.lowx = mt.low[0]; d.lowy = mt.low[1]; d.lowz = mt.low[2];
    d.hix = mt.hi[0]; d.hiy = mt.hi[1]; d.hiz = mt.hi[2];
    d.dx = (mt.hi[0] - mt.low[0]) / static_cast<double>(mt.N[0]);  // mesh_tally.cpp:66-101
    d.dy = (mt.hi[1] - mt.low[1]) / static_cast<double>(mt.N[1]);
    d.dz = (mt.hi[2] - mt.low[2]) / static_cast<double>(mt.N[2]);
    d.dx_inv = 1. / d.dx;
    d.dy_inv = 1. / d.dy;
    d.dz_inv = 1. / d.dz;
    d.net_weight = mt.net_weight;
    d.size = (uint64_t)d.Ne * d.Nx * d.Ny * d.Nz;
    if (mt.estimator == ABL_EST_COLLISION) P.n_coll_tallies++;
    if (mt.estimator == ABL_EST_TRACK_LENGTH) P.n_tl_tallies++;
    for (double** arr : {&d.gen, &d.avg, &d.var}) {
      if (CU(cudaMalloc(arr, d.size * sizeof(double)), "cudaMalloc(tally)")) return bail(ABL_ERR_CUDA);
      if (CU(cudaMemset(*arr, 0, d.size * sizeof(double)), "cudaMemset(tally)")) return bail(ABL_ERR_CUDA);
    }
  }
  if (p->ntallies > 0) UP(P.tally, p->ntallies, P.tally_dev);
  {
    std::vector<double> mid((size_t)G);
    for (int g = 0; g < G; g++) mid[(size_t)g] = 0.5 * (p->energy_bounds[g] + p->energy_bounds[g + 1]);
    UP(mid.data(), mid.size(), P.gmid);
    std::vector<int32_t> gb((size_t)(p->ntallies > 0 ? p->ntallies : 1) * G, -1);
    for (int t = 0; t < p->ntallies; t++) {
      const abl_mesh_tally& mt = p->tallies[t];
      const double* eb = p->tally_energy_bounds + mt.ebounds_offset;
      for (int g = 0; g < G; g++)
        for (int e = 0; e < mt.n_energy_bins; e++)  // "<= E <=", first match (collision_mesh_tally.cpp:50-56)
          if (eb[e] <= mid[(size_t)g] && mid[(size_t)g] <= eb[e + 1]) {
            gb[(size_t)t * G + g] = e;
            break;
          }
    }
    UP(gb.data(), gb.size(), P.tally_gbin);
  }
  {
    std::vector<double> rf(MG, 0.), sf(MG, 0.), is((size_t)(p->ntallies > 0 ? p->ntallies : 1) * MG, 0.);
    for (int m = 0; m < M; m++)
      for (int g = 0; g < G; g++) {
        const size_t mg = (size_t)m * G + g;
        const double Et = p->xs_total[mg], Ea = p->xs_absorption[mg] + 0.;
        if (p->sampling_xs) rf[mg] = Et / p->sampling_xs[g];
        sf[mg] = 1. - (Ea == 0. ? Ea : Ea / Et);  // ddiv_pos
        for (int t = 0; t < p->ntallies; t++) is[(size_t)t * MG + mg] = 1. / (Et * p->tallies[t].net_weight);
      }
    UP(rf.data(), rf.size(), P.real_frac);
    h->host_Et.assign(p->xs_total, p->xs_total + MG);
    h->k_site_max = 0.;
    for (size_t mg = 0; mg < MG; mg++) {
      const double Ea = p->xs_absorption[mg], nf = p->nu_total[mg] * p->xs_fission[mg];
      if (nf > 0.) h->k_site_max = std::max(h->k_site_max, Ea > 0. ? nf / Ea : 1e6);
    }
    UP(sf.data(), sf.size(), P.surv_frac);
    UP(is.data(), is.size(), P.inv_score);
  }
  P.nsources = p->nsources;
  UP(p->sources, p->nsources, P.sources);
  if (p->nsources >= 2) {
    std::vector<double> w;
    for (int s = 0; s < p->nsources; s++) w.push_back(p->sources[s].weight);
    std::vector<double> cp = discrete_table(w);
    UP(cp.data(), cp.size(), P.source_cp);
  }
  {  // square-oscillation noise sources; the frequency gate of every factor (square_oscillation_noise_source.cpp:85-170)
    std::vector<DevNoiseSrc> ns((size_t)(p->n_noise_sources > 0 ? p->n_noise_sources : 0));
    for (size_t i = 0; i < ns.size(); i++) {
      const abl_noise_source& f = p->noise_sources[i];
      for (int k = 0; k < 3; k++) {
        ns[i].low[k] = f.low[k];
        ns[i].hi[k] = f.hi[k];
      }
      const double w = p->w_noise, w0 = f.angular_frequency;
      const int32_t n = static_cast<int32_t>(std::round(w / w0));
      const double err = (n * w0 - w) / w;
      ns[i].on = ((n == 1 || n == -1) && std::abs(err) < 0.01) ? 1 : 0;
      ns[i].eps_t = f.eps_total;
      ns[i].eps_f_pi = f.eps_fission * ABL_PI;
      ns[i].eps_s_pi = f.eps_scatter * ABL_PI;
      ns[i].vibration = f.type == ABL_NOISE_FLAT_VIBRATION ? 1 : 0;
      if (ns[i].vibration) {  // flat_vibration_noise_source.cpp:66-82,222-244
        if (f.material_pos < 0 || f.material_pos >= M || f.material_neg < 0 || f.material_neg >= M) {
          h->error = "flat-vibration noise source: material index out of range";
          return bail(ABL_ERR_INVALID);
        }
        ns[i].basis = f.basis;
        ns[i].mat_pos = f.material_pos;
        ns[i].mat_neg = f.material_neg;
        ns[i].x0 = 0.5 * (f.low[f.basis] + f.hi[f.basis]);
        ns[i].eps = (f.hi[f.basis] - f.low[f.basis]) / 2.;
        const double verr = ((n * w0) - w) / w;
        ns[i].harmonic = std::abs(verr) > 0.01 ? 0 : n;
        if (ns[i].harmonic < 0) {  // (a negative noise frequency; n = 0 cannot pass the gate above: its relative error is 1)
          h->error = "flat-vibration noise source: negative noise frequency";
          return bail(ABL_ERR_UNSUPPORTED);
        }
      }
    }
    P.n_noise_src = (int32_t)ns.size();
    UP(ns.data(), ns.size(), P.noise_src);
  }
  P.entropy = make_mesh3(p->entropy, teb);
  P.cancel = make_mesh3(p->cancelator, teb);
  P.cancel.kind = p->cancelator.present ? (p->cancelator.kind == ABL_CANCEL_BASIC_EXACT || p->cancelator.kind == ABL_CANCEL_EXACT ? p->cancelator.kind : ABL_CANCEL_APPROXIMATE) : 0;
  P.cancel.beta = p->cancelator.beta;
  P.cancel.sobol = p->cancelator.sobol;
  P.cancel.nsamples = p->cancelator.n_samples;
  P.exact_cancel = (P.cancel.kind == ABL_CANCEL_BASIC_EXACT || P.cancel.kind == ABL_CANCEL_EXACT) ? 1 : 0;
  P.chi_matrix = p->chi_matrix;
  if (P.cancel.kind == ABL_CANCEL_EXACT && p->chi_pdf && p->exact_group_bins && p->n_exact_group_bins >= 1) {
    // ExactMGCancelator's tables: the chi rows themselves and the energy bins (Key::shape[3] is at least 1, exact_mg_cancelator.cpp:105-111)
    double* chi_d = nullptr;
    int32_t* egb_d = nullptr;
    if (CU(cudaMalloc(&chi_d, MGG * sizeof(double)), "cudaMalloc") || CU(cudaMalloc(&egb_d, (size_t)p->n_exact_group_bins * sizeof(int32_t)), "cudaMalloc") ||
        CU(cudaMemcpy(chi_d, p->chi_pdf, MGG * sizeof(double), cudaMemcpyHostToDevice), "cudaMemcpy") ||
        CU(cudaMemcpy(egb_d, p->exact_group_bins, (size_t)p->n_exact_group_bins * sizeof(int32_t), cudaMemcpyHostToDevice), "cudaMemcpy"))
      return bail(ABL_ERR_CUDA);
    P.chi_pdf = chi_d;
    P.egb = egb_d;
    P.cancel.Ne = p->exact_group_bins[0] > 0 ? p->exact_group_bins[0] : 1;
    h->exact_full_ready = true;
  }
#undef UP
  if (CU(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return bail(ABL_ERR_CUDA);
  *out = h;
  return ABL_OK;
}

int abl_device_info(abl_handle h, int* sm_count, int* cc_major, int* cc_minor, uint64_t* kernel_launches) {
  if (!h) return ABL_ERR_INVALID;
  if (sm_count) *sm_count = h->sm_count;
  if (cc_major) *cc_major = h->cc_major;
  if (cc_minor) *cc_minor = h->cc_minor;
  if (kernel_launches) *kernel_launches = h->launches;
  return ABL_OK;
}

int abl_last_transport_kernel(abl_handle h, float* milliseconds, int* grid_blocks, int* block_threads) {
  if (!h) return ABL_ERR_INVALID;
  if (milliseconds) *milliseconds = h->last_kernel_ms;
  if (grid_blocks) *grid_blocks = h->last_grid;
  if (block_threads) *block_threads = h->last_block;
  return ABL_OK;
}

int abl_transport_device(abl_handle h, const abl_bank* bank_dev, const abl_gen_params* params, abl_bank* fission_dev,
                         uint64_t* n_fission, double scores[6], uint64_t counters[8], void* stream) {
  if (!h || !bank_dev || !params || !fission_dev || !n_fission || !scores) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  return transport_impl(h, view_of(bank_dev), params, view_of(fission_dev), n_fission, scores, counters, s);
}

int abl_transport_noise_device(abl_handle h, const abl_bank* bank_dev, const abl_gen_params* params, abl_bank* fission_dev,
                               uint64_t* n_fission, abl_bank* noise_dev, uint64_t* n_noise, double scores[6], uint64_t counters[8],
                               void* stream) {
  if (!h || !bank_dev || !params || !fission_dev || !n_fission || !scores) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  BankView nv{};
  if (noise_dev) nv = view_of(noise_dev);
  return transport_impl(h, view_of(bank_dev), params, view_of(fission_dev), n_fission, scores, counters, s, noise_dev ? &nv : nullptr,
                        n_noise);
}

int abl_bank_weight_magnitude_device(abl_handle h, const abl_bank* bank_dev, double* sum, void* stream) {
  if (!h || !bank_dev || !sum || !bank_dev->wgt || !bank_dev->wgt2) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  ABL_CUDA(h, cudaMemsetAsync(h->small_dev->stats, 0, sizeof(double) * 4, s));
  if (bank_dev->n) {
    weight_magnitude_kernel<<<grid_for(h, bank_dev->n, 256), 256, 0, s>>>(bank_dev->wgt, bank_dev->wgt2, bank_dev->n, h->small_dev->stats);
    h->launches++;
  }
  ABL_CUDA(h, cudaMemcpyAsync(sum, h->small_dev->stats, sizeof(double), cudaMemcpyDeviceToHost, s));
  ABL_CUDA(h, cudaStreamSynchronize(s));
  return ABL_OK;
}

int abl_bank_divide_weights_device(abl_handle h, abl_bank* bank_dev, double divisor, void* stream) {
  if (!h || !bank_dev || !bank_dev->wgt || !bank_dev->wgt2) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  if (bank_dev->n) {
    divide_weights_kernel<<<grid_for(h, bank_dev->n, 256), 256, 0, s>>>(bank_dev->wgt, bank_dev->wgt2, bank_dev->n, divisor);
    h->launches++;
    ABL_CUDA(h, cudaGetLastError());
  }
  return ABL_OK;
}

namespace {
// host bank -> device (streamed behind the kernel for large k-eigenvalue banks), the history kernel, scan and placement: the
// fission bank is left in h->stage_out (capacity cap, wgt2 kept when want_wgt2)
int transport_upload_and_run(abl_handle h, const abl_bank* bank, const abl_gen_params* params, uint64_t cap, bool want_wgt2,
                             uint64_t* n_fission, double scores[6], uint64_t counters[8]) {
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, h->stream);
  const uint64_t N = bank->n;
  int rc;
  // (the recorded capacity is dropped BEFORE the old arrays are freed: a failed allocation must not leave a capacity that
  // lets a later, smaller call skip the reallocation and copy into freed pointers)
  if (N > h->stage_in_cap) {
    h->stage_in_cap = 0;
    if ((rc = alloc_bank(h, h->stage_in, N + N / 4 + 1024)) != 0) return rc;
    h->stage_in_cap = h->stage_in.n;
  }
  if (cap > h->stage_out_cap) {
    h->stage_out_cap = 0;
    if ((rc = alloc_bank(h, h->stage_out, cap)) != 0) return rc;
    h->stage_out_cap = h->stage_out.n;
  }
  BankView in = h->stage_in, out = h->stage_out;
  in.n = N;
  out.n = cap;
  const double* hin[9] = {bank->x, bank->y, bank->z, bank->ux, bank->uy, bank->uz, bank->E, bank->wgt, bank->wgt2};
  double* din[9] = {in.x, in.y, in.z, in.ux, in.uy, in.uz, in.E, in.wgt, in.wgt2};
  for (int k = 0; k < 8; k++)
    if (!hin[k]) return fail(h, ABL_ERR_INVALID, "null bank array");
  if (!bank->id_a) return fail(h, ABL_ERR_INVALID, "null history id array");
  if (!hin[8]) in.wgt2 = nullptr;
  if (!bank->id_b) in.id_b = nullptr;
  if (!bank->id_c) in.id_c = nullptr;
  // Large k-eigenvalue banks are STREAMED: the ids go first (the RNG streams are seeded from them), then the other
  // arrays in row chunks on a second stream, each chunk followed by an update of the arrival counter that the history
  // kernel polls before it loads a row -- the PCIe copy (17 ms for 1e7 particles) hides behind the kernel.
  constexpr int NCHUNK = 16;
  // (only the staged kernel polls the arrival counter of a streamed bank)
  const bool streamed = h->P.mode == ABL_MODE_K_EIGENVALUE && !h->P.exact_cancel && !params->noise && N >= (1u << 18) && h->P.tracking != ABL_TRACK_IMPLICIT_LEAKAGE;
  if (streamed) {
    if (!h->copy_stream) {
      ABL_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
      ABL_CUDA(h, cudaEventCreateWithFlags(&h->ev_ids, cudaEventDisableTiming));
      ABL_CUDA(h, cudaEventCreateWithFlags(&h->ev_zeroed, cudaEventDisableTiming));
      ABL_CUDA(h, cudaMallocHost(&h->chunk_ends, NCHUNK * sizeof(unsigned long long)));
    }
    cudaStream_t cs = h->copy_stream;
    ABL_CUDA(h, cudaMemsetAsync(&h->small_dev->avail, 0, sizeof(unsigned long long), s));
    ABL_CUDA(h, cudaEventRecord(h->ev_zeroed, s));
    ABL_CUDA(h, cudaStreamWaitEvent(cs, h->ev_zeroed, 0));
    ABL_CUDA(h, cudaMemcpyAsync(in.id_a, bank->id_a, N * 8, cudaMemcpyHostToDevice, cs));
    if (bank->id_b) ABL_CUDA(h, cudaMemcpyAsync(in.id_b, bank->id_b, N * 8, cudaMemcpyHostToDevice, cs));
    ABL_CUDA(h, cudaEventRecord(h->ev_ids, cs));
    ABL_CUDA(h, cudaStreamWaitEvent(s, h->ev_ids, 0));
    for (int c = 0; c < NCHUNK; c++) {
      const uint64_t r0 = N * c / NCHUNK, r1 = N * (c + 1) / NCHUNK;
      if (r1 > r0) {
        for (int k = 0; k < 9; k++)
          if (hin[k]) ABL_CUDA(h, cudaMemcpyAsync(din[k] + r0, hin[k] + r0, (r1 - r0) * sizeof(double), cudaMemcpyHostToDevice, cs));
        if (bank->id_c) ABL_CUDA(h, cudaMemcpyAsync(in.id_c + r0, bank->id_c + r0, (r1 - r0) * 8, cudaMemcpyHostToDevice, cs));
      }
      h->chunk_ends[c] = r1;
      ABL_CUDA(h, cudaMemcpyAsync(&h->small_dev->avail, &h->chunk_ends[c], sizeof(unsigned long long), cudaMemcpyHostToDevice, cs));
    }
  } else {
    for (int k = 0; k < 9; k++)
      if (hin[k] && N) ABL_CUDA(h, cudaMemcpyAsync(din[k], hin[k], N * sizeof(double), cudaMemcpyHostToDevice, s));
    if (N) ABL_CUDA(h, cudaMemcpyAsync(in.id_a, bank->id_a, N * 8, cudaMemcpyHostToDevice, s));
    if (bank->id_b && N) ABL_CUDA(h, cudaMemcpyAsync(in.id_b, bank->id_b, N * 8, cudaMemcpyHostToDevice, s));
    if (bank->id_c && N) ABL_CUDA(h, cudaMemcpyAsync(in.id_c, bank->id_c, N * 8, cudaMemcpyHostToDevice, s));
  }
  h->streamed_input = streamed;
  if (!want_wgt2) out.wgt2 = nullptr;
  rc = transport_impl(h, in, params, out, n_fission, scores, counters, s);
  h->streamed_input = false;
  if (rc) {
    if (streamed) cudaStreamSynchronize(h->copy_stream);
    return rc;
  }
  return ABL_OK;
}

// device fission bank (h->stage_out, m rows) -> host arrays
int transport_download(abl_handle h, uint64_t m, abl_bank* fission_out, cudaStream_t s) {
  BankView out = h->stage_out;
  if (m) {
    double* hout[9] = {fission_out->x, fission_out->y, fission_out->z, fission_out->ux, fission_out->uy, fission_out->uz,
                       fission_out->E, fission_out->wgt, fission_out->wgt2};
    double* dout[9] = {out.x, out.y, out.z, out.ux, out.uy, out.uz, out.E, out.wgt, out.wgt2};
    for (int k = 0; k < 9; k++)
      if (hout[k] && dout[k]) ABL_CUDA(h, cudaMemcpyAsync(hout[k], dout[k], m * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (fission_out->id_a) ABL_CUDA(h, cudaMemcpyAsync(fission_out->id_a, out.id_a, m * 8, cudaMemcpyDeviceToHost, s));
    if (fission_out->id_b) ABL_CUDA(h, cudaMemcpyAsync(fission_out->id_b, out.id_b, m * 8, cudaMemcpyDeviceToHost, s));
    if (fission_out->id_c) ABL_CUDA(h, cudaMemcpyAsync(fission_out->id_c, out.id_c, m * 8, cudaMemcpyDeviceToHost, s));
  }
  ABL_CUDA(h, cudaStreamSynchronize(s));
  return ABL_OK;
}

}  // namespace

int abl_transport(abl_handle h, const abl_bank* bank, const abl_gen_params* params, abl_bank* fission_out, uint64_t* n_fission,
                  double scores[6], uint64_t counters[8]) {
  if (!h || !bank || !params || !fission_out || !n_fission || !scores) return ABL_ERR_INVALID;
  const int rc = transport_upload_and_run(h, bank, params, fission_out->n, fission_out->wgt2 != nullptr, n_fission, scores, counters);
  if (rc) return rc;
  return transport_download(h, *n_fission, fission_out, use_stream(h, h->stream));
}

// abl_transport in two halves, for a caller that normalises the fission bank before it uses it (PowerIterator::run:
// normalize_weights and the hand-out of fresh history ids, src/power_iterator.cpp:397-399,538-569): begin leaves the fission
// bank on the device and returns its size, the scores and the weight sums the normalisation needs (positive / negative
// particle counts and weight sums, as abl_bank_weight_stats_device); finish applies the caller's factor and first history id
// on the device and copies the finished bank to the host arrays.  The bytes that cross PCIe are the same as in abl_transport;
// the caller makes no pass over the bank on the host.
int abl_transport_begin(abl_handle h, const abl_bank* bank, const abl_gen_params* params, uint64_t capacity, uint64_t* n_fission,
                        double scores[6], uint64_t counters[8], double weight_stats[4]) {
  if (!h || !bank || !params || !n_fission || !scores || !weight_stats) return ABL_ERR_INVALID;
  h->pending_rows = 0;
  int rc = transport_upload_and_run(h, bank, params, capacity, false, n_fission, scores, counters);
  if (rc) return rc;
  BankView out = h->stage_out;
  out.n = *n_fission;
  abl_bank ob{};
  ob.n = out.n;
  ob.wgt = out.wgt;
  if ((rc = abl_bank_weight_stats_device(h, &ob, weight_stats, nullptr)) != 0) return rc;
  h->pending_rows = *n_fission;
  return ABL_OK;
}

int abl_transport_finish(abl_handle h, double weight_factor, uint64_t first_history_id, abl_bank* fission_out) {
  if (!h || !fission_out) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, h->stream);
  const uint64_t m = h->pending_rows;
  if (m > fission_out->n) return fail(h, ABL_ERR_BANK_OVERFLOW, "abl_transport_finish: host bank smaller than the fission bank");
  if (m) {
    BankView out = h->stage_out;
    out.n = m;
    scale_weights_kernel<<<grid_for(h, m, 256), 256, 0, s>>>(out.wgt, m, weight_factor);
    to_particles_kernel<<<grid_for(h, m, 256), 256, 0, s>>>(out, first_history_id);
    h->launches += 2;
    ABL_CUDA(h, cudaGetLastError());
  }
  h->pending_rows = 0;
  abl_bank want = *fission_out;
  want.id_c = nullptr;  // (after to_particles: id_a = fresh history ids, id_b = family ids; the pcg32 column is not a result)
  want.wgt2 = nullptr;
  return transport_download(h, m, &want, s);
}

int abl_transport_noise(abl_handle h, const abl_bank* bank, const abl_gen_params* params, abl_bank* fission_out, uint64_t* n_fission,
                        abl_bank* noise_out, uint64_t* n_noise, double scores[6], uint64_t counters[8]) {
  if (!h || !bank || !params || !fission_out || !n_fission || !noise_out || !n_noise || !scores) return ABL_ERR_INVALID;
  for (const abl_bank* b : {(const abl_bank*)fission_out, (const abl_bank*)noise_out})
    if (!b->x || !b->y || !b->z || !b->ux || !b->uy || !b->uz || !b->E || !b->wgt || !b->wgt2 || !b->id_a || !b->id_b || !b->id_c)
      return fail(h, ABL_ERR_INVALID, "abl_transport_noise: output banks need all twelve arrays");
  if (!bank->x || !bank->y || !bank->z || !bank->ux || !bank->uy || !bank->uz || !bank->E || !bank->wgt || !bank->id_a)
    return fail(h, ABL_ERR_INVALID, "null bank array");
  ABL_CUDA(h, cudaSetDevice(h->device));
  // noise runs are small (the shipped deck: 1e5 particles per generation): plain staged copies through device banks
  abl_bank din{}, dfis{}, dnoise{};
  int rc = abl_bank_alloc_device(h, bank->n ? bank->n : 1, &din);
  if (!rc) rc = abl_bank_alloc_device(h, fission_out->n ? fission_out->n : 1, &dfis);
  if (!rc) rc = abl_bank_alloc_device(h, noise_out->n ? noise_out->n : 1, &dnoise);
  if (!rc) rc = abl_bank_upload(h, bank, &din);  // (null wgt2 -> zeros; null id_b / id_c are skipped)
  if (!rc) {
    abl_bank dview = din;
    dview.n = bank->n;
    if (!bank->id_b) dview.id_b = nullptr;
    if (!bank->id_c) dview.id_c = nullptr;  // the device seeds the streams from seed / stride / history id
    rc = abl_transport_noise_device(h, &dview, params, &dfis, n_fission, &dnoise, n_noise, scores, counters, nullptr);
  }
  if (!rc && *n_fission) rc = abl_bank_download(h, &dfis, *n_fission, fission_out);
  if (!rc && *n_noise) rc = abl_bank_download(h, &dnoise, *n_noise, noise_out);
  const std::string keep = h->error;
  if (din.x) abl_bank_free_device(h, &din);
  if (dfis.x) abl_bank_free_device(h, &dfis);
  if (dnoise.x) abl_bank_free_device(h, &dnoise);
  if (rc) h->error = keep;
  return rc;
}

int abl_get_trace(abl_handle h, uint64_t n, abl_trace* out) {
  if (!h || !out) return ABL_ERR_INVALID;
  if (n > h->trace_n) return fail(h, ABL_ERR_INVALID, "no trace of that size: pass params.trace = 1 to the transport call");
  ABL_CUDA(h, cudaSetDevice(h->device));
  if (out->flights) ABL_CUDA(h, cudaMemcpy(out->flights, h->tr_flights, n * 4, cudaMemcpyDeviceToHost));
  if (out->real) ABL_CUDA(h, cudaMemcpy(out->real, h->tr_real, n * 4, cudaMemcpyDeviceToHost));
  if (out->virt) ABL_CUDA(h, cudaMemcpy(out->virt, h->tr_virtual, n * 4, cudaMemcpyDeviceToHost));
  if (out->fission) ABL_CUDA(h, cudaMemcpy(out->fission, h->nfis, n * 4, cudaMemcpyDeviceToHost));
  if (out->hash) ABL_CUDA(h, cudaMemcpy(out->hash, h->tr_hash, n * 8, cudaMemcpyDeviceToHost));
  if (out->rng_state) ABL_CUDA(h, cudaMemcpy(out->rng_state, h->tr_rng, n * 8, cudaMemcpyDeviceToHost));
  return ABL_OK;
}

// ---- tallies -----------------------------------------------------------------------------------------------------------
int abl_tally_count(abl_handle h) { return h ? h->P.ntallies : ABL_ERR_INVALID; }

int abl_tally_shape(abl_handle h, int t, uint64_t shape4[4]) {
  if (!h || t < 0 || t >= h->P.ntallies) return ABL_ERR_INVALID;
  const DevTally& d = h->P.tally[t];
  shape4[0] = (uint64_t)d.Ne; shape4[1] = (uint64_t)d.Nx; shape4[2] = (uint64_t)d.Ny; shape4[3] = (uint64_t)d.Nz;
  return ABL_OK;
}

int abl_tallies_record(abl_handle h, double multiplier) {
  if (!h) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  use_stream(h, h->stream);
  for (int t = 0; t < h->P.ntallies; t++) {
    DevTally& d = h->P.tally[t];
    h->tally_g[(size_t)t]++;
    tally_record_kernel<<<grid_for(h, d.size, 256), 256, 0, h->stream>>>(d.gen, d.avg, d.var, d.size, multiplier,
                                                                          static_cast<double>(h->tally_g[(size_t)t]));
    h->launches++;
  }
  ABL_CUDA(h, cudaGetLastError());
  ABL_CUDA(h, cudaStreamSynchronize(h->stream));
  return ABL_OK;
}

int abl_tallies_clear(abl_handle h) {
  if (!h) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  use_stream(h, h->stream);
  for (int t = 0; t < h->P.ntallies; t++)
    ABL_CUDA(h, cudaMemsetAsync(h->P.tally[t].gen, 0, h->P.tally[t].size * sizeof(double), h->stream));
  ABL_CUDA(h, cudaStreamSynchronize(h->stream));
  return ABL_OK;
}

int abl_tally_device_ptr(abl_handle h, int t, int which, double** out_dev, uint64_t* n) {
  if (!h || t < 0 || t >= h->P.ntallies || which < 0 || which > 2 || !out_dev) return ABL_ERR_INVALID;
  DevTally& d = h->P.tally[t];
  *out_dev = which == 0 ? d.gen : (which == 1 ? d.avg : d.var);
  if (n) *n = d.size;
  return ABL_OK;
}

int abl_tally_fetch(abl_handle h, int t, int which, double* out_host) {
  if (!h || t < 0 || t >= h->P.ntallies || which < 0 || which > 3 || !out_host) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  use_stream(h, h->stream);
  DevTally& d = h->P.tally[t];
  if (which < 3) {
    const double* src = which == 0 ? d.gen : (which == 1 ? d.avg : d.var);
    ABL_CUDA(h, cudaMemcpy(out_host, src, d.size * sizeof(double), cudaMemcpyDeviceToHost));
    return ABL_OK;
  }
  double* tmp = nullptr;  // std = sqrt(var / g)  (mesh_tally.cpp:195-197)
  ABL_CUDA(h, cudaMalloc(&tmp, d.size * sizeof(double)));
  tally_std_kernel<<<grid_for(h, d.size, 256), 256, 0, h->stream>>>(d.var, tmp, d.size, static_cast<double>(h->tally_g[(size_t)t]));
  h->launches++;
  cudaError_t e = cudaMemcpyAsync(out_host, tmp, d.size * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  ABL_CUDA(h, e);
  return ABL_OK;
}

// ---- inter-generation pipeline --------------------------------------------------------------------------------------------
int abl_sample_source_device(abl_handle h, uint64_t n, uint64_t first_history_id, abl_bank* bank_dev, void* stream) {
  if (!h || !bank_dev) return ABL_ERR_INVALID;
  if (h->P.nsources < 1) return fail(h, ABL_ERR_INVALID, "problem has no sources");
  if (bank_dev->n < n) return fail(h, ABL_ERR_INVALID, "bank capacity too small");
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  BankView b = view_of(bank_dev);
  b.n = n;
  ABL_CUDA(h, cudaMemsetAsync(h->small_dev, 0, sizeof(DevSmall), s));
  sample_source_kernel<<<grid_for(h, n, 128), 128, 0, s>>>(h->P, b, first_history_id, h->small_dev->error);
  h->launches++;
  ABL_CUDA(h, cudaGetLastError());
  ABL_CUDA(h, cudaMemcpyAsync(h->small_host, h->small_dev, sizeof(DevSmall), cudaMemcpyDeviceToHost, s));
  ABL_CUDA(h, cudaStreamSynchronize(s));
  return status_from_device_error(h, *h->small_host);
}

int abl_bank_weight_stats_device(abl_handle h, const abl_bank* bank_dev, double stats[4], void* stream) {
  if (!h || !bank_dev || !stats) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  ABL_CUDA(h, cudaMemsetAsync(h->small_dev->stats, 0, sizeof(double) * 4, s));
  if (bank_dev->n) {
    weight_stats_kernel<<<grid_for(h, bank_dev->n, 256), 256, 0, s>>>(bank_dev->wgt, bank_dev->n, h->small_dev->stats);
    h->launches++;
  }
  ABL_CUDA(h, cudaMemcpyAsync(h->small_host->stats, h->small_dev->stats, sizeof(double) * 4, cudaMemcpyDeviceToHost, s));
  ABL_CUDA(h, cudaStreamSynchronize(s));
  for (int i = 0; i < 4; i++) stats[i] = h->small_host->stats[i];
  return ABL_OK;
}

int abl_bank_moments_device(abl_handle h, const abl_bank* bank_dev, const double origin[3], double moments[5], void* stream) {
  if (!h || !bank_dev || !origin || !moments) return ABL_ERR_INVALID;
  if (bank_dev->n && (!bank_dev->x || !bank_dev->y || !bank_dev->z || !bank_dev->wgt)) return fail(h, ABL_ERR_INVALID, "null bank array");
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  ABL_CUDA(h, cudaMemsetAsync(h->small_dev->moments, 0, sizeof(double) * 8, s));
  if (bank_dev->n) {
    bank_moments_kernel<<<grid_for(h, bank_dev->n, 256), 256, 0, s>>>(bank_dev->x, bank_dev->y, bank_dev->z, bank_dev->wgt, bank_dev->n, origin[0],
                                                                     origin[1], origin[2], h->small_dev->moments);
    h->launches++;
  }
  ABL_CUDA(h, cudaMemcpyAsync(h->small_host->moments, h->small_dev->moments, sizeof(double) * 8, cudaMemcpyDeviceToHost, s));
  ABL_CUDA(h, cudaStreamSynchronize(s));
  for (int i = 0; i < 5; i++) moments[i] = h->small_host->moments[i];
  return ABL_OK;
}

int abl_bank_scale_weights_device(abl_handle h, abl_bank* bank_dev, double factor, void* stream) {
  if (!h || !bank_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  if (bank_dev->n) {
    scale_weights_kernel<<<grid_for(h, bank_dev->n, 256), 256, 0, s>>>(bank_dev->wgt, bank_dev->n, factor);
    h->launches++;
  }
  ABL_CUDA(h, cudaGetLastError());
  return ABL_OK;
}

int abl_bank_to_particles_device(abl_handle h, abl_bank* bank_dev, uint64_t first_history_id, void* stream) {
  if (!h || !bank_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  if (bank_dev->n) {
    to_particles_kernel<<<grid_for(h, bank_dev->n, 256), 256, 0, s>>>(view_of(bank_dev), first_history_id);
    h->launches++;
  }
  ABL_CUDA(h, cudaGetLastError());
  return ABL_OK;
}

int abl_entropy_bin_device(abl_handle h, const abl_bank* bank_dev, double* bins_dev, double* total_dev, void* stream) {
  if (!h || !bank_dev || !bins_dev || !total_dev) return ABL_ERR_INVALID;
  if (!h->P.entropy.present) return fail(h, ABL_ERR_INVALID, "problem has no entropy mesh");
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  if (bank_dev->n) {
    const uint64_t nbins = (uint64_t)h->P.entropy.Nx * h->P.entropy.Ny * h->P.entropy.Nz;
    const int nshared = nbins <= 4096 ? (int)nbins : 0;
    entropy_bin_kernel<<<grid_for(h, bank_dev->n, 256), 256, nshared * sizeof(double), s>>>(h->P.entropy, view_of(bank_dev), bins_dev,
                                                                                           total_dev, nshared);
    h->launches++;
  }
  ABL_CUDA(h, cudaGetLastError());
  return ABL_OK;
}

int abl_score_source_device(abl_handle h, const abl_bank* bank_dev, int noise_source, void* stream) {
  if (!h || !bank_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  for (int t = 0; t < h->P.ntallies; t++) {
    const DevTally& d = h->P.tally[t];
    if (d.estimator != ABL_EST_SOURCE || (d.noise_source != 0) != (noise_source != 0) || bank_dev->n == 0) continue;
    score_source_kernel<<<grid_for(h, bank_dev->n, 256), 256, 0, s>>>(d, view_of(bank_dev));
    h->launches++;
  }
  ABL_CUDA(h, cudaGetLastError());
  return ABL_OK;
}

namespace {
int ensure_cancel_bins(abl_handle h, uint64_t* nbins_out) {
  const DevMesh3& m = h->P.cancel;
  if (!m.present || m.kind == ABL_CANCEL_BASIC_EXACT || m.kind == ABL_CANCEL_EXACT) return fail(h, ABL_ERR_INVALID, "problem has no approximate cancelator");
  const uint64_t nbins = (uint64_t)m.Nx * m.Ny * m.Nz * m.Ne;
  if (!h->cancel.count) {
    ABL_CUDA(h, cudaMalloc(&h->cancel.count, nbins * sizeof(uint32_t)));
    ABL_CUDA(h, cudaMemset(h->cancel.count, 0, nbins * sizeof(uint32_t)));
    for (double** p : {&h->cancel.sum_pos, &h->cancel.sum_neg, &h->cancel.sum_pos2, &h->cancel.sum_neg2}) {
      ABL_CUDA(h, cudaMalloc(p, nbins * sizeof(double)));
      ABL_CUDA(h, cudaMemset(*p, 0, nbins * sizeof(double)));
    }
  }
  if (nbins_out) *nbins_out = nbins;
  return ABL_OK;
}
}  // namespace

// accumulate -> [all-reduce of the bins across GPUs: abl_cancel_bins_device] -> apply (which also re-zeroes the bins)
int abl_cancel_accumulate_device(abl_handle h, const abl_bank* bank_dev, void* stream) {
  if (!h || !bank_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_cancel_bins(h, nullptr);
  if (rc) return rc;
  if (bank_dev->n == 0) return ABL_OK;
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  const BankView b = view_of(bank_dev);
  cancel_accumulate_kernel<<<grid_for(h, b.n, 256), 256, 0, s>>>(h->P.cancel, b, h->cancel);
  h->launches++;
  ABL_CUDA(h, cudaGetLastError());
  return ABL_OK;
}

int abl_cancel_apply_device(abl_handle h, abl_bank* bank_dev, void* stream) {
  if (!h || !bank_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_cancel_bins(h, nullptr);
  if (rc) return rc;
  if (bank_dev->n == 0) return ABL_OK;
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  const BankView b = view_of(bank_dev);
  const int grid = grid_for(h, b.n, 256);
  cancel_apply_kernel<<<grid, 256, 0, s>>>(h->P.cancel, b, h->cancel);
  cancel_reset_kernel<<<grid, 256, 0, s>>>(h->P.cancel, b, h->cancel);
  h->launches += 2;
  ABL_CUDA(h, cudaGetLastError());
  return ABL_OK;
}

int abl_cancel_bins_device(abl_handle h, double* sums_dev[4], uint32_t** count_dev, uint64_t* nbins) {
  if (!h || !sums_dev || !count_dev || !nbins) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_cancel_bins(h, nbins);
  if (rc) return rc;
  sums_dev[0] = h->cancel.sum_pos; sums_dev[1] = h->cancel.sum_neg; sums_dev[2] = h->cancel.sum_pos2; sums_dev[3] = h->cancel.sum_neg2;
  *count_dev = h->cancel.count;
  return ABL_OK;
}

int abl_cancel_device(abl_handle h, abl_bank* bank_dev, void* stream) {
  int rc = abl_cancel_accumulate_device(h, bank_dev, stream);
  if (rc) return rc;
  return abl_cancel_apply_device(h, bank_dev, stream);
}

int abl_bank_gather_device(abl_handle h, const abl_bank* src_dev, const uint32_t* rows_host, const double* wgts_host, uint64_t n,
                           abl_bank* dst_dev, void* stream) {
  if (!h || !src_dev || !dst_dev || (n && !rows_host)) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  dst_dev->n = n;
  if (n == 0) return ABL_OK;
  uint32_t* rows_d = nullptr;
  double* wgts_d = nullptr;
  ABL_CUDA(h, cudaMalloc(&rows_d, n * sizeof(uint32_t)));
  if (wgts_host) ABL_CUDA(h, cudaMalloc(&wgts_d, n * sizeof(double)));
  ABL_CUDA(h, cudaMemcpyAsync(rows_d, rows_host, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  if (wgts_host) ABL_CUDA(h, cudaMemcpyAsync(wgts_d, wgts_host, n * sizeof(double), cudaMemcpyHostToDevice, s));
  bank_gather_kernel<<<grid_for(h, n, 256), 256, 0, s>>>(view_of(src_dev), rows_d, wgts_d, n, view_of(dst_dev));
  h->launches++;
  ABL_CUDA(h, cudaGetLastError());
  ABL_CUDA(h, cudaStreamSynchronize(s));
  cudaFree(rows_d);
  if (wgts_d) cudaFree(wgts_d);
  return ABL_OK;
}

// ---- BasicExactMGCancelator ------------------------------------------------------------------------------------------------------
int abl_parent_info_download(abl_handle h, uint64_t n, double* x, double* y, double* z, double* esmp) {
  if (!h) return ABL_ERR_INVALID;
  if (!h->P.exact_cancel) return fail(h, ABL_ERR_INVALID, "problem has no exact cancelator");
  if (n > h->parent_n) return fail(h, ABL_ERR_INVALID, "more rows than the last fission bank holds");
  ABL_CUDA(h, cudaSetDevice(h->device));
  double* dst[4] = {x, y, z, esmp};
  for (int q = 0; q < 4; q++)
    if (dst[q] && n) ABL_CUDA(h, cudaMemcpy(dst[q], h->parent_info + q * h->parent_cap, n * sizeof(double), cudaMemcpyDeviceToHost));
  return ABL_OK;
}

int abl_parent_state_download(abl_handle h, uint64_t n, double* ux, double* uy, double* uz, double* e_before_last_scatter, double* e_parent,
                              double* was_virtual) {
  if (!h) return ABL_ERR_INVALID;
  if (!h->P.exact_cancel) return fail(h, ABL_ERR_INVALID, "problem has no exact cancelator");
  if (n > h->parent_n) return fail(h, ABL_ERR_INVALID, "more rows than the last fission bank holds");
  ABL_CUDA(h, cudaSetDevice(h->device));
  double* dst[6] = {ux, uy, uz, e_before_last_scatter, e_parent, was_virtual};
  for (int q = 0; q < 6; q++)
    if (dst[q] && n) ABL_CUDA(h, cudaMemcpy(dst[q], h->parent_info + (4 + q) * h->parent_cap, n * sizeof(double), cudaMemcpyDeviceToHost));
  return ABL_OK;
}

int abl_cancel_exact_device(abl_handle h, abl_bank* bank_dev, uint64_t capacity, uint64_t rng2[2], void* stream) {
  if (!h || !bank_dev || !rng2) return ABL_ERR_INVALID;
  const DevMesh3& m = h->P.cancel;
  const bool full = m.present && m.kind == ABL_CANCEL_EXACT;  // ExactMGCancelator: averages always, Sobol points, no engine offsets
  const bool timing = getenv("ABEILLE_B200_EXACT_TIMING") != nullptr;  // (development aid: where a call spends its time)
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
  const auto t_start = now();
  double t_cols = 0., t_bins = 0., t_avg = 0., t_cancel = 0.;
  if (full && !h->exact_full_ready)
    return fail(h, ABL_ERR_INVALID, "cancelator type exact needs abl_problem::chi_pdf and exact_group_bins");
  if (full && (double)m.Nx * m.Ny * m.Nz * m.Ne >= 2147483647.)
    return fail(h, ABL_ERR_UNSUPPORTED, "cancelator type exact: more than 2^31 bins");
  if (!m.present || (m.kind != ABL_CANCEL_BASIC_EXACT && !full)) return fail(h, ABL_ERR_INVALID, "problem has no exact cancelator");
  const uint64_t n = bank_dev->n;
  if (n != h->parent_n) return fail(h, ABL_ERR_INVALID, "abl_cancel_exact_device takes the fission bank of the last transport call");
  if (rng2[1] != 5ULL) return fail(h, ABL_ERR_UNSUPPORTED, "the global engine must be on stream 2 (settings::initialize_global_rng)");
  if ((!full && m.beta == ABL_BETA_ZERO) || n == 0) return ABL_OK;  // perform_cancellation / get_new_particles return at once (:458, :559)
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  const BankView b = view_of(bank_dev);
  // (1) per-particle columns on the device
  int32_t *key_d = nullptr, *mat_d = nullptr;
  double *f_d = nullptr, *fmin_d = nullptr;
  auto release = [&]() {
    for (void* p : {(void*)key_d, (void*)mat_d, (void*)f_d, (void*)fmin_d})
      if (p) cudaFree(p);
  };
  auto CU = [&](cudaError_t e, const char* what) {
    if (e == cudaSuccess) return false;
    h->error = std::string(what) + ": " + cudaGetErrorString(e);
    return true;
  };
  if (CU(cudaMalloc(&key_d, n * 4), "cudaMalloc") || CU(cudaMalloc(&mat_d, n * 4), "cudaMalloc") || CU(cudaMalloc(&f_d, n * 8), "cudaMalloc") ||
      CU(cudaMalloc(&fmin_d, n * 8), "cudaMalloc")) {
    release();
    return ABL_ERR_CUDA;
  }
  if (full) exact_full_prepare_kernel<<<grid_for(h, n, 128), 128, 0, s>>>(h->P, m, b, h->parent_info, h->parent_cap, n, key_d, mat_d, f_d);
  else exact_prepare_kernel<<<grid_for(h, n, 128), 128, 0, s>>>(h->P, m, b, h->parent_info, h->parent_cap, n, key_d, mat_d, f_d, fmin_d);
  h->launches++;
  std::vector<int32_t> key(n), mat(n);
  std::vector<double> f(n), fmin(n), w(n), w2(n, 0.);
  bool bad = CU(cudaMemcpyAsync(key.data(), key_d, n * 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync") ||
             CU(cudaMemcpyAsync(mat.data(), mat_d, n * 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync") ||
             CU(cudaMemcpyAsync(f.data(), f_d, n * 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync") ||
             (!full && CU(cudaMemcpyAsync(fmin.data(), fmin_d, n * 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync")) ||
             CU(cudaMemcpyAsync(w.data(), b.wgt, n * 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync") ||
             (b.wgt2 && CU(cudaMemcpyAsync(w2.data(), b.wgt2, n * 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync")) ||
             CU(cudaStreamSynchronize(s), "exact_prepare_kernel");
  release();
  if (bad) return ABL_ERR_CUDA;
  t_cols = ms_since(t_start);
  // (2) the bins, replayed in the reference's order (add_particle :68-108, perform_cancellation :457-555, cancel_bin :408-455)
  // The maps see exactly the insertions the reference's add_particle makes (a key when it first appears, a material when it first
  // appears under its key) in the same order, so they are walked in the same order afterwards; the per-particle lookups go through
  // a dense table key -> slot instead of the hash tables (node addresses of an unordered_map survive its rehashes).
  auto& bins = h->exact_bins;
  struct KeySlot {
    std::unordered_map<int, abl_context::ExactBin>* inner;
    std::vector<std::pair<int, abl_context::ExactBin*>> mats;
  };
  std::vector<KeySlot> slots;
  const uint64_t total_keys = (uint64_t)m.Nx * m.Ny * m.Nz * (full ? (uint64_t)m.Ne : 1ULL);
  std::vector<int32_t>& slot_of_key = h->exact_slot_of_key;
  if (slot_of_key.size() < total_keys) slot_of_key.assign(total_keys, -1);
  std::vector<int32_t> touched_keys;
  for (uint64_t i = 0; i < n; i++) {
    const int32_t k = key[i];
    if (k < 0) continue;
    abl_context::ExactBin* bin = nullptr;
    if ((uint64_t)k < total_keys) {
      int32_t sl = slot_of_key[(size_t)k];
      if (sl < 0) {
        sl = (int32_t)slots.size();
        slot_of_key[(size_t)k] = sl;
        touched_keys.push_back(k);
        auto it = bins.find(k);  // (cannot be there: clear()ed at the end of every call; kept for the shape of add_particle)
        if (it == bins.end()) it = bins.emplace(k, std::unordered_map<int, abl_context::ExactBin>()).first;
        slots.push_back(KeySlot{&it->second, {}});
      }
      KeySlot& ks = slots[(size_t)sl];
      for (auto& mp : ks.mats)
        if (mp.first == mat[i]) { bin = mp.second; break; }
      if (!bin) {
        bin = &ks.inner->emplace(mat[i], abl_context::ExactBin()).first->second;
        ks.mats.emplace_back(mat[i], bin);
      }
    } else {  // (a site exactly on the upper face of the `type: exact` mesh: an index one past the end, as in the reference)
      auto it = bins.find(k);
      if (it == bins.end()) it = bins.emplace(k, std::unordered_map<int, abl_context::ExactBin>()).first;
      auto jt = it->second.find(mat[i]);
      if (jt == it->second.end()) jt = it->second.emplace(mat[i], abl_context::ExactBin()).first;
      bin = &jt->second;
    }
    bin->particles.push_back(i);
    bin->W += w[i];
    bin->W2 += w2[i];
  }
  for (int32_t k : touched_keys) slot_of_key[(size_t)k] = -1;
  t_bins = ms_since(t_start);
  // bins with two or more particles of both signs are the ones cancel_bin works on
  struct Work { abl_context::ExactBin* bin; int key, mat; bool w1, w2; uint64_t first, advance; int can_cancel; };
  std::vector<Work> work;
  const bool averages = full || m.beta == ABL_BETA_AVERAGE_F || m.beta == ABL_BETA_AVERAGE_G;
  const int beta = full ? ABL_BETA_AVERAGE_G : m.beta;  // ExactMGCancelator::get_beta is the average-g rule (:237-247)
  uint64_t seed_advance = 0;  // per-bin offsets into the global engine's sequence (:476-490)
  const uint64_t max_rn_per_part = (uint64_t)m.nsamples * 100ULL * (beta == ABL_BETA_AVERAGE_G ? 2 : 1);
  std::vector<unsigned long long> rows;
  for (auto& kb : bins)
    for (auto& mb : kb.second) {
      abl_context::ExactBin& bin = mb.second;
      const uint64_t my_advance = seed_advance;
      seed_advance += bin.particles.size() * max_rn_per_part;
      if (bin.particles.size() > 1) {
        bool p1 = false, n1 = false, p2 = false, n2 = false;
        for (uint64_t i : bin.particles) {
          if (w[i] > 0.) p1 = true; else if (w[i] < 0.) n1 = true;
          if (w2[i] > 0.) p2 = true; else if (w2[i] < 0.) n2 = true;
          if (p1 && n1 && p2 && n2) break;
        }
        if ((p1 && n1) || (p2 && n2)) {
          work.push_back({&bin, kb.first, mb.first, p1 && n1, p2 && n2, rows.size(), my_advance, 1});
          rows.insert(rows.end(), bin.particles.begin(), bin.particles.end());
        }
      }
    }
  // get_averages / get_averages_sobol (:264-364) of those bins on the device
  std::vector<double> avg_f, avg_finv;
  if (averages && !work.empty()) {
    std::vector<unsigned long long> desc(5 * work.size());
    for (size_t q = 0; q < work.size(); q++) {
      desc[5 * q] = (unsigned long long)work[q].key; desc[5 * q + 1] = (unsigned long long)(long long)work[q].mat;
      desc[5 * q + 2] = work[q].first; desc[5 * q + 3] = work[q].bin->particles.size(); desc[5 * q + 4] = work[q].advance;
    }
    SobolMatrices3 SM;  // the first three dimensions of the Joe-Kuo sequence (vendor/sobol), from the recurrence (oracle/orc_main.cpp)
    {
      unsigned long long d1[53], d2[53];
      d1[1] = 1;
      for (int i = 2; i <= 52; i++) d1[i] = (2 * d1[i - 1]) ^ d1[i - 1];
      d2[1] = 1; d2[2] = 3;
      for (int i = 3; i <= 52; i++) d2[i] = (2 * d2[i - 1]) ^ (4 * d2[i - 2]) ^ d2[i - 2];
      unsigned long long d3[53];
      d3[1] = 1; d3[2] = 3; d3[3] = 1;
      for (int i = 4; i <= 52; i++) d3[i] = (4 * d3[i - 2]) ^ (8 * d3[i - 3]) ^ d3[i - 3];
      for (int i = 1; i <= 52; i++) {
        SM.m[0][i - 1] = 1ULL << (52 - i);
        SM.m[1][i - 1] = d1[i] << (52 - i);
        SM.m[2][i - 1] = d2[i] << (52 - i);
        SM.m[3][i - 1] = d3[i] << (52 - i);
      }
    }
    unsigned long long *desc_d = nullptr, *rows_d = nullptr;
    double *af_d = nullptr, *afi_d = nullptr;
    int32_t* cc_d = nullptr;
    auto release2 = [&]() {
      for (void* p : {(void*)desc_d, (void*)rows_d, (void*)af_d, (void*)afi_d, (void*)cc_d})
        if (p) cudaFree(p);
    };
    avg_f.resize(rows.size());
    avg_finv.resize(rows.size());
    std::vector<int32_t> cc(work.size());
    bool bad2 = CU(cudaMalloc(&desc_d, desc.size() * 8), "cudaMalloc") || CU(cudaMalloc(&rows_d, rows.size() * 8), "cudaMalloc") ||
                CU(cudaMalloc(&af_d, rows.size() * 8), "cudaMalloc") || CU(cudaMalloc(&afi_d, rows.size() * 8), "cudaMalloc") ||
                CU(cudaMalloc(&cc_d, work.size() * 4), "cudaMalloc") ||
                CU(cudaMemcpyAsync(desc_d, desc.data(), desc.size() * 8, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync") ||
                CU(cudaMemcpyAsync(rows_d, rows.data(), rows.size() * 8, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync") ||
                CU(cudaMemsetAsync(af_d, 0, rows.size() * 8, s), "cudaMemsetAsync") || CU(cudaMemsetAsync(afi_d, 0, rows.size() * 8, s), "cudaMemsetAsync");
    if (!bad2) {
      if (full)
        exact_full_average_kernel<<<grid_for(h, work.size(), 64), 64, 0, s>>>(h->P, m, SM, work.size(), desc_d, rows_d, h->parent_info, h->parent_cap,
                                                                              m.nsamples, af_d, afi_d, cc_d);
      else
        exact_average_kernel<<<grid_for(h, work.size(), 64), 64, 0, s>>>(h->P, m, SM, work.size(), desc_d, rows_d, b, h->parent_info, h->parent_cap,
                                                                         m.nsamples, m.sobol, rng2[0], af_d, afi_d, cc_d);
      h->launches++;
      bad2 = CU(cudaMemcpyAsync(avg_f.data(), af_d, rows.size() * 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync") ||
             CU(cudaMemcpyAsync(avg_finv.data(), afi_d, rows.size() * 8, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync") ||
             CU(cudaMemcpyAsync(cc.data(), cc_d, work.size() * 4, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync") ||
             CU(cudaStreamSynchronize(s), "exact_average_kernel");
    }
    release2();
    if (bad2) return ABL_ERR_CUDA;
    for (size_t q = 0; q < work.size(); q++) work[q].can_cancel = cc[q];
  }
  t_avg = ms_since(t_start);
  // cancel_bin (:408-455) with get_beta (:366-406), particle by particle in bank order
  bool touched = false, touched2 = false;
  for (Work& wk : work) {
    abl_context::ExactBin& bin = *wk.bin;
    const size_t np = bin.particles.size();
    double sum_c = 0., sum_c_wgt = 0., sum_c_wgt2 = 0.;
    if (beta == ABL_BETA_AVERAGE_G && wk.can_cancel) {
      auto C = [](double fa, double fi) { return 1. / (2. * fa * fi - 1.); };
      for (size_t e = 0; e < np; e++) sum_c += C(avg_f[wk.first + e], avg_finv[wk.first + e]);
      for (size_t e = 0; e < np; e++) {
        sum_c_wgt += C(avg_f[wk.first + e], avg_finv[wk.first + e]) * w[bin.particles[e]];
        sum_c_wgt2 += C(avg_f[wk.first + e], avg_finv[wk.first + e]) * w2[bin.particles[e]];
      }
    }
    for (int pass = 0; pass < 2; pass++) {
      const bool first = pass == 0;
      if (!(first ? wk.w1 : wk.w2)) continue;
      std::vector<double>& wv = first ? w : w2;
      for (size_t e = 0; e < np; e++) {
        const uint64_t i = bin.particles[e];
        const double wgt = wv[i];
        if (wgt == 0. && !full) break;  // (BasicExactMGCancelator::cancel_bin only, :432-434)
        double B = 0.;
        if (wk.can_cancel) {
          if (beta == ABL_BETA_MINIMUM) {
            B = fmin[i];
          } else if (beta == ABL_BETA_AVERAGE_F) {
            const double fa = avg_f[wk.first + e];
            const double N = static_cast<double>(np);
            const double W = first ? bin.W : bin.W2;
            B = fa * (1. - ((W) / ((N + 1.) * wgt)));
          } else {
            const double scw = first ? sum_c_wgt : sum_c_wgt2;
            const double S = scw / (1. + sum_c);
            const double fa = avg_f[wk.first + e], fi = avg_finv[wk.first + e];
            B = fa * (1. / (2. * fa * fi - 1.)) * (1. - (S / wgt));
          }
        }
        const double fi = f[i];
        const double P_p = (fi - B) / fi, P_u = B / fi;
        if (std::isinf(P_u) || std::isinf(P_p) || std::isnan(P_u) || std::isnan(P_p)) break;
        (first ? bin.uniform_wgt : bin.uniform_wgt2) += wv[i] * P_u;
        wv[i] *= P_p;
        (first ? touched : touched2) = true;
      }
    }
  }
  if (!full && m.beta != ABL_BETA_MINIMUM) {  // rng.advance(seed_advance) (:552-554; ExactMGCancelator does not touch the engine here)
    uint64_t acc_mult = 1, acc_plus = 0, cur_mult = 6364136223846793005ULL, cur_plus = rng2[1], delta = seed_advance;
    while (delta > 0) {
      if (delta & 1) {
        acc_mult *= cur_mult;
        acc_plus = acc_plus * cur_mult + cur_plus;
      }
      cur_plus = (cur_mult + 1) * cur_plus;
      cur_mult *= cur_mult;
      delta >>= 1;
    }
    rng2[0] = acc_mult * rng2[0] + acc_plus;
  }
  t_cancel = ms_since(t_start);
  // (3) get_new_particles (:557-617): which bins emit how many uniform particles, in the map's order
  std::vector<double> list;
  uint64_t n_new = 0;
  for (auto& kb : bins)
    for (auto& mb : kb.second) {
      abl_context::ExactBin& bin = mb.second;
      const uint32_t N = static_cast<uint32_t>(std::ceil(std::max(std::abs(bin.uniform_wgt), std::abs(bin.uniform_wgt2))));
      if (N > 0) {
        list.insert(list.end(), {(double)kb.first, (double)mb.first, (double)N, bin.uniform_wgt / N, bin.uniform_wgt2 / N});
        n_new += N;
      }
      bin.uniform_wgt = 0.;
      bin.uniform_wgt2 = 0.;
    }
  bins.clear();
  if (n + n_new > capacity) {
    char buf[160];
    snprintf(buf, sizeof buf, "exact cancelator: %llu particles + %llu uniform particles exceed the bank capacity %llu",
             (unsigned long long)n, (unsigned long long)n_new, (unsigned long long)capacity);
    return fail(h, ABL_ERR_BANK_OVERFLOW, buf);
  }
  if (touched) ABL_CUDA(h, cudaMemcpyAsync(b.wgt, w.data(), n * 8, cudaMemcpyHostToDevice, s));
  if (touched2 && b.wgt2) ABL_CUDA(h, cudaMemcpyAsync(b.wgt2, w2.data(), n * 8, cudaMemcpyHostToDevice, s));
  if (n_new) {
    int rc = ensure_parent_info(h, n + n_new, n);
    if (rc) return rc;
    double* list_d = nullptr;
    uint64_t* st_d = nullptr;
    ABL_CUDA(h, cudaMalloc(&list_d, list.size() * 8));
    ABL_CUDA(h, cudaMalloc(&st_d, 3 * 8));
    unsigned long long st[3] = {rng2[0], 0, 0};
    ABL_CUDA(h, cudaMemcpyAsync(list_d, list.data(), list.size() * 8, cudaMemcpyHostToDevice, s));
    ABL_CUDA(h, cudaMemcpyAsync(st_d, st, 3 * 8, cudaMemcpyHostToDevice, s));
    // (one warp: every lane draws the particle it would own if all before it took one position try -- bank_ops.cuh)
    std::vector<unsigned long long> cum(list.size() / 5 + 1, 0);
    for (size_t e = 0; e < list.size() / 5; e++) cum[e + 1] = cum[e] + (unsigned long long)list[5 * e + 2];
    unsigned long long* cum_d = nullptr;
    ABL_CUDA(h, cudaMalloc(&cum_d, cum.size() * 8));
    ABL_CUDA(h, cudaMemcpyAsync(cum_d, cum.data(), cum.size() * 8, cudaMemcpyHostToDevice, s));
    if (full)
      exact_uniform_warp_kernel<true><<<1, 32, 0, s>>>(h->P, m, list_d, cum_d, list.size() / 5, n_new, st_d, b, n, capacity, h->parent_info,
                                                       h->parent_cap, reinterpret_cast<unsigned long long*>(st_d + 1));
    else
      exact_uniform_warp_kernel<false><<<1, 32, 0, s>>>(h->P, m, list_d, cum_d, list.size() / 5, n_new, st_d, b, n, capacity, h->parent_info,
                                                        h->parent_cap, reinterpret_cast<unsigned long long*>(st_d + 1));
    h->launches++;
    ABL_CUDA(h, cudaMemcpyAsync(st, st_d, 3 * 8, cudaMemcpyDeviceToHost, s));
    ABL_CUDA(h, cudaStreamSynchronize(s));
    cudaFree(list_d);
    cudaFree(st_d);
    cudaFree(cum_d);
    if (st[2]) return fail(h, ABL_ERR_LOST, "Couldn't sample position for uniform particle.");
    rng2[0] = st[0];
    bank_dev->n = st[1];
    h->parent_n = st[1];
  } else {
    ABL_CUDA(h, cudaStreamSynchronize(s));
  }
  if (timing)
    fprintf(stderr, "abl_cancel_exact_device: n %llu, +%llu uniform; columns %.1f ms, bins %.1f, averages %.1f, cancel %.1f, total %.1f ms\n",
            (unsigned long long)n, (unsigned long long)n_new, t_cols, t_bins - t_cols, t_avg - t_bins, t_cancel - t_avg, ms_since(t_start));
  return ABL_OK;
}

// ---- device memory helpers ---------------------------------------------------------------------------------------------------
int abl_bank_alloc_device(abl_handle h, uint64_t capacity, abl_bank* out_dev) {
  if (!h || !out_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  BankView b{};
  int rc = alloc_bank(h, b, capacity ? capacity : 1);
  if (rc) {
    free_bank(b);
    return rc;
  }
  out_dev->n = capacity;
  out_dev->x = b.x; out_dev->y = b.y; out_dev->z = b.z; out_dev->ux = b.ux; out_dev->uy = b.uy; out_dev->uz = b.uz;
  out_dev->E = b.E; out_dev->wgt = b.wgt; out_dev->wgt2 = b.wgt2;
  out_dev->id_a = b.id_a; out_dev->id_b = b.id_b; out_dev->id_c = b.id_c;
  return ABL_OK;
}

int abl_bank_free_device(abl_handle h, abl_bank* bank_dev) {
  if (!h || !bank_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  ABL_CUDA(h, cudaDeviceSynchronize());
  BankView b = view_of(bank_dev);
  free_bank(b);
  *bank_dev = abl_bank{};
  return ABL_OK;
}

static int copy_bank(abl_handle h, const abl_bank* src, abl_bank* dst, uint64_t n, cudaMemcpyKind kind) {
  const double* s9[9] = {src->x, src->y, src->z, src->ux, src->uy, src->uz, src->E, src->wgt, src->wgt2};
  double* d9[9] = {dst->x, dst->y, dst->z, dst->ux, dst->uy, dst->uz, dst->E, dst->wgt, dst->wgt2};
  const uint64_t* s3[3] = {src->id_a, src->id_b, src->id_c};
  uint64_t* d3[3] = {dst->id_a, dst->id_b, dst->id_c};
  if (n == 0) return ABL_OK;
  for (int k = 0; k < 9; k++)
    if (s9[k] && d9[k]) ABL_CUDA(h, cudaMemcpy(d9[k], s9[k], n * sizeof(double), kind));
  for (int k = 0; k < 3; k++)
    if (s3[k] && d3[k]) ABL_CUDA(h, cudaMemcpy(d3[k], s3[k], n * sizeof(uint64_t), kind));
  return ABL_OK;
}

int abl_bank_upload(abl_handle h, const abl_bank* host, abl_bank* bank_dev) {
  if (!h || !host || !bank_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  ABL_CUDA(h, cudaDeviceSynchronize());
  int rc = copy_bank(h, host, bank_dev, host->n, cudaMemcpyHostToDevice);
  if (rc) return rc;
  if (!host->wgt2 && bank_dev->wgt2 && host->n) ABL_CUDA(h, cudaMemset(bank_dev->wgt2, 0, host->n * sizeof(double)));
  bank_dev->n = host->n;
  return ABL_OK;
}

int abl_bank_download(abl_handle h, const abl_bank* bank_dev, uint64_t n, abl_bank* host) {
  if (!h || !host || !bank_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  ABL_CUDA(h, cudaDeviceSynchronize());
  return copy_bank(h, bank_dev, host, n, cudaMemcpyDeviceToHost);
}

int abl_device_alloc(abl_handle h, uint64_t bytes, void** out_dev) {
  if (!h || !out_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  ABL_CUDA(h, cudaMalloc(out_dev, bytes ? bytes : 8));
  ABL_CUDA(h, cudaMemset(*out_dev, 0, bytes ? bytes : 8));
  return ABL_OK;
}

int abl_device_free(abl_handle h, void* dev) {
  if (!h) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  ABL_CUDA(h, cudaDeviceSynchronize());
  if (dev) ABL_CUDA(h, cudaFree(dev));
  return ABL_OK;
}

int abl_device_zero(abl_handle h, void* dev, uint64_t bytes, void* stream) {
  if (!h || !dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  ABL_CUDA(h, cudaMemsetAsync(dev, 0, bytes, use_stream(h, (cudaStream_t)stream)));
  return ABL_OK;
}

int abl_device_read(abl_handle h, void* dst_host, const void* src_dev, uint64_t bytes, void* stream) {
  if (!h || !dst_host || !src_dev) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = use_stream(h, (cudaStream_t)stream);
  ABL_CUDA(h, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, s));
  ABL_CUDA(h, cudaStreamSynchronize(s));
  return ABL_OK;
}

// ---- probes ------------------------------------------------------------------------------------------------------------------
static int ensure_probe(abl_handle h, uint64_t bytes) {
  use_stream(h, h->stream);
  if (bytes > h->probe_cap) {
    if (h->probe_buf) cudaFree(h->probe_buf);
    h->probe_buf = nullptr;
    h->probe_cap = 0;
    ABL_CUDA(h, cudaMalloc(&h->probe_buf, bytes));
    h->probe_cap = bytes;
  }
  return ABL_OK;
}

int abl_find_cells(abl_handle h, uint64_t n, const double* r3, const double* u3, int32_t* cell, int32_t* material) {
  if (!h || !r3 || !u3 || !cell || !material) return ABL_ERR_INVALID;
  if (n == 0) return ABL_OK;
  ABL_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_probe(h, n * (6 * sizeof(double) + 2 * sizeof(int32_t)));
  if (rc) return rc;
  double* dr = h->probe_buf;
  double* du = dr + 3 * n;
  int32_t* dc = reinterpret_cast<int32_t*>(du + 3 * n);
  int32_t* dm = dc + n;
  ABL_CUDA(h, cudaMemcpy(dr, r3, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
  ABL_CUDA(h, cudaMemcpy(du, u3, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
  find_cells_kernel<<<grid_for(h, n, 128), 128, 0, h->stream>>>(h->P, n, dr, du, dc, dm);
  h->launches++;
  ABL_CUDA(h, cudaGetLastError());
  ABL_CUDA(h, cudaStreamSynchronize(h->stream));
  ABL_CUDA(h, cudaMemcpy(cell, dc, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  ABL_CUDA(h, cudaMemcpy(material, dm, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return ABL_OK;
}

int abl_surface_probe(abl_handle h, int surface_index, uint64_t n, const double* r3, const double* u3, const int32_t* on_surf,
                      int32_t* sign, double* distance, double* norm3) {
  if (!h || !r3 || !u3 || !on_surf || !sign || !distance || !norm3) return ABL_ERR_INVALID;
  if (surface_index < 0 || surface_index >= h->P.nsurfaces) return ABL_ERR_INVALID;
  if (n == 0) return ABL_OK;
  ABL_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_probe(h, n * (10 * sizeof(double) + 2 * sizeof(int32_t)));
  if (rc) return rc;
  double* dr = h->probe_buf;
  double* du = dr + 3 * n;
  double* dd = du + 3 * n;
  double* dn = dd + n;
  int32_t* don = reinterpret_cast<int32_t*>(dn + 3 * n);
  int32_t* ds = don + n;
  ABL_CUDA(h, cudaMemcpy(dr, r3, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
  ABL_CUDA(h, cudaMemcpy(du, u3, 3 * n * sizeof(double), cudaMemcpyHostToDevice));
  ABL_CUDA(h, cudaMemcpy(don, on_surf, n * sizeof(int32_t), cudaMemcpyHostToDevice));
  surface_probe_kernel<<<grid_for(h, n, 128), 128, 0, h->stream>>>(h->P, surface_index, n, dr, du, don, ds, dd, dn);
  h->launches++;
  ABL_CUDA(h, cudaGetLastError());
  ABL_CUDA(h, cudaStreamSynchronize(h->stream));
  ABL_CUDA(h, cudaMemcpy(sign, ds, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  ABL_CUDA(h, cudaMemcpy(distance, dd, n * sizeof(double), cudaMemcpyDeviceToHost));
  ABL_CUDA(h, cudaMemcpy(norm3, dn, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
  return ABL_OK;
}

uint64_t abl_fission_capacity_hint(abl_handle h, uint64_t n_particles, double sum_abs_weight, double k_col) {
  if (!h) return 0;
  const double W = std::max(sum_abs_weight, static_cast<double>(n_particles));
  const double k = (k_col > 1e-3) ? k_col : 1e-3;
  const double mean = W * std::max(h->k_site_max, 0.05) / k;
  const double cap = 1.25 * mean + 8. * std::sqrt(mean) + 8192.;
  return cap < 1.8e19 ? static_cast<uint64_t>(cap) : ~0ULL;
}

int abl_set_sampling_xs(abl_handle h, const double* sampling_xs, int ngroups) {
  if (!h || !sampling_xs || ngroups != h->P.G || !h->P.smp) return ABL_ERR_INVALID;
  ABL_CUDA(h, cudaSetDevice(h->device));
  use_stream(h, h->stream);
  ABL_CUDA(h, cudaStreamSynchronize(h->stream));
  const int G = h->P.G, M = h->P.M;
  std::vector<double> rf((size_t)M * G);
  for (int m = 0; m < M; m++)
    for (int g = 0; g < G; g++) rf[(size_t)m * G + g] = h->host_Et[(size_t)m * G + g] / sampling_xs[g];  // delta_tracker.cpp:182
  ABL_CUDA(h, cudaMemcpy(const_cast<double*>(h->P.smp), sampling_xs, (size_t)G * sizeof(double), cudaMemcpyHostToDevice));
  ABL_CUDA(h, cudaMemcpy(const_cast<double*>(h->P.real_frac), rf.data(), rf.size() * sizeof(double), cudaMemcpyHostToDevice));
  return ABL_OK;
}

int abl_rng_probe(abl_handle h, uint64_t history_id, int n, uint32_t* out_u32, double* out_rand) {
  if (!h || n < 0 || !out_u32 || !out_rand) return ABL_ERR_INVALID;
  if (n == 0) return ABL_OK;
  ABL_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_probe(h, (uint64_t)n * 16);
  if (rc) return rc;
  double* dr = h->probe_buf;
  uint32_t* du = reinterpret_cast<uint32_t*>(dr + n);
  rng_probe_kernel<<<1, 32, 0, h->stream>>>(h->P, history_id, n, du, dr);
  h->launches++;
  ABL_CUDA(h, cudaGetLastError());
  ABL_CUDA(h, cudaStreamSynchronize(h->stream));
  ABL_CUDA(h, cudaMemcpy(out_u32, du, (size_t)n * 4, cudaMemcpyDeviceToHost));
  ABL_CUDA(h, cudaMemcpy(out_rand, dr, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return ABL_OK;
}

int abl_math_probe(abl_handle h, int n, const double* x, double* lg, double* sn, double* cs) {
  if (!h || n < 0 || !x || !lg || !sn || !cs) return ABL_ERR_INVALID;
  if (n == 0) return ABL_OK;
  ABL_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_probe(h, (uint64_t)n * 32);
  if (rc) return rc;
  double* dx = h->probe_buf;
  ABL_CUDA(h, cudaMemcpy(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice));
  math_probe_kernel<<<grid_for(h, (uint64_t)n, 128), 128, 0, h->stream>>>(n, dx, dx + n, dx + 2 * n, dx + 3 * n);
  h->launches++;
  ABL_CUDA(h, cudaGetLastError());
  ABL_CUDA(h, cudaStreamSynchronize(h->stream));
  ABL_CUDA(h, cudaMemcpy(lg, dx + n, (size_t)n * 8, cudaMemcpyDeviceToHost));
  ABL_CUDA(h, cudaMemcpy(sn, dx + 2 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
  ABL_CUDA(h, cudaMemcpy(cs, dx + 3 * n, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return ABL_OK;
}

}  // extern "C"
