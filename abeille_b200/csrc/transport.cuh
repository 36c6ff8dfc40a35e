/* transport.cuh -- particle state, collision physics and the PER-LANE history kernel (transport_kernel).
 *
 * The per-lane kernel is the direct form of Transporter::transport: every thread runs the reference's control flow for
 * its history, one whole flight per loop iteration.  It serves the noise modes (noise.cuh); k-eigenvalue runs use the
 * staged kernel (history.cuh), which shares everything in this file but the loop.
 *
 * One persistent grid, one lock-step CTA per SM.  Every thread owns ONE neutron history at a time and
 * keeps its whole state -- position, direction, group, weight, pcg32 state, geometry cursor -- in
 * registers / L1-resident local memory; when the history dies the thread takes the next bank index from
 * a global ticket counter (one aggregated atomic per warp), so lanes never idle while the bank has work.
 * HBM sees: 96 B read per history, 80 B written per fission site, one fp64 RED per tally score.
 *
 * Reference behaviour reproduced (paths in the reference tree):
 *   DeltaTracker::transport     src/delta_tracker.cpp:72-263
 *   SurfaceTracker::transport   src/surface_tracker.cpp:40-219
 *   CarterTracker::transport    src/carter_tracker.cpp:92-294
 *   Transporter::collision, branching_collision, make_fission_neutrons, russian_roulette, do_scatter
 *                               src/transporter.cpp:35-93,269-487
 *   MGNuclide samplers          src/mg_nuclide.cpp:442-543, include/materials/mg_angle_distribution.hpp:45-101
 *   MaterialHelper              include/materials/material_helper.hpp:47-224 (one nuclide, N = 1 in MG)
 *   Tracker::do_reflection      include/simulation/tracker.hpp:314-360
 * Random numbers are consumed in exactly the reference's order, including the
 * draws whose value is discarded (nuclide choice with one nuclide, wgt2 roulette in k-eigenvalue mode).
 *
 * The fission bank: sites are appended to a scratch array in arrival order (warp-aggregated ticket),
 * each tagged (parent bank index, daughter number); place_sites_kernel then writes them at
 * offset[parent] + daughter, offset = exclusive scan of the per-history site counts -- the reference's
 * order (bank order, then creation order; delta_tracker.cpp:246-253) without a sort.
 */
#pragma once
#include <cooperative_groups.h>

#include "tally.cuh"

namespace abl {
namespace cg = cooperative_groups;

struct alignas(16) Site {  // scratch fission site, 80 B
  double x, y, z, ux, uy, uz, E, w, w2;
  uint32_t parent, daughter;
};

struct BankView {  // SoA device arrays (abl_bank with device pointers)
  uint64_t n;
  double *x, *y, *z, *ux, *uy, *uz, *E, *wgt, *wgt2;
  uint64_t *id_a, *id_b, *id_c;
};

#define ABL_SEC_CAP 24     // LIFO entries a single history may hold at once (an entry is one secondary or `count` identical copies)
#define ABL_SEC_FIELDS 10  // r, u, E, wgt, wgt2, count
#define ABL_PARENT_FIELDS 10  // exact cancelators' side record per fission site (RunArgs::site_parent)

struct RunArgs {
  BankView bank;
  unsigned long long* ticket;      // next bank index
  Site* sites;                     // scratch fission sites
  unsigned long long* n_sites;     // sites appended (may exceed capacity: overflow is reported)
  uint64_t site_capacity;
  uint32_t* nfis;                  // per history: sites produced
  // trace (nullable)
  uint32_t *tr_flights, *tr_real, *tr_virtual;
  uint64_t *tr_hash, *tr_rng;
  double* scores;                  // [6] k_col, k_abs, k_trk, k_tot, leak, mig
  unsigned long long* counters;    // [8]
  int* error;                      // [0] code (min = most severe first seen), [1..2] history id lo/hi
  double* secondaries;             // [ABL_SEC_CAP][ABL_SEC_FIELDS][nthreads] or null
  double* site_parent;             // [site_capacity][ABL_PARENT_FIELDS] or null: what the exact cancelators read of every scratch site
                                   // (BankedParticle::parents_previous_position, Esmp_parent, parents_previous_direction,
                                   // parents_previous_previous_energy, parents_previous_energy, parents_previous_was_virtual)
  double k_col, keff;
  int converged;
  // rows of the bank that have arrived in HBM (host-buffer entry point: the copy overlaps the kernel), or null
  const unsigned long long* avail;
  // noise mode (noise.cuh)
  int sample_noise;                // this power-iteration generation samples the noise source
  Site* nsites;                    // scratch noise particles
  unsigned long long* n_nsites;
  uint64_t nsite_capacity;
  uint32_t* nnoise;                // per history: noise particles produced
  uint32_t *site_did, *nsite_did;  // daughter ids of the scratch sites (the Site itself carries the rank)
  // staged history kernel (history.cuh): histories per CTA and the geometry's nesting depth (frames, pads), which size
  // the per-history columns in shared memory
  int hk_slots, hk_nf, hk_np;
  // tables staged in shared memory: bytes (0 = none), their global-memory source, and the generic address the host assumed
  // for byte 0 of the dynamic shared memory when it redirected the table pointers (verified by the kernel)
  int hk_tables;
  const unsigned char* arena;
  unsigned long long smem_generic_base;
  unsigned long long* eq_stats;   // event kernel: dispatcher statistics (tasks and lanes per event; events.cuh ST_*), or null
  unsigned long long timeout_ns;  // the staged kernel's watchdog: a CTA that runs longer winds down with ABL_ERR_TIMEOUT
};

struct Hist {
  V3 r, u, rb;  // position, direction, birth position
  double E, w, w2;
  uint64_t rng, hash;
  uint32_t idx, daughter;
  uint32_t n_flights, n_real, n_virtual, n_fis, n_noise;
  int g, mat;  // energy group, material of the MaterialHelper
  int nsec;
  bool alive;
  bool emid;  // E is exactly the mid-point of group g (true after any scatter / for fission sites): tally bins by table
  // Particle::previous_position / reflected / Esmp_ (particle.hpp:104-133,229-237), kept when the problem has an exact cancelator
  bool refl, virt;   // virt: Particle::previous_collision_virtual (delta_tracker.cpp:188-195)
  V3 rprev, uprev;   // uprev / Eprev: Particle::previous_direction / previous_energy (set_direction / set_energy keep what they replace)
  double esmp, Eprev;
};

// the side record of a fission site (the order of abl_parent_info_download / abl_parent_state_download)
__device__ __forceinline__ void write_parent_record(double* site_parent, unsigned long long slot, const Hist& h) {
  double2* pp = reinterpret_cast<double2*>(site_parent + (size_t)ABL_PARENT_FIELDS * slot);
  pp[0] = make_double2(h.rprev.x, h.rprev.y);
  pp[1] = make_double2(h.rprev.z, h.esmp);
  pp[2] = make_double2(h.uprev.x, h.uprev.y);
  pp[3] = make_double2(h.uprev.z, h.Eprev);
  pp[4] = make_double2(h.E, h.virt ? 1. : 0.);
}

struct Acc {  // per-thread accumulators, reduced once at kernel exit
  double k_col, k_abs, k_trk, leak, mig;
  uint32_t flights, real, virt, tl_bins, sites, boundary, lost, coll_scores;
};

__device__ __forceinline__ void note(Hist& h, uint64_t v) { h.hash = (h.hash ^ v) * 1099511628211ULL; }

__device__ __forceinline__ int group_of(const DevProblem& P, double E) {  // mg_nuclide.cpp:382-392
  int i = 0;
  for (i = 0; i < P.G; i++)
    if (ldt(&P.ebounds[i]) <= E && E < ldt(&P.ebounds[i + 1])) break;
  return i;
}
__device__ __forceinline__ double group_mid(const DevProblem& P, int g) { return ldt(&P.gmid[g]); }  // 0.5*(b[g]+b[g+1])

__device__ __forceinline__ void raise_error(const RunArgs& A, int code, uint64_t history_id) {
  if (atomicCAS(&A.error[0], 0, code) == 0) {
    A.error[1] = (int)(history_id & 0xffffffffu);
    A.error[2] = (int)(history_id >> 32);
  }
}

// inverse CDF on a linearly interpolable pdf (mg_angle_distribution.hpp:45-60)
static __device__ __noinline__ double sample_mu_table(const double* __restrict__ acdf, const double* __restrict__ amu,
                                               const double* __restrict__ apdf, int off, int n, double xi) {
  const double* cdf = acdf + off;
  int lo = 0, len = n;
  while (len > 0) {  // std::lower_bound
    const int half = len >> 1;
    if (ldt(&cdf[lo + half]) < xi) {
      lo = lo + half + 1;
      len = len - half - 1;
    } else {
      len = half;
    }
  }
  int l = lo;
  const double* mu = amu + off;
  if (xi == ldt(&cdf[l])) return ldt(&mu[l]);
  l--;
  const double* pdf = apdf + off;
  const double p0 = ldt(&pdf[l]), p1 = ldt(&pdf[l + 1]);
  if (p0 == p1) return ldt(&mu[l]) + ((xi - ldt(&cdf[l])) / p0);
  const double m = (p1 - p0) / (ldt(&mu[l + 1]) - ldt(&mu[l]));
  return ldt(&mu[l]) + (1. / m) * (sqrt(p0 * p0 + 2. * m * (xi - ldt(&cdf[l]))) - p0);
}

// MGAngleDistribution::sample_mu (mg_angle_distribution.hpp:45-60,92-101)
template <class M = InlineMath>
__device__ __forceinline__ double sample_mu(const DevProblem& P, const abl_angle_table* at, uint64_t& rng) {
  const double xi = M::rand(rng);
  const int off = ldt(&at->offset), n = ldt(&at->n);
  // n < 0 marks the default isotropic table {mu:[-1,1], pdf:[.5,.5], cdf:[0,1]} (mg_angle_distribution.cpp:32-33):
  // the general formula below reduces to this expression, evaluated identically
  if (n < 0) return -1. + ((xi - 0.) / 0.5);
  return sample_mu_table(P.acdf, P.amu, P.apdf, off, n, xi);
}

// warp-aggregated append of one fission site
__device__ __forceinline__ void append_site(const RunArgs& A, const Site& s) {
  cg::coalesced_group grp = cg::coalesced_threads();
  unsigned long long base = 0;
  if (grp.thread_rank() == 0) base = atomicAdd(A.n_sites, (unsigned long long)grp.size());
  base = grp.shfl(base, 0);
  const unsigned long long slot = base + grp.thread_rank();
  if (slot < A.site_capacity) {
    // 80 B record written as five 16 B stores
    const double2* src = reinterpret_cast<const double2*>(&s);
    double2* dst = reinterpret_cast<double2*>(A.sites + slot);
#pragma unroll
    for (int q = 0; q < 5; q++) dst[q] = src[q];
  }
}

// n_new x MGNuclide::sample_fission (mg_nuclide.cpp:504-543) + Particle::add_fission_particle, as a real function
// call: about 2 % of the collisions bank sites, so this stays out of the hot loop's instruction footprint.
struct FissionTables {
  const double* chi_cp;
  const double* dg_cp;
  const int32_t* dg_off;
  const double* gmid;
  int G;
};
template <class M, bool PARENT = false>  // PARENT: also write the exact cancelators' side record (the per-lane kernel's calls)
static __device__ __noinline__ void bank_fission_sites(const FissionTables T, Site* sites, unsigned long long* n_sites, uint64_t capacity,
                                                uint64_t& rng, const V3 r, const V3 u, double w, uint32_t parent, uint32_t daughter0,
                                                int n_new, int mat, int mg, double P_delayed, double* site_parent = nullptr,
                                                const Hist* parent_state = nullptr) {
  const int dg0 = ldt(&T.dg_off[mat]), ndg = ldt(&T.dg_off[mat + 1]) - dg0;
  for (int i = 0; i < n_new; i++) {
    int ei = 0;
    if (T.G >= 2) ei = rng_discrete<M>(rng, T.chi_cp + (size_t)mg * T.G, T.G);
    const double E_out = ldt(&T.gmid[ei]);
    const double mu = 2. * M::rand(rng) - 1.;
    const double phi = 2. * ABL_PI * M::rand(rng);
    const V3 dir = rotate_dir<M>(u, mu, phi);
    if (M::rand(rng) < P_delayed) {
      if (ndg >= 2) (void)rng_discrete<M>(rng, T.dg_cp + dg0, ndg);  // delayed family: only matters in noise mode
    }
    Site s;
    s.x = r.x; s.y = r.y; s.z = r.z;
    s.ux = dir.x; s.uy = dir.y; s.uz = dir.z;
    s.E = E_out;
    s.w = w > 0. ? 1. : -1.;
    s.w2 = 0.;
    s.parent = parent;
    s.daughter = daughter0 + (uint32_t)i;
    cg::coalesced_group grp = cg::coalesced_threads();
    unsigned long long base = 0;
    if (grp.thread_rank() == 0) base = atomicAdd(n_sites, (unsigned long long)grp.size());
    base = grp.shfl(base, 0);
    const unsigned long long slot = base + grp.thread_rank();
    if (slot < capacity) {  // 80 B record written as five 16 B stores
      const double2* src = reinterpret_cast<const double2*>(&s);
      double2* dst = reinterpret_cast<double2*>(sites + slot);
#pragma unroll
      for (int q = 0; q < 5; q++) dst[q] = src[q];
      if (PARENT && site_parent) write_parent_record(site_parent, slot, *parent_state);
    }
  }
}

__device__ inline bool push_secondary(const RunArgs& A, Hist& h, const V3& u, double E, double w, double w2, uint32_t tid,
                                      uint32_t nthreads, int count = 1);  // (defined with the secondaries' LIFO below)

// Fixed-source problems: the n_new fission neutrons of a collision become secondaries of the history (Particle::make_secondary,
// transporter.cpp:460-463) instead of bank entries.  The same draws per neutron as bank_fission_sites.
template <class M>
static __device__ __noinline__ bool fission_to_secondaries(const FissionTables T, const RunArgs& A, Hist& h, int n_new, int mat, int mg,
                                                           double P_delayed, uint32_t tid, uint32_t nthreads) {
  const int dg0 = ldt(&T.dg_off[mat]), ndg = ldt(&T.dg_off[mat + 1]) - dg0;
  for (int i = 0; i < n_new; i++) {
    int ei = 0;
    if (T.G >= 2) ei = rng_discrete<M>(h.rng, T.chi_cp + (size_t)mg * T.G, T.G);
    const double E_out = ldt(&T.gmid[ei]);
    const double mu = 2. * M::rand(h.rng) - 1.;
    const double phi = 2. * ABL_PI * M::rand(h.rng);
    const V3 dir = rotate_dir<M>(h.u, mu, phi);
    if (M::rand(h.rng) < P_delayed) {
      if (ndg >= 2) (void)rng_discrete<M>(h.rng, T.dg_cp + dg0, ndg);
    }
    if (!push_secondary(A, h, dir, E_out, h.w > 0. ? 1. : -1., 0., tid, nthreads)) return false;
  }
  return true;
}

template <bool NOISE, class M = InlineMath>
__device__ __forceinline__ void russian_roulette(const DevProblem& P, Hist& h) {  // transporter.cpp:35-58
  if (fabs(h.w) < P.wgt_cutoff) {
    const double P_kill = 1.0 - M::div(fabs(h.w), P.wgt_survival);
    if (M::rand(h.rng) < P_kill) h.w = 0.;
    else h.w = copysign(P.wgt_survival, h.w);
  }
  if (fabs(h.w2) < P.wgt_cutoff) {  // w2 == 0 outside noise mode: the draw is still consumed
    const double P_kill = 1.0 - ddiv_pos<M>(fabs(h.w2), P.wgt_survival);
    if (M::rand(h.rng) < P_kill) h.w2 = 0.;
    else h.w2 = copysign(P.wgt_survival, h.w2);
  }
  if (h.w == 0. && h.w2 == 0.) h.alive = false;
}

// Transporter::branchless_collision_mat / _iso (transporter.cpp:104-267): the collision is EITHER a scatter that carries the
// multiplicity m = (nu Sigma_f + Sigma_s) / Sigma_t in its weight, OR a fission that banks one site of weight w*m and ends the
// particle.  One nuclide per material (atoms_bcm = 1), so sample_nuclide / sample_branchless_nuclide reduce to their single draw.
// A function call: only branchless-k-eigenvalue problems come here.
// (the tables it reads travel by value: a reference to the kernel's DevProblem parameter would make every thread copy the
// whole 2.5 KB structure to its stack)
struct BranchlessTables {
  const double *Et, *Ea, *Ef, *nu, *nud, *ps_cp, *chi_cp, *dg_cp, *gmid, *acdf, *amu, *apdf;
  const abl_angle_table* angle;
  const int32_t* dg_off;
  int G, branchless;
  double wgt_cutoff, wgt_survival, wgt_split, min_energy;
};
__device__ __forceinline__ BranchlessTables branchless_tables(const DevProblem& P) {
  return BranchlessTables{P.Et, P.Ea, P.Ef, P.nu, P.nud, P.ps_cp, P.chi_cp, P.dg_cp, P.gmid, P.acdf, P.amu, P.apdf, P.angle, P.dg_off,
                          P.G, P.branchless, P.wgt_cutoff, P.wgt_survival, P.wgt_split, P.min_energy};
}
template <class M>
static __device__ __noinline__ void branchless_collision(const BranchlessTables P, const RunArgs& A, Hist& h, Acc& acc, uint32_t tid,
                                                         uint32_t nthreads) {
  const int mg = h.mat * P.G + h.g;
  auto roulette = [&](Hist& q) {  // russian_roulette (transporter.cpp:35-58)
    if (fabs(q.w) < P.wgt_cutoff) {
      const double P_kill = 1.0 - M::div(fabs(q.w), P.wgt_survival);
      if (M::rand(q.rng) < P_kill) q.w = 0.;
      else q.w = copysign(P.wgt_survival, q.w);
    }
    if (fabs(q.w2) < P.wgt_cutoff) {
      const double P_kill = 1.0 - ddiv_pos<M>(fabs(q.w2), P.wgt_survival);
      if (M::rand(q.rng) < P_kill) q.w2 = 0.;
      else q.w2 = copysign(P.wgt_survival, q.w2);
    }
    if (q.w == 0. && q.w2 == 0.) q.alive = false;
  };
  const double Et = ldt(&P.Et[mg]), Ea = ldt(&P.Ea[mg]), Ef = ldt(&P.Ef[mg]), nu = ldt(&P.nu[mg]);
  const double vEf_i = nu * Ef, Es_i = Et - Ea;
  bool scatter = false;
  double m;
  if (P.branchless & ABL_BRANCHLESS_MATERIAL) {
    // MaterialHelper::Es / vEf / Et: sums over the components starting from 0. (material_helper.hpp:47-63,100-113)
    const double Es = 0. + 1. * fmax(Es_i, 0.);
    const double vEf = 0. + 1. * nu * Ef;
    const double Pscatter = ddiv_pos<M>(Es, vEf + Es);
    m = ddiv_pos<M>(vEf + Es, Et);
    if (M::rand(h.rng) < Pscatter) scatter = true;
    (void)M::rand(h.rng);  // sample_branchless_nuclide: xi = rand * sum (material_helper.hpp:249)
    const double m_i = ddiv_pos<M>(vEf_i + Es_i, Et);
    acc.k_abs += (m / m_i) * h.w * nu * Ef / Et;
    roulette(h);
    if (!h.alive) return;
  } else {
    (void)M::rand(h.rng);  // sample_nuclide (material_helper.hpp:189)
    acc.k_abs += h.w * nu * Ef / Et;
    const double Pscatter = Es_i / (vEf_i + Es_i);
    m = (vEf_i + Es_i) / Et;
    roulette(h);
    if (!h.alive) return;
    if (M::rand(h.rng) < Pscatter) scatter = true;
  }
  if (scatter) {
    int ei = 0;
    if (P.G >= 2) ei = rng_discrete<M>(h.rng, P.ps_cp + (size_t)mg * P.G, P.G);
    const double E_out = ldt(&P.gmid[ei]);
    double mu;
    {  // MGAngleDistribution::sample_mu (as sample_mu<M> above)
      const abl_angle_table* at = P.angle + (size_t)mg * P.G + ei;
      const double xi = M::rand(h.rng);
      const int off = ldt(&at->offset), n = ldt(&at->n);
      mu = n < 0 ? -1. + ((xi - 0.) / 0.5) : sample_mu_table(P.acdf, P.amu, P.apdf, off, n, xi);
    }
    const double phi = 2. * ABL_PI * M::rand(h.rng);
    h.uprev = h.u;
    h.Eprev = h.E;
    h.u = rotate_dir<M>(h.u, mu, phi);
    h.E = E_out;
    h.g = ei;
    h.emid = true;
    h.w = h.w * m * 1.;
    h.w2 = h.w2 * m * 1.;
    if (h.E < P.min_energy) h.alive = false;
    if ((P.branchless & ABL_BRANCHLESS_SPLITTING) && h.alive && fabs(h.w) >= P.wgt_split) {
      // the material flavour rounds up, the isotope flavour down (transporter.cpp:239-241 against :151-155)
      const int n_new = (int)((P.branchless & ABL_BRANCHLESS_MATERIAL) ? ceil(fabs(h.w)) : floor(fabs(h.w)));
      if (n_new > 1) {  // Particle::split (particle.hpp:165-173)
        h.w = h.w / (double)n_new;
        h.w2 = h.w2 / (double)n_new;
        if (!push_secondary(A, h, h.u, h.E, h.w, h.w2, tid, nthreads, n_new - 1)) raise_error(A, ABL_ERR_BANK_OVERFLOW, A.bank.id_a[h.idx]);
      }
    }
  } else {
    // MGNuclide::sample_fission (mg_nuclide.cpp:504-543), one neutron
    const int dg0 = ldt(&P.dg_off[h.mat]), ndg = ldt(&P.dg_off[h.mat + 1]) - dg0;
    int ei = 0;
    if (P.G >= 2) ei = rng_discrete<M>(h.rng, P.chi_cp + (size_t)mg * P.G, P.G);
    const double E_out = ldt(&P.gmid[ei]);
    const double mu = 2. * M::rand(h.rng) - 1.;
    const double phi = 2. * ABL_PI * M::rand(h.rng);
    const V3 dir = rotate_dir<M>(h.u, mu, phi);
    if (M::rand(h.rng) < ldt(&P.nud[mg]) / nu) {
      if (ndg >= 2) (void)rng_discrete<M>(h.rng, P.dg_cp + dg0, ndg);
    }
    Site s;
    s.x = h.r.x; s.y = h.r.y; s.z = h.r.z;
    s.ux = dir.x; s.uy = dir.y; s.uz = dir.z;
    s.E = E_out;
    s.w = h.w * m;
    s.w2 = h.w2 * m;
    s.parent = h.idx;
    s.daughter = h.daughter;
    cg::coalesced_group grp = cg::coalesced_threads();
    unsigned long long base = 0;
    if (grp.thread_rank() == 0) base = atomicAdd(A.n_sites, (unsigned long long)grp.size());
    base = grp.shfl(base, 0);
    const unsigned long long slot = base + grp.thread_rank();
    if (slot < A.site_capacity) {
      const double2* src = reinterpret_cast<const double2*>(&s);
      double2* dst = reinterpret_cast<double2*>(A.sites + slot);
#pragma unroll
      for (int q = 0; q < 5; q++) dst[q] = src[q];
      if (A.site_parent) write_parent_record(A.site_parent, slot, h);
    }
    h.daughter += 1u;
    h.n_fis += 1u;
    acc.sites += 1u;
    h.alive = false;
  }
}

// Transporter::collision + branching_collision (transporter.cpp:60-93,269-312), k-eigenvalue branch
template <bool NOISE, class M = InlineMath>
__device__ __forceinline__ void collision(const DevProblem& P, const RunArgs& A, Hist& h, Acc& acc, uint32_t tid = 0, uint32_t nthreads = 0) {
  const int mg = h.mat * P.G + h.g;
  const double Et = ldt(&P.Et[mg]), Ea = ldt(&P.Ea[mg]), Ef = ldt(&P.Ef[mg]), nu = ldt(&P.nu[mg]);
  acc.real++;
  h.n_real++;
  if (A.converged && P.n_coll_tallies) {
    const MatXS mx{Et, Ea, Ef, ldt(&P.Es[mg])};
    for (int t = 0; t < P.ntallies; t++)
      if (P.tally[t].estimator == ABL_EST_COLLISION) {
        const int l = h.emid ? ldt(&P.tally_gbin[t * P.G + h.g]) : tally_energy_bin(P.tally[t], h.E);
        acc.coll_scores += score_collision<M>(P.tally[t], h.r, l, h.w, h.w2, mx);
      }
  }
  {
    const double k_col_scr = ddiv_pos<M>(h.w * (nu * Ef), Et);
    const V3 dr{h.r.x - h.rb.x, h.r.y - h.rb.y, h.r.z - h.rb.z};
    const double mig_dist = norm3<M>(dr);
    const double mig_area_scr = ddiv_pos<M>(h.w * Ea, Et) * mig_dist * mig_dist;
    acc.k_col += k_col_scr;
    acc.mig += mig_area_scr;
  }
  if (!NOISE && P.mode == ABL_MODE_BRANCHLESS) {  // transporter.cpp:80-88
    branchless_collision<M>(branchless_tables(P), A, h, acc, tid, nthreads);
    note(h, 0x6000000000000000ULL | (h.alive ? (uint64_t)(h.g + 1) : 0ULL));
    return;
  }
  // MaterialHelper::sample_nuclide always draws, even with a single nuclide (material_helper.hpp:189)
  (void)M::rand(h.rng);
  const double k_abs_scr = ddiv_pos<M>(h.w * nu * Ef, Et);
  acc.k_abs += k_abs_scr;
  // make_fission_neutrons (transporter.cpp:358-487)
  const int n_new = (int)floor(ddiv_pos<M>(fabs(k_abs_scr), A.k_col) + M::rand(h.rng));
  if (n_new > 0) {
    const FissionTables ft{P.chi_cp, P.dg_cp, P.dg_off, P.gmid, P.G};
    if (P.mode == ABL_MODE_FIXED_SOURCE) {  // the neutrons continue this history (src/fixed_source.cpp: the returned bank is empty)
      if (!fission_to_secondaries<M>(ft, A, h, n_new, h.mat, mg, ldt(&P.nud[mg]) / nu, tid, nthreads))
        raise_error(A, ABL_ERR_BANK_OVERFLOW, A.bank.id_a[h.idx]);
      h.daughter += (uint32_t)n_new;  // (Particle::make_secondary does not number daughters; kept for the trace only)
      acc.sites += (uint32_t)n_new;
    } else {
      bank_fission_sites<M, true>(ft, A.sites, A.n_sites, A.site_capacity, h.rng, h.r, h.u, h.w, h.idx, h.daughter, n_new, h.mat, mg,
                                  ldt(&P.nud[mg]) / nu, A.site_parent, &h);
      h.daughter += (uint32_t)n_new;
      h.n_fis += (uint32_t)n_new;
      acc.sites += (uint32_t)n_new;
    }
  }
  note(h, 0x5000000000000000ULL | (uint64_t)(uint32_t)n_new);
  // implicit capture (transporter.cpp:295-298)
  const double surv = 1. - ddiv_pos<M>(Ea + 0., Et);
  h.w = h.w * surv;
  h.w2 = h.w2 * surv;
  russian_roulette<NOISE, M>(P, h);
  if (h.alive) {
    // MGNuclide::sample_scatter (mg_nuclide.cpp:442-461); the yield matrix is never applied
    int ei = 0;
    if (P.G >= 2) ei = rng_discrete<M>(h.rng, P.ps_cp + (size_t)mg * P.G, P.G);
    const double E_out = group_mid(P, ei);
    const double mu = sample_mu<M>(P, P.angle + (size_t)mg * P.G + ei, h.rng);
    const double phi = 2. * ABL_PI * M::rand(h.rng);
    h.uprev = h.u;  // Particle::set_direction / set_energy (particle.hpp:108-117)
    h.Eprev = h.E;
    h.u = rotate_dir<M>(h.u, mu, phi);
    h.E = E_out;
    h.g = ei;
    h.emid = true;
    h.w = h.w * 1.;
    h.w2 = h.w2 * 1.;
    if (h.E < P.min_energy) h.alive = false;
  }
  note(h, 0x6000000000000000ULL | (h.alive ? (uint64_t)(h.g + 1) : 0ULL));
}

// Tracker::do_reflection (tracker.hpp:314-360).  The reference sets the surface token and then wipes it
// again through set_r(): after a reflection the token is 0.
__device__ __forceinline__ bool do_reflection(const DevProblem& P, Cursor& c, Hist& h, const Boundary& b) {
  if (b.surface_index < 0) return false;
  const Surf s = load_surface(P, b.surface_index);
  const V3 r_on{h.r.x + b.distance * h.u.x, h.r.y + b.distance * h.u.y, h.r.z + b.distance * h.u.z};
  const V3 n = surf_norm(s, r_on);
  const double f = 2. * dot3(h.u, n);
  h.u = make_direction(h.u.x - n.x * f, h.u.y - n.y * f, h.u.z - n.z * f);
  h.r = r_on;
  c.token = 0;
  cursor_restart_nl(geo_tables(P), c, h.r, h.u);
  return true;
}

__device__ __forceinline__ void leak(Hist& h, Acc& acc, const Boundary& b) {
  h.alive = false;
  acc.leak += h.w;
  const V3 d{h.r.x + b.distance * h.u.x - h.rb.x, h.r.y + b.distance * h.u.y - h.rb.y, h.r.z + b.distance * h.u.z - h.rb.z};
  acc.mig += leak_mig_score(h.w, d);
}

__device__ __forceinline__ void score_flight_all(const DevProblem& P, const RunArgs& A, const Hist& h, double d, Acc& acc) {
  if (!(A.converged && P.n_tl_tallies)) return;
  const int mg = h.mat * P.G + h.g;
  const MatXS mx{ldt(&P.Et[mg]), ldt(&P.Ea[mg]), ldt(&P.Ef[mg]), ldt(&P.Es[mg])};
  for (int t = 0; t < P.ntallies; t++)
    if (P.tally[t].estimator == ABL_EST_TRACK_LENGTH) acc.tl_bins += score_flight(P.tally[t], h.r, h.u, d, h.E, h.w, h.w2, mx);
}

// secondaries (Particle::make_secondary / split / resurect, particle.hpp:150-186): a LIFO per thread
__device__ __forceinline__ double* sec_slot(const RunArgs& A, int e, int f, uint32_t tid, uint32_t nthreads) {
  return A.secondaries + ((size_t)(e * ABL_SEC_FIELDS + f) * nthreads + tid);
}
// `count` identical copies (Particle::split pushes n_new - 1 of them) take one entry: they pop one at a time, and whatever a
// copy pushes while it runs sits above the entry -- the order of the reference's vector of individual copies.
__device__ inline bool push_secondary(const RunArgs& A, Hist& h, const V3& u, double E, double w, double w2, uint32_t tid,
                                      uint32_t nthreads, int count) {
  if (count < 1) return true;
  if (h.nsec >= ABL_SEC_CAP || A.secondaries == nullptr) return false;
  const int e = h.nsec++;
  *sec_slot(A, e, 0, tid, nthreads) = h.r.x;
  *sec_slot(A, e, 1, tid, nthreads) = h.r.y;
  *sec_slot(A, e, 2, tid, nthreads) = h.r.z;
  *sec_slot(A, e, 3, tid, nthreads) = u.x;
  *sec_slot(A, e, 4, tid, nthreads) = u.y;
  *sec_slot(A, e, 5, tid, nthreads) = u.z;
  *sec_slot(A, e, 6, tid, nthreads) = E;
  *sec_slot(A, e, 7, tid, nthreads) = w;
  *sec_slot(A, e, 8, tid, nthreads) = w2;
  *sec_slot(A, e, 9, tid, nthreads) = (double)count;
  return true;
}
__device__ inline void pop_secondary(const DevProblem& P, const RunArgs& A, Hist& h, uint32_t tid, uint32_t nthreads) {
  const int e = h.nsec - 1;
  const double left = *sec_slot(A, e, 9, tid, nthreads) - 1.;
  if (left >= 1.) *sec_slot(A, e, 9, tid, nthreads) = left;
  else h.nsec = e;
  h.r.x = *sec_slot(A, e, 0, tid, nthreads);
  h.r.y = *sec_slot(A, e, 1, tid, nthreads);
  h.r.z = *sec_slot(A, e, 2, tid, nthreads);
  h.u.x = *sec_slot(A, e, 3, tid, nthreads);
  h.u.y = *sec_slot(A, e, 4, tid, nthreads);
  h.u.z = *sec_slot(A, e, 5, tid, nthreads);
  h.E = *sec_slot(A, e, 6, tid, nthreads);
  h.w = *sec_slot(A, e, 7, tid, nthreads);
  h.w2 = *sec_slot(A, e, 8, tid, nthreads);
  h.g = group_of(P, h.E);
  h.emid = h.g < P.G && h.E == group_mid(P, h.g);
  h.alive = true;
}

}  // namespace abl
#include "noise.cuh"
namespace abl {

// ---- one flight (+ collision) of the three trackers ---------------------------------------------------------
// MODE 0: k-eigenvalue run; 1: power-iteration generation of a noise run; 2: noise particles (see noise.cuh)
template <int MODE>
__device__ __forceinline__ void collide(const DevProblem& P, const RunArgs& A, Hist& h, Acc& acc, uint32_t tid, uint32_t nthreads) {
  if (MODE == 0) collision<false>(P, A, h, acc, tid, nthreads);
  else collision_nm<MODE>(P, A, h, acc, tid, nthreads);
}
// the copy "cross section" eta * omega / v of noise transport (material_helper.hpp:65-84)
template <int MODE>
__device__ __forceinline__ double noise_xs(const DevProblem& P, int mat, int g) {
  return MODE == 2 ? P.eta * P.w_noise / ldt(&P.speed[mat * P.G + g]) : 0.;
}

template <int TRK, int MODE>
__device__ __forceinline__ void flight(const DevProblem& P, const RunArgs& A, Hist& h, Cursor& c, Acc& acc, uint32_t tid,
                                       uint32_t nthreads) {
#define hid (A.bank.id_a[h.idx])
  bool had_collision = false;
  if (TRK == ABL_TRACK_SURFACE) {
    // SurfaceTracker::transport loop body (surface_tracker.cpp:72-146)
    const int mg = h.mat * P.G + h.g;
    const double d_coll = rng_exponential(h.rng, MODE == 2 ? ldt(&P.Et[mg]) + noise_xs<MODE>(P, h.mat, h.g) : ldt(&P.Et[mg]));
    const Boundary bound = cursor_nearest_boundary_nl(geo_tables(P), c, h.u);
    acc.flights++;
    h.n_flights++;
    const double d_min = fmin(d_coll, bound.distance);
    score_flight_all(P, A, h, d_min, acc);
    acc.k_trk += h.w * d_min * (ldt(&P.nu[mg]) * ldt(&P.Ef[mg]));
    if (bound.distance < d_coll || fabs(bound.distance - d_coll) < ABL_BOUNDRY_TOL) {
      acc.boundary++;
      if (bound.btype == ABL_BC_VACUUM) {
        note(h, 0x3000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
        leak(h, acc, bound);
      } else if (bound.btype == ABL_BC_REFLECTIVE) {
        if (!do_reflection(P, c, h, bound) || c.cell < 0) {
          raise_error(A, ABL_ERR_LOST, hid);
          h.alive = false;
          return;
        }
        note(h, 0x4000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
      } else {
        // Tracker::cross_surface + get_current (tracker.hpp:227-231)
        cursor_move(c, bound.distance, h.u);
        c.token = -bound.token;
        cursor_get_current_nl(geo_tables(P), c, h.u);
        h.r.x = h.r.x + bound.distance * h.u.x;
        h.r.y = h.r.y + bound.distance * h.u.y;
        h.r.z = h.r.z + bound.distance * h.u.z;
        if (c.cell < 0) {
          raise_error(A, ABL_ERR_LOST, hid);
          h.alive = false;
          return;
        }
        h.mat = c.mat;
        note(h, 0x7000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
      }
    } else {
      h.r.x = h.r.x + d_coll * h.u.x;
      h.r.y = h.r.y + d_coll * h.u.y;
      h.r.z = h.r.z + d_coll * h.u.z;
      cursor_move(c, d_coll, h.u);
      had_collision = true;
      note(h, 0x2000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
    }
    if (h.alive && had_collision) collide<MODE>(P, A, h, acc, tid, nthreads);
  } else if (TRK == ABL_TRACK_IMPLICIT_LEAKAGE) {
    // ImplicitLeakageDeltaTracker::transport loop body (implicit_leakage_delta_tracker.cpp:105-246): the boundary condition
    // is looked up before every flight; towards a vacuum boundary the leaking share of the weight is scored at once and the
    // flight distance is drawn from the exponential truncated at the boundary
    const double Emaj = MODE == 2 ? ldt(&P.smp[h.g]) + noise_xs<MODE>(P, h.mat, h.g) : ldt(&P.smp[h.g]);
    const Boundary bound = cursor_boundary_condition_nl(geo_tables(P), c, h.u);
    acc.flights++;
    h.n_flights++;
    double d_coll;
    if (bound.btype == ABL_BC_VACUUM) {
      const double P_leak = det_exp(-Emaj * bound.distance);
      const double P_no_leak = 1. - P_leak;
      const double w_leak = h.w * P_leak, w2_leak = h.w2 * P_leak;
      const double w_coll = h.w * P_no_leak, w2_coll = h.w2 * P_no_leak;
      acc.leak += w_leak;
      const V3 d{h.r.x + bound.distance * h.u.x - h.rb.x, h.r.y + bound.distance * h.u.y - h.rb.y,
                 h.r.z + bound.distance * h.u.z - h.rb.z};
      acc.mig += leak_mig_score(w_leak, d);
      d_coll = -det_log(1. - P_no_leak * rng_rand(h.rng)) / Emaj;
      h.w = w_leak;
      h.w2 = w2_leak;
      score_flight_all(P, A, h, bound.distance, acc);
      h.w = w_coll;
      h.w2 = w2_coll;
      score_flight_all(P, A, h, d_coll, acc);
    } else {
      d_coll = rng_exponential(h.rng, Emaj);
      score_flight_all(P, A, h, fmin(d_coll, bound.distance), acc);
    }
    bool crossed = false;
    if (bound.distance < d_coll || fabs(bound.distance - d_coll) < ABL_BOUNDRY_TOL) {
      crossed = true;
      acc.boundary++;
      if (bound.btype == ABL_BC_VACUUM) {
        note(h, 0x3000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
        leak(h, acc, bound);
      } else if (bound.btype == ABL_BC_REFLECTIVE) {
        if (!do_reflection(P, c, h, bound) || c.cell < 0) {
          raise_error(A, ABL_ERR_LOST, hid);
          h.alive = false;
          return;
        }
        note(h, 0x4000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
      } else {
        raise_error(A, ABL_ERR_LOST, hid);
        h.alive = false;
        return;
      }
    } else {
      h.r.x = h.r.x + d_coll * h.u.x;
      h.r.y = h.r.y + d_coll * h.u.y;
      h.r.z = h.r.z + d_coll * h.u.z;
      cursor_move(c, d_coll, h.u);
      cursor_get_current_nl(geo_tables(P), c, h.u);
      if (c.cell < 0) {
        raise_error(A, ABL_ERR_LOST, hid);
        h.alive = false;
        return;
      }
      h.mat = c.mat;
      const double Et = MODE == 2 ? ldt(&P.Et[h.mat * P.G + h.g]) + noise_xs<MODE>(P, h.mat, h.g) : ldt(&P.Et[h.mat * P.G + h.g]);
      if (Et - Emaj > 1.E-10) {
        raise_error(A, ABL_ERR_MAJORANT, hid);
        h.alive = false;
        return;
      }
      if (rng_rand(h.rng) < (Et / Emaj)) had_collision = true;
      note(h, (had_collision ? 0x2000000000000000ULL : 0x1000000000000000ULL) | (uint64_t)(uint32_t)(c.cell + 1));
    }
    if (h.alive && had_collision) {
      collide<MODE>(P, A, h, acc, tid, nthreads);
    } else if (h.alive && !crossed) {
      acc.virt++;
      h.n_virtual++;
    }
  } else {
    // DeltaTracker / CarterTracker loop body (delta_tracker.cpp:100-195, carter_tracker.cpp:120-230)
    const double Esample = MODE == 2 ? ldt(&P.smp[h.g]) + noise_xs<MODE>(P, h.mat, h.g) : ldt(&P.smp[h.g]);
    h.esmp = Esample;  // "Sampling XS saved for cancellation" (delta_tracker.cpp:111, carter_tracker.cpp:130)
    const double d_coll = rng_exponential(h.rng, Esample);
    Boundary bound{ABL_INF, -1, ABL_BC_NORMAL, 0};
    bool crossed = false;
    acc.flights++;
    h.n_flights++;
    cursor_move(c, d_coll, h.u);
    cursor_get_current_nl(geo_tables(P), c, h.u);
    if (c.cell < 0) {  // left the geometry: rewind to the pre-flight position and look for the boundary
      c.token = 0;
      cursor_restart_nl(geo_tables(P), c, h.r, h.u);
      bound = cursor_boundary_condition_nl(geo_tables(P), c, h.u);
      crossed = true;
    }
    score_flight_all(P, A, h, fmin(d_coll, bound.distance), acc);
    if (crossed) {
      acc.boundary++;
      if (bound.btype == ABL_BC_VACUUM) {
        note(h, 0x3000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
        leak(h, acc, bound);
      } else if (bound.btype == ABL_BC_REFLECTIVE) {
        // Tracker::do_reflection keeps the un-folded flight for the cancelators (tracker.hpp:330-352): the previous position
        // becomes the mirror image, a distance (this leg + what was flown since the last collision) behind the surface
        const V3 r_pre_refs = h.refl ? h.rprev : h.r;
        const V3 back{h.r.x - r_pre_refs.x, h.r.y - r_pre_refs.y, h.r.z - r_pre_refs.z};
        h.uprev = h.u;  // (p.set_direction(u_new), tracker.hpp:353)
        if (!do_reflection(P, c, h, bound) || c.cell < 0) {
          raise_error(A, ABL_ERR_LOST, hid);
          h.alive = false;
          return;
        }
        if (P.exact_cancel) {
          const double d = bound.distance + norm3<InlineMath>(back);
          h.rprev = V3{h.r.x - d * h.u.x, h.r.y - d * h.u.y, h.r.z - d * h.u.z};
          h.refl = true;
        }
        note(h, 0x4000000000000000ULL | (uint64_t)(uint32_t)(c.cell + 1));
      } else {
        raise_error(A, ABL_ERR_LOST, hid);
        h.alive = false;
        return;
      }
    } else {
      if (!h.refl) h.rprev = h.r;  // Particle::move (particle.hpp:125-133)
      h.refl = false;
      h.r.x = h.r.x + d_coll * h.u.x;
      h.r.y = h.r.y + d_coll * h.u.y;
      h.r.z = h.r.z + d_coll * h.u.z;
      h.mat = c.mat;
      const double Et = MODE == 2 ? ldt(&P.Et[h.mat * P.G + h.g]) + noise_xs<MODE>(P, h.mat, h.g) : ldt(&P.Et[h.mat * P.G + h.g]);
      if (TRK == ABL_TRACK_DELTA) {
        if (Et - Esample > 1.E-10) {
          raise_error(A, ABL_ERR_MAJORANT, hid);
          h.alive = false;
          return;
        }
        if (rng_rand(h.rng) < (Et / Esample)) had_collision = true;
      } else {
        if (Esample >= Et) {
          if (rng_rand(h.rng) < (Et / Esample)) had_collision = true;
        } else {  // under-estimated sampling xs: signed-weight branch (carter_tracker.cpp:192-207)
          const double D = Et / (2. * Et - Esample);
          const double F = Et / (D * Esample);
          if ((D - rng_rand(h.rng)) > 0.) {
            h.w = h.w * F;
            had_collision = true;
          } else {
            h.w = -h.w * F;
          }
        }
      }
      note(h, (had_collision ? 0x2000000000000000ULL : 0x1000000000000000ULL) | (uint64_t)(uint32_t)(c.cell + 1));
    }
    if (h.alive && had_collision) {
      collide<MODE>(P, A, h, acc, tid, nthreads);
      h.virt = false;  // set_previous_collision_real, whatever became of the particle (delta_tracker.cpp:188-192)
    } else if (h.alive) {
      h.virt = true;   // also after a boundary crossing (:193-195)
      if (!crossed) {
        acc.virt++;
        h.n_virtual++;
      }
    }
    if (TRK == ABL_TRACK_CARTER) {
      if (h.alive && fabs(h.w) >= P.wgt_split) {  // Particle::split (particle.hpp:165-173)
        const int n_new = (int)ceil(fabs(h.w));
        if (n_new > 1) {
          h.w = h.w / (double)n_new;
          h.w2 = h.w2 / (double)n_new;
          if (!push_secondary(A, h, h.u, h.E, h.w, h.w2, tid, nthreads, n_new - 1)) raise_error(A, ABL_ERR_BANK_OVERFLOW, hid);
        }
      }
    }
  }
}

#undef hid

// One CTA of TK_THREADS per SM whose warps meet at a block-wide vote once per flight: like the staged kernel
// (history.cuh), the loop body is far larger than the SM's instruction cache, and warps that stay within one iteration of
// each other share the lines they fetch.
#ifndef TK_THREADS
#define TK_THREADS 512
#endif
template <int TRK, int MODE>
__global__ void __launch_bounds__(TK_THREADS, 1) transport_kernel(const DevProblem P, const RunArgs A) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  Acc acc;
  acc.k_col = acc.k_abs = acc.k_trk = acc.leak = acc.mig = 0.;
  acc.flights = acc.real = acc.virt = acc.tl_bins = acc.sites = acc.boundary = acc.lost = acc.coll_scores = 0;
  Hist h;
  h.alive = false;
  h.nsec = 0;
  Cursor c;
  c.err = 0;
  c.np = 0;
  c.nf = 1;
  bool have = false, exhausted = false;
  const uint64_t N = A.bank.n;

  for (;;) {
    if (!have && !exhausted) {
      // take the next history (one atomic per converged group of lanes)
      unsigned long long idx;
      {
        cg::coalesced_group grp = cg::coalesced_threads();
        unsigned long long base = 0;
        if (grp.thread_rank() == 0) base = atomicAdd(A.ticket, (unsigned long long)grp.size());
        idx = grp.shfl(base, 0) + grp.thread_rank();
      }
      exhausted = idx >= N;
      if (!exhausted) {
      have = true;
      h.idx = (uint32_t)idx;
      h.r = {A.bank.x[idx], A.bank.y[idx], A.bank.z[idx]};
      h.u = {A.bank.ux[idx], A.bank.uy[idx], A.bank.uz[idx]};
      h.rb = h.r;
      h.E = A.bank.E[idx];
      h.w = A.bank.wgt[idx];
      h.w2 = (MODE == 2 && A.bank.wgt2) ? A.bank.wgt2[idx] : 0.;
      h.g = group_of(P, h.E);
      h.emid = h.g < P.G && h.E == group_mid(P, h.g);
      if (A.bank.id_c) h.rng = A.bank.id_c[idx];
      else h.rng = pcg_advance(P.seed_state, P.stride * A.bank.id_a[idx], P.jump);  // particle.hpp:188-193
      h.hash = 1469598103934665603ULL;
      h.daughter = 0;
      h.n_flights = h.n_real = h.n_virtual = h.n_fis = h.n_noise = 0;
      h.nsec = 0;
      h.alive = true;
      h.refl = false;
      h.virt = false;
      h.rprev = V3{0., 0., 0.};
      h.uprev = V3{1., 0., 0.};  // Direction() (direction.hpp:36)
      h.esmp = 0.;
      h.Eprev = 0.;
      c.token = 0;
      cursor_restart_nl(geo_tables(P), c, h.r, h.u);
      h.mat = c.mat;
      if (c.cell < 0) {  // lost at birth: warning + kill in the reference (delta_tracker.cpp:92-98)
        acc.lost++;
        h.alive = false;
      }
      }
    }
    if (__syncthreads_and(!have)) break;  // nobody holds a history and the bank is empty
    if (have && h.alive) flight<TRK, MODE>(P, A, h, c, acc, tid, nthreads);
    if (have && !h.alive) {
      if (h.nsec > 0) {  // Particle::resurect + Tracker restart (delta_tracker.cpp:197-229)
        pop_secondary(P, A, h, tid, nthreads);
        c.token = 0;
        cursor_restart_nl(geo_tables(P), c, h.r, h.u);
        if (c.cell < 0) {
          raise_error(A, ABL_ERR_LOST, A.bank.id_a[h.idx]);
          h.alive = false;
          h.nsec = 0;
        } else {
          h.mat = c.mat;
        }
      }
      if (!h.alive) {  // history finished
        A.nfis[h.idx] = h.n_fis;
        if (MODE == 1 && A.nnoise) A.nnoise[h.idx] = h.n_noise;
        if (A.tr_hash) {
          A.tr_flights[h.idx] = h.n_flights;
          A.tr_real[h.idx] = h.n_real;
          A.tr_virtual[h.idx] = h.n_virtual;
          A.tr_hash[h.idx] = h.hash;
          A.tr_rng[h.idx] = h.rng;
        }
        have = false;
      }
    }
    if (c.err) {
      raise_error(A, c.err, have ? A.bank.id_a[h.idx] : 0);
      c.err = 0;
    }
  }

  // ---- reduce the per-thread accumulators: warp shuffle, then one atomic per warp -----------------------
  __syncwarp();
  double dv[5] = {acc.k_col, acc.k_abs, acc.k_trk, acc.leak, acc.mig};
  unsigned long long cv[8] = {acc.flights, acc.real, acc.virt, acc.tl_bins, acc.sites, acc.boundary, acc.lost, acc.coll_scores};
  constexpr int NW = TK_THREADS / 32;
  __shared__ double sd[NW][5];
  __shared__ unsigned long long sc[NW][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 5; q++) {
    double v = dv[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sd[wid][q] = v;
  }
#pragma unroll
  for (int q = 0; q < 8; q++) {
    unsigned long long v = cv[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sc[wid][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    const int q = threadIdx.x;
    double v = 0.;
    for (int w = 0; w < NW; w++) v += sd[w][q];
    const int slot = q < 3 ? q : q + 1;  // scores layout: k_col,k_abs,k_trk,k_tot(unused),leak,mig
    atomicAdd(&A.scores[slot], v);
  } else if (threadIdx.x >= 32 && threadIdx.x < 40) {
    const int q = threadIdx.x - 32;
    unsigned long long v = 0;
    for (int w = 0; w < NW; w++) v += sc[w][q];
    atomicAdd(&A.counters[q], v);
  }
}

}  // namespace abl
