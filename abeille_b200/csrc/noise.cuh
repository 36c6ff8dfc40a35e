/* noise.cuh -- frequency-domain neutron noise on the device (simulation: noise).
 *
 * Two things happen in noise mode that a k-eigenvalue run does not have:
 *   MODE 1  power-iteration generations that also SAMPLE THE NOISE SOURCE: every real collision inside a noise source
 *           region emits noise particles with complex weights -- NoiseMaker::sample_noise_source
 *           (src/noise_maker.cpp:277-445) with SquareOscillationNoiseSource factors
 *           (src/square_oscillation_noise_source.cpp:76-170) -- into the history's noise bank;
 *   MODE 2  transport of the noise particles themselves (transport(bank, noise = true), src/noise.cpp:492): complex
 *           weights, the extra "cross section" eta*omega/v in the sampling and total cross sections
 *           (material_helper.hpp:47-84), the noise copy (transporter.cpp:349-356), fission neutrons that inherit the
 *           parent's complex weight times the delayed-neutron factor (transporter.cpp:418-438), roulette on both weights.
 * Complex arithmetic follows GCC's std::complex<double> (libgcc __muldc3 for finite values): (a+bi)(c+di) =
 * (ac - bd) + (ad + bc)i, complex / real and complex * real act on both parts.
 *
 * Every BankedParticle the reference constructs takes the next value of the history's daughter counter
 * (particle.hpp:98), fission sites and noise particles alike, so in these modes a site carries both its RANK in the
 * history's bank (for the placement at offset[parent] + rank) and the daughter id that is written to the output bank.
 */
#pragma once
// (included by transport.cuh, after the particle-state helpers it uses)

namespace abl {

struct Cplx {
  double re, im;
};
__device__ __forceinline__ Cplx cmul(double a, double b, double c, double d) { return Cplx{a * c - b * d, a * d + b * c}; }

__device__ __forceinline__ bool noise_src_contains(const DevNoiseSrc& ns, const V3& r) {
  return r.x > ns.low[0] && r.y > ns.low[1] && r.z > ns.low[2] && r.x < ns.hi[0] && r.y < ns.hi[1] && r.z < ns.hi[2];
}

// append to a site scratch: rank = position in the history's bank, did = daughter id
__device__ __forceinline__ void append_site_nm(Site* sites, unsigned long long* n_sites, uint64_t capacity, uint32_t* did_arr,
                                               const V3& r, const V3& dir, double E, double w, double w2, uint32_t parent,
                                               uint32_t rank, uint32_t did) {
  Site s;
  s.x = r.x; s.y = r.y; s.z = r.z;
  s.ux = dir.x; s.uy = dir.y; s.uz = dir.z;
  s.E = E;
  s.w = w;
  s.w2 = w2;
  s.parent = parent;
  s.daughter = rank;
  const unsigned long long slot = atomicAdd(n_sites, 1ULL);
  if (slot < capacity) {
    const double2* src = reinterpret_cast<const double2*>(&s);
    double2* dst = reinterpret_cast<double2*>(sites + slot);
#pragma unroll
    for (int q = 0; q < 5; q++) dst[q] = src[q];
    did_arr[slot] = did;
  }
}

struct FisSample {
  V3 dir;
  double E, lambda;
  bool delayed;
};
// MGNuclide::sample_fission (mg_nuclide.cpp:504-543)
static __device__ __noinline__ FisSample sample_fission_nm(const FissionTables T, const double* __restrict__ dg_lambda, uint64_t* rng_io,
                                                    const V3 u, int mat, int mg, double P_delayed) {
  uint64_t rng = *rng_io;
  int ei = 0;
  if (T.G >= 2) ei = rng_discrete(rng, T.chi_cp + (size_t)mg * T.G, T.G);
  FisSample f;
  f.E = __ldg(&T.gmid[ei]);
  const double mu = 2. * rng_rand(rng) - 1.;
  const double phi = 2. * ABL_PI * rng_rand(rng);
  f.dir = rotate_direction(u, mu, phi);
  f.delayed = false;
  f.lambda = 0.;
  if (rng_rand(rng) < P_delayed) {
    const int dg0 = __ldg(&T.dg_off[mat]), ndg = __ldg(&T.dg_off[mat + 1]) - dg0;
    int k = 0;
    if (ndg >= 2) k = rng_discrete(rng, T.dg_cp + dg0, ndg);
    f.delayed = true;
    f.lambda = ndg > 0 ? __ldg(&dg_lambda[dg0 + k]) : 0.;
  }
  *rng_io = rng;
  return f;
}

// delayed-neutron factor lambda^2/(lambda^2+w^2) - i lambda w/(lambda^2+w^2) (transporter.cpp:426-435, noise_maker.cpp:423-431)
__device__ __forceinline__ Cplx delayed_factor(double wr, double wi, double lambda, double w_noise) {
  const double denom = (lambda * lambda) + (w_noise * w_noise);
  return cmul(wr, wi, lambda * lambda / denom, -lambda * w_noise / denom);
}

// NoiseMaker::sample_noise_source (noise_maker.cpp:277-445), oscillation sources only
__device__ __forceinline__ void sample_noise_source_dev(const DevProblem& P, const RunArgs& A, Hist& h, Acc& acc) {
  bool inside = false;
  for (int s = 0; s < P.n_noise_src; s++)
    if (noise_src_contains(P.noise_src[s], h.r)) inside = true;
  if (!inside) return;
  const int mg = h.mat * P.G + h.g;
  const double Et = __ldg(&P.Et[mg]), Ea = __ldg(&P.Ea[mg]), Ef = __ldg(&P.Ef[mg]), nu = __ldg(&P.nu[mg]);
  {  // sample_noise_copy (:155-181): dEt sums eps_t * Sigma_t(material found again at r with direction (1,0,0)) * pi
    double dEt_re = 0., dEt_im = 0.;
    for (int s = 0; s < P.n_noise_src; s++) {
      const DevNoiseSrc& ns = P.noise_src[s];
      if (!noise_src_contains(ns, h.r)) continue;
      Cursor lc;
      lc.err = 0;
      lc.token = 0;
      cursor_restart(P, lc, h.r, V3{1., 0., 0.});
      if (lc.mat < 0) {
        raise_error(A, ABL_ERR_LOST, A.bank.id_a[h.idx]);
        return;
      }
      const double xs = __ldg(&P.Et[lc.mat * P.G + h.g]);
      dEt_re += ns.on ? ns.eps_t * xs * ABL_PI : 0.;
      dEt_im += 0.;
    }
    const double qre = dEt_re / Et, qim = dEt_im / Et;
    const Cplx wc = cmul(h.w, h.w2, -qre, -qim);
    append_site_nm(A.nsites, A.n_nsites, A.nsite_capacity, A.nsite_did, h.r, h.u, h.E, wc.re, wc.im, h.idx, h.n_noise++, h.daughter++);
  }
  // sample_oscillation_noise_source (:293-323)
  (void)rng_rand(h.rng);  // mat.sample_nuclide
  if (__ldg(&P.fissile[h.mat])) {  // sample_oscillation_noise_fission (:383-445)
    const double k_abs = nu * Ef / Et;
    const int n_new = (int)floor(k_abs / A.keff + rng_rand(h.rng));
    const double P_delayed = __ldg(&P.nud[mg]) / nu;
    double dEf_re = 0., dEf_im = 0.;
    for (int s = 0; s < P.n_noise_src; s++)
      if (noise_src_contains(P.noise_src[s], h.r)) {
        dEf_re += P.noise_src[s].on ? P.noise_src[s].eps_f_pi : 0.;
        dEf_im += 0.;
      }
    const FissionTables ft{P.chi_cp, P.dg_cp, P.dg_off, P.gmid, P.G};
    for (int i = 0; i < n_new; i++) {
      const FisSample f = sample_fission_nm(ft, P.dg_lambda, &h.rng, h.u, h.mat, mg, P_delayed);
      Cplx wgt{h.w, h.w2};
      if (f.delayed) wgt = delayed_factor(wgt.re, wgt.im, f.lambda, P.w_noise);
      wgt = cmul(wgt.re, wgt.im, dEf_re, dEf_im);
      append_site_nm(A.nsites, A.n_nsites, A.nsite_capacity, A.nsite_did, h.r, f.dir, f.E, wgt.re, wgt.im, h.idx, h.n_noise++, h.daughter++);
    }
  }
  const double P_scatter = 1. - (Ea / Et);
  {  // sample_oscillation_noise_scatter (:325-381); MG: mt == 2, yield == 1
    int ei = 0;
    if (P.G >= 2) ei = rng_discrete(h.rng, P.ps_cp + (size_t)mg * P.G, P.G);
    const double E_out = group_mid(P, ei);
    const double mu = sample_mu(P, P.angle + (size_t)mg * P.G + ei, h.rng);
    const double phi = 2. * ABL_PI * rng_rand(h.rng);
    const V3 dir = rotate_direction(h.u, mu, phi);
    Cplx wgt{h.w * 1., h.w2 * 1.};
    wgt.re = wgt.re * P_scatter;
    wgt.im = wgt.im * P_scatter;
    double dE_re = 0., dE_im = 0.;
    for (int s = 0; s < P.n_noise_src; s++)
      if (noise_src_contains(P.noise_src[s], h.r)) {
        dE_re += P.noise_src[s].on ? P.noise_src[s].eps_s_pi : 0.;
        dE_im += 0.;
      }
    wgt = cmul(wgt.re, wgt.im, dE_re, dE_im);
    append_site_nm(A.nsites, A.n_nsites, A.nsite_capacity, A.nsite_did, h.r, dir, E_out, wgt.re, wgt.im, h.idx, h.n_noise++, h.daughter++);
  }
  (void)acc;
}

// Transporter::collision + branching_collision (transporter.cpp:60-93,269-312) in noise mode.
// MODE 1: power-iteration generation of a noise run (noise = false); MODE 2: noise particles (noise = true).
template <int MODE>
__device__ __forceinline__ void collision_nm(const DevProblem& P, const RunArgs& A, Hist& h, Acc& acc, uint32_t tid, uint32_t nthreads) {
  constexpr bool NOISE = MODE == 2;
  const int mg = h.mat * P.G + h.g;
  const double Et0 = __ldg(&P.Et[mg]), Ea = __ldg(&P.Ea[mg]), Ef = __ldg(&P.Ef[mg]), nu = __ldg(&P.nu[mg]);
  acc.real++;
  h.n_real++;
  if (A.converged && P.n_coll_tallies) {  // mat.Et(E) without the noise term (collision_mesh_tally.cpp:35)
    const MatXS mx{Et0, Ea, Ef, __ldg(&P.Es[mg])};
    for (int t = 0; t < P.ntallies; t++)
      if (P.tally[t].estimator == ABL_EST_COLLISION) {
        const int l = h.emid ? __ldg(&P.tally_gbin[t * P.G + h.g]) : tally_energy_bin(P.tally[t], h.E);
        acc.coll_scores += score_collision(P.tally[t], h.r, l, h.w, h.w2, mx);
      }
  }
  if (!NOISE) {
    const double k_col_scr = ddiv_pos(h.w * (nu * Ef), Et0);
    const V3 dr{h.r.x - h.rb.x, h.r.y - h.rb.y, h.r.z - h.rb.z};
    const double mig_dist = norm3(dr);
    const double mig_area_scr = ddiv_pos(h.w * Ea, Et0) * mig_dist * mig_dist;
    acc.k_col += k_col_scr;
    acc.mig += mig_area_scr;
  }
  if (MODE == 1 && A.sample_noise) sample_noise_source_dev(P, A, h, acc);

  // MaterialHelper::sample_nuclide (material_helper.hpp:178-224): one draw; in noise transport the nuclide's total
  // carries the copy cross section eta*omega/v
  (void)rng_rand(h.rng);
  const double noise_copy = NOISE ? (P.eta * P.w_noise / __ldg(&P.speed[mg])) / (1. * 1.) : 0.;
  const double total = NOISE ? Et0 + noise_copy : Et0;
  const double k_abs_scr = ddiv_pos(h.w * nu * Ef, total);
  if (!NOISE) acc.k_abs += k_abs_scr;
  // make_fission_neutrons (transporter.cpp:358-487)
  int n_new;
  if (!NOISE) n_new = (int)floor(ddiv_pos(fabs(k_abs_scr), A.k_col) + rng_rand(h.rng));
  else n_new = (int)floor((nu * Ef / (total * A.keff)) + rng_rand(h.rng));
  if (n_new > 0) {
    const double P_delayed = __ldg(&P.nud[mg]) / nu;
    const FissionTables ft{P.chi_cp, P.dg_cp, P.dg_off, P.gmid, P.G};
    for (int i = 0; i < n_new; i++) {
      const FisSample f = sample_fission_nm(ft, P.dg_lambda, &h.rng, h.u, h.mat, mg, P_delayed);
      double wgt = h.w > 0. ? 1. : -1., wgt2 = 0.;
      if (NOISE) {
        wgt = h.w;
        wgt2 = h.w2;
        if (f.delayed) {
          const Cplx d = delayed_factor(wgt, wgt2, f.lambda, P.w_noise);
          wgt = d.re;
          wgt2 = d.im;
        }
      }
      const uint32_t did = h.daughter++;
      if (!NOISE || P.inner_generations) {
        append_site_nm(A.sites, A.n_sites, A.site_capacity, A.site_did, h.r, f.dir, f.E, wgt, wgt2, h.idx, h.n_fis++, did);
      } else {
        const V3 keep = h.u;
        if (!push_secondary(A, h, f.dir, f.E, wgt, wgt2, tid, nthreads)) raise_error(A, ABL_ERR_BANK_OVERFLOW, A.bank.id_a[h.idx]);
        (void)keep;
      }
      acc.sites++;
    }
  }
  note(h, 0x5000000000000000ULL | (uint64_t)(uint32_t)n_new);
  if (NOISE) {  // make_noise_copy (transporter.cpp:349-356)
    if (noise_copy / total + rng_rand(h.rng) >= 1.) {
      const Cplx wc = cmul(h.w, h.w2, 1., -1. / P.eta);
      if (!push_secondary(A, h, h.u, h.E, wc.re, wc.im, tid, nthreads)) raise_error(A, ABL_ERR_BANK_OVERFLOW, A.bank.id_a[h.idx]);
    }
  }
  const double surv = 1. - (Ea + noise_copy) / total;  // implicit capture (transporter.cpp:295-298)
  h.w = h.w * surv;
  h.w2 = h.w2 * surv;
  russian_roulette<NOISE>(P, h);
  if (h.alive) {  // do_scatter (transporter.cpp:314-347)
    int ei = 0;
    if (P.G >= 2) ei = rng_discrete(h.rng, P.ps_cp + (size_t)mg * P.G, P.G);
    const double E_out = group_mid(P, ei);
    const double mu = sample_mu(P, P.angle + (size_t)mg * P.G + ei, h.rng);
    const double phi = 2. * ABL_PI * rng_rand(h.rng);
    h.u = rotate_direction(h.u, mu, phi);
    h.E = E_out;
    h.g = ei;
    h.emid = true;
    h.w = h.w * 1.;
    h.w2 = h.w2 * 1.;
    if (h.E < P.min_energy) h.alive = false;
  }
  note(h, 0x6000000000000000ULL | (h.alive ? (uint64_t)(h.g + 1) : 0ULL));
}

}  // namespace abl
