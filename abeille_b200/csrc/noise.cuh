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
  f.E = ldt(&T.gmid[ei]);
  const double mu = 2. * rng_rand(rng) - 1.;
  const double phi = 2. * ABL_PI * rng_rand(rng);
  f.dir = rotate_direction(u, mu, phi);
  f.delayed = false;
  f.lambda = 0.;
  if (rng_rand(rng) < P_delayed) {
    const int dg0 = ldt(&T.dg_off[mat]), ndg = ldt(&T.dg_off[mat + 1]) - dg0;
    int k = 0;
    if (ndg >= 2) k = rng_discrete(rng, T.dg_cp + dg0, ndg);
    f.delayed = true;
    f.lambda = ndg > 0 ? ldt(&dg_lambda[dg0 + k]) : 0.;
  }
  *rng_io = rng;
  return f;
}

// delayed-neutron factor lambda^2/(lambda^2+w^2) - i lambda w/(lambda^2+w^2) (transporter.cpp:426-435, noise_maker.cpp:423-431)
__device__ __forceinline__ Cplx delayed_factor(double wr, double wi, double lambda, double w_noise) {
  const double denom = (lambda * lambda) + (w_noise * w_noise);
  return cmul(wr, wi, lambda * lambda / denom, -lambda * w_noise / denom);
}

// FlatVibrationNoiseSource::C_R / C_L (flat_vibration_noise_source.cpp:155-192); C_L == C_R for n != 0, and n = 0 never passes
// the frequency gate of dEt / dN.  n >= 3: (2 / n) sin(n acos(rel_diff)) exp(-i n pi / 2) -- a function call, off the usual path.
static __device__ __noinline__ Cplx vib_C_high(int n, double rel_diff) {
  const double dn = (double)n;
  double sn, cs;
  det_sincos(dn * det_acos(rel_diff), &sn, &cs);
  const double a = (2. / dn) * sn;
  det_sincos(-dn * ABL_PI * 0.5, &sn, &cs);  // std::exp of (-0, -n pi / 2): polar(1, -n pi / 2)
  return Cplx{a * cs, a * sn};
}
__device__ __forceinline__ Cplx vib_C(const DevNoiseSrc& ns, double x) {
  double rel_diff = (x - ns.x0) / ns.eps;
  if (rel_diff > 1.) rel_diff = 1.;
  else if (rel_diff < -1.) rel_diff = -1.;
  if (ns.harmonic >= 3) return vib_C_high(ns.harmonic, rel_diff);
  const double root = sqrt(1. - (rel_diff * rel_diff));
  if (ns.harmonic == 1) return Cplx{0., -2. * root};
  return Cplx{-2. * rel_diff * root, 0.};
}
__device__ __forceinline__ double vib_get_x(const DevNoiseSrc& ns, const V3& r) { return ns.basis == 0 ? r.x : (ns.basis == 1 ? r.y : r.z); }

// MGNuclide::sample_scatter of material pm (mg_nuclide.cpp:442-461): outgoing group mid-point energy and direction
struct ScatSample {
  V3 dir;
  double E;
};
__device__ __forceinline__ ScatSample sample_scatter_nm(const DevProblem& P, uint64_t& rng, const V3& u, int pmg) {
  int ei = 0;
  if (P.G >= 2) ei = rng_discrete(rng, P.ps_cp + (size_t)pmg * P.G, P.G);
  ScatSample sc;
  sc.E = group_mid(P, ei);
  const double mu = sample_mu(P, P.angle + (size_t)pmg * P.G + ei, rng);
  const double phi = 2. * ABL_PI * rng_rand(rng);
  sc.dir = rotate_direction(u, mu, phi);
  return sc;
}

// NoiseMaker::sample_noise_source (noise_maker.cpp:277-501): the noise copy, the vibration sources (with the homogenised
// "fake" material of :106-153), the oscillation sources.  In MG a material is one nuclide with atoms_bcm = 1 and the
// nuclide id is the material index.
__device__ __forceinline__ void sample_noise_source_dev(const DevProblem& P, const RunArgs& A, Hist& h, Acc& acc) {
  bool inside_vib = false, inside_osc = false;
  for (int s = 0; s < P.n_noise_src; s++)
    if (noise_src_contains(P.noise_src[s], h.r)) {
      if (P.noise_src[s].vibration) inside_vib = true;
      else inside_osc = true;
    }
  if (!inside_vib && !inside_osc) return;
  const int mg = h.mat * P.G + h.g;
  const double Et = ldt(&P.Et[mg]), Ea = ldt(&P.Ea[mg]), Ef = ldt(&P.Ef[mg]), nu = ldt(&P.nu[mg]);
  {  // sample_noise_copy (:155-181); NoiseMaker::dEt (:60-78): vibration sources first, then oscillation sources
    double dEt_re = 0., dEt_im = 0.;
    for (int s = 0; s < P.n_noise_src; s++) {  // FlatVibrationNoiseSource::dEt (:222-244): (Et_neg - Et_pos) * C
      const DevNoiseSrc& ns = P.noise_src[s];
      if (!ns.vibration || !noise_src_contains(ns, h.r)) continue;
      if (ns.harmonic == 0) {
        dEt_re += 0.;
        dEt_im += 0.;
        continue;
      }
      const double D_Et = ldt(&P.Et[ns.mat_neg * P.G + h.g]) - ldt(&P.Et[ns.mat_pos * P.G + h.g]);
      const Cplx C = vib_C(ns, vib_get_x(ns, h.r));
      dEt_re += C.re * D_Et;
      dEt_im += C.im * D_Et;
    }
    for (int s = 0; s < P.n_noise_src; s++) {  // SquareOscillationNoiseSource::dEt (:85-113): eps_t * Sigma_t(material
      const DevNoiseSrc& ns = P.noise_src[s];    // found again at r with direction (1,0,0)) * pi
      if (ns.vibration || !noise_src_contains(ns, h.r)) continue;
      Cursor lc;
      lc.err = 0;
      lc.token = 0;
      cursor_restart_nl(geo_tables(P), lc, h.r, V3{1., 0., 0.});
      if (lc.mat < 0) {
        raise_error(A, ABL_ERR_LOST, A.bank.id_a[h.idx]);
        return;
      }
      const double xs = ldt(&P.Et[lc.mat * P.G + h.g]);
      dEt_re += ns.on ? ns.eps_t * xs * ABL_PI : 0.;
      dEt_im += 0.;
    }
    const double qre = dEt_re / Et, qim = dEt_im / Et;
    const Cplx wc = cmul(h.w, h.w2, -qre, -qim);
    append_site_nm(A.nsites, A.n_nsites, A.nsite_capacity, A.nsite_did, h.r, h.u, h.E, wc.re, wc.im, h.idx, h.n_noise++, h.daughter++);
  }

  if (inside_vib) {  // sample_vibration_noise_source (:446-501)
    // make_fake_material (:106-153): sorted union of the nuclides of the sources we are in; with a single nuclide list per
    // source {neg, pos} the concentration is 1 for every member (the mean of equal entries).  At most ABL_VIB_NUC nuclides.
    constexpr int ABL_VIB_NUC = 8;
    int nuc[ABL_VIB_NUC];
    int nn = 0;
    bool bad = false;
    for (int s = 0; s < P.n_noise_src; s++) {
      const DevNoiseSrc& ns = P.noise_src[s];
      if (!ns.vibration || !noise_src_contains(ns, h.r)) continue;
      for (int q = 0; q < 2; q++) {
        const int m = q == 0 ? ns.mat_neg : ns.mat_pos;
        bool have = false;
        for (int k = 0; k < nn; k++) have = have || nuc[k] == m;
        if (!have) {
          if (nn < ABL_VIB_NUC) nuc[nn++] = m;
          else bad = true;
        }
      }
    }
    for (int i = 1; i < nn; i++)  // insertion sort (ids ascending)
      for (int k = i; k > 0 && nuc[k - 1] > nuc[k]; k--) {
        const int t = nuc[k];
        nuc[k] = nuc[k - 1];
        nuc[k - 1] = t;
      }
    // every source we are in must hold every nuclide of the union (the reference's map::at would throw otherwise)
    for (int s = 0; s < P.n_noise_src; s++) {
      const DevNoiseSrc& ns = P.noise_src[s];
      if (!ns.vibration || !noise_src_contains(ns, h.r)) continue;
      for (int k = 0; k < nn; k++)
        if (nuc[k] != ns.mat_neg && nuc[k] != ns.mat_pos) bad = true;
    }
    if (bad) {
      raise_error(A, ABL_ERR_UNSUPPORTED, A.bank.id_a[h.idx]);
      return;
    }
    double conc[ABL_VIB_NUC];
    {
      double num_sources = 0.;
      for (int s = 0; s < P.n_noise_src; s++)
        if (P.noise_src[s].vibration && noise_src_contains(P.noise_src[s], h.r)) num_sources += 1.;
      for (int k = 0; k < nn; k++) {
        double conc_sum = 0.;
        for (int s = 0; s < P.n_noise_src; s++) {
          const DevNoiseSrc& ns = P.noise_src[s];
          if (!ns.vibration || !noise_src_contains(ns, h.r)) continue;
          conc_sum += ns.mat_neg == ns.mat_pos ? (1. + 1.) / 2. : 1.;
        }
        conc[k] = conc_sum / num_sources;
      }
    }
    double Et_fake = 0.;
    for (int k = 0; k < nn; k++) Et_fake += conc[k] * ldt(&P.Et[nuc[k] * P.G + h.g]);
    // fake_mat.sample_nuclide (material_helper.hpp:178-224)
    const double invs_Et = 1. / Et_fake;
    const double xi = rng_rand(h.rng);
    int pick = nn - 1;
    double prob_sum = 0.;
    for (int k = 0; k < nn; k++) {
      const double nuc_prob = invs_Et * conc[k] * ldt(&P.Et[nuc[k] * P.G + h.g]);
      prob_sum += nuc_prob;
      if (xi <= prob_sum) {
        pick = k;
        break;
      }
    }
    const int pm = nuc[pick], pmg = pm * P.G + h.g;
    const double N = conc[pick];
    const double pEt = ldt(&P.Et[pmg]), pEa = ldt(&P.Ea[pmg]), pEf = ldt(&P.Ef[pmg]), pnu = ldt(&P.nu[pmg]);
    double dN_re = 0., dN_im = 0.;  // NoiseMaker::dN (:80-91), FlatVibrationNoiseSource::dN (:277-312)
    for (int s = 0; s < P.n_noise_src; s++) {
      const DevNoiseSrc& ns = P.noise_src[s];
      if (!ns.vibration || !noise_src_contains(ns, h.r)) continue;
      if ((pm != ns.mat_neg && pm != ns.mat_pos) || ns.harmonic == 0) {
        dN_re += 0.;
        dN_im += 0.;
        continue;
      }
      const double D_N = (ns.mat_neg == pm ? 1. : 0.) - (ns.mat_pos == pm ? 1. : 0.);
      const Cplx C = vib_C(ns, vib_get_x(ns, h.r));
      dN_re += C.re * D_N;
      dN_im += C.im * D_N;
    }
    const double dNN_re = dN_re / N, dNN_im = dN_im / N;
    const double Etfake_Et = Et_fake / Et;
    if (ldt(&P.fissile[pm])) {  // sample_vibration_noise_fission (:183-237)
      const double k_abs = pnu * pEf / pEt;
      const int n_new = (int)floor(k_abs / A.keff + rng_rand(h.rng));
      const double P_delayed = ldt(&P.nud[pmg]) / pnu;
      const FissionTables ft{P.chi_cp, P.dg_cp, P.dg_off, P.gmid, P.G};
      for (int i = 0; i < n_new; i++) {
        const FisSample f = sample_fission_nm(ft, P.dg_lambda, &h.rng, h.u, pm, pmg, P_delayed);
        Cplx wgt{h.w, h.w2};
        if (f.delayed) wgt = delayed_factor(wgt.re, wgt.im, f.lambda, P.w_noise);
        wgt = cmul(wgt.re, wgt.im, dNN_re, dNN_im);
        wgt.re = wgt.re * Etfake_Et;
        wgt.im = wgt.im * Etfake_Et;
        append_site_nm(A.nsites, A.n_nsites, A.nsite_capacity, A.nsite_did, h.r, f.dir, f.E, wgt.re, wgt.im, h.idx, h.n_noise++, h.daughter++);
      }
    }
    const double P_scatter = 1. - (pEa / pEt);
    {  // sample_vibration_noise_scatter (:239-275)
      const ScatSample sc = sample_scatter_nm(P, h.rng, h.u, pmg);
      Cplx wgt{h.w * 1., h.w2 * 1.};
      wgt.re = wgt.re * P_scatter;
      wgt.im = wgt.im * P_scatter;
      wgt = cmul(wgt.re, wgt.im, dNN_re * Etfake_Et, dNN_im * Etfake_Et);
      append_site_nm(A.nsites, A.n_nsites, A.nsite_capacity, A.nsite_did, h.r, sc.dir, sc.E, wgt.re, wgt.im, h.idx, h.n_noise++, h.daughter++);
    }
  }

  if (!inside_osc) return;  // sample_oscillation_noise_source (:293-323)
  (void)rng_rand(h.rng);    // mat.sample_nuclide
  if (ldt(&P.fissile[h.mat])) {  // sample_oscillation_noise_fission (:383-445)
    const double k_abs = nu * Ef / Et;
    const int n_new = (int)floor(k_abs / A.keff + rng_rand(h.rng));
    const double P_delayed = ldt(&P.nud[mg]) / nu;
    double dEf_re = 0., dEf_im = 0.;
    for (int s = 0; s < P.n_noise_src; s++)
      if (!P.noise_src[s].vibration && noise_src_contains(P.noise_src[s], h.r)) {
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
    const ScatSample sc = sample_scatter_nm(P, h.rng, h.u, mg);
    Cplx wgt{h.w * 1., h.w2 * 1.};
    wgt.re = wgt.re * P_scatter;
    wgt.im = wgt.im * P_scatter;
    double dE_re = 0., dE_im = 0.;
    for (int s = 0; s < P.n_noise_src; s++)
      if (!P.noise_src[s].vibration && noise_src_contains(P.noise_src[s], h.r)) {
        dE_re += P.noise_src[s].on ? P.noise_src[s].eps_s_pi : 0.;
        dE_im += 0.;
      }
    wgt = cmul(wgt.re, wgt.im, dE_re, dE_im);
    append_site_nm(A.nsites, A.n_nsites, A.nsite_capacity, A.nsite_did, h.r, sc.dir, sc.E, wgt.re, wgt.im, h.idx, h.n_noise++, h.daughter++);
  }
  (void)acc;
}

// Transporter::collision + branching_collision (transporter.cpp:60-93,269-312) in noise mode.
// MODE 1: power-iteration generation of a noise run (noise = false); MODE 2: noise particles (noise = true).
template <int MODE>
__device__ __forceinline__ void collision_nm(const DevProblem& P, const RunArgs& A, Hist& h, Acc& acc, uint32_t tid, uint32_t nthreads) {
  constexpr bool NOISE = MODE == 2;
  const int mg = h.mat * P.G + h.g;
  const double Et0 = ldt(&P.Et[mg]), Ea = ldt(&P.Ea[mg]), Ef = ldt(&P.Ef[mg]), nu = ldt(&P.nu[mg]);
  acc.real++;
  h.n_real++;
  if (A.converged && P.n_coll_tallies) {  // mat.Et(E) without the noise term (collision_mesh_tally.cpp:35)
    const MatXS mx{Et0, Ea, Ef, ldt(&P.Es[mg])};
    for (int t = 0; t < P.ntallies; t++)
      if (P.tally[t].estimator == ABL_EST_COLLISION) {
        const int l = h.emid ? ldt(&P.tally_gbin[t * P.G + h.g]) : tally_energy_bin(P.tally[t], h.E);
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
  const double noise_copy = NOISE ? (P.eta * P.w_noise / ldt(&P.speed[mg])) / (1. * 1.) : 0.;
  const double total = NOISE ? Et0 + noise_copy : Et0;
  const double k_abs_scr = ddiv_pos(h.w * nu * Ef, total);
  if (!NOISE) acc.k_abs += k_abs_scr;
  // make_fission_neutrons (transporter.cpp:358-487)
  int n_new;
  if (!NOISE) n_new = (int)floor(ddiv_pos(fabs(k_abs_scr), A.k_col) + rng_rand(h.rng));
  else n_new = (int)floor((nu * Ef / (total * A.keff)) + rng_rand(h.rng));
  if (n_new > 0) {
    const double P_delayed = ldt(&P.nud[mg]) / nu;
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
