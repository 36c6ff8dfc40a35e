/* simulation.cpp -- Tallies, GPUTransporter, PowerIterator.  See simulation.hpp. */
#include "simulation.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace abeille {

namespace {
[[noreturn]] void fatal_error(const std::string& m) { throw std::runtime_error(m); }
void check(abl_handle h, int rc, const char* what) {
  if (rc != ABL_OK) fatal_error(std::string(what) + ": " + abl_last_error(h));
}
}  // namespace

// ---- Tallies ------------------------------------------------------------------------------------------
void Tallies::clear_generation() {
  k_col_score = k_abs_score = k_trk_score = k_tot_score = leak_score = mig_area_score = 0.;
  if (backend) check(backend, abl_tallies_clear(backend), "abl_tallies_clear");
}

void Tallies::calc_gen_values() {
  k_col = k_col_score / total_weight;
  k_abs = k_abs_score / total_weight;
  k_trk = k_trk_score / total_weight;
  leak = leak_score / total_weight;
  k_tot = k_tot_score / total_weight;
  mig = mig_area_score / total_weight;
  k_col_vec.push_back(k_col);
  k_abs_vec.push_back(k_abs);
  k_trk_vec.push_back(k_trk);
  leak_vec.push_back(leak);
  mig_vec.push_back(mig);
}

void Tallies::update_avg_and_var(double x, double& x_avg, double& x_var) const {
  const double dgen = static_cast<double>(gen);
  const double x_avg_old = x_avg, x_var_old = x_var;
  x_avg = x_avg_old + (x - x_avg_old) / (dgen);
  if (gen > 1) x_var = x_var_old + ((x - x_avg_old) * (x - x_avg_old) / (dgen)) - ((x_var_old) / (dgen - 1.));
}

void Tallies::record_generation(double multiplier) {
  gen++;
  update_avg_and_var(k_col, k_col_avg, k_col_var);
  update_avg_and_var(k_abs, k_abs_avg, k_abs_var);
  update_avg_and_var(k_trk, k_trk_avg, k_trk_var);
  update_avg_and_var(leak, leak_avg, leak_var);
  update_avg_and_var(k_tot, k_tot_avg, k_tot_var);
  update_avg_and_var(mig, mig_avg, mig_var);
  if (backend) check(backend, abl_tallies_record(backend, multiplier), "abl_tallies_record");
}

// ---- GPUTransporter -------------------------------------------------------------------------------------
GPUTransporter::GPUTransporter(std::shared_ptr<Tallies> t, const Problem& problem, int device) : Transporter(std::move(t)) {
  FlatProblem flat;
  problem.flatten(flat);
  const int rc = abl_create(&flat.p, device, &h_);
  if (rc != ABL_OK) fatal_error(std::string("abl_create: ") + abl_last_error(nullptr));
  tallies->backend = h_;
}

GPUTransporter::~GPUTransporter() {
  if (tallies) tallies->backend = nullptr;
  abl_destroy(h_);
}

std::vector<BankedParticle> GPUTransporter::transport(std::vector<Particle>& bank, bool noise, std::vector<BankedParticle>* noise_bank,
                                                      const void* noise_maker) {
  // noise_maker != nullptr: this power-iteration generation also samples the noise source into *noise_bank
  // (Noise::power_iteration, src/noise.cpp:305-318); the sources themselves were flattened into the problem tables
  const bool sample_noise = noise_bank != nullptr && noise_maker != nullptr;
  const size_t N = bank.size();
  for (auto& b : buf_) b.resize(N);
  ida_.resize(N);
  idb_.resize(N);
  idc_.resize(N);
  bool any_state = false, all_state = true;
  for (size_t i = 0; i < N; i++) {
    const Particle& p = bank[i];
    buf_[0][i] = p.r().x; buf_[1][i] = p.r().y; buf_[2][i] = p.r().z;
    buf_[3][i] = p.u().x; buf_[4][i] = p.u().y; buf_[5][i] = p.u().z;
    buf_[6][i] = p.E(); buf_[7][i] = p.wgt(); buf_[8][i] = p.wgt2();
    ida_[i] = p.history_id();
    idb_[i] = p.family_id();
    idc_[i] = p.rng_state;
    any_state = any_state || p.has_rng_state;
    all_state = all_state && p.has_rng_state;
  }
  if (any_state && !all_state) fatal_error("GPUTransporter: bank mixes particles with and without an rng state.");
  abl_bank in{};
  in.n = N;
  in.x = buf_[0].data(); in.y = buf_[1].data(); in.z = buf_[2].data();
  in.ux = buf_[3].data(); in.uy = buf_[4].data(); in.uz = buf_[5].data();
  in.E = buf_[6].data(); in.wgt = buf_[7].data(); in.wgt2 = noise ? buf_[8].data() : nullptr;
  in.id_a = ida_.data(); in.id_b = idb_.data(); in.id_c = (N && all_state) ? idc_.data() : nullptr;
  abl_gen_params gp{};
  gp.k_col = tallies->kcol();
  gp.keff = tallies->keff();
  gp.converged = converged ? 1 : 0;
  gp.noise = noise ? 1 : 0;
  gp.sample_noise_source = sample_noise ? 1 : 0;
  const bool complex_out = noise || sample_noise;
  // the output bank is sized from the problem (most reactive material, bank weight, k_col); should a generation still bank
  // more, the call is repeated with the capacity it reported -- its scores come back per call and, in a k-eigenvalue run,
  // tally_gen holds nothing but this call's scores, so it is cleared first (the reference's vectors simply grow)
  double sum_abs_w = 0.;
  for (size_t i = 0; i < N; i++) sum_abs_w += std::fabs(buf_[7][i]);
  uint64_t cap = abl_fission_capacity_hint(h_, N, sum_abs_w, gp.k_col);
  if (complex_out) cap = std::max<uint64_t>(cap, 3 * static_cast<uint64_t>(N) + 4096);
  for (int attempt = 0;; attempt++) {
    for (auto& b : obuf_) b.resize(cap);
    oa_.resize(cap); ob_.resize(cap); oc_.resize(cap);
    abl_bank out{};
    out.n = cap;
    out.x = obuf_[0].data(); out.y = obuf_[1].data(); out.z = obuf_[2].data();
    out.ux = obuf_[3].data(); out.uy = obuf_[4].data(); out.uz = obuf_[5].data();
    out.E = obuf_[6].data(); out.wgt = obuf_[7].data(); out.wgt2 = complex_out ? obuf_[8].data() : nullptr;
    out.id_a = oa_.data(); out.id_b = ob_.data(); out.id_c = oc_.data();
    uint64_t n_fis = 0, n_noise = 0, cn[8];
    double scores[6];
    int rc;
    std::vector<double> nbuf[9];
    std::vector<uint64_t> nids[3];
    if (sample_noise) {
      const uint64_t ncap = 6 * static_cast<uint64_t>(N) + 4096;
      for (auto& b : nbuf) b.resize(ncap);
      for (auto& b : nids) b.resize(ncap);
      abl_bank nout{};
      nout.n = ncap;
      nout.x = nbuf[0].data(); nout.y = nbuf[1].data(); nout.z = nbuf[2].data();
      nout.ux = nbuf[3].data(); nout.uy = nbuf[4].data(); nout.uz = nbuf[5].data();
      nout.E = nbuf[6].data(); nout.wgt = nbuf[7].data(); nout.wgt2 = nbuf[8].data();
      nout.id_a = nids[0].data(); nout.id_b = nids[1].data(); nout.id_c = nids[2].data();
      rc = abl_transport_noise(h_, &in, &gp, &out, &n_fis, &nout, &n_noise, scores, cn);
    } else {
      rc = abl_transport(h_, &in, &gp, &out, &n_fis, scores, cn);
    }
    if (rc == ABL_ERR_BANK_OVERFLOW && n_fis > cap && attempt < 2 && !complex_out) {
      cap = n_fis + n_fis / 8 + 4096;
      if (converged) check(h_, abl_tallies_clear(h_), "abl_tallies_clear");
      continue;
    }
    check(h_, rc, "abl_transport");
    tallies->score_k_col(scores[0]);
    tallies->score_k_abs(scores[1]);
    tallies->score_k_trk(scores[2]);
    tallies->score_k_tot(scores[3]);
    tallies->score_leak(scores[4]);
    tallies->score_mig_area(scores[5]);
    counters.flights += cn[0]; counters.real_collisions += cn[1]; counters.virtual_collisions += cn[2];
    counters.tl_bins += cn[3]; counters.fission_sites += cn[4]; counters.boundary_events += cn[5];
    counters.lost_at_birth += cn[6]; counters.coll_scores += cn[7];
    std::vector<BankedParticle> fission(n_fis);
    for (uint64_t i = 0; i < n_fis; i++) {
      BankedParticle& f = fission[i];
      f.r = {obuf_[0][i], obuf_[1][i], obuf_[2][i]};
      f.u = {obuf_[3][i], obuf_[4][i], obuf_[5][i]};
      f.E = obuf_[6][i];
      f.wgt = obuf_[7][i];
      f.wgt2 = noise ? obuf_[8][i] : 0.;
      f.parent_history_id = oa_[i];
      f.parent_daughter_id = ob_[i];
      f.family_id = oc_[i];
    }
    for (uint64_t i = 0; i < n_noise; i++) {  // Particle::empty_noise_bank order: bank order, then creation order
      BankedParticle f;
      f.r = {nbuf[0][i], nbuf[1][i], nbuf[2][i]};
      f.u = {nbuf[3][i], nbuf[4][i], nbuf[5][i]};
      f.E = nbuf[6][i];
      f.wgt = nbuf[7][i];
      f.wgt2 = nbuf[8][i];
      f.parent_history_id = nids[0][i];
      f.parent_daughter_id = nids[1][i];
      f.family_id = nids[2][i];
      noise_bank->push_back(f);
    }
    bank.clear();  // delta_tracker.cpp:262
    return fission;
  }
}

// ---- PowerIterator --------------------------------------------------------------------------------------
struct PowerIterator::DeviceBank {
  abl_bank b{};
  uint64_t cap = 0;
};

namespace {
// a fission bank as the columns the C ABI takes
struct HostColumns {
  std::vector<double> f[9];
  std::vector<uint64_t> ia, ib, ic;
  void resize(uint64_t n) {
    for (auto& v : f) v.resize(n);
    ia.resize(n); ib.resize(n); ic.resize(n);
  }
  void from(const std::vector<BankedParticle>& bank) {
    resize(bank.size());
    for (uint64_t i = 0; i < bank.size(); i++) {
      const BankedParticle& p = bank[i];
      f[0][i] = p.r.x; f[1][i] = p.r.y; f[2][i] = p.r.z; f[3][i] = p.u.x; f[4][i] = p.u.y; f[5][i] = p.u.z;
      f[6][i] = p.E; f[7][i] = p.wgt; f[8][i] = p.wgt2;
      ia[i] = p.parent_history_id; ib[i] = p.parent_daughter_id; ic[i] = p.family_id;
    }
  }
  void to(std::vector<BankedParticle>& bank) const {
    bank.resize(ia.size());
    for (uint64_t i = 0; i < bank.size(); i++) {
      BankedParticle& p = bank[i];
      p.r = Position{f[0][i], f[1][i], f[2][i]}; p.u = Direction{f[3][i], f[4][i], f[5][i]};
      p.E = f[6][i]; p.wgt = f[7][i]; p.wgt2 = f[8][i];
      p.parent_history_id = ia[i]; p.parent_daughter_id = ib[i]; p.family_id = ic[i];
    }
  }
  abl_bank view() {
    abl_bank b{};
    b.n = ia.size();
    b.x = f[0].data(); b.y = f[1].data(); b.z = f[2].data(); b.ux = f[3].data(); b.uy = f[4].data(); b.uz = f[5].data();
    b.E = f[6].data(); b.wgt = f[7].data(); b.wgt2 = f[8].data();
    b.id_a = ia.data(); b.id_b = ib.data(); b.id_c = ic.data();
    return b;
  }
};
}  // namespace

double GlobalRng::rand() {
  double sum = 0.0, tmp = 1.0;
  sum += static_cast<double>((*this)()) * tmp;
  tmp *= 4294967296.0;
  sum += static_cast<double>((*this)()) * tmp;
  tmp *= 4294967296.0;
  double ret = sum / tmp;
  if (ret >= 1.0) ret = std::nextafter(1.0, 0.0);
  return ret;
}

void comb_rows(const std::vector<double>& wgt, GlobalRng& rng, std::vector<uint32_t>& rows, std::vector<double>& wgts) {
  // The three std::shuffle calls of the reference permute vectors of 112-byte particles; which elements a shuffle swaps depends on
  // the length of the range and on the engine only, so the same calls on index vectors of the same lengths give the same
  // permutations -- and the comb needs nothing of a particle but its weight.  pos / neg: rows of the positive / other particles.
  std::vector<uint32_t> pos, neg;
  pos.reserve(wgt.size());
  double Wpos = 0., Wneg = 0.;
  for (size_t i = 0; i < wgt.size(); i++) {
    if (wgt[i] > 0.) { Wpos += wgt[i]; pos.push_back(static_cast<uint32_t>(i)); }
    else { Wneg += wgt[i]; neg.push_back(static_cast<uint32_t>(i)); }
  }
  const size_t Npos = static_cast<size_t>(std::ceil(Wpos));
  const size_t Nneg = static_cast<size_t>(std::ceil(std::abs(Wneg)));
  std::vector<uint32_t> picked;  // combed bank before its shuffle: row of every tooth's particle
  std::vector<double> picked_wgt;
  picked.reserve(Npos + Nneg);
  picked_wgt.reserve(Npos + Nneg);
  // teeth every avg_pos_wgt along the shuffled positive weight, the first one at a random offset
  std::shuffle(pos.begin(), pos.end(), rng);
  const double avg_pos_wgt = Wpos / static_cast<double>(Npos);
  double comb_pos = rng.rand() * avg_pos_wgt;
  double current_particle = 0.;
  for (size_t i = 0; i < pos.size(); i++) {
    current_particle += wgt[pos[i]];
    while (comb_pos < current_particle) {
      picked.push_back(pos[i]);
      picked_wgt.push_back(avg_pos_wgt);
      comb_pos += avg_pos_wgt;
    }
  }
  // the negative comb as the reference has it: its tooth spacing divides by Npos and it copies the i-th POSITIVE particle
  // (branchless_power_iterator.cpp:637-646); with no negative weights only its one draw is taken
  std::shuffle(neg.begin(), neg.end(), rng);
  const double avg_neg_wgt = std::abs(Wneg) / static_cast<double>(Npos);
  comb_pos = rng.rand() * avg_neg_wgt;
  current_particle = 0.;
  for (size_t i = 0; i < neg.size(); i++) {
    current_particle -= wgt[neg[i]];
    while (comb_pos < current_particle) {
      if (i >= pos.size()) fatal_error("comb_particles: more negative than positive particles (the reference reads past its buffer here).");
      picked.push_back(pos[i]);
      picked_wgt.push_back(-avg_neg_wgt);
      comb_pos += avg_neg_wgt;
    }
  }
  // the last shuffle, of the combed bank
  std::vector<uint32_t> order(picked.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = static_cast<uint32_t>(i);
  std::shuffle(order.begin(), order.end(), rng);
  rows.resize(order.size());
  wgts.resize(order.size());
  for (size_t i = 0; i < order.size(); i++) {
    rows[i] = picked[order[i]];
    wgts[i] = picked_wgt[order[i]];
  }
}

void comb_particles(std::vector<BankedParticle>& next_gen, GlobalRng& rng) {
  std::vector<double> wgt(next_gen.size());
  for (size_t i = 0; i < next_gen.size(); i++) wgt[i] = next_gen[i].wgt;
  std::vector<uint32_t> rows;
  std::vector<double> wgts;
  comb_rows(wgt, rng, rows, wgts);
  std::vector<BankedParticle> combed(rows.size());
  for (size_t i = 0; i < rows.size(); i++) {
    combed[i] = next_gen[rows[i]];
    combed[i].wgt = wgts[i];
  }
  next_gen.swap(combed);
}

PowerIterator::PowerIterator(const Problem& p, int device) : problem(p), device_(device) {
  tallies = std::make_shared<Tallies>(static_cast<double>(p.settings.nparticles));
  transporter = std::make_shared<GPUTransporter>(tallies, p, device);
}
PowerIterator::~PowerIterator() = default;

void PowerIterator::alloc_device_bank(DeviceBank& b, uint64_t cap) {
  check(transporter->handle(), abl_bank_alloc_device(transporter->handle(), cap, &b.b), "abl_bank_alloc_device");
  b.cap = cap;
}
void PowerIterator::free_device_bank(DeviceBank& b) {
  if (b.cap) abl_bank_free_device(transporter->handle(), &b.b);
  b.cap = 0;
}

void PowerIterator::initialize() {
  // Simulation::sample_sources (simulation.cpp:55-77) runs on the device with the reference's streams;
  // the host-vector path downloads the result.
  abl_handle h = transporter->handle();
  const uint64_t N = static_cast<uint64_t>(problem.settings.nparticles);
  DeviceBank d;
  alloc_device_bank(d, N);
  histories_counter_ = 0;
  check(h, abl_sample_source_device(h, N, histories_counter_, &d.b, nullptr), "abl_sample_source_device");
  std::vector<double> f[9];
  std::vector<uint64_t> ia(N), ib(N), ic(N);
  abl_bank host{};
  host.n = N;
  for (auto& v : f) v.resize(N);
  host.x = f[0].data(); host.y = f[1].data(); host.z = f[2].data(); host.ux = f[3].data(); host.uy = f[4].data();
  host.uz = f[5].data(); host.E = f[6].data(); host.wgt = f[7].data(); host.wgt2 = f[8].data();
  host.id_a = ia.data(); host.id_b = ib.data(); host.id_c = ic.data();
  check(h, abl_bank_download(h, &d.b, N, &host), "abl_bank_download");
  free_device_bank(d);
  bank_.clear();
  bank_.reserve(N);
  for (uint64_t i = 0; i < N; i++) {
    Particle p(Position{f[0][i], f[1][i], f[2][i]}, Direction{f[3][i], f[4][i], f[5][i]}, f[6][i], f[7][i], ia[i]);
    p.set_family_id(ib[i]);
    p.has_rng_state = true;  // the particle continues the stream its source sampling used (simulation.cpp:70-73)
    p.rng_state = ic[i];
    bank_.push_back(p);
  }
  histories_counter_ += N;
  global_histories_counter_ = histories_counter_;
  global_rng_.initialize(problem.settings.rng_seed);  // Simulation::Simulation (simulation.cpp:52)
  initialized_ = true;
}

double PowerIterator::entropy_from_bins(const std::vector<double>& bins, double total) const {
  double sum = 0.;
  for (double b : bins) {
    const double p = std::fabs(b) / total;
    if (p > 1.0) {
    } else if (p != 0.) {
      sum -= p * std::log2(p);
    }
  }
  return sum;
}

void PowerIterator::run(int ngenerations, int nignored, bool resident) {
  if (!initialized_) initialize();
  if (resident) run_resident(ngenerations, nignored);
  else run_host(ngenerations, nignored);
}

// The reference's data flow: host vectors across transport(), serial host code in between
void PowerIterator::run_host(int ngenerations, int nignored) {
  const Settings& st = problem.settings;
  abl_handle h = transporter->handle();
  transporter->converged = (nignored == 0);
  const bool have_entropy = problem.entropy.present;
  const MeshSpec& em = problem.entropy;
  const double edx = have_entropy ? (em.hi[0] - em.low[0]) / static_cast<double>(em.N[0]) : 1.;
  const double edy = have_entropy ? (em.hi[1] - em.low[1]) / static_cast<double>(em.N[1]) : 1.;
  const double edz = have_entropy ? (em.hi[2] - em.low[2]) / static_cast<double>(em.N[2]) : 1.;
  std::vector<double> ebins(have_entropy ? static_cast<size_t>(em.N[0]) * em.N[1] * em.N[2] : 0, 0.);
  // upload scratch for source tallies / cancellation on the device
  bool have_source_tally = false;
  for (const auto& t : problem.tallies)
    if (t.flat.estimator == ABL_EST_SOURCE && !t.flat.noise_source) have_source_tally = true;
  const bool cancel = st.regional_cancellation && problem.cancelator.present;
  const bool exact = cancel && (problem.cancelator.kind == ABL_CANCEL_BASIC_EXACT || problem.cancelator.kind == ABL_CANCEL_EXACT);
  // branchless-k-eigenvalue: the normalised bank is combed before the source tally sees it (branchless_power_iterator.cpp:358-384)
  const bool comb = st.mode == ABL_MODE_BRANCHLESS && st.branchless_combing;
  const auto t0 = std::chrono::steady_clock::now();
  for (int g = 1; g <= ngenerations; g++) {
    if (transporter->converged) active_particles += static_cast<double>(bank_.size());
    nbank_vec.push_back(bank_.size());
    std::vector<BankedParticle> next_gen = transporter->transport(bank_);
    if (next_gen.empty()) fatal_error("No fission neutrons were produced.");
    double etotal = 0.;
    if (have_entropy) {  // compute_pre_cancellation_entropy (power_iterator.cpp:588-640), total sign
      std::fill(ebins.begin(), ebins.end(), 0.);
      for (const auto& p : next_gen) {
        const int nx = static_cast<int>(std::floor((p.r.x - em.low[0]) / edx));
        const int ny = static_cast<int>(std::floor((p.r.y - em.low[1]) / edy));
        const int nz = static_cast<int>(std::floor((p.r.z - em.low[2]) / edz));
        if (nx >= 0 && nx < em.N[0] && ny >= 0 && ny < em.N[1] && nz >= 0 && nz < em.N[2]) {
          etotal += p.wgt;
          ebins[static_cast<size_t>((em.N[1] * em.N[2]) * nx + em.N[2] * ny + nz)] += p.wgt;
        }
      }
    }
    tallies->calc_gen_values();
    // (with the comb the weights are normalised on the host in the reference's serial order: ceil(sum of weights) decides the
    // combed population, and that sum sits within rounding of an integer)
    const bool device_block = cancel || (have_source_tally && transporter->converged && !comb);
    if (device_block) {
      // cancellation and the source mesh tally run on the device (bank_ops.cuh) on an uploaded copy
      uint64_t M = next_gen.size();
      std::vector<double> f[9];
      std::vector<uint64_t> ia(M), ib(M), ic(M);
      for (auto& v : f) v.resize(M);
      for (uint64_t i = 0; i < M; i++) {
        const BankedParticle& p = next_gen[i];
        f[0][i] = p.r.x; f[1][i] = p.r.y; f[2][i] = p.r.z; f[3][i] = p.u.x; f[4][i] = p.u.y; f[5][i] = p.u.z;
        f[6][i] = p.E; f[7][i] = p.wgt; f[8][i] = p.wgt2;
        ia[i] = p.parent_history_id; ib[i] = p.parent_daughter_id; ic[i] = p.family_id;
      }
      abl_bank host{};
      host.n = M;
      host.x = f[0].data(); host.y = f[1].data(); host.z = f[2].data(); host.ux = f[3].data(); host.uy = f[4].data();
      host.uz = f[5].data(); host.E = f[6].data(); host.wgt = f[7].data(); host.wgt2 = f[8].data();
      host.id_a = ia.data(); host.id_b = ib.data(); host.id_c = ic.data();
      DeviceBank d;
      alloc_device_bank(d, exact ? 2 * M + 4096 : M);
      check(h, abl_bank_upload(h, &host, &d.b), "abl_bank_upload");
      if (cancel && exact) {
        // BasicExactMGCancelator on the device copy (its rows are the rows of the transport call's bank, whose parents' data the
        // handle holds); the uniform particles it appends come back with the weights
        uint64_t rng2[2] = {global_rng_.state, global_rng_.inc};
        check(h, abl_cancel_exact_device(h, &d.b, d.cap, rng2, nullptr), "abl_cancel_exact_device");
        global_rng_.state = rng2[0];
        if (d.b.n > M) {
          HostColumns cols;
          cols.resize(d.b.n);
          abl_bank all = cols.view();
          check(h, abl_bank_download(h, &d.b, d.b.n, &all), "abl_bank_download");
          std::vector<BankedParticle> grown;
          cols.to(grown);
          for (uint64_t i = M; i < d.b.n; i++) next_gen.push_back(grown[i]);
          M = d.b.n;
          for (auto& v : f) v.resize(M);
        }
      } else if (cancel) {
        check(h, abl_cancel_device(h, &d.b, nullptr), "abl_cancel_device");
      }
      if (!comb) {
        double ws[4];
        check(h, abl_bank_weight_stats_device(h, &d.b, ws, nullptr), "abl_bank_weight_stats_device");
        const double w_per_part = static_cast<double>(st.nparticles) / (ws[2] - ws[3]);
        check(h, abl_bank_scale_weights_device(h, &d.b, w_per_part, nullptr), "abl_bank_scale_weights_device");
        if (transporter->converged) check(h, abl_score_source_device(h, &d.b, 0, nullptr), "abl_score_source_device");
      }
      abl_bank wonly{};
      wonly.n = M;
      wonly.wgt = f[7].data();
      check(h, abl_bank_download(h, &d.b, M, &wonly), "abl_bank_download");
      free_device_bank(d);
      for (uint64_t i = 0; i < M; i++) next_gen[i].wgt = f[7][i];
    }
    if (!device_block || comb) {
      // normalize_weights (power_iterator.cpp:538-586)
      double W_neg = 0., W_pos = 0.;
      for (const auto& p : next_gen) {
        if (p.wgt > 0.) W_pos += p.wgt;
        else W_neg -= p.wgt;
      }
      const double w_per_part = static_cast<double>(st.nparticles) / (W_pos - W_neg);
      for (auto& p : next_gen) p.wgt *= w_per_part;
    }
    if (comb) {
      comb_particles(next_gen, global_rng_);
      if (have_source_tally && transporter->converged) {
        HostColumns cols;
        cols.from(next_gen);
        abl_bank host = cols.view();
        DeviceBank d;
        alloc_device_bank(d, host.n);
        check(h, abl_bank_upload(h, &host, &d.b), "abl_bank_upload");
        check(h, abl_score_source_device(h, &d.b, 0, nullptr), "abl_score_source_device");
        free_device_bank(d);
      }
    }
    if (transporter->converged) tallies->record_generation();
    tallies->clear_generation();
    entropy_vec.push_back(have_entropy ? entropy_from_bins(ebins, etotal) : 0.);
    // bank rebuild with fresh history ids and streams (power_iterator.cpp:386-404)
    bank_.clear();
    histories_counter_ = global_histories_counter_;
    bank_.reserve(next_gen.size());
    for (const auto& p : next_gen) {
      Particle np(p.r, p.u, p.E, p.wgt, histories_counter_++);
      np.set_family_id(p.family_id);
      bank_.push_back(np);
    }
    global_histories_counter_ += next_gen.size();
    if (out_of_time(g, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count())) break;  // check_time(g)
    if (g == nignored) transporter->converged = true;
  }
  seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Device-resident generation loop: the same steps, the bank never leaves HBM
void PowerIterator::run_resident(int ngenerations, int nignored) {
  const Settings& st = problem.settings;
  abl_handle h = transporter->handle();
  transporter->converged = (nignored == 0);
  const uint64_t N0 = bank_.size();
  // (the first generation runs with k_col = 1, so it banks about k_inf sites per particle: tallies.cpp:48)
  uint64_t cap = abl_fission_capacity_hint(h, std::max<uint64_t>(N0, static_cast<uint64_t>(st.nparticles)), 0., std::min(1., tallies->kcol()));
  const bool exact_cancel = st.regional_cancellation && problem.cancelator.present &&
                            (problem.cancelator.kind == ABL_CANCEL_BASIC_EXACT || problem.cancelator.kind == ABL_CANCEL_EXACT);
  if (exact_cancel) cap = 2 * cap + 4096;  // room for the uniform particles the cancelator appends
  // cur / nxt are views (n and id_c change per generation); *_alloc keep the full allocations
  DeviceBank cur, nxt;
  alloc_device_bank(cur, cap);
  alloc_device_bank(nxt, cap);
  DeviceBank cur_alloc = cur, nxt_alloc = nxt;
  {  // upload the current host bank once
    std::vector<double> f[9];
    std::vector<uint64_t> ia(N0), ib(N0), ic(N0);
    for (auto& v : f) v.resize(N0);
    bool all_state = true;
    for (uint64_t i = 0; i < N0; i++) {
      const Particle& p = bank_[i];
      f[0][i] = p.r().x; f[1][i] = p.r().y; f[2][i] = p.r().z; f[3][i] = p.u().x; f[4][i] = p.u().y; f[5][i] = p.u().z;
      f[6][i] = p.E(); f[7][i] = p.wgt(); f[8][i] = p.wgt2();
      ia[i] = p.history_id(); ib[i] = p.family_id(); ic[i] = p.rng_state;
      all_state = all_state && p.has_rng_state;
    }
    abl_bank host{};
    host.n = N0;
    host.x = f[0].data(); host.y = f[1].data(); host.z = f[2].data(); host.ux = f[3].data(); host.uy = f[4].data();
    host.uz = f[5].data(); host.E = f[6].data(); host.wgt = f[7].data(); host.wgt2 = f[8].data();
    host.id_a = ia.data(); host.id_b = ib.data(); host.id_c = ic.data();
    abl_bank dst = cur.b;
    check(h, abl_bank_upload(h, &host, &dst), "abl_bank_upload");
    cur.b.n = N0;
    if (!all_state) cur.b.id_c = nullptr;  // streams derived from the history ids
  }
  const bool have_entropy = problem.entropy.present;
  const size_t nebins = have_entropy ? static_cast<size_t>(problem.entropy.N[0]) * problem.entropy.N[1] * problem.entropy.N[2] : 0;
  double* ebins_dev = nullptr;
  std::vector<double> ebins(nebins + 1, 0.);
  if (have_entropy) check(h, abl_device_alloc(h, (nebins + 1) * sizeof(double), reinterpret_cast<void**>(&ebins_dev)), "abl_device_alloc");
  const bool cancel = st.regional_cancellation && problem.cancelator.present;
  const bool exact = cancel && (problem.cancelator.kind == ABL_CANCEL_BASIC_EXACT || problem.cancelator.kind == ABL_CANCEL_EXACT);
  const bool comb = st.mode == ABL_MODE_BRANCHLESS && st.branchless_combing;
  const auto t0 = std::chrono::steady_clock::now();
  for (int g = 1; g <= ngenerations; g++) {
    const uint64_t N = cur.b.n;
    if (transporter->converged) active_particles += static_cast<double>(N);
    nbank_vec.push_back(N);
    abl_gen_params gp{};
    gp.k_col = tallies->kcol();
    gp.keff = tallies->keff();
    gp.converged = transporter->converged ? 1 : 0;
    if (st.families) {  // the families that enter the generation (src/power_iterator.cpp:326-331): distinct family ids of the bank
      std::vector<uint64_t> fam(N);
      abl_bank ids{};
      ids.n = N;
      ids.id_b = fam.data();
      check(h, abl_bank_download(h, &cur.b, N, &ids), "abl_bank_download");
      std::sort(fam.begin(), fam.end());
      families_vec.push_back(static_cast<double>(std::unique(fam.begin(), fam.end()) - fam.begin()));
    }
    abl_bank out = nxt.b;
    out.n = nxt.cap;
    uint64_t n_fis = 0, cn[8];
    double scores[6];
    for (int attempt = 0;; attempt++) {
      const int rc = abl_transport_device(h, &cur.b, &gp, &out, &n_fis, scores, cn, nullptr);
      if (rc == ABL_ERR_BANK_OVERFLOW && n_fis > nxt.cap && attempt < 2) {
        // more sites than the output bank holds: grow it to what the call reported and repeat the generation (its scores
        // come back per call; tally_gen holds only this generation's scores and is cleared first)
        free_device_bank(nxt_alloc);
        alloc_device_bank(nxt_alloc, n_fis + n_fis / 8 + 4096);
        nxt.b = nxt_alloc.b;
        nxt.cap = nxt_alloc.cap;
        out = nxt.b;
        out.n = nxt.cap;
        if (transporter->converged) check(h, abl_tallies_clear(h), "abl_tallies_clear");
        continue;
      }
      check(h, rc, "abl_transport_device");
      break;
    }
    if (n_fis == 0) fatal_error("No fission neutrons were produced.");
    tallies->score_k_col(scores[0]); tallies->score_k_abs(scores[1]); tallies->score_k_trk(scores[2]);
    tallies->score_k_tot(scores[3]); tallies->score_leak(scores[4]); tallies->score_mig_area(scores[5]);
    Counters& c = transporter->counters;
    c.flights += cn[0]; c.real_collisions += cn[1]; c.virtual_collisions += cn[2]; c.tl_bins += cn[3];
    c.fission_sites += cn[4]; c.boundary_events += cn[5]; c.lost_at_birth += cn[6]; c.coll_scores += cn[7];
    out.n = n_fis;
    double entropy = 0.;
    if (have_entropy) {
      check(h, abl_device_zero(h, ebins_dev, (nebins + 1) * sizeof(double), nullptr), "abl_device_zero");
      check(h, abl_entropy_bin_device(h, &out, ebins_dev, ebins_dev + nebins, nullptr), "abl_entropy_bin_device");
      check(h, abl_device_read(h, ebins.data(), ebins_dev, (nebins + 1) * sizeof(double), nullptr), "abl_device_read");
      const double total = ebins[nebins];
      entropy = entropy_from_bins(std::vector<double>(ebins.begin(), ebins.begin() + static_cast<long>(nebins)), total);
      if (st.empty_entropy_bins)  // Entropy::calculate_empty_fraction (src/entropy.cpp:95-105)
        empty_entropy_frac_vec.push_back(static_cast<double>(std::count(ebins.begin(), ebins.begin() + static_cast<long>(nebins), 0.)) /
                                         static_cast<double>(nebins));
    }
    tallies->calc_gen_values();
    if (cancel && exact) {
      // (the uniform particles are appended to the bank: the banks of such a problem are allocated with room for as many again)
      uint64_t rng2[2] = {global_rng_.state, global_rng_.inc};
      check(h, abl_cancel_exact_device(h, &out, nxt.cap, rng2, nullptr), "abl_cancel_exact_device");
      global_rng_.state = rng2[0];
      n_fis = out.n;
    } else if (cancel) {
      check(h, abl_cancel_device(h, &out, nullptr), "abl_cancel_device");
    }
    if (!comb) {
      double ws[4];
      check(h, abl_bank_weight_stats_device(h, &out, ws, nullptr), "abl_bank_weight_stats_device");
      const double w_per_part = static_cast<double>(st.nparticles) / (ws[2] - ws[3]);
      check(h, abl_bank_scale_weights_device(h, &out, w_per_part, nullptr), "abl_bank_scale_weights_device");
    } else {
      // the comb is the reference's serial host step on the gathered bank (std::shuffle on the one global engine).  It is a function
      // of the weights alone, so only the weights come to the host; they are normalised there too, in the reference's serial order
      // (ceil(sum of weights) decides the combed population and that sum sits within rounding of an integer), and the combed bank
      // is a gather of the device bank's rows.
      std::vector<double> wgt(n_fis);
      abl_bank wonly{};
      wonly.n = n_fis;
      wonly.wgt = wgt.data();
      check(h, abl_bank_download(h, &out, n_fis, &wonly), "abl_bank_download");
      double W_neg = 0., W_pos = 0.;
      for (double v : wgt) {
        if (v > 0.) W_pos += v;
        else W_neg -= v;
      }
      const double w_per_part = static_cast<double>(st.nparticles) / (W_pos - W_neg);
      for (double& v : wgt) v *= w_per_part;
      std::vector<uint32_t> rows;
      std::vector<double> wgts;
      comb_rows(wgt, global_rng_, rows, wgts);
      // gather into the bank the generation came from (its particles are spent); it then already is the next `cur`
      if (rows.size() > cur_alloc.cap) {
        free_device_bank(cur_alloc);
        alloc_device_bank(cur_alloc, rows.size() + rows.size() / 8 + 4096);
      }
      abl_bank dst = cur_alloc.b;
      check(h, abl_bank_gather_device(h, &out, rows.data(), wgts.data(), rows.size(), &dst, nullptr), "abl_bank_gather_device");
      std::swap(cur_alloc, nxt_alloc);  // (undone by the swap at the end of the generation: the combed bank becomes `cur`)
      nxt.b = nxt_alloc.b;
      nxt.cap = nxt_alloc.cap;
      out = nxt.b;
      n_fis = rows.size();
      out.n = n_fis;
    }
    if (st.pair_distance_sqrd) {
      // PowerIterator::compute_pair_dist_sqrd over the normalised bank (src/power_iterator.cpp:362-365,637-663): the double sum over
      // all pairs equals the weighted second moment about the weighted centroid -- two passes of abl_bank_moments_device
      const double zero[3] = {0., 0., 0.};
      double m1[5], m2[5];
      check(h, abl_bank_moments_device(h, &out, zero, m1, nullptr), "abl_bank_moments_device");
      const double c[3] = {m1[1] / m1[0], m1[2] / m1[0], m1[3] / m1[0]};
      check(h, abl_bank_moments_device(h, &out, c, m2, nullptr), "abl_bank_moments_device");
      r_sqrd_vec.push_back((m2[4] - (m2[1] * m2[1] + m2[2] * m2[2] + m2[3] * m2[3]) / m2[0]) / m2[0]);
    }
    if (transporter->converged) {
      check(h, abl_score_source_device(h, &out, 0, nullptr), "abl_score_source_device");
      tallies->record_generation();
    }
    tallies->clear_generation();
    entropy_vec.push_back(entropy);
    check(h, abl_bank_to_particles_device(h, &out, global_histories_counter_, nullptr), "abl_bank_to_particles_device");
    global_histories_counter_ += n_fis;
    histories_counter_ = global_histories_counter_;
    // swap: the fission bank becomes the particle bank; its streams follow from the new history ids
    std::swap(cur_alloc, nxt_alloc);
    cur.b = cur_alloc.b;
    cur.cap = cur_alloc.cap;
    cur.b.n = n_fis;
    cur.b.id_c = nullptr;
    nxt.b = nxt_alloc.b;
    nxt.cap = nxt_alloc.cap;
    const uint64_t want = (exact_cancel ? 2 : 1) * abl_fission_capacity_hint(h, n_fis, static_cast<double>(st.nparticles), tallies->kcol());
    if (want > nxt.cap) {  // the population drifted upwards: regrow the output bank
      free_device_bank(nxt_alloc);
      alloc_device_bank(nxt_alloc, want + want / 8);
      nxt.b = nxt_alloc.b;
      nxt.cap = nxt_alloc.cap;
    }
    if (out_of_time(g, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count())) break;  // check_time(g)
    if (g == nignored) transporter->converged = true;
  }
  seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // bring the final bank back to the host side (write_source, simulation.cpp:137-175)
  {
    const uint64_t N = cur.b.n;
    std::vector<double> f[9];
    std::vector<uint64_t> ia(N), ib(N);
    for (auto& v : f) v.resize(N);
    abl_bank host{};
    host.n = N;
    host.x = f[0].data(); host.y = f[1].data(); host.z = f[2].data(); host.ux = f[3].data(); host.uy = f[4].data();
    host.uz = f[5].data(); host.E = f[6].data(); host.wgt = f[7].data(); host.wgt2 = f[8].data();
    host.id_a = ia.data(); host.id_b = ib.data(); host.id_c = nullptr;
    abl_bank src = cur_alloc.b;
    check(h, abl_bank_download(h, &src, N, &host), "abl_bank_download");
    bank_.clear();
    bank_.reserve(N);
    for (uint64_t i = 0; i < N; i++) {
      Particle p(Position{f[0][i], f[1][i], f[2][i]}, Direction{f[3][i], f[4][i], f[5][i]}, f[6][i], f[7][i], ia[i]);
      p.set_family_id(ib[i]);
      bank_.push_back(p);
    }
  }
  if (ebins_dev) abl_device_free(h, ebins_dev);
  free_device_bank(cur_alloc);
  free_device_bank(nxt_alloc);
}

// ---- output ---------------------------------------------------------------------------------------------
void write_npy(const std::string& path, const std::vector<double>& data, const std::vector<uint64_t>& shape) {
  std::string dict = "{'descr': '<f8', 'fortran_order': False, 'shape': (";
  for (size_t i = 0; i < shape.size(); i++) dict += std::to_string(shape[i]) + (shape.size() == 1 || i + 1 < shape.size() ? "," : "");
  dict += "), }";
  size_t total = 10 + dict.size() + 1;
  const size_t pad = (64 - total % 64) % 64;
  dict += std::string(pad, ' ');
  dict += "\n";
  std::ofstream f(path, std::ios::binary);
  if (!f) fatal_error("cannot write " + path);
  const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  f.write(reinterpret_cast<const char*>(magic), 8);
  const uint16_t hlen = static_cast<uint16_t>(dict.size());
  f.write(reinterpret_cast<const char*>(&hlen), 2);
  f.write(dict.data(), static_cast<std::streamsize>(dict.size()));
  f.write(reinterpret_cast<const char*>(data.data()), static_cast<std::streamsize>(data.size() * sizeof(double)));
}

bool PowerIterator::out_of_time(int gen, double loop_seconds) const {
  const double T_avg = loop_seconds / static_cast<double>(gen);
  const double T_used = seconds + loop_seconds;  // (the reference's alpha_omega_timer also holds the parse; here: every loop so far)
  return problem.settings.max_time - T_used < 2. * T_avg;
}

void PowerIterator::write_results(const std::string& dir) const {
  abl_handle h = transporter->handle();
  write_npy(dir + "/kcol.npy", tallies->k_col_vec, {tallies->k_col_vec.size()});
  write_npy(dir + "/ktrk.npy", tallies->k_trk_vec, {tallies->k_trk_vec.size()});
  write_npy(dir + "/kabs.npy", tallies->k_abs_vec, {tallies->k_abs_vec.size()});
  write_npy(dir + "/leakage.npy", tallies->leak_vec, {tallies->leak_vec.size()});
  write_npy(dir + "/mig-area.npy", tallies->mig_vec, {tallies->mig_vec.size()});
  write_npy(dir + "/entropy.npy", entropy_vec, {entropy_vec.size()});
  // results/families, results/pair-dist-sqrd, results/empty-entropy-frac (src/power_iterator.cpp:475-503)
  if (!families_vec.empty()) write_npy(dir + "/families.npy", families_vec, {families_vec.size()});
  if (!r_sqrd_vec.empty()) write_npy(dir + "/pair-dist-sqrd.npy", r_sqrd_vec, {r_sqrd_vec.size()});
  if (!empty_entropy_frac_vec.empty()) write_npy(dir + "/empty-entropy-frac.npy", empty_entropy_frac_vec, {empty_entropy_frac_vec.size()});
  for (int t = 0; t < abl_tally_count(h); t++) {
    uint64_t sh[4];
    check(h, abl_tally_shape(h, t, sh), "abl_tally_shape");
    std::vector<double> buf(sh[0] * sh[1] * sh[2] * sh[3]);
    const std::vector<uint64_t> shape{sh[0], sh[1], sh[2], sh[3]};
    check(h, abl_tally_fetch(h, t, 1, buf.data()), "abl_tally_fetch");
    write_npy(dir + "/" + problem.tallies[static_cast<size_t>(t)].name + "_avg.npy", buf, shape);
    check(h, abl_tally_fetch(h, t, 3, buf.data()), "abl_tally_fetch");
    write_npy(dir + "/" + problem.tallies[static_cast<size_t>(t)].name + "_std.npy", buf, shape);
    // the attributes MeshTally::write_tally stores with the arrays (src/mesh_tally.cpp:166-190): mesh coordinates, energy
    // bounds, quantity, estimator (HDF5 attributes in the reference; here one .npy per array and a text file)
    const MeshTallySpec& spec = problem.tallies[static_cast<size_t>(t)];
    const char* axis[3] = {"x", "y", "z"};
    for (int a = 0; a < 3; a++) {
      const uint64_t n = static_cast<uint64_t>(spec.flat.N[a]);
      const double d = (spec.flat.hi[a] - spec.flat.low[a]) / static_cast<double>(n);
      std::vector<double> bounds(n + 1);
      for (uint64_t i = 0; i <= n; i++) bounds[i] = (static_cast<double>(i) * d) + spec.flat.low[a];
      write_npy(dir + "/" + spec.name + "_" + axis[a] + "-bounds.npy", bounds, {n + 1});
    }
    write_npy(dir + "/" + spec.name + "_energy-bounds.npy", spec.energy_bounds, {spec.energy_bounds.size()});
    std::ofstream attrs(dir + "/" + spec.name + "_attributes.txt");
    if (!attrs) fatal_error("cannot write the attributes of tally " + spec.name);
    attrs << "quantity: " << spec.quantity_str << "\n";
    if (spec.quantity_str == "mt") attrs << "mt: " << spec.mt << "\n";
    attrs << "estimator: " << spec.estimator_str << "\n";
  }
  // the final source bank, one row per particle: x y z ux uy uz E wgt wgt2 (Simulation::write_source,
  // src/simulation.cpp:137-175)
  {
    std::vector<double> src(bank_.size() * 9);
    for (size_t i = 0; i < bank_.size(); i++) {
      const Particle& p = bank_[i];
      double* row = &src[i * 9];
      row[0] = p.r().x; row[1] = p.r().y; row[2] = p.r().z;
      row[3] = p.u().x; row[4] = p.u().y; row[5] = p.u().z;
      row[6] = p.E(); row[7] = p.wgt(); row[8] = p.wgt2();
    }
    write_npy(dir + "/source.npy", src, {static_cast<uint64_t>(bank_.size()), 9});
  }
}

}  // namespace abeille
