/* integration/gpu_transporter.hpp -- the adapter of INTEGRATION.md section 1, written against the REFERENCE'S OWN types.
 *
 * `GPUTransporter : Transporter` (include/simulation/transporter.hpp:39-47 of the reference) forwards
 * Transporter::transport(bank) to abl_transport() of libabeille_b200.so (include/abeille_b200.h) and hands the reference
 * back what its own trackers would: the fission bank as std::vector<BankedParticle> in the reference's order and the k /
 * leakage / migration-area scores through Tallies' public score_*() methods.  It is the file a maintainer of the reference
 * would add; here it is compiled into oracle/_ref/libabeille_ref.so next to the reference's sources (oracle/ref_probe.cpp,
 * ref_power_iteration_gpu), so that the reference's own PowerIterator::run() drives the B200 backend in
 * tests/test_gpu_reference_golden.py.  The C ABI is resolved with dlsym, so the library loads on machines without CUDA.
 *
 * Not done here (INTEGRATION.md section 1, table of edits): flatten_problem() from the reference's geometry / material
 * objects -- the handle comes from this repo's host library, which flattens the same YAML deck (ablh_open, ablh_backend);
 * the reference's MeshTally objects are not edited: the device tallies are recorded / cleared by the adapter itself (see
 * close_generation) and read back with tally().  k-eigenvalue mode only.
 */
#pragma once
#include <dlfcn.h>

#include <simulation/transporter.hpp>
#include <utils/error.hpp>
#include <utils/settings.hpp>

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../include/abeille_b200.h"

class GPUTransporter : public Transporter {
 public:
  GPUTransporter(std::shared_ptr<Tallies> tallies, const std::string& host_library, const std::string& yaml_deck, int device)
      : Transporter(tallies) {
    lib_ = dlopen(host_library.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!lib_) fatal_error(std::string("GPUTransporter: ") + dlerror());
    open_ = reinterpret_cast<open_fn>(dlsym(lib_, "ablh_open"));
    close_ = reinterpret_cast<close_fn>(dlsym(lib_, "ablh_close"));
    backend_ = reinterpret_cast<backend_fn>(dlsym(lib_, "ablh_backend"));
    transport_ = reinterpret_cast<transport_fn>(dlsym(lib_, "abl_transport"));
    last_error_ = reinterpret_cast<error_fn>(dlsym(lib_, "abl_last_error"));
    record_ = reinterpret_cast<record_fn>(dlsym(lib_, "abl_tallies_record"));
    clear_ = reinterpret_cast<clear_fn>(dlsym(lib_, "abl_tallies_clear"));
    fetch_ = reinterpret_cast<fetch_fn>(dlsym(lib_, "abl_tally_fetch"));
    shape_ = reinterpret_cast<shape_fn>(dlsym(lib_, "abl_tally_shape"));
    count_ = reinterpret_cast<count_fn>(dlsym(lib_, "abl_tally_count"));
    parent_ = reinterpret_cast<parent_fn>(dlsym(lib_, "abl_parent_info_download"));
    parent_state_ = reinterpret_cast<parent_state_fn>(dlsym(lib_, "abl_parent_state_download"));
    transport_noise_ = reinterpret_cast<transport_noise_fn>(dlsym(lib_, "abl_transport_noise"));
    if (!open_ || !close_ || !backend_ || !transport_ || !last_error_ || !record_ || !clear_ || !fetch_ || !shape_ || !count_)
      fatal_error("GPUTransporter: C ABI symbols missing");
    char err[512] = {0};
    ctx_ = open_(yaml_deck.c_str(), device, err, 512);
    if (!ctx_) fatal_error(std::string("GPUTransporter: ") + err);
    h_ = backend_(ctx_);
  }
  // From the reference's live objects: `problem` comes from flatten_problem() (integration/flatten_problem.hpp), the library is
  // libabeille_b200.so itself -- this repo's C++ host library and its YAML parser are not involved.
  GPUTransporter(std::shared_ptr<Tallies> tallies, const std::string& cuda_library, const abl_problem& problem, int device)
      : Transporter(tallies) {
    lib_ = dlopen(cuda_library.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!lib_) fatal_error(std::string("GPUTransporter: ") + dlerror());
    create_ = reinterpret_cast<create_fn>(dlsym(lib_, "abl_create"));
    destroy_ = reinterpret_cast<destroy_fn>(dlsym(lib_, "abl_destroy"));
    transport_ = reinterpret_cast<transport_fn>(dlsym(lib_, "abl_transport"));
    last_error_ = reinterpret_cast<error_fn>(dlsym(lib_, "abl_last_error"));
    record_ = reinterpret_cast<record_fn>(dlsym(lib_, "abl_tallies_record"));
    clear_ = reinterpret_cast<clear_fn>(dlsym(lib_, "abl_tallies_clear"));
    fetch_ = reinterpret_cast<fetch_fn>(dlsym(lib_, "abl_tally_fetch"));
    shape_ = reinterpret_cast<shape_fn>(dlsym(lib_, "abl_tally_shape"));
    count_ = reinterpret_cast<count_fn>(dlsym(lib_, "abl_tally_count"));
    parent_ = reinterpret_cast<parent_fn>(dlsym(lib_, "abl_parent_info_download"));
    parent_state_ = reinterpret_cast<parent_state_fn>(dlsym(lib_, "abl_parent_state_download"));
    transport_noise_ = reinterpret_cast<transport_noise_fn>(dlsym(lib_, "abl_transport_noise"));
    if (!create_ || !destroy_ || !transport_ || !last_error_ || !record_ || !clear_ || !fetch_ || !shape_ || !count_)
      fatal_error("GPUTransporter: C ABI symbols missing");
    if (create_(&problem, device, &h_) != 0) fatal_error(std::string("GPUTransporter: ") + last_error_(nullptr));
  }
  ~GPUTransporter() {
    if (ctx_) close_(ctx_);
    else if (h_ && destroy_) destroy_(h_);
  }

  // Mesh tallies live in HBM and are scored inside the kernels.  Between two transport() calls the reference does
  // tallies->record_generation() (if converged) and tallies->clear_generation() on its own MeshTally objects
  // (src/power_iterator.cpp:378-382); the edit of INTEGRATION.md forwards those to abl_tallies_record / _clear.  Without
  // touching the reference's MeshTally, the adapter does the same for the device arrays at the next transport() call --
  // nothing reads them in between -- and once more from finish() after the last generation.
  void close_generation() {
    if (scored_ && record_(h_, 1.0) != 0) fatal_error(std::string("GPUTransporter: ") + last_error_(h_));
    if (clear_(h_) != 0) fatal_error(std::string("GPUTransporter: ") + last_error_(h_));
    scored_ = false;
  }
  void finish() { close_generation(); }
  int ntallies() const { return count_(h_); }
  // average (which = 1) or error of the mean (which = 3) of mesh tally t, as MeshTally::write_tally would store them
  std::vector<double> tally(int t, int which) const {
    uint64_t sh[4];
    if (shape_(h_, t, sh) != 0) fatal_error("GPUTransporter: bad tally index");
    std::vector<double> out(sh[0] * sh[1] * sh[2] * sh[3]);
    if (fetch_(h_, t, which, out.data()) != 0) fatal_error(std::string("GPUTransporter: ") + last_error_(h_));
    return out;
  }

  std::vector<BankedParticle> transport(std::vector<Particle>& bank, bool noise = false,
                                        std::vector<BankedParticle>* noise_bank = nullptr,
                                        const NoiseMaker* noise_maker = nullptr) override {
    // transport(bank, false, &noise_bank, &noise_maker) also samples the noise source (src/noise.cpp:312-314); transport(nbank,
    // true, ...) moves noise particles with their complex weights (:492).  The noise sources themselves are in the problem tables.
    const bool sample_noise = noise_bank != nullptr && noise_maker != nullptr;
    const bool noise_run = settings::mode == settings::SimulationMode::NOISE;
    // (the Noise driver records its tallies per batch with a multiplier, src/noise.cpp: the per-generation bookkeeping below is
    // the k-eigenvalue one, so in a noise run the device's mesh tallies are left to the INTEGRATION.md forwarding edit)
    if (!noise_run) {
      close_generation();
      scored_ = settings::converged;
    }
    const std::size_t N = bank.size();
    for (auto* v : {&x_, &y_, &z_, &ux_, &uy_, &uz_, &E_, &w_, &w2_}) v->resize(N);
    id_.resize(N); fam_.resize(N); rng_.resize(N);
    for (std::size_t i = 0; i < N; i++) {  // AoS -> SoA (particle.hpp:68-243)
      const Particle& p = bank[i];
      x_[i] = p.r().x(); y_[i] = p.r().y(); z_[i] = p.r().z();
      ux_[i] = p.u().x(); uy_[i] = p.u().y(); uz_[i] = p.u().z();
      E_[i] = p.E(); w_[i] = p.wgt(); w2_[i] = p.wgt2();
      id_[i] = p.history_id(); fam_[i] = p.family_id();
      rng_[i] = p.rng.state_;  // source particles continue the stream they were sampled with (src/simulation.cpp:70-73)
    }
    const bool complex_out = noise || sample_noise;
    const std::size_t cap = 3 * N + 4096, ncap = sample_noise ? 6 * N + 4096 : 0;
    for (auto* v : {&ox_, &oy_, &oz_, &oux_, &ouy_, &ouz_, &oE_, &ow_, &ow2_}) v->resize(cap);
    oa_.resize(cap); ob_.resize(cap); oc_.resize(cap);
    abl_bank in{N, x_.data(), y_.data(), z_.data(), ux_.data(), uy_.data(), uz_.data(), E_.data(), w_.data(), noise ? w2_.data() : nullptr,
                id_.data(), fam_.data(), rng_.data()};
    abl_bank out{cap, ox_.data(), oy_.data(), oz_.data(), oux_.data(), ouy_.data(), ouz_.data(), oE_.data(), ow_.data(),
                 complex_out ? ow2_.data() : nullptr, oa_.data(), ob_.data(), oc_.data()};
    abl_gen_params gp{tallies->kcol(), tallies->keff(), settings::converged ? 1 : 0, noise ? 1 : 0, 0, sample_noise ? 1 : 0};
    uint64_t m = 0, mn = 0, counters[8];
    double s[6];
    if (sample_noise) {
      if (!transport_noise_) fatal_error("GPUTransporter: abl_transport_noise missing");
      for (auto* v : {&nx_, &ny_, &nz_, &nux_, &nuy_, &nuz_, &nE_, &nw_, &nw2_}) v->resize(ncap);
      na_.resize(ncap); nb_.resize(ncap); nc_.resize(ncap);
      abl_bank nout{ncap, nx_.data(), ny_.data(), nz_.data(), nux_.data(), nuy_.data(), nuz_.data(), nE_.data(), nw_.data(), nw2_.data(),
                    na_.data(), nb_.data(), nc_.data()};
      if (transport_noise_(h_, &in, &gp, &out, &m, &nout, &mn, s, counters) != 0) fatal_error(std::string("GPUTransporter: ") + last_error_(h_));
    } else if (transport_(h_, &in, &gp, &out, &m, s, counters) != 0) {
      fatal_error(std::string("GPUTransporter: ") + last_error_(h_));
    }
    tallies->score_k_col(s[0]); tallies->score_k_abs(s[1]); tallies->score_k_trk(s[2]); tallies->score_k_tot(s[3]);
    tallies->score_leak(s[4]); tallies->score_mig_area(s[5]);  // tallies.hpp:97-102
    std::vector<BankedParticle> fis(m);  // already in the reference's order (bank order, then creation order)
    for (std::size_t i = 0; i < m; i++) {
      BankedParticle& f = fis[i];
      f.r = Position(ox_[i], oy_[i], oz_[i]);
      const double u3[3] = {oux_[i], ouy_[i], ouz_[i]};
      std::memcpy(static_cast<void*>(&f.u), u3, sizeof(Direction));  // bit for bit: Direction(x, y, z) would renormalise
      f.E = oE_[i]; f.wgt = ow_[i]; f.wgt2 = noise ? ow2_[i] : 0.;
      f.parent_history_id = oa_[i]; f.parent_daughter_id = ob_[i]; f.family_id = oc_[i];
    }
    for (std::size_t i = 0; i < mn; i++) {  // Particle::empty_noise_bank order: bank order, then creation order
      BankedParticle f;
      f.r = Position(nx_[i], ny_[i], nz_[i]);
      const double u3[3] = {nux_[i], nuy_[i], nuz_[i]};
      std::memcpy(static_cast<void*>(&f.u), u3, sizeof(Direction));
      f.E = nE_[i]; f.wgt = nw_[i]; f.wgt2 = nw2_[i];
      f.parent_history_id = na_[i]; f.parent_daughter_id = nb_[i]; f.family_id = nc_[i];
      noise_bank->push_back(f);
    }
    bank.clear();  // as the reference's trackers leave it (src/delta_tracker.cpp:262)
    // what the reference's exact cancelators read from the bank (particle.hpp:52-57): kept by the kernels when the deck has one
    if (settings::regional_cancellation && parent_ && m > 0) {
      for (auto* v : {&px_, &py_, &pz_, &pe_}) v->resize(m);
      if (parent_(h_, m, px_.data(), py_.data(), pz_.data(), pe_.data()) == 0) {
        for (std::size_t i = 0; i < m; i++) {
          fis[i].parents_previous_position = Position(px_[i], py_[i], pz_[i]);
          fis[i].Esmp_parent = pe_[i];
        }
        for (auto* v : {&qx_, &qy_, &qz_, &qe1_, &qe3_, &qv_}) v->resize(m);
        if (parent_state_ && parent_state_(h_, m, qx_.data(), qy_.data(), qz_.data(), qe1_.data(), qe3_.data(), qv_.data()) == 0) {
          for (std::size_t i = 0; i < m; i++) {  // (`type: exact`, src/exact_mg_cancelator.cpp:319-327)
            const double u3[3] = {qx_[i], qy_[i], qz_[i]};
            std::memcpy(static_cast<void*>(&fis[i].parents_previous_direction), u3, sizeof(Direction));
            fis[i].parents_previous_previous_energy = qe1_[i];
            fis[i].parents_previous_energy = qe3_[i];
            fis[i].parents_previous_was_virtual = qv_[i] != 0.;
          }
        }
      }
    }
    return fis;
  }

 private:
  using open_fn = void* (*)(const char*, int, char*, int);
  using close_fn = void (*)(void*);
  using backend_fn = abl_handle (*)(void*);
  using transport_fn = int (*)(abl_handle, const abl_bank*, const abl_gen_params*, abl_bank*, uint64_t*, double*, uint64_t*);
  using error_fn = const char* (*)(abl_handle);
  using record_fn = int (*)(abl_handle, double);
  using clear_fn = int (*)(abl_handle);
  using fetch_fn = int (*)(abl_handle, int, int, double*);
  using shape_fn = int (*)(abl_handle, int, uint64_t*);
  using count_fn = int (*)(abl_handle);
  using create_fn = int (*)(const abl_problem*, int, abl_handle*);
  using destroy_fn = void (*)(abl_handle);
  create_fn create_ = nullptr;
  destroy_fn destroy_ = nullptr;
  using parent_fn = int (*)(abl_handle, uint64_t, double*, double*, double*, double*);
  using transport_noise_fn = int (*)(abl_handle, const abl_bank*, const abl_gen_params*, abl_bank*, uint64_t*, abl_bank*, uint64_t*, double*, uint64_t*);
  transport_noise_fn transport_noise_ = nullptr;
  std::vector<double> w2_, ow2_, nx_, ny_, nz_, nux_, nuy_, nuz_, nE_, nw_, nw2_;
  std::vector<uint64_t> na_, nb_, nc_;
  parent_fn parent_ = nullptr;
  using parent_state_fn = int (*)(abl_handle, uint64_t, double*, double*, double*, double*, double*, double*);
  parent_state_fn parent_state_ = nullptr;
  std::vector<double> px_, py_, pz_, pe_, qx_, qy_, qz_, qe1_, qe3_, qv_;
  record_fn record_ = nullptr; clear_fn clear_ = nullptr; fetch_fn fetch_ = nullptr; shape_fn shape_ = nullptr; count_fn count_ = nullptr;
  bool scored_ = false;
  void* lib_ = nullptr;
  void* ctx_ = nullptr;
  abl_handle h_ = nullptr;
  open_fn open_ = nullptr; close_fn close_ = nullptr; backend_fn backend_ = nullptr; transport_fn transport_ = nullptr;
  error_fn last_error_ = nullptr;
  std::vector<double> x_, y_, z_, ux_, uy_, uz_, E_, w_, ox_, oy_, oz_, oux_, ouy_, ouz_, oE_, ow_;
  std::vector<uint64_t> id_, fam_, rng_, oa_, ob_, oc_;
};
