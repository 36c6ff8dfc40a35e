/* simulation.hpp -- the caller side of the drop-in boundary, kept in the reference's shape:
 *
 *   Particle / BankedParticle   include/simulation/particle.hpp:38-243 (the fields transport consumes)
 *   Tallies                     include/simulation/tallies.hpp, src/tallies.cpp:100-203,268-280
 *   Transporter (abstract)      include/simulation/transporter.hpp:39-47
 *   GPUTransporter              the new backend: marshals to SoA and calls the C ABI (abl_transport)
 *   PowerIterator               src/power_iterator.cpp:46-133,305-473,538-586 (k-eigenvalue driver)
 *
 * A maintainer of the reference adds GPUTransporter next to DeltaTracker/SurfaceTracker/CarterTracker
 * and selects it in make_transporter() (src/parser.cpp:889-911); see INTEGRATION.md.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "model.hpp"

namespace abeille {

struct Position { double x = 0, y = 0, z = 0; };
struct Direction { double x = 0, y = 0, z = 1; };

struct BankedParticle {  // particle.hpp:38-66
  Position r;
  Direction u;
  double E = 0, wgt = 0, wgt2 = 0;
  uint64_t parent_history_id = 0, parent_daughter_id = 0, family_id = 0;
  bool operator<(const BankedParticle& rhs) const {  // particle.hpp:59-65
    if (parent_history_id < rhs.parent_history_id) return true;
    if (parent_history_id == rhs.parent_history_id && parent_daughter_id < rhs.parent_daughter_id) return true;
    return false;
  }
};

class Particle {  // particle.hpp:68-243 (state that crosses the transport() boundary)
 public:
  Particle(Position r, Direction u, double E, double wgt, uint64_t id = 0) : r_(r), u_(u), E_(E), wgt_(wgt), history_id_(id), family_id_(id) {}
  const Position& r() const { return r_; }
  const Direction& u() const { return u_; }
  double E() const { return E_; }
  double wgt() const { return wgt_; }
  double wgt2() const { return wgt2_; }
  uint64_t history_id() const { return history_id_; }
  uint64_t family_id() const { return family_id_; }
  void set_weight2(double w) { wgt2_ = w; }
  void set_family_id(uint64_t f) { family_id_ = f; }
  // pcg32 state.  initialize_rng() == seed(seed); advance(stride*history_id) (particle.hpp:188-193) is
  // evaluated on the device; a particle that already consumed draws (source sampling) carries its state.
  bool has_rng_state = false;
  uint64_t rng_state = 0;

 private:
  Position r_;
  Direction u_;
  double E_, wgt_, wgt2_ = 0.;
  uint64_t history_id_, family_id_;
};

class Tallies {  // scalar generation scores + statistics; mesh tallies live on the device
 public:
  explicit Tallies(double tot_wgt) : total_weight(tot_wgt) {}
  void set_total_weight(double w) { total_weight = w; }
  void score_k_col(double s) { k_col_score += s; }
  void score_k_abs(double s) { k_abs_score += s; }
  void score_k_trk(double s) { k_trk_score += s; }
  void score_k_tot(double s) { k_tot_score += s; }
  void score_leak(double s) { leak_score += s; }
  void score_mig_area(double s) { mig_area_score += s; }
  void clear_generation();                     // tallies.cpp:143-157
  void calc_gen_values();                      // tallies.cpp:159-181
  void record_generation(double multiplier = 1.);  // tallies.cpp:183-203
  double kcol() const { return k_col; }
  double ktrk() const { return k_trk; }
  double kabs() const { return k_abs; }
  double leakage() const { return leak; }
  double mig_area() const { return mig; }
  double kcol_avg() const { return k_col_avg; }
  double kcol_err() const { return err(k_col_var); }
  double ktrk_avg() const { return k_trk_avg; }
  double ktrk_err() const { return err(k_trk_var); }
  double kabs_avg() const { return k_abs_avg; }
  double kabs_err() const { return err(k_abs_var); }
  double leakage_avg() const { return leak_avg; }
  double leakage_err() const { return err(leak_var); }
  double keff() const { return keff_; }
  void set_keff(double k) { keff_ = k; }
  void set_kcol(double k) { k_col = k; }
  int generations() const { return gen; }
  std::vector<double> k_col_vec, k_abs_vec, k_trk_vec, leak_vec, mig_vec;
  abl_handle backend = nullptr;  // mesh arrays: abl_tallies_record / abl_tallies_clear

 private:
  double err(double var) const { return gen > 0 ? std::sqrt(var / static_cast<double>(gen)) : 0.; }
  void update_avg_and_var(double x, double& x_avg, double& x_var) const;  // tallies.cpp:268-280
  double total_weight;
  double k_col_score = 0, k_abs_score = 0, k_trk_score = 0, leak_score = 0, k_tot_score = 0, mig_area_score = 0;
  double k_col = 1., k_col_avg = 0, k_col_var = 0, k_abs = 1., k_abs_avg = 0, k_abs_var = 0, k_trk = 1., k_trk_avg = 0,
         k_trk_var = 0, leak = 0, leak_avg = 0, leak_var = 0, k_tot = 1., k_tot_avg = 0, k_tot_var = 0, mig = 0, mig_avg = 0,
         mig_var = 0;
  double keff_ = 1.;
  int gen = 0;
};

class Transporter {  // include/simulation/transporter.hpp:39-47
 public:
  explicit Transporter(std::shared_ptr<Tallies> tallies) : tallies(std::move(tallies)) {}
  virtual ~Transporter() = default;
  virtual std::vector<BankedParticle> transport(std::vector<Particle>& bank, bool noise = false,
                                                std::vector<BankedParticle>* noise_bank = nullptr,
                                                const void* noise_maker = nullptr) = 0;

 protected:
  std::shared_ptr<Tallies> tallies;
};

struct Counters {
  uint64_t flights = 0, real_collisions = 0, virtual_collisions = 0, tl_bins = 0, fission_sites = 0, boundary_events = 0,
           lost_at_birth = 0, coll_scores = 0;
};

class GPUTransporter : public Transporter {
 public:
  GPUTransporter(std::shared_ptr<Tallies> tallies, const Problem& problem, int device);
  ~GPUTransporter() override;
  std::vector<BankedParticle> transport(std::vector<Particle>& bank, bool noise = false,
                                        std::vector<BankedParticle>* noise_bank = nullptr,
                                        const void* noise_maker = nullptr) override;
  abl_handle handle() const { return h_; }
  bool converged = false;  // settings::converged
  Counters counters;       // accumulated over calls

 private:
  abl_handle h_ = nullptr;
  std::vector<double> buf_[9], obuf_[9];
  std::vector<uint64_t> ida_, idb_, idc_, oa_, ob_, oc_;
};

// settings::rng (src/settings.cpp:59,116-119): pcg32 seeded with rng_seed on the default stream, then switched to stream 2 (the
// increment becomes (2 << 1) | 1, the state is kept).  The Simulation constructor re-initialises it (simulation.cpp:52), so a run
// starts from exactly this state.  It is the UniformRandomBitGenerator of the comb's std::shuffle calls.
struct GlobalRng {
  using result_type = uint32_t;
  static constexpr result_type min() { return 0u; }
  static constexpr result_type max() { return 0xffffffffu; }
  uint64_t state = 0, inc = 1442695040888963407ULL;
  void initialize(uint64_t seed) {
    state = (seed + 1442695040888963407ULL) * 6364136223846793005ULL + 1442695040888963407ULL;
    inc = (2ULL << 1) | 1ULL;
  }
  result_type operator()() {
    const uint64_t old = state;
    state = old * 6364136223846793005ULL + inc;
    const uint32_t xorshifted = static_cast<uint32_t>(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = static_cast<uint32_t>(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
  }
  double rand();  // RNG::rand (rng.hpp:41): libstdc++ generate_canonical<double, 53>
};

// BranchlessPowerIterator::comb_particles (src/branchless_power_iterator.cpp:592-651): a serial host step on the gathered bank in
// the reference too; std::shuffle of this libstdc++ on the global engine, so the combed bank is the reference's particle for particle.
void comb_particles(std::vector<BankedParticle>& next_gen, GlobalRng& rng);
// ... the same from the weights alone: the combed bank is row rows[k] of the bank with weight wgts[k], k = 0 .. rows.size() - 1
void comb_rows(const std::vector<double>& wgt, GlobalRng& rng, std::vector<uint32_t>& rows, std::vector<double>& wgts);

class PowerIterator {  // src/power_iterator.cpp (and src/branchless_power_iterator.cpp: the same loop plus the comb)
 public:
  PowerIterator(const Problem& problem, int device);
  ~PowerIterator();
  void initialize();  // sample the initial source (power_iterator.cpp:46-133)
  // Generation loop (power_iterator.cpp:305-473).  resident = false: the reference's data flow, the bank
  // crosses the Transporter::transport() boundary as host vectors every generation.  resident = true:
  // the bank never leaves HBM (device entry points of the C ABI).
  void run(int ngenerations, int nignored, bool resident);
  void write_results(const std::string& dir) const;  // tallies as .npy ([Ne,Nx,Ny,Nz], avg and std)

  std::shared_ptr<Tallies> tallies;
  std::shared_ptr<GPUTransporter> transporter;
  std::vector<double> entropy_vec;
  // settings: pair-distance-sqrd, families, empty-entropy-bins (src/power_iterator.cpp:283-297; the device-resident loop fills them)
  std::vector<double> r_sqrd_vec, families_vec, empty_entropy_frac_vec;
  // PowerIterator::out_of_time / check_time (src/power_iterator.cpp:715-749): true when less than two average generations of
  // settings: max-run-time are left; the loop then ends as if `gen` had been the last generation
  bool out_of_time(int gen, double loop_seconds) const;
  std::vector<uint64_t> nbank_vec;
  double seconds = 0., active_particles = 0.;
  const Problem& problem;

 private:
  struct DeviceBank;
  void alloc_device_bank(DeviceBank& b, uint64_t cap);
  void free_device_bank(DeviceBank& b);
  void run_host(int ngenerations, int nignored);
  void run_resident(int ngenerations, int nignored);
  double entropy_from_bins(const std::vector<double>& bins, double total) const;  // entropy.cpp:62-93
  std::vector<Particle> bank_;
  GlobalRng global_rng_;
  uint64_t histories_counter_ = 0, global_histories_counter_ = 0;
  bool initialized_ = false;
  int device_ = 0;
};

void write_npy(const std::string& path, const std::vector<double>& data, const std::vector<uint64_t>& shape);

}  // namespace abeille
