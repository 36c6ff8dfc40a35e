/* distributed_main.cpp -- the multi-GPU k-eigenvalue generation loop in C++, NCCL called directly (one process per GPU).
 *
 * What abeille_b200/distributed.py does over torch.distributed, for a host that has no Python: PowerIterator::run
 * (src/power_iterator.cpp:305-473) with the bank sharded by global history id -- rank r of W owns the contiguous slice
 * [r N/W, (r+1) N/W) of the source and, from then on, the fission sites its histories made (the global bank is the concatenation
 * of the slices in rank order, which is the reference's bank order).  Nothing of the bank crosses NVLink; per generation the ranks
 * exchange
 *   - one all-gather of 12 doubles per rank (site count, the six generation scores, the four weight sums),
 *   - one all-reduce of the entropy bins (when the deck has an entropy mesh),
 *   - the dense bins of the approximate cancelator (all-reduce of four double arrays and the counts), when cancellation is on,
 *   - one all-reduce of every mesh tally's generation array in active generations,
 * and every rank ends a generation with the same k, the same statistics and its slice of the next bank with the global history
 * ids the reference would hand out.  Integer outcomes do not depend on W (the histories are the same histories); floating-point
 * sums are taken in rank order.  The slices are not rebalanced here (the Python driver does that when they drift).
 *
 * Usage (the launcher starts one per GPU):
 *   abl_pi_nccl <deck.yaml> <rank> <world> <id-file> <ngenerations> <nignored> [device = rank]
 * rank 0 writes the ncclUniqueId to <id-file>; the others wait for it.  Rank 0 prints one JSON line with the k series, bank sizes
 * and entropy.  The deck's nparticles is the GLOBAL population and must be divisible by <world>.
 */
#include <cuda_runtime.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "simulation.hpp"

using namespace abeille;

namespace {
[[noreturn]] void die(const std::string& m) {
  std::fprintf(stderr, "abl_pi_nccl: %s\n", m.c_str());
  std::exit(1);
}
void ck(abl_handle h, int rc, const char* what) {
  if (rc != ABL_OK) die(std::string(what) + ": " + abl_last_error(h));
}
#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) die(std::string(#call) + ": " + cudaGetErrorString(e_));    \
  } while (0)
#define NC(call)                                                                       \
  do {                                                                                 \
    ncclResult_t r_ = (call);                                                          \
    if (r_ != ncclSuccess) die(std::string(#call) + ": " + ncclGetErrorString(r_));    \
  } while (0)

struct DevBank {
  abl_bank b{};
  uint64_t cap = 0;
};
}  // namespace

int main(int argc, char** argv) {
  if (argc < 7) die("usage: abl_pi_nccl <deck.yaml> <rank> <world> <id-file> <ngenerations> <nignored> [device]");
  const std::string deck = argv[1], id_file = argv[4];
  const int rank = std::atoi(argv[2]), world = std::atoi(argv[3]), ngen = std::atoi(argv[5]), nign = std::atoi(argv[6]);
  const int device = argc > 7 ? std::atoi(argv[7]) : rank;
  if (world < 1 || rank < 0 || rank >= world) die("bad rank / world");
  try {
    CU(cudaSetDevice(device));
    // ---- communicator -----------------------------------------------------------------------------------------------------
    ncclUniqueId id;
    if (rank == 0) {
      NC(ncclGetUniqueId(&id));
      std::ofstream f(id_file + ".tmp", std::ios::binary);
      f.write(reinterpret_cast<const char*>(&id), sizeof id);
      f.close();
      std::rename((id_file + ".tmp").c_str(), id_file.c_str());
    } else {
      for (int tries = 0;; tries++) {
        std::ifstream f(id_file, std::ios::binary);
        if (f && f.read(reinterpret_cast<char*>(&id), sizeof id)) break;
        if (tries > 600) die("no ncclUniqueId in " + id_file);
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
      }
    }
    ncclComm_t comm;
    NC(ncclCommInitRank(&comm, world, id, rank));
    cudaStream_t stream;
    CU(cudaStreamCreate(&stream));

    // ---- problem ----------------------------------------------------------------------------------------------------------
    const Problem problem = Problem::from_yaml(yaml_lite::LoadFile(deck));
    const Settings& st = problem.settings;
    if (st.mode != ABL_MODE_K_EIGENVALUE || st.fixed_source) die("k-eigenvalue decks only");
    if (st.nparticles % world != 0) die("nparticles must be divisible by the number of ranks");
    const uint64_t n_total = static_cast<uint64_t>(st.nparticles), n_local = n_total / static_cast<uint64_t>(world);
    auto tallies = std::make_shared<Tallies>(static_cast<double>(n_total));
    GPUTransporter transporter(tallies, problem, device);
    abl_handle h = transporter.handle();
    const bool cancel = st.regional_cancellation && problem.cancelator.present;
    if (cancel && problem.cancelator.kind != ABL_CANCEL_APPROXIMATE && problem.cancelator.kind != 0)
      die("the sharded loop runs the approximate cancelator (exact cancelators work on the gathered bank: one GPU)");
    const bool have_entropy = problem.entropy.present;
    const size_t nebins = have_entropy ? static_cast<size_t>(problem.entropy.N[0]) * problem.entropy.N[1] * problem.entropy.N[2] : 0;

    auto alloc_bank = [&](DevBank& d, uint64_t cap) {
      ck(h, abl_bank_alloc_device(h, cap, &d.b), "abl_bank_alloc_device");
      d.cap = cap;
    };
    DevBank cur, nxt;
    uint64_t cap = abl_fission_capacity_hint(h, n_local, 0., 1.0);
    alloc_bank(cur, cap);
    alloc_bank(nxt, cap);
    // scratch: [world][12] gather buffer, entropy bins
    double *gather_d = nullptr, *mine_d = nullptr, *ebins_d = nullptr;
    CU(cudaMalloc(&gather_d, sizeof(double) * 12 * static_cast<size_t>(world)));
    CU(cudaMalloc(&mine_d, sizeof(double) * 12));
    if (have_entropy) CU(cudaMalloc(&ebins_d, sizeof(double) * (nebins + 1)));
    std::vector<double> gathered(12 * static_cast<size_t>(world)), ebins(nebins + 1);

    // Simulation::sample_sources: rank r samples the ids [r n, (r+1) n)
    ck(h, abl_sample_source_device(h, n_local, static_cast<uint64_t>(rank) * n_local, &cur.b, nullptr), "abl_sample_source_device");
    uint64_t n_cur = n_local, global_counter = n_total;
    bool use_state = true, converged = (nign == 0);
    std::vector<double> kcol_series, entropy_series;
    std::vector<uint64_t> nbank_series;
    CU(cudaDeviceSynchronize());
    // (the first collective of a communicator sets its channels up: done here, outside the timed loop, and it lines the ranks up)
    NC(ncclAllGather(mine_d, gather_d, 12, ncclDouble, comm, stream));
    CU(cudaStreamSynchronize(stream));
    const auto t0 = std::chrono::steady_clock::now();
    double active_particles = 0.;

    for (int g = 1; g <= ngen; g++) {
      abl_gen_params gp{};
      gp.k_col = tallies->kcol();
      gp.keff = tallies->keff();
      gp.converged = converged ? 1 : 0;
      abl_bank in = cur.b;
      in.n = n_cur;
      if (!use_state) in.id_c = nullptr;  // streams from the history ids (particle.hpp:188-193)
      uint64_t m = 0, cn[8];
      double scores[6];
      for (int attempt = 0;; attempt++) {
        abl_bank out = nxt.b;
        out.n = nxt.cap;
        const int rc = abl_transport_device(h, &in, &gp, &out, &m, scores, cn, nullptr);
        if (rc == ABL_ERR_BANK_OVERFLOW && m > nxt.cap && attempt < 2) {
          abl_bank_free_device(h, &nxt.b);
          alloc_bank(nxt, m + m / 8 + 4096);
          if (converged) ck(h, abl_tallies_clear(h), "abl_tallies_clear");
          continue;
        }
        ck(h, rc, "abl_transport_device");
        break;
      }
      abl_bank fis = nxt.b;
      fis.n = m;
      // entropy of the un-normalised fission bank (power_iterator.cpp:341-353)
      double entropy = 0.;
      if (have_entropy) {
        CU(cudaMemsetAsync(ebins_d, 0, sizeof(double) * (nebins + 1), stream));
        CU(cudaStreamSynchronize(stream));
        ck(h, abl_entropy_bin_device(h, &fis, ebins_d, ebins_d + nebins, nullptr), "abl_entropy_bin_device");
        CU(cudaDeviceSynchronize());
        NC(ncclAllReduce(ebins_d, ebins_d, nebins + 1, ncclDouble, ncclSum, comm, stream));
        CU(cudaMemcpyAsync(ebins.data(), ebins_d, sizeof(double) * (nebins + 1), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        const double total = ebins[nebins];
        for (size_t b = 0; b < nebins; b++) {  // Entropy::calculate_entropy (entropy.cpp:62-93)
          const double p = std::fabs(ebins[b]) / total;
          if (p != 0. && p <= 1.0) entropy -= p * std::log2(p);
        }
      }
      // regional cancellation: the dense bins are summed over the ranks, every rank applies them to its slice
      if (cancel) {
        ck(h, abl_cancel_accumulate_device(h, &fis, nullptr), "abl_cancel_accumulate_device");
        double* sums[4];
        uint32_t* count = nullptr;
        uint64_t nb = 0;
        ck(h, abl_cancel_bins_device(h, sums, &count, &nb), "abl_cancel_bins_device");
        CU(cudaDeviceSynchronize());
        NC(ncclGroupStart());
        for (double* sp : sums) NC(ncclAllReduce(sp, sp, nb, ncclDouble, ncclSum, comm, stream));
        NC(ncclAllReduce(count, count, nb, ncclUint32, ncclSum, comm, stream));
        NC(ncclGroupEnd());
        CU(cudaStreamSynchronize(stream));
        ck(h, abl_cancel_apply_device(h, &fis, nullptr), "abl_cancel_apply_device");
        // (apply re-zeroes the bins this rank's particles touched; the others still hold the global sums)
        for (double* sp : sums) CU(cudaMemsetAsync(sp, 0, sizeof(double) * nb, stream));
        CU(cudaMemsetAsync(count, 0, sizeof(uint32_t) * nb, stream));
        CU(cudaStreamSynchronize(stream));
      }
      // counts, scores and weight sums of every rank
      double ws[4];
      ck(h, abl_bank_weight_stats_device(h, &fis, ws, nullptr), "abl_bank_weight_stats_device");
      double mine[12] = {static_cast<double>(m), scores[0], scores[1], scores[2], scores[3], scores[4], scores[5], ws[0], ws[1], ws[2], ws[3],
                         static_cast<double>(n_cur)};
      CU(cudaMemcpyAsync(mine_d, mine, sizeof mine, cudaMemcpyHostToDevice, stream));
      NC(ncclAllGather(mine_d, gather_d, 12, ncclDouble, comm, stream));
      CU(cudaMemcpyAsync(gathered.data(), gather_d, sizeof(double) * 12 * static_cast<size_t>(world), cudaMemcpyDeviceToHost, stream));
      CU(cudaStreamSynchronize(stream));
      uint64_t m_total = 0, before_me = 0, n_in_total = 0;
      double sc[6] = {0, 0, 0, 0, 0, 0}, wpos = 0., wneg = 0.;
      for (int r = 0; r < world; r++) {
        const double* row = &gathered[12 * static_cast<size_t>(r)];
        if (r < rank) before_me += static_cast<uint64_t>(row[0]);
        m_total += static_cast<uint64_t>(row[0]);
        for (int q = 0; q < 6; q++) sc[q] += row[1 + q];
        wpos += row[9];
        wneg += row[10];
        n_in_total += static_cast<uint64_t>(row[11]);
      }
      if (m_total == 0) die("No fission neutrons were produced.");
      if (converged) active_particles += static_cast<double>(n_in_total);
      nbank_series.push_back(n_in_total);
      tallies->score_k_col(sc[0]); tallies->score_k_abs(sc[1]); tallies->score_k_trk(sc[2]);
      tallies->score_k_tot(sc[3]); tallies->score_leak(sc[4]); tallies->score_mig_area(sc[5]);
      tallies->calc_gen_values();
      // normalize_weights over the global bank (power_iterator.cpp:538-586)
      ck(h, abl_bank_scale_weights_device(h, &fis, static_cast<double>(n_total) / (wpos - wneg), nullptr), "abl_bank_scale_weights_device");
      if (converged) {
        ck(h, abl_score_source_device(h, &fis, 0, nullptr), "abl_score_source_device");
        if (world > 1) {
          CU(cudaDeviceSynchronize());
          for (int t = 0; t < abl_tally_count(h); t++) {
            double* gen = nullptr;
            uint64_t nbins = 0;
            ck(h, abl_tally_device_ptr(h, t, 0, &gen, &nbins), "abl_tally_device_ptr");
            NC(ncclAllReduce(gen, gen, nbins, ncclDouble, ncclSum, comm, stream));
          }
          CU(cudaStreamSynchronize(stream));
        }
        tallies->record_generation();
      }
      tallies->clear_generation();
      kcol_series.push_back(tallies->kcol());
      entropy_series.push_back(entropy);
      // fresh global history ids: this rank's slice starts after the slices of the ranks before it
      ck(h, abl_bank_to_particles_device(h, &fis, global_counter + before_me, nullptr), "abl_bank_to_particles_device");
      global_counter += m_total;
      std::swap(cur, nxt);
      n_cur = m;
      use_state = false;
      const uint64_t want = abl_fission_capacity_hint(h, m, static_cast<double>(n_total) / world, tallies->kcol());
      if (want > nxt.cap) {
        abl_bank_free_device(h, &nxt.b);
        alloc_bank(nxt, want + want / 8);
      }
      if (g == nign) converged = true;
    }
    CU(cudaDeviceSynchronize());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rank == 0) {
      std::printf("{\"world\": %d, \"nparticles\": %llu, \"seconds\": %.6f, \"active_particles\": %.0f, \"kcol_avg\": %.17g, \"kcol_err\": %.17g, \"kcol\": [",
                  world, static_cast<unsigned long long>(n_total), secs, active_particles, tallies->kcol_avg(),
                  tallies->generations() > 0 ? tallies->kcol_err() : 0.);
      for (size_t i = 0; i < kcol_series.size(); i++) std::printf("%s%.17g", i ? ", " : "", kcol_series[i]);
      std::printf("], \"nbank\": [");
      for (size_t i = 0; i < nbank_series.size(); i++) std::printf("%s%llu", i ? ", " : "", static_cast<unsigned long long>(nbank_series[i]));
      std::printf("], \"entropy\": [");
      for (size_t i = 0; i < entropy_series.size(); i++) std::printf("%s%.17g", i ? ", " : "", entropy_series[i]);
      std::printf("]}\n");
      std::fflush(stdout);
    }
    abl_bank_free_device(h, &cur.b);
    abl_bank_free_device(h, &nxt.b);
    NC(ncclCommDestroy(comm));
    return 0;
  } catch (const std::exception& e) {
    die(e.what());
  }
}
