/* capi.cpp -- C entry points of libabeille_host.so: the C++ host (YAML deck -> object model ->
 * GPUTransporter / PowerIterator) for callers without a C++ toolchain (the Python tests and bench.py
 * load it with ctypes).  Errors never cross as exceptions: functions return 0 / non-zero and
 * ablh_last_error() gives the text (the reference's fatal_error() exits the process instead,
 * src/error.cpp:36-45).
 */
#include <algorithm>
#include <cstring>
#include <functional>
#include <memory>
#include <string>

#include "simulation.hpp"

using namespace abeille;

namespace {
struct HostContext {
  Problem problem;
  std::unique_ptr<PowerIterator> sim;
  std::string error;
};
thread_local std::string g_open_error;
}  // namespace

extern "C" {

void* ablh_open(const char* yaml_path, int device, char* errbuf, int errlen) {
  try {
    auto ctx = std::make_unique<HostContext>();
    ctx->problem = Problem::from_yaml(yaml_lite::LoadFile(yaml_path));
    ctx->sim = std::make_unique<PowerIterator>(ctx->problem, device);
    return ctx.release();
  } catch (const std::exception& e) {
    g_open_error = e.what();
    if (errbuf && errlen > 0) {
      std::strncpy(errbuf, e.what(), static_cast<size_t>(errlen) - 1);
      errbuf[errlen - 1] = 0;
    }
    return nullptr;
  }
}

void ablh_close(void* c) { delete static_cast<HostContext*>(c); }

const char* ablh_last_error(void* c) { return c ? static_cast<HostContext*>(c)->error.c_str() : g_open_error.c_str(); }

abl_handle ablh_backend(void* c) { return static_cast<HostContext*>(c)->sim->transporter->handle(); }

/* Parses and flattens only (no device): used by the CPU-side tests of the host logic.
 * info[0..11] = ngroups, nparticles, ngenerations, nignored, ntallies, tracking, mode, nsurfaces, ncells,
 *               nuniverses, nmaterials, max stack depth */
int ablh_parse_only(const char* yaml_path, int64_t info[12], double* majorant, int majorant_cap, char* errbuf, int errlen) {
  try {
    Problem P = Problem::from_yaml(yaml_lite::LoadFile(yaml_path));
    FlatProblem F;
    P.flatten(F);
    info[0] = P.settings.ngroups; info[1] = P.settings.nparticles; info[2] = P.settings.ngenerations;
    info[3] = P.settings.nignored; info[4] = static_cast<int64_t>(P.tallies.size()); info[5] = P.settings.tracking;
    info[6] = P.settings.mode; info[7] = F.p.nsurfaces; info[8] = F.p.ncells; info[9] = F.p.nuniverses;
    info[10] = F.p.nmaterials; info[11] = P.max_stack_depth();
    for (int g = 0; g < P.settings.ngroups && g < majorant_cap; g++) majorant[g] = F.smp[static_cast<size_t>(g)];
    return 0;
  } catch (const std::exception& e) {
    if (errbuf && errlen > 0) {
      std::strncpy(errbuf, e.what(), static_cast<size_t>(errlen) - 1);
      errbuf[errlen - 1] = 0;
    }
    return 1;
  }
}

/* Dumps the flattened tables of a deck as text (one table per line) -- compared against the oracle's
 * independently built tables in tests/test_host.py. */
int ablh_dump_tables(const char* yaml_path, char* out, int64_t out_cap, char* errbuf, int errlen) {
  try {
    Problem P = Problem::from_yaml(yaml_lite::LoadFile(yaml_path));
    FlatProblem F;
    P.flatten(F);
    std::string s;
    char buf[64];
    auto dump = [&](const char* name, const std::vector<double>& v) {
      s += name;
      for (double x : v) {
        snprintf(buf, sizeof buf, " %.17g", x);
        s += buf;
      }
      s += "\n";
    };
    auto dumpi = [&](const char* name, const std::vector<int32_t>& v) {
      s += name;
      for (int32_t x : v) s += " " + std::to_string(x);
      s += "\n";
    };
    dump("Et", F.Et); dump("Ea", F.Ea); dump("Ef", F.Ef); dump("Es", F.Es); dump("nu", F.nu); dump("nud", F.nud);
    dump("chi_cdf", F.chi_cdf); dump("scatter_cdf", F.scatter_cdf); dump("amu", F.amu); dump("apdf", F.apdf); dump("acdf", F.acdf);
    dump("smp", F.smp);
    dumpi("rpn", F.rpn); dumpi("universe_cells", F.universe_cells); dumpi("lattice_tiles", F.lattice_tiles);
    std::vector<int32_t> ang;
    for (const auto& a : F.angle) { ang.push_back(a.offset); ang.push_back(a.n); }
    dumpi("angle", ang);
    // geometry records (compared with integration/flatten_problem.hpp, which builds them from the reference's live objects)
    std::vector<double> geo;
    for (const auto& sf : F.surfaces) { geo.push_back(sf.type); geo.push_back(sf.bc); for (double v : sf.p) geo.push_back(v); }
    dump("surfaces", geo);
    std::vector<int32_t> cl;
    for (const auto& c : F.cells) { cl.push_back(c.rpn_offset); cl.push_back(c.rpn_len); cl.push_back(c.simple); cl.push_back(c.vac_or_refl); cl.push_back(c.fill_universe); cl.push_back(c.material); }
    dumpi("cells", cl);
    std::vector<double> un;
    for (const auto& u : F.universes) {
      for (double v : {(double)u.type, (double)u.has_bc, (double)u.cell_offset, (double)u.ncells, (double)u.N[0], (double)u.N[1], (double)u.N[2],
                       (double)u.tile_offset, (double)u.outer, u.P[0], u.P[1], u.P[2], u.Pinv[0], u.Pinv[1], u.Pinv[2], u.Xl[0], u.Xl[1], u.Xl[2]})
        un.push_back(v);
    }
    dump("universes", un);
    dumpi("root", {F.p.root_universe});
    if (static_cast<int64_t>(s.size()) + 1 > out_cap) throw std::runtime_error("dump buffer too small");
    std::memcpy(out, s.c_str(), s.size() + 1);
    return 0;
  } catch (const std::exception& e) {
    if (errbuf && errlen > 0) {
      std::strncpy(errbuf, e.what(), static_cast<size_t>(errlen) - 1);
      errbuf[errlen - 1] = 0;
    }
    return 1;
  }
}

/* The flattened source records of a deck (host only): 20 doubles per source -- weight, fissile_only, is_box, low[3], hi[3], energy,
 * direction_kind, dir[3], cos_aperture, energy_kind, en_a, en_b.  Returns the number of sources, or -1 with the parser's message. */
int ablh_sources(const char* yaml_path, double* out, int64_t cap, char* errbuf, int errlen) {
  try {
    Problem P = Problem::from_yaml(yaml_lite::LoadFile(yaml_path));
    int64_t k = 0;
    for (const Source& src : P.sources) {
      const abl_source& f = src.flat;
      const double row[20] = {f.weight, (double)f.fissile_only, (double)f.is_box, f.low[0], f.low[1], f.low[2], f.hi[0], f.hi[1], f.hi[2],
                              f.energy, (double)f.direction_kind, f.dir[0], f.dir[1], f.dir[2], f.cos_aperture, (double)f.energy_kind,
                              f.en_a, f.en_b, 0., 0.};
      if (k + 20 > cap) throw std::runtime_error("source buffer too small");
      for (double v : row) out[k++] = v;
    }
    return static_cast<int>(P.sources.size());
  } catch (const std::exception& e) {
    if (errbuf && errlen > 0) {
      std::strncpy(errbuf, e.what(), static_cast<size_t>(errlen) - 1);
      errbuf[errlen - 1] = 0;
    }
    return -1;
  }
}

int ablh_info(void* c, int64_t info[12]) {
  HostContext* ctx = static_cast<HostContext*>(c);
  const Problem& P = ctx->problem;
  info[0] = P.settings.ngroups; info[1] = P.settings.nparticles; info[2] = P.settings.ngenerations;
  info[3] = P.settings.nignored; info[4] = static_cast<int64_t>(P.tallies.size()); info[5] = P.settings.tracking;
  info[6] = P.settings.mode; info[7] = static_cast<int64_t>(P.surfaces.size()); info[8] = static_cast<int64_t>(P.cells.size());
  info[9] = static_cast<int64_t>(P.universes.size()); info[10] = static_cast<int64_t>(P.materials.size());
  info[11] = P.max_stack_depth();
  return 0;
}

/* Transporter::transport through the C++ adapter (vector<Particle> in, vector<BankedParticle> out).
 * in: SoA of n particles (id_c = rng state or NULL); out: capacity out->n, *n_out the real count. */
int ablh_transport(void* c, const abl_bank* in, int converged, double k_col, abl_bank* out, uint64_t* n_out, double scores[6]) {
  HostContext* ctx = static_cast<HostContext*>(c);
  try {
    std::vector<Particle> bank;
    bank.reserve(in->n);
    for (uint64_t i = 0; i < in->n; i++) {
      Particle p(Position{in->x[i], in->y[i], in->z[i]}, Direction{in->ux[i], in->uy[i], in->uz[i]}, in->E[i], in->wgt[i], in->id_a[i]);
      if (in->id_b) p.set_family_id(in->id_b[i]);
      if (in->id_c) {
        p.has_rng_state = true;
        p.rng_state = in->id_c[i];
      }
      bank.push_back(p);
    }
    auto tallies = std::make_shared<Tallies>(static_cast<double>(ctx->problem.settings.nparticles));
    // score through a scratch Tallies so that the raw sums of this call can be returned
    Tallies& T = *ctx->sim->tallies;
    T.set_kcol(k_col);
    ctx->sim->transporter->converged = converged != 0;
    const size_t before = T.k_col_vec.size();
    (void)before;
    std::vector<BankedParticle> fis = ctx->sim->transporter->transport(bank);
    T.calc_gen_values();  // k = score / total_weight
    const double tw = static_cast<double>(ctx->problem.settings.nparticles);
    scores[0] = T.kcol() * tw; scores[1] = T.kabs() * tw; scores[2] = T.ktrk() * tw; scores[3] = 0.;
    scores[4] = T.leakage() * tw; scores[5] = T.mig_area() * tw;
    T.k_col_vec.pop_back(); T.k_abs_vec.pop_back(); T.k_trk_vec.pop_back(); T.leak_vec.pop_back(); T.mig_vec.pop_back();
    // zero the scalar scores without clearing the mesh arrays
    abl_handle keep = T.backend;
    T.backend = nullptr;
    T.clear_generation();
    T.backend = keep;
    *n_out = fis.size();
    if (fis.size() > out->n) throw std::runtime_error("ablh_transport: output capacity too small");
    for (size_t i = 0; i < fis.size(); i++) {
      out->x[i] = fis[i].r.x; out->y[i] = fis[i].r.y; out->z[i] = fis[i].r.z;
      out->ux[i] = fis[i].u.x; out->uy[i] = fis[i].u.y; out->uz[i] = fis[i].u.z;
      out->E[i] = fis[i].E; out->wgt[i] = fis[i].wgt;
      if (out->wgt2) out->wgt2[i] = fis[i].wgt2;
      out->id_a[i] = fis[i].parent_history_id; out->id_b[i] = fis[i].parent_daughter_id; out->id_c[i] = fis[i].family_id;
    }
    return bank.empty() ? 0 : 2;  // transport() must leave the input bank empty
  } catch (const std::exception& e) {
    ctx->error = e.what();
    return 1;
  }
}

/* Transporter::transport(bank, noise, &noise_bank, &noise_maker) through the C++ adapter (noise mode, src/noise.cpp:312-314,492):
 * noise != 0 transports noise particles (in->wgt2 is their second weight); sample_noise != 0 is a power-iteration generation
 * that also fills noise_out with the sampled noise source.  Output banks need all twelve arrays. */
int ablh_transport_noise(void* c, const abl_bank* in, int converged, double k_col, double keff, int noise, int sample_noise,
                         abl_bank* out, uint64_t* n_out, abl_bank* noise_out, uint64_t* n_noise) {
  HostContext* ctx = static_cast<HostContext*>(c);
  try {
    std::vector<Particle> bank;
    bank.reserve(in->n);
    for (uint64_t i = 0; i < in->n; i++) {
      Particle p(Position{in->x[i], in->y[i], in->z[i]}, Direction{in->ux[i], in->uy[i], in->uz[i]}, in->E[i], in->wgt[i], in->id_a[i]);
      if (in->wgt2) p.set_weight2(in->wgt2[i]);
      if (in->id_b) p.set_family_id(in->id_b[i]);
      if (in->id_c) {
        p.has_rng_state = true;
        p.rng_state = in->id_c[i];
      }
      bank.push_back(p);
    }
    Tallies& T = *ctx->sim->tallies;
    T.set_kcol(k_col);
    T.set_keff(keff);
    ctx->sim->transporter->converged = converged != 0;
    std::vector<BankedParticle> nb;
    const int maker = 1;  // stands for the NoiseMaker: the sources live in the flattened problem tables
    std::vector<BankedParticle> fis = ctx->sim->transporter->transport(bank, noise != 0, sample_noise ? &nb : nullptr, sample_noise ? &maker : nullptr);
    abl_handle keep = T.backend;  // zero the scalar scores without clearing the mesh arrays
    T.backend = nullptr;
    T.clear_generation();
    T.backend = keep;
    const auto put = [](const std::vector<BankedParticle>& v, abl_bank* o) {
      if (v.size() > o->n) throw std::runtime_error("ablh_transport_noise: output capacity too small");
      for (size_t i = 0; i < v.size(); i++) {
        o->x[i] = v[i].r.x; o->y[i] = v[i].r.y; o->z[i] = v[i].r.z;
        o->ux[i] = v[i].u.x; o->uy[i] = v[i].u.y; o->uz[i] = v[i].u.z;
        o->E[i] = v[i].E; o->wgt[i] = v[i].wgt; o->wgt2[i] = v[i].wgt2;
        o->id_a[i] = v[i].parent_history_id; o->id_b[i] = v[i].parent_daughter_id; o->id_c[i] = v[i].family_id;
      }
    };
    *n_out = fis.size();
    put(fis, out);
    *n_noise = nb.size();
    if (noise_out) put(nb, noise_out);
    return bank.empty() ? 0 : 2;
  } catch (const std::exception& e) {
    ctx->error = e.what();
    return 1;
  }
}

/* PowerIterator::run.  Per generation: kcol, ktrk, leak, mig, entropy, bank size.
 * summary[0..9] = kcol_avg, kcol_err, ktrk_avg, ktrk_err, leak_avg, leak_err, seconds, active particles,
 *                 real collisions, flights */
int ablh_run_power_iteration(void* c, int ngen, int nignored, int resident, double* kcol, double* ktrk, double* leak, double* mig,
                             double* entropy, uint64_t* nbank, double* summary) {
  HostContext* ctx = static_cast<HostContext*>(c);
  try {
    PowerIterator& S = *ctx->sim;
    const size_t g0 = S.tallies->k_col_vec.size();
    S.run(ngen, nignored, resident != 0);
    const Tallies& T = *S.tallies;
    // (settings: max-run-time may have ended the loop early: the generations that did not run keep the caller's zeros)
    const int ran = static_cast<int>(std::min<size_t>(static_cast<size_t>(ngen), T.k_col_vec.size() - g0));
    for (int g = 0; g < ran; g++) {
      const size_t k = g0 + static_cast<size_t>(g);
      kcol[g] = T.k_col_vec[k]; ktrk[g] = T.k_trk_vec[k]; leak[g] = T.leak_vec[k]; mig[g] = T.mig_vec[k];
      entropy[g] = S.entropy_vec[k]; nbank[g] = S.nbank_vec[k];
    }
    summary[0] = T.kcol_avg(); summary[1] = T.kcol_err(); summary[2] = T.ktrk_avg(); summary[3] = T.ktrk_err();
    summary[4] = T.leakage_avg(); summary[5] = T.leakage_err(); summary[6] = S.seconds; summary[7] = S.active_particles;
    summary[8] = static_cast<double>(S.transporter->counters.real_collisions);
    summary[9] = static_cast<double>(S.transporter->counters.flights);
    return 0;
  } catch (const std::exception& e) {
    ctx->error = e.what();
    return 1;
  }
}

int ablh_write_results(void* c, const char* dir) {
  HostContext* ctx = static_cast<HostContext*>(c);
  try {
    ctx->sim->write_results(dir);
    return 0;
  } catch (const std::exception& e) {
    ctx->error = e.what();
    return 1;
  }
}

/* BranchlessPowerIterator::comb_particles on host columns (no device involved): in->n particles in, at most out->n (capacity)
 * written, *n_out = the combed population; rng2 = {state, increment} of settings::rng, updated. */
int ablh_comb_particles(const abl_bank* in, abl_bank* out, uint64_t* n_out, uint64_t rng2[2]) {
  try {
    std::vector<abeille::BankedParticle> v(in->n);
    for (uint64_t i = 0; i < in->n; i++) {
      v[i].r = abeille::Position{in->x[i], in->y[i], in->z[i]};
      v[i].u = abeille::Direction{in->ux[i], in->uy[i], in->uz[i]};
      v[i].E = in->E[i]; v[i].wgt = in->wgt[i]; v[i].wgt2 = in->wgt2 ? in->wgt2[i] : 0.;
      v[i].parent_history_id = in->id_a[i]; v[i].parent_daughter_id = in->id_b[i]; v[i].family_id = in->id_c[i];
    }
    abeille::GlobalRng rng;
    rng.state = rng2[0];
    rng.inc = rng2[1];
    abeille::comb_particles(v, rng);
    rng2[0] = rng.state;
    rng2[1] = rng.inc;
    *n_out = v.size();
    for (uint64_t i = 0; i < v.size() && i < out->n; i++) {
      out->x[i] = v[i].r.x; out->y[i] = v[i].r.y; out->z[i] = v[i].r.z;
      out->ux[i] = v[i].u.x; out->uy[i] = v[i].u.y; out->uz[i] = v[i].u.z;
      out->E[i] = v[i].E; out->wgt[i] = v[i].wgt;
      if (out->wgt2) out->wgt2[i] = v[i].wgt2;
      out->id_a[i] = v[i].parent_history_id; out->id_b[i] = v[i].parent_daughter_id; out->id_c[i] = v[i].family_id;
    }
    return v.size() > out->n ? 2 : 0;
  } catch (const std::exception& e) {
    g_open_error = e.what();
    return 1;
  }
}

/* ... from the weights alone (abeille::comb_rows): the combed bank is row rows[k] of the bank with weight wgts[k]; at most cap
 * entries are written, *n_out = the combed population */
int ablh_comb_rows(const double* wgt, uint64_t n, uint64_t rng2[2], uint32_t* rows, double* wgts, uint64_t cap, uint64_t* n_out) {
  try {
    abeille::GlobalRng rng;
    rng.state = rng2[0];
    rng.inc = rng2[1];
    std::vector<uint32_t> r;
    std::vector<double> w;
    abeille::comb_rows(std::vector<double>(wgt, wgt + n), rng, r, w);
    rng2[0] = rng.state;
    rng2[1] = rng.inc;
    *n_out = r.size();
    for (uint64_t i = 0; i < r.size() && i < cap; i++) { rows[i] = r[i]; wgts[i] = w[i]; }
    return r.size() > cap ? 2 : 0;
  } catch (const std::exception& e) {
    g_open_error = e.what();
    return 1;
  }
}

/* yaml_lite self-test hook: parses text, returns a canonical one-line rendering */
int ablh_yaml_roundtrip(const char* text, char* out, int64_t out_cap) {
  try {
    const yaml_lite::Node n = yaml_lite::Load(std::string(text));
    std::string s;
    std::function<void(const yaml_lite::Node&)> emit = [&](const yaml_lite::Node& x) {
      switch (x.type) {
        case yaml_lite::Node::Null: s += "null"; break;
        case yaml_lite::Node::Scalar: s += "\"" + x.scalar + "\""; break;
        case yaml_lite::Node::Sequence:
          s += "[";
          for (size_t i = 0; i < x.seq.size(); i++) {
            if (i) s += ",";
            emit(x.seq[i]);
          }
          s += "]";
          break;
        case yaml_lite::Node::Map:
          s += "{";
          for (size_t i = 0; i < x.map.size(); i++) {
            if (i) s += ",";
            s += "\"" + x.map[i].first + "\":";
            emit(x.map[i].second);
          }
          s += "}";
          break;
      }
    };
    emit(n);
    if (static_cast<int64_t>(s.size()) + 1 > out_cap) return 2;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return 0;
  } catch (const std::exception& e) {
    std::strncpy(out, e.what(), static_cast<size_t>(out_cap) - 1);
    out[out_cap - 1] = 0;
    return 1;
  }
}

}  // extern "C"
