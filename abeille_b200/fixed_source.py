"""Fixed-source and modified-fixed-source simulations on the B200 backend.

`FixedSource` (src/fixed_source.cpp:81-175): every batch samples the source and transports it once; the fission neutrons of
a collision continue the history that made them as secondaries (src/transporter.cpp:374-379,460-463: n_new = floor(|w nu Sf /
St| + xi), Particle::make_secondary), so the transport call returns an empty bank.  The deck's `simulation: fixed-source`
reaches the kernels as ABL_MODE_FIXED_SOURCE; the per-lane kernel runs it (its secondaries LIFO is the one carter splitting
and noise copies use).

`ModifiedFixedSource`: the reference's `ModifiedFixedSource` driver
(src/modified_fixed_source.cpp:59-141) over the device entry points of the C ABI.  Banks never leave HBM.

Every batch samples `nparticles` source particles (Simulation::sample_sources, src/simulation.cpp:55-77) and follows the whole
fission chain of the batch: the fission bank a transport call returns becomes the bank of the next call -- weights kept, fresh
history ids, RNG streams re-seeded from the ids (Particle(p.r, p.u, p.E, p.wgt, p.wgt2, histories_counter++) +
initialize_rng) -- until a call returns an empty bank; then Tallies::calc_gen_values / record_generation and the mesh tallies'
record_generation.  The problem must be subcritical.  In this mode make_fission_neutrons does not divide the number of fission
neutrons by k_col (src/transporter.cpp:381-386); the k-eigenvalue kernels do the same when k_col is 1 (x / 1 is x exactly),
so no kernel is special to this driver.
"""
import numpy as np
import yaml

from .backend import Backend


class FixedSource:
    def __init__(self, deck_path: str, device: int = 0):
        with open(deck_path) as f:
            deck = yaml.safe_load(f)
        st = deck.get("settings", {})
        if st.get("simulation") != "fixed-source":
            raise ValueError("FixedSource needs a deck with `simulation: fixed-source`")
        self.gpu = Backend(deck_path, device)
        self.nparticles = int(st.get("nparticles", 100000))
        self.nbatches = int(st.get("ngenerations", 120))
        self.tally_names = [str(t.get("name", f"tally{i}")) for i, t in enumerate(deck.get("tallies", []) or [])]
        self.cur, self.nxt = self.gpu.new_device_bank(self.nparticles), self.gpu.new_device_bank(4096)
        self.history_counter = 0
        self.k_col, self.leakage, self.mig_area = [], [], []

    def batch(self):
        g, n = self.gpu, self.nparticles
        g.sample_source_device(self.cur, n, self.history_counter)
        self.history_counter += n  # (FixedSource::mpi_advance: the next batch starts nparticles further)
        m, scores, _ = g.transport_device(self.cur, n, self.nxt, k_col=1.0, converged=True, use_rng_state=True)
        if m:
            raise RuntimeError("Returned bank not empty on fixed-source transport.")
        self.k_col.append(scores[0] / n)
        self.leakage.append(scores[4] / n)
        self.mig_area.append(scores[5] / n)
        g.tallies_record(1.0)
        g.tallies_clear()

    def run(self):
        for _ in range(self.nbatches):
            self.batch()
        return {"kcol": np.array(self.k_col), "leak": np.array(self.leakage), "mig": np.array(self.mig_area)}

    def tally(self, t: int, which: str = "avg") -> np.ndarray:
        return self.gpu.tally(t, which)

    def close(self):
        self.gpu.close()


class ModifiedFixedSource:
    def __init__(self, deck_path: str, device: int = 0):
        with open(deck_path) as f:
            deck = yaml.safe_load(f)
        st = deck.get("settings", {})
        if st.get("simulation") != "modified-fixed-source":
            raise ValueError("ModifiedFixedSource needs a deck with `simulation: modified-fixed-source`")
        self.gpu = Backend(deck_path, device)
        self.nparticles = int(st.get("nparticles", 100000))
        self.nbatches = int(st.get("ngenerations", 120))
        self.tally_names = [str(t.get("name", f"tally{i}")) for i, t in enumerate(deck.get("tallies", []) or [])]
        self.cap = self.gpu.fission_capacity(self.nparticles, float(self.nparticles), 1.0) + 4 * self.nparticles
        self.cur, self.nxt = self.gpu.new_device_bank(self.cap), self.gpu.new_device_bank(self.cap)
        self.history_counter = 0
        self.k_col, self.leakage, self.mig_area, self.chain_generations, self.transported = [], [], [], [], 0

    def batch(self):
        g, n = self.gpu, self.nparticles
        g.sample_source_device(self.cur, n, self.history_counter)
        self.history_counter += n
        totals = np.zeros(6)
        m_in, use_state, gens = n, True, 0
        while m_in:
            m, scores, _ = g.transport_device(self.cur, m_in, self.nxt, k_col=1.0, converged=True, use_rng_state=use_state)
            use_state = False
            self.transported += m_in
            totals += scores
            gens += 1
            if m > self.cap:
                raise RuntimeError("modified-fixed-source: fission chain outgrew the bank (is the problem subcritical?)")
            if m:
                g.to_particles_device(self.nxt, m, self.history_counter)  # fresh history ids; the weights are kept
                self.history_counter += m
            self.cur, self.nxt = self.nxt, self.cur
            m_in = m
        self.k_col.append(totals[0] / n)        # Tallies::calc_gen_values: score / total_weight
        self.leakage.append(totals[4] / n)
        self.mig_area.append(totals[5] / n)
        self.chain_generations.append(gens)
        g.tallies_record(1.0)
        g.tallies_clear()

    def run(self):
        for _ in range(self.nbatches):
            self.batch()
        return {"kcol": np.array(self.k_col), "leak": np.array(self.leakage), "mig": np.array(self.mig_area),
                "chain_generations": list(self.chain_generations), "transported": self.transported}

    def tally(self, t: int, which: str = "avg") -> np.ndarray:
        return self.gpu.tally(t, which)

    def write_results(self, directory: str):
        import os
        os.makedirs(directory, exist_ok=True)
        np.save(os.path.join(directory, "leakage.npy"), np.array(self.leakage))
        for t, name in enumerate(self.tally_names):
            np.save(os.path.join(directory, f"{name}_avg.npy"), self.gpu.tally(t, "avg"))
            np.save(os.path.join(directory, f"{name}_std.npy"), self.gpu.tally(t, "std"))

    def close(self):
        self.gpu.close()
