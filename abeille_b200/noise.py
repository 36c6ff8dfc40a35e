"""Frequency-domain neutron-noise simulation on the B200 backend: the reference's `Noise` driver
(src/noise.cpp:211-559) over the device entry points of the C ABI.  Banks never leave HBM.

    Noise::run               noise.cpp:211-303   nignored power-iteration generations, then ngenerations noise batches
                                                 of (nskip - 1) plain generations + one that samples the noise source
    Noise::power_iteration   noise.cpp:305-372   transport -> k values -> (regional cancellation) -> normalise -> new bank
    Noise::noise_simulation  noise.cpp:425-559   normalise the noise source by its mean |w|, score it, transport the noise
                                                 particles generation by generation (inner generations) until none is
                                                 left, record the tallies scaled back by the mean |w|
Settings that only this driver reads (nskip, keff, noise-cancellation ...) come straight from the deck with the
reference's defaults (src/settings.cpp:37-82); everything the kernels need went through the C++ host already.
"""
import numpy as np
import yaml

from .backend import Backend

_INT_MAX = 2147483647


class NoiseSimulation:
    def __init__(self, deck_path: str, device: int = 0):
        with open(deck_path) as f:
            deck = yaml.safe_load(f)
        st = deck.get("settings", {})
        self.tally_names = [str(t.get("name", f"tally{i}")) for i, t in enumerate(deck.get("tallies", []) or [])]
        if st.get("simulation") != "noise":
            raise ValueError("NoiseSimulation needs a deck with `simulation: noise`")
        self.gpu = Backend(deck_path, device)
        self.nparticles = int(st.get("nparticles", 100000))
        self.nbatches = int(st.get("ngenerations", 120))
        self.nignored = int(st.get("nignored", 20))
        self.nskip = int(st.get("nskip", 10))
        self.keff = float(st.get("keff", 1.0))
        self.cancel_pi = bool(st.get("cancellation", False))
        self.cancel_noise = bool(st.get("noise-cancellation", False))
        self.n_cancel_noise_gens = int(st.get("cancel-noise-gens", _INT_MAX))
        self.normalize_noise_source = bool(st.get("normalize-noise-source", True))
        n = self.nparticles
        self.cap = int(3.0 * n) + 4096
        self.noise_cap = int(6.0 * n) + 4096
        g = self.gpu
        self.bank, self.next = g.new_device_bank(self.cap), g.new_device_bank(self.cap)
        self.noise_bank = g.new_device_bank(self.noise_cap)
        self.nb_a, self.nb_b = g.new_device_bank(self.noise_cap), g.new_device_bank(self.noise_cap)
        self.n_bank = 0
        self.k_col = 1.0
        self.history_counter = 0
        self.pi_gen = 0
        self.converged = False
        self.k_history, self.noise_generations, self.noise_particles = [], [], []

    # ---- PowerIterator-like initial source (noise.cpp:75-133 does what power_iterator.cpp:46-133 does) ----
    def initialize(self):
        self.gpu.sample_source_device(self.bank, self.nparticles, 0)
        self.n_bank = self.nparticles
        self.history_counter = self.nparticles
        self._use_state = True  # the first generation continues the source-sampling streams (simulation.cpp:70-73)

    def power_iteration(self, sample_noise: bool) -> int:
        """One generation (noise.cpp:305-372).  Returns the number of noise particles sampled."""
        g = self.gpu
        m, n_noise, scores, _ = g.transport_noise_device(
            self.bank, self.n_bank, self.next, self.noise_bank if sample_noise else None, k_col=self.k_col, keff=self.keff,
            converged=self.converged, noise=False, sample_noise=sample_noise, use_rng_state=self._use_state)
        self._use_state = False
        if m == 0:
            raise RuntimeError("No fission neutrons were produced.")
        self.k_col = scores[0] / self.nparticles  # tallies->calc_gen_values(): score / total_weight
        g.tallies_clear()                         # tallies->clear_generation()
        if self.cancel_pi:
            g.cancel_device(self.next, m)
        ws = g.weight_stats_device(self.next, m)  # Noise::normalize_weights (noise.cpp:631-674)
        g.scale_weights_device(self.next, m, self.nparticles / (ws[2] - ws[3]))
        g.to_particles_device(self.next, m, self.history_counter)
        self.history_counter += m
        self.bank, self.next = self.next, self.bank
        self.n_bank = m
        self.k_history.append(self.k_col)
        return n_noise

    def noise_simulation(self, n_noise: int):
        """Noise::noise_simulation (noise.cpp:425-559) on the noise bank sampled by the last generation."""
        g = self.gpu
        original_kcol = self.k_col
        avg_wgt_mag = 1.0
        if self.normalize_noise_source and n_noise:
            avg_wgt_mag = g.weight_magnitude_device(self.noise_bank, n_noise) / float(n_noise)
            g.divide_weights_device(self.noise_bank, n_noise, avg_wgt_mag)
        if self.converged and n_noise:
            g.score_source_device(self.noise_bank, n_noise, noise_source=True)
        cur, nxt = self.noise_bank, self.nb_a
        g.to_particles_device(cur, n_noise, self.history_counter)
        self.history_counter += n_noise
        n, gen, total = n_noise, 0, 0
        while n != 0:
            gen += 1
            total += n
            m, _, _, _ = g.transport_noise_device(cur, n, nxt, None, k_col=self.k_col, keff=self.keff,
                                                  converged=self.converged, noise=True, sample_noise=False)
            if self.cancel_noise and gen <= self.n_cancel_noise_gens and m:
                g.cancel_device(nxt, m)
            g.to_particles_device(nxt, m, self.history_counter)
            self.history_counter += m
            cur, nxt = nxt, (self.nb_b if nxt is self.nb_a else self.nb_a)
            n = m
        g.tallies_record(avg_wgt_mag)  # tallies->record_generation(avg_wgt_mag); calc_gen_values' k is discarded
        g.tallies_clear()
        self.k_col = original_kcol     # tallies->set_kcol(original_kcol)
        self.noise_generations.append(gen)
        self.noise_particles.append(total)

    def run(self):
        """Noise::run (noise.cpp:211-303)."""
        self.initialize()
        self.converged = False
        for _ in range(self.nignored):
            self.pi_gen += 1
            self.power_iteration(False)
        self.converged = True
        for _ in range(self.nbatches):
            for _ in range(self.nskip - 1):
                self.pi_gen += 1
                self.power_iteration(False)
            self.pi_gen += 1
            n_noise = self.power_iteration(True)
            self.noise_simulation(n_noise)
        return {"k_col": np.array(self.k_history), "noise_generations": self.noise_generations,
                "noise_particles": self.noise_particles}

    def tally(self, t: int, which: str = "avg") -> np.ndarray:
        return self.gpu.tally(t, which)

    def write_results(self, directory: str):
        """Tallies as .npy, [Ne,Nx,Ny,Nz] mean and standard deviation over the noise batches (MeshTally::write_tally,
        src/mesh_tally.cpp:154-206: std = sqrt(var / g)), named <tally>_avg.npy / <tally>_std.npy like the k-eigenvalue
        writer (host/simulation.cpp), plus the k_col series of the power-iteration generations."""
        import os
        os.makedirs(directory, exist_ok=True)
        np.save(os.path.join(directory, "kcol.npy"), np.array(self.k_history))
        for t, name in enumerate(self.tally_names):
            np.save(os.path.join(directory, f"{name}_avg.npy"), self.gpu.tally(t, "avg"))
            np.save(os.path.join(directory, f"{name}_std.npy"), self.gpu.tally(t, "std"))

    def close(self):
        self.gpu.close()


class DistributedNoiseSimulation(NoiseSimulation):
    """The noise simulation sharded over the ranks of a torch.distributed group, one process per GPU (the reference shards
    it over MPI ranks the same way: src/noise.cpp:312-372,425-559 with `mpi::Allreduce_sum` / `sync_banks`,
    src/simulation.cpp:84-111).  Every bank -- the power-iteration bank, the noise-source bank, the noise fission banks of
    the inner generations -- is split into contiguous slices in rank order, and history ids are handed out by an exclusive
    scan over the ranks' counts, so the global order and every history's RNG stream are those of a one-rank run: the integer
    outcomes (bank sizes, noise generations per batch, particles per generation) do not depend on the number of GPUs; sums
    (k, tallies, cancellation bins) differ in rounding order only.  Per generation the ranks exchange one small vector
    (all_gather), with regional cancellation the dense cancellation bins (all_reduce), and once per noise batch the tally
    arrays (all_reduce before MeshTally::record_generation)."""

    def __init__(self, deck_path: str, device: int = 0, group=None):
        import torch
        import torch.distributed as dist
        super().__init__(deck_path, device)
        self._torch, self._dist = torch, dist
        self.group = group
        self.use_dist = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.use_dist else 0
        self.world = dist.get_world_size(group) if self.use_dist else 1
        self.tdev = torch.device("cuda", device)
        self._cancel_views = None
        self._tally_views = None
        self._noise_counts = [0] * self.world

    # ---- collectives ----
    def _gather(self, values) -> np.ndarray:
        from .distributed import gather_vector
        return gather_vector(np.asarray(values, dtype=np.float64), self.world, self.tdev, self.group)

    def _views(self):
        from .distributed import _DevPtr
        if self._tally_views is None:
            self._tally_views = []
            for t in range(self.gpu.ntallies()):
                ptr, n = self.gpu.tally_device_ptr(t, 0)
                self._tally_views.append(self._torch.as_tensor(_DevPtr(ptr, n), device=self.tdev))
        return self._tally_views

    def _cancel(self, bank, m: int):
        """Regional cancellation over the GLOBAL bank (noise.cpp:387,449,649-653): dense bins, one all_reduce per bin array."""
        from .distributed import _DevPtr
        gpu, torch = self.gpu, self._torch
        if self.world == 1:
            if m:
                gpu.cancel_device(bank, m)
            return
        if m:
            gpu.cancel_accumulate_device(bank, m)
        if self._cancel_views is None:
            sums, count, nb = gpu.cancel_bins_device()
            self._cancel_views = [torch.as_tensor(_DevPtr(p, nb), device=self.tdev) for p in sums]
            self._cancel_views.append(torch.as_tensor(_DevPtr(count, nb, "<i4"), device=self.tdev))
        torch.cuda.current_stream().synchronize()
        for t in self._cancel_views:
            self._dist.all_reduce(t, group=self.group)
        if m:
            gpu.cancel_apply_device(bank, m)
        for t in self._cancel_views:  # bins only other ranks touched still hold the global sums
            t.zero_()

    def _reduce_tallies(self):
        if self.world > 1:
            self._torch.cuda.current_stream().synchronize()
            for t in self._views():
                self._dist.all_reduce(t, group=self.group)

    # ---- the driver, rank-local slices ----
    def initialize(self):
        from .distributed import even_split
        counts, bounds = even_split(self.nparticles, self.world)
        self.gpu.sample_source_device(self.bank, counts[self.rank], int(bounds[self.rank]))
        self.n_bank = counts[self.rank]
        self.history_counter = self.nparticles
        self._use_state = True

    def power_iteration(self, sample_noise: bool) -> int:
        g = self.gpu
        m = n_noise = 0
        scores = np.zeros(6)
        if self.n_bank:
            m, n_noise, scores, _ = g.transport_noise_device(
                self.bank, self.n_bank, self.next, self.noise_bank if sample_noise else None, k_col=self.k_col, keff=self.keff,
                converged=self.converged, noise=False, sample_noise=sample_noise, use_rng_state=self._use_state)
        self._use_state = False
        allv = self._gather([scores[0], float(m), float(n_noise)])
        tot = allv.sum(axis=0)
        counts = [int(v) for v in allv[:, 1]]
        if sum(counts) == 0:
            raise RuntimeError("No fission neutrons were produced.")
        self.k_col = tot[0] / self.nparticles
        g.tallies_clear()
        if self.cancel_pi:
            self._cancel(self.next, m)
        ws = g.weight_stats_device(self.next, m) if m else np.zeros(4)
        wall = self._gather([ws[2], ws[3]]).sum(axis=0)
        if m:
            g.scale_weights_device(self.next, m, self.nparticles / (wall[0] - wall[1]))
            g.to_particles_device(self.next, m, self.history_counter + int(sum(counts[: self.rank])))
        self.history_counter += int(sum(counts))
        self.bank, self.next = self.next, self.bank
        self.n_bank = m
        self.k_history.append(self.k_col)
        self._noise_counts = [int(v) for v in allv[:, 2]]
        return n_noise

    def noise_simulation(self, n_noise: int):
        g = self.gpu
        original_kcol = self.k_col
        counts = self._noise_counts
        n_total = int(sum(counts))
        avg_wgt_mag = 1.0
        if self.normalize_noise_source and n_total:
            mag = g.weight_magnitude_device(self.noise_bank, n_noise) if n_noise else 0.0
            avg_wgt_mag = float(self._gather([mag]).sum()) / float(n_total)
            if n_noise:
                g.divide_weights_device(self.noise_bank, n_noise, avg_wgt_mag)
        if self.converged and n_noise:
            g.score_source_device(self.noise_bank, n_noise, noise_source=True)
        cur, nxt = self.noise_bank, self.nb_a
        if n_noise:
            g.to_particles_device(cur, n_noise, self.history_counter + int(sum(counts[: self.rank])))
        self.history_counter += n_total
        n, n_glob, gen, total = n_noise, n_total, 0, 0
        while n_glob != 0:
            gen += 1
            total += n_glob
            m = 0
            if n:
                m, _, _, _ = g.transport_noise_device(cur, n, nxt, None, k_col=self.k_col, keff=self.keff,
                                                      converged=self.converged, noise=True, sample_noise=False)
            mc = [int(v) for v in self._gather([float(m)])[:, 0]]
            m_glob = int(sum(mc))
            if self.cancel_noise and gen <= self.n_cancel_noise_gens and m_glob:
                self._cancel(nxt, m)
            if m:
                g.to_particles_device(nxt, m, self.history_counter + int(sum(mc[: self.rank])))
            self.history_counter += m_glob
            cur, nxt = nxt, (self.nb_b if nxt is self.nb_a else self.nb_a)
            n, n_glob = m, m_glob
        self._reduce_tallies()          # mpi::Reduce_sum of the generation's scores before record_generation
        g.tallies_record(avg_wgt_mag)
        g.tallies_clear()
        self.k_col = original_kcol
        self.noise_generations.append(gen)
        self.noise_particles.append(total)
