"""Multi-GPU k-eigenvalue generation loop: one process per GPU, histories sharded by global history id.

The reference shards the same way across MPI ranks (src/power_iterator.cpp:135-166,386-404) but funnels the
fission bank through rank 0 every generation (Gatherv / Scatterv, src/simulation.cpp:113-135) and reduces every
mesh tally to rank 0 (src/mesh_tally.cpp:129).  Here each GPU keeps its slice of the bank resident in HBM and
the only exchange steps are (NCCL over NVLink, via torch.distributed):

  1. one all_gather of a 20-double vector per generation: 6 scores, 8 event counters, site count, W+/W-/N+/N-
     (replaces the 6 Allreduce of src/tallies.cpp:162-167 and the size Bcasts of the bank funnel);
  2. global history ids from an exclusive scan of the gathered site counts, so the RNG streams -- and therefore
     every integer outcome -- are independent of the number of GPUs, as in the reference;
  3. an order-preserving all_to_all_single that moves partition boundaries back to an even split when the
     slices drift apart (replaces gather-to-root + scatter);
  4. on active generations an all_reduce of every tally_gen array before the Welford update (after the source-estimator
     tallies were scored from the normalised fission bank), and a small all_reduce of the entropy bins + total weight.

With world_size == 1 (or no process group) every collective degenerates to a no-op and the loop is exactly the
single-GPU device-resident loop of the C++ PowerIterator.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .backend import BANK_F64, BANK_U64, Backend


def even_split(total: int, world: int):
    """Partition boundaries of `total` items over `world` ranks, remainder to the first ranks
    (src/power_iterator.cpp:138-157 distributes the same way)."""
    base, rem = divmod(int(total), int(world))
    counts = [base + (1 if r < rem else 0) for r in range(world)]
    bounds = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return counts, bounds


def rebalance_plan(counts, rank: int, target=None):
    """Order-preserving re-partition: rank r holds global range [B_r, B_r+1) and must end with [B'_r, B'_r+1) -- an even split, or
    the counts `target` (e.g. everything on rank 0: the reference's particles_to_master).
    Returns (send_splits, recv_splits) for all_to_all_single; pieces arrive in source-rank order == global order."""
    world = len(counts)
    old = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    if target is None:
        _, new = even_split(int(old[-1]), world)
    else:
        new = np.concatenate([[0], np.cumsum(target)]).astype(np.int64)
    send = [int(max(0, min(old[rank + 1], new[q + 1]) - max(old[rank], new[q]))) for q in range(world)]
    recv = [int(max(0, min(old[q + 1], new[rank + 1]) - max(old[q], new[rank]))) for q in range(world)]
    return send, recv


def rebalance_bank(src: dict, dst: dict, keys, counts, rank: int, group=None, target=None) -> int:
    """Order-preserving all_to_all of every bank array: src[k][:counts[rank]] -> dst[k][:new count].  Works on any
    backend (NCCL on the GPUs, gloo in the CPU tests); returns this rank's new count."""
    send, recv = rebalance_plan(counts, rank, target)
    m_new = int(sum(recv))
    via_host = dist.get_backend(group) == "gloo"  # (gloo exchanges host tensors only: two test ranks sharing one GPU)
    for k in keys:
        if via_host and dst[k].is_cuda:
            out = torch.empty(m_new, dtype=dst[k].dtype)
            dist.all_to_all_single(out, src[k][: int(sum(send))].cpu(), recv, send, group=group)
            dst[k][:m_new].copy_(out)
        else:
            dist.all_to_all_single(dst[k][:m_new], src[k][: int(sum(send))], recv, send, group=group)
    return m_new


def gather_vector(vec: np.ndarray, world: int, device, group=None) -> np.ndarray:
    """all_gather of a small fp64 vector -> [world, len] on the host."""
    if world == 1:
        return vec[None, :].copy()
    if dist.get_backend(group) == "gloo":  # (gloo gathers host tensors only: the CPU tests, and two ranks sharing one GPU)
        device = torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64)).to(device)
    out = torch.empty(world * len(vec), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape(world, len(vec))


def pair_distance_sqrd(moments, gather) -> float:
    """sum_ij w_i w_j |r_i - r_j|^2 / (2 W^2) over the bank all ranks hold together (PowerIterator::compute_pair_dist_sqrd,
    src/power_iterator.cpp:637-663) from two passes of this rank's `moments(origin)` = [sum w, sum w (r - o) (3), sum w |r - o|^2]
    (abl_bank_moments_device) and two gathers: about the origin for the weighted centroid c, then about c."""
    m1 = gather(np.asarray(moments((0., 0., 0.)), dtype=np.float64)).sum(axis=0)
    c = m1[1:4] / m1[0]
    m2 = gather(np.asarray(moments(c), dtype=np.float64)).sum(axis=0)
    # (the centroid of the second pass is exact only to rounding: remove what is left of it)
    return float((m2[4] - (m2[1:4] ** 2).sum() / m2[0]) / m2[0])


def global_first_ids(counts, rank: int, global_counter: int) -> int:
    """First history id of this rank's slice: exclusive scan of the gathered counts (power_iterator.cpp:389-394)."""
    return int(global_counter) + int(sum(counts[:rank]))


class _DevPtr:
    """Wraps a raw device pointer for torch.as_tensor through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class DistributedPowerIterator:
    def __init__(self, deck_path: str, device: int, nparticles_per_rank: int, group=None, single: bool = False):
        self.gpu = Backend(deck_path, device)
        self.device = torch.device("cuda", device)
        self.n_local = int(nparticles_per_rank)
        # single = True: this process runs the whole population alone even inside a process group (the 1-rank replay that
        # bench.py compares a multi-rank run with)
        self.use_dist = dist.is_available() and dist.is_initialized() and not single
        self.group = group
        self.rank = dist.get_rank(group) if self.use_dist else 0
        self.world = dist.get_world_size(group) if self.use_dist else 1
        self.n_total = self.n_local * self.world  # tallies->total_weight; the deck's nparticles must equal this
        if self.gpu.info["nparticles"] != self.n_total:
            raise ValueError(f"deck nparticles {self.gpu.info['nparticles']} != world*n_local {self.n_total}")
        self.comb = False
        if self.gpu.info["mode"] == 3:  # ABL_MODE_BRANCHLESS
            import yaml
            from .backend import global_rng_state
            with open(deck_path) as f:
                st = yaml.safe_load(f)["settings"]
            # comb_particles works on the gathered bank with the one global engine on rank 0 (branchless_power_iterator.cpp:353-363,
            # 592-651): the slices go to rank 0 in bank order, are combed there and come back as an even split
            self.comb = bool(st.get("branchless-combing", True))
            self.global_rng = global_rng_state(int(st.get("seed", 19073486328125)))
        # output bank sized from the problem (the first generation runs with k_col = 1 and banks ~ k_inf sites per particle)
        self.cap = self.gpu.fission_capacity(self.n_local, k_col=1.0)
        self.cur = self.gpu.new_device_bank(self.cap)
        self.nxt = self.gpu.new_device_bank(self.cap)
        self.n_cur = 0
        self.use_state = True
        self.k_col = 1.0
        self.global_counter = 0
        self.converged = False
        self.gen = 0
        self.kcol_series, self.nbank_series = [], []
        self.counters = np.zeros(8)
        self.active_particles = 0.0
        self._tally_views = None
        self._cancel_views = None
        # regional cancellation (settings: cancellation + an approximate cancelator; power_iterator.cpp:355-364)
        import yaml
        with open(deck_path) as f:
            deck = yaml.safe_load(f)
        self.cancellation = bool(deck.get("settings", {}).get("cancellation", False)) and "cancelator" in deck
        # optional per-generation diagnostics of PowerIterator (settings: pair-distance-sqrd, families, empty-entropy-bins;
        # src/parser.cpp:833-858, src/power_iterator.cpp:283-297,326-331,362-365,613-615)
        st = deck.get("settings", {})
        self.pair_distance = bool(st.get("pair-distance-sqrd", False))
        self.families = bool(st.get("families", False))
        self.empty_entropy_bins = bool(st.get("empty-entropy-bins", False)) and "entropy" in deck
        self.r_sqrd_series, self.families_series, self.empty_entropy_frac_series = [], [], []
        # Shannon entropy of the fission source (src/entropy.cpp:32-93): bins + total weight, summed over the ranks
        self.entropy_series = []
        self._ebins = None
        if "entropy" in deck:
            nb = int(np.prod([int(v) for v in deck["entropy"]["shape"]]))
            self._ebins = torch.zeros(nb + 1, dtype=torch.float64, device=self.device)

    # ---- helpers ----
    def _gather(self, vec: np.ndarray) -> np.ndarray:
        return gather_vector(vec, self.world, self.device, self.group)

    def tally_tensors(self):
        if self._tally_views is None:
            self._tally_views = []
            for t in range(self.gpu.ntallies()):
                ptr, n = self.gpu.tally_device_ptr(t, 0)
                self._tally_views.append(torch.as_tensor(_DevPtr(ptr, n), device=self.device))
        return self._tally_views

    def _cancel(self, m: int):
        """PowerIterator::perform_regional_cancellation (power_iterator.cpp:751-777) over the GLOBAL fission bank: every rank
        accumulates its slice into its dense bins, one all-reduce (sum) per bin array replaces the reference's gather of the
        whole bank on the master, then every rank applies the bin means to its own particles."""
        gpu = self.gpu
        if self.world == 1:
            gpu.cancel_device(self.nxt, m)
            return
        gpu.cancel_accumulate_device(self.nxt, m)
        if self._cancel_views is None:
            sums, count, nb = gpu.cancel_bins_device()
            self._cancel_views = [torch.as_tensor(_DevPtr(p, nb), device=self.device) for p in sums]
            self._cancel_views.append(torch.as_tensor(_DevPtr(count, nb, "<i4"), device=self.device))  # counts stay < 2^31
        torch.cuda.current_stream().synchronize()
        for t in self._cancel_views:
            dist.all_reduce(t, group=self.group)
        gpu.cancel_apply_device(self.nxt, m)
        # the bins this rank touched are zero again; bins only other ranks touched still hold the global sums
        for t in self._cancel_views:
            t.zero_()

    def _grow(self, need: int):
        """Reallocates both ping-pong banks when the (global) population outgrows them; contents of self.nxt[:keep] survive."""
        if need <= self.cap:
            return
        new_cap = int(need + need // 4 + 4096)
        for name in ("cur", "nxt"):
            old = getattr(self, name)
            new = self.gpu.new_device_bank(new_cap)
            for k in old:
                new[k][: self.cap].copy_(old[k][: self.cap])
            setattr(self, name, new)
        self.cap = new_cap

    def _entropy(self, m: int) -> float:
        if self._ebins is None:
            return 0.0
        self._ebins.zero_()
        self.gpu.entropy_bin_device(self.nxt, m, self._ebins[:-1], self._ebins[-1:])
        if self.world > 1:
            torch.cuda.current_stream().synchronize()
            dist.all_reduce(self._ebins, group=self.group)
        b = self._ebins.cpu().numpy()
        if self.empty_entropy_bins:  # Entropy::calculate_empty_fraction (src/entropy.cpp:95-105)
            self.empty_entropy_frac_series.append(float((b[:-1] == 0.).sum()) / (len(b) - 1))
        total = b[-1]
        p = np.abs(b[:-1]) / total
        p = p[(p != 0.) & (p <= 1.0)]
        return float(-(p * np.log2(p)).sum())

    def _pair_distance_sqrd(self, m: int) -> float:
        """PowerIterator::compute_pair_dist_sqrd (src/power_iterator.cpp:637-663) over the global normalised fission bank: the
        reference's double sum over all pairs, sum_ij w_i w_j |r_i - r_j|^2 / (2 W^2), is sum_i w_i |r_i - c|^2 / W about the
        weighted centroid c -- two passes of abl_bank_moments_device and two small gathers instead of N^2 terms."""
        return pair_distance_sqrd(lambda origin: self.gpu.moments_device(self.nxt, m, origin), self._gather)

    # ---- Simulation::sample_sources: rank r samples the ids [r*n, (r+1)*n) ----
    def initialize(self):
        first = self.rank * self.n_local
        self.gpu.sample_source_device(self.cur, self.n_local, first)
        self.n_cur = self.n_local
        self.use_state = True
        self.global_counter = self.n_total

    # ---- PowerIterator::load_source_from_file (src/power_iterator.cpp:60-133): restart from a saved source ----
    @classmethod
    def from_source(cls, deck_path: str, source, device: int, group=None):
        """The reference's `settings: insource`: `source` is the [N, 9] array Simulation::write_source saves (x y z ux uy uz E
        wgt wgt2; here the `source.npy` of write_results, or an array).  nparticles and the tallies' total weight become
        round(sum of wgt), every rank takes its contiguous share of the rows (the remainder goes to the first ranks), history ids
        are the row numbers, family id = history id, and the RNG streams are seeded from the ids."""
        import tempfile
        import yaml
        src = np.load(source) if isinstance(source, (str, bytes)) or hasattr(source, "__fspath__") else np.asarray(source, dtype=np.float64)
        if src.ndim != 2 or src.shape[1] != 9:
            raise ValueError("Invalid source from file dimensions.")
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        counts, bounds = even_split(len(src), world)
        n_total = int(round(float(src[:, 7].sum())))
        with open(deck_path) as f:
            deck = yaml.safe_load(f)
        deck["settings"]["nparticles"] = n_total
        deck["settings"].pop("insource", None)
        tmp = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
        yaml.safe_dump(deck, tmp, default_flow_style=None, sort_keys=False, width=200)
        tmp.close()
        sim = cls.__new__(cls)
        # (the constructor sizes everything from n_local * world == nparticles; a restart's row count is not its weight)
        if n_total % world != 0:
            raise ValueError("restart: round(total weight) must be divisible by the number of ranks")
        cls.__init__(sim, tmp.name, device, n_total // world, group)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        m = hi - lo
        sim._grow(m)
        part = torch.from_numpy(np.ascontiguousarray(src[lo:hi].T)).to(sim.device)
        for k, key in enumerate(("x", "y", "z", "ux", "uy", "uz", "E", "wgt", "wgt2")):
            sim.cur[key][:m] = part[k]
        ids = torch.arange(lo, hi, dtype=torch.int64, device=sim.device)
        sim.cur["id_a"][:m] = ids
        sim.cur["id_b"][:m] = ids  # PowerIterator::initialize: family id = history id
        sim.n_cur = m
        sim.use_state = False  # initialize_rng(seed, stride) from the history id
        sim.global_counter = len(src)
        return sim

    def source_array(self) -> np.ndarray:
        """This rank's bank as Simulation::write_source lays it out ([n, 9]: x y z ux uy uz E wgt wgt2)."""
        m = self.n_cur
        out = np.zeros((m, 9))
        for k, key in enumerate(("x", "y", "z", "ux", "uy", "uz", "E", "wgt", "wgt2")):
            out[:, k] = self.cur[key][:m].cpu().numpy()
        return out

    def _comb(self, counts):
        """normalize_weights + comb_particles on rank 0 (branchless_power_iterator.cpp:353-363): returns the new counts (even split
        of the combed bank) and this rank's count.  Only the weights come to the host; the combed bank is a gather of rows."""
        from .backend import comb_rows
        m_total = int(sum(counts))
        if self.world > 1:  # particles_to_master, in bank order
            target = [m_total] + [0] * (self.world - 1)
            self._grow(m_total if self.rank == 0 else 1)
            self._rebalance(counts, target)
        n_comb = np.zeros(1)
        if self.rank == 0:
            w = self.nxt["wgt"][:m_total].cpu().numpy()
            # serial sums, as normalize_weights takes them (a cumulative sum adds in order; ceil(sum) decides the combed population)
            w_pos = float(np.cumsum(np.where(w > 0., w, 0.))[-1])
            w_neg = float(np.cumsum(np.where(w > 0., 0., -w))[-1])
            w = w * (self.n_total / (w_pos - w_neg))
            rows, wgts, self.global_rng = comb_rows(w, self.global_rng)
            self._grow(len(rows))
            self.gpu.bank_gather_device(self.nxt, rows, wgts, self.cur)
            self.cur, self.nxt = self.nxt, self.cur  # the combed bank is the fission bank now
            n_comb[0] = len(rows)
        n_comb = int(self._gather(n_comb).sum())
        new_counts, _ = even_split(n_comb, self.world)
        m = n_comb
        if self.world > 1:  # distribute_particles: an even split, in bank order
            self._grow(max(new_counts))
            m = self._rebalance([n_comb] + [0] * (self.world - 1), new_counts)
        return new_counts, m

    def _rebalance(self, counts, target=None):
        """Moves partition boundaries back to an even split (or to `target`), preserving global order."""
        keys = [k for k in BANK_F64 if k != "wgt2"] + ["id_a", "id_b", "id_c"]
        m_new = rebalance_bank(self.nxt, self.cur, keys, counts, self.rank, self.group, target)  # the consumed bank receives
        self.cur, self.nxt = self.nxt, self.cur  # keep the invariant: the fission bank lives in self.nxt
        return m_new

    def generation(self, converged: bool | None = None) -> dict:
        """One power-iteration generation (src/power_iterator.cpp:325-428) on this rank's slice."""
        if converged is not None:
            self.converged = converged
        gpu = self.gpu
        n_in = self.n_cur
        from .backend import BackendError
        if self.families:
            # the families that enter the generation: every rank counts its own set and the counts are summed, as
            # mpi::Reduce_sum(families.size()) does (src/power_iterator.cpp:326-331,288-292)
            nfam = torch.unique(self.cur["id_b"][:n_in]).numel() if n_in else 0
            self.families_series.append(int(self._gather(np.array([float(nfam)])).sum()))
        for attempt in range(3):
            try:
                m, scores, cn = gpu.transport_device(self.cur, n_in, self.nxt, k_col=self.k_col, converged=self.converged,
                                                     use_rng_state=self.use_state)
                break
            except BackendError as e:
                # more sites than the output bank holds: the message carries the count; grow and repeat the generation
                # (scores come back per call; tally_gen holds only this generation's scores and is cleared first)
                if e.code != -3 or "fission bank overflow" not in str(e) or attempt == 2:
                    raise
                need = int(str(e).split("overflow: ")[1].split(" sites")[0])
                self._grow(need)
                if self.converged:
                    gpu.tallies_clear()
        local = np.concatenate([scores, np.array([cn[k] for k in cn], dtype=np.float64), [float(m), float(n_in)]])
        allv = self._gather(local)
        tot = allv.sum(axis=0)
        counts = [int(v) for v in allv[:, 14]]
        n_in_total = int(tot[15])
        self.k_col = tot[0] / self.n_total  # Tallies::calc_gen_values: score / total_weight
        self.counters += tot[6:14]
        if self.converged:
            self.active_particles += n_in_total
        self.kcol_series.append(self.k_col)
        self.nbank_series.append(n_in_total)
        m_total = int(sum(counts))
        m_pre = m
        if m_total == 0:
            raise RuntimeError("No fission neutrons were produced.")
        self.entropy_series.append(self._entropy(m))  # Entropy::add_point over the un-normalised fission bank (power_iterator.cpp:341-353)
        if self.cancellation:
            self._cancel(m)
        if self.comb:
            # branchless-k-eigenvalue: normalisation and comb on the gathered bank, then an even split again
            counts, m = self._comb(counts)
            m_total = int(sum(counts))
        else:
            # weight normalisation over the global bank (src/power_iterator.cpp:538-586)
            ws = gpu.weight_stats_device(self.nxt, m)
            wall = self._gather(ws).sum(axis=0)
            gpu.scale_weights_device(self.nxt, m, self.n_total / (wall[2] - wall[3]))
        if self.pair_distance:
            self.r_sqrd_series.append(self._pair_distance_sqrd(m))  # after normalize_weights (src/power_iterator.cpp:358-366)
        if self.converged:
            gpu.score_source_device(self.nxt, m)  # SourceMeshTally::score_source on the normalised bank (power_iterator.cpp:366-372)
            if self.world > 1:
                torch.cuda.current_stream().synchronize()
                for t in self.tally_tensors():
                    dist.all_reduce(t, group=self.group)
            gpu.tallies_record(1.0)
        gpu.tallies_clear()
        # even out the slices when they drift (order-preserving), then assign global ids
        if self.world > 1:
            target, _ = even_split(m_total, self.world)
            if max(abs(c - t) for c, t in zip(counts, target)) > max(0.01 * m_total / self.world, 64):
                self._grow(max(target))
                m = self._rebalance(counts)
                counts = target
        # next generation's output must hold what this population can bank at the new k_col
        self._grow(gpu.fission_capacity(m, float(self.n_total) / self.world, self.k_col))
        first = global_first_ids(counts, self.rank, self.global_counter)
        gpu.to_particles_device(self.nxt, m, first)
        self.global_counter += m_total
        self.cur, self.nxt = self.nxt, self.cur
        self.n_cur = m
        self.use_state = False
        self.gen += 1
        return {"k_col": self.k_col, "n_in": n_in, "n_in_total": n_in_total, "m": m, "m_pre": m_pre, "m_total": m_total,
                "real_collisions": tot[7], "flights": tot[6], "coll_scores": tot[13], "tl_bins": tot[9],
                "local_real_collisions": cn["real_collisions"], "local_coll_scores": cn["coll_scores"],
                "local_tl_bins": cn["tl_bins"], "local_flights": cn["flights"]}


def _positive_negative_sums(w: "torch.Tensor"):
    """(sum of the positive weights, minus the sum of the negative ones) of a host bank (power_iterator.cpp:538-560),
    without bank-sized temporaries: all-positive banks (every tracker but carter) need one min and one sum, mixed banks
    are reduced in cache-sized chunks."""
    if float(w.min()) >= 0.:
        return float(w.sum()), 0.0
    pos = neg = 0.0
    for c in w.split(1 << 18):
        pos += float(torch.clamp(c, min=0.).sum())
        neg -= float(torch.clamp(c, max=0.).sum())
    return pos, neg


class HostBufferLoop:
    """The reference's own data flow (src/power_iterator.cpp:325-428): every generation the bank crosses the
    Transporter::transport() boundary as HOST arrays -- abl_transport copies it to the device, runs the kernels and
    copies the fission bank and the scores back.  The caller-side steps (k_col, weight normalisation, fresh history
    ids) are done on the host, as PowerIterator does.  Host arrays are pinned.  bench.py's `e2e` leg."""

    F64_OUT = ("x", "y", "z", "ux", "uy", "uz", "E", "wgt")

    def __init__(self, deck_path: str, device: int, nparticles_per_rank: int, group=None):
        self.gpu = Backend(deck_path, device)
        self.device = torch.device("cuda", device)
        self.n_local = int(nparticles_per_rank)
        self.use_dist = dist.is_available() and dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group) if self.use_dist else 0
        self.world = dist.get_world_size(group) if self.use_dist else 1
        self.n_total = self.n_local * self.world
        self.cap = self.gpu.fission_capacity(self.n_local, k_col=1.0)
        self.bufs = [self._pinned_bank(self.cap), self._pinned_bank(self.cap)]
        self.bank = None
        self.k_col = 1.0
        self.global_counter = 0
        self.h2d_bytes = self.d2h_bytes = 0
        self.gens = 0
        self._tally_views = None

    @staticmethod
    def _pinned_bank(cap):
        b = {k: torch.zeros(cap, dtype=torch.float64).pin_memory().numpy() for k in HostBufferLoop.F64_OUT}
        b.update({k: torch.zeros(cap, dtype=torch.int64).pin_memory().numpy().view(np.uint64) for k in BANK_U64})
        return b

    _gather = DistributedPowerIterator._gather
    tally_tensors = DistributedPowerIterator.tally_tensors

    def initialize(self):
        db = self.gpu.new_device_bank(self.n_local)
        self.gpu.sample_source_device(db, self.n_local, self.rank * self.n_local)
        torch.cuda.synchronize()
        b = self.bufs[0]
        n = self.n_local
        for k in self.F64_OUT:
            b[k][:n] = db[k].cpu().numpy()
        for k in BANK_U64:
            b[k][:n] = db[k].cpu().numpy().view(np.uint64)
        self.bank = {k: b[k][:n] for k in self.F64_OUT + BANK_U64}
        self.bank["wgt2"] = None
        self.cur = 0
        self.global_counter = self.n_total

    def generation(self, converged: bool = True) -> dict:
        gpu = self.gpu
        n_in = len(self.bank["x"])
        out = self.bufs[1 - self.cur]
        out_full = dict(out)
        out_full["wgt2"] = None
        # abl_transport_begin: H2D (streamed behind the kernel), kernels; the fission bank waits on the device while the ranks
        # exchange what normalize_weights sums (power_iterator.cpp:538-569)
        m, scores, cn, ws = gpu.transport_begin(self.bank, k_col=self.k_col, converged=converged, capacity=self.cap)
        self.h2d_bytes += n_in * 8 * (8 + sum(1 for k in BANK_U64 if self.bank.get(k) is not None))
        local = np.concatenate([scores, [cn["real_collisions"], float(m), float(n_in), ws[2], ws[3]]])
        allv = self._gather(local)
        tot = allv.sum(axis=0)
        counts = [int(v) for v in allv[:, 7]]
        self.k_col = tot[0] / self.n_total
        if sum(counts) == 0:
            raise RuntimeError("No fission neutrons were produced.")
        if converged:
            if self.world > 1:
                for t in self.tally_tensors():
                    dist.all_reduce(t, group=self.group)
            gpu.tallies_record(1.0)
        gpu.tallies_clear()
        # abl_transport_finish: weights normalised and fresh history ids (power_iterator.cpp:397-399) on the device, D2H
        first = self.global_counter + int(sum(counts[: self.rank]))
        out_full.update({"id_c": None})
        fis = gpu.transport_finish(m, self.n_total / (tot[9] - tot[10]), first, out_full)
        self.d2h_bytes += m * 8 * 10 + 6 * 8 + 8 * 8 + 4 * 8
        self.bank = {k: fis[k] for k in self.F64_OUT}
        self.bank.update({"wgt2": None, "id_a": fis["id_a"], "id_b": fis["id_b"], "id_c": None})
        self.global_counter += int(sum(counts))
        self.cur = 1 - self.cur
        self.gens += 1
        return {"k_col": self.k_col, "n_in": n_in, "n_in_total": int(tot[8]), "m": m, "real_collisions": tot[6]}
