"""oracle/api.py -- TEST INFRASTRUCTURE: ctypes front-end of the CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package (abeille_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from . import deck as _deck

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_PD = C.POINTER(C.c_double)
_PU64 = C.POINTER(C.c_uint64)


class _Bank(C.Structure):
    _fields_ = [("n", C.c_uint64)] + [(k, _PD) for k in ("x", "y", "z", "ux", "uy", "uz", "E", "wgt", "wgt2")] + [
        ("id_a", _PU64), ("id_b", _PU64), ("id_c", _PU64)]


BANK_F64 = ("x", "y", "z", "ux", "uy", "uz", "E", "wgt", "wgt2")
BANK_U64 = ("id_a", "id_b", "id_c")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_load.restype = C.c_void_p
        L.orc_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_rng_exponential.restype = C.c_double
        L.orc_rng_exponential.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_double]
        L.orc_tally_size.restype = C.c_uint64
        for name in ("orc_free", "orc_set_nparticles", "orc_set_converged", "orc_set_kcol", "orc_set_trace",
                     "orc_reset_counters", "orc_get_counters", "orc_get_majorant", "orc_ngroups", "orc_nparticles",
                     "orc_find_cells", "orc_sample_source", "orc_set_history_counter", "orc_transport",
                     "orc_get_trace", "orc_ntallies", "orc_tally_size", "orc_tally_shape", "orc_tally_get",
                     "orc_tallies_record", "orc_tallies_clear", "orc_tallies_calc_gen", "orc_cancel_and_normalize",
                     "orc_run_power_iteration", "orc_pi_init", "orc_pi_run"):
            getattr(L, name).argtypes = None
        _lib = L
    return _lib


def new_bank(n: int) -> dict:
    b = {k: np.zeros(n, dtype=np.float64) for k in BANK_F64}
    b.update({k: np.zeros(n, dtype=np.uint64) for k in BANK_U64})
    return b


def _as_struct(b: dict, n: int | None = None, allow_null=()):
    s = _Bank()
    s.n = int(len(b["x"]) if n is None else n)
    for k in BANK_F64:
        a = b.get(k)
        if a is None:
            a = np.zeros(s.n)
            b[k] = a
        assert a.dtype == np.float64 and a.flags.c_contiguous
        setattr(s, k, a.ctypes.data_as(_PD))
    for k in BANK_U64:
        a = b.get(k)
        if a is None and k in allow_null:
            setattr(s, k, None)
            continue
        assert a.dtype == np.uint64 and a.flags.c_contiguous
        setattr(s, k, a.ctypes.data_as(_PU64))
    return s


def set_math(mode: str) -> None:
    """'libm' = glibc log/sin/cos like the reference; 'det' = the shared deterministic kernels."""
    lib().orc_set_math(C.c_int(0 if mode == "libm" else 1))


def get_math() -> str:
    return "libm" if int(lib().orc_get_math()) == 0 else "det"


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(int(n)))


def max_threads() -> int:
    return int(lib().orc_max_threads())


class Oracle:
    """One loaded problem (deck) on the CPU oracle."""

    def __init__(self, yaml_path: str, overrides: dict | None = None):
        L = lib()
        with tempfile.NamedTemporaryFile("w", suffix=".orcdeck", delete=False) as f:
            f.write(_deck.deck_to_text(_deck.apply_overrides(_deck.load_yaml(yaml_path), overrides)))
            path = f.name
        err = C.create_string_buffer(512)
        self.h = L.orc_load(path.encode(), err, 512)
        os.unlink(path)
        if not self.h:
            raise RuntimeError("oracle: " + err.value.decode())
        self.h = C.c_void_p(self.h)
        self.G = int(L.orc_ngroups(self.h))

    def close(self):
        if self.h:
            lib().orc_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return lib().orc_last_error(self.h).decode()

    # --- settings ---
    def set_nparticles(self, n): lib().orc_set_nparticles(self.h, C.c_int(int(n)))
    def set_converged(self, c): lib().orc_set_converged(self.h, C.c_int(int(bool(c))))
    def set_kcol(self, k): lib().orc_set_kcol(self.h, C.c_double(float(k)))
    def set_trace(self, t): lib().orc_set_trace(self.h, C.c_int(int(bool(t))))
    def set_history_counter(self, c): lib().orc_set_history_counter(self.h, C.c_uint64(int(c)))
    def nparticles(self): return int(lib().orc_nparticles(self.h))

    def majorant(self):
        maj = np.zeros(self.G)
        smp = np.zeros(self.G)
        lib().orc_get_majorant(self.h, maj.ctypes.data_as(_PD), smp.ctypes.data_as(_PD))
        return maj, smp

    def counters(self) -> dict:
        out = np.zeros(8, dtype=np.uint64)
        lib().orc_get_counters(self.h, out.ctypes.data_as(_PU64))
        keys = ("flights", "real_collisions", "virtual_collisions", "tl_bins", "fission_sites",
                "boundary_events", "lost_at_birth", "coll_scores")
        return {k: int(v) for k, v in zip(keys, out)}

    def reset_counters(self): lib().orc_reset_counters(self.h)

    # --- geometry probe ---
    def find_cells(self, r: np.ndarray, u: np.ndarray):
        r = np.ascontiguousarray(r, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        n = r.shape[0]
        cell = np.zeros(n, dtype=np.int32)
        mat = np.zeros(n, dtype=np.int32)
        lib().orc_find_cells(self.h, C.c_int(n), r.ctypes.data_as(_PD), u.ctypes.data_as(_PD),
                             cell.ctypes.data_as(C.POINTER(C.c_int32)), mat.ctypes.data_as(C.POINTER(C.c_int32)))
        return cell, mat

    # --- source / transport ---
    def sample_source(self, n: int) -> dict:
        """Bank dict: id_a = history id, id_b = family id, id_c = rng state after source sampling."""
        b = new_bank(n)
        s = _as_struct(b)
        if lib().orc_sample_source(self.h, C.byref(s)) != 0:
            raise RuntimeError("oracle: " + self._err())
        return b

    def transport(self, bank: dict, noise: bool = False, capacity: int | None = None):
        """Returns (fission_bank dict, scores[6], n_out). Input id_c=None => rng from seed/stride/history id."""
        n = len(bank["x"])
        cap = int(capacity if capacity is not None else max(4 * n, 1024))
        out = new_bank(cap)
        sin = _as_struct(bank, allow_null=("id_c",))
        sout = _as_struct(out)
        nout = C.c_uint64(0)
        scores = np.zeros(6)
        rc = lib().orc_transport(self.h, C.byref(sin), C.c_int(int(noise)), C.byref(sout), C.byref(nout),
                                 scores.ctypes.data_as(_PD))
        if rc != 0:
            raise RuntimeError("oracle: " + self._err())
        m = int(nout.value)
        if m > cap:
            raise RuntimeError(f"oracle: fission bank capacity {cap} < {m}")
        return {k: v[:m].copy() for k, v in out.items()}, scores, m

    def transport_noise(self, bank: dict, noise: bool, sample_noise: bool, capacity: int | None = None):
        """Transporter::transport(bank, noise, &noise_bank, &noise_maker).  Returns (fission bank, noise-source bank,
        scores[6]); the noise-source bank is empty unless sample_noise."""
        n = len(bank["x"])
        cap = int(capacity if capacity is not None else max(6 * n, 1024))
        out, nout_bank = new_bank(cap), new_bank(cap)
        sin = _as_struct(bank, allow_null=("id_c",))
        sout, snoise = _as_struct(out), _as_struct(nout_bank)
        nout, nnoise = C.c_uint64(0), C.c_uint64(0)
        scores = np.zeros(6)
        rc = lib().orc_transport_noise(self.h, C.byref(sin), C.c_int(int(noise)), C.c_int(int(sample_noise)), C.byref(sout),
                                       C.byref(nout), C.byref(snoise), C.byref(nnoise), scores.ctypes.data_as(_PD))
        if rc != 0:
            raise RuntimeError("oracle: " + self._err())
        m, mn = int(nout.value), int(nnoise.value)
        if m > cap or mn > cap:
            raise RuntimeError(f"oracle: bank capacity {cap} < {max(m, mn)}")
        return {k: v[:m].copy() for k, v in out.items()}, {k: v[:mn].copy() for k, v in nout_bank.items()}, scores

    def set_keff(self, k): lib().orc_set_keff(self.h, C.c_double(float(k)))

    def score_source(self, bank: dict, noise_source: bool = False):
        s = _as_struct(bank, allow_null=("id_c",))
        lib().orc_score_source(self.h, C.byref(s), C.c_int(int(noise_source)))

    def trace(self, n: int) -> dict:
        t = {k: np.zeros(n, dtype=np.uint32) for k in ("flights", "real", "virtual", "fission")}
        t["hash"] = np.zeros(n, dtype=np.uint64)
        t["rng_state"] = np.zeros(n, dtype=np.uint64)
        P32 = C.POINTER(C.c_uint32)
        lib().orc_get_trace(self.h, t["flights"].ctypes.data_as(P32), t["real"].ctypes.data_as(P32),
                            t["virtual"].ctypes.data_as(P32), t["fission"].ctypes.data_as(P32),
                            t["hash"].ctypes.data_as(_PU64), t["rng_state"].ctypes.data_as(_PU64))
        return t

    # --- tallies ---
    def ntallies(self): return int(lib().orc_ntallies(self.h))

    def tally_estimator(self, t): return int(lib().orc_tally_estimator(self.h, C.c_int(t)))  # 0 collision, 1 track-length, 2 source

    def tally_shape(self, t):
        sh = np.zeros(4, dtype=np.uint64)
        lib().orc_tally_shape(self.h, C.c_int(t), sh.ctypes.data_as(_PU64))
        return tuple(int(v) for v in sh)

    def tally(self, t: int, which: str = "gen") -> np.ndarray:
        n = int(lib().orc_tally_size(self.h, C.c_int(t)))
        out = np.zeros(n)
        lib().orc_tally_get(self.h, C.c_int(t), C.c_int({"gen": 0, "avg": 1, "var": 2, "std": 3}[which]),
                            out.ctypes.data_as(_PD))
        return out.reshape(self.tally_shape(t))

    def tallies_record(self, mult=1.0): lib().orc_tallies_record(self.h, C.c_double(mult))
    def tallies_clear(self): lib().orc_tallies_clear(self.h)

    def calc_gen_values(self):
        out = np.zeros(6)
        lib().orc_tallies_calc_gen(self.h, out.ctypes.data_as(_PD))
        return out

    def cancel_and_normalize(self, bank: dict, do_cancel: bool):
        s = _as_struct(bank)
        stats = np.zeros(6)
        lib().orc_cancel_and_normalize(self.h, C.byref(s), C.c_int(int(do_cancel)), stats.ctypes.data_as(_PD))
        return stats

    def cancel(self, bank: dict):
        s = _as_struct(bank)
        lib().orc_cancel(self.h, C.byref(s))

    def last_parent_info(self, n: int) -> np.ndarray:
        """[n, 4]: BankedParticle::parents_previous_position and Esmp_parent of the fission bank the last transport() returned."""
        out = np.zeros((n, 4))
        lib().orc_last_parent_info.restype = C.c_uint64
        m = lib().orc_last_parent_info(self.h, out.ctypes.data_as(_PD), C.c_uint64(n))
        assert m == n, (m, n)
        return out

    def last_parent_state(self, n: int) -> np.ndarray:
        """[n, 6]: parents_previous_direction, parents_previous_previous_energy, parents_previous_energy, parents_previous_was_virtual
        (what `type: exact` reads, exact_mg_cancelator.cpp:319-327) of the fission bank the last transport() returned."""
        out = np.zeros((n, 6))
        lib().orc_last_parent_state.restype = C.c_uint64
        m = lib().orc_last_parent_state(self.h, out.ctypes.data_as(_PD), C.c_uint64(n))
        assert m == n, (m, n)
        return out

    def cancel_exact(self, bank: dict, parent_info: np.ndarray, rng2, parent_state: np.ndarray | None = None):
        """PowerIterator::perform_regional_cancellation with the deck's BasicExactMGCancelator: weights reduced in place, uniform
        particles appended.  Returns (bank, (state, increment) of the global engine afterwards)."""
        n = len(bank["x"])
        out = new_bank(3 * n + 4096)
        r = np.array(rng2, dtype=np.uint64)
        nout = C.c_uint64(0)
        pi = np.ascontiguousarray(parent_info, dtype=np.float64)
        ps = np.ascontiguousarray(parent_state, dtype=np.float64) if parent_state is not None else None
        if lib().orc_cancel_exact_state(self.h, C.byref(_as_struct(bank)), pi.ctypes.data_as(_PD), ps.ctypes.data_as(_PD) if ps is not None else None,
                                        C.byref(_as_struct(out)), C.byref(nout), r.ctypes.data_as(_PU64)) != 0:
            raise RuntimeError("oracle: " + self._err())
        m = int(nout.value)
        return {k: v[:m].copy() for k, v in out.items()}, (int(r[0]), int(r[1]))

    def comb(self, bank: dict, rng2):
        """BranchlessPowerIterator::comb_particles on a (normalised) fission bank; rng2 = (state, increment) of the global
        engine.  Returns (combed bank, (state, increment) afterwards)."""
        n = len(bank["x"])
        out = new_bank(2 * n + int(np.ceil(np.abs(bank["wgt"]).sum())) + 16)
        r = np.array(rng2, dtype=np.uint64)
        nout = C.c_uint64(0)
        if lib().orc_comb(self.h, C.byref(_as_struct(bank)), C.byref(_as_struct(out)), C.byref(nout),
                          r.ctypes.data_as(_PU64)) != 0:
            raise RuntimeError("oracle: " + self._err())
        m = int(nout.value)
        return {k: v[:m].copy() for k, v in out.items()}, (int(r[0]), int(r[1]))

    def run_noise(self, settings: dict) -> dict:
        """The reference's Noise driver (src/noise.cpp:211-559) over the oracle's transport: nignored power-iteration
        generations, then `ngenerations` noise batches of (nskip - 1) plain generations, one generation that samples
        the noise source, and the noise simulation of that source (inner generations until no particle is left)."""
        st = settings
        N = int(st.get("nparticles", 100000))
        nbatches, nignored, nskip = int(st.get("ngenerations", 120)), int(st.get("nignored", 20)), int(st.get("nskip", 10))
        keff = float(st.get("keff", 1.0))
        cancel_pi, cancel_noise = bool(st.get("cancellation", False)), bool(st.get("noise-cancellation", False))
        n_cancel_gens = int(st.get("cancel-noise-gens", 2147483647))
        normalize_src = bool(st.get("normalize-noise-source", True))
        self.set_keff(keff)
        bank = self.sample_source(N)
        counter = N
        k_col = 1.0
        out = {"k_col": [], "noise_generations": [], "noise_particles": []}

        def to_particles(fis, first):  # Particle(p.r, p.u, p.E, p.wgt[, p.wgt2], histories_counter++) + initialize_rng
            b = {k: fis[k] for k in BANK_F64}
            m = len(fis["x"])
            b["id_a"] = np.arange(first, first + m, dtype=np.uint64)
            b["id_b"] = fis["id_c"].copy()
            b["id_c"] = None
            return b

        def power_iteration(sample_noise):  # noise.cpp:305-372
            nonlocal bank, counter, k_col
            self.set_kcol(k_col)
            fis, nb, scores = self.transport_noise(bank, False, sample_noise)
            if len(fis["x"]) == 0:
                raise RuntimeError("No fission neutrons were produced.")
            k_col = scores[0] / N
            self.tallies_clear()
            self.cancel_and_normalize(fis, cancel_pi)
            bank = to_particles(fis, counter)
            counter += len(fis["x"])
            out["k_col"].append(k_col)
            return nb

        def noise_simulation(nb):  # noise.cpp:425-559
            nonlocal counter
            n = len(nb["x"])
            avg = 1.0
            if normalize_src and n:
                # summed in bank order as the reference does (src/noise.cpp:443-447); numpy's pairwise sum differs in the
                # last bits, and the normalised weights feed every later comparison
                total_mag = 0.0
                for mag in np.sqrt(nb["wgt"] * nb["wgt"] + nb["wgt2"] * nb["wgt2"]).tolist():
                    total_mag += mag
                avg = total_mag / float(n)
                nb["wgt"] = nb["wgt"] / avg
                nb["wgt2"] = nb["wgt2"] / avg
            if n:
                self.score_source(nb, noise_source=True)
            cur = to_particles(nb, counter)
            counter += n
            gen = total = 0
            while n != 0:
                gen += 1
                total += n
                fis, _, _ = self.transport_noise(cur, True, False)
                m = len(fis["x"])
                if cancel_noise and gen <= n_cancel_gens and m:
                    self.cancel(fis)
                cur = to_particles(fis, counter)
                counter += m
                n = m
            self.tallies_record(avg)
            self.tallies_clear()
            out["noise_generations"].append(gen)
            out["noise_particles"].append(total)

        self.set_converged(False)
        for _ in range(nignored):
            power_iteration(False)
        self.set_converged(True)
        for _ in range(nbatches):
            for _ in range(nskip - 1):
                power_iteration(False)
            nb = power_iteration(True)
            noise_simulation(nb)
        out["k_col"] = np.array(out["k_col"])
        out["final_bank"] = (len(bank["x"]), int(bank["id_a"][0]) if len(bank["x"]) else 0, int(counter))
        return out

    def run_fixed_source(self, nbatches: int) -> dict:
        """FixedSource::run (src/fixed_source.cpp:81-175), one rank: every batch samples the source and transports it; the
        fission neutrons are secondaries of their history (the deck's mode makes make_fission_neutrons do that), so the
        returned bank must be empty; then Tallies::calc_gen_values / record_generation."""
        n = self.nparticles()
        self.set_converged(True)
        counter = 0
        kcol, leak, mig = [], [], []
        for _ in range(nbatches):
            self.set_history_counter(counter)
            bank = self.sample_source(n)
            counter += n
            fis, scores, _ = self.transport(bank)
            if len(fis["x"]):
                raise RuntimeError("Returned bank not empty on fixed-source transport.")
            kcol.append(scores[0] / n)
            leak.append(scores[4] / n)
            mig.append(scores[5] / n)
            self.tallies_record(1.0)
            self.tallies_clear()
        return {"kcol": np.array(kcol), "leak": np.array(leak), "mig": np.array(mig)}

    def run_modified_fixed_source(self, nbatches: int) -> dict:
        """ModifiedFixedSource::run (src/modified_fixed_source.cpp:59-141): every batch samples the source and follows the whole
        fission chain -- the fission bank of a transport call becomes the next call's bank with its weights kept and fresh
        history ids (Particle(p.r, p.u, p.E, p.wgt, p.wgt2, histories_counter++) + initialize_rng) until it is empty; then
        Tallies::calc_gen_values / record_generation.  make_fission_neutrons does not divide by k_col in this mode
        (transporter.cpp:381-386), which is k_col = 1 here (x / 1 is x exactly)."""
        n = self.nparticles()
        self.set_converged(True)
        counter = 0
        kcol, leak, mig = [], [], []
        transported = 0
        for _ in range(nbatches):
            self.set_history_counter(counter)
            bank = self.sample_source(n)
            counter += n
            totals = np.zeros(6)
            while len(bank["x"]):
                self.set_kcol(1.0)
                m_in = len(bank["x"])
                fis, scores, _ = self.transport(bank)
                transported += m_in
                totals += scores
                m = len(fis["x"])
                bank = {k: fis[k].copy() for k in ("x", "y", "z", "ux", "uy", "uz", "E", "wgt")}
                bank["wgt2"] = fis["wgt2"].copy() if fis.get("wgt2") is not None else np.zeros(m)
                bank["id_a"] = np.arange(counter, counter + m, dtype=np.uint64)
                bank["id_b"] = np.zeros(m, dtype=np.uint64)
                bank["id_c"] = None
                counter += m
            kcol.append(totals[0] / n)
            leak.append(totals[4] / n)
            mig.append(totals[5] / n)
            self.tallies_record(1.0)
            self.tallies_clear()
        return {"kcol": np.array(kcol), "leak": np.array(leak), "mig": np.array(mig), "transported": transported}

    def run_power_iteration(self, ngen: int, nignored: int) -> dict:
        arr = {k: np.zeros(ngen) for k in ("kcol", "ktrk", "leak", "mig", "entropy")}
        nbank = np.zeros(ngen, dtype=np.uint64)
        summ = np.zeros(8)
        rc = lib().orc_run_power_iteration(self.h, C.c_int(ngen), C.c_int(nignored),
                                           *[arr[k].ctypes.data_as(_PD) for k in ("kcol", "ktrk", "leak", "mig", "entropy")],
                                           nbank.ctypes.data_as(_PU64), summ.ctypes.data_as(_PD))
        if rc != 0:
            raise RuntimeError("oracle: " + self._err())
        arr["nbank"] = nbank
        arr.update(kcol_avg=summ[0], kcol_err=summ[1], ktrk_avg=summ[2], ktrk_err=summ[3], leak_avg=summ[4],
                   leak_err=summ[5], seconds=summ[6], active_particles=summ[7])
        return arr


    def set_pi_diagnostics(self, on: bool = True):
        """settings: pair-distance-sqrd, families, empty-entropy-bins for the generation loops below (the pair distance is the
        reference's sum over all pairs: keep nparticles small)."""
        lib().orc_pi_set_diagnostics(self.h, C.c_int(int(bool(on))))

    def pi_diagnostics(self, ngen: int) -> dict:
        a = [np.zeros(ngen) for _ in range(3)]
        n = lib().orc_pi_diagnostics(self.h, *[v.ctypes.data_as(_PD) for v in a], C.c_int(ngen))
        return {"r_sqrd": a[0][:n], "families": a[1][:n], "empty": a[2][:n]}

    # stateful generation loop (bench.py CPU legs)
    def pi_init(self, nignored: int):
        if lib().orc_pi_init(self.h, C.c_int(int(nignored))) != 0:
            raise RuntimeError("oracle: " + self._err())

    def pi_run(self, ngen: int) -> dict:
        out = np.zeros(4)
        if lib().orc_pi_run(self.h, C.c_int(int(ngen)), out.ctypes.data_as(_PD)) != 0:
            raise RuntimeError("oracle: " + self._err())
        return {"seconds": out[0], "particles": out[1], "real_collisions": out[2], "k_col": out[3]}


# --- RNG / math known-answer helpers ---
def rng_stream(seed, stride, hid, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_rng_stream(C.c_uint64(seed), C.c_uint64(stride), C.c_uint64(hid), C.c_int(n),
                         out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def rng_rand(seed, stride, hid, n):
    out = np.zeros(n)
    lib().orc_rng_rand(C.c_uint64(seed), C.c_uint64(stride), C.c_uint64(hid), C.c_int(n), out.ctypes.data_as(_PD))
    return out


def rng_exponential(seed, stride, hid, lam):
    return float(lib().orc_rng_exponential(C.c_uint64(seed), C.c_uint64(stride), C.c_uint64(hid), C.c_double(lam)))


def rng_discrete(seed, stride, hid, weights, ndraws):
    w = np.ascontiguousarray(weights, dtype=np.float64)
    out = np.zeros(ndraws, dtype=np.int32)
    nd = lib().orc_rng_discrete(C.c_uint64(seed), C.c_uint64(stride), C.c_uint64(hid), w.ctypes.data_as(_PD),
                                C.c_int(len(w)), C.c_int(ndraws), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out, int(nd)


def acos_probe(x):
    """acos of the current math mode (libm, or the deterministic sequence the kernels share)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros(len(x))
    lib().orc_acos_eval(C.c_int(len(x)), x.ctypes.data_as(_PD), out.ctypes.data_as(_PD))
    return out


def math_eval(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    lg, sn, cs = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    lib().orc_math_eval(C.c_int(len(x)), x.ctypes.data_as(_PD), lg.ctypes.data_as(_PD), sn.ctypes.data_as(_PD),
                        cs.ctypes.data_as(_PD))
    return lg, sn, cs
