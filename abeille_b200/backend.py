"""ctypes bindings of the two native libraries (see include/abeille_b200.h and host/capi.cpp)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.path.join(_HERE, "lib")

_PD = C.POINTER(C.c_double)
_PU64 = C.POINTER(C.c_uint64)
_PU32 = C.POINTER(C.c_uint32)
_PI32 = C.POINTER(C.c_int32)

BANK_F64 = ("x", "y", "z", "ux", "uy", "uz", "E", "wgt", "wgt2")
BANK_U64 = ("id_a", "id_b", "id_c")
COUNTER_KEYS = ("flights", "real_collisions", "virtual_collisions", "tl_bins", "fission_sites", "boundary_events",
                "lost_at_birth", "coll_scores")

# every symbol include/abeille_b200.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = (
    "abl_create", "abl_destroy", "abl_last_error", "abl_device_info", "abl_last_transport_kernel", "abl_transport",
    "abl_transport_noise", "abl_transport_begin", "abl_transport_finish", "abl_get_trace",
    "abl_transport_device", "abl_transport_noise_device", "abl_bank_weight_magnitude_device", "abl_bank_divide_weights_device",
    "abl_tally_count", "abl_tally_shape", "abl_tallies_record", "abl_tallies_clear",
    "abl_tally_fetch", "abl_tally_device_ptr", "abl_sample_source_device", "abl_bank_weight_stats_device", "abl_bank_moments_device",
    "abl_bank_scale_weights_device", "abl_bank_to_particles_device", "abl_entropy_bin_device",
    "abl_score_source_device", "abl_cancel_device", "abl_cancel_accumulate_device", "abl_cancel_apply_device",
    "abl_cancel_bins_device", "abl_cancel_exact_device", "abl_parent_info_download", "abl_parent_state_download", "abl_bank_alloc_device", "abl_bank_free_device",
    "abl_bank_upload", "abl_bank_download", "abl_bank_gather_device", "abl_device_alloc", "abl_device_free", "abl_device_zero",
    "abl_device_read", "abl_find_cells", "abl_rng_probe", "abl_math_probe", "abl_surface_probe", "abl_set_sampling_xs", "abl_fission_capacity_hint")


class BackendError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"abeille_b200 backend error {code}: {msg}")
        self.code = code


class AblBank(C.Structure):
    _fields_ = [("n", C.c_uint64)] + [(k, _PD) for k in BANK_F64] + [(k, _PU64) for k in BANK_U64]


class AblGenParams(C.Structure):
    _fields_ = [("k_col", C.c_double), ("keff", C.c_double), ("converged", C.c_int32), ("noise", C.c_int32),
                ("trace", C.c_int32), ("sample_noise_source", C.c_int32)]


class AblTrace(C.Structure):
    _fields_ = [("flights", _PU32), ("real", _PU32), ("virt", _PU32), ("fission", _PU32), ("hash", _PU64),
                ("rng_state", _PU64)]


def lib_paths():
    """(CUDA library, host library).  ABEILLE_B200_LIBDIR selects another build of the pair (kernel tuning experiments:
    same ABI, different compile-time constants).  It must be a whole directory: the host library names
    libabeille_b200.so as a dependency (rpath $ORIGIN), and two builds of the CUDA library in one process would
    register their kernels under the same host symbols."""
    d = os.environ.get("ABEILLE_B200_LIBDIR") or _LIBDIR
    return os.path.join(d, "libabeille_b200.so"), os.path.join(d, "libabeille_host.so")


_backend_lib = None
_host_lib = None


def load_backend_lib():
    """libabeille_b200.so; raises if it has not been built (there is no fallback)."""
    global _backend_lib
    if _backend_lib is None:
        path = lib_paths()[0]
        if not os.path.exists(path):
            raise BackendError(-2, f"{path} is missing: build it with __graft_entry__.build() (make -C abeille_b200/csrc)")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.abl_last_error.restype = C.c_char_p
        L.abl_last_error.argtypes = [C.c_void_p]
        L.abl_fission_capacity_hint.restype = C.c_uint64
        L.abl_fission_capacity_hint.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double]
        _backend_lib = L
    return _backend_lib


def load_host_lib():
    global _host_lib
    if _host_lib is None:
        load_backend_lib()
        path = lib_paths()[1]
        if not os.path.exists(path):
            raise BackendError(-2, f"{path} is missing: build it with __graft_entry__.build() (make -C abeille_b200/host)")
        L = C.CDLL(path)
        L.ablh_open.restype = C.c_void_p
        L.ablh_open.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.ablh_close.argtypes = [C.c_void_p]
        L.ablh_last_error.restype = C.c_char_p
        L.ablh_last_error.argtypes = [C.c_void_p]
        L.ablh_backend.restype = C.c_void_p
        L.ablh_backend.argtypes = [C.c_void_p]
        _host_lib = L
    return _host_lib


def new_bank(n: int, wgt2: bool = True) -> dict:
    b = {k: np.zeros(n, dtype=np.float64) for k in BANK_F64 if wgt2 or k != "wgt2"}
    b.update({k: np.zeros(n, dtype=np.uint64) for k in BANK_U64})
    return b


def _host_struct(b: dict, n=None) -> AblBank:
    """abl_bank over numpy arrays (missing / None entries become NULL)."""
    s = AblBank()
    s.n = int(len(b["x"]) if n is None else n)
    for k in BANK_F64:
        a = b.get(k)
        if a is None:
            setattr(s, k, None)
            continue
        assert a.dtype == np.float64 and a.flags.c_contiguous, k
        setattr(s, k, a.ctypes.data_as(_PD))
    for k in BANK_U64:
        a = b.get(k)
        if a is None:
            setattr(s, k, None)
            continue
        assert a.dtype == np.uint64 and a.flags.c_contiguous, k
        setattr(s, k, a.ctypes.data_as(_PU64))
    return s


def _device_struct(b: dict, n: int) -> AblBank:
    """abl_bank over torch CUDA tensors (float64 / int64 viewed as uint64)."""
    s = AblBank()
    s.n = int(n)
    for k in BANK_F64:
        t = b.get(k)
        setattr(s, k, C.cast(C.c_void_p(t.data_ptr()), _PD) if t is not None else None)
    for k in BANK_U64:
        t = b.get(k)
        setattr(s, k, C.cast(C.c_void_p(t.data_ptr()), _PU64) if t is not None else None)
    return s


def parse_only(yaml_path: str):
    """Host-only: parse + flatten a deck (no device).  Returns (info dict, sampling xs per group)."""
    L = load_host_lib()
    info = (C.c_int64 * 12)()
    maj = (C.c_double * 64)()
    err = C.create_string_buffer(1024)
    rc = L.ablh_parse_only(yaml_path.encode(), info, maj, 64, err, 1024)
    if rc != 0:
        raise BackendError(rc, err.value.decode())
    keys = ("ngroups", "nparticles", "ngenerations", "nignored", "ntallies", "tracking", "mode", "nsurfaces", "ncells",
            "nuniverses", "nmaterials", "max_stack_depth")
    d = {k: int(v) for k, v in zip(keys, info)}
    return d, np.array(maj[: d["ngroups"]])


def source_records(yaml_path: str) -> np.ndarray:
    """Host-only: the flattened abl_source records of a deck, [nsources, 18]: weight, fissile_only, is_box, low[3], hi[3], energy,
    direction_kind (ABL_DIR_*), dir[3] (normalised), cos_aperture, energy_kind (ABL_EN_*), en_a, en_b."""
    L = load_host_lib()
    out = np.zeros(20 * 64)
    err = C.create_string_buffer(1024)
    n = L.ablh_sources(yaml_path.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(len(out)), err, 1024)
    if n < 0:
        raise BackendError(1, err.value.decode())
    return out[: 20 * n].reshape(n, 20)[:, :18].copy()


def dump_tables(yaml_path: str) -> dict:
    L = load_host_lib()
    cap = 1 << 24
    out = C.create_string_buffer(cap)
    err = C.create_string_buffer(1024)
    rc = L.ablh_dump_tables(yaml_path.encode(), out, C.c_int64(cap), err, 1024)
    if rc != 0:
        raise BackendError(rc, err.value.decode())
    tables = {}
    for line in out.value.decode().splitlines():
        name, *vals = line.split()
        if name in ("rpn", "universe_cells", "lattice_tiles", "angle", "cells", "root"):
            tables[name] = np.array([int(v) for v in vals], dtype=np.int64)
        else:
            tables[name] = np.array([float(v) for v in vals], dtype=np.float64)
    return tables


def global_rng_state(seed: int = 19073486328125):
    """(state, increment) of the reference's settings::rng when a simulation starts (settings.cpp:116-119, simulation.cpp:52)."""
    mask = (1 << 64) - 1
    inc = 1442695040888963407
    return ((seed + inc) * 6364136223846793005 + inc) & mask, 5


def comb_particles(bank: dict, rng2):
    """BranchlessPowerIterator::comb_particles (src/branchless_power_iterator.cpp:592-651) on a host bank: the C++ host's
    serial restatement (std::shuffle on the global engine).  Returns (combed bank, (state, increment) afterwards)."""
    L = load_host_lib()
    n = len(bank["x"])
    out = new_bank(2 * n + int(np.ceil(np.abs(bank["wgt"]).sum())) + 16)
    r = (C.c_uint64 * 2)(int(rng2[0]), int(rng2[1]))
    nout = C.c_uint64(0)
    L.ablh_comb_particles.argtypes = [C.POINTER(AblBank), C.POINTER(AblBank), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    rc = L.ablh_comb_particles(C.byref(_host_struct(bank)), C.byref(_host_struct(out)), C.byref(nout), r)
    if rc != 0:
        raise BackendError(rc, L.ablh_last_error(None).decode())
    m = int(nout.value)
    return {k: v[:m].copy() for k, v in out.items()}, (int(r[0]), int(r[1]))


def comb_rows(wgt: np.ndarray, rng2):
    """The comb from the weights alone (abeille::comb_rows): returns (rows uint32, weights, (state, increment) afterwards) -- the
    combed bank is row rows[k] of the bank with weight weights[k]."""
    L = load_host_lib()
    wgt = np.ascontiguousarray(wgt, dtype=np.float64)
    cap = int(np.ceil(np.abs(wgt).sum())) + len(wgt) + 16
    rows, wgts = np.zeros(cap, dtype=np.uint32), np.zeros(cap)
    r = (C.c_uint64 * 2)(int(rng2[0]), int(rng2[1]))
    nout = C.c_uint64(0)
    rc = L.ablh_comb_rows(wgt.ctypes.data_as(_PD), C.c_uint64(len(wgt)), r, rows.ctypes.data_as(C.POINTER(C.c_uint32)), wgts.ctypes.data_as(_PD),
                          C.c_uint64(cap), C.byref(nout))
    if rc != 0:
        raise BackendError(rc, L.ablh_last_error(None).decode())
    m = int(nout.value)
    return rows[:m].copy(), wgts[:m].copy(), (int(r[0]), int(r[1]))


def yaml_roundtrip(text: str) -> str:
    L = load_host_lib()
    cap = 1 << 22
    out = C.create_string_buffer(cap)
    rc = L.ablh_yaml_roundtrip(text.encode(), out, C.c_int64(cap))
    if rc != 0:
        raise BackendError(rc, out.value.decode())
    return out.value.decode()


class Backend:
    """One deck loaded through the C++ host onto one CUDA device."""

    def __init__(self, yaml_path: str, device: int = 0):
        self.H = load_host_lib()
        self.L = load_backend_lib()
        err = C.create_string_buffer(2048)
        ctx = self.H.ablh_open(str(yaml_path).encode(), int(device), err, 2048)
        if not ctx:
            raise BackendError(-1, err.value.decode())
        self.ctx = C.c_void_p(ctx)
        self.h = C.c_void_p(self.H.ablh_backend(self.ctx))
        self.device = int(device)
        info = (C.c_int64 * 12)()
        self.H.ablh_info(self.ctx, info)
        keys = ("ngroups", "nparticles", "ngenerations", "nignored", "ntallies", "tracking", "mode", "nsurfaces",
                "ncells", "nuniverses", "nmaterials", "max_stack_depth")
        self.info = {k: int(v) for k, v in zip(keys, info)}

    # ---- lifetime ----
    def close(self):
        if getattr(self, "ctx", None):
            self.H.ablh_close(self.ctx)
            self.ctx = None
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BackendError(rc, self.L.abl_last_error(self.h).decode())

    def _hcheck(self, rc):
        if rc != 0:
            raise BackendError(rc, self.H.ablh_last_error(self.ctx).decode())

    def device_info(self) -> dict:
        sm, ma, mi, nl = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
        self._check(self.L.abl_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(nl)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "kernel_launches": int(nl.value)}

    def last_transport_kernel(self) -> dict:
        ms, g, b = C.c_float(), C.c_int(), C.c_int()
        self._check(self.L.abl_last_transport_kernel(self.h, C.byref(ms), C.byref(g), C.byref(b)))
        return {"ms": float(ms.value), "grid": g.value, "block": b.value}

    # ---- Transporter::transport, host buffers (C ABI) ----
    def transport(self, bank: dict, k_col: float = 1.0, converged: bool = False, trace: bool = False,
                  capacity: int | None = None, out: dict | None = None):
        """abl_transport.  Returns (fission bank dict, scores[6], counters dict)."""
        n = len(bank["x"])
        cap = int(capacity if capacity is not None else self.fission_capacity(n, float(np.abs(bank["wgt"]).sum()), k_col))
        if out is None:
            out = new_bank(cap, wgt2=False)
        sin = _host_struct(bank)
        sout = _host_struct(out, cap)
        gp = AblGenParams(float(k_col), 1.0, int(bool(converged)), 0, int(bool(trace)), 0)
        nout = C.c_uint64(0)
        scores = np.zeros(6)
        cn = np.zeros(8, dtype=np.uint64)
        rc = self.L.abl_transport(self.h, C.byref(sin), C.byref(gp), C.byref(sout), C.byref(nout),
                                  scores.ctypes.data_as(_PD), cn.ctypes.data_as(_PU64))
        self._check(rc)
        m = int(nout.value)
        fis = {k: v[:m] for k, v in out.items() if v is not None}
        return fis, scores, {k: int(v) for k, v in zip(COUNTER_KEYS, cn)}

    def transport_begin(self, bank: dict, k_col: float = 1.0, converged: bool = False, capacity: int | None = None):
        """abl_transport_begin: host bank in, the fission bank stays on the device.  Returns (n_fission, scores[6], counters
        dict, weight stats [n+, n-, sum w+, -sum w-])."""
        n = len(bank["x"])
        cap = int(capacity if capacity is not None else self.fission_capacity(n, float(np.abs(bank["wgt"]).sum()), k_col))
        sin = _host_struct(bank)
        gp = AblGenParams(float(k_col), 1.0, int(bool(converged)), 0, 0, 0)
        nout = C.c_uint64(0)
        scores, cn, ws = np.zeros(6), np.zeros(8, dtype=np.uint64), np.zeros(4)
        rc = self.L.abl_transport_begin(self.h, C.byref(sin), C.byref(gp), C.c_uint64(cap), C.byref(nout), scores.ctypes.data_as(_PD),
                                        cn.ctypes.data_as(_PU64), ws.ctypes.data_as(_PD))
        self._check(rc)
        return int(nout.value), scores, {k: int(v) for k, v in zip(COUNTER_KEYS, cn)}, ws

    def transport_finish(self, n_fission: int, weight_factor: float, first_history_id: int, out: dict):
        """abl_transport_finish: normalise and number the fission bank on the device, then copy it to the host arrays `out`
        (x y z ux uy uz E wgt id_a id_b).  Returns the bank as views of `out`."""
        cap = len(out["x"])
        sout = _host_struct(out, cap)
        self._check(self.L.abl_transport_finish(self.h, C.c_double(weight_factor), C.c_uint64(first_history_id), C.byref(sout)))
        return {k: v[:n_fission] for k, v in out.items() if v is not None}

    def trace(self, n: int) -> dict:
        t = {k: np.zeros(n, dtype=np.uint32) for k in ("flights", "real", "virtual", "fission")}
        t["hash"] = np.zeros(n, dtype=np.uint64)
        t["rng_state"] = np.zeros(n, dtype=np.uint64)
        s = AblTrace(t["flights"].ctypes.data_as(_PU32), t["real"].ctypes.data_as(_PU32),
                     t["virtual"].ctypes.data_as(_PU32), t["fission"].ctypes.data_as(_PU32),
                     t["hash"].ctypes.data_as(_PU64), t["rng_state"].ctypes.data_as(_PU64))
        self._check(self.L.abl_get_trace(self.h, C.c_uint64(n), C.byref(s)))
        return t

    # ---- Transporter::transport through the C++ GPUTransporter adapter (vector<Particle>) ----
    def transport_vectors(self, bank: dict, k_col: float = 1.0, converged: bool = False, capacity: int | None = None):
        n = len(bank["x"])
        cap = int(capacity if capacity is not None else self.fission_capacity(n, float(np.abs(bank["wgt"]).sum()), k_col))
        out = new_bank(cap)
        sin = _host_struct(bank)
        sout = _host_struct(out, cap)
        nout = C.c_uint64(0)
        scores = np.zeros(6)
        rc = self.H.ablh_transport(self.ctx, C.byref(sin), C.c_int(int(bool(converged))), C.c_double(float(k_col)),
                                   C.byref(sout), C.byref(nout), scores.ctypes.data_as(_PD))
        self._hcheck(rc)
        m = int(nout.value)
        return {k: v[:m] for k, v in out.items()}, scores

    def transport_vectors_noise(self, bank: dict, k_col: float = 1.0, keff: float = 1.0, converged: bool = False,
                                noise: bool = False, sample_noise: bool = False, capacity: int | None = None):
        """GPUTransporter::transport(bank, noise, &noise_bank, &noise_maker) through the C++ adapter (host vectors)."""
        n = len(bank["x"])
        cap = int(capacity if capacity is not None else max(6 * n + 4096, 4096))
        out, nout_bank = new_bank(cap), new_bank(cap)
        sin = _host_struct(bank)
        sout, snoise = _host_struct(out, cap), _host_struct(nout_bank, cap)
        nout, nnoise = C.c_uint64(0), C.c_uint64(0)
        rc = self.H.ablh_transport_noise(self.ctx, C.byref(sin), C.c_int(int(bool(converged))), C.c_double(float(k_col)),
                                         C.c_double(float(keff)), C.c_int(int(bool(noise))), C.c_int(int(bool(sample_noise))),
                                         C.byref(sout), C.byref(nout), C.byref(snoise), C.byref(nnoise))
        self._hcheck(rc)
        m, mn = int(nout.value), int(nnoise.value)
        return {k: v[:m] for k, v in out.items()}, {k: v[:mn] for k, v in nout_bank.items()}

    def run_power_iteration(self, ngen: int, nignored: int, resident: bool = True) -> dict:
        arr = {k: np.zeros(ngen) for k in ("kcol", "ktrk", "leak", "mig", "entropy")}
        nbank = np.zeros(ngen, dtype=np.uint64)
        summ = np.zeros(10)
        rc = self.H.ablh_run_power_iteration(self.ctx, C.c_int(ngen), C.c_int(nignored), C.c_int(int(bool(resident))),
                                             *[arr[k].ctypes.data_as(_PD) for k in ("kcol", "ktrk", "leak", "mig", "entropy")],
                                             nbank.ctypes.data_as(_PU64), summ.ctypes.data_as(_PD))
        self._hcheck(rc)
        ran = int(np.count_nonzero(nbank))  # fewer than ngen when settings: max-run-time ended the run (PowerIterator::check_time)
        if ran < ngen:
            arr = {k: v[:ran] for k, v in arr.items()}
            nbank = nbank[:ran]
        arr["nbank"] = nbank
        arr.update(kcol_avg=summ[0], kcol_err=summ[1], ktrk_avg=summ[2], ktrk_err=summ[3], leak_avg=summ[4],
                   leak_err=summ[5], seconds=summ[6], active_particles=summ[7], real_collisions=summ[8], flights=summ[9])
        return arr

    def write_results(self, directory: str):
        os.makedirs(directory, exist_ok=True)
        self._hcheck(self.H.ablh_write_results(self.ctx, str(directory).encode()))

    # ---- tallies ----
    def ntallies(self) -> int:
        return int(self.L.abl_tally_count(self.h))

    def tally_shape(self, t: int):
        sh = (C.c_uint64 * 4)()
        self._check(self.L.abl_tally_shape(self.h, C.c_int(t), sh))
        return tuple(int(v) for v in sh)

    def tally(self, t: int, which: str = "gen") -> np.ndarray:
        shape = self.tally_shape(t)
        out = np.zeros(int(np.prod(shape)))
        self._check(self.L.abl_tally_fetch(self.h, C.c_int(t), C.c_int({"gen": 0, "avg": 1, "var": 2, "std": 3}[which]),
                                           out.ctypes.data_as(_PD)))
        return out.reshape(shape)

    def tallies_record(self, multiplier: float = 1.0):
        self._check(self.L.abl_tallies_record(self.h, C.c_double(multiplier)))

    def tallies_clear(self):
        self._check(self.L.abl_tallies_clear(self.h))

    def tally_device_ptr(self, t: int, which: int = 0):
        p = C.POINTER(C.c_double)()
        n = C.c_uint64()
        self._check(self.L.abl_tally_device_ptr(self.h, C.c_int(t), C.c_int(which), C.byref(p), C.byref(n)))
        return C.cast(p, C.c_void_p).value, int(n.value)

    # ---- probes ----
    def find_cells(self, r: np.ndarray, u: np.ndarray):
        r = np.ascontiguousarray(r, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        n = r.shape[0]
        cell = np.zeros(n, dtype=np.int32)
        mat = np.zeros(n, dtype=np.int32)
        self._check(self.L.abl_find_cells(self.h, C.c_uint64(n), r.ctypes.data_as(_PD), u.ctypes.data_as(_PD),
                                          cell.ctypes.data_as(_PI32), mat.ctypes.data_as(_PI32)))
        return cell, mat

    def surface_probe(self, surface_index: int, r: np.ndarray, u: np.ndarray, on_surf: np.ndarray):
        """Surface::sign / distance / norm of one surface of the problem at n points (include/abeille_b200.h)."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        on = np.ascontiguousarray(on_surf, dtype=np.int32)
        n = r.shape[0]
        sign = np.zeros(n, dtype=np.int32)
        dist, norm = np.zeros(n), np.zeros((n, 3))
        self._check(self.L.abl_surface_probe(self.h, C.c_int(surface_index), C.c_uint64(n), r.ctypes.data_as(_PD),
                                             u.ctypes.data_as(_PD), on.ctypes.data_as(_PI32), sign.ctypes.data_as(_PI32),
                                             dist.ctypes.data_as(_PD), norm.ctypes.data_as(_PD)))
        return sign, dist, norm

    def fission_capacity(self, n: int, sum_abs_weight: float | None = None, k_col: float = 1.0) -> int:
        """Output-bank capacity for n histories (abl_fission_capacity_hint: sized from the problem's most reactive material)."""
        w = float(n if sum_abs_weight is None else sum_abs_weight)
        return int(self.L.abl_fission_capacity_hint(self.h, C.c_uint64(int(n)), C.c_double(w), C.c_double(float(k_col))))

    def set_sampling_xs(self, xs):
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        self._check(self.L.abl_set_sampling_xs(self.h, xs.ctypes.data_as(_PD), C.c_int(len(xs))))

    def rng_probe(self, history_id: int, n: int):
        u32 = np.zeros(n, dtype=np.uint32)
        rnd = np.zeros(n)
        self._check(self.L.abl_rng_probe(self.h, C.c_uint64(history_id), C.c_int(n), u32.ctypes.data_as(_PU32),
                                         rnd.ctypes.data_as(_PD)))
        return u32, rnd

    def math_probe(self, x: np.ndarray):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lg, sn, cs = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        self._check(self.L.abl_math_probe(self.h, C.c_int(len(x)), x.ctypes.data_as(_PD), lg.ctypes.data_as(_PD),
                                          sn.ctypes.data_as(_PD), cs.ctypes.data_as(_PD)))
        return lg, sn, cs

    # ---- device-resident banks (torch tensors own the memory) ----
    def new_device_bank(self, capacity: int) -> dict:
        import torch
        dev = torch.device("cuda", self.device)
        b = {k: torch.zeros(capacity, dtype=torch.float64, device=dev) for k in BANK_F64}
        b.update({k: torch.zeros(capacity, dtype=torch.int64, device=dev) for k in BANK_U64})
        return b

    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def sample_source_device(self, bank: dict, n: int, first_history_id: int = 0):
        s = _device_struct(bank, len(bank["x"]))
        self._check(self.L.abl_sample_source_device(self.h, C.c_uint64(n), C.c_uint64(first_history_id), C.byref(s),
                                                    self._stream()))

    def transport_device(self, bank: dict, n: int, out: dict, k_col: float = 1.0, converged: bool = False,
                         trace: bool = False, use_rng_state: bool = False):
        """abl_transport_device on torch tensors.  Returns (n_fission, scores[6], counters dict)."""
        sin = _device_struct(bank, n)
        if not use_rng_state:
            sin.id_c = None
        sout = _device_struct(out, len(out["x"]))
        gp = AblGenParams(float(k_col), 1.0, int(bool(converged)), 0, int(bool(trace)), 0)
        nout = C.c_uint64(0)
        scores = np.zeros(6)
        cn = np.zeros(8, dtype=np.uint64)
        rc = self.L.abl_transport_device(self.h, C.byref(sin), C.byref(gp), C.byref(sout), C.byref(nout),
                                         scores.ctypes.data_as(_PD), cn.ctypes.data_as(_PU64), self._stream())
        self._check(rc)
        return int(nout.value), scores, {k: int(v) for k, v in zip(COUNTER_KEYS, cn)}

    def transport_noise_device(self, bank: dict, n: int, out: dict, noise_out: dict | None = None, k_col: float = 1.0,
                               keff: float = 1.0, converged: bool = False, noise: bool = False, sample_noise: bool = False,
                               trace: bool = False, use_rng_state: bool = False):
        """abl_transport_noise_device (simulation: noise).  noise=True transports noise particles (complex weights);
        sample_noise=True is a power-iteration generation that fills noise_out with the sampled noise source.
        Returns (n_fission, n_noise, scores[6], counters dict)."""
        sin = _device_struct(bank, n)
        if not use_rng_state:
            sin.id_c = None
        sout = _device_struct(out, len(out["x"]))
        snoise = _device_struct(noise_out, len(noise_out["x"])) if noise_out is not None else None
        gp = AblGenParams(float(k_col), float(keff), int(bool(converged)), int(bool(noise)), int(bool(trace)), int(bool(sample_noise)))
        nout, nnoise = C.c_uint64(0), C.c_uint64(0)
        scores = np.zeros(6)
        cn = np.zeros(8, dtype=np.uint64)
        rc = self.L.abl_transport_noise_device(self.h, C.byref(sin), C.byref(gp), C.byref(sout), C.byref(nout),
                                               C.byref(snoise) if snoise is not None else None, C.byref(nnoise),
                                               scores.ctypes.data_as(_PD), cn.ctypes.data_as(_PU64), self._stream())
        self._check(rc)
        return int(nout.value), int(nnoise.value), scores, {k: int(v) for k, v in zip(COUNTER_KEYS, cn)}

    def weight_magnitude_device(self, bank: dict, n: int) -> float:
        s = _device_struct(bank, n)
        out = C.c_double(0.)
        self._check(self.L.abl_bank_weight_magnitude_device(self.h, C.byref(s), C.byref(out), self._stream()))
        return float(out.value)

    def divide_weights_device(self, bank: dict, n: int, divisor: float):
        s = _device_struct(bank, n)
        self._check(self.L.abl_bank_divide_weights_device(self.h, C.byref(s), C.c_double(divisor), self._stream()))

    def weight_stats_device(self, bank: dict, n: int):
        s = _device_struct(bank, n)
        st = np.zeros(4)
        self._check(self.L.abl_bank_weight_stats_device(self.h, C.byref(s), st.ctypes.data_as(_PD), self._stream()))
        return st

    def moments_device(self, bank: dict, n: int, origin=(0., 0., 0.)) -> np.ndarray:
        """abl_bank_moments_device: [sum w, sum w (r - o) (3), sum w |r - o|^2] of the first n rows."""
        s = _device_struct(bank, n)
        o = np.asarray(origin, dtype=np.float64)
        out = np.zeros(5)
        self._check(self.L.abl_bank_moments_device(self.h, C.byref(s), o.ctypes.data_as(_PD), out.ctypes.data_as(_PD), self._stream()))
        return out

    def scale_weights_device(self, bank: dict, n: int, factor: float):
        s = _device_struct(bank, n)
        self._check(self.L.abl_bank_scale_weights_device(self.h, C.byref(s), C.c_double(factor), self._stream()))

    def to_particles_device(self, bank: dict, n: int, first_history_id: int):
        s = _device_struct(bank, n)
        self._check(self.L.abl_bank_to_particles_device(self.h, C.byref(s), C.c_uint64(first_history_id), self._stream()))

    def entropy_bin_device(self, bank: dict, n: int, bins, total):
        s = _device_struct(bank, n)
        self._check(self.L.abl_entropy_bin_device(self.h, C.byref(s), C.c_void_p(bins.data_ptr()),
                                                  C.c_void_p(total.data_ptr()), self._stream()))

    def score_source_device(self, bank: dict, n: int, noise_source: bool = False):
        s = _device_struct(bank, n)
        self._check(self.L.abl_score_source_device(self.h, C.byref(s), C.c_int(int(noise_source)), self._stream()))

    def cancel_device(self, bank: dict, n: int):
        s = _device_struct(bank, n)
        self._check(self.L.abl_cancel_device(self.h, C.byref(s), self._stream()))

    def parent_info(self, n: int) -> np.ndarray:
        """[n, 4]: the parents' previous position and sampling cross section of the rows of the last fission bank (problems
        with an exact cancelator; BankedParticle::parents_previous_position / Esmp_parent)."""
        cols = [np.zeros(n) for _ in range(4)]
        self._check(self.L.abl_parent_info_download(self.h, C.c_uint64(n), *[c.ctypes.data_as(_PD) for c in cols]))
        return np.stack(cols, axis=1)

    def parent_state(self, n: int) -> np.ndarray:
        """[n, 6]: parents_previous_direction, the parent's energy before its last scatter, its energy at the fission, and whether
        its collision before was virtual (what the reference's `type: exact` cancelator reads on top of parent_info)."""
        cols = [np.zeros(n) for _ in range(6)]
        self._check(self.L.abl_parent_state_download(self.h, C.c_uint64(n), *[c.ctypes.data_as(_PD) for c in cols]))
        return np.stack(cols, axis=1)

    def cancel_exact_device(self, bank: dict, n: int, rng2):
        """abl_cancel_exact_device: BasicExactMGCancelator on the fission bank of the last transport call (torch tensors of
        capacity len(bank['x'])).  Returns (rows afterwards, (state, increment) of the global engine afterwards)."""
        s = _device_struct(bank, n)
        r = (C.c_uint64 * 2)(int(rng2[0]), int(rng2[1]))
        self._check(self.L.abl_cancel_exact_device(self.h, C.byref(s), C.c_uint64(len(bank["x"])), r, self._stream()))
        return int(s.n), (int(r[0]), int(r[1]))

    def bank_gather_device(self, src: dict, rows: np.ndarray, wgts, dst: dict) -> int:
        """abl_bank_gather_device: dst row k = src row rows[k] (weight wgts[k] when given); torch banks; returns the row count."""
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        ssrc, sdst = _device_struct(src, len(src["x"])), _device_struct(dst, len(dst["x"]))
        wp = np.ascontiguousarray(wgts, dtype=np.float64).ctypes.data_as(_PD) if wgts is not None else None
        self._check(self.L.abl_bank_gather_device(self.h, C.byref(ssrc), rows.ctypes.data_as(C.POINTER(C.c_uint32)), wp, C.c_uint64(len(rows)),
                                                  C.byref(sdst), self._stream()))
        return len(rows)

    def cancel_accumulate_device(self, bank: dict, n: int):
        s = _device_struct(bank, n)
        self._check(self.L.abl_cancel_accumulate_device(self.h, C.byref(s), self._stream()))

    def cancel_apply_device(self, bank: dict, n: int):
        s = _device_struct(bank, n)
        self._check(self.L.abl_cancel_apply_device(self.h, C.byref(s), self._stream()))

    def cancel_bins_device(self):
        """Device pointers of the dense cancellation bins: ([4 fp64 sum arrays], u32 member counts, number of bins)."""
        sums = (C.c_void_p * 4)()
        count = C.c_void_p()
        nb = C.c_uint64(0)
        self._check(self.L.abl_cancel_bins_device(self.h, sums, C.byref(count), C.byref(nb)))
        return [int(p) for p in sums], int(count.value), int(nb.value)
