"""oracle/ref_pins.py -- TEST INFRASTRUCTURE: pins the oracle against the reference's own compiled code.

oracle/_ref/libabeille_ref.so (oracle/Makefile target `ref`, driver oracle/ref_probe.cpp) is built from the reference's
own translation units where they lie under /root/reference: the nine surface classes, Direction / rotate_direction, the
RNG helpers over pcg32, MGAngleDistribution::sample_mu and LegendreDistribution::linearize.  `evaluate("reference")`
runs the seeded cases below through it, `evaluate("oracle")` through the oracle's restatement (orc_*_probe entry points
of oracle/orc_main.cpp).  Both return {name: ndarray}; tests/test_reference_pins.py demands bit-equality, and
scripts/make_ref_pins.py stores the reference's outputs as tests/golden/ref_pins.npz so that the comparison also runs on
machines without /root/reference (the GPU box).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess

import numpy as np

from . import api

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libabeille_ref.so")
REFERENCE_ROOT = "/root/reference"

_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int)
_ref = None

SEED, STRIDE = 19073486328125, 152917  # settings.hpp defaults (rng_seed, rng_stride)


def reference_available() -> bool:
    return os.path.exists(REF_LIB) or os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_LIB):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
        _ref = C.CDLL(REF_LIB)
        _ref.ref_rng_exponential.restype = C.c_double
        _ref.ref_rng_discrete.restype = C.c_double
    return _ref


def _d(a):
    return a.ctypes.data_as(_PD)


@contextlib.contextmanager
def _reference_math(impl):
    """The reference calls glibc's log / sin / cos; the oracle is compared with it in its "libm" math mode and on one thread
    (score sums in bank order).  Its "det" mode (the operation sequence shared with the CUDA kernels) differs from glibc by
    an ulp on a fraction of the arguments, which tests/test_oracle.py quantifies."""
    if impl == "reference":
        yield
        return
    mode = api.get_math()
    api.set_math("libm")
    api.set_threads(1)
    try:
        yield
    finally:
        api.set_math(mode)
        api.set_threads(api.max_threads())


# ---------------------------------------------------------------------------------------------------------- cases
SURFACE_NAMES = ("xplane", "yplane", "zplane", "plane", "xcyl", "ycyl", "zcyl", "cyl", "sphere")


def surface_cases(n=1500):
    """Per surface type: 7 parameters in the order of the constructors (xplane.hpp ... sphere.hpp), n positions,
    directions and on_surf flags.  A third of the positions are projected onto (or within an ulp-scale distance of) the
    surface, so the SURFACE_COINCIDENT / on_surf branches of distance() and sign() are exercised."""
    rng = np.random.default_rng(20261017)
    out = []
    for t, name in enumerate(SURFACE_NAMES):
        p = np.zeros(7)
        if t <= 2:
            p[0] = rng.uniform(-2, 2)
        elif t == 3:
            p[:3] = rng.normal(size=3)
            p[:3] /= np.linalg.norm(p[:3])
            p[3] = rng.uniform(-1, 1)
        elif t <= 6:
            p[:2] = rng.uniform(-1, 1, 2)
            p[2] = rng.uniform(0.3, 2.5)
        elif t == 7:
            p[:3] = rng.uniform(-1, 1, 3)
            p[3:6] = rng.normal(size=3)
            p[6] = rng.uniform(0.3, 2.5)
        else:
            p[:3] = rng.uniform(-1, 1, 3)
            p[3] = rng.uniform(0.3, 2.5)
        r = rng.uniform(-4, 4, (n, 3))
        u = rng.normal(size=(n, 3))
        # axis-aligned and grazing directions
        u[::17] = np.eye(3)[rng.integers(0, 3, len(u[::17]))] * rng.choice([-1., 1.], (len(u[::17]), 1))
        near = slice(0, n // 3)
        r[near] = _project(t, p, r[near])
        r[near.start:near.stop:2] += rng.normal(scale=1e-10, size=(len(r[near.start:near.stop:2]), 3))
        on = (rng.random(n) < 0.25).astype(np.int32)
        out.append((t, name, p, np.ascontiguousarray(r), np.ascontiguousarray(u), on))
    return out


def _project(t, p, r):
    r = r.copy()
    if t <= 2:
        r[:, t] = p[0]
    elif t == 3:
        nrm = p[:3]
        r -= np.outer(r @ nrm - p[3], nrm)
    elif t <= 6:
        ax = t - 4
        oth = [k for k in range(3) if k != ax]
        c = np.zeros(3)
        c[oth] = p[:2]
        d = r - c
        d[:, ax] = 0
        d *= (p[2] / np.linalg.norm(d, axis=1))[:, None]
        r[:, oth] = (c + d)[:, oth]
    elif t == 7:
        a = p[3:6] / np.linalg.norm(p[3:6])
        d = r - p[:3]
        par = np.outer(d @ a, a)
        perp = d - par
        perp *= (p[6] / np.linalg.norm(perp, axis=1))[:, None]
        r = p[:3] + par + perp
    else:
        d = r - p[:3]
        d *= (p[3] / np.linalg.norm(d, axis=1))[:, None]
        r = p[:3] + d
    return r


def direction_cases(n=4000):
    rng = np.random.default_rng(7)
    xyz = rng.normal(size=(n, 3)) * rng.choice([1e-3, 1., 1e3], (n, 1))
    u = rng.normal(size=(n, 3))
    u[:50] = np.array([0., 0., 1.]) * rng.choice([-1., 1.], (50, 1))  # the |w| == 1 pole of rotate_direction
    u[50:100, :2] *= 1e-9
    mu = rng.uniform(-1, 1, n)
    mu[:10] = [-1., 1., 0., -1., 1., 0., 0.5, -0.5, 1., -1.]
    phi = rng.uniform(0, 2 * np.pi, n)
    return np.ascontiguousarray(xyz), np.ascontiguousarray(u), mu, phi


LEGENDRE_CASES = ([0.3], [0.3, 0.1], [-0.2, 0.05, 0.01], [0.1, 0.2, 0.05, 0.02, 0.01], [0.9, 0.7, 0.5, 0.3, 0.2, 0.1])
DISCRETE_WEIGHTS = ([1.0], [0.2, 0.8], [0.0, 0.3, 0.0, 0.7], [1e-3, 0.5, 0.25, 0.249],
                    [0.1, 0.2, 0.3, 0.05, 0.15, 0.1, 0.1])
HISTORY_IDS = (0, 1, 2, 1000, 123456789, 10 ** 12 + 7)


def mg_material_cases():
    """Materials of the shipped multigroup decks (tests/decks) plus one synthetic 3-group material with P1..P3 moments and
    delayed groups, as the flat arrays ref_mg_nuclide / orc_mg_nuclide_probe take.  A single chi row is replicated for
    every incoming group (src/mg_nuclide.cpp:836-852); non-fissile materials get a flat chi so that both sides normalise
    finite numbers (their fission branch is never reached in transport)."""
    from . import deck as _deck
    decks = os.path.join(os.path.dirname(_HERE), "tests", "decks")
    cases = []
    for fname in ("c5g7_delta_collision.yaml", "UD2O-2-1-SL.yaml", "PUa-1-2-SL.yaml", "Ua-1-1-CY.yaml", "noise_oscillation.yaml"):
        d = _deck.load_yaml(os.path.join(decks, fname))
        G = int(d["settings"]["ngroups"])
        eb = np.asarray(d["settings"]["energy-bounds"], dtype=np.float64)
        for m in d["materials"]:
            cases.append((f"{fname.split('.')[0]}_m{m['id']}", _material_arrays(m, G, eb)))
    G = 3
    rng = np.random.default_rng(3)
    Es = rng.uniform(0.05, 0.6, (G, G))
    Es[2, 0] = 0.0  # a zero-probability transfer
    m = {"total": (Es.sum(1) + 0.3).tolist(), "absorption": [0.3] * G, "fission": [0.1, 0.12, 0.2],
         "nu_prompt": [2.4, 2.45, 2.5], "nu_delayed": [0.02, 0.018, 0.016],
         "chi": [[0.7, 0.3, 0.0], [0.6, 0.3, 0.1], [0.5, 0.25, 0.25]], "scatter": Es.tolist(),
         "P1": (0.3 * rng.random((G, G))).tolist(), "P2": (0.1 * rng.random((G, G))).tolist(),
         "P3": (0.03 * rng.random((G, G))).tolist(),
         "delayed_groups": {"probabilities": [0.1, 0.2, 0.3, 0.4], "constants": [0.012, 0.03, 0.11, 0.3]}}
    cases.append(("synthetic_3g", _material_arrays(m, G, np.array([0., 1., 2., 3.]))))
    return cases


def _material_arrays(m, G, eb):
    f = lambda k, dflt: np.asarray(m.get(k, dflt), dtype=np.float64)  # noqa: E731
    if "nu" in m:
        nup, nud = f("nu", None), np.zeros(G)  # nu is nu_prompt, nu_delayed = 0 (mg_nuclide.cpp:742-752)
    elif "nu_prompt" in m:
        nup, nud = f("nu_prompt", None), f("nu_delayed", None)
    else:
        nup, nud = np.zeros(G), np.zeros(G)
    chi = np.asarray(m.get("chi", [[1.0] * G]), dtype=np.float64)
    if chi.shape[0] == 1:
        chi = np.repeat(chi, G, axis=0)
    if not chi.any():
        chi = np.ones((G, G))
    legs = [np.asarray(m[f"P{l}"], dtype=np.float64) for l in range(1, 6) if f"P{l}" in m]
    assert all(f"P{l}" in m for l in range(1, len(legs) + 1)), "moments must be contiguous for the probe"
    dg = m.get("delayed_groups") or {"probabilities": [], "constants": []}
    return dict(G=G, eb=eb, Et=f("total", None), Ea=f("absorption", None), Ef=f("fission", [0.0] * G), nup=nup, nud=nud,
                chi=np.ascontiguousarray(chi), Es=np.ascontiguousarray(f("scatter", None)),
                leg=np.ascontiguousarray(np.stack(legs)) if legs else np.zeros(1), nleg=len(legs),
                Pd=np.asarray(dg["probabilities"], dtype=np.float64), lam=np.asarray(dg["constants"], dtype=np.float64))


NHIST_MG, NDRAW_MG = 8, 250

# geometry walks: deck, bounding box of the ray origins, number of rays, steps per ray, mean flight of the delta walk
GEOMETRY_CASES = (
    ("c5g7_delta_collision.yaml", (-32.13, -32.13, -0.5), (32.13, 32.13, 0.5), 300, 50, 1.2),       # nested RectLattices
    ("ref_sqr_c5g7_surface_tl.yaml", None, None, 300, 50, 1.2),                                       # reflective quarter core
    ("Ua-1-1-CY.yaml", None, None, 300, 12, 2.0),                                                     # z-cylinder, vacuum
    ("PUa-1-0-SL.yaml", None, None, 300, 12, 1.0),                                                    # slab
    ("UD2O-2-1-SL.yaml", None, None, 300, 12, 3.0),
    ("noise_oscillation.yaml", None, None, 300, 20, 2.0),
)


def geometry_text(deck: dict):
    """The surfaces / cells / universes / root lines of the oracle's flat deck (oracle/deck.py), and the material ids."""
    from . import deck as _deck
    lines = _deck.deck_to_text(deck).splitlines()
    a = next(i for i, l in enumerate(lines) if l.startswith("nsurf "))
    b = next(i for i, l in enumerate(lines) if l.startswith("root "))
    return "\n".join(lines[a:b + 1]) + "\n", np.asarray([int(m["id"]) for m in deck["materials"]], dtype=np.int32)


def geometry_rays(deck: dict, low, hi, n, nsteps, mean_flight, seed):
    """Seeded ray origins inside the source box of the deck (or the given box), isotropic directions with axis-aligned and
    lattice-diagonal ones mixed in, exponential flight lengths and post-collision directions for the delta walk."""
    rng = np.random.default_rng(seed)
    if low is None:
        sp = deck["sources"][0]["spatial"]
        low, hi = (sp["low"], sp["hi"]) if sp["type"] == "box" else (np.asarray(sp["position"]) - 0.3, np.asarray(sp["position"]) + 0.3)
    r = rng.uniform(low, hi, (n, 3))
    u = rng.normal(size=(n, 3))
    u[::11] = np.eye(3)[rng.integers(0, 2, len(u[::11]))] * rng.choice([-1., 1.], (len(u[::11]), 1))
    u[5::13, 2] = 0.0                                   # in-plane flights of a 2D core
    u[7::29] = np.array([1., 1., 0.]) * rng.choice([-1., 1.], (len(u[7::29]), 1))  # through lattice corners
    r[3::31, 0] = np.round(r[3::31, 0] / 1.26) * 1.26  # born on a pin-cell (tile) boundary
    d = rng.exponential(mean_flight, (n, nsteps))
    unew = rng.normal(size=(n, nsteps, 3))
    return np.ascontiguousarray(r), np.ascontiguousarray(u), np.ascontiguousarray(d), np.ascontiguousarray(unew)


# whole transport() calls: deck, particles, k_col of the previous generation
TRANSPORT_CASES = (
    ("PUa-1-0-IN.yaml", 1000, 1.0), ("PUa-1-0-SL.yaml", 1000, 0.97), ("PUa-1-1-SL.yaml", 1000, 1.0), ("PUa-1-2-SL.yaml", 1000, 1.02),
    ("PUb-1-0-IN.yaml", 1000, 1.0), ("PUb-1-0-SL.yaml", 1000, 1.0), ("UD2O-2-1-SL.yaml", 1000, 1.01), ("Ua-1-1-CY.yaml", 1000, 0.99),
    ("Ua-1-1-IN.yaml", 1000, 1.0),
    ("c5g7_delta_collision.yaml", 2500, 1.17), ("c5g7_carter_cancel.yaml", 2500, 1.17), ("c5g7_surface_tracklength.yaml", 1500, 1.17),
    ("c5g7_delta_tracklength.yaml", 1500, 1.17), ("ref_sqr_c5g7_surface_tl.yaml", 1500, 1.1),
)


# HexLattice (src/hex_lattice.cpp) through DeltaTracker::transport: pointy top with the origin at zero, flat top with the origin
# off zero.  (The reference's SurfaceTracker does not terminate on a hexagonal lattice: tests/decks/make_hex_decks.py.)
HEX_CASES = (("hex_delta_collision.yaml", 2000, 1.1), ("hex_delta_flat_offset.yaml", 2000, 0.95))


def transport_bank(deck: dict, n: int, seed: int, negative: bool):
    """A seeded bank inside the deck's source region: unit directions, the source energy, weights in [0.3, 1.7] (a tenth of
    them negative for the carter deck, as after an under-estimated majorant), history ids with gaps."""
    rng = np.random.default_rng(seed)
    src = deck["sources"][0]
    sp = src["spatial"]
    low, hi = (sp["low"], sp["hi"]) if sp["type"] == "box" else (np.asarray(sp["position"]) - 0.25, np.asarray(sp["position"]) + 0.25)
    r = rng.uniform(low, hi, (n, 3))
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    G = int(deck["settings"]["ngroups"])
    eb = np.asarray(deck["settings"]["energy-bounds"], dtype=np.float64)
    g = rng.integers(0, G, n)
    E = 0.5 * (eb[g] + eb[g + 1])  # particles of every group, at the group mid-point as the samplers leave them
    w = rng.uniform(0.3, 1.7, n)
    if negative:
        w[::10] *= -1.0
    hid = (np.arange(n, dtype=np.uint64) * np.uint64(3) + np.uint64(11))
    return np.ascontiguousarray(r), np.ascontiguousarray(u), E, w, hid


# ------------------------------------------------------------------------------------------------------ evaluation
def evaluate(impl: str) -> dict:
    """Run every case through `impl` ("reference": oracle/_ref, "oracle": the restatement)."""
    with _reference_math(impl):
        return _evaluate(impl)


def _evaluate(impl: str) -> dict:
    ref = impl == "reference"
    L = ref_lib() if ref else api.lib()
    out = {}
    # surfaces
    fn = L.ref_surface if ref else L.orc_surface_probe
    for t, name, p, r, u, on in surface_cases():
        n = len(r)
        sign = np.zeros(n, dtype=np.int32)
        dist, norm = np.zeros(n), np.zeros((n, 3))
        rc = fn(C.c_int(t), _d(p), C.c_int(n), _d(r), _d(u), on.ctypes.data_as(_PI), sign.ctypes.data_as(_PI), _d(dist),
                _d(norm))
        assert rc == 0
        out[f"surf_{name}_sign"], out[f"surf_{name}_dist"], out[f"surf_{name}_norm"] = sign, dist, norm
    # directions
    xyz, u, mu, phi = direction_cases()
    n = len(xyz)
    d = np.zeros((n, 3))
    (L.ref_direction if ref else L.orc_direction_probe)(C.c_int(n), _d(xyz), _d(d))
    out["direction"] = d
    d2 = np.zeros((n, 3))
    (L.ref_rotate_direction if ref else L.orc_rotate_direction_probe)(C.c_int(n), _d(u), _d(mu), _d(phi), _d(d2))
    out["rotate_direction"] = d2
    # RNG
    for hid in HISTORY_IDS:
        v = np.zeros(64)
        (L.ref_rng_rand if ref else L.orc_rng_rand)(C.c_uint64(SEED), C.c_uint64(STRIDE), C.c_uint64(hid), C.c_int(64), _d(v))
        out[f"rand_{hid}"] = v
        f = L.ref_rng_exponential if ref else L.orc_rng_exponential
        f.restype = C.c_double
        out[f"exp_{hid}"] = np.array([f(C.c_uint64(SEED), C.c_uint64(STRIDE), C.c_uint64(hid), C.c_double(lam))
                                      for lam in (0.1, 1.0, 2.5, 37.0)])
        for k, w in enumerate(DISCRETE_WEIGHTS):
            w = np.asarray(w, dtype=np.float64)
            draws = np.zeros(200, dtype=np.int32)
            f = L.ref_rng_discrete if ref else L.orc_rng_discrete_probe
            f.restype = C.c_double
            nxt = f(C.c_uint64(SEED), C.c_uint64(STRIDE), C.c_uint64(hid), _d(w), C.c_int(len(w)), C.c_int(200),
                    draws.ctypes.data_as(_PI))
            out[f"discrete_{hid}_{k}"] = draws
            out[f"discrete_next_{hid}_{k}"] = np.array([nxt])
    # Legendre linearisation, then sample_mu from the table the REFERENCE produced (so a table mismatch does not hide
    # behind a sampling mismatch: the golden table is the input of both)
    lin = L.ref_legendre_linearize if ref else L.orc_legendre_linearize_probe
    smp = L.ref_sample_mu if ref else L.orc_sample_mu_probe
    cap = 1 << 16
    for k, a in enumerate(LEGENDRE_CASES):
        a = np.asarray(a, dtype=np.float64)
        mu_t, pdf_t, cdf_t = np.zeros(cap), np.zeros(cap), np.zeros(cap)
        npts = lin(_d(a), C.c_int(len(a)), C.c_int(cap), _d(mu_t), _d(pdf_t), _d(cdf_t))
        assert npts > 1, f"linearize failed for {a}: {npts}"
        out[f"legendre_{k}_mu"], out[f"legendre_{k}_pdf"], out[f"legendre_{k}_cdf"] = mu_t[:npts].copy(), pdf_t[:npts].copy(), \
            cdf_t[:npts].copy()
    # MGNuclide: micro XS, sample_scatter, sample_fission
    fn = L.ref_mg_nuclide if ref else L.orc_mg_nuclide_probe
    for name, a in mg_material_cases():
        G = a["G"]
        micro = np.zeros((G, 6))
        scat, fis = np.zeros((NHIST_MG, NDRAW_MG, 4)), np.zeros((NHIST_MG, NDRAW_MG, 6))
        Pd, lam = (a["Pd"], a["lam"]) if len(a["Pd"]) else (np.zeros(1), np.zeros(1))
        rc = fn(C.c_int(G), _d(a["eb"]), _d(a["Et"]), _d(a["Ea"]), _d(a["Ef"]), _d(a["nup"]),
                _d(a["nud"]) if a["nud"] is not None else None, _d(a["chi"]), _d(a["Es"]), C.c_int(a["nleg"]), _d(a["leg"]),
                C.c_int(len(a["Pd"])), _d(Pd), _d(lam), C.c_uint64(SEED), C.c_uint64(STRIDE), C.c_int(NHIST_MG),
                C.c_int(NDRAW_MG), _d(micro), _d(scat), _d(fis))
        assert rc == 0, f"mg nuclide probe failed for {name}: {rc}"
        out[f"mg_{name}_micro"], out[f"mg_{name}_scatter"], out[f"mg_{name}_fission"] = micro, scat, fis
    # geometry: the Tracker walking the decks' CSG trees (surface-tracking and delta-tracking call sequences)
    from . import deck as _deck
    decks = os.path.join(os.path.dirname(_HERE), "tests", "decks")
    for ci, (fname, low, hi, n, nsteps, flight) in enumerate(GEOMETRY_CASES):
        deck = _deck.load_yaml(os.path.join(decks, fname))
        r, u, d, unew = geometry_rays(deck, low, hi, n, nsteps, flight, 100 + ci)
        ws, wd = np.zeros((n, nsteps + 1, 8)), np.zeros((n, nsteps + 1, 8))
        if ref:
            text, mids = geometry_text(deck)
            assert L.ref_geometry_load(text.encode(), C.c_int(len(mids)), mids.ctypes.data_as(_PI)) == 0
            assert L.ref_geometry_walk_surface(C.c_int(n), _d(r), _d(u), C.c_int(nsteps), _d(ws)) == 0
            assert L.ref_geometry_walk_delta(C.c_int(n), _d(r), _d(u), C.c_int(nsteps), _d(d), _d(unew), _d(wd)) == 0
        else:
            o = api.Oracle(os.path.join(decks, fname))
            assert L.orc_geometry_walk_surface(o.h, C.c_int(n), _d(r), _d(u), C.c_int(nsteps), _d(ws)) == 0, o._err()
            assert L.orc_geometry_walk_delta(o.h, C.c_int(n), _d(r), _d(u), C.c_int(nsteps), _d(d), _d(unew), _d(wd)) == 0, o._err()
            o.close()
        name = fname.split(".")[0]
        out[f"geo_{name}_surface_walk"], out[f"geo_{name}_delta_walk"] = ws, wd
    return out


def sample_mu(impl: str, tables: dict) -> dict:
    """sample_mu over the given (golden) tables: 500 draws of three histories per Legendre case."""
    with _reference_math(impl):
        return _sample_mu(impl, tables)


def _sample_mu(impl: str, tables: dict) -> dict:
    ref = impl == "reference"
    L = ref_lib() if ref else api.lib()
    smp = L.ref_sample_mu if ref else L.orc_sample_mu_probe
    out = {}
    for k in range(len(LEGENDRE_CASES)):
        mu_t, pdf_t, cdf_t = (np.ascontiguousarray(tables[f"legendre_{k}_{q}"]) for q in ("mu", "pdf", "cdf"))
        for hid in HISTORY_IDS[:3]:
            v = np.zeros(500)
            smp(_d(mu_t), _d(pdf_t), _d(cdf_t), C.c_int(len(mu_t)), C.c_uint64(SEED), C.c_uint64(STRIDE), C.c_uint64(hid),
                C.c_int(500), _d(v))
            out[f"sample_mu_{k}_{hid}"] = v
    return out


# ImplicitLeakageDeltaTracker::transport (src/implicit_leakage_delta_tracker.cpp:73-263): vacuum slabs (every flight splits
# its weight into a leaking and a colliding part) and c5g7 (reflective sides, vacuum top / right) with both estimators
IMPLICIT_CASES = (
    ("PUa-1-0-SL_implicit.yaml", 1000, 0.97), ("UD2O-2-1-SL_implicit.yaml", 1000, 1.01),
    ("c5g7_implicit_collision.yaml", 2500, 1.17), ("c5g7_implicit_tracklength.yaml", 1500, 1.17),
)


# transport calls whose fission banks are compared down to the fields the exact cancelators read (tests/golden/ref_pins_exact.npz)
EXACT_TRANSPORT_CASES = (("c5g7_carter_cancel.yaml", 3000, 1.1), ("PUa-cube_carter_exact_min.yaml", 3000, 2.5))


def evaluate_transport(impl: str, cases=None, seed0: int = 500, parents: bool = False) -> dict:
    """One Transporter::transport(bank) per case through the reference's own SurfaceTracker / DeltaTracker / CarterTracker
    (oracle/_ref, one OpenMP thread) or the oracle (glibc math, one thread): the fission bank in the order it is returned
    (9 doubles and parent history id, daughter id, family id per site), the six generation values of
    Tallies::calc_gen_values (k_col, k_abs, k_trk, k_tot, leakage, migration area) and, with settings::converged set, the
    generation scores of every collision / track-length mesh tally of the deck."""
    from . import deck as _deck
    ref = impl == "reference"
    L = ref_lib() if ref else api.lib()
    decks = os.path.join(os.path.dirname(_HERE), "tests", "decks")
    PU = C.POINTER(C.c_uint64)
    out = {}
    with _reference_math(impl):
        for ci, (fname, n, k_col) in enumerate(TRANSPORT_CASES if cases is None else cases):
            path = os.path.join(decks, fname)
            ov = {"settings": {"nparticles": n}}
            deck = _deck.apply_overrides(_deck.load_yaml(path), ov)
            r, u, E, w, hid = transport_bank(deck, n, seed0 + ci, "carter" in fname)
            name = fname.split(".")[0]
            if ref:
                assert L.ref_problem_load(_deck.deck_to_text(deck).encode()) == 0
                cap = 16 * n
                f9, ids, nout, k6 = np.zeros((cap, 9)), np.zeros((cap, 3), dtype=np.uint64), C.c_uint64(0), np.zeros(6)
                rc = L.ref_transport(C.c_uint64(n), _d(r), _d(u), _d(E), _d(w), hid.ctypes.data_as(PU), hid.ctypes.data_as(PU),
                                     C.c_double(k_col), C.c_int(1), C.c_uint64(cap), _d(f9), ids.ctypes.data_as(PU), C.byref(nout), _d(k6))
                assert rc == 0 and nout.value <= cap
                f9, ids = f9[:nout.value].copy(), ids[:nout.value].copy()
                if parents:
                    par = np.zeros((nout.value, 10))
                    L.ref_last_parents.restype = C.c_uint64
                    assert L.ref_last_parents(_d(par), C.c_uint64(nout.value)) == nout.value
                L.ref_tally_size.restype = C.c_uint64
                tallies = []
                for t in range(L.ref_ntallies()):
                    a = np.zeros(int(L.ref_tally_size(C.c_int(t))))
                    if a.size:
                        L.ref_tally_get(C.c_int(t), _d(a))
                    tallies.append(a)
            else:
                o = api.Oracle(path, ov)
                bank = {k: np.ascontiguousarray(v) for k, v in zip(("x", "y", "z"), r.T)}
                bank.update({k: np.ascontiguousarray(v) for k, v in zip(("ux", "uy", "uz"), u.T)})
                bank.update(E=E, wgt=w, wgt2=np.zeros(n), id_a=hid, id_b=hid.copy(), id_c=None)
                o.set_kcol(k_col)
                o.set_converged(True)
                o.tallies_clear()
                fb, scores, m = o.transport(bank, False, capacity=16 * n)
                tallies = [np.ravel(o.tally(t, "gen")) if o.tally_estimator(t) != 2 else np.zeros(0) for t in range(o.ntallies())]
                f9 = np.ascontiguousarray(np.stack([fb[k] for k in api.BANK_F64], 1))
                ids = np.ascontiguousarray(np.stack([fb["id_a"], fb["id_b"], fb["id_c"]], 1))
                k6 = scores / float(n)
                if parents:
                    par = np.concatenate([o.last_parent_info(m), o.last_parent_state(m)], axis=1)
                o.close()
            out[f"transport_{name}_sites"], out[f"transport_{name}_ids"], out[f"transport_{name}_k"] = f9, ids, k6
            if parents:  # parents_previous_position, Esmp_parent, parents_previous_direction, previous previous energy, previous energy, was_virtual
                out[f"transport_{name}_parents"] = par
            for t, a in enumerate(tallies):
                out[f"transport_{name}_tally{t}"] = a
    return out


# noise mode: deck, particles
NOISE_CASES = (("noise_oscillation.yaml", 4000), ("noise_oscillation_delta.yaml", 4000), ("noise_vibration.yaml", 4000))


IMPLICIT_NOISE_CASES = (("noise_oscillation_implicit.yaml", 4000),)
# the flat-vibration source at its third harmonic (tests/golden/ref_pins_vibration.npz, scripts/make_ref_pins_vibration.py)
HARMONIC_NOISE_CASES = (("noise_vibration_h3.yaml", 4000),)


def evaluate_noise(impl: str, cases=None, seed0: int = 900) -> dict:
    """Noise mode through the reference's own code (oracle/_ref) or the oracle.  Per deck two transport calls:
    (A) a power-iteration generation that samples the noise source at its collisions -- transport(bank, noise = false,
        &noise_bank, &noise_maker), NoiseMaker::sample_noise_source with the deck's square-oscillation / flat-vibration
        sources -> fission bank, noise-source bank, generation values;
    (B) a noise generation -- transport(bank, noise = true) over a seeded bank with complex weights -> fission bank
        (complex weights, delayed-neutron factor), generation values."""
    from . import deck as _deck
    ref = impl == "reference"
    L = ref_lib() if ref else api.lib()
    decks = os.path.join(os.path.dirname(_HERE), "tests", "decks")
    PU = C.POINTER(C.c_uint64)
    out = {}
    with _reference_math(impl):
        for ci, (fname, n) in enumerate(NOISE_CASES if cases is None else cases):
            path = os.path.join(decks, fname)
            ov = {"settings": {"nparticles": n}}
            deck = _deck.apply_overrides(_deck.load_yaml(path), ov)
            keff = float(deck["settings"].get("keff", 1.0))
            r, u, E, w, hid = transport_bank(deck, n, seed0 + ci, False)
            rng = np.random.default_rng(seed0 + 50 + ci)
            w2 = rng.uniform(-0.8, 0.8, n)
            name = fname.split(".")[0]
            if ref:
                assert L.ref_problem_load(_deck.deck_to_text(deck).encode()) == 0
            else:
                o = api.Oracle(path, ov)
            for phase, noise, sample, wa, wb in (("A", 0, 1, w, np.zeros(n)), ("B", 1, 0, w, w2)):
                cap = 24 * n
                if ref:
                    f9, fi, b9, bi = np.zeros((cap, 9)), np.zeros((cap, 3), dtype=np.uint64), np.zeros((cap, 9)), np.zeros((cap, 3), dtype=np.uint64)
                    nf, nb, k6 = C.c_uint64(0), C.c_uint64(0), np.zeros(6)
                    rc = L.ref_transport_noise(C.c_uint64(n), _d(r), _d(u), _d(E), _d(wa), _d(wb), hid.ctypes.data_as(PU),
                                               hid.ctypes.data_as(PU), C.c_double(1.0), C.c_double(keff), C.c_int(noise), C.c_int(sample),
                                               C.c_uint64(cap), _d(f9), fi.ctypes.data_as(PU), C.byref(nf), _d(b9), bi.ctypes.data_as(PU),
                                               C.byref(nb), _d(k6))
                    assert rc == 0 and nf.value <= cap and nb.value <= cap
                    f9, fi, b9, bi = f9[:nf.value].copy(), fi[:nf.value].copy(), b9[:nb.value].copy(), bi[:nb.value].copy()
                else:
                    bank = {k: np.ascontiguousarray(v) for k, v in zip(("x", "y", "z"), r.T)}
                    bank.update({k: np.ascontiguousarray(v) for k, v in zip(("ux", "uy", "uz"), u.T)})
                    bank.update(E=E, wgt=np.ascontiguousarray(wa), wgt2=np.ascontiguousarray(wb), id_a=hid, id_b=hid.copy(), id_c=None)
                    o.set_kcol(1.0)
                    o.set_keff(keff)
                    o.set_converged(False)
                    o.tallies_clear()
                    fb, nbk, scores = o.transport_noise(bank, bool(noise), bool(sample), capacity=cap)
                    st = lambda b: (np.ascontiguousarray(np.stack([b[k] for k in api.BANK_F64], 1)).reshape(-1, 9),  # noqa: E731
                                    np.ascontiguousarray(np.stack([b["id_a"], b["id_b"], b["id_c"]], 1)).reshape(-1, 3))
                    (f9, fi), (b9, bi) = st(fb), st(nbk)
                    k6 = scores / float(n)
                out[f"noise_{name}_{phase}_sites"], out[f"noise_{name}_{phase}_ids"] = f9, fi
                out[f"noise_{name}_{phase}_source"], out[f"noise_{name}_{phase}_source_ids"] = b9, bi
                out[f"noise_{name}_{phase}_k"] = k6
            if not ref:
                o.close()
    return out


# whole k-eigenvalue simulations: deck, particles, generations, ignored generations
POWER_ITERATION_CASES = (
    ("PUa-1-0-SL.yaml", 2000, 12, 4), ("PUa-1-0-IN.yaml", 2000, 10, 3), ("Ua-1-1-CY.yaml", 2000, 10, 3), ("UD2O-2-1-SL.yaml", 2000, 10, 3),
    ("c5g7_delta_collision.yaml", 3000, 8, 3), ("c5g7_carter_cancel.yaml", 3000, 8, 3), ("ref_sqr_c5g7_surface_tl.yaml", 2000, 6, 2),
)


# whole simulations with the fourth tracker (golden vectors in tests/golden/ref_pins_implicit.npz): addressed by `only` = an
# index into ALL_PI_CASES; `only` = None keeps meaning the cases of tests/golden/ref_pins.npz
IMPLICIT_POWER_ITERATION_CASES = (("c5g7_implicit_collision.yaml", 3000, 8, 3), ("PUa-1-0-SL_implicit.yaml", 2000, 12, 4))
# branchless-k-eigenvalue (src/branchless_power_iterator.cpp + Transporter::branchless_collision_{mat,iso}): material flavour with the
# comb, isotope flavour with splitting and no comb, material flavour with splitting and the comb (tests/golden/ref_pins_branchless.npz)
BRANCHLESS_PI_CASES = (("c5g7_delta_branchless.yaml", 3000, 8, 3), ("PUa-1-0-SL_branchless_iso_split.yaml", 2000, 10, 3),
                       ("UD2O-2-1-SL_branchless_split_comb.yaml", 2000, 10, 3))
# carter tracking with negative weights and the basic-exact regional cancelator (src/basic_exact_mg_cancelator.cpp): beta minimum
# in a reflective cube, average-f with sampled points and average-g with Sobol points in a vacuum cube
# (tests/golden/ref_pins_exact.npz).  One material each: the reference orders its bins by Material pointer otherwise.
EXACT_PI_CASES = (("PUa-cube_carter_exact_min.yaml", 2000, 8, 3), ("PUa-cube_carter_exact_avgf.yaml", 2000, 8, 3),
                  ("PUa-cube_carter_exact_avgg.yaml", 2000, 8, 3))
# ... and with the reference's second exact cancelator (`type: exact`, src/exact_mg_cancelator.cpp): one-group cube (which counts as a
# chi matrix: energy bins, the fourth Sobol dimension) and c5g7 (chi vector, several materials per mesh cell)
EXACT_FULL_PI_CASES = (("PUa-cube_carter_exact_full.yaml", 2000, 8, 3), ("c5g7_carter_exact_full.yaml", 3000, 6, 2))
ALL_PI_CASES = POWER_ITERATION_CASES + IMPLICIT_POWER_ITERATION_CASES + BRANCHLESS_PI_CASES + EXACT_PI_CASES + EXACT_FULL_PI_CASES
IMPLICIT_PI_RANGE = range(len(POWER_ITERATION_CASES), len(POWER_ITERATION_CASES) + len(IMPLICIT_POWER_ITERATION_CASES))
BRANCHLESS_PI_RANGE = range(IMPLICIT_PI_RANGE.stop, IMPLICIT_PI_RANGE.stop + len(BRANCHLESS_PI_CASES))
EXACT_PI_RANGE = range(BRANCHLESS_PI_RANGE.stop, BRANCHLESS_PI_RANGE.stop + len(EXACT_PI_CASES))
EXACT_FULL_PI_RANGE = range(EXACT_PI_RANGE.stop, len(ALL_PI_CASES))


def pi_golden_file(ci: int) -> str:
    """The file under tests/golden/ that holds the reference's output for ALL_PI_CASES[ci]."""
    if ci in EXACT_PI_RANGE or ci in EXACT_FULL_PI_RANGE:
        return "ref_pins_exact.npz"
    return "ref_pins_branchless.npz" if ci in BRANCHLESS_PI_RANGE else ("ref_pins_implicit.npz" if ci in IMPLICIT_PI_RANGE else "ref_pins.npz")


def evaluate_power_iteration(impl: str, only: int | None = None) -> dict:
    """The reference's own PowerIterator::initialize() + run() (oracle/_ref: source sampling through Source / Box / Point /
    Isotropic / MonoEnergetic, transport, Entropy, ApproximateMeshCancelator, weight normalisation, history-id hand-out,
    Tallies statistics, MeshTally::record_generation) against the oracle's driver: per generation k_col, k_trk, leakage,
    migration area and entropy; the final averages and errors; the average and the error of the mean of the deck's
    collision / track-length mesh tallies as MeshTally::write_tally leaves them."""
    from . import deck as _deck
    ref = impl == "reference"
    L = ref_lib() if ref else api.lib()
    decks = os.path.join(os.path.dirname(_HERE), "tests", "decks")
    keys = ("kcol", "ktrk", "leak", "mig", "entropy")
    out = {}
    if ref and only is None:
        # The reference is a one-simulation-per-process program (settings, MPI bookkeeping and id counters are process
        # globals that PowerIterator::run leaves changed): every case gets a fresh process, as the reference itself would.
        import subprocess
        import sys
        import tempfile
        for i in range(len(POWER_ITERATION_CASES)):
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "pi.npz")
                code = (f"import sys; sys.path.insert(0, {os.path.dirname(_HERE)!r}); import numpy as np; "
                        f"from oracle import ref_pins; np.savez({path!r}, **ref_pins.evaluate_power_iteration('reference', only={i}))")
                subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
                out.update(dict(np.load(path)))
        return out
    with _reference_math(impl):
        for fname, n, ngen, nign in (POWER_ITERATION_CASES if only is None else ALL_PI_CASES[only:only + 1]):
            path = os.path.join(decks, fname)
            ov = {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}}
            name = fname.split(".")[0]
            tal = {}
            if ref:
                deck = _deck.apply_overrides(_deck.load_yaml(path), ov)
                a = {k: np.zeros(ngen) for k in keys}
                summ = np.zeros(6)
                L.ref_set_threads(C.c_int(1))
                rc = L.ref_power_iteration(_deck.deck_to_text(deck).encode(), C.c_int(ngen), C.c_int(nign), *[_d(a[k]) for k in keys],
                                           _d(summ))
                assert rc == 0
                L.ref_tally_size.restype = C.c_uint64
                for t in range(L.ref_ntallies()):
                    size = int(L.ref_tally_size(C.c_int(t)))
                    if size:
                        for which, wname in ((1, "avg"), (2, "std")):  # write_tally has turned var into the error of the mean
                            v = np.zeros(size)
                            L.ref_tally_get_stat(C.c_int(t), C.c_int(which), _d(v))
                            tal[f"tally{t}_{wname}"] = v
            else:
                o = api.Oracle(path, ov)
                r = o.run_power_iteration(ngen, nign)
                a = {k: r[k] for k in keys}
                summ = np.array([r[k] for k in ("kcol_avg", "kcol_err", "ktrk_avg", "ktrk_err", "leak_avg", "leak_err")])
                for t in range(o.ntallies()):
                    if o.tally_estimator(t) != 2:
                        tal[f"tally{t}_avg"], tal[f"tally{t}_std"] = np.ravel(o.tally(t, "avg")), np.ravel(o.tally(t, "std"))
                o.close()
            for k in keys:
                out[f"pi_{name}_{k}"] = np.ascontiguousarray(a[k])
            out[f"pi_{name}_summary"] = summ
            for k, v in tal.items():
                out[f"pi_{name}_{k}"] = np.ascontiguousarray(v)
    return out


# the power iteration's optional diagnostics: deck, particles, generations, ignored generations (carter: negative weights)
DIAG_CASES = (("c5g7_delta_collision.yaml", 1500, 5, 2), ("c5g7_carter_cancel.yaml", 1500, 4, 1))


def evaluate_pi_diagnostics(only: int) -> dict:
    """The reference's own PowerIterator::run() with settings pair-distance-sqrd, families and empty-entropy-bins on
    (src/power_iterator.cpp:283-297,326-331,362-365,613-615,637-663): the three per-generation series, k_col, and the final
    bank [n, 4] = x y z wgt the last pair distance was taken over.  One case per process (oracle/_ref only)."""
    from . import deck as _deck
    L = ref_lib()
    fname, n, ngen, nign = DIAG_CASES[only]
    decks = os.path.join(os.path.dirname(_HERE), "tests", "decks")
    deck = _deck.apply_overrides(_deck.load_yaml(os.path.join(decks, fname)), {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}})
    keys = ("kcol", "ktrk", "leak", "mig", "entropy")
    a = {k: np.zeros(ngen) for k in keys}
    summ = np.zeros(6)
    L.ref_set_threads(C.c_int(1))
    L.ref_set_diagnostics(C.c_int(1), C.c_int(1), C.c_int(1))
    rc = L.ref_power_iteration(_deck.deck_to_text(deck).encode(), C.c_int(ngen), C.c_int(nign), *[_d(a[k]) for k in keys], _d(summ))
    assert rc == 0
    ser = [np.zeros(ngen) for _ in range(3)]
    n3 = (C.c_uint64 * 3)()
    L.ref_pi_diagnostics(_d(ser[0]), _d(ser[1]), _d(ser[2]), C.c_uint64(ngen), n3)
    assert list(n3) == [ngen, ngen, ngen]
    L.ref_last_bank_get.restype = C.c_uint64
    bank = np.zeros((4 * n, 4))
    nb = int(L.ref_last_bank_get(_d(bank), C.c_uint64(len(bank))))
    assert 0 < nb <= len(bank)
    name = fname.split(".")[0]
    return {f"diag_{name}_r_sqrd": ser[0], f"diag_{name}_families": ser[1], f"diag_{name}_empty": ser[2], f"diag_{name}_kcol": a["kcol"],
            f"diag_{name}_entropy": a["entropy"], f"diag_{name}_bank": np.ascontiguousarray(bank[:nb])}


# whole noise simulations: deck, particles, noise batches, ignored generations, nskip
NOISE_DRIVER_CASES = (("noise_oscillation.yaml", 600, 2, 1, 2), ("noise_oscillation_delta.yaml", 500, 2, 2, 1),
                      ("noise_vibration.yaml", 500, 2, 1, 2))


# modified-fixed-source: deck, particles per batch, batches
MFS_CASES = (("PUa-1-0-SL_subcritical_mfs.yaml", 3000, 6),)
# fixed-source (fission neutrons as secondaries): deck, particles per batch, batches
FS_CASES = (("PUa-1-0-SL_subcritical_fs.yaml", 3000, 6),)
# the same driver with beam sources (mono-directional, cone from a box, cone about the pole, isotropic; four sources picked by weight)
BEAM_CASES = (("PUa-1-0-SL_subcritical_fs_beam.yaml", 3000, 6),
              # ... and with energy distributions: Maxwellian, Watt with a rejected tail, mono-energetic, on a two-group slab
              ("UD2O-2-1-SL_subcritical_fs_spectra.yaml", 3000, 6))


def evaluate_fixed_source(impl: str) -> dict:
    """The reference's own FixedSource::run() (oracle/_ref) against the oracle's driver: k_col, leakage and migration area of
    every batch and the mesh tallies' average and error of the mean."""
    return evaluate_modified_fixed_source(impl, FS_CASES, "fs")


def evaluate_beam_sources(impl: str, only: int | None = None) -> dict:
    """FixedSource::run() over sources with MonoDirectional and Cone direction distributions (src/mono_directional.cpp,
    src/cone.cpp, include/simulation/mono_directional.hpp:38) and Maxwellian and Watt energy distributions (src/maxwellian.cpp,
    src/watt.cpp), the reference's own against the oracle's.  `only`: one case (the reference keeps its state in process
    globals: one case per process)."""
    return evaluate_modified_fixed_source(impl, BEAM_CASES if only is None else BEAM_CASES[only:only + 1], "fs")


def evaluate_modified_fixed_source(impl: str, cases=None, kind: str = "mfs") -> dict:
    """The reference's own ModifiedFixedSource::run() (oracle/_ref) against the oracle's driver: k_col, leakage and migration
    area of every batch, the histories transported (source + every fission generation of every chain), and the mesh tallies'
    average and error of the mean.  One case: the reference keeps its state in process globals."""
    from . import deck as _deck
    ref = impl == "reference"
    decks = os.path.join(os.path.dirname(_HERE), "tests", "decks")
    out = {}
    with _reference_math(impl):
        for fname, n, nb in (MFS_CASES if cases is None else cases):
            path = os.path.join(decks, fname)
            ov = {"settings": {"nparticles": n, "ngenerations": nb}}
            name = fname.split(".")[0]
            if ref:
                L = ref_lib()
                deck = _deck.apply_overrides(_deck.load_yaml(path), ov)
                a = {k: np.zeros(nb) for k in ("kcol", "leak", "mig")}
                tr = C.c_uint64(0)
                L.ref_set_threads(C.c_int(1))
                if kind == "mfs":
                    rc = L.ref_modified_fixed_source(_deck.deck_to_text(deck).encode(), C.c_int(nb), _d(a["kcol"]), _d(a["leak"]),
                                                     _d(a["mig"]), C.byref(tr))
                else:
                    rc = L.ref_fixed_source(_deck.deck_to_text(deck).encode(), C.c_int(nb), _d(a["kcol"]), _d(a["leak"]), _d(a["mig"]))
                assert rc == 0
                # (tr is 0 in the reference: it adds bank.size() after transport() has cleared the bank; not compared)
                L.ref_tally_size.restype = C.c_uint64
                for t in range(L.ref_ntallies()):
                    size = int(L.ref_tally_size(C.c_int(t)))
                    for which, wname in ((1, "avg"), (2, "std")):
                        v = np.zeros(size)
                        L.ref_tally_get_stat(C.c_int(t), C.c_int(which), _d(v))
                        a[f"tally{t}_{wname}"] = v
            else:
                o = api.Oracle(path, ov)
                r = o.run_modified_fixed_source(nb) if kind == "mfs" else o.run_fixed_source(nb)
                a = {k: r[k] for k in ("kcol", "leak", "mig")}
                for t in range(o.ntallies()):
                    a[f"tally{t}_avg"], a[f"tally{t}_std"] = np.ravel(o.tally(t, "avg")), np.ravel(o.tally(t, "std"))
                o.close()
            for k, v in a.items():
                out[f"{kind}_{name}_{k}"] = np.ascontiguousarray(v)
    return out


def evaluate_noise_driver(impl: str, only: int | None = None) -> dict:
    """The reference's own Noise::initialize() + run() (src/noise.cpp: power-iteration generations, noise-source sampling
    and normalisation, inner noise generations with regional cancellation of the noise fission banks, tally statistics)
    against the oracle's driver (oracle/api.py run_noise): k_col of every power-iteration generation, the final bank size /
    first history id / history counter, and average and error of the mean of every mesh tally (noise-source tallies,
    real / imaginary flux)."""
    from . import deck as _deck
    ref = impl == "reference"
    out = {}
    if ref and only is None:  # one simulation per process, as for the power iteration
        import subprocess
        import sys
        import tempfile
        for i in range(len(NOISE_DRIVER_CASES)):
            with tempfile.TemporaryDirectory() as td:
                path = os.path.join(td, "nd.npz")
                code = (f"import sys; sys.path.insert(0, {os.path.dirname(_HERE)!r}); import numpy as np; "
                        f"from oracle import ref_pins; np.savez({path!r}, **ref_pins.evaluate_noise_driver('reference', only={i}))")
                subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
                out.update(dict(np.load(path)))
        return out
    L = ref_lib() if ref else api.lib()
    decks = os.path.join(os.path.dirname(_HERE), "tests", "decks")
    with _reference_math(impl):
        for fname, n, nb, nign, nskip in (NOISE_DRIVER_CASES if only is None else NOISE_DRIVER_CASES[only:only + 1]):
            path = os.path.join(decks, fname)
            ov = {"settings": {"nparticles": n, "ngenerations": nb, "nignored": nign, "nskip": nskip}}
            deck = _deck.apply_overrides(_deck.load_yaml(path), ov)
            name = fname.split(".")[0]
            if ref:
                kc, nk, fb3 = np.zeros(4096), C.c_int(0), np.zeros(3, dtype=np.uint64)
                L.ref_set_threads(C.c_int(1))
                rc = L.ref_noise_run(_deck.deck_to_text(deck).encode(), C.c_int(nb), C.c_int(nign), C.c_int(nskip), _d(kc), C.byref(nk),
                                     fb3.ctypes.data_as(C.POINTER(C.c_uint64)))
                assert rc == 0
                kcol = kc[:nk.value]
                kcol = kcol[kcol != 0.0]  # Tallies::calc_gen_values after a noise batch appends a 0 (no k score in noise mode)
                o = api.Oracle(path, ov)  # only for the tally shapes
                tal = []
                for t in range(o.ntallies()):
                    size = int(np.prod(o.tally_shape(t)))
                    a, e = np.zeros(size), np.zeros(size)
                    L.ref_tally_get_stat(C.c_int(t), C.c_int(1), _d(a))
                    L.ref_tally_get_stat(C.c_int(t), C.c_int(2), _d(e))
                    tal.append((a, e))
                o.close()
            else:
                o = api.Oracle(path, ov)
                r = o.run_noise(deck["settings"])
                kcol, fb3 = np.asarray(r["k_col"]), np.asarray(r["final_bank"], dtype=np.uint64)
                tal = [(np.ravel(o.tally(t, "avg")), np.ravel(o.tally(t, "std"))) for t in range(o.ntallies())]
                o.close()
            out[f"nd_{name}_kcol"], out[f"nd_{name}_final_bank"] = np.ascontiguousarray(kcol), fb3
            for t, (a, e) in enumerate(tal):
                out[f"nd_{name}_tally{t}_avg"], out[f"nd_{name}_tally{t}_std"] = a, e
    return out


def flatten_dump(yaml_deck: str) -> dict:
    """integration/flatten_problem.hpp on the reference's live objects for this deck (oracle/_ref): the tables of abl_problem, in
    the layout of abeille_b200.dump_tables (which flattens the same deck through this repo's host library)."""
    from . import deck as _deck
    L = ref_lib()
    out = C.create_string_buffer(1 << 24)
    rc = L.ref_flatten_dump(_deck.deck_to_text(_deck.load_yaml(yaml_deck)).encode(), out, C.c_longlong(1 << 24))
    assert rc == 0, rc
    tables = {}
    for line in out.value.decode().splitlines():
        name, *vals = line.split()
        tables[name] = np.array([float(v) for v in vals])
    return tables


def power_iteration_through_gpu_transporter(only: int, host_library: str, yaml_deck: str, device: int = 0, from_objects: bool = False) -> dict:
    """The reference's own PowerIterator::run() with its transporter replaced by GPUTransporter
    (integration/gpu_transporter.hpp): the drop-in of INTEGRATION.md, live.  Needs a GPU; one case per process.
    from_objects: `host_library` is libabeille_b200.so and the problem tables come from flatten_problem() on the reference's live
    objects instead of this repo's host library parsing the YAML deck."""
    from . import deck as _deck
    fname, n, ngen, nign = ALL_PI_CASES[only]
    deck = _deck.load_yaml(yaml_deck)
    if from_objects:
        yaml_deck = ""
    L = ref_lib()
    keys = ("kcol", "ktrk", "leak", "mig", "entropy")
    a = {k: np.zeros(ngen) for k in keys}
    summ = np.zeros(6)
    L.ref_set_threads(C.c_int(1))
    rc = L.ref_power_iteration_gpu(_deck.deck_to_text(deck).encode(), host_library.encode(), yaml_deck.encode(), C.c_int(device),
                                   C.c_int(ngen), C.c_int(nign), *[_d(a[k]) for k in keys], _d(summ))
    assert rc == 0
    name = fname.split(".")[0]
    out = {f"pi_{name}_{k}": a[k] for k in keys}
    out[f"pi_{name}_summary"] = summ
    L.ref_gpu_tally_get.restype = C.c_uint64
    for t in range(L.ref_gpu_ntallies()):  # the device's mesh tallies, recorded / cleared by the adapter between generations
        size = int(L.ref_gpu_tally_get(C.c_int(t), C.c_int(1), None, C.c_uint64(0)))
        for which, wname in ((1, "avg"), (3, "std")):
            v = np.zeros(size)
            L.ref_gpu_tally_get(C.c_int(t), C.c_int(which), _d(v), C.c_uint64(size))
            out[f"pi_{name}_tally{t}_{wname}"] = v
    L.ref_gpu_release()
    return out


def noise_through_gpu_transporter(only: int, host_library: str, yaml_deck: str, device: int = 0, from_objects: bool = False) -> dict:
    """The reference's own Noise::run() with its transporter replaced by GPUTransporter (integration/gpu_transporter.hpp):
    transport(bank), transport(bank, false, &noise_bank, &noise_maker) and transport(nbank, true) all through the C ABI.
    Needs a GPU; one case per process."""
    from . import deck as _deck
    fname, n, nb, nign, nskip = NOISE_DRIVER_CASES[only]
    deck = _deck.load_yaml(yaml_deck)
    if from_objects:  # host_library is libabeille_b200.so; flatten_problem() on the reference's live objects
        yaml_deck = ""
    L = ref_lib()
    kc, nk, fb3 = np.zeros(4096), C.c_int(0), np.zeros(3, dtype=np.uint64)
    L.ref_set_threads(C.c_int(1))
    rc = L.ref_noise_run_gpu(_deck.deck_to_text(deck).encode(), host_library.encode(), yaml_deck.encode(), C.c_int(device), C.c_int(nb),
                             C.c_int(nign), C.c_int(nskip), _d(kc), C.byref(nk), fb3.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == 0
    kcol = kc[:nk.value]
    name = fname.split(".")[0]
    L.ref_gpu_release()
    return {f"nd_{name}_kcol": np.ascontiguousarray(kcol[kcol != 0.0]), f"nd_{name}_final_bank": fb3}


def sobol_points(impl: str, n: int = 5000) -> np.ndarray:
    """The first n points of the 4-d Sobol sequence the exact cancelators sample their bins with (three coordinates; ExactMGCancelator
    takes the energy group from the fourth): the reference's vendored table (vendor/sobol) or the oracle's matrices generated from
    the Joe-Kuo recurrence."""
    out = np.zeros((n, 4))
    L = ref_lib() if impl == "reference" else api.lib()
    getattr(L, "ref_sobol_points" if impl == "reference" else "orc_sobol_points")(C.c_int(n), _d(out))
    return out
