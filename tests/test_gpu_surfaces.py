"""Every surface class on the device: sign / distance / norm of the nine types against the reference's own outputs
(tests/golden/ref_pins.npz `surf_*`: plane.cpp:33-60, sphere.cpp:33-80, cylinder.cpp:60-110, x/y/zcylinder.cpp,
x/y/zplane.cpp compiled from /root/reference by oracle/Makefile), and a deck whose cells are cut by a sphere, a general
cylinder, axis cylinders and general planes driven through Transporter::transport bit-exact against the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_deck, write_deck

pytestmark = pytest.mark.gpu

YAML_TYPE = {"xplane": "xplane", "yplane": "yplane", "zplane": "zplane", "plane": "plane", "xcyl": "xcylinder",
             "ycyl": "ycylinder", "zcyl": "zcylinder", "cyl": "cylinder", "sphere": "sphere"}
YAML_KEYS = {"xplane": ("x0",), "yplane": ("y0",), "zplane": ("z0",), "plane": ("A", "B", "C", "D"),
             "xcyl": ("y0", "z0", "r"), "ycyl": ("x0", "z0", "r"), "zcyl": ("x0", "y0", "r"),
             "cyl": ("x0", "y0", "z0", "u0", "v0", "w0", "r"), "sphere": ("x0", "y0", "z0", "r")}


@pytest.fixture(scope="module")
def ab(native_libs):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return native_libs


def _unit(u):
    """Direction(x, y, z) (include/utils/direction.hpp:37-43) in the same IEEE operations."""
    m = np.sqrt(u[:, 0] * u[:, 0] + u[:, 1] * u[:, 1] + u[:, 2] * u[:, 2])
    return np.ascontiguousarray(u / m[:, None])


def _nine_surface_deck(cases):
    deck = load_deck("PUa-1-0-IN.yaml")
    surfs = list(deck["surfaces"])
    first = len(surfs)
    for t, name, p, r, u, on in cases:
        s = {"type": YAML_TYPE[name], "id": 100 + t}
        for k, key in enumerate(YAML_KEYS[name]):
            s[key] = float(p[k])
        surfs.append(s)
    deck["surfaces"] = surfs
    return deck, first


def test_nine_surface_types_match_the_reference(ab, oracle_api, tmp_path):
    from oracle import ref_pins
    golden = np.load(os.path.join(GOLDEN, "ref_pins.npz"))
    cases = ref_pins.surface_cases()
    deck, first = _nine_surface_deck(cases)
    gpu = ab.Backend(write_deck(deck, tmp_path / "nine.yaml"), 0)
    L = oracle_api.lib()
    import ctypes as C
    _d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _i = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    for t, name, p, r, u, on in cases:
        uu = _unit(u)
        sign, dist, norm = gpu.surface_probe(first + t, r, uu, on)
        # the oracle (pinned bit for bit to the reference in tests/test_reference_pins.py): only sqrt and division, so
        # the device reproduces it to the last bit
        n = len(r)
        osign = np.zeros(n, dtype=np.int32)
        odist, onorm = np.zeros(n), np.zeros((n, 3))
        assert L.orc_surface_probe(C.c_int(t), _d(p), C.c_int(n), _d(r), _d(u), _i(on), _i(osign), _d(odist), _d(onorm)) == 0
        assert np.array_equal(sign, osign), name
        assert np.array_equal(dist, odist), name
        assert np.array_equal(norm, onorm), name
        # and the reference's own outputs directly
        assert np.array_equal(sign, golden[f"surf_{name}_sign"]), name
        assert np.array_equal(dist, golden[f"surf_{name}_dist"]), name
        assert np.array_equal(norm, golden[f"surf_{name}_norm"]), name
        finite = dist < 1e300
        assert finite.sum() > n // 10 and (~finite).sum() > n // 10, name  # both branches of distance()


def mixed_surface_deck(transport, tally_estimator):
    """Fuel sphere inside a tilted (general) cylinder of water inside the intersection of an x- and a y-cylinder of MOX,
    the rest water (a union region: RPN evaluation), all inside a reflective sphere cut by two vacuum planes."""
    d = load_deck("c5g7_delta_collision.yaml")
    d["surfaces"] = [
        {"type": "sphere", "x0": 0.3, "y0": -0.2, "z0": 0.1, "r": 2.0, "id": 1},
        {"type": "cylinder", "x0": 0.0, "y0": 0.0, "z0": 0.0, "u0": 1.0, "v0": 2.0, "w0": 3.0, "r": 3.5, "id": 2},
        {"type": "xcylinder", "y0": 0.2, "z0": -0.1, "r": 5.0, "id": 3},
        {"type": "ycylinder", "x0": 0.1, "z0": 0.3, "r": 5.5, "id": 4},
        {"type": "plane", "A": 1.0, "B": 1.0, "C": 1.0, "D": 8.0, "boundary": "vacuum", "id": 5},
        {"type": "plane", "A": -1.0, "B": -0.5, "C": -1.0, "D": 8.0, "boundary": "vacuum", "id": 6},
        {"type": "sphere", "x0": 0.0, "y0": 0.0, "z0": 0.0, "r": 7.0, "boundary": "reflective", "id": 7},
    ]
    mats = [m["id"] for m in d["materials"]]
    fuel, mox, water = mats[0], mats[2], mats[-1]
    d["cells"] = [
        {"id": 1, "name": "fuel", "region": "-1", "material": fuel},
        {"id": 2, "name": "sleeve", "region": "+1 & -2 & -7 & -5 & -6", "material": water},
        {"id": 3, "name": "cross", "region": "+2 & -3 & -4 & -7 & -5 & -6", "material": mox},
        {"id": 4, "name": "rest", "region": "+2 & (+3 U +4) & -7 & -5 & -6", "material": water},
    ]
    d["universes"] = [{"id": 1, "name": "all", "cells": [1, 2, 3, 4]}]
    d["root-universe"] = 1
    d["sources"] = [{"spatial": {"type": "box", "low": [-1.0, -1.2, -0.9], "hi": [1.3, 0.9, 1.1]},
                     "direction": {"type": "isotropic"}, "energy": {"type": "mono-energetic", "energy": 6.5}, "weight": 1.0}]
    d["tallies"] = [{"name": "flux", "low": [-7.0, -7.0, -7.0], "hi": [7.0, 7.0, 7.0], "shape": [28, 28, 28],
                     "energy-bounds": [0, 1, 2, 3, 4, 5, 6, 7], "quantity": "flux", "estimator": tally_estimator}]
    d.pop("entropy", None)
    d.pop("cancelator", None)
    d["settings"]["transport"] = transport
    d["settings"].pop("cancellation", None)
    return d


@pytest.mark.parametrize("transport,estimator", [("surface-tracking", "track-length"), ("delta-tracking", "collision"),
                                                 ("carter-tracking", "track-length")])
def test_mixed_surface_deck_histories_bit_exact(ab, oracle_api, tmp_path, transport, estimator):
    from test_gpu_parity import _assert_same_histories, _transport_both
    deck = mixed_surface_deck(transport, estimator)
    if transport == "carter-tracking":
        deck["sampling-xs-ratio"] = [0.9, 1, 1, 1, 1, 1, 1]
    path = write_deck(deck, tmp_path / "mixed.yaml", {"settings": {"nparticles": 6000}})
    orc, gpu = oracle_api.Oracle(path), ab.Backend(path, 0)
    for second in (False, True):
        orc.set_converged(False)  # (the oracle's extra first generation of a second-generation case must not score)
        bank, o, g = _transport_both(orc, gpu, 6000, converged=True, k_col=1.05, second_generation=second)
        _assert_same_histories(o, g)
        assert o[3]["boundary_events"] > 200  # reflections off the sphere and leaks through the planes both happen
        og, gg = orc.tally(0, "gen"), gpu.tally(0, "gen")
        assert og.max() > 0 and np.allclose(gg, og, rtol=1e-9, atol=1e-12 * np.abs(og).max())
        orc.tallies_clear()
        gpu.tallies_clear()


def test_mixed_surface_deck_find_cells(ab, oracle_api, tmp_path):
    path = write_deck(mixed_surface_deck("surface-tracking", "collision"), tmp_path / "mixed.yaml")
    orc, gpu = oracle_api.Oracle(path), ab.Backend(path, 0)
    rng = np.random.default_rng(11)
    n = 100000
    r = rng.uniform(-7.5, 7.5, (n, 3))
    # points on the sphere / the tilted cylinder / the planes: the direction-dependent tie-breaks
    r[:5000] = np.array([0.3, -0.2, 0.1]) + 2.0 * _unit(rng.normal(size=(5000, 3)))
    q = r[5000:8000]
    q[:, 2] = 8.0 - q[:, 0] - q[:, 1]
    u = _unit(rng.normal(size=(n, 3)))
    oc, om = orc.find_cells(r, u)
    gc, gm = gpu.find_cells(r, u)
    assert np.array_equal(gc, oc) and np.array_equal(gm, om)
    assert len(set(oc.tolist())) == 5  # four cells and "outside"
