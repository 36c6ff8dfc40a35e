"""GPU: parity of the CUDA path against the CPU oracle, called through the C ABI (abl_*) and through the C++
GPUTransporter adapter.  Integer / index / byte results are compared bit for bit; floating-point sums whose
order differs between a serial CPU loop and GPU atomics (scores, mesh bins) to 1e-10 relative."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, deck_path, load_deck, write_deck

pytestmark = pytest.mark.gpu

BANK_EXACT = ("x", "y", "z", "ux", "uy", "uz", "E", "wgt", "id_a", "id_b", "id_c")


@pytest.fixture(scope="module")
def ab(native_libs):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return native_libs


def _pair(ab, oracle_api, tmp_path, deck, overrides, name="deck.yaml"):
    path = write_deck(load_deck(deck), tmp_path / name, overrides)
    return oracle_api.Oracle(path), ab.Backend(path, 0)


def _next_bank(fis, first_id):
    """Fission bank -> particle bank of the next generation (power_iterator.cpp:386-404): fresh history ids,
    family kept, RNG streams from seed / stride / history id."""
    m = len(fis["x"])
    nxt = {k: fis[k].copy() for k in ("x", "y", "z", "ux", "uy", "uz", "E", "wgt")}
    nxt["wgt2"] = np.zeros(m)
    nxt["id_a"] = np.arange(first_id, first_id + m, dtype=np.uint64)
    nxt["id_b"] = fis["id_c"].copy()
    nxt["id_c"] = None
    return nxt


def _transport_both(orc, gpu, n, converged=True, k_col=1.0, second_generation=False):
    bank = orc.sample_source(n)
    if second_generation:  # fission sites are born fast: reaches the groups the source never visits
        fis, _, _ = orc.transport({k: v.copy() for k, v in bank.items()})
        bank = _next_bank(fis, n)
        n = len(bank["x"])
    orc.set_trace(True)
    orc.set_converged(converged)
    orc.set_kcol(k_col)
    orc.reset_counters()
    ofis, oscores, on = orc.transport({k: (v.copy() if v is not None else None) for k, v in bank.items()})
    otr = orc.trace(n)
    gfis, gscores, gcn = gpu.transport(bank, k_col=k_col, converged=converged, trace=True)
    gtr = gpu.trace(n)
    return bank, (ofis, oscores, otr, orc.counters()), (gfis, gscores, gtr, gcn)


def _assert_same_histories(o, g):
    ofis, oscores, otr, ocn = o
    gfis, gscores, gtr, gcn = g
    for k in ("flights", "real", "virtual", "fission", "hash", "rng_state"):
        bad = np.nonzero(gtr[k] != otr[k])[0]
        assert bad.size == 0, f"per-history {k} differs for {bad.size} histories, first {bad[:5]}"
    assert len(gfis["x"]) == len(ofis["x"])
    for k in BANK_EXACT:
        assert np.array_equal(gfis[k], ofis[k]), f"fission bank field {k} differs"
    for k in ("flights", "real_collisions", "virtual_collisions", "tl_bins", "fission_sites", "boundary_events",
              "lost_at_birth", "coll_scores"):
        assert gcn[k] == ocn[k], (k, gcn[k], ocn[k])
    assert np.allclose(gscores, oscores, rtol=1e-11, atol=1e-300), (gscores, oscores)


# ---- building blocks ------------------------------------------------------------------------------------------
def test_rng_streams_bit_exact(ab, oracle_api, tmp_path):
    with open(os.path.join(GOLDEN, "rng_kat.json")) as f:
        kat = json.load(f)
    gpu = ab.Backend(deck_path("PUa-1-0-IN.yaml"), 0)
    for h in kat["histories"]:
        u32, rnd = gpu.rng_probe(h["id"], 8)
        assert list(u32) == h["u32"]
        assert np.array_equal(rnd[:6], np.array(h["rand"]))
    for hid in (3, 99999, 2**40 + 17, 2**63 + 5):
        u32, rnd = gpu.rng_probe(hid, 64)
        assert np.array_equal(u32, oracle_api.rng_stream(kat["seed"], kat["stride"], hid, 64))
        assert np.array_equal(rnd, oracle_api.rng_rand(kat["seed"], kat["stride"], hid, 64))


def test_device_math_bit_exact(ab, oracle_api):
    gpu = ab.Backend(deck_path("PUa-1-0-IN.yaml"), 0)
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(1e-300, 1.0, 200000), rng.uniform(0.0, 2 * np.pi, 200000),
                        [1.0, 0.5, 2 * np.pi, np.pi, np.pi / 2, np.pi / 4, 1e-9, 2.0 ** -30, 0.7853981633974483]])
    lg, sn, cs = gpu.math_probe(x)
    olg, osn, ocs = oracle_api.math_eval(x)
    assert np.array_equal(lg, olg) and np.array_equal(sn, osn) and np.array_equal(cs, ocs)


@pytest.mark.parametrize("deck,box", [
    ("c5g7_delta_collision.yaml", ([-33, -33, -108], [33, 33, 108])),
    ("ref_sqr_c5g7_surface_tl.yaml", ([-1, -1, -1], [65, 65, 215])),
    ("PUa-1-0-IN.yaml", ([-6, -6, -6], [6, 6, 6])),
    ("Ua-1-1-CY.yaml", ([-10, -10, -10], [10, 10, 10])),
])
def test_find_cells_matches_oracle(ab, oracle_api, deck, box):
    orc = oracle_api.Oracle(deck_path(deck))
    gpu = ab.Backend(deck_path(deck), 0)
    rng = np.random.default_rng(3)
    n = 200000
    r = rng.uniform(box[0], box[1], size=(n, 3))
    # include points exactly on lattice planes / surfaces
    r[:2000] = np.round(r[:2000] / 1.26) * 1.26
    r[2000:3000, 0] = np.round(r[2000:3000, 0] / 21.42) * 21.42
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    oc, om = orc.find_cells(r, u)
    gc, gm = gpu.find_cells(r, u)
    assert np.array_equal(gc, oc) and np.array_equal(gm, om)
    assert (oc >= 0).sum() > n // 10


def test_source_sampling_bit_exact(ab, oracle_api, tmp_path):
    import torch
    # (the beam deck: mono-directional, cone from a box, cone about the pole, isotropic -- four sources picked by weight)
    # (the spectra deck: Maxwellian and Watt energies with the rejection loop of Source::generate_particle, a cone, a point)
    for deck in ("c5g7_delta_collision.yaml", "PUa-1-0-IN.yaml", "PUa-1-0-SL_subcritical_fs_beam.yaml",
                 "UD2O-2-1-SL_subcritical_fs_spectra.yaml"):
        orc, gpu = _pair(ab, oracle_api, tmp_path, deck, {"settings": {"nparticles": 20000}})
        ob = orc.sample_source(20000)
        db = gpu.new_device_bank(20000)
        gpu.sample_source_device(db, 20000, 0)
        torch.cuda.synchronize()
        for k in ("x", "y", "z", "ux", "uy", "uz", "E", "wgt"):
            assert np.array_equal(db[k].cpu().numpy(), ob[k]), k
        for k in ("id_a", "id_b", "id_c"):
            assert np.array_equal(db[k].cpu().numpy().view(np.uint64), ob[k]), k


# ---- Transporter::transport ----------------------------------------------------------------------------------------
CASES = [
    ("PUa-1-0-IN.yaml", {}, 20000),                       # S1: surface tracking, reflective cube, 1 group
    ("c5g7_delta_collision.yaml", {}, 20000),             # S2: delta tracking, 2 lattice levels, collision tally
    ("c5g7_delta_tracklength.yaml", {}, 8000),            # delta tracking + track-length mesh walk
    ("c5g7_surface_tracklength.yaml", {}, 8000),          # surface tracking through lattices + cylinders
    ("ref_sqr_c5g7_surface_tl.yaml", {}, 8000),           # S3: complement (RPN) cells, surface tracking, TLE
    ("c5g7_carter_cancel.yaml", {"second_generation": True}, 8000),  # S4: carter tracking, negative weights, splitting
    ("PUa-1-1-SL.yaml", {}, 20000),                       # P1 anisotropic scattering, vacuum slab
    ("PUa-1-2-SL.yaml", {}, 20000),                       # P2
    ("UD2O-2-1-SL.yaml", {}, 20000),                      # 2 groups, P1, tally
    ("Ua-1-1-CY.yaml", {}, 20000),                        # infinite cylinder (zcylinder distance)
    ("PUb-1-0-SL.yaml", {}, 20000),
    ("PUa-1-0-SL_implicit.yaml", {}, 20000),              # implicit-leakage delta tracking, vacuum slab: every flight splits
    ("UD2O-2-1-SL_implicit.yaml", {}, 20000),             # ... two groups, P1, tally
    ("c5g7_implicit_collision.yaml", {}, 20000),          # ... lattices, reflective + vacuum sides, collision tally
    ("c5g7_implicit_tracklength.yaml", {}, 8000),         # ... track-length tally of the leaking and the colliding share
    ("hex_delta_collision.yaml", {}, 20000),              # HexLattice (pointy top) under delta tracking, an empty position, outer universe
    ("hex_delta_flat_offset.yaml", {}, 20000),            # ... flat top, origin off zero (the reference's un-shifted tile check)
    ("c5g7_delta_branchless.yaml", {"second_generation": True}, 8000),            # branchless collisions on the material (weights carry m)
    ("PUa-1-0-SL_branchless_iso_split.yaml", {"second_generation": True}, 8000),  # ... on the isotope, with splitting (floor)
    ("UD2O-2-1-SL_branchless_split_comb.yaml", {"second_generation": True}, 8000),  # ... on the material, with splitting (ceil), 2 groups
]


@pytest.mark.parametrize("deck,overrides,n", CASES, ids=[c[0] for c in CASES])
def test_transport_bit_exact_against_oracle(ab, oracle_api, tmp_path, deck, overrides, n):
    ov = {"settings": {"nparticles": n}}
    ov.update({k: v for k, v in overrides.items() if k != "second_generation"})
    orc, gpu = _pair(ab, oracle_api, tmp_path, deck, ov)
    bank, o, g = _transport_both(orc, gpu, n, second_generation=overrides.get("second_generation", False))
    _assert_same_histories(o, g)
    if deck == "c5g7_carter_cancel.yaml":
        assert (bank["wgt"] > 0).all() and (g[0]["wgt"] < 0).any(), "carter tracking must bank negative sites here"
    for t in range(gpu.ntallies()):
        og, gg = orc.tally(t, "gen"), gpu.tally(t, "gen")
        assert og.shape == gg.shape
        assert np.count_nonzero(og) > 0 or orc.tally_shape(t) is None
        assert np.array_equal(og != 0, gg != 0)
        assert np.allclose(gg, og, rtol=1e-10, atol=1e-300)


def test_transport_not_converged_scores_no_mesh_tally(ab, oracle_api, tmp_path):
    orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_delta_collision.yaml", {"settings": {"nparticles": 3000}})
    bank = orc.sample_source(3000)
    gpu.transport(bank, converged=False)
    assert np.count_nonzero(gpu.tally(0, "gen")) == 0


def test_second_generation_streams_from_history_ids(ab, oracle_api, tmp_path):
    """Generation g+1: fresh history ids, streams = seed + advance(stride*id) evaluated on the device."""
    n = 10000
    orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_delta_collision.yaml", {"settings": {"nparticles": n}})
    bank = orc.sample_source(n)
    fis, _, _ = gpu.transport(bank)
    m = len(fis["x"])
    nxt = _next_bank(fis, n)
    orc.set_trace(True)
    orc.set_kcol(1.17)
    ofis, oscores, on = orc.transport({k: (v.copy() if v is not None else None) for k, v in nxt.items()})
    otr = orc.trace(m)
    gfis, gscores, gcn = gpu.transport(nxt, k_col=1.17, trace=True)
    gtr = gpu.trace(m)
    for k in ("flights", "real", "virtual", "fission", "hash", "rng_state"):
        assert np.array_equal(gtr[k], otr[k]), k
    for k in BANK_EXACT:
        assert np.array_equal(gfis[k], ofis[k]), k


def test_empty_and_tiny_banks(ab, oracle_api, tmp_path):
    orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_delta_collision.yaml", {"settings": {"nparticles": 100}})
    empty = ab.new_bank(0)
    fis, scores, cn = gpu.transport(empty)
    assert len(fis["x"]) == 0 and np.all(scores == 0) and cn["flights"] == 0
    for n in (1, 31, 33, 129):
        orc.set_history_counter(0)
        bank, o, g = _transport_both(orc, gpu, n)
        _assert_same_histories(o, g)


def test_particle_born_outside_geometry_is_killed(ab, oracle_api, tmp_path):
    """Lost at birth = warning + kill in the reference (delta_tracker.cpp:92-98)."""
    orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_delta_collision.yaml", {"settings": {"nparticles": 100}})
    bank = orc.sample_source(64)
    bank["x"][::2] = 500.0
    fis, scores, cn = gpu.transport(bank, trace=True)
    tr = gpu.trace(64)
    assert cn["lost_at_birth"] == 32
    assert np.all(tr["flights"][::2] == 0) and np.all(tr["flights"][1::2] > 0)
    orc.set_trace(True)
    orc.reset_counters()
    ofis, oscores, on = orc.transport({k: v.copy() for k, v in bank.items()})
    assert orc.counters()["lost_at_birth"] == 32
    assert np.array_equal(fis["x"], ofis["x"])


def test_fission_bank_overflow_is_reported(ab, oracle_api, tmp_path):
    from abeille_b200 import BackendError
    orc, gpu = _pair(ab, oracle_api, tmp_path, "PUa-1-0-IN.yaml", {"settings": {"nparticles": 5000}})
    bank = orc.sample_source(5000)
    with pytest.raises(BackendError) as e:
        gpu.transport(bank, capacity=100)
    assert e.value.code == -3 and "overflow" in str(e.value)


def test_majorant_violation_is_fatal(ab, oracle_api, tmp_path):
    """A total cross section above the majorant must abort the run with the offending history's id, as
    delta_tracker.cpp:174-180 does (`if (Et - Emaj > 1e-10) fatal_error(...)`): ABL_ERR_MAJORANT through every entry point."""
    from abeille_b200 import BackendError
    n = 3000
    path = write_deck(load_deck("c5g7_delta_collision.yaml"), tmp_path / "maj.yaml", {"settings": {"nparticles": n}})
    gpu = ab.Backend(path, 0)
    orc = oracle_api.Oracle(path)
    bank = orc.sample_source(n)
    gpu.transport(bank)  # the deck's own majorant: fine
    from abeille_b200 import backend
    smp = np.array(backend.dump_tables(path)["smp"])
    low = smp.copy()
    low[6] *= 0.5  # thermal group: water's total xs (2.65) is now above the "majorant" (1.33)
    gpu.set_sampling_xs(low)
    for trace in (False, True):
        with pytest.raises(BackendError) as e:
            gpu.transport(bank, trace=trace)
        assert e.value.code == -5 and "majorant" in str(e.value)
        hid = int(str(e.value).split("history ")[1].rstrip(")"))
        assert 0 <= hid < n  # one of this bank's histories (which one is first is a race, as under OpenMP)
    gpu.set_sampling_xs(smp)  # restored: the same handle runs again, bit-exact
    fis, _, _ = gpu.transport(bank)
    ofis, _, _ = orc.transport({k: v.copy() for k, v in bank.items()})
    for k in BANK_EXACT:
        assert np.array_equal(fis[k], ofis[k]), k


def _holed_deck(transport):
    """A fuel sphere inside a reflective box whose cells do not fill the box: the shell 2 < |r| < 3 belongs to no cell."""
    d = load_deck("PUa-1-0-IN.yaml")
    d["surfaces"] = d["surfaces"] + [{"type": "sphere", "x0": 0.0, "y0": 0.0, "z0": 0.0, "r": 2.0, "id": 7},
                                     {"type": "sphere", "x0": 0.0, "y0": 0.0, "z0": 0.0, "r": 3.0, "id": 8}]
    d["cells"] = [{"region": "-7", "material": 1, "name": "ball", "id": 1},
                  {"region": "+8 & +1 & -2 & +3 & -4 & +5 & -6", "material": 1, "name": "rest", "id": 2}]
    d["universes"] = [{"cells": [1, 2], "id": 1}]
    d["settings"]["transport"] = transport
    return d


def test_lost_particle_is_fatal(ab, oracle_api, tmp_path):
    """A particle that crosses a surface into a point no cell covers is lost: fatal in the reference
    (surface_tracker.cpp:126-133 "Particle became lost" after cross_surface) -> ABL_ERR_LOST with its history id."""
    from abeille_b200 import BackendError
    n = 2000
    path = write_deck(_holed_deck("surface-tracking"), tmp_path / "hole.yaml", {"settings": {"nparticles": n}})
    gpu = ab.Backend(path, 0)
    bank = oracle_api.Oracle(path).sample_source(n)
    for trace in (False, True):
        with pytest.raises(BackendError) as e:
            gpu.transport(bank, trace=trace)
        assert e.value.code == -4 and "lost" in str(e.value)
        assert 0 <= int(str(e.value).split("history ")[1].rstrip(")")) < n


def test_lost_at_birth_is_a_warning_not_an_error(ab, oracle_api, tmp_path):
    """Source particles born where no cell is are killed with a warning (delta_tracker.cpp:92-98) and counted."""
    n = 2000
    path = write_deck(_holed_deck("delta-tracking"), tmp_path / "hole.yaml", {"settings": {"nparticles": n}})
    orc, gpu = oracle_api.Oracle(path), ab.Backend(path, 0)
    bank = orc.sample_source(n)
    rng = np.random.default_rng(5)
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    bank["x"][:], bank["y"][:], bank["z"][:] = (2.5 * dirs).T  # inside the hole
    fis, scores, cn = gpu.transport(bank)
    assert cn["lost_at_birth"] == n and len(fis["x"]) == 0 and cn["flights"] == 0


def test_secondary_stack_overflow_is_reported(ab, oracle_api, tmp_path):
    """Carter tracking splits a particle of weight w into ceil(|w|) copies (particle.hpp:165-173); the device keeps at most
    ABL_SEC_CAP of them per history and reports an overflow instead of dropping copies."""
    from abeille_b200 import BackendError
    n = 500
    path = write_deck(load_deck("c5g7_carter_cancel.yaml"), tmp_path / "carter.yaml", {"settings": {"nparticles": n}})
    gpu = ab.Backend(path, 0)
    bank = oracle_api.Oracle(path).sample_source(n)
    bank["wgt"][:] = 40.0
    with pytest.raises(BackendError) as e:
        gpu.transport(bank)
    assert e.value.code == -3 and "secondary" in str(e.value)
    assert 0 <= int(str(e.value).split("history ")[1].rstrip(")")) < n


def test_cpp_adapter_matches_c_abi(ab, oracle_api, tmp_path):
    """GPUTransporter::transport(vector<Particle>&) (the reference-shaped C++ entry) == abl_transport."""
    n = 6000
    orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_delta_collision.yaml", {"settings": {"nparticles": n}})
    bank = orc.sample_source(n)
    fis_c, scores_c, _ = gpu.transport(bank, k_col=1.1)
    fis_v, scores_v = gpu.transport_vectors(bank, k_col=1.1)
    for k in BANK_EXACT:
        assert np.array_equal(fis_c[k], fis_v[k]), k
    assert np.allclose(scores_c, scores_v, rtol=1e-11)


# ---- tallies -----------------------------------------------------------------------------------------------------------
def test_tally_statistics_match_oracle(ab, oracle_api, tmp_path):
    n = 5000
    orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_delta_collision.yaml", {"settings": {"nparticles": n}})
    orc.set_converged(True)
    for gen in range(3):
        orc.set_history_counter(gen * n)
        bank = orc.sample_source(n)
        orc.transport({k: v.copy() for k, v in bank.items()})
        gpu.transport(bank, converged=True)
        orc.tallies_record(1.0)
        orc.tallies_clear()
        gpu.tallies_record(1.0)
        gpu.tallies_clear()
    for which in ("avg", "var", "std"):
        o, g = orc.tally(0, which), gpu.tally(0, which)
        assert np.allclose(g, o, rtol=1e-9, atol=1e-300), which
    assert np.count_nonzero(gpu.tally(0, "gen")) == 0


# ---- inter-generation pipeline -----------------------------------------------------------------------------------------
def test_cancellation_and_normalisation_match_oracle(ab, oracle_api, tmp_path):
    import torch
    n = 12000
    orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_carter_cancel.yaml",
                     {"settings": {"nparticles": n}, "cancelator": {"type": "approximate", "shape": [10, 10, 8],
                                                                    "low": [-32.13, -32.13, -107.11], "hi": [10.71, 10.71, 85.69]}})
    fis1, _, _ = gpu.transport(orc.sample_source(n))
    fis, _, _ = gpu.transport(_next_bank(fis1, n))  # second generation: fast neutrons meet the under-estimated majorant
    m = len(fis["x"])
    assert (fis["wgt"] < 0).any(), "carter tracking with an under-estimated majorant should bank negative sites"
    ob = {k: fis[k].copy() for k in fis}
    ob["wgt2"] = np.zeros(m)
    stats = orc.cancel_and_normalize(ob, True)
    db = gpu.new_device_bank(m)
    for k in ("x", "y", "z", "ux", "uy", "uz", "E", "wgt"):
        db[k].copy_(torch.from_numpy(fis[k]))
    gpu.cancel_device(db, m)
    ws = gpu.weight_stats_device(db, m)
    gpu.scale_weights_device(db, m, n / (ws[2] - ws[3]))
    torch.cuda.synchronize()
    assert np.allclose(db["wgt"].cpu().numpy(), ob["wgt"], rtol=1e-11, atol=1e-300)
    assert abs(db["wgt"].sum().item() - n) < 1e-6 * n


def test_power_iteration_resident_and_host_paths_match_oracle(ab, oracle_api, tmp_path):
    n, ngen, nign = 6000, 8, 3
    ov = {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}}
    for resident in (False, True):
        orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_delta_collision.yaml", ov, name=f"pi{int(resident)}.yaml")
        o = orc.run_power_iteration(ngen, nign)
        g = gpu.run_power_iteration(ngen, nign, resident=resident)
        assert np.array_equal(g["nbank"], o["nbank"]), (resident, g["nbank"], o["nbank"])
        assert np.allclose(g["kcol"], o["kcol"], rtol=1e-10)
        assert np.allclose(g["leak"], o["leak"], rtol=1e-10)
        assert np.allclose(g["entropy"], o["entropy"], rtol=1e-10)
        assert np.allclose(gpu.tally(0, "avg"), orc.tally(0, "avg"), rtol=1e-9, atol=1e-300)
        assert np.allclose(gpu.tally(0, "std"), orc.tally(0, "std"), rtol=1e-7, atol=1e-300)


def test_power_iteration_implicit_leakage_matches_oracle(ab, oracle_api, tmp_path):
    """Whole k-eigenvalue runs with transport: implicit-leakage-delta-tracking (per-lane kernel in mode 0), resident and
    host-buffer paths, against the oracle's driver: bank sizes exactly, k / leakage / entropy series and tallies to 1e-10."""
    n, ngen, nign = 6000, 8, 3
    ov = {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}}
    for resident in (False, True):
        orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_implicit_collision.yaml", ov, name=f"pil{int(resident)}.yaml")
        o = orc.run_power_iteration(ngen, nign)
        g = gpu.run_power_iteration(ngen, nign, resident=resident)
        assert np.array_equal(g["nbank"], o["nbank"]), (resident, g["nbank"], o["nbank"])
        assert np.allclose(g["kcol"], o["kcol"], rtol=1e-10)
        assert np.allclose(g["leak"], o["leak"], rtol=1e-10) and o["leak"].min() > 0.0
        assert np.allclose(g["entropy"], o["entropy"], rtol=1e-10)
        assert np.allclose(gpu.tally(0, "avg"), orc.tally(0, "avg"), rtol=1e-9, atol=1e-300)


def test_power_iteration_surface_tracking_sood(ab, oracle_api, tmp_path):
    """S1 on the device: k_inf of PUa-1-0-IN within 3 sigma of 2.612903 (and equal to the oracle's series)."""
    n, ngen, nign = 20000, 40, 10
    ov = {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}}
    orc, gpu = _pair(ab, oracle_api, tmp_path, "PUa-1-0-IN.yaml", ov)
    g = gpu.run_power_iteration(ngen, nign, resident=True)
    assert abs(g["kcol_avg"] - 2.612903) < 3 * g["kcol_err"] + 1e-9
    assert abs(g["ktrk_avg"] - 2.612903) < 3 * g["ktrk_err"] + 1e-9
    o = orc.run_power_iteration(ngen, nign)
    assert np.array_equal(g["nbank"], o["nbank"])
    assert np.allclose(g["kcol"], o["kcol"], rtol=1e-10) and np.allclose(g["ktrk"], o["ktrk"], rtol=1e-10)


def test_results_written_as_npy(ab, tmp_path):
    path = write_deck(load_deck("c5g7_delta_collision.yaml"), tmp_path / "w.yaml",
                      {"settings": {"nparticles": 2000, "ngenerations": 4, "nignored": 2}})
    gpu = ab.Backend(path, 0)
    gpu.run_power_iteration(4, 2, resident=True)
    out = tmp_path / "results"
    gpu.write_results(str(out))
    avg = np.load(out / "flux_pin_avg.npy")
    std = np.load(out / "flux_pin_std.npy")
    assert avg.shape == (7, 51, 51, 1) and std.shape == avg.shape
    assert np.array_equal(avg, gpu.tally(0, "avg")) and np.load(out / "kcol.npy").shape == (4,)
    # what MeshTally::write_tally stores beside the arrays (src/mesh_tally.cpp:166-190) and Simulation::write_source's
    # [N, 9] source array (src/simulation.cpp:137-175)
    xb = np.load(out / "flux_pin_x-bounds.npy")
    deck = load_deck("c5g7_delta_collision.yaml")
    t = [m for m in deck["tallies"] if m["name"] == "flux_pin"][0]
    assert xb.shape == (52,) and xb[0] == t["low"][0] and abs(xb[-1] - t["hi"][0]) < 1e-12
    assert np.load(out / "flux_pin_z-bounds.npy").shape == (2,)
    assert np.array_equal(np.load(out / "flux_pin_energy-bounds.npy"), np.array(t["energy-bounds"], dtype=float))
    attrs = dict(l.split(": ") for l in (out / "flux_pin_attributes.txt").read_text().splitlines())
    assert attrs == {"quantity": t["quantity"], "estimator": t.get("estimator", "collision")}
    src = np.load(out / "source.npy")
    assert src.ndim == 2 and src.shape[1] == 9 and src.shape[0] > 1000
    assert np.allclose(np.linalg.norm(src[:, 3:6], axis=1), 1.0, atol=1e-12) and np.all(src[:, 8] == 0.0)
    assert abs(src[:, 7].sum() - 2000.0) < 1e-6  # the bank leaves the last generation normalised to nparticles


def test_max_run_time_ends_the_run_after_a_whole_generation(ab, tmp_path):
    """settings: max-run-time (minutes; src/parser.cpp:703-711).  PowerIterator::check_time (src/power_iterator.cpp:715-749) ends
    the loop after a generation -- recorded, normalised, written -- once less than two average generations of the budget are
    left; with no budget at all that is the first generation, on the host-buffer path and on the device-resident one."""
    for resident in (True, False):
        path = write_deck(load_deck("c5g7_delta_collision.yaml"), tmp_path / f"t{int(resident)}.yaml",
                          {"settings": {"nparticles": 2000, "ngenerations": 6, "nignored": 2, "max-run-time": 1.0e-9}})
        gpu = ab.Backend(path, 0)
        r = gpu.run_power_iteration(6, 2, resident=resident)
        assert r["kcol"].shape == (1,) and r["nbank"][0] == 2000 and 0.5 < r["kcol"][0] < 2.0
        out = tmp_path / f"results{int(resident)}"
        gpu.write_results(str(out))
        assert np.load(out / "kcol.npy").shape == (1,)
        src = np.load(out / "source.npy")
        assert abs(src[:, 7].sum() - 2000.0) < 1e-6  # the bank of the generation that ran, normalised
        gpu.close()
    # a generous budget changes nothing
    path = write_deck(load_deck("c5g7_delta_collision.yaml"), tmp_path / "t2.yaml",
                      {"settings": {"nparticles": 2000, "ngenerations": 4, "nignored": 2, "max-run-time": 600.0}})
    gpu = ab.Backend(path, 0)
    assert gpu.run_power_iteration(4, 2, resident=True)["kcol"].shape == (4,)
    gpu.close()


def test_full_size_generation_properties(ab, tmp_path):
    """BASELINE size (10^7 histories, the deck's own 1224x1224x10x7 mesh): properties that do not need the oracle.
    Every flight ends in exactly one of a real collision, a virtual collision or a boundary event; the collision
    estimator scores once per real collision; k_col == k_abs (one nuclide per material); the fission bank comes out in
    the reference's order (parent bank index, then daughter number 0..n-1) and is bit-identical between two runs."""
    import torch
    n = 10_000_000
    path = write_deck(load_deck("c5g7_delta_collision_fullmesh.yaml"), tmp_path / "full.yaml", {"settings": {"nparticles": n}})
    gpu = ab.Backend(path, 0)
    cap = int(2.0 * n)
    src, out1, out2 = gpu.new_device_bank(n), gpu.new_device_bank(cap), gpu.new_device_bank(cap)
    gpu.sample_source_device(src, n, 0)
    m1, s1, c1 = gpu.transport_device(src, n, out1, k_col=1.0, converged=True, use_rng_state=True)
    tally_sum = float(torch.as_tensor(gpu.tally(0, "gen")).sum())
    gpu.tallies_clear()
    m2, s2, c2 = gpu.transport_device(src, n, out2, k_col=1.0, converged=True, use_rng_state=True)
    assert m1 == m2 == c1["fission_sites"] and c1 == c2
    assert c1["flights"] == c1["real_collisions"] + c1["virtual_collisions"] + c1["boundary_events"]
    assert c1["coll_scores"] == c1["real_collisions"]
    assert abs(s1[0] - s1[1]) <= 1e-12 * abs(s1[0])  # k_col == k_abs
    assert np.allclose(s1, s2, rtol=1e-11)
    for k in BANK_EXACT:
        assert torch.equal(out1[k][:m1], out2[k][:m1]), f"two runs differ in {k}"
    ida, idb = out1["id_a"][:m1], out1["id_b"][:m1]
    assert bool((ida[1:] >= ida[:-1]).all()), "fission bank not in bank order"
    same = ida[1:] == ida[:-1]
    assert bool((idb[1:][same] == idb[:-1][same] + 1).all()) and bool((idb[1:][~same] == 0).all()) and int(idb[0]) == 0
    assert tally_sum > 0.


def test_streamed_host_bank_matches_resident_bank(ab, oracle_api, tmp_path):
    """abl_transport streams banks of >= 2^18 particles to the device in row chunks while the history kernel runs (the
    kernel waits for a row before loading it).  Same bank through the device-resident entry point: identical fission
    bank, scores and counters; and a second call reuses the staging buffers."""
    import torch
    n = 400_000
    path = write_deck(load_deck("c5g7_delta_collision.yaml"), tmp_path / "streamed.yaml", {"settings": {"nparticles": n}})
    gpu = ab.Backend(path, 0)
    dsrc, dout = gpu.new_device_bank(n), gpu.new_device_bank(4 * n)
    gpu.sample_source_device(dsrc, n, 0)
    host = {k: dsrc[k].cpu().numpy() for k in ("x", "y", "z", "ux", "uy", "uz", "E", "wgt")}
    host.update({k: dsrc[k].cpu().numpy().view(np.uint64) for k in ("id_a", "id_b", "id_c")})
    host["wgt2"] = None
    m, s_dev, c_dev = gpu.transport_device(dsrc, n, dout, k_col=1.0, converged=True, use_rng_state=True)
    for rep in range(2):
        gpu.tallies_clear()
        fis, s_host, c_host = gpu.transport(host, k_col=1.0, converged=True, capacity=4 * n)
        assert len(fis["x"]) == m and c_host == c_dev
        assert np.allclose(s_host, s_dev, rtol=1e-11)
        for k in BANK_EXACT:
            ref = dout[k][:m].cpu().numpy()
            got = fis[k] if fis[k].dtype == ref.dtype else fis[k].view(ref.dtype)
            assert np.array_equal(got, ref), f"streamed call {rep}: {k} differs"


def test_distributed_iterator_with_cancellation_matches_oracle(ab, oracle_api, tmp_path):
    """Config 4 (carter tracking, under-estimated majorant, approximate regional cancellation) through the multi-GPU
    driver on one rank: k and bank sizes of every generation against the oracle's PowerIterator."""
    from abeille_b200.distributed import DistributedPowerIterator
    n, ngen = 20000, 4
    path = write_deck(load_deck("c5g7_carter_cancel.yaml"), tmp_path / "carter.yaml",
                      {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": 1}})
    ref = oracle_api.Oracle(path).run_power_iteration(ngen, 1)
    sim = DistributedPowerIterator(path, 0, n)
    assert sim.cancellation
    sim.initialize()
    for g in range(ngen):
        sim.generation(converged=g >= 1)
    assert [int(v) for v in sim.nbank_series] == [int(v) for v in ref["nbank"]]
    assert np.allclose(sim.kcol_series, ref["kcol"], rtol=1e-10)


def test_sood_k_inf_deck_at_large_n_does_not_overflow(ab, oracle_api, tmp_path):
    """PUa-1-0-IN (k_inf = 2.61) banks 2.61 sites per particle in the first generation (k_col starts at 1,
    tallies.cpp:48): the output banks are sized from the problem (abl_fission_capacity_hint), not as a fixed multiple
    of the bank -- with N = 2*10^5 a 2.5 N bank would overflow (522 000 sites vs 504 096)."""
    n, ngen, nign = 200000, 3, 1
    ov = {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}}
    for resident in (True, False):
        path = write_deck(load_deck("PUa-1-0-IN.yaml"), tmp_path / f"big{int(resident)}.yaml", ov)
        gpu = ab.Backend(path, 0)
        g = gpu.run_power_iteration(ngen, nign, resident=resident)
        assert g["nbank"][1] > 2.55 * n and abs(g["kcol"][0] - 2.612903) < 0.02
    assert gpu.fission_capacity(n) > 2.7 * n
    gpu1 = ab.Backend(write_deck(load_deck("c5g7_delta_collision.yaml"), tmp_path / "c.yaml"), 0)
    assert 1.3 * n < gpu1.fission_capacity(n) < 3.5 * n


def test_distributed_iterator_regrows_its_banks(ab, oracle_api, tmp_path):
    """An output bank that turns out too small is grown and the generation repeated (scores come back per call, tally_gen
    is cleared before the retry): same bank sizes, k and tallies as a run whose banks were large enough from the start,
    entropy and source-estimator tallies included."""
    from abeille_b200.distributed import DistributedPowerIterator
    n, ngen = 8000, 4
    deck = load_deck("c5g7_delta_collision.yaml")
    deck["tallies"] = deck["tallies"] + [{"name": "src", "low": [-32.13, -32.13, -107.1], "hi": [32.13, 32.13, 107.1], "shape": [17, 17, 1],
                                          "energy-bounds": [0, 7], "quantity": "source", "estimator": "source"}]
    path = write_deck(deck, tmp_path / "d.yaml", {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": 1}})
    ref = oracle_api.Oracle(path).run_power_iteration(ngen, 1)
    runs = []
    for small in (False, True):
        sim = DistributedPowerIterator(path, 0, n)
        if small:  # banks that hold the source but not the fission sites of the first generation
            sim.cap = n + 64
            sim.cur, sim.nxt = sim.gpu.new_device_bank(sim.cap), sim.gpu.new_device_bank(sim.cap)
        sim.initialize()
        for g in range(ngen):
            sim.generation(converged=g >= 1)
        runs.append(sim)
        assert [int(v) for v in sim.nbank_series] == [int(v) for v in ref["nbank"]]
        assert np.allclose(sim.kcol_series, ref["kcol"], rtol=1e-10)
        assert np.allclose(sim.entropy_series, ref["entropy"], rtol=1e-10)
    assert runs[1].cap > n + 64
    for t in range(runs[0].gpu.ntallies()):
        a, b = runs[0].gpu.tally(t, "avg"), runs[1].gpu.tally(t, "avg")
        assert np.allclose(a, b, rtol=1e-12, atol=0) and a.sum() > 0


# ---- the sharded power iteration: two ranks against one (the multi-GPU path where a one-GPU test run can see it) ---------------------
# Two processes, one rank each, both on GPU 0 with the gloo backend (NCCL refuses two ranks on one device): slices, global
# history ids, the rebalancing all_to_all, the tally / cancellation all_reduce are those of a two-GPU run.
def _sharded_pi_worker(rank, world, port, path, n_local, gens, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from abeille_b200.distributed import DistributedPowerIterator
        sim = DistributedPowerIterator(path, 0, n_local)
        sim.initialize()
        for g in range(gens):
            sim.generation(converged=g >= 1)
        q.put((rank, {"nbank": [int(v) for v in sim.nbank_series], "k_col": [float(v) for v in sim.kcol_series],
                      "collisions": float(sim.counters[1]),
                      "tallies": [sim.gpu.tally(t, "avg") for t in range(sim.gpu.ntallies())]}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("deck", ["c5g7_delta_collision.yaml", "c5g7_carter_cancel.yaml", "c5g7_delta_branchless.yaml"])
def test_power_iteration_sharded_over_two_ranks_matches_one_rank(ab, tmp_path, deck):
    import socket
    import torch.multiprocessing as mp
    from abeille_b200.distributed import DistributedPowerIterator
    n, gens, world = 40000, 4, 2
    path = write_deck(load_deck(deck), tmp_path / deck, {"settings": {"nparticles": n}})
    one = DistributedPowerIterator(path, 0, n)
    one.initialize()
    for g in range(gens):
        one.generation(converged=g >= 1)
    ref = {"nbank": [int(v) for v in one.nbank_series], "k_col": [float(v) for v in one.kcol_series],
           "collisions": float(one.counters[1]), "tallies": [one.gpu.tally(t, "avg") for t in range(one.gpu.ntallies())]}
    del one
    if "branchless" in deck:
        # the comb in the sharded driver (slices to rank 0 in bank order, normalised and combed there from the weights, an even
        # split back) against the one-GPU C++ loop, which is pinned on the oracle and the reference
        host = ab.Backend(path, 0).run_power_iteration(gens, 1, resident=True)
        assert ref["nbank"] == [int(v) for v in host["nbank"]] and np.allclose(ref["k_col"], host["kcol"], rtol=1e-10)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_pi_worker, args=(r, world, port, path, n // world, gens, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        got = res[r]
        if "carter" in deck:
            # regional cancellation sums the bins in another order on two ranks: weights move in the last bit, which can flip
            # a roulette or a split a generation later -- sizes and k agree statistically, not bit for bit
            assert np.allclose(got["nbank"], ref["nbank"], rtol=5e-3) and np.allclose(got["k_col"], ref["k_col"], rtol=5e-3)
            continue
        # histories shard by global id and RNG streams are a function of the id: integer outcomes are identical
        assert got["nbank"] == ref["nbank"] and got["collisions"] == ref["collisions"]
        assert np.allclose(got["k_col"], ref["k_col"], rtol=1e-12)
        for t, (a, b) in enumerate(zip(got["tallies"], ref["tallies"])):
            scale = np.abs(b).max()
            assert np.allclose(a, b, rtol=1e-7, atol=1e-9 * scale), f"tally {t}"


def test_million_histories_full_mesh_generation_against_oracle(ab, oracle_api, tmp_path):
    """The bench workload at 1e6 histories against the oracle: the deck's 1224 x 1224 x 10 x 7 mesh (105 M bins), a fission
    source, the host-buffer entry point -- so the streamed input bank, the ticket refill over many waves per thread and the
    32-bit site offsets run at scale -- with every history's integer outcome and the whole fission bank compared bit for bit
    and the mesh tally bin by bin."""
    n = 1_000_000
    orc, gpu = _pair(ab, oracle_api, tmp_path, "c5g7_delta_collision_fullmesh.yaml", {"settings": {"nparticles": n}})
    oracle_api.set_threads(os.cpu_count() or 1)
    bank, o, g = _transport_both(orc, gpu, n, converged=True, k_col=1.17, second_generation=True)
    assert len(bank["x"]) > n  # the fission source of the first generation (k_col = 1 there banks ~1.6 sites per history)
    _assert_same_histories(o, g)
    og, gg = orc.tally(0, "gen"), gpu.tally(0, "gen")
    assert og.shape == gg.shape == (7, 1224, 1224, 10)
    assert np.array_equal(og != 0, gg != 0)
    assert np.allclose(gg, og, rtol=1e-10, atol=1e-300)
    assert abs(gg.sum() - og.sum()) <= 1e-9 * abs(og.sum())


def test_modified_fixed_source_driver_matches_oracle_and_reference(ab, oracle_api, tmp_path):
    """abeille_b200.fixed_source.ModifiedFixedSource (the reference's ModifiedFixedSource::run over the device entry points)
    on a subcritical slab: every batch's k_col, leakage and migration area and the mesh tally statistics against the oracle's
    driver (det math: 1e-10) and against the reference's own run (tests/golden/ref_pins_mfs.npz: 1e-9)."""
    from abeille_b200.fixed_source import ModifiedFixedSource
    from oracle import ref_pins
    fname, n, nb = ref_pins.MFS_CASES[0]
    path = write_deck(load_deck(fname), tmp_path / fname, {"settings": {"nparticles": n, "ngenerations": nb}})
    orc = oracle_api.Oracle(path)
    ref = orc.run_modified_fixed_source(nb)
    sim = ModifiedFixedSource(path, 0)
    got = sim.run()
    assert got["transported"] == ref["transported"] and min(got["chain_generations"]) > 3
    for k in ("kcol", "leak", "mig"):
        assert np.allclose(got[k], ref[k], rtol=1e-10), (k, got[k], ref[k])
    for t in range(orc.ntallies()):
        for which in ("avg", "std"):
            a, b = sim.tally(t, which), orc.tally(t, which)
            assert np.allclose(a, b, rtol=1e-8, atol=1e-12 * np.abs(b).max()), (t, which)
    gold = dict(np.load(os.path.join(GOLDEN, "ref_pins_mfs.npz")))
    name = fname.split(".")[0]
    for k in ("kcol", "leak", "mig"):
        assert np.allclose(got[k], gold[f"mfs_{name}_{k}"], rtol=1e-9), k
    assert np.allclose(np.ravel(sim.tally(0, "avg")), gold[f"mfs_{name}_tally0_avg"], rtol=1e-7, atol=1e-12)
    sim.close()


def test_two_phase_host_transport_matches_the_resident_loop(ab, tmp_path):
    """abl_transport_begin / abl_transport_finish (the bank on the host before and after, normalize_weights and the fresh
    history ids applied on the device in between) through HostBufferLoop against the device-resident loop: the same k_col and
    bank size every generation, the same mesh tally."""
    from abeille_b200.distributed import DistributedPowerIterator, HostBufferLoop
    n, gens = 300_000, 4  # (>= 2^18 histories: the input bank is streamed behind the kernel)
    path = write_deck(load_deck("c5g7_delta_collision.yaml"), tmp_path / "two_phase.yaml", {"settings": {"nparticles": n}})
    res = DistributedPowerIterator(path, 0, n)
    res.initialize()
    for g in range(gens):
        res.generation(converged=g >= 1)
    ks, ns = list(res.kcol_series), [int(v) for v in res.nbank_series]
    tally = res.gpu.tally(0, "avg")
    del res
    loop = HostBufferLoop(path, 0, n)
    loop.initialize()
    got_k, got_n = [], []
    for g in range(gens):
        r = loop.generation(converged=g >= 1)
        got_k.append(r["k_col"])
        got_n.append(r["n_in"])
    assert got_n == ns
    assert np.allclose(got_k, ks, rtol=1e-12)
    assert np.allclose(loop.gpu.tally(0, "avg"), tally, rtol=1e-9, atol=1e-12 * np.abs(tally).max())
    assert loop.h2d_bytes > 0 and loop.d2h_bytes > 0


@pytest.mark.parametrize("cases,ci,golden_file", [("FS_CASES", 0, "ref_pins_mfs.npz"), ("BEAM_CASES", 0, "ref_pins_beam.npz"),
                                                  ("BEAM_CASES", 1, "ref_pins_beam.npz")])
def test_fixed_source_driver_matches_oracle_and_reference(ab, oracle_api, tmp_path, cases, ci, golden_file):
    """abeille_b200.fixed_source.FixedSource (the reference's FixedSource::run: fission neutrons continue their history as
    secondaries, transport returns an empty bank) on a subcritical slab against the oracle's driver and the reference's own
    run (tests/golden/ref_pins_mfs.npz); and the same slab lit by mono-directional and cone sources (src/mono_directional.cpp,
    src/cone.cpp; tests/golden/ref_pins_beam.npz), and a two-group slab lit by Maxwellian and Watt spectra (src/maxwellian.cpp,
    src/watt.cpp)."""
    from abeille_b200.fixed_source import FixedSource
    from oracle import ref_pins
    fname, n, nb = getattr(ref_pins, cases)[ci]
    path = write_deck(load_deck(fname), tmp_path / fname, {"settings": {"nparticles": n, "ngenerations": nb}})
    orc = oracle_api.Oracle(path)
    ref = orc.run_fixed_source(nb)
    sim = FixedSource(path, 0)
    got = sim.run()
    for k in ("kcol", "leak", "mig"):
        assert np.allclose(got[k], ref[k], rtol=1e-10), (k, got[k], ref[k])
    for t in range(orc.ntallies()):
        for which in ("avg", "std"):
            a, b = sim.tally(t, which), orc.tally(t, which)
            assert np.allclose(a, b, rtol=1e-8, atol=1e-12 * np.abs(b).max()), (t, which)
    gold = dict(np.load(os.path.join(GOLDEN, golden_file)))
    name = fname.split(".")[0]
    for k in ("kcol", "leak", "mig"):
        assert np.allclose(got[k], gold[f"fs_{name}_{k}"], rtol=1e-9), k
    assert got["leak"].min() > 1.0  # every source neutron's chain leaks more than one neutron's weight: the slab multiplies
    sim.close()


def test_power_iterator_diagnostics_match_the_references_definitions(ab, tmp_path):
    """settings: pair-distance-sqrd, families, empty-entropy-bins (src/parser.cpp:833-858).  The pair distance is the reference's
    double sum over all pairs of the normalised fission bank (PowerIterator::compute_pair_dist_sqrd, src/power_iterator.cpp:637-663),
    evaluated here term by term in numpy; the device gets it from two passes of abl_bank_moments_device.  Families: the distinct
    family ids entering a generation (src/power_iterator.cpp:326-331).  Empty entropy bins: Entropy::calculate_empty_fraction
    (src/entropy.cpp:95-105) against a histogram of the bank's positions."""
    from abeille_b200.distributed import DistributedPowerIterator
    n = 2500
    deck = load_deck("c5g7_delta_collision.yaml")
    path = write_deck(deck, tmp_path / "d.yaml", {"settings": {"nparticles": n, "ngenerations": 5, "nignored": 2, "pair-distance-sqrd": True,
                                                               "families": True, "empty-entropy-bins": True}})
    sim = DistributedPowerIterator(path, 0, n)
    sim.initialize()
    ent = deck["entropy"]
    edges = [np.linspace(ent["low"][k], ent["hi"][k], int(ent["shape"][k]) + 1) for k in range(3)]
    for g in range(4):
        fam_in = len(np.unique(sim.cur["id_b"][:sim.n_cur].cpu().numpy()))
        sim.generation(converged=g >= 2)
        assert sim.families_series[-1] == fam_in
        m = sim.n_cur
        r = np.stack([sim.cur[k][:m].cpu().numpy() for k in ("x", "y", "z")], 1)
        w = sim.cur["wgt"][:m].cpu().numpy()
        d2 = ((r[:, None, :] - r[None, :, :]) ** 2).sum(axis=2)               # r.dot(r) of every pair
        ref = float((d2 * w[:, None] * w[None, :]).sum() / (2.0 * w.sum() ** 2))
        assert abs(sim.r_sqrd_series[-1] - ref) < 1e-10 * ref, (g, sim.r_sqrd_series[-1], ref)
        hist, _ = np.histogramdd(r, bins=edges)
        assert abs(sim.empty_entropy_frac_series[-1] - float((hist == 0).sum()) / hist.size) < 2.0 / hist.size
    assert sim.families_series[0] == n and sim.families_series[-1] < sim.families_series[1] <= n   # families die out
    assert 0.0 < sim.empty_entropy_frac_series[-1] < 1.0 and sim.r_sqrd_series[-1] > 100.0
    # moments about a far origin and about the centroid give the same second central moment
    mo = sim.gpu.moments_device(sim.cur, sim.n_cur, (1.0e3, -2.0e3, 5.0e2))
    c = np.array([1.0e3, -2.0e3, 5.0e2]) + mo[1:4] / mo[0]
    mc = sim.gpu.moments_device(sim.cur, sim.n_cur, c)
    assert abs(mc[4] / mc[0] - sim.r_sqrd_series[-1]) < 1e-9 * sim.r_sqrd_series[-1] and np.abs(mc[1:4]).max() < 1e-6 * mo[0]
    empty = sim.gpu.moments_device(sim.cur, 0)
    assert not empty.any()
    # the C++ host's device-resident PowerIterator keeps the same three series and writes them under the reference's dataset
    # names (results/families, results/pair-dist-sqrd, results/empty-entropy-frac: src/power_iterator.cpp:475-503)
    gpu = ab.Backend(path, 0)
    gpu.run_power_iteration(4, 2, resident=True)
    out = tmp_path / "results"
    gpu.write_results(str(out))
    assert np.array_equal(np.load(out / "families.npy"), np.array(sim.families_series, dtype=float))
    assert np.allclose(np.load(out / "pair-dist-sqrd.npy"), sim.r_sqrd_series, rtol=1e-10, atol=0.0)
    assert np.allclose(np.load(out / "empty-entropy-frac.npy"), sim.empty_entropy_frac_series, rtol=0.0, atol=1e-15)


def test_restart_from_a_saved_source_matches_oracle(ab, oracle_api, tmp_path):
    """settings: insource (PowerIterator::load_source_from_file, power_iterator.cpp:60-133): the saved [N, 9] source
    becomes the bank, history ids are the row numbers, the streams are seeded from them and nparticles is the rounded
    total weight.  The first generation after the restart against the oracle transporting the same rows."""
    from abeille_b200.distributed import DistributedPowerIterator
    n = 6000
    path = write_deck(load_deck("c5g7_delta_collision.yaml"), tmp_path / "d.yaml",
                      {"settings": {"nparticles": n, "ngenerations": 4, "nignored": 1}})
    a = DistributedPowerIterator(path, 0, n)
    a.initialize()
    for g in range(2):
        a.generation(converged=False)
    src = a.source_array()
    assert src.shape == (a.n_cur, 9) and src.shape[0] != n and int(round(src[:, 7].sum())) == n
    np.save(tmp_path / "source.npy", src)

    b = DistributedPowerIterator.from_source(path, str(tmp_path / "source.npy"), 0)
    assert b.n_total == n and b.n_cur == len(src) and b.global_counter == len(src)
    r = b.generation(converged=False)

    orc = oracle_api.Oracle(path)
    bank = {k: np.ascontiguousarray(src[:, i]) for i, k in enumerate(("x", "y", "z", "ux", "uy", "uz", "E", "wgt", "wgt2"))}
    bank["id_a"] = np.arange(len(src), dtype=np.uint64)
    bank["id_b"] = bank["id_a"].copy()
    bank["id_c"] = None
    orc.set_converged(False)
    orc.set_kcol(1.0)
    ofis, oscores, om = orc.transport(bank)
    assert r["m_total"] == om
    assert np.isclose(b.k_col, oscores[0] / n, rtol=1e-11)
    got = b.source_array()
    for i, k in enumerate(("x", "y", "z", "ux", "uy", "uz", "E")):
        assert np.array_equal(got[:, i], ofis[k]), k
    assert np.allclose(got[:, 7], ofis["wgt"] * n / ofis["wgt"].sum(), rtol=1e-12)
    with pytest.raises(ValueError):
        DistributedPowerIterator.from_source(path, src[:, :8], 0)


@pytest.mark.parametrize("deck", ["c5g7_delta_branchless.yaml", "PUa-1-0-SL_branchless_iso_split.yaml", "UD2O-2-1-SL_branchless_split_comb.yaml"])
def test_branchless_power_iteration_matches_oracle(ab, oracle_api, tmp_path, deck):
    """simulation: branchless-k-eigenvalue (src/branchless_power_iterator.cpp) -- branchless collisions on the device, the comb as
    the serial host step it is in the reference -- with the bank resident in HBM and through host buffers, against the oracle's
    driver (itself bit-for-bit the reference's, tests/test_reference_pins.py): bank sizes exactly, the series to 1e-10."""
    n, ngen, nign = 5000, 7, 2
    ov = {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}}
    for resident in (False, True):
        orc, gpu = _pair(ab, oracle_api, tmp_path, deck, ov, name=f"bl{int(resident)}.yaml")
        o = orc.run_power_iteration(ngen, nign)
        g = gpu.run_power_iteration(ngen, nign, resident=resident)
        assert np.array_equal(g["nbank"], o["nbank"]), (resident, g["nbank"], o["nbank"])
        assert np.allclose(g["kcol"], o["kcol"], rtol=1e-10)
        assert np.allclose(g["leak"], o["leak"], rtol=1e-10, atol=1e-300)
        assert np.allclose(g["mig"], o["mig"], rtol=1e-10)
        assert np.allclose(g["entropy"], o["entropy"], rtol=1e-10)
        for t in range(gpu.ntallies()):
            assert np.allclose(gpu.tally(t, "avg"), orc.tally(t, "avg"), rtol=1e-9, atol=1e-300)
        if "comb" in deck or deck.startswith("c5g7"):  # a combed bank holds nparticles (or one more) particles of equal weight
            assert all(abs(int(v) - n) <= 1 for v in g["nbank"][1:])


@pytest.mark.parametrize("deck", ["PUa-cube_carter_exact_min.yaml", "PUa-cube_carter_exact_avgf.yaml", "PUa-cube_carter_exact_avgg.yaml",
                                  "PUa-cube_carter_exact_full.yaml"])
def test_exact_cancellation_matches_oracle(ab, oracle_api, tmp_path, deck):
    """cancelator: {type: basic-exact} (src/basic_exact_mg_cancelator.cpp; beta minimum, average-f with points drawn from the global
    engine, average-g with Sobol points) and {type: exact} (src/exact_mg_cancelator.cpp) under carter tracking with negative weights: the per-site parent data the kernels keep for
    it (BankedParticle::parents_previous_position through reflections, Esmp_parent), the cancelled weights, the appended uniform
    particles and the engine state afterwards, all bit for bit against the oracle (whose driver is pinned on the reference's,
    tests/test_reference_pins.py); then whole simulations on both host paths."""
    import torch
    n = 6000
    orc, gpu = _pair(ab, oracle_api, tmp_path, deck, {"settings": {"nparticles": n}})
    bank = orc.sample_source(n)
    fis, _, _ = orc.transport({k: v.copy() for k, v in bank.items()})
    nxt = _next_bank(fis, n)                     # a second generation: both signs, reflections behind the parents
    m_in = len(nxt["x"])
    ofis, _, om = orc.transport({k: (v.copy() if v is not None else None) for k, v in nxt.items()})
    opar = orc.last_parent_info(om)
    assert (ofis["wgt"] < 0).any() and (ofis["wgt"] > 0).any()
    dev_in, dev_out = gpu.new_device_bank(m_in), gpu.new_device_bank(3 * om + 4096)
    for k, v in nxt.items():
        if v is not None:
            dev_in[k].copy_(torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else v))
    gm, _, _ = gpu.transport_device(dev_in, m_in, dev_out)
    assert gm == om
    gpar = gpu.parent_info(gm)
    assert np.array_equal(gpar, opar)
    opst = orc.last_parent_state(om)             # previous direction, energy before the last scatter, energy, was_virtual
    assert np.array_equal(gpu.parent_state(gm), opst)
    assert 0.02 < opst[:, 5].mean() < 0.98 and (opst[:, 0] != 1.0).any()
    if "min" in deck:
        assert np.abs(opar[:, :3]).max() > 5.0   # mirror images behind the reflective faces of the +-5 cm cube
    state = ab.global_rng_state()
    ocb, ostate = orc.cancel_exact({k: v.copy() for k, v in ofis.items()}, opar, state, orc.last_parent_state(om))
    gn, gstate = gpu.cancel_exact_device(dev_out, gm, state)
    assert gn == len(ocb["x"]) and gn > om and gstate == ostate
    for k in ("x", "y", "z", "ux", "uy", "uz", "E", "wgt"):
        assert np.array_equal(dev_out[k][:gn].cpu().numpy(), ocb[k]), k
    assert np.abs(ocb["wgt"][:om]).sum() < np.abs(ofis["wgt"]).sum()   # weight was cancelled
    n, ngen, nign = 4000, 7, 2
    ov = {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}}
    for resident in (False, True):
        orc, gpu = _pair(ab, oracle_api, tmp_path, deck, ov, name=f"ex{int(resident)}.yaml")
        o = orc.run_power_iteration(ngen, nign)
        g = gpu.run_power_iteration(ngen, nign, resident=resident)
        assert np.array_equal(g["nbank"], o["nbank"]), (resident, g["nbank"], o["nbank"])
        assert np.allclose(g["kcol"], o["kcol"], rtol=1e-10)
        assert np.allclose(g["mig"], o["mig"], rtol=1e-10)
        assert np.allclose(g["entropy"], o["entropy"], rtol=1e-10)


def test_exact_cancellation_in_bins_that_hold_several_materials(ab, oracle_api, tmp_path):
    """The shipped c5g7.yaml's cancelator kind (basic-exact, average-g, Sobol points) on the carter-tracking c5g7 deck, on a mesh
    whose bins span fuel, cladding-free moderator and guide tubes: the rejection sampling of bin points on the bin's material, bins
    keyed by (mesh cell, material), whole simulations against the oracle (bank sizes -- uniform particles included -- exactly)."""
    n, ngen, nign = 8000, 5, 2
    ov = {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}}
    for deck in ("c5g7_carter_exact_avgg.yaml", "c5g7_carter_exact_full.yaml"):   # ... and the same mesh under `type: exact`
        for resident in (True, False):
            orc, gpu = _pair(ab, oracle_api, tmp_path, deck, ov, name=f"c5e{int(resident)}.yaml")
            o = orc.run_power_iteration(ngen, nign)
            g = gpu.run_power_iteration(ngen, nign, resident=resident)
            assert np.array_equal(g["nbank"], o["nbank"]), (deck, resident, g["nbank"], o["nbank"])
            assert np.allclose(g["kcol"], o["kcol"], rtol=1e-10)
            assert np.allclose(g["entropy"], o["entropy"], rtol=1e-10)
