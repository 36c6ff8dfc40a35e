"""GPU: the CUDA path against the REFERENCE'S OWN outputs (tests/golden/ref_pins.npz, made on the CPU box by
scripts/make_ref_pins.py from oracle/_ref -- the reference's SurfaceTracker / DeltaTracker / CarterTracker::transport compiled
from its sources), through the C ABI (abl_transport).  No oracle in between: the same seeded banks (oracle/ref_pins.py)
go to the kernels, and the fission bank they return is compared with the one the reference returned.

Integer results -- number of sites, parent history id, daughter id, family id, order -- and the values no libm call
enters (energy = group mid-point, weight) must be identical.  Positions and directions carry glibc's log / sin / cos on the
reference side and the shared fdlibm sequence on the device, which differ by an ulp on a fraction of
the arguments: they are compared to 1e-9 (absolute, cm / unit vector), generation values to 1e-9 relative, mesh-tally
bins to 1e-9 of the largest bin."""
import os

import numpy as np
import pytest

from conftest import load_deck, write_deck
from oracle import ref_pins

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_pins.npz")


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="module")
def ab(native_libs):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return native_libs


@pytest.mark.parametrize("ci", range(len(ref_pins.TRANSPORT_CASES)), ids=[c[0].split(".")[0] for c in ref_pins.TRANSPORT_CASES])
def test_kernels_reproduce_the_references_transport(ab, golden, tmp_path, ci):
    fname, n, k_col = ref_pins.TRANSPORT_CASES[ci]
    _transport_case(ab, golden, tmp_path, fname, n, k_col, 500 + ci)


GOLDEN_IMPLICIT = os.path.join(os.path.dirname(__file__), "golden", "ref_pins_implicit.npz")


@pytest.mark.parametrize("ci", range(len(ref_pins.IMPLICIT_CASES)), ids=[c[0].split(".")[0] for c in ref_pins.IMPLICIT_CASES])
def test_kernels_reproduce_the_references_implicit_leakage_transport(ab, tmp_path, ci):
    """ImplicitLeakageDeltaTracker::transport of the reference (scripts/make_ref_pins_implicit.py) against the per-lane
    kernel: here the reference's weights carry glibc's exp, the device's the shared fdlibm sequence, so the implicit
    leakage / migration scores and the tallies agree to 1e-9 and the integer outcomes exactly."""
    fname, n, k_col = ref_pins.IMPLICIT_CASES[ci]
    _transport_case(ab, dict(np.load(GOLDEN_IMPLICIT)), tmp_path, fname, n, k_col, 700 + ci)


GOLDEN_HEX = os.path.join(os.path.dirname(__file__), "golden", "ref_pins_hex.npz")


@pytest.mark.parametrize("ci", range(len(ref_pins.HEX_CASES)), ids=[c[0].split(".")[0] for c in ref_pins.HEX_CASES])
def test_kernels_reproduce_the_references_hex_lattice_transport(ab, tmp_path, ci):
    """DeltaTracker::transport of the reference over its HexLattice (scripts/make_ref_pins_hex.py) against the kernels."""
    fname, n, k_col = ref_pins.HEX_CASES[ci]
    # (these histories are long -- a small leaky core -- so most of them meet one of the arguments where glibc's log / sin / cos
    # and the shared fdlibm sequence differ in the last bit: positions agree to 1e-9, fewer of them to the bit)
    _transport_case(ab, dict(np.load(GOLDEN_HEX)), tmp_path, fname, n, k_col, 1300 + ci, min_identical=0.05)


def _transport_case(ab, golden, tmp_path, fname, n, k_col, seed, min_identical=0.5):
    name = fname.split(".")[0]
    deck = load_deck(fname)
    path = write_deck(deck, tmp_path / fname, {"settings": {"nparticles": n}})
    deck["settings"]["nparticles"] = n
    r, u, E, w, hid = ref_pins.transport_bank(deck, n, seed, "carter" in fname)
    bank = {k: np.ascontiguousarray(v) for k, v in zip(("x", "y", "z"), r.T)}
    bank.update({k: np.ascontiguousarray(v) for k, v in zip(("ux", "uy", "uz"), u.T)})
    bank.update(E=np.ascontiguousarray(E), wgt=np.ascontiguousarray(w), wgt2=np.zeros(n), id_a=hid, id_b=hid.copy(), id_c=None)
    gpu = ab.Backend(path, 0)
    gpu.tallies_clear()
    fis, scores, _ = gpu.transport(bank, k_col=k_col, converged=True, capacity=16 * n)

    ref_sites, ref_ids, ref_k = golden[f"transport_{name}_sites"], golden[f"transport_{name}_ids"], golden[f"transport_{name}_k"]
    assert len(fis["x"]) == len(ref_sites), "number of fission sites differs from the reference"
    got_ids = np.stack([fis["id_a"], fis["id_b"], fis["id_c"]], 1)
    assert np.array_equal(got_ids, ref_ids), "parent history id / daughter id / family id or their order differ"
    assert np.array_equal(fis["E"], ref_sites[:, 6]), "site energies differ"
    assert np.array_equal(fis["wgt"], ref_sites[:, 7]), "site weights differ"
    got_ru = np.stack([fis[k] for k in ("x", "y", "z", "ux", "uy", "uz")], 1)
    assert np.abs(got_ru - ref_sites[:, :6]).max() < 1e-9
    # identical to the last bit wherever no libm difference entered the history: the large majority
    assert (got_ru == ref_sites[:, :6]).all(1).mean() > min_identical
    assert np.allclose(scores / float(n), ref_k, rtol=1e-9, atol=1e-300), (scores / float(n), ref_k)
    for t in range(gpu.ntallies()):
        key = f"transport_{name}_tally{t}"
        if key in golden and golden[key].size:
            got = np.ravel(gpu.tally(t, "gen"))
            assert got.shape == golden[key].shape
            assert np.abs(got - golden[key]).max() <= 1e-9 * golden[key].max()
    gpu.close()


@pytest.mark.parametrize("ci", range(len(ref_pins.NOISE_CASES)), ids=[c[0].split(".")[0] for c in ref_pins.NOISE_CASES])
def test_kernels_reproduce_the_references_noise_mode(ab, golden, tmp_path, ci):
    """Noise mode against the reference's own NoiseMaker / noise sources / complex-weight transport: (A) a power-iteration
    generation that samples the noise source, (B) a noise generation with complex weights -- through the C++ adapter
    GPUTransporter::transport(bank, noise, &noise_bank, &noise_maker)."""
    fname, n = ref_pins.NOISE_CASES[ci]
    _noise_case(ab, golden, tmp_path, fname, n, 900 + ci, 950 + ci)


def test_kernels_reproduce_the_references_vibration_at_its_third_harmonic(ab, tmp_path):
    """The flat-vibration source observed at three times its frequency (FlatVibrationNoiseSource::C_R for n >= 3: acos, sin and the
    complex exponential, src/flat_vibration_noise_source.cpp:176-179; scripts/make_ref_pins_vibration.py)."""
    gold = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pins_vibration.npz")))
    fname, n = ref_pins.HARMONIC_NOISE_CASES[0]
    _noise_case(ab, gold, tmp_path, fname, n, 1500, 1550)
    assert len(gold["noise_noise_vibration_h3_A_source"]) > 20


def _noise_case(ab, golden, tmp_path, fname, n, seed_bank, seed_w2):
    name = fname.split(".")[0]
    deck = load_deck(fname)
    path = write_deck(deck, tmp_path / fname, {"settings": {"nparticles": n}})
    deck["settings"]["nparticles"] = n
    keff = float(deck["settings"].get("keff", 1.0))
    r, u, E, w, hid = ref_pins.transport_bank(deck, n, seed_bank, False)
    w2 = np.random.default_rng(seed_w2).uniform(-0.8, 0.8, n)
    gpu = ab.Backend(path, 0)
    for phase, noise, sample, wb in (("A", False, True, np.zeros(n)), ("B", True, False, w2)):
        bank = {k: np.ascontiguousarray(v) for k, v in zip(("x", "y", "z"), r.T)}
        bank.update({k: np.ascontiguousarray(v) for k, v in zip(("ux", "uy", "uz"), u.T)})
        bank.update(E=np.ascontiguousarray(E), wgt=np.ascontiguousarray(w), wgt2=np.ascontiguousarray(wb), id_a=hid, id_b=hid.copy(),
                    id_c=None)
        fis, nb = gpu.transport_vectors_noise(bank, k_col=1.0, keff=keff, converged=False, noise=noise, sample_noise=sample,
                                              capacity=24 * n)
        for got, what in ((fis, "sites"), (nb, "source")):
            ref9 = golden[f"noise_{name}_{phase}_{what}"]
            ref_ids = golden[f"noise_{name}_{phase}_{what}_ids" if what == "source" else f"noise_{name}_{phase}_ids"]
            assert len(got["x"]) == len(ref9), f"{phase} {what}: number of particles differs from the reference"
            if not len(ref9):
                continue
            assert np.array_equal(np.stack([got["id_a"], got["id_b"], got["id_c"]], 1), ref_ids), f"{phase} {what}: ids / order"
            assert np.array_equal(got["E"], ref9[:, 6])
            g9 = np.stack([got[k] for k in ("x", "y", "z", "ux", "uy", "uz", "E", "wgt", "wgt2")], 1)
            assert np.abs(g9[:, :6] - ref9[:, :6]).max() < 1e-9
            assert np.allclose(g9[:, 7:], ref9[:, 7:], rtol=1e-9, atol=1e-12), f"{phase} {what}: complex weights"
    gpu.close()


@pytest.mark.parametrize("resident", [True, False], ids=["resident", "host-buffers"])
@pytest.mark.parametrize("ci", range(len(ref_pins.ALL_PI_CASES)), ids=[c[0].split(".")[0] for c in ref_pins.ALL_PI_CASES])
def test_power_iteration_reproduces_the_references_power_iterator(ab, golden, tmp_path, ci, resident):
    """Whole k-eigenvalue simulations -- the C++ host's PowerIterator over the GPU transporter, bank resident in HBM or
    through host buffers -- against the reference's own PowerIterator::run(): every generation's k_col, k_trk, leakage,
    migration area and entropy, and the final averages and errors, to 1e-9 relative (floating-point sums are taken in a
    different order on the device; the histories themselves are the same)."""
    fname, n, ngen, nign = ref_pins.ALL_PI_CASES[ci]
    if ci >= len(ref_pins.POWER_ITERATION_CASES):  # the implicit-leakage tracker's and the branchless iterator's simulations
        golden = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", ref_pins.pi_golden_file(ci))))
    name = fname.split(".")[0]
    path = write_deck(load_deck(fname), tmp_path / fname, {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}})
    gpu = ab.Backend(path, 0)
    g = gpu.run_power_iteration(ngen, nign, resident=resident)
    for k in ("kcol", "ktrk", "leak", "mig", "entropy"):
        ref = golden[f"pi_{name}_{k}"]
        assert np.allclose(g[k], ref, rtol=1e-9, atol=1e-12), (k, g[k], ref)
    summ = np.array([g[k] for k in ("kcol_avg", "kcol_err", "ktrk_avg", "ktrk_err", "leak_avg", "leak_err")])
    assert np.allclose(summ, golden[f"pi_{name}_summary"], rtol=1e-7, atol=1e-12), (summ, golden[f"pi_{name}_summary"])
    for t in range(gpu.ntallies()):
        key = f"pi_{name}_tally{t}_avg"
        if key in golden:
            got = np.ravel(gpu.tally(t, "avg"))
            assert np.abs(got - golden[key]).max() <= 1e-9 * golden[key].max()
    gpu.close()


@pytest.mark.parametrize("ci", range(len(ref_pins.NOISE_DRIVER_CASES)), ids=[c[0].split(".")[0] for c in ref_pins.NOISE_DRIVER_CASES])
def test_noise_simulation_reproduces_the_references_noise_driver(ab, golden, tmp_path, ci):
    """Whole noise simulations (abeille_b200.noise.NoiseSimulation over the GPU transporter) against the reference's own
    Noise::run(): k_col of every power-iteration generation to 1e-9 and the noise-source tallies to 1e-6 of their largest
    bin.  The flux tallies integrate the inner noise generations -- hundreds of generations of long histories in a
    reflective lattice, a chain in which a last-bit difference of log / sin / cos grows (5e-7 cm after 50 inner
    generations) until a boundary-or-collision comparison flips: the oracle itself parts ways between its glibc and its
    fdlibm math mode in the second batch of the first case.  They are therefore compared to 1e-6
    where the chain stayed on the reference's path and otherwise only for their support and integral."""
    from abeille_b200.noise import NoiseSimulation
    fname, n, nb, nign, nskip = ref_pins.NOISE_DRIVER_CASES[ci]
    name = fname.split(".")[0]
    path = write_deck(load_deck(fname), tmp_path / fname,
                      {"settings": {"nparticles": n, "ngenerations": nb, "nignored": nign, "nskip": nskip}})
    sim = NoiseSimulation(path, 0)
    got = sim.run()
    ref_k = golden[f"nd_{name}_kcol"]
    assert np.allclose(got["k_col"], ref_k, rtol=1e-9), (got["k_col"], ref_k)
    on_path = 0
    for t, tname in enumerate(sim.tally_names):
        ref = golden[f"nd_{name}_tally{t}_avg"]
        a = np.ravel(sim.tally(t, "avg"))
        scale = np.abs(ref).max()
        close = np.abs(a - ref).max() <= 1e-6 * scale + 1e-300
        if "source" in tname:
            assert close, f"noise-source tally {tname}: max diff {np.abs(a - ref).max()} of {scale}"
        else:
            on_path += int(close)
            assert np.array_equal(a != 0, ref != 0) or abs(a.sum() - ref.sum()) <= 0.5 * np.abs(ref).sum(), tname
    assert len(sim.tally_names) == 4
    sim.close()


@pytest.mark.parametrize("ci", range(len(ref_pins.ALL_PI_CASES)), ids=[c[0].split(".")[0] for c in ref_pins.ALL_PI_CASES])
def test_references_power_iterator_drives_the_gpu_transporter(ab, golden, tmp_path, ci):
    """The drop-in, live: the reference's OWN PowerIterator::run() (compiled from its sources into oracle/_ref) with its
    transporter replaced by integration/gpu_transporter.hpp -- `GPUTransporter : Transporter` written against the
    reference's types, forwarding transport() to abl_transport() of libabeille_b200.so.  The reference's source sampling,
    entropy, cancellation, normalisation and statistics run unchanged; every generation's k_col, k_trk, leakage, migration
    area and entropy must equal what the reference obtained with its own CPU trackers (golden vectors) to 1e-9."""
    import subprocess
    import sys
    if not os.path.exists(ref_pins.REF_LIB):
        pytest.skip("oracle/_ref/libabeille_ref.so was not built (needs /root/reference at build time)")
    from abeille_b200 import backend
    _, host_lib = backend.lib_paths()
    fname, n, ngen, nign = ref_pins.ALL_PI_CASES[ci]
    if ci >= len(ref_pins.POWER_ITERATION_CASES):
        golden = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", ref_pins.pi_golden_file(ci))))
    name = fname.split(".")[0]
    path = write_deck(load_deck(fname), tmp_path / fname, {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}})
    out = str(tmp_path / "pi_gpu.npz")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (f"import sys; sys.path.insert(0, {root!r}); import numpy as np; from oracle import ref_pins; "
            f"np.savez({out!r}, **ref_pins.power_iteration_through_gpu_transporter({ci}, {host_lib!r}, {str(path)!r}))")
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)  # one simulation per process
    got = dict(np.load(out))
    for k in ("kcol", "ktrk", "leak", "mig", "entropy"):
        assert np.allclose(got[f"pi_{name}_{k}"], golden[f"pi_{name}_{k}"], rtol=1e-9, atol=1e-12), (k, got[f"pi_{name}_{k}"])
    assert np.allclose(got[f"pi_{name}_summary"], golden[f"pi_{name}_summary"], rtol=1e-7, atol=1e-12)
    # the device's mesh tallies (recorded / cleared by the adapter where the reference records / clears its own) against the
    # average and error of the mean the reference's MeshTally objects ended with
    for key in [k for k in got if "_tally" in k and k in golden]:
        scale = np.abs(golden[key]).max()
        assert got[key].shape == golden[key].shape
        assert np.abs(got[key] - golden[key]).max() <= 1e-9 * scale, (key, np.abs(got[key] - golden[key]).max(), scale)


@pytest.mark.parametrize("ci", range(len(ref_pins.NOISE_DRIVER_CASES)), ids=[c[0].split(".")[0] for c in ref_pins.NOISE_DRIVER_CASES])
def test_references_noise_driver_drives_the_gpu_transporter(ab, oracle_api, golden, tmp_path, ci):
    """The drop-in in noise mode: the reference's OWN Noise::initialize() + run() (compiled into oracle/_ref) over
    integration/gpu_transporter.hpp -- plain generations through abl_transport, the generations that sample the noise source
    through abl_transport_noise (transport(bank, false, &noise_bank, &noise_maker)), the noise particles' inner generations
    through abl_transport with their complex weights (transport(nbank, true)).  Source normalisation, regional cancellation of
    the noise banks and the bank hand-over are the reference's, unchanged.  k_col of every power-iteration generation against
    what the reference obtained with its own CPU trackers (1e-9); the final bank, its first history id and the history counter
    -- which counts every noise particle of every inner generation -- EXACTLY against the oracle's Noise driver in the device's
    math mode (the inner-generation chain follows fdlibm on the device and glibc in the reference: the golden counter of the
    first case differs for that reason alone, the other two agree with the golden file as well)."""
    import subprocess
    import sys
    from oracle import deck as _deck
    if not os.path.exists(ref_pins.REF_LIB):
        pytest.skip("oracle/_ref/libabeille_ref.so was not built (needs /root/reference at build time)")
    from abeille_b200 import backend
    _, host_lib = backend.lib_paths()
    fname, n, nb, nign, nskip = ref_pins.NOISE_DRIVER_CASES[ci]
    name = fname.split(".")[0]
    ov = {"settings": {"nparticles": n, "ngenerations": nb, "nignored": nign, "nskip": nskip}}
    path = write_deck(load_deck(fname), tmp_path / fname, ov)
    out = str(tmp_path / "nd_gpu.npz")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (f"import sys; sys.path.insert(0, {root!r}); import numpy as np; from oracle import ref_pins; "
            f"np.savez({out!r}, **ref_pins.noise_through_gpu_transporter({ci}, {host_lib!r}, {str(path)!r}))")
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)  # one simulation per process
    got = dict(np.load(out))
    ref_k, ref_fb = golden[f"nd_{name}_kcol"], golden[f"nd_{name}_final_bank"]
    assert np.allclose(got[f"nd_{name}_kcol"], ref_k, rtol=1e-9), (got[f"nd_{name}_kcol"], ref_k)
    fb = [int(v) for v in got[f"nd_{name}_final_bank"]]
    # ... and the same run over a transporter built from the reference's live objects (flatten_problem(): noise sources from the
    # NoiseMaker, no YAML, no host library)
    cuda_lib, _ = backend.lib_paths()
    code = (f"import sys; sys.path.insert(0, {root!r}); import numpy as np; from oracle import ref_pins; "
            f"np.savez({out!r}, **ref_pins.noise_through_gpu_transporter({ci}, {cuda_lib!r}, {str(path)!r}, from_objects=True))")
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
    flat = dict(np.load(out))
    # (k is a sum of atomically added block sums: equal to rounding between two runs; the counters are integers)
    assert np.allclose(flat[f"nd_{name}_kcol"], got[f"nd_{name}_kcol"], rtol=1e-12) and [int(v) for v in flat[f"nd_{name}_final_bank"]] == fb
    orc = oracle_api.Oracle(str(path))
    o = orc.run_noise(_deck.load_yaml(str(path))["settings"])
    assert fb == [int(v) for v in o["final_bank"]], (fb, o["final_bank"])
    assert fb[:2] == [int(v) for v in ref_fb[:2]]                        # the last power-iteration bank, as the reference's
    assert fb[2] > 1.5 * n * (nign + nb * nskip)                         # (most histories of the run are noise particles)
    if ci > 0:
        assert fb[2] == int(ref_fb[2])


@pytest.mark.parametrize("fname", ["c5g7_delta_collision.yaml", "ref_sqr_c5g7_surface_tl.yaml", "c5g7_carter_cancel.yaml", "PUa-1-0-IN.yaml",
                                   "PUa-cube_carter_exact_avgg.yaml", "UD2O-2-1-SL_branchless_split_comb.yaml", "c5g7_carter_exact_full.yaml"])
def test_references_power_iterator_drives_a_transporter_built_from_its_own_objects(ab, golden, tmp_path, fname):
    """The drop-in without this repo's host library: abl_problem comes from integration/flatten_problem.hpp -- the reference's
    settings, geometry::, materials and mesh tallies, read where they live -- and goes straight to abl_create of
    libabeille_b200.so; the reference's own PowerIterator (or BranchlessPowerIterator, or its exact cancelator) runs over that
    transporter and reproduces its CPU results to 1e-9, mesh tallies included."""
    import subprocess
    import sys
    if not os.path.exists(ref_pins.REF_LIB):
        pytest.skip("oracle/_ref/libabeille_ref.so was not built (needs /root/reference at build time)")
    from abeille_b200 import backend
    cuda_lib, _ = backend.lib_paths()
    ci = [c[0] for c in ref_pins.ALL_PI_CASES].index(fname)
    _, n, ngen, nign = ref_pins.ALL_PI_CASES[ci]
    if ci >= len(ref_pins.POWER_ITERATION_CASES):
        golden = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", ref_pins.pi_golden_file(ci))))
    name = fname.split(".")[0]
    path = write_deck(load_deck(fname), tmp_path / fname, {"settings": {"nparticles": n, "ngenerations": ngen, "nignored": nign}})
    out = str(tmp_path / "pi_flat.npz")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (f"import sys; sys.path.insert(0, {root!r}); import numpy as np; from oracle import ref_pins; "
            f"np.savez({out!r}, **ref_pins.power_iteration_through_gpu_transporter({ci}, {cuda_lib!r}, {str(path)!r}, from_objects=True))")
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
    got = dict(np.load(out))
    for k in ("kcol", "ktrk", "leak", "mig", "entropy"):
        assert np.allclose(got[f"pi_{name}_{k}"], golden[f"pi_{name}_{k}"], rtol=1e-9, atol=1e-12), (k, got[f"pi_{name}_{k}"])
    assert np.allclose(got[f"pi_{name}_summary"], golden[f"pi_{name}_summary"], rtol=1e-7, atol=1e-12)
    for key in [k for k in got if "_tally" in k and k in golden]:
        scale = np.abs(golden[key]).max()
        assert np.abs(got[key] - golden[key]).max() <= 1e-9 * scale, key
