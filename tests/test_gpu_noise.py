"""GPU: noise mode (BASELINE config 5, input_files/noise_oscillation.yaml) against the CPU oracle.

Three layers, as for the k-eigenvalue path: (1) one power-iteration generation that samples the noise source
(NoiseMaker::sample_noise_source) -- fission bank and noise bank bit for bit; (2) one transport call of noise particles
(complex weights, noise copies, delayed factor) -- per-history outcomes and the next noise bank bit for bit; (3) the whole
Noise::run driver -- generation counts and particle counts identical, tallies to 1e-9 (sums accumulated by atomics and a
device tree reduction for the mean |w| differ from the serial CPU sums in rounding order)."""
import numpy as np
import pytest
import yaml

from conftest import deck_path, load_deck, write_deck

pytestmark = pytest.mark.gpu

F64 = ("x", "y", "z", "ux", "uy", "uz", "E", "wgt", "wgt2")
U64 = ("id_a", "id_b", "id_c")


@pytest.fixture(scope="module")
def ab(native_libs):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return native_libs


def _to_device(gpu, bank, cap):
    import torch
    d = gpu.new_device_bank(cap)
    n = len(bank["x"])
    for k in F64:
        d[k][:n] = torch.from_numpy(np.ascontiguousarray(bank[k] if bank.get(k) is not None else np.zeros(n)))
    for k in U64:
        if bank.get(k) is not None:
            d[k][:n] = torch.from_numpy(bank[k].view(np.int64).copy())
    return d


def _from_device(d, n):
    out = {k: d[k][:n].cpu().numpy() for k in F64}
    out.update({k: d[k][:n].cpu().numpy().view(np.uint64) for k in U64})
    return out


def _assert_banks_equal(g, o, what):
    assert len(g["x"]) == len(o["x"]), (what, len(g["x"]), len(o["x"]))
    for k in F64 + U64:
        bad = np.nonzero(g[k] != o[k])[0]
        assert bad.size == 0, f"{what}: field {k} differs for {bad.size} entries, first {bad[:4]}: {g[k][bad[:2]]} vs {o[k][bad[:2]]}"


def _particles(fis, first):
    b = {k: fis[k].copy() for k in F64}
    m = len(fis["x"])
    b["id_a"] = np.arange(first, first + m, dtype=np.uint64)
    b["id_b"] = fis["id_c"].copy()
    b["id_c"] = None
    return b


@pytest.mark.parametrize("deck", ["noise_oscillation.yaml", "noise_oscillation_delta.yaml", "noise_vibration.yaml",
                                  "noise_oscillation_implicit.yaml", "noise_vibration_h3.yaml"])
def test_noise_source_and_noise_transport_bit_exact(ab, oracle_api, tmp_path, deck):
    path = write_deck(load_deck(deck), tmp_path / deck, {"settings": {"nparticles": 6000}})
    keff = float(load_deck(deck)["settings"]["keff"])
    orc, gpu = oracle_api.Oracle(path), ab.Backend(path, 0)
    orc.set_keff(keff)
    n = 6000
    bank = orc.sample_source(n)
    cap = 8 * n
    # (1) power-iteration generation that samples the noise source
    orc.set_converged(True)
    orc.set_kcol(1.0)
    orc.set_trace(True)
    ofis, onoise, oscores = orc.transport_noise({k: v.copy() for k, v in bank.items()}, False, True)
    otr = orc.trace(n)
    dbank, dout, dnoise = _to_device(gpu, bank, n), gpu.new_device_bank(cap), gpu.new_device_bank(cap)
    m, mn, gscores, _ = gpu.transport_noise_device(dbank, n, dout, dnoise, k_col=1.0, keff=keff, converged=True, noise=False,
                                                   sample_noise=True, trace=True, use_rng_state=True)
    gtr = gpu.trace(n)
    assert mn == len(onoise["x"]) and mn > 0, (mn, len(onoise["x"]))
    for k in ("flights", "real", "virtual", "hash", "rng_state"):
        assert np.array_equal(gtr[k], otr[k]), f"per-history {k} differs"
    _assert_banks_equal(_from_device(dout, m), ofis, "fission bank")
    _assert_banks_equal(_from_device(dnoise, mn), onoise, "noise-source bank")
    assert np.allclose(gscores, oscores, rtol=1e-11, atol=1e-300)
    # (2) two generations of noise particles (complex weights)
    cur = _particles(onoise, n)
    first = n + len(cur["x"])
    for gen in range(2):
        nn = len(cur["x"])
        ofis2, _, _ = orc.transport_noise({k: (v.copy() if v is not None else None) for k, v in cur.items()}, True, False)
        otr2 = orc.trace(nn)
        dcur, dout2 = _to_device(gpu, cur, nn), gpu.new_device_bank(cap)
        m2, _, _, _ = gpu.transport_noise_device(dcur, nn, dout2, None, k_col=1.0, keff=keff, converged=True, noise=True, trace=True)
        gtr2 = gpu.trace(nn)
        for k in ("flights", "real", "virtual", "hash", "rng_state"):
            assert np.array_equal(gtr2[k], otr2[k]), f"noise generation {gen}: per-history {k} differs"
        _assert_banks_equal(_from_device(dout2, m2), ofis2, f"noise fission bank, generation {gen}")
        cur = _particles(ofis2, first)
        first += m2
    # mesh tallies scored so far (real / imaginary flux with complex weights)
    for t in range(gpu.ntallies()):
        assert np.allclose(gpu.tally(t, "gen"), orc.tally(t, "gen"), rtol=1e-9, atol=1e-300), f"tally {t}"


@pytest.mark.parametrize("deck", ["noise_oscillation.yaml", "noise_vibration.yaml"])
def test_noise_run_matches_oracle(ab, oracle_api, tmp_path, deck):
    from abeille_b200.noise import NoiseSimulation
    path = deck_path(deck)
    st = yaml.safe_load(open(path))["settings"]
    orc = oracle_api.Oracle(path)
    ref = orc.run_noise(st)
    sim = NoiseSimulation(path, 0)
    got = sim.run()
    assert np.allclose(got["k_col"], ref["k_col"], rtol=1e-12), (got["k_col"], ref["k_col"])
    assert got["noise_generations"] == ref["noise_generations"]
    assert got["noise_particles"] == ref["noise_particles"]
    for t in range(orc.ntallies()):
        a, b = sim.tally(t, "avg"), orc.tally(t, "avg")
        scale = np.abs(b).max()
        assert np.allclose(a, b, rtol=1e-7, atol=1e-9 * scale), f"tally {t}: max diff {np.abs(a - b).max()} of {scale}"
    sim.write_results(str(tmp_path / "out"))
    for t, name in enumerate(sim.tally_names):
        assert np.array_equal(np.load(tmp_path / "out" / f"{name}_avg.npy"), sim.tally(t, "avg"))
        assert np.load(tmp_path / "out" / f"{name}_std.npy").shape == orc.tally_shape(t)
    sim.close()


def test_cpp_adapter_noise_bank_matches_device_entry_point(ab, oracle_api, tmp_path):
    """GPUTransporter::transport(bank, false, &noise_bank, &noise_maker) and transport(bank, true) through the C++ adapter
    (host vectors, abl_transport_noise / abl_transport) against the oracle."""
    deck = "noise_oscillation.yaml"
    n = 3000
    path = write_deck(load_deck(deck), tmp_path / deck, {"settings": {"nparticles": n}})
    keff = float(load_deck(deck)["settings"]["keff"])
    orc, gpu = oracle_api.Oracle(path), ab.Backend(path, 0)
    orc.set_keff(keff)
    orc.set_kcol(1.0)
    bank = orc.sample_source(n)
    ofis, onoise, _ = orc.transport_noise({k: v.copy() for k, v in bank.items()}, False, True)
    gfis, gnoise = gpu.transport_vectors_noise(bank, k_col=1.0, keff=keff, sample_noise=True)
    _assert_banks_equal(gfis, ofis, "adapter fission bank")
    _assert_banks_equal(gnoise, onoise, "adapter noise bank")
    cur = _particles(onoise, n)
    ofis2, _, _ = orc.transport_noise({k: (v.copy() if v is not None else None) for k, v in cur.items()}, True, False)
    gfis2, _ = gpu.transport_vectors_noise(cur, k_col=1.0, keff=keff, noise=True)
    _assert_banks_equal(gfis2, ofis2, "adapter noise fission bank")


# ---- the noise simulation sharded over two ranks (BASELINE config 5: "sharded across 8 x B200") ---------------------------------
# Two processes, one rank each, both on GPU 0 with the gloo backend (NCCL refuses two ranks on one device; on the multi-GPU
# box the same class runs over NCCL): bank slices, history ids and the collectives are exactly those of a two-GPU run.
def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sharded_noise_worker(rank, world, port, path, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from abeille_b200.noise import DistributedNoiseSimulation
        sim = DistributedNoiseSimulation(path, 0)
        got = sim.run()
        out = {"k_col": got["k_col"], "noise_generations": got["noise_generations"], "noise_particles": got["noise_particles"],
               "history_counter": sim.history_counter, "tallies": [sim.tally(t, "avg") for t in range(sim.gpu.ntallies())]}
        sim.close()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("deck", ["noise_oscillation.yaml"])
def test_noise_run_sharded_over_two_ranks_matches_one_rank(ab, tmp_path, deck):
    import torch.multiprocessing as mp
    from abeille_b200.noise import NoiseSimulation
    path = deck_path(deck)
    one = NoiseSimulation(path, 0)
    ref = one.run()
    ref_tallies = [one.tally(t, "avg") for t in range(one.gpu.ntallies())]
    ref_counter = one.history_counter
    one.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_noise_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        got = res[r]
        # integer outcomes do not depend on the number of ranks
        assert got["noise_generations"] == ref["noise_generations"] and got["noise_particles"] == ref["noise_particles"]
        assert got["history_counter"] == ref_counter
        assert np.allclose(got["k_col"], ref["k_col"], rtol=1e-12)
        for t, (a, b) in enumerate(zip(got["tallies"], ref_tallies)):
            scale = np.abs(b).max()
            assert np.allclose(a, b, rtol=1e-7, atol=1e-9 * scale), f"rank {r} tally {t}: max diff {np.abs(a - b).max()} of {scale}"
