"""CPU: host-side logic of the product (YAML reader, object model, flattening) and the C-ABI export check.
No compute call is made here (there is no GPU in the build container)."""
import copy
import ctypes
import re
import os

import numpy as np
import pytest
import yaml

from conftest import DECKS, ROOT, deck_path, load_deck, write_deck


def test_c_abi_library_exports_every_declared_symbol(native_libs):
    from abeille_b200 import backend
    header = open(os.path.join(ROOT, "include", "abeille_b200.h")).read()
    declared = set(re.findall(r"\b(abl_[a-z0-9_]+)\s*\(", header))
    assert declared == set(backend.ABI_SYMBOLS)
    lib = backend.load_backend_lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libabeille_b200.so does not export {name}"


def test_c_abi_struct_layout_matches_ctypes(native_libs):
    from abeille_b200 import backend
    assert ctypes.sizeof(backend.AblBank) == 8 + 12 * 8
    assert ctypes.sizeof(backend.AblGenParams) == 32
    assert ctypes.sizeof(backend.AblTrace) == 48


def test_no_device_fails_loudly(native_libs, tmp_path):
    """The product has no CPU fallback: opening a backend without a CUDA device must raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from abeille_b200 import Backend, BackendError
    with pytest.raises(BackendError) as e:
        Backend(deck_path("PUa-1-0-IN.yaml"), 0)
    assert "no CUDA device" in str(e.value) or "cuda" in str(e.value).lower()


def _py(node):
    """PyYAML tree -> the canonical rendering ablh_yaml_roundtrip produces (scalars as written)."""
    if node is None:
        return "null"
    if isinstance(node, dict):
        return "{" + ",".join(f'"{k}":{_py(v)}' for k, v in node.items()) + "}"
    if isinstance(node, list):
        return "[" + ",".join(_py(v) for v in node) + "]"
    return None  # scalars are compared numerically below


def _compare(a, b):
    """a: PyYAML object, b: json-ish object parsed back from the roundtrip string."""
    if isinstance(a, dict):
        assert isinstance(b, dict) and list(a.keys()) == list(b.keys())
        for k in a:
            _compare(a[k], b[k])
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b)
        for x, y in zip(a, b):
            _compare(x, y)
    elif a is None:
        assert b is None
    elif isinstance(a, bool):
        assert b in ("true", "false", "True", "False") and (b.lower() == "true") == a
    elif isinstance(a, (int, float)):
        assert float(b) == float(a)
    else:
        assert str(a) == b


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(DECKS) if f.endswith(".yaml")))
def test_yaml_lite_agrees_with_pyyaml_on_every_deck(native_libs, name):
    import json
    text = open(deck_path(name)).read()
    _compare(yaml.safe_load(text), json.loads(native_libs.yaml_roundtrip(text)))


def test_yaml_lite_block_and_flow_styles(native_libs):
    import json
    text = """
# comment
a: [1, 2,
    3]   # trailing
b:
  - {x: 1, y: 'q r'}
  - z: 2
    w: [[1, 2], [3]]
  -
    k: v
c: "-1 & +2"
d:
- 1
- 2
e: {p: [1, 2], q: {r: s}}
"""
    _compare(yaml.safe_load(text), json.loads(native_libs.yaml_roundtrip(text)))


def test_yaml_lite_rejects_bad_input(native_libs):
    from abeille_b200 import BackendError
    with pytest.raises(BackendError):
        native_libs.yaml_roundtrip("a: [1, 2")
    with pytest.raises(BackendError):
        native_libs.yaml_roundtrip("a: 1\n   b: 2\n")


def test_parse_c5g7(native_libs):
    info, smp = native_libs.parse_only(deck_path("c5g7_delta_collision.yaml"))
    assert info["ngroups"] == 7 and info["nsurfaces"] == 7 and info["ncells"] == 8 and info["nuniverses"] == 10
    assert info["nmaterials"] == 7 and info["max_stack_depth"] == 4 and info["tracking"] == 1
    # majorant: water dominates all groups but 0 and 3 (8.7% MOX) -- SURVEY appendix C
    assert np.allclose(smp, [0.183045, 0.41297, 0.59031, 0.606174, 0.718, 1.25445, 2.65038], rtol=0, atol=0)


def test_implicit_leakage_tracking_is_parsed(native_libs):
    """settings: transport: implicit-leakage-delta-tracking (src/parser.cpp:415-418) -> ABL_TRACK_IMPLICIT_LEAKAGE, sampling
    cross section = the majorant, as for delta tracking (src/implicit_leakage_delta_tracker.cpp:52-57)"""
    _, maj = native_libs.parse_only(deck_path("c5g7_delta_collision.yaml"))
    info, smp = native_libs.parse_only(deck_path("c5g7_implicit_collision.yaml"))
    assert info["tracking"] == 3 and np.array_equal(smp, maj)
    info, _ = native_libs.parse_only(deck_path("noise_oscillation_implicit.yaml"))
    assert info["tracking"] == 3 and info["mode"] == 1


def test_carter_sampling_xs_is_ratio_times_majorant(native_libs):
    _, maj = native_libs.parse_only(deck_path("c5g7_delta_collision.yaml"))
    _, smp = native_libs.parse_only(deck_path("c5g7_carter_cancel.yaml"))
    assert smp[0] == maj[0] * 0.9 and np.array_equal(smp[1:], maj[1:])


def test_flattened_tables_agree_with_the_oracle(native_libs, oracle_api):
    """The oracle reads decks through an independent path (PyYAML -> token file -> its own table builder)."""
    for name in ("c5g7_delta_collision.yaml", "PUa-1-2-SL.yaml", "UD2O-2-1-SL.yaml"):
        t = native_libs.dump_tables(deck_path(name))
        o = oracle_api.Oracle(deck_path(name))
        maj, smp = o.majorant()
        assert np.array_equal(t["smp"], smp)
        G = o.G
        assert len(t["Et"]) % G == 0 and len(t["chi_cdf"]) == len(t["Et"]) * G
        # cumulative tables end at exactly 1.0 (libstdc++ forces the last partial sum)
        if G >= 2:
            assert np.all(t["scatter_cdf"].reshape(-1, G)[:, -1] == 1.0)


def test_legendre_tables_for_anisotropic_deck(native_libs):
    t = native_libs.dump_tables(deck_path("PUa-1-2-SL.yaml"))
    n = int(t["angle"][1])
    assert n > 2  # linearised P2 distribution has interior points
    mu, pdf, cdf = t["amu"][:n], t["apdf"][:n], t["acdf"][:n]
    assert mu[0] == -1.0 and mu[-1] == 1.0 and np.all(np.diff(mu) > 0)
    assert cdf[0] == 0.0 and cdf[-1] == 1.0 and np.all(np.diff(cdf) >= 0) and np.all(pdf >= 0)


@pytest.mark.parametrize("bad,msg", [
    ({"settings": {"transport": "warp-drive"}}, "Invalid tracking method"),
    ({"settings": {"energy-mode": "continuous-energy"}}, "multi-group"),
    ({"root-universe": 12345}, "Could not find universe"),
    ({"settings": {"nignored": 5000}}, "ignored"),
    # the old `lattice: n` universe form of the shipped ref_sqr_c5g7.yaml, which the reference's own parser refuses (src/parser.cpp:302-308)
    ({"universes": [{"lattice": 1, "id": 10, "name": "old form"}], "root-universe": 10}, "Invalid universe definition."),
])
def test_parser_errors(native_libs, tmp_path, bad, msg):
    from abeille_b200 import BackendError
    path = write_deck(load_deck("c5g7_delta_collision.yaml"), tmp_path / "bad.yaml", bad)
    with pytest.raises(BackendError) as e:
        native_libs.parse_only(path)
    assert msg in str(e.value)


@pytest.mark.parametrize("energy", [7.0, 9.5, 0.0])
def test_source_energy_outside_the_group_structure_is_fatal(native_libs, tmp_path, energy):
    """Source::generate_particle redraws the energy while E <= min_energy or max_energy <= E and stops the run after 200 draws
    (src/source.cpp:48-58); a mono-energetic source can only end there, and the parser says so with the reference's words."""
    from abeille_b200 import BackendError
    deck = load_deck("c5g7_delta_collision.yaml")   # energy-bounds [0, ..., 7]
    deck["sources"][0]["energy"]["energy"] = energy
    path = write_deck(deck, tmp_path / "bad.yaml")
    with pytest.raises(BackendError) as e:
        native_libs.parse_only(path)
    assert "200 samplings of energy" in str(e.value)


def test_source_direction_distributions(native_libs, tmp_path):
    """direction: isotropic | mono-directional | cone (src/direction_distribution.cpp:36-56): the axis is normalised as
    Direction(x, y, z) does, the cone keeps cos(aperture) (src/cone.cpp:31-32); the reference's messages for what is missing."""
    from abeille_b200 import BackendError
    r = native_libs.source_records(deck_path("PUa-1-0-SL_subcritical_fs_beam.yaml"))
    assert r.shape == (4, 18) and not r[:, 15:].any()   # all mono-energetic
    assert list(r[:, 10]) == [1, 2, 2, 0] and list(r[:, 0]) == [2.0, 1.5, 0.5, 1.0] and list(r[:, 2]) == [0, 1, 0, 1]
    ax = np.array([3.0, 1.0, -0.5])
    assert np.array_equal(r[0, 11:14], ax / np.sqrt((ax * ax).sum()))
    assert r[1, 14] == np.cos(0.4) and r[2, 14] == np.cos(0.05) and np.array_equal(r[2, 11:14], [0.0, 0.0, -1.0])
    deck = load_deck("PUa-1-0-SL_subcritical_fs_beam.yaml")
    for edit, msg in ((lambda d: d["sources"][0]["direction"].pop("direction"), "No valid direction entry for mono-directional distribution."),
                      (lambda d: d["sources"][1]["direction"].pop("aperture"), "No valid aperture entry for cone distribution."),
                      (lambda d: d["sources"][1]["direction"].update(direction=[1.0, 0.0]), "No valid direction entry for cone distribution."),
                      (lambda d: d["sources"][3]["direction"].update(type="lambertian"), "Invalid direction distribution type lambertian.")):
        bad = copy.deepcopy(deck)
        edit(bad)
        with pytest.raises(BackendError) as e:
            native_libs.source_records(write_deck(bad, tmp_path / "bad.yaml"))
        assert msg in str(e.value)


def test_source_energy_distributions(native_libs, tmp_path):
    """energy: mono-energetic | maxwellian | watt (src/energy_distribution.cpp:36-58, src/maxwellian.cpp:44-55, src/watt.cpp:49-66);
    `tabulated` needs PapillonNDL's PCTable and is refused by name."""
    from abeille_b200 import BackendError
    r = native_libs.source_records(deck_path("UD2O-2-1-SL_subcritical_fs_spectra.yaml"))
    assert r.shape == (3, 18)
    assert list(r[:, 15]) == [1, 2, 0] and list(r[:, 16]) == [0.45, 0.35, 0.0] and list(r[:, 17]) == [0.0, 2.0, 0.0] and r[2, 9] == 1.5
    assert list(r[:, 10]) == [0, 2, 0] and r[1, 14] == np.cos(0.6)
    deck = load_deck("UD2O-2-1-SL_subcritical_fs_spectra.yaml")
    for edit, msg in ((lambda d: d["sources"][0]["energy"].pop("a"), 'No valid "a" entry in maxwellian distribution.'),
                      (lambda d: d["sources"][0]["energy"].update(a=-1.0), "Maxwellian parameter a must be >= 0."),
                      (lambda d: d["sources"][1]["energy"].pop("b"), 'No valid "b" entry in watt distribution.'),
                      (lambda d: d["sources"][1]["energy"].update(b=0.0), "Watt parameter b must be >= 0."),
                      (lambda d: d["sources"][2]["energy"].update(type="tabulated"), "tabulated"),
                      (lambda d: d["sources"][2]["energy"].update(type="thermal"), "Invalid energy distribution type thermal.")):
        bad = copy.deepcopy(deck)
        edit(bad)
        with pytest.raises(BackendError) as e:
            native_libs.source_records(write_deck(bad, tmp_path / "bad.yaml"))
        assert msg in str(e.value)


def test_unknown_surface_in_region_is_rejected(native_libs, tmp_path):
    from abeille_b200 import BackendError
    deck = load_deck("PUa-1-0-IN.yaml")
    deck["cells"][0]["region"] = "+1 & -99"
    path = write_deck(deck, tmp_path / "bad.yaml")
    with pytest.raises(BackendError) as e:
        native_libs.parse_only(path)
    assert "surface" in str(e.value)


def test_cpp_multi_gpu_driver_is_built_and_links():
    """abeille_b200/lib/abl_pi_nccl (host/distributed_main.cpp: the sharded generation loop in C++, NCCL called directly) resolves its
    libraries and explains itself when started without arguments; it is exercised on 2 and 8 GPUs by scripts/nccl_pi_check.py
    (profiles/t9b_*, t9c_*, t9d_*)."""
    import subprocess
    from abeille_b200 import backend
    cuda_lib, _ = backend.lib_paths()
    binary = os.path.join(os.path.dirname(cuda_lib), "abl_pi_nccl")
    if not os.path.exists(binary):
        pytest.skip("abl_pi_nccl was not built (nccl.h / libnccl missing at build time)")
    p = subprocess.run([binary], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "usage: abl_pi_nccl" in p.stderr
