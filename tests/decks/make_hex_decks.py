"""Generates the hexagonal-lattice deck fixtures (HexLattice, src/hex_lattice.cpp): the C5G7 pin universes and materials of
tests/decks/c5g7_delta_collision.yaml arranged in a four-ring hexagonal lattice inside a moderator box with reflective and
vacuum faces.  The reference ships no multigroup hexagonal deck; these exercise the class through the YAML keys its factory
reads (make_hex_lattice: shape [rings, nz], pitch [p, pz], origin, top, universes, outer).

    python tests/decks/make_hex_decks.py
"""
import copy
import os

import yaml

OUT = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(OUT, "c5g7_delta_collision.yaml")) as f:
    base = yaml.safe_load(f)


def hexdeck(top, origin, transport, estimator, half=9.0):
    d = copy.deepcopy(base)
    d["surfaces"] = [s for s in d["surfaces"] if s["id"] == 1] + [
        {"type": "xplane", "x0": -half, "boundary": "reflective", "id": 11}, {"type": "xplane", "x0": half, "boundary": "vacuum", "id": 12},
        {"type": "yplane", "y0": -half, "boundary": "reflective", "id": 13}, {"type": "yplane", "y0": half, "boundary": "vacuum", "id": 14},
        {"type": "zplane", "z0": -20.0, "boundary": "reflective", "id": 15}, {"type": "zplane", "z0": 20.0, "boundary": "vacuum", "id": 16}]
    rings = 4
    nhex = 1 + sum(6 * r for r in range(1, rings))
    pattern = [1, 2, 3, 6, 1, 4, 1, 5, 2, 1, 3]
    ids = [pattern[i % len(pattern)] for i in range(nhex)]
    ids[5] = -1  # one empty position: the outer universe shows through
    d["universes"] = [u for u in d["universes"] if "cells" in u] + [
        {"id": 20, "type": "hexagonal", "name": "hex core", "shape": [rings, 1], "pitch": [1.4, 40.0], "origin": origin, "top": top,
         "outer": 7, "universes": ids}]
    d["root-universe"] = 20
    d["sources"] = [{"spatial": {"type": "box", "low": [-4.0, -4.0, -15.0], "hi": [4.0, 4.0, 15.0], "fissile-only": True},
                     "direction": {"type": "isotropic"}, "energy": {"type": "mono-energetic", "energy": 6.5}, "weight": 1.0}]
    d["tallies"] = [{"name": "flux_hex", "low": [-half, -half, -20.0], "hi": [half, half, 20.0], "shape": [24, 24, 2],
                     "energy-bounds": [0, 1, 2, 3, 4, 5, 6, 7], "quantity": "flux", "estimator": estimator}]
    d["entropy"] = {"low": [-half, -half, -20.0], "hi": [half, half, 20.0], "shape": [4, 4, 2]}
    d["settings"].update({"transport": transport, "nparticles": 4000, "ngenerations": 8, "nignored": 3})
    return d


def dump(deck, name, header):
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        f.write("# " + header.replace("\n", "\n# ") + "\n")
        yaml.safe_dump(deck, f, default_flow_style=None, sort_keys=False, width=200)
    print("wrote", path)


dump(hexdeck("pointy", [0.0, 0.0, 0.0], "delta-tracking", "collision"), "hex_delta_collision.yaml",
     "HexLattice, pointy top, origin at zero: C5G7 pins and materials in a four-ring hexagonal lattice, delta tracking,\n"
     "collision-estimator flux tally (tests/decks/make_hex_decks.py).")
dump(hexdeck("flat", [0.35, -0.2, 0.0], "delta-tracking", "collision"), "hex_delta_flat_offset.yaml",
     "HexLattice, flat top, origin off zero, delta tracking: the reference shifts by the origin in get_cell but not in the\n"
     "Tracker's tile check (hex_lattice.cpp:140-150 against tracker.hpp:262) -- reproduced as is (tests/decks/make_hex_decks.py).")
# Surface tracking through a HexLattice is not a fixture: the reference's own SurfaceTracker does not terminate on these decks
# (HexLattice::get_tile has no tie-break at a tile boundary, unlike RectLattice::get_tile: a particle that lands on one is
# handed back to the tile it came from and crosses the same boundary again).  The device and oracle code for
# distance_to_tile_boundary_{pointy,flat} is written from hex_lattice.cpp:347-455 but cannot be pinned against a run.
