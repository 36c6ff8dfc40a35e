"""Generates the deck fixtures under tests/decks/ from the reference's shipped input files.

Run in the BUILD container only (it reads /root/reference/input_files, which does not exist on the GPU
box); the generated YAML files are committed.  Every derived deck states what was changed and why
(SURVEY.md section 0 / 8d lists the discrepancies between BASELINE.json's configs and the shipped files).

    python tests/decks/make_decks.py
"""
import copy
import os

import yaml

REF = "/root/reference/input_files"
OUT = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with open(os.path.join(REF, name)) as f:
        return yaml.safe_load(f)


def dump(deck, name, header):
    deck = copy.deepcopy(deck)
    deck.pop("plots", None)  # plotting is out of scope
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        f.write("# " + header.replace("\n", "\n# ") + "\n")
        yaml.safe_dump(deck, f, default_flow_style=None, sort_keys=False, width=200)
    print("wrote", path)


# S1: Sood PUa-1-0-IN, verbatim (k_inf = 2.612903, input_files/PUa-1-0-IN.yaml:1-2)
for name in ("PUa-1-0-IN", "PUb-1-0-IN", "PUa-1-0-SL", "PUb-1-0-SL", "PUa-1-1-SL", "PUa-1-2-SL", "Ua-1-1-CY", "Ua-1-1-IN",
             "UD2O-2-1-SL"):
    d = load(name + ".yaml")
    dump(d, name + ".yaml", f"{name}: values of reference input_files/{name}.yaml, re-serialised (Sood analytic benchmark)")

# S2: c5g7 geometry + materials verbatim, delta tracking, COLLISION-estimator flux tally on a pin mesh
c5 = load("c5g7.yaml")
c5.pop("cancelator", None)      # shipped deck configures basic-exact (out of scope) but never enables it
c5.pop("sampling-xs", None)     # key the parser never reads (src/parser.cpp:426 wants sampling-xs-ratio)
s2 = copy.deepcopy(c5)
s2["tallies"] = [{"name": "flux_pin", "low": [-32.13, -32.13, -107.1], "hi": [32.13, 32.13, 107.1], "shape": [51, 51, 1],
                  "energy-bounds": [0, 1, 2, 3, 4, 5, 6, 7], "quantity": "flux", "estimator": "collision"}]
s2["entropy"] = {"low": [-32.13, -32.13, -107.1], "hi": [32.13, 32.13, 107.1], "shape": [8, 8, 4]}
s2["settings"].update({"transport": "delta-tracking", "nparticles": 100000, "ngenerations": 2500, "nignored": 500})
dump(s2, "c5g7_delta_collision.yaml",
     "S2: reference input_files/c5g7.yaml geometry/materials/source, transport: delta-tracking,\n"
     "collision-estimator flux tally on a 51x51x1 pin mesh (BASELINE config 2), entropy mesh added.")

# S2-full: the shipped 1224x1224x10x7 mesh, collision estimator (throughput configuration)
s2f = copy.deepcopy(s2)
s2f["tallies"] = [{"name": "flux_c5g7", "low": [-32.13, -32.13, -107.1], "hi": [32.13, 32.13, 107.1], "shape": [1224, 1224, 10],
                   "energy-bounds": [0, 1, 2, 3, 4, 5, 6, 7], "quantity": "flux", "estimator": "collision"}]
dump(s2f, "c5g7_delta_collision_fullmesh.yaml",
     "S2 (throughput): as c5g7_delta_collision.yaml with the shipped 1224x1224x10 mesh (104.9 M bins, 839 MB per array).")

# shipped tally (track-length, full mesh) kept as a variant of S2
s2t = copy.deepcopy(s2)
s2t["tallies"] = [{"name": "flux_tl", "low": [-32.13, -32.13, -107.1], "hi": [32.13, 32.13, 107.1], "shape": [102, 102, 4],
                   "energy-bounds": [0, 1, 2, 3, 4, 5, 6, 7], "quantity": "flux", "estimator": "track-length"}]
dump(s2t, "c5g7_delta_tracklength.yaml", "c5g7 delta-tracking with a track-length flux tally (the estimator the shipped deck uses) on a 102x102x4 mesh.")

# S3: ref_sqr_c5g7 re-expressed in the current input schema (lattices as pitch-keyed universes),
#     transport: surface-tracking, estimator: track-length
rs = load("ref_sqr_c5g7.yaml")
lat = {l["id"]: l for l in rs.pop("lattices")}
unis = []
for u in rs["universes"]:
    if "lattice" in u:
        l = copy.deepcopy(lat[u["lattice"]])
        l["id"] = u["id"]
        l["name"] = u.get("name", l.get("name", ""))
        unis.append(l)
    else:
        unis.append(u)
rs["universes"] = unis
rs.pop("cancelator", None)
rs["tallies"] = [{"name": "flux", "low": [0., 0., 0.], "hi": [64.26, 64.26, 214.2], "shape": [102, 102, 5],
                  "energy-bounds": [0., 1., 2., 3., 4., 5., 6., 7.], "quantity": "flux", "estimator": "track-length"}]
rs["settings"].pop("sourcefile", None)
rs["settings"].update({"transport": "surface-tracking", "nparticles": 100000})
dump(rs, "ref_sqr_c5g7_surface_tl.yaml",
     "S3: reference input_files/ref_sqr_c5g7.yaml (square pins from 4 planes, complement moderator cell) re-expressed in the\n"
     "current schema (src/parser.cpp:295-313 accepts `cells` or `pitch` universes only), transport: surface-tracking,\n"
     "track-length flux tally (BASELINE config 3; the shipped tally key `estimatory` is a typo). Mesh reduced to 102x102x5.")

# S4: c5g7 carter tracking, under-estimated sampling xs in group 0, approximate mesh cancellation
s4 = copy.deepcopy(s2)
s4["sampling-xs-ratio"] = [0.9, 1., 1., 1., 1., 1., 1.]
s4["cancelator"] = {"type": "approximate", "shape": [170, 170, 765], "low": [-32.13, -32.13, -107.11], "hi": [10.71, 10.71, 85.69]}
s4["settings"].update({"transport": "carter-tracking", "cancellation": True})
dump(s4, "c5g7_carter_cancel.yaml",
     "S4: c5g7 with transport: carter-tracking, sampling-xs-ratio [0.9,1,...] (under-estimated majorant in group 0),\n"
     "approximate mesh cancellation on the shipped cancelator mesh (BASELINE config 4).")

# c5g7 with surface tracking (exercises lattice + cylinder distances)
s5 = copy.deepcopy(s2t)
s5["settings"].update({"transport": "surface-tracking"})
dump(s5, "c5g7_surface_tracklength.yaml", "c5g7 geometry with transport: surface-tracking and a track-length flux tally.")

# S5: noise_oscillation.yaml verbatim (frequency-domain noise, square-oscillation source, approximate cancellation of the
# noise fission banks) with run sizes small enough for a parity test; plus a delta-tracking variant of the same problem
ns = load("noise_oscillation.yaml")
ns["settings"].update({"nparticles": 4000, "ngenerations": 2, "nignored": 2, "nskip": 2})
dump(ns, "noise_oscillation.yaml",
     "S5: reference input_files/noise_oscillation.yaml (BASELINE config 5), run sizes reduced for the parity tests\n"
     "(nparticles 100000 -> 4000, 2000 noise batches -> 2, nignored 10 -> 2, nskip 3 -> 2).")
nsd = copy.deepcopy(ns)
nsd["settings"]["transport"] = "delta-tracking"
dump(nsd, "noise_oscillation_delta.yaml", "S5 variant: noise_oscillation.yaml with transport: delta-tracking.")

# noise_vibration.yaml verbatim (flat-vibration noise sources: a fuel pin moving in y), run sizes reduced
nv = load("noise_vibration.yaml")
nv["settings"].update({"nparticles": 4000, "ngenerations": 2, "nignored": 2, "nskip": 2})
dump(nv, "noise_vibration.yaml",
     "reference input_files/noise_vibration.yaml (flat-vibration noise sources), run sizes reduced for the parity tests\n"
     "(nparticles 100000 -> 4000, 2100 noise batches -> 2, nignored 10 -> 2, nskip 3 -> 2).")
