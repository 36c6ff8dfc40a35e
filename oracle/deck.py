"""oracle/deck.py -- TEST INFRASTRUCTURE.

Independent input path for the CPU oracle: reads an Abeille YAML deck with
PyYAML and writes the flat token file `orc_main.cpp:load_problem` reads.  It is
deliberately separate from the product's C++ YAML reader
(abeille_b200/host/yaml_lite.*), so that a parsing or flattening bug on either
side shows up as a parity failure instead of cancelling out.

Keys follow the reference parser:
  settings   src/parser.cpp:341-866      materials  src/mg_nuclide.cpp:579-922
  surfaces   src/parser.cpp:248-293      cells      src/cell.cpp:301-452
  universes  src/parser.cpp:295-313, src/rect_lattice.cpp:312-424
  sources    src/source.cpp:92-140       tallies    src/collision_mesh_tally.cpp:150-270
  cancelator src/cancelator.cpp:32-78    entropy    src/parser.cpp:1008-1054
"""
from __future__ import annotations

import copy
import yaml

EST = {"collision": 0, "track-length": 1, "source": 2}
QTY = {"flux": 0, "total": 1, "elastic": 2, "absorption": 3, "fission": 4, "mt": 5,
       "real-flux": 6, "imag-flux": 7, "source": 8, "real-source": 9, "imag-source": 10}

SURF_PARAMS = {
    "xplane": ["x0"], "yplane": ["y0"], "zplane": ["z0"], "plane": ["A", "B", "C", "D"],
    "xcylinder": ["y0", "z0", "r"], "ycylinder": ["x0", "z0", "r"], "zcylinder": ["x0", "y0", "r"],
    "cylinder": ["x0", "y0", "z0", "u0", "v0", "w0", "r"], "sphere": ["x0", "y0", "z0", "r"],
}


def load_yaml(path):
    with open(path) as f:
        return yaml.safe_load(f)


def apply_overrides(deck: dict, overrides: dict | None) -> dict:
    """Shallow section overrides: {'settings': {...}, 'tallies': [...], 'sampling-xs-ratio': [...]}"""
    deck = copy.deepcopy(deck)
    for k, v in (overrides or {}).items():
        if isinstance(v, dict) and isinstance(deck.get(k), dict):
            deck[k].update(v)
        else:
            deck[k] = v
    return deck


def _f(x):
    return repr(float(x))


def _fl(xs):
    return " ".join(_f(x) for x in xs)


def deck_to_text(deck: dict) -> str:
    st = deck["settings"]
    if st.get("energy-mode") != "multi-group":
        raise ValueError("oracle handles multi-group decks only")
    G = int(st["ngroups"])
    out = ["ORCDECK 1"]
    sim = st["simulation"]
    if sim not in ("k-eigenvalue", "noise", "modified-fixed-source", "fixed-source", "branchless-k-eigenvalue"):
        raise ValueError(f"unsupported simulation {sim}")
    tr = {"surface-tracking": "surface", "delta-tracking": "delta", "carter-tracking": "carter",
          "implicit-leakage-delta-tracking": "implicit"}[
        st.get("transport", "surface-tracking")]
    mode = {"noise": "noise", "modified-fixed-source": "mfs", "fixed-source": "fs"}.get(sim, "k")
    if sim == "branchless-k-eigenvalue":  # parser.cpp:367-409; defaults settings.cpp:87-89
        mode = (f"branchless {int(bool(st.get('branchless-material', True)))} {int(bool(st.get('branchless-splitting', False)))} "
                f"{int(bool(st.get('branchless-combing', True)))}")
    out.append(f"mode {mode} tracking {tr}")
    out.append(f"ngroups {G}")
    eb = st["energy-bounds"]
    assert len(eb) == G + 1
    out.append("ebounds " + _fl(eb))
    out.append(f"nparticles {int(st['nparticles'])} ngenerations {int(st['ngenerations'])} "
               f"nignored {int(st.get('nignored', 0))} nskip {int(st.get('nskip', 10))}")
    out.append(f"wgt {_f(st.get('wgt-cutoff', 0.25))} {_f(st.get('wgt-survival', 1.0))} {_f(st.get('wgt-split', 2.0))}")
    out.append(f"seed {int(st.get('seed', 19073486328125))} stride {int(st.get('stride', 152917))}")
    if tr == "carter":
        ratios = deck["sampling-xs-ratio"]
        assert len(ratios) == G
    else:
        ratios = []
    out.append(f"ratios {len(ratios)} " + _fl(ratios))
    ncg = int(st.get("cancel-noise-gens", 2147483647))
    out.append(f"cancel {int(bool(st.get('cancellation', False)))} {int(bool(st.get('noise-cancellation', False)))} {ncg}")
    out.append(f"noise {_f(st.get('noise-angular-frequency', -1.0))} {_f(st.get('keff', 1.0))} "
               f"{int(bool(st.get('inner-generations', True)))} {int(bool(st.get('normalize-noise-source', True)))}")

    mats = deck["materials"]
    out.append(f"nmat {len(mats)}")
    for m in mats:
        out.append(f"mat {int(m['id'])}")
        out.append("total " + _fl(m["total"]))
        out.append("absorption " + _fl(m["absorption"]))
        fis = m.get("fission", [0.0] * G)
        out.append("fission " + _fl(fis))
        if "nu" in m:
            nup, nud = m["nu"], [0.0] * G
        elif "nu_prompt" in m and "nu_delayed" in m:
            nup, nud = m["nu_prompt"], m["nu_delayed"]
        else:
            nup, nud = [0.0] * G, [0.0] * G
        out.append("nu_p " + _fl(nup))
        out.append("nu_d " + _fl(nud))
        out.append("speeds " + _fl(m.get("group-speeds", [1.0] * G)))
        chi = m.get("chi", [[0.0] * G])
        out.append(f"chi {len(chi)}")
        for row in chi:
            assert len(row) == G
            out.append(_fl(row))
        out.append("scatter")
        for row in m["scatter"]:
            assert len(row) == G
            out.append(_fl(row))
        legs = [l for l in range(1, 6) if f"P{l}" in m]
        out.append(f"nleg {len(legs)}")
        for l in legs:
            out.append(f"P {l}")
            for row in m[f"P{l}"]:
                out.append(_fl(row))
        dg = m.get("delayed_groups")
        if dg:
            out.append(f"ndg {len(dg['probabilities'])} " + _fl(dg["probabilities"]) + " " + _fl(dg["constants"]))
        else:
            out.append("ndg 0")

    surfs = deck["surfaces"]
    out.append(f"nsurf {len(surfs)}")
    for s in surfs:
        names = SURF_PARAMS[s["type"]]
        vals = [s.get(n, 0.0) for n in names]
        out.append(f"surf {int(s['id'])} {s['type']} {s.get('boundary', 'normal')} {len(vals)} " + _fl(vals))

    cells = deck["cells"]
    out.append(f"ncell {len(cells)}")
    for c in cells:
        region = str(c["region"]).replace(" ", "")
        if "material" in c:
            out.append(f"cell {int(c['id'])} m {int(c['material'])} {region}")
        else:
            out.append(f"cell {int(c['id'])} u {int(c['universe'])} {region}")

    unis = deck["universes"]
    out.append(f"nuni {len(unis)}")
    for u in unis:
        if "cells" in u:
            out.append(f"uni {int(u['id'])} cells {len(u['cells'])} " + " ".join(str(int(c)) for c in u["cells"]))
        elif "pitch" in u and u.get("type") == "hexagonal":  # make_hex_lattice, src/hex_lattice.cpp:457-577
            sh, pt, org = u["shape"], u["pitch"], u["origin"]
            top = {"pointy": 0, "flat": 1}[u.get("top", "pointy")]
            ids = u["universes"]
            out.append(f"uni {int(u['id'])} hex {int(sh[0])} {int(sh[1])} {_fl(pt)} {_fl(org)} {top} "
                       f"{int(u.get('outer', -1))} {len(ids)} " + " ".join(str(int(i)) for i in ids))
        elif "pitch" in u:
            if u.get("type", "rectlinear") != "rectlinear":
                raise ValueError("oracle: only rectlinear and hexagonal lattices")
            sh, pt = u["shape"], u["pitch"]
            org = u.get("origin", [0.0, 0.0, 0.0])
            ids = u["universes"]
            assert len(ids) == sh[0] * sh[1] * sh[2]
            out.append(f"uni {int(u['id'])} rect {int(sh[0])} {int(sh[1])} {int(sh[2])} {_fl(pt)} {_fl(org)} "
                       f"{int(u.get('outer', -1))} {len(ids)} " + " ".join(str(int(i)) for i in ids))
        else:
            raise ValueError("universe needs cells or pitch")
    out.append(f"root {int(deck['root-universe'])}")

    srcs = deck.get("sources", [])
    out.append(f"nsrc {len(srcs)}")
    for s in srcs:
        sp = s["spatial"]
        ee = s["energy"]
        if ee["type"] == "mono-energetic":
            e_src = float(ee["energy"])
            if e_src <= float(eb[0]) or float(eb[-1]) <= e_src:  # Source::generate_particle's rejection loop (src/source.cpp:48-58)
                raise ValueError("Exceded 200 samplings of energy.")
            en = f"energy {_f(ee['energy'])}"
        elif ee["type"] == "maxwellian":  # src/maxwellian.cpp
            en = f"energy 0 maxwellian {_f(ee['a'])}"
        elif ee["type"] == "watt":  # src/watt.cpp
            en = f"energy 0 watt {_f(ee['a'])} {_f(ee['b'])}"
        else:
            raise ValueError("oracle: mono-energetic, maxwellian and watt sources only")
        dd = s["direction"]
        if dd["type"] == "isotropic":
            dirn = "dir iso"
        elif dd["type"] == "mono-directional":  # src/mono_directional.cpp:28-42
            dirn = "dir mono " + _fl(dd["direction"])
        elif dd["type"] == "cone":  # src/cone.cpp:46-65 (aperture in radians)
            dirn = "dir cone " + _fl(dd["direction"]) + " " + _f(dd["aperture"])
        else:
            raise ValueError("Invalid direction distribution type " + str(dd["type"]) + ".")
        fo = int(bool(s.get("fissile-only", False)))  # read at SOURCE level only (src/source.cpp:104-110)
        if sp["type"] == "box":
            pos = "box " + _fl(sp["low"]) + " " + _fl(sp["hi"])
        elif sp["type"] == "point":
            pos = "point " + _fl(sp["position"])
        else:
            raise ValueError("oracle: box/point sources only")
        out.append(f"src {_f(s['weight'])} {fo} {pos} {en} {dirn}")

    tallies = deck.get("tallies", []) or []
    out.append(f"ntally {len(tallies)}")
    for t in tallies:
        est = t.get("estimator", "collision")
        q = t["quantity"]
        shape = t.get("shape", [1, 1, 1])
        noise_like = int(est == "source" and q in ("real-source", "imag-source"))
        teb = t["energy-bounds"]
        out.append(f"tally {str(t['name']).replace(' ', '_')} {EST[est]} {QTY[q]} {noise_like} "
                   f"{int(shape[0])} {int(shape[1])} {int(shape[2])} {_fl(t['low'])} {_fl(t['hi'])} {len(teb)} {_fl(teb)}")

    c = deck.get("cancelator")
    if c and c.get("type") == "approximate":
        ceb = c.get("energy-bounds", [])
        sh = c["shape"]
        out.append(f"cancelator 1 {int(sh[0])} {int(sh[1])} {int(sh[2])} {_fl(c['low'])} {_fl(c['hi'])} {len(ceb)} {_fl(ceb)}")
    elif c and c.get("type") == "basic-exact":  # src/basic_exact_mg_cancelator.cpp:610-705 (sobol defaults to true, n-samples to 10)
        sh = c["shape"]
        beta = {"zero": 0, "minimum": 1, "average-f": 2, "average-g": 3}[c["beta"]]
        out.append(f"cancelator 2 {int(sh[0])} {int(sh[1])} {int(sh[2])} {_fl(c['low'])} {_fl(c['hi'])} {beta} "
                   f"{int(bool(c.get('sobol', True)))} {int(c.get('n-samples', 10))}")
    elif c and c.get("type") == "exact":  # src/exact_mg_cancelator.cpp:594-686 (read by the compiled reference only: oracle/ref_probe.cpp)
        sh = c["shape"]
        gb = c.get("group-bins", [])
        bins = " ".join(f"{len(b)} " + " ".join(str(int(g)) for g in b) for b in gb)
        out.append(f"cancelator 3 {int(sh[0])} {int(sh[1])} {int(sh[2])} {_fl(c['low'])} {_fl(c['hi'])} {int(c.get('n-samples', 10))} {len(gb)} {bins}".rstrip())
    else:
        out.append("cancelator 0")
    e = deck.get("entropy")
    if e:
        sh = e["shape"]
        out.append(f"entropy 1 {_fl(e['low'])} {_fl(e['hi'])} {int(sh[0])} {int(sh[1])} {int(sh[2])}")
    else:
        out.append("entropy 0")
    ns = deck.get("noise-sources", []) or []
    out.append(f"nnoise {len(ns)}")
    mat_ids = [int(m["id"]) for m in deck["materials"]]
    for n in ns:
        if n["type"] == "square-oscillation":
            out.append(f"sqosc {_fl(n['low'])} {_fl(n['hi'])} {_f(n['angular-frequency'])} {_f(n['epsilon-total'])} "
                       f"{_f(n['epsilon-fission'])} {_f(n['epsilon-scatter'])}")
        elif n["type"] == "flat-vibration":
            basis = {"x": 0, "y": 1, "z": 2}[str(n["direction"]).lower()]
            out.append(f"flatvib {_fl(n['low'])} {_fl(n['hi'])} {_f(n['angular-frequency'])} {basis} "
                       f"{mat_ids.index(int(n['positive-material']))} {mat_ids.index(int(n['negative-material']))}")
        else:
            raise ValueError("oracle: square-oscillation and flat-vibration noise sources only")
    return "\n".join(out) + "\n"


def write_deck_text(yaml_path: str, out_path: str, overrides: dict | None = None) -> None:
    deck = apply_overrides(load_yaml(yaml_path), overrides)
    with open(out_path, "w") as f:
        f.write(deck_to_text(deck))
