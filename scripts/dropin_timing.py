"""Times the reference's OWN PowerIterator::run() (oracle/_ref) on the bench deck, once with its own DeltaTracker on all host
threads and once with GPUTransporter (integration/gpu_transporter.hpp) forwarding transport() to the B200 backend.  The
time is the reference's own simulation_timer (generation loop).  Usage: python scripts/dropin_timing.py [N] [ngen] [nignored]"""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(mode, n, ngen, nign, yaml_path):
    import numpy as np
    from oracle import ref_pins, deck as _deck
    from abeille_b200 import backend
    L = ref_pins.ref_lib()
    L.ref_last_simulation_seconds.restype = C.c_double
    deck = _deck.load_yaml(yaml_path)
    a = [np.zeros(ngen) for _ in range(5)]
    summ = np.zeros(6)
    L.ref_set_threads(C.c_int(os.cpu_count() or 1))
    text = _deck.deck_to_text(deck).encode()
    P = [x.ctypes.data_as(C.POINTER(C.c_double)) for x in a] + [summ.ctypes.data_as(C.POINTER(C.c_double))]
    if mode == "gpu":
        rc = L.ref_power_iteration_gpu(text, backend.lib_paths()[1].encode(), yaml_path.encode(), 0, ngen, nign, *P)
    else:
        rc = L.ref_power_iteration(text, ngen, nign, *P)
    assert rc == 0
    secs = float(L.ref_last_simulation_seconds())
    return {"mode": mode, "particles_per_generation": n, "generations": ngen, "seconds": secs,
            "particles_per_s": n * ngen / secs, "k_col": a[0].tolist()}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in ("cpu", "gpu"):
        mode, n, ngen, nign, path = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
        saved = os.dup(1); os.dup2(2, 1)          # the reference's Output writes to stdout
        r = one(mode, n, ngen, nign, path)
        C.CDLL(None).fflush(None); os.dup2(saved, 1)
        print(json.dumps(r), flush=True)
        sys.exit(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    ngen = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    nign = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    import yaml
    deck = yaml.safe_load(open(os.path.join(ROOT, "tests", "decks", "c5g7_delta_collision_fullmesh.yaml")))
    deck["settings"].update(nparticles=n, ngenerations=ngen, nignored=nign)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "deck.yaml")
        yaml.safe_dump(deck, open(path, "w"), default_flow_style=None, sort_keys=False, width=200)
        for mode, nn in (("cpu", min(n, 200_000)), ("gpu", n)):   # one simulation per process (process globals)
            d2 = dict(deck); d2["settings"] = dict(deck["settings"], nparticles=nn)
            yaml.safe_dump(d2, open(path, "w"), default_flow_style=None, sort_keys=False, width=200)
            out = subprocess.run([sys.executable, os.path.abspath(__file__), mode, str(nn), str(ngen), str(nign), path],
                                 check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
            print(out, flush=True)
