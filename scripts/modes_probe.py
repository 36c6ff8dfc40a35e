"""Resident generation rate of the one-GPU C++ PowerIterator (Backend.run_power_iteration) for the simulation modes / cancelators that
do not go through the staged history kernel, beside the ones that do: particles per second over whole generations."""
import json
import os
import sys
import tempfile

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import abeille_b200 as ab  # noqa: E402

N, NGEN, NIGN = 1_000_000, 6, 2
out = {}
for fname in ("c5g7_delta_collision.yaml", "c5g7_delta_branchless.yaml", "c5g7_carter_cancel.yaml", "c5g7_carter_exact_avgg.yaml",
              "c5g7_carter_exact_full.yaml", "PUa-1-0-SL_subcritical_mfs.yaml"):
    with open(os.path.join(ROOT, "tests", "decks", fname)) as f:
        deck = yaml.safe_load(f)
    if "fixed-source" in deck["settings"]["simulation"]:
        continue
    deck["settings"].update({"nparticles": N, "ngenerations": NGEN, "nignored": NIGN})
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        yaml.safe_dump(deck, f, default_flow_style=None, sort_keys=False, width=200)
        path = f.name
    gpu = ab.Backend(path, 0)
    gpu.run_power_iteration(2, 1, resident=True)  # warm-up (allocations, kernel images)
    gpu.close()
    gpu = ab.Backend(path, 0)
    r = gpu.run_power_iteration(NGEN, NIGN, resident=True)
    total = float(sum(int(v) for v in r["nbank"]))
    out[fname] = {"particles_per_s": total / r["seconds"], "seconds": r["seconds"], "kcol_avg": r["kcol_avg"], "nbank_last": int(r["nbank"][-1])}
    gpu.close()
    os.unlink(path)
print(json.dumps(out))
