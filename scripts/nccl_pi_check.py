"""The C++ multi-GPU generation loop (abeille_b200/lib/abl_pi_nccl: one process per GPU, NCCL called directly) against itself on one
GPU and against the one-GPU C++ PowerIterator of the host library: bank sizes identical, k and entropy series equal to rounding.

    python scripts/nccl_pi_check.py [--world 2]        (needs `world` GPUs; every process is run under `timeout`)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BIN = os.path.join(ROOT, "abeille_b200", "lib", "abl_pi_nccl")


def run(deck_path, world, ngen, nign, limit=240):
    with tempfile.TemporaryDirectory() as td:
        idf = os.path.join(td, "nccl_id")
        procs = [subprocess.Popen(["timeout", str(limit), BIN, deck_path, str(r), str(world), idf, str(ngen), str(nign), str(r)],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
        outs = [p.communicate() for p in procs]
        for r, (p, (o, e)) in enumerate(zip(procs, outs)):
            if p.returncode != 0:
                raise RuntimeError(f"rank {r} of {world} exited {p.returncode}: {e[-2000:]}")
        return json.loads(outs[0][0].strip().splitlines()[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--nparticles", type=int, default=200000)
    args = ap.parse_args()
    import abeille_b200 as ab
    report = {}
    for fname, ngen, nign in (("c5g7_delta_collision.yaml", 6, 2), ("c5g7_carter_cancel.yaml", 6, 2)):
        with open(os.path.join(ROOT, "tests", "decks", fname)) as f:
            deck = yaml.safe_load(f)
        deck["settings"].update({"nparticles": args.nparticles, "ngenerations": ngen, "nignored": nign})
        with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
            yaml.safe_dump(deck, f, default_flow_style=None, sort_keys=False, width=200)
            path = f.name
        one = run(path, 1, ngen, nign)
        many = run(path, args.world, ngen, nign)
        gpu = ab.Backend(path, 0)
        host = gpu.run_power_iteration(ngen, nign, resident=True)
        assert one["nbank"] == many["nbank"] == [int(v) for v in host["nbank"]], (one["nbank"], many["nbank"], host["nbank"])
        assert np.allclose(one["kcol"], many["kcol"], rtol=1e-12) and np.allclose(one["kcol"], host["kcol"], rtol=1e-10)
        assert np.allclose(one["entropy"], many["entropy"], rtol=1e-12) and np.allclose(one["entropy"], host["entropy"], rtol=1e-10)
        assert abs(one["kcol_avg"] - many["kcol_avg"]) < 1e-12
        report[fname] = {"world": args.world, "nbank": many["nbank"], "kcol_1": one["kcol"], f"kcol_{args.world}": many["kcol"],
                         "max_rel_diff_k": float(np.max(np.abs(np.array(one["kcol"]) / np.array(many["kcol"]) - 1.0))),
                         "seconds_1": one["seconds"], f"seconds_{args.world}": many["seconds"]}
        os.unlink(path)
    print(json.dumps(report))


if __name__ == "__main__":
    main()
