"""Quick device-side throughput probe (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, yaml, tempfile
import abeille_b200 as ab

def run(deck, n, gens=4, converged=True):
    with open(os.path.join("tests/decks", deck)) as f:
        d = yaml.safe_load(f)
    d["settings"].update({"nparticles": n})
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        yaml.safe_dump(d, f, default_flow_style=None, sort_keys=False, width=200)
        path = f.name
    gpu = ab.Backend(path, 0)
    cap = int(4 * n) + 4096
    a, b = gpu.new_device_bank(cap), gpu.new_device_bank(cap)
    gpu.sample_source_device(a, n, 0)
    cur_n, use_state, k, first = n, True, 1.0, n
    for g in range(gens):
        torch.cuda.synchronize(); t0 = time.time()
        m, scores, cn = gpu.transport_device(a, cur_n, b, k_col=k, converged=converged, use_rng_state=use_state)
        torch.cuda.synchronize(); t1 = time.time()
        kk = gpu.last_transport_kernel()
        k = scores[0] / n
        ws = gpu.weight_stats_device(b, m)
        gpu.scale_weights_device(b, m, n / (ws[2] - ws[3]))
        gpu.to_particles_device(b, m, first)
        if converged:
            gpu.tallies_record(1.0); gpu.tallies_clear()
        torch.cuda.synchronize(); t2 = time.time()
        print(f"{deck} n={cur_n} gen{g}: kernel {kk['ms']:.1f} ms grid {kk['grid']} transport call {1e3*(t1-t0):.1f} ms, rest {1e3*(t2-t1):.1f} ms, "
              f"k={k:.5f} m={m} flights/hist={cn['flights']/cur_n:.1f} coll/hist={cn['real_collisions']/cur_n:.1f} "
              f"-> {cur_n/(kk['ms']*1e-3)/1e6:.2f} M particles/s (kernel), {cn['real_collisions']/(kk['ms']*1e-3)/1e6:.1f} M coll/s", flush=True)
        first += m
        a, b = b, a
        cur_n, use_state = m, False
    gpu.close()

if __name__ == "__main__":
    if "--implicit" in sys.argv:  # implicit-leakage delta tracking (per-lane kernel) next to plain delta tracking (staged kernel)
        run("c5g7_implicit_collision.yaml", 1_000_000, 3, True)
        run("c5g7_delta_collision.yaml", 1_000_000, 3, True)
        run("PUa-1-0-SL_implicit.yaml", 1_000_000, 3, True)
        sys.exit(0)
    if "--others" not in sys.argv:
        run("c5g7_delta_collision.yaml", 1_000_000, 4, True)
        run("c5g7_delta_collision.yaml", 10_000_000, 3, False)
        run("c5g7_delta_collision_fullmesh.yaml", 10_000_000, 3, True)
    run("PUa-1-0-IN.yaml", 400_000, 3, True)
    run("c5g7_surface_tracklength.yaml", 1_000_000, 2, True)
    run("c5g7_carter_cancel.yaml", 1_000_000, 3, True)
