"""Wall time of a short noise run on the device (development aid): python scripts/noise_probe.py [deck] [nparticles] [batches]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, yaml
from abeille_b200.noise import NoiseSimulation
deck = sys.argv[1] if len(sys.argv) > 1 else "noise_oscillation.yaml"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
batches = int(sys.argv[3]) if len(sys.argv) > 3 else 2
with open(os.path.join("tests/decks", deck)) as f:
    d = yaml.safe_load(f)
d["settings"].update({"nparticles": n, "ngenerations": batches, "nignored": 3, "nskip": 3})
with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
    yaml.safe_dump(d, f, default_flow_style=None, sort_keys=False, width=200)
sim = NoiseSimulation(f.name, 0)
torch.cuda.synchronize(); t0 = time.perf_counter()
sim.initialize()
for _ in range(sim.nignored): sim.power_iteration(False)
sim.converged = True
torch.cuda.synchronize(); t1 = time.perf_counter()
tp = tn = 0.0
for b in range(batches):
    torch.cuda.synchronize(); a = time.perf_counter()
    for _ in range(sim.nskip - 1): sim.power_iteration(False)
    nn = sim.power_iteration(True)
    torch.cuda.synchronize(); c = time.perf_counter()
    sim.noise_simulation(nn)
    torch.cuda.synchronize(); e = time.perf_counter()
    tp += c - a; tn += e - c
print(f"{deck} n={n}: {sim.nignored} inactive generations {t1-t0:.3f} s; per batch: {sim.nskip} power-iteration generations {tp/batches:.3f} s, "
      f"noise simulation {tn/batches:.3f} s ({sim.noise_generations} generations, {sim.noise_particles} noise histories)")
