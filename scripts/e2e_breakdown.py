"""Where the end-to-end (host-buffer) generation spends its wall time (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tempfile, torch, yaml
from abeille_b200.distributed import HostBufferLoop
import abeille_b200.backend as be

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
with open("tests/decks/c5g7_delta_collision_fullmesh.yaml") as f:
    d = yaml.safe_load(f)
d["settings"]["nparticles"] = n
with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
    yaml.safe_dump(d, f, default_flow_style=None, sort_keys=False, width=200)
loop = HostBufferLoop(f.name, 0, n)
loop.initialize()
orig = loop.gpu.transport
acc = {"transport": 0.0, "calls": 0}
def timed(*a, **k):
    t0 = time.perf_counter(); r = orig(*a, **k); acc["transport"] += time.perf_counter() - t0; acc["calls"] += 1; return r
loop.gpu.transport = timed
for _ in range(3): loop.generation(True)
acc.update(transport=0.0, calls=0)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4): loop.generation(True)
torch.cuda.synchronize(); tot = time.perf_counter() - t0
k = loop.gpu.last_transport_kernel()
print(f"per generation: total {1e3*tot/4:.1f} ms, abl_transport call {1e3*acc['transport']/4:.1f} ms (history kernel {k['ms']:.1f} ms), "
      f"host-side steps + tally record {1e3*(tot-acc['transport'])/4:.1f} ms, threads {torch.get_num_threads()}")
