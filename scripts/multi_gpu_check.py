"""Multi-GPU consistency check (run under torchrun on a B200 box): the same deck on N ranks and on one rank must give the
same per-generation k and bank sizes (integer outcomes do not depend on the GPU count; sums differ in rounding order).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/multi_gpu_check.py [deck] [n_total] [gens]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist, yaml
from abeille_b200.distributed import DistributedPowerIterator

deck = sys.argv[1] if len(sys.argv) > 1 else "c5g7_carter_cancel.yaml"
n_total = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
gens = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
with open(os.path.join("tests/decks", deck)) as f:
    d = yaml.safe_load(f)
d["settings"]["nparticles"] = n_total
path = os.path.join(tempfile.mkdtemp(), f"deck_{rank}.yaml")
with open(path, "w") as f:
    yaml.safe_dump(d, f, default_flow_style=None, sort_keys=False, width=200)

def run(group_world):
    sim = DistributedPowerIterator(path, local, n_total // group_world) if group_world > 1 else None
    return sim

sim = DistributedPowerIterator(path, local, n_total // world)
sim.initialize()
for g in range(gens):
    sim.generation(converged=g >= 1)
k_multi, nb_multi = np.array(sim.kcol_series), np.array(sim.nbank_series)
tallies = [t.clone() for t in sim.tally_tensors()] if sim.gpu.ntallies() else []
del sim
if world > 1:
    dist.barrier()
if rank == 0:
    # single-rank reference in the same process (no process group involvement: use_dist is evaluated per instance)
    import abeille_b200.distributed as D
    saved = D.dist.is_initialized
    D.dist.is_initialized = lambda: False
    ref = DistributedPowerIterator(path, local, n_total)
    D.dist.is_initialized = saved
    ref.initialize()
    for g in range(gens):
        ref.generation(converged=g >= 1)
    ok_n = [int(a) for a in nb_multi] == [int(a) for a in ref.nbank_series]
    ok_k = bool(np.allclose(k_multi, ref.kcol_series, rtol=1e-10))
    print(f"{deck} n={n_total} ranks={world}: bank sizes equal: {ok_n}; k equal to 1e-10: {ok_k}; k = {k_multi}")
    print("MULTI_GPU_CHECK", "OK" if ok_n and ok_k else "FAILED")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
