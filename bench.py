#!/usr/bin/env python
"""bench.py -- throughput of the transport hot path on the BASELINE.json workload.

Workload (configs[1]): C5G7, 7 groups, delta tracking, collision-estimator flux mesh tally on the deck's own
1224 x 1224 x 10 x 7 mesh (tests/decks/c5g7_delta_collision_fullmesh.yaml), 10^7 particles per generation per
GPU.  A "step" is one ACTIVE power-iteration generation: Transporter::transport over the whole bank, then the
per-generation bookkeeping the reference does between transport calls (k_col, weight normalisation, tally
record/clear, history-id / RNG re-seeding).  Histories shard across GPUs by global history id (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

* value        active particles/s, bank resident in HBM, CUDA events around the K steps, max over ranks.
* e2e          the same metric through the host-buffer C ABI (abl_transport): every step copies the bank
               host->device from pinned memory and the fission bank + scores device->host.
* roofline     the history kernel: algorithmic bytes / measured kernel time.
* cpu_baseline the reference's own DeltaTracker::transport (oracle/_ref, compiled from the reference's sources; kind
               "reference") on this box's cores; the CPU oracle port (oracle/) when that library was not built.
* --impl reference: that CPU path as its own arm (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DECK = os.path.join(ROOT, "tests", "decks", "c5g7_delta_collision_fullmesh.yaml")
METRIC = "active particles/sec (+ collisions/sec), C5G7 delta-tracking"
WORKLOAD = "c5g7 7-group delta-tracking k-eigenvalue, collision-estimator flux mesh tally 1224x1224x10x7"

# The other configurations of BASELINE.json (`--config N`; the default, 2, is the contract's bench line).  Each prints one
# JSON line of the same shape for its own deck (resident leg only) and runs its parity tests (tests/, the only place besides
# the CPU legs that uses the oracle) as a subprocess in the same run.
CONFIGS = {
    1: {"deck": "PUa-1-0-IN.yaml", "particles": 1_000_000, "kernel": "history_kernel<surface>",
        "metric": "active particles/sec (+ collisions/sec), Sood PUa-1-0-IN surface-tracking",
        "workload": "Sood PUa-1-0-IN one-group infinite medium, surface tracking (k_inf = 2.612903)",
        "tests": ["tests/test_gpu_parity.py", "-k", "PUa-1-0-IN or sood"]},
    3: {"deck": "ref_sqr_c5g7_surface_tl.yaml", "particles": 2_000_000, "kernel": "history_kernel<surface, track-length>",
        "metric": "active particles/sec (+ collisions/sec), C5G7 surface-tracking + track-length tally",
        "workload": "ref_sqr c5g7 7-group surface-tracking k-eigenvalue, track-length flux mesh tally 102x102x5x7",
        "tests": ["tests/test_gpu_parity.py", "tests/test_gpu_reference_golden.py", "-k", "surface or ref_sqr"]},
    4: {"deck": "c5g7_carter_cancel.yaml", "particles": 12_500_000, "kernel": "history_kernel<carter>",
        "metric": "active particles/sec (+ collisions/sec), C5G7 carter-tracking + approximate mesh cancellation",
        "workload": "c5g7 7-group carter tracking with an under-estimated sampling cross section, approximate mesh weight "
                    "cancellation (170x170x765 bins, dense-bin all-reduce across ranks), 1.25e7 particles per generation per GPU",
        "tests": ["tests/test_gpu_parity.py", "-k", "carter or cancel"]},
    5: {"deck": "noise_oscillation.yaml", "particles": 100_000, "kernel": "transport_kernel<surface, noise>",
        "metric": "noise batch: particles/sec (power-iteration + noise particles), frequency-domain neutron noise",
        "workload": "noise_oscillation.yaml: nskip power-iteration generations (the last one samples the noise source), then the "
                    "inner noise generations (complex weights, regional cancellation) until the noise bank is empty",
        "tests": ["tests/test_gpu_noise.py", "-k", "noise"]},
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _write_deck(nparticles_total, path, deck_path=None):
    import yaml
    with open(deck_path or DECK) as f:
        deck = yaml.safe_load(f)
    deck["settings"]["nparticles"] = int(nparticles_total)
    with open(path, "w") as f:
        yaml.safe_dump(deck, f, default_flow_style=None, sort_keys=False, width=200)
    return path


def bind_to_gpu_numa_node(gpu_index):
    """Run this process on the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated (first touch
    puts the pages on that node): with 8 ranks on a two-socket host the bank copies (1.7 GB per rank and generation) otherwise
    cross the socket interconnect for half the GPUs.  Returns a short description for the bench line, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml pads the domain to 8 digits, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"gpu": gpu_index, "numa_node": node, "cpus": len(allowed)}
    except Exception:  # noqa: BLE001 -- placement is an optimisation, never a reason to fail the bench
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, smax, reasons, pw = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
def cpu_leg(nparticles, inactive, active, threads=None):
    """The CPU oracle (restatement of the reference's OpenMP trackers) on a bounded sample of the workload."""
    from oracle import api
    api.build()
    api.set_math("libm")  # glibc log/sin/cos, as the reference
    nthreads = threads or (os.cpu_count() or 1)
    api.set_threads(nthreads)
    with tempfile.TemporaryDirectory() as td:
        orc = api.Oracle(_write_deck(nparticles, os.path.join(td, "deck.yaml")))
    orc.pi_init(inactive)
    if inactive:
        orc.pi_run(inactive)
    r = orc.pi_run(active)
    orc.close()
    return {"particles_per_s": r["particles"] / r["seconds"], "collisions_per_s": r["real_collisions"] / r["seconds"],
            "seconds": r["seconds"], "particles": r["particles"], "cores": nthreads,
            "sample": f"{active} active generations of {nparticles} particles after {inactive} inactive, same deck and mesh, "
                      f"OpenMP schedule(dynamic) over histories, {nthreads} threads, g++ -O2"}


class _StdoutToStderr:
    """The reference's Output singleton writes its progress lines to stdout; bench.py's stdout is the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        import ctypes
        ctypes.CDLL(None).fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)


def ref_leg(nparticles, warm, active, threads=None):
    with _StdoutToStderr():
        try:
            return _ref_leg(nparticles, warm, active, threads)
        except Exception as e:  # library missing its dependencies, deck rejected ...: the caller falls back to the port
            print(f"bench.py: oracle/_ref not usable ({type(e).__name__}: {e}); falling back to the oracle port", file=sys.stderr)
            return None


def _ref_leg(nparticles, warm, active, threads=None):
    """The REFERENCE'S OWN DeltaTracker::transport (oracle/_ref/libabeille_ref.so: the reference's multigroup sources compiled
    in place by oracle/Makefile, see oracle/ref_probe.cpp) on a bounded sample of the workload, all host threads, its own
    CollisionMeshTally on the deck's full mesh scoring.  A step is one active generation: the timed region is the
    transport() call; between calls the bank of the next generation is made from the fission bank as PowerIterator::run
    does (weights normalised to the particle count, new history ids).  Returns None when the library is not there."""
    import ctypes as C
    import numpy as np
    from oracle import ref_pins, deck as _deck
    if not os.path.exists(ref_pins.REF_LIB):
        return None
    L = C.CDLL(ref_pins.REF_LIB)
    L.ref_last_transport_seconds.restype = C.c_double
    nthreads = threads or (os.cpu_count() or 1)
    deck = _deck.apply_overrides(_deck.load_yaml(DECK), {"settings": {"nparticles": int(nparticles)}})
    if L.ref_problem_load(_deck.deck_to_text(deck).encode()) != 0:
        raise RuntimeError("oracle/_ref: ref_problem_load failed")
    L.ref_set_threads(C.c_int(nthreads))
    PD, PU = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    rng = np.random.default_rng(1)
    sp = deck["sources"][0]["spatial"]
    n = int(nparticles)
    r = np.ascontiguousarray(rng.uniform(sp["low"], sp["hi"], (n, 3)))
    u = rng.normal(size=(n, 3))
    u = np.ascontiguousarray(u / np.linalg.norm(u, axis=1)[:, None])
    eb = deck["settings"]["energy-bounds"]
    E = np.full(n, 0.5 * (eb[0] + eb[1]))
    w = np.ones(n)
    k_col, next_id, seconds, particles = 1.0, 0, 0.0, 0
    cap = 4 * n
    f9, ids, k6 = np.zeros((cap, 9)), np.zeros((cap, 3), dtype=np.uint64), np.zeros(6)
    for g in range(warm + active):
        m = len(w)
        hid = np.arange(next_id, next_id + m, dtype=np.uint64)
        next_id += m
        nout = C.c_uint64(0)
        rc = L.ref_transport(C.c_uint64(m), r.ctypes.data_as(PD), u.ctypes.data_as(PD), E.ctypes.data_as(PD), w.ctypes.data_as(PD),
                             hid.ctypes.data_as(PU), hid.ctypes.data_as(PU), C.c_double(k_col), C.c_int(int(g >= warm)),
                             C.c_uint64(cap), f9.ctypes.data_as(PD), ids.ctypes.data_as(PU), C.byref(nout), k6.ctypes.data_as(PD))
        if rc != 0 or nout.value > cap or nout.value == 0:
            raise RuntimeError(f"oracle/_ref: ref_transport failed (rc {rc}, {nout.value} sites)")
        if g >= warm:
            seconds += float(L.ref_last_transport_seconds())
            particles += m
        k_col = float(k6[0])
        k = int(nout.value)
        r, u = np.ascontiguousarray(f9[:k, 0:3]), np.ascontiguousarray(f9[:k, 3:6])
        E = np.ascontiguousarray(f9[:k, 6])
        w = np.ascontiguousarray(f9[:k, 7] * (n / f9[:k, 7].sum()))  # PowerIterator::normalize_weights
    return {"particles_per_s": particles / seconds, "collisions_per_s": None, "seconds": seconds, "particles": particles,
            "cores": nthreads, "k_col": k_col,
            "sample": f"{active} active generations of ~{n} particles after {warm} inactive, same deck and 1224x1224x10x7 mesh, "
                      f"the reference's own DeltaTracker::transport + CollisionMeshTally compiled from /root/reference "
                      f"(oracle/_ref), OpenMP schedule(dynamic), {nthreads} threads, g++ -O3; transport() calls timed"}


def run_reference(args):
    """--impl reference: the reference's own CPU transport (oracle/_ref; the oracle port if that library is absent) on all
    host threads.  A step is one active generation of a bounded particle count."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_particles
    t0 = time.time()
    r, kind = ref_leg(n, max(args.warmup, 1), args.steps), "reference"
    if r is None:  # oracle/_ref was not built (no /root/reference at build time): the oracle port
        r, kind = cpu_leg(n, max(args.warmup, 1), args.steps), "port"
    out = {"impl": "reference", "metric": METRIC, "value": r["particles_per_s"], "unit": "particles/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "collisions_per_s": r["collisions_per_s"],
           "config": {"workload": WORKLOAD, "particles_per_step": n, "host_threads": r["cores"]},
           "cpu_baseline": {"value": r["particles_per_s"], "unit": "particles/s", "cores": r["cores"], "kind": kind,
                            "sample": r["sample"]},
           "e2e": {"value": r["particles_per_s"], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.time() - t0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 backend has no CPU fallback)")
    numa = bind_to_gpu_numa_node(local) if not args.no_numa else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # torchrun exports OMP_NUM_THREADS=1; the e2e leg's caller-side steps (PowerIterator::normalize_weights on the host
    # bank) are CPU work that the reference spreads over the host's cores: give every rank its share
    torch.set_num_threads(max(1, min(len(os.sched_getaffinity(0)), (os.cpu_count() or 1) // max(world, 1))))
    # NCCL prints its version banner on stdout when the first communicator is made (N > 1): the file descriptor is pointed at
    # stderr for the whole run and restored for the one JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as ge
    if rank == 0:
        ge.build_if_missing()
    if world > 1:
        dist.barrier()
    import abeille_b200 as ab
    from abeille_b200.distributed import DistributedPowerIterator, HostBufferLoop

    cfg = CONFIGS.get(args.config)
    if args.config == 5:
        return run_noise_config(args, rank, world, local, dev, saved_stdout)
    metric, workload, kernel_name = (cfg["metric"], cfg["workload"], cfg["kernel"]) if cfg else (METRIC, WORKLOAD, "history_kernel<delta>")
    if cfg and args.particles == 10_000_000:
        args.particles = cfg["particles"]
    if cfg:  # the other configurations: the resident leg and their parity tests
        args.no_e2e = args.no_cpu = args.no_ncu = args.no_ranks_check = True
    n_local = args.particles
    n_total = n_local * world
    td = tempfile.mkdtemp()
    deck = _write_deck(n_total, os.path.join(td, f"bench_{rank}.yaml"), os.path.join(ROOT, "tests", "decks", cfg["deck"]) if cfg else None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident leg: `value` -------------------------------------------------------------------------------
    sim = DistributedPowerIterator(deck, local, n_local)
    sim.initialize()
    for _ in range(args.warmup):
        sim.generation(converged=True)
    launches0 = sim.gpu.device_info()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    particles = collisions = sites = coll_scores = flights = 0.0
    local_particles = local_coll = local_sites = local_scores = 0.0
    kernel_ms = []
    for _ in range(args.steps):
        g = sim.generation(converged=True)
        particles += g["n_in_total"]; collisions += g["real_collisions"]; sites += g["m_total"]; coll_scores += g["coll_scores"]
        local_particles += g["n_in"]; local_coll += g["local_real_collisions"]; local_sites += g["m_pre"]
        local_scores += g["local_coll_scores"] + g["local_tl_bins"]
        flights += g["local_flights"]
        kernel_ms.append(sim.gpu.last_transport_kernel()["ms"])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.gpu.device_info()["kernel_launches"] - launches0
    kinfo = sim.gpu.last_transport_kernel()
    k_col = sim.k_col
    # roofline of the history kernel on this rank: SURVEY.md 8(d) algorithmic bytes per launch
    alg_bytes = (96.0 * local_particles + 72.0 * local_sites + 16.0 * local_scores) / args.steps
    k_ms = float(np.mean(kernel_ms))
    peak, peak_src = _peaks()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    tally_bins = sum(int(np.prod(sim.gpu.tally_shape(t))) for t in range(sim.gpu.ntallies()))
    del sim
    torch.cuda.empty_cache()

    # ---- e2e leg: host buffers through abl_transport -----------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        loop = HostBufferLoop(deck, local, n_local)
        loop.initialize()
        for _ in range(min(args.warmup, 3)):
            loop.generation(converged=True)
        barrier()
        t0 = time.perf_counter()
        ep = 0.0
        for _ in range(args.e2e_steps):
            ep += loop.generation(converged=True)["n_in_total"]
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": ep / dt, "unit": "particles/s", "h2d_bytes_per_step": loop.h2d_bytes / loop.gens,
               "d2h_bytes_per_step": loop.d2h_bytes / loop.gens, "steps": args.e2e_steps,
               "api": "abl_transport_begin + abl_transport_finish (host buffers, pinned): the bank crosses PCIe both ways every generation; normalize_weights and the fresh history ids of PowerIterator::run are applied on the device between the two calls"}
        del loop

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r, kind = ref_leg(args.cpu_particles, 2, args.cpu_steps), "reference"
        if r is None:
            r, kind = cpu_leg(args.cpu_particles, 2, args.cpu_steps), "port"
        cpu = {"value": r["particles_per_s"], "unit": "particles/s", "cores": r["cores"], "kind": kind, "sample": r["sample"],
               "collisions_per_s": r["collisions_per_s"]}
        if kind == "reference":  # the oracle port beside it, for the record
            rp = cpu_leg(args.cpu_particles, 2, args.cpu_steps)
            cpu["port_value"] = rp["particles_per_s"]

    # ---- one kernel launch under ncu (N = 1): DRAM traffic, lanes, issue utilisation -----------------------------------
    prof = None
    if rank == 0 and world == 1 and not args.no_ncu:
        torch.cuda.synchronize()
        prof = ncu_kernel_metrics(args)
    # ---- N > 1: sharded run == 1-rank replay -------------------------------------------------------------------------------
    rcheck = None
    if world > 1 and not args.no_ranks_check:
        rcheck = ranks_check(world, rank, local, dev)

    if rank == 0:
        out = {"metric": metric, "value": particles / (ms * 1e-3), "unit": "particles/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "collisions_per_s": collisions / (ms * 1e-3),
               "config": {"workload": workload, "particles_per_generation_per_gpu": n_local, "particles_per_generation": n_total,
                          "tally_bins": tally_bins, "source": "deck box source, then the fission source after the warm-up generations",
                          "l2": "inputs larger than L2 (bank 96 B x %.3g = %.2f GB)" % (n_local, 96e-9 * n_local), "k_col": k_col,
                          "collisions_per_particle": collisions / max(particles, 1)},
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": prof["dram_bytes"] if prof else None, "kernel": kernel_name, "kernel_ms": k_ms,
                            "profile": prof, "flights_per_launch": flights / args.steps,
                            "warp_instructions_per_flight": (prof["warp_instructions"] / (flights / args.steps))
                            if prof and prof.get("warp_instructions") and flights else None,
                            "grid": [kinfo["grid"], kinfo["block"]], "algorithmic_bytes_per_launch": alg_bytes,
                            "peak_source": peak_src, "kernel_share_of_step": k_ms * args.steps / ms},
               "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "host_placement": numa}
        if rcheck is not None:
            out["ranks_check"] = rcheck
        if cfg and not args.no_parity:  # the configuration's parity tests, in the same run (rank 0's GPU)
            t0 = time.time()
            r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x"] + cfg["tests"], cwd=ROOT, capture_output=True,
                               text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(local)))
            tail = [l for l in r.stdout.strip().splitlines() if l.strip()][-1:] or [""]
            out["parity"] = {"passed": r.returncode == 0, "summary": tail[0], "seconds": time.time() - t0,
                             "what": "pytest -m gpu " + " ".join(cfg["tests"]) + " (bit-exact against the oracle / the reference's golden vectors)"}
    import ctypes
    sys.stdout.flush()
    ctypes.CDLL(None).fflush(None)
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def run_noise_config(args, rank, world, local, dev, saved_stdout):
    """--config 5: noise batches of the shipped noise_oscillation deck through DistributedNoiseSimulation (one process per GPU,
    every bank sharded in rank order).  A step is one noise batch: nskip power-iteration generations and the inner noise
    generations that follow; value = (power-iteration particles + noise particles) of the timed batches / device time."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import yaml
    from abeille_b200.noise import DistributedNoiseSimulation
    cfg = CONFIGS[5]
    n_total = (cfg["particles"] if args.particles == 10_000_000 else args.particles) * world
    td = tempfile.mkdtemp()
    with open(os.path.join(ROOT, "tests", "decks", cfg["deck"])) as f:
        deck = yaml.safe_load(f)
    deck["settings"].update({"nparticles": int(n_total), "nignored": 2, "nskip": 2, "ngenerations": args.steps})
    path = os.path.join(td, f"noise_{rank}.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(deck, f, default_flow_style=None, sort_keys=False, width=200)
    sim = DistributedNoiseSimulation(path, local)
    sim.initialize()
    sim.converged = False
    for _ in range(sim.nignored):
        sim.power_iteration(False)
    sim.converged = True

    def batch():
        pi = 0
        for _ in range(sim.nskip - 1):
            sim.power_iteration(False)
            pi += sim.nparticles
        n_noise = sim.power_iteration(True)
        pi += sim.nparticles
        sim.noise_simulation(n_noise)
        return pi + sim.noise_particles[-1]

    for _ in range(min(args.warmup, 1)):
        batch()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sim.gpu.device_info()["kernel_launches"]
    e0.record()
    particles = 0
    first = len(sim.noise_generations)
    for _ in range(args.steps):
        particles += batch()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = None
    if rank == 0:
        out = {"metric": cfg["metric"], "value": particles / (ms * 1e-3), "unit": "particles/s", "n_gpus": world, "steps": args.steps,
               "warmup": min(args.warmup, 1), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": cfg["workload"], "particles_per_generation": int(n_total), "nskip": sim.nskip,
                          "noise_generations_per_batch": sim.noise_generations[first:],
                          "noise_particles_per_batch": sim.noise_particles[first:], "k_col": float(sim.k_col)},
               "gpu_launches": int(sim.gpu.device_info()["kernel_launches"] - launches0)}
        if not args.no_parity:
            t0 = time.time()
            r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x"] + cfg["tests"], cwd=ROOT, capture_output=True,
                               text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=str(local)))
            tail = [l for l in r.stdout.strip().splitlines() if l.strip()][-1:] or [""]
            out["parity"] = {"passed": r.returncode == 0, "summary": tail[0], "seconds": time.time() - t0,
                             "what": "pytest -m gpu " + " ".join(cfg["tests"])}
    import ctypes
    sys.stdout.flush()
    ctypes.CDLL(None).fflush(None)
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def ncu_child(args):
    """The resident leg alone, a few generations, no output: the process `ncu_kernel_metrics` profiles."""
    import torch
    import __graft_entry__ as ge
    ge.build_if_missing()
    from abeille_b200.distributed import DistributedPowerIterator
    td = tempfile.mkdtemp()
    sim = DistributedPowerIterator(_write_deck(args.particles, os.path.join(td, "ncu.yaml")), 0, args.particles)
    sim.initialize()
    for _ in range(4):
        sim.generation(converged=True)
    torch.cuda.synchronize()


def ncu_kernel_metrics(args):
    """DRAM bytes, lanes per instruction, issue utilisation and instruction count of ONE history-kernel launch of this very
    workload, captured by running this script's resident leg under ncu (4th generation: the fission source, tallies on).
    Returns None when ncu is not there or the capture fails -- the bench line then carries no traffic figure."""
    import csv, io, shutil
    ncu = shutil.which("ncu") or ("/usr/local/cuda/bin/ncu" if os.path.exists("/usr/local/cuda/bin/ncu") else None)
    if ncu is None:
        return None
    metrics = ["dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
               "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
               "gpu__time_duration.sum"]
    cmd = [ncu, "--metrics", ",".join(metrics), "--clock-control", "none", "-k", "regex:history_kernel|event_kernel", "-s", "3", "-c", "1", "--csv",
           sys.executable, os.path.abspath(__file__), "--ncu-child", "--particles", str(args.particles)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        rows = [row for row in csv.reader(io.StringIO(r.stdout)) if len(row) > 5]
        hdr = next(row for row in rows if "Metric Name" in row)
        iname, ival, iunit = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        vals = {}
        for row in rows:
            if row is hdr or len(row) <= ival:
                continue
            try:
                v = float(row[ival].replace(",", ""))
            except ValueError:
                continue
            unit = row[iunit].lower()
            scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "usecond": 1e-3, "msecond": 1.0, "second": 1e3,
                     "nsecond": 1e-6, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
            vals[row[iname]] = v * scale
        if "dram__bytes_read.sum" not in vals:
            return None
        return {"dram_bytes": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
                "warp_instructions": vals.get("smsp__inst_executed.sum"),
                "lanes_per_instruction": (vals["smsp__thread_inst_executed.sum"] / vals["smsp__inst_executed.sum"])
                if vals.get("smsp__inst_executed.sum") else None,
                "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": vals.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "kernel_ms_under_ncu": vals.get("gpu__time_duration.sum"),
                "how": "ncu --metrics ... -k regex:history_kernel -s 3 -c 1 on `bench.py --ncu-child` (same deck and bank size, 4th "
                       "generation), run by this bench after the timed legs"}
    except Exception as e:  # noqa: BLE001 -- a failed capture must not cost the bench line
        print(f"bench.py: ncu capture failed ({type(e).__name__}: {e})", file=sys.stderr)
        return None


def ranks_check(world, rank, local, dev, n_total=1_600_000, gens=3):
    """N > 1: the sharded run against a 1-rank replay.  Histories shard by global history id and RNG streams are a function
    of that id, so bank sizes must be IDENTICAL for any GPU count and k_col equal up to summation order.
    Every rank runs `gens` generations of n_total particles sharded; rank 0 then replays the same generations alone."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from abeille_b200.distributed import DistributedPowerIterator
    td = tempfile.mkdtemp()
    n_total -= n_total % world
    deck = _write_deck(n_total, os.path.join(td, f"check_{rank}.yaml"))
    sim = DistributedPowerIterator(deck, local, n_total // world)
    sim.initialize()
    for g in range(gens):
        sim.generation(converged=g >= 1)
    torch.cuda.synchronize()
    sharded = (list(sim.nbank_series), list(sim.kcol_series), float(sim.counters[1]))
    del sim
    out = None
    if rank == 0:
        one = DistributedPowerIterator(deck, local, n_total, single=True)
        one.initialize()
        for g in range(gens):
            one.generation(converged=g >= 1)
        torch.cuda.synchronize()
        k1, kn = np.array(one.kcol_series), np.array(sharded[1])
        out = {"particles": n_total, "generations": gens, "ranks": world,
               "bank_sizes_equal": [int(v) for v in one.nbank_series] == [int(v) for v in sharded[0]],
               "collisions_equal": float(one.counters[1]) == sharded[2],
               "k_col_max_rel_diff": float(np.max(np.abs(k1 - kn) / np.abs(k1))),
               "what": "sharded run vs a 1-rank replay of the same generations on rank 0 (deck source, then fission source)"}
        del one
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=10_000_000, help="particles per generation per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-particles", type=int, default=100_000)
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--ref-particles", type=int, default=100_000, help="--impl reference: particles per step")
    ap.add_argument("--no-ncu", action="store_true", help="skip the ncu capture of one kernel launch (roofline.traffic = null)")
    ap.add_argument("--no-ranks-check", action="store_true", help="N > 1: skip the sharded-vs-1-rank consistency run")
    ap.add_argument("--ncu-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configuration: 2 = the bench line (default); 1, 3, 4 = the other k-eigenvalue configurations; 5 = noise batches")
    ap.add_argument("--no-parity", action="store_true", help="--config 1|3|4: skip the configuration's parity tests")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.ncu_child:
        ncu_child(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
