"""Runs the bench workload's history kernel for every CUDA-library build under abeille_b200/lib/variants (kernel tuning
experiments: same ABI, different compile-time constants) and prints the kernel time of each (development aid)."""
import glob, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "10000000"
libs = [("default", None)] + [(os.path.basename(p), p) for p in sorted(glob.glob(os.path.join(root, "abeille_b200/lib/variants/*"))) if os.path.isdir(p)]
code = "import sys; sys.path.insert(0, %r); from scripts.perf_probe import run; run('c5g7_delta_collision_fullmesh.yaml', %s, 4, True)" % (root, n)
for name, path in libs:
    env = dict(os.environ)
    if path: env["ABEILLE_B200_LIBDIR"] = path
    try:
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=100, cwd=root)
        lines = [l for l in out.stdout.splitlines() if "kernel" in l]
        print(f"== {name}: rc {out.returncode}")
        for l in lines[-2:]: print("   ", l[l.index("gen"):][:200])
        for l in out.stderr.splitlines():
            if "event kernel" in l: print("   ", l[:500])
        for l in out.stderr.splitlines():
            if "event kernel" in l: print("   ", l[:400])
        if out.returncode: print(out.stderr[-800:])
    except subprocess.TimeoutExpired:
        print(f"== {name}: TIMEOUT")
    sys.stdout.flush()
