"""ncu target: a few converged C5G7 delta-tracking generations at 1e6 particles (device-resident path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.perf_probe import run
if __name__ == "__main__":
    deck = sys.argv[1] if len(sys.argv) > 1 else "c5g7_delta_collision_fullmesh.yaml"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    gens = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    run(deck, n, gens, True)
