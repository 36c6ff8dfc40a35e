"""compute-sanitizer target: a few small generations through the device-resident path (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.perf_probe import run
if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "c5g7_delta_collision.yaml", int(sys.argv[2]) if len(sys.argv) > 2 else 200000, 3, True)
