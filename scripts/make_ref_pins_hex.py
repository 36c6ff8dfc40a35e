"""Generates tests/golden/ref_pins_hex.npz from the REFERENCE's own DeltaTracker::transport over its HexLattice
(src/hex_lattice.cpp compiled in place into oracle/_ref/libabeille_ref.so by `make -C oracle ref`).  Run in the container
that has /root/reference:

    python scripts/make_ref_pins_hex.py

Cases: oracle/ref_pins.py HEX_CASES; tests/test_reference_pins.py compares the oracle with this file bit for bit,
tests/test_gpu_reference_golden.py the kernels.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_pins  # noqa: E402

out = ref_pins.evaluate_transport("reference", ref_pins.HEX_CASES, seed0=1300)
path = os.path.join(ROOT, "tests", "golden", "ref_pins_hex.npz")
np.savez_compressed(path, **out)
print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
