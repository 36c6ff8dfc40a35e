"""Generates tests/golden/ref_pins.npz from the REFERENCE's own compiled code (oracle/_ref/libabeille_ref.so, built by
`make -C oracle ref` from the sources under /root/reference).  Run in the container that has /root/reference:

    python scripts/make_ref_pins.py

The cases are defined in oracle/ref_pins.py; tests/test_reference_pins.py compares the oracle with this file bit for bit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_pins  # noqa: E402

out = ref_pins.evaluate("reference")
out.update(ref_pins.sample_mu("reference", out))
out.update(ref_pins.evaluate_transport("reference"))
out.update(ref_pins.evaluate_noise("reference"))
path = os.path.join(ROOT, "tests", "golden", "ref_pins.npz")
np.savez_compressed(path, **out)
print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
