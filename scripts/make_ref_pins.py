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
out.update(ref_pins.evaluate_power_iteration("reference"))
out.update(ref_pins.evaluate_noise_driver("reference"))
# mesh-tally arrays of more than 100 000 elements are kept as the SHA-256 of their bytes plus shape: the comparison in
# tests/test_reference_pins.py is bit for bit either way, and the fixture stays small
import hashlib  # noqa: E402

small = {}
for k, v in out.items():
    if v.size > 100000 and "tally" in k:
        small[k + "__sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(v).tobytes()).digest(), dtype=np.uint8)
        small[k + "__shape"] = np.asarray(v.shape, dtype=np.int64)
    else:
        small[k] = v
out = small
path = os.path.join(ROOT, "tests", "golden", "ref_pins.npz")
np.savez_compressed(path, **out)
print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
