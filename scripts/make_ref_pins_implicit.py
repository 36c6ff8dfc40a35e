"""Generates tests/golden/ref_pins_implicit.npz from the REFERENCE's own ImplicitLeakageDeltaTracker::transport
(src/implicit_leakage_delta_tracker.cpp compiled in place into oracle/_ref/libabeille_ref.so by `make -C oracle ref`).  Run in
the container that has /root/reference:

    python scripts/make_ref_pins_implicit.py

Cases: oracle/ref_pins.py IMPLICIT_CASES; tests/test_reference_pins.py compares the oracle with this file bit for bit.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_pins  # noqa: E402

out = ref_pins.evaluate_transport("reference", ref_pins.IMPLICIT_CASES, seed0=700)
out.update(ref_pins.evaluate_noise("reference", ref_pins.IMPLICIT_NOISE_CASES, seed0=1100))  # noise mode, complex weights
# whole simulations through the reference's own PowerIterator::run() over its ImplicitLeakageDeltaTracker: one per process
import subprocess  # noqa: E402
import tempfile  # noqa: E402

for i in ref_pins.IMPLICIT_PI_RANGE:
    with tempfile.TemporaryDirectory() as td:
        tmp = os.path.join(td, "pi.npz")
        code = (f"import sys; sys.path.insert(0, {ROOT!r}); import numpy as np; from oracle import ref_pins; "
                f"np.savez({tmp!r}, **ref_pins.evaluate_power_iteration('reference', only={i}))")
        subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
        out.update(dict(np.load(tmp)))
path = os.path.join(ROOT, "tests", "golden", "ref_pins_implicit.npz")
np.savez_compressed(path, **out)
print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
