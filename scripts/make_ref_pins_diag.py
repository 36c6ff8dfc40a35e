"""Generates tests/golden/ref_pins_diag.npz from the REFERENCE's own PowerIterator::run() (oracle/_ref/libabeille_ref.so, built by
`make -C oracle ref`) with settings pair-distance-sqrd, families and empty-entropy-bins on: the three per-generation series and the
final bank the last pair distance (the reference's double sum over all pairs, src/power_iterator.cpp:637-663) was taken over.
Run in the container that has /root/reference:

    python scripts/make_ref_pins_diag.py

Cases: oracle/ref_pins.py DIAG_CASES; tests/test_reference_pins.py holds abeille_b200.distributed.pair_distance_sqrd (two passes of
weighted moments, what abl_bank_moments_device computes) and the other definitions against this file.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_pins  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as td:
    for i in range(len(ref_pins.DIAG_CASES)):  # a process per case: the reference keeps its state in process globals
        tmp = os.path.join(td, f"case{i}.npz")
        code = (f"import sys; sys.path.insert(0, {ROOT!r}); import numpy as np; from oracle import ref_pins; "
                f"np.savez({tmp!r}, **ref_pins.evaluate_pi_diagnostics({i}))")
        subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
        out.update(dict(np.load(tmp)))
path = os.path.join(ROOT, "tests", "golden", "ref_pins_diag.npz")
np.savez_compressed(path, **out)
print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
