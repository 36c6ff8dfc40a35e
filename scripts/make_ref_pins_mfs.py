"""Generates tests/golden/ref_pins_mfs.npz from the REFERENCE's own ModifiedFixedSource::run() and FixedSource::run() (src/modified_fixed_source.cpp,
src/fixed_source.cpp compiled in place into oracle/_ref/libabeille_ref.so by `make -C oracle ref`).  Run in the container that has /root/reference:

    python scripts/make_ref_pins_mfs.py

Cases: oracle/ref_pins.py MFS_CASES; tests/test_reference_pins.py compares the oracle with this file bit for bit,
tests/test_gpu_parity.py the device driver (abeille_b200/fixed_source.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_pins  # noqa: E402

out = ref_pins.evaluate_modified_fixed_source("reference")
# FixedSource::run() in a process of its own (the reference keeps its state in process globals)
import subprocess  # noqa: E402
import tempfile  # noqa: E402

with tempfile.TemporaryDirectory() as td:
    tmp = os.path.join(td, "fs.npz")
    code = (f"import sys; sys.path.insert(0, {ROOT!r}); import numpy as np; from oracle import ref_pins; "
            f"np.savez({tmp!r}, **ref_pins.evaluate_fixed_source('reference'))")
    subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
    out.update(dict(np.load(tmp)))
path = os.path.join(ROOT, "tests", "golden", "ref_pins_mfs.npz")
np.savez_compressed(path, **out)
print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
