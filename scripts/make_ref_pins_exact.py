"""Generates tests/golden/ref_pins_exact.npz from the REFERENCE's own PowerIterator::run() with its BasicExactMGCancelator
(src/basic_exact_mg_cancelator.cpp and the vendored Sobol table vendor/sobol/src/sobol.cpp, compiled in place into
oracle/_ref/libabeille_ref.so by `make -C oracle ref`).  Run in the container that has /root/reference:

    python scripts/make_ref_pins_exact.py

Cases: oracle/ref_pins.py EXACT_PI_CASES (carter tracking with negative weights; beta minimum, average-f with sampled points,
average-g with Sobol points).  tests/test_reference_pins.py compares the oracle with this file bit for bit,
tests/test_gpu_reference_golden.py the device path.
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
for i in list(ref_pins.EXACT_PI_RANGE) + list(ref_pins.EXACT_FULL_PI_RANGE):  # one simulation per process (the reference keeps its state in process globals)
    with tempfile.TemporaryDirectory() as td:
        tmp = os.path.join(td, "pi.npz")
        code = (f"import sys; sys.path.insert(0, {ROOT!r}); import numpy as np; from oracle import ref_pins; "
                f"np.savez({tmp!r}, **ref_pins.evaluate_power_iteration('reference', only={i}))")
        subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
        out.update(dict(np.load(tmp)))
out["sobol_points"] = ref_pins.sobol_points("reference")
# the bank fields the exact cancelators read, from the reference's own trackers: one transport call per case
ex = ref_pins.evaluate_transport("reference", ref_pins.EXACT_TRANSPORT_CASES, seed0=1300, parents=True)
out.update({k.replace("transport_", "extransport_"): v for k, v in ex.items()})
path = os.path.join(ROOT, "tests", "golden", "ref_pins_exact.npz")
np.savez_compressed(path, **out)
print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
