"""Generates tests/golden/ref_pins_beam.npz from the REFERENCE's own FixedSource::run() (src/fixed_source.cpp) over sources with
mono-directional and cone direction distributions and Maxwellian and Watt energy distributions (src/mono_directional.cpp,
src/cone.cpp, src/maxwellian.cpp, src/watt.cpp, compiled in place into oracle/_ref/libabeille_ref.so by `make -C oracle ref`).
Run in the container that has /root/reference:

    python scripts/make_ref_pins_beam.py

Cases: oracle/ref_pins.py BEAM_CASES; tests/test_reference_pins.py compares the oracle with this file bit for bit,
tests/test_gpu_parity.py the device driver (abeille_b200/fixed_source.py).
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
    for i in range(len(ref_pins.BEAM_CASES)):  # a process per case: the reference keeps its state in process globals
        tmp = os.path.join(td, f"case{i}.npz")
        code = (f"import sys; sys.path.insert(0, {ROOT!r}); import numpy as np; from oracle import ref_pins; "
                f"np.savez({tmp!r}, **ref_pins.evaluate_beam_sources('reference', {i}))")
        subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.DEVNULL)
        out.update(dict(np.load(tmp)))
path = os.path.join(ROOT, "tests", "golden", "ref_pins_beam.npz")
np.savez_compressed(path, **out)
print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
