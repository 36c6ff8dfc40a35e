"""Generates tests/golden/ref_pins_vibration.npz from the REFERENCE's own noise-mode transport (NoiseMaker + FlatVibrationNoiseSource compiled in
place into oracle/_ref/libabeille_ref.so by `make -C oracle ref`) with the noise frequency at the third harmonic of the vibration.  Run in
the container that has /root/reference:

    python scripts/make_ref_pins_vibration.py

Cases: oracle/ref_pins.py HARMONIC_NOISE_CASES; tests/test_reference_pins.py compares the oracle with this file bit for bit,
tests/test_gpu_reference_golden.py the kernels.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_pins  # noqa: E402

out = ref_pins.evaluate_noise("reference", ref_pins.HARMONIC_NOISE_CASES, seed0=1500)
path = os.path.join(ROOT, "tests", "golden", "ref_pins_vibration.npz")
np.savez_compressed(path, **out)
print(f"{path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")
