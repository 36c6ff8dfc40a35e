export ABEILLE_B200_KERNEL_TIMEOUT_S=60
timeout 900 python -m pytest tests -m gpu -x -q -k "noise_driver_drives" 2>&1 | tail -25
