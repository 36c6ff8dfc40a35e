export ABEILLE_B200_KERNEL_TIMEOUT_S=60
timeout 900 python -m pytest tests -m gpu -x -q -k "vibration or noise_mode" 2>&1 | tail -15
