export ABEILLE_B200_KERNEL_TIMEOUT_S=60
timeout 300 python -m pytest tests -m gpu -x -q -k "fixed_source or two_phase" 2>&1 | tail -12
