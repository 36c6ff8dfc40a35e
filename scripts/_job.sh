export ABEILLE_B200_KERNEL_TIMEOUT_S=60
timeout 400 python -m pytest tests -m gpu -x -q -k "sharded_over_two_ranks" 2>&1 | tail -20
