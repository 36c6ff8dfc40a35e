export ABEILLE_B200_KERNEL_TIMEOUT_S=60
timeout 900 python -m pytest tests -m gpu -x -q -k "branchless or restart or noise or carter or fixed_source or implicit" 2>&1 | tail -15
