export ABEILLE_B200_KERNEL_TIMEOUT_S=60
( time timeout 600 python -m pytest tests -m gpu -x -q -k "million" 2>&1 | tail -6 ) 2>&1 | tail -10
