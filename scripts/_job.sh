export ABEILLE_B200_KERNEL_TIMEOUT_S=30
for mode in fixed general fixed general; do
  if [ $mode = general ]; then export ABEILLE_B200_NO_FIXED_SHAPE=1; else unset ABEILLE_B200_NO_FIXED_SHAPE; fi
  timeout 300 python bench.py --no-e2e --no-cpu --no-ncu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', 'value %.4g ms/step %.2f kernel_ms %.2f share %.3f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step']))"
done
unset ABEILLE_B200_NO_FIXED_SHAPE
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
