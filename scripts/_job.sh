export ABEILLE_B200_KERNEL_TIMEOUT_S=30
timeout 300 python -m pytest tests -m gpu -x -q -k "hex" 2>&1 | tail -6
timeout 200 python bench.py --no-e2e --no-cpu --no-ncu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', 'value %.4g ms/step %.2f kernel_ms %.2f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
