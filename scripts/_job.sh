export ABEILLE_B200_KERNEL_TIMEOUT_S=60
timeout 900 python -m pytest tests -m gpu -x -q -k "surface or ref_sqr or sood or PUa or golden" 2>&1 | tail -3
timeout 300 python bench.py --config 3 --steps 4 --warmup 3 --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('config3 lazy', 'value %.4g ms/step %.2f'%(d['value'], d['ms_per_step']))"
ABEILLE_B200_NO_BC_BOUND=1 timeout 300 python bench.py --config 3 --steps 4 --warmup 3 --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('config3 full', 'value %.4g ms/step %.2f'%(d['value'], d['ms_per_step']))"
