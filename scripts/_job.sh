export ABEILLE_B200_KERNEL_TIMEOUT_S=30
export ABEILLE_B200_NO_SMEM_TABLES=1
for v in a768 a896; do
  export ABEILLE_B200_LIBDIR=$PWD/abeille_b200/lib/variants/$v
  timeout 200 python bench.py --no-e2e --no-cpu --no-ncu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'value %.4g ms/step %.2f kernel_ms %.2f grid %s'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['grid']))"
done
