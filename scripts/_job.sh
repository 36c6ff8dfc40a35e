export ABEILLE_B200_KERNEL_TIMEOUT_S=60
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in 1 3 4; do
  timeout 600 python bench.py --config $c --steps 4 --warmup 3 > gpurun_out/t5a_bench_config$c.json 2> gpurun_out/t5a_bench_config$c.err
  python -c "import json; d=json.load(open('gpurun_out/t5a_bench_config$c.json')); print('config $c', 'value %.4g ms/step %.2f kernel_ms %.2f coll/part %.1f k %.5f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['collisions_per_particle'], d['config']['k_col']), d.get('parity'))" || tail -5 gpurun_out/t5a_bench_config$c.err
done
