export ABEILLE_B200_KERNEL_TIMEOUT_S=60
( time timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t7a_pytest.log 2>&1 ) 2>&1 | grep real; tail -3 gpurun_out/t7a_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/t7a_bench_n1.json 2> gpurun_out/t7a_bench.err ) 2>&1 | grep real
python -c "import json; d=json.load(open('gpurun_out/t7a_bench_n1.json')); print('bench value %.4g ms/step %.2f kernel %.2f e2e %.4g cpu %.4g traffic %s lanes %s'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['traffic'], d['roofline']['profile'] and d['roofline']['profile']['lanes_per_instruction']))"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/t7a_bench_reference.json 2>/dev/null; head -c 400 gpurun_out/t7a_bench_reference.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/t7a_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-ncu > /dev/null 2>&1; grep -c . gpurun_out/t7a_launches.csv
