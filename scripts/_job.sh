export ABEILLE_B200_KERNEL_TIMEOUT_S=60
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/t9a_pytest_gpu.log
cat gpurun_out/t9a_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/t9a_bench_n1.json 2> gpurun_out/t9a_bench_n1.err; tail -c 600 gpurun_out/t9a_bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t9a_bench_reference_arm.json 2>/dev/null; tail -c 300 gpurun_out/t9a_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t9a_launches.csv python bench.py --steps 2 --warmup 1 --no-parity > gpurun_out/t9a_ncu_bench.log 2>&1; tail -3 gpurun_out/t9a_launches.csv | cut -c1-200
