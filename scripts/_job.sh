export ABEILLE_B200_KERNEL_TIMEOUT_S=60
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/t9j_pytest_gpu.log
cat gpurun_out/t9j_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/t9j_bench_n1.json 2> gpurun_out/t9j_bench_n1.err; tail -c 300 gpurun_out/t9j_bench_n1.json
