export ABEILLE_B200_KERNEL_TIMEOUT_S=60
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/t8a_pytest_gpu.log
cat gpurun_out/t8a_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/t8a_bench_n1.json 2> gpurun_out/t8a_bench_n1.err; tail -c 1800 gpurun_out/t8a_bench_n1.json
timeout 400 python bench.py --config 5 --no-parity > gpurun_out/t8a_bench_config5_n1.json 2> gpurun_out/t8a_c5.err; tail -c 500 gpurun_out/t8a_bench_config5_n1.json
