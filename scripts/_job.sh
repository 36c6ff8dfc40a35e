set -x
export ABEILLE_B200_KERNEL_TIMEOUT_S=30
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t3b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t3b_pytest.log
tail -12 gpurun_out/t3b_pytest.log
( time timeout 900 python bench.py > gpurun_out/t3b_bench_n1.json 2> gpurun_out/t3b_bench.err ) 2>&1 | tail -3
cat gpurun_out/t3b_bench_n1.json
tail -3 gpurun_out/t3b_bench.err
