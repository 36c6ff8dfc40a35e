set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s1c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s1c_pytest.log
tail -15 gpurun_out/s1c_pytest.log
timeout 600 python scripts/variant_probe.py > gpurun_out/s1c_variants.log 2>&1
grep -E "^==|gen3" gpurun_out/s1c_variants.log
